"""The drop-in test of SURVEY.md section 8b: the reference's OWN programs on the new library.

oracle/Makefile (target `dropin`, run by __graft_entry__.build() where /root/reference is present) compiles the
reference's unmodified self-test (src/test.cpp) and command line tool (src/opal_aligner.cpp + ScoreMatrix.cpp) against
include/opal.h and links them with libopal_b200.so instead of opal.cpp.  Here they run on the GPU:

  * `test SW|NW|HW|OV` (src/test.cpp:97-99: OPAL_SEARCH_ALIGNMENT over 200 random targets, compared by the program
    itself with its scalar DP and replayed by its checkAlignment, :154-167, :348-422) must print the known maxima
    573 / 460 / 567 / 567 and not a single mismatch line;
  * `opal_aligner` (src/opal_aligner.cpp:158-160) must print, line by line, what the unmodified reference printed for
    the same files and options (tests/golden/cli/*.txt) -- the timing lines excepted.
"""
import json
import os
import re
import subprocess

import pytest

from _util import GOLDEN_DIR, ROOT

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
TEST_BIN = os.path.join(REF_DIR, "test_on_b200")
ALIGNER_BIN = os.path.join(REF_DIR, "opal_aligner_on_b200")
FIX = os.path.join(GOLDEN_DIR, "cli")
INDEX = json.load(open(os.path.join(FIX, "index.json")))
TIMING = ("Cpu time of searching:", "GCUPS (giga cell updates per second):")
MISMATCH = re.compile(r"^#\d+:|Alignment went outside|Should be m|Alignment ended at|Wrong score|Error")


def _need(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} not built (needs /root/reference at build time: make -C oracle)")


@pytest.mark.parametrize("mode,maximum", [("SW", 573), ("NW", 460), ("HW", 567), ("OV", 567)])
def test_reference_self_test_passes_on_the_new_library(mode, maximum):
    _need(TEST_BIN)
    r = subprocess.run([TEST_BIN, mode], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-300:])
    lines = r.stdout.split("\n")
    assert "Starting Opal!" in lines and "Starting normal!" in lines
    assert [l for l in lines if l.startswith("Maximum:")] == [f"Maximum: {maximum}"] * 2  # the library's, then the scalar DP's
    bad = [l for l in lines if MISMATCH.search(l)]
    assert not bad, bad[:5]
    assert any(l.startswith("Times faster:") for l in lines)  # the program ran to its end


@pytest.mark.parametrize("case", sorted(INDEX))
def test_reference_cli_on_the_new_library_prints_the_golden_output(case):
    _need(ALIGNER_BIN)
    c = INDEX[case]
    opts = [os.path.join(FIX, o) if o.endswith(".mat") else o for o in c["options"]]
    r = subprocess.run([ALIGNER_BIN] + opts + [os.path.join(FIX, c["query"]), os.path.join(FIX, c["db"])],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-300:], r.stderr[-300:])
    want = open(os.path.join(FIX, case + ".txt")).read()
    strip = lambda text: [l for l in text.split("\n") if not l.startswith(TIMING)]
    assert strip(r.stdout) == strip(want)
