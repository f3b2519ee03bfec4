"""The drop-in test: a plain C program including include/opal.h, linked against libopal_b200.so, must
reproduce the reference's README answers (tests/golden/readme.json)."""
import os
import subprocess

import pytest

from _util import MODES, ROOT
from test_oracle_golden import golden

pytestmark = pytest.mark.gpu


def test_plain_c_caller(tmp_path):
    exe = str(tmp_path / "c_linkage")
    libdir = os.path.join(ROOT, "opal_b200", "csrc")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_linkage_main.c"), "-L", libdir, "-lopal_b200",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    g = golden("readme.json")
    for mode, code in MODES.items():
        out = subprocess.run([exe, str(code)], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        want = [f"{r[1]} {r[4]} {r[5]} {r[2]} {r[3]} {r[7]}" for r in g[f"{mode}/2/1"]["results"]]
        assert out.stdout.strip().split("\n") == want, mode
