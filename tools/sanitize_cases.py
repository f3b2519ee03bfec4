"""Small searches that together touch every kernel path (run under compute-sanitizer by tools/sanitize.sh):
the README example (all modes, alignments), a folded latency class beside a bulk group, a multi-pass query with
boundary rows through HBM, chained passes, a 16 -> 32 bit re-run, a-priori 32-bit routing, the alignment stage, a batch with several
queries in flight, the top-k pipeline and two shards on one device.  Every result is compared with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import MODES, ORACLE_SO, README_DB, README_MATRIX, README_QUERY, OpalCLibrary, SequenceDB, dump_results, free_alignments  # noqa: E402
from opal_b200 import datasets, matrices  # noqa: E402
from opal_b200.handle import OpalB200  # noqa: E402

eng, oracle = OpalB200(), OpalCLibrary(ORACLE_SO)
checked = 0


def same(q, db, go, ge, m, a, mode, st):
    global checked
    rc1, want = oracle.search_database(q, db, go, ge, m, a, None, st, MODES[mode])
    rc2, got = eng.search_database(q, db, go, ge, m, a, None, st, MODES[mode])
    assert rc1 == rc2 == 0, (rc1, rc2, eng.last_error())
    w, g = dump_results(want), dump_results(got)
    if mode == "SW":  # reference quirk Q2: end of a zero-score hit
        for x in w:
            if x[1] == 0:
                x[2] = x[3] = -1
    assert g == w, (mode, st, [(i, a_, b_) for i, (a_, b_) in enumerate(zip(g, w)) if a_ != b_][:3])
    free_alignments(want)
    free_alignments(got)
    checked += len(db)


rng = np.random.default_rng(11)
sm = matrices.blosum62()
readme = SequenceDB.from_sequences([np.array(s, np.uint8) for s in README_DB])
for mode in ("NW", "HW", "OV", "SW"):
    for st in (0, 1, 2):
        same(README_QUERY, readme, 3, 1, README_MATRIX, 4, mode, st)
# folded latency class + bulk, one pass (Q = 300) and several passes (Q = 1300)
seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(1, 120, 260)]
seqs[0] = datasets.random_residues(1500, rng, sm)
seqs[1] = datasets.random_residues(900, rng, sm)
q300 = datasets.random_residues(300, rng, sm)
seqs[2] = datasets.mutate(q300, 0.8, rng, sm)
db = SequenceDB.from_sequences(seqs)
for mode in ("SW", "NW", "HW", "OV"):
    same(q300, db, 11, 1, sm.flat(), 23, mode, 1)
same(datasets.random_residues(1300, rng, sm), SequenceDB.from_sequences(seqs[:40]), 11, 1, sm.flat(), 23, "SW", 1)
same(datasets.random_residues(1300, rng, sm), SequenceDB.from_sequences(seqs[:40]), 11, 1, sm.flat(), 23, "OV", 1)
# chained passes (every pass of a long query on a warp of its own, boundary rows handed over through L2)
os.environ["OPAL_B200_CHAIN"] = "1"
q1300 = datasets.random_residues(1300, rng, sm)
for mode in ("SW", "NW", "HW", "OV"):
    same(q1300, SequenceDB.from_sequences(seqs[:40]), 11, 1, sm.flat(), 23, mode, 1)
del os.environ["OPAL_B200_CHAIN"]
# 16 -> 32 bit re-run (8 x BLOSUM62 on near-identical sequences) and a-priori 32-bit routing (huge gap penalties)
big = (sm.flat() * 40).astype(np.int32)
qq = datasets.random_residues(400, rng, sm)
same(qq, SequenceDB.from_sequences([qq.copy(), datasets.mutate(qq, 0.9, rng, sm)] + seqs[3:20]), 440, 40, big, 23, "SW", 1)
same(q300[:80], SequenceDB.from_sequences(seqs[3:30]), 3000, 900, sm.flat(), 23, "NW", 1)
# alignment stage, every mode
for mode in ("SW", "NW", "HW", "OV"):
    same(q300[:90], SequenceDB.from_sequences([datasets.mutate(q300[:90], 0.7, rng, sm) for _ in range(12)] + seqs[40:60]), 11, 1, sm.flat(), 23, mode, 2)
# handle: batch in flight, top-k, two shards on one device
h = eng.create_db(db, [0, 0])
qs = [q300, q300[:50], datasets.random_residues(200, rng, sm)]
rc, S, Q, T, _ = h.search_batch(qs, 11, 1, sm.flat(), 23, 1, "SW", in_flight=3)
assert rc == 0
for k, q in enumerate(qs):
    rc, want = oracle.search_database(q, db, 11, 1, sm.flat(), 23, None, 1, MODES["SW"])
    assert rc == 0 and (want["score"] == S[k]).all()
rc, idx, res = h.search_topk(q300, 11, 1, sm.flat(), 23, 2, "SW", 10)
assert rc == 0 and len(idx) == 10 and idx[0] == 2
free_alignments(res)
h.close()
print(f"sanitize_cases ok: {checked} records compared with the oracle")
