"""Per-step cost probe on a few very long targets (development aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import SequenceDB, datasets, matrices
from opal_b200.handle import OpalB200
eng = OpalB200(); sm = matrices.blosum62(); rng = np.random.default_rng(1)
T = int(os.environ.get("TLEN", "30000")); n = int(os.environ.get("NSEQ", "64"))
db = SequenceDB.from_sequences([datasets.random_residues(T, rng, sm) for _ in range(n)])
h = eng.create_db(db, 0)
for qlen in (144, 513):
    q = datasets.random_residues(qlen, rng, sm)
    for mode in ("SW", "NW"):
        best = 1e9
        for _ in range(3):
            rc, sc, eq, et, ms = h.search(q, 11, 1, sm.flat(), 23, 1, mode)
            best = min(best, ms)
        st = h.last_stats()
        steps = T + st["G"] - 1
        print(f"Q={qlen} {mode}: {best:.3f} ms  -> {best*1e-3*1.965e9/steps/st['passes']:.0f} clk/step  {st}")
