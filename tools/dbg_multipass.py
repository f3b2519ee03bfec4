import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import MODES, ORACLE_SO, OpalCLibrary, SequenceDB
from opal_b200 import datasets, matrices
from opal_b200.handle import OpalB200
qlen = 2600
rng = np.random.default_rng(qlen)
sm = matrices.blosum62()
q = datasets.random_residues(qlen, rng, sm)
seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(1, 400, 21)]
seqs[2] = datasets.mutate(q, 0.7, rng, sm)
seqs[5] = q[700:1900].copy()
db = SequenceDB.from_sequences(seqs)
eng = OpalB200(); ora = OpalCLibrary(ORACLE_SO)
h = eng.create_db(db, 0)
for mode in ("NW", "SW", "HW", "OV"):
    rc, want = ora.search_database(q, db, 11, 1, sm.flat(), 23, None, 1, MODES[mode])
    rc, sc, eq, et, ms = h.search(q, 11, 1, sm.flat(), 23, 1, mode)
    bad = np.nonzero(sc != want["score"])[0]
    print(mode, h.last_stats(), "bad", bad.tolist(), [(int(sc[i]), int(want["score"][i]), int(db.lengths[i])) for i in bad[:4]])
