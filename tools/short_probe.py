"""Geometry comparison on the short targets of the configs[1] database only (development aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import SequenceDB, datasets, matrices
from opal_b200.handle import OpalB200
eng = OpalB200(); sm = matrices.blosum62(); q = sm.encode(datasets.P18080)
lim = int(sys.argv[1]) if len(sys.argv) > 1 else 600
full = datasets.config2_db(sm, q) if len(sys.argv) < 3 else datasets.config3_db(sm, n=int(sys.argv[2]))
keep = [i for i in range(len(full)) if full.lengths[i] < lim]
db = full.subset(keep * int(os.environ.get("REPL", "1")))
print(f"{len(db)} of {len(full)} sequences below {lim}: {db.total_residues} of {full.total_residues} residues")
print("longer ones:", sorted((int(x) for x in full.lengths if x >= 1900), reverse=True)[:40])
h = eng.create_db(db, 0)
for geo in [None] + os.environ.get("GEOS", "32,17,2 16,33,2 16,33,1 32,17,1").split():
    if geo: os.environ["OPAL_B200_GEOMETRY"] = geo
    best = 1e9
    for _ in range(4):
        rc, sc, eq, et, ms = h.search(q, 11, 1, sm.flat(), 23, 1, "SW")
        best = min(best, ms)
    print(geo, f"{best:.3f} ms  {len(q)*db.total_residues/best/1e6:.0f} GCUPS", h.last_stats())
