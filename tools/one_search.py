"""One resident search (development aid for ncu): python tools/one_search.py [config2|N] [mode] [type] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import datasets, matrices  # noqa: E402
from opal_b200.handle import OpalB200  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "config2"
mode = sys.argv[2] if len(sys.argv) > 2 else "SW"
st = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
eng = OpalB200()
sm = matrices.blosum62()
q = sm.encode(datasets.P18080)
qlen = int(os.environ.get("QLEN", "0"))
if qlen:
    q = [x for x in datasets.config3_queries(sm) if len(x) == qlen][0]
db = datasets.config2_db(sm, q) if which == "config2" else datasets.config3_db(sm, n=int(which))
h = eng.create_db(db, 0)
for _ in range(reps):
    rc, sc, eq, et, ms = h.search(q, 11, 1, sm.flat(), 23, st, mode)
    print(rc, ms, "ms", len(q) * db.total_residues / ms / 1e6, "GCUPS", h.last_stats())
h.close()
