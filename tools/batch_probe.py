"""Multi-query batch throughput through opalb200_db_search_batch (development aid):
python tools/batch_probe.py [config2|N] [mode] [searchType] [numQueries]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import datasets, matrices  # noqa: E402
from opal_b200.handle import OpalB200  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "config2"
mode = sys.argv[2] if len(sys.argv) > 2 else "SW"
st = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 32
eng = OpalB200()
sm = matrices.blosum62()
q = sm.encode(datasets.P18080)
rng = np.random.default_rng(1)
if which == "config2":
    db = datasets.config2_db(sm, q)
    queries = [q] + [datasets.mutate(q, 0.5, rng, sm)[:len(q)] for _ in range(nq - 1)]
else:
    db = datasets.config3_db(sm, n=int(which))
    queries = datasets.config3_queries(sm)[:nq]
h = eng.create_db(db, 0)
cells = sum(len(x) for x in queries) * db.total_residues
single = 0.0
for x in queries:
    best = 1e9
    for _ in range(2):
        rc, *_, ms = h.search(x, 11, 1, sm.flat(), 23, st, mode)
        assert rc == 0
        best = min(best, ms)
    single += best
print(f"{len(queries)} queries, one at a time: device {single:.3f} ms  {cells/single/1e6:.0f} GCUPS")
for k in (1, 2, 3, 4, 6):
    best_ms, best_wall = 1e9, 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        rc, S, EQ, ET, ms = h.search_batch(queries, 11, 1, sm.flat(), 23, st, mode, in_flight=k)
        wall = (time.perf_counter() - t0) * 1e3
        assert rc == 0, eng.last_error()
        best_ms, best_wall = min(best_ms, ms), min(best_wall, wall)
    print(f"in_flight={k}: device {best_ms:.3f} ms  {cells/best_ms/1e6:.0f} GCUPS   wall {best_wall:.3f} ms  {cells/best_wall/1e6:.0f} GCUPS", flush=True)
h.close()
