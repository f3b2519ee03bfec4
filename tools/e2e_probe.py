"""Drop-in call phase timings (development aid): OPAL_B200_TRACE=1 python tools/e2e_probe.py [config2|config3]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import MODES, datasets, matrices, new_results
from opal_b200.handle import OpalB200
eng = OpalB200()
sm = matrices.blosum62()
q = sm.encode(datasets.P18080)
which = sys.argv[1] if len(sys.argv) > 1 else "config2"
db = datasets.config2_db(sm, q) if which == "config2" else datasets.config3_db(sm, query=q)
for i in range(5):
    t0 = time.perf_counter()
    res = new_results(len(db))
    t1 = time.perf_counter()
    rc, res = eng.search_database(q, db, 11, 1, sm.flat(), 23, res, 1, MODES["SW"])
    t2 = time.perf_counter()
    print(f"call {i}: new_results {1e3*(t1-t0):.3f} ms, opalSearchDatabase {1e3*(t2-t1):.3f} ms rc={rc}", flush=True)
