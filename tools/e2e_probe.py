"""Drop-in call phase timings (development aid):
OPAL_B200_TRACE=1 [SHARD_OF=8] [QLEN=2005] [MODE=NW] python tools/e2e_probe.py [config2|config3]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import MODES, datasets, matrices, new_results, sharding
from opal_b200.handle import OpalB200
eng = OpalB200()
sm = matrices.blosum62()
p18080 = sm.encode(datasets.P18080)
which = sys.argv[1] if len(sys.argv) > 1 else "config2"
db = datasets.config2_db(sm, p18080) if which == "config2" else datasets.config3_db(sm, query=p18080)
shard_of = int(os.environ.get("SHARD_OF", "1"))
if shard_of > 1:
    db = sharding.shard_db(db, sharding.deal_shards(db.lengths, shard_of)[0])
q = p18080
qlen = int(os.environ.get("QLEN", "0"))
if qlen:
    q = [x for x in datasets.config3_queries(sm) if len(x) == qlen][0]
mode = os.environ.get("MODE", "SW")
print(f"{len(db)} sequences, {db.total_residues} residues, Q = {len(q)}, {mode}", flush=True)
for i in range(5):
    t0 = time.perf_counter()
    res = new_results(len(db))
    t1 = time.perf_counter()
    rc, res = eng.search_database(q, db, 11, 1, sm.flat(), 23, res, 1, MODES[mode])
    t2 = time.perf_counter()
    print(f"call {i}: new_results {1e3*(t1-t0):.3f} ms, opalSearchDatabase {1e3*(t2-t1):.3f} ms rc={rc}", flush=True)
