"""BASELINE configs[4]: DNA (alphabet 4, +5/-4, gap 16/4: logarithmic regime for random DNA, Go >= 2 Ge), 10 kb query vs a heavy-tailed database with targets up to
100 kb and a few planted near-copies of the query (score > 32767 -> 32-bit re-run); many passes over the query.
python tools/config5_probe.py [numSequences] [check]   -- `check` compares a sample with the oracle / reference."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import MODES, OpalCLibrary, datasets, matrices  # noqa: E402
from opal_b200.handle import OpalB200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
check = len(sys.argv) > 2
eng = OpalB200()
GO, GE = int(os.environ.get("GO", "16")), int(os.environ.get("GE", "4"))
sm = matrices.simple(4, 5, -4)
rng = np.random.default_rng(20261019)
q = rng.integers(0, 4, 10000, dtype=np.uint8)
t0 = time.time()
db = datasets.dna_db(n, 20261019, query=q)
print(f"db: {len(db)} seqs, {db.total_residues} residues, longest {int(db.lengths.max())}, gen {time.time()-t0:.1f}s", flush=True)
t0 = time.time()
h = eng.create_db(db, 0)
print(f"pack+upload {time.time()-t0:.3f}s", flush=True)
cells = len(q) * db.total_residues
for mode in (os.environ.get("MODES", "SW").split(",")):
    for st in (0, 1):
        best = 1e9
        for _ in range(2):
            rc, sc, eq, et, ms = h.search(q, GO, GE, sm.flat(), 4, st, mode)
            assert rc == 0, (rc, eng.last_error())
            best = min(best, ms)
        print(f"{mode} type={st}: {best:.1f} ms  {cells/best/1e6:.0f} GCUPS  max score {int(sc.max())}  stats={h.last_stats()}", flush=True)
        if check:
            ref_so = os.path.join(ROOT, "oracle", "_ref", "libopal_ref.so")
            lib = OpalCLibrary(ref_so if os.path.exists(ref_so) else os.path.join(ROOT, "oracle", "liboracle.so"))
            order = np.argsort(-sc.astype(np.int64), kind="stable")
            sample = sorted(set(order[:6].tolist()) | set(np.argsort(-db.lengths, kind="stable")[:3].tolist()) | set(range(0, n, max(1, n // 40))))
            sub = db.subset(sample)
            t1 = time.time()
            rc2, want = lib.search_database(q, sub, GO, GE, sm.flat(), 4, None, st, MODES[mode])
            ok = rc2 == 0 and (want["score"] == sc[sample]).all()
            if st:
                ok = ok and (want["endLocationQuery"] == eq[sample]).all() and (want["endLocationTarget"] == et[sample]).all()
            print(f"   {len(sample)} sampled targets ({sub.total_residues} residues) vs {os.path.basename(lib.path)}: identical={bool(ok)}  ({time.time()-t1:.1f}s on the host)", flush=True)
h.close()
