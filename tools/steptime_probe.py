"""Step time of the wavefront kernel (development aid; feeds the planner's timing model in engine.cu).

A database of exactly one target pair per resident warp, all of one length, swept with a forced geometry
(OPAL_B200_GEOMETRY=32,R,k): the device time divided by the steps of one pair is the time of one step of a warp that
runs alone on its scheduler partition (k = 1) or shares it with another warp (k = 2).
usage: python tools/steptime_probe.py [length]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import SequenceDB, datasets, matrices  # noqa: E402
from opal_b200.handle import OpalB200  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
eng = OpalB200()
sm = matrices.blosum62()
rng = np.random.default_rng(3)
SMS, MHZ = 148, 1965.0
for k in [int(x) for x in os.environ.get("KS", "1,2").split(",")]:
    pairs = SMS * 4 * k
    db = SequenceDB.from_sequences([datasets.random_residues(T, rng, sm) for _ in range(2 * pairs)])
    h = eng.create_db(db, 0)
    for mode, st in [(m.split(":")[0], int(m.split(":")[1])) for m in os.environ.get("MODES", "SW:1,SW:0,NW:1,HW:1").split(",")]:
        row = []
        for R in (4, 8, 9, 12, 17, 24, 33):
            q = datasets.random_residues(32 * R, rng, sm)
            os.environ["OPAL_B200_GEOMETRY"] = f"32,{R},{k}"
            best = 1e9
            for _ in range(3):
                rc, sc, eq, et, ms = h.search(q, 11, 1, sm.flat(), 23, st, mode)
                assert rc == 0
                best = min(best, ms)
            stats = h.last_stats()
            if not (stats["R"] == R and stats["G"] == 32 and stats["warps_per_partition"] == k):
                row.append(f"R={R}: n/a")
                continue
            row.append(f"R={R}: {best * 1e-3 * MHZ * 1e6 / (T + 31):.0f}")
        print(f"k={k} {mode} type={st}  cycles per step  " + "  ".join(row), flush=True)
    h.close()

# Desynchronised variant: pairs of mixed lengths, several per warp, so that the warps of a partition are at different
# points of their steps (the equal-length runs above keep them in lockstep, the worst case for pipe contention).
if os.environ.get("MIXED", "1") == "1":
    for k in [int(x) for x in os.environ.get("KS", "1,2").split(",")]:
        warps = SMS * 4 * k
        lens = np.sort(rng.integers(T // 4, T, 2 * 10 * warps))[::-1]
        db = SequenceDB.from_sequences([datasets.random_residues(int(x), rng, sm) for x in lens])
        steps = float(sum(int(lens[2 * i]) + 31 for i in range(len(lens) // 2)))
        h = eng.create_db(db, 0)
        for mode, st in [(m.split(":")[0], int(m.split(":")[1])) for m in os.environ.get("MODES", "SW:1,SW:0,NW:1,HW:1").split(",")]:
            row = []
            for R in (9, 17, 33):
                q = datasets.random_residues(32 * R, rng, sm)
                os.environ["OPAL_B200_GEOMETRY"] = f"32,{R},{k}"
                best = 1e9
                for _ in range(3):
                    rc, sc, eq, et, ms = h.search(q, 11, 1, sm.flat(), 23, st, mode)
                    assert rc == 0
                    best = min(best, ms)
                row.append(f"R={R}: {best * 1e-3 * MHZ * 1e6 * warps / steps:.0f}")
            print(f"mixed k={k} {mode} type={st}  effective cycles per warp-step  " + "  ".join(row), flush=True)
        h.close()
