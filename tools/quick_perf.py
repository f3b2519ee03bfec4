"""Quick device-timed GCUPS probe through the resident-handle API (development aid)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import datasets, matrices  # noqa: E402
from opal_b200.handle import OpalB200  # noqa: E402


def main():
    eng = OpalB200()
    gc, ips, ms = eng.measure_dpx_peak(0)
    print(f"DPX peak: {gc:.0f} GCUPS-equivalent (SW, 6 instr / 2 cells), {ips/1e12:.2f} T thread-instr/s, {ms:.3f} ms")
    sm = matrices.blosum62()
    q = sm.encode(datasets.P18080)
    which = sys.argv[1] if len(sys.argv) > 1 else "config2"
    t0 = time.time()
    if which == "config2":
        db = datasets.config2_db(sm, q)
    else:
        db = datasets.config3_db(sm, n=int(sys.argv[2]) if len(sys.argv) > 2 else 570000)
    print(f"db: {len(db)} seqs, {db.total_residues} residues, gen {time.time()-t0:.1f}s")
    t0 = time.time()
    h = eng.create_db(db, 0)
    print(f"pack+upload {time.time()-t0:.3f}s")
    queries = [("P18080", q)]
    if which != "config2":
        want = [int(x) for x in os.environ.get("QLENS", "144,375,1000,2005,5478").split(",")]
        queries += [(f"Q{len(x)}", x) for x in datasets.config3_queries(sm) if len(x) in want]
    modes = tuple(os.environ["MODES"].split(",")) if "MODES" in os.environ else (("SW", "NW", "HW", "OV") if which == "config2" else ("SW", "NW"))
    for name, qq in queries:
        for mode in modes:
            for st in (0, 1):
                best = 1e9
                for _ in range(3):
                    rc, sc, eq, et, ms = h.search(qq, 11, 1, sm.flat(), 23, st, mode)
                    assert rc == 0, (rc, eng.last_error())
                    best = min(best, ms)
                stats = h.last_stats()
                print(f"{name} Q={len(qq)} {mode} type={st}: {best:.3f} ms  "
                      f"{len(qq)*db.total_residues/best/1e6:.0f} GCUPS  stats={stats}", flush=True)
    h.close()


if __name__ == "__main__":
    main()
