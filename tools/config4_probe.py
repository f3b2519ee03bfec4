"""BASELINE configs[3]: SW score pass, top-1000 hits, OPAL_SEARCH_ALIGNMENT on the sub-database with prefilled
results (reuse path), timed end to end through opalSearchDatabase; optional comparison with oracle/_ref."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opal_b200 import MODES, OpalCLibrary, datasets, matrices, new_results, free_alignments, get_alignment
from opal_b200.handle import OpalB200
eng = OpalB200(); sm = matrices.blosum62(); q = sm.encode(datasets.P18080)
which = sys.argv[1] if len(sys.argv) > 1 else "config2"
db = datasets.config2_db(sm, q) if which == "config2" else datasets.config3_db(sm, query=q)
for scale, go, ge in ((1, 11, 1), (8, 88, 8)):
    m = (sm.matrix * scale).ravel()
    t0 = time.perf_counter()
    rc, res = eng.search_database(q, db, go, ge, m, 23, None, 0, MODES["SW"])
    t1 = time.perf_counter()
    top = np.argsort(-res["score"].astype(np.int64), kind="stable")[:1000]
    sub = db.subset(top)
    pre = new_results(len(sub)); pre["scoreSet"] = 1; pre["score"] = res["score"][top]
    for rep in range(3):
        r2 = pre.copy()
        t2 = time.perf_counter()
        rc2, out = eng.search_database(q, sub, go, ge, m, 23, r2, 2, MODES["SW"], entry="opalSearchDatabaseRescore")
        t3 = time.perf_counter()
        if rep < 2: free_alignments(out)
    cells = int(((out["endLocationQuery"] - out["startLocationQuery"] + 1).astype(np.int64) * (out["endLocationTarget"] - out["startLocationTarget"] + 1)).sum())
    print(f"x{scale}: score pass {1e3*(t1-t0):.2f} ms (rc {rc}), max score {res['score'].max()}; top-1000 alignment call {1e3*(t3-t2):.2f} ms (rc {rc2}), "
          f"{sub.total_residues} residues, ops total {int(out['alignmentLength'].sum())}, aligned-rectangle cells {cells/1e6:.1f} M")
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libopal_ref.so")
    if os.path.exists(ref_so):
        ref = OpalCLibrary(ref_so)
        r3 = pre.copy()
        t4 = time.perf_counter()
        rc3, want = ref.search_database(q, sub, go, ge, m, 23, r3, 2, MODES["SW"])
        t5 = time.perf_counter()
        same = all((out[f] == want[f]).all() for f in ("score", "endLocationQuery", "endLocationTarget", "startLocationQuery", "startLocationTarget", "alignmentLength"))
        same = same and all((get_alignment(out, i) == get_alignment(want, i)).all() for i in range(len(sub)))
        print(f"     reference (1 thread) {1e3*(t5-t4):.1f} ms, identical records + operation strings: {same}")
        free_alignments(want)
    # the same protocol as ONE call on the resident database (opalb200_db_search_topk)
    h = eng.create_db(db, 0)
    for rep in range(3):
        t6 = time.perf_counter()
        rck, idx, topres = h.search_topk(q, go, ge, m, 23, 2, "SW", 1000)
        t7 = time.perf_counter()
        same_k = (idx == top).all() and all((topres[f] == out[f]).all() for f in ("score", "endLocationQuery", "endLocationTarget", "startLocationQuery", "startLocationTarget", "alignmentLength"))
        free_alignments(topres)
    print(f"     resident top-k pipeline (score+end over {len(db)} + top-1000 alignment): {1e3*(t7-t6):.2f} ms (rc {rck}), same records as the two-call protocol: {bool(same_k)}")
    h.close()
    free_alignments(out)
