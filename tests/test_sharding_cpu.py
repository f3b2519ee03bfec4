"""N > 1 host logic on CPU: two gloo ranks each search a residue-balanced shard (with the oracle standing in
for the device) and the merged result must equal the unsharded search.  Also the C-ABI export check."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np

from _util import MODES, ORACLE_SO, PRODUCT_SO, ROOT, OpalCLibrary, SequenceDB
from opal_b200 import datasets, matrices, sharding

WORKER = r"""
import os, sys, pickle
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from _util import MODES, ORACLE_SO, OpalCLibrary
from opal_b200 import datasets, matrices, sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sm = matrices.blosum62()
rng = np.random.default_rng(11)
q = datasets.random_residues(120, rng, sm)
db = datasets.protein_db(301, 5, sm, query=q, homolog_fraction=0.05)
shards = sharding.deal_shards(db.lengths, world)
mine = sharding.shard_db(db, shards[rank])
lib = OpalCLibrary(ORACLE_SO)
rc, res = lib.search_database(q, mine, 11, 1, sm.flat(), 23, None, 1, MODES[{mode!r}])
assert rc == 0
gathered = [None] * world
dist.all_gather_object(gathered, (shards[rank], res))
if rank == 0:
    merged = sharding.merge_results(len(db), [g[0] for g in gathered], [g[1] for g in gathered])
    rc, full = lib.search_database(q, db, 11, 1, sm.flat(), 23, None, 1, MODES[{mode!r}])
    for f in ("scoreSet", "score", "endLocationQuery", "endLocationTarget"):
        assert (merged[f] == full[f]).all(), f
    sizes = [int(db.lengths[s].sum()) for s in shards]
    assert max(sizes) - min(sizes) <= db.lengths.max(), sizes
    print("MERGE_OK", sizes)
dist.destroy_process_group()
"""


def test_two_rank_sharded_search_equals_unsharded(tmp_path):
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    for mode in ("SW", "NW"):
        script = tmp_path / f"worker_{mode}.py"
        script.write_text(WORKER.format(root=ROOT, mode=mode))
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "MERGE_OK" in out.stdout


def test_deal_shards_balances_residues():
    rng = np.random.default_rng(3)
    lengths = datasets.lognormal_lengths(5000, rng)
    for world in (1, 2, 4, 8):
        shards = sharding.deal_shards(lengths, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(5000))
        sums = [int(lengths[s].sum()) for s in shards]
        assert max(sums) - min(sums) <= int(lengths.max())


def test_c_abi_exports_every_declared_symbol():
    """The product library loads without a GPU and exports every function include/*.h declares."""
    assert os.path.exists(PRODUCT_SO), "build the library first: make -C opal_b200/csrc (or __graft_entry__.build())"
    lib = ctypes.CDLL(PRODUCT_SO)
    declared = set()
    for header in ("opal.h", "opal_b200.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(opal[A-Za-z0-9_]+)\s*\(", text))
    assert {"opalSearchDatabase", "opalSearchDatabaseCharSW", "opalSearchDatabaseRescore", "opalInitSearchResult",
            "opalSearchResultIsEmpty", "opalSearchResultSetScore", "opalb200_db_create", "opalb200_db_search",
            "opalb200_db_create_sorted", "opalb200_db_search_batch", "opalb200_db_search_results",
            "opalb200_db_search_topk"} <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/ but not exported"


def test_product_fails_loudly_without_a_device():
    """No CPU fallback: on a box without CUDA devices a search returns OPAL_ERR_NO_SIMD_SUPPORT."""
    from opal_b200.handle import OpalB200
    eng = OpalB200()
    if eng.device_count() > 0:
        return
    sm = matrices.simple(4, 3, -1)
    db = SequenceDB.from_sequences([[0, 1, 2, 3]])
    rc, _ = eng.search_database(np.array([0, 1], dtype=np.uint8), db, 3, 1, sm.flat(), 4)
    assert rc == 2 and "CUDA" in eng.last_error()
