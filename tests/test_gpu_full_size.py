"""Parity at BASELINE.json's full sizes.

configs[1] (12,071 sequences, Q = 513) is small enough to compare every record with the unmodified reference
(or, without oracle/_ref, with the scalar oracle).  configs[2]'s 570k-sequence database is checked through
size-independent properties -- invariance under permutation and sharding of the database, agreement between the
search levels, NW symmetry -- plus an oracle comparison on a random sample that includes the longest targets."""
import os

import numpy as np
import pytest

from _util import MODES, OPAL_OVERFLOW_BUCKETS, REF_SO, OpalCLibrary, SequenceDB, new_results
from opal_b200 import datasets, matrices, sharding

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["SW", "NW", "HW", "OV"])
def test_config2_every_record(product, oracle, mode):
    sm = matrices.blosum62()
    q = sm.encode(datasets.P18080)
    db = datasets.config2_db(sm, q)
    checker = OpalCLibrary(REF_SO) if os.path.exists(REF_SO) else oracle
    rc1, want = checker.search_database(q, db, 11, 1, sm.flat(), 23, None, 1, MODES[mode], OPAL_OVERFLOW_BUCKETS)
    rc2, got = product.search_database(q, db, 11, 1, sm.flat(), 23, None, 1, MODES[mode], OPAL_OVERFLOW_BUCKETS)
    assert rc1 == rc2 == 0
    for f in ("scoreSet", "score", "endLocationQuery", "endLocationTarget"):
        w = want[f].copy()
        if mode == "SW" and f.startswith("end"):
            w[want["score"] == 0] = -1  # reference quirk Q2
        assert (got[f] == w).all(), (mode, f, int((got[f] != w).sum()))


@pytest.fixture(scope="module")
def big():
    sm = matrices.blosum62()
    q = sm.encode(datasets.P18080)
    return sm, q, datasets.config3_db(sm, query=q)


def test_config3_properties(product, oracle, big):
    sm, q, db = big
    assert len(db) == 570000 and abs(db.total_residues / 206e6 - 1) < 0.02
    h = product.create_db(db, 0)
    rc, sc1, eq1, et1, _ = h.search(q, 11, 1, sm.flat(), 23, 1, "SW")
    rc0, sc0, _, _, _ = h.search(q, 11, 1, sm.flat(), 23, 0, "SW")
    assert rc == rc0 == 0
    assert (sc0 == sc1).all()                                   # search levels agree
    assert ((eq1 >= 0) == (sc1 > 0)).all() and (et1 < db.lengths).all() and (eq1 < len(q)).all()
    h.close()
    # permutation + sharding invariance: two residue-balanced shards, shuffled, merged by index
    rng = np.random.default_rng(0)
    merged = np.zeros(len(db), dtype=np.int32)
    merged_et = np.zeros(len(db), dtype=np.int32)
    for idx in sharding.deal_shards(db.lengths, 2):
        idx = rng.permutation(idx)
        hs = product.create_db(sharding.shard_db(db, idx), 0)
        rc, s, e, t, _ = hs.search(q, 11, 1, sm.flat(), 23, 1, "SW")
        assert rc == 0
        merged[idx], merged_et[idx] = s, t
        hs.close()
    assert (merged == sc1).all() and (merged_et == et1).all()
    # oracle on a sample: the 40 longest, the 40 best hits and 200 random targets
    order = np.argsort(-db.lengths, kind="stable")
    sample = np.unique(np.concatenate([order[:40], np.argsort(-sc1, kind="stable")[:40], rng.integers(0, len(db), 200)]))
    sub = db.subset(sample)
    rc, want = oracle.search_database(q, sub, 11, 1, sm.flat(), 23, None, 1, MODES["SW"])
    assert rc == 0
    assert (want["score"] == sc1[sample]).all()
    ok = want["score"] > 0
    assert (want["endLocationQuery"][ok] == eq1[sample][ok]).all() and (want["endLocationTarget"][ok] == et1[sample][ok]).all()


def test_config3_global_modes_sample_and_nw_symmetry(product, oracle, big):
    sm, _, db = big
    rng = np.random.default_rng(5)
    q = datasets.config3_queries(sm)[3]  # Q = 375
    h = product.create_db(db, 0)
    order = np.argsort(-db.lengths, kind="stable")
    sample = np.unique(np.concatenate([order[:12], rng.integers(0, len(db), 150)]))
    sub = db.subset(sample)
    for mode in ("NW", "HW", "OV"):
        rc, sc, eq, et, _ = h.search(q, 11, 1, sm.flat(), 23, 1, mode)
        assert rc == 0
        rc, want = oracle.search_database(q, sub, 11, 1, sm.flat(), 23, None, 1, MODES[mode])
        assert rc == 0
        assert (want["score"] == sc[sample]).all(), mode
        assert (want["endLocationQuery"] == eq[sample]).all() and (want["endLocationTarget"] == et[sample]).all(), mode
        if mode == "NW":
            nw = sc
    h.close()
    # NW with a symmetric matrix is symmetric in its arguments: swap roles for a few targets
    for i in sample[:6]:
        t = db.sequence(int(i))
        one = SequenceDB.from_sequences([q])
        rc, res = product.search_database(t, one, 11, 1, sm.flat(), 23, None, 0, MODES["NW"])
        assert rc == 0 and int(res["score"][0]) == int(nw[i])
