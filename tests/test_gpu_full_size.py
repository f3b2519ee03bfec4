"""Parity at BASELINE.json's full sizes.

configs[1] (12,071 sequences, Q = 513) is small enough to compare every record with the unmodified reference
(or, without oracle/_ref, with the scalar oracle).  configs[2]'s 570k-sequence database is checked through
size-independent properties -- invariance under permutation and sharding of the database, agreement between the
search levels, NW symmetry -- plus an oracle comparison on a random sample that includes the longest targets."""
import os

import numpy as np
import pytest

from _util import (MODES, OPAL_OVERFLOW_BUCKETS, REF_SO, OpalCLibrary, SequenceDB, dump_results, free_alignments, get_alignment,
                   new_results, search_sample_parallel)
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


# ------------------------------------------------------------------ BASELINE configs[2] at size, every strip-height class
@pytest.fixture(scope="module")
def big_handle(product, big):
    sm, _, db = big
    h = product.create_db(db, 0)
    yield h
    h.close()


def _sample(db, rng, extra=()):
    """>= 500 targets: the 40 longest (the 35,213-residue one among them), 460 random, and `extra`."""
    order = np.argsort(-db.lengths.astype(np.int64), kind="stable")
    assert db.lengths[order[0]] == 35213
    return np.unique(np.concatenate([order[:40], rng.choice(len(db), 480, replace=False), np.asarray(extra, dtype=np.int64)]))


@pytest.mark.parametrize("qlen", [144, 1000, 2005, 5478])
@pytest.mark.parametrize("mode", ["NW", "HW", "OV", "SW"])
def test_config3_at_size_sampled_against_reference_and_oracle(product, oracle, big, big_handle, qlen, mode):
    """The whole 570k-sequence database is searched (score + end); >= 500 targets -- the 40 longest with the one of
    35,213 residues, the 20 best hits, 480 random ones -- are compared record by record with the unmodified reference
    where its passes are defined and with the scalar oracle elsewhere (the reference's 32-bit NW/HW/OV
    pass, reached from about Q + T > 30,000, is undefined behaviour).  Q = 144: strips of 9 rows; 1000: one pass of 32;
    2005: two passes; 5478: six passes with boundary rows through HBM and the 35,213-residue target at 32 bits."""
    sm, _, db = big
    q = [x for x in datasets.config3_queries(sm) if len(x) == qlen][0]
    rng = np.random.default_rng(qlen)
    rc, sc, eq, et, _ = big_handle.search(q, 11, 1, sm.flat(), 23, 1, mode)
    assert rc == 0, product.last_error()
    rc0, sc0, _, _, _ = big_handle.search(q, 11, 1, sm.flat(), 23, 0, mode)
    assert rc0 == 0 and np.array_equal(sc0, sc)                      # the two search levels agree on every target
    sample = _sample(db, rng, np.argsort(-sc.astype(np.int64), kind="stable")[:20])
    assert len(sample) >= 500
    use_ref = os.path.exists(REF_SO)
    defined = (db.lengths[sample].astype(np.int64) + qlen < 28000) if mode != "SW" else np.ones(len(sample), bool)
    want = new_results(len(sample))
    if use_ref and defined.any():
        rc, part = search_sample_parallel(OpalCLibrary(REF_SO), q, db, sample[defined], 11, 1, sm.flat(), 23, 1, MODES[mode])
        assert rc == 0
        want[defined] = part
    rest = ~defined if use_ref else np.ones(len(sample), bool)
    if rest.any():
        rc, part = search_sample_parallel(oracle, q, db, sample[rest], 11, 1, sm.flat(), 23, 1, MODES[mode])
        assert rc == 0
        want[rest] = part
    assert np.array_equal(want["score"], sc[sample]), (mode, qlen, sample[want["score"] != sc[sample]][:5])
    ok = want["score"] > 0 if mode == "SW" else np.ones(len(sample), bool)    # reference quirk Q2: end of a zero-score SW hit
    assert np.array_equal(want["endLocationQuery"][ok], eq[sample][ok]) and np.array_equal(want["endLocationTarget"][ok], et[sample][ok])


# ------------------------------------------------------------------ BASELINE configs[3] at size
def test_config4_top_1000_alignment_with_32_bit_rescoring_at_size(product, oracle, big, big_handle):
    """SW score pass over the 570k-sequence database, the 1000 best hits, OPAL_SEARCH_ALIGNMENT on that sub-database
    with the prefilled records (the reuse path: opalSearchDatabaseRescore) -- with 16 x BLOSUM62 and gaps 176/16, so
    that the best hits leave 16 bits and are re-run at 32.  EVERY record and operation string is compared with the
    unmodified reference (SW with gapOpen >= 2 gapExt is where its alignment stage is reliable), or
    with the oracle when oracle/_ref is absent; the one-call pipeline opalb200_db_search_topk must give the same."""
    sm, q, db = big
    m16 = (sm.matrix * 16).ravel().astype(np.int32)
    go, ge = 176, 16
    rc, sc, _, _, _ = big_handle.search(q, go, ge, m16, 23, 0, "SW")
    assert rc == 0 and sc.max() > 32767 and big_handle.last_stats()["rerun32"] > 0
    top = np.argsort(-sc.astype(np.int64), kind="stable")[:1000]
    sub = db.subset(top)
    pre = new_results(1000)
    pre["scoreSet"] = 1
    pre["score"] = sc[top]
    checker = OpalCLibrary(REF_SO) if os.path.exists(REF_SO) else oracle
    rc1, want = checker.search_database(q, sub, go, ge, m16, 23, pre.copy(), 2, MODES["SW"])
    rc2, got = product.search_database(q, sub, go, ge, m16, 23, pre.copy(), 2, MODES["SW"], entry="opalSearchDatabaseRescore")
    assert rc1 == 0 and rc2 == 0
    w, g = dump_results(want, digest=True), dump_results(got, digest=True)
    assert g == w, [(i, a, b) for i, (a, b) in enumerate(zip(g, w)) if a != b][:3]
    assert (got["alignmentLength"] > 0).sum() >= 990
    rck, idx, topres = big_handle.search_topk(q, go, ge, m16, 23, 2, "SW", 1000)
    assert rck == 0 and np.array_equal(idx, top)
    assert dump_results(topres, digest=True) == g
    for r in (want, got, topres):
        free_alignments(r)


# ------------------------------------------------------------------ BASELINE configs[4] at size
def test_config5_long_dna_targets_at_size(product, oracle):
    """DNA alphabet, 10 kb query (ten passes over the query), 100,000 sequences with Pareto lengths up to 100 kb --
    ten of them exactly 100 kb -- and planted near-copies of the query whose score leaves 16 bits.  Every target
    beyond 50 kb, every target that was re-run at 32 bits (score > 32,767 - 5) and 500 random ones are compared with
    the unmodified reference (the oracle without it), score and end location."""
    rng = np.random.default_rng(20261019)
    sm = matrices.simple(4, 5, -4)
    q = rng.integers(0, 4, 10000, dtype=np.uint8)
    db = datasets.dna_db(100000, 20261019, query=q)
    assert (db.lengths == 100000).sum() >= 10
    h = product.create_db(db, 0)
    try:
        rc, sc, eq, et, _ = h.search(q, 16, 4, sm.flat(), 4, 1, "SW")
        assert rc == 0, product.last_error()
        assert h.last_stats()["rerun32"] > 0 and sc.max() > 32767
        rc0, sc0, _, _, _ = h.search(q, 16, 4, sm.flat(), 4, 0, "SW")
        assert rc0 == 0 and np.array_equal(sc0, sc)
    finally:
        h.close()
    sample = np.unique(np.concatenate([np.nonzero(db.lengths > 50000)[0], np.nonzero(sc > 32767 - 6)[0],
                                       rng.choice(len(db), 500, replace=False)]))
    assert (db.lengths[sample] > 50000).sum() >= 10
    checker = OpalCLibrary(REF_SO) if os.path.exists(REF_SO) else oracle
    rc, want = search_sample_parallel(checker, q, db, sample, 16, 4, sm.flat(), 4, 1, MODES["SW"])
    assert rc == 0
    assert np.array_equal(want["score"], sc[sample]), sample[want["score"] != sc[sample]][:5]
    ok = want["score"] > 0
    assert np.array_equal(want["endLocationQuery"][ok], eq[sample][ok]) and np.array_equal(want["endLocationTarget"][ok], et[sample][ok])
