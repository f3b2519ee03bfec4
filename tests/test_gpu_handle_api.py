"""GPU tests of the resident-database extensions of include/opal_b200.h: the packed
("sorted") constructor, the multi-query batch and the handle form of opalSearchDatabase.  Every result is compared
with the oracle through the C ABI, and the three routes are compared with each other."""
import numpy as np
import pytest

from _util import MODES, SequenceDB, dump_results, free_alignments, new_results, search_dump
from opal_b200 import datasets, matrices

pytestmark = pytest.mark.gpu


def _random_db(rng, sm, n, lo, hi, planted=None):
    seqs = [datasets.random_residues(int(x), rng, sm) for x in rng.integers(lo, hi, n)]
    if planted is not None:
        for k in range(0, n, max(1, n // 8)):
            seqs[k] = datasets.mutate(planted, 0.7, rng, sm)
    return SequenceDB.from_sequences(seqs)


def _packed(db):
    """Longest-first packing of a SequenceDB, as opal_makedb_b200 writes it: (residues, sortedLengths, order)."""
    order = np.argsort(-db.lengths.astype(np.int64), kind="stable").astype(np.int32)
    residues = np.concatenate([db.sequence(int(i)) for i in order]) if len(order) else np.zeros(0, np.uint8)
    return residues, db.lengths[order], order


@pytest.mark.parametrize("mode", ["NW", "HW", "OV", "SW"])
def test_sorted_constructor_matches_pointer_constructor_and_oracle(product, oracle, mode):
    rng = np.random.default_rng(5)
    sm = matrices.blosum62()
    q = datasets.random_residues(130, rng, sm)
    db = _random_db(rng, sm, 333, 1, 400, planted=q)
    residues, lens, order = _packed(db)
    h1 = product.create_db(db, 0)
    h2 = product.create_db_sorted(residues, lens, order, 0)
    try:
        for st in (0, 1):
            rc1, s1, eq1, et1, _ = h1.search(q, 11, 1, sm.flat(), 23, st, mode)
            rc2, s2, eq2, et2, _ = h2.search(q, 11, 1, sm.flat(), 23, st, mode)
            assert rc1 == 0 and rc2 == 0
            assert np.array_equal(s1, s2) and np.array_equal(eq1, eq2) and np.array_equal(et1, et2)
            rc, want = search_dump(oracle, q, db, 11, 1, sm.flat(), 23, st, MODES[mode])
            assert rc == 0
            assert [int(x) for x in s2] == [w[1] for w in want]
            if st:
                assert [int(x) for x in eq2] == [w[2] for w in want]
                assert [int(x) for x in et2] == [w[3] for w in want]
    finally:
        h1.close()
        h2.close()


def test_sorted_constructor_identity_order_and_rejections(product):
    rng = np.random.default_rng(6)
    sm = matrices.simple(4, 2, -3)
    lens = np.array([9, 7, 7, 3, 0], dtype=np.int32)
    residues = rng.integers(0, 4, int(lens.sum())).astype(np.uint8)
    h = product.create_db_sorted(residues, lens, None, 0)
    db = SequenceDB(residues, np.concatenate([[0], np.cumsum(lens)]))
    h0 = product.create_db(db, 0)
    q = rng.integers(0, 4, 8).astype(np.uint8)
    a, b = h.search(q, 5, 2, sm.flat(), 4, 1, "HW"), h0.search(q, 5, 2, sm.flat(), 4, 1, "HW")
    assert a[0] == 0 and all(np.array_equal(x, y) for x, y in zip(a[1:4], b[1:4]))
    h.close()
    h0.close()
    with pytest.raises(RuntimeError, match="sorted"):
        product.create_db_sorted(residues, lens[::-1].copy(), None, 0)
    with pytest.raises(RuntimeError, match="permutation"):
        product.create_db_sorted(residues, lens, np.array([0, 1, 1, 3, 4], dtype=np.int32), 0)


@pytest.mark.parametrize("mode,in_flight", [("SW", 3), ("SW", 1), ("NW", 2), ("OV", 4), ("HW", 8)])
def test_batch_equals_single_searches_and_oracle(product, oracle, mode, in_flight):
    rng = np.random.default_rng(7)
    sm = matrices.blosum62()
    queries = [datasets.random_residues(int(n), rng, sm) for n in (33, 150, 600, 1, 97, 1300, 64)]
    db = _random_db(rng, sm, 420, 1, 700, planted=queries[2])
    h = product.create_db(db, 0)
    try:
        for st in (0, 1):
            rc, S, EQ, ET, ms = h.search_batch(queries, 11, 1, sm.flat(), 23, st, mode, in_flight=in_flight)
            assert rc == 0, product.last_error()
            assert ms > 0
            assert h.last_stats()["kernel_launches"] >= len(queries)
            for k, q in enumerate(queries):
                rc1, s, eq, et, _ = h.search(q, 11, 1, sm.flat(), 23, st, mode)
                assert rc1 == 0
                assert np.array_equal(S[k], s) and np.array_equal(EQ[k], eq) and np.array_equal(ET[k], et), (k, st)
            # one query of the batch against the oracle as well (the single-search route is covered by test_gpu_parity)
            k = 2
            rc, want = search_dump(oracle, queries[k], db, 11, 1, sm.flat(), 23, st, MODES[mode])
            assert [int(x) for x in S[k]] == [w[1] for w in want]
    finally:
        h.close()


def test_batch_reports_errors(product):
    sm = matrices.simple(4, 2, -3)
    db = SequenceDB.from_sequences([[0, 1, 2, 3], [1, 1]])
    h = product.create_db(db, 0)
    try:
        rc, *_ = h.search_batch([np.array([0, 1], np.uint8), np.array([0, 9], np.uint8)], 5, 2, sm.flat(), 4, 0, "SW")
        assert rc == 4 and "alphabetLength" in product.last_error()
        rc, S, *_ = h.search_batch([], 5, 2, sm.flat(), 4, 0, "SW")
        assert rc == 0 and S.shape == (0, 2)
    finally:
        h.close()


@pytest.mark.parametrize("mode", ["SW", "NW", "HW", "OV"])
def test_handle_records_equal_drop_in_records(product, mode):
    """opalb200_db_search_results == opalSearchDatabase on the same inputs, all three search levels, including the
    reuse of prefilled records and the alignment strings."""
    rng = np.random.default_rng(8)
    sm = matrices.blosum62()
    q = datasets.random_residues(90, rng, sm)
    db = _random_db(rng, sm, 150, 1, 260, planted=q)
    residues, lens, order = _packed(db)
    for make in (lambda: product.create_db(db, 0), lambda: product.create_db_sorted(residues, lens, order, 0)):
        h = make()
        try:
            for st in (0, 1, 2):
                rc, res = h.search_results(q, 11, 1, sm.flat(), 23, st, mode)
                assert rc == 0, product.last_error()
                got = dump_results(res)
                free_alignments(res)
                rc, want = search_dump(product, q, db, 11, 1, sm.flat(), 23, st, MODES[mode])
                assert rc == 0 and got == want, [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w][:3]
            # reuse: score + end prefilled by one call, alignment added by the next (reference src/opal.h:118-122)
            rc, res = h.search_results(q, 11, 1, sm.flat(), 23, 1, mode)
            rc, res = h.search_results(q, 11, 1, sm.flat(), 23, 2, mode, results=res)
            assert rc == 0
            got = dump_results(res)
            free_alignments(res)
            assert got == want
        finally:
            h.close()


def test_handle_records_empty_and_invalid_mode(product):
    sm = matrices.simple(4, 2, -3)
    h = product.create_db(SequenceDB.from_sequences([[0, 1, 2, 3]]), 0)
    try:
        res = new_results(1)
        rc, res = h.search_results(np.array([0, 1], np.uint8), 5, 2, sm.flat(), 4, 1, 17, results=res)
        assert rc == 3 and res["scoreSet"][0] == 0  # OPAL_ERR_INVALID_MODE, results untouched
    finally:
        h.close()


@pytest.mark.parametrize("mode", ["SW", "NW", "HW", "OV"])
def test_topk_pipeline_equals_the_two_call_protocol(product, mode):
    """BASELINE configs[3] in one call: score -> top-k -> alignment on the resident database must give the records
    the reference protocol gives (score+end over the database, top k by (score desc, index asc), ALIGNMENT over the
    k-entry sub-database with the prefilled records)."""
    rng = np.random.default_rng(9)
    sm = matrices.blosum62()
    q = datasets.random_residues(110, rng, sm)
    db = _random_db(rng, sm, 300, 1, 330, planted=q)
    k = 37
    rc, res = product.search_database(q, db, 11, 1, sm.flat(), 23, None, 1, MODES[mode])
    assert rc == 0
    order = sorted(range(len(db)), key=lambda i: (-int(res["score"][i]), i))[:k]
    pre = res[order].copy()
    rc, res2 = product.search_database(q, db.subset(order), 11, 1, sm.flat(), 23, pre, 2, MODES[mode])
    assert rc == 0
    want = dump_results(res2)
    free_alignments(res2)
    h = product.create_db(db, 0)
    try:
        rc, idx, top = h.search_topk(q, 11, 1, sm.flat(), 23, 2, mode, k)
        assert rc == 0, product.last_error()
        assert idx.tolist() == order
        got = dump_results(top)
        free_alignments(top)
        assert got == want, [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w][:3]
        # lower search levels: the same records as opalSearchDatabase gives the selected entries
        for st in (0, 1):
            rc, idx, top = h.search_topk(q, 11, 1, sm.flat(), 23, st, mode, k)
            rc2, full = product.search_database(q, db, 11, 1, sm.flat(), 23, None, st, MODES[mode])
            assert rc == 0 and rc2 == 0 and idx.tolist() == order
            assert dump_results(top) == [dump_results(full)[i] for i in order]
        rc, idx, top = h.search_topk(q, 11, 1, sm.flat(), 23, 1, mode, 10 ** 6)  # k beyond the database
        assert rc == 0 and len(idx) == len(db) and sorted(idx.tolist()) == list(range(len(db)))
        rc, idx, top = h.search_topk(q, 11, 1, sm.flat(), 23, 1, mode, 0)
        assert rc == 0 and len(idx) == 0
    finally:
        h.close()


@pytest.mark.parametrize("mode", ["SW", "NW", "HW", "OV"])
def test_topk_selection_on_the_device_ties_empties_and_reruns(product, mode):
    """The device-side selection (radix select over score | ~index): tie-heavy scores (tiny alphabet), zero-length
    targets, targets that are re-run at 32 bits (scaled matrix, copies of the query) and every k from 1 to n."""
    rng = np.random.default_rng(19)
    a = 3
    m = (matrices.simple(a, 2, -1).matrix * 300).ravel().astype(np.int32)
    q = rng.integers(0, a, 150).astype(np.uint8)
    seqs = [rng.integers(0, a, int(n)).astype(np.uint8) for n in rng.integers(0, 40, 260)]
    for i in (3, 17, 101):
        seqs[i] = np.zeros(0, np.uint8)
    seqs[50] = q.copy()                                 # score 150 * 600 = 90,000: beyond 16 bits
    seqs[51] = np.concatenate([q[:100], q[:100]])
    for i in range(200, 230):
        seqs[i] = seqs[199].copy()                      # 31 identical targets: equal scores, ordered by index
    db = SequenceDB.from_sequences(seqs)
    rc, full = product.search_database(q, db, 900, 300, m, a, None, 1, MODES[mode])
    assert rc == 0
    order = sorted(range(len(db)), key=lambda i: (-int(full["score"][i]), i))
    ref = dump_results(full)
    h = product.create_db(db, 0)
    try:
        for k in (1, 2, 31, 64, 229, len(db) - 1, len(db)):
            rc, idx, top = h.search_topk(q, 900, 300, m, a, 1, mode, k)
            assert rc == 0, product.last_error()
            assert idx.tolist() == order[:k], (mode, k)
            assert dump_results(top) == [ref[i] for i in order[:k]]
    finally:
        h.close()
