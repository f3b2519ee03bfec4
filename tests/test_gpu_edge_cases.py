"""GPU edge cases the reference's tests never see (SURVEY.md section 4 "coverage gap") plus BASELINE configs[3]/[4]
in miniature: ragged / empty inputs, prefilled results, argument ranges, long targets, top-k alignment."""
import numpy as np
import pytest

from _util import (MODES, OPAL_OVERFLOW_BUCKETS, OPAL_OVERFLOW_SIMPLE, SequenceDB, dump_results, free_alignments,
                   get_alignment, new_results, search_dump)
from opal_b200 import datasets, matrices

pytestmark = pytest.mark.gpu


def _same(product, oracle, q, db, go, ge, m, a, mode, st, results_p=None, results_o=None):
    rc1, want = search_dump(oracle, q, db, go, ge, m, a, st, MODES[mode], results=results_o)
    rc2, got = search_dump(product, q, db, go, ge, m, a, st, MODES[mode], results=results_p)
    assert rc1 == rc2, (rc1, rc2)
    assert got == want, [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w][:3]


@pytest.mark.parametrize("mode", ["NW", "HW", "OV", "SW"])
def test_zero_length_targets_and_single_residues(product, oracle, mode):
    sm = matrices.simple(4, 2, -3)
    q = np.array([0, 1, 2, 3, 0, 1], dtype=np.uint8)
    db = SequenceDB.from_sequences([[], [0], [1, 2], [], [3, 3, 3, 3, 3, 3, 3], [0, 1, 2, 3, 0, 1]])
    for st in (0, 1):
        _same(product, oracle, q, db, 5, 2, sm.flat(), 4, mode, st)


def test_empty_database_and_invalid_arguments(product):
    sm = matrices.simple(4, 2, -3)
    q = np.array([0, 1], dtype=np.uint8)
    rc, res = product.search_database(q, SequenceDB.from_sequences([]), 5, 2, sm.flat(), 4)
    assert rc == 0 and len(res) == 0
    db = SequenceDB.from_sequences([[0, 1, 2]])
    big = sm.flat().copy()
    big[0] = 2 ** 30  # outside (INT_MIN/2, INT_MAX/2): reference src/opal.cpp:183-198 -> OPAL_ERR_OVERFLOW
    rc, _ = product.search_database(q, db, 5, 2, big, 4)
    assert rc == 1
    rc, _ = product.search_database(np.array([0, 7], dtype=np.uint8), db, 5, 2, sm.flat(), 4)
    assert rc == 4 and "alphabetLength" in product.last_error()  # OPAL_ERR_INVALID_ARGUMENT instead of the reference's wild read
    rc, _ = product.search_database(q, db, -1, 2, sm.flat(), 4)
    assert rc == 4 and "gap" in product.last_error()
    rc, _ = product.search_database(q, SequenceDB.from_sequences([[0, 1, 5]]), 5, 2, sm.flat(), 4)
    assert rc == 4 and "database" in product.last_error()


@pytest.mark.parametrize("mode", ["SW", "NW"])
def test_prefilled_results_are_skipped_even_beyond_1024(product, oracle, mode):
    """The documented reuse rule (src/opal.h:118-122); the reference's BUCKETS path mis-indexes its skip mask
    beyond 1024 entries, the documented semantic is what is implemented."""
    rng = np.random.default_rng(21)
    sm = matrices.blosum62()
    q = datasets.random_residues(60, rng, sm)
    db = SequenceDB.from_sequences([datasets.random_residues(int(n), rng, sm) for n in rng.integers(5, 90, 2100)])
    rp, ro = new_results(len(db)), new_results(len(db))
    for r in (rp, ro):
        for i in (0, 7, 1500, 2099):
            r["scoreSet"][i] = 1
            r["score"][i] = 7777
    _same(product, oracle, q, db, 11, 1, sm.flat(), 23, mode, 0, rp, ro)
    assert rp["score"][1500] == 7777


def test_matrix_beyond_16_bits_goes_straight_to_32(product, oracle):
    rng = np.random.default_rng(4)
    a = 4
    m = matrices.simple(a, 40000, -30000).flat()
    q = rng.integers(0, a, 50).astype(np.uint8)
    db = SequenceDB.from_sequences([rng.integers(0, a, int(n)).astype(np.uint8) for n in rng.integers(1, 80, 33)])
    for mode in ("SW", "NW", "HW", "OV"):
        _same(product, oracle, q, db, 50000, 3000, m, a, mode, 1)


def test_score_that_leaves_the_supported_range_reports_overflow(product):
    a = 2
    m = matrices.simple(a, (1 << 28) - 1, -5).flat()
    q = np.zeros(40, dtype=np.uint8)
    db = SequenceDB.from_sequences([np.zeros(40, dtype=np.uint8)])
    rc, _ = product.search_database(q, db, 3, 1, m, a, None, 0, MODES["SW"])
    assert rc == 1  # OPAL_ERR_OVERFLOW (src/opal.h:17)


@pytest.mark.parametrize("alen", [1, 2, 64, 200])
def test_alphabet_sizes(product, oracle, alen):
    rng = np.random.default_rng(alen)
    m = rng.integers(-6, 9, (alen, alen)).astype(np.int32)
    q = rng.integers(0, alen, 77).astype(np.uint8)
    db = SequenceDB.from_sequences([rng.integers(0, alen, int(n)).astype(np.uint8) for n in rng.integers(1, 150, 41)])
    for mode in ("SW", "OV"):
        _same(product, oracle, q, db, 7, 2, m.ravel(), alen, mode, 1)


def test_zero_gap_penalties(product, oracle):
    rng = np.random.default_rng(8)
    a = 3
    m = matrices.simple(a, 2, -2).flat()
    q = rng.integers(0, a, 30).astype(np.uint8)
    db = SequenceDB.from_sequences([rng.integers(0, a, int(n)).astype(np.uint8) for n in rng.integers(1, 50, 25)])
    for mode in ("SW", "NW", "HW", "OV"):
        for go, ge in ((0, 0), (2, 0), (0, 1)):
            _same(product, oracle, q, db, go, ge, m, a, mode, 1)


def test_long_dna_targets_config5_in_miniature(product, oracle):
    """BASELINE configs[4] shape: DNA alphabet, long query, heavy-tailed target lengths; one planted near-copy of the
    query pushes SW past 16 bits (score > 32767 needs Q >= 16384 at +2, so the matrix is +5/-4 here)."""
    rng = np.random.default_rng(19)
    m = matrices.simple(4, 5, -4).flat()
    q = rng.integers(0, 4, 8000).astype(np.uint8)
    lens = [20000, 9000, 8000, 2500, 700, 300, 200, 60, 9]
    seqs = [rng.integers(0, 4, n).astype(np.uint8) for n in lens]
    seqs[1][500:8500] = np.where(rng.random(8000) < 0.97, q, rng.integers(0, 4, 8000))
    db = SequenceDB.from_sequences(seqs)
    rc, want = search_dump(oracle, q, db, 5, 2, m, 4, 1, MODES["SW"])
    assert max(r[1] for r in want) > 32767
    rc2, got = search_dump(product, q, db, 5, 2, m, 4, 1, MODES["SW"])
    assert rc == rc2 == 0 and got == want


def test_top_hits_alignment_with_32_bit_rescoring_config4_in_miniature(product, oracle):
    """BASELINE configs[3] shape: score pass, top-k by score, OPAL_SEARCH_ALIGNMENT on the sub-database with the
    prefilled results (reuse path), 8 x BLOSUM62 with gaps 88/8 so that some hits exceed 16 bits."""
    rng = np.random.default_rng(23)
    sm = matrices.blosum62()
    m8 = (sm.matrix * 8).ravel()
    q = datasets.random_residues(1100, rng, sm)
    seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(30, 500, 400)]
    for i in range(0, 400, 25):
        seqs[i] = datasets.mutate(q, float(rng.uniform(0.5, 0.98)), rng, sm)
    seqs[3] = q.copy()
    db = SequenceDB.from_sequences(seqs)
    rc, res = product.search_database(q, db, 88, 8, m8, 23, None, 0, MODES["SW"], OPAL_OVERFLOW_BUCKETS)
    assert rc == 0 and res["score"].max() > 32767
    top = np.argsort(-res["score"].astype(np.int64), kind="stable")[:50]
    sub = db.subset(top)
    pre_p, pre_o = new_results(50), new_results(50)
    for pre in (pre_p, pre_o):
        pre["scoreSet"] = 1
        pre["score"] = res["score"][top]
    # end locations are missing, so they are recomputed; then start + alignment are added
    rc1, want = search_dump(oracle, q, sub, 88, 8, m8, 23, 2, MODES["SW"], results=pre_o)
    rc2, got = search_dump(product, q, sub, 88, 8, m8, 23, 2, MODES["SW"], results=pre_p, entry="opalSearchDatabaseRescore")
    assert rc1 == rc2 == 0
    assert got == want


@pytest.mark.parametrize("mode,go,ge", [("SW", 10, 2), ("NW", 10, 2), ("HW", 10, 2), ("OV", 10, 2), ("SW", 16, 4)])
def test_config5_miniature_long_dna_query_many_passes(product, oracle, mode, go, ge):
    """BASELINE configs[4] in miniature: DNA alphabet, a query far longer than one strip of rows (many passes over
    the query with boundary rows in HBM), heavy-tailed target lengths, and planted near-copies of the query whose
    score leaves 16 bits (32-bit re-run for SW, a-priori 32-bit routing for the global modes).  With gaps 10/2 random
    DNA scores grow linearly with length (every SW end location needs the exact re-sweep); 16/4 is the logarithmic
    regime tools/config5_probe.py measures."""
    rng = np.random.default_rng(55)
    sm = matrices.simple(4, 5, -4)
    q = rng.integers(0, 4, 7200, dtype=np.uint8)
    db = datasets.dna_db(48, 55, query=q, xmin=60, max_len=9000, n_at_max=2, n_planted=3)
    for st in (0, 1):
        rc1, want = search_dump(oracle, q, db, go, ge, sm.flat(), 4, st, MODES[mode])
        rc2, got = search_dump(product, q, db, go, ge, sm.flat(), 4, st, MODES[mode])
        assert rc1 == rc2 == 0
        assert got == want, [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w][:3]
    if mode == "SW":
        assert max(w[1] for w in want) > 32767  # the planted copies really leave the 16-bit range


@pytest.mark.parametrize("mode", ["HW", "OV", "NW"])
def test_long_query_against_unequal_short_pair_16_bit_pad_columns(product, oracle, mode):
    """a pair of unequal lengths sweeps pad columns for its shorter member; with H below -16384 the pad
    cell (diag - 16384) used to wrap into a large positive last-row value at 16 bits.  Q * gapExt > 17000 here."""
    rng = np.random.default_rng(77)
    sm = matrices.blosum62()
    q = datasets.random_residues(20000, rng, sm)
    seqs = [datasets.random_residues(n, rng, sm) for n in (500, 400, 900, 350, 120, 119, 64, 3)]
    db = SequenceDB.from_sequences(seqs)
    for st in (0, 1):
        _same(product, oracle, q, db, 11, 1, sm.flat(), 23, mode, st)


@pytest.mark.parametrize("top", [254, 255])
def test_alphabet_255_and_256(product, oracle, top):
    """Any unsigned char alphabet (reference src/opal.h:96-98): alphabetLength 255 and 256, with and without the
    residue code 255 (which cannot ride in the 16-bit streams and is searched by the 32-bit class)."""
    rng = np.random.default_rng(5 + top)
    for A in (255, 256):
        if top >= A:
            continue
        matrix = rng.integers(-4, 3, (A, A)).astype(np.int32)
        matrix[np.arange(A), np.arange(A)] = 5
        q = rng.integers(0, A, 90).astype(np.uint8)
        q[3] = A - 1
        seqs = [rng.integers(0, top + 1, int(n)).astype(np.uint8) for n in rng.integers(1, 300, 60)]
        seqs[7][0] = top
        seqs[11] = q[10:80].copy()
        db = SequenceDB.from_sequences(seqs)
        for mode in ("SW", "NW", "HW", "OV"):
            _same(product, oracle, q, db, 7, 2, matrix.ravel(), A, mode, 1)
