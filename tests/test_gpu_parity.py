"""GPU parity: libopal_b200.so (through its C ABI) against the oracle and the committed golden vectors.

Everything here is bit-exact integer comparison of every OpalSearchResult field.
"""
import numpy as np
import pytest

from _util import (MODES, OPAL_OVERFLOW_BUCKETS, OPAL_OVERFLOW_SIMPLE, README_DB, README_MATRIX, README_QUERY,
                   SequenceDB, dump_results, glibc_testcpp_data, new_results, search_dump)
from opal_b200 import datasets, matrices
from test_oracle_golden import fix_sw_zero, golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["NW", "HW", "OV", "SW"])
@pytest.mark.parametrize("search_type", [0, 1])
def test_readme_example(product, mode, search_type):
    g = golden("readme.json")
    db = SequenceDB.from_sequences(README_DB)
    for ovf in (OPAL_OVERFLOW_SIMPLE, OPAL_OVERFLOW_BUCKETS):
        rc, d = search_dump(product, README_QUERY, db, 3, 1, README_MATRIX, 4, search_type, MODES[mode], ovf)
        want = g[f"{mode}/{search_type}/{ovf}"]
        assert rc == want["rc"] == 0
        assert d == fix_sw_zero(want["results"], mode, search_type)


@pytest.mark.parametrize("key", ["SW/0", "SW/1", "NW/1", "HW/1", "OV/1"])
def test_config1_golden(product, key):
    g = golden("config1.json")
    b50 = matrices.blosum50()
    mode, st = key.split("/")
    rc, d = search_dump(product, np.array(g["query"], dtype=np.uint8), SequenceDB.from_sequences(g["db"]),
                        3, 1, b50.flat(), 24, int(st), MODES[mode], OPAL_OVERFLOW_BUCKETS)
    assert rc == 0
    assert d == fix_sw_zero(g[key]["results"], mode, int(st))


@pytest.mark.parametrize("key", ["NW/0", "NW/1", "HW/0", "HW/1", "OV/0", "OV/1", "SW/0", "SW/1"])
def test_protein_golden(product, key):
    g = golden("protein.json")
    b62 = matrices.blosum62()
    mode, st = key.split("/")
    rc, d = search_dump(product, np.array(g["query"], dtype=np.uint8), SequenceDB.from_sequences(g["db"]),
                        11, 1, b62.flat(), 23, int(st), MODES[mode], OPAL_OVERFLOW_BUCKETS)
    assert rc == 0
    assert d == fix_sw_zero(g[key]["results"], mode, int(st))


@pytest.mark.parametrize("mode,maximum", [("SW", 573), ("NW", 460), ("HW", 567), ("OV", 567)])
def test_reference_selftest_inputs(product, mode, maximum):
    """Inputs of the reference's ./test (src/test.cpp:35-99): score + end of all 200 targets."""
    g = golden("testcpp.json")[mode]
    q, db = glibc_testcpp_data()
    m = matrices.simple(4, 3, -1).flat()
    rc, d = search_dump(product, q, db, 11, 1, m, 4, 1, MODES[mode], OPAL_OVERFLOW_SIMPLE)
    assert rc == 0
    assert max(r[1] for r in d) == maximum
    assert [r[:4] for r in d] == [r[:4] for r in g["results"]]


def _compare(product, oracle, q, db, go, ge, m, a, modes=("NW", "HW", "OV", "SW"), types=(0, 1)):
    for mode in modes:
        for st in types:
            rc1, want = search_dump(oracle, q, db, go, ge, m, a, st, MODES[mode])
            rc2, got = search_dump(product, q, db, go, ge, m, a, st, MODES[mode])
            assert rc1 == rc2 == 0, (mode, st, rc1, rc2)
            bad = [i for i in range(len(want)) if want[i] != got[i]]
            assert not bad, (mode, st, bad[:5], [want[i] for i in bad[:3]], [got[i] for i in bad[:3]])


@pytest.mark.parametrize("seed", range(6))
def test_random_small_vs_oracle(product, oracle, seed):
    """Short / ragged / tie-heavy inputs: lengths from 1, tiny alphabets, zero gap extension."""
    rng = np.random.default_rng(100 + seed)
    a = int(rng.integers(2, 6))
    m = matrices.simple(a, int(rng.integers(1, 5)), -int(rng.integers(0, 4))).flat()
    go, ge = int(rng.integers(0, 8)), int(rng.integers(0, 3))
    q = rng.integers(0, a, int(rng.integers(1, 80))).astype(np.uint8)
    seqs = [rng.integers(0, a, int(n)).astype(np.uint8) for n in rng.integers(1, 120, 101)]
    _compare(product, oracle, q, SequenceDB.from_sequences(seqs), go, ge, m, a)


@pytest.mark.parametrize("qlen", [1, 7, 8, 9, 31, 32, 33, 64, 65, 127, 128, 129, 255, 256, 257, 513, 1024, 1025])
def test_query_lengths_around_strip_boundaries(product, oracle, qlen):
    rng = np.random.default_rng(qlen)
    sm = matrices.blosum62()
    q = datasets.random_residues(qlen, rng, sm)
    seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(1, 200, 37)]
    seqs[3] = datasets.mutate(q, 0.8, rng, sm)
    seqs[11] = q.copy()
    _compare(product, oracle, q, SequenceDB.from_sequences(seqs), 11, 1, sm.flat(), 23)


@pytest.mark.parametrize("qlen", [1500, 2600])
def test_multi_pass_queries(product, oracle, qlen):
    """Queries longer than one pass of G*R rows: the boundary row goes through HBM between passes."""
    rng = np.random.default_rng(qlen)
    sm = matrices.blosum62()
    q = datasets.random_residues(qlen, rng, sm)
    seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(1, 400, 21)]
    seqs[2] = datasets.mutate(q, 0.7, rng, sm)
    seqs[5] = q[700:1900].copy()
    _compare(product, oracle, q, SequenceDB.from_sequences(seqs), 11, 1, sm.flat(), 23)


def test_overflow_escalation_to_32_bit(product, oracle):
    """Scores beyond the 16-bit lanes: scaled matrix (8 x BLOSUM62, gaps 88/8) with near-identical targets,
    the analogue of the reference's char -> short -> int ladder (src/opal.cpp:512-530)."""
    rng = np.random.default_rng(5)
    sm = matrices.blosum62()
    q = datasets.random_residues(900, rng, sm)
    seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(50, 600, 40)]
    seqs[0] = q.copy()
    seqs[7] = datasets.mutate(q, 0.95, rng, sm)
    seqs[20] = np.concatenate([seqs[20], q, seqs[21]])
    db = SequenceDB.from_sequences(seqs)
    m8 = (sm.matrix * 8).ravel()
    rc, want = search_dump(oracle, q, db, 88, 8, m8, 23, 1, MODES["SW"])
    assert rc == 0 and max(r[1] for r in want) > 32767
    _compare(product, oracle, q, db, 88, 8, m8, 23)


def test_global_modes_route_long_targets_to_32_bit(product, oracle):
    """NW first-row values -gapOpen - c*gapExt leave the 16-bit range for long targets:
    the correct 32-bit result is expected, not the reference's undefined behaviour."""
    rng = np.random.default_rng(9)
    a = 4
    m = matrices.simple(a, 3, -1).flat()
    q = rng.integers(0, a, 300).astype(np.uint8)
    seqs = [rng.integers(0, a, n).astype(np.uint8) for n in (40000, 33000, 29000, 150, 90, 1)]
    _compare(product, oracle, q, SequenceDB.from_sequences(seqs), 11, 1, m, a, modes=("NW", "HW", "OV"), types=(1,))


def test_reuse_rule_and_rescore_entry(product):
    g = golden("api.json")
    db = SequenceDB.from_sequences(README_DB)
    res = new_results(4)
    args = (README_QUERY, db, 3, 1, README_MATRIX, 4, res)
    product.search_database(*args, 0, MODES["SW"], OPAL_OVERFLOW_SIMPLE)
    res["score"][1] = 9999
    product.search_database(*args, 0, MODES["SW"], OPAL_OVERFLOW_SIMPLE)
    assert dump_results(res) == g["reuse_score_then_score"]
    product.search_database(*args, 1, MODES["SW"], OPAL_OVERFLOW_SIMPLE, entry="opalSearchDatabaseRescore")
    assert dump_results(res) == g["reuse_then_score_end"]


def test_char_sw_and_invalid_mode(product):
    g = golden("api.json")
    c = g["char_sw"]
    rc, res = product.search_database_char_sw(np.array(c["query"], dtype=np.uint8), SequenceDB.from_sequences(c["db"]),
                                              3, 1, np.array(c["matrix"], dtype=np.int32), 4)
    assert rc == c["rc"] == 1
    assert dump_results(res, with_alignment=False) == c["results"]
    db = SequenceDB.from_sequences(README_DB)
    res = new_results(4)
    rc, res = product.search_database(README_QUERY, db, 3, 1, README_MATRIX, 4, res, 0, 7, OPAL_OVERFLOW_SIMPLE)
    assert rc == 3
    assert dump_results(res) == g["invalid_mode"]["results"]


@pytest.mark.parametrize("mode", ["NW", "HW", "OV"])
def test_range_tracking_flags_scores_that_leave_16_bits(product, oracle, mode):
    """NW / HW / OV at 16 bits are guarded by sampled range tracking instead of an a-priori bound: targets whose cells
    come near the ends of the 16-bit range -- upwards (scaled matrix, near-identical sequences) or downwards (a long
    query against short targets, very long gaps) -- must be re-run at 32 bits, everything else stays at 16."""
    rng = np.random.default_rng(12)
    sm = matrices.blosum62()
    # upwards: 6 x BLOSUM62 on a 1,300-residue query and copies of it (true scores up to ~45,000)
    q = datasets.random_residues(1300, rng, sm)
    seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(50, 900, 30)]
    seqs[0] = q.copy()
    seqs[1] = datasets.mutate(q, 0.97, rng, sm)
    seqs[2] = q[:700].copy()                       # ~24,000: inside 16 bits but near enough to the limit region
    seqs[3] = datasets.mutate(q, 0.6, rng, sm)
    m6 = (sm.matrix * 6).ravel()
    db = SequenceDB.from_sequences(seqs)
    rc, want = search_dump(oracle, q, db, 66, 6, m6, 23, 1, MODES[mode])
    assert rc == 0 and max(r[1] for r in want) > 32767
    _compare(product, oracle, q, db, 66, 6, m6, 23, modes=(mode,))
    # downwards: queries of 12,000 ... 27,000 against short targets (NW / HW scores around -Q) and unequal pairs
    for qlen in (12000, 16300, 21000, 27000):
        ql = datasets.random_residues(qlen, rng, sm)
        seqs = [datasets.random_residues(int(n), rng, sm) for n in (700, 650, 300, 299, 90, 40, 5, 1)]
        _compare(product, oracle, ql, SequenceDB.from_sequences(seqs), 11, 1, sm.flat(), 23, modes=(mode,), types=(1,))
    # long targets against a short query: NW's first row runs down to -T
    qs = datasets.random_residues(200, rng, sm)
    seqs = [datasets.random_residues(int(n), rng, sm) for n in (27500, 26000, 16500, 16000, 9000, 8999, 120)]
    _compare(product, oracle, qs, SequenceDB.from_sequences(seqs), 11, 1, sm.flat(), 23, modes=(mode,), types=(1,))
