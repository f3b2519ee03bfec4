"""Chained passes (SearchParams::chain): for the longest targets of a database every pass over a long query runs on a
warp of its own, all at once, the boundary row handed from pass to pass through L2 with progress marks.  Forced here on
small databases (OPAL_B200_CHAIN) in every mode and flavor, compared with the oracle and with the unchained plan."""
import numpy as np
import pytest

from _util import MODES, SequenceDB, search_dump
from opal_b200 import datasets, matrices

pytestmark = pytest.mark.gpu


def _db(rng, sm, q, longest=(2600, 2300, 1900, 1500, 1200, 1100, 900, 800)):
    seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(1, 300, 180)]
    for k, n in enumerate(longest):
        seqs[3 * k] = datasets.random_residues(n, rng, sm)
    seqs[1] = np.concatenate([datasets.random_residues(400, rng, sm), datasets.mutate(q, 0.8, rng, sm), datasets.random_residues(700, rng, sm)])
    seqs[4] = q[100:].copy()
    return SequenceDB.from_sequences(seqs)


@pytest.mark.parametrize("mode", ["SW", "NW", "HW", "OV"])
@pytest.mark.parametrize("qlen", [700, 1600, 3300])
def test_chained_passes_match_the_oracle(product, oracle, mode, qlen, monkeypatch):
    rng = np.random.default_rng(qlen)
    sm = matrices.blosum62()
    q = datasets.random_residues(qlen, rng, sm)
    db = _db(rng, sm, q)
    monkeypatch.setenv("OPAL_B200_CHAIN", "1")
    h = product.create_db(db, 0)
    try:
        for st in (0, 1):
            rc, sc, eq, et, _ = h.search(q, 11, 1, sm.flat(), 23, st, mode)
            assert rc == 0, product.last_error()
            assert h.last_stats()["chained"] > 0, h.last_stats()
            rc2, want = search_dump(oracle, q, db, 11, 1, sm.flat(), 23, st, MODES[mode])
            assert rc2 == 0
            assert [int(x) for x in sc] == [w[1] for w in want], (mode, st)
            if st:
                assert [int(x) for x in eq] == [w[2] for w in want] and [int(x) for x in et] == [w[3] for w in want], (mode, st)
    finally:
        h.close()


def test_chained_passes_at_32_bits_and_with_exact_end_resweep(product, oracle, monkeypatch):
    """16 x BLOSUM62: the copies of the query leave 16 bits (re-run by the 32-bit class, chained as well), and scores
    beyond the fast end-location key are swept twice inside a chained pass (the second sweep must not park the
    boundary row again).  Gap penalties of thousands send NW / HW / OV to the 32-bit class a priori."""
    rng = np.random.default_rng(8)
    sm = matrices.blosum62()
    q = datasets.random_residues(2100, rng, sm)
    db = _db(rng, sm, q)
    m16 = (sm.matrix * 16).ravel().astype(np.int32)
    monkeypatch.setenv("OPAL_B200_CHAIN", "1")
    for mode, go, ge, m in (("SW", 176, 16, m16), ("NW", 3000, 900, sm.flat()), ("OV", 176, 16, m16)):
        for st in (0, 1):
            rc1, want = search_dump(oracle, q, db, go, ge, m, 23, st, MODES[mode])
            rc2, got = search_dump(product, q, db, go, ge, m, 23, st, MODES[mode])
            assert rc1 == rc2 == 0, (mode, st, product.last_error())
            assert got == want, (mode, st, [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w][:3])
    assert max(w[1] for w in search_dump(oracle, q, db, 176, 16, m16, 23, 0, MODES["SW"])[1]) > 32767


def test_chained_and_unchained_plans_agree_in_a_batch(product, monkeypatch):
    rng = np.random.default_rng(9)
    sm = matrices.blosum62()
    qs = [datasets.random_residues(n, rng, sm) for n in (2500, 900, 1300, 4000)]
    db = _db(rng, sm, qs[0])
    monkeypatch.setenv("OPAL_B200_NO_CHAIN", "1")
    h = product.create_db(db, 0)
    rc, S0, Q0, T0, _ = h.search_batch(qs, 11, 1, sm.flat(), 23, 1, ["HW", "SW", "NW", "OV"], in_flight=1)
    assert rc == 0 and h.last_stats()["chained"] == 0
    monkeypatch.delenv("OPAL_B200_NO_CHAIN")
    monkeypatch.setenv("OPAL_B200_CHAIN", "1")
    rc, S1, Q1, T1, _ = h.search_batch(qs, 11, 1, sm.flat(), 23, 1, ["HW", "SW", "NW", "OV"], in_flight=4)
    h.close()
    assert rc == 0
    assert np.array_equal(S0, S1) and np.array_equal(Q0, Q1) and np.array_equal(T0, T1)
