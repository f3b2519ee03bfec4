"""GPU parity of the folded sweep (SearchParams::folded in opal_b200/csrc/search_kernel.cuh): the longest targets of
a database are swept one per warp with both 16-bit lanes on the SAME target -- the second half of the query rows
rides 32 columns behind the first -- so that the targets which bound the search time finish sooner.  The databases
here have a long tail so that the planner takes that route (asserted through opalb200_db_last_folded); every score and
end location is compared with the oracle, and with the same search with folding switched off."""
import os

import numpy as np
import pytest

from _util import MODES, SequenceDB, search_dump
from opal_b200 import datasets, matrices

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_chained_passes(monkeypatch):
    """These tests are about the folded sweep: keep the planner from giving the longest targets of a query of several
    strips chained passes instead (tests/test_gpu_chained.py covers those)."""
    monkeypatch.setenv("OPAL_B200_NO_CHAIN", "1")


def _tailed_db(rng, sm, n_short, n_long, long_lo, long_hi, planted=None, planted_long=None):
    seqs = [datasets.random_residues(int(x), rng, sm) for x in rng.integers(30, 420, n_short)]
    longs = [datasets.random_residues(int(x), rng, sm) for x in rng.integers(long_lo, long_hi, n_long)]
    if planted is not None:
        for k in range(0, n_short, max(1, n_short // 6)):
            seqs[k] = datasets.mutate(planted, 0.75, rng, sm)
    if planted_long is not None:
        # homologs INSIDE long targets, at both ends and in the middle, so that the best cell lies in either half
        # of the query rows and on either side of the 32-column lag
        for k in range(0, n_long, max(1, n_long // 5)):
            t = longs[k]
            h = datasets.mutate(planted_long, 0.9, rng, sm)
            at = [0, max(0, len(t) - len(h)), len(t) // 3][(k // max(1, n_long // 5)) % 3]
            t[at:at + len(h)] = h[:max(0, len(t) - at)]
    order = rng.permutation(n_short + n_long)
    allseqs = seqs + longs
    return SequenceDB.from_sequences([allseqs[i] for i in order])


def _check(product, oracle, q, db, go, ge, m, alen, st, expect_folded=True, mode="SW"):
    h = product.create_db(db, 0)
    try:
        rc, s, eq, et, _ = h.search(q, go, ge, m, alen, st, mode)
        assert rc == 0, product.last_error()
        folded = h.last_stats()["folded"]
        if expect_folded:
            assert folded > 0, "planner did not fold the tail of this database"
        rc, want = search_dump(oracle, q, db, go, ge, m, alen, st, MODES[mode])
        assert rc == 0
        assert [int(x) for x in s] == [w[1] for w in want]
        if st:
            weq = [w[2] if (w[1] > 0 or mode != "SW") else -1 for w in want]
            wet = [w[3] if (w[1] > 0 or mode != "SW") else -1 for w in want]
            assert [int(x) for x in eq] == weq
            assert [int(x) for x in et] == wet
        return folded, (s.copy(), eq.copy(), et.copy())
    finally:
        h.close()


@pytest.mark.parametrize("qlen", [97, 300, 513, 1100])
@pytest.mark.parametrize("st", [0, 1])
def test_folded_tail_matches_oracle(product, oracle, qlen, st):
    rng = np.random.default_rng(100 + qlen)
    sm = matrices.blosum62()
    q = sm.encode(datasets.P18080)[:qlen] if qlen <= 513 else datasets.random_residues(qlen, rng, sm)
    db = _tailed_db(rng, sm, 3000, 70, 2000, 5000, planted=q, planted_long=q)
    # (a query much shorter than 64 x the smallest strip height is left to the ordinary latency class)
    _check(product, oracle, q, db, 11, 1, sm.flat(), 23, st, expect_folded=qlen >= 300)


def test_folded_matches_unfolded_and_overflow_reruns(product, oracle):
    """8 x BLOSUM62 (gaps 88 / 8): planted near-copies inside long targets score above the 16-bit range, so folded
    targets are flagged and re-run at 32 bits; scores between the key-tracking limit and the 16-bit limit take the
    exact re-sweep inside the folded warp."""
    rng = np.random.default_rng(7)
    sm = matrices.blosum62()
    m8 = (sm.flat() * 8).astype(np.int32)
    q = datasets.random_residues(1000, rng, sm)
    db = _tailed_db(rng, sm, 2500, 64, 3000, 6000, planted=q[:300], planted_long=q)
    folded, got = _check(product, oracle, q, db, 88, 8, m8, 23, 1)
    os.environ["OPAL_B200_NO_FOLD"] = "1"
    try:
        folded2, got2 = _check(product, oracle, q, db, 88, 8, m8, 23, 1, expect_folded=False)
    finally:
        del os.environ["OPAL_B200_NO_FOLD"]
    assert folded2 == 0
    for a, b in zip(got, got2):
        assert np.array_equal(a, b)
    assert int(got[0].max()) > 32767


def test_folded_dna_zero_extension_gaps(product, oracle):
    """Alphabet 4, gapExt 0 and a query shorter than one strip: ties everywhere, first-row / first-column rule."""
    rng = np.random.default_rng(8)
    sm = matrices.simple(4, 2, -3)
    q = rng.integers(0, 4, 70).astype(np.uint8)
    seqs = [rng.integers(0, 4, int(x)).astype(np.uint8) for x in rng.integers(20, 300, 4000)]
    seqs += [rng.integers(0, 4, int(x)).astype(np.uint8) for x in rng.integers(3000, 9000, 66)]
    db = SequenceDB.from_sequences(seqs)
    for st in (0, 1):
        _check(product, oracle, q, db, 3, 0, sm.flat(), 4, st, expect_folded=False)


def test_folded_with_skipped_targets_and_batches(product, oracle):
    """A skipped target among the longest ones leaves fewer foldable targets (the folded class takes a prefix of the
    length-sorted database whose members are all wanted); a batch with several queries in flight shares the folded
    stream between its search contexts.  Both must give what single full searches give."""
    rng = np.random.default_rng(11)
    sm = matrices.blosum62()
    q = sm.encode(datasets.P18080)
    db = _tailed_db(rng, sm, 3000, 70, 2000, 5000, planted=q, planted_long=q)
    h = product.create_db(db, 0)
    try:
        rc, s0, eq0, et0, _ = h.search(q, 11, 1, sm.flat(), 23, 1, "SW")
        assert rc == 0 and h.last_stats()["folded"] > 0
        order = np.argsort(-db.lengths.astype(np.int64), kind="stable")
        for skipped in ([int(order[0])], [int(order[5]), int(order[40])], [int(i) for i in order[:70]]):
            skip = np.zeros(len(db), dtype=np.uint8)
            skip[skipped] = 1
            rc, s, eq, et, _ = h.search(q, 11, 1, sm.flat(), 23, 1, "SW", skip=skip)
            assert rc == 0
            keep = skip == 0
            assert np.array_equal(s[keep], s0[keep]) and np.array_equal(eq[keep], eq0[keep]) and np.array_equal(et[keep], et0[keep])
        queries = [q, q[:300], datasets.mutate(q, 0.6, rng, sm)[:450], q[100:]]
        rc, S, EQ, ET, _ = h.search_batch(queries, 11, 1, sm.flat(), 23, 1, "SW", in_flight=3)
        assert rc == 0
        for k, x in enumerate(queries):
            rc, s, eq, et, _ = h.search(x, 11, 1, sm.flat(), 23, 1, "SW")
            assert rc == 0
            assert np.array_equal(S[k], s) and np.array_equal(EQ[k], eq) and np.array_equal(ET[k], et)
    finally:
        h.close()


@pytest.mark.parametrize("mode", ["NW", "HW", "OV"])
@pytest.mark.parametrize("qlen", [300, 513, 700])
def test_folded_global_modes_match_oracle(product, oracle, mode, qlen):
    """NW / HW / OV: the high half-words of a folded task idle for 32 columns and are then set to their column -1;
    the last query row (HW, OV), the last target column (OV: low half-words scanned inside the sweep, high ones after
    it) and NW's final cell (in whichever half holds query row Q - 1) are picked up from the right half."""
    rng = np.random.default_rng(300 + qlen)
    sm = matrices.blosum62()
    q = sm.encode(datasets.P18080)[:qlen] if qlen <= 513 else datasets.random_residues(qlen, rng, sm)
    db = _tailed_db(rng, sm, 3000, 70, 1500, 3500, planted=q, planted_long=q)
    for st in (0, 1):
        folded, got = _check(product, oracle, q, db, 11, 1, sm.flat(), 23, st, mode=mode)
    os.environ["OPAL_B200_NO_FOLD"] = "1"
    try:
        folded2, got2 = _check(product, oracle, q, db, 11, 1, sm.flat(), 23, 1, expect_folded=False, mode=mode)
    finally:
        del os.environ["OPAL_B200_NO_FOLD"]
    assert folded2 == 0
    for a, b in zip(got, got2):
        assert np.array_equal(a, b)


def test_folded_global_modes_ties_and_short_query(product, oracle):
    """Alphabet 4 (ties everywhere) with a query that fits the low half-words alone (NW's last row then sits there)."""
    rng = np.random.default_rng(12)
    sm = matrices.simple(4, 2, -3)
    seqs = [rng.integers(0, 4, int(x)).astype(np.uint8) for x in rng.integers(20, 300, 3000)]
    seqs += [rng.integers(0, 4, int(x)).astype(np.uint8) for x in rng.integers(1500, 3000, 66)]
    db = SequenceDB.from_sequences(seqs)
    for qlen in (250, 330):
        q = rng.integers(0, 4, qlen).astype(np.uint8)
        for mode in ("NW", "HW", "OV"):
            _check(product, oracle, q, db, 5, 2, sm.flat(), 4, 1, expect_folded=False, mode=mode)
