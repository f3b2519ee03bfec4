"""Seeded random parity sweep of the CUDA path against the oracle through the C ABI: alphabets from 2 to 256 letters,
asymmetric matrices with small to huge entries (16-bit lanes, overflow re-runs and a-priori 32-bit routing), gap
penalties including gapExt = 0 and gapOpen < gapExt, query lengths from 1 to several strips of rows, ragged targets
including empty ones.  Every record field is compared, for all four modes and both score levels."""
import numpy as np
import pytest

from _util import MODES, SequenceDB, search_dump

pytestmark = pytest.mark.gpu

ALPHABETS = (2, 4, 7, 20, 23, 24, 60, 254, 255, 256)
MAGNITUDES = (1, 5, 20, 127, 2000, 40000)


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    A = int(rng.choice(ALPHABETS))
    mag = int(rng.choice(MAGNITUDES))
    matrix = rng.integers(-mag, mag + 1, (A, A)).astype(np.int32)
    diag = rng.integers(max(1, mag // 2), mag + 1, A)
    matrix[np.arange(A), np.arange(A)] = diag  # matches are rewarded, everything else is arbitrary (asymmetric)
    go = int(rng.integers(0, 3 * mag + 2))
    ge = int(rng.integers(0, mag + 2))
    Q = int(rng.choice([1, 2, 17, 33, 100, 257, 513, 545, 1056, 1057, 1500]))
    n = int(rng.integers(1, 220))
    lens = rng.integers(0, 400, n)
    lens[rng.integers(0, n)] = int(rng.integers(1500, 5000))
    q = rng.integers(0, A, Q).astype(np.uint8)
    seqs = [rng.integers(0, A, int(x)).astype(np.uint8) for x in lens]
    for k in range(0, n, 7):  # related sequences, so that scores are not all tiny
        if Q > 4:
            a, b = sorted(rng.integers(0, Q, 2))
            piece = q[a:b + 1].copy()
            flip = rng.random(len(piece)) < 0.15
            piece[flip] = rng.integers(0, A, int(flip.sum()))
            seqs[k] = np.concatenate([seqs[k][:20], piece, seqs[k][20:]])
    return q, SequenceDB.from_sequences(seqs), go, ge, matrix.ravel(), A


@pytest.mark.parametrize("seed", range(48))
def test_random_parameters_match_the_oracle(product, oracle, seed):
    q, db, go, ge, matrix, A = _case(seed)
    mode = ("NW", "HW", "OV", "SW")[seed % 4]
    for st in (0, 1):
        rc1, want = search_dump(oracle, q, db, go, ge, matrix, A, st, MODES[mode])
        rc2, got = search_dump(product, q, db, go, ge, matrix, A, st, MODES[mode])
        assert rc1 == rc2, (rc1, rc2, product.last_error())
        if rc1 == 0:
            assert got == want, (mode, st, go, ge, A, len(q), [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w][:3])
