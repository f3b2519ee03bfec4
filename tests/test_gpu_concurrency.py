"""Concurrent drop-in calls (SURVEY.md section 8b "Threading": the reference is re-entrant bar one static,
src/opal.cpp:604; a replacement must be thread-safe, each call on streams of its own).  Eight host threads call
opalSearchDatabase at once on different databases, queries, modes and search levels; every record must equal what
the same call returns when it runs alone."""
import threading

import numpy as np
import pytest

from _util import MODES, SequenceDB, dump_results, free_alignments
from opal_b200 import datasets, matrices

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(91)
    sm = matrices.blosum62()
    dna = matrices.simple(4, 2, -3)
    cases = []
    for k in range(8):
        if k % 4 == 3:  # DNA alphabet, other gap penalties
            q = rng.integers(0, 4, int(rng.integers(40, 700))).astype(np.uint8)
            seqs = [rng.integers(0, 4, int(n)).astype(np.uint8) for n in rng.integers(1, 900, 150 + 40 * k)]
            seqs[1] = q[5:].copy()
            cases.append((q, SequenceDB.from_sequences(seqs), 5, 2, dna.flat(), 4, ("SW", "NW", "HW", "OV")[k % 4 - 3], 1))
            continue
        q = datasets.random_residues(int(rng.integers(30, 1300)), rng, sm)
        seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(1, 600, 200 + 60 * k)]
        seqs[0] = datasets.random_residues(2500, rng, sm)
        seqs[7] = datasets.mutate(q, 0.8, rng, sm)
        cases.append((q, SequenceDB.from_sequences(seqs), 11, 1, sm.flat(), 23, ("SW", "NW", "HW", "OV")[k % 4], (1, 0, 2)[k % 3] if k % 4 == 0 else 1))
    return cases


def test_eight_threads_calling_the_drop_in_entry_point(product):
    cases = _cases()

    def run(c):
        q, db, go, ge, m, a, mode, st = c
        rc, res = product.search_database(q, db, go, ge, m, a, None, st, MODES[mode])
        out = (rc, dump_results(res))
        free_alignments(res)
        return out

    alone = [run(c) for c in cases]
    assert all(rc == 0 for rc, _ in alone)
    for _ in range(3):
        together = [None] * len(cases)

        def work(k):
            for _ in range(4):
                together[k] = run(cases[k])

        threads = [threading.Thread(target=work, args=(k,)) for k in range(len(cases))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for k in range(len(cases)):
            assert together[k] == alone[k], (k, cases[k][6])


def test_threads_sharing_one_resident_handle_through_batches(product):
    """search_batch drives several host threads over one handle; two batches from two caller threads on two handles
    of the same database must not disturb each other either."""
    rng = np.random.default_rng(92)
    sm = matrices.blosum62()
    seqs = [datasets.random_residues(int(n), rng, sm) for n in rng.integers(1, 500, 700)]
    db = SequenceDB.from_sequences(seqs)
    qs = [datasets.random_residues(int(n), rng, sm) for n in (64, 300, 513, 900, 31, 1100)]
    h1, h2 = product.create_db(db, 0), product.create_db(db, 0)
    try:
        rc, want, wq, wt, _ = h1.search_batch(qs, 11, 1, sm.flat(), 23, 1, "SW", in_flight=1)
        assert rc == 0
        got = {}

        def work(name, h, mode):
            got[name] = h.search_batch(qs, 11, 1, sm.flat(), 23, 1, mode, in_flight=4)

        t1 = threading.Thread(target=work, args=("a", h1, "SW"))
        t2 = threading.Thread(target=work, args=("b", h2, "SW"))
        t1.start(); t2.start(); t1.join(); t2.join()
        for name in ("a", "b"):
            rc, s, q, t, _ = got[name]
            assert rc == 0 and np.array_equal(s, want) and np.array_equal(q, wq) and np.array_equal(t, wt)
    finally:
        h1.close()
        h2.close()
