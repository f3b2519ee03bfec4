"""One call, several devices (reference src/opal.h:150-154: one call covers the whole database).

The library deals a database over a list of devices and searches the shards concurrently, one host thread per device.
On a box with several GPUs the list is the distinct devices; on a one-GPU box the same ordinal is listed several times
(several shards on one device), which exercises the same dealing, fan-out, scatter and per-shard alignment code.
Everything is compared with the single-device result and with the oracle."""
import os

import numpy as np
import pytest

from _util import MODES, SequenceDB, dump_results, free_alignments, new_results, search_dump
from opal_b200 import datasets, matrices

pytestmark = pytest.mark.gpu


def _devices(product, want=3):
    n = product.device_count()
    return list(range(n)) if n >= 2 else [0] * want


def _db(rng, sm, n, q):
    seqs = [datasets.random_residues(int(x), rng, sm) for x in rng.integers(0, 500, n)]
    seqs[3] = datasets.random_residues(4000, rng, sm)
    for k in range(0, n, 9):
        seqs[k] = datasets.mutate(q, 0.7, rng, sm)
    return SequenceDB.from_sequences(seqs)


@pytest.mark.parametrize("mode", ["NW", "HW", "OV", "SW"])
def test_handle_on_several_devices_equals_one_device_and_oracle(product, oracle, mode):
    rng = np.random.default_rng(31)
    sm = matrices.blosum62()
    q = datasets.random_residues(150, rng, sm)
    db = _db(rng, sm, 401, q)
    devs = _devices(product)
    h1 = product.create_db(db, 0)
    hn = product.create_db(db, devs)
    try:
        assert hn.devices() == len(devs) and h1.devices() == 1
        skip = (rng.random(len(db)) < 0.2).astype(np.uint8)
        for st in (0, 1):
            for sk in (None, skip):
                rc1, s1, q1, t1, _ = h1.search(q, 11, 1, sm.flat(), 23, st, mode, skip=sk)
                rcn, sn, qn, tn, ms = hn.search(q, 11, 1, sm.flat(), 23, st, mode, skip=sk)
                assert rc1 == 0 and rcn == 0 and ms > 0
                keep = np.ones(len(db), bool) if sk is None else sk == 0
                assert np.array_equal(s1[keep], sn[keep]) and np.array_equal(q1[keep], qn[keep]) and np.array_equal(t1[keep], tn[keep])
        rc, want = search_dump(oracle, q, db, 11, 1, sm.flat(), 23, 1, MODES[mode])
        rcn, sn, qn, tn, _ = hn.search(q, 11, 1, sm.flat(), 23, 1, mode)
        assert rc == 0 and rcn == 0
        assert [int(x) for x in sn] == [w[1] for w in want]
        assert [int(x) for x in qn] == [w[2] for w in want] and [int(x) for x in tn] == [w[3] for w in want]
        # several queries in flight on every device
        qs = [q, datasets.random_residues(33, rng, sm), datasets.mutate(q, 0.5, rng, sm), datasets.random_residues(700, rng, sm)]
        rc1, S1, Q1, T1, _ = h1.search_batch(qs, 11, 1, sm.flat(), 23, 1, mode, in_flight=2)
        rcn, Sn, Qn, Tn, _ = hn.search_batch(qs, 11, 1, sm.flat(), 23, 1, mode, in_flight=2)
        assert rc1 == 0 and rcn == 0
        assert np.array_equal(S1, Sn) and np.array_equal(Q1, Qn) and np.array_equal(T1, Tn)
    finally:
        h1.close()
        hn.close()


@pytest.mark.parametrize("mode", ["SW", "NW"])
def test_drop_in_call_on_several_devices(product, oracle, mode, monkeypatch):
    """opalSearchDatabase with OPAL_B200_DEVICES set: one call, every listed device, identical records -- all three
    search levels, prefilled records included (the reuse rule is applied per shard)."""
    rng = np.random.default_rng(32)
    sm = matrices.blosum62()
    q = datasets.random_residues(120, rng, sm)
    db = _db(rng, sm, 230, q)
    devs = _devices(product)
    for st in (0, 1, 2):
        monkeypatch.delenv("OPAL_B200_DEVICES", raising=False)
        rc1, one = product.search_database(q, db, 11, 1, sm.flat(), 23, None, st, MODES[mode])
        monkeypatch.setenv("OPAL_B200_DEVICES", ",".join(str(d) for d in devs))
        pre = new_results(len(db))
        pre["scoreSet"][5] = 1
        pre["score"][5] = 4242 if st == 0 else one["score"][5]
        if st:
            pre["endLocationQuery"][5], pre["endLocationTarget"][5] = one["endLocationQuery"][5], one["endLocationTarget"][5]
        rcn, many = product.search_database(q, db, 11, 1, sm.flat(), 23, pre, st, MODES[mode])
        assert rc1 == 0 and rcn == 0
        a, b = dump_results(one), dump_results(many)
        if st == 0:
            assert b[5][1] == 4242
            a[5] = b[5]
        assert a == b
        free_alignments(one)
        free_alignments(many)
    monkeypatch.setenv("OPAL_B200_DEVICES", "all")
    rc, res = product.search_database(q, db, 11, 1, sm.flat(), 23, None, 1, MODES[mode])
    rc2, want = search_dump(oracle, q, db, 11, 1, sm.flat(), 23, 1, MODES[mode])
    assert rc == 0 and rc2 == 0 and dump_results(res) == want


def test_topk_and_results_on_several_devices(product):
    rng = np.random.default_rng(33)
    sm = matrices.blosum62()
    q = datasets.random_residues(200, rng, sm)
    db = _db(rng, sm, 500, q)
    devs = _devices(product, want=4)
    h1 = product.create_db(db, 0)
    hn = product.create_db(db, devs)
    try:
        for st in (1, 2):
            rc1, i1, r1 = h1.search_topk(q, 11, 1, sm.flat(), 23, st, "SW", 40)
            rcn, i_n, rn = hn.search_topk(q, 11, 1, sm.flat(), 23, st, "SW", 40)
            assert rc1 == 0 and rcn == 0
            assert np.array_equal(i1, i_n) and dump_results(r1) == dump_results(rn)
            free_alignments(r1)
            free_alignments(rn)
        rc1, a = h1.search_results(q, 11, 1, sm.flat(), 23, 2, "HW")
        rcn, b = hn.search_results(q, 11, 1, sm.flat(), 23, 2, "HW")
        assert rc1 == 0 and rcn == 0 and dump_results(a) == dump_results(b)
        free_alignments(a)
        free_alignments(b)
    finally:
        h1.close()
        hn.close()


def test_packed_database_on_several_devices(product):
    rng = np.random.default_rng(34)
    sm = matrices.blosum62()
    q = datasets.random_residues(90, rng, sm)
    db = _db(rng, sm, 300, q)
    order = np.argsort(-db.lengths.astype(np.int64), kind="stable").astype(np.int32)
    residues = np.concatenate([db.sequence(int(i)) for i in order])
    os.environ["OPAL_B200_DEVICES"] = ",".join(str(d) for d in _devices(product))
    try:
        hn = product.create_db_sorted(residues, db.lengths[order], order, -1)
    finally:
        del os.environ["OPAL_B200_DEVICES"]
    h1 = product.create_db(db, 0)
    try:
        assert hn.devices() >= 2
        for mode in ("SW", "OV"):
            rc1, s1, q1, t1, _ = h1.search(q, 11, 1, sm.flat(), 23, 1, mode)
            rcn, sn, qn, tn, _ = hn.search(q, 11, 1, sm.flat(), 23, 1, mode)
            assert rc1 == 0 and rcn == 0
            assert np.array_equal(s1, sn) and np.array_equal(q1, qn) and np.array_equal(t1, tn)
    finally:
        h1.close()
        hn.close()


@pytest.mark.parametrize("mode", ["SW", "HW"])
def test_drop_in_call_in_pipelined_slices(product, oracle, mode, monkeypatch):
    """A large database goes through the drop-in call in slices (longest sequences first) whose uploads and searches
    overlap; OPAL_B200_SLICES forces that on a small one.  Records must not depend on the number of slices, alone or
    combined with several devices."""
    rng = np.random.default_rng(35)
    sm = matrices.blosum62()
    q = datasets.random_residues(140, rng, sm)
    db = _db(rng, sm, 350, q)
    monkeypatch.delenv("OPAL_B200_DEVICES", raising=False)
    monkeypatch.delenv("OPAL_B200_SLICES", raising=False)
    for st in (0, 1, 2):
        rc1, one = product.search_database(q, db, 11, 1, sm.flat(), 23, None, st, MODES[mode])
        assert rc1 == 0
        want = dump_results(one)
        free_alignments(one)
        for slices, devs in ((3, None), (2, "0,0"), (16, None)):
            monkeypatch.setenv("OPAL_B200_SLICES", str(slices))
            if devs:
                monkeypatch.setenv("OPAL_B200_DEVICES", devs)
            rc, res = product.search_database(q, db, 11, 1, sm.flat(), 23, None, st, MODES[mode])
            assert rc == 0, product.last_error()
            assert dump_results(res) == want, (st, slices, devs)
            free_alignments(res)
            monkeypatch.delenv("OPAL_B200_SLICES")
            monkeypatch.delenv("OPAL_B200_DEVICES", raising=False)
    rc, want = search_dump(oracle, q, db, 11, 1, sm.flat(), 23, 1, MODES[mode])
    monkeypatch.setenv("OPAL_B200_SLICES", "4")
    rc2, got = search_dump(product, q, db, 11, 1, sm.flat(), 23, 1, MODES[mode])
    assert rc == 0 and rc2 == 0 and got == want
    # invalid arguments still come back as such from every slice
    rc, _ = product.search_database(np.array([0, 99], dtype=np.uint8), db, 11, 1, sm.flat(), 23, None, 0, MODES[mode])
    assert rc == 4
