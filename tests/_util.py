"""Shared helpers for the test-suite (loading the three opal.h libraries, result dumps)."""
from __future__ import annotations

import ctypes
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from opal_b200 import (MODES, OPAL_OVERFLOW_BUCKETS, OPAL_OVERFLOW_SIMPLE, OPAL_SEARCH_ALIGNMENT,  # noqa: E402
                       OPAL_SEARCH_SCORE, OPAL_SEARCH_SCORE_END, OpalCLibrary, SequenceDB,
                       free_alignments, get_alignment, new_results)

ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libopal_ref.so")
PRODUCT_SO = os.path.join(ROOT, "opal_b200", "csrc", "libopal_b200.so")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

README_MATRIX = np.array([2, -1, -3, 0, -1, 4, -5, -1, -3, -5, 1, -10, 0, -1, -10, 4], dtype=np.int32)
README_QUERY = np.array([0, 1, 3, 2, 1, 0, 3, 0, 1, 1], dtype=np.uint8)
README_DB = [[1, 3, 2, 3, 0, 0, 1, 0, 2, 2, 1, 2, 3, 2], [2, 1, 1, 3, 2, 0, 0, 2, 2, 0, 2, 1],
             [0, 0, 2, 1, 0, 3, 1, 1, 2, 3, 2, 1, 0], [2, 3, 3, 3, 1, 1, 2, 2, 0]]


def dump_results(res, with_alignment=True, digest=False):
    """List of per-sequence tuples covering every OpalSearchResult field."""
    out = []
    for i in range(len(res)):
        rec = [int(res[f][i]) for f in ("scoreSet", "score", "endLocationQuery", "endLocationTarget",
                                        "startLocationQuery", "startLocationTarget", "alignmentLength")]
        if with_alignment:
            ops = get_alignment(res, i)
            rec.append(hashlib.sha1(ops.tobytes()).hexdigest()[:16] if digest else "".join(map(str, ops)))
        out.append(rec)
    return out


def search_dump(lib, query, db, go, ge, matrix, alen, search_type, mode, ovf=OPAL_OVERFLOW_SIMPLE,
                results=None, digest=False, entry="opalSearchDatabase"):
    rc, res = lib.search_database(query, db, go, ge, matrix, alen, results, search_type, mode, ovf, entry=entry)
    out = dump_results(res, digest=digest)
    free_alignments(res)
    return rc, out


def glibc_testcpp_data():
    """The inputs of the reference's self-test (reference src/test.cpp:35-54, 185-188): glibc
    srand(42), alphabet 4, query of 1000, 200 targets of 800 + rand() % 4000."""
    libc = ctypes.CDLL(None)
    libc.srand(42)
    rand = libc.rand
    query = np.array([rand() % 4 for _ in range(1000)], dtype=np.uint8)
    seqs = []
    for _ in range(200):
        n = 800 + rand() % 4000
        seqs.append(np.array([rand() % 4 for _ in range(n)], dtype=np.uint8))
    return query, SequenceDB.from_sequences(seqs)


def run_forked(fn, timeout=600):
    """Run fn() in a forked child and return its (picklable) result, or None if the child died
    (the reference's alignment stage can SIGSEGV / assert on NW/HW/OV)."""
    import pickle
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        try:
            os.close(r)
            import faulthandler
            faulthandler.disable()  # a dying reference is an expected outcome here, not a report
            os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
            data = pickle.dumps(fn())
            with os.fdopen(w, "wb") as f:
                f.write(data)
            os._exit(0)
        except BaseException:
            os._exit(1)
    os.close(w)
    with os.fdopen(r, "rb") as f:
        data = f.read()
    _, status = os.waitpid(pid, 0)
    if status != 0 or not data:
        return None
    return pickle.loads(data)


def search_sample_parallel(lib, query, db, sample, go, ge, matrix, alen, search_type, mode, threads=None, results=None):
    """`lib`.opalSearchDatabase over the sub-database db[sample], split into residue-balanced parts that run on host
    threads of their own (the checkers are single-threaded per call and re-entrant; ctypes releases the GIL).
    Returns (rc, records in `sample` order)."""
    import threading
    sample = np.asarray(sample, dtype=np.int64)
    threads = max(1, min(threads or (os.cpu_count() or 1), len(sample)))
    by_len = np.argsort(-db.lengths[sample].astype(np.int64), kind="stable")
    parts = [by_len[k::threads] for k in range(threads)]  # dealt longest first: equal work, equal length mix
    out = new_results(len(sample)) if results is None else results
    rcs = [0] * threads

    def work(k):
        sub = db.subset(sample[parts[k]])
        pre = None if results is None else np.ascontiguousarray(results[parts[k]])
        rcs[k], res = lib.search_database(query, sub, go, ge, matrix, alen, pre, search_type, mode)
        out[parts[k]] = res

    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return max(rcs), out
