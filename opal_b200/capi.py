"""ctypes binding of the opal.h C ABI (include/opal.h).

The same binding drives three shared libraries that all export the reference's
symbols (reference src/opal.h:82-165):

  * the product, ``opal_b200/csrc/libopal_b200.so`` (sm_100a CUDA path),
  * the parity checker ``oracle/liboracle.so`` (scalar C restatement), and
  * ``oracle/_ref/libopal_ref.so`` (the unmodified reference).

Only the first is product code; the other two are loaded by tests/ and bench.py
through :class:`OpalCLibrary` with an explicit path.  Data is marshalled with
numpy so that half-million-sequence databases cross the boundary without Python
loops: a database is one concatenated ``uint8`` residue buffer plus offsets, and
``OpalSearchResult`` records live in one structured array.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

# Constants of include/opal.h (reference src/opal.h:17-40).
OPAL_ERR_OVERFLOW = 1
OPAL_ERR_NO_SIMD_SUPPORT = 2
OPAL_ERR_INVALID_MODE = 3
OPAL_ERR_INVALID_ARGUMENT = 4  # addition of opal-b200 (include/opal.h)
OPAL_MODE_NW, OPAL_MODE_HW, OPAL_MODE_OV, OPAL_MODE_SW = 0, 1, 2, 3
OPAL_OVERFLOW_SIMPLE, OPAL_OVERFLOW_BUCKETS = 0, 1
OPAL_SEARCH_SCORE, OPAL_SEARCH_SCORE_END, OPAL_SEARCH_ALIGNMENT = 0, 1, 2
OPAL_ALIGN_MATCH, OPAL_ALIGN_DEL, OPAL_ALIGN_INS, OPAL_ALIGN_MISMATCH = 0, 1, 2, 3
MODES = {"NW": OPAL_MODE_NW, "HW": OPAL_MODE_HW, "OV": OPAL_MODE_OV, "SW": OPAL_MODE_SW}

# struct OpalSearchResult, LP64 layout (reference src/opal.h:47-74): 40 bytes.
RESULT_DTYPE = np.dtype(
    {
        "names": ["scoreSet", "score", "endLocationTarget", "endLocationQuery",
                  "startLocationTarget", "startLocationQuery", "alignment", "alignmentLength"],
        "formats": ["<i4", "<i4", "<i4", "<i4", "<i4", "<i4", "<u8", "<i4"],
        "offsets": [0, 4, 8, 12, 16, 20, 24, 32],
        "itemsize": 40,
    }
)


class OpalSearchResultStruct(ctypes.Structure):
    _fields_ = [
        ("scoreSet", ctypes.c_int), ("score", ctypes.c_int),
        ("endLocationTarget", ctypes.c_int), ("endLocationQuery", ctypes.c_int),
        ("startLocationTarget", ctypes.c_int), ("startLocationQuery", ctypes.c_int),
        ("alignment", ctypes.c_void_p), ("alignmentLength", ctypes.c_int),
    ]


assert ctypes.sizeof(OpalSearchResultStruct) == RESULT_DTYPE.itemsize == 40

_libc = ctypes.CDLL(None)
_libc.free.argtypes = [ctypes.c_void_p]
_libc.free.restype = None


@dataclass
class SequenceDB:
    """A database in the form the C API takes it: residues are alphabet indices."""

    residues: np.ndarray  # uint8, all sequences back to back
    offsets: np.ndarray   # int64, n+1 entries

    def __post_init__(self):
        self.residues = np.ascontiguousarray(self.residues, dtype=np.uint8)
        if self.residues.size == 0:  # keep a valid base address for all-empty databases
            self.residues = np.zeros(1, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(self.offsets, dtype=np.int64)
        self.lengths = np.ascontiguousarray(np.diff(self.offsets), dtype=np.int32)
        self.pointers = (self.residues.ctypes.data + self.offsets[:-1]).astype(np.uint64)

    @classmethod
    def from_sequences(cls, seqs):
        seqs = [np.asarray(s, dtype=np.uint8) for s in seqs]
        offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        if seqs:
            offsets[1:] = np.cumsum([len(s) for s in seqs])
        residues = np.concatenate(seqs) if seqs and offsets[-1] > 0 else np.zeros(0, dtype=np.uint8)
        return cls(residues, offsets)

    def __len__(self):
        return len(self.lengths)

    def sequence(self, i):
        return self.residues[self.offsets[i]:self.offsets[i + 1]]

    @property
    def total_residues(self):
        return int(self.offsets[-1])

    def subset(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        return SequenceDB.from_sequences([self.sequence(int(i)) for i in idx])


def new_results(n):
    """n records in the state opalInitSearchResult leaves them (reference src/opal.cpp:1549-1555)."""
    res = np.zeros(n, dtype=RESULT_DTYPE)
    for f in ("endLocationTarget", "endLocationQuery", "startLocationTarget", "startLocationQuery"):
        res[f] = -1
    return res


def result_pointers(results):
    assert results.dtype == RESULT_DTYPE and results.flags["C_CONTIGUOUS"]
    return (results.ctypes.data + 40 * np.arange(len(results), dtype=np.uint64)).astype(np.uint64)


def get_alignment(results, i):
    """Copy of the operation string of record i (empty array when there is none)."""
    n = int(results["alignmentLength"][i])
    p = int(results["alignment"][i])
    if p == 0 or n <= 0:
        return np.zeros(0, dtype=np.uint8)
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_ubyte)), shape=(n,)).copy()


def free_alignments(results):
    """free() every alignment the library malloc()ed (reference src/opal.h:69-70)."""
    for i in np.nonzero(results["alignment"])[0]:
        _libc.free(ctypes.c_void_p(int(results["alignment"][i])))
        results["alignment"][i] = 0


class OpalCLibrary:
    """One loaded shared library exporting the opal.h symbols."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(f"shared library not found: {path}")
        self.path = path
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        search_args = [vp, ci, vp, ci, vp, ci, ci, vp, ci, vp, ci, ci, ci]
        self.lib.opalSearchDatabase.argtypes = search_args
        self.lib.opalSearchDatabase.restype = ci
        if hasattr(self.lib, "opalSearchDatabaseRescore"):
            self.lib.opalSearchDatabaseRescore.argtypes = search_args
            self.lib.opalSearchDatabaseRescore.restype = ci
        self.lib.opalSearchDatabaseCharSW.argtypes = [vp, ci, vp, ci, vp, ci, ci, vp, ci, vp]
        self.lib.opalSearchDatabaseCharSW.restype = ci
        self.lib.opalInitSearchResult.argtypes = [ctypes.POINTER(OpalSearchResultStruct)]
        self.lib.opalInitSearchResult.restype = None
        self.lib.opalSearchResultIsEmpty.argtypes = [OpalSearchResultStruct]
        self.lib.opalSearchResultIsEmpty.restype = ci
        self.lib.opalSearchResultSetScore.argtypes = [ctypes.POINTER(OpalSearchResultStruct), ci]
        self.lib.opalSearchResultSetScore.restype = None

    def has(self, symbol):
        return hasattr(self.lib, symbol)

    def search_database(self, query, db: SequenceDB, gap_open, gap_ext, score_matrix, alphabet_length,
                        results=None, search_type=OPAL_SEARCH_SCORE, mode=OPAL_MODE_SW,
                        overflow_method=OPAL_OVERFLOW_BUCKETS, entry="opalSearchDatabase", result_ptrs=None):
        """opalSearchDatabase (reference src/opal.h:150-154). Returns (rc, results).  `result_ptrs`: the array of
        record pointers a caller that reuses its records keeps (result_pointers(results)), instead of rebuilding it."""
        query = np.ascontiguousarray(query, dtype=np.uint8)
        sm = np.ascontiguousarray(score_matrix, dtype=np.int32).ravel()
        assert sm.size == alphabet_length * alphabet_length
        if results is None:
            results = new_results(len(db))
        rp = result_pointers(results) if result_ptrs is None else result_ptrs
        qbuf = query if query.size else np.zeros(1, dtype=np.uint8)
        rc = getattr(self.lib, entry)(
            qbuf.ctypes.data, int(query.size), db.pointers.ctypes.data, len(db), db.lengths.ctypes.data,
            int(gap_open), int(gap_ext), sm.ctypes.data, int(alphabet_length), rp.ctypes.data,
            int(search_type), int(mode), int(overflow_method))
        return rc, results

    def search_database_char_sw(self, query, db: SequenceDB, gap_open, gap_ext, score_matrix,
                                alphabet_length, results=None):
        """opalSearchDatabaseCharSW (reference src/opal.h:162-165). Returns (rc, results)."""
        query = np.ascontiguousarray(query, dtype=np.uint8)
        sm = np.ascontiguousarray(score_matrix, dtype=np.int32).ravel()
        if results is None:
            results = new_results(len(db))
        rp = result_pointers(results)
        rc = self.lib.opalSearchDatabaseCharSW(
            query.ctypes.data, int(query.size), db.pointers.ctypes.data, len(db), db.lengths.ctypes.data,
            int(gap_open), int(gap_ext), sm.ctypes.data, int(alphabet_length), rp.ctypes.data)
        return rc, results
