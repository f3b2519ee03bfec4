"""Python mirror of include/opal_b200_io.h (libopal_b200_io.so): score matrices, FASTA and the packed database
format, i.e. the data formats either side of the search path (reference src/ScoreMatrix.cpp:17-35,
src/opal_aligner.cpp:247-301).  Host-only code; used by the tests and by callers that want the C++ readers."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .capi import SequenceDB
from .matrices import ScoreMatrix

IO_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli", "libopal_b200_io.so")


class OpalIO:
    def __init__(self, path=IO_LIB_PATH):
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C opal_b200/cli` (or __graft_entry__.build())")
        L = self.lib = ctypes.CDLL(path)
        vp, ci, ll, cp = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_char_p
        L.opalio_last_error.restype = cp
        L.opalio_load_matrix.argtypes = [cp, cp, vp, ctypes.POINTER(ci), vp, ci]
        L.opalio_load_matrix.restype = ci
        L.opalio_read_fasta.argtypes = [cp, vp, ci, ll, ctypes.POINTER(ci)]
        L.opalio_read_fasta.restype = vp
        for name, res in (("count", ci), ("residues", ll), ("offsets", vp), ("data", vp)):
            f = getattr(L, "opalio_sequences_" + name)
            f.argtypes, f.restype = [vp], res
        L.opalio_sequences_free.argtypes, L.opalio_sequences_free.restype = [vp], None
        L.opalio_pack_fasta.argtypes = [cp, vp, ci, cp]
        L.opalio_pack_fasta.restype = ci
        L.opalio_packed_open.argtypes, L.opalio_packed_open.restype = [cp], vp
        for name, res in (("count", ci), ("residues", ll), ("lengths", vp), ("order", vp), ("data", vp)):
            f = getattr(L, "opalio_packed_" + name)
            f.argtypes, f.restype = [vp], res
        L.opalio_packed_alphabet.argtypes, L.opalio_packed_alphabet.restype = [vp, vp], ci
        L.opalio_packed_free.argtypes, L.opalio_packed_free.restype = [vp], None

    def last_error(self):
        return self.lib.opalio_last_error().decode()

    def load_matrix(self, name=None, path=None) -> ScoreMatrix:
        alphabet = np.zeros(256, dtype=np.uint8)
        matrix = np.zeros(254 * 254, dtype=np.int32)
        n = ctypes.c_int(0)
        rc = self.lib.opalio_load_matrix(None if name is None else name.encode(), None if path is None else str(path).encode(),
                                         alphabet.ctypes.data, ctypes.byref(n), matrix.ctypes.data, matrix.size)
        if rc:
            raise ValueError(self.last_error())
        a = n.value
        return ScoreMatrix(alphabet[:a].tobytes().decode("latin1"), matrix[:a * a].copy())

    def read_fasta(self, path, sm: ScoreMatrix, max_residues=0):
        """Returns (SequenceDB, whole_file)."""
        alphabet = np.frombuffer(sm.alphabet.encode("latin1"), dtype=np.uint8).copy()
        whole = ctypes.c_int(-1)
        h = self.lib.opalio_read_fasta(str(path).encode(), alphabet.ctypes.data, len(alphabet), int(max_residues), ctypes.byref(whole))
        if not h:
            raise ValueError(self.last_error())
        try:
            n = self.lib.opalio_sequences_count(h)
            total = self.lib.opalio_sequences_residues(h)
            offsets = np.ctypeslib.as_array(ctypes.cast(self.lib.opalio_sequences_offsets(h), ctypes.POINTER(ctypes.c_longlong)), shape=(n + 1,)).copy()
            data = (np.ctypeslib.as_array(ctypes.cast(self.lib.opalio_sequences_data(h), ctypes.POINTER(ctypes.c_ubyte)), shape=(total,)).copy()
                    if total else np.zeros(0, dtype=np.uint8))
        finally:
            self.lib.opalio_sequences_free(h)
        return SequenceDB(data, offsets), bool(whole.value)

    def pack_fasta(self, fasta_path, sm: ScoreMatrix, out_path):
        alphabet = np.frombuffer(sm.alphabet.encode("latin1"), dtype=np.uint8).copy()
        if self.lib.opalio_pack_fasta(str(fasta_path).encode(), alphabet.ctypes.data, len(alphabet), str(out_path).encode()):
            raise ValueError(self.last_error())

    def read_packed(self, path):
        """Returns dict(alphabet, lengths, order, residues): the arguments of OpalB200.create_db_sorted."""
        h = self.lib.opalio_packed_open(str(path).encode())
        if not h:
            raise ValueError(self.last_error())
        try:
            n = self.lib.opalio_packed_count(h)
            total = self.lib.opalio_packed_residues(h)
            alphabet = np.zeros(256, dtype=np.uint8)
            a = self.lib.opalio_packed_alphabet(h, alphabet.ctypes.data)
            as_int = lambda p: np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_int)), shape=(n,)).copy() if n else np.zeros(0, np.int32)  # noqa: E731
            lengths, order = as_int(self.lib.opalio_packed_lengths(h)), as_int(self.lib.opalio_packed_order(h))
            residues = (np.ctypeslib.as_array(ctypes.cast(self.lib.opalio_packed_data(h), ctypes.POINTER(ctypes.c_ubyte)), shape=(total,)).copy()
                        if total else np.zeros(0, dtype=np.uint8))
        finally:
            self.lib.opalio_packed_free(h)
        return {"alphabet": alphabet[:a].tobytes().decode("latin1"), "lengths": lengths, "order": order, "residues": residues}
