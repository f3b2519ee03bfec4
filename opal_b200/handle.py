"""Python mirror of include/opal_b200.h: the resident-database handle of libopal_b200.so.

No CPU implementation exists: constructing :class:`OpalB200` raises if the CUDA library has not
been built, and every search fails loudly (non-zero return code) when no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .capi import MODES, OpalCLibrary, SequenceDB, new_results, result_pointers

LIB_PATH = os.environ.get("OPAL_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libopal_b200.so")


class OpalB200(OpalCLibrary):
    """libopal_b200.so: the opal.h entry points (inherited) plus the opal_b200.h extensions."""

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C opal_b200/csrc` "
                               "(or __graft_entry__.build()); opal-b200 has no CPU fallback")
        super().__init__(path)
        L, vp, ci = self.lib, ctypes.c_void_p, ctypes.c_int
        L.opalb200_device_count.restype = ci
        L.opalb200_last_error.restype = ctypes.c_char_p
        L.opalb200_db_create.argtypes = [vp, ci, vp, ci]
        L.opalb200_db_create.restype = vp
        L.opalb200_db_create_multi.argtypes = [vp, ci, vp, vp, ci]
        L.opalb200_db_create_multi.restype = vp
        L.opalb200_db_devices.argtypes = [vp]
        L.opalb200_db_devices.restype = ci
        L.opalb200_db_create_sorted.argtypes = [vp, vp, vp, ci, ci]
        L.opalb200_db_create_sorted.restype = vp
        L.opalb200_db_search_batch.argtypes = [vp, ci, vp, vp, ci, ci, vp, ci, ci, ci, vp, vp, vp, ci, vp]
        L.opalb200_db_search_batch.restype = ci
        L.opalb200_db_search_batch_modes.argtypes = [vp, ci, vp, vp, vp, ci, ci, vp, ci, ci, vp, vp, vp, ci, vp]
        L.opalb200_db_search_batch_modes.restype = ci
        L.opalb200_db_search_results.argtypes = [vp, vp, ci, ci, ci, vp, ci, vp, ci, ci]
        L.opalb200_db_search_results.restype = ci
        L.opalb200_db_search_topk.argtypes = [vp, vp, ci, ci, ci, vp, ci, ci, ci, ci, vp, vp, ctypes.POINTER(ci)]
        L.opalb200_db_search_topk.restype = ci
        L.opalb200_db_destroy.argtypes = [vp]
        L.opalb200_db_destroy.restype = None
        L.opalb200_db_length.argtypes = [vp]
        L.opalb200_db_length.restype = ci
        L.opalb200_db_residues.argtypes = [vp]
        L.opalb200_db_residues.restype = ctypes.c_longlong
        L.opalb200_db_search.argtypes = [vp, vp, ci, ci, ci, vp, ci, ci, ci, vp, vp, vp, vp, vp]
        L.opalb200_db_search.restype = ci
        L.opalb200_db_last_stats.argtypes = [vp] + [ctypes.POINTER(ci)] * 7
        L.opalb200_db_last_stats.restype = None
        L.opalb200_db_last_folded.argtypes = [vp]
        L.opalb200_db_last_folded.restype = ci
        L.opalb200_db_last_chained.argtypes = [vp]
        L.opalb200_db_last_chained.restype = ci
        L.opalb200_measure_dpx_peak.argtypes = [ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)]
        L.opalb200_measure_dpx_peak.restype = ctypes.c_double
        L.opalb200_measure_dpx_peak_mix.argtypes = [ci, ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)]
        L.opalb200_measure_dpx_peak_mix.restype = ctypes.c_double
        L.opalb200_trim_cache.argtypes = []
        L.opalb200_trim_cache.restype = None

    def device_count(self):
        return int(self.lib.opalb200_device_count())

    def last_error(self):
        return self.lib.opalb200_last_error().decode()

    def create_db(self, db: SequenceDB, device=0):
        """opalb200_db_create on one device, or -- `device` a list of ordinals -- opalb200_db_create_multi."""
        if isinstance(device, (list, tuple)):
            devs = np.ascontiguousarray(device, dtype=np.int32)
            h = self.lib.opalb200_db_create_multi(db.pointers.ctypes.data, len(db), db.lengths.ctypes.data, devs.ctypes.data, int(devs.size))
        else:
            h = self.lib.opalb200_db_create(db.pointers.ctypes.data, len(db), db.lengths.ctypes.data, int(device))
        if not h:
            raise RuntimeError("opalb200_db_create failed: " + self.last_error())
        return ResidentDb(self, h, len(db))

    def create_db_sorted(self, residues, sorted_lengths, order=None, device=0):
        """opalb200_db_create_sorted: a database that is already packed (longest first, contiguous)."""
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        lens = np.ascontiguousarray(sorted_lengths, dtype=np.int32)
        buf = residues if residues.size else np.zeros(1, dtype=np.uint8)
        order_arr = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
        h = self.lib.opalb200_db_create_sorted(buf.ctypes.data, lens.ctypes.data,
                                               None if order_arr is None else order_arr.ctypes.data, int(lens.size), int(device))
        if not h:
            raise RuntimeError("opalb200_db_create_sorted failed: " + self.last_error())
        return ResidentDb(self, h, int(lens.size))

    def trim_cache(self):
        self.lib.opalb200_trim_cache()

    def measure_dpx_peak(self, device=0, mix=0):
        """(GCUPS the integer pipe can sustain, packed thread-instructions/s, kernel ms); mix 0 = SW, 1 = NW/HW/OV."""
        ips, ms = ctypes.c_double(0), ctypes.c_float(0)
        g = self.lib.opalb200_measure_dpx_peak_mix(int(device), int(mix), ctypes.byref(ips), ctypes.byref(ms))
        if g <= 0:
            raise RuntimeError("DPX probe failed: " + self.last_error())
        return float(g), float(ips.value), float(ms.value)


class ResidentDb:
    """A database packed once into one GPU's HBM (opalb200_db_create)."""

    def __init__(self, eng: OpalB200, handle, n):
        self.eng, self.handle, self.n = eng, handle, n

    def search(self, query, gap_open, gap_ext, score_matrix, alphabet_length, search_type, mode, skip=None):
        """Returns (rc, scores, endQuery, endTarget, device_ms); arrays are in caller order."""
        query = np.ascontiguousarray(query, dtype=np.uint8)
        sm = np.ascontiguousarray(score_matrix, dtype=np.int32).ravel()
        sc = np.zeros(self.n, dtype=np.int32)
        eq = np.full(self.n, -1, dtype=np.int32)
        et = np.full(self.n, -1, dtype=np.int32)
        ms = ctypes.c_float(0)
        if isinstance(mode, str):
            mode = MODES[mode]
        skp = None if skip is None else np.ascontiguousarray(skip, dtype=np.uint8)
        rc = self.eng.lib.opalb200_db_search(
            self.handle, query.ctypes.data, int(query.size), int(gap_open), int(gap_ext), sm.ctypes.data,
            int(alphabet_length), int(search_type), int(mode), None if skp is None else skp.ctypes.data,
            sc.ctypes.data, eq.ctypes.data, et.ctypes.data, ctypes.byref(ms))
        return rc, sc, eq, et, float(ms.value)

    def search_batch(self, queries, gap_open, gap_ext, score_matrix, alphabet_length, search_type, mode, in_flight=3, out=None):
        """opalb200_db_search_batch. Returns (rc, scores[nq, n], endQuery, endTarget, batch_ms); `out` = three
        preallocated int32 [nq, n] arrays to write into (entries that are not computed keep their old values)."""
        qs = [np.ascontiguousarray(q, dtype=np.uint8) for q in queries]
        qs = [q if q.size else np.zeros(1, dtype=np.uint8) for q in qs]
        nq = len(qs)
        ptrs = np.array([q.ctypes.data for q in qs], dtype=np.uint64)
        qlens = np.array([len(q) for q in queries], dtype=np.int32)
        sm = np.ascontiguousarray(score_matrix, dtype=np.int32).ravel()
        if out is not None:
            sc, eq, et = out
            assert all(a.shape == (nq, self.n) and a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] for a in out)
        else:
            sc = np.zeros((nq, self.n), dtype=np.int32)
            eq = np.full((nq, self.n), -1, dtype=np.int32)
            et = np.full((nq, self.n), -1, dtype=np.int32)
        ms = ctypes.c_float(0)
        if isinstance(mode, (list, tuple)):  # one mode per search: opalb200_db_search_batch_modes
            modes = np.array([MODES[m] if isinstance(m, str) else int(m) for m in mode], dtype=np.int32)
            assert len(modes) == nq
            rc = self.eng.lib.opalb200_db_search_batch_modes(
                self.handle, nq, ptrs.ctypes.data, qlens.ctypes.data, modes.ctypes.data, int(gap_open), int(gap_ext), sm.ctypes.data,
                int(alphabet_length), int(search_type), sc.ctypes.data, eq.ctypes.data, et.ctypes.data,
                int(in_flight), ctypes.byref(ms))
            return rc, sc, eq, et, float(ms.value)
        if isinstance(mode, str):
            mode = MODES[mode]
        rc = self.eng.lib.opalb200_db_search_batch(
            self.handle, nq, ptrs.ctypes.data, qlens.ctypes.data, int(gap_open), int(gap_ext), sm.ctypes.data,
            int(alphabet_length), int(search_type), int(mode), sc.ctypes.data, eq.ctypes.data, et.ctypes.data,
            int(in_flight), ctypes.byref(ms))
        return rc, sc, eq, et, float(ms.value)

    def search_results(self, query, gap_open, gap_ext, score_matrix, alphabet_length, search_type, mode, results=None):
        """opalb200_db_search_results: opalSearchDatabase's records against the resident database. Returns (rc, results)."""
        query = np.ascontiguousarray(query, dtype=np.uint8)
        sm = np.ascontiguousarray(score_matrix, dtype=np.int32).ravel()
        if results is None:
            results = new_results(self.n)
        rp = result_pointers(results)
        qbuf = query if query.size else np.zeros(1, dtype=np.uint8)
        if isinstance(mode, str):
            mode = MODES[mode]
        rc = self.eng.lib.opalb200_db_search_results(
            self.handle, qbuf.ctypes.data, int(query.size), int(gap_open), int(gap_ext), sm.ctypes.data,
            int(alphabet_length), rp.ctypes.data, int(search_type), int(mode))
        return rc, results

    def search_topk(self, query, gap_open, gap_ext, score_matrix, alphabet_length, search_type, mode, k):
        """opalb200_db_search_topk. Returns (rc, indices[found], results[found])."""
        query = np.ascontiguousarray(query, dtype=np.uint8)
        sm = np.ascontiguousarray(score_matrix, dtype=np.int32).ravel()
        results = new_results(max(int(k), 1))
        rp = result_pointers(results)
        idx = np.full(max(int(k), 1), -1, dtype=np.int32)
        found = ctypes.c_int(0)
        qbuf = query if query.size else np.zeros(1, dtype=np.uint8)
        if isinstance(mode, str):
            mode = MODES[mode]
        rc = self.eng.lib.opalb200_db_search_topk(
            self.handle, qbuf.ctypes.data, int(query.size), int(gap_open), int(gap_ext), sm.ctypes.data,
            int(alphabet_length), int(search_type), int(mode), int(k), idx.ctypes.data, rp.ctypes.data, ctypes.byref(found))
        return rc, idx[:found.value], results[:found.value]

    def devices(self):
        return int(self.eng.lib.opalb200_db_devices(self.handle))

    def last_stats(self):
        v = [ctypes.c_int(0) for _ in range(7)]
        self.eng.lib.opalb200_db_last_stats(self.handle, *[ctypes.byref(x) for x in v])
        d = dict(zip(("kernel_launches", "rerun32", "G", "R", "passes", "warps_per_partition", "groups"), (x.value for x in v)))
        d["folded"] = int(self.eng.lib.opalb200_db_last_folded(self.handle))
        d["chained"] = int(self.eng.lib.opalb200_db_last_chained(self.handle))
        return d

    def close(self):
        if self.handle:
            self.eng.lib.opalb200_db_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
