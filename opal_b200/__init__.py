"""opal-b200: the database-search hot path of Martinsos/opal on NVIDIA B200 (sm_100a).

The product is ``opal_b200/csrc/libopal_b200.so`` -- hand-written CUDA kernels behind the
reference's own C API (include/opal.h).  This package is the thin Python mirror of that API
used by tests and bench.py; it has no CPU implementation and raises when the CUDA library is
missing.
"""
from .capi import (  # noqa: F401
    OPAL_ALIGN_DEL, OPAL_ALIGN_INS, OPAL_ALIGN_MATCH, OPAL_ALIGN_MISMATCH, OPAL_ERR_INVALID_MODE,
    OPAL_ERR_NO_SIMD_SUPPORT, OPAL_ERR_OVERFLOW, OPAL_MODE_HW, OPAL_MODE_NW, OPAL_MODE_OV,
    OPAL_MODE_SW, OPAL_OVERFLOW_BUCKETS, OPAL_OVERFLOW_SIMPLE, OPAL_SEARCH_ALIGNMENT,
    OPAL_SEARCH_SCORE, OPAL_SEARCH_SCORE_END, MODES, OpalCLibrary, SequenceDB, free_alignments,
    get_alignment, new_results, result_pointers)
from . import sharding  # noqa: F401,E402
