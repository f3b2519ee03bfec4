/*
 * opal.h -- drop-in C boundary of opal-b200.
 *
 * This header declares exactly the C API that callers of Martinsos/opal bind
 * (reference: src/opal.h:12-169): the same symbol names, argument order,
 * constant values and OpalSearchResult layout, so that reference callers
 * (src/opal_aligner.cpp:158-160, src/test.cpp:97-99, README.md:64-66) compile
 * and link against libopal_b200.so without edits.  The implementation behind
 * it is the sm_100a CUDA path in opal_b200/csrc/, not the SSE4.1/AVX2 code.
 *
 * One addition: opalSearchDatabaseRescore (same signature as
 * opalSearchDatabase), the entry point for the result-reuse rule of
 * reference src/opal.h:118-122 / src/opal.cpp:1446-1451.
 */
#ifndef OPAL_H
#define OPAL_H

#ifdef __cplusplus
extern "C" {
#endif

/* Return codes (reference src/opal.h:17-19). 0 means success. */
#define OPAL_ERR_OVERFLOW 1        /* a score does not fit the 32-bit range the library guarantees */
#define OPAL_ERR_NO_SIMD_SUPPORT 2 /* reference: no SSE4.1/AVX2; here: no usable sm_100 device / CUDA failure */
#define OPAL_ERR_INVALID_MODE 3    /* mode is none of NW/HW/OV/SW */
/* Addition of this library (the reference reads out of bounds instead, src/opal.cpp:265): alphabetLength outside
 * [1, 256], a residue code >= alphabetLength in the query or the database, a negative gap penalty.  The text is
 * available from opalb200_last_error() (opal_b200.h). */
#define OPAL_ERR_INVALID_ARGUMENT 4

/* Alignment modes (reference src/opal.h:22-25). */
#define OPAL_MODE_NW 0 /* global */
#define OPAL_MODE_HW 1 /* semi-global: target prefix/suffix free */
#define OPAL_MODE_OV 2 /* overlap: prefix/suffix of both sequences free */
#define OPAL_MODE_SW 3 /* local */

/* Overflow scheduling (reference src/opal.h:28-29). Results never depend on it. */
#define OPAL_OVERFLOW_SIMPLE 0
#define OPAL_OVERFLOW_BUCKETS 1

/* Search levels (reference src/opal.h:32-34). */
#define OPAL_SEARCH_SCORE 0     /* score only */
#define OPAL_SEARCH_SCORE_END 1 /* score + end location */
#define OPAL_SEARCH_ALIGNMENT 2 /* score + end + start location + operation string */

/* Alignment operation codes (reference src/opal.h:37-40). */
#define OPAL_ALIGN_MATCH 0
#define OPAL_ALIGN_DEL 1      /* query residue aligned to a gap in the target */
#define OPAL_ALIGN_INS 2      /* target residue aligned to a gap in the query */
#define OPAL_ALIGN_MISMATCH 3

/*
 * One result record per database sequence (reference src/opal.h:47-74).
 * Field order and types are part of the ABI: 6 ints, a pointer, an int.
 * Among equally scoring end cells the one with the smallest target position,
 * then the smallest query position, is reported.  `alignment` is malloc()ed
 * by the library and owned (free()d) by the caller.
 */
struct OpalSearchResult {
    int scoreSet;            /* 1 once the record holds at least a score */
    int score;
    int endLocationTarget;   /* 0-based, -1 when unset */
    int endLocationQuery;    /* 0-based, -1 when unset */
    int startLocationTarget; /* 0-based, -1 when unset */
    int startLocationQuery;  /* 0-based, -1 when unset */
    unsigned char* alignment; /* OPAL_ALIGN_* codes from alignment start to end, or NULL */
    int alignmentLength;
};
#ifndef __cplusplus
typedef struct OpalSearchResult OpalSearchResult;
#endif

/* Reset a record to "empty" (reference src/opal.h:82, src/opal.cpp:1549-1555). */
void opalInitSearchResult(OpalSearchResult* result);

/* 1 if the record holds no score (reference src/opal.h:87, src/opal.cpp:1557-1559). */
int opalSearchResultIsEmpty(const OpalSearchResult result);

/* Store a score and mark the record non-empty (reference src/opal.h:89, src/opal.cpp:1561-1564). */
void opalSearchResultSetScore(OpalSearchResult* result, int score);

/*
 * Align `query` against every sequence of `db` with affine gaps (a gap of
 * length n costs gapOpen + (n-1)*gapExt) and fill results[i] for each i
 * (reference src/opal.h:150-154, src/opal.cpp:1435-1519).
 *
 * Sequences are arrays of alphabet indices in [0, alphabetLength);
 * scoreMatrix[q*alphabetLength + t] scores query letter q against target
 * letter t.  A record that already holds a score (and, for searchType above
 * OPAL_SEARCH_SCORE, an end location) is not recomputed; under
 * OPAL_SEARCH_ALIGNMENT its score and end are used to derive start+alignment.
 *
 * Returns 0, or OPAL_ERR_OVERFLOW / OPAL_ERR_NO_SIMD_SUPPORT /
 * OPAL_ERR_INVALID_MODE / OPAL_ERR_INVALID_ARGUMENT.
 *
 * Inputs the reference leaves undefined (src/opal.cpp:263-265, 330-331) are defined here: a zero-length target (or
 * query) scores as an alignment against nothing -- 0 for SW and OV, the cost of one gap over the other sequence for NW
 * (and for HW when the target is empty) -- with end location (queryLength - 1, targetLength - 1), i.e. -1 on the empty
 * side; OPAL_SEARCH_ALIGNMENT leaves such a record without start, end and alignment (alignmentLength 0).
 *
 * With several GPUs (environment OPAL_B200_DEVICES=all or a comma-separated list of ordinals) one call deals the
 * database over those devices, balanced by residue count, and searches them concurrently; the records are the same.
 */
int opalSearchDatabase(
    unsigned char query[], int queryLength, unsigned char* db[], int dbLength,
    int dbSeqLengths[], int gapOpen, int gapExt, int* scoreMatrix,
    int alphabetLength, OpalSearchResult* results[],
    const int searchType, int mode, int overflowMethod);

/*
 * 8-bit-range SW scores only (reference src/opal.h:162-165, src/opal.cpp:1522-1546):
 * a sequence whose SW score does not fit the signed 8-bit range gets
 * score = -1, scoreSet = 0 and the call returns OPAL_ERR_OVERFLOW.
 */
int opalSearchDatabaseCharSW(
    unsigned char query[], int queryLength, unsigned char** db, int dbLength,
    int dbSeqLengths[], int gapOpen, int gapExt, int* scoreMatrix,
    int alphabetLength, OpalSearchResult* results[]);

/*
 * Addition of this library: explicit name for the result-reuse path.  Same
 * arguments and semantics as opalSearchDatabase; intended to be called with
 * records prefilled by an earlier, cheaper search level.
 */
int opalSearchDatabaseRescore(
    unsigned char query[], int queryLength, unsigned char* db[], int dbLength,
    int dbSeqLengths[], int gapOpen, int gapExt, int* scoreMatrix,
    int alphabetLength, OpalSearchResult* results[],
    const int searchType, int mode, int overflowMethod);

#ifdef __cplusplus
}
#endif

#endif /* OPAL_H */
