// api.cu -- the C ABI of libopal_b200.so: opal.h (drop-in) and opal_b200.h (resident handle).
//
// opalSearchDatabase mirrors reference src/opal.cpp:1435-1519 step by step -- skip mask from
// prefilled records, score/end search, early return on error, then either the alignment stage
// or the "no alignment" field fill -- with the SIMD passes replaced by DeviceDb::search.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/opal.h"
#include "../../include/opal_b200.h"
#include "align.h"
#include "engine.h"

using namespace opalb200;

// The CUDA "current device" is per-thread state of the caller: entry points put it back on return.
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) saved = -1; }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

static int default_device() {
    const char* e = getenv("OPAL_B200_DEVICE");
    return e ? atoi(e) : 0;
}

extern "C" {

void opalInitSearchResult(OpalSearchResult* r) {  // reference src/opal.cpp:1549-1555
    r->scoreSet = 0;
    r->startLocationTarget = r->startLocationQuery = -1;
    r->endLocationTarget = r->endLocationQuery = -1;
    r->alignment = NULL;
    r->alignmentLength = 0;
}

int opalSearchResultIsEmpty(const OpalSearchResult r) { return !r.scoreSet; }  // :1557-1559

void opalSearchResultSetScore(OpalSearchResult* r, int score) {  // :1561-1564
    r->scoreSet = 1;
    r->score = score;
}

// The body of opalSearchDatabase once the database is (or can be made) resident.  `ddb` may be NULL: it is
// then created from db / dbSeqLengths if any entry needs work.  db / dbSeqLengths may be NULL for handle calls.
static int search_into_results(DeviceDb* ddb, const unsigned char* query, int queryLength, unsigned char* const* db, int dbLength,
                               const int* dbSeqLengths, int gapOpen, int gapExt, const int* scoreMatrix, int alphabetLength,
                               OpalSearchResult* results[], const int searchType, int mode) {
    if (mode != OPAL_MODE_NW && mode != OPAL_MODE_HW && mode != OPAL_MODE_OV && mode != OPAL_MODE_SW)
        return OPAL_ERR_INVALID_MODE;  // :1469-1473, results untouched
    if (dbLength <= 0) return 0;

    // Entries that already hold what this search level needs are not recomputed (:1446-1451).
    // (The loops over all records run on the host pool for large databases: half a million 40-byte records cost
    // milliseconds per pass on one core.)
    std::vector<unsigned char> skip(dbLength);
    std::atomic<bool> anyWorkFlag(false);
    parallel_for(dbLength, 65536, [&](long long lo, long long hi) {
        bool any = false;
        for (long long i = lo; i < hi; i++) {
            const OpalSearchResult* r = results[i];
            skip[i] = r->scoreSet && (searchType == OPAL_SEARCH_SCORE || (r->endLocationQuery >= 0 && r->endLocationTarget >= 0));
            any |= !skip[i];
        }
        if (any) anyWorkFlag.store(true);
    });
    const bool anyWork = anyWorkFlag.load();
    const int wantEnd = searchType != OPAL_SEARCH_SCORE;
    const bool trace = getenv("OPAL_B200_TRACE") != nullptr;  // phase timings of the call on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto t0 = now();
    DeviceDb* owned = nullptr;
    if (!ddb && (anyWork || searchType == OPAL_SEARCH_ALIGNMENT)) {
        ddb = owned = DeviceDb::create(db, dbLength, dbSeqLengths, default_device());
        if (!ddb) return OPAL_ERR_NO_SIMD_SUPPORT;
    }
    const auto t1 = now();
    int status = 0;
    if (anyWork) {
        std::vector<int> sc(dbLength), eq(dbLength, -1), et(dbLength, -1);
        status = ddb->search(query, queryLength, gapOpen, gapExt, scoreMatrix, alphabetLength, wantEnd, mode, skip.data(),
                             sc.data(), eq.data(), et.data(), nullptr);
        if (status == 0) {
            parallel_for(dbLength, 65536, [&](long long lo, long long hi) {
                for (long long i = lo; i < hi; i++) {
                    if (skip[i]) continue;
                    opalSearchResultSetScore(results[i], sc[i]);
                    results[i]->endLocationQuery = wantEnd ? eq[i] : -1;  // :420-426, 869-905
                    results[i]->endLocationTarget = wantEnd ? et[i] : -1;
                    if (searchType != OPAL_SEARCH_ALIGNMENT) {  // :1508-1515, same pass
                        results[i]->alignment = NULL;
                        results[i]->alignmentLength = -1;
                        results[i]->startLocationQuery = results[i]->startLocationTarget = -1;
                    }
                }
            });
        }
    }
    const auto t2 = now();
    if (status == 0 && searchType == OPAL_SEARCH_ALIGNMENT)
        status = align_database(ddb, query, queryLength, db, dbLength, dbSeqLengths, gapOpen, gapExt, scoreMatrix, alphabetLength,
                                results, mode);
    const auto t3 = now();
    delete owned;
    if (trace)
        fprintf(stderr, "[opal-b200] pack+upload %.3f ms, search %.3f ms, alignment %.3f ms, release %.3f ms\n", ms(t0, t1), ms(t1, t2),
                ms(t2, t3), ms(t3, now()));
    if (status) return status;  // :1473
    if (searchType != OPAL_SEARCH_ALIGNMENT) {  // :1508-1515: every entry, skipped ones included (the others were done above)
        parallel_for(dbLength, 65536, [&](long long lo, long long hi) {
            for (long long i = lo; i < hi; i++) {
                if (anyWork && !skip[i]) continue;
                results[i]->alignment = NULL;
                results[i]->alignmentLength = -1;
                results[i]->startLocationQuery = -1;
                results[i]->startLocationTarget = -1;
            }
        });
    }
    return 0;
}

int opalSearchDatabase(unsigned char query[], int queryLength, unsigned char* db[], int dbLength, int dbSeqLengths[],
                       int gapOpen, int gapExt, int* scoreMatrix, int alphabetLength, OpalSearchResult* results[],
                       const int searchType, int mode, int overflowMethod) {
    DeviceGuard guard;
    (void)overflowMethod;  // OPAL_OVERFLOW_SIMPLE / _BUCKETS only schedule the reference's passes; results are equal
    return search_into_results(nullptr, query, queryLength, db, dbLength, dbSeqLengths, gapOpen, gapExt, scoreMatrix,
                               alphabetLength, results, searchType, mode);
}

int opalSearchDatabaseRescore(unsigned char query[], int queryLength, unsigned char* db[], int dbLength, int dbSeqLengths[],
                              int gapOpen, int gapExt, int* scoreMatrix, int alphabetLength, OpalSearchResult* results[],
                              const int searchType, int mode, int overflowMethod) {
    return opalSearchDatabase(query, queryLength, db, dbLength, dbSeqLengths, gapOpen, gapExt, scoreMatrix, alphabetLength,
                              results, searchType, mode, overflowMethod);
}

int opalSearchDatabaseCharSW(unsigned char query[], int queryLength, unsigned char** db, int dbLength, int dbSeqLengths[],
                             int gapOpen, int gapExt, int* scoreMatrix, int alphabetLength, OpalSearchResult* results[]) {
    DeviceGuard guard;
    // reference src/opal.cpp:1522-1546: SW scores that fit 8 bits; the others come back unset (-1).
    if (dbLength <= 0) return 0;
    bool argsFit = !(gapOpen < -128 || 127 < gapOpen || gapExt < -128 || 127 < gapExt);  // :178-180
    for (int i = 0; argsFit && i < alphabetLength * alphabetLength; i++)
        if (scoreMatrix[i] < -128 || 127 < scoreMatrix[i]) argsFit = false;               // :188-193
    std::vector<int> sc(dbLength, -1);
    int rc = argsFit ? 0 : OPAL_ERR_OVERFLOW;
    if (argsFit) {
        DeviceDb* ddb = DeviceDb::create(db, dbLength, dbSeqLengths, default_device());
        if (!ddb) return OPAL_ERR_NO_SIMD_SUPPORT;
        const int st = ddb->search(query, queryLength, gapOpen, gapExt, scoreMatrix, alphabetLength, 0, OPAL_MODE_SW, nullptr,
                                   sc.data(), nullptr, nullptr, nullptr);
        delete ddb;
        if (st == OPAL_ERR_NO_SIMD_SUPPORT) return st;
        if (st) std::fill(sc.begin(), sc.end(), -1);
    }
    for (int i = 0; i < dbLength; i++) {
        if (argsFit && sc[i] >= 0 && sc[i] <= 127) {
            opalSearchResultSetScore(results[i], sc[i]);
            results[i]->endLocationQuery = results[i]->endLocationTarget = -1;  // :423-426
        } else {
            results[i]->score = -1;  // :1538-1541
            results[i]->scoreSet = 0;
            rc = OPAL_ERR_OVERFLOW;
        }
    }
    return rc;
}

// ------------------------------------------------------------------ opal_b200.h
int opalb200_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

const char* opalb200_last_error(void) { return last_error(); }

void opalb200_trim_cache(void) {
    DeviceGuard guard;
    trim_cache();
}

OpalB200Db* opalb200_db_create(unsigned char* db[], int dbLength, const int dbSeqLengths[], int device) {
    DeviceGuard guard;
    return reinterpret_cast<OpalB200Db*>(DeviceDb::create(db, dbLength, dbSeqLengths, device));
}

OpalB200Db* opalb200_db_create_sorted(const unsigned char* residues, const int sortedLengths[], const int order[], int dbLength,
                                      int device) {
    DeviceGuard guard;
    if (dbLength < 0 || (dbLength > 0 && (!residues || !sortedLengths))) { set_error("invalid packed database"); return nullptr; }
    return reinterpret_cast<OpalB200Db*>(DeviceDb::create_sorted(residues, sortedLengths, order, dbLength, device));
}

void opalb200_db_destroy(OpalB200Db* h) {
    DeviceGuard guard;
    delete reinterpret_cast<DeviceDb*>(h);
}

int opalb200_db_length(const OpalB200Db* h) { return reinterpret_cast<const DeviceDb*>(h)->size(); }

long long opalb200_db_residues(const OpalB200Db* h) { return reinterpret_cast<const DeviceDb*>(h)->residues(); }

int opalb200_db_search(OpalB200Db* h, const unsigned char query[], int queryLength, int gapOpen, int gapExt,
                       const int* scoreMatrix, int alphabetLength, int searchType, int mode, const unsigned char* skip,
                       int* scores, int* endQuery, int* endTarget, float* deviceMs) {
    if (!h || !scores) return OPAL_ERR_NO_SIMD_SUPPORT;
    DeviceGuard guard;
    const int wantEnd = searchType != OPAL_SEARCH_SCORE && endQuery && endTarget;
    return reinterpret_cast<DeviceDb*>(h)->search(query, queryLength, gapOpen, gapExt, scoreMatrix, alphabetLength, wantEnd,
                                                  mode, skip, scores, endQuery, endTarget, deviceMs);
}

int opalb200_db_search_batch(OpalB200Db* h, int numQueries, const unsigned char* const queries[], const int queryLengths[],
                             int gapOpen, int gapExt, const int* scoreMatrix, int alphabetLength, int searchType, int mode,
                             int* scores, int* endQuery, int* endTarget, int inFlight, float* batchMs) {
    if (!h || !scores || (numQueries > 0 && (!queries || !queryLengths))) return OPAL_ERR_NO_SIMD_SUPPORT;
    DeviceGuard guard;
    const int wantEnd = searchType != OPAL_SEARCH_SCORE && endQuery && endTarget;
    return reinterpret_cast<DeviceDb*>(h)->search_batch(numQueries, queries, queryLengths, gapOpen, gapExt, scoreMatrix,
                                                        alphabetLength, wantEnd, mode, scores, endQuery, endTarget,
                                                        inFlight <= 0 ? 3 : inFlight, batchMs);
}

int opalb200_db_search_results(OpalB200Db* h, const unsigned char query[], int queryLength, int gapOpen, int gapExt,
                               const int* scoreMatrix, int alphabetLength, OpalSearchResult* results[], int searchType, int mode) {
    if (!h || !results) return OPAL_ERR_NO_SIMD_SUPPORT;
    DeviceGuard guard;
    DeviceDb* ddb = reinterpret_cast<DeviceDb*>(h);
    return search_into_results(ddb, query, queryLength, nullptr, ddb->size(), nullptr, gapOpen, gapExt, scoreMatrix, alphabetLength,
                               results, searchType, mode);
}

int opalb200_db_search_topk(OpalB200Db* h, const unsigned char query[], int queryLength, int gapOpen, int gapExt,
                            const int* scoreMatrix, int alphabetLength, int searchType, int mode, int k, int* indices,
                            OpalSearchResult* results[], int* found) {
    if (found) *found = 0;
    if (!h || !indices || !results || k < 0) return OPAL_ERR_NO_SIMD_SUPPORT;
    DeviceGuard guard;
    DeviceDb* ddb = reinterpret_cast<DeviceDb*>(h);
    const int n = ddb->size(), kk = std::min(k, n);
    std::vector<int> sc((size_t)n), eq((size_t)n, -1), et((size_t)n, -1);
    const int wantEnd = searchType != OPAL_SEARCH_SCORE;
    int status = ddb->search(query, queryLength, gapOpen, gapExt, scoreMatrix, alphabetLength, wantEnd, mode, nullptr, sc.data(),
                             eq.data(), et.data(), nullptr);
    if (status) return status;
    // the k best: score descending, caller index ascending among equals
    std::vector<int> idx((size_t)n);
    for (int i = 0; i < n; i++) idx[i] = i;
    auto better = [&](int a, int b) { return sc[a] != sc[b] ? sc[a] > sc[b] : a < b; };
    if (kk < n) std::nth_element(idx.begin(), idx.begin() + kk, idx.end(), better);
    std::sort(idx.begin(), idx.begin() + kk, better);
    for (int j = 0; j < kk; j++) {
        const int i = idx[j];
        indices[j] = i;
        opalInitSearchResult(results[j]);
        opalSearchResultSetScore(results[j], sc[i]);
        results[j]->endLocationQuery = wantEnd ? eq[i] : -1;
        results[j]->endLocationTarget = wantEnd ? et[i] : -1;
        results[j]->alignmentLength = -1;  // as opalSearchDatabase leaves it below OPAL_SEARCH_ALIGNMENT (:1508-1515)
    }
    if (found) *found = kk;
    if (searchType == OPAL_SEARCH_ALIGNMENT && kk > 0)
        status = align_database(ddb, query, queryLength, nullptr, kk, nullptr, gapOpen, gapExt, scoreMatrix, alphabetLength, results,
                                mode, indices);
    return status;
}

void opalb200_db_last_stats(const OpalB200Db* h, int* kernelLaunches, int* rerun32, int* G, int* R, int* passes,
                            int* warpsPerPartition, int* groups) {
    const SearchStats& s = reinterpret_cast<const DeviceDb*>(h)->stats();
    if (kernelLaunches) *kernelLaunches = s.kernelLaunches;
    if (rerun32) *rerun32 = s.rerun32;
    if (G) *G = s.G;
    if (R) *R = s.R;
    if (passes) *passes = s.passes;
    if (warpsPerPartition) *warpsPerPartition = s.warpsPerPartition;
    if (groups) *groups = s.groups;
}

int opalb200_db_last_folded(const OpalB200Db* h) { return reinterpret_cast<const DeviceDb*>(h)->stats().foldedTasks; }

double opalb200_measure_dpx_peak(int device, double* threadInstrPerSec, float* ms) {
    DeviceGuard guard;
    return measure_dpx_peak(device, 0, threadInstrPerSec, ms);
}

double opalb200_measure_dpx_peak_mix(int device, int mix, double* threadInstrPerSec, float* ms) {
    DeviceGuard guard;
    return measure_dpx_peak(device, mix != 0 ? 1 : 0, threadInstrPerSec, ms);
}

}  // extern "C"
