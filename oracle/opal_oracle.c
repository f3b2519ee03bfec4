/*
 * opal_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar, single-threaded CPU restatement of the algorithm behind
 * Martinsos/opal's database search, used only as the parity checker by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing
 * under opal_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * (a) the known answers of the reference's own programs (README example,
 * `opal_aligner` on test_data, `./test` maxima; see tests/golden/) and
 * (b) when oracle/_ref/libopal_ref.so is present, the unmodified reference
 * compiled from /root/reference/src/opal.cpp, field by field on random inputs.
 *
 * What is restated (reference file:line):
 *   score + end location, SW ............ src/test.cpp:199-251, src/opal.cpp:280-328,384-402
 *   score + end location, NW/HW/OV ...... src/test.cpp:253-328, src/opal.cpp:671-689,716-774,843-905
 *   band for the alignment stage ........ src/opal.cpp:1046-1179
 *   reverse DP + traceback .............. src/opal.cpp:1236-1431
 *   orchestration, reuse rule, fills .... src/opal.cpp:1435-1519
 *   8-bit-only SW entry ................. src/opal.cpp:1522-1546
 *   result helpers ...................... src/opal.cpp:1549-1564
 *   alignment replay check .............. src/test.cpp:348-422
 *
 * Deliberate differences from the reference, all on inputs where the
 * reference has undefined behaviour:
 *   - arithmetic is 64-bit with a non-wrapping -infinity, so the NW/HW/OV
 *     "32-bit path" gives the mathematically correct score (Q1);
 *   - SW with score 0 under SCORE_END reports end = (-1,-1) (Q2), as
 *     src/test.cpp:242-244 does;
 *   - the skip mask is applied to exactly the prefilled entries (Q7);
 *   - the alignment stage never indexes outside its matrix, never divides by
 *     zero and falls back to the full band when the reference's band would
 *     not exist (Q3, Q4, Q8).
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "../include/opal.h"

typedef long long i64;
#define NEG_INF (LLONG_MIN / 4)

static i64 max2(i64 a, i64 b) { return a > b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* ---------------------------------------------------------------- helpers */

void opalInitSearchResult(OpalSearchResult* r) { /* src/opal.cpp:1549-1555 */
    r->scoreSet = 0;
    r->endLocationTarget = r->endLocationQuery = -1;
    r->startLocationTarget = r->startLocationQuery = -1;
    r->alignment = NULL;
    r->alignmentLength = 0;
}

int opalSearchResultIsEmpty(const OpalSearchResult r) { return !r.scoreSet; } /* :1557-1559 */

void opalSearchResultSetScore(OpalSearchResult* r, int score) { /* :1561-1564 */
    r->scoreSet = 1;
    r->score = score;
}

/* Argument range rule of the widest (int) pass: src/opal.cpp:178-198, 610-630. */
static int args_out_of_int_half_range(int Go, int Ge, const int* S, int A) {
    if (Go <= INT_MIN / 2 || INT_MAX / 2 <= Go || Ge <= INT_MIN / 2 || INT_MAX / 2 <= Ge) return 1;
    for (int i = 0; i < A * A; i++)
        if (S[i] <= INT_MIN / 2 || INT_MAX / 2 <= S[i]) return 1;
    return 0;
}

/* ------------------------------------------------ score + end, one target */

/*
 * Plain Gotoh over the whole Q x T matrix, column by column (target index c
 * outer, query index r inner), exactly the visiting order of the reference.
 * Returns 0, or 1 when the score (SW: any diagonal sum, src/opal.cpp:339-351)
 * leaves the int range.
 *
 * End-location key (src/opal.h:43-45): among the cells the mode allows an
 * alignment to end in, the maximal H; ties -> smallest c, then smallest r.
 *   SW: strict '>' while scanning column-major (src/test.cpp:218-222).
 *   NW: the last cell (src/opal.cpp:873-876).
 *   HW: last row, first column attaining the max (src/opal.cpp:824-832,877-880).
 *   OV: last row first; the last column only wins with a strictly larger
 *       value, then its first maximal row (src/opal.cpp:881-904, test.cpp:293-313).
 */
int oracle_score_end(const unsigned char* q, int Q, const unsigned char* t, int T,
                     int Go, int Ge, const int* S, int A, int mode,
                     int* score, int* endQ, int* endT) {
    i64* prevH = (i64*)malloc(sizeof(i64) * (size_t)(Q > 0 ? Q : 1));
    i64* prevE = (i64*)malloc(sizeof(i64) * (size_t)(Q > 0 ? Q : 1));
    int rc = 0;
    i64 best;
    int bq = -1, bt = -1;

    if (mode == OPAL_MODE_SW) {
        for (int r = 0; r < Q; r++) prevH[r] = prevE[r] = 0; /* src/opal.cpp:247-249 */
        best = 0;
        for (int c = 0; c < T; c++) {
            i64 uF = 0, uH = 0, ulH = 0; /* src/opal.cpp:272-273 */
            for (int r = 0; r < Q; r++) {
                i64 E = max2(prevH[r] - Go, prevE[r] - Ge);
                i64 F = max2(uH - Go, uF - Ge);
                i64 d = ulH + S[q[r] * A + t[c]];
                if (d > INT_MAX) rc = 1;
                i64 H = max2(max2(E, F), max2(d, 0));
                if (H > best) { best = H; bq = r; bt = c; }
                uF = F; uH = H; ulH = prevH[r];
                prevH[r] = H; prevE[r] = E;
            }
        }
    } else {
        /* Column -1: src/opal.cpp:671-682.  Row -1: src/opal.cpp:714-732. */
        for (int r = 0; r < Q; r++) {
            prevH[r] = (mode == OPAL_MODE_OV) ? 0 : -(i64)Go - (i64)r * Ge;
            prevE[r] = NEG_INF;
        }
        best = NEG_INF;
        i64 H = NEG_INF;
        for (int c = 0; c < T; c++) {
            i64 uF = NEG_INF, uH, ulH;
            if (mode == OPAL_MODE_NW) {
                uH = -(i64)Go - (i64)c * Ge;
                ulH = (c == 0) ? 0 : uH + Ge;
            } else {
                uH = ulH = 0;
            }
            for (int r = 0; r < Q; r++) {
                i64 E = max2(prevH[r] - Go, prevE[r] - Ge);
                i64 F = max2(uH - Go, uF - Ge);
                H = max2(max2(E, F), ulH + S[q[r] * A + t[c]]);
                if (mode == OPAL_MODE_OV && c == T - 1 && H > best) { best = H; bq = r; bt = c; }
                uF = F; uH = H; ulH = prevH[r];
                prevH[r] = H; prevE[r] = E;
            }
            if (mode != OPAL_MODE_NW && Q > 0 && H > best) { best = H; bq = Q - 1; bt = c; }
        }
        if (mode == OPAL_MODE_NW) {
            if (T > 0 && Q > 0) best = H;
            else if (Q > 0) best = -(i64)Go - (i64)(Q - 1) * Ge; /* empty target: query against gaps */
            else if (T > 0) best = -(i64)Go - (i64)(T - 1) * Ge;  /* empty query: target against gaps */
            else best = 0;
            bq = Q - 1; bt = T - 1;
        } else if (best == NEG_INF) { /* empty target or query: nothing aligned */
            best = (mode == OPAL_MODE_HW && Q > 0) ? -(i64)Go - (i64)(Q - 1) * Ge : 0;
            bq = Q - 1; bt = T - 1;
        }
        if (best > INT_MAX || best < INT_MIN) rc = 1;
    }
    free(prevH); free(prevE);
    if (rc) return 1;
    *score = (int)best;
    if (mode == OPAL_MODE_SW && best == 0) { *endQ = -1; *endT = -1; }
    else { *endQ = bq; *endT = bt; }
    return 0;
}

/* ------------------------------------------------------------ band borders */

static int gap_penalty(int len, int Go, int Ge) { return len > 0 ? Go + Ge * (len - 1) : 0; } /* :1046-1052 */

/* Truncating division as in the reference; a zero divisor marks "no usable bound". */
static int tdiv(i64 num, i64 den, int* bad) {
    if (den == 0) { *bad = 1; return 0; }
    i64 v = num / den;
    if (v > INT_MAX) v = INT_MAX;
    if (v < INT_MIN) v = INT_MIN;
    return (int)v;
}

static int bottom_border_ov(int k, int Q, int T, int Go, int Ge, int M, int* bad) { /* :1057-1070 */
    int border = 0;
    border = imax(border, imin(Q - T, tdiv(-1 * ((i64)k + Go - Ge - (i64)M * T), Ge, bad)));
    int cand = tdiv(-1 * ((i64)k - (i64)M * Q + Go - Ge), (i64)Ge + M, bad);
    if (cand > Q - T) border = imax(border, cand);
    return imin(border, Q - 1);
}

static int top_border_hw(int k, int Q, int T, int Go, int Ge, int M, int* bad) { /* :1072-1085 */
    int border = 0;
    int v = tdiv(-1 * ((i64)k - (i64)M * Q + Go), Ge, bad);
    border = imax(border, imin(T - Q, v == INT_MAX ? v : v + 1));
    int cand = tdiv(-1 * ((i64)k - (i64)T * M + 2 * (i64)Go + (i64)Ge * (Q - T - 2)), 2 * (i64)Ge + M, bad);
    if (cand > T - Q) border = imax(border, cand);
    return imin(border, T - 1);
}

static int bottom_border_hw(int k, int Q, int T, int Go, int Ge, int M, int* bad) { /* :1087-1102 */
    int border = 0;
    int cand = tdiv(-1 * ((i64)k + Go - Ge - (i64)Q * M), (i64)Ge + M, bad);
    if (cand >= Q - T) border = imax(border, cand);
    if (-2 * (i64)Go - (i64)Ge * (Q - T - 2) + (i64)M * T >= k) border = imax(border, Q - T - 1);
    return imin(border, Q - 1);
}

static int bottom_border_nw(int k, int Q, int T, int Go, int Ge, int M, int* bad) { /* :1104-1124 */
    int border = 0;
    int cand = tdiv(-1 * ((i64)k + 2 * (i64)Go - (i64)M * Q + (i64)Ge * (T - Q - 2)), 2 * (i64)Ge + M, bad);
    if (cand > Q - T) border = imax(border, cand);
    if (Q - T <= tdiv(-1 * ((i64)k + Go - (i64)M * T - Ge), Ge, bad)) border = imax(border, Q - T);
    if (-2 * (i64)Go - (i64)Ge * (Q - T - 2) + (i64)M * T >= k) border = imax(border, Q - T - 1);
    return imin(border, Q - 1);
}

/*
 * (bottom, top) diagonal offsets of the band that holds every alignment of
 * score >= k (src/opal.cpp:1151-1179).  Returns 0 and the pair, or -1 when the
 * reference would report "no band" / divide by zero; callers then use the
 * whole matrix.
 */
int oracle_band_borders(int k, int mode, int Q, int T, int Go, int Ge, int M, int* bottom, int* top) {
    int bad = 0;
    int minQT = imin(Q, T);
    if (mode == OPAL_MODE_OV || mode == OPAL_MODE_SW) {
        if ((i64)M * minQT < k) return -1;
        *bottom = bottom_border_ov(k, Q, T, Go, Ge, M, &bad);
        *top = bottom_border_ov(k, T, Q, Go, Ge, M, &bad);
    } else if (mode == OPAL_MODE_HW) {
        if ((i64)M * minQT - gap_penalty(Q - minQT, Go, Ge) < k) return -1;
        *bottom = bottom_border_hw(k, Q, T, Go, Ge, M, &bad);
        *top = top_border_hw(k, Q, T, Go, Ge, M, &bad);
    } else if (mode == OPAL_MODE_NW) {
        if ((i64)M * minQT - gap_penalty(abs(Q - T), Go, Ge) < k) return -1;
        *bottom = bottom_border_nw(k, Q, T, Go, Ge, M, &bad);
        *top = bottom_border_nw(k, T, Q, Go, Ge, M, &bad);
    } else {
        return -1;
    }
    if (bad || *bottom < 0 || *bottom >= Q || *top < 0 || *top >= T) return -1;
    return 0;
}

/* ------------------------------------------------- reverse DP + traceback */

typedef struct { i64 H, E, F; } Cell;

/*
 * Restatement of findAlignment (src/opal.cpp:1236-1431): NW-anchored banded
 * Gotoh over (q, t) until the first column in which a cell that `mode` allows
 * as an end reaches `scoreLimit`; then a traceback with the fixed preference
 * E, then F, then the diagonal.  Outputs the end cell and the operations in
 * origin -> end order.  `*ops` is malloc()ed.  Returns 0, or -1 if no cell
 * reaches scoreLimit (the reference's behaviour is then undefined).
 */
int oracle_find_alignment(const unsigned char* q, int Q, const unsigned char* t, int T,
                          int Go, int Ge, const int* S, int A, int scoreLimit, int mode,
                          int* outScore, int* outEndQ, int* outEndT,
                          unsigned char** ops, int* opsLen) {
    int M = S[0];
    for (int i = 1; i < A * A; i++) if (S[i] > M) M = S[i]; /* arrayMax, :1029-1038 */
    int bottom, top;
    if (oracle_band_borders(scoreLimit, mode, Q, T, Go, Ge, M, &bottom, &top) != 0) {
        bottom = Q - 1; top = T - 1;
    }

    Cell** mat = (Cell**)calloc((size_t)T, sizeof(Cell*));
    Cell* init = (Cell*)malloc(sizeof(Cell) * (size_t)Q);
    for (int r = 0; r < Q; r++) { init[r].H = -(i64)Go - (i64)r * Ge; init[r].E = NEG_INF; init[r].F = NEG_INF; } /* :1266-1269 */

    Cell* prev = init;
    i64 maxScore = NEG_INF, H = NEG_INF;
    int c;
    for (c = 0; c < T && maxScore < scoreLimit; c++) { /* :1275 */
        Cell* col = mat[c] = (Cell*)malloc(sizeof(Cell) * (size_t)Q);
        int r0 = imax(0, c - top), r1 = imin(Q - 1, c + bottom); /* :1279-1280 */
        i64 uF, uH, ulH;
        if (r0 == 0) { uF = NEG_INF; uH = -(i64)Go - (i64)c * Ge; ulH = (c == 0) ? 0 : uH + Ge; } /* :1283-1286 */
        else { uH = uF = NEG_INF; ulH = prev[r0 - 1].H; }                                      /* :1288-1289 */
        for (int r = r0; r <= r1; r++) {
            i64 E = max2(prev[r].H > NEG_INF ? prev[r].H - Go : NEG_INF, prev[r].E > NEG_INF ? prev[r].E - Ge : NEG_INF);
            i64 F = max2(uH > NEG_INF ? uH - Go : NEG_INF, uF > NEG_INF ? uF - Ge : NEG_INF);
            i64 d = ulH > NEG_INF ? ulH + S[q[r] * A + t[c]] : NEG_INF;
            H = max2(E, max2(F, d));
            if (mode == OPAL_MODE_SW || (mode == OPAL_MODE_OV && c == T - 1)) maxScore = max2(maxScore, H); /* :1307-1310 */
            uF = F; uH = H; ulH = prev[r].H;
            col[r].H = H; col[r].E = E; col[r].F = F;
        }
        for (int r = 0; r < r0; r++) col[r].H = col[r].E = col[r].F = NEG_INF;       /* :1322-1324 */
        for (int r = r1 + 1; r < Q; r++) col[r].H = col[r].E = col[r].F = NEG_INF;   /* :1325-1327 */
        /* NB: H is the last IN-BAND row's value here, as in the reference (:1329-1331, quirk Q9). */
        if (mode == OPAL_MODE_HW || mode == OPAL_MODE_OV) maxScore = max2(maxScore, H);
        prev = col;
    }
    int lastCol = c - 1;
    int rc = 0, eq = -1, et = -1;
    i64 sc = NEG_INF;
    if (lastCol < 0) rc = -1;
    else if (mode == OPAL_MODE_NW) { sc = H; et = T - 1; eq = Q - 1; if (lastCol != T - 1) rc = -1; } /* :1341-1345 */
    else if (mode == OPAL_MODE_HW) { sc = maxScore; et = lastCol; eq = Q - 1; }                      /* :1346-1350 */
    else { /* SW, OV: first row of the stop column holding maxScore, :1351-1358 */
        sc = maxScore; et = lastCol;
        int r = 0;
        while (r < Q && mat[lastCol][r].H != maxScore) r++;
        if (r >= Q) rc = -1;
        eq = r;
    }
    if (rc == 0 && sc < scoreLimit && mode != OPAL_MODE_NW) rc = -1;

    unsigned char* a = NULL;
    int n = 0;
    if (rc == 0) {
        a = (unsigned char*)malloc((size_t)(Q + T + 2));
        int ri = eq, ci = et;
        int field = 0; /* 0 = H, 1 = E, 2 = F */
        while (ri >= 0 && ci >= 0) { /* :1372-1399 */
            Cell cell = mat[ci][ri];
            if (field == 0) {
                if (cell.H == cell.E) field = 1;
                else if (cell.H == cell.F) field = 2;
                else { a[n++] = (q[ri] == t[ci]) ? OPAL_ALIGN_MATCH : OPAL_ALIGN_MISMATCH; ci--; ri--; }
            } else if (field == 1) {
                i64 leftH = (ci > 0) ? mat[ci - 1][ri].H : init[ri].H; /* the reference reads matrix[-1] here (Q4) */
                field = (cell.E == leftH - Go) ? 0 : 1;
                a[n++] = OPAL_ALIGN_INS; ci--;
            } else {
                i64 upH = (ri > 0) ? mat[ci][ri - 1].H : -(i64)Go - (i64)ci * Ge;
                field = (cell.F == upH - Go) ? 0 : 2;
                a[n++] = OPAL_ALIGN_DEL; ri--;
            }
        }
        while (ri >= 0) { a[n++] = OPAL_ALIGN_DEL; ri--; } /* :1402-1405 */
        while (ci >= 0) { a[n++] = OPAL_ALIGN_INS; ci--; } /* :1406-1409 */
        for (int i = 0; i < n / 2; i++) { unsigned char x = a[i]; a[i] = a[n - 1 - i]; a[n - 1 - i] = x; } /* :1413 */
    }
    for (int i = 0; i <= lastCol; i++) free(mat[i]);
    free(mat); free(init);
    if (rc) return rc;
    *outScore = (int)sc; *outEndQ = eq; *outEndT = et; *ops = a; *opsLen = n;
    return 0;
}

/* --------------------------------------------------------------- the API */

static unsigned char* reversed_copy(const unsigned char* s, int n) { /* :1186-1192 */
    unsigned char* r = (unsigned char*)malloc((size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) r[i] = s[n - 1 - i];
    return r;
}

int opalSearchDatabase(unsigned char query[], int Q, unsigned char* db[], int N, int lens[],
                       int Go, int Ge, int* S, int A, OpalSearchResult* results[],
                       const int searchType, int mode, int overflowMethod) {
    (void)overflowMethod; /* scheduling only: results never depend on it */
    if (mode != OPAL_MODE_NW && mode != OPAL_MODE_HW && mode != OPAL_MODE_OV && mode != OPAL_MODE_SW)
        return OPAL_ERR_INVALID_MODE; /* :1469-1471 */
    if (args_out_of_int_half_range(Go, Ge, S, A)) return OPAL_ERR_OVERFLOW;

    int status = 0;
    for (int i = 0; i < N; i++) {
        OpalSearchResult* r = results[i];
        int skip = r->scoreSet && (searchType == OPAL_SEARCH_SCORE ||
                                   (r->endLocationQuery >= 0 && r->endLocationTarget >= 0)); /* :1448-1450 */
        if (skip) continue;
        int sc, eq, et;
        if (oracle_score_end(query, Q, db[i], lens[i], Go, Ge, S, A, mode, &sc, &eq, &et)) { status = OPAL_ERR_OVERFLOW; continue; }
        opalSearchResultSetScore(r, sc);
        if (searchType == OPAL_SEARCH_SCORE) { r->endLocationQuery = -1; r->endLocationTarget = -1; } /* :423-426, 869-871 */
        else { r->endLocationQuery = eq; r->endLocationTarget = et; }
    }
    if (status) return status; /* :1473 */

    if (searchType == OPAL_SEARCH_ALIGNMENT) { /* :1475-1507 */
        unsigned char* rq = reversed_copy(query, Q);
        for (int i = 0; i < N; i++) {
            OpalSearchResult* r = results[i];
            if ((mode == OPAL_MODE_SW && r->score == 0) || r->endLocationQuery < 0 || r->endLocationTarget < 0) {
                r->alignment = NULL; r->alignmentLength = 0;
                r->startLocationQuery = r->startLocationTarget = -1;
                r->endLocationQuery = r->endLocationTarget = -1;
                continue;
            }
            int aq = r->endLocationQuery + 1, at = r->endLocationTarget + 1;
            unsigned char* rt = reversed_copy(db[i], at);
            int sc, eq, et, n; unsigned char* ops;
            if (oracle_find_alignment(rq + Q - aq, aq, rt, at, Go, Ge, S, A, r->score, mode, &sc, &eq, &et, &ops, &n) == 0) {
                r->startLocationQuery = aq - eq - 1;   /* :1499-1500 */
                r->startLocationTarget = at - et - 1;
                for (int k = 0; k < n / 2; k++) { unsigned char x = ops[k]; ops[k] = ops[n - 1 - k]; ops[n - 1 - k] = x; } /* :1503 */
                r->alignment = (unsigned char*)realloc(ops, (size_t)(n > 0 ? n : 1));
                r->alignmentLength = n;
            } else {
                r->alignment = NULL; r->alignmentLength = 0;
                r->startLocationQuery = r->startLocationTarget = -1;
                status = -1; /* inconsistent prefilled score/end: "behavior is undefined" in the reference */
            }
            free(rt);
        }
        free(rq);
    } else { /* :1508-1515 */
        for (int i = 0; i < N; i++) {
            results[i]->alignment = NULL;
            results[i]->alignmentLength = -1;
            results[i]->startLocationQuery = -1;
            results[i]->startLocationTarget = -1;
        }
    }
    return status < 0 ? 0 : status;
}

int opalSearchDatabaseRescore(unsigned char query[], int Q, unsigned char* db[], int N, int lens[],
                              int Go, int Ge, int* S, int A, OpalSearchResult* results[],
                              const int searchType, int mode, int overflowMethod) {
    return opalSearchDatabase(query, Q, db, N, lens, Go, Ge, S, A, results, searchType, mode, overflowMethod);
}

int opalSearchDatabaseCharSW(unsigned char query[], int Q, unsigned char** db, int N, int lens[],
                             int Go, int Ge, int* S, int A, OpalSearchResult* results[]) { /* :1522-1546 */
    int argsFit = !(Go < SCHAR_MIN || SCHAR_MAX < Go || Ge < SCHAR_MIN || SCHAR_MAX < Ge); /* :178-180 */
    for (int i = 0; argsFit && i < A * A; i++) if (S[i] < SCHAR_MIN || SCHAR_MAX < S[i]) argsFit = 0; /* :188-193 */
    int rc = argsFit ? 0 : 1;
    for (int i = 0; i < N; i++) {
        int sc = 0, eq, et, ok = 0;
        if (argsFit && oracle_score_end(query, Q, db[i], lens[i], Go, Ge, S, A, OPAL_MODE_SW, &sc, &eq, &et) == 0)
            ok = sc <= SCHAR_MAX; /* H reaching 128 trips the 8-bit overflow test, :300-301, 353-362 */
        if (ok) {
            opalSearchResultSetScore(results[i], sc);
            results[i]->endLocationQuery = results[i]->endLocationTarget = -1; /* :423-426 */
        } else {
            results[i]->score = -1; results[i]->scoreSet = 0; /* :1538-1541 */
            rc = OPAL_ERR_OVERFLOW;
        }
    }
    return rc;
}

/* -------------------------------------------- alignment replay (checker) */

/*
 * Replays an operation string the way src/test.cpp:348-422 does.  Returns 0
 * when it is a valid alignment for `res` (labels agree with the residues, the
 * path ends at the recorded end cell, the affine score equals res->score);
 * otherwise a small positive code naming the first failed check.
 */
int oracle_check_alignment(const unsigned char* q, int Q, const unsigned char* t, int T,
                           const OpalSearchResult* res, int Go, int Ge, const int* S, int A) {
    i64 sc = 0;
    int qi = res->startLocationQuery, ti = res->startLocationTarget, prevOp = -1;
    if (qi < 0 || ti < 0) return 6;
    for (int i = 0; i < res->alignmentLength; i++) {
        int op = res->alignment[i];
        if ((op != OPAL_ALIGN_DEL && ti >= T) || (op != OPAL_ALIGN_INS && qi >= Q)) return 1;
        switch (op) {
        case OPAL_ALIGN_MATCH:    if (q[qi] != t[ti]) return 2; sc += S[q[qi] * A + t[ti]]; qi++; ti++; break;
        case OPAL_ALIGN_MISMATCH: if (q[qi] == t[ti]) return 3; sc += S[q[qi] * A + t[ti]]; qi++; ti++; break;
        case OPAL_ALIGN_DEL: sc -= (prevOp == OPAL_ALIGN_DEL ? Ge : Go); qi++; break;
        case OPAL_ALIGN_INS: sc -= (prevOp == OPAL_ALIGN_INS ? Ge : Go); ti++; break;
        default: return 7;
        }
        prevOp = op;
    }
    if (qi - 1 != res->endLocationQuery || ti - 1 != res->endLocationTarget) return 4;
    if (sc != res->score) return 5;
    return 0;
}
