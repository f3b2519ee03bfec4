/* Plain C caller of the drop-in API: the README example of the reference (README.md:33-66), printed as
 * "score startQ startT endQ endT ops" per target. Compiled and run by tests/test_gpu_c_linkage.py. */
#include <stdio.h>
#include <stdlib.h>

#include "opal.h"

int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : OPAL_MODE_SW;
    int scoreMatrix[16] = {2, -1, -3, 0, -1, 4, -5, -1, -3, -5, 1, -10, 0, -1, -10, 4};
    unsigned char query[10] = {0, 1, 3, 2, 1, 0, 3, 0, 1, 1};
    unsigned char s1[14] = {1, 3, 2, 3, 0, 0, 1, 0, 2, 2, 1, 2, 3, 2}, s2[12] = {2, 1, 1, 3, 2, 0, 0, 2, 2, 0, 2, 1};
    unsigned char s3[13] = {0, 0, 2, 1, 0, 3, 1, 1, 2, 3, 2, 1, 0}, s4[9] = {2, 3, 3, 3, 1, 1, 2, 2, 0};
    unsigned char* db[4] = {s1, s2, s3, s4};
    int lens[4] = {14, 12, 13, 9};
    OpalSearchResult recs[4];
    OpalSearchResult* results[4];
    for (int i = 0; i < 4; i++) { results[i] = &recs[i]; opalInitSearchResult(results[i]); }
    int rc = opalSearchDatabase(query, 10, db, 4, lens, 3, 1, scoreMatrix, 4, results, OPAL_SEARCH_ALIGNMENT, mode,
                                OPAL_OVERFLOW_BUCKETS);
    if (rc) { printf("rc=%d\n", rc); return 1; }
    for (int i = 0; i < 4; i++) {
        printf("%d %d %d %d %d ", recs[i].score, recs[i].startLocationQuery, recs[i].startLocationTarget,
               recs[i].endLocationQuery, recs[i].endLocationTarget);
        for (int k = 0; k < recs[i].alignmentLength; k++) printf("%d", recs[i].alignment[k]);
        printf("\n");
        if (!opalSearchResultIsEmpty(recs[i])) free(recs[i].alignment);
    }
    return 0;
}
