// aligner_main.cpp -- opal_aligner_b200: the reference's command line tool on the B200 library.
//
// Same options, same flow and the same output lines as the reference CLI (reference src/opal_aligner.cpp:20-245,
// usage :44-60, result lines :169-198, "Cpu time of searching" / "GCUPS" :203-207), so that scripts written
// against it (reference test/perf, test/compare_aligners grep the "Cpu time of searching:" line) run unchanged.
// Differences: the time reported is the wall-clock time of the search calls (the reference reports clock(),
// which for its single-threaded search is the same thing); <db> may also be a database packed by
// opal_makedb_b200, which is uploaded once and searched through the resident handle; -q searches every
// record of the query file, not only the first (score levels 0/1 go through the multi-query batch call).
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/opal.h"
#include "../../include/opal_b200.h"
#include "fasta.h"
#include "packed_db.h"
#include "scoring.h"

using namespace opalcli;

namespace {

double seconds_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void usage() {
    fprintf(stderr,
            "\nUsage: opal_aligner_b200 [options...] <query.fasta> <db.fasta | db.opdb>\n"
            "Options:\n"
            "  -o N  N is gap opening penalty (-g N is accepted too). [default: 3]\n"
            "  -e N  N is gap extension penalty. [default: 1]\n"
            "    Gap of length n will have penalty of g + (n - 1) * e.\n"
            "  -m Blosum50|Blosum62  Built-in score matrix to be used. [default: Blosum50]\n"
            "  -f FILE  FILE contains score matrix and some additional data. Overrides -m.\n"
            "  -s  If set, there will be no score output (silent mode).\n"
            "  -a SW|NW|HW|OV  Alignment mode that will be used. [default: SW]\n"
            "  -x search_level  Following search levels are available [default: %d]:\n"
            "    %d - score\n"
            "    %d - score, end location\n"
            "    %d - score, end and start location and alignment\n"
            "  -q  Search with every sequence of <query.fasta>, not only the first one.\n"
            "  -D N  CUDA device to use. [default: 0]\n"
            "  <db> may be a FASTA file or a database packed by opal_makedb_b200.\n",
            OPAL_SEARCH_SCORE, OPAL_SEARCH_SCORE, OPAL_SEARCH_SCORE_END, OPAL_SEARCH_ALIGNMENT);
}

// Two text rows per 50 operations, target above query, '_' for the gap side, then the index range covered
// (layout of the reference's printAlignment, src/opal_aligner.cpp:304-339).
void print_alignment(const unsigned char* query, const unsigned char* target, const OpalSearchResult& r,
                     const unsigned char* alphabet) {
    int t = r.startLocationTarget, q = r.startLocationQuery;
    for (int from = 0; from < r.alignmentLength; from += 50) {
        const int to = std::min(from + 50, r.alignmentLength);
        const int t0 = t, q0 = q;
        printf("T: ");
        for (int j = from; j < to; j++) {
            if (r.alignment[j] == OPAL_ALIGN_DEL) putchar('_');
            else putchar(alphabet[target[t++]]);
        }
        printf(" (%d - %d)\n", std::max(t0, 0), t - 1);
        printf("Q: ");
        for (int j = from; j < to; j++) {
            if (r.alignment[j] == OPAL_ALIGN_INS) putchar('_');
            else putchar(alphabet[query[q++]]);
        }
        printf(" (%d - %d)\n\n", std::max(q0, 0), q - 1);
    }
}

void print_results(const unsigned char* query, int firstIndex, int n, OpalSearchResult* const* results,
                   const std::vector<const unsigned char*>& targets, const unsigned char* alphabet) {
    printf("\n#<i>: <score> (<query start>, <target start>) (<query end>, <target end>)\n");
    for (int i = 0; i < n; i++) {
        const OpalSearchResult& r = *results[i];
        printf("#%d: %d", firstIndex + i, r.score);
        if (r.startLocationQuery >= 0) printf(" (%d, %d)", r.startLocationQuery, r.startLocationTarget);
        else printf(" (?, ?)");
        if (r.endLocationQuery >= 0) printf(" (%d, %d)", r.endLocationQuery, r.endLocationTarget);
        else printf(" (?, ?)");
        printf("\n");
        if (r.alignment) print_alignment(query, targets[i], r, alphabet);
    }
}

struct ResultSet {
    std::vector<OpalSearchResult> records;
    std::vector<OpalSearchResult*> pointers;
    explicit ResultSet(int n) : records((size_t)n), pointers((size_t)n) {
        for (int i = 0; i < n; i++) { opalInitSearchResult(&records[i]); pointers[i] = &records[i]; }
    }
    ~ResultSet() {
        for (auto& r : records) free(r.alignment);
    }
};

}  // namespace

int main(int argc, char* const argv[]) {
    int gapOpen = 3, gapExt = 1, searchType = OPAL_SEARCH_SCORE, device = 0;
    std::string matrixName = "Blosum50", matrixPath, modeName = "SW";
    bool silent = false, allQueries = false;
    int option;
    while ((option = getopt(argc, argv, "a:o:g:e:m:f:x:sqD:")) >= 0) {
        switch (option) {
            case 'a': modeName = optarg; break;
            case 'o': case 'g': gapOpen = atoi(optarg); break;
            case 'e': gapExt = atoi(optarg); break;
            case 'm': matrixName = optarg; break;
            case 'f': matrixPath = optarg; break;
            case 's': silent = true; break;
            case 'x': searchType = atoi(optarg); break;
            case 'q': allQueries = true; break;
            case 'D': device = atoi(optarg); break;
            default: usage(); return 1;
        }
    }
    if (optind + 2 != argc) { usage(); return 1; }

    Scoring scoring;
    if (!Scoring::builtin(matrixName, &scoring)) {
        fprintf(stderr, "Given score matrix name is not valid\n");
        return 1;
    }
    if (!matrixPath.empty()) {
        std::string error;
        if (!Scoring::load(matrixPath.c_str(), &scoring, &error)) { fprintf(stderr, "Error: %s\n", error.c_str()); return 1; }
    }
    const unsigned char* alphabet = scoring.alphabet.data();
    const int alphabetLength = scoring.size();
    int16_t codes[256];
    scoring.letter_codes(codes);

    int modeCode;
    if (modeName == "SW") modeCode = OPAL_MODE_SW;
    else if (modeName == "HW") modeCode = OPAL_MODE_HW;
    else if (modeName == "NW") modeCode = OPAL_MODE_NW;
    else if (modeName == "OV") modeCode = OPAL_MODE_OV;
    else { printf("Invalid mode!\n"); return 1; }
    printf("Using %s alignment mode.\n", modeName.c_str());

    // ---- query
    const char* queryPath = argv[optind];
    FILE* queryFile = fopen(queryPath, "r");
    if (!queryFile) { printf("Error: There is no file with name %s\n", queryPath); return 1; }
    printf("Reading query fasta file...\n");
    SequenceBatch queries;
    std::string error;
    {
        FastaReader reader(queryFile, codes);
        if (reader.next(&queries, &error) < 0) { printf("Error: %s: %s\n", queryPath, error.c_str()); return 1; }
    }
    fclose(queryFile);
    if (queries.count() == 0) { printf("Error: %s holds no sequence\n", queryPath); return 1; }
    const int numQueries = allQueries ? queries.count() : 1;
    if (allQueries) printf("Read %d query sequences, %lld residues.\n", queries.count(), queries.total());
    else printf("Read query sequence, %d residues.\n", queries.length(0));

    // ---- database
    const char* dbPath = argv[optind + 1];
    FILE* dbFile = fopen(dbPath, "r");
    if (!dbFile) { printf("Error: There is no file with name %s\n", dbPath); return 1; }
    const bool packedInput = is_packed_file(dbPath);
    char deviceEnv[32];
    snprintf(deviceEnv, sizeof(deviceEnv), "%d", device);
    setenv("OPAL_B200_DEVICE", deviceEnv, 1);  // the drop-in entry points take their device from the environment

    double searchSeconds = 0;
    long long dbTotalResidues = 0, queryResidues = 0;
    int dbTotalSequences = 0, exitCode = 0;
    for (int k = 0; k < numQueries; k++) queryResidues += queries.length(k);

    // One chunk of the database: search with every selected query and print.
    auto process = [&](OpalB200Db* handle, unsigned char** db, int* lengths, int n, const std::vector<const unsigned char*>& targets) {
        const int firstIndex = dbTotalSequences - n;
        // score levels of several queries against a resident database: one batched call
        if (handle && numQueries > 1 && searchType != OPAL_SEARCH_ALIGNMENT) {
            std::vector<const unsigned char*> qptr((size_t)numQueries);
            std::vector<int> qlen((size_t)numQueries);
            for (int k = 0; k < numQueries; k++) { qptr[k] = queries.sequence(k); qlen[k] = queries.length(k); }
            std::vector<int> sc((size_t)numQueries * n), eq((size_t)numQueries * n, -1), et((size_t)numQueries * n, -1);
            printf("\nComparing %d queries to database...", numQueries);
            fflush(stdout);
            const auto t0 = std::chrono::steady_clock::now();
            const int rc = opalb200_db_search_batch(handle, numQueries, qptr.data(), qlen.data(), gapOpen, gapExt, scoring.matrix.data(),
                                                    alphabetLength, searchType, modeCode, sc.data(), eq.data(), et.data(), 0, nullptr);
            searchSeconds += seconds_since(t0);
            if (rc) { printf("\nDatabase search failed with error code: %d\n", rc); exitCode = 1; }
            printf("\nFinished!\n");
            if (!silent && !rc)
                for (int k = 0; k < numQueries; k++) {
                    ResultSet rs(n);
                    for (int i = 0; i < n; i++) {
                        opalSearchResultSetScore(rs.pointers[i], sc[(size_t)k * n + i]);
                        rs.records[i].endLocationQuery = eq[(size_t)k * n + i];
                        rs.records[i].endLocationTarget = et[(size_t)k * n + i];
                    }
                    printf("\nQuery #%d, %d residues:", k, queries.length(k));
                    print_results(queries.sequence(k), firstIndex, n, rs.pointers.data(), targets, alphabet);
                }
            return;
        }
        for (int k = 0; k < numQueries; k++) {
            ResultSet rs(n);
            printf(numQueries > 1 ? "\nComparing query #%d to database..." : "\nComparing query to database...", k);
            fflush(stdout);
            const auto t0 = std::chrono::steady_clock::now();
            const int rc = handle ? opalb200_db_search_results(handle, queries.sequence(k), queries.length(k), gapOpen, gapExt,
                                                               scoring.matrix.data(), alphabetLength, rs.pointers.data(), searchType, modeCode)
                                  : opalSearchDatabase(queries.sequence(k), queries.length(k), db, n, lengths, gapOpen, gapExt,
                                                       scoring.matrix.data(), alphabetLength, rs.pointers.data(), searchType, modeCode,
                                                       OPAL_OVERFLOW_BUCKETS);
            searchSeconds += seconds_since(t0);
            if (rc) {
                printf("\nDatabase search failed with error code: %d\n", rc);
                if (*opalb200_last_error()) fprintf(stderr, "%s\n", opalb200_last_error());
                exitCode = 1;
            }
            printf("\nFinished!\n");
            if (!silent) print_results(queries.sequence(k), firstIndex, n, rs.pointers.data(), targets, alphabet);
        }
    };

    if (packedInput) {
        fclose(dbFile);
        printf("\nReading packed database file...\n");
        PackedDb packed;
        if (!read_packed(dbPath, &packed, &error)) { printf("Error: %s\n", error.c_str()); return 1; }
        if (packed.alphabet != scoring.alphabet) { printf("Error: %s was packed with a different alphabet than the score matrix\n", dbPath); return 1; }
        const int n = packed.count();
        dbTotalSequences = n;
        dbTotalResidues = packed.total();
        printf("Read %d database sequences, %lld residues total.\n", n, dbTotalResidues);
        printf("Whole database read: %d database sequences, %lld residues in total.\n", n, dbTotalResidues);
        const auto t0 = std::chrono::steady_clock::now();
        OpalB200Db* handle = opalb200_db_create_sorted(packed.residues.data(), packed.lengths.data(), packed.order.data(), n, device);
        if (!handle) { printf("Error: %s\n", opalb200_last_error()); return 1; }
        printf("Database uploaded to device %d in %.6lf s.\n", device, seconds_since(t0));
        std::vector<const unsigned char*> targets((size_t)n);
        for (int p = 0; p < n; p++) targets[packed.order[p]] = packed.residues.data() + packed.offsets[p];
        process(handle, nullptr, nullptr, n, targets);
        opalb200_db_destroy(handle);
    } else {
        FastaReader reader(dbFile, codes);
        bool wholeDbRead = false;
        while (!wholeDbRead) {
            SequenceBatch chunk;
            printf("\nReading database fasta file...\n");
            // Read and process the database chunk by chunk (one chunk unless the database is huge).
            const int state = reader.next(&chunk, &error);
            if (state < 0) { printf("Error: %s: %s\n", dbPath, error.c_str()); return 1; }
            wholeDbRead = state == 1;
            const int n = chunk.count();
            std::vector<unsigned char*> db((size_t)n);
            std::vector<int> lengths((size_t)n);
            std::vector<const unsigned char*> targets((size_t)n);
            for (int i = 0; i < n; i++) { db[i] = chunk.sequence(i); lengths[i] = chunk.length(i); targets[i] = db[i]; }
            printf("Read %d database sequences, %lld residues total.\n", n, chunk.total());
            dbTotalResidues += chunk.total();
            dbTotalSequences += n;
            if (wholeDbRead)
                printf("Whole database read: %d database sequences, %lld residues in total.\n", dbTotalSequences, dbTotalResidues);
            // several queries: keep the chunk resident instead of packing it once per query
            OpalB200Db* handle = nullptr;
            if (numQueries > 1) {
                handle = opalb200_db_create(db.data(), n, lengths.data(), device);
                if (!handle) { printf("Error: %s\n", opalb200_last_error()); return 1; }
            }
            process(handle, db.data(), lengths.data(), n, targets);
            if (handle) opalb200_db_destroy(handle);
        }
        fclose(dbFile);
    }

    printf("\nCpu time of searching: %.2lf\n", searchSeconds);
    printf("Wall time of searching (s): %.6lf\n", searchSeconds);
    if (searchType != OPAL_SEARCH_ALIGNMENT)
        printf("GCUPS (giga cell updates per second): %.2lf\n", dbTotalResidues / 1000000000.0 * queryResidues / searchSeconds);
    return exitCode;
}
