// makedb_main.cpp -- opal_makedb_b200: FASTA -> packed database (packed_db.h).
//
// Does once what the reference CLI does on every run before it can search (readFastaSequences,
// reference src/opal_aligner.cpp:247-301) plus the length sort the GPU engine wants.
#include <getopt.h>
#include <stdio.h>

#include <string>

#include "fasta.h"
#include "packed_db.h"
#include "scoring.h"

using namespace opalcli;

int main(int argc, char* const argv[]) {
    std::string matrixName = "Blosum50", matrixPath;
    int option;
    while ((option = getopt(argc, argv, "m:f:")) >= 0) {
        if (option == 'm') matrixName = optarg;
        else if (option == 'f') matrixPath = optarg;
        else { optind = argc + 1; break; }
    }
    if (optind + 2 != argc) {
        fprintf(stderr,
                "\nUsage: opal_makedb_b200 [-m Blosum50|Blosum62 | -f matrix.mat] <db.fasta> <out.opdb>\n"
                "  Only the alphabet of the score matrix is used: it fixes the residue codes stored in <out.opdb>.\n");
        return 1;
    }
    Scoring scoring;
    std::string error;
    if (!Scoring::builtin(matrixName, &scoring)) { fprintf(stderr, "Given score matrix name is not valid\n"); return 1; }
    if (!matrixPath.empty() && !Scoring::load(matrixPath.c_str(), &scoring, &error)) { fprintf(stderr, "Error: %s\n", error.c_str()); return 1; }
    int16_t codes[256];
    scoring.letter_codes(codes);
    FILE* in = fopen(argv[optind], "r");
    if (!in) { fprintf(stderr, "Error: There is no file with name %s\n", argv[optind]); return 1; }
    SequenceBatch all;
    FastaReader reader(in, codes);
    const int state = reader.next(&all, &error, (1LL << 62));
    fclose(in);
    if (state < 0) { fprintf(stderr, "Error: %s: %s\n", argv[optind], error.c_str()); return 1; }
    PackedDb packed;
    pack_sequences(all, scoring.alphabet, &packed);
    if (!write_packed(argv[optind + 1], packed, &error)) { fprintf(stderr, "Error: %s\n", error.c_str()); return 1; }
    printf("Packed %d sequences, %lld residues, longest %d, alphabet of %d letters -> %s\n", packed.count(), packed.total(),
           packed.count() ? packed.lengths[0] : 0, scoring.size(), argv[optind + 1]);
    return 0;
}
