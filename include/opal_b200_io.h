/*
 * opal_b200_io.h -- C ABI of libopal_b200_io.so: the host-side data formats either side of the search path.
 *
 * The reference keeps these in its command line tool: score matrix files and the built-in BLOSUM50
 * (reference src/ScoreMatrix.cpp:17-35, 57-84) and FASTA -> alphabet codes (readFastaSequences,
 * reference src/opal_aligner.cpp:247-301).  Here they are a small library of their own (no CUDA in it), used
 * by opal_aligner_b200 / opal_makedb_b200 and bindable from anywhere, plus the packed on-disk database that
 * feeds opalb200_db_create_sorted (opal_b200.h) without any per-run parsing or sorting.
 *
 * Functions returning int return 0 on success and non-zero on failure; opalio_last_error() has the text.
 */
#ifndef OPAL_B200_IO_H
#define OPAL_B200_IO_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OpalioSequences OpalioSequences;
typedef struct OpalioPacked OpalioPacked;

const char* opalio_last_error(void);

/*
 * Score matrix by built-in name ("Blosum50", "Blosum62"; path == NULL) or from a .mat file (path != NULL:
 * first line the letters, then one row of integers per line).  alphabet must hold 256 bytes, matrix
 * matrixCapacity ints (alphabetLength^2 are written; 254*254 always suffices).
 */
int opalio_load_matrix(const char* name, const char* path, unsigned char* alphabet, int* alphabetLength,
                       int* matrix, int matrixCapacity);

/*
 * Parses a FASTA file into alphabet codes.  Letters outside the alphabet map to '*' when the alphabet has
 * it and are an error otherwise.  maxResidues <= 0 reads the whole file; otherwise reading stops before
 * the first record that starts after more than maxResidues residues (the reference's chunking,
 * src/opal_aligner.cpp:282-285).  *wholeFile (nullable) is set to 1 when the end of the file was reached, else 0.
 */
OpalioSequences* opalio_read_fasta(const char* path, const unsigned char* alphabet, int alphabetLength,
                                   long long maxResidues, int* wholeFile);
int opalio_sequences_count(const OpalioSequences* s);
long long opalio_sequences_residues(const OpalioSequences* s);
/* n + 1 offsets into opalio_sequences_data(): record i = data[offsets[i] .. offsets[i+1]) */
const long long* opalio_sequences_offsets(const OpalioSequences* s);
const unsigned char* opalio_sequences_data(const OpalioSequences* s);
void opalio_sequences_free(OpalioSequences* s);

/* FASTA -> packed database file (sorted longest first; format in opal_b200/cli/packed_db.h). */
int opalio_pack_fasta(const char* fastaPath, const unsigned char* alphabet, int alphabetLength, const char* outPath);

/* Packed database file -> memory, validated.  The three arrays are exactly the arguments of
 * opalb200_db_create_sorted(residues, sortedLengths, order, count, device). */
OpalioPacked* opalio_packed_open(const char* path);
int opalio_packed_count(const OpalioPacked* p);
long long opalio_packed_residues(const OpalioPacked* p);
int opalio_packed_alphabet(const OpalioPacked* p, unsigned char* alphabet /* 256 bytes */);
const int* opalio_packed_lengths(const OpalioPacked* p);
const int* opalio_packed_order(const OpalioPacked* p);
const unsigned char* opalio_packed_data(const OpalioPacked* p);
void opalio_packed_free(OpalioPacked* p);

#ifdef __cplusplus
}
#endif
#endif /* OPAL_B200_IO_H */
