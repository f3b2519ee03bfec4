// fasta.h -- FASTA -> sequences of alphabet codes, read in bounded chunks.
//
// Plays the role of readFastaSequences (reference src/opal_aligner.cpp:247-301) with the same record
// rules: '>' starts a header that runs to the end of the line, a record exists once its first residue has
// been seen (headers without residues yield nothing), '\r' and '\n' are dropped, and reading stops before a
// new record once more than `maxResidues` residues are held (the reference's 1 GiB chunking, :282-285).
// Differences, all on malformed input: bytes outside the alphabet are an error when the alphabet has no
// '*' (the reference reads an uninitialised table there), spaces and tabs inside sequence lines are
// dropped, and residues are stored back to back with offsets instead of one heap vector per record.
#pragma once
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

namespace opalcli {

struct SequenceBatch {
    std::vector<unsigned char> residues;       // all records back to back (alphabet codes)
    std::vector<long long> offsets = {0};      // record i = residues[offsets[i] .. offsets[i+1])

    int count() const { return (int)offsets.size() - 1; }
    int length(int i) const { return (int)(offsets[i + 1] - offsets[i]); }
    const unsigned char* sequence(int i) const { return residues.data() + offsets[i]; }
    unsigned char* sequence(int i) { return residues.data() + offsets[i]; }
    long long total() const { return offsets.back(); }
    void clear() { residues.clear(); offsets.assign(1, 0); }
};

class FastaReader {
public:
    static constexpr long long kDefaultChunk = 1073741824LL;  // residues per chunk, as the reference

    FastaReader(FILE* file, const int16_t codes[256]);
    // Reads the next chunk into *out (cleared first).  Returns 1 when the end of the file was reached,
    // 0 when the chunk limit stopped the read (call again), -1 on a parse error (*error says where).
    int next(SequenceBatch* out, std::string* error, long long maxResidues = kDefaultChunk);

private:
    bool fill();
    FILE* file_;
    int16_t codes_[256];
    std::vector<unsigned char> buffer_;
    size_t pos_ = 0, end_ = 0;
    long long line_ = 1;
    bool inHeader_ = false, eof_ = false;
};

}  // namespace opalcli
