// packed_db.h -- the on-disk form of a database that is ready for the GPU engine.
//
// What the reference CLI redoes on every run -- parse FASTA, map letters to codes (src/opal_aligner.cpp:247-301) --
// and what the engine would redo on every upload -- sort by length, concatenate -- is done once by
// opal_makedb_b200 and stored.  A packed file loads with two reads and goes to opalb200_db_create_sorted as is.
//
//   offset  size          field
//   0       8             magic "OPALB2DB"
//   8       4             version (1), little endian like every integer below
//   12      4             alphabetLength A
//   16      8             numSequences n
//   24      8             numResidues
//   32      256           alphabet letters (first A bytes used, rest 0)
//   288     4 n           lengths, longest first (ties: original order)
//   ..      4 n           order: original index of the p-th sorted sequence
//   ..      numResidues   residues (alphabet codes), sorted sequences back to back
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "fasta.h"

namespace opalcli {

struct PackedDb {
    std::vector<unsigned char> alphabet;
    std::vector<int> lengths;   // sorted, longest first
    std::vector<int> order;     // sorted position -> original index
    std::vector<unsigned char> residues;
    std::vector<long long> offsets;  // derived: n + 1 entries into residues (sorted positions)

    int count() const { return (int)lengths.size(); }
    long long total() const { return (long long)residues.size(); }
};

// Sorts a parsed batch longest first (stable) and lays it out as above.
void pack_sequences(const SequenceBatch& batch, const std::vector<unsigned char>& alphabet, PackedDb* out);
bool write_packed(const char* path, const PackedDb& db, std::string* error);
bool read_packed(const char* path, PackedDb* out, std::string* error);
// True if the file starts with the packed-database magic.
bool is_packed_file(const char* path);

}  // namespace opalcli
