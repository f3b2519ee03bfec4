// scoring.h -- alphabet + substitution matrix as the opal C API takes them (int* scoreMatrix, alphabetLength).
//
// Plays the role of the reference CLI's ScoreMatrix class (reference src/ScoreMatrix.hpp:8-31,
// src/ScoreMatrix.cpp:17-35): same ".mat" text format (first line: the letters naming the columns,
// then one row of integers per line), same built-in "Blosum50" (src/ScoreMatrix.cpp:57-84).
// Unlike the reference, a malformed file is an error instead of a short matrix.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace opalcli {

struct Scoring {
    std::vector<unsigned char> alphabet;  // letter of code i
    std::vector<int> matrix;              // alphabet.size()^2 entries, row-major: matrix[a * A + b]

    int size() const { return (int)alphabet.size(); }

    // Built-in matrices by name (case-insensitive): "Blosum50", the reference's only one, plus "Blosum62".
    static bool builtin(const std::string& name, Scoring* out);
    // The reference's .mat format.  Returns false and fills *error on unreadable / ragged / non-square input.
    static bool load(const char* path, Scoring* out, std::string* error);

    // Letter -> code table used when parsing FASTA (reference src/opal_aligner.cpp:250-258): a letter of the
    // alphabet maps to its index; every other byte maps to the index of '*' if the alphabet has one and to
    // -1 (parse error) otherwise -- the reference leaves those entries uninitialised.
    void letter_codes(int16_t codes[256]) const;
};

}  // namespace opalcli
