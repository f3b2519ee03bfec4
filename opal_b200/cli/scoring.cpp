// scoring.cpp -- see scoring.h.
#include "scoring.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace opalcli {

// Standard NCBI BLOSUM tables as lower triangles (the matrices are symmetric), in the letter order the
// reference uses (src/ScoreMatrix.cpp:51-55; BLOSUM62 as src/score_matrices/blosum62.mat, which has no '*').
static const char* const kBlosum50Triangle[] = {
    "5",
    "-2 7",
    "-1 -1 7",
    "-2 -2 2 8",
    "-1 -4 -2 -4 13",
    "-1 1 0 0 -3 7",
    "-1 0 0 2 -3 2 6",
    "0 -3 0 -1 -3 -2 -3 8",
    "-2 0 1 -1 -3 1 0 -2 10",
    "-1 -4 -3 -4 -2 -3 -4 -4 -4 5",
    "-2 -3 -4 -4 -2 -2 -3 -4 -3 2 5",
    "-1 3 0 -1 -3 2 1 -2 0 -3 -3 6",
    "-1 -2 -2 -4 -2 0 -2 -3 -1 2 3 -2 7",
    "-3 -3 -4 -5 -2 -4 -3 -4 -1 0 1 -4 0 8",
    "-1 -3 -2 -1 -4 -1 -1 -2 -2 -3 -4 -1 -3 -4 10",
    "1 -1 1 0 -1 0 -1 0 -1 -3 -3 0 -2 -3 -1 5",
    "0 -1 0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1 2 5",
    "-3 -3 -4 -5 -5 -1 -3 -3 -3 -3 -2 -3 -1 1 -4 -4 -3 15",
    "-2 -1 -2 -3 -3 -1 -2 -3 2 -1 -1 -2 0 4 -3 -2 -2 2 8",
    "0 -3 -3 -4 -1 -3 -3 -4 -4 4 1 -3 1 -1 -3 -2 0 -3 -1 5",
    "-2 -1 5 6 -3 0 1 -1 0 -4 -4 0 -3 -4 -2 0 0 -5 -3 -3 6",
    "-1 0 0 1 -3 4 5 -2 0 -3 -3 1 -1 -4 -1 0 -1 -2 -2 -3 1 5",
    "-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1",
    "-5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 1",
};
static const char* const kBlosum62Triangle[] = {
    "4",
    "-1 5",
    "-2 0 6",
    "-2 -2 1 6",
    "0 -3 -3 -3 9",
    "-1 1 0 0 -3 5",
    "-1 0 0 2 -4 2 5",
    "0 -2 0 -1 -3 -2 -2 6",
    "-2 0 1 -1 -3 0 0 -2 8",
    "-1 -3 -3 -3 -1 -3 -3 -4 -3 4",
    "-1 -2 -3 -4 -1 -2 -3 -4 -3 2 4",
    "-1 2 0 -1 -3 1 1 -2 -1 -3 -2 5",
    "-1 -1 -2 -3 -1 0 -2 -3 -2 1 2 -1 5",
    "-2 -3 -3 -3 -2 -3 -3 -3 -1 0 0 -3 0 6",
    "-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4 7",
    "1 -1 1 0 -1 0 0 0 -1 -2 -2 0 -1 -2 -1 4",
    "0 -1 0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1 1 5",
    "-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1 1 -4 -3 -2 11",
    "-2 -2 -2 -3 -2 -1 -2 -3 2 -1 -1 -2 -1 3 -3 -2 -2 2 7",
    "0 -3 -3 -3 -1 -2 -2 -3 -3 3 1 -2 1 -1 -2 -2 0 -3 -1 4",
    "-2 -1 3 4 -3 0 1 -1 0 -3 -4 0 -3 -3 -2 0 -1 -4 -3 -3 4",
    "-1 0 0 1 -3 3 4 -2 0 -3 -3 1 -1 -3 -1 0 -1 -3 -2 -2 1 4",
    "0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2 0 0 -2 -1 -1 -1 -1 -1",
};

// Whitespace-separated tokens of one line.  (No iostreams here: this code is also loaded into processes that
// already hold another copy of the C++ runtime, and plain C parsing keeps the two from meeting.)
static std::vector<std::string> tokens_of(const std::string& line) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && isspace((unsigned char)line[i])) i++;
        size_t j = i;
        while (j < line.size() && !isspace((unsigned char)line[j])) j++;
        if (j > i) out.push_back(line.substr(i, j - i));
        i = j;
    }
    return out;
}

static bool parse_int(const std::string& tok, int* out) {
    char* end = nullptr;
    const long v = std::strtol(tok.c_str(), &end, 10);
    if (end == tok.c_str() || *end != 0) return false;
    *out = (int)v;
    return true;
}

static bool from_triangle(const char* letters, const char* const* rows, Scoring* out) {
    const int A = (int)strlen(letters);
    out->alphabet.assign(letters, letters + A);
    out->matrix.assign((size_t)A * A, 0);
    for (int r = 0; r < A; r++) {
        const std::vector<std::string> toks = tokens_of(rows[r]);
        if ((int)toks.size() != r + 1) return false;
        for (int c = 0; c <= r; c++) {
            int v;
            if (!parse_int(toks[c], &v)) return false;
            out->matrix[(size_t)r * A + c] = out->matrix[(size_t)c * A + r] = v;
        }
    }
    return true;
}

bool Scoring::builtin(const std::string& name, Scoring* out) {
    std::string n = name;
    std::transform(n.begin(), n.end(), n.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (n == "blosum50") return from_triangle("ARNDCQEGHILKMFPSTWYVBZX*", kBlosum50Triangle, out);
    if (n == "blosum62") return from_triangle("ARNDCQEGHILKMFPSTWYVBZX", kBlosum62Triangle, out);
    return false;
}

static bool read_line(FILE* f, std::string* line) {
    line->clear();
    int c;
    while ((c = fgetc(f)) != EOF) {
        if (c == '\n') return true;
        line->push_back((char)c);
    }
    return !line->empty();
}

bool Scoring::load(const char* path, Scoring* out, std::string* error) {
    FILE* file = fopen(path, "r");
    if (!file) { *error = std::string("cannot open score matrix file ") + path; return false; }
    auto fail = [&](const std::string& why) { fclose(file); *error = "score matrix file: " + why; return false; };
    out->alphabet.clear();
    out->matrix.clear();
    std::string line;
    while (read_line(file, &line)) {  // first non-empty line: the letters
        for (const std::string& tok : tokens_of(line)) out->alphabet.push_back((unsigned char)tok[0]);
        if (!out->alphabet.empty()) break;
    }
    const size_t A = out->alphabet.size();
    if (A == 0 || A > 254) return fail("the first line must name 1..254 letters");
    size_t rows = 0;
    while (read_line(file, &line)) {
        size_t inRow = 0;
        const std::vector<std::string> toks = tokens_of(line);
        for (size_t k = 0; k < toks.size(); k++) {
            int v;
            if (!parse_int(toks[k], &v)) {
                if (k == 0 && toks[k].size() == 1) continue;  // a row may start with its letter (NCBI style)
                return fail("'" + toks[k] + "' is not an integer");
            }
            out->matrix.push_back(v);
            inRow++;
        }
        if (inRow == 0) continue;
        if (inRow != A) return fail("row " + std::to_string(rows + 1) + " has " + std::to_string(inRow) + " entries, expected " + std::to_string(A));
        rows++;
    }
    if (rows != A) return fail(std::to_string(rows) + " rows for " + std::to_string(A) + " letters");
    fclose(file);
    return true;
}

void Scoring::letter_codes(int16_t codes[256]) const {
    int wildcard = -1;
    for (int i = 0; i < size(); i++)
        if (alphabet[i] == '*') { wildcard = i; break; }
    for (int c = 0; c < 256; c++) codes[c] = (int16_t)wildcard;
    for (int i = 0; i < size(); i++) codes[alphabet[i]] = (int16_t)i;
}

}  // namespace opalcli
