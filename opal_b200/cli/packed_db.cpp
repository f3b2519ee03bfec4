// packed_db.cpp -- see packed_db.h.
#include "packed_db.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>

namespace opalcli {

static const char kMagic[8] = {'O', 'P', 'A', 'L', 'B', '2', 'D', 'B'};
static const uint32_t kVersion = 1;

void pack_sequences(const SequenceBatch& batch, const std::vector<unsigned char>& alphabet, PackedDb* out) {
    const int n = batch.count();
    out->alphabet = alphabet;
    out->order.resize(n);
    std::iota(out->order.begin(), out->order.end(), 0);
    std::stable_sort(out->order.begin(), out->order.end(), [&](int a, int b) { return batch.length(a) > batch.length(b); });
    out->lengths.resize(n);
    out->offsets.assign((size_t)n + 1, 0);
    out->residues.resize((size_t)batch.total());
    long long at = 0;
    for (int p = 0; p < n; p++) {
        const int i = out->order[p], len = batch.length(i);
        out->lengths[p] = len;
        out->offsets[p] = at;
        if (len > 0) memcpy(out->residues.data() + at, batch.sequence(i), (size_t)len);
        at += len;
    }
    out->offsets[n] = at;
}

namespace {
struct Header {
    char magic[8];
    uint32_t version, alphabetLength;
    uint64_t numSequences, numResidues;
    unsigned char alphabet[256];
};
static_assert(sizeof(Header) == 288, "packed header layout");
}  // namespace

bool write_packed(const char* path, const PackedDb& db, std::string* error) {
    FILE* f = fopen(path, "wb");
    if (!f) { *error = std::string("cannot create ") + path; return false; }
    Header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, kMagic, 8);
    h.version = kVersion;
    h.alphabetLength = (uint32_t)db.alphabet.size();
    h.numSequences = (uint64_t)db.count();
    h.numResidues = (uint64_t)db.total();
    memcpy(h.alphabet, db.alphabet.data(), std::min<size_t>(db.alphabet.size(), 256));
    const size_t n = (size_t)db.count();
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && (n == 0 || fwrite(db.lengths.data(), sizeof(int), n, f) == n);
    ok = ok && (n == 0 || fwrite(db.order.data(), sizeof(int), n, f) == n);
    ok = ok && (db.residues.empty() || fwrite(db.residues.data(), 1, db.residues.size(), f) == db.residues.size());
    ok = (fclose(f) == 0) && ok;
    if (!ok) *error = std::string("short write to ") + path;
    return ok;
}

bool read_packed(const char* path, PackedDb* out, std::string* error) {
    FILE* f = fopen(path, "rb");
    if (!f) { *error = std::string("cannot open ") + path; return false; }
    Header h;
    auto fail = [&](const std::string& why) { fclose(f); *error = std::string(path) + ": " + why; return false; };
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, kMagic, 8) != 0) return fail("not a packed opal-b200 database");
    if (h.version != kVersion) return fail("unsupported packed database version " + std::to_string(h.version));
    if (h.alphabetLength == 0 || h.alphabetLength > 254 || h.numSequences > 0x7fffffffULL) return fail("corrupt header");
    const size_t n = (size_t)h.numSequences;
    out->alphabet.assign(h.alphabet, h.alphabet + h.alphabetLength);
    out->lengths.resize(n);
    out->order.resize(n);
    out->residues.resize((size_t)h.numResidues);
    if (n && (fread(out->lengths.data(), sizeof(int), n, f) != n || fread(out->order.data(), sizeof(int), n, f) != n)) return fail("truncated index");
    if (h.numResidues && fread(out->residues.data(), 1, out->residues.size(), f) != out->residues.size()) return fail("truncated residues");
    fclose(f);
    out->offsets.assign(n + 1, 0);
    std::vector<char> seen(n, 0);
    for (size_t p = 0; p < n; p++) {
        const int len = out->lengths[p], idx = out->order[p];
        if (len < 0 || (p > 0 && len > out->lengths[p - 1]) || idx < 0 || (size_t)idx >= n || seen[idx]) { *error = std::string(path) + ": corrupt index"; return false; }
        seen[idx] = 1;
        out->offsets[p + 1] = out->offsets[p] + len;
    }
    if ((uint64_t)out->offsets[n] != h.numResidues) { *error = std::string(path) + ": lengths do not add up to the residue count"; return false; }
    for (unsigned char c : out->residues)
        if (c >= h.alphabetLength) { *error = std::string(path) + ": residue code outside the alphabet"; return false; }
    return true;
}

bool is_packed_file(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char magic[8];
    const bool yes = fread(magic, 1, 8, f) == 8 && memcmp(magic, kMagic, 8) == 0;
    fclose(f);
    return yes;
}

}  // namespace opalcli
