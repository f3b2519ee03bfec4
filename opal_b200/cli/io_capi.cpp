// io_capi.cpp -- the C ABI of include/opal_b200_io.h over scoring / fasta / packed_db.
#include "../../include/opal_b200_io.h"

#include <string.h>

#include <string>

#include "fasta.h"
#include "packed_db.h"
#include "scoring.h"

using namespace opalcli;

struct OpalioSequences { SequenceBatch batch; };
struct OpalioPacked { PackedDb db; };

static thread_local std::string g_error;

static void codes_of(const unsigned char* alphabet, int alphabetLength, int16_t codes[256]) {
    Scoring s;
    s.alphabet.assign(alphabet, alphabet + alphabetLength);
    s.letter_codes(codes);
}

extern "C" {

const char* opalio_last_error(void) { return g_error.c_str(); }

int opalio_load_matrix(const char* name, const char* path, unsigned char* alphabet, int* alphabetLength, int* matrix,
                       int matrixCapacity) {
    Scoring s;
    if (path) {
        if (!Scoring::load(path, &s, &g_error)) return 1;
    } else if (!name || !Scoring::builtin(name, &s)) {
        g_error = "Given score matrix name is not valid";
        return 1;
    }
    if ((int)s.matrix.size() > matrixCapacity) { g_error = "matrix buffer too small"; return 1; }
    memset(alphabet, 0, 256);
    memcpy(alphabet, s.alphabet.data(), s.alphabet.size());
    *alphabetLength = s.size();
    memcpy(matrix, s.matrix.data(), sizeof(int) * s.matrix.size());
    return 0;
}

OpalioSequences* opalio_read_fasta(const char* path, const unsigned char* alphabet, int alphabetLength, long long maxResidues,
                                   int* wholeFile) {
    FILE* f = fopen(path, "r");
    if (!f) { g_error = std::string("There is no file with name ") + path; return nullptr; }
    int16_t codes[256];
    codes_of(alphabet, alphabetLength, codes);
    FastaReader reader(f, codes);
    OpalioSequences* out = new OpalioSequences();
    const int state = reader.next(&out->batch, &g_error, maxResidues > 0 ? maxResidues : (1LL << 62));
    fclose(f);
    if (state < 0) { delete out; return nullptr; }
    if (wholeFile) *wholeFile = state;
    return out;
}

int opalio_sequences_count(const OpalioSequences* s) { return s->batch.count(); }
long long opalio_sequences_residues(const OpalioSequences* s) { return s->batch.total(); }
const long long* opalio_sequences_offsets(const OpalioSequences* s) { return s->batch.offsets.data(); }
const unsigned char* opalio_sequences_data(const OpalioSequences* s) { return s->batch.residues.data(); }
void opalio_sequences_free(OpalioSequences* s) { delete s; }

int opalio_pack_fasta(const char* fastaPath, const unsigned char* alphabet, int alphabetLength, const char* outPath) {
    OpalioSequences* s = opalio_read_fasta(fastaPath, alphabet, alphabetLength, 0, nullptr);
    if (!s) return 1;
    PackedDb packed;
    pack_sequences(s->batch, std::vector<unsigned char>(alphabet, alphabet + alphabetLength), &packed);
    delete s;
    return write_packed(outPath, packed, &g_error) ? 0 : 1;
}

OpalioPacked* opalio_packed_open(const char* path) {
    OpalioPacked* p = new OpalioPacked();
    if (!read_packed(path, &p->db, &g_error)) { delete p; return nullptr; }
    return p;
}
int opalio_packed_count(const OpalioPacked* p) { return p->db.count(); }
long long opalio_packed_residues(const OpalioPacked* p) { return p->db.total(); }
int opalio_packed_alphabet(const OpalioPacked* p, unsigned char* alphabet) {
    memset(alphabet, 0, 256);
    memcpy(alphabet, p->db.alphabet.data(), p->db.alphabet.size());
    return (int)p->db.alphabet.size();
}
const int* opalio_packed_lengths(const OpalioPacked* p) { return p->db.lengths.data(); }
const int* opalio_packed_order(const OpalioPacked* p) { return p->db.order.data(); }
const unsigned char* opalio_packed_data(const OpalioPacked* p) { return p->db.residues.data(); }
void opalio_packed_free(OpalioPacked* p) { delete p; }

}  // extern "C"
