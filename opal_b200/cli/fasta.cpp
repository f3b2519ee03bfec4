// fasta.cpp -- see fasta.h.
#include "fasta.h"

#include <string.h>

namespace opalcli {

FastaReader::FastaReader(FILE* file, const int16_t codes[256]) : file_(file), buffer_(1 << 20) {
    memcpy(codes_, codes, sizeof(codes_));
}

bool FastaReader::fill() {
    if (eof_) return false;
    end_ = fread(buffer_.data(), 1, buffer_.size(), file_);
    pos_ = 0;
    if (end_ == 0) { eof_ = true; return false; }
    return true;
}

int FastaReader::next(SequenceBatch* out, std::string* error, long long maxResidues) {
    out->clear();
    bool inRecord = false;  // residues of the current record are being appended
    for (;;) {
        if (pos_ == end_ && !fill()) break;
        const unsigned char c = buffer_[pos_];
        if (inHeader_) {  // skip to the end of the header line in one go
            const void* nl = memchr(buffer_.data() + pos_, '\n', end_ - pos_);
            if (!nl) { pos_ = end_; continue; }
            pos_ = (size_t)((const unsigned char*)nl - buffer_.data()) + 1;
            line_++;
            inHeader_ = false;
            continue;
        }
        if (c == '>') {
            if (inRecord) { out->offsets.push_back((long long)out->residues.size()); inRecord = false; }
            inHeader_ = true;
            pos_++;
            continue;
        }
        if (c == '\n') line_++;
        if (c == '\n' || c == '\r' || c == ' ' || c == '\t') { pos_++; continue; }
        if (!inRecord) {
            // a new record starts here: stop first if the chunk is full (the byte stays unread)
            if ((long long)out->residues.size() > maxResidues) return 0;
            inRecord = true;
        }
        const int code = codes_[c];
        if (code < 0) {
            *error = "line " + std::to_string(line_) + ": byte " + std::to_string((int)c) + " ('" + std::string(1, (char)c) +
                     "') is not in the alphabet and the alphabet has no '*'";
            return -1;
        }
        out->residues.push_back((unsigned char)code);
        pos_++;
    }
    if (inRecord) out->offsets.push_back((long long)out->residues.size());
    return 1;
}

}  // namespace opalcli
