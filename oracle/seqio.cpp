// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Follows the first-pass parsing rules of sequence/seqio.go:188-267 (readFasta, no cache, no ignore list).
#include "oracle.hpp"

#include <fstream>
#include <sstream>
#include <stdexcept>

namespace dpo {

static std::string trim_space(const std::string& s) {  // strings.TrimSpace (ASCII subset)
    size_t a = 0, b = s.size();
    auto sp = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; };
    while (a < b && sp(s[a])) a++;
    while (b > a && sp(s[b - 1])) b--;
    return s.substr(a, b - a);
}

std::vector<FastaRecord> ParseFasta(const std::string& content, gint minLength) {
    std::vector<FastaRecord> out;
    size_t pos = 0;
    // bin.ReadBytes('\n'): returns the line including '\n'; at EOF returns the rest with err != nil
    auto read_line = [&](std::string& line, bool& err) {
        if (pos >= content.size()) {
            line.clear();
            err = true;
            return;
        }
        size_t nl = content.find('\n', pos);
        if (nl == std::string::npos) {
            line = content.substr(pos);
            pos = content.size();
            err = true;
        } else {
            line = content.substr(pos, nl + 1 - pos);
            pos = nl + 1;
            err = false;
        }
    };
    std::string buf;
    bool err = false;
    bool isFastq = false;
    std::string lastName;
    read_line(buf, err);  // seqio.go:189-203: the first line is always a header
    if (err) return out;
    if (buf[0] == '@') isFastq = true;
    lastName = buf.substr(1);
    for (read_line(buf, err); buf.size() > 0 || !err; read_line(buf, err)) {  // seqio.go:207
        if (buf.empty()) break;
        if (buf[0] >= 'A' && buf[0] <= 'T') {
            bool readSeq = (gint)buf.size() >= minLength;
            if (readSeq) {
                FastaRecord r;
                r.name = trim_space(lastName);
                r.seq = buf.substr(0, buf.size() - 1);  // buf[:len(buf)-1] — drops the last byte even without '\n'
                out.push_back(std::move(r));
            }
            if (isFastq) {
                std::string plus;
                bool e2;
                read_line(plus, e2);
                if (e2 || plus.empty() || plus[0] != '+') throw std::runtime_error("Invalid fastq format (on + line)");
                std::string qual;
                read_line(qual, e2);
                buf = qual;  // the reference's `buf` now holds the quality line; only its length matters below
            }
        } else if (buf[0] == '@') {
            isFastq = true;
            lastName = buf.substr(1);
        } else {
            lastName = buf.substr(1);
        }
        if (err) break;
    }
    return out;
}

std::vector<FastaRecord> ReadFasta(const std::string& filename, gint minLength) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + filename);
    std::stringstream ss;
    ss << f.rdbuf();
    return ParseFasta(ss.str(), minLength);
}

}  // namespace dpo
