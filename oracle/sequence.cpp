// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Follows sequence/sequence.go and sequence/asm_amd64.s of the reference.
#include "oracle.hpp"

#include <cstring>
#include <stdexcept>

namespace dpo {

static const size_t kPad = 16;  // zero bytes appended to every store so asm over-reads are defined

// sequence.go:59 / :80 — ((b >> 1) ^ ((b & 4) >> 2)) & 3 : A0 C1 G2 T3, N->2, case-insensitive
uint8_t base_code(uint8_t b) { return (uint8_t)(((b >> 1) ^ ((b & 4) >> 2)) & 3); }

// asm_amd64.s:33-78 packBytes. do-while over 4-byte groups; the caller guarantees n >= 4.
// The PSHUFB/shift/OR dance computes, per group, c0<<6 | c1<<4 | c2<<2 | c3.
void packBytes(const uint8_t* seq, size_t n, uint8_t* data) {
    long long r8 = (long long)n;
    const uint8_t* ax = seq;
    uint8_t* bx = data;
    do {
        // 16-bit lanes [s3, s2, s1, s0] after PSHUFB; per-lane ((x>>1) ^ ((x&4)>>2)) & 3
        uint64_t x1 = (uint64_t)ax[3] | ((uint64_t)ax[2] << 16) | ((uint64_t)ax[1] << 32) | ((uint64_t)ax[0] << 48);
        uint64_t r10 = x1;
        uint64_t r9 = x1 >> 1;
        r10 &= 0x0004000400040004ULL;
        r10 >>= 2;
        r9 ^= r10;
        r9 &= 0x0003000300030003ULL;
        uint8_t dl = (uint8_t)r9;
        r9 >>= 14;
        dl |= (uint8_t)r9;
        r9 >>= 14;
        dl |= (uint8_t)r9;
        r9 >>= 14;
        dl |= (uint8_t)r9;
        *bx = dl;
        ax += 4;
        bx += 1;
        r8 -= 4;
    } while (r8 >= 4);
}

static std::shared_ptr<std::vector<uint8_t>> make_store(size_t nbytes) {
    return std::make_shared<std::vector<uint8_t>>(nbytes + kPad, (uint8_t)0);
}

// sequence.go:67-93
PackedSeq NewPackedSequence(gint id, const std::string& seq, std::shared_ptr<std::string> name) {
    gint length = (gint)seq.size() / 4;
    gint internalLength = length * 4;
    gint finalLength = (gint)seq.size() - internalLength;
    size_t nb = (seq.size() + 3) / 4;
    auto store = make_store(nb);
    uint8_t* data = store->data();
    if (internalLength >= 4) {
        packBytes((const uint8_t*)seq.data(), (size_t)internalLength, data);
    }
    if (finalLength > 0) {
        uint8_t b = 0;
        for (gint i = finalLength; i > 0; i--) {
            uint8_t nbb = (uint8_t)seq[seq.size() - i];
            nbb = base_code(nbb);
            b = (uint8_t)((b << 2) | nbb);
        }
        if (finalLength < 4) {
            b = (uint8_t)(b << (unsigned)(8 - finalLength * 2));
        }
        data[nb - 1] = b;
    }
    PackedSeq s;
    s.store = store;
    s.off = 0;
    s.nbytes = nb;
    s.id = id;
    s.offset = 0;
    s.inset = 0;
    s.name = name;
    s.length = (gint)seq.size();
    s.firstLen = 4;
    s.finalLen = finalLength;
    if (s.finalLen > s.length) s.finalLen = s.length;
    return s;
}

// sequence.go:353-370 (note Q3: `end--` before inset is computed => inset is one too large)
PackedSeq SubSequence(const PackedSeq& s, gint start, gint end) {
    if (end > s.length) end = s.length;
    end--;
    gint off = start + 4 - s.firstLen;
    gint offByte = off / 4;
    off -= offByte * 4;
    gint in = end + 4 - s.firstLen;
    gint inByte = in / 4;
    in -= inByte * 4;
    if (offByte < 0 || inByte + 1 > (gint)s.nbytes || offByte > inByte + 1)
        throw std::runtime_error("oracle: SubSequence slice bounds out of range (Go would panic)");
    PackedSeq ss;
    ss.store = s.store;
    ss.off = s.off + (size_t)offByte;
    ss.nbytes = (size_t)(inByte + 1 - offByte);
    ss.id = s.id;
    ss.offset = s.offset + start;
    ss.inset = s.inset + s.length - end;
    ss.name = s.name;
    ss.firstLen = 4 - off;
    ss.finalLen = in + 1;
    ss.length = end - start + 1;
    return ss;
}

// sequence.go:179-198
PackedSeq ReverseComplement(const PackedSeq& s) {
    auto store = make_store(s.nbytes);
    uint8_t* bs = store->data();
    const uint8_t* d = s.data();
    size_t n = s.nbytes;
    for (size_t i = 0; i < n; i++) {
        uint8_t b = (uint8_t)~d[i];
        bs[n - 1 - i] = (uint8_t)(((b & 3) << 6) | ((b & 12) << 2) | ((b & 48) >> 2) | ((b & 192) >> 6));
    }
    PackedSeq rc;
    rc.store = store;
    rc.off = 0;
    rc.nbytes = n;
    rc.id = s.id;
    rc.offset = s.inset;
    rc.inset = s.offset;
    rc.firstLen = s.finalLen;
    rc.finalLen = s.firstLen;
    rc.name = s.name;
    rc.length = s.length;
    return rc;
}

// sequence.go:242-276
std::string String(const PackedSeq& s) {
    std::vector<uint8_t> buf(s.nbytes * 4 + 8, 0);
    gint j = s.firstLen * 2 - 2;
    size_t count = 0;
    const uint8_t* d = s.data();
    for (size_t bi = 0; bi + 1 < s.nbytes; bi++) {
        uint8_t b = d[bi];
        while (j >= 0) {
            buf[count++] = (uint8_t)((b >> (unsigned)j) & 3);
            j -= 2;
        }
        j = 6;
    }
    uint8_t b = d[s.nbytes - 1];
    gint last = 8 - s.finalLen * 2;
    if (last == 8) last = 0;
    while (j >= last) {
        buf[count++] = (uint8_t)((b >> (unsigned)j) & 3);
        j -= 2;
    }
    static const char L[4] = {'A', 'C', 'G', 'T'};
    std::string out((size_t)s.length, 'A');
    for (gint i = 0; i < s.length; i++) out[(size_t)i] = L[buf[(size_t)i]];
    return out;
}

// sequence.go:164-177 — Append re-packs the concatenated strings (a raw NewPackedSequence:
// firstLen=4, finalLen=len%4), offset from the left part, inset from the right part.
PackedSeq Append(const PackedSeq& s, gint id, const PackedSeq& other) {
    std::string str = String(s) + String(other);
    PackedSeq seq = NewPackedSequence(id, str, nullptr);
    seq.offset = s.offset;
    seq.inset = other.inset;
    return seq;
}

static inline uint64_t load64be(const uint8_t* p, size_t avail) {
    // MOVQ (AX), R ; BSWAPQ R — bytes beyond the store read as zero (reference: undefined)
    uint64_t v = 0;
    for (size_t i = 0; i < 8; i++) {
        uint8_t b = (i < avail) ? p[i] : 0;
        v = (v << 8) | b;
    }
    return v;
}

// asm_amd64.s:3-30
gint packedKmerAt(const uint8_t* data, size_t avail, gint offset, gint k) {
    uint64_t bx = (uint64_t)offset;
    uint64_t cx = bx & 3;
    bx >>= 2;
    uint64_t ax = load64be(data + bx, avail > bx ? avail - bx : 0);
    cx <<= 1;
    ax <<= cx;
    uint64_t sh = 64 - (uint64_t)k * 2;
    ax >>= sh;
    return (gint)(int32_t)(uint32_t)ax;  // MOVL AX, ret (int32)
}

gint KmerAt(const PackedSeq& s, gint index, gint k) {  // sequence.go:440-442
    return packedKmerAt(s.data(), s.store->size() - s.off, index + 4 - s.firstLen, k);
}

gint NextKmer(const PackedSeq& s, gint current, gint mask, gint nextBaseIndex) {  // sequence.go:447-453
    nextBaseIndex += 4 - s.firstLen;
    uint8_t b = s.data()[nextBaseIndex / 4];
    unsigned subIndex = (unsigned)(3 - (nextBaseIndex & 3)) << 1;
    b = (uint8_t)((b >> subIndex) & 3);
    return ((current << 2) | (gint)b) & mask;
}

// asm_amd64.s:81-203 — register-level emulation. Q1: `initial` and `internal` are do-while loops.
gint packedCountKmers(const uint8_t* data, gint nbytes, size_t avail, gint upTo, gint skipFront, gint skipBack,
                      gint k, const uint8_t* seeds) {
    const uint8_t* ax = data;
    const uint8_t* const base = data;
    auto ld = [&](const uint8_t* p) -> uint64_t {
        size_t o = (size_t)(p - base);
        return load64be(p, avail > o ? avail - o : 0);
    };
    int64_t r8 = nbytes;
    int64_t si = upTo;
    int64_t bx = skipBack;
    int64_t dx = k;
    r8 -= 1;
    r8 <<= 2;
    r8 -= bx;
    r8 -= dx;
    r8 += 1;
    int64_t r15 = r8;
    r8 &= (int64_t)0xFFFFFFFFFFFFFFFCULL;
    r15 &= 3;
    dx <<= 1;
    unsigned cl = (unsigned)(64 - dx);  // CX: right shift isolating a k-mer
    int64_t r9 = 0;                     // count
    uint64_t r10 = ld(ax);
    bx = skipFront;
    bx <<= 1;
    r10 = (bx >= 64) ? 0 : (r10 << (unsigned)bx);
    // initial:
    do {
        uint64_t r12 = r10 >> cl;
        uint8_t e = seeds[r12];
        r9 = (r9 & ~(int64_t)0xFF) | (int64_t)(uint8_t)((uint8_t)r9 + e);  // ADDB R12, R9
        r10 <<= 2;
        bx += 2;
    } while (bx <= 6);
    // internal:
    for (;;) {
        ax += 1;
        uint64_t r14 = ld(ax);
        uint8_t a = seeds[r14 >> cl];
        uint8_t b = seeds[(r14 << 2) >> cl];
        uint8_t c = seeds[(r14 << 4) >> cl];
        uint8_t d = seeds[(r14 << 6) >> cl];
        uint8_t sum = (uint8_t)((uint8_t)(a + b) + (uint8_t)(c + d));
        r9 += (int64_t)sum;
        if (r9 >= si) return r9;  // CMPQ R9, SI ; JGE endtail
        r8 -= 4;
        if (!(r8 >= 4)) break;
    }
    ax += 1;
    r10 = ld(ax);
    while (r15 != 0) {
        uint64_t r12 = r10 >> cl;
        r9 += (int64_t)seeds[r12];
        r10 <<= 2;
        r15 -= 1;
    }
    return r9;
}

// asm_amd64.s:206-394 — register-level emulation.
void packedWriteSegments(const uint8_t* data, gint nbytes, size_t avail, gint skipFront, gint skipBack, gint k,
                         const uint8_t* seeds, gint* segments) {
    const uint8_t* ax = data;
    const uint8_t* const base = data;
    auto ld = [&](const uint8_t* p) -> uint64_t {
        size_t o = (size_t)(p - base);
        return load64be(p, avail > o ? avail - o : 0);
    };
    int64_t r8 = nbytes;
    int64_t bx = skipBack;
    int64_t dx = k;
    gint* r14 = segments;
    const int64_t x2 = -dx;  // -k
    r8 -= 1;
    r8 <<= 2;
    r8 -= bx;
    r8 -= dx;
    r8 += 1;
    int64_t r15 = r8;
    r8 &= (int64_t)0xFFFFFFFFFFFFFFFCULL;
    r15 &= 3;
    dx <<= 1;
    unsigned cl = (unsigned)(64 - dx);
    int64_t r9 = 0;  // running gap
    uint64_t r10 = ld(ax);
    bx = skipFront;
    bx <<= 1;
    r10 = (bx >= 64) ? 0 : (r10 << (unsigned)bx);
    auto visit = [&](uint64_t kmer) {
        if (seeds[kmer]) {
            r14[0] = r9;
            r14[1] = (gint)kmer;
            r14 += 2;
            r9 = x2;
        }
        r9 += 1;
    };
    // initial:
    do {
        visit(r10 >> cl);
        r10 <<= 2;
        bx += 2;
    } while (bx <= 6);
    // internal:
    for (;;) {
        ax += 1;
        uint64_t x1 = ld(ax);
        visit(x1 >> cl);
        visit((x1 << 2) >> cl);
        visit((x1 << 4) >> cl);
        visit((x1 << 6) >> cl);
        r8 -= 4;
        if (!(r8 >= 4)) break;
    }
    ax += 1;
    r10 = ld(ax);
    while ((int32_t)r15 != 0) {  // CMPL R15, $0
        visit(r10 >> cl);
        r10 <<= 2;
        r15 -= 1;
    }
    r9 -= x2;
    r9 -= 1;
    r14[0] = r9;
}

gint CountKmers(const PackedSeq& s, gint upTo, gint k, const uint8_t* seeds) {  // sequence.go:329-331
    return packedCountKmers(s.data(), (gint)s.nbytes, s.store->size() - s.off, upTo, 4 - s.firstLen, 4 - s.finalLen, k,
                            seeds);
}

// sequence.go:332-337 — Q4: the whole sequence's skipBack is applied to a byte-rounded sub-slice.
gint CountKmersBetween(const PackedSeq& s, gint from, gint to, gint upTo, gint k, const uint8_t* seeds) {
    gint start = (from + 4 - s.firstLen + 3) / 4;
    gint end = (to + 4 - s.firstLen) / 4;
    if (start < 0 || end > (gint)s.nbytes || start > end)
        throw std::runtime_error("oracle: CountKmersBetween slice bounds out of range (Go would panic)");
    return packedCountKmers(s.data() + start, end - start, s.store->size() - s.off - (size_t)start, upTo,
                            4 - s.firstLen, 4 - s.finalLen, k, seeds);
}

void WriteSegments(const PackedSeq& s, gint* segments, gint k, const uint8_t* seeds) {  // sequence.go:338-340
    packedWriteSegments(s.data(), (gint)s.nbytes, s.store->size() - s.off, 4 - s.firstLen, 4 - s.finalLen, k, seeds,
                        segments);
}

// sequence.go:482-504
std::vector<uint16_t> ShortKmers(const PackedSeq& s, gint k, bool collapse) {
    gint length = s.length - k + 1;
    std::vector<uint16_t> kmers((size_t)length);
    gint v = KmerAt(s, 0, k);
    gint mask = 0;
    for (gint i = 0; i < k; i++) mask = (mask << 2) | 3;
    gint prev = 0;
    gint index = 0;
    for (gint i = k; i < s.length; i++) {
        if (!collapse || v != prev || index == 0) {
            kmers[(size_t)index] = (uint16_t)v;
            prev = v;
            index++;
        }
        v = NextKmer(s, v, mask, i);
    }
    kmers[(size_t)index] = (uint16_t)v;
    index++;
    kmers.resize((size_t)index);
    return kmers;
}

// ---------------------------------------------------------------------------
// byteSequence
// ---------------------------------------------------------------------------
ByteSeq NewByteSequence(const std::string& seq) {  // sequence.go:55-63
    ByteSeq s;
    s.data.resize(seq.size());
    for (size_t i = 0; i < seq.size(); i++) s.data[i] = base_code((uint8_t)seq[i]);
    return s;
}
ByteSeq SubSequence(const ByteSeq& s, gint start, gint end) {  // sequence.go:342-351
    if (end > (gint)s.data.size()) end = (gint)s.data.size();
    ByteSeq ss;
    ss.data.assign(s.data.begin() + start, s.data.begin() + end);
    ss.offset = s.offset + start;
    ss.inset = s.inset + (gint)s.data.size() - end;
    return ss;
}
ByteSeq ReverseComplement(const ByteSeq& s) {  // sequence.go:134-148
    ByteSeq rc;
    rc.data.resize(s.data.size());
    for (size_t i = 0; i < s.data.size(); i++) rc.data[s.data.size() - 1 - i] = s.data[i] ^ 3;
    rc.offset = s.inset;
    rc.inset = s.offset;
    return rc;
}
std::string String(const ByteSeq& s) {  // sequence.go:227-241
    static const char L[4] = {'A', 'C', 'G', 'T'};
    std::string out(s.data.size(), 'A');
    for (size_t i = 0; i < s.data.size(); i++) out[i] = L[s.data[i] & 3];
    return out;
}
gint KmerAt(const ByteSeq& s, gint index, gint k) {  // sequence.go:429-435
    gint v = 0;
    for (gint i = index; i < index + k; i++) v = (v << 2) | (gint)s.data[(size_t)i];
    return v;
}
gint NextKmer(const ByteSeq& s, gint current, gint mask, gint nextBaseIndex) {  // sequence.go:444-446
    return ((current << 2) | (gint)s.data[(size_t)nextBaseIndex]) & mask;
}
gint CountKmers(const ByteSeq& s, gint upTo, gint k, gint mask, const uint8_t* kmers) {  // sequence.go:278-291
    gint seed = KmerAt(s, 0, k) >> 2;
    gint count = 0;
    for (gint i = k - 1; i < (gint)s.data.size(); i++) {
        seed = NextKmer(s, seed, mask, i);
        if (kmers[seed]) {
            count++;
            if (count >= upTo) break;
        }
    }
    return count;
}
gint CountKmersBetween(const ByteSeq& s, gint from, gint to, gint upTo, gint k, gint mask, const uint8_t* kmers) {
    gint seed = KmerAt(s, from, k) >> 2;  // sequence.go:293-306
    gint count = 0;
    for (gint i = from + k - 1; i < to; i++) {
        seed = NextKmer(s, seed, mask, i);
        if (kmers[seed]) {
            count++;
            if (count >= upTo) break;
        }
    }
    return count;
}
void WriteSegments(const ByteSeq& s, gint* segments, gint k, gint mask, const uint8_t* seeds) {  // :308-324
    gint seed = KmerAt(s, 0, k) >> 2;
    gint kmerIndex = 0;
    gint prev = 0;
    gint count = 0;
    for (gint i = k - 1; i < (gint)s.data.size(); i++) {
        seed = NextKmer(s, seed, mask, i);
        if (seeds[seed]) {
            segments[count] = kmerIndex - prev;
            segments[count + 1] = seed;
            prev = kmerIndex + k;
            count += 2;
        }
        kmerIndex++;
    }
    segments[count] = (gint)s.data.size() - prev;
}
std::vector<uint16_t> ShortKmers(const ByteSeq& s, gint k, bool collapse) {  // sequence.go:456-480
    gint length = (gint)s.data.size() - k + 1;
    std::vector<uint16_t> kmers((size_t)length);
    uint16_t mask = 0, v = 0;
    for (gint i = 0; i < k; i++) {
        mask = (uint16_t)((mask << 2) | 3);
        v = (uint16_t)((v << 2) | (uint16_t)s.data[(size_t)i]);
    }
    uint16_t prev = 0;
    gint index = 0;
    for (gint i = k; i < (gint)s.data.size(); i++) {
        if (!collapse || v != prev || index == 0) {
            kmers[(size_t)index] = v;
            prev = v;
            index++;
        }
        v = (uint16_t)(((v << 2) | (uint16_t)s.data[(size_t)i]) & mask);
    }
    kmers[(size_t)index] = v;
    index++;
    kmers.resize((size_t)index);
    return kmers;
}
gint KmerValue(const std::string& s) {  // sequence.go:520-528
    gint value = 0;
    for (char c : s) value = (value << 2) | (gint)base_code((uint8_t)c);
    return value;
}

}  // namespace dpo
