// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Follows util/bitset.go and util/asm_amd64.s of the reference.
#include "oracle.hpp"

#include <stdexcept>

namespace dpo {

static const uint64_t Bit = 1;

IntSet NewIntSet() {  // bitset.go:20-23
    IntSet s;
    s.vs.assign(50, 0);
    s.start = 1;
    s.end = 0;
    s.count = 0;
    return s;
}

IntSet NewIntSetCapacity(gint capacity) {  // bitset.go:25-28
    IntSet s;
    s.vs.assign((size_t)(capacity / 64 + 1), 0);
    s.start = 1;
    s.end = 0;
    s.count = 0;
    return s;
}

bool Contains(const IntSet& set, uint64_t x) {  // bitset.go:65-72
    uint64_t index = x >> 6;
    if (index < set.start || index > set.end) return false;
    uint64_t subIndex = x & 0x3F;
    return (set.vs[index] & (Bit << subIndex)) != 0;
}

void Add(IntSet& set, uint64_t x) {  // bitset.go:74-108
    uint64_t index = x >> 6;
    uint64_t subIndex = x & 0x3F;
    uint64_t bit = Bit << subIndex;
    if ((gint)index >= (gint)set.vs.size()) {
        set.vs.resize(index + 2, 0);  // newVs := make([]uint64, index+2); copy
    }
    if (set.end < set.start) {
        set.start = index;
        set.end = index;
        set.vs[index] = bit;
        set.count = 1;
        return;
    }
    if (index < set.start) {
        set.start = index;
        set.vs[index] = bit;
        set.count++;
        return;
    }
    if (index > set.end) {
        set.end = index;
        set.vs[index] = bit;
        set.count++;
        return;
    }
    uint64_t old = set.vs[index];
    if ((old & bit) != 0) return;
    set.vs[index] = old | bit;
    set.count++;
}

void Clear(IntSet& set) {  // bitset.go:145-153
    while (set.start <= set.end) {
        set.vs[set.start] = 0;
        set.start++;
    }
    set.end = 0;
    set.start = (uint64_t)set.vs.size() + 1;
    set.count = 0;
}

uint64_t CountIntersection(const IntSet& set, const IntSet& other) {  // bitset.go:163-177
    uint64_t start = set.start, end = set.end;
    if (other.start > start) start = other.start;
    if (end > other.end) end = other.end;
    uint64_t count = 0;
    for (; start <= end; start++) count += (uint64_t)__builtin_popcountll(set.vs[start] & other.vs[start]);
    return count;
}

// asm_amd64.s:14-117. Blocks of 8 words with an early exit tested only at block starts, then a word-wise tail.
uint64_t countIntersectionToAsm(const uint64_t* a, const uint64_t* b, gint n, gint maxCount) {
    int64_t cx = n;
    int64_t si = maxCount;
    int64_t dx = 0;
    for (;;) {
        if (cx <= 7) break;       // CMPQ CX,$7 ; JLE tail
        if (dx >= si) return dx;  // CMPQ DX,SI ; JGE finished
        for (int i = 0; i < 8; i++) dx += __builtin_popcountll(a[i] & b[i]);
        a += 8;
        b += 8;
        cx -= 8;
    }
    while (cx > 0) {
        dx += __builtin_popcountll(a[0] & b[0]);
        a++;
        b++;
        cx--;
    }
    return (uint64_t)dx;
}

uint64_t CountIntersectionTo(const IntSet& set, const IntSet& other, gint maxCount) {  // bitset.go:179-195
    uint64_t start = set.start, end = set.end;
    if (other.start > start) start = other.start;
    if (end > other.end) end = other.end;
    // set.vs[start:end+1] — Go panics if start > end+1 (or end+1 beyond capacity)
    if (start > end + 1 || end + 1 > set.vs.size() || end + 1 > other.vs.size())
        throw std::runtime_error("oracle: CountIntersectionTo slice bounds out of range (Go would panic)");
    return countIntersectionToAsm(set.vs.data() + start, other.vs.data() + start, (gint)(end + 1 - start), maxCount);
}

// asm_amd64.s:121-193
void getSoftUnion4Asm(const uint64_t* ax, gint n, uint64_t out[4]) {
    int64_t bx = n;
    uint64_t r8 = 0, r9 = 0, r10 = 0, r11 = 0, dx, cx;
    if (!(bx <= 3)) {
        r8 = ax[0];
        dx = ax[1];
        r9 = r8;
        r9 &= dx;
        r8 |= dx;
        dx = ax[2];
        r10 = r9;
        r10 &= dx;
        cx = r8 & dx;
        r9 |= cx;
        r8 |= dx;
        dx = ax[3];
        r11 = r10;
        r11 &= dx;
        cx = r9 & dx;
        r10 |= cx;
        cx = r8 & dx;
        r9 |= cx;
        r8 |= dx;
        bx -= 4;
        ax += 4;
    }
    while (!(bx <= 0)) {
        dx = ax[0];
        cx = r10 & dx;
        r11 |= cx;
        cx = r9 & dx;
        r10 |= cx;
        cx = r8 & dx;
        r9 |= cx;
        r8 |= dx;
        bx--;
        ax++;
    }
    out[0] = r8;
    out[1] = r9;
    out[2] = r10;
    out[3] = r11;
}

// asm_amd64.s:196-314. Q7: with n <= 5 the routine jumps to the loop with R8-R11 uninitialised
// (whatever the Go caller left in them). Canonical choice here: zeros.
void getSoftUnion8Asm(const uint64_t* ax, gint n, uint64_t out[4]) {
    int64_t bx = n;
    uint64_t r8 = 0, r9 = 0, r10 = 0, r11 = 0;  // undefined in the reference when n <= 5
    uint64_t r12 = 0, r13 = 0, r14 = 0, r15 = 0, dx, cx;
    if (!(bx <= 5)) {
        r8 = ax[0];
        dx = ax[1];
        r9 = r8 & dx;
        r8 |= dx;
        dx = ax[2];
        r10 = r9 & dx;
        cx = r8 & dx;
        r9 |= cx;
        r8 |= dx;
        dx = ax[3];
        r11 = r10 & dx;
        cx = r9 & dx;
        r10 |= cx;
        cx = r8 & dx;
        r9 |= cx;
        r8 |= dx;
        dx = ax[4];
        r12 = r11 & dx;
        cx = r10 & dx;
        r11 |= cx;
        cx = r9 & dx;
        r10 |= cx;
        cx = r8 & dx;
        r9 |= cx;
        r8 |= dx;
        dx = ax[5];
        r13 = r12 & dx;
        cx = r11 & dx;
        r12 |= cx;
        cx = r10 & dx;
        r11 |= cx;
        cx = r9 & dx;
        r10 |= cx;
        cx = r8 & dx;
        r9 |= cx;
        r8 |= dx;
        r14 = 0;
        r15 = 0;
        bx -= 6;
        ax += 6;
    }
    while (!(bx <= 0)) {
        dx = ax[0];
        cx = r14 & dx;
        r15 |= cx;
        cx = r13 & dx;
        r14 |= cx;
        cx = r12 & dx;
        r13 |= cx;
        cx = r11 & dx;
        r12 |= cx;
        cx = r10 & dx;
        r11 |= cx;
        cx = r9 & dx;
        r10 |= cx;
        cx = r8 & dx;
        r9 |= cx;
        r8 |= dx;
        ax++;
        bx--;
    }
    out[0] = r12;
    out[1] = r13;
    out[2] = r14;
    out[3] = r15;
}

// asm_amd64.s:317-509. Unconditional 8-step unroll (the caller guarantees n >= 13). Q6: step 8 has no
// `ORQ DX, R8`, so v1 misses the 8th word. X0..X3 are two-lane registers: hi lane = v12,v11,v10,v9 and
// lo lane = v16,v15,v14,v13 respectively.
void getSoftUnion16Asm(const uint64_t* ax, gint n, uint64_t out[4]) {
    int64_t bx = n;
    uint64_t x0lo = 0, x0hi = 0, x1lo = 0, x1hi = 0, x2lo = 0, x2hi = 0, x3lo = 0, x3hi = 0;
    uint64_t r8, r9, r10, r11, r12, r13, r14, r15, dx, cx;
    if (n < 8) throw std::runtime_error("oracle: getSoftUnion16Asm with n < 8 reads out of bounds in the reference");
    r8 = ax[0];
    dx = ax[1];
    r9 = r8 & dx;
    r8 |= dx;
    dx = ax[2];
    r10 = r9 & dx;
    cx = r8 & dx;
    r9 |= cx;
    r8 |= dx;
    dx = ax[3];
    r11 = r10 & dx;
    cx = r9 & dx;
    r10 |= cx;
    cx = r8 & dx;
    r9 |= cx;
    r8 |= dx;
    dx = ax[4];
    r12 = r11 & dx;
    cx = r10 & dx;
    r11 |= cx;
    cx = r9 & dx;
    r10 |= cx;
    cx = r8 & dx;
    r9 |= cx;
    r8 |= dx;
    dx = ax[5];
    r13 = r12 & dx;
    cx = r11 & dx;
    r12 |= cx;
    cx = r10 & dx;
    r11 |= cx;
    cx = r9 & dx;
    r10 |= cx;
    cx = r8 & dx;
    r9 |= cx;
    r8 |= dx;
    dx = ax[6];
    r14 = r13 & dx;
    cx = r12 & dx;
    r13 |= cx;
    cx = r11 & dx;
    r12 |= cx;
    cx = r10 & dx;
    r11 |= cx;
    cx = r9 & dx;
    r10 |= cx;
    cx = r8 & dx;
    r9 |= cx;
    r8 |= dx;
    dx = ax[7];
    r15 = r14 & dx;
    cx = r13 & dx;
    r14 |= cx;
    cx = r12 & dx;
    r13 |= cx;
    cx = r11 & dx;
    r12 |= cx;
    cx = r10 & dx;
    r11 |= cx;
    cx = r9 & dx;
    r10 |= cx;
    cx = r8 & dx;
    r9 |= cx;
    // (no ORQ DX, R8 here — Q6)
    bx -= 8;
    ax += 8;
    while (!(bx <= 0)) {
        uint64_t x4lo = x1lo, x4hi = x1hi;  // MOVOA X1, X4
        uint64_t x5lo = x2lo, x5hi = x2hi;
        uint64_t x6lo = x3lo, x6hi = x3hi;
        cx = x0hi;  // PEXTRQ $1, X0, CX
        dx = ax[0];
        uint64_t x7lo = dx, x7hi = dx;  // MOVQ DX,X7 ; MOVLHPS X7,X7
        x4lo &= x7lo;
        x4hi &= x7hi;
        x5lo &= x7lo;
        x5hi &= x7hi;
        x6lo &= x7lo;
        x6hi &= x7hi;
        x0lo |= x4lo;
        x0hi |= x4hi;
        x1lo |= x5lo;
        x1hi |= x5hi;
        x2lo |= x6lo;
        x2hi |= x6hi;
        x4lo = cx;   // MOVQ CX, X4 (upper lane zeroed)
        x4hi = r15;  // PINSRQ $1, R15, X4
        x7lo &= x4lo;
        x7hi &= x4hi;
        x3lo |= x7lo;
        x3hi |= x7hi;
        cx = r14 & dx;
        r15 |= cx;
        cx = r13 & dx;
        r14 |= cx;
        cx = r12 & dx;
        r13 |= cx;
        cx = r11 & dx;
        r12 |= cx;
        cx = r10 & dx;
        r11 |= cx;
        cx = r9 & dx;
        r10 |= cx;
        cx = r8 & dx;
        r9 |= cx;
        r8 |= dx;
        ax++;
        bx--;
    }
    out[0] = x3lo;  // v13
    out[1] = x2lo;  // v14
    out[2] = x1lo;  // v15
    out[3] = x0lo;  // v16
}

// bitset.go:509-538
static void addSoftUnionIDs(uint64_t v, const std::vector<uint64_t>& vs, gint n, gint minCount,
                            std::vector<uint64_t>& ids, uint64_t offset) {
    uint64_t bit = Bit;
    uint64_t zs = (uint64_t)__builtin_ctzll(v);
    bit <<= zs;
    v >>= zs;
    for (uint64_t j = zs; j < 64 && v != 0; j++) {
        if ((Bit & v) != 0) {
            gint count = 0;
            for (gint k = 0; k < n; k++) {
                if ((vs[(size_t)k] & bit) != 0) {
                    count++;
                    if (count >= minCount) {
                        ids.push_back(offset + j);
                        break;
                    }
                } else if (n - k + count <= minCount) {
                    break;
                }
            }
        }
        v >>= 1;
        bit <<= 1;
    }
}

// bitset.go:308-411
std::vector<uint64_t> GetSharedIDs(const std::vector<const IntSet*>& sets, gint minCount, bool fast) {
    std::vector<uint64_t> ids;
    if (minCount > 24) fast = false;
    gint n = (gint)sets.size();
    uint64_t start = (uint64_t)sets[0]->vs.size();
    uint64_t end = 0;
    std::vector<uint64_t> lens((size_t)n);
    std::vector<const std::vector<uint64_t>*> vs((size_t)n);
    uint64_t shortest = start;
    for (gint i = 0; i < n; i++) {
        vs[(size_t)i] = &sets[(size_t)i]->vs;
        lens[(size_t)i] = sets[(size_t)i]->end + 1;
        if (sets[(size_t)i]->start < start) start = sets[(size_t)i]->start;
        if (sets[(size_t)i]->end > end) end = sets[(size_t)i]->end;
        if (lens[(size_t)i] < shortest) shortest = lens[(size_t)i];
    }
    std::vector<uint64_t> nextVs((size_t)n);
    for (uint64_t i = start; i <= end; i++) {
        if (shortest <= i) {
            uint64_t nextShortest = end;
            for (gint j = 0; j < n; j++) {
                if (lens[(size_t)j] <= i) {
                    gint last = n - 1;
                    if (last < minCount) return ids;
                    vs[(size_t)j] = vs[(size_t)last];
                    lens[(size_t)j] = lens[(size_t)last];
                    n = last;
                    j--;
                } else if (lens[(size_t)j] < nextShortest) {
                    nextShortest = lens[(size_t)j];
                }
            }
            shortest = nextShortest;
        }
        for (gint j = 0; j < n; j++) {
            const std::vector<uint64_t>& w = *vs[(size_t)j];
            if (i >= w.size()) throw std::runtime_error("oracle: GetSharedIDs index out of range (Go would panic)");
            nextVs[(size_t)j] = w[i];
        }
        uint64_t v = 0;
        uint64_t o[4];
        if (minCount >= 13) {
            getSoftUnion16Asm(nextVs.data(), n, o);
            if (minCount >= 16) v = o[3];
            else if (minCount == 15) v = o[2];
            else if (minCount == 14) v = o[1];
            else v = o[0];
        } else if (minCount >= 5) {
            getSoftUnion8Asm(nextVs.data(), n, o);
            if (minCount >= 8) v = o[3];
            else if (minCount == 7) v = o[2];
            else if (minCount == 6) v = o[1];
            else v = o[0];
        } else {
            getSoftUnion4Asm(nextVs.data(), n, o);
            if (minCount == 4) v = o[3];
            else if (minCount == 3) v = o[2];
            else if (minCount == 2) v = o[1];
            else v = o[0];
        }
        if (v != 0) {
            if (fast) {
                uint64_t shifted = 0;
                while (v != 0) {
                    uint64_t zs = (uint64_t)__builtin_ctzll(v);
                    ids.push_back((i << 6) + shifted + zs);
                    v = (zs + 1 >= 64) ? 0 : (v >> (zs + 1));  // Go: shift >= 64 yields 0
                    shifted += zs + 1;
                }
            } else {
                addSoftUnionIDs(v, nextVs, n, minCount, ids, i << 6);
            }
        }
    }
    return ids;
}

uint64_t CountMembers(IntSet& set) {  // bitset.go:584-591
    uint64_t count = 0;
    for (uint64_t i = set.start; i <= set.end; i++) count += (uint64_t)__builtin_popcountll(set.vs[i]);
    set.count = count;
    return count;
}

std::vector<uint64_t> AsUints(const IntSet& set) {
    std::vector<uint64_t> ids;
    if (set.start > set.end) return ids;
    for (uint64_t i = set.start; i <= set.end; i++) {
        uint64_t v = set.vs[i];
        while (v) {
            ids.push_back(i * 64 + (uint64_t)__builtin_ctzll(v));
            v &= v - 1;
        }
    }
    return ids;
}

}  // namespace dpo
