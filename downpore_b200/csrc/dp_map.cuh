// downpore_b200 — map-side kernels (sm_100a): seed extraction, seed-index lookup, chaining.
// One performMapping() call (mapping/mapping.go:489-611) = one DpWindow; the three kernels below are its stages.
#pragma once
#include "dp_common.cuh"

// ===============================================================================================================
// Stage 1 — seed extraction: SeedIndex.NewSeedSequence on the window and on its reverse complement
// (seeds/seeds.go:33-50; the asm scans sequence/asm_amd64.s:81-394).
//
// One warp per window, persistent CTAs (one per SM) that keep a prefix filter of the seed table in shared memory.
// The window is cut into blocks of 1024 forward k-mer positions; inside a block every lane owns 32 CONSECUTIVE
// positions, so its 32+k-1 bases live in three registers (two coalesced loads, two shuffles, three funnel shifts)
// and every k-mer of either strand is one funnel shift away:
//   pass 1 (fully unrolled, no global memory): both strands' k-mers -> filter bit -> two 32-bit candidate masks;
//   pass A: only the filter positives (~6 %) gather the exact flag word from the L2-resident table -> hit masks;
//   one atomicAdd allocates the window's slice of the compact output;
//   pass B: the hits gather {flags, rank} once more and write (seed rank, scan position) in scan order.
// The reverse-complement strand visits the same positions backwards (rc position = L-k-j), so one pass serves both.
//
// Scan position = visit index of the reference's asm scan, which equals the base offset except for the raw-sequence
// quirks (Q2): an un-sliced read with len%4==0 loses its last four bases on the forward strand, and on the
// reverse-complement strand skips four bases and visits its first k-mer twice.
// ===============================================================================================================
struct DpExtractOut {
    unsigned* wsOff;  // [2*nWin] first entry of each window strand
    int* wsN;         // [2*nWin] seeds per window strand
    unsigned* qSeed;  // compact seed ranks
    int* qPos;        // compact scan positions
    unsigned long long* cursor;  // bump allocator over qSeed/qPos
};

// reverse complement of 16 packed bases (first base in the top bits)
__device__ __forceinline__ unsigned dp_revcomp16(unsigned w) {
    unsigned y = __brev(~w);
    return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}

// The 48 bases a lane owns in one block: fwd a0:a1:a2 (first base on top), and their reverse complement pre-shifted
// so that the reverse complement of the k-mer at lane position i starts at base 32-i of r0:r1:r2.
struct DpLaneBases {
    unsigned a0, a1, a2, r0, r1, r2;
};

__device__ __forceinline__ DpLaneBases dp_lane_bases(const unsigned* __restrict__ words, long long firstBase, bool live,
                                                     int k, unsigned lane) {
    // words g..g+3 of the lane; g advances by two per lane, so the upper pair is the next lane's lower pair
    // (`live`: this lane or its predecessor owns a visited position, so its words lie inside the packed window)
    unsigned w0 = 0, w1 = 0, w2, w3;
    const unsigned* p = words + (firstBase >> 4);
    if (live) {
        w0 = __ldg(p);
        w1 = __ldg(p + 1);
    }
    w2 = __shfl_down_sync(DP_FULL, w0, 1);
    w3 = __shfl_down_sync(DP_FULL, w1, 1);
    if (lane == 31) {
        w2 = live ? __ldg(p + 2) : 0u;
        w3 = live ? __ldg(p + 3) : 0u;
    }
    const unsigned sh = ((unsigned)firstBase & 15u) * 2u;
    DpLaneBases B;
    B.a0 = __funnelshift_l(w1, w0, sh);
    B.a1 = __funnelshift_l(w2, w1, sh);
    B.a2 = __funnelshift_l(w3, w2, sh);
    const unsigned c0 = dp_revcomp16(B.a2), c1 = dp_revcomp16(B.a1), c2 = dp_revcomp16(B.a0);
    const unsigned rs = (unsigned)(16 - k) * 2u;  // k <= 15
    B.r0 = __funnelshift_l(c1, c0, rs);
    B.r1 = __funnelshift_l(c2, c1, rs);
    B.r2 = c2 << rs;
    return B;
}

// k-mer (top-aligned in 32 bits) at lane position i, forward strand / reverse complement; i may be a run-time value
__device__ __forceinline__ unsigned dp_fwd_at(const DpLaneBases& B, unsigned i) {
    return __funnelshift_l(i < 16 ? B.a1 : B.a2, i < 16 ? B.a0 : B.a1, 2u * i);
}
__device__ __forceinline__ unsigned dp_rc_at(const DpLaneBases& B, unsigned i) {
    // base 32-i: i == 0 -> r2 itself; 1..16 -> inside r1; 17..31 -> inside r0; shift = 2*((32-i) & 15)
    if (i == 0) return B.r2;
    return __funnelshift_l(i <= 16 ? B.r2 : B.r1, i <= 16 ? B.r1 : B.r0, (64u - 2u * i));
}

__global__ void __launch_bounds__(1024, 1) dp_extract_kernel(DpIndexDev I, const unsigned* __restrict__ readWords,
                                                             const long long* __restrict__ readWordOff,
                                                             const DpWindow* __restrict__ wins, int nWin,
                                                             DpExtractOut O, int maskWords,
                                                             DpCounters* __restrict__ ctr) {
    // shared memory: [ prefix filter over the seed k-mers | per-warp hit masks: maskWords x {fwd, rc} ]
    extern __shared__ unsigned dp_smem[];
    const unsigned lane = dp_lane();
    const int warpInBlock = threadIdx.x >> 5;
    const int fBits = I.filterBits;
    const unsigned fWords = fBits ? ((1u << fBits) + 31u) >> 5 : 0u;
    const unsigned fWordsPad = (fWords + 3u) & ~3u;
    const unsigned* filt = dp_smem;
    if (fBits) {
        for (unsigned i = threadIdx.x; i < fWords; i += blockDim.x) dp_smem[i] = __ldg(I.filter + i);
        __syncthreads();
    }
    unsigned* mF = dp_smem + fWordsPad + (size_t)warpInBlock * 2 * maskWords;
    unsigned* mR = mF + maskWords;
    const int k = I.k;
    const unsigned kShift = 32u - 2u * (unsigned)k;
    const unsigned fShift = 32u - (unsigned)fBits;  // top-aligned k-mer -> filter bit index
    const uint2* __restrict__ table = I.table;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nWarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long lookups = 0, seeds = 0;
    for (int w = warp; w < nWin; w += nWarps) {
        DpWindow win = wins[w];
        const int L = win.len;
        if (L <= 0) {  // empty second slot of a short read
            if (lane == 0) {
                O.wsOff[2 * w] = 0;
                O.wsN[2 * w] = 0;
                O.wsOff[2 * w + 1] = 0;
                O.wsN[2 * w + 1] = 0;
            }
            continue;
        }
        const unsigned* words = readWords + readWordOff[win.read];
        const bool q2 = win.whole && ((L & 3) == 0);
        const int nJ = L - k + 1 - (q2 ? 4 : 0);  // forward positions 0..nJ-1 are visited by both strands
        const int nBlk = (nJ + 1023) >> 10;
        unsigned cnt = 0;  // per lane: forward hits | rc hits << 16, over all blocks
        for (int b = 0; b < nBlk; b++) {
            const int pb = (b << 10) + ((int)lane << 5);  // first forward position of this lane
            const int nValid = min(32, max(0, nJ - pb));
            DpLaneBases B = dp_lane_bases(words, (long long)win.start + pb, pb < nJ + 32, k, lane);
            unsigned fm, rm;
            if (fBits) {
                fm = 0;
                rm = 0;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    unsigned x = i < 16 ? __funnelshift_l(B.a1, B.a0, 2 * i) : __funnelshift_l(B.a2, B.a1, 2 * (i - 16));
                    unsigned y = i == 0 ? B.r2
                                        : (i <= 16 ? __funnelshift_l(B.r2, B.r1, 2 * (16 - i))
                                                   : __funnelshift_l(B.r1, B.r0, 2 * (32 - i)));
                    unsigned hx = x >> fShift, hy = y >> fShift;
                    unsigned bx = __funnelshift_r(filt[hx >> 5], 0u, hx) & 1u;
                    unsigned by = __funnelshift_r(filt[hy >> 5], 0u, hy) & 1u;
                    fm |= bx << i;
                    rm |= by << i;
                }
                const unsigned valid = nValid >= 32 ? 0xffffffffu : ((1u << nValid) - 1u);
                fm &= valid;
                rm &= valid;
            } else {
                fm = rm = nValid >= 32 ? 0xffffffffu : ((1u << nValid) - 1u);
            }
            // pass A: exact flags for the filter positives
            // (two positives per strand per trip: all four table loads are issued before the first is used)
            unsigned hf = 0, hr = 0;
            while (fm | rm) {
                unsigned wF0 = 0, wF1 = 0, wR0 = 0, wR1 = 0;  // flag words; a missing positive contributes none
                unsigned sF0 = 0, sF1 = 0, sR0 = 0, sR1 = 0;  // (kmer & 31) | position << 8
                if (fm) {
                    const unsigned i = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const unsigned kmer = dp_fwd_at(B, i) >> kShift;
                    wF0 = __ldg(&table[kmer >> 5].x);
                    sF0 = (kmer & 31u) | (i << 8);
                }
                if (rm) {
                    const unsigned i = __ffs(rm) - 1;
                    rm &= rm - 1;
                    const unsigned kmer = dp_rc_at(B, i) >> kShift;
                    wR0 = __ldg(&table[kmer >> 5].x);
                    sR0 = (kmer & 31u) | (i << 8);
                }
                if (fm) {
                    const unsigned i = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const unsigned kmer = dp_fwd_at(B, i) >> kShift;
                    wF1 = __ldg(&table[kmer >> 5].x);
                    sF1 = (kmer & 31u) | (i << 8);
                }
                if (rm) {
                    const unsigned i = __ffs(rm) - 1;
                    rm &= rm - 1;
                    const unsigned kmer = dp_rc_at(B, i) >> kShift;
                    wR1 = __ldg(&table[kmer >> 5].x);
                    sR1 = (kmer & 31u) | (i << 8);
                }
                hf |= ((wF0 >> (sF0 & 31u)) & 1u) << (sF0 >> 8);
                hf |= ((wF1 >> (sF1 & 31u)) & 1u) << (sF1 >> 8);
                hr |= ((wR0 >> (sR0 & 31u)) & 1u) << (sR0 >> 8);
                hr |= ((wR1 >> (sR1 & 31u)) & 1u) << (sR1 >> 8);
            }
            mF[(b << 5) + lane] = hf;
            mR[(b << 5) + lane] = hr;
            cnt += __popc(hf) | (__popc(hr) << 16);
        }
        __syncwarp();
        unsigned tot = cnt;
        for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(DP_FULL, tot, d);
        const int nF = tot & 0xffff, nR = tot >> 16;
        // Q2 on the rc strand: its first visited k-mer (forward position nJ-1) is visited twice
        int dup = 0;
        if (q2) dup = (mR[(nJ - 1) >> 5] >> ((nJ - 1) & 31)) & 1;
        unsigned base = 0;
        if (lane == 0) {
            base = (unsigned)atomicAdd(O.cursor, (unsigned long long)(nF + nR + dup));
            O.wsOff[2 * w] = base;
            O.wsN[2 * w] = nF;
            O.wsOff[2 * w + 1] = base + nF;
            O.wsN[2 * w + 1] = nR + dup;
        }
        base = __shfl_sync(DP_FULL, base, 0);
        const unsigned baseR = base + nF;
        const int rcShift = q2 ? 3 : 0;  // rc scan position = (L-k-j) - rcShift
        // pass B: ranks of the hits, written in scan order (forward ascending, rc descending in j)
        unsigned blockBase = 0;  // hits of earlier blocks: fwd | rc << 16
        for (int b = 0; b < nBlk; b++) {
            unsigned hf = mF[(b << 5) + lane], hr = mR[(b << 5) + lane];
            const unsigned mine = __popc(hf) | (__popc(hr) << 16);
            unsigned incl = mine;
            for (int d = 1; d < 32; d <<= 1) {
                unsigned y = __shfl_up_sync(DP_FULL, incl, d);
                if ((int)lane >= d) incl += y;
            }
            const unsigned before = blockBase + incl - mine;
            blockBase += __shfl_sync(DP_FULL, incl, 31);
            if (__any_sync(DP_FULL, (hf | hr) != 0)) {
                const int pb = (b << 10) + ((int)lane << 5);
                DpLaneBases B = dp_lane_bases(words, (long long)win.start + pb, pb < nJ + 32, k, lane);
                unsigned idxF = base + (before & 0xffff);
                int below = (int)(before >> 16);  // rc hits at smaller forward positions
                while (hf | hr) {
                    if (hf) {
                        unsigned i = __ffs(hf) - 1;
                        hf &= hf - 1;
                        unsigned kmer = dp_fwd_at(B, i) >> kShift;
                        uint2 e = __ldg(table + (kmer >> 5));
                        O.qSeed[idxF] = e.y + __popc(e.x & ((1u << (kmer & 31)) - 1u));
                        O.qPos[idxF] = pb + (int)i;
                        idxF++;
                    }
                    if (hr) {
                        unsigned i = __ffs(hr) - 1;
                        hr &= hr - 1;
                        unsigned kmer = dp_rc_at(B, i) >> kShift;
                        uint2 e = __ldg(table + (kmer >> 5));
                        unsigned rank = e.y + __popc(e.x & ((1u << (kmer & 31)) - 1u));
                        const int j = pb + (int)i;
                        unsigned idx = baseR + dup + (unsigned)(nR - 1 - below);
                        O.qSeed[idx] = rank;
                        O.qPos[idx] = (L - k - j) - rcShift;
                        if (dup && j == nJ - 1) {  // the double visit: scan positions 0 and 1
                            O.qSeed[baseR] = rank;
                            O.qPos[baseR] = 0;
                        }
                        below++;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            lookups += 2ull * (unsigned)nJ + (q2 ? 1u : 0u);  // Q2: the rc scan visits one k-mer twice
            seeds += (unsigned)(nF + nR + dup);
        }
    }
    if (lane == 0 && (lookups | seeds)) {
        atomicAdd(&ctr->kmer_lookups, lookups);
        atomicAdd(&ctr->query_seeds, seeds);
    }
}

// ===============================================================================================================
// Stage 2 — seed-index lookup: SeedIndex.Matches -> util.GetSharedIDs -> getSoftUnion{4,8,16}Asm
// (seeds/seeds.go:335-353; util/bitset.go:308-411; util/asm_amd64.s:121-509), restated over posting lists
// (SURVEY.md Appendix C).  One warp per window strand.
//
//   E        = included seed occurrences: |D(s)| < C and s != previous eligible seed
//   n < 5    -> no candidates;  minCount = (n+2)>>2  (= int(0.25*n + 0.5))
//   count[c] = #{ j in E : c in D(E_j) }   gathered from the posting runs with shared-memory atomics
//   level    : T = minCount for <=8 and 13..16; 9..12 -> 8; 17..24 -> 16; >24 -> exact minCount
//   clamped levels also need live(c>>6) = #{ j : max(D(E_j))>>6 >= c>>6 } >= minCount (whole-search early stop, Q11)
//   minCount >= 13: the level-16 routine under-counts by one a chunk that is in column slot 7 but in none of slots
//   0..6 of the current column order (Q6); the column order follows the swap-with-last drops of bitset.go:335-349.
// The same pass counts, per chunk, the DISTINCT query seeds it contains (upper 16 bits of the counter), which is what
// IntSet.CountIntersectionTo decides on later (mapping.go:520-523).
// ===============================================================================================================
#define DP_ECAP 128   // included seed occurrences per window strand held in shared memory
#define DP_TCAP 256   // touched chunks per window strand held in shared memory
#define DP_LWARPS 4   // warps per lookup CTA
#define DP_DUPCAP 32  // included runs that repeat an earlier run's seed, listed per window strand

struct DpLookupScratch {  // per-warp global scratch for oversized window strands, `stride` entries per array
    unsigned* eSeed;
    unsigned* eOff;
    unsigned* ePre;    // exclusive prefix of posting run lengths (stride+1)
    unsigned* eEndW;   // last 64-chunk word of each included run (0 if empty), lens = eEndW+1
    unsigned char* eFirst;  // first occurrence of the seed among E
    unsigned* allSeeds;     // seeds present in every chunk (|D| >= C)
    unsigned short* order;  // Q6 column-order simulation
    unsigned* touched;      // chunks whose counter left zero [tStride]
    unsigned long long* cand;  // (chunk << 32 | counter) of chunks over the threshold [2*tStride]: list + sort space
    unsigned* counters;     // [(C+1)/2] per warp when they do not fit shared memory; all zero between window strands
    int stride;
    int tStride;
};

__device__ __forceinline__ bool dp_run_contains(const unsigned* __restrict__ chunks, unsigned off, unsigned cnt,
                                                unsigned c) {
    unsigned lo = 0, hi = cnt;
    while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        unsigned v = __ldg(chunks + off + mid);
        if (v < c) lo = mid + 1;
        else hi = mid;
    }
    return lo < cnt && __ldg(chunks + off + lo) == c;
}

// Everything the candidate refinement needs about the included runs of one window strand
struct DpRefineCtx {
    const unsigned* eOff;   // [nInc] first posting of each included run (in seedChunks)
    const unsigned* ePre;   // [nInc+1] exclusive prefix of run lengths
    const unsigned* eEndW;  // [nInc] last 64-chunk word of each run (valid when clamped || q6)
    const unsigned char* eFirst;  // [nInc] first occurrence of the seed among the included runs
    const unsigned short* dup;    // the runs that are NOT first occurrences (valid when 0 <= nDup)
    int nDup;                     // -1: too many to list, use eFirst
    unsigned short* order;        // [nInc] scratch of the Q6 column-order simulation
    int* sim;                     // [2] shared memory: simulation state (word reached, live sets)
    int nInc, minCount, T;
    bool clamped, q6;
    int nAllDistinct;
};

// Q6 column-order simulation: advances the drop simulation of bitset.go:332-353 to word `wword`. The reference visits
// every word i and drops (swap with the last live set, in scan order) the sets whose last word is before i; nothing
// happens at a word where no set ends, so only the words where a live set ends are visited. Warp-collective; the state
// (simWord, simLive) is warp-uniform and persists over the candidates of one window strand (ascending words).
__device__ __noinline__ void dp_q6_advance(const DpIndexDev& I, const DpRefineCtx& X, unsigned wword, bool simInit) {
    const unsigned lane = dp_lane();
    int simWord = X.sim[0], simLive = simInit ? X.sim[1] : X.nInc;
    const int nInc = X.nInc;
    const unsigned* eOff = X.eOff;
    const unsigned* ePre = X.ePre;
    const unsigned* eEndW = X.eEndW;
    unsigned short* order = X.order;
    if (!simInit) {
        for (int j = (int)lane; j < nInc; j += 32) order[j] = (unsigned short)j;
        unsigned st = 0xffffffffu;  // min over sets of IntSet.start (1 if empty)
        for (int j = (int)lane; j < nInc; j += 32) {
            unsigned len = ePre[j + 1] - ePre[j];
            unsigned s0 = len ? (__ldg(I.seedChunks + eOff[j]) >> 6) : 1u;
            if (s0 < st) st = s0;
        }
        for (int d = 16; d; d >>= 1) st = min(st, __shfl_xor_sync(DP_FULL, st, d));
        simWord = (int)st - 1;
        __syncwarp();
    }
    for (;;) {
        // the next word at which a live set is dropped: the smallest (last word + 1), but not before
        // simWord + 1 (sets that ended earlier are all dropped by the first pass)
        unsigned nxt = 0xffffffffu;
        for (int t = (int)lane; t < simLive; t += 32) nxt = min(nxt, eEndW[order[t]] + 1u);
        for (int d = 16; d; d >>= 1) nxt = min(nxt, __shfl_xor_sync(DP_FULL, nxt, d));
        if (nxt == 0xffffffffu) break;
        const int i = max((int)nxt, simWord + 1);
        if (i > (int)wword) break;
        if (lane == 0) {
            int t = 0, live = simLive;
            while (t < live) {
                if (eEndW[order[t]] + 1 <= (unsigned)i) {
                    order[t] = order[live - 1];
                    live--;
                } else {
                    t++;
                }
            }
            simLive = live;
        }
        simLive = __shfl_sync(DP_FULL, simLive, 0);
        simWord = i;
        __syncwarp();
    }
    if ((int)wword > simWord) simWord = (int)wword;
    __syncwarp();
    if (lane == 0) {
        X.sim[0] = simWord;
        X.sim[1] = simLive;
    }
    __syncwarp();
}

// Candidates over the count threshold, ascending by chunk id, as (chunk << 32 | soft count): applies the level clamp's
// early stop (Q11) and the level-16 under-count (Q6), computes the DISTINCT query seeds present in each survivor
// (IntSet.CountIntersectionTo's operand, mapping.go:520-523) and writes (chunk, distinct) pairs. Warp-collective;
// returns the number of survivors (which may exceed candStride: the caller flags that).
template <bool DUP_IN_COUNT>  // bits 16-31 of a candidate's count word = the repeated-seed runs containing it (no search)
__device__ __forceinline__ int dp_refine_emit(const DpIndexDev& I, const DpRefineCtx& X, const unsigned long long* sorted, int nCand,
                              unsigned* outChunk, unsigned short* outDist, int candStride) {
    const unsigned lane = dp_lane();
    const int nInc = X.nInc, minCount = X.minCount, T = X.T;
    const bool clamped = X.clamped, q6 = X.q6;
    const unsigned* eOff = X.eOff;
    const unsigned* ePre = X.ePre;
    const unsigned* eEndW = X.eEndW;
    unsigned short* order = X.order;
    int nCandOut = 0;
    bool simInit = false;
    for (int r0 = 0; r0 < nCand; r0 += 32) {
        const bool have = r0 + (int)lane < nCand;
        const unsigned long long mine = have ? sorted[r0 + lane] : 0ull;
        unsigned c = (unsigned)(mine >> 32);
        unsigned v = (unsigned)mine;
        int soft = (int)(v & 0xffffu);
        unsigned mp = __ballot_sync(DP_FULL, have);
        if (mp && (clamped || q6)) {
            unsigned todo = mp;
            while (todo) {
                int l = __ffs(todo) - 1;
                todo &= todo - 1;
                unsigned cc = __shfl_sync(DP_FULL, c, l);
                unsigned wword = cc >> 6;
                int sft = __shfl_sync(DP_FULL, soft, l);
                bool keep = true;
                if (clamped) {
                    int live = 0;
                    for (int j0 = 0; j0 < nInc; j0 += 32) {
                        int j = j0 + (int)lane;
                        live += __popc(__ballot_sync(DP_FULL, j < nInc && eEndW[j] >= wword));
                    }
                    if (live < minCount) keep = false;
                }
                if (keep && q6 && sft == T) {
                    dp_q6_advance(I, X, wword, simInit);
                    simInit = true;
                    __syncwarp();
                    bool in = false;
                    if (lane < 8) {
                        unsigned j = order[lane];
                        in = dp_run_contains(I.seedChunks, eOff[j], ePre[j + 1] - ePre[j], cc);
                    }
                    unsigned mi = __ballot_sync(DP_FULL, in);
                    if ((mi & 0x80u) && !(mi & 0x7fu)) keep = false;  // count-1 < T
                }
                if (!keep) mp &= ~(1u << l);
            }
        }
        // distinct query seeds present in each survivor = its soft count (one per included run containing it) minus
        // the runs that repeat an earlier run's seed and contain it
        unsigned todo2 = mp;
        while (todo2) {
            int l = __ffs(todo2) - 1;
            todo2 &= todo2 - 1;
            unsigned cc = __shfl_sync(DP_FULL, c, l);
            int distinct;
            if (DUP_IN_COUNT) {
                distinct = __shfl_sync(DP_FULL, soft - (int)(v >> 16), l);
            } else if (X.nDup >= 0) {
                int rep = 0;
                for (int d0 = 0; d0 < X.nDup; d0 += 32) {
                    int d = d0 + (int)lane;
                    bool in = false;
                    if (d < X.nDup) {
                        int j = X.dup[d];
                        in = dp_run_contains(I.seedChunks, eOff[j], ePre[j + 1] - ePre[j], cc);
                    }
                    rep += __popc(__ballot_sync(DP_FULL, in));
                }
                distinct = __shfl_sync(DP_FULL, soft, l) - rep;
            } else {
                distinct = 0;
                for (int j0 = 0; j0 < nInc; j0 += 32) {
                    int j = j0 + (int)lane;
                    bool in = false;
                    if (j < nInc && X.eFirst[j]) in = dp_run_contains(I.seedChunks, eOff[j], ePre[j + 1] - ePre[j], cc);
                    distinct += __popc(__ballot_sync(DP_FULL, in));
                }
            }
            int idx = nCandOut + __popc(mp & ((1u << l) - 1));
            if (lane == 0 && idx < candStride) {
                outChunk[idx] = cc;
                outDist[idx] = (unsigned short)(distinct + X.nAllDistinct);
            }
        }
        nCandOut += __popc(mp);
    }
    return nCandOut;
}


__global__ void __launch_bounds__(32 * DP_LWARPS, 8) dp_lookup_kernel(DpIndexDev I, DpExtractOut Q, int nWS,
                                                                  const int* __restrict__ wsList,
                                                                  const int* __restrict__ nWsList,
                                                                  DpLookupScratch S, int countersInSmem,
                                                                  int* __restrict__ candN,
                                                                  unsigned* __restrict__ candChunk,
                                                                  unsigned short* __restrict__ candDistinct,
                                                                  int candStride, DpCounters* __restrict__ ctr) {
    extern __shared__ unsigned dp_smem[];  // per-warp chunk counters when they fit
    __shared__ unsigned shSeed[DP_LWARPS][DP_ECAP];
    __shared__ unsigned shOff[DP_LWARPS][DP_ECAP];
    __shared__ unsigned shPre[DP_LWARPS][DP_ECAP + 1];
    __shared__ unsigned shEndW[DP_LWARPS][DP_ECAP];
    __shared__ unsigned char shFirst[DP_LWARPS][DP_ECAP];
    __shared__ unsigned shTouched[DP_LWARPS][DP_TCAP];
    __shared__ unsigned long long shCand[DP_LWARPS][DP_TCAP];
    __shared__ unsigned short shDup[DP_LWARPS][DP_DUPCAP];
    __shared__ int shSim[DP_LWARPS][2];
    const unsigned lane = dp_lane();
    const unsigned lt = dp_lanemask_lt();
    const int wib = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned C = I.numChunks;
    // per-chunk soft counters, 16 bits each, two per word; zero between window strands (cleared once per kernel)
    const unsigned cWords = (C + 1) >> 1;
    unsigned* cnt = countersInSmem ? dp_smem + (size_t)wib * cWords : S.counters + (size_t)gwarp * cWords;
    if (countersInSmem) {
        for (unsigned c = lane; c < cWords; c += 32) cnt[c] = 0;
        __syncwarp();
    }
    const size_t so = (size_t)gwarp * S.stride;
    unsigned* allSeeds = S.allSeeds + so;
    unsigned short* order = S.order + so;
    unsigned long long cRuns = 0, cEntries = 0, cCand = 0;

    const int nTodo = wsList ? min(*nWsList, nWS) : nWS;  // a list: the window strands the block kernel deferred
    for (int wi = gwarp; wi < nTodo; wi += nWarps) {
        const int ws = wsList ? wsList[wi] : wi;
        const int n = Q.wsN[ws];
        const unsigned qb = Q.wsOff[ws];
        int nCandOut = 0;
        unsigned* outChunk = candChunk + (size_t)ws * candStride;
        unsigned short* outDist = candDistinct + (size_t)ws * candStride;
        if (n >= 5) {
            const bool eSmall = n <= DP_ECAP;
            unsigned* eSeed = eSmall ? shSeed[wib] : S.eSeed + so;
            unsigned* eOff = eSmall ? shOff[wib] : S.eOff + so;
            unsigned* ePre = eSmall ? shPre[wib] : S.ePre + (size_t)gwarp * (S.stride + 1);
            unsigned* eEndW = eSmall ? shEndW[wib] : S.eEndW + so;
            unsigned char* eFirst = eSmall ? shFirst[wib] : S.eFirst + so;
            // ---- inclusion filter (seeds.go:340-346), ordered ----
            int nInc = 0, nAll = 0;
            unsigned prevElig = 0xffffffffu;
            for (int j0 = 0; j0 < n; j0 += 32) {
                int j = j0 + (int)lane;
                bool valid = j < n;
                unsigned s = 0xffffffffu, o = 0, c = 0;
                if (valid) {
                    s = Q.qSeed[qb + j];
                    o = __ldg(I.seedOff + s);
                    c = __ldg(I.seedOff + s + 1) - o;
                }
                bool elig = valid && c < C;
                unsigned me = __ballot_sync(DP_FULL, elig);
                unsigned lower = me & lt;
                int src = lower ? 31 - __clz(lower) : 0;
                unsigned ps = __shfl_sync(DP_FULL, s, src);
                if (!lower) ps = prevElig;
                bool inc = elig && s != ps;
                unsigned mi = __ballot_sync(DP_FULL, inc);
                if (inc) {
                    int idx = nInc + __popc(mi & lt);
                    eSeed[idx] = s;
                    eOff[idx] = o;
                    ePre[idx] = c;  // run length for now; prefix-summed below
                }
                nInc += __popc(mi);
                if (me) prevElig = __shfl_sync(DP_FULL, s, 31 - __clz(me));
                unsigned ma = __ballot_sync(DP_FULL, valid && c >= C);
                if (valid && c >= C) allSeeds[nAll + __popc(ma & lt)] = s;
                nAll += __popc(ma);
            }
            __syncwarp();
            if (nInc >= 5) {
                const int minCount = (nInc + 2) >> 2;
                int T;
                bool clamped = false;
                if (minCount >= 9 && minCount <= 12) {
                    T = 8;
                    clamped = true;
                } else if (minCount >= 17 && minCount <= 24) {
                    T = 16;
                    clamped = true;
                } else {
                    T = minCount;
                }
                const bool q6 = minCount >= 13 && minCount <= 24;  // level-16 plane decides alone
                // ---- distinct seeds present in every chunk (tiny references only) ----
                int nAllDistinct = 0;
                for (int a0 = 0; a0 < nAll; a0 += 32) {
                    int a = a0 + (int)lane;
                    bool first = false;
                    if (a < nAll) {
                        first = true;
                        unsigned s = allSeeds[a];
                        for (int b = 0; b < a; b++)
                            if (allSeeds[b] == s) {
                                first = false;
                                break;
                            }
                    }
                    nAllDistinct += __popc(__ballot_sync(DP_FULL, first));
                }
                // ---- first occurrences among E (__match_any_sync), last word of each run, run-length prefix ----
                unsigned total = 0;
                int nDup = 0;
                for (int j0 = 0; j0 < nInc; j0 += 32) {
                    int j = j0 + (int)lane;
                    unsigned c = 0;
                    unsigned s = j < nInc ? eSeed[j] : (0x80000000u | lane);
                    unsigned mm = __match_any_sync(DP_FULL, s);
                    bool isDup = false;
                    if (j < nInc) {
                        c = ePre[j];
                        bool first = (__ffs(mm) - 1) == (int)lane;
                        if (first && j0 > 0) {
                            for (int b = 0; b < j0; b++)
                                if (eSeed[b] == s) {
                                    first = false;
                                    break;
                                }
                        }
                        eFirst[j] = first ? 1 : 0;
                        isDup = !first;
                        if (clamped || q6) eEndW[j] = c ? (__ldg(I.seedChunks + eOff[j] + c - 1) >> 6) : 0u;
                    }
                    {   // list of the runs that repeat an earlier run's seed (few): what the distinct count subtracts
                        unsigned md = __ballot_sync(DP_FULL, isDup);
                        if (nDup >= 0) {
                            if (nDup + __popc(md) > DP_DUPCAP) nDup = -1;
                            else {
                                if (isDup) shDup[wib][nDup + __popc(md & lt)] = (unsigned short)j;
                                nDup += __popc(md);
                            }
                        }
                    }
                    unsigned x = c;  // warp inclusive scan
                    for (int d = 1; d < 32; d <<= 1) {
                        unsigned y = __shfl_up_sync(DP_FULL, x, d);
                        if ((int)lane >= d) x += y;
                    }
                    __syncwarp();
                    if (j < nInc) ePre[j] = total + x - c;
                    total += __shfl_sync(DP_FULL, x, 31);
                }
                if (lane == 0) ePre[nInc] = total;
                __syncwarp();
                // ---- gather the posting runs into the per-chunk counters; remember which counters left zero ----
                const bool tSmall = total <= DP_TCAP;
                unsigned* touched = tSmall ? shTouched[wib] : S.touched + (size_t)gwarp * S.tStride;
                unsigned long long* cand = tSmall ? shCand[wib] : S.cand + (size_t)gwarp * 2 * S.tStride;
                int nTouched = 0;
                if (total >= 12u * (unsigned)nInc) {
                    // long runs: one run at a time, lanes across its postings (coalesced, no search)
                    for (int j = 0; j < nInc; j++) {
                        const unsigned off = eOff[j], len = ePre[j + 1] - ePre[j];
                        for (unsigned p0 = 0; p0 < len; p0 += 32) {
                            unsigned p = p0 + lane;
                            bool owner = false;
                            unsigned chunk = 0;
                            if (p < len) {
                                chunk = __ldg(I.seedChunks + off + p);
                                unsigned sh = (chunk & 1u) * 16u;
                                unsigned old = atomicAdd(cnt + (chunk >> 1), 1u << sh);
                                owner = ((old >> sh) & 0xffffu) == 0;
                            }
                            unsigned mo = __ballot_sync(DP_FULL, owner);
                            if (owner) touched[nTouched + __popc(mo & lt)] = chunk;
                            nTouched += __popc(mo);
                        }
                    }
                } else {
                    // short runs: flatten all postings over the lanes, each finds its run by binary search
                    for (unsigned p0 = 0; p0 < total; p0 += 32) {
                        unsigned p = p0 + lane;
                        bool owner = false;
                        unsigned chunk = 0;
                        if (p < total) {
                            int lo = 0, hi = nInc;  // largest j with ePre[j] <= p
                            while (hi - lo > 1) {
                                int mid = (lo + hi) >> 1;
                                if (ePre[mid] <= p) lo = mid;
                                else hi = mid;
                            }
                            chunk = __ldg(I.seedChunks + eOff[lo] + (p - ePre[lo]));
                            unsigned sh = (chunk & 1u) * 16u;
                            unsigned old = atomicAdd(cnt + (chunk >> 1), 1u << sh);
                            owner = ((old >> sh) & 0xffffu) == 0;
                        }
                        unsigned mo = __ballot_sync(DP_FULL, owner);
                        if (owner) touched[nTouched + __popc(mo & lt)] = chunk;
                        nTouched += __popc(mo);
                    }
                }
                __syncwarp();
                cRuns += (unsigned)nInc;
                cEntries += total;
                // ---- threshold on the touched chunks; counters go back to zero ----
                int nCand = 0;
                for (int t0 = 0; t0 < nTouched; t0 += 32) {
                    int t = t0 + (int)lane;
                    unsigned chunk = 0, v = 0;
                    if (t < nTouched) {
                        chunk = touched[t];
                        unsigned sh = (chunk & 1u) * 16u;
                        v = (cnt[chunk >> 1] >> sh) & 0xffffu;
                    }
                    __syncwarp();  // two lanes may own the two halves of one word: read everything, then clear
                    if (t < nTouched) atomicAnd(cnt + (chunk >> 1), ~(0xffffu << ((chunk & 1u) * 16u)));
                    bool pass = (int)v >= T;
                    unsigned mp = __ballot_sync(DP_FULL, pass);
                    if (pass) cand[nCand + __popc(mp & lt)] = ((unsigned long long)chunk << 32) | v;
                    nCand += __popc(mp);
                }
                __syncwarp();
                // ---- ascending chunk id (rank sort: chunk ids are distinct), then the rare refinements ----
                const unsigned long long* sorted = cand;
                if (nCand > 1) {
                    if (nCand <= 32) {  // in place: every lane holds its element before anyone writes
                        unsigned long long e = (int)lane < nCand ? cand[lane] : ~0ull;
                        int rank = 0;
                        for (int y = 0; y < nCand; y++) rank += (cand[y] >> 32) < (e >> 32);
                        __syncwarp();
                        if ((int)lane < nCand) cand[rank] = e;
                    } else {
                        unsigned long long* dst = S.cand + (size_t)gwarp * 2 * S.tStride + S.tStride;
                        for (int x0 = 0; x0 < nCand; x0 += 32) {
                            int x = x0 + (int)lane;
                            if (x < nCand) {
                                unsigned long long e = cand[x];
                                int rank = 0;
                                for (int y = 0; y < nCand; y++) rank += (cand[y] >> 32) < (e >> 32);
                                dst[rank] = e;
                            }
                        }
                        sorted = dst;
                    }
                    __syncwarp();
                }
                DpRefineCtx X;
                X.eOff = eOff;
                X.ePre = ePre;
                X.eEndW = eEndW;
                X.eFirst = eFirst;
                X.dup = shDup[wib];
                X.nDup = nDup;
                X.order = order;
                X.sim = shSim[wib];
                X.nInc = nInc;
                X.minCount = minCount;
                X.T = T;
                X.clamped = clamped;
                X.q6 = q6;
                X.nAllDistinct = nAllDistinct;
                nCandOut = dp_refine_emit<false>(I, X, sorted, nCand, outChunk, outDist, candStride);
                if (nCandOut > candStride) {
                    if (lane == 0) atomicOr(&ctr->overflow, DP_OV_CANDS);
                    nCandOut = candStride;
                }
            }
        }
        if (lane == 0) candN[ws] = nCandOut;
        cCand += (unsigned)nCandOut;
        __syncwarp();
    }
    if (lane == 0 && (cRuns | cCand)) {
        atomicAdd(&ctr->posting_runs, cRuns);
        atomicAdd(&ctr->posting_entries, cEntries);
        atomicAdd(&ctr->candidates, cCand);
    }
}

// ===============================================================================================================
// Stage 2, common case of small references (dp_lookup_small_kernel): a window strand with at most 32 seeds against an
// index of at most DP_SMALL_CHUNKS chunks. Everything dp_lookup_kernel keeps in shared-memory lists lives in registers
// here (one included run per lane), and the generality is gone:
//   * n <= 32 means minCount = (nInc + 2) / 4 <= 8, so the threshold is exact: no level clamp (Q11), no level-16
//     under-count (Q6), no early stop;
//   * one 32-bit counter per chunk holds the soft count (every included run) in its lower and the DISTINCT count
//     (first occurrences of a seed only) in its upper half-word, so the distinct count needs no search;
//   * a chunk becomes a candidate when its soft count reaches T (counts only grow: once); the table is blanked
//     wholesale afterwards.
// Window strands outside the common case (more than 32 seeds, a seed present in every chunk, more than 32 candidates)
// are handed to dp_lookup_kernel through a list, with identical results.
// ===============================================================================================================
#define DP_SMALL_CHUNKS 2048
#define DP_SMALL_WARPS 4

__global__ void __launch_bounds__(32 * DP_SMALL_WARPS, 12) dp_lookup_small_kernel(DpIndexDev I, DpExtractOut Q, int nWS,
                                                                                int* __restrict__ deferList,
                                                                                int* __restrict__ nDefer,
                                                                                int* __restrict__ candN,
                                                                                unsigned* __restrict__ candChunk,
                                                                                unsigned short* __restrict__ candDistinct,
                                                                                int candStride, DpCounters* __restrict__ ctr) {
    extern __shared__ unsigned dp_smem[];  // DP_SMALL_WARPS x C counters
    const unsigned lane = dp_lane();
    const unsigned lt = dp_lanemask_lt();
    const int wib = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned C = I.numChunks;
    unsigned* cnt = dp_smem + (size_t)wib * C;
    for (unsigned c = lane; c < C; c += 32) cnt[c] = 0;
    __syncwarp();
    unsigned long long cRuns = 0, cEntries = 0, cCand = 0;
    for (int ws = gwarp; ws < nWS; ws += nWarps) {
        const int n = Q.wsN[ws];
        int nOut = 0;
        bool defer = n > 32;
        if (n >= 5 && !defer) {
            const unsigned qb = Q.wsOff[ws];
            const bool valid = (int)lane < n;
            unsigned s = 0xffffffffu, o = 0, c = 0;
            if (valid) {
                s = Q.qSeed[qb + lane];
                o = __ldg(I.seedOff + s);
                c = __ldg(I.seedOff + s + 1) - o;
            }
            defer = __any_sync(DP_FULL, valid && c >= C);  // a seed present in every chunk: the general kernel's business
            if (!defer) {
                // ---- inclusion filter (seeds.go:340-346): every occurrence is eligible here, repeats of the
                //      previous occurrence's seed are skipped ----
                const unsigned sPrev = __shfl_up_sync(DP_FULL, s, 1);
                const bool inc = valid && (lane == 0 || s != sPrev);
                const unsigned mi = __ballot_sync(DP_FULL, inc);
                const int nInc = __popc(mi);
                if (nInc >= 5) {
                    const int T = (nInc + 2) >> 2;  // <= 8: exact level
                    // compact the included runs: lane t holds the t-th
                    const int src = (int)lane < nInc ? (int)__fns(mi, 0, (int)lane + 1) : 0;
                    const unsigned es = __shfl_sync(DP_FULL, s, src);
                    const unsigned eo = __shfl_sync(DP_FULL, o, src);
                    const unsigned ecAll = __shfl_sync(DP_FULL, c, src);
                    const unsigned ec = (int)lane < nInc ? ecAll : 0u;
                    // first occurrence of the seed among the included runs
                    const unsigned mm = __match_any_sync(DP_FULL, (int)lane < nInc ? es : (0x80000000u | lane));
                    const unsigned add = (__ffs(mm) - 1) == (int)lane ? 0x10001u : 1u;
                    // exclusive prefix of the run lengths
                    unsigned x = ec;
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned y = __shfl_up_sync(DP_FULL, x, d);
                        if ((int)lane >= d) x += y;
                    }
                    const unsigned pre = x - ec;
                    const unsigned total = __shfl_sync(DP_FULL, x, 31);
                    // ---- flattened gather: posting p belongs to the run r with pre_r <= p < pre_r + ec_r ----
                    unsigned candChunkReg = 0;  // lane t: the t-th chunk whose soft count reached T
                    int nCand = 0;
                    for (unsigned p0 = 0; p0 < total && nCand <= 32; p0 += 32) {
                        const unsigned p = p0 + lane;
                        int lo = 0, hi = nInc;
                        for (int step = 0; step < 5; step++) {  // largest r with pre_r <= p (nInc <= 32)
                            const int mid = (lo + hi) >> 1;
                            const unsigned v = __shfl_sync(DP_FULL, pre, mid);
                            if (hi - lo > 1) {
                                if (v <= p) lo = mid;
                                else hi = mid;
                            }
                        }
                        const unsigned ro = __shfl_sync(DP_FULL, eo, lo);
                        const unsigned rp = __shfl_sync(DP_FULL, pre, lo);
                        const unsigned ra = __shfl_sync(DP_FULL, add, lo);
                        bool reached = false;
                        unsigned chunk = 0;
                        if (p < total) {
                            chunk = __ldg(I.seedChunks + ro + (p - rp));
                            const unsigned old = atomicAdd(cnt + chunk, ra);
                            reached = (int)(old & 0xffffu) + 1 == T;
                        }
                        const unsigned mr = __ballot_sync(DP_FULL, reached);
                        if (mr) {
                            // append this round's chunks to the candidate registers (lane nCand + rank)
                            const int base = nCand;
                            nCand += __popc(mr);
                            if (nCand <= 32) {
                                const int want = (int)lane - base;  // which of this round's chunks lands on this lane
                                const int from = want >= 0 && want < __popc(mr) ? (int)__fns(mr, 0, want + 1) : 0;
                                const unsigned got = __shfl_sync(DP_FULL, chunk, from);
                                if (want >= 0 && want < __popc(mr)) candChunkReg = got;
                            }
                        }
                    }
                    __syncwarp();
                    if (nCand > 32) {
                        defer = true;  // (the general kernel lists candidates in memory)
                    } else {
                        cRuns += (unsigned)nInc;
                        cEntries += total;
                        // final counts, ascending chunk id, output
                        const unsigned w = (int)lane < nCand ? cnt[candChunkReg] : 0u;
                        int rank = 0;
                        for (int y = 0; y < nCand; y++) rank += __shfl_sync(DP_FULL, candChunkReg, y) < candChunkReg;
                        if ((int)lane < nCand && rank < candStride) {
                            candChunk[(size_t)ws * candStride + rank] = candChunkReg;
                            candDistinct[(size_t)ws * candStride + rank] = (unsigned short)(w >> 16);
                        }
                        nOut = nCand;
                        if (nOut > candStride) {
                            if (lane == 0) atomicOr(&ctr->overflow, DP_OV_CANDS);
                            nOut = candStride;
                        }
                    }
                    __syncwarp();
                    if (total) for (unsigned c2 = lane; c2 < C; c2 += 32) cnt[c2] = 0;
                    __syncwarp();
                }
            }
        }
        if (lane == 0) {
            if (defer) deferList[atomicAdd(nDefer, 1)] = ws;
            else candN[ws] = nOut;
        }
        if (!defer) cCand += (unsigned)nOut;
    }
    if (lane == 0 && (cRuns | cCand)) {
        atomicAdd(&ctr->posting_runs, cRuns);
        atomicAdd(&ctr->posting_entries, cEntries);
        atomicAdd(&ctr->candidates, cCand);
    }
}

// ===============================================================================================================
// Stage 2 for indexes with many chunks — one CTA per window strand (dp_lookup_block_kernel).
//
// With C chunks a warp-private set of counters costs 2C bytes; beyond a few thousand chunks that either leaves the SM
// nearly empty or spills the counters to global memory, where every posting becomes a random read-modify-write.
// Here a CTA of 256 threads owns ONE set of 32-bit counters in shared memory and all its warps stream the posting runs
// of one window strand into it. To keep several CTAs resident per SM at human-genome scale (10^5 chunks and more) the
// counters are COARSE: counter g counts the postings of the 2^gShift adjacent chunks g*2^gShift ... (a group's count
// is >= the count of each of its chunks, so every chunk over the threshold lies in a group over the threshold).
//
//   1. inclusion filter, thread per query seed, block-wide ordered compaction                 (seeds.go:340-346)
//   2. the runs are cut into items of 128 postings (32 for indexes with short runs) aligned to 16 bytes; a warp takes
//      NI = 6 warp-wide loads' worth of items at once, issues its six 16-byte loads per lane (3 KB per warp in flight),
//      then adds into the group counters with shared-memory reductions (no return value, no per-posting test or
//      branch — a posting outside its item's range goes to the lane's dummy counter: ~6 instructions per posting)
//   3. one pass over the counters finds the groups that reached the threshold and blanks the counters (16-byte
//      loads/stores); with gShift = 0 these are the candidates and their exact counts
//   4. gShift > 0: for the few groups over the threshold (the true locus; random groups stay far below it) every run
//      binary-searches the group's first chunk and adds its <= 2^gShift postings to an exact per-chunk table
//   5. candidates sorted by chunk id -> dp_refine_emit, as in the warp kernel
//
// At human-genome scale a window strand gathers 10^4-10^5 postings (hundreds of KB): this is the HBM-bound kernel of
// the path; its traffic is sequential inside each run and every posting is read once.
// Window strands that contain a seed present in EVERY chunk (tiny references only) are deferred to dp_lookup_kernel
// through a list.
// ===============================================================================================================
#define DP_BITEMS 512   // gather items listed in shared memory at a time
#define DP_BCAND 256    // candidates over the threshold held in shared memory
#define DP_BDUP 64      // repeated-seed runs listed per window strand
#define DP_BGLIST 256   // groups over the threshold listed in shared memory (more: global scratch)
#define DP_BEXACT 2048  // words of the exact recount table: DP_BEXACT >> gShift groups are recounted at a time

struct DpLookupBlockCfg {
    int gShift;      // a counter covers 2^gShift adjacent chunks
    int cntWords;    // 32-bit group counters (multiple of 4)
    int eCap;        // included runs held in shared memory (more: global scratch)
    int gListCap;    // groups listed in shared memory before the list spills (<= DP_BGLIST; tests shrink it)
    int gBatch;      // groups recounted at a time (<= exactWords >> gShift; tests shrink it)
    int exactWords;  // words of the exact recount table (<= DP_BEXACT)
    unsigned* work;  // dynamic work counter, zero at launch
    int* deferList;  // window strands left to dp_lookup_kernel
    int* nDefer;
};

struct DpBlockShared {
    unsigned wTot[32];
    unsigned wLast[32];
    int ws;
    int nCand;
    int nDup;
    int nCandOut;
    int nGroup;
    unsigned nextItem;
    int sim[2];
};

// exclusive prefix sum over the CTA (blockDim.x <= 1024); total returned in `total`
__device__ __forceinline__ unsigned dp_block_excl_scan(unsigned v, unsigned* wTot, unsigned& total) {
    const unsigned lane = dp_lane();
    const unsigned warp = threadIdx.x >> 5, nWarp = blockDim.x >> 5;
    unsigned x = v;
    for (int d = 1; d < 32; d <<= 1) {
        unsigned y = __shfl_up_sync(DP_FULL, x, d);
        if ((int)lane >= d) x += y;
    }
    if (lane == 31) wTot[warp] = x;
    __syncthreads();
    unsigned t = lane < nWarp ? wTot[lane] : 0u;
    for (int d = 1; d < 32; d <<= 1) {
        unsigned y = __shfl_up_sync(DP_FULL, t, d);
        if ((int)lane >= d) t += y;
    }
    const unsigned base = warp ? __shfl_sync(DP_FULL, t, warp - 1) : 0u;
    total = __shfl_sync(DP_FULL, t, nWarp - 1);
    __syncthreads();
    return base + x - v;
}

// 16 bytes of a posting run: streamed once, never reused by this SM (no L1 allocation)
__device__ __forceinline__ uint4 dp_load_postings(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// 16 bytes of an index table that should outlive the streams passing through L2 beside it
__device__ __forceinline__ unsigned long long dp_policy_evict_last() {
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 dp_load_index16(const uint4* p, unsigned long long pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}

// one posting into the group counters: shared-memory reduction (no return value); a posting outside the item's range
// goes to the lane's private dummy counter instead (no branch, no bank conflict among the dummies)
__device__ __forceinline__ void dp_group_count(unsigned cntAddr, unsigned dummyAddr, unsigned chunk, int gShift, bool valid) {
    const unsigned addr = valid ? cntAddr + ((chunk >> gShift) << 2) : dummyAddr;
    asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory");
}

// THREADS: CTA size; MINB: resident CTAs per SM the register budget is cut for; NI: 16-byte loads a lane has in flight;
// SEG: postings per gather item — 128 (a warp per item) for long runs, 32 (8 lanes per item, four items per warp-wide
// load) for indexes whose runs are a few dozen postings long, where 128-posting items would be mostly empty
template <int THREADS, int MINB, int NI, int SEG>
__global__ void __launch_bounds__(THREADS, MINB) dp_lookup_block_kernel(DpIndexDev I, DpExtractOut Q, int nWS,
                                                                         DpLookupScratch S, DpLookupBlockCfg G,
                                                                         int* __restrict__ candN,
                                                                         unsigned* __restrict__ candChunk,
                                                                         unsigned short* __restrict__ candDistinct,
                                                                         int candStride, DpCounters* __restrict__ ctr) {
    extern __shared__ unsigned dp_smem[];
    __shared__ DpBlockShared sh;
    __shared__ unsigned long long shCand[DP_BCAND];
    __shared__ unsigned short shDup[DP_BDUP];
    __shared__ unsigned shGroup[DP_BGLIST];
    const int tid = threadIdx.x, nT = blockDim.x;
    const unsigned lane = dp_lane();
    const int warp = tid >> 5, nWarp = nT >> 5;
    const unsigned C = I.numChunks;
    const int gShift = G.gShift;
    // dynamic shared memory: counters | 32 dummy counters | eSeed | eOff | ePre | eItem | eEndW | item starts |
    // item ranges | exact | eFirst
    unsigned* cnt = dp_smem;
    unsigned* smSeed = cnt + G.cntWords + 32;
    unsigned* smOff = smSeed + G.eCap;
    unsigned* smPre = smOff + G.eCap;
    unsigned* smItem = smPre + G.eCap + 1;
    unsigned* smEndW = smItem + G.eCap + 1;
    unsigned* smItemStart = smEndW + G.eCap;
    unsigned* smItemRange = smItemStart + DP_BITEMS;
    unsigned* exact = smItemRange + DP_BITEMS;
    unsigned char* smFirst = reinterpret_cast<unsigned char*>(exact + (gShift ? G.exactWords : 0));
    for (int i = tid; i < G.cntWords; i += nT) cnt[i] = 0;
    // global scratch of this CTA for oversized window strands
    const size_t so = (size_t)blockIdx.x * S.stride;
    unsigned short* order = S.order + so;
    unsigned long long* gCand = S.cand + (size_t)blockIdx.x * 2 * S.tStride;
    unsigned* gGroup = reinterpret_cast<unsigned*>(gCand + S.tStride);  // (the sort space: free until the sort)
    const uint4* chunks4 = reinterpret_cast<const uint4*>(I.seedChunks);
    const unsigned cntAddr = (unsigned)__cvta_generic_to_shared(cnt);
    const unsigned dummyAddr = cntAddr + ((unsigned)G.cntWords + lane) * 4u;  // 32 words after the counters, never read
    unsigned long long cRuns = 0, cEntries = 0, cCand = 0;  // thread 0 only
    unsigned nextWs = tid == 0 ? atomicAdd(G.work, 1u) : 0u;  // (the next one is fetched while this one is processed)
    __syncthreads();

    for (;;) {
        if (tid == 0) {
            sh.ws = (int)min(nextWs, 0x7fffffffu);
            sh.nCand = 0;
            sh.nDup = 0;
            sh.nCandOut = 0;
            sh.nGroup = 0;
        }
        __syncthreads();
        const int ws = sh.ws;
        if (ws >= nWS) break;
        if (tid == 0) nextWs = atomicAdd(G.work, 1u);
        const int n = Q.wsN[ws];
        const unsigned qb = Q.wsOff[ws];
        bool defer = false;
        if (n >= 5) {
            const bool eSmall = n <= G.eCap;
            unsigned* eSeed = eSmall ? smSeed : S.eSeed + so;
            unsigned* eOff = eSmall ? smOff : S.eOff + so;
            unsigned* ePre = eSmall ? smPre : S.ePre + (size_t)blockIdx.x * (S.stride + 1);
            unsigned* eItem = eSmall ? smItem : S.touched + (size_t)blockIdx.x * S.tStride;
            unsigned* eEndW = eSmall ? smEndW : S.eEndW + so;
            unsigned char* eFirst = eSmall ? smFirst : S.eFirst + so;
            // ---- inclusion filter (seeds.go:340-346): thread per seed, ordered compaction over the CTA ----
            int nInc = 0;
            unsigned carry = 0xffffffffu;  // seed of the last eligible occurrence of earlier rounds
            int sawAll = 0;
            for (int j0 = 0; j0 < n; j0 += nT) {
                const int j = j0 + tid;
                const bool valid = j < n;
                unsigned s = 0xffffffffu, o = 0, c = 0;
                if (valid) {
                    s = Q.qSeed[qb + j];
                    o = __ldg(I.seedOff + s);
                    c = __ldg(I.seedOff + s + 1) - o;
                }
                const bool elig = valid && c < C;
                if (valid && c >= C) sawAll = 1;
                const unsigned me = __ballot_sync(DP_FULL, elig);
                const unsigned lastSeed = __shfl_sync(DP_FULL, s, me ? 31 - __clz(me) : 0);
                if (lane == 0) sh.wLast[warp] = me ? lastSeed : 0xffffffffu;
                __syncthreads();
                // seed of the nearest eligible occurrence before this warp
                unsigned pv = (int)lane < warp ? sh.wLast[lane] : 0xffffffffu;
                unsigned mv = __ballot_sync(DP_FULL, pv != 0xffffffffu);
                unsigned warpCarry = __shfl_sync(DP_FULL, pv, mv ? 31 - __clz(mv) : 0);
                if (!mv) warpCarry = carry;
                const unsigned lower = me & dp_lanemask_lt();
                unsigned ps = __shfl_sync(DP_FULL, s, lower ? 31 - __clz(lower) : 0);
                if (!lower) ps = warpCarry;
                const bool inc = elig && s != ps;
                // carry for the next round: the last eligible occurrence of this round
                unsigned av = (int)lane < nWarp ? sh.wLast[lane] : 0xffffffffu;
                unsigned ma = __ballot_sync(DP_FULL, av != 0xffffffffu);
                if (ma) carry = __shfl_sync(DP_FULL, av, 31 - __clz(ma));
                unsigned roundTot;
                const unsigned base = dp_block_excl_scan(inc ? 1u : 0u, sh.wTot, roundTot);  // two barriers inside
                if (inc) {
                    const unsigned idx = (unsigned)nInc + base;
                    eSeed[idx] = s;
                    eOff[idx] = o;
                    ePre[idx] = c;  // run length for now; prefix-summed below
                }
                nInc += (int)roundTot;
            }
            defer = __syncthreads_or(sawAll) != 0;
            if (!defer && nInc >= 5) {
                const int minCount = (nInc + 2) >> 2;
                int T;
                bool clamped = false;
                if (minCount >= 9 && minCount <= 12) {
                    T = 8;
                    clamped = true;
                } else if (minCount >= 17 && minCount <= 24) {
                    T = 16;
                    clamped = true;
                } else {
                    T = minCount;
                }
                const bool q6 = minCount >= 13 && minCount <= 24;  // level-16 plane decides alone
                // ---- run-length prefix, item prefix, last word of each run ----
                // items of run [off, off+c): the 128-posting pieces of [off & ~3, off+c) (16-byte aligned starts)
                unsigned total = 0, nItems = 0;
                for (int j0 = 0; j0 < nInc; j0 += nT) {
                    const int j = j0 + tid;
                    unsigned c = 0, items = 0;
                    if (j < nInc) {
                        c = ePre[j];
                        const unsigned off = eOff[j];
                        if (c) items = (off + c - (off & ~3u) + SEG - 1) / SEG;
                        if (clamped || q6) eEndW[j] = c ? (__ldg(I.seedChunks + off + c - 1) >> 6) : 0u;
                    }
                    unsigned tA, tB;
                    const unsigned pa = dp_block_excl_scan(c, sh.wTot, tA);
                    const unsigned pb = dp_block_excl_scan(items, sh.wTot, tB);
                    if (j < nInc) {
                        ePre[j] = total + pa;
                        eItem[j] = nItems + pb;
                    }
                    total += tA;
                    nItems += tB;
                }
                if (tid == 0) {
                    ePre[nInc] = total;
                    eItem[nInc] = nItems;
                }
                __syncthreads();
                // ---- stream the runs into the group counters ----
                for (unsigned d0 = 0; d0 < nItems; d0 += DP_BITEMS) {
                    const unsigned dN = min((unsigned)DP_BITEMS, nItems - d0);
                    if (tid == 0) sh.nextItem = 0;
                    for (int j = tid; j < nInc; j += nT) {
                        const unsigned first = eItem[j], last = eItem[j + 1];  // this run's items
                        if (last <= d0 || first >= d0 + dN) continue;
                        const unsigned off = eOff[j], end = off + (ePre[j + 1] - ePre[j]);
                        const unsigned a = off & ~3u;
                        for (unsigned it = max(first, d0); it < min(last, d0 + dN); it++) {
                            const unsigned st = a + (it - first) * SEG;
                            const unsigned lo = max(off, st) - st, hi = min(end, st + SEG) - st;
                            smItemStart[it - d0] = st >> 2;
                            smItemRange[it - d0] = lo | (hi << 8);
                        }
                    }
                    __syncthreads();
                    constexpr unsigned LPI = SEG / 4;   // lanes per item (16 bytes = 4 postings per lane)
                    constexpr unsigned IPR = 32 / LPI;  // items per warp-wide load
                    const unsigned sub = lane / LPI, ln = lane % LPI;
                    for (;;) {
                        unsigned it = 0;
                        if (lane == 0) it = atomicAdd(&sh.nextItem, (unsigned)NI * IPR);
                        it = __shfl_sync(DP_FULL, it, 0);
                        if (it >= dN) break;
                        uint4 v[NI];
                        unsigned rg[NI];
#pragma unroll
                        for (int d = 0; d < NI; d++) {
                            const unsigned idx = it + (unsigned)d * IPR + sub;
                            const bool have = idx < dN;
                            rg[d] = have ? smItemRange[idx] : 0u;
                            const unsigned st4 = have ? smItemStart[idx] : 0u;
                            v[d] = make_uint4(0, 0, 0, 0);
                            if (4u * ln < (rg[d] >> 8)) v[d] = dp_load_postings(chunks4 + st4 + ln);
                        }
#pragma unroll
                        for (int d = 0; d < NI; d++) {
                            if (IPR == 1 && !rg[d]) continue;  // (warp-uniform: fewer than NI items were left)
                            // posting 4*ln+e of the item counts iff lo <= 4*ln+e < hi (unsigned wrap-around compare)
                            const unsigned lo = rg[d] & 0xffu, span = (rg[d] >> 8) - lo;
                            const unsigned bl = 4u * ln - lo;
                            dp_group_count(cntAddr, dummyAddr, v[d].x, gShift, bl < span);
                            dp_group_count(cntAddr, dummyAddr, v[d].y, gShift, bl + 1u < span);
                            dp_group_count(cntAddr, dummyAddr, v[d].z, gShift, bl + 2u < span);
                            dp_group_count(cntAddr, dummyAddr, v[d].w, gShift, bl + 3u < span);
                        }
                    }
                    __syncthreads();
                }
                // ---- counters over the threshold; blank the counters ----
                {
                    uint4* cnt4 = reinterpret_cast<uint4*>(cnt);
                    for (int w4 = tid; w4 < (G.cntWords >> 2); w4 += nT) {
                        const uint4 x = cnt4[w4];
                        cnt4[w4] = make_uint4(0, 0, 0, 0);
                        const unsigned xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            if ((int)xs[i] >= T) {
                                const unsigned g = (unsigned)w4 * 4u + (unsigned)i;
                                if (gShift == 0) {
                                    const int slot = atomicAdd(&sh.nCand, 1);
                                    const unsigned long long e = ((unsigned long long)g << 32) | xs[i];
                                    if (slot < DP_BCAND) shCand[slot] = e;
                                    else if (slot < S.tStride) gCand[slot] = e;
                                } else {
                                    const int slot = atomicAdd(&sh.nGroup, 1);
                                    if (slot < G.gListCap) shGroup[slot] = g;
                                    else gGroup[slot] = g;
                                }
                            }
                        }
                    }
                    __syncthreads();
                }
                // ---- runs that repeat an earlier run's seed (for the distinct counts) ----
                if (gShift ? sh.nGroup > 0 : sh.nCand > 0) {
                    // within a warp's 32 runs by __match_any_sync; against the earlier runs four seeds per load
                    const uint4* eSeed4 = reinterpret_cast<const uint4*>(eSeed);
                    for (int j0 = warp * 32; j0 < nInc; j0 += nT) {  // (warp-uniform bounds)
                        const int j = j0 + (int)lane;
                        const unsigned s = j < nInc ? eSeed[j] : (0x80000000u | lane);  // (seed ranks are below 2^31)
                        const unsigned mm = __match_any_sync(DP_FULL, s);
                        bool first = (__ffs(mm) - 1) == (int)lane;
                        if (j < nInc && first)
                            for (int b4 = 0; b4 < (j0 >> 2) && first; b4++) {
                                const uint4 q = eSeed4[b4];
                                first = !(q.x == s || q.y == s || q.z == s || q.w == s);
                            }
                        if (j < nInc) {
                            eFirst[j] = first ? 1 : 0;
                            if (!first) {
                                const int slot = atomicAdd(&sh.nDup, 1);
                                if (slot < DP_BDUP) shDup[slot] = (unsigned short)j;
                            }
                        }
                    }
                    __syncthreads();
                }
                // ---- exact per-chunk counts of the groups over the threshold ----
                if (gShift) {
                    const int nG = sh.nGroup;
                    const unsigned gSize = 1u << gShift;
                    for (int b0 = 0; b0 < nG; b0 += G.gBatch) {
                        const int bn = min(G.gBatch, nG - b0);
                        for (int x = tid; x < (bn << gShift); x += nT) exact[x] = 0;
                        __syncthreads();
                        for (int idx = tid; idx < bn * nInc; idx += nT) {
                            const int gi = idx / nInc, j = idx - gi * nInc;
                            const int gx = b0 + gi;
                            const unsigned g = gx < G.gListCap ? shGroup[gx] : gGroup[gx];
                            const unsigned cLo = g << gShift, cHi = cLo + gSize;
                            const unsigned off = eOff[j], len = ePre[j + 1] - ePre[j];
                            const unsigned one = eFirst[j] ? 1u : 0x10001u;  // upper half: runs repeating an earlier seed
                            unsigned lo = 0, hi = len;
                            while (lo < hi) {
                                const unsigned mid = (lo + hi) >> 1;
                                if (__ldg(I.seedChunks + off + mid) < cLo) lo = mid + 1;
                                else hi = mid;
                            }
                            for (unsigned p = lo; p < len; p++) {
                                const unsigned c = __ldg(I.seedChunks + off + p);
                                if (c >= cHi) break;
                                atomicAdd(exact + ((unsigned)gi << gShift) + (c - cLo), one);
                            }
                        }
                        __syncthreads();
                        for (int x = tid; x < (bn << gShift); x += nT) {
                            const unsigned count = exact[x];  // soft count | repeated-seed runs << 16
                            if ((int)(count & 0xffffu) >= T) {
                                const int gx = b0 + (x >> gShift);
                                const unsigned g = gx < G.gListCap ? shGroup[gx] : gGroup[gx];
                                const unsigned chunk = (g << gShift) + ((unsigned)x & (gSize - 1u));
                                const int slot = atomicAdd(&sh.nCand, 1);
                                const unsigned long long e = ((unsigned long long)chunk << 32) | count;
                                if (slot < DP_BCAND) shCand[slot] = e;
                                else if (slot < S.tStride) gCand[slot] = e;
                            }
                        }
                        __syncthreads();
                    }
                }
                if (tid == 0) {
                    cRuns += (unsigned)nInc;
                    cEntries += total;
                }
                int nCand = sh.nCand;
                if (nCand > S.tStride) nCand = S.tStride;  // cannot happen: at most C chunks
                if (nCand > 0) {
                    // ---- ascending chunk id (rank sort: chunk ids are distinct) ----
                    const unsigned long long* sorted = shCand;
                    if (nCand <= DP_BCAND && nCand <= nT) {
                        unsigned long long e = 0;
                        int rank = 0;
                        if (tid < nCand) {
                            e = shCand[tid];
                            for (int y = 0; y < nCand; y++) rank += (shCand[y] >> 32) < (e >> 32);
                        }
                        __syncthreads();
                        if (tid < nCand) shCand[rank] = e;
                    } else {
                        for (int x = tid; x < DP_BCAND; x += nT) gCand[x] = shCand[x];
                        __syncthreads();
                        unsigned long long* dst = gCand + S.tStride;
                        for (int x = tid; x < nCand; x += nT) {
                            const unsigned long long e = gCand[x];
                            int rank = 0;
                            for (int y = 0; y < nCand; y++) rank += (gCand[y] >> 32) < (e >> 32);
                            dst[rank] = e;
                        }
                        sorted = dst;
                    }
                    if (warp == 0) {
                        DpRefineCtx X;
                        X.eOff = eOff;
                        X.ePre = ePre;
                        X.eEndW = eEndW;
                        X.eFirst = eFirst;
                        X.dup = shDup;
                        X.nDup = sh.nDup <= DP_BDUP ? sh.nDup : -1;
                        X.order = order;
                        X.sim = sh.sim;
                        X.nInc = nInc;
                        X.minCount = minCount;
                        X.T = T;
                        X.clamped = clamped;
                        X.q6 = q6;
                        X.nAllDistinct = 0;
                        int nOut = gShift ? dp_refine_emit<true>(I, X, sorted, nCand, candChunk + (size_t)ws * candStride,
                                                                 candDistinct + (size_t)ws * candStride, candStride)
                                          : dp_refine_emit<false>(I, X, sorted, nCand, candChunk + (size_t)ws * candStride,
                                                                  candDistinct + (size_t)ws * candStride, candStride);
                        if (nOut > candStride) {
                            if (lane == 0) atomicOr(&ctr->overflow, DP_OV_CANDS);
                            nOut = candStride;
                        }
                        if (lane == 0) sh.nCandOut = nOut;
                    }
                    __syncthreads();
                }
            }
        }
        if (tid == 0) {
            if (defer) {
                G.deferList[atomicAdd(G.nDefer, 1)] = ws;
            } else {
                candN[ws] = sh.nCandOut;
                cCand += (unsigned)sh.nCandOut;
            }
        }
        __syncthreads();
    }
    if (tid == 0 && (cRuns | cCand)) {
        atomicAdd(&ctr->posting_runs, cRuns);
        atomicAdd(&ctr->posting_entries, cEntries);
        atomicAdd(&ctr->candidates, cCand);
    }
}

// ===============================================================================================================
// Stage 2 for indexes of a few thousand chunks — one WARP per window strand with many loads in flight
// (dp_lookup_mid_kernel; BASELINE config 3: 6 465 chunks, ~70 seeds and ~1 650 postings per window strand).
//
// The whole index (tens of MB) sits in L2 here, so the kernel is bound by the chain of dependent loads of one window
// strand and by the shared-memory reductions, not by bytes. dp_lookup_kernel walks that chain one posting per lane at a
// time (~50 round trips); the CTA-per-window-strand kernel spends ~20 barriers on 1 650 postings. This kernel keeps the
// warp as the unit (no barrier anywhere) and takes the round trips out of the chain:
//   * the header of the next window strand is loaded an iteration ahead and its seeds are copied into shared memory by
//     cp.async while the current window strand is counted;
//   * ONE 16-byte gather per seed (DpIndexDev::midSeed: first posting, run length, first block of the padded copy, last
//     chunk >> 6) replaces the two CSR offsets and the run's last posting, four seeds per lane issued together;
//   * the runs are read from a copy made when the mapper is opened (DpIndexDev::midPost): every run starts on a 16-byte
//     block and is padded to whole blocks, a posting is stored ready to count — (byte offset of the chunk's counter
//     word) << 16 | (PRMT selector placing a 1 in the chunk's byte) — so a posting costs LEA.HI + PRMT + ATOMS and no
//     range test. Items of up to 8 blocks (8 lanes x uint4, four items per warp-wide load); a lane keeps a ring of NI
//     loads in flight and refills a slot as soon as it has counted it. The index loads carry an L2 evict_last policy:
//     the seed lists of 10^5 window strands stream through L2 beside them;
//   * at most DP_MID_ECAP = 128 included runs means a chunk's count stays below 256: ONE BYTE per chunk, four chunks
//     per 32-bit word of the warp's slice of shared memory (6.5 KB for config 3, so 20 warps stay resident per SM).
//     Postings are added with shared-memory reductions (no return value: nothing waits on them); one pass of 16-byte
//     loads finds the bytes that reached the threshold (adding 128 - T cannot carry: bit 7 = "count >= T") and blanks
//     the counters on the way;
//   * repeated seeds among the included runs (needed for the distinct-seed count) are found through a hash table that
//     borrows the still blank counter array: slot -> index of the run that claimed it, always verified against the
//     run's seed; the rare slot taken by a different seed is settled by a warp-wide compare.
// Level clamp (Q11), level-16 under-count (Q6) and the distinct counts: dp_refine_emit, as in the other kernels.
// Window strands outside its bounds (more than DP_MID_ECAP seeds, a seed present in every chunk, more than DP_MID_CAND
// chunks over the threshold) go to dp_lookup_kernel through a list, with identical results.
// Measured (profiles/r2i_*): 9.13 -> 4.83 ms per 262 144 reads of config 3 against dp_lookup_block_kernel.
// ===============================================================================================================
#define DP_MID_WARPS 4
#define DP_MID_ECAP 128   // query seeds per window strand (a byte counter holds up to 255)
#define DP_MID_ITEMS 128  // gather items listed at a time
#define DP_MID_CAND 64    // chunks over the threshold
#define DP_MID_DUP 32     // repeated-seed runs listed

struct DpMidWarp {  // shared memory of one warp
    unsigned eSeed[DP_MID_ECAP];
    unsigned eOff[DP_MID_ECAP];
    unsigned ePre[DP_MID_ECAP + 1];
    unsigned eItem[DP_MID_ECAP + 1];
    unsigned eEndW[DP_MID_ECAP];
    unsigned eMid[DP_MID_ECAP];         // first 16-byte block of the run in the mid-lookup copy of the postings
    unsigned itemStart[DP_MID_ITEMS];   // first 16-byte block of the item | (blocks - 1) << 28
    unsigned long long cand[DP_MID_CAND];
    unsigned short order[DP_MID_ECAP];
    unsigned short dup[DP_MID_DUP];
    unsigned char eFirst[DP_MID_ECAP];
    int sim[2];
    int nCand;
    int pad;
};

// words of one warp's byte counters (+ 32 dummy words), a multiple of 4
__host__ __device__ inline unsigned dp_mid_words(unsigned numChunks) { return ((((numChunks + 3u) >> 2) + 3u) & ~3u) + 32u; }

// the seeds of a window strand into a warp's shared memory, four per lane, without passing through registers
__device__ __forceinline__ void dp_mid_stage_seeds(unsigned seedAddr, const unsigned* __restrict__ qSeed, unsigned qb, int n,
                                                   unsigned lane) {
    if (n < 5 || n > DP_MID_ECAP) return;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int j = r * 32 + (int)lane;
        if (j < n) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(seedAddr + 4u * (unsigned)j), "l"(qSeed + qb + j) : "memory");
    }
}

template <int NI>
__global__ void __launch_bounds__(32 * DP_MID_WARPS, 5) dp_lookup_mid_kernel(DpIndexDev I, DpExtractOut Q, int nWS,
                                                                             int* __restrict__ deferList,
                                                                             int* __restrict__ nDefer,
                                                                             int* __restrict__ candN,
                                                                             unsigned* __restrict__ candChunk,
                                                                             unsigned short* __restrict__ candDistinct,
                                                                             int candStride, DpCounters* __restrict__ ctr) {
    extern __shared__ unsigned dp_smem[];  // DP_MID_WARPS x dp_mid_words(C): byte counters | 32 dummy words
    __shared__ DpMidWarp shW[DP_MID_WARPS];
    const unsigned lane = dp_lane();
    const unsigned lt = dp_lanemask_lt();
    const int wib = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned C = I.numChunks;
    const unsigned cWords = dp_mid_words(C) - 32u;
    unsigned* cnt = dp_smem + (size_t)wib * (cWords + 32u);
    unsigned short* h16 = reinterpret_cast<unsigned short*>(cnt);  // the blank counters double as the seed hash table
    const unsigned hSlots = C >> 1;                                 // 16-bit slots inside the C counter bytes
    DpMidWarp& W = shW[wib];
    for (unsigned c = lane; c < cWords; c += 32) cnt[c] = 0;
    __syncwarp();
    const unsigned cntAddr = (unsigned)__cvta_generic_to_shared(cnt);
    const uint4* post4 = I.midPost;
    const unsigned long long keep = dp_policy_evict_last();  // the index outlives the seed lists streaming through L2
    unsigned long long cRuns = 0, cEntries = 0, cCand = 0;
    // The seeds of the NEXT window strand are copied into W.eSeed (cp.async, no registers) while the current one is
    // counted, and its header is loaded an iteration ahead: two of the dependent round trips leave the chain.
    const unsigned seedAddr = (unsigned)__cvta_generic_to_shared(W.eSeed);
    int ws = gwarp;
    int n = 0;
    unsigned qb = 0;
    if (ws < nWS) {
        n = Q.wsN[ws];
        qb = Q.wsOff[ws];
        dp_mid_stage_seeds(seedAddr, Q.qSeed, qb, n, lane);
    }
    for (; ws < nWS; ws += nWarps) {
        int nNext = 0;
        unsigned qbNext = 0;
        if (ws + nWarps < nWS) {
            nNext = Q.wsN[ws + nWarps];
            qbNext = Q.wsOff[ws + nWarps];
        }
        bool staged = false;
        int nOut = 0;
        bool defer = n > DP_MID_ECAP;
        if (n >= 5 && !defer) {
            // ---- the window strand's seeds and posting runs: 4 per lane, all loads of a level issued together ----
            unsigned s[4], o[4], c[4], m[4], ew[4];
            asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int j = r * 32 + (int)lane;
                s[r] = j < n ? W.eSeed[j] : 0xffffffffu;  // (a lane reads what it copied itself)
            }
            __syncwarp();  // the compaction below rewrites W.eSeed
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const bool valid = r * 32 + (int)lane < n;
                // one 16-byte gather per seed: {first posting, run length, first block of the padded copy, last chunk >> 6}
                uint4 e = make_uint4(0, 0, 0, 0);
                if (valid) e = dp_load_index16(I.midSeed + s[r], keep);
                o[r] = e.x;
                c[r] = e.y;
                m[r] = e.z;
                ew[r] = e.w;
            }
            bool sawAll = false;
#pragma unroll
            for (int r = 0; r < 4; r++) sawAll |= (r * 32 + (int)lane < n) && c[r] >= C;
            defer = __any_sync(DP_FULL, sawAll);  // a seed present in every chunk: the general kernel's business
            if (!defer) {
                // ---- inclusion filter (seeds.go:340-346): every occurrence is eligible here; repeats of the previous
                //      occurrence's seed are skipped. Ordered compaction into shared memory ----
                int nInc = 0;
                unsigned carry = 0xffffffffu;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const bool valid = r * 32 + (int)lane < n;
                    unsigned prev = __shfl_up_sync(DP_FULL, s[r], 1);
                    if (lane == 0) prev = carry;
                    const bool inc = valid && s[r] != prev;
                    carry = __shfl_sync(DP_FULL, s[r], 31);
                    const unsigned mi = __ballot_sync(DP_FULL, inc);
                    if (inc) {
                        const int idx = nInc + __popc(mi & lt);
                        W.eSeed[idx] = s[r];
                        W.eOff[idx] = o[r];
                        W.ePre[idx] = c[r];  // run length for now; prefix-summed below
                        W.eMid[idx] = m[r];
                        W.eEndW[idx] = ew[r];  // bounds the run for the level clamp's early stop (Q11) and the Q6 simulation
                    }
                    nInc += __popc(mi);
                }
                __syncwarp();
                if (nInc >= 5) {
                    const int minCount = (nInc + 2) >> 2;
                    int T;
                    bool clamped = false;
                    if (minCount >= 9 && minCount <= 12) {
                        T = 8;
                        clamped = true;
                    } else if (minCount >= 17 && minCount <= 24) {
                        T = 16;
                        clamped = true;
                    } else {
                        T = minCount;
                    }
                    const bool q6 = minCount >= 13 && minCount <= 24;  // level-16 plane decides alone
                    // ---- per included run: repeated seed?, prefixes of postings and items ----
                    unsigned total = 0, nItems = 0;
                    int nDup = 0;
                    const int nRounds = (nInc + 31) >> 5;
                    for (int r = 0; r < nRounds; r++) {
                        const int j = r * 32 + (int)lane;
                        const bool have = j < nInc;
                        const unsigned es = have ? W.eSeed[j] : (0x80000000u | lane);  // (seed ranks are below 2^31)
                        const unsigned ec = have ? W.ePre[j] : 0u;
                        // repeated seed: inside the round by __match_any_sync, against earlier rounds through the hash
                        const unsigned mm = __match_any_sync(DP_FULL, es);
                        const bool leader = have && (__ffs(mm) - 1) == (int)lane;
                        const unsigned slot = __umulhi(es * 0x9E3779B1u, hSlots);
                        const unsigned t = leader ? (unsigned)h16[slot] : 0u;
                        __syncwarp();
                        bool isDup = have && !leader, unsure = false;
                        if (leader) {
                            if (t == 0) h16[slot] = (unsigned short)(j + 1);
                            else if (W.eSeed[t - 1] == es) isDup = true;
                            else unsure = true;  // the slot belongs to another seed: compare with every earlier run
                        }
                        unsigned mu = __ballot_sync(DP_FULL, unsure);
                        while (mu) {
                            const int l = __ffs(mu) - 1;
                            mu &= mu - 1;
                            const unsigned ss = __shfl_sync(DP_FULL, es, l);
                            bool found = false;
                            for (int q = 0; q < r; q++) found |= W.eSeed[q * 32 + (int)lane] == ss;
                            if (__any_sync(DP_FULL, found) && (int)lane == l) isDup = true;
                        }
                        if (have) W.eFirst[j] = isDup ? 0 : 1;
                        const unsigned md = __ballot_sync(DP_FULL, isDup);
                        if (nDup >= 0) {
                            if (nDup + __popc(md) > DP_MID_DUP) nDup = -1;
                            else {
                                if (isDup) W.dup[nDup + __popc(md & lt)] = (unsigned short)j;
                                nDup += __popc(md);
                            }
                        }
                        // items of the run: pieces of 8 blocks (32 postings) of its padded copy
                        const unsigned items = (ec + 31u) >> 5;
                        unsigned xa = ec, xb = items;
                        for (int d = 1; d < 32; d <<= 1) {
                            const unsigned ya = __shfl_up_sync(DP_FULL, xa, d), yb = __shfl_up_sync(DP_FULL, xb, d);
                            if ((int)lane >= d) {
                                xa += ya;
                                xb += yb;
                            }
                        }
                        if (have) {
                            W.ePre[j] = total + xa - ec;
                            W.eItem[j] = nItems + xb - items;
                        }
                        total += __shfl_sync(DP_FULL, xa, 31);
                        nItems += __shfl_sync(DP_FULL, xb, 31);
                    }
                    if (lane == 0) {
                        W.ePre[nInc] = total;
                        W.eItem[nInc] = nItems;
                        W.nCand = 0;
                    }
                    __syncwarp();
                    // hand the borrowed hash slots back blank
                    for (int j = (int)lane; j < nInc; j += 32) h16[__umulhi(W.eSeed[j] * 0x9E3779B1u, hSlots)] = 0;
                    __syncwarp();
                    // W.eSeed is dead from here on: the next window strand's seeds arrive while this one is counted
                    dp_mid_stage_seeds(seedAddr, Q.qSeed, qbNext, nNext, lane);
                    staged = true;
                    // ---- stream the runs into the byte counters ----
                    for (unsigned d0 = 0; d0 < nItems; d0 += DP_MID_ITEMS) {
                        const unsigned dN = min((unsigned)DP_MID_ITEMS, nItems - d0);
                        for (int j = (int)lane; j < nInc; j += 32) {
                            const unsigned first = W.eItem[j], last = W.eItem[j + 1];  // this run's items
                            if (last <= d0 || first >= d0 + dN) continue;
                            const unsigned blocks = (W.ePre[j + 1] - W.ePre[j] + 3u) >> 2, mo = W.eMid[j];
                            for (unsigned it = max(first, d0); it < min(last, d0 + dN); it++) {
                                const unsigned b0 = (it - first) << 3;
                                W.itemStart[it - d0] = (mo + b0) | ((min(8u, blocks - b0) - 1u) << 28);
                            }
                        }
                        __syncwarp();
                        // 8 lanes x 16 bytes per item, 4 items per warp-wide load; a ring of NI loads per lane: a slot is
                        // refilled as soon as it has been counted, so NI - 1 loads stay in flight throughout
                        const unsigned sub = lane >> 3, ln = lane & 7u;
                        uint4 v[NI];
                        unsigned ok = 0;  // bit d: slot d holds a block of postings
#pragma unroll
                        for (int d = 0; d < NI; d++) {
                            const unsigned idx = (unsigned)d * 4u + sub;
                            const unsigned desc = idx < dN ? W.itemStart[idx] : 0u;
                            if (idx < dN && ln <= (desc >> 28)) {
                                v[d] = dp_load_index16(post4 + (desc & 0x0fffffffu) + ln, keep);
                                ok |= 1u << d;
                            }
                        }
                        for (unsigned it = 0; it < dN; it += NI * 4) {
#pragma unroll
                            for (int d = 0; d < NI; d++) {
                                if (ok & (1u << d)) {
                                    const unsigned vv[4] = {v[d].x, v[d].y, v[d].z, v[d].w};
#pragma unroll
                                    for (int e = 0; e < 4; e++) {  // a posting = counter word offset << 16 | byte selector
                                        unsigned one;
                                        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(one) : "r"(1u), "r"(0u), "r"(vv[e]));
                                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(cntAddr + (vv[e] >> 16)), "r"(one) : "memory");
                                    }
                                }
                                // refill the slot
                                const unsigned idx = it + (unsigned)(NI * 4) + (unsigned)d * 4u + sub;
                                const unsigned desc = idx < dN ? W.itemStart[idx] : 0u;
                                ok &= ~(1u << d);
                                if (idx < dN && ln <= (desc >> 28)) {
                                    v[d] = dp_load_index16(post4 + (desc & 0x0fffffffu) + ln, keep);
                                    ok |= 1u << d;
                                }
                            }
                        }
                        __syncwarp();
                    }
                    // ---- bytes that reached the threshold (SIMD byte compare); the counters are blanked on the way ----
                    if (total) {
                        uint4* cnt4 = reinterpret_cast<uint4*>(cnt);
                        // a chunk is counted at most once per run and there are at most 128 runs: adding 128 - T to every
                        // byte cannot carry, and bit 7 of a byte then says "count >= T"
                        const unsigned K4 = (unsigned)(128 - T) * 0x01010101u;
                        for (unsigned c4 = lane; c4 < (cWords >> 2); c4 += 32) {
                            const uint4 x = cnt4[c4];
                            cnt4[c4] = make_uint4(0, 0, 0, 0);
                            const unsigned xs[4] = {x.x, x.y, x.z, x.w};
                            const unsigned hit = ((x.x + K4) | (x.y + K4) | (x.z + K4) | (x.w + K4)) & 0x80808080u;
                            if (hit) {
#pragma unroll
                                for (int i = 0; i < 4; i++) {
#pragma unroll
                                    for (int b8 = 0; b8 < 4; b8++) {
                                        const unsigned v = (xs[i] >> (8 * b8)) & 0xffu;
                                        if ((int)v >= T) {
                                            const int slotC = atomicAdd(&W.nCand, 1);
                                            if (slotC < DP_MID_CAND)
                                                W.cand[slotC] = ((unsigned long long)(c4 * 16u + (unsigned)i * 4u + (unsigned)b8) << 32) | v;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    const int nCand = W.nCand;
                    if (nCand > DP_MID_CAND) {
                        defer = true;  // (the general kernel lists candidates in memory)
                    } else {
                        cRuns += (unsigned)nInc;
                        cEntries += total;
                        if (nCand > 0) {
                            // ascending chunk id (rank sort in registers: chunk ids are distinct)
                            unsigned long long e0 = (int)lane < nCand ? W.cand[lane] : ~0ull;
                            unsigned long long e1 = (int)lane + 32 < nCand ? W.cand[lane + 32] : ~0ull;
                            int r0 = 0, r1 = 0;
                            for (int y = 0; y < nCand; y++) {
                                const unsigned cy = (unsigned)(W.cand[y] >> 32);
                                r0 += cy < (unsigned)(e0 >> 32);
                                r1 += cy < (unsigned)(e1 >> 32);
                            }
                            __syncwarp();
                            if ((int)lane < nCand) W.cand[r0] = e0;
                            if ((int)lane + 32 < nCand) W.cand[r1] = e1;
                            __syncwarp();
                            DpRefineCtx X;
                            X.eOff = W.eOff;
                            X.ePre = W.ePre;
                            X.eEndW = W.eEndW;
                            X.eFirst = W.eFirst;
                            X.dup = W.dup;
                            X.nDup = nDup;
                            X.order = W.order;
                            X.sim = W.sim;
                            X.nInc = nInc;
                            X.minCount = minCount;
                            X.T = T;
                            X.clamped = clamped;
                            X.q6 = q6;
                            X.nAllDistinct = 0;
                            nOut = dp_refine_emit<false>(I, X, W.cand, nCand, candChunk + (size_t)ws * candStride,
                                                         candDistinct + (size_t)ws * candStride, candStride);
                            if (nOut > candStride) {
                                if (lane == 0) atomicOr(&ctr->overflow, DP_OV_CANDS);
                                nOut = candStride;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (lane == 0) {
            if (defer) deferList[atomicAdd(nDefer, 1)] = ws;
            else candN[ws] = nOut;
        }
        if (!defer) cCand += (unsigned)nOut;
        if (!staged) {
            __syncwarp();
            dp_mid_stage_seeds(seedAddr, Q.qSeed, qbNext, nNext, lane);
        }
        n = nNext;
        qb = qbNext;
    }
    if (lane == 0 && (cRuns | cCand)) {
        atomicAdd(&ctr->posting_runs, cRuns);
        atomicAdd(&ctr->posting_entries, cEntries);
        atomicAdd(&ctr->candidates, cCand);
    }
}

// ===============================================================================================================
// Stage 3 — chaining: the candidate loop of performMapping (mapping/mapping.go:518-608) with
// SeedSequence.Match -> Reduced -> dynamicMatch -> extendChain (seeds/sequence.go:85-123, 361-576) and the
// coordinate arithmetic of GetSeedOffset / GetSeedOffsetFromEnd / GetBasesCovered (sequence.go:830-858, 1239-1276)
// done on scan positions. One warp per window: candidates are visited in ascending chunk id, forward strand first,
// because the thresholds minMatches / minRCMatches escalate as chains are accepted (Q12).
//
// chunk.Reduced(querySet) never streams the chunk: each distinct query seed binary-searches its position-carrying
// posting run for the candidate chunk; the few (position, seed) pairs found are rank-sorted by position and the
// same-as-previous seeds collapsed. query.Reduced(chunkSet) is the query list filtered by "seed found in the chunk".
// A seed is identified by the index of its first occurrence in the strand's list (__match_any_sync), so no hashing.
// dynamicMatch's search for chain starts is a warp ballot per query seed; lane 0 runs the greedy extension.
// ===============================================================================================================
#define DP_QCAP 128  // query entries per strand held in shared memory (longer lists use the global scratch)
#define DP_MCAP 128  // chunk-side entries per candidate held in shared memory

struct DpChainScratch {  // per-warp global scratch for lists that do not fit shared memory, and for rare big records
    unsigned short* qFirst;  // [qStride]
    unsigned short* qCnt;    // [qStride]
    unsigned* qLo;           // [qStride]
    int* rqPos;              // [qStride]
    unsigned short* rqId;    // [qStride]
    int* chainLen;           // [qStride]
    int* lastB;              // [qStride]
    unsigned long long* ent; // [sStride]
    int* rsPos;              // [sStride]
    unsigned short* rsId;    // [sStride]
    int* chains;             // accepted chains of the current candidate: 6 ints each [chainCap*6]
    DpMappingDev* results;   // window results before sort/dedupe [resultCap]
    int qStride;
    int sStride;
    int chainCap;
    int resultCap;
};

struct DpChainLists {  // the working lists of one candidate (shared or global memory)
    const int* rqPos;
    const unsigned short* rqId;
    int nq, qScanLen;
    const int* rsPos;
    const unsigned short* rsId;
    int ns, sScanLen;
    int* chainLen;
    int* lastB;
};

// Warp-collective dynamicMatch (sequence.go:401-471). Accepted chains -> ch[] as {len, firstA, lastA, firstB, lastB,
// ids} (indices into the reduced lists); returns their number, or -1 on chain list overflow. Uniform across lanes.
__device__ int dp_dynamic_match(const DpChainLists& T, int minMatch, int k, int* ch, int chainCap) {
    const unsigned lane = dp_lane();
    const int nq = T.nq, ns = T.ns;
    if (minMatch == 0) minMatch = 1;
    int nGood = 0;
    int nilCount = nq;  // lane 0's copy is authoritative
    for (int x = lane; x < nq; x += 32) T.chainLen[x] = 0;
    __syncwarp();
#define GAPQ(i) (((i) + 1 < nq ? T.rqPos[(i) + 1] : T.qScanLen) - T.rqPos[(i)] - k)
#define GAPS(i) (((i) + 1 < ns ? T.rsPos[(i) + 1] : T.sScanLen) - T.rsPos[(i)] - k)
    for (int qi = 0; qi <= nq - minMatch; qi++) {
        // sequence.go:409: internal to closely spaced repeats (cannot fire on reduced lists; kept for fidelity)
        if (qi > 0 && qi + 1 < nq && GAPQ(qi - 1) < 0 && GAPQ(qi) < 0 && T.rqId[qi] == T.rqId[qi - 1] &&
            T.rqId[qi] == T.rqId[qi + 1])
            continue;
        if (T.chainLen[qi] != 0) continue;  // uniform: shared/global memory made visible by the __syncwarp below
        const unsigned short qid = T.rqId[qi];
        for (int s0 = 0; s0 < ns; s0 += 32) {
            int si = s0 + (int)lane;
            bool cand = si < ns && T.rsId[si] == qid && (si == 0 || T.rsId[si - 1] != qid);
            unsigned m = __ballot_sync(DP_FULL, cand);
            while (m) {
                const int si0 = s0 + __ffs(m) - 1;
                m &= m - 1;
                int ret = 0;  // 1: early return requested, -1: overflow
                if (lane == 0) {
                    // conditions re-evaluated with the state as it is now (minMatch may have grown)
                    if (si0 <= ns - minMatch && (T.chainLen[qi] == 0 || T.lastB[qi] != si0)) {
                        if (T.chainLen[qi] == 0) nilCount--;
                        T.chainLen[qi] = 1;
                        T.lastB[qi] = si0;
                        // ---- extendChain (sequence.go:476-576) ----
                        int curLen = 1;
                        int ids = k;
                        int lastA = qi, lastBi = si0;
                        int offsetA = GAPQ(qi);
                        int offsetB = GAPS(si0);
                        int ai = qi + 1, bi = si0 + 1;
                        bool done = false;
                        while (!done && ai < nq && bi < ns) {
                            int minB, maxB;
                            if (offsetA < 0) {
                                minB = -k;
                                maxB = 0;
                            } else {
                                minB = (offsetA * 2) / 3 - k;
                                maxB = (offsetA * 3) / 2 + k;
                            }
                            while (maxB < offsetB) {
                                offsetA += GAPQ(ai) + k;
                                ai++;
                                if (ai >= nq) {
                                    done = true;
                                    break;
                                }
                                minB = (offsetA * 2) / 3 - k;
                                maxB = (offsetA * 3) / 2 + k;
                            }
                            if (done) break;
                            while (offsetB < minB) {
                                offsetB += GAPS(bi) + k;
                                bi++;
                                if (bi >= ns) {
                                    done = true;
                                    break;
                                }
                            }
                            if (done) break;
                            int oldBi = bi, oldBOffset = offsetB;
                            bool matched = false;
                            const unsigned short seedA = T.rqId[ai];
                            while (offsetB <= maxB) {
                                if (seedA == T.rsId[bi]) {
                                    if (T.chainLen[ai] != 0) {
                                        if (bi == T.lastB[ai] && T.chainLen[ai] > curLen) {
                                            done = true;  // they have a better chain already
                                            break;
                                        }
                                    } else {
                                        nilCount--;
                                    }
                                    curLen++;
                                    T.chainLen[ai] = curLen;
                                    T.lastB[ai] = bi;
                                    int d2 = T.rsPos[bi] - T.rsPos[lastBi] - k;  // GetBasesCovered, reference side
                                    ids += k + (d2 < 0 ? d2 : 0);
                                    lastA = ai;
                                    lastBi = bi;
                                    offsetA = GAPQ(ai);
                                    offsetB = GAPS(bi);
                                    ai++;
                                    bi++;
                                    matched = true;
                                    break;
                                } else {
                                    offsetB += GAPS(bi) + k;
                                    bi++;
                                    if (bi >= ns) break;
                                }
                            }
                            if (done) break;
                            if (!matched) {
                                offsetA += GAPQ(ai) + k;
                                ai++;
                                offsetB = oldBOffset;
                                bi = oldBi;
                            }
                        }
                        // ---- dynamicMatch bookkeeping (sequence.go:435-465) ----
                        if (curLen >= minMatch) {
                            int nextLength = (curLen * 2) / 3;
                            if (nextLength > minMatch) {
                                minMatch = nextLength;
                                for (int j = nGood - 1; j >= 0; j--) {
                                    if (ch[j * 6] < nextLength) {
                                        for (int z = 0; z < 6; z++) ch[j * 6 + z] = ch[(nGood - 1) * 6 + z];
                                        nGood--;
                                    }
                                }
                            }
                            if (nGood >= chainCap) {
                                ret = -1;
                            } else {
                                int* r = ch + nGood * 6;
                                r[0] = curLen;
                                r[1] = qi;
                                r[2] = lastA;
                                r[3] = si0;
                                r[4] = lastBi;
                                r[5] = ids;
                                nGood++;
                                if (nilCount < curLen) ret = 1;
                            }
                        }
                    }
                }
                __syncwarp();
                ret = __shfl_sync(DP_FULL, ret, 0);
                minMatch = __shfl_sync(DP_FULL, minMatch, 0);
                nGood = __shfl_sync(DP_FULL, nGood, 0);
                if (ret < 0) return -1;
                if (ret > 0) return nGood;
            }
        }
    }
#undef GAPQ
#undef GAPS
    return nGood;
}

// Working arrays of one warp for the lists of one candidate (shared memory; global scratch for oversized lists)
struct DpListBuf {
    unsigned short* qFirst;  // [n] index of the first occurrence of each query entry's seed (seed identity)
    unsigned short* qCnt;    // [n] postings of the seed inside the candidate chunk (first occurrences only)
    unsigned* qLo;           // [n] first such posting
    int* rqPos;              // reduced query list
    unsigned short* rqId;
    unsigned long long* shEnt;  // chunk side, DP_MCAP entries of shared memory ...
    int* shRsPos;
    unsigned short* shRsId;
    unsigned long long* gEnt;   // ... or sStride entries of global scratch
    int* gRsPos;
    unsigned short* gRsId;
    int sStride;
};

// seed identity = index of the first occurrence of the seed in this strand's list (warp-collective)
__device__ __forceinline__ void dp_seed_identity(const DpExtractOut& Q, unsigned qb, int n, unsigned short* qFirst) {
    const unsigned lane = dp_lane();
    for (int j0 = 0; j0 < n; j0 += 32) {
        int j = j0 + (int)lane;
        unsigned s = j < n ? Q.qSeed[qb + j] : (0x80000000u | lane);  // padding lanes never match
        unsigned mm = __match_any_sync(DP_FULL, s);
        int first = j0 + __ffs(mm) - 1;
        if (j < n && j0 > 0) {
            for (int b = 0; b < j0; b++)
                if (Q.qSeed[qb + b] == s) {
                    first = b;
                    break;
                }
        }
        if (j < n) qFirst[j] = (unsigned short)first;
    }
    __syncwarp();
}

// query.Reduced(chunkSet) and chunk.Reduced(querySet) for candidate chunk c (sequence.go:85-123), warp-collective.
// Returns false if the chunk side does not fit the scratch. Outputs: nq entries in B.rqPos/B.rqId; ns entries in
// *rsPosOut / *rsIdOut (shared or global, depending on size).
__device__ bool dp_build_lists(const DpIndexDev& I, const DpExtractOut& Q, unsigned qb, int n, unsigned c,
                               const DpListBuf& B, int& nqOut, int& nsOut, int*& rsPosOut, unsigned short*& rsIdOut) {
    const unsigned lane = dp_lane();
    const unsigned lt = dp_lanemask_lt();
    // every distinct query seed looks up its postings inside chunk c
    int m = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
        int j = j0 + (int)lane;
        unsigned lo = 0;
        int cnt = 0;
        if (j < n && B.qFirst[j] == j) {
            unsigned s = Q.qSeed[qb + j];
            unsigned b = __ldg(I.postOff + s), e = __ldg(I.postOff + s + 1);
            unsigned hi = e;
            lo = b;
            while (lo < hi) {
                unsigned mid = (lo + hi) >> 1;
                if (__ldg(I.postChunk + mid) < c) lo = mid + 1;
                else hi = mid;
            }
            while (lo + cnt < e && cnt < 65535 && __ldg(I.postChunk + lo + cnt) == c) cnt++;
        }
        if (j < n) {
            B.qLo[j] = lo;
            B.qCnt[j] = (unsigned short)cnt;
        }
        unsigned x = (unsigned)cnt;
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(DP_FULL, x, d);
        m += (int)x;
    }
    __syncwarp();
    // query.Reduced(chunkSet): entries whose seed occurs in the chunk, same-as-previous-member collapsed.
    // (Both Reduced calls precede the nil test, sequence.go:366-374.)
    int nq = 0;
    int prevMember = -1;
    for (int j0 = 0; j0 < n; j0 += 32) {
        int j = j0 + (int)lane;
        int id = -2 - (int)lane;
        bool member = false;
        if (j < n) {
            id = B.qFirst[j];
            member = B.qCnt[id] != 0;
        }
        unsigned mm = __ballot_sync(DP_FULL, member);
        unsigned lower = mm & lt;
        int src = lower ? 31 - __clz(lower) : 0;
        int pid = __shfl_sync(DP_FULL, id, src);
        if (!lower) pid = prevMember;
        bool keep = member && id != pid;
        unsigned mk = __ballot_sync(DP_FULL, keep);
        if (keep) {
            int idx = nq + __popc(mk & lt);
            B.rqId[idx] = (unsigned short)id;
            B.rqPos[idx] = Q.qPos[qb + j];
        }
        nq += __popc(mk);
        if (mm) prevMember = __shfl_sync(DP_FULL, id, 31 - __clz(mm));
    }
    nqOut = nq;
    // chunk.Reduced(querySet): gather (position, seed) pairs, rank-sort by position, collapse
    if (m > B.sStride) return false;  // cannot happen: a chunk has at most maxChunkSeeds entries
    const bool sSmall = m <= DP_MCAP;
    unsigned long long* ent = sSmall ? B.shEnt : B.gEnt;
    int* rsPos = sSmall ? B.shRsPos : B.gRsPos;
    unsigned short* rsId = sSmall ? B.shRsId : B.gRsId;
    {
        int base = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            int j = j0 + (int)lane;
            int cnt = j < n ? (int)B.qCnt[j] : 0;
            int x = cnt;  // inclusive scan
            for (int d = 1; d < 32; d <<= 1) {
                int y = __shfl_up_sync(DP_FULL, x, d);
                if ((int)lane >= d) x += y;
            }
            int off = base + x - cnt;
            if (cnt) {
                unsigned lo = B.qLo[j];
                for (int t = 0; t < cnt; t++)
                    ent[off + t] = ((unsigned long long)(unsigned)__ldg(I.postPos + lo + t) << 32) | (unsigned)j;
            }
            base += __shfl_sync(DP_FULL, x, 31);
        }
    }
    __syncwarp();
    for (int i0 = 0; i0 < m; i0 += 32) {  // scan positions inside a chunk are distinct: ranks are unique
        int i = i0 + (int)lane;
        if (i < m) {
            unsigned long long mine = ent[i];
            int rank = 0;
            for (int x = 0; x < m; x++) rank += (ent[x] >> 32) < (mine >> 32);
            rsPos[rank] = (int)(mine >> 32);
            rsId[rank] = (unsigned short)(mine & 0xffffu);
        }
    }
    __syncwarp();
    int ns = 0;
    prevMember = -1;
    for (int i0 = 0; i0 < m; i0 += 32) {
        int i = i0 + (int)lane;
        int id = -2 - (int)lane, pos = 0;
        if (i < m) {
            id = rsId[i];
            pos = rsPos[i];
        }
        int pid = __shfl_up_sync(DP_FULL, id, 1);
        if (lane == 0) pid = prevMember;
        bool keep = i < m && id != pid;
        unsigned mk = __ballot_sync(DP_FULL, keep);
        __syncwarp();
        if (keep) {
            int idx = ns + __popc(mk & lt);
            rsId[idx] = (unsigned short)id;
            rsPos[idx] = pos;
        }
        ns += __popc(mk);
        prevMember = __shfl_sync(DP_FULL, id, 31);
        __syncwarp();
    }
    nsOut = ns;
    rsPosOut = rsPos;
    rsIdOut = rsId;
    return true;
}

// The two Reduced lists of dp_build_lists for a window strand with at most 32 seeds, in registers: lane j holds seed
// occurrence j (`s`, its scan position `qPos`, `first` = lane of the seed's first occurrence, and for first occurrences
// the bounds [pb, pe) of the seed's position-carrying posting run). Returns false when the chunk side has more than 32
// entries (the caller then takes the general routine). On success lane t < nq holds the t-th query-side word in
// `rqWord`, lane t < ns the t-th chunk-side word in `rsWord` (scan position << 16 | seed identity).
__device__ __forceinline__ bool dp_build_lists_small(const DpIndexDev& I, int n, unsigned c, int first, int qPos, unsigned pb,
                                                     unsigned pe, unsigned* sortScratch, int& nq, int& ns,
                                                     unsigned& rqWord, unsigned& rsWord) {
    const unsigned lane = dp_lane();
    const unsigned lt = dp_lanemask_lt();
    const bool isFirst = (int)lane < n && first == (int)lane;
    // every distinct query seed looks up its postings inside chunk c
    unsigned lo = pb;
    unsigned cnt = 0;
    if (isFirst) {
        unsigned hi = pe;
        while (lo < hi) {
            const unsigned mid = (lo + hi) >> 1;
            if (__ldg(I.postChunk + mid) < c) lo = mid + 1;
            else hi = mid;
        }
        while (lo + cnt < pe && __ldg(I.postChunk + lo + cnt) == c) cnt++;
    }
    unsigned x = cnt;  // inclusive scan of the counts
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(DP_FULL, x, d);
        if ((int)lane >= d) x += y;
    }
    const unsigned pre = x - cnt;
    const unsigned m = __shfl_sync(DP_FULL, x, 31);
    if (m > 32u) return false;
    // query.Reduced(chunkSet): occurrences whose seed is in the chunk, same-as-previous-member collapsed
    const unsigned cntFirst = __shfl_sync(DP_FULL, cnt, first & 31);
    const bool member = (int)lane < n && cntFirst != 0;
    const unsigned mmask = __ballot_sync(DP_FULL, member);
    const unsigned lower = mmask & lt;
    int pid = __shfl_sync(DP_FULL, first, lower ? 31 - __clz(lower) : 0);
    if (!lower) pid = -1;
    const bool keep = member && first != pid;
    const unsigned mk = __ballot_sync(DP_FULL, keep);
    nq = __popc(mk);
    {
        const unsigned word = ((unsigned)qPos << 16) | (unsigned)first;
        const int src = (int)lane < nq ? (int)__fns(mk, 0, (int)lane + 1) : 0;
        rqWord = __shfl_sync(DP_FULL, word, src);
    }
    // chunk.Reduced(querySet): (position, seed) pairs of the chunk side, sorted by position, collapsed
    unsigned ent = 0xffffffffu;  // position << 16 | seed identity (lane of the first occurrence)
    {
        int rl = 0, rh = 32;  // largest lane r with pre_r <= lane (runs of zero length are never selected, see below)
        for (int step = 0; step < 5; step++) {
            const int mid = (rl + rh) >> 1;
            const unsigned v = __shfl_sync(DP_FULL, pre, mid);
            if (v <= lane) rl = mid;
            else rh = mid;
        }
        const unsigned rLo = __shfl_sync(DP_FULL, lo, rl);
        const unsigned rPre = __shfl_sync(DP_FULL, pre, rl);
        if (lane < m) ent = ((unsigned)__ldg(I.postPos + rLo + (lane - rPre)) << 16) | (unsigned)rl;
    }
    // (scan positions inside a chunk are distinct: ranks are unique)
    int rank = 0;
    for (unsigned y = 0; y < m; y++) rank += (__shfl_sync(DP_FULL, ent, (int)y) >> 16) < (ent >> 16);
    __syncwarp();
    if (lane < m) sortScratch[rank] = ent;
    __syncwarp();
    const unsigned sorted = lane < m ? sortScratch[lane] : 0xffffffffu;
    __syncwarp();
    const unsigned id = sorted & 0xffffu;
    unsigned prevId = __shfl_up_sync(DP_FULL, id, 1);
    if (lane == 0) prevId = 0xffffffffu;
    const bool keepS = lane < m && id != prevId;
    const unsigned ms = __ballot_sync(DP_FULL, keepS);
    ns = __popc(ms);
    {
        const int src = (int)lane < ns ? (int)__fns(ms, 0, (int)lane + 1) : 0;
        rsWord = __shfl_sync(DP_FULL, sorted, src);
    }
    return true;
}

// Exact general path, one warp per window: used for the windows the fast path (dp_reduce_kernel + dp_chain_thread_kernel
// below) hands back (`winList` / `nWinList` set), and for everything when the fast path is switched off.
__global__ void __launch_bounds__(128, 6) dp_chain_kernel(DpIndexDev I, const DpWindow* __restrict__ wins,
                                                       const int* __restrict__ winList,
                                                       const int* __restrict__ nWinList,
                                                       const int* __restrict__ readLen, int nWin, DpExtractOut Q,
                                                       const int* __restrict__ candN,
                                                       const unsigned* __restrict__ candChunk,
                                                       const unsigned short* __restrict__ candDistinct, int candStride,
                                                       DpChainScratch S, int* __restrict__ outN,
                                                       unsigned* __restrict__ outOff,
                                                       DpMappingDev* __restrict__ outMaps,
                                                       unsigned long long* __restrict__ outCursor,
                                                       unsigned long long outCapacity,
                                                       DpCounters* __restrict__ ctr) {
    // shared memory per warp
    __shared__ unsigned short shFirst[4][DP_QCAP];
    __shared__ unsigned short shCnt[4][DP_QCAP];
    __shared__ unsigned shLo[4][DP_QCAP];
    __shared__ int shRqPos[4][DP_QCAP];
    __shared__ unsigned short shRqId[4][DP_QCAP];
    __shared__ int shChainLen[4][DP_QCAP];
    __shared__ int shLastB[4][DP_QCAP];
    __shared__ unsigned long long shEnt[4][DP_MCAP];
    __shared__ int shRsPos[4][DP_MCAP];
    __shared__ unsigned short shRsId[4][DP_MCAP];
    const unsigned lane = dp_lane();
    const unsigned lt = dp_lanemask_lt();
    const int wib = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    const int k = I.k;
    int* chains = S.chains + (size_t)gwarp * S.chainCap * 6;
    DpMappingDev* results = S.results + (size_t)gwarp * S.resultCap;
    unsigned long long cCells = 0, cMaps = 0;

    const int nTodo = winList ? min(*nWinList, nWin) : nWin;
    for (int wi = gwarp; wi < nTodo; wi += nWarps) {
        const int w = winList ? winList[wi] : wi;
        DpWindow win = wins[w];
        const int L = win.len;
        const int rlen = readLen[win.read];
        const bool q2 = win.whole && ((L & 3) == 0);
        // SeedSequence.offset / inset of the window (Q3: SubSequence stores inset one too large)
        const int wOffset = win.whole ? 0 : win.start;
        const int wInset = win.whole ? 0 : (rlen - (win.start + L) + 1);
        int nRes = 0;
        int minMatches = Q.wsN[2 * w] / 5;
        int minRCMatches = Q.wsN[2 * w + 1] / 5;
        if (minMatches < 5) minMatches = 5;
        if (minRCMatches < 5) minRCMatches = 5;
        bool overflow = false;

        for (int strand = 0; strand < 2; strand++) {
            const int ws = 2 * w + strand;
            const int nc = candN[ws];
            if (nc == 0) continue;
            const int n = Q.wsN[ws];
            const unsigned qb = Q.wsOff[ws];
            const int qScanLen = strand == 0 ? (L - (q2 ? 4 : 0)) : (L - (q2 ? 3 : 0));
            const bool qSmall = n <= DP_QCAP;
            unsigned short* qFirst = qSmall ? shFirst[wib] : S.qFirst + (size_t)gwarp * S.qStride;
            unsigned short* qCnt = qSmall ? shCnt[wib] : S.qCnt + (size_t)gwarp * S.qStride;
            unsigned* qLo = qSmall ? shLo[wib] : S.qLo + (size_t)gwarp * S.qStride;
            int* rqPos = qSmall ? shRqPos[wib] : S.rqPos + (size_t)gwarp * S.qStride;
            unsigned short* rqId = qSmall ? shRqId[wib] : S.rqId + (size_t)gwarp * S.qStride;
            int* chainLen = qSmall ? shChainLen[wib] : S.chainLen + (size_t)gwarp * S.qStride;
            int* lastB = qSmall ? shLastB[wib] : S.lastB + (size_t)gwarp * S.qStride;
            DpListBuf LB;
            LB.qFirst = qFirst;
            LB.qCnt = qCnt;
            LB.qLo = qLo;
            LB.rqPos = rqPos;
            LB.rqId = rqId;
            LB.shEnt = shEnt[wib];
            LB.shRsPos = shRsPos[wib];
            LB.shRsId = shRsId[wib];
            LB.gEnt = S.ent + (size_t)gwarp * S.sStride;
            LB.gRsPos = S.rsPos + (size_t)gwarp * S.sStride;
            LB.gRsId = S.rsId + (size_t)gwarp * S.sStride;
            LB.sStride = S.sStride;
            dp_seed_identity(Q, qb, n, qFirst);
            for (int ci = 0; ci < nc; ci++) {
                const int thr = strand == 0 ? minMatches : minRCMatches;
                // 1. CountIntersectionTo(...) < threshold (mapping.go:520-523, 559-562)
                if ((int)candDistinct[(size_t)ws * candStride + ci] < thr) continue;
                const unsigned c = candChunk[(size_t)ws * candStride + ci];
                // 2.-4. query.Reduced(chunkSet) and chunk.Reduced(querySet)
                int nq, ns;
                int* rsPos;
                unsigned short* rsId;
                if (!dp_build_lists(I, Q, qb, n, c, LB, nq, ns, rsPos, rsId)) {
                    overflow = true;
                    continue;
                }
                if (ns < thr || nq < thr) continue;  // Reduced returned nil
                cCells += (unsigned)(ns + nq);
                // 5. dynamicMatch
                const int sScanLen = __ldg(I.chunkScanLen + c);
                DpChainLists T;
                T.rqPos = rqPos;
                T.rqId = rqId;
                T.nq = nq;
                T.qScanLen = qScanLen;
                T.rsPos = rsPos;
                T.rsId = rsId;
                T.ns = ns;
                T.sScanLen = sScanLen;
                T.chainLen = chainLen;
                T.lastB = lastB;
                int nGood = dp_dynamic_match(T, thr, k, chains, S.chainCap);
                if (nGood < 0) {  // (the host reruns the launch with a longer chain list)
                    if (lane == 0) atomicOr(&ctr->overflow, DP_OV_CHAINS);
                    nGood = 0;
                }
                __syncwarp();
                // 6. chains -> mappings (mapping.go:528-549 / 567-587), in allGoodChains order; lane 0 keeps the state
                if (lane == 0) {
                    const long long cOffset = __ldg(I.chunkOffset + c);
                    const long long cInset = __ldg(I.chunkInset + c);
                    for (int g = 0; g < nGood; g++) {
                        const int* r = chains + g * 6;
                        long long start = cOffset + rsPos[r[3]];
                        long long end = I.refLen - cInset - (long long)(sScanLen - rsPos[r[4]] - k);
                        if (I.circular && start > I.refLen) start -= I.refLen;
                        int first = rqPos[r[1]];                   // GetSeedOffset(MatchA[0])
                        int fromEnd = qScanLen - rqPos[r[2]] - k;  // GetSeedOffsetFromEnd(MatchA[last])
                        if (first + fromEnd > (L * 2) / 3) continue;
                        DpMappingDev mp;
                        mp.start = start;
                        mp.end = end;
                        if (strand == 0) {
                            mp.qOffset = first + wOffset;
                            mp.qInset = fromEnd + wInset;
                        } else {  // rcQuery.offset = window inset, rcQuery.inset = window offset
                            mp.qInset = first + wInset;
                            mp.qOffset = fromEnd + wOffset;
                        }
                        mp.ids = r[5];
                        mp.rc = strand;
                        if (nRes < S.resultCap) results[nRes] = mp;
                        else overflow = true;
                        nRes++;
                        int limit = (r[0] * 4) / 5;
                        if (strand == 0) {
                            if (limit > minMatches) minMatches = limit;
                            if (limit > minRCMatches) minRCMatches = limit;
                        } else {
                            if (limit > minRCMatches) minRCMatches = limit;
                        }
                    }
                }
                nRes = __shfl_sync(DP_FULL, nRes, 0);
                minMatches = __shfl_sync(DP_FULL, minMatches, 0);
                minRCMatches = __shfl_sync(DP_FULL, minRCMatches, 0);
                overflow = __shfl_sync(DP_FULL, (int)overflow, 0) != 0;
            }
            __syncwarp();
        }
        // ---- sort by Start + overlap dedupe (mapping.go:590-608); lane 0 ----
        if (lane == 0) {
            if (overflow || nRes > S.resultCap) {
                atomicOr(&ctr->overflow, DP_OV_RESULTS);
                if (nRes > S.resultCap) nRes = S.resultCap;
            }
            if (nRes > 1) {
                for (int i = 1; i < nRes; i++) {  // stable insertion sort (= Go's sort.Sort for n <= 12)
                    DpMappingDev x = results[i];
                    int j = i;
                    while (j > 0 && x.start < results[j - 1].start) {
                        results[j] = results[j - 1];
                        j--;
                    }
                    results[j] = x;
                }
                for (int i = nRes - 1; i > 0; i--) {
                    DpMappingDev ra = results[i - 1], rb = results[i];
                    if (ra.rc == rb.rc && rb.start < ra.end) {
                        if (ra.end - ra.start > rb.end - rb.start) {
                            results[i] = results[nRes - 1];
                            nRes--;
                        } else {
                            results[i - 1] = results[i];
                            results[i] = results[nRes - 1];
                            nRes--;
                        }
                    }
                }
            }
            // compact output: one bump allocation per window
            unsigned long long base = nRes ? atomicAdd(outCursor, (unsigned long long)nRes) : 0ull;
            if (base + (unsigned)nRes > outCapacity) {
                atomicOr(&ctr->overflow, DP_OV_OUTPOOL);
                nRes = 0;
                base = 0;
            }
            outN[w] = nRes;
            outOff[w] = (unsigned)base;
            for (int i = 0; i < nRes; i++) outMaps[base + i] = results[i];
            cMaps += (unsigned)nRes;
        }
        __syncwarp();
    }
    if (lane == 0 && (cCells | cMaps)) {
        atomicAdd(&ctr->chain_cells, cCells);
        atomicAdd(&ctr->mappings, cMaps);
    }
}

// ===============================================================================================================
// Stage 3, fast path = the same computation as dp_chain_kernel split where its parallelism changes shape:
//   dp_reduce_kernel        one WARP per window: for every candidate that passes the window's INITIAL thresholds,
//                           builds the two reduced lists (warp-parallel) and stores them in a compact pool;
//   dp_chain_thread_kernel  one THREAD per window: the sequential part — the threshold escalation over candidates
//                           (Q12), dynamicMatch/extendChain, chains -> mappings, sort + dedupe — which in
//                           dp_chain_kernel occupies one lane of a warp while 31 idle.
// Thresholds only grow while a window is processed, so a candidate that fails the initial ones fails the final ones:
// building lists under the initial thresholds is a superset of what the sequential pass needs.
// Every bounded resource of the fast path (list pool, 8 chains per candidate, 8 mappings per window) has an exact way
// out: the window is appended to `slowList` and recomputed by dp_chain_kernel, whose scratch is sized for the worst
// case. Results are identical either way.
// ===============================================================================================================
#define DP_FAST_CHAINS 8
#define DP_FAST_RESULTS 8

struct DpChainTask {  // one candidate whose lists were built
    unsigned listOff;  // pool offset: nq query entries, ns chunk entries, nq memo words
    unsigned short nq, ns;  // nq == 0xffff: not built (failed the initial thresholds)
};

struct DpFastChain {
    DpChainTask* tasks;        // [taskCap]
    unsigned* taskBase;        // [2*nWin] first task of each window strand (valid where candN > 0)
    unsigned* pool;            // packed entries: scan position << 16 | seed identity
    unsigned long long* cursors;  // [0] tasks, [1] pool words
    unsigned long long taskCap, poolCap;
    unsigned taskSlab, poolSlab;  // a warp reserves this many task slots / pool words per global atomic
    unsigned char* slow;       // [nWin] window handed to the general path
    int* slowList;             // [nWin]
    int* nSlow;
};

__device__ __forceinline__ void dp_mark_slow(const DpFastChain& F, int w) {
    F.slow[w] = 1;
    F.slowList[atomicAdd(F.nSlow, 1)] = w;
}

__global__ void __launch_bounds__(128, 8) dp_reduce_kernel(DpIndexDev I, const DpWindow* __restrict__ wins, int nWin,
                                                           DpExtractOut Q, const int* __restrict__ candN,
                                                           const unsigned* __restrict__ candChunk,
                                                           const unsigned short* __restrict__ candDistinct,
                                                           int candStride, DpChainScratch S, DpFastChain F) {
    __shared__ unsigned short shFirst[4][DP_QCAP];
    __shared__ unsigned short shCnt[4][DP_QCAP];
    __shared__ unsigned shLo[4][DP_QCAP];
    __shared__ int shRqPos[4][DP_QCAP];
    __shared__ unsigned short shRqId[4][DP_QCAP];
    __shared__ unsigned long long shEnt[4][DP_MCAP];
    __shared__ int shRsPos[4][DP_MCAP];
    __shared__ unsigned short shRsId[4][DP_MCAP];
    const unsigned lane = dp_lane();
    const int wib = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    // Task slots and pool words are bump-allocated from two global cursors. A round trip to a global atomic costs a
    // warp about a microsecond, so a warp reserves a slab per atomic and hands out pieces of it locally (the unused
    // tail of a slab is lost: the pool is sized for it).
    unsigned long long tCur = 0, tEnd = 0, pCur = 0, pEnd = 0;  // warp-uniform
    auto take = [&](unsigned long long& cur, unsigned long long& end, unsigned long long* cursor, unsigned need,
                    unsigned slab) {
        if (cur + need > end) {
            const unsigned long long want = need > slab ? need : slab;
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(cursor, want);
            b = __shfl_sync(DP_FULL, b, 0);
            cur = b;
            end = b + want;
        }
        const unsigned long long off = cur;
        cur += need;
        return off;
    };
    for (int w = gwarp; w < nWin; w += nWarps) {
        if (lane == 0) F.slow[w] = 0;
        bool slow = false;
        for (int strand = 0; strand < 2 && !slow; strand++) {
            const int ws = 2 * w + strand;
            const int nc = candN[ws];
            if (nc == 0) continue;
            const int n = Q.wsN[ws];
            const unsigned qb = Q.wsOff[ws];
            int thr0 = n / 5;
            if (thr0 < 5) thr0 = 5;
            // the window strand's task slots
            const unsigned long long tb = take(tCur, tEnd, F.cursors + 0, (unsigned)nc, F.taskSlab);
            if (tb + (unsigned)nc > F.taskCap) {
                slow = true;
                break;
            }
            if (lane == 0) F.taskBase[ws] = (unsigned)tb;
            const bool qSmall = n <= DP_QCAP;
            DpListBuf LB;
            LB.qFirst = qSmall ? shFirst[wib] : S.qFirst + (size_t)gwarp * S.qStride;
            LB.qCnt = qSmall ? shCnt[wib] : S.qCnt + (size_t)gwarp * S.qStride;
            LB.qLo = qSmall ? shLo[wib] : S.qLo + (size_t)gwarp * S.qStride;
            LB.rqPos = qSmall ? shRqPos[wib] : S.rqPos + (size_t)gwarp * S.qStride;
            LB.rqId = qSmall ? shRqId[wib] : S.rqId + (size_t)gwarp * S.qStride;
            LB.shEnt = shEnt[wib];
            LB.shRsPos = shRsPos[wib];
            LB.shRsId = shRsId[wib];
            LB.gEnt = S.ent + (size_t)gwarp * S.sStride;
            LB.gRsPos = S.rsPos + (size_t)gwarp * S.sStride;
            LB.gRsId = S.rsId + (size_t)gwarp * S.sStride;
            LB.sStride = S.sStride;
            // window strands with at most 32 seeds (the bulk on small references) build their lists in registers
            const bool small = n <= 32;
            int firstR = 0, qPosR = 0;
            unsigned pbR = 0, peR = 0;
            if (small) {
                const unsigned sR = (int)lane < n ? Q.qSeed[qb + lane] : (0x80000000u | lane);  // padding lanes never match
                const unsigned mm = __match_any_sync(DP_FULL, sR);
                firstR = __ffs(mm) - 1;
                if ((int)lane < n) {
                    qPosR = Q.qPos[qb + lane];
                    if (firstR == (int)lane) {
                        pbR = __ldg(I.postOff + sR);
                        peR = __ldg(I.postOff + sR + 1);
                    }
                }
            }
            bool identDone = false;
            for (int ci = 0; ci < nc; ci++) {
                DpChainTask t;
                t.listOff = 0;
                t.nq = 0xffff;
                t.ns = 0;
                if ((int)candDistinct[(size_t)ws * candStride + ci] >= thr0) {
                    const unsigned c = candChunk[(size_t)ws * candStride + ci];
                    int nq, ns;
                    int* rsPos = nullptr;
                    unsigned short* rsId = nullptr;
                    unsigned rqWord = 0, rsWord = 0;
                    bool inRegs = small && dp_build_lists_small(I, n, c, firstR, qPosR, pbR, peR,
                                                                reinterpret_cast<unsigned*>(shEnt[wib]), nq, ns, rqWord, rsWord);
                    if (!inRegs) {
                        if (!identDone) {
                            dp_seed_identity(Q, qb, n, LB.qFirst);
                            identDone = true;
                        }
                        if (!dp_build_lists(I, Q, qb, n, c, LB, nq, ns, rsPos, rsId) || nq >= 0xffff || ns > 0xffff) {
                            slow = true;
                            break;
                        }
                    }
                    if (ns >= thr0 && nq >= thr0) {
                        const unsigned need = 2u * (unsigned)nq + (unsigned)ns;
                        const unsigned long long off = take(pCur, pEnd, F.cursors + 1, need, F.poolSlab);
                        if (off + need > F.poolCap) {
                            slow = true;
                            break;
                        }
                        unsigned* dst = F.pool + off;
                        if (inRegs) {
                            if ((int)lane < nq) {
                                dst[lane] = rqWord;
                                dst[nq + ns + lane] = 0;  // memo: chain length << 16 | last chunk-side index
                            }
                            if ((int)lane < ns) dst[nq + lane] = rsWord;
                        }
                        for (int x = lane; !inRegs && x < nq; x += 32) {
                            dst[x] = ((unsigned)LB.rqPos[x] << 16) | LB.rqId[x];
                            dst[nq + ns + x] = 0;  // memo: chain length << 16 | last chunk-side index
                        }
                        for (int x = lane; !inRegs && x < ns; x += 32) dst[nq + x] = ((unsigned)rsPos[x] << 16) | rsId[x];
                        t.listOff = (unsigned)off;
                        t.nq = (unsigned short)nq;
                        t.ns = (unsigned short)ns;
                    }
                    __syncwarp();
                }
                if (lane == 0) F.tasks[tb + ci] = t;
            }
        }
        if (slow && lane == 0) dp_mark_slow(F, w);
        __syncwarp();
    }
}

// dynamicMatch + extendChain (sequence.go:401-576) for one candidate, one thread. qe/se: packed reduced lists
// (position << 16 | seed identity); memo: per query entry chain length << 16 | last chunk-side index, zero on entry.
// Accepted chains -> ch[] as {len, firstA, lastA, firstB, lastB, ids}; returns their number or -1 if more than
// DP_FAST_CHAINS would have to be kept.
__device__ int dp_dynamic_match_serial(const unsigned* __restrict__ qe, int nq, int qScanLen,
                                       const unsigned* __restrict__ se, int ns, int sScanLen, unsigned* memo,
                                       int minMatch, int k, int (*ch)[6]) {
    if (minMatch == 0) minMatch = 1;
    int nGood = 0;
    int nilCount = nq;
#define QPOS(i) ((int)(qe[(i)] >> 16))
#define SPOS(i) ((int)(se[(i)] >> 16))
#define QID(i) (qe[(i)] & 0xffffu)
#define SID(i) (se[(i)] & 0xffffu)
#define GAPQ(i) (((i) + 1 < nq ? QPOS((i) + 1) : qScanLen) - QPOS(i) - k)
#define GAPS(i) (((i) + 1 < ns ? SPOS((i) + 1) : sScanLen) - SPOS(i) - k)
#define CLEN(i) ((int)(memo[(i)] >> 16))
#define LASTB(i) ((int)(memo[(i)] & 0xffffu))
    for (int qi = 0; qi <= nq - minMatch; qi++) {
        // sequence.go:409: internal to closely spaced repeats (cannot fire on reduced lists; kept for fidelity)
        if (qi > 0 && qi + 1 < nq && GAPQ(qi - 1) < 0 && GAPQ(qi) < 0 && QID(qi) == QID(qi - 1) && QID(qi) == QID(qi + 1))
            continue;
        if (CLEN(qi) != 0) continue;
        const unsigned qid = QID(qi);
        for (int si0 = 0; si0 < ns; si0++) {
            if (SID(si0) != qid || (si0 > 0 && SID(si0 - 1) == qid)) continue;
            // conditions evaluated with the state as it is now (minMatch may have grown)
            if (!(si0 <= ns - minMatch && (CLEN(qi) == 0 || LASTB(qi) != si0))) continue;
            if (CLEN(qi) == 0) nilCount--;
            memo[qi] = (1u << 16) | (unsigned)si0;
            // ---- extendChain (sequence.go:476-576) ----
            int curLen = 1;
            int ids = k;
            int lastA = qi, lastBi = si0;
            int offsetA = GAPQ(qi);
            int offsetB = GAPS(si0);
            int ai = qi + 1, bi = si0 + 1;
            bool done = false;
            while (!done && ai < nq && bi < ns) {
                int minB, maxB;
                if (offsetA < 0) {
                    minB = -k;
                    maxB = 0;
                } else {
                    minB = (offsetA * 2) / 3 - k;
                    maxB = (offsetA * 3) / 2 + k;
                }
                while (maxB < offsetB) {
                    offsetA += GAPQ(ai) + k;
                    ai++;
                    if (ai >= nq) {
                        done = true;
                        break;
                    }
                    minB = (offsetA * 2) / 3 - k;
                    maxB = (offsetA * 3) / 2 + k;
                }
                if (done) break;
                while (offsetB < minB) {
                    offsetB += GAPS(bi) + k;
                    bi++;
                    if (bi >= ns) {
                        done = true;
                        break;
                    }
                }
                if (done) break;
                int oldBi = bi, oldBOffset = offsetB;
                bool matched = false;
                const unsigned seedA = QID(ai);
                while (offsetB <= maxB) {
                    if (seedA == SID(bi)) {
                        if (CLEN(ai) != 0) {
                            if (bi == LASTB(ai) && CLEN(ai) > curLen) {
                                done = true;  // they have a better chain already
                                break;
                            }
                        } else {
                            nilCount--;
                        }
                        curLen++;
                        memo[ai] = ((unsigned)curLen << 16) | (unsigned)bi;
                        int d2 = SPOS(bi) - SPOS(lastBi) - k;  // GetBasesCovered, reference side
                        ids += k + (d2 < 0 ? d2 : 0);
                        lastA = ai;
                        lastBi = bi;
                        offsetA = GAPQ(ai);
                        offsetB = GAPS(bi);
                        ai++;
                        bi++;
                        matched = true;
                        break;
                    } else {
                        offsetB += GAPS(bi) + k;
                        bi++;
                        if (bi >= ns) break;
                    }
                }
                if (done) break;
                if (!matched) {
                    offsetA += GAPQ(ai) + k;
                    ai++;
                    offsetB = oldBOffset;
                    bi = oldBi;
                }
            }
            // ---- dynamicMatch bookkeeping (sequence.go:435-465) ----
            if (curLen >= minMatch) {
                int nextLength = (curLen * 2) / 3;
                if (nextLength > minMatch) {
                    minMatch = nextLength;
                    for (int j = nGood - 1; j >= 0; j--) {
                        if (ch[j][0] < nextLength) {
                            for (int z = 0; z < 6; z++) ch[j][z] = ch[nGood - 1][z];
                            nGood--;
                        }
                    }
                }
                if (nGood >= DP_FAST_CHAINS) return -1;
                ch[nGood][0] = curLen;
                ch[nGood][1] = qi;
                ch[nGood][2] = lastA;
                ch[nGood][3] = si0;
                ch[nGood][4] = lastBi;
                ch[nGood][5] = ids;
                nGood++;
                if (nilCount < curLen) return nGood;
            }
        }
    }
#undef QPOS
#undef SPOS
#undef QID
#undef SID
#undef GAPQ
#undef GAPS
#undef CLEN
#undef LASTB
    return nGood;
}

__global__ void __launch_bounds__(128) dp_chain_thread_kernel(DpIndexDev I, const DpWindow* __restrict__ wins,
                                                              const int* __restrict__ readLen, int nWin,
                                                              DpExtractOut Q, const int* __restrict__ candN,
                                                              const unsigned* __restrict__ candChunk,
                                                              const unsigned short* __restrict__ candDistinct,
                                                              int candStride, DpFastChain F, int* __restrict__ outN,
                                                              unsigned* __restrict__ outOff,
                                                              DpMappingDev* __restrict__ outMaps,
                                                              unsigned long long* __restrict__ outCursor,
                                                              unsigned long long outCapacity,
                                                              DpCounters* __restrict__ ctr) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = dp_lane();
    const int k = I.k;
    DpMappingDev results[DP_FAST_RESULTS];
    int nRes = 0;
    unsigned cells = 0;
    bool live = w < nWin && F.slow[w] == 0;
    bool giveUp = false;
    if (live) {
        DpWindow win = wins[w];
        const int L = win.len;
        if (L > 0) {
            const int rlen = readLen[win.read];
            const bool q2 = win.whole && ((L & 3) == 0);
            // SeedSequence.offset / inset of the window (Q3: SubSequence stores inset one too large)
            const int wOffset = win.whole ? 0 : win.start;
            const int wInset = win.whole ? 0 : (rlen - (win.start + L) + 1);
            int minMatches = Q.wsN[2 * w] / 5;
            int minRCMatches = Q.wsN[2 * w + 1] / 5;
            if (minMatches < 5) minMatches = 5;
            if (minRCMatches < 5) minRCMatches = 5;
            for (int strand = 0; strand < 2 && !giveUp; strand++) {
                const int ws = 2 * w + strand;
                const int nc = candN[ws];
                if (nc == 0) continue;
                const unsigned tb = F.taskBase[ws];
                const int qScanLen = strand == 0 ? (L - (q2 ? 4 : 0)) : (L - (q2 ? 3 : 0));
                for (int ci = 0; ci < nc && !giveUp; ci++) {
                    const int thr = strand == 0 ? minMatches : minRCMatches;
                    // CountIntersectionTo(...) < threshold (mapping.go:520-523, 559-562)
                    if ((int)candDistinct[(size_t)ws * candStride + ci] < thr) continue;
                    const DpChainTask t = F.tasks[tb + ci];
                    if (t.nq == 0xffff) continue;  // a Reduced list was shorter than the initial threshold already
                    const int nq = t.nq, ns = t.ns;
                    if (ns < thr || nq < thr) continue;  // Reduced returned nil
                    cells += (unsigned)(ns + nq);
                    const unsigned c = candChunk[(size_t)ws * candStride + ci];
                    const int sScanLen = __ldg(I.chunkScanLen + c);
                    const unsigned* qe = F.pool + t.listOff;
                    const unsigned* se = qe + nq;
                    unsigned* memo = F.pool + t.listOff + nq + ns;
                    int ch[DP_FAST_CHAINS][6];
                    const int nGood = dp_dynamic_match_serial(qe, nq, qScanLen, se, ns, sScanLen, memo, thr, k, ch);
                    if (nGood < 0) {
                        giveUp = true;
                        break;
                    }
                    // chains -> mappings (mapping.go:528-549 / 567-587), in allGoodChains order
                    const long long cOffset = __ldg(I.chunkOffset + c);
                    const long long cInset = __ldg(I.chunkInset + c);
                    for (int g = 0; g < nGood; g++) {
                        const int* r = ch[g];
                        const int sFirst = (int)(se[r[3]] >> 16), sLast = (int)(se[r[4]] >> 16);
                        long long start = cOffset + sFirst;
                        long long end = I.refLen - cInset - (long long)(sScanLen - sLast - k);
                        if (I.circular && start > I.refLen) start -= I.refLen;
                        int first = (int)(qe[r[1]] >> 16);                   // GetSeedOffset(MatchA[0])
                        int fromEnd = qScanLen - (int)(qe[r[2]] >> 16) - k;  // GetSeedOffsetFromEnd(MatchA[last])
                        if (first + fromEnd > (L * 2) / 3) continue;
                        if (nRes >= DP_FAST_RESULTS) {
                            giveUp = true;
                            break;
                        }
                        DpMappingDev mp;
                        mp.start = start;
                        mp.end = end;
                        if (strand == 0) {
                            mp.qOffset = first + wOffset;
                            mp.qInset = fromEnd + wInset;
                        } else {  // rcQuery.offset = window inset, rcQuery.inset = window offset
                            mp.qInset = first + wInset;
                            mp.qOffset = fromEnd + wOffset;
                        }
                        mp.ids = r[5];
                        mp.rc = strand;
                        results[nRes++] = mp;
                        int limit = (r[0] * 4) / 5;
                        if (strand == 0) {
                            if (limit > minMatches) minMatches = limit;
                            if (limit > minRCMatches) minRCMatches = limit;
                        } else {
                            if (limit > minRCMatches) minRCMatches = limit;
                        }
                    }
                }
            }
            if (giveUp) {  // the general path recomputes this window from the candidates
                dp_mark_slow(F, w);
                nRes = 0;
                cells = 0;
            } else if (nRes > 1) {
                // ---- sort by Start + overlap dedupe (mapping.go:590-608) ----
                for (int i = 1; i < nRes; i++) {  // stable insertion sort (= Go's sort.Sort for n <= 12)
                    DpMappingDev x = results[i];
                    int j = i;
                    while (j > 0 && x.start < results[j - 1].start) {
                        results[j] = results[j - 1];
                        j--;
                    }
                    results[j] = x;
                }
                for (int i = nRes - 1; i > 0; i--) {
                    DpMappingDev ra = results[i - 1], rb = results[i];
                    if (ra.rc == rb.rc && rb.start < ra.end) {
                        if (ra.end - ra.start > rb.end - rb.start) {
                            results[i] = results[nRes - 1];
                            nRes--;
                        } else {
                            results[i - 1] = results[i];
                            results[i] = results[nRes - 1];
                            nRes--;
                        }
                    }
                }
            }
        }
    }
    // compact output: one bump allocation per warp
    const bool writes = live && !giveUp;
    int incl = writes ? nRes : 0;
    const int mine = incl;
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(DP_FULL, incl, d);
        if ((int)lane >= d) incl += y;
    }
    const int warpTot = __shfl_sync(DP_FULL, incl, 31);
    unsigned long long warpBase = 0;
    if (lane == 31 && warpTot) warpBase = atomicAdd(outCursor, (unsigned long long)warpTot);
    warpBase = __shfl_sync(DP_FULL, warpBase, 31);
    unsigned wCells = cells, wMaps = 0;
    if (writes) {
        unsigned long long base = warpBase + (unsigned)(incl - mine);
        int cnt = mine;
        if (base + (unsigned)cnt > outCapacity) {
            atomicOr(&ctr->overflow, DP_OV_OUTPOOL);
            cnt = 0;
            base = 0;
        }
        outN[w] = cnt;
        outOff[w] = (unsigned)base;
        for (int i = 0; i < cnt; i++) outMaps[base + i] = results[i];
        wMaps = (unsigned)cnt;
    }
    for (int d = 16; d > 0; d >>= 1) {
        wCells += __shfl_xor_sync(DP_FULL, wCells, d);
        wMaps += __shfl_xor_sync(DP_FULL, wMaps, d);
    }
    if (lane == 0 && (wCells | wMaps)) {
        atomicAdd(&ctr->chain_cells, (unsigned long long)wCells);
        atomicAdd(&ctr->mappings, (unsigned long long)wMaps);
    }
}
