// downpore_b200 — map-side kernels (sm_100a): seed extraction, seed-index lookup, chaining.
// One performMapping() call (mapping/mapping.go:489-611) = one DpWindow; the three kernels below are its stages.
#pragma once
#include "dp_common.cuh"

// ===============================================================================================================
// Stage 1 — seed extraction: SeedIndex.NewSeedSequence on the window and on its reverse complement
// (seeds/seeds.go:33-50; the asm scans sequence/asm_amd64.s:81-394).
//
// One warp per window. Iteration t looks at forward k-mer positions j = 32t + lane: the k-mer comes from two packed
// words and a funnel shift, its reverse complement from brev; both are looked up in the 8-byte {flags, rank} table.
// The reverse-complement strand visits the same positions backwards (rc position = L-k-j), so a single pass serves
// both strands. Pass 1 stores the two ballot masks per iteration in shared memory and counts; one atomicAdd
// allocates the window's slice of the compact output; pass 2 re-gathers only the hits and writes them in order.
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

__global__ void __launch_bounds__(256) dp_extract_kernel(DpIndexDev I, const unsigned* __restrict__ readWords,
                                                         const long long* __restrict__ readWordOff,
                                                         const DpWindow* __restrict__ wins, int nWin, DpExtractOut O,
                                                         int maskWords, DpCounters* __restrict__ ctr) {
    extern __shared__ unsigned dp_smem[];
    const unsigned lane = dp_lane();
    const unsigned lt = dp_lanemask_lt();
    const int warpInBlock = threadIdx.x >> 5;
    unsigned* mF = dp_smem + (size_t)warpInBlock * 2 * maskWords;
    unsigned* mR = mF + maskWords;
    const int k = I.k;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nWarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long lookups = 0, seeds = 0;
    for (int w = warp; w < nWin; w += nWarps) {
        DpWindow win = wins[w];
        const unsigned* words = readWords + readWordOff[win.read];
        const int L = win.len;
        const bool q2 = win.whole && ((L & 3) == 0);
        const int nJ = L - k + 1 - (q2 ? 4 : 0);  // forward positions 0..nJ-1 are visited by both strands
        int nF = 0, nR = 0;
        int nIter = (nJ + 31) >> 5;
        for (int t = 0; t < nIter; t++) {
            int j = t * 32 + (int)lane;
            bool hf = false, hr = false;
            if (j < nJ) {
                unsigned kmer = dp_kmer_at(words, (long long)win.start + j, k);
                hf = dp_seed_flag(I.table, kmer);
                hr = dp_seed_flag(I.table, dp_revcomp(kmer, k));
            }
            unsigned bf = __ballot_sync(DP_FULL, hf);
            unsigned br = __ballot_sync(DP_FULL, hr);
            if (lane == 0) {
                mF[t] = bf;
                mR[t] = br;
            }
            nF += __popc(bf);
            nR += __popc(br);
        }
        __syncwarp();
        // Q2 on the rc strand: its first visited k-mer (forward position nJ-1) is visited twice
        int dup = 0;
        if (q2 && nJ > 0) dup = (mR[(nJ - 1) >> 5] >> ((nJ - 1) & 31)) & 1;
        unsigned base = 0;
        if (lane == 0) base = (unsigned)atomicAdd(O.cursor, (unsigned long long)(nF + nR + dup));
        base = __shfl_sync(DP_FULL, base, 0);
        if (lane == 0) {
            O.wsOff[2 * w] = base;
            O.wsN[2 * w] = nF;
            O.wsOff[2 * w + 1] = base + nF;
            O.wsN[2 * w + 1] = nR + dup;
        }
        const unsigned baseR = base + nF;
        const int rcShift = q2 ? 3 : 0;  // rc scan position = (L-k-j) - rcShift
        int cumF = 0, cumR = 0;
        for (int t = 0; t < nIter; t++) {
            unsigned bf = mF[t], br = mR[t];
            if ((bf | br) != 0) {
                int j = t * 32 + (int)lane;
                bool hf = (bf >> lane) & 1, hr = (br >> lane) & 1;
                if (hf | hr) {
                    unsigned kmer = dp_kmer_at(words, (long long)win.start + j, k);
                    unsigned rank;
                    if (hf) {
                        dp_seed_lookup(I.table, kmer, &rank);
                        unsigned idx = base + cumF + __popc(bf & lt);
                        O.qSeed[idx] = rank;
                        O.qPos[idx] = j;
                    }
                    if (hr) {
                        dp_seed_lookup(I.table, dp_revcomp(kmer, k), &rank);
                        int below = cumR + __popc(br & lt);        // rc hits at smaller forward positions
                        unsigned idx = baseR + dup + (nR - 1 - below);  // rc order is descending in j
                        O.qSeed[idx] = rank;
                        O.qPos[idx] = (L - k - j) - rcShift;
                        if (dup && j == nJ - 1) {  // the double visit: scan positions 0 and 1
                            O.qSeed[baseR] = rank;
                            O.qPos[baseR] = 0;
                        }
                    }
                }
                cumF += __popc(bf);
                cumR += __popc(br);
            }
        }
        __syncwarp();
        lookups += 2ull * (unsigned)(nJ > 0 ? nJ : 0) + (q2 ? 1u : 0u);  // Q2: the rc scan visits one k-mer twice
        seeds += (unsigned)(nF + nR + dup);
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
struct DpLookupScratch {  // per-warp global scratch, `stride` entries per array
    unsigned* eSeed;
    unsigned* eOff;
    unsigned* ePre;    // exclusive prefix of posting run lengths (stride+1)
    unsigned* eEndW;   // last 64-chunk word of each included run (0 if empty), lens = eEndW+1
    unsigned char* eFirst;  // first occurrence of the seed among E
    unsigned* allSeeds;     // seeds present in every chunk (|D| >= C)
    unsigned short* order;  // Q6 column-order simulation
    unsigned* counters;     // [C] when C does not fit shared memory
    int stride;
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

__global__ void __launch_bounds__(128) dp_lookup_kernel(DpIndexDev I, DpExtractOut Q, int nWS, DpLookupScratch S,
                                                        int countersInSmem, int* __restrict__ candN,
                                                        unsigned* __restrict__ candChunk,
                                                        unsigned short* __restrict__ candDistinct, int candStride,
                                                        DpCounters* __restrict__ ctr) {
    extern __shared__ unsigned dp_smem[];
    const unsigned lane = dp_lane();
    const unsigned lt = dp_lanemask_lt();
    const int warpInBlock = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned C = I.numChunks;
    unsigned* cnt = countersInSmem ? dp_smem + (size_t)warpInBlock * C : S.counters + (size_t)gwarp * C;
    const size_t so = (size_t)gwarp * S.stride;
    unsigned* eSeed = S.eSeed + so;
    unsigned* eOff = S.eOff + so;
    unsigned* ePre = S.ePre + (size_t)gwarp * (S.stride + 1);
    unsigned* eEndW = S.eEndW + so;
    unsigned char* eFirst = S.eFirst + so;
    unsigned* allSeeds = S.allSeeds + so;
    unsigned short* order = S.order + so;
    unsigned long long cRuns = 0, cEntries = 0, cCand = 0;

    for (int ws = gwarp; ws < nWS; ws += nWarps) {
        const int n = Q.wsN[ws];
        const unsigned qb = Q.wsOff[ws];
        int nCandOut = 0;
        unsigned* outChunk = candChunk + (size_t)ws * candStride;
        unsigned short* outDist = candDistinct + (size_t)ws * candStride;
        if (n >= 5) {
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
                // ---- distinct seeds present in every chunk ----
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
                // ---- first occurrences among E, last word of each run, prefix of run lengths ----
                unsigned total = 0;
                for (int j0 = 0; j0 < nInc; j0 += 32) {
                    int j = j0 + (int)lane;
                    unsigned c = 0;
                    if (j < nInc) {
                        unsigned s = eSeed[j];
                        c = ePre[j];
                        bool first = true;
                        for (int b = 0; b < j; b++)
                            if (eSeed[b] == s) {
                                first = false;
                                break;
                            }
                        eFirst[j] = first ? 1 : 0;
                        eEndW[j] = c ? (__ldg(I.seedChunks + eOff[j] + c - 1) >> 6) : 0u;
                    }
                    // warp inclusive scan of c
                    unsigned x = c;
                    for (int d = 1; d < 32; d <<= 1) {
                        unsigned y = __shfl_up_sync(DP_FULL, x, d);
                        if ((int)lane >= d) x += y;
                    }
                    __syncwarp();
                    if (j < nInc) ePre[j] = total + x - c;
                    total += __shfl_sync(DP_FULL, x, 31);
                }
                if (lane == 0) ePre[nInc] = total;
                // ---- clear counters ----
                for (unsigned c = lane; c < C; c += 32) cnt[c] = 0;
                __syncwarp();
                // ---- gather the posting runs: the HBM/L2 gather this kernel is about ----
                for (unsigned p0 = 0; p0 < total; p0 += 32) {
                    unsigned p = p0 + lane;
                    if (p < total) {
                        int lo = 0, hi = nInc;  // largest j with ePre[j] <= p
                        while (hi - lo > 1) {
                            int mid = (lo + hi) >> 1;
                            if (ePre[mid] <= p) lo = mid;
                            else hi = mid;
                        }
                        unsigned chunk = __ldg(I.seedChunks + eOff[lo] + (p - ePre[lo]));
                        atomicAdd(cnt + chunk, 1u + ((unsigned)eFirst[lo] << 16));
                    }
                }
                __syncwarp();
                cRuns += (unsigned)nInc;
                cEntries += total;
                // ---- threshold, ascending chunk id ----
                int simWord = -1, simLive = nInc;  // Q6 column-order simulation state (lane 0 owns it)
                bool simInit = false;
                for (unsigned c0 = 0; c0 < C; c0 += 32) {
                    unsigned c = c0 + lane;
                    unsigned v = (c < C) ? cnt[c] : 0;
                    int soft = (int)(v & 0xffffu);
                    bool pass = soft >= T;
                    unsigned mp = __ballot_sync(DP_FULL, pass);
                    if (mp && (clamped || q6)) {
                        // refine borderline lanes one at a time (rare)
                        unsigned todo = mp;
                        while (todo) {
                            int l = __ffs(todo) - 1;
                            todo &= todo - 1;
                            unsigned cc = c0 + l;
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
                                // advance the drop simulation of bitset.go:332-353 to word `wword`
                                if (lane == 0) {
                                    if (!simInit) {
                                        for (int j = 0; j < nInc; j++) order[j] = (unsigned short)j;
                                        // start = min over sets of IntSet.start (1 for an empty set)
                                        unsigned st = 0xffffffffu;
                                        for (int j = 0; j < nInc; j++) {
                                            unsigned len = ePre[j + 1] - ePre[j];
                                            unsigned s0 = len ? (__ldg(I.seedChunks + eOff[j]) >> 6) : 1u;
                                            if (s0 < st) st = s0;
                                        }
                                        simWord = (int)st - 1;
                                    }
                                    for (int i = simWord + 1; i <= (int)wword; i++) {
                                        int t = 0;
                                        while (t < simLive) {
                                            if (eEndW[order[t]] + 1 <= (unsigned)i) {
                                                order[t] = order[simLive - 1];
                                                simLive--;
                                            } else {
                                                t++;
                                            }
                                        }
                                    }
                                    if ((int)wword > simWord) simWord = (int)wword;
                                }
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
                        pass = (mp >> lane) & 1;
                    }
                    if (pass) {
                        int idx = nCandOut + __popc(mp & lt);
                        if (idx < candStride) {
                            outChunk[idx] = c;
                            outDist[idx] = (unsigned short)((v >> 16) + nAllDistinct);
                        }
                    }
                    nCandOut += __popc(mp);
                }
                if (nCandOut > candStride) {
                    if (lane == 0) atomicOr(&ctr->overflow, 4u);
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
// Stage 3 — chaining: the candidate loop of performMapping (mapping/mapping.go:518-608) with
// SeedSequence.Match -> Reduced -> dynamicMatch -> extendChain (seeds/sequence.go:85-123, 361-576) and the
// coordinate arithmetic of GetSeedOffset / GetSeedOffsetFromEnd / GetBasesCovered (sequence.go:830-858, 1239-1276)
// done on scan positions. One warp per window: candidates are visited in ascending chunk id, forward strand first,
// because the thresholds minMatches / minRCMatches escalate as chains are accepted (Q12). The warp builds both
// reduced lists cooperatively (ordered ballot compaction); lane 0 runs the greedy chainer.
// ===============================================================================================================
struct DpChainScratch {  // per-warp global scratch
    unsigned* hashKey;        // [hashSize] query seed ranks (0xffffffff = empty)
    unsigned char* hashFlag;  // [hashSize] bit0: present in the current chunk
    unsigned* rqSeed;         // reduced query [qStride]
    int* rqPos;
    unsigned* rsSeed;         // reduced chunk [sStride]
    int* rsPos;
    int* chainLen;            // memo per reduced query seed: 0 = nil  [qStride]
    int* lastB;               // memo: last chunk index of the chain through this query seed [qStride]
    int* chains;              // accepted chains of the current candidate: 6 ints each [chainCap*6]
    DpMappingDev* results;    // window results before sort/dedupe [resultCap]
    int hashSize;
    int qStride;
    int sStride;
    int chainCap;
    int resultCap;
};

__device__ __forceinline__ unsigned dp_hash(unsigned s) { return s * 2654435761u; }

// lane 0 only. Reduced lists: (qs,qp)[nq] query, (ss,sp)[ns] chunk. Accepted chains -> ch[] as
// {len, firstA, lastA, firstB, lastB, ids}; returns their number (or -1 on chain list overflow).
__device__ int dp_dynamic_match(const unsigned* qs, const int* qp, int nq, int qScanLen, const unsigned* ss,
                                const int* sp, int ns, int sScanLen, int minMatch, int k, int* chainLen, int* lastB,
                                int* ch, int chainCap) {
    if (minMatch == 0) minMatch = 1;
    int nGood = 0;
    int nilCount = nq;
    for (int x = 0; x < nq; x++) chainLen[x] = 0;
#define GAPQ(i) (((i) + 1 < nq ? qp[(i) + 1] : qScanLen) - qp[(i)] - k)
#define GAPS(i) (((i) + 1 < ns ? sp[(i) + 1] : sScanLen) - sp[(i)] - k)
    for (int qi = 0; qi <= nq - minMatch; qi++) {
        if (qi > 0 && qi + 1 < nq && GAPQ(qi - 1) < 0 && GAPQ(qi) < 0 && qs[qi] == qs[qi - 1] && qs[qi] == qs[qi + 1])
            continue;  // sequence.go:409 (cannot fire on reduced lists; kept for fidelity)
        if (chainLen[qi] != 0) continue;
        unsigned prevSeed = 0xffffffffu;
        const unsigned qseed = qs[qi];
        for (int si = 0; si <= ns - minMatch; si++) {
            unsigned nextSeed = ss[si];
            if (nextSeed == qseed && nextSeed != prevSeed && (chainLen[qi] == 0 || lastB[qi] != si)) {
                if (chainLen[qi] == 0) nilCount--;
                chainLen[qi] = 1;
                lastB[qi] = si;
                // ---- extendChain (sequence.go:476-576) ----
                int curLen = 1;
                int ids = k;
                int lastA = qi, lastBi = si;
                int offsetA = GAPQ(qi);
                int offsetB = GAPS(si);
                int ai = qi + 1, bi = si + 1;
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
                    unsigned seedA = qs[ai];
                    while (offsetB <= maxB) {
                        if (seedA == ss[bi]) {
                            if (chainLen[ai] != 0) {
                                if (bi == lastB[ai] && chainLen[ai] > curLen) {
                                    done = true;  // they have a better chain already
                                    break;
                                }
                            } else {
                                nilCount--;
                            }
                            curLen++;
                            chainLen[ai] = curLen;
                            lastB[ai] = bi;
                            int d2 = sp[bi] - sp[lastBi] - k;  // GetBasesCovered: overlap on the reference side
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
                    if (nGood >= chainCap) return -1;
                    int* r = ch + nGood * 6;
                    r[0] = curLen;
                    r[1] = qi;
                    r[2] = lastA;
                    r[3] = si;
                    r[4] = lastBi;
                    r[5] = ids;
                    nGood++;
                    if (nilCount < curLen) return nGood;
                }
            }
            prevSeed = nextSeed;
        }
    }
#undef GAPQ
#undef GAPS
    return nGood;
}

__global__ void __launch_bounds__(128) dp_chain_kernel(DpIndexDev I, const DpWindow* __restrict__ wins,
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
    const unsigned lane = dp_lane();
    const unsigned lt = dp_lanemask_lt();
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    const int k = I.k;
    unsigned* hashKey = S.hashKey + (size_t)gwarp * S.hashSize;
    unsigned char* hashFlag = S.hashFlag + (size_t)gwarp * S.hashSize;
    unsigned* rqSeed = S.rqSeed + (size_t)gwarp * S.qStride;
    int* rqPos = S.rqPos + (size_t)gwarp * S.qStride;
    unsigned* rsSeed = S.rsSeed + (size_t)gwarp * S.sStride;
    int* rsPos = S.rsPos + (size_t)gwarp * S.sStride;
    int* chainLen = S.chainLen + (size_t)gwarp * S.qStride;
    int* lastB = S.lastB + (size_t)gwarp * S.qStride;
    int* chains = S.chains + (size_t)gwarp * S.chainCap * 6;
    DpMappingDev* results = S.results + (size_t)gwarp * S.resultCap;
    unsigned long long cCells = 0, cMaps = 0;

    for (int w = gwarp; w < nWin; w += nWarps) {
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
            // ---- hash set of the strand's seeds (open addressing, at most half full) ----
            int hsize = 64;
            while (hsize < 2 * n) hsize <<= 1;
            const unsigned hmask = (unsigned)hsize - 1;
            for (int h = lane; h < hsize; h += 32) hashKey[h] = 0xffffffffu;
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
                unsigned s = Q.qSeed[qb + j];
                unsigned h = dp_hash(s) & hmask;
                for (;;) {
                    unsigned old = atomicCAS(hashKey + h, 0xffffffffu, s);
                    if (old == 0xffffffffu || old == s) break;
                    h = (h + 1) & hmask;
                }
            }
            __syncwarp();
            for (int ci = 0; ci < nc; ci++) {
                const int thr = strand == 0 ? minMatches : minRCMatches;
                // 1. CountIntersectionTo(...) < threshold (mapping.go:520-523, 559-562)
                if ((int)candDistinct[(size_t)ws * candStride + ci] < thr) continue;
                const unsigned c = candChunk[(size_t)ws * candStride + ci];
                const unsigned cb = __ldg(I.chunkOff + c);
                const int cn = (int)(__ldg(I.chunkOff + c + 1) - cb);
                // 2. chunk.Reduced(querySet) (sequence.go:85-123): members of the query set, same-as-previous collapsed
                for (int h = lane; h < hsize; h += 32) hashFlag[h] = 0;
                __syncwarp();
                int ns = 0;
                unsigned prevMember = 0xffffffffu;
                for (int e0 = 0; e0 < cn; e0 += 32) {
                    int e = e0 + (int)lane;
                    unsigned s = 0xffffffffu;
                    int pos = 0;
                    bool member = false;
                    if (e < cn) {
                        s = __ldg(I.chunkSeed + cb + e);
                        pos = __ldg(I.chunkPos + cb + e);
                        unsigned h = dp_hash(s) & hmask;
                        for (;;) {
                            unsigned key = hashKey[h];
                            if (key == s) {
                                member = true;
                                hashFlag[h] = 1;
                                break;
                            }
                            if (key == 0xffffffffu) break;
                            h = (h + 1) & hmask;
                        }
                    }
                    unsigned mm = __ballot_sync(DP_FULL, member);
                    unsigned lower = mm & lt;
                    int src = lower ? 31 - __clz(lower) : 0;
                    unsigned ps = __shfl_sync(DP_FULL, s, src);
                    if (!lower) ps = prevMember;
                    bool keep = member && s != ps;
                    unsigned mk = __ballot_sync(DP_FULL, keep);
                    if (keep) {
                        int idx = ns + __popc(mk & lt);
                        rsSeed[idx] = s;
                        rsPos[idx] = pos;
                    }
                    ns += __popc(mk);
                    if (mm) prevMember = __shfl_sync(DP_FULL, s, 31 - __clz(mm));
                }
                __syncwarp();
                // 3. query.Reduced(chunkSet): both Reduced calls run before the nil test (sequence.go:366-374)
                int nq = 0;
                prevMember = 0xffffffffu;
                for (int j0 = 0; j0 < n; j0 += 32) {
                    int j = j0 + (int)lane;
                    unsigned s = 0xffffffffu;
                    int pos = 0;
                    bool member = false;
                    if (j < n) {
                        s = Q.qSeed[qb + j];
                        pos = Q.qPos[qb + j];
                        unsigned h = dp_hash(s) & hmask;
                        for (;;) {
                            unsigned key = hashKey[h];
                            if (key == s) {
                                member = hashFlag[h] != 0;
                                break;
                            }
                            h = (h + 1) & hmask;
                        }
                    }
                    unsigned mm = __ballot_sync(DP_FULL, member);
                    unsigned lower = mm & lt;
                    int src = lower ? 31 - __clz(lower) : 0;
                    unsigned ps = __shfl_sync(DP_FULL, s, src);
                    if (!lower) ps = prevMember;
                    bool keep = member && s != ps;
                    unsigned mk = __ballot_sync(DP_FULL, keep);
                    if (keep) {
                        int idx = nq + __popc(mk & lt);
                        rqSeed[idx] = s;
                        rqPos[idx] = pos;
                    }
                    nq += __popc(mk);
                    if (mm) prevMember = __shfl_sync(DP_FULL, s, 31 - __clz(mm));
                }
                __syncwarp();
                if (ns < thr || nq < thr) continue;  // Reduced returned nil
                cCells += (unsigned)(ns + nq);
                // 4. dynamicMatch on lane 0
                const int sScanLen = __ldg(I.chunkScanLen + c);
                int nGood = 0;
                if (lane == 0)
                    nGood = dp_dynamic_match(rqSeed, rqPos, nq, qScanLen, rsSeed, rsPos, ns, sScanLen, thr, k, chainLen,
                                             lastB, chains, S.chainCap);
                nGood = __shfl_sync(DP_FULL, nGood, 0);
                if (nGood < 0) {
                    overflow = true;
                    if (lane == 0) atomicOr(&ctr->overflow, 2u);
                    nGood = 0;
                }
                __syncwarp();
                // 5. chains -> mappings (mapping.go:528-549 / 567-587), in allGoodChains order; lane 0 keeps the state
                if (lane == 0) {
                    const long long cOffset = __ldg(I.chunkOffset + c);
                    const long long cInset = __ldg(I.chunkInset + c);
                    for (int g = 0; g < nGood; g++) {
                        const int* r = chains + g * 6;
                        long long start = cOffset + rsPos[r[3]];
                        long long end = I.refLen - cInset - (long long)(sScanLen - rsPos[r[4]] - k);
                        if (I.circular && start > I.refLen) start -= I.refLen;
                        int first = rqPos[r[1]];                      // GetSeedOffset(MatchA[0])
                        int fromEnd = qScanLen - rqPos[r[2]] - k;     // GetSeedOffsetFromEnd(MatchA[last])
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
        }
        // ---- sort by Start + overlap dedupe (mapping.go:590-608); lane 0 ----
        if (lane == 0) {
            if (nRes > S.resultCap) {
                atomicOr(&ctr->overflow, 1u);
                nRes = S.resultCap;
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
                atomicOr(&ctr->overflow, 1u);
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
