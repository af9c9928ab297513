// downpore_b200 — kernels of the `overlap` path (sm_100a): one round of `downpore overlap` up to the stream of seed
// matches (commands/overlap.go:115-160): PrepareQueries (overlap/overlap.go:157-215), AddSequences (:218-318) and
// FindOverlaps / matchWorker (:320-387).
//
// Where the reference leaves an order to the goroutine scheduler the canonical choice is num_workers = 1 (the order
// the test oracle states): reads in file order, chunks numbered in emission order, queries in slice order with the
// forward query in front of its reverse complement, candidates in ascending chunk order.
//
// Data in HBM for one round
//   reads        : packed words, every read on a 16-base word boundary (the sequence set, himem cached: every read
//                  reaches a round as cached[id].SubSequence(0, Len()), sequence/seqio.go:118-125)
//   seed table   : uint2[4^k/32] {32 flags, seeds before this word}; a seed's device id is its RANK among the seed k-mers
//                  (ids are labels in the reference: only equality is ever tested); regOfRank[] gives the reference's id
//                  (registration order, seeds/seeds.go:131-154) for everything that is exported
//   read seeds   : CSR (rOff[n+1], rPos[], rSeed[]) of every seed occurrence of every non-ignored read in scan order plus
//                  one sentinel per read (pos = read length): the gap after a seed is pos[i+1] - pos[i] - k everywhere
//   chunks       : seed-space pieces of the reads (overlap.go:253-318) as views into the read seeds {first, n, ...}
//   seed -> chunks : CSR of the distinct chunks containing each seed, ascending (SeedIndex.sequenceSets)
//   queries      : 2 per slice (forward, reverse complement): CSR (qOff, qPos, qSeed) + sorted distinct seeds + the slot
//                  of every seed among them
#pragma once
#include "dp_common.cuh"
#include "dp_index.cuh"
#include "dp_map.cuh"

#define OV_MAXSLICE 8192   // bases of one query slice (< 2 * overlap_size)
#define OV_MAXBLK 512      // k-blocks of one slice in AddSeeds (a slice of 8192 bases holds at most 8192 / 18)
#define OV_MAXN 256        // num_seeds
#define OV_QMAX 512        // seeds of one query
#define OV_AMAX 512        // reduced query seeds PairwiseAlignments can hold: seedAligner.reduced has overlap/2 ints
#define OV_OPEN 500        // seedAligner.open (seeds/alignment.go:299)

struct OvParams {
    int k;
    int overlap;
    int numSeeds;
    int seedLimit;
    int chunkSize;
    int queryBatch;
    int redCap;  // overlap / 2 = len(seedAligner.reduced)
    int pad;
    double hitFraction;
};

struct OvSlice {
    int read;
    int start;
    int len;
    int pad;
};

struct OvChunk {  // one SeedSequence handed to index.AddSequence
    unsigned first;  // absolute index of its first seed in rPos / rSeed
    unsigned n;      // seeds
    int read;
    int length;      // SeedSequence.length (bases)
    int offset;      // SeedSequence.offset
    int inset;       // SeedSequence.inset
};

// header of the selection kernel's output
struct OvSelectOut {
    int size;      // index.Size(): registered seeds
    int nSlices;   // cached slices = queries / 2
    int lastRead;  // SequenceID of the last slice (-1: none)
    int sent;      // reads taken from GetNSequencesFrom
};

__device__ __forceinline__ unsigned ov_ldcg(const unsigned* p) { return __ldcg(p); }

// ---------------------------------------------------------------------------------------------------------------
// PrepareQueries' seed selection: getEdges (overlap/overlap.go:56-94) feeding AddSeeds (seeds/seeds.go:62-156), one
// worker. Inherently ordered: a slice's choice depends on the seed table as the previous slices left it, and the loop
// ends when index.Size() reaches seed_batch_size. One warp walks the reads; inside a slice the work is parallel:
//   A  seed flags of all k-mer positions of the slice -> bit masks in shared memory (lanes own 32 positions each);
//   B  the k-block walk (all lanes, uniform): a block = the k k-mers starting at p0+1..p0+k; a block that contains a
//      seed "resets" (the next block starts 2k behind that seed), else the next block starts 3k further on;
//   C  best-valued k-mer (first maximum, value > 0) of every block that did not reset, one lane per block;
//   D  topN by insertion (seeds.go:108-120) = the numSeeds best blocks by (value desc, block asc), listed ascending by
//      value with later blocks first among equals; unfilled slots keep k-mer 0 at the front;
//   E  registration of the list, every k-mer followed by its reverse complement (seeds.go:131-154): duplicates inside
//      the list by __match_any_sync, ids by ballot prefix.
// ---------------------------------------------------------------------------------------------------------------
#define OV_SEL_THREADS 256
__global__ void __launch_bounds__(OV_SEL_THREADS) ov_select_kernel(const unsigned* __restrict__ words,
                                                                   const long long* __restrict__ readBase,
                                                                   const int* __restrict__ readLen,
                                                                   const unsigned char* __restrict__ ignore, int nReads,
                                                                   int firstSequence, OvParams P,
                                                                   const double* __restrict__ values, unsigned* bits,
                                                                   unsigned* __restrict__ regKmer, int regCap,
                                                                   OvSlice* __restrict__ slices, int sliceCap,
                                                                   OvSelectOut* __restrict__ out,
                                                                   unsigned* __restrict__ err) {
    __shared__ unsigned sw[2][OV_MAXSLICE / 16 + 4];       // the packed words of the read's (one or two) slices
    __shared__ unsigned shMask[2][OV_MAXSLICE / 32 + 2];   // seed flags of their k-mer positions
    __shared__ int shBlk[OV_MAXBLK];
    __shared__ double shVal[OV_MAXBLK];
    __shared__ unsigned shKmer[OV_MAXBLK];
    __shared__ double shV2[OV_MAXSLICE / 3 + 16];          // value of every position of every kept block
    __shared__ unsigned shTop[OV_MAXN];
    __shared__ unsigned shNew[2 * OV_MAXN];                // k-mers the read's first slice registered
    __shared__ int shNb, shNewN, shSize;
    const unsigned tid = threadIdx.x, lane = dp_lane(), wib = tid >> 5;
    const unsigned lt = dp_lanemask_lt();
    const int k = P.k;
    const unsigned kShift = 32u - 2u * (unsigned)k;
    int size = 0, nSlices = 0, sent = 0, lastRead = -1;
    for (int id = firstSequence; id < nReads && sent < P.queryBatch; id++) {
        if (ignore && ignore[id]) continue;
        sent++;
        if (size >= P.seedLimit) break;
        const int rlen = readLen[id];
        const int nsl = rlen < P.overlap * 2 ? 1 : 2;
        const int len = nsl == 1 ? rlen : P.overlap;
        if (nSlices + nsl > sliceCap || len > OV_MAXSLICE) {
            if (tid == 0) atomicOr(err, 1u);
            break;
        }
        const int nPos = len - k + 1;
        const int nW = nPos > 0 ? (nPos + 31) >> 5 : 0;
        __syncthreads();
        // ---- the slices' words and the seed flags of all their positions (the table as the previous read left it) ----
        for (int sl = 0; sl < nsl; sl++) {
            const long long base = readBase[id] + (sl ? rlen - P.overlap : 0);
            const int nw = (int)(((base & 15) + len + 15) >> 4) + 1;
            for (int i = (int)tid; i < nw; i += OV_SEL_THREADS) sw[sl][i] = __ldg(words + (base >> 4) + i);
            if (tid < 2) shMask[sl][nW + tid] = 0;
        }
        if (tid == 0) {
            for (int sl = 0; sl < nsl; sl++) {
                OvSlice s;
                s.read = id;
                s.start = sl ? rlen - P.overlap : 0;
                s.len = len;
                s.pad = 0;
                slices[nSlices + sl] = s;
            }
            shNewN = 0;
        }
        __syncthreads();
        for (int sl = 0; sl < nsl; sl++) {
            const unsigned bo = (unsigned)((readBase[id] + (sl ? rlen - P.overlap : 0)) & 15);
            for (int p0 = 0; p0 < nPos; p0 += OV_SEL_THREADS) {
                const int p = p0 + (int)tid;
                unsigned f = 0;
                if (p < nPos) {
                    const unsigned o = bo + (unsigned)p;
                    const unsigned km = __funnelshift_l(sw[sl][(o >> 4) + 1], sw[sl][o >> 4], (o & 15u) * 2u) >> kShift;
                    f = (ov_ldcg(bits + (km >> 5)) >> (km & 31)) & 1u;
                }
                const unsigned m = __ballot_sync(DP_FULL, f);
                if (lane == 0 && p0 + (int)(wib * 32) < nPos) shMask[sl][(p0 >> 5) + wib] = m;
            }
        }
        __syncthreads();
        for (int sl = 0; sl < nsl; sl++) {
            const unsigned bo = (unsigned)((readBase[id] + (sl ? rlen - P.overlap : 0)) & 15);
            auto kmer_at = [&](int p) -> unsigned {
                const unsigned o = bo + (unsigned)p;
                return __funnelshift_l(sw[sl][(o >> 4) + 1], sw[sl][o >> 4], (o & 15u) * 2u) >> kShift;
            };
            if (sl == 1) {
                // the first slice's seeds: flag their occurrences in the second slice
                const int nn = shNewN;
                for (int p = (int)tid; p < nPos; p += OV_SEL_THREADS) {
                    const unsigned km = kmer_at(p);
                    bool hit = false;
                    for (int j = 0; j < nn; j++) hit |= shNew[j] == km;
                    if (hit) atomicOr(&shMask[1][p >> 5], 1u << (p & 31));
                }
                __syncthreads();
            }
            // ---- B: block walk ----
            if (tid == 0) {
                int nb = 0;
                for (int p0 = 0; p0 + k < len - k;) {
                    const int lo = p0 + 1;
                    const int w = lo >> 5, sh = lo & 31;
                    unsigned v = shMask[sl][w] >> sh;
                    if (sh) v |= shMask[sl][w + 1] << (32 - sh);
                    v &= (1u << k) - 1u;
                    if (v) {
                        p0 = lo + (__ffs(v) - 1) + 2 * k;
                    } else {
                        if (nb < OV_MAXBLK) shBlk[nb] = p0;
                        else atomicOr(err, 1u);
                        nb++;
                        p0 += 3 * k;
                    }
                }
                shNb = nb > OV_MAXBLK ? OV_MAXBLK : nb;
            }
            __syncthreads();
            const int nb = shNb;
            // ---- C: best-valued k-mer of every block (first maximum, value > 0) ----
            for (int i = (int)tid; i < nb * k; i += OV_SEL_THREADS) {
                const int b = i / k, s = i - b * k;
                shV2[i] = __ldg(values + kmer_at(shBlk[b] + 1 + s));
            }
            for (int t = (int)tid; t < P.numSeeds; t += OV_SEL_THREADS) shTop[t] = 0;
            __syncthreads();
            for (int b = (int)tid; b < nb; b += OV_SEL_THREADS) {
                double bestValue = 0.0;
                unsigned bestSeed = 0;
                for (int s = 0; s < k; s++) {
                    const double val = shV2[b * k + s];
                    if (val > bestValue) {
                        bestValue = val;
                        bestSeed = kmer_at(shBlk[b] + 1 + s);
                    }
                }
                shVal[b] = bestValue;
                shKmer[b] = bestSeed;
            }
            __syncthreads();
            // ---- D: topN ----
            for (int b = (int)tid; b < nb; b += OV_SEL_THREADS) {
                const double v = shVal[b];
                if (v > 0.0) {
                    int rank = 0;
                    for (int c = 0; c < nb; c++) {
                        const double u = shVal[c];
                        rank += (u > v) || (u == v && c < b);
                    }
                    if (rank < P.numSeeds) shTop[P.numSeeds - 1 - rank] = shKmer[b];
                }
            }
            __syncthreads();
            // ---- E: registration (warp 0) ----
            if (wib == 0) {
                int sz = size, nn = shNewN;
                for (int e0 = 0; e0 < 2 * P.numSeeds; e0 += 32) {
                    const int e = e0 + (int)lane;
                    const bool valid = e < 2 * P.numSeeds;
                    unsigned km = 0x80000000u | lane;
                    if (valid) {
                        km = shTop[e >> 1];
                        if (e & 1) km = dp_revcomp(km, k);
                    }
                    const unsigned mm = __match_any_sync(DP_FULL, km);
                    bool isNew = false;
                    if (valid && (unsigned)(__ffs(mm) - 1) == lane)
                        isNew = ((ov_ldcg(bits + (km >> 5)) >> (km & 31)) & 1u) == 0;
                    const unsigned mn = __ballot_sync(DP_FULL, isNew);
                    if (isNew) {
                        const int pre = __popc(mn & lt);
                        if (sz + pre < regCap) regKmer[sz + pre] = km;
                        else atomicOr(err, 1u);
                        atomicOr(bits + (km >> 5), 1u << (km & 31));
                        if (sl == 0 && nn + pre < 2 * OV_MAXN) shNew[nn + pre] = km;
                    }
                    sz += __popc(mn);
                    nn += __popc(mn);
                    __syncwarp();
                }
                if (lane == 0) {
                    shSize = sz;
                    if (sl == 0) shNewN = nn;
                }
            }
            __syncthreads();
            size = shSize;
        }
        nSlices += nsl;
        lastRead = id;
    }
    if (tid == 0) {
        out->size = size;
        out->nSlices = nSlices;
        out->lastRead = lastRead;
        out->sent = sent;
    }
}

// registration order <-> rank
__global__ void ov_rank_kernel(const uint2* __restrict__ table, const unsigned* __restrict__ regKmer, int size,
                               unsigned* __restrict__ kmerOfRank, unsigned* __restrict__ regOfRank,
                               unsigned* __restrict__ rankOfReg) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= size) return;
    unsigned km = regKmer[r], rank;
    dp_seed_lookup(table, km, &rank);
    kmerOfRank[rank] = km;
    regOfRank[rank] = (unsigned)r;
    rankOfReg[r] = rank;
}

// descriptors for dp_chunk_scan_kernel: a slice or a whole read, scanned as a SubSequence view (every k-mer visited)
__global__ void ov_slice_descs_kernel(const OvSlice* __restrict__ slices, int n, const long long* __restrict__ readBase,
                                      int k, DpChunkDesc* __restrict__ descs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    OvSlice s = slices[i];
    DpChunkDesc d;
    d.base = readBase[s.read] + s.start;
    d.nVisit = max(0, s.len - k + 1);
    d.pad = 0;
    descs[i] = d;
}
__global__ void ov_read_descs_kernel(const long long* __restrict__ readBase, const int* __restrict__ readLen,
                                     const unsigned char* __restrict__ ignore, int n, int k,
                                     DpChunkDesc* __restrict__ descs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DpChunkDesc d;
    d.base = readBase[i];
    d.nVisit = (ignore && ignore[i]) ? 0 : max(0, readLen[i] - k + 1);
    d.pad = 0;
    descs[i] = d;
}
__global__ void ov_add_one_kernel(unsigned* __restrict__ counts, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] += 1u;
}
__global__ void ov_sentinel_kernel(const unsigned* __restrict__ rOff, const int* __restrict__ readLen, int n,
                                   int* __restrict__ rPos, unsigned* __restrict__ rSeed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned e = rOff[i + 1] - 1;
    rPos[e] = readLen[i];
    rSeed[e] = 0xffffffffu;
}

// ---------------------------------------------------------------------------------------------------------------
// AddSequences' scan: SeedIndex.NewSeedSequence on every read that is not ignored (seeds/seeds.go:33-50 through
// AddSequenceWorker, :357-363). The per-round cost that grows with the read set: every base of every read is visited.
// One warp per read, persistent CTAs (one per SM) that keep the seed flags in shared memory — all of them for k <= 10
// (4^k bits <= 128 KiB), their 2^20-bit k-mer-prefix fold above that (positives are confirmed in the L2-resident table).
// Inside a block of 1024 positions every lane owns 32 consecutive ones (dp_map.cuh's lane-contiguous scheme: the
// lane's 32+k-1 bases sit in three registers, every k-mer is one funnel shift away):
//   pass A  hit masks of all blocks -> shared memory, count;
//   one atomicAdd allocates the read's slice of the seed arrays (count + 1: the sentinel), in whatever order the
//           warps arrive (rStart / rCount instead of a CSR);
//   pass B  the hits gather {flags, rank} and write (position, rank) in scan order.
// ---------------------------------------------------------------------------------------------------------------
#define OV_SCAN_CACHE 12  // blocks of hit masks kept per warp between the passes (longer reads are scanned twice)

template <bool EXACT>
__global__ void __launch_bounds__(1024, 1) ov_scan_kernel(const unsigned* __restrict__ words,
                                                          const long long* __restrict__ readBase,
                                                          const int* __restrict__ readLen,
                                                          const unsigned char* __restrict__ ignore, int nReads, int k,
                                                          const uint2* __restrict__ table,
                                                          const unsigned* __restrict__ filter, int fBits,
                                                          unsigned long long* cursor, unsigned long long cap,
                                                          unsigned* __restrict__ rStart, unsigned* __restrict__ rCount,
                                                          int* __restrict__ rPos, unsigned* __restrict__ rSeed) {
    extern __shared__ unsigned dp_smem[];
    const unsigned lane = dp_lane();
    const int wib = threadIdx.x >> 5;
    const unsigned fWords = ((1u << fBits) + 31u) >> 5;
    for (unsigned i = threadIdx.x; i < fWords; i += blockDim.x) dp_smem[i] = __ldg(filter + i);
    __syncthreads();
    const unsigned* filt = dp_smem;
    unsigned* hm = dp_smem + ((fWords + 3u) & ~3u) + (size_t)wib * OV_SCAN_CACHE * 32;
    const unsigned kShift = 32u - 2u * (unsigned)k;
    const unsigned fShift = 32u - (unsigned)fBits;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < nReads; r += nWarps) {
        const long long base = readBase[r];
        const int nJ = (ignore && ignore[r]) ? 0 : max(0, readLen[r] - k + 1);
        const int nBlk = (nJ + 1023) >> 10;
        auto block_mask = [&](int b) -> unsigned {
            const int pb = (b << 10) + ((int)lane << 5);
            const int nValid = min(32, max(0, nJ - pb));
            DpLaneBases B = dp_lane_bases(words, base + pb, pb < nJ + 32, k, lane);
            unsigned fm = 0;
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const unsigned x = i < 16 ? __funnelshift_l(B.a1, B.a0, 2 * i) : __funnelshift_l(B.a2, B.a1, 2 * (i - 16));
                const unsigned hx = x >> fShift;
                fm |= (__funnelshift_r(filt[hx >> 5], 0u, hx) & 1u) << i;
            }
            fm &= nValid >= 32 ? 0xffffffffu : ((1u << nValid) - 1u);
            if (!EXACT) {  // confirm the filter positives
                unsigned h = 0;
                while (fm) {
                    const unsigned i = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const unsigned kmer = dp_fwd_at(B, i) >> kShift;
                    h |= ((__ldg(&table[kmer >> 5].x) >> (kmer & 31u)) & 1u) << i;
                }
                fm = h;
            }
            return fm;
        };
        unsigned cnt = 0;
        for (int b = 0; b < nBlk; b++) {
            const unsigned m = block_mask(b);
            if (b < OV_SCAN_CACHE) hm[(b << 5) + lane] = m;
            cnt += __popc(m);
        }
        __syncwarp();
        unsigned tot = cnt;
        for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(DP_FULL, tot, d);
        unsigned long long at = 0;
        if (lane == 0) {
            at = atomicAdd(cursor, (unsigned long long)tot + 1ull);
            rStart[r] = (unsigned)at;
            rCount[r] = tot;
        }
        at = __shfl_sync(DP_FULL, at, 0);
        if (at + tot + 1ull > cap) continue;  // the host sees the cursor beyond the capacity and grows the arrays
        if (lane == 0) {
            rPos[at + tot] = readLen[r];
            rSeed[at + tot] = 0xffffffffu;
        }
        unsigned blockBase = 0;
        for (int b = 0; b < nBlk; b++) {
            unsigned m = b < OV_SCAN_CACHE ? hm[(b << 5) + lane] : block_mask(b);
            const unsigned mine = __popc(m);
            unsigned incl = mine;
            for (int d = 1; d < 32; d <<= 1) {
                unsigned y = __shfl_up_sync(DP_FULL, incl, d);
                if ((int)lane >= d) incl += y;
            }
            unsigned long long idx = at + blockBase + incl - mine;
            blockBase += __shfl_sync(DP_FULL, incl, 31);
            // (the neighbour's words come by shuffle inside dp_lane_bases: every lane takes part when any lane has a hit)
            if (__any_sync(DP_FULL, m != 0)) {
                const int pb = (b << 10) + ((int)lane << 5);
                DpLaneBases B = dp_lane_bases(words, base + pb, pb < nJ + 32, k, lane);
                while (m) {
                    const unsigned i = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned kmer = dp_fwd_at(B, i) >> kShift;
                    const uint2 e = __ldg(table + (kmer >> 5));
                    rSeed[idx] = e.y + __popc(e.x & ((1u << (kmer & 31)) - 1u));
                    rPos[idx] = pb + (int)i;
                    idx++;
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Queries (overlap/overlap.go:175-205): NewSeedSequence of every slice and its SeedSequence.ReverseComplement
// (seeds/sequence.go:134-159: gaps reversed, every seed replaced by the seed of its reverse-complement k-mer; a k-mer
// that is no seed maps to kmerMap's zero value, seed 0). One warp per slice. Also: the sorted distinct seeds of each
// query and every seed's slot among them (what the pairwise chainer compares instead of seed ids).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ov_build_queries_kernel(const OvSlice* __restrict__ slices, int nSlices, int k,
                                                               const uint2* __restrict__ table,
                                                               const unsigned* __restrict__ kmerOfRank,
                                                               const unsigned* __restrict__ rankOfReg,
                                                               const unsigned* __restrict__ fOff,
                                                               const int* __restrict__ fPos,
                                                               const unsigned* __restrict__ fSeed,
                                                               unsigned* __restrict__ qOff, int* __restrict__ qPos,
                                                               unsigned* __restrict__ qSeed,
                                                               unsigned* __restrict__ qDistinct,
                                                               unsigned short* __restrict__ qSlot,
                                                               int* __restrict__ qND, unsigned* __restrict__ err) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = dp_lane();
    if (warp >= nSlices) return;
    const unsigned fo = fOff[warp], n = fOff[warp + 1] - fo;
    const int len = slices[warp].len;
    if (lane == 0) {
        qOff[2 * warp] = 2 * fo;
        qOff[2 * warp + 1] = 2 * fo + n;
        if (warp == nSlices - 1) qOff[2 * nSlices] = 2 * (fo + n);
    }
    if (n > OV_QMAX) {
        if (lane == 0) {
            atomicOr(err, 2u);
            qND[2 * warp] = 0;
            qND[2 * warp + 1] = 0;
        }
        return;
    }
    const unsigned seed0 = rankOfReg[0];
    for (unsigned j = lane; j < n; j += 32) {
        qPos[2 * fo + j] = fPos[fo + j];
        qSeed[2 * fo + j] = fSeed[fo + j];
        const unsigned src = fo + n - 1 - j;
        unsigned rc = dp_revcomp(kmerOfRank[fSeed[src]], k), rank;
        if (!dp_seed_lookup(table, rc, &rank)) rank = seed0;
        qPos[2 * fo + n + j] = len - k - fPos[src];
        qSeed[2 * fo + n + j] = rank;
    }
    __syncwarp();
    for (int s = 0; s < 2; s++) {
        const unsigned qo = 2 * fo + s * n;
        // distinct seeds, ascending: an element is kept when no earlier element equals it (flag parked in qSlot); its
        // place = the number of kept elements below it
        for (unsigned j = lane; j < n; j += 32) {
            const unsigned v = qSeed[qo + j];
            bool first = true;
            for (unsigned b = 0; b < j; b++)
                if (qSeed[qo + b] == v) {
                    first = false;
                    break;
                }
            qSlot[qo + j] = first ? 1 : 0;
        }
        __syncwarp();
        unsigned nd = 0;
        for (unsigned j0 = 0; j0 < n; j0 += 32) {
            const unsigned j = j0 + lane;
            bool first = false;
            if (j < n && qSlot[qo + j]) {
                first = true;
                const unsigned v = qSeed[qo + j];
                unsigned below = 0;
                for (unsigned b = 0; b < n; b++) below += (qSlot[qo + b] != 0) && qSeed[qo + b] < v;
                qDistinct[qo + below] = v;
            }
            nd += __popc(__ballot_sync(DP_FULL, first));
        }
        __syncwarp();
        if (lane == 0) qND[2 * warp + s] = (int)nd;
        __syncwarp();
        for (unsigned j = lane; j < n; j += 32) {
            const unsigned v = qSeed[qo + j];
            unsigned lo = 0, hi = nd;
            while (lo < hi) {
                unsigned mid = (lo + hi) >> 1;
                if (qDistinct[qo + mid] < v) lo = mid + 1;
                else hi = mid;
            }
            qSlot[qo + j] = (unsigned short)lo;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// chunkWorker (overlap/overlap.go:253-318): a read's seed sequence cut in SEED space. One thread per read; pass 0
// counts the pieces, pass 1 writes them. With pos[] = start of every seed (and the sentinel pos[n] = length):
//   GetSeedOffset(0)         = pos[0]
//   GetNextSeedOffset(i)     = pos[i+1] - pos[i]           (i = -1: pos[0] + k)
//   GetSeedOffsetFromEnd(i)  = length - pos[i] - k
// ---------------------------------------------------------------------------------------------------------------
__global__ void ov_chunk_kernel(const unsigned* __restrict__ rStart, const unsigned* __restrict__ rCount,
                                const int* __restrict__ rPos,
                                const int* __restrict__ readLen, const unsigned char* __restrict__ ignore, int nReads,
                                OvParams P, int pass, unsigned* __restrict__ counts,
                                const unsigned* __restrict__ pieceOff, OvChunk* __restrict__ chunks,
                                unsigned* __restrict__ err) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nReads) return;
    unsigned cnt = 0;
    const unsigned ro = rStart[r];
    const int n = (int)rCount[r];
    const int length = readLen[r];
    const int k = P.k;
    OvChunk* outp = pass ? chunks + pieceOff[r] : nullptr;
    const int* pos = rPos + ro;
    auto emit = [&](int start, int end, int len, int off, int ins) {
        if (pass) {
            OvChunk c;
            c.first = ro + (unsigned)start;
            c.n = (unsigned)(end - start + 1);
            c.read = r;
            c.length = len;
            c.offset = off;
            c.inset = ins;
            outp[cnt] = c;
        }
        cnt++;
    };
    auto nextOff = [&](int i) -> int { return i < 0 ? pos[0] + k : pos[i + 1] - pos[i]; };
    if (!(ignore && ignore[r]) && n >= 0) {
        // the read's own SeedSequence: SubSequence(0, Len()) view -> offset 0, inset 1 (Q3), length = Len()
        const int numChunks = length / P.chunkSize + 1;
        if (numChunks == 1 || n < P.numSeeds * 3) {
            if (n >= P.numSeeds && n > 0) emit(0, n - 1, length, 0, 1);
        } else {
            int prevSeedIndex = 0;
            int totalOffset = pos[0];
            int lengthInBases = 0;
            for (int guard = 0;; guard++) {
                if (guard > 8 * n + 64) {  // the reference does not terminate on this read (a seedless stretch of
                    atomicOr(err, 128u);   // ~chunk_size bases followed by a few close seeds walks back and forth)
                    break;
                }
                int seedCount = 0;
                if (prevSeedIndex >= n - 150) {
                    if (prevSeedIndex == 0) {
                        emit(0, n - 1, length, 0, 1);
                    } else {
                        const int newFirstGap = nextOff(prevSeedIndex - 1) - k;
                        lengthInBases += (length - pos[prevSeedIndex] - k) + k + newFirstGap;
                        emit(prevSeedIndex, n - 1, lengthInBases, totalOffset - newFirstGap, 0);
                    }
                    break;
                }
                for (; lengthInBases < P.chunkSize && seedCount < 100 && prevSeedIndex + seedCount < n; seedCount++)
                    lengthInBases += nextOff(prevSeedIndex + seedCount);
                if (seedCount >= P.numSeeds) {
                    const int newFirstGap = nextOff(prevSeedIndex - 1) - k;
                    lengthInBases += newFirstGap;
                    emit(prevSeedIndex, prevSeedIndex + seedCount - 1, lengthInBases, totalOffset - newFirstGap,
                         length - totalOffset - lengthInBases + newFirstGap);
                    totalOffset += lengthInBases - newFirstGap;
                    lengthInBases = 0;
                    prevSeedIndex += seedCount;
                    if (prevSeedIndex >= n) break;
                    for (seedCount = 0; seedCount < 5 && lengthInBases < P.overlap / 2 && prevSeedIndex > 0; seedCount++) {
                        prevSeedIndex--;
                        const int step = nextOff(prevSeedIndex);
                        lengthInBases += step;
                        totalOffset -= step;
                    }
                    lengthInBases = 0;
                } else {
                    prevSeedIndex += seedCount;
                    for (seedCount = 0; lengthInBases < P.overlap / 2 && prevSeedIndex > 0; seedCount++) {
                        prevSeedIndex--;
                        const int step = nextOff(prevSeedIndex);
                        lengthInBases += step;
                        totalOffset -= step;
                    }
                    lengthInBases = 0;
                }
            }
        }
    }
    if (!pass) counts[r] = cnt;
}

__global__ void ov_chunk_n_kernel(const OvChunk* __restrict__ chunks, unsigned nChunks, unsigned* __restrict__ cn) {
    unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nChunks) cn[c] = chunks[c].n;
}

// (seed, chunk) keys in chunk order: one warp per chunk
__global__ void __launch_bounds__(256) ov_keys_kernel(const OvChunk* __restrict__ chunks, unsigned nChunks,
                                                      const unsigned* __restrict__ keyOff,
                                                      const unsigned* __restrict__ rSeed,
                                                      unsigned long long* __restrict__ keys) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = dp_lane();
    const unsigned nWarps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned c = warp; c < nChunks; c += nWarps) {
        const unsigned first = chunks[c].first, n = chunks[c].n, o = keyOff[c];
        for (unsigned i = lane; i < n; i += 32) keys[o + i] = ((unsigned long long)rSeed[first + i] << 32) | c;
    }
}

// sorted distinct (seed, chunk) keys -> CSR: the chunk column, and seedOff[] from the places where the seed changes (no
// atomics: a counter per seed under 60 M increments is a serial queue)
__global__ void ov_seed_bounds_kernel(const unsigned long long* __restrict__ keys, unsigned long long n, unsigned numSeeds,
                                      unsigned* __restrict__ seedOff, unsigned* __restrict__ seedChunks) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = keys[i];
    const unsigned s = (unsigned)(key >> 32);
    seedChunks[i] = (unsigned)key;
    const long long prev = i ? (long long)(keys[i - 1] >> 32) : -1ll;
    for (long long t = prev + 1; t <= (long long)s; t++) seedOff[t] = (unsigned)i;
    if (i == n - 1)
        for (unsigned t = s + 1; t <= numSeeds; t++) seedOff[t] = (unsigned)n;
}

// ---------------------------------------------------------------------------------------------------------------
// SeedIndex.Matches (seeds/seeds.go:335-353) -> util.GetSharedIDs (util/bitset.go:308-411) for the queries of a round,
// restated over posting runs (the refinement of the clamped levels and of the level-16
// under-count is dp_map.cuh's dp_refine_emit). One CTA per query; the chunk ids are walked in tiles of 65536 whose
// 16-bit counters live in shared memory (a round's index holds 10^5..10^6 chunks: counters in HBM cost one L2 atomic per
// posting, 5 ms per round; in shared memory the postings are read once and that is all the traffic):
//   inclusion filter (ordered, one thread on prefetched run lengths)
//   per tile: every run's postings inside the tile (runs ascend: one lower_bound from the previous tile's end) ->
//             shared-memory counters -> ordered threshold scan, compaction in chunk order
//   refinement + distinct counts by warp 0 into the round's candidate pool.
// minCount = int(hitFraction * n + 0.5) in fp64 (no fused multiply-add).
// ---------------------------------------------------------------------------------------------------------------
#define OV_LTHREADS 512
#define OV_TILE 65536u

__global__ void __launch_bounds__(OV_LTHREADS, 1) ov_lookup_kernel(DpIndexDev I, const unsigned* __restrict__ qOff,
                                                                   const unsigned* __restrict__ qSeed, int nQueries,
                                                                   double hitFraction,
                                                                   unsigned long long* candScratch, unsigned candCap,
                                                                   unsigned* __restrict__ qCandOff,
                                                                   int* __restrict__ qCandN,
                                                                   unsigned* __restrict__ poolChunk,
                                                                   unsigned short* __restrict__ poolDist,
                                                                   unsigned long long* cursor, unsigned long long poolCap,
                                                                   unsigned long long* __restrict__ postingCount,
                                                                   unsigned* __restrict__ err) {
    extern __shared__ unsigned cntw[];  // OV_TILE / 2 words: two 16-bit counters each
    __shared__ unsigned tSeed[OV_QMAX], tOff[OV_QMAX], tCnt[OV_QMAX];
    __shared__ unsigned eSeed[OV_QMAX], eOff[OV_QMAX], ePre[OV_QMAX + 1], eEndW[OV_QMAX];
    __shared__ unsigned cur[OV_QMAX], tPre[OV_QMAX + 1];
    __shared__ unsigned char eFirst[OV_QMAX];
    __shared__ unsigned short dupList[OV_QMAX], order[OV_QMAX];
    __shared__ int sim[2];
    __shared__ int shN[4];  // nInc, nAllDistinct, candidates out, nDup
    __shared__ unsigned shRange[2];
    __shared__ unsigned warpTot[OV_LTHREADS / 32];
    __shared__ unsigned shBase;
    const unsigned tid = threadIdx.x, lane = dp_lane(), wib = tid >> 5;
    const unsigned C = I.numChunks;
    unsigned long long* cand = candScratch + (size_t)blockIdx.x * candCap;
    unsigned long long postings = 0;
    for (int q = blockIdx.x; q < nQueries; q += gridDim.x) {
        const unsigned qb = qOff[q];
        const int n = (int)(qOff[q + 1] - qb);
        int nCandOut = 0;
        unsigned region = 0;
        __syncthreads();
        if (n >= 5 && n <= OV_QMAX) {
            for (int j = (int)tid; j < n; j += OV_LTHREADS) {
                const unsigned s = qSeed[qb + j];
                const unsigned o = __ldg(I.seedOff + s);
                tSeed[j] = s;
                tOff[j] = o;
                tCnt[j] = __ldg(I.seedOff + s + 1) - o;
            }
            if (tid == 0) {
                shRange[0] = 0xffffffffu;
                shRange[1] = 0;
            }
            __syncthreads();
            if (tid == 0) {
                // inclusion filter (seeds.go:340-346) and the distinct seeds present in every chunk
                int nInc = 0, nAllD = 0;
                unsigned prev = 0xffffffffu;
                for (int j = 0; j < n; j++) {
                    const unsigned s = tSeed[j], c = tCnt[j];
                    if (c >= C) {
                        bool first = true;
                        for (int b = 0; b < j; b++)
                            if (tSeed[b] == s) {
                                first = false;
                                break;
                            }
                        nAllD += first;
                    } else if (s != prev) {
                        eSeed[nInc] = s;
                        eOff[nInc] = tOff[j];
                        ePre[nInc] = c;
                        nInc++;
                        prev = s;
                    }
                }
                shN[0] = nInc;
                shN[1] = nAllD;
            }
            __syncthreads();
            const int nInc = shN[0];
            if (nInc >= 5) {
                const int minCount = (int)__dadd_rn(__dmul_rn(hitFraction, (double)nInc), 0.5);
                int T;
                bool clamped = false;
                if (minCount >= 9 && minCount <= 12) {
                    T = 8;
                    clamped = true;
                } else if (minCount >= 17 && minCount <= 24) {
                    T = 16;
                    clamped = true;
                } else {
                    T = minCount > 1 ? minCount : 1;
                }
                const bool q6 = minCount >= 13 && minCount <= 24;
                // first occurrences among the included runs, first / last chunk of each run
                unsigned cmin = 0xffffffffu, cmax = 0;
                for (int j = (int)tid; j < nInc; j += OV_LTHREADS) {
                    const unsigned s = eSeed[j];
                    bool dup = false;
                    for (int b = 0; b < j; b++)
                        if (eSeed[b] == s) {
                            dup = true;
                            break;
                        }
                    eFirst[j] = dup ? 0 : 1;
                    const unsigned c = ePre[j];
                    unsigned last = 0;
                    if (c) {
                        last = __ldg(I.seedChunks + eOff[j] + c - 1);
                        cmin = min(cmin, __ldg(I.seedChunks + eOff[j]));
                        cmax = max(cmax, last);
                    }
                    eEndW[j] = c ? (last >> 6) : 0u;
                    cur[j] = 0;
                }
                if (cmin != 0xffffffffu) {
                    atomicMin(&shRange[0], cmin);
                    atomicMax(&shRange[1], cmax);
                }
                __syncthreads();
                if (tid == 0) {
                    unsigned tot = 0;
                    int nDup = 0;
                    for (int j = 0; j < nInc; j++) {
                        const unsigned c = ePre[j];
                        ePre[j] = tot;
                        tot += c;
                        if (!eFirst[j]) dupList[nDup++] = (unsigned short)j;
                    }
                    ePre[nInc] = tot;
                    shN[3] = nDup;
                }
                __syncthreads();
                postings += tid == 0 ? ePre[nInc] : 0u;
                unsigned nCand = 0;
                if (shRange[0] != 0xffffffffu) {
                    const unsigned tile0 = shRange[0] / OV_TILE, tile1 = shRange[1] / OV_TILE;
                    for (unsigned tile = tile0; tile <= tile1; tile++) {
                        const unsigned tStart = tile * OV_TILE;
                        const unsigned tEnd = min(C, tStart + OV_TILE);
                        const unsigned tWords = (tEnd - tStart + 1) >> 1;
                        // the part of every run that lies inside the tile
                        for (int j = (int)tid; j < nInc; j += OV_LTHREADS) {
                            const unsigned len = ePre[j + 1] - ePre[j];
                            unsigned lo = cur[j], hi = len;
                            const unsigned* run = I.seedChunks + eOff[j];
                            while (lo < hi) {
                                const unsigned mid = (lo + hi) >> 1;
                                if (__ldg(run + mid) < tEnd) lo = mid + 1;
                                else hi = mid;
                            }
                            tPre[j] = lo - cur[j];
                        }
                        for (unsigned i = tid; i < tWords; i += OV_LTHREADS) cntw[i] = 0;
                        __syncthreads();
                        if (wib == 0) {  // exclusive prefix of the in-tile lengths
                            unsigned carry = 0;
                            for (int j0 = 0; j0 < nInc; j0 += 32) {
                                const int j = j0 + (int)lane;
                                const unsigned c = j < nInc ? tPre[j] : 0u;
                                unsigned x = c;
                                for (int d = 1; d < 32; d <<= 1) {
                                    unsigned y = __shfl_up_sync(DP_FULL, x, d);
                                    if ((int)lane >= d) x += y;
                                }
                                if (j < nInc) tPre[j] = carry + x - c;
                                carry += __shfl_sync(DP_FULL, x, 31);
                            }
                            if (lane == 0) tPre[nInc] = carry;
                        }
                        __syncthreads();
                        const unsigned tTot = tPre[nInc];
                        if (tTot == 0) continue;  // (uniform)
                        for (unsigned p = tid; p < tTot; p += OV_LTHREADS) {
                            int lo = 0, hi = nInc;
                            while (hi - lo > 1) {
                                const int mid = (lo + hi) >> 1;
                                if (tPre[mid] <= p) lo = mid;
                                else hi = mid;
                            }
                            const unsigned chunk = __ldg(I.seedChunks + eOff[lo] + cur[lo] + (p - tPre[lo]));
                            const unsigned loc = chunk - tStart;
                            atomicAdd(&cntw[loc >> 1], 1u << ((loc & 1u) * 16u));
                        }
                        __syncthreads();
                        for (int j = (int)tid; j < nInc; j += OV_LTHREADS) cur[j] += tPre[j + 1] - tPre[j];
                        // ordered threshold scan of the tile
                        for (unsigned wb = 0; wb < tWords; wb += OV_LTHREADS * 2) {
                            const unsigned wi = wb + tid * 2;
                            unsigned v[4] = {0, 0, 0, 0};
                            if (wi < tWords) {
                                const unsigned x = cntw[wi];
                                v[0] = x & 0xffffu;
                                v[1] = x >> 16;
                                if (wi + 1 < tWords) {
                                    const unsigned y = cntw[wi + 1];
                                    v[2] = y & 0xffffu;
                                    v[3] = y >> 16;
                                }
                            }
                            unsigned mine = 0;
                            for (int i = 0; i < 4; i++) mine += (int)v[i] >= T;
                            unsigned x = mine;
                            for (int d = 1; d < 32; d <<= 1) {
                                unsigned y = __shfl_up_sync(DP_FULL, x, d);
                                if ((int)lane >= d) x += y;
                            }
                            if (lane == 31) warpTot[wib] = x;
                            __syncthreads();
                            unsigned before = 0, all = 0;
                            for (unsigned w = 0; w < OV_LTHREADS / 32; w++) {
                                if (w < wib) before += warpTot[w];
                                all += warpTot[w];
                            }
                            if (all) {
                                unsigned idx = nCand + before + x - mine;
                                for (int i = 0; i < 4; i++)
                                    if ((int)v[i] >= T) {
                                        if (idx < candCap) cand[idx] = ((unsigned long long)(tStart + wi * 2 + i) << 32) | v[i];
                                        idx++;
                                    }
                                nCand += all;
                            }
                            __syncthreads();
                        }
                    }
                }
                if (nCand > candCap) {
                    if (tid == 0) atomicOr(err, 4u);
                    nCand = 0;
                }
                // ---- refinement, distinct counts, pool ----
                if (tid == 0) shBase = 0xffffffffu;
                __syncthreads();
                if (wib == 0 && nCand > 0) {
                    unsigned long long at = 0;
                    if (lane == 0) at = atomicAdd(cursor, (unsigned long long)nCand);
                    at = __shfl_sync(DP_FULL, at, 0);
                    if (at + nCand <= poolCap) {
                        DpRefineCtx X;
                        X.eOff = eOff;
                        X.ePre = ePre;
                        X.eEndW = eEndW;
                        X.eFirst = eFirst;
                        X.dup = dupList;
                        X.nDup = shN[3];
                        X.order = order;
                        X.sim = sim;
                        X.nInc = nInc;
                        X.minCount = minCount;
                        X.T = T;
                        X.clamped = clamped;
                        X.q6 = q6;
                        X.nAllDistinct = shN[1];
                        __threadfence_block();
                        __syncwarp();
                        int no = dp_refine_emit<false>(I, X, cand, (int)nCand, poolChunk + at, poolDist + at, (int)nCand);
                        if (lane == 0) {
                            shBase = (unsigned)at;
                            shN[2] = no;
                        }
                    } else if (lane == 0) {
                        atomicOr(err, 8u);
                    }
                }
                __syncthreads();
                if (shBase != 0xffffffffu) {
                    region = shBase;
                    nCandOut = shN[2];
                }
            }
        } else if (n > OV_QMAX) {
            if (tid == 0) atomicOr(err, 2u);
        }
        if (tid == 0) {
            qCandOff[q] = region;
            qCandN[q] = nCandOut;
        }
    }
    if (postings) atomicAdd(postingCount, postings);
}

// ---------------------------------------------------------------------------------------------------------------
// matchWorker (overlap/overlap.go:346-387) with seedAligner.PairwiseAlignments (seeds/alignment.go:426-616).
//
// One CTA per query, one lane per candidate chunk of the current wave. The reference walks a query's candidates in
// ascending chunk order and raises minMatches after a long alignment (overlap.go:382-384), so a candidate's result
// depends on the ones before it. The wave speculates with the current minMatches; the first lane whose alignment raises
// it ends the wave: lanes up to it commit, the ones behind it are recomputed with the new value (a query raises its
// threshold a handful of times at most).
//
// What a lane computes is the reference's state machine statement by statement, with three observations:
//   * the match matchWorker keeps is sMatches' LAST element with a non-zero cover (bestCount is never updated,
//     overlap.go:369-372); covers are positive (seed starts strictly increase), and sMatches lists `results` backwards
//     (alignment.go:609), so the match is results[0]: the FIRST chain that reaches the results — the lane stops there;
//   * the two removeOpenState calls with their arguments in the wrong slots (alignment.go:490,497) sit behind
//     `found != -1` inside the open-list loop, but every assignment to `found` in that loop is followed by
//     `break searchMatch`: dead code;
//   * states are identified by nothing but their links, so the 10 000-entry pool with its recycling stack is replaced
//     by a bump-allocated node list per lane (a chain that is dropped just stays behind).
// Capacities of the reference that panic there (reduced / open / results overflow) are reported through `err`.
// ---------------------------------------------------------------------------------------------------------------
struct OvAlignScratch {
    unsigned short* oAPos;   // [threads][OV_OPEN] open list, structure of arrays
    unsigned short* oBPos;
    unsigned short* oAGapIndex;
    unsigned short* oLength;
    int* oAGap;
    int* oBGap;
    int* oNode;
    unsigned* nodes;   // [threads][nodeCap] (aPos << 16 | bPos)
    int* nodePrev;     // [threads][nodeCap]
    int nodeCap;
};

__device__ __forceinline__ void ov_gap_range(int gap, int k, int* minGap, int* maxGap) {  // alignment.go:411-424
    int mn = (gap * 2) / 3 - k;
    int mx = (gap * 3) / 2 + k + 1;
    if (mn < 0) {
        mn = -k;
        if (mx < 0) mx = 0;
    } else if (mx < 20) {
        mx = 20;
        mn = 0;
    }
    *minGap = mn;
    *maxGap = mx;
}

// returns the chain length (0: nil) and its head node
template <int W>
__device__ int ov_pairwise(const OvParams& P, const int* __restrict__ aPosArr, const unsigned short* __restrict__ aSlotArr,
                           int nA, const unsigned* __restrict__ qd, int nd, const int* __restrict__ bpos,
                           const unsigned* __restrict__ bseed, int nB, int minMatches, const OvAlignScratch& S,
                           size_t th, unsigned short* aMapOut, int* headOut, unsigned* err) {
    const int k = P.k;
    if (minMatches == 0) minMatches = 1;
    unsigned short* oAPos = S.oAPos + th * OV_OPEN;
    unsigned short* oBPos = S.oBPos + th * OV_OPEN;
    unsigned short* oAGI = S.oAGapIndex + th * OV_OPEN;
    unsigned short* oLen = S.oLength + th * OV_OPEN;
    int* oAGap = S.oAGap + th * OV_OPEN;
    int* oBGap = S.oBGap + th * OV_OPEN;
    int* oNode = S.oNode + th * OV_OPEN;
    unsigned* nodes = S.nodes + th * (size_t)S.nodeCap;
    int* nodePrev = S.nodePrev + th * (size_t)S.nodeCap;
    int nNodes = 0;
    // ---- bSet restricted to the query's seeds ----
    unsigned mask[OV_QMAX / 32];
#pragma unroll
    for (int i = 0; i < OV_QMAX / 32; i++) mask[i] = 0;
    auto slotOf = [&](unsigned s) -> int {
        int lo = 0, hi = nd;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (qd[mid] < s) lo = mid + 1;
            else hi = mid;
        }
        return (lo < nd && qd[lo] == s) ? lo : -1;
    };
    for (int i = 0; i < nB; i++) {
        const int sl = slotOf(__ldg(bseed + i));
        if (sl >= 0) mask[sl >> 5] |= 1u << (sl & 31);
    }
    // ---- prepareInitial (alignment.go:341-388) ----
    int aGapArr[OV_AMAX + 1];
    unsigned short aSl[OV_AMAX];
    int maxAIndex = (2 * nA + 1) - minMatches * 2 + 1;
    int aLen = 0, offset = -k, startSize = 0, prevSlot = -1;
    for (int ai = 0; ai < nA; ai++) {
        const int gapBefore = ai == 0 ? aPosArr[0] : aPosArr[ai] - aPosArr[ai - 1] - k;
        const int sl = aSlotArr[ai];
        if (!((mask[sl >> 5] >> (sl & 31)) & 1u)) {
            offset += gapBefore + k;
            maxAIndex--;
            continue;
        }
        if (sl == prevSlot && (ai >= nA - 1 || aSlotArr[ai + 1] == prevSlot)) {
            offset += gapBefore + k;
            maxAIndex--;
            continue;
        }
        prevSlot = sl;
        offset += gapBefore + k;
        if (aLen * 2 + 1 >= P.redCap || aLen >= OV_AMAX) {  // Go: index out of range on align.reduced
            atomicOr(err, 16u);
            return 0;
        }
        aGapArr[aLen] = offset;
        aSl[aLen] = (unsigned short)sl;
        aMapOut[aLen] = (unsigned short)ai;
        offset = -k;
        if (aLen <= maxAIndex) startSize++;
        aLen++;
    }
    if (aLen * 2 >= P.redCap) {
        atomicOr(err, 16u);
        return 0;
    }
    aGapArr[aLen] = 0;
    while (startSize > 0 && 2 * (startSize - 1) + 1 > maxAIndex) startSize--;
    const int initialSize = startSize;
    const int aRedLen = aLen * 2 + 1;
#define OV_AR_SEED(j) ((int)aSl[(j) >> 1])   /* aRed[j], j odd */
#define OV_AR_GAP(j) (aGapArr[(j) >> 1])     /* aRed[j], j even */
    // ---- main loop over b ----
    int openSize = 0;
    const int bLen = 2 * nB + 1;
    int maxBIndex = bLen - minMatches * 2 + 1;
    int bOffset = 0;
    unsigned prevSeed = 0xfffffffeu;
    for (int bi = 0; bi < nB; bi++) {
        const int bIndex = 2 * bi + 1;
        const unsigned bs = __ldg(bseed + bi);
        const int gapAfter = __ldg(bpos + bi + 1) - __ldg(bpos + bi) - k;
        const int slot = slotOf(bs);
        if (slot < 0) {
            bOffset += gapAfter + k;
            continue;
        }
        if (bs == prevSeed && (bi >= nB - 1 || __ldg(bseed + bi + 1) == prevSeed)) {
            bOffset += gapAfter + k;
            continue;
        }
        prevSeed = bs;
        int found = -1;
        for (int i = openSize - 1; i >= 0; i--) {
            int sAGap = oAGap[i], sAGI = oAGI[i];
            const int sLen = oLen[i];
            int sBGap = oBGap[i] + bOffset;
            oBGap[i] = sBGap;
            int minGap, maxGap;
            ov_gap_range(sBGap, k, &minGap, &maxGap);
            bool ended = false;
            while (sAGap < minGap) {
                if (sAGI >= aRedLen) {
                    ended = true;
                    break;
                }
                sAGap += OV_AR_GAP(sAGI + 1) + k;
                sAGI += 2;
            }
            oAGap[i] = sAGap;
            oAGI[i] = (unsigned short)sAGI;
            if (ended) {
                // removeOpenState(i, minMatches, ...): a chain that is long enough is results[0]
                if (sLen >= minMatches) {
                    *headOut = oNode[i];
                    return sLen;
                }
                openSize--;
                oAPos[i] = oAPos[openSize];
                oBPos[i] = oBPos[openSize];
                oAGI[i] = oAGI[openSize];
                oLen[i] = oLen[openSize];
                oAGap[i] = oAGap[openSize];
                oBGap[i] = oBGap[openSize];
                oNode[i] = oNode[openSize];
                break;  // break searchMatch
            }
            bool extended = false;
            if (sAGap <= maxGap) {
                int g = sAGap;
                for (int j = sAGI; j < aRedLen && g <= maxGap; j += 2) {
                    if (OV_AR_SEED(j) == slot) {
                        found = j;
                        if (nNodes >= S.nodeCap) {
                            atomicOr(err, 32u);
                            return 0;
                        }
                        nodes[nNodes] = ((unsigned)j << 16) | (unsigned)bIndex;
                        nodePrev[nNodes] = oNode[i];
                        oNode[i] = nNodes++;
                        oAPos[i] = (unsigned short)j;
                        oBPos[i] = (unsigned short)bIndex;
                        oAGI[i] = (unsigned short)(j + 2);
                        oAGap[i] = OV_AR_GAP(j + 1);
                        oBGap[i] = gapAfter;
                        oLen[i] = (unsigned short)(sLen + 1);
                        if (((sLen + 1) * 2) / 3 > minMatches) {
                            minMatches = ((sLen + 1) * 2) / 3;
                            maxBIndex = bLen - minMatches * 2 + 1;
                        }
                        extended = true;
                        break;
                    }
                    g += OV_AR_GAP(j + 1) + k;
                }
            }
            if (extended) break;  // break searchMatch
            if (sLen + (bLen - bIndex) < minMatches) {
                // removeOpenState: too short to ever reach minMatches (sLen < minMatches here), the chain is dropped
                openSize--;
                oAPos[i] = oAPos[openSize];
                oBPos[i] = oBPos[openSize];
                oAGI[i] = oAGI[openSize];
                oLen[i] = oLen[openSize];
                oAGap[i] = oAGap[openSize];
                oBGap[i] = oBGap[openSize];
                oNode[i] = oNode[openSize];
            } else {
                oBGap[i] = sBGap + gapAfter + k;
            }
        }
        bOffset = 0;
        if (bIndex <= maxBIndex) {
            for (int i = 0; i < initialSize; i++) {
                const int aPos = 2 * i + 1;
                if (aPos != found && (int)aSl[i] == slot) {
                    if (found != -1) {
                        for (int j = 0; j < openSize; j++)
                            if (oBPos[j] == bIndex && oAPos[j] == aPos) {
                                found = aPos;
                                break;
                            }
                    }
                    if (found == aPos || openSize >= OV_OPEN) continue;
                    if (nNodes >= S.nodeCap) {
                        atomicOr(err, 32u);
                        return 0;
                    }
                    nodes[nNodes] = ((unsigned)aPos << 16) | (unsigned)bIndex;
                    nodePrev[nNodes] = -1;
                    oNode[openSize] = nNodes++;
                    oAPos[openSize] = (unsigned short)aPos;
                    oBPos[openSize] = (unsigned short)bIndex;
                    oAGI[openSize] = (unsigned short)(aPos + 2);
                    oAGap[openSize] = OV_AR_GAP(aPos + 1);
                    oBGap[openSize] = gapAfter;
                    oLen[openSize] = 1;
                    openSize++;
                }
            }
        }
    }
    for (int i = 0; i < openSize; i++)
        if ((int)oLen[i] >= minMatches) {
            *headOut = oNode[i];
            return oLen[i];
        }
    return 0;
#undef OV_AR_SEED
#undef OV_AR_GAP
}

template <int W>
__global__ void __launch_bounds__(W) ov_align_kernel(OvParams P, int nQueries, const unsigned* __restrict__ qOff,
                                                     const int* __restrict__ qPos,
                                                     const unsigned short* __restrict__ qSlot,
                                                     const unsigned* __restrict__ qDistinct,
                                                     const int* __restrict__ qND,
                                                     const unsigned* __restrict__ qCandOff,
                                                     const int* __restrict__ qCandN,
                                                     const unsigned* __restrict__ poolChunk,
                                                     const unsigned short* __restrict__ poolDist,
                                                     const OvChunk* __restrict__ chunks,
                                                     const int* __restrict__ rPos, const unsigned* __restrict__ rSeed,
                                                     OvAlignScratch S, int* __restrict__ hitLen,
                                                     unsigned long long* __restrict__ hitOff,
                                                     unsigned short* __restrict__ matchPool,
                                                     unsigned long long* matchCursor, unsigned long long matchCap,
                                                     unsigned long long* __restrict__ pairCount,
                                                     unsigned* __restrict__ err) {
    __shared__ int aPosArr[OV_QMAX];
    __shared__ unsigned short aSlotArr[OV_QMAX];
    __shared__ unsigned qd[OV_QMAX];
    __shared__ int shFirstEsc[W / 32];
    __shared__ int shEscLen;
    const int tid = threadIdx.x;
    const unsigned lane = dp_lane();
    const size_t th = (size_t)blockIdx.x * W + tid;
    unsigned long long pairs = 0;
    for (int q = blockIdx.x; q < nQueries; q += gridDim.x) {
        const unsigned qb = qOff[q];
        const int nA = (int)(qOff[q + 1] - qb);
        const int nCand = qCandN[q];
        const unsigned region = qCandOff[q];
        __syncthreads();
        if (nCand <= 0 || nA > OV_QMAX) continue;
        const int nd = qND[q];
        for (int j = tid; j < nA; j += W) {
            aPosArr[j] = qPos[qb + j];
            aSlotArr[j] = qSlot[qb + j];
        }
        for (int j = tid; j < nd; j += W) qd[j] = qDistinct[qb + j];
        __syncthreads();
        int minMatches = (int)__dadd_rn(__dmul_rn(P.hitFraction, (double)nA), 0.5);
        for (int base = 0; base < nCand;) {
            const int ci = base + tid;
            int len = 0, head = -1;
            unsigned short aMap[OV_AMAX];
            unsigned chunk = 0;
            if (ci < nCand && (int)poolDist[region + ci] >= minMatches) {
                chunk = poolChunk[region + ci];
                const OvChunk ck = chunks[chunk];
                len = ov_pairwise<W>(P, aPosArr, aSlotArr, nA, qd, nd, rPos + ck.first, rSeed + ck.first, (int)ck.n,
                                     minMatches, S, th, aMap, &head, err);
                pairs++;
            }
            const bool esc = len * 2 > minMatches * 3;
            const unsigned me = __ballot_sync(DP_FULL, esc);
            if (lane == 0) shFirstEsc[tid >> 5] = me ? (tid + __ffs(me) - 1) : W;
            __syncthreads();
            int firstEsc = W;
            for (int w = 0; w < W / 32; w++) firstEsc = min(firstEsc, shFirstEsc[w]);
            if (tid == firstEsc) shEscLen = len;
            if (ci < nCand && tid <= firstEsc) {
                hitLen[region + ci] = len;
                if (len > 0) {
                    const unsigned long long at = atomicAdd(matchCursor, (unsigned long long)(2 * len));
                    hitOff[region + ci] = at;
                    if (at + 2 * len <= matchCap) {
                        const unsigned* nodes = S.nodes + th * (size_t)S.nodeCap;
                        const int* nodePrev = S.nodePrev + th * (size_t)S.nodeCap;
                        int s = head;
                        for (int i = len - 1; i >= 0 && s >= 0; i--) {
                            const unsigned nd2 = nodes[s];
                            matchPool[at + i] = aMap[(nd2 >> 16) >> 1];
                            matchPool[at + len + i] = (unsigned short)((nd2 & 0xffffu) >> 1);
                            s = nodePrev[s];
                        }
                    } else {
                        atomicOr(err, 64u);
                    }
                }
            }
            __syncthreads();
            if (firstEsc < W) {
                minMatches = (shEscLen * 2) / 3;
                base += firstEsc + 1;
            } else {
                base += W;
            }
            __syncthreads();
        }
    }
    if (pairs) atomicAdd(pairCount, pairs);
}

// ---------------------------------------------------------------------------------------------------------------
// Hits in delivery order: queries ascending, candidates ascending. Pass 0 counts per query, pass 1 (after the scans)
// writes the records and moves the match lists next to each other.
// ---------------------------------------------------------------------------------------------------------------
struct OvHit {      // one *seeds.SeedMatch of FindOverlaps' channel
    int queryId;    // SeedMatch.QueryID (= slice index: shared by the forward and reverse-complement query)
    int rc;         // SeedMatch.ReverseComplementQuery
    int target;     // chunk id: index.GetSeedSequence(target) is SeqB
    int n;          // len(MatchA) = len(MatchB)
    long long at;   // first element of MatchA in the match array; MatchB follows at at + n
};

__global__ void __launch_bounds__(128) ov_hit_count_kernel(int nQueries, const unsigned* __restrict__ qCandOff,
                                                           const int* __restrict__ qCandN,
                                                           const int* __restrict__ hitLen,
                                                           unsigned* __restrict__ qHits,
                                                           unsigned long long* __restrict__ qMatch) {
    __shared__ unsigned shH[4];
    __shared__ unsigned long long shM[4];
    const int q = blockIdx.x;
    if (q >= nQueries) return;
    const unsigned region = qCandOff[q];
    const int nCand = qCandN[q];
    unsigned h = 0;
    unsigned long long m = 0;
    for (int i = threadIdx.x; i < nCand; i += 128) {
        const int l = hitLen[region + i];
        if (l > 0) {
            h++;
            m += 2ull * (unsigned)l;
        }
    }
    for (int d = 16; d; d >>= 1) {
        h += __shfl_xor_sync(DP_FULL, h, d);
        m += __shfl_xor_sync(DP_FULL, m, d);
    }
    if (dp_lane() == 0) {
        shH[threadIdx.x >> 5] = h;
        shM[threadIdx.x >> 5] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        qHits[q] = shH[0] + shH[1] + shH[2] + shH[3];
        qMatch[q] = shM[0] + shM[1] + shM[2] + shM[3];
    }
}

__global__ void __launch_bounds__(32) ov_hit_write_kernel(int nQueries, const unsigned* __restrict__ qCandOff,
                                                          const int* __restrict__ qCandN,
                                                          const unsigned* __restrict__ poolChunk,
                                                          const int* __restrict__ hitLen,
                                                          const unsigned long long* __restrict__ hitOff,
                                                          const unsigned short* __restrict__ matchPool,
                                                          const unsigned* __restrict__ qHitOff,
                                                          const unsigned long long* __restrict__ qMatchOff,
                                                          OvHit* __restrict__ hits,
                                                          unsigned short* __restrict__ matches) {
    const int q = blockIdx.x;
    if (q >= nQueries) return;
    const unsigned lane = dp_lane(), lt = dp_lanemask_lt();
    const unsigned region = qCandOff[q];
    const int nCand = qCandN[q];
    unsigned hAt = qHitOff[q];
    unsigned long long mAt = qMatchOff[q];
    for (int i0 = 0; i0 < nCand; i0 += 32) {
        const int i = i0 + (int)lane;
        const int l = i < nCand ? hitLen[region + i] : 0;
        const unsigned mh = __ballot_sync(DP_FULL, l > 0);
        // exclusive scan of 2*l over the lanes
        unsigned long long x = l > 0 ? 2ull * (unsigned)l : 0ull, mine = x;
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long y = __shfl_up_sync(DP_FULL, x, d);
            if ((int)lane >= d) x += y;
        }
        const unsigned long long tot = __shfl_sync(DP_FULL, x, 31);
        if (l > 0) {
            OvHit h;
            h.queryId = q >> 1;
            h.rc = q & 1;
            h.target = (int)poolChunk[region + i];
            h.n = l;
            h.at = (long long)(mAt + x - mine);
            hits[hAt + __popc(mh & lt)] = h;
            const unsigned long long src = hitOff[region + i];
            for (int t = 0; t < 2 * l; t++) matches[h.at + t] = matchPool[src + t];
        }
        hAt += __popc(mh);
        mAt += tot;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Export of seed sequences as the reference's segments (gap, seed, gap, ..., gap) with its seed ids
// ---------------------------------------------------------------------------------------------------------------
// chunks: ids[] -> per requested chunk {read, length, offset, inset, nSeeds} and its 2n+1 segments at segOff[i]
__global__ void ov_export_chunk_meta_kernel(const OvChunk* __restrict__ chunks, const unsigned* __restrict__ ids, int n,
                                            long long* __restrict__ meta) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    OvChunk c = chunks[ids[i]];
    meta[i * 5 + 0] = c.read;
    meta[i * 5 + 1] = c.length;
    meta[i * 5 + 2] = c.offset;
    meta[i * 5 + 3] = c.inset;
    meta[i * 5 + 4] = c.n;
}
__global__ void ov_export_chunk_segs_kernel(const OvChunk* __restrict__ chunks, const unsigned* __restrict__ ids, int n,
                                            const unsigned* __restrict__ rStart, const int* __restrict__ rPos,
                                            const unsigned* __restrict__ rSeed, const unsigned* __restrict__ regOfRank,
                                            int k, const long long* __restrict__ segOff, long long* __restrict__ segs) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = dp_lane();
    if (w >= n) return;
    const OvChunk c = chunks[ids[w]];
    long long* out = segs + segOff[w];
    const unsigned r0 = rStart[c.read];
    for (unsigned i = lane; i <= c.n; i += 32) {
        const unsigned g = c.first + i;
        // gap before seed i of the piece (i == n: the gap behind its last seed)
        long long gap;
        if (g == r0) gap = rPos[g];
        else gap = (long long)rPos[g] - rPos[g - 1] - k;
        if (c.n == 0) gap = rPos[r0];  // a seedless read: one segment = the read length
        out[2 * i] = gap;
        if (i < c.n) out[2 * i + 1] = regOfRank[rSeed[g]];
    }
}
__global__ void ov_export_query_segs_kernel(int nQueries, const unsigned* __restrict__ qOff, const int* __restrict__ qPos,
                                            const unsigned* __restrict__ qSeed, const OvSlice* __restrict__ slices,
                                            const unsigned* __restrict__ regOfRank, int k, long long* __restrict__ segs) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = dp_lane();
    if (q >= nQueries) return;
    const unsigned qb = qOff[q], n = qOff[q + 1] - qb;
    const int len = slices[q >> 1].len;
    long long* out = segs + 2ll * qb + q;  // 2n+1 segments per query
    for (unsigned i = lane; i <= n; i += 32) {
        long long gap;
        if (n == 0) gap = len;
        else if (i == 0) gap = qPos[qb];
        else if (i == n) gap = (long long)len - qPos[qb + n - 1] - k;
        else gap = (long long)qPos[qb + i] - qPos[qb + i - 1] - k;
        out[2 * i] = gap;
        if (i < n) out[2 * i + 1] = regOfRank[qSeed[qb + i]];
    }
}
