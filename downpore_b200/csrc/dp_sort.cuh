// downpore_b200 — the sort of index construction on the device: a stable LSD radix sort of 64-bit (seed, chunk) keys
// with optional 32-bit values, the exclusive scan it needs, and unique-of-sorted. They replace util/sort.go's sort
// (the reference sorts seed ids on the CPU while it builds SeedIndex.sequenceSets, seeds/seeds.go:372-384) and are the
// only sort / scan / select primitives of the library.
//
// Scan: two kernels for any n. Tiles of 2048 items; the first kernel sums every tile and the LAST tile to finish (a
// ticket counter) scans the tile sums; the second kernel scans inside the tiles and adds the tile offsets. The input is
// read through a functor, so "flags of a sorted array" (unique) or "counts as another type" need no array of their own.
//
// Radix sort: 8 bits per pass, two kernels + one scan per pass. A tile is 4096 keys = 8 warps x 16 rounds of 32
// CONSECUTIVE keys, which is what makes it stable with warp primitives only: in a round the lanes with equal digits
// find each other with __match_any_sync, their order is the lane order, and a per-warp counter of the digit carries the
// order from round to round; the warps' counters are then prefixed per digit (warp order = key order). The tile is
// staged in shared memory in sorted order, so that every run of equal digits leaves as one contiguous, coalesced piece
// at the tile's base for the digit, which comes from the scan of the digit-major (digit, tile) histogram.
#pragma once
#include "dp_common.cuh"
#include "dp_host.hpp"

#define DP_SCAN_TILE 2048
#define DP_SORT_TILE 4096

template <class T>
struct DpLoadPtr {
    const T* p;
    template <class TO>
    __device__ __forceinline__ TO get(long long i) const {
        return (TO)p[i];
    }
};
// 1 where a sorted array starts a new value
struct DpLoadFirstFlags {
    const unsigned long long* k;
    template <class TO>
    __device__ __forceinline__ TO get(long long i) const {
        return (TO)((i == 0 || k[i] != k[i - 1]) ? 1 : 0);
    }
};

template <class TO>
__device__ __forceinline__ TO dp_block_exclusive(TO v, TO* warpSums, TO* total) {  // 256 threads; returns the exclusive prefix
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TO inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        TO o = __shfl_up_sync(DP_FULL, inc, d);
        if (lane >= (unsigned)d) inc += o;
    }
    if (lane == 31) warpSums[warp] = inc;
    __syncthreads();
    TO base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const TO s = warpSums[w];
        if ((unsigned)w < warp) base += s;
        all += s;
    }
    __syncthreads();
    *total = all;
    return base + inc - v;
}

template <class TO, class F>
__global__ void __launch_bounds__(256) dp_scan_sums_kernel(F load, long long n, TO* __restrict__ tileSums,
                                                           TO* __restrict__ tileOffs, unsigned* __restrict__ ticket) {
    __shared__ TO warpSums[8];
    __shared__ bool last;
    const long long base = (long long)blockIdx.x * DP_SCAN_TILE;
    TO s = 0;
#pragma unroll
    for (int j = 0; j < DP_SCAN_TILE / 256; j++) {
        const long long i = base + j * 256 + threadIdx.x;
        if (i < n) s += load.template get<TO>(i);
    }
    TO total;
    dp_block_exclusive<TO>(s, warpSums, &total);
    if (threadIdx.x == 0) {
        tileSums[blockIdx.x] = total;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last tile to arrive scans the tile sums (a few hundred trips at most: 2^31 items are 2^20 tiles)
    TO carry = 0;
    const long long nTiles = gridDim.x;
    for (long long t0 = 0; t0 < nTiles; t0 += 256) {
        const long long t = t0 + threadIdx.x;
        const TO v = t < nTiles ? __ldcg(tileSums + t) : (TO)0;
        TO all;
        const TO ex = dp_block_exclusive<TO>(v, warpSums, &all);
        if (t < nTiles) tileOffs[t] = carry + ex;
        carry += all;
    }
    if (threadIdx.x == 0) *ticket = 0;  // ready for the next scan on this stream
}

template <class TO, class F>
__global__ void __launch_bounds__(256) dp_scan_write_kernel(F load, long long n, const TO* __restrict__ tileOffs,
                                                            TO* __restrict__ out) {
    __shared__ TO warpSums[8];
    const long long first = (long long)blockIdx.x * DP_SCAN_TILE + (long long)threadIdx.x * (DP_SCAN_TILE / 256);
    TO v[DP_SCAN_TILE / 256];
    TO s = 0;
#pragma unroll
    for (int j = 0; j < DP_SCAN_TILE / 256; j++) {
        v[j] = first + j < n ? load.template get<TO>(first + j) : (TO)0;
        s += v[j];
    }
    TO total;
    TO run = dp_block_exclusive<TO>(s, warpSums, &total) + tileOffs[blockIdx.x];
#pragma unroll
    for (int j = 0; j < DP_SCAN_TILE / 256; j++) {
        if (first + j < n) out[first + j] = run;
        run += v[j];
    }
}

// out[i] = load(0) + .. + load(i-1), i in [0, n)
template <class TO, class F>
void dp_exclusive_scan(F load, TO* out, long long n, DBuf<unsigned char>& tmp, cudaStream_t st) {
    if (n <= 0) return;
    const long long nTiles = (n + DP_SCAN_TILE - 1) / DP_SCAN_TILE;
    const size_t need = 256 + (size_t)(2 * nTiles + 2) * sizeof(TO);
    if (tmp.cap < need) {
        tmp.reserve(need);
        CK(cudaMemsetAsync(tmp.p, 0, 256, st));  // the ticket (every scan leaves it at zero)
    }
    unsigned* ticket = reinterpret_cast<unsigned*>(tmp.p);
    TO* tileSums = reinterpret_cast<TO*>(tmp.p + 256);
    TO* tileOffs = tileSums + nTiles + 1;
    dp_scan_sums_kernel<TO, F><<<(unsigned)nTiles, 256, 0, st>>>(load, n, tileSums, tileOffs, ticket);
    dp_scan_write_kernel<TO, F><<<(unsigned)nTiles, 256, 0, st>>>(load, n, tileOffs, out);
    CK(cudaGetLastError());
}

template <class TI, class TO>
void dp_exclusive_sum(const TI* in, TO* out, long long n, DBuf<unsigned char>& tmp, cudaStream_t st) {
    DpLoadPtr<TI> f{in};
    dp_exclusive_scan<TO>(f, out, n, tmp, st);
}

// ---------------------------------------------------------------------------------------------------------------
// radix sort
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dp_radix_hist_kernel(const unsigned long long* __restrict__ keys, long long n, int shift,
                                                            unsigned mask, unsigned* __restrict__ hist, unsigned nTiles) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * DP_SORT_TILE;
#pragma unroll
    for (int j = 0; j < DP_SORT_TILE / 256; j++) {
        const long long i = base + j * 256 + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nTiles + blockIdx.x] = h[threadIdx.x];  // digit-major: the scan orders by (digit, tile)
}

template <bool VALUES>
__global__ void __launch_bounds__(256) dp_radix_scatter_kernel(const unsigned long long* __restrict__ keys,
                                                               const unsigned* __restrict__ vals, long long n, int shift,
                                                               unsigned mask, const unsigned* __restrict__ histScan,
                                                               unsigned nTiles, unsigned long long* __restrict__ keysOut,
                                                               unsigned* __restrict__ valsOut) {
    constexpr int ROUNDS = DP_SORT_TILE / 256;
    __shared__ unsigned long long staged[DP_SORT_TILE];  // the tile in sorted order: runs of equal digits leave coalesced
    __shared__ unsigned short warpCount[8][256];
    __shared__ unsigned tileBase[256], digitStart[256], warpSums[8];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w = 0; w < 8; w++) warpCount[w][threadIdx.x] = 0;
    tileBase[threadIdx.x] = histScan[(size_t)threadIdx.x * nTiles + blockIdx.x];
    __syncthreads();
    const long long tile0 = (long long)blockIdx.x * DP_SORT_TILE;
    const long long first = tile0 + (long long)warp * (32 * ROUNDS);
    const int tileN = (int)min((long long)DP_SORT_TILE, n - tile0);
    unsigned long long key[ROUNDS];
    unsigned val[VALUES ? ROUNDS : 1];
    unsigned short off[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const long long i = first + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys[i] : 0ull;
        if (VALUES) val[r] = valid ? vals[i] : 0u;
        const unsigned d = (unsigned)(key[r] >> shift) & mask;
        const unsigned active = __ballot_sync(DP_FULL, valid);
        off[r] = 0;
        if (valid) {
            const unsigned peers = __match_any_sync(active, d);
            const unsigned rank = __popc(peers & dp_lanemask_lt());
            const unsigned before = warpCount[warp][d];
            __syncwarp(active);
            if (rank == 0) warpCount[warp][d] = (unsigned short)(before + __popc(peers));
            __syncwarp(active);
            off[r] = (unsigned short)(before + rank);
        }
    }
    __syncthreads();
    {   // per digit: the warps' counts -> exclusive prefix in warp order; the digit's total -> start of its run in the tile
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const unsigned c = warpCount[w][threadIdx.x];
            warpCount[w][threadIdx.x] = (unsigned short)run;
            run += c;
        }
        unsigned all;
        digitStart[threadIdx.x] = dp_block_exclusive<unsigned>(run, warpSums, &all);
    }
    __syncthreads();
    unsigned short at[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const long long i = first + r * 32 + lane;
        at[r] = 0;
        if (i < n) {
            const unsigned d = (unsigned)(key[r] >> shift) & mask;
            at[r] = (unsigned short)(digitStart[d] + warpCount[warp][d] + off[r]);
            staged[at[r]] = key[r];
        }
    }
    __syncthreads();
    unsigned char dig[ROUNDS];  // digit of the slots this thread writes out
#pragma unroll
    for (int j = 0; j < ROUNDS; j++) {
        const int i = (int)threadIdx.x + 256 * j;
        dig[j] = 0;
        if (i < tileN) {
            const unsigned long long kk = staged[i];
            const unsigned d = (unsigned)(kk >> shift) & mask;
            dig[j] = (unsigned char)d;
            keysOut[tileBase[d] + ((unsigned)i - digitStart[d])] = kk;
        }
    }
    if (VALUES) {  // the values ride through the same buffer
        unsigned* stagedV = reinterpret_cast<unsigned*>(staged);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ROUNDS; r++)
            if (first + r * 32 + lane < n) stagedV[at[r]] = val[r];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ROUNDS; j++) {
            const int i = (int)threadIdx.x + 256 * j;
            if (i < tileN) valsOut[tileBase[dig[j]] + ((unsigned)i - digitStart[dig[j]])] = stagedV[i];
        }
    }
}

// Stable sort of keysA[0..n) by bits [beginBit, endBit) into keysB (keysA is only read; keysC: scratch of n keys when more
// than one pass is needed). valsIn (may be null) travels with the keys into valsOut (valsTmp: scratch of n values when
// more than one pass is needed).
inline void dp_radix_sort(const unsigned long long* keysA, unsigned long long* keysB, unsigned long long* keysC,
                          const unsigned* valsIn, unsigned* valsOut, unsigned* valsTmp, long long n, int beginBit, int endBit,
                          DBuf<unsigned>& hist, DBuf<unsigned char>& tmp, cudaStream_t st) {
    if (n <= 0) return;
    if (n >= 0xffffffffll) throw std::runtime_error("radix sort: more than 2^32 - 1 keys");
    const int passes = std::max(1, (endBit - beginBit + 7) / 8);
    const unsigned nTiles = (unsigned)((n + DP_SORT_TILE - 1) / DP_SORT_TILE);
    hist.reserve((size_t)256 * nTiles + 1);
    const unsigned long long* kin = keysA;
    const unsigned* vin = valsIn;
    for (int p = 0; p < passes; p++) {
        const int shift = beginBit + 8 * p;
        const int bits = std::min(8, endBit - shift);
        const unsigned mask = bits >= 8 ? 255u : (bits <= 0 ? 0u : ((1u << bits) - 1u));
        unsigned long long* kout = ((passes - 1 - p) % 2 == 0) ? keysB : keysC;
        unsigned* vout = ((passes - 1 - p) % 2 == 0) ? valsOut : valsTmp;
        dp_radix_hist_kernel<<<nTiles, 256, 0, st>>>(kin, n, shift, mask, hist.p, nTiles);
        dp_exclusive_sum<unsigned, unsigned>(hist.p, hist.p, (long long)256 * nTiles, tmp, st);
        if (valsIn)
            dp_radix_scatter_kernel<true><<<nTiles, 256, 0, st>>>(kin, vin, n, shift, mask, hist.p, nTiles, kout, vout);
        else
            dp_radix_scatter_kernel<false><<<nTiles, 256, 0, st>>>(kin, nullptr, n, shift, mask, hist.p, nTiles, kout, nullptr);
        CK(cudaGetLastError());
        kin = kout;
        vin = vout;
    }
}

// unique of a sorted array: out = the first key of every run, *count = how many (device)
__global__ void dp_unique_write_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ pos, long long n,
                                       unsigned long long* __restrict__ out, unsigned long long* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool first = i == 0 || keys[i] != keys[i - 1];
    if (first) out[pos[i]] = keys[i];
    if (i == n - 1) *count = (unsigned long long)pos[i] + (first ? 1ull : 0ull);
}

inline void dp_unique_sorted(const unsigned long long* keys, unsigned long long* out, unsigned long long* count, long long n,
                             DBuf<unsigned>& pos, DBuf<unsigned char>& tmp, cudaStream_t st) {
    if (n <= 0) {
        CK(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
        return;
    }
    pos.reserve((size_t)n + 1);
    DpLoadFirstFlags f{keys};
    dp_exclusive_scan<unsigned>(f, pos.p, n, tmp, st);
    dp_unique_write_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, pos.p, n, out, count);
    CK(cudaGetLastError());
}
