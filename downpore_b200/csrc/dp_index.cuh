// downpore_b200 — index-construction kernels (sm_100a): pack, seed selection (AddSingleSeeds), chunk scan.
#pragma once
#include "dp_common.cuh"

// ---------------------------------------------------------------------------------------------------------------
// 2-bit pack (sequence.NewPackedSequence + asm packBytes, sequence/sequence.go:67-93, sequence/asm_amd64.s:33-78).
// One warp per sequence, one lane per 16-base output word and iteration: each lane reads its 16 ASCII bytes as five
// 4-byte-aligned words (the neighbouring lane's loads hit the same 32 B sectors, so DRAM traffic stays 1 B/base) and
// realigns with funnel shifts. Bases past the end of the sequence are packed as zero (the reference zero-pads its
// tail byte too).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned dp_pack4(unsigned w) {
    // four ASCII bytes (first base in the low byte) -> 8 bits, first base in the high bits
    unsigned c = ((w >> 1) ^ ((w & 0x04040404u) >> 2)) & 0x03030303u;
    return ((c & 0xff) << 6) | (((c >> 8) & 0xff) << 4) | (((c >> 16) & 0xff) << 2) | ((c >> 24) & 0xff);
}

__global__ void __launch_bounds__(256) dp_pack_kernel(const unsigned char* __restrict__ ascii,
                                                      const long long* __restrict__ seqOff,    // [n+1] byte offsets
                                                      const long long* __restrict__ wordOff,   // [n] first output word
                                                      unsigned* __restrict__ words, long long nSeq) {
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nWarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned lane = dp_lane();
    for (long long r = warp; r < nSeq; r += nWarps) {
        long long b0 = seqOff[r];
        long long len = seqOff[r + 1] - b0;
        long long nw = (len + 15) >> 4;
        unsigned* out = words + wordOff[r];
        const unsigned char* src = ascii + b0;
        unsigned mis = (unsigned)((unsigned long long)src & 3ull);
        const unsigned* src4 = (const unsigned*)(src - mis);
        unsigned sh = mis * 8;
        for (long long j = lane; j < nw; j += 32) {
            const unsigned* p = src4 + j * 4;
            long long remain = len - j * 16;  // bases available for this word (>0)
            unsigned a0 = __ldg(p), a1 = 0, a2 = 0, a3 = 0, a4 = 0;
            // never read past the last 4-byte word that holds a base of this sequence
            long long lastByte = (long long)mis + (remain < 16 ? remain : 16) - 1;  // offset from p, in bytes
            if (lastByte >= 4) a1 = __ldg(p + 1);
            if (lastByte >= 8) a2 = __ldg(p + 2);
            if (lastByte >= 12) a3 = __ldg(p + 3);
            if (lastByte >= 16) a4 = __ldg(p + 4);
            unsigned w0 = __funnelshift_r(a0, a1, sh);
            unsigned w1 = __funnelshift_r(a1, a2, sh);
            unsigned w2 = __funnelshift_r(a2, a3, sh);
            unsigned w3 = __funnelshift_r(a3, a4, sh);
            unsigned v = (dp_pack4(w0) << 24) | (dp_pack4(w1) << 16) | (dp_pack4(w2) << 8) | dp_pack4(w3);
            if (remain < 16) v &= ~0u << (unsigned)(2 * (16 - remain));
            out[j] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Windowed pack: only the windows that are queried are ever packed (Map() touches 2..12 windows of `edge` bases per
// read, never the middle of a long read: mapping/mapping.go:430-487). `ascii` may be device memory, pinned host
// memory mapped into the device address space (coalesced zero-copy reads of just the queried bytes: the few windows
// of later Map() rounds), or — `stage` — the HBM staging buffer dp_pull_windows_kernel has filled from such host
// memory (the bulk of the round-0 windows). One warp per window; per iteration the lanes load 32 consecutive 16-byte aligned blocks
// (one LDG.128 each, 512 contiguous bytes per warp), neighbours exchange blocks by shuffle, and 31 packed words are
// written. Each source byte crosses the bus once (+1/31 overlap).
// ---------------------------------------------------------------------------------------------------------------
// `byteOff` != null (dp_mapper_map_batch_spans): read r starts at ascii + byteOff[r] instead of ascii + seqOff[r] (reads
// mapped where they lie in a file image, name lines between them).
__global__ void __launch_bounds__(256, 6) dp_pack_windows_kernel(const unsigned char* __restrict__ ascii,
                                                              const long long* __restrict__ seqOff,
                                                              const long long* __restrict__ byteOff,
                                                              const long long* __restrict__ wordOff,
                                                              const DpWindow* __restrict__ wins, int nWin,
                                                              unsigned* __restrict__ words,
                                                              const unsigned char* __restrict__ stage,
                                                              const unsigned* __restrict__ stagePos) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nWarps = (gridDim.x * blockDim.x) >> 5;
    unsigned lane = dp_lane();
    for (int w = warp; w < nWin; w += nWarps) {
        DpWindow win = wins[w];
        if (win.len <= 0) continue;
        const long long readLen = seqOff[win.read + 1] - seqOff[win.read];
        const long long readBase = byteOff ? byteOff[win.read] : seqOff[win.read];
        // packed words covering the window, plus the following word when it still holds bases of the read
        long long w0 = win.start >> 4;
        long long w1 = ((long long)win.start + win.len - 1) >> 4;  // last word holding a window base
        long long wLast = (readLen - 1) >> 4;                      // last word holding a read base
        if (w1 < wLast) w1++;                                      // k-mer extraction reads one word ahead
        const long long nWords = w1 - w0 + 1;
        unsigned* out = words + wordOff[win.read] + w0;
        const unsigned char* src = ascii + readBase + w0 * 16;     // first byte of word w0
        const unsigned long long addr = (unsigned long long)src;
        const unsigned mis = (unsigned)(addr & 15ull);
        const uint4* blk = (const uint4*)(src - mis);              // aligned block holding the first byte
        // (staged: dp_pull_windows_kernel has copied these blocks to stage + 16 * stagePos[w])
        if (stage) blk = (const uint4*)stage + stagePos[w];
        const unsigned char* srcEnd = ascii + readBase + readLen;  // one past the last byte of the read
        const long long lastBlk = ((long long)((unsigned long long)(srcEnd - 1) - (unsigned long long)(src - mis))) >> 4;
        const unsigned q = mis >> 2, sh = (mis & 3) * 8;
        // three iterations per trip, all three block loads issued before the first is consumed: the pull is bound by
        // the latency of the link, so bytes in flight per warp are what buys bandwidth (a 1000-base window = one trip)
        for (long long g00 = 0; g00 < nWords; g00 += 93) {
            uint4 vv[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
                long long b = g00 + 31 * u + lane;  // aligned block index this lane loads
                vv[u] = make_uint4(0, 0, 0, 0);
                if (b <= lastBlk && b <= nWords) vv[u] = __ldg(blk + b);  // never past the block holding the read's last byte
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const long long g0 = g00 + 31 * u;
                if (g0 >= nWords) break;  // warp-uniform
                const uint4 v = vv[u];
                uint4 nx;
                nx.x = __shfl_down_sync(DP_FULL, v.x, 1);
                nx.y = __shfl_down_sync(DP_FULL, v.y, 1);
                nx.z = __shfl_down_sync(DP_FULL, v.z, 1);
                nx.w = __shfl_down_sync(DP_FULL, v.w, 1);
                // 16 source bytes of word g = bytes [mis, mis+16) of (v, nx)
                unsigned a0, a1, a2, a3, a4;
                switch (q) {  // warp-uniform
                    case 0: a0 = v.x; a1 = v.y; a2 = v.z; a3 = v.w; a4 = nx.x; break;
                    case 1: a0 = v.y; a1 = v.z; a2 = v.w; a3 = nx.x; a4 = nx.y; break;
                    case 2: a0 = v.z; a1 = v.w; a2 = nx.x; a3 = nx.y; a4 = nx.z; break;
                    default: a0 = v.w; a1 = nx.x; a2 = nx.y; a3 = nx.z; a4 = nx.w; break;
                }
                long long g = g0 + lane;
                if (lane < 31 && g < nWords) {
                    unsigned x0 = __funnelshift_r(a0, a1, sh);
                    unsigned x1 = __funnelshift_r(a1, a2, sh);
                    unsigned x2 = __funnelshift_r(a2, a3, sh);
                    unsigned x3 = __funnelshift_r(a3, a4, sh);
                    unsigned pv = (dp_pack4(x0) << 24) | (dp_pack4(x1) << 16) | (dp_pack4(x2) << 8) | dp_pack4(x3);
                    long long remain = readLen - (w0 + g) * 16;  // bases of the read in this word
                    if (remain < 16) pv &= remain > 0 ? (~0u << (unsigned)(2 * (16 - remain))) : 0u;
                    out[g] = pv;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Windowed pack for reads that ARRIVE PACKED (dp_mapper_map_batch_packed): the caller hands over what the reference's
// own reader produces, sequence.packedSequence bytes (sequence/sequence.go:67-93: four bases per byte, first base in the
// two most significant bits, tail byte left-aligned and zero-padded) — a quarter of a byte per base crosses the link
// instead of one. A read starts at any byte of the batch buffer, so the window's bytes are realigned with funnel shifts
// and byte-swapped into the 16-bases-per-word layout (first base in the top bits of the word). One warp per window; a
// lane loads one aligned 16-byte block (out of device memory, or mapped pinned host memory: coalesced zero-copy reads
// of just the queried bytes) and writes four words; neighbours exchange blocks by shuffle.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 6) dp_pack_windows_packed_kernel(const unsigned char* __restrict__ packed,
                                                                        const long long* __restrict__ byteOff,  // [n] first byte of each read
                                                                        const long long* __restrict__ seqOff,   // [n+1] cumulative bases
                                                                        const long long* __restrict__ wordOff,
                                                                        const DpWindow* __restrict__ wins, int nWin,
                                                                        unsigned* __restrict__ words,
                                                                        const unsigned char* __restrict__ stage,
                                                                        const unsigned* __restrict__ stagePos) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nWarps = (gridDim.x * blockDim.x) >> 5;
    unsigned lane = dp_lane();
    for (int w = warp; w < nWin; w += nWarps) {
        DpWindow win = wins[w];
        if (win.len <= 0) continue;
        const long long readLen = seqOff[win.read + 1] - seqOff[win.read];
        long long w0 = win.start >> 4;
        long long w1 = ((long long)win.start + win.len - 1) >> 4;  // last word holding a window base
        long long wLast = (readLen - 1) >> 4;                      // last word holding a read base
        if (w1 < wLast) w1++;                                      // k-mer extraction reads one word ahead
        const long long nWords = w1 - w0 + 1;
        unsigned* out = words + wordOff[win.read] + w0;
        const unsigned char* rd = packed + byteOff[win.read];
        const unsigned char* src = rd + w0 * 4;                    // first byte of word w0
        const unsigned mis = (unsigned)((unsigned long long)src & 15ull);
        const uint4* blk = (const uint4*)(src - mis);              // aligned block holding that byte
        // (staged: dp_pull_windows_kernel has copied these blocks to stage + 16 * stagePos[w])
        if (stage) blk = (const uint4*)stage + stagePos[w];
        const unsigned char* srcEnd = rd + ((readLen + 3) >> 2);   // one past the read's last byte
        const long long lastBlk = ((long long)((unsigned long long)(srcEnd - 1) - (unsigned long long)(src - mis))) >> 4;
        const long long needBlk = ((long long)mis + 4 * nWords - 1) >> 4;  // last block holding a byte of these words
        const unsigned q = mis >> 2, sh = (mis & 3) * 8;
        for (long long b0 = 0; b0 * 4 < nWords; b0 += 31) {  // 31 output blocks of four words per trip
            const long long b = b0 + lane;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (b <= lastBlk && b <= needBlk) v = __ldg(blk + b);  // never past the block holding the read's last byte
            uint4 nx;
            nx.x = __shfl_down_sync(DP_FULL, v.x, 1);
            nx.y = __shfl_down_sync(DP_FULL, v.y, 1);
            nx.z = __shfl_down_sync(DP_FULL, v.z, 1);
            nx.w = __shfl_down_sync(DP_FULL, v.w, 1);
            unsigned a0, a1, a2, a3, a4;
            switch (q) {  // warp-uniform
                case 0: a0 = v.x; a1 = v.y; a2 = v.z; a3 = v.w; a4 = nx.x; break;
                case 1: a0 = v.y; a1 = v.z; a2 = v.w; a3 = nx.x; a4 = nx.y; break;
                case 2: a0 = v.z; a1 = v.w; a2 = nx.x; a3 = nx.y; a4 = nx.z; break;
                default: a0 = v.w; a1 = nx.x; a2 = nx.y; a3 = nx.z; a4 = nx.w; break;
            }
            if (lane < 31) {
                const unsigned x[4] = {__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh),
                                       __funnelshift_r(a3, a4, sh)};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const long long g = 4 * b + t;
                    if (g < nWords) {
                        unsigned pv = __byte_perm(x[t], 0u, 0x0123);  // first byte of the read on top
                        long long remain = readLen - (w0 + g) * 16;   // bases of the read in this word
                        if (remain < 16) pv &= remain > 0 ? (~0u << (unsigned)(2 * (16 - remain))) : 0u;
                        out[g] = pv;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The PCIe leg of the windowed pack as a copy-engine-like kernel: the 16-byte blocks that dp_pack_windows_kernel reads
// for window w are moved from the caller's pinned buffer (host memory mapped into the device address space) to slot w
// of a staging buffer in HBM by TMA bulk copies — host -> shared memory (cp.async.bulk, completion on an mbarrier) ->
// HBM (bulk store) — issued by ONE thread per CTA over a ring of shared-memory slots. A few dozen such CTAs saturate
// the link (measured: 16 CTAs x 8 slots = 49.8 GB/s of 2 KB pieces, the same as 148 x 8 warps of LDG.128,
// scripts/pcie_probe2.cu), so the SMs stay free for the compute kernels of the other lanes; the pack itself then runs
// from HBM. The other 31 lanes of the warp only prepare the copy descriptors, 32 windows at a time.
// ---------------------------------------------------------------------------------------------------------------
// ring slots per CTA / loads in flight per CTA (the other slots are being stored). Alone, 16 CTAs pull 250-560 byte pieces
// at 45 GB/s (scripts/pcie_probe3.cu); next to the compute kernels of the other lanes the pull of a step takes 31 ms with
// 32 CTAs, 21 with 64, 24 with 128. Measured and rejected: a deeper ring (32 slots, 28 loads in flight: same), eight
// lanes of the warp issuing copies side by side (16 CTAs x 8 lanes: 27 ms, 64 x 8: 22-23) — what the pull runs out of is
// the memory path of the SMs it shares with the compute kernels, neither bytes in flight nor issue slots: it wants to be
// spread over about 64 SMs.
#define DP_PULL_SLOTS_ASCII 16
#define DP_PULL_SLOTS_PACKED 16

// `byteOff` != null: read r starts at ascii + byteOff[r]; `packed`: the reads are packedSequence bytes (four bases per
// byte) and the blocks moved are the ones dp_pack_windows_packed_kernel reads.
template <int DP_PULL_SLOTS, int DP_PULL_AHEAD>
__global__ void __launch_bounds__(32) dp_pull_windows_kernel(const unsigned char* __restrict__ ascii,
                                                             const long long* __restrict__ seqOff,
                                                             const long long* __restrict__ byteOff, bool packed,
                                                             const DpWindow* __restrict__ wins, int nWin,
                                                             unsigned char* __restrict__ stage, int stageStride,
                                                             unsigned* __restrict__ stagePos, unsigned* __restrict__ work) {
    extern __shared__ __align__(128) unsigned char dp_pull_smem[];  // DP_PULL_SLOTS x 2 * stageStride
    __shared__ __align__(8) unsigned long long bars[DP_PULL_SLOTS];
    __shared__ unsigned long long dSrc[2][32];
    __shared__ unsigned dBytes[2][32];
    __shared__ unsigned dDst[2][32];  // destination in the staging buffer, in 16-byte blocks
    const unsigned lane = dp_lane();
    const unsigned slot0 = (unsigned)__cvta_generic_to_shared(dp_pull_smem);
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bars);
    if (lane == 0) {
        for (int s = 0; s < DP_PULL_SLOTS; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // Work is handed out in chunks of 32 consecutive windows through a counter (`work`, zero at launch): the CTAs share
    // their SMs with the compute kernels of the other lanes and run at different speeds.
    const unsigned nChunks = ((unsigned)nWin + 31u) >> 5;
    auto grab = [&]() {
        unsigned c = 0;
        if (lane == 0) c = atomicAdd(work, 1u);
        return __shfl_sync(DP_FULL, c, 0);
    };
    // copy descriptors of the windows of chunk c into buffer `buf`
    auto describe = [&](unsigned c, int buf) {
        const int w = (int)(32u * c + lane);
        unsigned long long src = 0;
        unsigned bytes = 0;
        if (w < nWin) {
            const DpWindow win = wins[w];
            if (win.len > 0) {
                const long long readLen = seqOff[win.read + 1] - seqOff[win.read];
                const long long readBase = byteOff ? byteOff[win.read] : seqOff[win.read];
                const long long w0 = win.start >> 4;
                long long w1 = ((long long)win.start + win.len - 1) >> 4;
                if (w1 < ((readLen - 1) >> 4)) w1++;
                const long long nWords = w1 - w0 + 1;
                if (packed) {
                    const unsigned char* rd = ascii + readBase;
                    const unsigned char* s0 = rd + w0 * 4;
                    const unsigned mis = (unsigned)((unsigned long long)s0 & 15ull);
                    const unsigned char* blk = s0 - mis;
                    const long long lastBlk = (long long)((unsigned long long)(rd + ((readLen + 3) >> 2) - 1) - (unsigned long long)blk) >> 4;
                    const long long needBlk = ((long long)mis + 4 * nWords - 1) >> 4;
                    src = (unsigned long long)blk;
                    bytes = (unsigned)(min(lastBlk, needBlk) + 1) * 16u;
                } else {
                    const unsigned char* s0 = ascii + readBase + w0 * 16;
                    const unsigned mis = (unsigned)((unsigned long long)s0 & 15ull);
                    const unsigned char* blk = s0 - mis;
                    const long long lastBlk = (long long)((unsigned long long)(ascii + readBase + readLen - 1) - (unsigned long long)blk) >> 4;
                    src = (unsigned long long)blk;
                    bytes = (unsigned)(min(lastBlk, nWords) + 1) * 16u;
                }
            }
        }
        // A window that starts where its predecessor ends (the tail window of one read and the head window of the
        // next are neighbours in the caller's buffer) rides in the predecessor's copy: half as many, twice as long
        // PCIe reads. Pairs only: a follower's predecessor is never a follower itself.
        const unsigned long long pSrc = __shfl_up_sync(DP_FULL, src, 1);
        const unsigned pBytes = __shfl_up_sync(DP_FULL, bytes, 1);
        const bool adj = lane > 0 && bytes && pBytes && src >= pSrc && src <= pSrc + pBytes + 64ull &&
                         (src - pSrc) + bytes <= 2ull * (unsigned long long)stageStride;
        const bool pAdj = __shfl_up_sync(DP_FULL, adj, 1);
        const bool follower = adj && !pAdj;
        const unsigned long long nSrc = __shfl_down_sync(DP_FULL, src, 1);
        const unsigned nBytes = __shfl_down_sync(DP_FULL, bytes, 1);
        const bool nFollower = __shfl_down_sync(DP_FULL, follower, 1) && lane < 31;
        const unsigned base = (unsigned)w * (unsigned)(stageStride >> 4);  // in 16-byte blocks
        const unsigned dst = follower ? base - (unsigned)(stageStride >> 4) + (unsigned)((src - pSrc) >> 4) : base;
        if (w < nWin) stagePos[w] = dst;
        if (nFollower) bytes = (unsigned)(nSrc - src) + nBytes;  // leader: one copy covers both windows
        if (follower) bytes = 0;
        dSrc[buf][lane] = src;
        dBytes[buf][lane] = bytes;
        dDst[buf][lane] = dst;
    };
    // lane 0 numbers the windows it handles 0, 1, 2, ... (32 per chunk taken, bytes = 0 where there is nothing to copy);
    // window i uses ring slot i % DP_PULL_SLOTS and descriptor buffer (i / 32) & 1
    int issued = 0;           // lane 0: windows whose load has been issued
    unsigned phaseBits = 0;   // lane 0: parity each slot's mbarrier completes next (windows without bytes skip a use)
    unsigned cur = grab();
    if (cur < nChunks) describe(cur, 0);
    for (int k = 0; cur < nChunks; k++) {
        const unsigned nxt = grab();
        if (nxt < nChunks) describe(nxt, (k + 1) & 1);  // (ready before lane 0 runs ahead into it)
        __syncwarp();
        if (lane == 0) {
            const int ahead = nxt < nChunks ? 32 * (k + 2) : 32 * (k + 1);
            for (int i = 32 * k; i < 32 * (k + 1); i++) {
                const int lim = min(ahead, i + DP_PULL_AHEAD);
                while (issued < lim) {
                    const int j = issued;
                    const unsigned bytes = dBytes[(j >> 5) & 1][j & 31];
                    if (bytes) {
                        if (j >= DP_PULL_SLOTS)  // the store that last read this slot has finished reading it
                            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DP_PULL_SLOTS - DP_PULL_AHEAD - 1) : "memory");
                        const unsigned s = (unsigned)(j % DP_PULL_SLOTS);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(bytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(slot0 + s * 2u * (unsigned)stageStride), "l"(dSrc[(j >> 5) & 1][j & 31]), "r"(bytes), "r"(bar0 + 8u * s)
                                     : "memory");
                    }
                    issued++;
                }
                const unsigned bytes = dBytes[(i >> 5) & 1][i & 31];
                const unsigned s = (unsigned)(i % DP_PULL_SLOTS);
                if (bytes) {
                    const unsigned parity = (phaseBits >> s) & 1u;
                    phaseBits ^= 1u << s;
                    unsigned ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(ok) : "r"(bar0 + 8u * s), "r"(parity) : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(stage + (size_t)dDst[(i >> 5) & 1][i & 31] * 16u), "r"(slot0 + s * 2u * (unsigned)stageStride), "r"(bytes)
                                 : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");  // one group per window, empty or not
            }
        }
        __syncwarp();
        cur = nxt;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// packed words (16 bases, MSB first) -> the reference's byte layout (4 bases per byte, MSB first): byte i of the
// sequence is bits [31-8*(i%4) .. 24-8*(i%4)] of word i/4.
__global__ void dp_words_to_bytes_kernel(const unsigned* __restrict__ words, unsigned char* __restrict__ out,
                                         long long nBytes) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nBytes) out[i] = (unsigned char)(words[i >> 2] >> (24 - 8 * (unsigned)(i & 3)));
}

// ---------------------------------------------------------------------------------------------------------------
// sequtil.KmerOccurrences (util/sequtil/kmers.go:53-69): dense 4^k histogram of every k-mer of one sequence.
// ---------------------------------------------------------------------------------------------------------------
__global__ void dp_kmer_hist_kernel(const unsigned* __restrict__ words, long long len, int k,
                                    unsigned long long* __restrict__ counts) {
    long long n = len - k + 1;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        atomicAdd(counts + dp_kmer_at(words, p, k), 1ull);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// AddSingleSeeds (seeds/seeds.go:160-200) as a parallel fixed point.
//
// Window w covers bases [w*rate, (w+1)*rate). The reference walks windows in order; window w adds its best k-mer
// (argmax of values over all k-mers starting in the window, first maximum wins) iff none of the k-mers it *counts*
// (A_w, the Q4 subset visited by packedCountKmers on the byte-rounded sub-slice) is a seed yet. best_w does not
// depend on the evolving seed set; only need_w does:
//     need_w  <=>  for all x in A_w :  f(x) >= w,     f(x) = min { w' : best_w' = x and need_w' }
// (f(x) = the window at which x becomes a seed). need_0 is true, and need_w only depends on need_w' for w' < w, so
// iterating need <- F(need) from any start reaches the unique solution after (dependency depth) rounds.
// ---------------------------------------------------------------------------------------------------------------
struct DpSeedSelParams {
    long long refLen;
    long long nWindows;
    int rate;
    int k;
    int skipBack;  // 4 - finalLen of the raw reference (finalLen = refLen % 4, 0 when divisible: Q2)
};

// the k-mer positions A_w = [aStart, aStart + aCount) counted for window w (sequence.go:332-337 + asm :81-203)
__device__ __forceinline__ void dp_counted_range(const DpSeedSelParams& P, long long w, long long* aStart, int* aCount) {
    long long i = w * P.rate;
    long long startByte = (i + 3) / 4;        // firstLen = 4 for the raw reference
    long long endByte = (i + P.rate) / 4;
    long long nb = endByte - startByte;
    long long r8 = (nb - 1) * 4 - P.skipBack - P.k + 1;
    long long r15 = r8 & 3;
    r8 &= ~3ll;
    long long groups = (r8 >= 8) ? (r8 >> 2) : 1;  // do-while: at least one block of four (Q1)
    *aStart = startByte * 4;
    *aCount = (int)(4 + groups * 4 + r15);         // 4 "initial" k-mers (skipFront = 0) + internal + tail
}

__global__ void dp_seed_best_kernel(const unsigned* __restrict__ words, const double* __restrict__ values,
                                    DpSeedSelParams P, unsigned* __restrict__ best) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P.nWindows) return;
    long long i = w * P.rate;
    unsigned kmer = dp_kmer_at(words, i, P.k);
    double bestValue = __ldg(values + kmer);
    unsigned bestKmer = kmer;
    for (long long j = i + 1; j + P.k <= i + P.rate; j++) {  // k-mers starting at i+1 .. i+rate-k
        kmer = dp_kmer_at(words, j, P.k);
        double v = __ldg(values + kmer);
        if (v > bestValue) {
            bestValue = v;
            bestKmer = kmer;
        }
    }
    best[w] = bestKmer;
}

__global__ void dp_seed_fmin_kernel(const unsigned* __restrict__ best, const unsigned char* __restrict__ need,
                                    long long nWindows, unsigned* __restrict__ f) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nWindows && need[w]) atomicMin(f + best[w], (unsigned)w);
}

__global__ void dp_seed_need_kernel(const unsigned* __restrict__ words, const unsigned* __restrict__ f,
                                    DpSeedSelParams P, const unsigned char* __restrict__ needIn,
                                    unsigned char* __restrict__ needOut, unsigned* __restrict__ changed) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P.nWindows) return;
    long long aStart;
    int aCount;
    dp_counted_range(P, w, &aStart, &aCount);
    bool need = true;
    for (int t = 0; t < aCount && need; t++) {
        unsigned x = dp_kmer_at(words, aStart + t, P.k);
        if (__ldg(f + x) < (unsigned)w) need = false;
    }
    needOut[w] = need ? 1 : 0;
    if ((needIn[w] != 0) != need) *changed = 1;
}

__global__ void dp_seed_setbits_kernel(const unsigned* __restrict__ best, const unsigned char* __restrict__ need,
                                       long long nWindows, unsigned* __restrict__ bits) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nWindows && need[w]) atomicOr(bits + (best[w] >> 5), 1u << (best[w] & 31));
}

__global__ void dp_popc_kernel(const unsigned* __restrict__ bits, unsigned* __restrict__ pc, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pc[i] = __popc(bits[i]);
}
__global__ void dp_table_kernel(const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix,
                                uint2* __restrict__ table, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) table[i] = make_uint2(bits[i], prefix[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// Chunk scan = NewSeedSequence on every reference chunk (seeds/seeds.go:33-50; mapping/mapping.go:79-106).
// One warp per chunk; 32 consecutive k-mer positions per iteration; ballot + popc give the ordered compaction.
// pass 0 counts, pass 1 writes (pos, seed rank) in scan order and the (seed, chunk) sort keys.
// ---------------------------------------------------------------------------------------------------------------
struct DpChunkDesc {
    long long base;  // first base of the chunk in the packed reference array
    int nVisit;      // k-mers visited by the scan (Q2: raw sequences with len%4==0 lose four)
    int pad;
};

__global__ void __launch_bounds__(256) dp_chunk_scan_kernel(const unsigned* __restrict__ words,
                                                            const uint2* __restrict__ table,
                                                            const DpChunkDesc* __restrict__ chunks, unsigned nChunks,
                                                            int k, int pass, unsigned* __restrict__ counts,
                                                            const unsigned* __restrict__ chunkOff,
                                                            int* __restrict__ chunkPos, unsigned* __restrict__ chunkSeed,
                                                            unsigned long long* __restrict__ keys) {
    unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned nWarps = (gridDim.x * blockDim.x) >> 5;
    unsigned lane = dp_lane();
    unsigned lt = dp_lanemask_lt();
    for (unsigned c = warp; c < nChunks; c += nWarps) {
        DpChunkDesc d = chunks[c];
        unsigned n = 0;
        unsigned off = pass ? chunkOff[c] : 0;
        for (int j0 = 0; j0 < d.nVisit; j0 += 32) {
            int j = j0 + (int)lane;
            bool hit = false;
            unsigned rank = 0;
            if (j < d.nVisit) hit = dp_seed_lookup(table, dp_kmer_at(words, d.base + j, k), &rank);
            unsigned m = __ballot_sync(DP_FULL, hit);
            if (pass && hit) {
                unsigned idx = off + n + __popc(m & lt);
                chunkPos[idx] = j;
                chunkSeed[idx] = rank;
                if (keys) keys[idx] = ((unsigned long long)rank << 32) | c;
            }
            n += __popc(m);
        }
        if (!pass && lane == 0) counts[c] = n;
    }
}

// every (seed, chunk) key (sorted, with duplicates) -> per-seed occurrence counts and the chunk column of the
// position-carrying postings
__global__ void dp_posting_all_kernel(const unsigned long long* __restrict__ keys, long long n,
                                      unsigned* __restrict__ seedCount, unsigned* __restrict__ postChunk) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned long long key = keys[i];
        atomicAdd(seedCount + (unsigned)(key >> 32), 1u);
        postChunk[i] = (unsigned)key;
    }
}

// prefix filter over the seed k-mers (shared-memory prefilter of the extract kernel)
__global__ void dp_filter_build_kernel(const uint2* __restrict__ table, long long nTable, int k, int bits,
                                       unsigned* __restrict__ filter) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nTable) return;
    unsigned f = table[w].x;
    while (f) {
        int b = __ffs(f) - 1;
        f &= f - 1;
        unsigned h = dp_filter_hash((unsigned)(w * 32 + b), k, bits);
        atomicOr(filter + (h >> 5), 1u << (h & 31));
    }
}

// sorted unique (seed, chunk) keys -> per-seed run lengths and the chunk column
__global__ void dp_posting_fill_kernel(const unsigned long long* __restrict__ keys, long long n,
                                       unsigned* __restrict__ seedCount, unsigned* __restrict__ seedChunks) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned long long key = keys[i];
        atomicAdd(seedCount + (unsigned)(key >> 32), 1u);
        seedChunks[i] = (unsigned)key;
    }
}

// ----------------------------------------------------------------------------------------------------------------
// The mid-lookup copy of the seed -> chunks runs (DpIndexDev::midOff / midPost)
// ----------------------------------------------------------------------------------------------------------------
__global__ void dp_mid_blocks_kernel(const unsigned* __restrict__ seedOff, unsigned numSeeds, unsigned* __restrict__ blocks) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s <= numSeeds) blocks[s] = s < numSeeds ? (seedOff[s + 1] - seedOff[s] + 3u) >> 2 : 0u;
}

// padWord: byte offset of the first dummy counter word (a padding posting adds zero to one of the 32 dummy words)
__global__ void dp_mid_fill_kernel(const unsigned* __restrict__ seedOff, const unsigned* __restrict__ seedChunks,
                                   const unsigned* __restrict__ midOff, unsigned numSeeds, unsigned padWord,
                                   unsigned* __restrict__ midPost, uint4* __restrict__ midSeed) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const unsigned nWarps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned s = warp; s < numSeeds; s += nWarps) {
        const unsigned o = seedOff[s], n = seedOff[s + 1] - o;
        const unsigned b = midOff[s] * 4u, slots = (midOff[s + 1] - midOff[s]) * 4u;
        if (lane == 0) midSeed[s] = make_uint4(o, n, midOff[s], n ? seedChunks[o + n - 1] >> 6 : 0u);
        for (unsigned i = lane; i < slots; i += 32) {
            unsigned p;
            if (i < n) {
                const unsigned c = seedChunks[o + i];
                p = ((c & ~3u) << 16) | (0x1111u ^ (1u << ((c & 3u) << 2)));
            } else {
                p = ((padWord + (i & 31u) * 4u) << 16) | 0x1111u;
            }
            midPost[b + i] = p;
        }
    }
}
