// downpore_b200 — host driver and C ABI (include/downpore_b200.h). No CPU fallback: everything fails loudly without
// a usable CUDA device.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/downpore_b200.h"
#include "dp_common.cuh"
#include "dp_host.hpp"
#include "dp_finish.cuh"
#include "dp_index.cuh"
#include "dp_io.cuh"
#include "dp_map.cuh"
#include "dp_rounds.cuh"
#include "dp_sort.cuh"

namespace {

thread_local std::string g_err;


double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Large result arrays are recycled through dp_free: a fresh 40 MB malloc per call costs ten thousand page faults on the
// threads that fill it (milliseconds of a 20 ms step). Blocks of a megabyte and more that come back through dp_free are
// kept (a handful, bounded in bytes) and handed out again with their pages in place; everything else is plain malloc/free.
struct ResultBlocks {
    struct Block {
        void* p;
        size_t bytes;
    };
    std::mutex mu;
    std::vector<Block> live, idle;
    size_t idleBytes = 0;
};
ResultBlocks& result_blocks() {
    static ResultBlocks* rb = new ResultBlocks();  // (never destroyed: dp_free may run from finalisers at exit)
    return *rb;
}
const size_t kResultBlockMin = 1u << 20, kResultIdleMax = 8, kResultIdleBytes = 2ull << 30;

void* result_alloc(size_t bytes) {
    if (bytes < kResultBlockMin) return malloc(bytes ? bytes : 1);
    ResultBlocks& rb = result_blocks();
    {
        std::lock_guard<std::mutex> lk(rb.mu);
        size_t best = rb.idle.size();
        for (size_t i = 0; i < rb.idle.size(); i++)
            if (rb.idle[i].bytes >= bytes && rb.idle[i].bytes / 4 <= bytes && (best == rb.idle.size() || rb.idle[i].bytes < rb.idle[best].bytes))
                best = i;
        if (best < rb.idle.size()) {
            const ResultBlocks::Block b = rb.idle[best];
            rb.idle.erase(rb.idle.begin() + (long)best);
            rb.idleBytes -= b.bytes;
            rb.live.push_back(b);
            return b.p;
        }
    }
    void* p = malloc(bytes);
    if (p) {
        std::lock_guard<std::mutex> lk(rb.mu);
        rb.live.push_back({p, bytes});
    }
    return p;
}

void result_free(void* p) {
    if (!p) return;
    ResultBlocks& rb = result_blocks();
    {
        std::lock_guard<std::mutex> lk(rb.mu);
        for (size_t i = 0; i < rb.live.size(); i++)
            if (rb.live[i].p == p) {
                const ResultBlocks::Block b = rb.live[i];
                rb.live.erase(rb.live.begin() + (long)i);
                if (rb.idle.size() < kResultIdleMax && rb.idleBytes + b.bytes <= kResultIdleBytes) {
                    rb.idle.push_back(b);
                    rb.idleBytes += b.bytes;
                    return;
                }
                break;
            }
    }
    free(p);
}

}  // namespace

// Per-lane workspace: one stream plus every buffer a sub-batch needs. Two lanes let the host work of one sub-batch
// (result assembly, waiting for the rare later rounds of Map()) overlap the kernels of the next.
// Capacities of one launch of the window kernels. None is a limit of the library: a launch that runs out of one flags
// it (DpCounters.overflow), the host grows it and recomputes the affected reads (map_range below) — the reference keeps
// every hit (mapping/mapping.go:504 sizes `results` by the candidate count, :518-552 append without bound).
struct Caps {
    int outStride = 16;         // window results kept per window ON AVERAGE over a launch (one bump-allocated pool)
    int resultCap = 128;        // mappings of one window before sort/dedupe (general chain kernel)
    int chainCap = 64;          // chains kept for one candidate (general chain kernel)
    unsigned candStride = 1024; // candidate chunks per window strand (min(C, .))
    int roundsScale = 1;        // later rounds of Map(): scratch per read and window cache, in units of their defaults
};
const int kCapMax = 1 << 20;

struct Lane {
    double hp[6] = {0, 0, 0, 0, 0, 0};  // DP_HOST_PROFILE: host ms in tables / launches / wait / post / later rounds / assemble
    cudaStream_t stream = nullptr;
    Caps caps;
    HBuf<DpCounters> hCtr;  // counters + overflow flags of the current attempt, as of its last synchronise
    cudaEvent_t evSync = nullptr;  // cudaEventBlockingSync: the lane's thread sleeps instead of spinning (sync_mode_blocking)
    DBuf<unsigned char> dAscii, dStage;
    DBuf<long long> dSeqOff, dWordOff, dWordsNeeded;
    DBuf<int> dReadLen;
    DBuf<unsigned> dWords;
    DBuf<DpWindow> dWins;
    DBuf<unsigned> wsOff, qSeed, candChunk, outOff, dFinOff, dStagePos, dPullWork;
    DBuf<int> wsN, qPos, candN, outN, dFinN;
    DBuf<unsigned short> candDistinct;
    DBuf<DpMappingDev> outMaps;
    DBuf<unsigned long long> cursor;
    DBuf<DpCounters> dCtr;
    DBuf<unsigned char> scanTmp;
    // lookup scratch
    DBuf<unsigned> lsSeed, lsOff, lsPre, lsEndW, lsAll, lsCounters, lsTouched;
    DBuf<unsigned long long> lsCand;
    bool lsCountersZeroed = false;
    DBuf<int> lbDefer;        // block lookup kernel: deferred window strands
    DBuf<unsigned> lbWork;    // [0] work counter, [1] deferred count
    DBuf<unsigned char> lsFirst;
    DBuf<unsigned short> lsOrder;
    // chain scratch
    DBuf<unsigned short> csQFirst, csQCnt, csRqId, csRsId;
    DBuf<unsigned> csQLo;
    DBuf<unsigned long long> csEnt;
    DBuf<int> csRqPos, csRsPos, csChainLen, csLastB, csChains;
    DBuf<DpMappingDev> csResults;
    // fast chain path: tasks, list pool, hand-back list
    DBuf<DpChainTask> fcTasks;
    DBuf<unsigned> fcTaskBase, fcPool;
    DBuf<unsigned long long> fcCursors;  // [0] tasks, [1] pool words, [2] low word = hand-back count
    DBuf<unsigned char> fcSlow;
    DBuf<int> fcSlowList;
    size_t fcTaskCap = 0, fcPoolCap = 0;
    HBuf<int> hOutN, hUnresN;
    HBuf<unsigned> hOutOff, hFinOff;
    HBuf<unsigned char> hStage;
    HBuf<DpUnresolved> hUnres;
    DBuf<DpUnresolved> dUnres;
    // later rounds of Map() on the device (dp_rounds.cuh): window cache, requests, results, per-thread scratch
    DBuf<int> rdHead, rdWinSlot, rdResN, rdLists;
    DBuf<int4> rdEnt;
    DBuf<int2> rdEnt2;
    DBuf<DpMappingDev> rdCache, rdResMaps;
    DBuf<unsigned char> rdDone;
    DBuf<unsigned> rdResOff, rdCur;
    DBuf<DpHit> rdHits;
    DpRoundsDev roundsDev{};  // the rounds in flight on this lane (rounds_begin -> rounds_run)
    int rdThreads = 0;
    std::unique_ptr<Lane> rounds;    // workspace + stream for the later rounds of this lane's previous sub-batch
    cudaEvent_t evRounds = nullptr;  // round-0 hits filed in the rounds workspace
    HBuf<unsigned> hRdCur, hRdResOff;
    HBuf<int> hRdResN;
    HBuf<DpMappingDev> hRdResMaps;
    HBuf<DpMappingDev> hOutMaps, hFinMaps;
    DBuf<DpMappingDev> dFinMaps;  // the finished records of a sub-batch in read order, before their copy to hFinMaps
    HBuf<long long> hRel;
    size_t hOutTotal = 0;
    bool pendingStageTimes = false;
    bool reduceTimed = false;
    const unsigned char* curAscii = nullptr;  // device-visible reads of the current sub-batch (device or mapped host)
    bool curAsciiIsHost = false;
    // Buffers are sized for the largest sub-batch of the current call the first time a lane touches them (floors set by
    // map_batch_impl), not grown piece by piece as the lane happens to meet larger sub-batches: a cudaFree in the middle
    // of a call synchronises the device under every other lane.
    size_t floorReads = 0, floorWins = 0, floorSeeds = 0, floorBytes = 0;
    bool curSpans = false;   // ASCII reads at arbitrary places of the caller's buffer (dp_mapper_map_batch_spans): dByteOff
    bool curPacked = false;  // the reads are sequence.packedSequence bytes (dp_mapper_map_batch_packed), not ASCII
    DBuf<long long> dByteOff;  // packed input: first byte of each read, relative to curAscii
    HBuf<long long> hByteRel;
    cudaEvent_t evReady = nullptr, evPulled = nullptr;  // hand-over to / from the mapper's pull stream
    std::vector<Timer> timers;
    dp_stats stats{};
    int candStride = 0;
    int extractWarps = 0, lookupWarps = 0, chainWarps = 0;
    ~Lane() {
        rounds.reset();
        if (evRounds) cudaEventDestroy(evRounds);
        if (evReady) cudaEventDestroy(evReady);
        if (evPulled) cudaEventDestroy(evPulled);
        if (evSync) cudaEventDestroy(evSync);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct dp_mapper {
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;  // index construction
    // Reads that live in pinned host memory are pulled over PCIe by ONE stream shared by all lanes: pulls run back to
    // back at link speed while the lanes' compute kernels and host work overlap them (lanes pulling independently fall
    // into lock step: both wait on the link, then both wait on the SMs).
    cudaStream_t pullStream = nullptr;
    std::mutex pullMu;
    cudaEvent_t evTrace = nullptr;  // DP_TRACE: time zero of the call's device timeline (recorded on the pull stream)
    double traceHost0 = 0;
    // parameters
    int k = 0, circular = 0, seedRate = 0, edge = 0, chunkSize = 0;
    long long refLen = 0;
    // index
    DBuf<unsigned> refWords;
    DBuf<unsigned char> image;  // set when the mapper was opened from an index image: I points into it
    DBuf<uint2> table;
    DBuf<unsigned> seedOff, seedChunks, chunkOff, chunkSeed, postOff, postChunk, filter;
    DBuf<int> chunkPos, chunkScanLen, postPos;
    int filterBits = 0;
    DBuf<long long> chunkOffset, chunkInset;
    DBuf<unsigned> midPost;
    DBuf<uint4> midSeed;  // derived from seedOff / seedChunks when the mapper is opened (not part of the image)
    std::vector<long long> hChunkOffset, hChunkInset;
    std::vector<int> hChunkLen, hChunkScanLen;
    DpIndexDev I{};
    long long nChunkPostings = 0, nSeedPostings = 0;
    size_t indexBytes = 0;
    // Lanes are handed out per call (LaneSet): concurrent callers of one mapper — the reference runs num_workers goroutines
    // against one Mapper, commands/map.go:84-86 — each work on their own lanes.
    std::vector<std::unique_ptr<Lane>> lanes;
    std::vector<char> laneBusy;
    std::mutex laneMu;
    std::condition_variable laneCv;
    dp_stats stats{};  // of the call that finished last (guarded by laneMu)
    // blocks the sub-batches' records are assembled in, kept between sub-batches and calls (no fresh pages per call)
    std::vector<std::vector<dp_mapping>> spare;
    std::mutex spareMu;

    ~dp_mapper() {
        lanes.clear();
        if (evTrace) cudaEventDestroy(evTrace);
        if (pullStream) cudaStreamDestroy(pullStream);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

// ----------------------------------------------------------------------------------------------------------------
// index construction (mapping.NewMapper, mapping/mapping.go:67-109)
// ----------------------------------------------------------------------------------------------------------------
void build_mid_postings(dp_mapper& M);
void build_postings(dp_mapper& M, DBuf<unsigned long long>& keys, unsigned long long P2, unsigned numSeeds, unsigned C,
                    unsigned maxChunkSeeds);

void build_index(dp_mapper& M, const uint8_t* ref, const double* values) {
    const int k = M.k;
    const long long L = M.refLen;
    const int e = M.edge;
    cudaStream_t st = M.stream;
    const long long nTable = (1ll << (2 * k)) / 32;

    // ---- pack the reference (+ the circular join sequence, mapping.go:93-95) ----
    long long refWordsN = (L + 15) / 16;
    long long joinWord = refWordsN + 4;  // zero gap: scans that over-read the reference end see zeros, like the oracle
    long long joinLen = M.circular ? 2ll * e : 0;
    long long totalWords = joinWord + (joinLen + 15) / 16 + 4;
    M.refWords.reserve((size_t)totalWords);
    CK(cudaMemsetAsync(M.refWords.p, 0, M.refWords.cap * sizeof(unsigned), st));
    {
        DBuf<unsigned char> dA;
        dA.reserve((size_t)(L + joinLen + 64));
        CK(cudaMemcpyAsync(dA.p, ref, (size_t)L, cudaMemcpyHostToDevice, st));
        if (M.circular) {
            CK(cudaMemcpyAsync(dA.p + L, ref + (L - e), (size_t)e, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(dA.p + L + e, ref, (size_t)e, cudaMemcpyHostToDevice, st));
        }
        DBuf<long long> dOff, dWord;
        // one warp per sequence would serialise a whole genome on one warp: split the reference into slices
        // (the kernel takes (offset, word) pairs, so slices of 16-base multiples are independent "sequences")
        const long long slice = 1 << 14;  // bases per slice, multiple of 16
        long long nSl = (L + slice - 1) / slice;
        std::vector<long long> so((size_t)nSl + 2), sw((size_t)nSl + 1);
        for (long long i = 0; i < nSl; i++) {
            so[(size_t)i] = i * slice;
            sw[(size_t)i] = i * slice / 16;
        }
        so[(size_t)nSl] = L;
        long long nSeq = nSl;
        if (M.circular) {
            sw[(size_t)nSl] = joinWord;
            so[(size_t)nSl + 1] = L + joinLen;
            nSeq = nSl + 1;
        }
        dOff.reserve(so.size());
        dWord.reserve(sw.size());
        CK(cudaMemcpyAsync(dOff.p, so.data(), so.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dWord.p, sw.data(), sw.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
        int blocks = std::min<long long>((nSeq + 7) / 8, (long long)M.smCount * 8);
        dp_pack_kernel<<<blocks, 256, 0, st>>>(dA.p, dOff.p, dWord.p, M.refWords.p, nSeq);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
    }

    // ---- AddSingleSeeds (seeds/seeds.go:160-200) ----
    DBuf<unsigned> bits;
    bits.reserve((size_t)nTable);
    CK(cudaMemsetAsync(bits.p, 0, (size_t)nTable * sizeof(unsigned), st));
    {
        DpSeedSelParams P;
        P.refLen = L;
        P.rate = M.seedRate;
        P.k = k;
        P.nWindows = (L - M.seedRate > 0) ? (L - M.seedRate + M.seedRate - 1) / M.seedRate : 0;
        int finalLen = (int)(L % 4);
        P.skipBack = 4 - finalLen;
        if (P.nWindows > 0) {
            if (P.nWindows >= 0xffffffffll) throw std::runtime_error("reference too long for 32-bit window ids");
            DBuf<double> dVal;
            dVal.reserve((size_t)(1ull << (2 * k)));
            CK(cudaMemcpyAsync(dVal.p, values, sizeof(double) << (2 * k), cudaMemcpyHostToDevice, st));
            DBuf<unsigned> best, f;
            DBuf<unsigned char> needA, needB;
            DBuf<unsigned> changed;
            best.reserve((size_t)P.nWindows);
            f.reserve((size_t)(1ull << (2 * k)));
            needA.reserve((size_t)P.nWindows);
            needB.reserve((size_t)P.nWindows);
            changed.reserve(1);
            int blocks = div_up(P.nWindows, 256);
            dp_seed_best_kernel<<<blocks, 256, 0, st>>>(M.refWords.p, dVal.p, P, best.p);
            CK(cudaGetLastError());
            CK(cudaMemsetAsync(needA.p, 1, (size_t)P.nWindows, st));
            unsigned char* cur = needA.p;
            unsigned char* nxt = needB.p;
            int iters = 0;
            for (;;) {
                CK(cudaMemsetAsync(f.p, 0xff, sizeof(unsigned) << (2 * k), st));
                CK(cudaMemsetAsync(changed.p, 0, sizeof(unsigned), st));
                dp_seed_fmin_kernel<<<blocks, 256, 0, st>>>(best.p, cur, P.nWindows, f.p);
                dp_seed_need_kernel<<<blocks, 256, 0, st>>>(M.refWords.p, f.p, P, cur, nxt, changed.p);
                CK(cudaGetLastError());
                unsigned hc = 0;
                CK(cudaMemcpyAsync(&hc, changed.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                std::swap(cur, nxt);
                iters++;
                if (!hc) break;
                if (iters > 1000000) throw std::runtime_error("seed selection did not converge");
            }
            dp_seed_setbits_kernel<<<blocks, 256, 0, st>>>(best.p, cur, P.nWindows, bits.p);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));
        }
    }
    // ---- {flags, rank} table ----
    M.table.reserve((size_t)nTable);
    unsigned numSeeds = 0;
    {
        DBuf<unsigned> pc, prefix;
        pc.reserve((size_t)nTable + 1);
        prefix.reserve((size_t)nTable + 1);
        dp_popc_kernel<<<div_up(nTable, 256), 256, 0, st>>>(bits.p, pc.p, nTable);
        DBuf<unsigned char> tmp;
        CK(cudaMemsetAsync(pc.p + nTable, 0, sizeof(unsigned), st));
        dp_exclusive_sum(pc.p, prefix.p, (long long)nTable + 1, tmp, st);
        dp_table_kernel<<<div_up(nTable, 256), 256, 0, st>>>(bits.p, prefix.p, M.table.p, nTable);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&numSeeds, prefix.p + nTable, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }

    // ---- chunk list in the producer's emission order (mapping.go:80-95; canonical ids) ----
    std::vector<DpChunkDesc> descs;
    M.hChunkOffset.clear();
    M.hChunkInset.clear();
    M.hChunkLen.clear();
    M.hChunkScanLen.clear();
    for (int j = 0; j < 10; j++) {
        long long start = (long long)j * M.chunkSize;
        long long step = (long long)M.chunkSize * 10 - e;
        for (long long i = start; i < L - M.chunkSize / 2; i += step) {
            long long end = std::min<long long>(i + M.chunkSize, L);
            DpChunkDesc d;
            d.base = i;
            d.nVisit = (int)(end - i - k + 1);
            d.pad = 0;
            descs.push_back(d);
            M.hChunkOffset.push_back(i);
            M.hChunkInset.push_back(L - (end - 1));  // SubSequence: inset + length - (end-1)  (Q3)
            M.hChunkLen.push_back((int)(end - i));
            M.hChunkScanLen.push_back((int)(end - i));
        }
    }
    if (M.circular) {
        DpChunkDesc d;
        d.base = joinWord * 16;
        int jl = 2 * e;
        d.nVisit = jl - k + 1 - ((jl % 4 == 0) ? 4 : 0);  // raw packedSequence: Q2
        d.pad = 0;
        descs.push_back(d);
        M.hChunkOffset.push_back(L - e);      // Append keeps the left part's offset ...
        M.hChunkInset.push_back(L - (e - 1)); // ... and the right part's inset (sequence.go:173-174)
        M.hChunkLen.push_back(jl);
        M.hChunkScanLen.push_back(d.nVisit + k - 1);
    }
    const unsigned C = (unsigned)descs.size();
    if (C == 0) throw std::runtime_error("reference shorter than chunk_size/2: no chunks to index");
    DBuf<DpChunkDesc> dDescs;
    dDescs.reserve(C);
    CK(cudaMemcpyAsync(dDescs.p, descs.data(), C * sizeof(DpChunkDesc), cudaMemcpyHostToDevice, st));
    M.chunkOffset.reserve(C);
    M.chunkInset.reserve(C);
    M.chunkScanLen.reserve(C);
    CK(cudaMemcpyAsync(M.chunkOffset.p, M.hChunkOffset.data(), C * sizeof(long long), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(M.chunkInset.p, M.hChunkInset.data(), C * sizeof(long long), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(M.chunkScanLen.p, M.hChunkScanLen.data(), C * sizeof(int), cudaMemcpyHostToDevice, st));

    // ---- chunk scan: count, exclusive scan, write ----
    DBuf<unsigned> counts;
    counts.reserve(C + 1);
    M.chunkOff.reserve(C + 1);
    CK(cudaMemsetAsync(counts.p, 0, (C + 1) * sizeof(unsigned), st));
    int scanBlocks = std::min<int>(div_up(C, 8), M.smCount * 8);
    dp_chunk_scan_kernel<<<scanBlocks, 256, 0, st>>>(M.refWords.p, M.table.p, dDescs.p, C, k, 0, counts.p, nullptr,
                                                     nullptr, nullptr, nullptr);
    CK(cudaGetLastError());
    {
        DBuf<unsigned char> tmp;
        dp_exclusive_sum(counts.p, M.chunkOff.p, (long long)C + 1, tmp, st);
        CK(cudaStreamSynchronize(st));
    }
    std::vector<unsigned> hCounts(C + 1);
    CK(cudaMemcpy(hCounts.data(), counts.p, (C + 1) * sizeof(unsigned), cudaMemcpyDeviceToHost));
    unsigned long long P2 = 0;
    unsigned maxChunkSeeds = 0;
    for (unsigned c = 0; c < C; c++) {
        P2 += hCounts[c];
        maxChunkSeeds = std::max(maxChunkSeeds, hCounts[c]);
    }
    if (P2 >= 0xffffffffull) throw std::runtime_error("index too large for 32-bit posting offsets");
    M.chunkPos.reserve((size_t)P2 + 1);
    M.chunkSeed.reserve((size_t)P2 + 1);
    DBuf<unsigned long long> keys;
    keys.reserve((size_t)P2 + 1);
    dp_chunk_scan_kernel<<<scanBlocks, 256, 0, st>>>(M.refWords.p, M.table.p, dDescs.p, C, k, 1, nullptr, M.chunkOff.p,
                                                     M.chunkPos.p, M.chunkSeed.p, keys.p);
    CK(cudaGetLastError());

    build_postings(M, keys, P2, numSeeds, C, maxChunkSeeds);
    M.refWords.release();  // only index construction reads the packed reference
}

// Second half of index construction, shared by `map` (chunks of the reference) and `overlap` (seed-space chunks of the
// reads): from the chunk -> seeds lists (M.chunkOff / chunkPos / chunkSeed) and their (seed, chunk) keys to the two
// seed -> ... CSRs, the extract kernel's prefix filter, the mid-lookup copy, and M.I.
void build_postings(dp_mapper& M, DBuf<unsigned long long>& keys, unsigned long long P2, unsigned numSeeds, unsigned C,
                    unsigned maxChunkSeeds) {
    const int k = M.k;
    const int e = M.edge;
    const long long L = M.refLen;
    cudaStream_t st = M.stream;
    const long long nTable = (1ll << (2 * k)) / 32;
    DBuf<unsigned long long> keysSorted;
    keysSorted.reserve((size_t)P2 + 1);
    // ---- seed -> distinct chunks: radix sort of (seed, chunk) keys, unique, CSR ----
    M.seedOff.reserve((size_t)numSeeds + 2);
    unsigned long long P1 = 0;
    if (P2 > 0) {
        int endBit = 32;
        while ((1ull << (endBit - 32)) < (unsigned long long)numSeeds + 1 && endBit < 64) endBit++;
        // stable LSD radix sort: equal (seed, chunk) keys keep their input order, which is ascending scan position
        M.postPos.reserve((size_t)P2 + 1);
        M.postChunk.reserve((size_t)P2 + 1);
        DBuf<unsigned char> tmp;
        DBuf<unsigned> hist;
        {
            DBuf<unsigned long long> keysScratch;
            DBuf<unsigned> valsScratch;
            if (endBit > 8) {
                keysScratch.reserve((size_t)P2 + 1);
                valsScratch.reserve((size_t)P2 + 1);
            }
            dp_radix_sort(keys.p, keysSorted.p, keysScratch.p, reinterpret_cast<const unsigned*>(M.chunkPos.p),
                          reinterpret_cast<unsigned*>(M.postPos.p), valsScratch.p, (long long)P2, 0, endBit, hist, tmp, st);
            CK(cudaStreamSynchronize(st));  // (the scratch arrays go out of scope)
        }
        {
            DBuf<unsigned> seedCountAll;
            seedCountAll.reserve((size_t)numSeeds + 2);
            M.postOff.reserve((size_t)numSeeds + 2);
            CK(cudaMemsetAsync(seedCountAll.p, 0, ((size_t)numSeeds + 2) * sizeof(unsigned), st));
            dp_posting_all_kernel<<<div_up((long long)P2, 256), 256, 0, st>>>(keysSorted.p, (long long)P2,
                                                                             seedCountAll.p, M.postChunk.p);
            CK(cudaGetLastError());
            dp_exclusive_sum(seedCountAll.p, M.postOff.p, (long long)numSeeds + 1, tmp, st);
            CK(cudaStreamSynchronize(st));
        }
        DBuf<unsigned long long> nSel;
        nSel.reserve(1);
        {
            DBuf<unsigned> upos;
            dp_unique_sorted(keysSorted.p, keys.p, nSel.p, (long long)P2, upos, tmp, st);
            CK(cudaStreamSynchronize(st));
        }
        CK(cudaMemcpyAsync(&P1, nSel.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    M.seedChunks.reserve((size_t)P1 + 4);  // (the block lookup reads aligned 16-byte pieces)
    {
        DBuf<unsigned> seedCount;
        seedCount.reserve((size_t)numSeeds + 2);
        CK(cudaMemsetAsync(seedCount.p, 0, ((size_t)numSeeds + 2) * sizeof(unsigned), st));
        if (P1 > 0) {
            dp_posting_fill_kernel<<<div_up((long long)P1, 256), 256, 0, st>>>(keys.p, (long long)P1, seedCount.p,
                                                                              M.seedChunks.p);
            CK(cudaGetLastError());
        }
        DBuf<unsigned char> tmp;
        dp_exclusive_sum(seedCount.p, M.seedOff.p, (long long)numSeeds + 1, tmp, st);
        CK(cudaStreamSynchronize(st));
    }

    if (P2 == 0) {
        M.postOff.reserve((size_t)numSeeds + 2);
        M.postChunk.reserve(1);
        M.postPos.reserve(1);
        CK(cudaMemsetAsync(M.postOff.p, 0, ((size_t)numSeeds + 2) * sizeof(unsigned), st));
    }
    // ---- shared-memory prefilter for the extract kernel: the seed table OR-folded to 2^bits k-mer prefixes (exact when
    //      2k <= 20); worthwhile while at most half of its bits are set ----
    M.filterBits = 0;
    {
        const int bits = std::min(20, 2 * k);  // 128 KiB at most: one CTA per SM next to the hit masks
        if (2 * k <= 20 || (unsigned long long)numSeeds * 2 <= (1ull << bits)) {
            size_t fw = (((size_t)1 << bits) + 31) / 32;
            M.filter.reserve(fw);
            CK(cudaMemsetAsync(M.filter.p, 0, fw * sizeof(unsigned), st));
            dp_filter_build_kernel<<<div_up(nTable, 256), 256, 0, st>>>(M.table.p, nTable, k, bits, M.filter.p);
            CK(cudaGetLastError());
            M.filterBits = bits;
        }
    }
    CK(cudaStreamSynchronize(st));

    DpIndexDev& I = M.I;
    I.k = k;
    I.postOff = M.postOff.p;
    I.postChunk = M.postChunk.p;
    I.postPos = M.postPos.p;
    I.filter = M.filter.p;
    I.filterBits = M.filterBits;
    I.circular = M.circular;
    I.edge = e;
    I.maxWindow = 2 * e;
    I.refLen = L;
    I.numSeeds = numSeeds;
    I.numChunks = C;
    I.maxChunkSeeds = maxChunkSeeds;
    I.table = M.table.p;
    I.seedOff = M.seedOff.p;
    I.seedChunks = M.seedChunks.p;
    I.chunkOff = M.chunkOff.p;
    I.chunkPos = M.chunkPos.p;
    I.chunkSeed = M.chunkSeed.p;
    I.chunkOffset = M.chunkOffset.p;
    I.chunkInset = M.chunkInset.p;
    I.chunkScanLen = M.chunkScanLen.p;
    M.nChunkPostings = (long long)P2;
    M.nSeedPostings = (long long)P1;
    M.indexBytes = M.table.bytes() + M.seedOff.bytes() + M.seedChunks.bytes() +
                   M.chunkOff.bytes() + M.chunkPos.bytes() + M.chunkSeed.bytes() + M.chunkOffset.bytes() +
                   M.postOff.bytes() + M.postChunk.bytes() + M.postPos.bytes() + M.filter.bytes() +
                   M.chunkInset.bytes() + M.chunkScanLen.bytes();
    build_mid_postings(M);
}

// dp_lookup_mid_kernel: a byte counter per chunk for each warp, at least three CTAs of four warps per SM
const unsigned kLookupMidMaxChunks = 16384;

// The mid-lookup copy of the seed -> chunks runs (DpIndexDev::midOff / midPost): derived data, rebuilt whenever a mapper
// is opened (from a reference or from an index image), for indexes dp_lookup_mid_kernel can take.
void build_mid_postings(dp_mapper& M) {
    DpIndexDev& I = M.I;
    I.midSeed = nullptr;
    I.midPost = nullptr;
    if (I.numChunks > kLookupMidMaxChunks || I.numSeeds == 0) return;
    cudaStream_t st = M.stream;
    const unsigned S = I.numSeeds;
    DBuf<unsigned> blocks;
    blocks.reserve((size_t)S + 1);
    DBuf<unsigned> midOff;
    midOff.reserve((size_t)S + 1);
    M.midSeed.reserve((size_t)S);
    dp_mid_blocks_kernel<<<div_up((long long)S + 1, 256), 256, 0, st>>>(I.seedOff, S, blocks.p);
    CK(cudaGetLastError());
    DBuf<unsigned char> tmp;
    dp_exclusive_sum(blocks.p, midOff.p, (long long)S + 1, tmp, st);
    unsigned total = 0;
    CK(cudaMemcpyAsync(&total, midOff.p + S, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (total >= (1u << 28)) return;  // (item descriptors carry a 28-bit block index; far beyond any index of <= 16384 chunks)
    M.midPost.reserve((size_t)total * 4 + 4);
    const unsigned padWord = (dp_mid_words(I.numChunks) - 32u) * 4u;
    dp_mid_fill_kernel<<<M.smCount * 8, 256, 0, st>>>(I.seedOff, I.seedChunks, midOff.p, S, padWord, M.midPost.p, M.midSeed.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    I.midSeed = M.midSeed.p;
    I.midPost = reinterpret_cast<const uint4*>(M.midPost.p);
    M.indexBytes += M.midSeed.bytes() + M.midPost.bytes();
}

// ----------------------------------------------------------------------------------------------------------------
// window rounds
// ----------------------------------------------------------------------------------------------------------------
const unsigned kLookupSmemChunksDefault = 24000;  // 16-bit counters: 4 warps x 48 KB of shared memory at most
unsigned lookup_smem_chunks() {  // DP_LOOKUP_SMEM_CHUNKS: tests force the warp kernel's global-memory counters
    const char* env = getenv("DP_LOOKUP_SMEM_CHUNKS");
    return env ? (unsigned)atoi(env) : kLookupSmemChunksDefault;
}
#define kLookupSmemChunks lookup_smem_chunks()
const unsigned kLookupBlockMinChunks = 2048;  // from here on a CTA per window strand (dp_lookup_block_kernel)

// per-warp / per-CTA stride of the lookup scratch lists: a window strand touches at most C chunks; the CTA kernel also
// keeps the item prefix of an oversized window strand (<= maxWindow + 1 entries) there
size_t lookup_tstride(const DpIndexDev& I) { return std::max<size_t>(I.numChunks, (size_t)I.maxWindow + 1) + 8; }

struct LookupBlockPlan {
    bool use = false;
    int gShift = 0, cntWords = 0, eCap = 512, gListCap = DP_BGLIST, gBatch = DP_BEXACT, ctasPerSm = 1, seg = 128, threads = 256, exactWords = DP_BEXACT;
    size_t smem = 0;
};

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

bool use_mid_lookup(const DpIndexDev& I) {
    if (getenv("DP_LOOKUP_BLOCK") && atoi(getenv("DP_LOOKUP_BLOCK")) != 0) return false;
    if (!I.midPost) return false;
    const int mid = env_int("DP_LOOKUP_MID", -1);  // tests: 1 forces it on small references, 0 switches it off
    if (mid >= 0) return mid != 0;
    return I.numChunks > DP_SMALL_CHUNKS;
}

LookupBlockPlan plan_block_lookup(const DpIndexDev& I, double avgRun) {
    LookupBlockPlan P;
    const char* env = getenv("DP_LOOKUP_BLOCK");
    bool want = I.numChunks >= kLookupBlockMinChunks && !use_mid_lookup(I);
    if (env) want = atoi(env) != 0;
    if (!want) return P;
    // what a window strand is expected to look like: seeds per strand, postings per strand
    const double estSeeds = (double)(I.edge - I.k + 1) * (double)I.numSeeds / (double)(1ull << (2 * I.k));
    const double estPostings = estSeeds * avgRun;
    // runs of a few dozen postings (k = 13, or a few thousand chunks) would leave 128-posting items mostly empty
    P.seg = avgRun < 64.0 ? 32 : 128;
    if (getenv("DP_LOOKUP_SEG")) P.seg = env_int("DP_LOOKUP_SEG", 128) == 32 ? 32 : 128;  // tests
    // A few thousand postings per window strand (BASELINE config 3) are streamed in a few microseconds; what bounds
    // the kernel then is how many window strands an SM works on at once: 128-thread CTAs, eight per SM, with the
    // shared-memory footprint cut to fit. Human-scale indexes (tens of thousands of postings per window strand) do
    // better with 256 threads and four CTAs per SM (measured).
    const bool small = P.seg == 32 && estPostings < 8192.0;
    P.threads = small ? 128 : 256;
    if (getenv("DP_LOOKUP_THREADS")) P.threads = P.seg == 32 && env_int("DP_LOOKUP_THREADS", 256) == 128 ? 128 : 256;  // tests
    P.eCap = estSeeds * 2.0 <= 256.0 ? 256 : 512;
    P.eCap = std::max(64, std::min(2048, env_int("DP_LOOKUP_ECAP", P.eCap)));  // measurements
    P.exactWords = P.threads == 128 ? DP_BEXACT / 2 : DP_BEXACT;
    const size_t maxSmem = 220 * 1024;  // of 227 KB per CTA; the kernel has ~4 KB of static shared memory
    auto counter_words = [&](int gs) { return (size_t)((((size_t)I.numChunks - 1) >> gs) + 1 + 3) / 4 * 4; };
    auto smem_bytes = [&](int gs) {
        // counters | dummy counters | five uint32 arrays (+2 sentinels) | item list | exact recount table | one byte array
        return (counter_words(gs) * 4 + 128 + (size_t)P.eCap * 20 + 8 + 2 * DP_BITEMS * 4 + (gs ? P.exactWords * 4 : 0) + P.eCap + 15) / 16 * 16;
    };
    // a 32-bit counter covers 2^gShift adjacent chunks. The smallest gShift whose counters take <= 20 KB (8 KB for the
    // small CTAs) keeps the CTAs resident and the groups selective: a random group collects
    // ~4 * (chunks per seed / C) * 2^gShift of the threshold — ~1/3 at 32 chunks per group on the synthetic 3.1 Gb
    // genome, past 64 the random groups start to cross it (measured: 5.3 ms -> 26 ms -> 580 ms at gShift 5, 6, 7 on
    // 1 Gb); more than 32 chunks per group only when shared memory leaves no choice
    const int maxShift = 10;
    const size_t target = (size_t)(P.threads == 128 ? 8 : 20) * 1024;
    int gs = 0;
    while (gs < 5 && counter_words(gs) * 4 > target) gs++;
    while (gs < maxShift && smem_bytes(gs) > maxSmem) gs++;
    if (getenv("DP_LOOKUP_GSHIFT")) gs = std::min(maxShift, std::max(0, env_int("DP_LOOKUP_GSHIFT", 0)));  // tests
    if (smem_bytes(gs) > maxSmem) throw std::runtime_error("reference has too many chunks for the lookup kernel's shared-memory counters");
    P.gShift = gs;
    P.cntWords = (int)counter_words(gs);
    P.smem = smem_bytes(gs);
    P.gListCap = std::min(DP_BGLIST, std::max(1, env_int("DP_LOOKUP_GLIST", DP_BGLIST)));                        // tests
    P.gBatch = std::min(P.exactWords >> gs, std::max(1, env_int("DP_LOOKUP_GBATCH", P.exactWords >> gs)));     // tests
    if (P.gBatch < 1) P.gBatch = 1;
    size_t fit = (227 * 1024) / (P.smem + 5 * 1024);  // ~4 KB static + 1 KB the driver reserves per CTA
    P.ctasPerSm = (int)std::max<size_t>(1, std::min<size_t>(fit, P.threads == 128 ? 8 : 4));  // 64 registers per thread
    P.ctasPerSm = std::max(1, std::min(P.ctasPerSm, env_int("DP_LOOKUP_CTAS", 8)));           // measurements
    P.use = true;
    return P;
}

void ensure_window_capacity(dp_mapper& M, Lane& W, size_t nWinNow, size_t seedEntriesNow) {
    const DpIndexDev& I = M.I;
    const size_t nWin = std::max(nWinNow, W.floorWins), seedEntries = std::max(seedEntriesNow, W.floorSeeds);  // what is reserved
    W.dWins.reserve(nWin);
    W.wsOff.reserve(2 * nWin);
    W.wsN.reserve(2 * nWin);
    W.qSeed.reserve(seedEntries + 64);
    W.qPos.reserve(seedEntries + 64);
    W.cursor.reserve(4);
    W.dCtr.reserve(1);
    W.candStride = (int)std::min<unsigned>(I.numChunks, W.caps.candStride);
    W.candN.reserve(2 * nWin);
    W.candChunk.reserve(2 * nWin * (size_t)W.candStride);
    W.candDistinct.reserve(2 * nWin * (size_t)W.candStride);
    W.outN.reserve(nWin);
    W.outOff.reserve(nWin);
    W.outMaps.reserve(nWin * (size_t)W.caps.outStride);
    W.hCtr.reserve(1);
    // per-warp scratch
    const int qStride = I.maxWindow + 8;
    W.extractWarps = M.smCount * 8 * 8;
    // lookup scratch is indexed by warp in dp_lookup_kernel and by CTA in dp_lookup_block_kernel (<= 8 per SM)
    const LookupBlockPlan LP = plan_block_lookup(I, (double)M.nSeedPostings / std::max(1u, I.numSeeds));
    W.lookupWarps = LP.use ? M.smCount * 4 * 2 : M.smCount * 4 * 8;
    W.lbDefer.reserve(2 * nWin);
    W.lbWork.reserve(4);
    {   // the general chain kernel's per-warp scratch grows with the capacities: fewer warps then (2 GB of scratch)
        const size_t perWarp = (size_t)W.caps.resultCap * sizeof(DpMappingDev) + (size_t)W.caps.chainCap * 6 * sizeof(int);
        const size_t fit = ((size_t)2 << 30) / perWarp;
        W.chainWarps = (int)std::max<size_t>(128, std::min<size_t>((size_t)M.smCount * 4 * 8, fit / 4 * 4));
    }
    size_t lw = (size_t)W.lookupWarps;
    W.lsSeed.reserve(lw * qStride);
    W.lsOff.reserve(lw * qStride);
    W.lsPre.reserve(lw * (qStride + 1));
    W.lsEndW.reserve(lw * qStride);
    W.lsAll.reserve(lw * qStride);
    W.lsFirst.reserve(lw * qStride);
    W.lsOrder.reserve(lw * qStride);
    // a window strand touches at most min(C, total postings) chunks
    const size_t tStride = lookup_tstride(I);
    W.lsTouched.reserve(lw * tStride);
    W.lsCand.reserve(lw * 2 * tStride);
    if (I.numChunks > kLookupSmemChunks && !W.lsCountersZeroed) {
        W.lsCounters.reserve(lw * ((I.numChunks + 1) / 2));
        // the kernel keeps the invariant "all counters zero between window strands": clear once
        CK(cudaMemsetAsync(W.lsCounters.p, 0, W.lsCounters.cap * sizeof(unsigned), W.stream));
        W.lsCountersZeroed = true;
    }
    size_t cw = (size_t)W.chainWarps;
    const int sStride = (int)I.maxChunkSeeds + 8;
    W.csQFirst.reserve(cw * qStride);
    W.csQCnt.reserve(cw * qStride);
    W.csQLo.reserve(cw * qStride);
    W.csRqPos.reserve(cw * qStride);
    W.csRqId.reserve(cw * qStride);
    W.csChainLen.reserve(cw * qStride);
    W.csLastB.reserve(cw * qStride);
    W.csEnt.reserve(cw * sStride);
    W.csRsPos.reserve(cw * sStride);
    W.csRsId.reserve(cw * sStride);
    W.csChains.reserve(cw * W.caps.chainCap * 6);
    W.csResults.reserve(cw * W.caps.resultCap);
    // fast chain path: ~1.5 candidates and ~40 list entries per window on ONT-like reads; whatever does not fit is
    // handed back to the general kernel, so these sizes are a speed knob, not a limit
    W.fcTaskCap = nWin * 8 + 1024;
    W.fcPoolCap = getenv("DP_FAST_POOL_WORDS") ? (size_t)atoll(getenv("DP_FAST_POOL_WORDS")) : nWin * 256 + 4096;
    if (W.fcPoolCap > 0xfff00000ull) W.fcPoolCap = 0xfff00000ull;
    W.fcTasks.reserve(W.fcTaskCap);
    W.fcTaskBase.reserve(2 * nWin);
    W.fcPool.reserve(W.fcPoolCap);
    W.fcCursors.reserve(4);
    W.fcSlow.reserve(nWin);
    W.fcSlowList.reserve(nWin);
}

enum { T_PACK = 0, T_EXTRACT, T_LOOKUP, T_CHAIN, T_FINISH, T_REDUCE, T_PULL, T_N };
enum { CUR_SEEDS = 0, CUR_OUT = 1, CUR_FIN = 2 };

// Launches the three performMapping stages for the `nWin` windows already in W.dWins (device). Results stay on the
// device: W.outN / W.outOff / W.outMaps (compact, bump-allocated through cursor[CUR_OUT]).
void launch_windows(dp_mapper& M, Lane& W, size_t nWin, size_t seedEntries, const unsigned* dWords, const long long* dWordOff,
                    const int* dReadLen, bool bulkPull = false) {
    const DpIndexDev& I = M.I;
    cudaStream_t st = W.stream;
    if (seedEntries >= 0xffffffffull) throw std::runtime_error("window round too large for 32-bit seed offsets");
    CK(cudaMemsetAsync(W.cursor.p, 0, 4 * sizeof(unsigned long long), st));
    // compute kernels leave part of every SM free while another lane's pull kernel reads host memory (measured with
    // the zero-copy pack kernel; with the TMA pull it makes no measurable difference; DP_HEADROOM=0/1 overrides)
    const bool headroom = W.curAsciiIsHost && env_int("DP_HEADROOM", 1) != 0;
    {   // pack exactly the queried windows (out of pinned host memory when that is where the reads live)
        // PCIe-bound when the reads are pulled from host memory: keep its footprint at two CTAs per SM so the lanes'
        // compute kernels stay resident beside it
        int perSm = W.curAsciiIsHost ? 2 : 6;
        const long long* spanOff = W.curSpans ? W.dByteOff.p : nullptr;  // ASCII reads addressed by their own offsets
        int blocks = (int)std::min<size_t>((nWin + 7) / 8, (size_t)M.smCount * perSm);
        const int stageStride = ((I.maxWindow + 15) / 16 + 3) * 16;  // the 16-byte blocks of the longest window
        const bool tmaPull = env_int("DP_PULL_TMA", 1) != 0 && (size_t)DP_PULL_SLOTS_ASCII * 2 * stageStride <= 96 * 1024;
        if (W.curPacked && W.curAsciiIsHost && bulkPull && env_int("DP_PULL_TMA", 1) != 0 &&
            (size_t)DP_PULL_SLOTS_PACKED * 2 * (((I.maxWindow / 4 + 15) / 16 + 3) * 16) <= 96 * 1024) {
            // reads that arrive packed, in pinned host memory: the same TMA pull (a quarter of the bytes), then the
            // realign + byte swap from the HBM staging buffer
            const int pkStride = ((I.maxWindow / 4 + 15) / 16 + 3) * 16;
            W.dStage.reserve((std::max(nWin, W.floorWins) + 1) * (size_t)pkStride);
            W.dStagePos.reserve(std::max(nWin, W.floorWins));
            W.dPullWork.reserve(1);
            {
                std::lock_guard<std::mutex> lk(M.pullMu);  // (wait, kernel, record) must enter the pull stream as one unit
                CK(cudaEventRecord(W.evReady, st));
                CK(cudaStreamWaitEvent(M.pullStream, W.evReady, 0));
                CK(cudaEventRecord(W.timers[T_PACK].a, M.pullStream));
                // (pieces of 250-500 bytes: the link is bound by the read requests in flight, 64 single-warp CTAs measured
                // best — 24.3 ms per 1M reads against 35.5 with 32, 25.5 with 128 and 29 with zero-copy loads)
                const int pullCtas = (int)std::min<size_t>((nWin + 31) / 32, (size_t)std::max(1, env_int("DP_PULL_CTAS", 64)));
                CK(cudaMemsetAsync(W.dPullWork.p, 0, sizeof(unsigned), M.pullStream));
                dp_pull_windows_kernel<DP_PULL_SLOTS_PACKED, 12><<<pullCtas, 32, (size_t)DP_PULL_SLOTS_PACKED * 2 * pkStride, M.pullStream>>>(
                    W.curAscii, W.dSeqOff.p, W.dByteOff.p, true, W.dWins.p, (int)nWin, W.dStage.p, pkStride, W.dStagePos.p,
                    W.dPullWork.p);
                CK(cudaGetLastError());
                CK(cudaEventRecord(W.timers[T_PULL].b, M.pullStream));
                CK(cudaEventRecord(W.evPulled, M.pullStream));
            }
            CK(cudaStreamWaitEvent(st, W.evPulled, 0));
            int fullBlocks = (int)std::min<size_t>((nWin + 7) / 8, (size_t)M.smCount * 6);
            dp_pack_windows_packed_kernel<<<fullBlocks, 256, 0, st>>>(W.curAscii, W.dByteOff.p, W.dSeqOff.p, dWordOff, W.dWins.p,
                                                                      (int)nWin, const_cast<unsigned*>(dWords), W.dStage.p,
                                                                      W.dStagePos.p);
            CK(cudaGetLastError());
            CK(cudaEventRecord(W.timers[T_PACK].b, st));
            W.stats.kernel_launches += 1;
        } else if (W.curPacked) {
            // reads that arrive packed: realign + byte swap of just the queried bytes (zero-copy out of mapped host
            // memory when that is where they live: the round-0 pulls of all lanes go through the shared pull stream)
            const bool viaPull = W.curAsciiIsHost && bulkPull;
            cudaStream_t ps = viaPull ? M.pullStream : st;
            std::unique_lock<std::mutex> lk(M.pullMu, std::defer_lock);
            if (viaPull) {
                lk.lock();  // (wait, kernel, record) must enter the pull stream as one unit
                CK(cudaEventRecord(W.evReady, st));
                CK(cudaStreamWaitEvent(M.pullStream, W.evReady, 0));
            }
            CK(cudaEventRecord(W.timers[T_PACK].a, ps));
            dp_pack_windows_packed_kernel<<<blocks, 256, 0, ps>>>(W.curAscii, W.dByteOff.p, W.dSeqOff.p, dWordOff, W.dWins.p,
                                                                  (int)nWin, const_cast<unsigned*>(dWords), nullptr, nullptr);
            CK(cudaGetLastError());
            CK(cudaEventRecord(W.timers[T_PACK].b, ps));
            if (viaPull) {
                CK(cudaEventRecord(W.evPulled, M.pullStream));
                CK(cudaStreamWaitEvent(st, W.evPulled, 0));
            }
        } else if (W.curAsciiIsHost && bulkPull && tmaPull) {
            // the PCIe leg as TMA bulk copies into an HBM staging buffer (a few single-warp CTAs on the shared pull
            // stream), then the pack from HBM at full width on the lane's own stream
            W.dStage.reserve((std::max(nWin, W.floorWins) + 1) * (size_t)stageStride);
            W.dStagePos.reserve(std::max(nWin, W.floorWins));
            W.dPullWork.reserve(1);
            {
                std::lock_guard<std::mutex> lk(M.pullMu);  // (wait, kernel, record) must enter the pull stream as one unit
                CK(cudaEventRecord(W.evReady, st));
                CK(cudaStreamWaitEvent(M.pullStream, W.evReady, 0));
                CK(cudaEventRecord(W.timers[T_PACK].a, M.pullStream));
                const int pullCtas = (int)std::min<size_t>((nWin + 31) / 32, (size_t)std::max(1, env_int("DP_PULL_CTAS", 32)));
                CK(cudaMemsetAsync(W.dPullWork.p, 0, sizeof(unsigned), M.pullStream));
                dp_pull_windows_kernel<DP_PULL_SLOTS_ASCII, 12><<<pullCtas, 32, (size_t)DP_PULL_SLOTS_ASCII * 2 * stageStride, M.pullStream>>>(
                    W.curAscii, W.dSeqOff.p, spanOff, false, W.dWins.p, (int)nWin, W.dStage.p, stageStride, W.dStagePos.p,
                    W.dPullWork.p);
                CK(cudaGetLastError());
                CK(cudaEventRecord(W.timers[T_PULL].b, M.pullStream));
                CK(cudaEventRecord(W.evPulled, M.pullStream));
            }
            CK(cudaStreamWaitEvent(st, W.evPulled, 0));
            int fullBlocks = (int)std::min<size_t>((nWin + 7) / 8, (size_t)M.smCount * 6);
            dp_pack_windows_kernel<<<fullBlocks, 256, 0, st>>>(W.curAscii, W.dSeqOff.p, spanOff, dWordOff, W.dWins.p, (int)nWin,
                                                               const_cast<unsigned*>(dWords), W.dStage.p, W.dStagePos.p);
            CK(cudaGetLastError());
            CK(cudaEventRecord(W.timers[T_PACK].b, st));
            W.stats.kernel_launches += 1;
        } else if (W.curAsciiIsHost && bulkPull) {
            std::lock_guard<std::mutex> lk(M.pullMu);  // (wait, kernel, record) must enter the pull stream as one unit
            CK(cudaEventRecord(W.evReady, st));
            CK(cudaStreamWaitEvent(M.pullStream, W.evReady, 0));
            CK(cudaEventRecord(W.timers[T_PACK].a, M.pullStream));
            dp_pack_windows_kernel<<<blocks, 256, 0, M.pullStream>>>(W.curAscii, W.dSeqOff.p, spanOff, dWordOff, W.dWins.p, (int)nWin,
                                                                    const_cast<unsigned*>(dWords), nullptr, nullptr);
            CK(cudaGetLastError());
            CK(cudaEventRecord(W.timers[T_PACK].b, M.pullStream));
            CK(cudaEventRecord(W.evPulled, M.pullStream));
            CK(cudaStreamWaitEvent(st, W.evPulled, 0));
        } else {
            CK(cudaEventRecord(W.timers[T_PACK].a, st));
            dp_pack_windows_kernel<<<blocks, 256, 0, st>>>(W.curAscii, W.dSeqOff.p, spanOff, dWordOff, W.dWins.p, (int)nWin,
                                                           const_cast<unsigned*>(dWords), nullptr, nullptr);
            CK(cudaGetLastError());
            CK(cudaEventRecord(W.timers[T_PACK].b, st));
        }
        W.stats.kernel_launches += 1;
    }
    const int qStride = I.maxWindow + 8;
    DpExtractOut Q;
    Q.wsOff = W.wsOff.p;
    Q.wsN = W.wsN.p;
    Q.qSeed = W.qSeed.p;
    Q.qPos = W.qPos.p;
    Q.cursor = W.cursor.p + CUR_SEEDS;
    const int maskWords = ((I.maxWindow + 1023) / 1024) * 32;  // 32 lanes x blocks of 1024 positions
    {
        // one persistent CTA per SM. When the other lane may be pulling reads over PCIe, every compute kernel leaves
        // a quarter of the SM's registers and thread slots free so that the pull kernel stays resident beside it.
        int warpsPerBlock = headroom ? 20 : 32;
        const size_t fWords = I.filterBits ? ((((size_t)1 << I.filterBits) + 31) / 32 + 3) / 4 * 4 : 0;
        const size_t perWarp = (size_t)2 * maskWords * sizeof(unsigned);
        const size_t room = 200 * 1024 - fWords * sizeof(unsigned);
        if (perWarp * warpsPerBlock > room) warpsPerBlock = (int)std::max<size_t>(1, room / perWarp);  // very long windows
        int blocks = (int)std::min<size_t>((nWin + warpsPerBlock - 1) / warpsPerBlock, (size_t)M.smCount);
        size_t smem = fWords * sizeof(unsigned) + perWarp * warpsPerBlock;
        if (smem > 200 * 1024) throw std::runtime_error("query_size too large for the extract kernel's shared memory");
        CK(cudaEventRecord(W.timers[T_EXTRACT].a, st));
        dp_extract_kernel<<<blocks, 32 * warpsPerBlock, smem, st>>>(I, dWords, dWordOff, W.dWins.p, (int)nWin, Q, maskWords,
                                                               W.dCtr.p);
        CK(cudaGetLastError());
        CK(cudaEventRecord(W.timers[T_EXTRACT].b, st));
    }
    {
        DpLookupScratch S;
        S.eSeed = W.lsSeed.p;
        S.eOff = W.lsOff.p;
        S.ePre = W.lsPre.p;
        S.eEndW = W.lsEndW.p;
        S.eFirst = W.lsFirst.p;
        S.allSeeds = W.lsAll.p;
        S.order = W.lsOrder.p;
        S.counters = W.lsCounters.p;
        S.touched = W.lsTouched.p;
        S.cand = W.lsCand.p;
        S.stride = qStride;
        S.tStride = (int)lookup_tstride(I);
        int inSmem = I.numChunks <= kLookupSmemChunks ? 1 : 0;
        int warpsPerBlock = DP_LWARPS;
        size_t smem = inSmem ? (size_t)warpsPerBlock * ((I.numChunks + 1) / 2) * sizeof(unsigned) : 0;
        const LookupBlockPlan LP = plan_block_lookup(I, (double)M.nSeedPostings / std::max(1u, I.numSeeds));
        CK(cudaEventRecord(W.timers[T_LOOKUP].a, st));
        if (LP.use) {
            DpLookupBlockCfg G;
            G.gShift = LP.gShift;
            G.cntWords = LP.cntWords;
            G.eCap = LP.eCap;
            G.gListCap = LP.gListCap;
            G.gBatch = LP.gBatch;
            G.exactWords = LP.exactWords;
            G.work = W.lbWork.p;
            G.deferList = W.lbDefer.p;
            G.nDefer = reinterpret_cast<int*>(W.lbWork.p + 1);
            CK(cudaMemsetAsync(W.lbWork.p, 0, 4 * sizeof(unsigned), st));
            int ctas = LP.ctasPerSm;
            if (headroom && ctas > 1) ctas -= ctas / 4 ? ctas / 4 : 0;  // headroom for the pull kernel
            int blocks = (int)std::min<size_t>(2 * nWin, (size_t)M.smCount * ctas);
            // 256 threads, four CTAs per SM (64 registers), six 16-byte loads per lane in flight: measured best of
            // {128, 256} threads x {2..6} CTAs x {4, 6, 8} items on a 1 Gb reference
            auto kern = LP.seg == 32 ? (LP.threads == 128 ? dp_lookup_block_kernel<128, 8, 6, 32> : dp_lookup_block_kernel<256, 4, 6, 32>)
                                     : dp_lookup_block_kernel<256, 4, 6, 128>;
            kern<<<blocks, LP.threads, LP.smem, st>>>(I, Q, (int)(2 * nWin), S, G, W.candN.p, W.candChunk.p, W.candDistinct.p,
                                              W.candStride, W.dCtr.p);
            CK(cudaGetLastError());
            // window strands the CTA kernel deferred (a seed present in every chunk): none on real references
            int dBlocks = (int)std::min<size_t>((2 * nWin + warpsPerBlock - 1) / warpsPerBlock, (size_t)M.smCount * 2);
            dp_lookup_kernel<<<dBlocks, 32 * DP_LWARPS, smem, st>>>(I, Q, (int)(2 * nWin), G.deferList, G.nDefer, S, inSmem,
                                                                   W.candN.p, W.candChunk.p, W.candDistinct.p,
                                                                   W.candStride, W.dCtr.p);
            CK(cudaGetLastError());
            W.stats.kernel_launches += 1;
        } else if (use_mid_lookup(I)) {
            // a few thousand chunks (BASELINE config 3): warp per window strand, many loads in flight, 16-bit counters
            // in shared memory; the general kernel takes what it hands back
            int* deferList = W.lbDefer.p;
            int* nDefer = reinterpret_cast<int*>(W.lbWork.p + 1);
            CK(cudaMemsetAsync(W.lbWork.p, 0, 4 * sizeof(unsigned), st));
            const size_t mSmem = (size_t)DP_MID_WARPS * dp_mid_words(I.numChunks) * sizeof(unsigned);
            int ctas = (int)std::max<size_t>(1, std::min<size_t>(5, (227 * 1024) / (mSmem + sizeof(DpMidWarp) * DP_MID_WARPS + 64 + 1024)));
            ctas = std::max(1, std::min(ctas, env_int("DP_LOOKUP_CTAS", 8)));  // measurements
            int mBlocks = (int)std::min<size_t>((2 * nWin + DP_MID_WARPS - 1) / DP_MID_WARPS, (size_t)M.smCount * ctas);
            dp_lookup_mid_kernel<6><<<mBlocks, 32 * DP_MID_WARPS, mSmem, st>>>(
                I, Q, (int)(2 * nWin), deferList, nDefer, W.candN.p, W.candChunk.p, W.candDistinct.p, W.candStride, W.dCtr.p);
            CK(cudaGetLastError());
            int dBlocks = (int)std::min<size_t>((2 * nWin + warpsPerBlock - 1) / warpsPerBlock,
                                                (size_t)M.smCount * (headroom ? 6 : 8));
            dp_lookup_kernel<<<dBlocks, 32 * DP_LWARPS, smem, st>>>(I, Q, (int)(2 * nWin), deferList, nDefer, S, inSmem,
                                                                   W.candN.p, W.candChunk.p, W.candDistinct.p,
                                                                   W.candStride, W.dCtr.p);
            CK(cudaGetLastError());
            W.stats.kernel_launches += 1;
        } else if (I.numChunks <= DP_SMALL_CHUNKS && env_int("DP_LOOKUP_SMALL", 1) != 0) {
            // small references: the lean register-resident kernel takes the window strands with at most 32 seeds (the
            // bulk), the general kernel the ones it hands back
            int* deferList = W.lbDefer.p;
            int* nDefer = reinterpret_cast<int*>(W.lbWork.p + 1);
            CK(cudaMemsetAsync(W.lbWork.p, 0, 4 * sizeof(unsigned), st));
            int sBlocks = (int)std::min<size_t>((2 * nWin + DP_SMALL_WARPS - 1) / DP_SMALL_WARPS,
                                                (size_t)M.smCount * (headroom ? 9 : 12));
            dp_lookup_small_kernel<<<sBlocks, 32 * DP_SMALL_WARPS, (size_t)DP_SMALL_WARPS * I.numChunks * sizeof(unsigned), st>>>(
                I, Q, (int)(2 * nWin), deferList, nDefer, W.candN.p, W.candChunk.p, W.candDistinct.p, W.candStride, W.dCtr.p);
            CK(cudaGetLastError());
            // (about a tenth of config 2's window strands have more than 32 seeds: few per warp, so the pass lasts as
            // long as one warp's serial chain of them — spread them over as many warps as fit)
            int dBlocks = (int)std::min<size_t>((2 * nWin + warpsPerBlock - 1) / warpsPerBlock,
                                                (size_t)M.smCount * (headroom ? 6 : 8));
            dp_lookup_kernel<<<dBlocks, 32 * DP_LWARPS, smem, st>>>(I, Q, (int)(2 * nWin), deferList, nDefer, S, inSmem,
                                                                   W.candN.p, W.candChunk.p, W.candDistinct.p,
                                                                   W.candStride, W.dCtr.p);
            CK(cudaGetLastError());
            W.stats.kernel_launches += 1;
        } else {
            int blocks = (int)std::min<size_t>((2 * nWin + warpsPerBlock - 1) / warpsPerBlock,
                                               (size_t)M.smCount * (headroom ? 6 : 8));
            dp_lookup_kernel<<<blocks, 32 * DP_LWARPS, smem, st>>>(I, Q, (int)(2 * nWin), nullptr, nullptr, S, inSmem,
                                                                   W.candN.p, W.candChunk.p, W.candDistinct.p,
                                                                   W.candStride, W.dCtr.p);
            CK(cudaGetLastError());
        }
        CK(cudaEventRecord(W.timers[T_LOOKUP].b, st));
    }
    {
        DpChainScratch S;
        S.qFirst = W.csQFirst.p;
        S.qCnt = W.csQCnt.p;
        S.qLo = W.csQLo.p;
        S.rqPos = W.csRqPos.p;
        S.rqId = W.csRqId.p;
        S.chainLen = W.csChainLen.p;
        S.lastB = W.csLastB.p;
        S.ent = W.csEnt.p;
        S.rsPos = W.csRsPos.p;
        S.rsId = W.csRsId.p;
        S.chains = W.csChains.p;
        S.results = W.csResults.p;
        S.qStride = qStride;
        S.sStride = (int)I.maxChunkSeeds + 8;
        S.chainCap = W.caps.chainCap;
        S.resultCap = W.caps.resultCap;
        int warpsPerBlock = 4;
        int blocks = (int)std::min<size_t>((nWin + warpsPerBlock - 1) / warpsPerBlock,
                                           (size_t)M.smCount * (headroom ? 4 : 6));
        blocks = std::min(blocks, W.chainWarps / warpsPerBlock);  // (the scratch is sized for chainWarps warps)
        const unsigned long long outCap = (unsigned long long)nWin * W.caps.outStride;
        const bool fast = !(getenv("DP_CHAIN_FAST") && atoi(getenv("DP_CHAIN_FAST")) == 0);
        CK(cudaEventRecord(W.timers[T_CHAIN].a, st));
        W.reduceTimed = fast;
        if (fast) {
            DpFastChain F;
            F.tasks = W.fcTasks.p;
            F.taskBase = W.fcTaskBase.p;
            F.pool = W.fcPool.p;
            F.cursors = W.fcCursors.p;
            F.taskCap = W.fcTaskCap;
            F.poolCap = W.fcPoolCap;
            F.slow = W.fcSlow.p;
            F.slowList = W.fcSlowList.p;
            F.nSlow = reinterpret_cast<int*>(W.fcCursors.p + 2);
            CK(cudaMemsetAsync(W.fcCursors.p, 0, 4 * sizeof(unsigned long long), st));
            int rBlocks = (int)std::min<size_t>((nWin + 3) / 4, (size_t)M.smCount * (headroom ? 6 : 8));
            {   // slab sizes of the warps' local allocators: at most half of either pool can be lost in slab tails
                const unsigned long long activeWarps = std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)rBlocks * 4, nWin));
                F.taskSlab = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>(64, F.taskCap / (2 * activeWarps)));
                F.poolSlab = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>(2048, F.poolCap / (2 * activeWarps)));
            }
            dp_reduce_kernel<<<rBlocks, 128, 0, st>>>(I, W.dWins.p, (int)nWin, Q, W.candN.p, W.candChunk.p,
                                                      W.candDistinct.p, W.candStride, S, F);
            CK(cudaGetLastError());
            CK(cudaEventRecord(W.timers[T_REDUCE].b, st));
            dp_chain_thread_kernel<<<div_up((long long)nWin, 128), 128, 0, st>>>(
                I, W.dWins.p, dReadLen, (int)nWin, Q, W.candN.p, W.candChunk.p, W.candDistinct.p, W.candStride, F,
                W.outN.p, W.outOff.p, W.outMaps.p, W.cursor.p + CUR_OUT, outCap, W.dCtr.p);
            CK(cudaGetLastError());
            // windows the fast path handed back (none on typical data: the kernel then exits at once)
            int sBlocks = std::min(blocks, M.smCount * 2);
            dp_chain_kernel<<<sBlocks, 128, 0, st>>>(I, W.dWins.p, F.slowList, F.nSlow, dReadLen, (int)nWin, Q, W.candN.p,
                                                     W.candChunk.p, W.candDistinct.p, W.candStride, S, W.outN.p,
                                                     W.outOff.p, W.outMaps.p, W.cursor.p + CUR_OUT, outCap, W.dCtr.p);
            CK(cudaGetLastError());
            W.stats.kernel_launches += 2;
        } else {
            dp_chain_kernel<<<blocks, 128, 0, st>>>(I, W.dWins.p, nullptr, nullptr, dReadLen, (int)nWin, Q, W.candN.p,
                                                    W.candChunk.p, W.candDistinct.p, W.candStride, S, W.outN.p,
                                                    W.outOff.p, W.outMaps.p, W.cursor.p + CUR_OUT, outCap, W.dCtr.p);
            CK(cudaGetLastError());
        }
        CK(cudaEventRecord(W.timers[T_CHAIN].b, st));
    }
    W.stats.kernel_launches += 3;
    W.stats.rounds += 1;
    W.stats.windows += (int64_t)nWin;
    W.pendingStageTimes = true;
}

void collect_stage_times(Lane& W) {  // call after a stream synchronize
    if (!W.pendingStageTimes) return;
    float ms;
    CK(cudaEventElapsedTime(&ms, W.timers[T_PACK].a, W.timers[T_PACK].b));
    W.stats.ms_pack += ms;
    CK(cudaEventElapsedTime(&ms, W.timers[T_EXTRACT].a, W.timers[T_EXTRACT].b));
    W.stats.ms_extract += ms;
    CK(cudaEventElapsedTime(&ms, W.timers[T_LOOKUP].a, W.timers[T_LOOKUP].b));
    W.stats.ms_lookup += ms;
    CK(cudaEventElapsedTime(&ms, W.timers[T_CHAIN].a, W.timers[T_CHAIN].b));
    W.stats.ms_chain += ms;
    if (W.reduceTimed) {  // fast path: [a .. reduce.b] is the list reduction, the rest the sequential chaining
        float mr;
        CK(cudaEventElapsedTime(&mr, W.timers[T_CHAIN].a, W.timers[T_REDUCE].b));
        W.stats.ms_reduce += mr;
        W.stats.ms_chain -= mr;
    }
    W.pendingStageTimes = false;
}

// How a lane's host thread waits for its stream. Spinning (the runtime's default with few contexts) answers fastest,
// but lanes x ranks-per-node threads spinning on fewer cores starve each other. DP_SYNC=spin|block
// overrides; by default block when LOCAL_WORLD_SIZE (torchrun) x lanes exceeds the cores this process may run on.
int lane_count();
bool sync_mode_blocking() {
    static const bool blocking = [] {
        const char* e = getenv("DP_SYNC");
        if (e && !strcmp(e, "spin")) return false;
        if (e && !strcmp(e, "block")) return true;
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        const int ranks = lw ? std::max(1, atoi(lw)) : 1;
        const unsigned cores = std::max(1u, std::thread::hardware_concurrency());
        return (unsigned)(ranks * lane_count()) > cores;
    }();
    return blocking;
}

void lane_sync(Lane& W) {
    if (sync_mode_blocking()) {
        CK(cudaEventRecord(W.evSync, W.stream));
        CK(cudaEventSynchronize(W.evSync));
    } else {
        CK(cudaStreamSynchronize(W.stream));
    }
}

// Copies the window results of the last launch_windows() to the pinned host mirrors (hOutN, hOutOff, hOutMaps) together
// with the attempt's counters. Returns the launch's overflow bits: non-zero means the results are incomplete and the
// caller must recompute with more room (nothing is copied then).
unsigned download_windows(Lane& W, size_t nWin) {
    cudaStream_t st = W.stream;
    unsigned long long cur[4];
    CK(cudaMemcpyAsync(cur, W.cursor.p, sizeof(cur), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(W.hCtr.p, W.dCtr.p, sizeof(DpCounters), cudaMemcpyDeviceToHost, st));
    W.hOutN.reserve(nWin);
    W.hOutOff.reserve(nWin);
    CK(cudaMemcpyAsync(W.hOutN.p, W.outN.p, nWin * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(W.hOutOff.p, W.outOff.p, nWin * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    lane_sync(W);
    collect_stage_times(W);
    if (W.hCtr.p->overflow) return W.hCtr.p->overflow;
    size_t total = (size_t)cur[CUR_OUT];
    if (total > W.outMaps.cap) throw std::runtime_error("internal error: window result pool overrun without a flag");
    W.hOutMaps.reserve(total + 1);
    if (total) {
        CK(cudaMemcpyAsync(W.hOutMaps.p, W.outMaps.p, total * sizeof(DpMappingDev), cudaMemcpyDeviceToHost, st));
        lane_sync(W);
    }
    W.hOutTotal = total;
    return 0;
}

// Host-supplied window list (rounds after the first, and the test probe). Returns the overflow bits (see above).
unsigned run_windows(dp_mapper& M, Lane& W, const DpWindow* wins, size_t nWin, const unsigned* dWords, const long long* dWordOff,
                     const int* dReadLen) {
    if (nWin == 0) return 0;
    size_t seedEntries = 0;
    for (size_t i = 0; i < nWin; i++) seedEntries += 2 * (size_t)(wins[i].len + 2);
    ensure_window_capacity(M, W, nWin, seedEntries);
    if (W.curAsciiIsHost)
        for (size_t i = 0; i < nWin; i++) W.stats.h2d_bytes += W.curPacked ? wins[i].len / 4 + 32 : wins[i].len + 32;
    CK(cudaMemcpyAsync(W.dWins.p, wins, nWin * sizeof(DpWindow), cudaMemcpyHostToDevice, W.stream));
    launch_windows(M, W, nWin, seedEntries, dWords, dWordOff, dReadLen);
    return download_windows(W, nWin);
}

// The device counters belong to one attempt at one sub-batch: zeroed when it starts, delivered to W.hCtr with its
// results (finish kernel / download_windows), added to the lane's statistics only when the attempt succeeded.
void reset_counters(Lane& W) {
    W.dCtr.reserve(1);
    W.hCtr.reserve(1);
    CK(cudaMemsetAsync(W.dCtr.p, 0, sizeof(DpCounters), W.stream));
}

void absorb_counters(Lane& W) {
    const DpCounters& c = *W.hCtr.p;
    W.stats.kmer_lookups += (int64_t)c.kmer_lookups;
    W.stats.query_seeds += (int64_t)c.query_seeds;
    W.stats.posting_runs += (int64_t)c.posting_runs;
    W.stats.posting_entries += (int64_t)c.posting_entries;
    W.stats.candidates += (int64_t)c.candidates;
    W.stats.chain_cells += (int64_t)c.chain_cells;
}

// An abandoned attempt keeps its time and launches in the statistics (they were spent) but not its work counts.
void forget_attempt(Lane& W, const dp_stats& before) {
    W.stats.windows = before.windows;
    W.stats.h2d_bytes = before.h2d_bytes;
    W.stats.retries += 1;
}

// Grows the flagged capacities fourfold. Returns false when one is at its ceiling already.
bool grow_caps(Caps& c, unsigned bits, unsigned numChunks) {
    bool ok = true;
    auto up = [&](int& v) {
        if (v >= kCapMax) ok = false;
        v = (int)std::min<long long>((long long)v * 4, kCapMax);
    };
    if (bits & DP_OV_OUTPOOL) up(c.outStride);
    if (bits & DP_OV_RESULTS) up(c.resultCap);
    if (bits & DP_OV_CHAINS) up(c.chainCap);
    if (bits & DP_OV_ROUNDS) {
        if (c.roundsScale >= 4096) ok = false;
        c.roundsScale = std::min(c.roundsScale * 4, 4096);
    }
    if (bits & DP_OV_CANDS) {
        if (c.candStride >= numChunks) ok = false;  // (a window strand has at most C candidates)
        c.candStride = (unsigned)std::min<unsigned long long>((unsigned long long)c.candStride * 4, numChunks);
    }
    return ok;
}

// Output of one sub-batch: mappings grouped per read in input order.
struct SubOut {
    std::vector<dp_mapping> maps;
};

static_assert(sizeof(dp_mapping) == sizeof(DpMappingDev), "ABI record and device record share one layout");

// Upper bound of the seed entries the round-0 windows of one read can produce (both strands; a strand visits at most
// len - k + 1 k-mers, one of them twice on a raw rc strand, Q2): what the compact seed lists of a launch are sized by.
inline size_t round0_seed_bound(long long len, int e, int minLen) {
    if (len < minLen) return 0;
    return len <= 2ll * e ? 2 * (size_t)(len + 2) : 4 * (size_t)(e + 2);
}

// Later rounds of Mapper.Map for the `nUn` reads round 0 left open (W.dUnres), entirely on the device: strategy kernel
// (one replay of Map() per open read against its window cache) -> the windows it asks for, through the same window
// kernels as round 0 -> collect kernel -> strategy kernel again, until no read asks for a window. Per round the host
// reads one counter block; the reads' final records arrive in hRdResN / hRdResOff / hRdResMaps of the lane that ran the
// rounds (indexed by slot = position in dUnres).
//
// Two steps, so that the rounds of one sub-batch can run NEXT TO round 0 of the lane's next sub-batch (a few hundred
// windows make ~40 tiny, latency-bound launches and three synchronisations: done in line they cost 16 % of a step while
// the big kernels wait): rounds_begin files the round-0 hits of the open reads in the window cache of lane R (on W's
// stream: they live in W's window-result pool, which W's next launch overwrites); rounds_run does the rest on R's
// stream with R's window buffers. R == W runs them in line (pieces, retries).
void rounds_begin(dp_mapper& M, Lane& W, Lane& R, int nUn, int64_t n, int minLen) {
    cudaStream_t st = W.stream;
    const size_t scale = (size_t)R.caps.roundsScale;
    // (sizes in powers of two from 1024 slots: the number of open reads changes from sub-batch to sub-batch, the buffers
    // should not — a cudaFree in the middle of a call synchronises the device under every lane)
    size_t slots = 1024;
    while (slots < (size_t)nUn) slots *= 2;
    const int nThreads = (int)std::min<size_t>(((size_t)std::max(nUn, 64) + 63) / 64 * 64, 4096);
    DpRoundsDev& D = R.roundsDev;
    memset(&D, 0, sizeof(D));
    // (DP_ROUNDS_HITS / DP_ROUNDS_LIST / DP_ROUNDS_CACHE: tests start from tiny capacities to walk the retries)
    D.hitCap = (int)std::min<size_t>((size_t)std::max(1, env_int("DP_ROUNDS_HITS", 256)) * scale, 1u << 20);
    D.listCap = (int)std::min<size_t>((size_t)std::max(1, env_int("DP_ROUNDS_LIST", 128)) * scale, 1u << 20);
    D.entCap = (unsigned)std::min<size_t>((32 * slots + 64) * scale, 0x7fffffffu);
    D.cacheCap = (unsigned)std::min<size_t>(((size_t)std::max(1, env_int("DP_ROUNDS_CACHE", 256)) * slots + 4096) * scale, 0xfffffff0u);
    D.resCap = (unsigned)std::min<size_t>((16 * slots + 1024) * scale, 0xfffffff0u);
    if (&R != &W) {
        // a rounds workspace sizes its window buffers for two windows per slot (what the replays of one round can ask for)
        // and their worst-case seed lists, in the same powers of two as the slots: they settle after the first sub-batches
        // and — the window list holds the requests — never move while the rounds run
        R.floorWins = std::max(R.floorWins, 2 * slots);
        R.floorSeeds = std::max(R.floorSeeds, R.floorWins * 2 * (size_t)(2 * M.edge + 2));
        R.dWins.reserve(R.floorWins + 64);
    }
    R.rdHead.reserve(slots);
    R.rdEnt.reserve(D.entCap);
    R.rdEnt2.reserve(D.entCap);
    R.rdCache.reserve(D.cacheCap);
    R.rdWinSlot.reserve(R.dWins.cap);
    R.rdDone.reserve(slots);
    R.rdResN.reserve(slots);
    R.rdResOff.reserve(slots);
    R.rdResMaps.reserve(D.resCap);
    R.rdHits.reserve((size_t)std::min<size_t>(slots, 4096) * D.hitCap);
    R.rdLists.reserve((size_t)std::min<size_t>(slots, 4096) * DP_RL_LISTS * D.listCap);
    R.rdCur.reserve(DP_RC_N);
    R.hRdCur.reserve(DP_RC_N);
    R.hRdResN.reserve(slots);
    R.hRdResOff.reserve(slots);
    R.rdThreads = nThreads;
    D.unres = W.dUnres.p;  // (these two tables move to R with the rest of the sub-batch's tables: the pointers stay valid)
    D.nSlots = nUn;
    D.readLen = W.dReadLen.p;
    D.head = R.rdHead.p;
    D.ent = R.rdEnt.p;
    D.ent2 = R.rdEnt2.p;
    D.cacheMaps = R.rdCache.p;
    D.wins = R.dWins.p;
    D.winSlot = R.rdWinSlot.p;
    D.winCap = (unsigned)std::min<size_t>(R.dWins.cap, (size_t)2 * (size_t)n);
    D.done = R.rdDone.p;
    D.resN = R.rdResN.p;
    D.resOff = R.rdResOff.p;
    D.resMaps = R.rdResMaps.p;
    D.hits = R.rdHits.p;
    D.lists = R.rdLists.p;
    D.cur = R.rdCur.p;
    D.edge = M.edge;
    D.circular = M.circular;
    D.refLen = M.refLen;
    CK(cudaMemsetAsync(R.rdCur.p, 0, DP_RC_N * sizeof(unsigned), st));
    CK(cudaMemsetAsync(R.rdHead.p, 0xff, (size_t)nUn * sizeof(int), st));
    dp_rounds_seed_kernel<<<div_up(nUn, 128), 128, 0, st>>>(D, minLen, W.outN.p, W.outOff.p, W.outMaps.p);
    CK(cudaGetLastError());
    W.stats.kernel_launches += 1;
}

// Returns the overflow bits of a capacity that was too small (nothing is delivered then).
unsigned rounds_run(dp_mapper& M, Lane& W, int nUn, int minLen) {
    cudaStream_t st = W.stream;
    const DpRoundsDev& R = W.roundsDev;
    const int nThreads = W.rdThreads;
    for (;;) {
        CK(cudaMemsetAsync(W.rdCur.p + DP_RC_REQ, 0, sizeof(unsigned), st));
        CK(cudaMemsetAsync(W.rdCur.p + DP_RC_OPEN, 0, 3 * sizeof(unsigned), st));  // open reads, summed window lengths
        dp_rounds_replay_kernel<<<nThreads / 64, 64, 0, st>>>(R, minLen);
        CK(cudaGetLastError());
        W.stats.kernel_launches += 1;
        CK(cudaMemcpyAsync(W.hRdCur.p, W.rdCur.p, DP_RC_N * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(W.hCtr.p, W.dCtr.p, sizeof(DpCounters), cudaMemcpyDeviceToHost, st));
        {
            const double tw = now_ms();
            lane_sync(W);
            W.hp[2] += now_ms() - tw;
        }
        collect_stage_times(W);
        if (W.hCtr.p->overflow) return W.hCtr.p->overflow;  // (of the window launch in front of this replay)
        if (W.hRdCur.p[DP_RC_OVF]) return DP_OV_ROUNDS;
        const size_t nReq = W.hRdCur.p[DP_RC_REQ];
        if (nReq == 0) break;
        if (nReq > R.winCap) throw std::runtime_error("internal error: more window requests than two per open read");
        const unsigned long long lenSum = (unsigned long long)W.hRdCur.p[DP_RC_LEN_LO] | ((unsigned long long)W.hRdCur.p[DP_RC_LEN_HI] << 32);
        const size_t seedEntries = 2 * ((size_t)lenSum + 2 * nReq);
        ensure_window_capacity(M, W, nReq, seedEntries);
        if (W.dWins.p != R.wins) throw std::runtime_error("internal error: the window list moved under the rounds");
        if (W.curAsciiIsHost) W.stats.h2d_bytes += (int64_t)(W.curPacked ? lenSum / 4 + 32 * nReq : lenSum + 32 * nReq);
        launch_windows(M, W, nReq, seedEntries, W.dWords.p, W.dWordOff.p, W.dReadLen.p);
        dp_rounds_collect_kernel<<<div_up((long long)nReq, 128), 128, 0, st>>>(R, (int)nReq, W.outN.p, W.outOff.p, W.outMaps.p,
                                                                              (unsigned long long)nReq * W.caps.outStride);
        CK(cudaGetLastError());
        W.stats.kernel_launches += 1;
    }
    const size_t total = W.hRdCur.p[DP_RC_RES];
    {
        size_t room = 4096;
        while (room < total + 1) room *= 2;
        W.hRdResMaps.reserve(room);
    }
    CK(cudaMemcpyAsync(W.hRdResN.p, W.rdResN.p, (size_t)nUn * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(W.hRdResOff.p, W.rdResOff.p, (size_t)nUn * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    if (total) CK(cudaMemcpyAsync(W.hRdResMaps.p, W.rdResMaps.p, total * sizeof(DpMappingDev), cudaMemcpyDeviceToHost, st));
    lane_sync(W);
    return 0;
}

// One sub-batch between the two halves of its processing (sub_begin: everything of round 0 enqueued; sub_complete: the
// results taken) and, when its later rounds were handed to the lane's rounds workspace, until they have run.
struct SubState {
    int64_t r0 = 0, n = 0;
    dp_stats before{};
    int64_t nShort = 0;
    int minLen = 0;
    bool viaHbm = false;
    size_t copyHead = 0;
    double tEnq = 0;
    // deferred rounds
    bool pending = false;
    size_t sI = 0;
    int nUn = 0;
    size_t devTotal = 0;
    std::vector<DpUnresolved> unres;
    dp_stats round0{};   // what round 0 added to the lane's statistics (taken back if the sub-batch has to be redone)
    dp_stats roundsBefore{};
};

template <class T>
void swap_buf(DBuf<T>& a, DBuf<T>& b) {
    std::swap(a.p, b.p);
    std::swap(a.cap, b.cap);
}
template <class T>
void swap_buf(HBuf<T>& a, HBuf<T>& b) {
    std::swap(a.p, b.p);
    std::swap(a.d, b.d);
    std::swap(a.cap, b.cap);
}

void finish_write_pass(dp_mapper& M, Lane& W, int64_t n, int minLen, bool viaHbm) {
    cudaStream_t st = W.stream;
    CK(cudaMemsetAsync(W.cursor.p + CUR_FIN, 0, sizeof(unsigned long long), st));
    dp_finish_round0_kernel<true><<<div_up(n, 128), 128, 0, st>>>(
        M.I, W.dReadLen.p, n, minLen, W.outN.p, W.outOff.p, W.outMaps.p, nullptr, W.dFinOff.p,
        viaHbm ? W.dFinMaps.p : W.hFinMaps.d, (unsigned long long)W.hFinMaps.cap, W.dUnres.p,
        reinterpret_cast<int*>(W.cursor.p + CUR_FIN), (int)n + 1, W.dCtr.p, W.hCtr.d);
    CK(cudaGetLastError());
}

const int kUnresHead = 4096;  // unresolved reads copied with the results; a longer list is fetched afterwards

// Enqueues round 0 of reads [r0, r1) (whose ASCII lives at dAscii + (offsets[i] - offsets[r0]) on the device) with the
// lane's current capacities (W.caps): read tables, windows, performMapping stages, Map()'s first decision, the copies of
// its results. Nothing is waited for.
void sub_begin(dp_mapper& M, Lane& W, const unsigned char* dAscii, const int64_t* offsets, const int64_t* byteOff, bool packed,
               int64_t r0, int64_t r1, SubState& S) {
    const int64_t n = r1 - r0;
    cudaStream_t st = W.stream;
    S.r0 = r0;
    S.n = n;
    S.before = W.stats;
    S.pending = false;
    reset_counters(W);
    const int k = M.k;
    const int e = M.edge;
    const int minLen = k + 12;  // shorter reads: the reference's scans over-read their slice (undefined); no mappings
    S.minLen = minLen;
    // ---- read tables on the device: lengths, packed-word offsets ----
    double t0 = now_ms();
    S.tEnq = t0;
    const size_t nR = std::max<size_t>((size_t)n, W.floorReads);  // what is reserved
    W.hRel.reserve(nR + 1);
    long long maxLen = 0;
    int64_t nWinReal = 0, nShort = 0;
    size_t seedEntries0 = 0;  // worst case: every k-mer of every round-0 window is a seed on both strands
    for (int64_t i = 0; i <= n; i++) W.hRel.p[i] = offsets[r0 + i] - offsets[r0];
    for (int64_t i = 0; i < n; i++) {
        long long len = W.hRel.p[i + 1] - W.hRel.p[i];
        if (len < 0) throw std::runtime_error("read offsets must be non-decreasing");
        maxLen = std::max(maxLen, len);
        if (len >= minLen) nWinReal += (len <= 2ll * e) ? 1 : 2;
        else nShort++;
        seedEntries0 += round0_seed_bound(len, e, minLen);
    }
    S.nShort = nShort;
    if (maxLen > 0x7fffff00ll) throw std::runtime_error("read too long");
    W.hp[0] += now_ms() - t0;
    const long long totalBytes = W.hRel.p[n];
    const size_t wordCap = std::max<size_t>((size_t)totalBytes, W.floorBytes) / 16 + 2 * nR + 16;
    W.dSeqOff.reserve(nR + 1);
    W.dWordsNeeded.reserve(nR + 1);
    W.dWordOff.reserve(nR + 1);
    W.dReadLen.reserve(nR);
    W.dWords.reserve(wordCap);
    CK(cudaMemcpyAsync(W.dSeqOff.p, W.hRel.p, ((size_t)n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    dp_read_table_kernel<<<div_up(n + 1, 256), 256, 0, st>>>(W.dSeqOff.p, n, W.dReadLen.p, W.dWordsNeeded.p);
    CK(cudaGetLastError());
    dp_exclusive_sum(W.dWordsNeeded.p, W.dWordOff.p, (long long)n + 1, W.scanTmp, st);
    W.stats.kernel_launches += 3;
    W.curAscii = dAscii;
    W.curPacked = byteOff != nullptr && packed;
    W.curSpans = byteOff != nullptr && !packed;
    if (byteOff) {  // reads addressed by their own offsets: dAscii points at the first byte of read r0
        W.hByteRel.reserve(nR);
        W.dByteOff.reserve(nR);
        for (int64_t i = 0; i < n; i++) W.hByteRel.p[i] = byteOff[r0 + i] - byteOff[r0];
        CK(cudaMemcpyAsync(W.dByteOff.p, W.hByteRel.p, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, st));
    }
    // ---- round 0 entirely on the device: windows, performMapping stages, Map()'s first decision ----
    const size_t nWin0 = 2 * (size_t)n;
    if (seedEntries0 >= 0xffffffffull) throw std::runtime_error("internal error: sub-batch cut too large");  // (map_batch_impl cuts by this bound)
    ensure_window_capacity(M, W, nWin0, seedEntries0);
    dp_round0_windows_kernel<<<div_up(n, 256), 256, 0, st>>>(W.dReadLen.p, n, e, minLen, W.dWins.p);
    CK(cudaGetLastError());
    launch_windows(M, W, nWin0, seedEntries0, W.dWords.p, W.dWordOff.p, W.dReadLen.p, /*bulkPull=*/true);
    W.stats.windows += nWinReal - (int64_t)nWin0;  // empty second slots of short reads are not window queries
    if (W.curAsciiIsHost) {  // bytes the windowed pack pulls over the link: the round-0 windows (+ one word ahead)
        for (int64_t i = 0; i < n; i++) {
            long long len = W.hRel.p[i + 1] - W.hRel.p[i];
            if (len >= minLen) {
                const long long b = (len <= 2ll * e) ? len : std::min<long long>(len, 2ll * (e + 32));
                W.stats.h2d_bytes += W.curPacked ? b / 4 + 16 : b;
            }
        }
    }
    // Map()'s first decision per read, delivered in read order: count pass, device-wide scan, write pass into HBM and one
    // copy of the block to page-locked host memory (DP_FINISH_HBM=0: the write pass stores straight into mapped host
    // memory instead — posted writes over PCIe from a kernel that then sits on its SMs for the length of the transfer)
    W.dFinN.reserve(nR + 1);
    W.dFinOff.reserve(nR + 1);
    W.hFinOff.reserve(nR + 1);
    W.dUnres.reserve(nR + 1);
    W.hUnres.reserve((size_t)kUnresHead);
    W.hUnresN.reserve(1);
    if (W.hFinMaps.cap < nR * 4 + 64) W.hFinMaps.reserve(nR * 4 + 64);
    S.viaHbm = env_int("DP_FINISH_HBM", 1) != 0;
    if (S.viaHbm) W.dFinMaps.reserve(W.hFinMaps.cap);
    S.copyHead = std::min<size_t>(W.hFinMaps.cap, (size_t)n + (size_t)n / 4 + 64);  // records copied before their count is known
    CK(cudaEventRecord(W.timers[T_FINISH].a, st));
    dp_finish_round0_kernel<false><<<div_up(n + 1, 128), 128, 0, st>>>(M.I, W.dReadLen.p, n, minLen, W.outN.p, W.outOff.p,
                                                                       W.outMaps.p, W.dFinN.p, nullptr, nullptr, 0, nullptr,
                                                                       nullptr, 0, W.dCtr.p, nullptr);
    CK(cudaGetLastError());
    dp_exclusive_sum(W.dFinN.p, W.dFinOff.p, (long long)n + 1, W.scanTmp, st);
    finish_write_pass(M, W, n, minLen, S.viaHbm);
    CK(cudaEventRecord(W.timers[T_FINISH].b, st));
    if (S.viaHbm) CK(cudaMemcpyAsync(W.hFinMaps.p, W.dFinMaps.p, S.copyHead * sizeof(DpMappingDev), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(W.hFinOff.p, W.dFinOff.p, ((size_t)n + 1) * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(W.hUnresN.p, W.cursor.p + CUR_FIN, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(W.hUnres.p, W.dUnres.p, (size_t)std::min<int64_t>(n, kUnresHead) * sizeof(DpUnresolved),
                       cudaMemcpyDeviceToHost, st));
    W.stats.kernel_launches += 5;
    W.stats.ms_host_logic += now_ms() - t0;
    W.hp[1] += now_ms() - t0;
}

// The sub-batch's output: the delivered block as it is, with the late results spliced in where the unresolved reads sit
// (they delivered nothing in round 0). F holds the block (hFinOff / hFinMaps), L the late results (hRdRes*). `dest` is
// asked for room once the number of records is known: the sub-batch's place in the call's result array when every
// sub-batch in front of it has reported its size already, else a block of its own.
typedef std::function<dp_mapping*(size_t)> SubDest;
void sub_assemble(Lane& F, Lane& L, const SubState& S, const std::vector<DpUnresolved>& unres, int64_t* counts, const SubDest& dest) {
    const int64_t n = S.n, r0 = S.r0;
    const int nUn = (int)unres.size();
    const unsigned* fOff = F.hFinOff.p;
    const size_t devTotal = fOff[n];
    const dp_mapping* fin = reinterpret_cast<const dp_mapping*>(F.hFinMaps.p);
    std::vector<int> order((size_t)nUn);  // slots by read
    for (int a = 0; a < nUn; a++) order[(size_t)a] = a;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return unres[(size_t)x].read < unres[(size_t)y].read; });
    size_t total = devTotal;
    for (int a = 0; a < nUn; a++) total += (size_t)L.hRdResN.p[a];
    dp_mapping* out = dest(total);
    size_t pos = 0, src = 0;
    int64_t prevRead = 0;
    for (int a = 0; a <= nUn; a++) {
        const int slot = a < nUn ? order[(size_t)a] : -1;
        const int64_t u = a < nUn ? unres[(size_t)slot].read : n;
        const size_t segEnd = fOff[u];  // finished reads [prevRead, u): one block
        if (segEnd > src) memcpy(out + pos, fin + src, (segEnd - src) * sizeof(dp_mapping));
        pos += segEnd - src;
        src = segEnd;
        for (int64_t i = prevRead; i < u; i++) counts[r0 + i] = (int64_t)(fOff[i + 1] - fOff[i]);
        if (a == nUn) break;
        const int nLate = L.hRdResN.p[slot];
        counts[r0 + u] = nLate;
        if (nLate) memcpy(out + pos, L.hRdResMaps.p + L.hRdResOff.p[slot], (size_t)nLate * sizeof(dp_mapping));
        pos += (size_t)nLate;
        prevRead = u + 1;
    }
    if (pos != total) throw std::runtime_error("internal error: sub-batch assembly mismatch");
}

dp_stats stats_diff(const dp_stats& a, const dp_stats& b) {
    dp_stats d = a;
    d.ms_pack -= b.ms_pack;
    d.ms_extract -= b.ms_extract;
    d.ms_lookup -= b.ms_lookup;
    d.ms_chain -= b.ms_chain;
    d.ms_reduce -= b.ms_reduce;
    d.ms_finish -= b.ms_finish;
    d.ms_host_logic -= b.ms_host_logic;
    d.ms_h2d -= b.ms_h2d;
    d.rounds -= b.rounds;
    d.windows -= b.windows;
    d.kmer_lookups -= b.kmer_lookups;
    d.query_seeds -= b.query_seeds;
    d.posting_runs -= b.posting_runs;
    d.posting_entries -= b.posting_entries;
    d.candidates -= b.candidates;
    d.chain_cells -= b.chain_cells;
    d.mappings -= b.mappings;
    d.kernel_launches -= b.kernel_launches;
    d.h2d_bytes -= b.h2d_bytes;
    d.retries -= b.retries;
    d.short_reads -= b.short_reads;
    return d;
}
void take_back_work(dp_stats& st, const dp_stats& d) {
    st.windows -= d.windows;
    st.h2d_bytes -= d.h2d_bytes;
    st.kmer_lookups -= d.kmer_lookups;
    st.query_seeds -= d.query_seeds;
    st.posting_runs -= d.posting_runs;
    st.posting_entries -= d.posting_entries;
    st.candidates -= d.candidates;
    st.chain_cells -= d.chain_cells;
    st.short_reads -= d.short_reads;
}

// Takes the results of round 0 (waits for the lane's stream). Returns the DP_OV_* bits of a capacity that was too small
// (nothing is delivered then and the caller reruns the range with more room, map_range); else 0 with counts[r0..r1) and
// `out` filled — or, when `defer` and the sub-batch has open reads, with S.pending set: their later rounds have been
// handed to the lane's rounds workspace (sub_finish_rounds takes them up) and nothing is filled yet.
unsigned sub_complete(dp_mapper& M, Lane& W, SubState& S, bool defer, int64_t* counts, const SubDest& dest) {
    cudaStream_t st = W.stream;
    const int64_t n = S.n;
    const int minLen = S.minLen;
    {
        const double tw = now_ms();
        lane_sync(W);
        W.hp[2] += now_ms() - tw;
    }
    collect_stage_times(W);
    if (M.evTrace) {  // DP_TRACE=1: the sub-batch's device timeline, ms since the call started
        auto at = [&](cudaEvent_t e) {
            float ms = -1;
            if (cudaEventElapsedTime(&ms, M.evTrace, e) != cudaSuccess) cudaGetLastError();
            return ms;
        };
        fprintf(stderr, "[dp trace] reads %lld+%lld host: enq %.2f synced %.2f | dev: pull %.2f-%.2f pack -%.2f extract %.2f-%.2f "
                        "lookup -%.2f reduce -%.2f chain -%.2f finish %.2f-%.2f\n",
                (long long)S.r0, (long long)n, S.tEnq - M.traceHost0, now_ms() - M.traceHost0, at(W.timers[T_PACK].a),
                at(W.timers[T_PULL].b), at(W.timers[T_PACK].b), at(W.timers[T_EXTRACT].a), at(W.timers[T_EXTRACT].b),
                at(W.timers[T_LOOKUP].b), W.reduceTimed ? at(W.timers[T_REDUCE].b) : -1.f, at(W.timers[T_CHAIN].b),
                at(W.timers[T_FINISH].a), at(W.timers[T_FINISH].b));
    }
    {
        float ms;
        CK(cudaEventElapsedTime(&ms, W.timers[T_FINISH].a, W.timers[T_FINISH].b));
        W.stats.ms_chain += ms;  // Map()'s pairing step is accounted with the chaining stage
        W.stats.ms_finish += ms;
    }
    if (W.hCtr.p->overflow) {  // a capacity of round 0 was too small: the caller grows it and reruns the range
        forget_attempt(W, S.before);
        return W.hCtr.p->overflow;
    }
    double t0 = now_ms();
    const unsigned* fOff = W.hFinOff.p;
    const size_t devTotal = fOff[n];
    if (devTotal > W.hFinMaps.cap) {  // more records than the delivery buffer holds (repeat-rich reads): grow, write again
        W.hFinMaps.reserve(devTotal + 64);
        if (S.viaHbm) W.dFinMaps.reserve(W.hFinMaps.cap);
        finish_write_pass(M, W, n, minLen, S.viaHbm);
        if (S.viaHbm) CK(cudaMemcpyAsync(W.hFinMaps.p, W.dFinMaps.p, devTotal * sizeof(DpMappingDev), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(W.hUnresN.p, W.cursor.p + CUR_FIN, sizeof(int), cudaMemcpyDeviceToHost, st));
        W.stats.kernel_launches += 1;
        lane_sync(W);
    } else if (S.viaHbm && devTotal > S.copyHead) {  // the tail the first copy did not cover
        CK(cudaMemcpyAsync(W.hFinMaps.p + S.copyHead, W.dFinMaps.p + S.copyHead, (devTotal - S.copyHead) * sizeof(DpMappingDev),
                           cudaMemcpyDeviceToHost, st));
        lane_sync(W);
    }
    const int nUn = *W.hUnresN.p;
    S.unres.resize((size_t)nUn);
    if (nUn > kUnresHead) {
        CK(cudaMemcpyAsync(S.unres.data(), W.dUnres.p, (size_t)nUn * sizeof(DpUnresolved), cudaMemcpyDeviceToHost, st));
        lane_sync(W);
    } else if (nUn > 0) {
        memcpy(S.unres.data(), W.hUnres.p, (size_t)nUn * sizeof(DpUnresolved));
    }
    W.stats.short_reads += S.nShort;
    W.stats.ms_host_logic += now_ms() - t0;
    W.hp[3] += now_ms() - t0;

    // ---- unresolved reads: the later rounds of Map() on the device (dp_rounds.cuh) ----
    if (nUn > 0 && defer && W.rounds) {
        // handed to the rounds workspace: the hits of the open reads are filed there (on this stream, before this lane's
        // next launch overwrites them), the sub-batch's tables and its delivered block move over, and the lane is free
        // for its next sub-batch; the rounds run on the other stream next to it
        Lane& R = *W.rounds;
        absorb_counters(W);  // round 0's counters are final
        S.round0 = stats_diff(W.stats, S.before);
        S.roundsBefore = R.stats;
        R.caps = W.caps;
        R.floorReads = R.floorBytes = 0;  // (its window buffers are sized by what the rounds ask for: floors that only double)
        R.curAscii = W.curAscii;
        R.curAsciiIsHost = W.curAsciiIsHost;
        R.curPacked = W.curPacked;
        R.curSpans = W.curSpans;
        // the buffers that trade places below: the workspace's copies as large as the lane's, once (equal capacities:
        // after the first sub-batches nothing is allocated in the middle of a call any more)
        R.dUnres.reserve_exact(W.dUnres.cap);
        R.dReadLen.reserve_exact(W.dReadLen.cap);
        R.dSeqOff.reserve_exact(W.dSeqOff.cap);
        R.dWordOff.reserve_exact(W.dWordOff.cap);
        R.dWords.reserve_exact(W.dWords.cap);
        R.dByteOff.reserve_exact(W.dByteOff.cap);
        R.hFinOff.reserve_exact(W.hFinOff.cap);
        R.hFinMaps.reserve_exact(W.hFinMaps.cap);
        reset_counters(R);
        rounds_begin(M, W, R, nUn, n, minLen);
        CK(cudaEventRecord(W.evRounds, st));
        CK(cudaStreamWaitEvent(R.stream, W.evRounds, 0));
        swap_buf(W.dUnres, R.dUnres);
        swap_buf(W.dReadLen, R.dReadLen);
        swap_buf(W.dSeqOff, R.dSeqOff);
        swap_buf(W.dWordOff, R.dWordOff);
        swap_buf(W.dWords, R.dWords);
        swap_buf(W.dByteOff, R.dByteOff);
        swap_buf(W.hFinOff, R.hFinOff);
        swap_buf(W.hFinMaps, R.hFinMaps);
        S.nUn = nUn;
        S.devTotal = devTotal;
        S.pending = true;
        return 0;
    }
    const double tRounds = now_ms();
    if (nUn > 0) {
        rounds_begin(M, W, W, nUn, n, minLen);
        if (unsigned ov = rounds_run(M, W, nUn, minLen)) {
            forget_attempt(W, S.before);
            W.stats.short_reads -= S.nShort;
            return ov;
        }
        if (M.evTrace)
            fprintf(stderr, "[dp trace] reads %lld+%lld later rounds of %d reads in line: host %.2f - %.2f ms\n", (long long)S.r0,
                    (long long)n, nUn, tRounds - M.traceHost0, now_ms() - M.traceHost0);
    }
    W.hp[4] += now_ms() - tRounds;
    t0 = now_ms();
    sub_assemble(W, W, S, S.unres, counts, dest);
    absorb_counters(W);
    W.stats.ms_host_logic += now_ms() - t0;
    W.hp[5] += now_ms() - t0;
    return 0;
}

// The later rounds of a sub-batch that sub_complete handed to the rounds workspace, then its assembly. Returns the
// overflow bits when a capacity of the rounds was too small: the statistics of the sub-batch are taken back and the
// caller maps it again in line (map_range), where every capacity has its way out.
unsigned sub_finish_rounds(dp_mapper& M, Lane& W, SubState& S, int64_t* counts, const SubDest& dest) {
    Lane& R = *W.rounds;
    S.pending = false;
    const double tRounds = now_ms();
    const unsigned ov = rounds_run(M, R, S.nUn, S.minLen);
    W.hp[4] += now_ms() - tRounds;
    if (M.evTrace)
        fprintf(stderr, "[dp trace] reads %lld+%lld later rounds of %d reads next to the lane's next sub-batch: host %.2f - %.2f ms\n",
                (long long)S.r0, (long long)S.n, S.nUn, tRounds - M.traceHost0, now_ms() - M.traceHost0);
    if (ov) {  // (an abandoned attempt keeps its time and launches in the statistics but not its work counts)
        take_back_work(R.stats, stats_diff(R.stats, S.roundsBefore));
        take_back_work(W.stats, S.round0);
        W.stats.retries += 1;
        return ov;
    }
    absorb_counters(R);
    const double t0 = now_ms();
    sub_assemble(R, R, S, S.unres, counts, dest);
    W.stats.ms_host_logic += now_ms() - t0;
    W.hp[5] += now_ms() - t0;
    return 0;
}

// One sub-batch in line: round 0, the later rounds, the assembly.
unsigned map_subbatch(dp_mapper& M, Lane& W, const unsigned char* dAscii, const int64_t* offsets, const int64_t* byteOff,
                      bool packed, int64_t r0, int64_t r1, int64_t* counts, SubOut& out) {
    SubState S;
    sub_begin(M, W, dAscii, offsets, byteOff, packed, r0, r1, S);
    return sub_complete(M, W, S, false, counts, [&](size_t total) {
        out.maps.resize(total);
        return out.maps.data();
    });
}

// (DP_CAP_OUT / DP_CAP_RESULTS / DP_CAP_CHAINS / DP_CAP_CANDS: tests start from tiny capacities to walk the retries)
Caps default_caps() {
    Caps c;
    c.outStride = std::max(1, std::min(kCapMax, env_int("DP_CAP_OUT", c.outStride)));
    c.resultCap = std::max(1, std::min(kCapMax, env_int("DP_CAP_RESULTS", c.resultCap)));
    c.chainCap = std::max(1, std::min(kCapMax, env_int("DP_CAP_CHAINS", c.chainCap)));
    c.candStride = (unsigned)std::max(1, env_int("DP_CAP_CANDS", (int)c.candStride));
    return c;
}

// map_subbatch with the exact way out of every device capacity: on overflow the flagged capacity grows fourfold and the
// range is recomputed; when the candidate lists would outgrow the lane's memory budget the range is cut into pieces
// first (a single read always fits: a window strand has at most C candidates). The lane's capacities return to their
// defaults afterwards (the buffers stay grown).
void map_range(dp_mapper& M, Lane& W, const unsigned char* dAscii, const int64_t* offsets, const int64_t* byteOff, bool packed,
               int64_t r0, int64_t r1, int64_t* counts, SubOut& out) {
    const size_t candBudget = std::max<size_t>(1, (size_t)env_int("DP_CAND_BUDGET_MB", 8192) << 20);  // (tests: 0 = single reads)
    for (;;) {
        const int64_t n = r1 - r0;
        const size_t candBytes = (size_t)4 * (size_t)n * std::min<size_t>(M.I.numChunks, W.caps.candStride) * 6;
        if (candBytes > candBudget && n > 1) {
            const int64_t pieces = std::min<int64_t>(n, (int64_t)((candBytes + candBudget - 1) / candBudget));
            out.maps.clear();
            for (int64_t p = 0; p < pieces; p++) {
                const int64_t a = r0 + n * p / pieces, b = r0 + n * (p + 1) / pieces;
                if (a == b) continue;
                SubOut part;
                const Caps keep = W.caps;
                const int64_t* so = byteOff ? byteOff : offsets;  // where a read starts in the caller's buffer
                map_range(M, W, dAscii + (so[a] - so[r0]), offsets, byteOff, packed, a, b, counts, part);
                W.caps = keep;
                out.maps.insert(out.maps.end(), part.maps.begin(), part.maps.end());
            }
            return;
        }
        const unsigned ov = map_subbatch(M, W, dAscii, offsets, byteOff, packed, r0, r1, counts, out);
        if (!ov) return;
        if (!grow_caps(W.caps, ov, M.I.numChunks))
            throw std::runtime_error("a window of this batch needs more than 2^20 mappings or chains on the device");
    }
}

void add_stats(dp_stats& a, const dp_stats& b) {
    a.ms_pack += b.ms_pack;
    a.ms_extract += b.ms_extract;
    a.ms_lookup += b.ms_lookup;
    a.ms_chain += b.ms_chain;
    a.ms_reduce += b.ms_reduce;
    a.ms_finish += b.ms_finish;
    a.ms_host_logic += b.ms_host_logic;
    a.ms_h2d += b.ms_h2d;
    a.rounds += b.rounds;
    a.windows += b.windows;
    a.kmer_lookups += b.kmer_lookups;
    a.query_seeds += b.query_seeds;
    a.posting_runs += b.posting_runs;
    a.posting_entries += b.posting_entries;
    a.candidates += b.candidates;
    a.chain_cells += b.chain_cells;
    a.mappings += b.mappings;
    a.kernel_launches += b.kernel_launches;
    a.h2d_bytes += b.h2d_bytes;
    a.retries += b.retries;
    a.short_reads += b.short_reads;
}

const size_t kMaxLanes = 12;  // per mapper, over all concurrent calls

// The lanes one call works with: up to `want` idle ones (created on demand while the mapper has fewer than kMaxLanes);
// a caller that finds every lane taken waits for the first to come back.
struct LaneSet {
    dp_mapper& M;
    std::vector<Lane*> lanes;
    std::vector<size_t> idx;
    LaneSet(dp_mapper& m, size_t want) : M(m) {
        std::unique_lock<std::mutex> lk(M.laneMu);
        for (;;) {
            for (size_t i = 0; i < M.lanes.size() && lanes.size() < want; i++)
                if (!M.laneBusy[i]) take(i);
            while (lanes.size() < want && M.lanes.size() < kMaxLanes) {
                std::unique_ptr<Lane> L(new Lane());
                CK(cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking));
                CK(cudaEventCreateWithFlags(&L->evReady, cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&L->evPulled, cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&L->evSync, cudaEventDisableTiming | cudaEventBlockingSync));
                L->timers.resize(T_N);
                for (auto& t : L->timers) t.init();
                CK(cudaEventCreateWithFlags(&L->evRounds, cudaEventDisableTiming));
                L->rounds.reset(new Lane());
                {   // the rounds' launches are tiny and latency-bound: in front of the queued blocks of the big kernels
                    int lo = 0, hi = 0;
                    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                    if (env_int("DP_ROUNDS_PRIO", 1))
                        CK(cudaStreamCreateWithPriority(&L->rounds->stream, cudaStreamNonBlocking, hi));
                    else
                        CK(cudaStreamCreateWithFlags(&L->rounds->stream, cudaStreamNonBlocking));
                }
                CK(cudaEventCreateWithFlags(&L->rounds->evReady, cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&L->rounds->evPulled, cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&L->rounds->evSync, cudaEventDisableTiming | cudaEventBlockingSync));
                L->rounds->timers.resize(T_N);
                for (auto& t : L->rounds->timers) t.init();
                M.lanes.push_back(std::move(L));
                M.laneBusy.push_back(0);
                take(M.lanes.size() - 1);
            }
            if (!lanes.empty()) return;
            M.laneCv.wait(lk);
        }
    }
    ~LaneSet() {
        {
            std::lock_guard<std::mutex> lk(M.laneMu);
            for (size_t i : idx) M.laneBusy[i] = 0;
        }
        M.laneCv.notify_all();
    }
    LaneSet(const LaneSet&) = delete;
    LaneSet& operator=(const LaneSet&) = delete;

   private:
    void take(size_t i) {
        M.laneBusy[i] = 1;
        lanes.push_back(M.lanes[i].get());
        idx.push_back(i);
    }
};

const int64_t kSubBatchReads = 1 << 16;
const int64_t kSubBatchBytes = 1ll << 30;
const size_t kSubBatchSeedEntries = 400u << 20;  // 3.2 GB of seed lists per lane at most

int lane_count() {
    const char* env = getenv("DP_LANES");
    int v = env ? atoi(env) : 6;
    return std::max(1, std::min(v, 8));
}

// Copies reads [r0,r1) to the lane's device buffer. Pinned (or registered) caller memory is copied directly;
// pageable memory goes through the lane's pinned staging buffer in two alternating pieces.
void upload_ascii(Lane& W, const uint8_t* bases, const int64_t* offsets, int64_t r0, int64_t r1, bool pinned) {
    size_t nbytes = (size_t)(offsets[r1] - offsets[r0]);
    W.dAscii.reserve(nbytes + 64);
    double t0 = now_ms();
    if (pinned) {
        CK(cudaMemcpyAsync(W.dAscii.p, bases + offsets[r0], nbytes, cudaMemcpyHostToDevice, W.stream));
    } else {
        const size_t piece = 16u << 20;
        W.hStage.reserve(2 * piece);
        cudaEvent_t ev[2];
        CK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        int slot = 0;
        for (size_t o = 0; o < nbytes; o += piece, slot ^= 1) {
            size_t len = std::min(piece, nbytes - o);
            if (o >= 2 * piece) CK(cudaEventSynchronize(ev[slot]));
            memcpy(W.hStage.p + (size_t)slot * piece, bases + offsets[r0] + o, len);
            CK(cudaMemcpyAsync(W.dAscii.p + o, W.hStage.p + (size_t)slot * piece, len, cudaMemcpyHostToDevice,
                               W.stream));
            CK(cudaEventRecord(ev[slot], W.stream));
        }
        CK(cudaStreamSynchronize(W.stream));
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
    }
    CK(cudaStreamSynchronize(W.stream));
    W.stats.ms_h2d += now_ms() - t0;
}

// Shared driver of the two batch entry points. `hostBases` xor `devBases` is set.
// `offsets`: cumulative bases (= byte offsets of back-to-back ASCII reads). `byteOff` (packed reads, or ASCII reads mapped where
// they lie in a file image): first byte of each read in the caller's buffer, n_reads + 1 entries (the last one bounds the buffer).
void map_batch_impl(dp_mapper& M, int64_t n_reads, const uint8_t* hostBases, const uint8_t* devBases,
                    const int64_t* offsets, dp_mapping** out, int64_t** out_offsets, const int64_t* byteOff = nullptr,
                    bool packed = false) {
    const int64_t* srcOff = byteOff ? byteOff : offsets;  // where a read starts in the caller's buffer
    CK(cudaSetDevice(M.device));
    double tStart = now_ms();
    const bool trace = getenv("DP_TRACE") != nullptr;
    if (trace) {
        if (!M.evTrace) CK(cudaEventCreate(&M.evTrace));
        CK(cudaEventRecord(M.evTrace, M.pullStream));
        M.traceHost0 = tStart;
    } else if (M.evTrace) {
        cudaEventDestroy(M.evTrace);
        M.evTrace = nullptr;
    }
    auto mark = [&](const char* what) {
        if (trace) fprintf(stderr, "[dp trace] host %.2f ms: %s\n", now_ms() - tStart, what);
    };
    // sub-batch boundaries. Sizes ramp up at the start and down at the end: the first pull / first kernels start
    // (and the last host pass ends) on a quarter-size piece, so less of the pipeline's fill and drain is exposed.
    std::vector<int64_t> cuts;
    cuts.push_back(0);
    size_t floorReads = 0, floorBytes = 0, floorSeeds = 0;  // the largest sub-batch: what every lane sizes its buffers for
    {
        const bool ramp = !(getenv("DP_RAMP") && atoi(getenv("DP_RAMP")) == 0);  // DP_RAMP=0: equal pieces (profiling)
        const int64_t full = std::max<int64_t>(4, std::min<int64_t>(kSubBatchReads, env_int("DP_SUB_READS", (int)kSubBatchReads)));  // (tests: many small sub-batches)
        const int64_t tail = ramp ? full / 4 + full / 2 : 0;
        int idx = ramp ? 0 : 2;
        for (int64_t r0 = 0; r0 < n_reads; idx++) {
            int64_t remaining = n_reads - r0;
            int64_t want = std::min<int64_t>(full, (full / 4) << std::min(idx, 2));
            if (remaining > tail) want = std::min(want, remaining - tail);
            else if (ramp && remaining > full / 4) want = remaining - full / 4;
            else want = remaining;
            // (also bounded by the worst-case seed entries of its round-0 windows: 8 bytes each in the lane's compact
            // seed lists, 32-bit offsets — large query_size values cut smaller sub-batches instead of failing)
            int64_t r1 = r0;
            size_t seedBound = 0;
            // (no read's round-0 windows produce more than 4 * (e + 2) entries: when the piece fits both limits by that
            // bound it is cut without looking at its reads — a million offsets walked one by one cost a millisecond
            // before the first kernel of the call could start)
            const int64_t rq = std::min<int64_t>(n_reads, r0 + want);
            const size_t quick = (size_t)(rq - r0) * 4 * (size_t)(M.edge + 2);
            if (offsets[rq] - offsets[r0] <= kSubBatchBytes && quick <= kSubBatchSeedEntries) {
                r1 = rq;
                seedBound = quick;
            }
            while (r1 < rq && offsets[r1 + 1] - offsets[r0] <= kSubBatchBytes) {
                seedBound += round0_seed_bound(offsets[r1 + 1] - offsets[r1], M.edge, M.k + 12);
                if (seedBound > kSubBatchSeedEntries && r1 > r0) break;
                r1++;
            }
            if (r1 == r0) {
                seedBound = round0_seed_bound(offsets[r0 + 1] - offsets[r0], M.edge, M.k + 12);
                r1 = r0 + 1;
            }
            floorSeeds = std::max(floorSeeds, seedBound);
            cuts.push_back(r1);
            r0 = r1;
        }
    }
    mark("sub-batch cuts done");
    const size_t nSub = cuts.size() - 1;
    LaneSet held(M, std::min<size_t>((size_t)lane_count(), std::max<size_t>(nSub, 1)));
    const int nLanes = (int)held.lanes.size();
    for (size_t sI = 0; sI < nSub; sI++) {
        floorReads = std::max(floorReads, (size_t)(cuts[sI + 1] - cuts[sI]));
        floorBytes = std::max(floorBytes, (size_t)(offsets[cuts[sI + 1]] - offsets[cuts[sI]]));
    }
    // Pinned (or registered) caller memory is read in place through its device mapping; pageable memory is staged.
    const unsigned char* mappedBase = nullptr;
    if (hostBases && !getenv("DP_NO_ZEROCOPY")) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, hostBases) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
            void* dptr = nullptr;
            if (cudaHostGetDevicePointer(&dptr, const_cast<uint8_t*>(hostBases), 0) == cudaSuccess)
                mappedBase = static_cast<const unsigned char*>(dptr);
        }
        cudaGetLastError();
    }
    for (int l = 0; l < nLanes; l++) {
        Lane& W = *held.lanes[(size_t)l];
        memset(&W.stats, 0, sizeof(W.stats));
        if (W.rounds) memset(&W.rounds->stats, 0, sizeof(W.rounds->stats));
        W.floorReads = floorReads;
        W.floorWins = 2 * floorReads;
        W.floorBytes = floorBytes;
        W.floorSeeds = floorSeeds;
    }
    std::vector<SubOut> subs(nSub);
    std::atomic<size_t> nextSub(0);
    std::vector<std::string> errs((size_t)nLanes);
    // The result arrays are filled while the call runs. A sub-batch reports its number of records as soon as it is known;
    // its place is known once every sub-batch in front of it has reported. In the common case (sub-batches finish roughly
    // in order) that is before its block is assembled, and the block is assembled where it belongs; else it is assembled
    // in a block of its own and copied later, by whichever lane thread has a moment (none is left with a queue of
    // copies when the last kernels have finished). Room for two records per read is reserved up front (untouched pages
    // cost nothing); a batch that needs more (repeat-rich reads) is finished the slow way below.
    struct {
        std::mutex mu;
        std::vector<char> reported, based, direct, ready, placed;
        std::vector<int64_t> total, base;
        size_t frontier = 0;       // first sub-batch that has not been given its place
        int64_t frontierBase = 0;  // records in front of it
        bool spilled = false;
        std::vector<size_t> waiting;  // given a place, assembled in a block of their own, not copied yet
    } pl;
    for (auto* v : {&pl.reported, &pl.based, &pl.direct, &pl.ready, &pl.placed}) v->assign(nSub, 0);
    pl.total.assign(nSub, 0);
    pl.base.assign(nSub, 0);
    size_t mapsCap = (size_t)n_reads * 2 + 4096;
    if (getenv("DP_RESULT_CAP")) mapsCap = (size_t)std::max(1, env_int("DP_RESULT_CAP", 1));  // (tests: force the slow way)
    // (off[i] holds read i's record COUNT, written by the lane that maps it, until its sub-batch is given its place)
    int64_t* off = (int64_t*)result_alloc(sizeof(int64_t) * ((size_t)n_reads + 1));
    dp_mapping* maps = (dp_mapping*)result_alloc(sizeof(dp_mapping) * mapsCap);
    if (!off || !maps) {
        dp_free(off);
        dp_free(maps);
        throw std::runtime_error("out of host memory for the result");
    }
    std::atomic<int> bad(0);
    auto counts_to_offsets = [&](size_t sI, int64_t base) {
        int64_t run = base;
        for (int64_t i = cuts[sI]; i < cuts[sI + 1]; i++) {
            const int64_t c = off[i];
            off[i] = run;
            run += c;
        }
        if (run != base + pl.total[sI]) bad.store(1);
    };
    auto place_one = [&](size_t sI, int64_t base) {  // a block of its own -> its place
        if (!subs[sI].maps.empty()) memcpy(maps + base, subs[sI].maps.data(), subs[sI].maps.size() * sizeof(dp_mapping));
        counts_to_offsets(sI, base);
        subs[sI].maps.clear();  // the block goes back to the mapper for the next sub-batch (of this or a later call)
        std::lock_guard<std::mutex> lk(M.spareMu);
        if (M.spare.size() < 2 * kMaxLanes) M.spare.push_back(std::move(subs[sI].maps));
        std::vector<dp_mapping>().swap(subs[sI].maps);
    };
    // (call with pl.mu held) gives places to the reported sub-batches at the frontier
    auto advance = [&]() {
        while (pl.frontier < nSub && pl.reported[pl.frontier] && !pl.spilled) {
            const size_t q = pl.frontier;
            if ((size_t)(pl.frontierBase + pl.total[q]) > mapsCap) {
                pl.spilled = true;
                break;
            }
            pl.base[q] = pl.frontierBase;
            pl.based[q] = 1;
            pl.frontierBase += pl.total[q];
            pl.frontier++;
            if (pl.ready[q] && !pl.placed[q]) pl.waiting.push_back(q);
        }
    };
    // the size of sub-batch q is known: its place in the result array if it has one already, else null
    auto report = [&](size_t q, size_t total) -> dp_mapping* {
        std::lock_guard<std::mutex> lk(pl.mu);
        pl.total[q] = (int64_t)total;
        pl.reported[q] = 1;
        advance();
        if (!pl.based[q]) return nullptr;
        pl.direct[q] = 1;
        return maps + pl.base[q];
    };
    // copies of sub-batches that wait for one (any lane thread, one at a time)
    auto drain = [&]() {
        for (;;) {
            size_t q;
            {
                std::lock_guard<std::mutex> lk(pl.mu);
                if (pl.waiting.empty()) return;
                q = pl.waiting.back();
                pl.waiting.pop_back();
                pl.placed[q] = 1;
            }
            place_one(q, pl.base[q]);
        }
    };
    // sub-batch q is assembled (in place, or in subs[q])
    auto deliver = [&](size_t q) {
        bool inPlace = false, copyNow = false;
        {
            std::lock_guard<std::mutex> lk(pl.mu);
            if (!pl.reported[q]) {  // (came through map_range: assembled in a block of its own before its size was reported)
                pl.total[q] = (int64_t)subs[q].maps.size();
                pl.reported[q] = 1;
                advance();
            }
            inPlace = pl.direct[q] != 0;
            pl.ready[q] = 1;
            if (!inPlace && pl.based[q] && !pl.placed[q]) {
                // (advance() may have queued it a moment ago: take it back, this thread copies it now)
                for (size_t i = 0; i < pl.waiting.size(); i++)
                    if (pl.waiting[i] == q) {
                        pl.waiting.erase(pl.waiting.begin() + (long)i);
                        break;
                    }
                pl.placed[q] = 1;
                copyNow = true;
            }
            if (inPlace) pl.placed[q] = 1;
        }
        if (inPlace) counts_to_offsets(q, pl.base[q]);
        if (copyNow) place_one(q, pl.base[q]);
        drain();
    };
    // The later rounds of a sub-batch run next to round 0 of the lane's next one when the reads are resident on the
    // device (DP_ROUNDS_DEFER=0/1 overrides): in line, all lanes reach their rounds at about the same time and the GPU
    // sees nothing but a few hundred windows' worth of tiny launches for 3 ms of an 18 ms step. Reads pulled out of
    // host memory keep them in line — there the step is the PCIe pulls and the rounds hide under them (measured: deferring
    // costs 1.5 ms per step). Never for pageable reads: they are staged in one buffer per lane.
    // And not for indexes whose lookup scratch lives in HBM (more than 24 000 chunks: 8 GB per workspace at human scale, and
    // a step of hundreds of milliseconds in which the rounds do not show): a second workspace per lane would be waste.
    const bool deferRounds = (hostBases ? env_int("DP_ROUNDS_DEFER", 0) != 0 && mappedBase : env_int("DP_ROUNDS_DEFER", 1) != 0) &&
                             M.I.numChunks <= kLookupSmemChunks;
    const size_t candBudget = std::max<size_t>(1, (size_t)env_int("DP_CAND_BUDGET_MB", 8192) << 20);
    auto work = [&](int l) {
        try {
            CK(cudaSetDevice(M.device));
            Lane& W = *held.lanes[(size_t)l];
            SubState pend;               // the sub-batch whose later rounds are in the lane's rounds workspace
            std::vector<size_t> redo;    // sub-batches whose rounds ran out of a capacity
            auto take_spare = [&](size_t q) {
                std::lock_guard<std::mutex> lk(M.spareMu);
                if (subs[q].maps.capacity() == 0 && !M.spare.empty()) {
                    subs[q].maps = std::move(M.spare.back());
                    M.spare.pop_back();
                }
            };
            auto deliver_timed = [&](size_t q) {
                const double tp = now_ms();
                deliver(q);
                W.hp[5] += now_ms() - tp;
            };
            auto source_of = [&](size_t q) -> const unsigned char* {
                const int64_t r0 = cuts[q], r1 = cuts[q + 1];
                W.curAsciiIsHost = false;
                if (hostBases && mappedBase) {
                    W.curAsciiIsHost = true;
                    return mappedBase + srcOff[r0];  // the kernels read the caller's pinned buffer in place
                }
                if (hostBases) {
                    upload_ascii(W, hostBases, srcOff, r0, r1, false);
                    W.stats.h2d_bytes += srcOff[r1] - srcOff[r0];
                    return W.dAscii.p;
                }
                return devBases + srcOff[r0];
            };
            auto in_line = [&](size_t q, const unsigned char* src) {  // with the way out of every capacity (map_range)
                W.caps = default_caps();
                take_spare(q);
                map_range(M, W, src, offsets, byteOff, packed, cuts[q], cuts[q + 1], off, subs[q]);
                W.caps = default_caps();
                deliver_timed(q);
            };
            auto dest_of = [&](size_t q) -> SubDest {
                return [&, q](size_t total) -> dp_mapping* {
                    if (dp_mapping* there = report(q, total)) return there;
                    take_spare(q);
                    subs[q].maps.resize(total);
                    return subs[q].maps.data();
                };
            };
            auto finish_pending = [&]() {  // the later rounds of the lane's previous sub-batch, then its delivery
                if (!pend.pending) return;
                const size_t q = pend.sI;
                if (sub_finish_rounds(M, W, pend, off, dest_of(q))) redo.push_back(q);  // (mapped again in line below)
                else deliver_timed(q);
            };
            for (;;) {
                size_t sI = nextSub.fetch_add(1);
                if (sI >= nSub) break;
                const int64_t r0 = cuts[sI], r1 = cuts[sI + 1];
                W.caps = default_caps();
                const size_t candBytes = (size_t)4 * (size_t)(r1 - r0) * std::min<size_t>(M.I.numChunks, W.caps.candStride) * 6;
                if (!deferRounds || candBytes > candBudget) {
                    finish_pending();
                    in_line(sI, source_of(sI));
                    continue;
                }
                // round 0 of this sub-batch is enqueued, THEN the rounds of the previous one run next to it
                const unsigned char* dA = source_of(sI);
                SubState cur;
                sub_begin(M, W, dA, offsets, byteOff, packed, r0, r1, cur);
                finish_pending();
                if (sub_complete(M, W, cur, true, off, dest_of(sI))) {  // a capacity of round 0 was too small
                    in_line(sI, dA);
                } else if (cur.pending) {
                    cur.sI = sI;
                    pend = std::move(cur);
                } else {
                    deliver_timed(sI);
                }
                W.caps = default_caps();
            }
            finish_pending();
            for (size_t q : redo) in_line(q, source_of(q));
            lane_sync(W);
        } catch (const std::exception& ex) {
            errs[(size_t)l] = ex.what();
            if (errs[(size_t)l].empty()) errs[(size_t)l] = "unknown error";
        }
    };
    mark("lanes start");
    std::vector<std::thread> th;
    for (int l = 1; l < nLanes; l++) th.emplace_back(work, l);
    work(0);
    for (auto& t : th) t.join();
    mark("lanes joined");
    for (auto& e : errs)
        if (!e.empty()) {
            dp_free(maps);
            dp_free(off);
            throw std::runtime_error(e);
        }
    drain();  // (copies that became possible after the lane that could have made them had finished)
    int64_t total = pl.frontierBase;
    if (pl.frontier < nSub) {  // more than two records per read: the rest is placed now, in an array of the exact size
        for (size_t sI = pl.frontier; sI < nSub; sI++) total += pl.total[sI];
        dp_mapping* grown = (dp_mapping*)result_alloc(sizeof(dp_mapping) * (size_t)(total ? total : 1));
        if (!grown) {
            dp_free(maps);
            dp_free(off);
            throw std::runtime_error("out of host memory for the result");
        }
        if (pl.frontierBase) memcpy(grown, maps, sizeof(dp_mapping) * (size_t)pl.frontierBase);
        dp_free(maps);
        maps = grown;
        int64_t base = pl.frontierBase;
        for (size_t sI = pl.frontier; sI < nSub; sI++) {
            place_one(sI, base);
            base += pl.total[sI];
        }
    }
    off[n_reads] = total;
    mark("results in place");
    if (bad.load()) {
        dp_free(maps);
        dp_free(off);
        throw std::runtime_error("internal error: result assembly mismatch");
    }
    if (getenv("DP_HOST_PROFILE")) {  // where the lanes' host threads spent the call (ms, summed over sub-batches)
        for (int l = 0; l < nLanes; l++) {
            Lane& W = *held.lanes[(size_t)l];
            fprintf(stderr, "[dp host] lane %d: tables %.2f launch %.2f wait %.2f post %.2f replay+rounds %.2f assemble %.2f | call %.2f ms\n",
                    l, W.hp[0], W.hp[1] - W.hp[0], W.hp[2], W.hp[3], W.hp[4], W.hp[5], now_ms() - tStart);
            for (double& x : W.hp) x = 0;
        }
    }
    {
        dp_stats st2;
        memset(&st2, 0, sizeof(st2));
        for (int l = 0; l < nLanes; l++) {
            add_stats(st2, held.lanes[(size_t)l]->stats);
            if (held.lanes[(size_t)l]->rounds) add_stats(st2, held.lanes[(size_t)l]->rounds->stats);
        }
        st2.bases = offsets[n_reads] - offsets[0];
        st2.mappings = total;
        st2.ms_total = now_ms() - tStart;
        std::lock_guard<std::mutex> lk(M.laneMu);
        M.stats = st2;
    }
    *out = maps;
    *out_offsets = off;
}

}  // namespace

namespace {
// Random 32-byte-sector gather over a table far larger than L2: the "HBM gather roofline" the lookup kernel is held
// against. Every thread reads `perThread` independent random sectors (one 4-byte word of each, so each
// gather costs a full 32 B sector of DRAM traffic, exactly like a short posting run).
__global__ void dp_gather_probe_kernel(const unsigned* __restrict__ table, unsigned long long nSectors, int perThread,
                                       unsigned* __restrict__ sink) {
    unsigned long long x = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 1;
    unsigned acc = 0;
    for (int i = 0; i < perThread; i += 4) {
        unsigned long long idx[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            x ^= x << 13;
            x ^= x >> 7;
            x ^= x << 17;
            idx[u] = (x % nSectors) * 8;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += __ldg(table + idx[u]);
    }
    if (acc == 0x12345678u) *sink = acc;
}

}  // namespace

namespace {

std::unique_ptr<dp_mapper> open_mapper(int device) {
    int nDev = 0;
    CK(cudaGetDeviceCount(&nDev));
    if (nDev <= 0) throw std::runtime_error("no CUDA device");
    if (device < 0 || device >= nDev) throw std::runtime_error("bad device ordinal");
    CK(cudaSetDevice(device));
    std::unique_ptr<dp_mapper> M(new dp_mapper());
    M->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    M->smCount = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&M->stream, cudaStreamNonBlocking));
    int lo = 0, hi = 0;  // numerically lowest value = highest priority
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&M->pullStream, cudaStreamNonBlocking, hi));
    // (per device; set here, before any lane thread exists)
    CK(cudaFuncSetAttribute(dp_extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(dp_pull_windows_kernel<DP_PULL_SLOTS_ASCII, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(dp_lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(dp_lookup_mid_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(dp_lookup_block_kernel<256, 4, 6, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(dp_lookup_block_kernel<256, 4, 6, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(dp_lookup_block_kernel<128, 8, 6, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    return M;
}

// ----------------------------------------------------------------------------------------------------------------
// Index image: every array the map path reads, packed into one relocatable byte image behind a small header. One
// image = one cudaMemcpy / one NCCL broadcast / one file write; a mapper opened from an image points into its copy.
// ----------------------------------------------------------------------------------------------------------------
enum { IX_TABLE = 0, IX_FILTER, IX_SEEDOFF, IX_SEEDCHUNKS, IX_POSTOFF, IX_POSTCHUNK, IX_POSTPOS, IX_CHUNKOFF,
       IX_CHUNKPOS, IX_CHUNKSEED, IX_CHUNKOFFSET, IX_CHUNKINSET, IX_CHUNKSCANLEN, IX_CHUNKLEN, IX_N };

struct DpImageHeader {
    char magic[8];  // "DPB200IX"
    uint32_t version;
    int32_t k, circular, seedRate, edge, chunkSize, filterBits;
    int64_t refLen;
    uint32_t numSeeds, numChunks, maxChunkSeeds, pad;
    int64_t nChunkPostings, nSeedPostings;
    uint64_t totalBytes;
    uint64_t off[IX_N], bytes[IX_N];
};
const uint32_t kImageVersion = 1;
const size_t kImageAlign = 256;

void image_layout(const dp_mapper& M, DpImageHeader& H) {
    memset(&H, 0, sizeof(H));
    memcpy(H.magic, "DPB200IX", 8);
    H.version = kImageVersion;
    H.k = M.k;
    H.circular = M.circular;
    H.seedRate = M.seedRate;
    H.edge = M.edge;
    H.chunkSize = M.chunkSize;
    H.filterBits = M.filterBits;
    H.refLen = M.refLen;
    H.numSeeds = M.I.numSeeds;
    H.numChunks = M.I.numChunks;
    H.maxChunkSeeds = M.I.maxChunkSeeds;
    H.nChunkPostings = M.nChunkPostings;
    H.nSeedPostings = M.nSeedPostings;
    const uint64_t S = M.I.numSeeds, C = M.I.numChunks, P1 = (uint64_t)M.nSeedPostings, P2 = (uint64_t)M.nChunkPostings;
    H.bytes[IX_TABLE] = (((uint64_t)1 << (2 * M.k)) / 32) * sizeof(uint2);
    H.bytes[IX_FILTER] = M.filterBits ? ((((uint64_t)1 << M.filterBits) + 31) / 32) * 4 : 0;
    H.bytes[IX_SEEDOFF] = (S + 1) * 4;
    H.bytes[IX_SEEDCHUNKS] = P1 * 4;
    H.bytes[IX_POSTOFF] = (S + 1) * 4;
    H.bytes[IX_POSTCHUNK] = P2 * 4;
    H.bytes[IX_POSTPOS] = P2 * 4;
    H.bytes[IX_CHUNKOFF] = (C + 1) * 4;
    H.bytes[IX_CHUNKPOS] = P2 * 4;
    H.bytes[IX_CHUNKSEED] = P2 * 4;
    H.bytes[IX_CHUNKOFFSET] = C * 8;
    H.bytes[IX_CHUNKINSET] = C * 8;
    H.bytes[IX_CHUNKSCANLEN] = C * 4;
    H.bytes[IX_CHUNKLEN] = C * 4;
    uint64_t at = (sizeof(DpImageHeader) + kImageAlign - 1) / kImageAlign * kImageAlign;
    for (int i = 0; i < IX_N; i++) {
        H.off[i] = at;
        at += (H.bytes[i] + kImageAlign - 1) / kImageAlign * kImageAlign;
    }
    H.totalBytes = at;
}

void check_image_header(const DpImageHeader& H, int64_t bytes) {
    if (memcmp(H.magic, "DPB200IX", 8) != 0) throw std::runtime_error("not a downpore_b200 index image");
    if (H.version != kImageVersion) throw std::runtime_error("index image version mismatch");
    if ((int64_t)H.totalBytes > bytes) throw std::runtime_error("index image truncated");
    // the same parameter ranges as dp_mapper_create, and the layout exactly as image_layout() derives it from the
    // header's counts: an image (it is also the on-disk index) that was cut, edited or written by another version is
    // refused here instead of giving out-of-bounds device reads later
    if (H.k < 5 || H.k > 15 || H.numChunks == 0 || H.numSeeds == 0) throw std::runtime_error("index image header corrupt");
    if (H.seedRate < H.k + 4 || H.edge < 4 * H.k || H.edge > 16000 || H.chunkSize > 60000 || H.chunkSize < 2 * H.edge ||
        H.refLen < 2ll * H.edge || H.refLen < H.seedRate || H.filterBits < 0 || H.filterBits > 2 * H.k || H.filterBits > 24 ||
        (H.circular != 0 && H.circular != 1))
        throw std::runtime_error("index image parameters out of range");
    if (H.nSeedPostings <= 0 || H.nChunkPostings <= 0 || H.nSeedPostings > H.nChunkPostings ||
        (uint64_t)H.nChunkPostings >= 0xffffffffull || H.maxChunkSeeds == 0 || (uint64_t)H.maxChunkSeeds > (uint64_t)H.nChunkPostings)
        throw std::runtime_error("index image counts corrupt");
    dp_mapper probe;
    probe.k = H.k;
    probe.filterBits = H.filterBits;
    probe.I.numSeeds = H.numSeeds;
    probe.I.numChunks = H.numChunks;
    probe.nSeedPostings = H.nSeedPostings;
    probe.nChunkPostings = H.nChunkPostings;
    DpImageHeader E;
    image_layout(probe, E);
    if (E.totalBytes != H.totalBytes) throw std::runtime_error("index image layout corrupt");
    for (int i = 0; i < IX_N; i++)
        if (H.off[i] != E.off[i] || H.bytes[i] != E.bytes[i]) throw std::runtime_error("index image layout corrupt");
}

// the CSR arrays of an image must end where the header says they do (a cheap whole-payload consistency check)
void check_image_payload(const dp_mapper& M, const DpImageHeader& H) {
    unsigned ends[3];
    CK(cudaMemcpy(&ends[0], M.I.seedOff + H.numSeeds, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&ends[1], M.I.postOff + H.numSeeds, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&ends[2], M.I.chunkOff + H.numChunks, 4, cudaMemcpyDeviceToHost));
    if (ends[0] != (unsigned)H.nSeedPostings || ends[1] != (unsigned)H.nChunkPostings || ends[2] != (unsigned)H.nChunkPostings)
        throw std::runtime_error("index image payload corrupt (posting offsets do not match the header)");
}

}  // namespace

#define API_TRY try {
#define API_CATCH                      \
    }                                  \
    catch (const std::exception& ex) { \
        g_err = ex.what();             \
        return 1;                      \
    }                                  \
    return 0;

void dp_set_last_error(const char* msg) { g_err = msg ? msg : ""; }

extern "C" {

const char* dp_last_error(void) { return g_err.c_str(); }
const char* dp_version(void) { return "downpore_b200 0.2 (sm_100a)"; }
void dp_free(void* p) { result_free(p); }

int dp_host_alloc(void** out, size_t bytes) {
    API_TRY
    if (!out) throw std::runtime_error("null argument");
    unsigned flags = cudaHostAllocPortable | cudaHostAllocMapped;
    if (getenv("DP_HOST_WC") && atoi(getenv("DP_HOST_WC"))) flags |= cudaHostAllocWriteCombined;
    CK(cudaHostAlloc(out, bytes ? bytes : 1, flags));
    API_CATCH
}

void dp_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int dp_mapper_create(const uint8_t* ref_ascii, int64_t ref_len, int circular, int k, const double* kmer_values,
                     int seed_rate, int edge_size, int chunk_size, int device, dp_mapper** out) {
    API_TRY
    if (!ref_ascii || !kmer_values || !out) throw std::runtime_error("null argument");
    if (k < 5 || k > 15) throw std::runtime_error("k must be in [5, 15] (packedKmerAt returns int32: sequence/asm_amd64.s:29)");
    if (seed_rate < k + 4) throw std::runtime_error("seed_rate must be at least k+4");
    if (edge_size < 4 * k || edge_size > 16000) throw std::runtime_error("query_size must be in [4k, 16000]");
    if (chunk_size > 60000) throw std::runtime_error("chunk_size must be at most 60000");
    if (chunk_size < 2 * edge_size || (long long)chunk_size * 10 - edge_size <= 0)
        throw std::runtime_error("chunk_size must be at least 2*query_size");
    if (ref_len < 2ll * edge_size || ref_len < seed_rate) throw std::runtime_error("reference shorter than 2*query_size");
    std::unique_ptr<dp_mapper> M = open_mapper(device);
    M->k = k;
    M->circular = circular ? 1 : 0;
    M->seedRate = seed_rate;
    M->edge = edge_size;
    M->chunkSize = chunk_size;
    M->refLen = ref_len;
    build_index(*M, ref_ascii, kmer_values);
    *out = M.release();
    API_CATCH
}

int dp_mapper_index_image_size(const dp_mapper* m, int64_t* bytes) {
    API_TRY
    if (!m || !bytes) throw std::runtime_error("null argument");
    DpImageHeader H;
    image_layout(*m, H);
    *bytes = (int64_t)H.totalBytes;
    API_CATCH
}

int dp_mapper_index_export(const dp_mapper* m, void* image, int64_t bytes) {
    API_TRY
    if (!m || !image) throw std::runtime_error("null argument");
    CK(cudaSetDevice(m->device));
    DpImageHeader H;
    image_layout(*m, H);
    if ((int64_t)H.totalBytes > bytes) throw std::runtime_error("image buffer too small");
    const void* src[IX_N] = {m->I.table,       m->I.filter,      m->I.seedOff,     m->I.seedChunks,  m->I.postOff,
                             m->I.postChunk,   m->I.postPos,     m->I.chunkOff,    m->I.chunkPos,    m->I.chunkSeed,
                             m->I.chunkOffset, m->I.chunkInset,  m->I.chunkScanLen, m->hChunkLen.data()};
    char* dst = static_cast<char*>(image);
    CK(cudaMemcpy(dst, &H, sizeof(H), cudaMemcpyDefault));
    for (int i = 0; i < IX_N; i++)
        if (H.bytes[i]) CK(cudaMemcpy(dst + H.off[i], src[i], H.bytes[i], cudaMemcpyDefault));
    CK(cudaDeviceSynchronize());
    API_CATCH
}

int dp_mapper_create_from_index(const void* image, int64_t bytes, int device, dp_mapper** out) {
    API_TRY
    if (!image || !out || bytes < (int64_t)sizeof(DpImageHeader)) throw std::runtime_error("bad argument");
    std::unique_ptr<dp_mapper> M = open_mapper(device);
    DpImageHeader H;
    CK(cudaMemcpy(&H, image, sizeof(H), cudaMemcpyDefault));
    check_image_header(H, bytes);
    M->image.reserve((size_t)H.totalBytes);
    CK(cudaMemcpy(M->image.p, image, (size_t)H.totalBytes, cudaMemcpyDefault));
    M->k = H.k;
    M->circular = H.circular;
    M->seedRate = H.seedRate;
    M->edge = H.edge;
    M->chunkSize = H.chunkSize;
    M->filterBits = H.filterBits;
    M->refLen = H.refLen;
    M->nChunkPostings = H.nChunkPostings;
    M->nSeedPostings = H.nSeedPostings;
    M->indexBytes = M->image.bytes();
    const unsigned char* b = M->image.p;
    DpIndexDev& I = M->I;
    I.k = H.k;
    I.circular = H.circular;
    I.edge = H.edge;
    I.maxWindow = 2 * H.edge;
    I.refLen = H.refLen;
    I.numSeeds = H.numSeeds;
    I.numChunks = H.numChunks;
    I.maxChunkSeeds = H.maxChunkSeeds;
    I.filterBits = H.filterBits;
    I.table = reinterpret_cast<const uint2*>(b + H.off[IX_TABLE]);
    I.filter = reinterpret_cast<const unsigned*>(b + H.off[IX_FILTER]);
    I.seedOff = reinterpret_cast<const unsigned*>(b + H.off[IX_SEEDOFF]);
    I.seedChunks = reinterpret_cast<const unsigned*>(b + H.off[IX_SEEDCHUNKS]);
    I.postOff = reinterpret_cast<const unsigned*>(b + H.off[IX_POSTOFF]);
    I.postChunk = reinterpret_cast<const unsigned*>(b + H.off[IX_POSTCHUNK]);
    I.postPos = reinterpret_cast<const int*>(b + H.off[IX_POSTPOS]);
    I.chunkOff = reinterpret_cast<const unsigned*>(b + H.off[IX_CHUNKOFF]);
    I.chunkPos = reinterpret_cast<const int*>(b + H.off[IX_CHUNKPOS]);
    I.chunkSeed = reinterpret_cast<const unsigned*>(b + H.off[IX_CHUNKSEED]);
    I.chunkOffset = reinterpret_cast<const long long*>(b + H.off[IX_CHUNKOFFSET]);
    I.chunkInset = reinterpret_cast<const long long*>(b + H.off[IX_CHUNKINSET]);
    I.chunkScanLen = reinterpret_cast<const int*>(b + H.off[IX_CHUNKSCANLEN]);
    const size_t C = H.numChunks;
    M->hChunkOffset.resize(C);
    M->hChunkInset.resize(C);
    M->hChunkScanLen.resize(C);
    M->hChunkLen.resize(C);
    CK(cudaMemcpy(M->hChunkOffset.data(), I.chunkOffset, C * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(M->hChunkInset.data(), I.chunkInset, C * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(M->hChunkScanLen.data(), I.chunkScanLen, C * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(M->hChunkLen.data(), b + H.off[IX_CHUNKLEN], C * 4, cudaMemcpyDeviceToHost));
    check_image_payload(*M, H);
    build_mid_postings(*M);
    *out = M.release();
    API_CATCH
}

void dp_mapper_destroy(dp_mapper* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    delete m;
}

int dp_mapper_map_batch_device(dp_mapper* m, int64_t n_reads, const uint8_t* d_bases, const int64_t* offsets,
                               dp_mapping** out, int64_t** out_offsets) {
    API_TRY
    if (!m || !offsets || !out || !out_offsets || n_reads < 0 || (!d_bases && n_reads > 0))
        throw std::runtime_error("bad argument");
    map_batch_impl(*m, n_reads, nullptr, d_bases, offsets, out, out_offsets);
    API_CATCH
}

int dp_mapper_map_batch(dp_mapper* m, int64_t n_reads, const uint8_t* bases, const int64_t* offsets, dp_mapping** out,
                        int64_t** out_offsets) {
    API_TRY
    if (!m || !offsets || !out || !out_offsets || n_reads < 0 || (!bases && n_reads > 0))
        throw std::runtime_error("bad argument");
    map_batch_impl(*m, n_reads, bases, nullptr, offsets, out, out_offsets);
    API_CATCH
}

int dp_mapper_map_batch_packed(dp_mapper* m, int64_t n_reads, const uint8_t* packed, const int64_t* byte_offsets,
                               const int64_t* lengths, dp_mapping** out, int64_t** out_offsets) {
    API_TRY
    if (!m || !byte_offsets || !lengths || !out || !out_offsets || n_reads < 0 || (!packed && n_reads > 0))
        throw std::runtime_error("bad argument");
    // cumulative bases of the reads (what the ASCII entry is handed): checked and summed by a few threads in pieces, in
    // a recycled block — a million reads walked by one thread into fresh memory cost 2-3 ms before the first kernel
    const double tIn = now_ms();
    struct Scratch {
        int64_t* p;
        explicit Scratch(size_t n) : p((int64_t*)result_alloc(sizeof(int64_t) * n)) {
            if (!p) throw std::runtime_error("out of host memory");
        }
        ~Scratch() { result_free(p); }
    } bases((size_t)n_reads + 1);
    {
        const int T = (int)std::max<int64_t>(1, std::min<int64_t>(8, n_reads / 65536));
        std::vector<int64_t> partial((size_t)T + 1, 0);
        std::atomic<int> badRead(0);
        auto piece = [&](int t, bool second) {
            const int64_t a = n_reads * t / T, b = n_reads * (t + 1) / T;
            if (!second) {
                int64_t sum = 0;
                bool ok = true;
                for (int64_t i = a; i < b; i++) {
                    ok &= lengths[i] >= 0 && byte_offsets[i + 1] - byte_offsets[i] >= (lengths[i] + 3) / 4;
                    sum += lengths[i];
                }
                partial[(size_t)t + 1] = sum;
                if (!ok) badRead.store(1);
            } else {
                int64_t run = partial[(size_t)t];
                for (int64_t i = a; i < b; i++) {
                    bases.p[i] = run;
                    run += lengths[i];
                }
            }
        };
        for (int pass = 0; pass < 2; pass++) {
            std::vector<std::thread> th;
            for (int t = 1; t < T; t++) th.emplace_back(piece, t, pass == 1);
            piece(0, pass == 1);
            for (auto& x : th) x.join();
            if (pass == 0) {
                if (badRead.load())
                    throw std::runtime_error("a packed read needs (length + 3) / 4 bytes between its offset and the next");
                for (int t = 0; t < T; t++) partial[(size_t)t + 1] += partial[(size_t)t];
            }
        }
        bases.p[n_reads] = partial[(size_t)T];
    }
    if (getenv("DP_TRACE")) fprintf(stderr, "[dp trace] packed entry: read table in %.2f ms\n", now_ms() - tIn);
    cudaPointerAttributes attr;
    const bool onDevice = n_reads > 0 && cudaPointerGetAttributes(&attr, packed) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    map_batch_impl(*m, n_reads, onDevice ? nullptr : packed, onDevice ? packed : nullptr, bases.p, out, out_offsets,
                   byte_offsets, true);
    API_CATCH
}

int dp_mapper_paf_line(const dp_mapper* m, const dp_mapping* mp, const char* query_name, int64_t query_len,
                       const char* ref_name, char* buf, int buf_len) {
    if (!m || !mp || !buf) return -1;
    long long mappedLength = mp->end - mp->start;
    if (m->circular && mappedLength < 0) mappedLength = m->refLen - mp->start + mp->end;
    int n = snprintf(buf, (size_t)buf_len, "%s\t%lld\t%d\t%lld\t%s\t%s\t%lld\t%lld\t%lld\t%d\t%lld\t255", query_name,
                     (long long)query_len, mp->q_offset, (long long)query_len - mp->q_inset, mp->rc ? "-" : "+",
                     ref_name, m->refLen, (long long)mp->start, (long long)mp->end, mp->ids, mappedLength);
    return (n < 0 || n >= buf_len) ? -1 : n;
}

// ---- input / output side on the device (dp_io.cuh) --------------------------------------------------------------
extern "C++" {
namespace {

template <class T>
void exclusive_sum(const int* in, T* out, size_t n, DBuf<unsigned char>& tmp, cudaStream_t st) {
    dp_exclusive_sum(in, out, (long long)n, tmp, st);
}

// the pointer a kernel of `device` can read `p` through; `staged` receives a device copy when `p` is pageable host memory
const unsigned char* device_view(const uint8_t* p, size_t bytes, DBuf<unsigned char>& staged, bool forceCopy = false) {
    cudaPointerAttributes attr;
    const bool known = cudaPointerGetAttributes(&attr, p) == cudaSuccess;
    cudaGetLastError();
    if (known && attr.type == cudaMemoryTypeDevice) return p;
    if (known && attr.type == cudaMemoryTypeManaged) return p;
    if (known && attr.type == cudaMemoryTypeHost && !forceCopy) {
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, const_cast<uint8_t*>(p), 0) == cudaSuccess) return static_cast<const unsigned char*>(d);
        cudaGetLastError();
    }
    staged.reserve(bytes + 64);
    CK(cudaMemcpy(staged.p, p, bytes, cudaMemcpyDefault));
    return staged.p;
}

}  // namespace
}  // extern "C++"

int dp_device_alloc(void** out, size_t bytes, int device) {
    API_TRY
    if (!out) throw std::runtime_error("bad argument");
    CK(cudaSetDevice(device));
    CK(cudaMalloc(out, bytes ? bytes : 1));
    API_CATCH
}

void dp_device_free(void* p, int device) {
    if (!p) return;
    cudaSetDevice(device);
    cudaFree(p);
}

int dp_device_copy(void* dst, const void* src, size_t bytes, int device) {
    API_TRY
    if ((!dst || !src) && bytes) throw std::runtime_error("bad argument");
    CK(cudaSetDevice(device));
    if (bytes) CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    API_CATCH
}

int dp_split_records(const uint8_t* image, int64_t bytes, int64_t min_length, int final_piece, int* is_fastq, int device,
                     dp_record** records, int64_t* n_records, int64_t* consumed) {
    API_TRY
    if (!records || !n_records || bytes < 0 || (!image && bytes > 0)) throw std::runtime_error("bad argument");
    static_assert(sizeof(dp_record) == sizeof(DpRecordDev), "dp_record layout");
    CK(cudaSetDevice(device));
    *records = nullptr;
    *n_records = 0;
    if (consumed) *consumed = final_piece ? bytes : 0;
    if (bytes == 0) {
        *records = (dp_record*)malloc(sizeof(dp_record));
        return 0;
    }
    cudaStream_t st = nullptr;  // (a one-off pass per file piece: the default stream keeps it simple)
    DBuf<unsigned char> staged, scanTmp, cls;
    // host memory is copied once (two passes read every byte; the link should carry it once)
    const unsigned char* buf = device_view(image, (size_t)bytes, staged, true);
    const unsigned mis = (unsigned)((unsigned long long)buf & 15ull);
    const long long nTiles = ((long long)bytes + mis + DP_IO_TILE - 1) / DP_IO_TILE;
    if (nTiles + 1 >= (1ll << 31)) throw std::runtime_error("file piece too large: at most 2^31 tiles of 64 bytes");
    DBuf<int> cnt, off;
    cnt.reserve((size_t)nTiles + 1);
    off.reserve((size_t)nTiles + 1);
    const int tb = 256;
    dp_line_starts_kernel<false><<<div_up(nTiles + 1, tb), tb, 0, st>>>(buf, bytes, nTiles, cnt.p, nullptr, nullptr);
    CK(cudaGetLastError());
    exclusive_sum(cnt.p, off.p, (size_t)nTiles + 1, scanTmp, st);
    int extra = 0;
    CK(cudaMemcpy(&extra, off.p + nTiles, sizeof(int), cudaMemcpyDeviceToHost));
    const long long nLines = 1 + (long long)extra;
    DBuf<long long> lineStart, candSeq, candName, res;
    lineStart.reserve((size_t)nLines);
    dp_line_starts_kernel<true><<<div_up(nTiles + 1, tb), tb, 0, st>>>(buf, bytes, nTiles, nullptr, off.p, lineStart.p);
    CK(cudaGetLastError());
    cls.reserve((size_t)nLines);
    dp_line_class_kernel<<<div_up(nLines, tb), tb, 0, st>>>(buf, lineStart.p, nLines, cls.p);
    CK(cudaGetLastError());
    candSeq.reserve((size_t)nLines);
    candName.reserve((size_t)nLines);
    res.reserve(4);
    unsigned char lastByte = 0;
    CK(cudaMemcpy(&lastByte, buf + bytes - 1, 1, cudaMemcpyDeviceToHost));
    dp_record_walk_kernel<<<1, 32, 0, st>>>(cls.p, nLines, lastByte == '\n', final_piece != 0, is_fastq && *is_fastq, candSeq.p,
                                            candName.p, res.p);
    CK(cudaGetLastError());
    long long hres[4];
    CK(cudaMemcpy(hres, res.p, sizeof(hres), cudaMemcpyDeviceToHost));
    if (hres[1]) throw std::runtime_error("Invalid fastq format (on + line)");  // the reference's log.Fatal (seqio.go:226,242)
    const long long nCand = hres[0];
    long long dropName = -1;
    if (!final_piece) {
        dropName = hres[2];
        if (dropName == 0) throw std::runtime_error("the piece holds no complete record after its first name line: pass a larger piece");
        long long cut = 0;
        CK(cudaMemcpy(&cut, lineStart.p + dropName, sizeof(long long), cudaMemcpyDeviceToHost));
        if (consumed) *consumed = cut;
        if (is_fastq) *is_fastq = (int)hres[3];
    } else if (is_fastq) {
        *is_fastq = 0;
    }
    DBuf<int> keep, koff;
    keep.reserve((size_t)nCand + 1);
    koff.reserve((size_t)nCand + 1);
    dp_record_finish_kernel<false><<<div_up(nCand + 1, tb), tb, 0, st>>>(buf, bytes, lineStart.p, nLines, candSeq.p, candName.p, nCand,
                                                                        min_length, dropName, keep.p, nullptr, nullptr);
    CK(cudaGetLastError());
    exclusive_sum(keep.p, koff.p, (size_t)nCand + 1, scanTmp, st);
    int nKept = 0;
    CK(cudaMemcpy(&nKept, koff.p + nCand, sizeof(int), cudaMemcpyDeviceToHost));
    DBuf<DpRecordDev> recs;
    recs.reserve((size_t)nKept + 1);
    dp_record_finish_kernel<true><<<div_up(nCand + 1, tb), tb, 0, st>>>(buf, bytes, lineStart.p, nLines, candSeq.p, candName.p, nCand,
                                                                       min_length, dropName, nullptr, koff.p, recs.p);
    CK(cudaGetLastError());
    dp_record* outRecs = (dp_record*)malloc(sizeof(dp_record) * (size_t)(nKept ? nKept : 1));
    if (!outRecs) throw std::runtime_error("out of host memory for the records");
    cudaError_t e = cudaMemcpy(outRecs, recs.p, sizeof(dp_record) * (size_t)nKept, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        free(outRecs);
        CK(e);
    }
    *records = outRecs;
    *n_records = nKept;
    API_CATCH
}

int dp_mapper_map_batch_spans(dp_mapper* m, int64_t n_reads, const uint8_t* image, const dp_record* records, dp_mapping** out,
                              int64_t** out_offsets) {
    API_TRY
    if (!m || !out || !out_offsets || n_reads < 0 || ((!image || !records) && n_reads > 0)) throw std::runtime_error("bad argument");
    std::vector<int64_t> bases((size_t)n_reads + 1, 0), starts((size_t)n_reads + 1, 0);
    int64_t bound = 0;
    for (int64_t i = 0; i < n_reads; i++) {
        const dp_record& r = records[i];
        if (r.seq_len < 0 || r.seq_start < 0 || (i > 0 && r.seq_start < records[i - 1].seq_start + records[i - 1].seq_len))
            throw std::runtime_error("records must lie in the image in ascending order without overlap");
        bases[(size_t)i + 1] = bases[(size_t)i] + r.seq_len;
        starts[(size_t)i] = r.seq_start;
        bound = r.seq_start + r.seq_len;
    }
    starts[(size_t)n_reads] = bound;
    cudaPointerAttributes attr;
    const bool onDevice = n_reads > 0 && cudaPointerGetAttributes(&attr, image) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    map_batch_impl(*m, n_reads, onDevice ? nullptr : image, onDevice ? image : nullptr, bases.data(), out, out_offsets,
                   starts.data(), false);
    API_CATCH
}

int dp_mapper_paf_block(const dp_mapper* m, int64_t n_reads, const uint8_t* image, const dp_record* records,
                        const dp_mapping* maps, const int64_t* out_offsets, const char* ref_name, char** text,
                        int64_t* text_bytes) {
    API_TRY
    if (!m || !text || !text_bytes || !ref_name || n_reads < 0 || ((!records || !out_offsets || !image) && n_reads > 0))
        throw std::runtime_error("bad argument");
    static_assert(sizeof(dp_mapping) == sizeof(DpMappingDev), "dp_mapping layout");
    CK(cudaSetDevice(m->device));
    const long long nMaps = n_reads > 0 ? out_offsets[n_reads] : 0;
    if (nMaps > 0 && !maps) throw std::runtime_error("bad argument");
    *text = nullptr;
    *text_bytes = 0;
    if (nMaps == 0) {
        *text = (char*)malloc(1);
        return 0;
    }
    DpPafParams P;
    memset(&P, 0, sizeof(P));
    P.refLen = m->refLen;
    P.circular = m->circular;
    const size_t rl = strlen(ref_name);
    if (rl >= sizeof(P.refName)) throw std::runtime_error("reference name longer than 255 bytes");
    memcpy(P.refName, ref_name, rl);
    P.refNameLen = (int)rl;
    // the names: read where they lie when the image is device-visible, else gathered into one block first
    DBuf<unsigned char> staged;
    std::vector<dp_record> recs(records, records + n_reads);
    const unsigned char* names = nullptr;
    {
        cudaPointerAttributes attr;
        const bool known = cudaPointerGetAttributes(&attr, image) == cudaSuccess;
        cudaGetLastError();
        if (known && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) names = image;
        else if (known && attr.type == cudaMemoryTypeHost) {
            void* d = nullptr;
            if (cudaHostGetDevicePointer(&d, const_cast<uint8_t*>(image), 0) == cudaSuccess) names = static_cast<const unsigned char*>(d);
            cudaGetLastError();
        }
        if (!names) {
            std::vector<unsigned char> blk;
            for (int64_t i = 0; i < n_reads; i++) {
                const int64_t a = (int64_t)blk.size();
                if (records[i].name_len < 0 || records[i].name_start < 0) throw std::runtime_error("bad record");
                blk.insert(blk.end(), image + records[i].name_start, image + records[i].name_start + records[i].name_len);
                recs[(size_t)i].name_start = a;
            }
            staged.reserve(blk.size() + 64);
            CK(cudaMemcpy(staged.p, blk.data(), blk.size(), cudaMemcpyHostToDevice));
            names = staged.p;
        }
    }
    DBuf<DpRecordDev> dRecs;
    DBuf<DpMappingDev> dMaps;
    DBuf<long long> dOutOff, lineOff;
    DBuf<int> lineLen;
    DBuf<unsigned char> scanTmp;
    DBuf<char> dText;
    dRecs.reserve((size_t)n_reads);
    dMaps.reserve((size_t)nMaps);
    dOutOff.reserve((size_t)n_reads + 1);
    lineLen.reserve((size_t)nMaps + 1);
    lineOff.reserve((size_t)nMaps + 1);
    CK(cudaMemcpy(dRecs.p, recs.data(), sizeof(dp_record) * (size_t)n_reads, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dMaps.p, maps, sizeof(dp_mapping) * (size_t)nMaps, cudaMemcpyDefault));
    CK(cudaMemcpy(dOutOff.p, out_offsets, sizeof(int64_t) * ((size_t)n_reads + 1), cudaMemcpyHostToDevice));
    const int tb = 128;
    dp_paf_kernel<false><<<div_up(nMaps + 1, tb), tb>>>(P, dMaps.p, nMaps, dOutOff.p, n_reads, names, dRecs.p, lineLen.p, nullptr, nullptr);
    CK(cudaGetLastError());
    exclusive_sum(lineLen.p, lineOff.p, (size_t)nMaps + 1, scanTmp, nullptr);
    long long total = 0;
    CK(cudaMemcpy(&total, lineOff.p + nMaps, sizeof(long long), cudaMemcpyDeviceToHost));
    dText.reserve((size_t)total + 1);
    dp_paf_kernel<true><<<div_up(nMaps + 1, tb), tb>>>(P, dMaps.p, nMaps, dOutOff.p, n_reads, names, dRecs.p, nullptr, lineOff.p, dText.p);
    CK(cudaGetLastError());
    char* outText = (char*)malloc((size_t)total + 1);
    if (!outText) throw std::runtime_error("out of host memory for the PAF text");
    cudaError_t e = cudaMemcpy(outText, dText.p, (size_t)total, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        free(outText);
        CK(e);
    }
    outText[total] = 0;
    *text = outText;
    *text_bytes = total;
    API_CATCH
}

int dp_mapper_get_stats(const dp_mapper* m, dp_stats* out) {
    if (!m || !out) return 1;
    *out = m->stats;
    return 0;
}

int dp_mapper_index_info(const dp_mapper* m, int64_t* out5) {
    if (!m || !out5) return 1;
    out5[0] = m->I.numSeeds;
    out5[1] = m->I.numChunks;
    out5[2] = m->nChunkPostings;
    out5[3] = m->nSeedPostings;
    out5[4] = (int64_t)m->indexBytes;
    return 0;
}

int dp_mapper_params(const dp_mapper* m, int64_t* out8) {
    if (!m || !out8) return 1;
    out8[0] = m->k;
    out8[1] = m->circular;
    out8[2] = m->refLen;
    out8[3] = m->edge;
    out8[4] = m->seedRate;
    out8[5] = m->chunkSize;
    out8[6] = m->filterBits;
    out8[7] = m->device;
    return 0;
}

int dp_mapper_seed_kmers(const dp_mapper* m, int64_t* out) {
    API_TRY
    if (!m || !out) throw std::runtime_error("null argument");
    CK(cudaSetDevice(m->device));
    size_t nTable = (size_t)((1ll << (2 * m->k)) / 32);
    std::vector<uint2> t(nTable);
    CK(cudaMemcpy(t.data(), m->I.table, nTable * sizeof(uint2), cudaMemcpyDeviceToHost));
    size_t p = 0;
    for (size_t w = 0; w < nTable; w++) {
        unsigned bits = t[w].x;
        while (bits) {
            int b = __builtin_ctz(bits);
            out[p++] = (int64_t)(w * 32 + (size_t)b);
            bits &= bits - 1;
        }
    }
    if (p != m->I.numSeeds) throw std::runtime_error("seed table inconsistent");
    API_CATCH
}

int dp_mapper_chunk(const dp_mapper* m, int64_t c, int64_t* fields4, int32_t* pos, int64_t* kmer) {
    API_TRY
    if (!m || !fields4) throw std::runtime_error("null argument");
    if (c < 0 || c >= (int64_t)m->I.numChunks) throw std::runtime_error("chunk id out of range");
    CK(cudaSetDevice(m->device));
    unsigned off[2];
    CK(cudaMemcpy(off, m->I.chunkOff + c, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost));
    unsigned n = off[1] - off[0];
    fields4[0] = m->hChunkOffset[(size_t)c];
    fields4[1] = m->hChunkInset[(size_t)c];
    fields4[2] = m->hChunkLen[(size_t)c];
    fields4[3] = n;
    if (pos && n) CK(cudaMemcpy(pos, m->I.chunkPos + off[0], n * sizeof(int), cudaMemcpyDeviceToHost));
    if (kmer && n) {
        std::vector<unsigned> ranks(n);
        CK(cudaMemcpy(ranks.data(), m->I.chunkSeed + off[0], n * sizeof(unsigned), cudaMemcpyDeviceToHost));
        std::vector<int64_t> seeds(m->I.numSeeds);
        if (dp_mapper_seed_kmers(m, seeds.data())) throw std::runtime_error(g_err);
        for (unsigned i = 0; i < n; i++) kmer[i] = seeds[ranks[i]];
    }
    API_CATCH
}

int dp_mapper_probe_window(dp_mapper* m, const uint8_t* read_ascii, int64_t read_len, int64_t start, int64_t end,
                           int whole, int32_t* n_seeds2, int32_t* seed_pos, int64_t* seed_kmer, int64_t seed_cap,
                           int32_t* n_cand2, int32_t* cand, int64_t cand_cap, int32_t* n_map, dp_mapping* maps,
                           int64_t map_cap) {
    API_TRY
    if (!m || !read_ascii) throw std::runtime_error("null argument");
    CK(cudaSetDevice(m->device));
    if (whole) {
        start = 0;
        end = read_len;
    }
    if (start < 0 || end > read_len || end - start < m->k + 12 || end - start > 2 * m->edge)
        throw std::runtime_error("bad probe window");
    LaneSet held(*m, 1);
    Lane& W = *held.lanes[0];
    memset(&W.stats, 0, sizeof(W.stats));
    cudaStream_t st = W.stream;
    W.dAscii.reserve((size_t)read_len + 64);
    CK(cudaMemcpyAsync(W.dAscii.p, read_ascii, (size_t)read_len, cudaMemcpyHostToDevice, st));
    long long seqOff[2] = {0, read_len};
    long long wordOff[1] = {0};
    int rl[1] = {(int)read_len};
    W.dSeqOff.reserve(2);
    W.dWordOff.reserve(1);
    W.dReadLen.reserve(1);
    W.dWords.reserve((size_t)read_len / 16 + 8);
    CK(cudaMemcpyAsync(W.dSeqOff.p, seqOff, sizeof(seqOff), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(W.dWordOff.p, wordOff, sizeof(wordOff), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(W.dReadLen.p, rl, sizeof(rl), cudaMemcpyHostToDevice, st));
    W.curAscii = W.dAscii.p;
    W.curAsciiIsHost = false;
    W.curPacked = false;
    DpWindow w;
    w.read = 0;
    w.start = (int)start;
    w.len = (int)(end - start);
    w.whole = whole ? 1 : 0;
    W.caps = default_caps();
    for (;;) {  // (the same exact way out of the device capacities as map_range)
        reset_counters(W);
        const unsigned ov = run_windows(*m, W, &w, 1, W.dWords.p, W.dWordOff.p, W.dReadLen.p);
        if (!ov) break;
        if (!grow_caps(W.caps, ov, m->I.numChunks)) throw std::runtime_error("window needs more than 2^20 mappings or chains");
    }
    absorb_counters(W);
    W.caps = default_caps();
    // seeds
    unsigned wsOff[2];
    int wsN[2];
    CK(cudaMemcpy(wsOff, W.wsOff.p, sizeof(wsOff), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(wsN, W.wsN.p, sizeof(wsN), cudaMemcpyDeviceToHost));
    if (n_seeds2) {
        n_seeds2[0] = wsN[0];
        n_seeds2[1] = wsN[1];
    }
    if (seed_pos || seed_kmer) {
        int tot = wsN[0] + wsN[1];
        if (tot > seed_cap) throw std::runtime_error("seed_cap too small");
        std::vector<int64_t> seeds(m->I.numSeeds);
        if (dp_mapper_seed_kmers(m, seeds.data())) throw std::runtime_error(g_err);
        std::vector<unsigned> r((size_t)tot + 1);
        std::vector<int> p((size_t)tot + 1);
        // strand 0 then strand 1 (they are adjacent in the compact buffer)
        CK(cudaMemcpy(r.data(), W.qSeed.p + wsOff[0], (size_t)tot * sizeof(unsigned), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(p.data(), W.qPos.p + wsOff[0], (size_t)tot * sizeof(int), cudaMemcpyDeviceToHost));
        for (int i = 0; i < tot; i++) {
            if (seed_pos) seed_pos[i] = p[(size_t)i];
            if (seed_kmer) seed_kmer[i] = seeds[r[(size_t)i]];
        }
    }
    if (n_cand2 || cand) {
        int cn[2];
        CK(cudaMemcpy(cn, W.candN.p, sizeof(cn), cudaMemcpyDeviceToHost));
        if (n_cand2) {
            n_cand2[0] = cn[0];
            n_cand2[1] = cn[1];
        }
        if (cand) {
            if (cn[0] + cn[1] > cand_cap) throw std::runtime_error("cand_cap too small");
            std::vector<unsigned> c0((size_t)cn[0] + 1), c1((size_t)cn[1] + 1);
            CK(cudaMemcpy(c0.data(), W.candChunk.p, (size_t)cn[0] * sizeof(unsigned), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(c1.data(), W.candChunk.p + W.candStride, (size_t)cn[1] * sizeof(unsigned),
                          cudaMemcpyDeviceToHost));
            for (int i = 0; i < cn[0]; i++) cand[i] = (int32_t)c0[(size_t)i];
            for (int i = 0; i < cn[1]; i++) cand[cn[0] + i] = (int32_t)c1[(size_t)i];
        }
    }
    if (n_map) *n_map = W.hOutN.p[0];
    if (maps) {
        int cnt = W.hOutN.p[0];
        if (cnt > map_cap) throw std::runtime_error("map_cap too small");
        for (int i = 0; i < cnt; i++) {
            const DpMappingDev& d = W.hOutMaps.p[W.hOutOff.p[0] + i];
            dp_mapping o;
            memset(&o, 0, sizeof(o));
            o.start = d.start;
            o.end = d.end;
            o.q_offset = d.qOffset;
            o.q_inset = d.qInset;
            o.ids = d.ids;
            o.rc = (uint8_t)(d.rc & 0xff);
            maps[i] = o;
        }
    }
    API_CATCH
}

int dp_probe_gather_gbs(int device, int64_t table_bytes, double* sector_gbs) {
    API_TRY
    if (!sector_gbs || table_bytes < (1 << 20)) throw std::runtime_error("bad argument");
    CK(cudaSetDevice(device));
    DBuf<unsigned> tab, sink;
    size_t nWords = (size_t)table_bytes / 4;
    tab.reserve(nWords);
    sink.reserve(1);
    CK(cudaMemset(tab.p, 1, nWords * 4));
    const unsigned long long nSectors = nWords / 8;
    const int perThread = 64, threads = 256, blocks = 148 * 64;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {
        CK(cudaEventRecord(a));
        dp_gather_probe_kernel<<<blocks, threads>>>(tab.p, nSectors, perThread, sink.p);
        CK(cudaGetLastError());
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (it && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *sector_gbs = 32.0 * perThread * threads * blocks / (best * 1e-3) / 1e9;
    API_CATCH
}

int dp_pack(const uint8_t* ascii, int64_t len, uint8_t* out, int device) {
    API_TRY
    if (!ascii || !out || len <= 0) throw std::runtime_error("bad argument");
    CK(cudaSetDevice(device));
    DBuf<unsigned char> dA, dOut;
    DBuf<unsigned> dW;
    DBuf<long long> dOff, dWord;
    size_t nWords = (size_t)(len + 15) / 16;
    size_t nBytes = (size_t)(len + 3) / 4;
    dA.reserve((size_t)len + 64);
    dW.reserve(nWords + 8);
    dOut.reserve(nBytes + 8);
    // slices of 16-base multiples are independent sequences for the kernel
    const long long slice = 1 << 14;
    long long nSl = (len + slice - 1) / slice;
    std::vector<long long> so((size_t)nSl + 1), sw((size_t)nSl);
    for (long long i = 0; i < nSl; i++) {
        so[(size_t)i] = i * slice;
        sw[(size_t)i] = i * slice / 16;
    }
    so[(size_t)nSl] = len;
    dOff.reserve(so.size());
    dWord.reserve(sw.size());
    CK(cudaMemcpy(dA.p, ascii, (size_t)len, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dOff.p, so.data(), so.size() * sizeof(long long), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dWord.p, sw.data(), sw.size() * sizeof(long long), cudaMemcpyHostToDevice));
    dp_pack_kernel<<<(int)std::min<long long>((nSl + 7) / 8, 148 * 16), 256>>>(dA.p, dOff.p, dWord.p, dW.p, nSl);
    CK(cudaGetLastError());
    dp_words_to_bytes_kernel<<<div_up((long long)nBytes, 256), 256>>>(dW.p, dOut.p, (long long)nBytes);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dOut.p, nBytes, cudaMemcpyDeviceToHost));
    API_CATCH
}

int dp_kmer_counts(const uint8_t* ascii, int64_t len, int k, uint64_t* counts, int device) {
    API_TRY
    if (!ascii || !counts || len < k || k < 1 || k > 15) throw std::runtime_error("bad argument");
    CK(cudaSetDevice(device));
    DBuf<unsigned char> dA;
    DBuf<unsigned> dW;
    DBuf<long long> dOff, dWord;
    DBuf<unsigned long long> dC;
    size_t nWords = (size_t)(len + 15) / 16;
    size_t nK = (size_t)1 << (2 * k);
    dA.reserve((size_t)len + 64);
    dW.reserve(nWords + 8);
    dC.reserve(nK);
    CK(cudaMemset(dW.p, 0, dW.cap * sizeof(unsigned)));
    const long long slice = 1 << 14;
    long long nSl = (len + slice - 1) / slice;
    std::vector<long long> so((size_t)nSl + 1), sw((size_t)nSl);
    for (long long i = 0; i < nSl; i++) {
        so[(size_t)i] = i * slice;
        sw[(size_t)i] = i * slice / 16;
    }
    so[(size_t)nSl] = len;
    dOff.reserve(so.size());
    dWord.reserve(sw.size());
    CK(cudaMemcpy(dA.p, ascii, (size_t)len, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dOff.p, so.data(), so.size() * sizeof(long long), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dWord.p, sw.data(), sw.size() * sizeof(long long), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC.p, counts, nK * sizeof(uint64_t), cudaMemcpyHostToDevice));
    dp_pack_kernel<<<(int)std::min<long long>((nSl + 7) / 8, 148 * 16), 256>>>(dA.p, dOff.p, dWord.p, dW.p, nSl);
    CK(cudaGetLastError());
    dp_kmer_hist_kernel<<<148 * 8, 256>>>(dW.p, len, k, dC.p);
    CK(cudaGetLastError());
    CK(cudaMemcpy(counts, dC.p, nK * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    API_CATCH
}

}  // extern "C"

#include "dp_overlap_api.cuh"
