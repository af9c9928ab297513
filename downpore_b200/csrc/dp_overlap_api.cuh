// downpore_b200 — host driver and C ABI of the `overlap` path (dp_overlapper_*, include/downpore_b200.h). Included at the
// end of dp_api.cu (one translation unit: the kernels of dp_map.cuh / dp_index.cuh are shared). No CPU fallback.
#pragma once
#include "dp_overlap.cuh"

struct dp_overlapper {
    int device = 0;
    int smCount = 148;
    cudaStream_t st = nullptr;
    OvParams P{};
    long long nReads = 0, totalBases = 0, totalWords = 0;
    // the sequence set (sequence.NewFastaSequenceSet with himem: every read cached, packed)
    DBuf<unsigned> words;
    DBuf<long long> readBase;
    DBuf<int> readLen;
    DBuf<double> values;
    bool haveValues = false;
    // one round
    DBuf<unsigned char> ignore;
    DBuf<unsigned> bits, regKmer, kmerOfRank, regOfRank, rankOfReg, pc, prefix;
    DBuf<uint2> table;
    DBuf<OvSlice> slices;
    DBuf<OvSelectOut> selOut;
    DBuf<unsigned> err;
    DBuf<DpChunkDesc> descs;
    DBuf<unsigned> counts, fOff, fSeed;
    DBuf<int> fPos;
    DBuf<unsigned> qOff, qSeed, qDistinct;
    DBuf<int> qPos, qND;
    DBuf<unsigned short> qSlot;
    DBuf<unsigned> rStart, rCount, rSeed, filter;
    unsigned long long seedCap = 0;
    bool scanAttr = false;
    DBuf<int> rPos;
    DBuf<unsigned> pieceOff, keyOff;
    DBuf<OvChunk> chunks;
    DBuf<unsigned long long> keys, keysSorted, keysScratch, nSel;
    DBuf<unsigned> sortHist, uniquePos;
    DBuf<unsigned char> scanTmp;  // scratch of dp_exclusive_scan only (its first bytes are the scan's ticket)
    DBuf<unsigned> seedOff, seedChunks, seedCount;
    DBuf<unsigned char> tmp;
    // lookup
    DBuf<unsigned> qCandOff, poolChunk;
    bool lookupAttr = false;
    DBuf<unsigned long long> candScratch, cursors;  // cursors: [0] candidate pool, [1] match pool, [2] pairs
    DBuf<int> qCandN;
    DBuf<unsigned short> poolDist;
    unsigned long long poolCap = 0, matchCap = 0;
    unsigned candCap = 0;
    // align
    DBuf<unsigned short> oAPos, oBPos, oAGapIndex, oLength, matchPool, matches;
    DBuf<int> oAGap, oBGap, oNode, nodePrev, hitLen;
    DBuf<unsigned> nodes;
    DBuf<unsigned long long> hitOff, qMatch, qMatchOff;
    DBuf<unsigned> qHits, qHitOff;
    DBuf<OvHit> hits;
    int nodeCap = 2048;
    // state of the last round (the getters read it)
    int S = 0, nSlices = 0, nQueries = 0;
    unsigned nChunks = 0;
    unsigned long long nReadSeeds = 0, nKeys = 0, nSeedPostings = 0;
    std::vector<OvSlice> hSlices;
    cudaEvent_t ev[10] = {nullptr};

    ~dp_overlapper() {
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        if (st) cudaStreamDestroy(st);
    }
};

namespace {

void ov_scan_u32(dp_overlapper& O, const unsigned* in, unsigned* out, long long n) {
    dp_exclusive_sum(in, out, n, O.scanTmp, O.st);
}
void ov_scan_u64(dp_overlapper& O, const unsigned long long* in, unsigned long long* out, long long n) {
    dp_exclusive_sum(in, out, n, O.scanTmp, O.st);
}

unsigned ov_fetch_err(dp_overlapper& O) {
    unsigned e = 0;
    CK(cudaMemcpyAsync(&e, O.err.p, sizeof(unsigned), cudaMemcpyDeviceToHost, O.st));
    CK(cudaStreamSynchronize(O.st));
    return e;
}

void ov_check_fatal(unsigned e) {
    if (e & 1u) throw std::runtime_error("overlap: a capacity of the seed selection was exceeded (slice longer than 8192 bases, "
                                         "more than 512 k-blocks in a slice, or more slices than 2 x query_batch_size)");
    if (e & 2u) throw std::runtime_error("overlap: a query slice holds more than 512 seeds (unsupported)");
    if (e & 16u) throw std::runtime_error("overlap: the reference panics on this input (seedAligner.reduced overflows, "
                                          "seeds/alignment.go:341-388)");
    if (e & 128u) throw std::runtime_error("overlap: the reference does not terminate on this input (chunkWorker walks back "
                                           "and forth over a seedless stretch, overlap/overlap.go:266-314)");
}

void ov_create(dp_overlapper& O, const uint8_t* bases, const int64_t* offsets, int64_t nReads) {
    cudaStream_t st = O.st;
    O.nReads = nReads;
    std::vector<long long> hBase((size_t)nReads), hWord((size_t)nReads);
    std::vector<int> hLen((size_t)nReads);
    long long w = 0, total = 0;
    for (int64_t i = 0; i < nReads; i++) {
        const long long len = offsets[i + 1] - offsets[i];
        if (len < 0 || len > 0x7fffffffll) throw std::runtime_error("bad read offsets");
        hWord[(size_t)i] = w;
        hBase[(size_t)i] = w * 16;
        hLen[(size_t)i] = (int)len;
        w += (len + 15) / 16 + 1;  // a zero word behind every read: k-mer loads one word past the end see zeros
        total += len;
    }
    O.totalBases = total;
    O.totalWords = w + 4;
    O.words.reserve((size_t)O.totalWords);
    CK(cudaMemsetAsync(O.words.p, 0, O.words.cap * sizeof(unsigned), st));
    O.readBase.reserve((size_t)nReads + 1);
    O.readLen.reserve((size_t)nReads + 1);
    CK(cudaMemcpyAsync(O.readBase.p, hBase.data(), (size_t)nReads * sizeof(long long), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(O.readLen.p, hLen.data(), (size_t)nReads * sizeof(int), cudaMemcpyHostToDevice, st));
    // pack in pieces of at most 1 GiB of ASCII (the staging buffer is transient)
    DBuf<unsigned char> dA;
    DBuf<long long> dOff, dWord;
    const long long pieceBytes = 1ll << 30;
    int64_t i0 = 0;
    while (i0 < nReads) {
        int64_t i1 = i0;
        while (i1 < nReads && (i1 == i0 || offsets[i1 + 1] - offsets[i0] <= pieceBytes)) i1++;
        const long long nb = offsets[i1] - offsets[i0];
        const int64_t cnt = i1 - i0;
        dA.reserve((size_t)nb + 64);
        std::vector<long long> so((size_t)cnt + 1);
        for (int64_t i = 0; i <= cnt; i++) so[(size_t)i] = offsets[i0 + i] - offsets[i0];
        dOff.reserve((size_t)cnt + 1);
        dWord.reserve((size_t)cnt);
        CK(cudaMemcpyAsync(dA.p, bases + offsets[i0], (size_t)nb, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dOff.p, so.data(), so.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dWord.p, hWord.data() + i0, (size_t)cnt * sizeof(long long), cudaMemcpyHostToDevice, st));
        const int blocks = (int)std::min<long long>((cnt + 7) / 8, (long long)O.smCount * 16);
        dp_pack_kernel<<<blocks, 256, 0, st>>>(dA.p, dOff.p, dWord.p, O.words.p, cnt);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        i0 = i1;
    }
}

// sequtil.KmerOccurrences over all reads (commands/overlap.go:43): one warp per read
__global__ void ov_kmer_hist_kernel(const unsigned* __restrict__ words, const long long* __restrict__ readBase,
                                    const int* __restrict__ readLen, long long nReads, int k,
                                    unsigned long long* __restrict__ counts) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nWarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    for (long long r = warp; r < nReads; r += nWarps) {
        const long long base = readBase[r];
        const int n = readLen[r] - k + 1;
        for (int p = (int)lane; p < n; p += 32) atomicAdd(counts + dp_kmer_at(words, base + p, k), 1ull);
    }
}

struct OvTimes {
    double select = 0, queries = 0, scan = 0, chunk = 0, index = 0, lookup = 0, align = 0, collect = 0;
};

void ov_round(dp_overlapper& O, const uint8_t* ignoreHost, long long firstSequence, dp_overlap_round* R) {
    cudaStream_t st = O.st;
    const OvParams& P = O.P;
    const int k = P.k;
    const long long nReads = O.nReads;
    const long long nTable = (1ll << (2 * k)) / 32;
    memset(R, 0, sizeof(*R));
    if (!O.haveValues) throw std::runtime_error("overlap: no k-mer values set (dp_overlapper_set_values)");
    if (firstSequence < 0) throw std::runtime_error("bad first_sequence");
    const double t0 = now_ms();
    auto mark = [&](int i) { CK(cudaEventRecord(O.ev[i], st)); };
    // ---- PrepareQueries: seed selection ----
    mark(0);
    O.ignore.reserve((size_t)nReads + 1);
    if (ignoreHost) CK(cudaMemcpyAsync(O.ignore.p, ignoreHost, (size_t)nReads, cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(O.ignore.p, 0, (size_t)nReads, st));
    O.bits.reserve((size_t)nTable);
    CK(cudaMemsetAsync(O.bits.p, 0, (size_t)nTable * sizeof(unsigned), st));
    O.err.reserve(1);
    O.cursors.reserve(8);
    CK(cudaMemsetAsync(O.err.p, 0, sizeof(unsigned), st));
    const int regCap = P.seedLimit + 4 * P.numSeeds + 64;
    const int sliceCap = 2 * P.queryBatch + 2;
    O.regKmer.reserve((size_t)regCap);
    O.slices.reserve((size_t)sliceCap);
    O.selOut.reserve(1);
    ov_select_kernel<<<1, OV_SEL_THREADS, 0, st>>>(O.words.p, O.readBase.p, O.readLen.p, O.ignore.p, (int)nReads, (int)firstSequence, P,
                                       O.values.p, O.bits.p, O.regKmer.p, regCap, O.slices.p, sliceCap, O.selOut.p, O.err.p);
    CK(cudaGetLastError());
    OvSelectOut so;
    CK(cudaMemcpyAsync(&so, O.selOut.p, sizeof(so), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ov_check_fatal(ov_fetch_err(O));
    O.S = so.size;
    O.nSlices = so.nSlices;
    O.nQueries = 2 * so.nSlices;
    O.nChunks = 0;
    R->num_seeds = so.size;
    R->num_queries = O.nQueries;
    R->num_query_seqs = so.nSlices;
    R->next_first_sequence = so.nSlices ? so.lastRead + 1 : firstSequence;
    O.hSlices.resize((size_t)so.nSlices);
    if (so.nSlices == 0) {
        R->ms_total = now_ms() - t0;
        return;  // commands/overlap.go:132-134: the command ends here
    }
    CK(cudaMemcpyAsync(O.hSlices.data(), O.slices.p, (size_t)so.nSlices * sizeof(OvSlice), cudaMemcpyDeviceToHost, st));
    const int S = so.size;
    // ---- {flags, rank} table, registration order <-> rank ----
    O.table.reserve((size_t)nTable);
    O.pc.reserve((size_t)nTable + 1);
    O.prefix.reserve((size_t)nTable + 1);
    dp_popc_kernel<<<div_up(nTable, 256), 256, 0, st>>>(O.bits.p, O.pc.p, nTable);
    CK(cudaMemsetAsync(O.pc.p + nTable, 0, sizeof(unsigned), st));
    ov_scan_u32(O, O.pc.p, O.prefix.p, nTable + 1);
    dp_table_kernel<<<div_up(nTable, 256), 256, 0, st>>>(O.bits.p, O.prefix.p, O.table.p, nTable);
    O.kmerOfRank.reserve((size_t)S + 1);
    O.regOfRank.reserve((size_t)S + 1);
    O.rankOfReg.reserve((size_t)S + 1);
    ov_rank_kernel<<<div_up(S, 256), 256, 0, st>>>(O.table.p, O.regKmer.p, S, O.kmerOfRank.p, O.regOfRank.p, O.rankOfReg.p);
    CK(cudaGetLastError());
    mark(1);
    // ---- queries: NewSeedSequence of every slice, and its reverse complement ----
    const int nSl = so.nSlices;
    O.descs.reserve((size_t)std::max<long long>(nSl, nReads) + 1);
    O.counts.reserve((size_t)std::max<long long>(nSl, nReads) + 2);
    O.fOff.reserve((size_t)nSl + 2);
    ov_slice_descs_kernel<<<div_up(nSl, 256), 256, 0, st>>>(O.slices.p, nSl, O.readBase.p, k, O.descs.p);
    CK(cudaMemsetAsync(O.counts.p, 0, ((size_t)nSl + 1) * sizeof(unsigned), st));
    int scanBlocks = std::min<int>(div_up(nSl, 8), O.smCount * 8);
    dp_chunk_scan_kernel<<<scanBlocks, 256, 0, st>>>(O.words.p, O.table.p, O.descs.p, (unsigned)nSl, k, 0, O.counts.p, nullptr,
                                                     nullptr, nullptr, nullptr);
    ov_scan_u32(O, O.counts.p, O.fOff.p, nSl + 1);
    unsigned F = 0;
    CK(cudaMemcpyAsync(&F, O.fOff.p + nSl, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    O.fPos.reserve((size_t)F + 1);
    O.fSeed.reserve((size_t)F + 1);
    dp_chunk_scan_kernel<<<scanBlocks, 256, 0, st>>>(O.words.p, O.table.p, O.descs.p, (unsigned)nSl, k, 1, nullptr, O.fOff.p,
                                                     O.fPos.p, O.fSeed.p, nullptr);
    const int nQ = 2 * nSl;
    O.qOff.reserve((size_t)nQ + 2);
    O.qPos.reserve((size_t)2 * F + 2);
    O.qSeed.reserve((size_t)2 * F + 2);
    O.qDistinct.reserve((size_t)2 * F + 2);
    O.qSlot.reserve((size_t)2 * F + 2);
    O.qND.reserve((size_t)nQ + 2);
    ov_build_queries_kernel<<<div_up((long long)nSl * 32, 128), 128, 0, st>>>(O.slices.p, nSl, k, O.table.p, O.kmerOfRank.p,
                                                                             O.rankOfReg.p, O.fOff.p, O.fPos.p, O.fSeed.p,
                                                                             O.qOff.p, O.qPos.p, O.qSeed.p, O.qDistinct.p,
                                                                             O.qSlot.p, O.qND.p, O.err.p);
    CK(cudaGetLastError());
    mark(2);
    // ---- AddSequences: every read's seed sequence ----
    {
        const int fBits = std::min(20, 2 * k);
        const bool exact = 2 * k <= 20;
        const unsigned* filt = O.bits.p;
        if (!exact) {
            const size_t fw = (((size_t)1 << fBits) + 31) / 32;
            O.filter.reserve(fw);
            CK(cudaMemsetAsync(O.filter.p, 0, fw * sizeof(unsigned), st));
            dp_filter_build_kernel<<<div_up(nTable, 256), 256, 0, st>>>(O.table.p, nTable, k, fBits, O.filter.p);
            filt = O.filter.p;
        }
        const size_t fWords = (((size_t)1 << fBits) + 31) / 32;
        const size_t smem = (((fWords + 3) & ~(size_t)3) + (size_t)32 * OV_SCAN_CACHE * 32) * sizeof(unsigned);
        if (!O.scanAttr) {
            CK(cudaFuncSetAttribute(ov_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            CK(cudaFuncSetAttribute(ov_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            O.scanAttr = true;
        }
        O.rStart.reserve((size_t)nReads + 1);
        O.rCount.reserve((size_t)nReads + 1);
        if (O.seedCap == 0) O.seedCap = (unsigned long long)(O.totalBases / 48) + (unsigned long long)nReads + 4096;
        for (;;) {
            if (O.seedCap >= 0xfffffff0ull) throw std::runtime_error("overlap: more than 2^32 seed occurrences in the read set");
            O.rPos.reserve((size_t)O.seedCap + 2);
            O.rSeed.reserve((size_t)O.seedCap + 2);
            CK(cudaMemsetAsync(O.cursors.p + 3, 0, sizeof(unsigned long long), st));
            const int grid = (int)std::min<long long>(O.smCount, (nReads + 31) / 32);
            if (exact)
                ov_scan_kernel<true><<<grid, 1024, smem, st>>>(O.words.p, O.readBase.p, O.readLen.p, O.ignore.p, (int)nReads, k,
                                                             O.table.p, filt, fBits, O.cursors.p + 3, O.seedCap, O.rStart.p,
                                                             O.rCount.p, O.rPos.p, O.rSeed.p);
            else
                ov_scan_kernel<false><<<grid, 1024, smem, st>>>(O.words.p, O.readBase.p, O.readLen.p, O.ignore.p, (int)nReads, k,
                                                              O.table.p, filt, fBits, O.cursors.p + 3, O.seedCap, O.rStart.p,
                                                              O.rCount.p, O.rPos.p, O.rSeed.p);
            CK(cudaGetLastError());
            unsigned long long used = 0;
            CK(cudaMemcpyAsync(&used, O.cursors.p + 3, sizeof(used), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            O.nReadSeeds = used;
            if (used <= O.seedCap) break;
            O.seedCap = used + used / 16 + 4096;
        }
    }
    const unsigned long long Rn = O.nReadSeeds;
    mark(3);
    // ---- chunkWorker ----
    O.pieceOff.reserve((size_t)nReads + 2);
    CK(cudaMemsetAsync(O.counts.p, 0, ((size_t)nReads + 1) * sizeof(unsigned), st));
    ov_chunk_kernel<<<div_up(nReads, 128), 128, 0, st>>>(O.rStart.p, O.rCount.p, O.rPos.p, O.readLen.p, O.ignore.p, (int)nReads, P, 0,
                                                         O.counts.p, nullptr, nullptr, O.err.p);
    ov_scan_u32(O, O.counts.p, O.pieceOff.p, nReads + 1);
    unsigned C = 0;
    CK(cudaMemcpyAsync(&C, O.pieceOff.p + nReads, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ov_check_fatal(ov_fetch_err(O));
    O.nChunks = C;
    R->num_chunks = C;
    R->read_seeds = (int64_t)Rn - nReads;
    if (C == 0) {
        R->ms_total = now_ms() - t0;
        return;
    }
    O.chunks.reserve((size_t)C + 1);
    ov_chunk_kernel<<<div_up(nReads, 128), 128, 0, st>>>(O.rStart.p, O.rCount.p, O.rPos.p, O.readLen.p, O.ignore.p, (int)nReads, P, 1, nullptr,
                                                         O.pieceOff.p, O.chunks.p, O.err.p);
    O.keyOff.reserve((size_t)C + 2);
    if (O.counts.cap < (size_t)C + 2) O.counts.reserve((size_t)C + 2);
    ov_chunk_n_kernel<<<div_up(C, 256), 256, 0, st>>>(O.chunks.p, C, O.counts.p);
    CK(cudaMemsetAsync(O.counts.p + C, 0, sizeof(unsigned), st));
    ov_scan_u32(O, O.counts.p, O.keyOff.p, (long long)C + 1);
    unsigned P2 = 0;
    CK(cudaMemcpyAsync(&P2, O.keyOff.p + C, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    O.nKeys = P2;
    mark(4);
    // ---- IndexSequences: seed -> distinct chunks ----
    O.keys.reserve((size_t)P2 + 1);
    O.keysSorted.reserve((size_t)P2 + 1);
    ov_keys_kernel<<<std::min<int>(div_up(C, 8), O.smCount * 16), 256, 0, st>>>(O.chunks.p, C, O.keyOff.p, O.rSeed.p, O.keys.p);
    unsigned long long P1 = 0;
    {
        // keys are written in chunk order, so a STABLE sort on the seed bits alone orders them by (seed, chunk)
        int endBit = 33;
        while ((1ull << (endBit - 32)) < (unsigned long long)S + 1 && endBit < 64) endBit++;
        if (endBit - 32 > 8) O.keysScratch.reserve((size_t)P2 + 1);
        dp_radix_sort(O.keys.p, O.keysSorted.p, O.keysScratch.p, nullptr, nullptr, nullptr, (long long)P2, 32, endBit, O.sortHist,
                      O.scanTmp, st);
        O.nSel.reserve(1);
        dp_unique_sorted(O.keysSorted.p, O.keys.p, O.nSel.p, (long long)P2, O.uniquePos, O.scanTmp, st);
        CK(cudaMemcpyAsync(&P1, O.nSel.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    O.nSeedPostings = P1;
    O.seedChunks.reserve((size_t)P1 + 4);
    O.seedOff.reserve((size_t)S + 2);
    if (P1 > 0)
        ov_seed_bounds_kernel<<<div_up((long long)P1, 256), 256, 0, st>>>(O.keys.p, P1, (unsigned)S, O.seedOff.p, O.seedChunks.p);
    else
        CK(cudaMemsetAsync(O.seedOff.p, 0, ((size_t)S + 2) * sizeof(unsigned), st));
    CK(cudaGetLastError());
    mark(5);
    DpIndexDev I{};
    I.k = k;
    I.numSeeds = (unsigned)S;
    I.numChunks = C;
    I.seedOff = O.seedOff.p;
    I.seedChunks = O.seedChunks.p;
    I.table = O.table.p;
    // ---- Matches: candidate chunks of every query ----
    const int lookupGrid = std::min(nQ, O.smCount);
    const size_t lookupSmem = (size_t)OV_TILE / 2 * sizeof(unsigned);
    if (!O.lookupAttr) {
        CK(cudaFuncSetAttribute(ov_lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lookupSmem));
        O.lookupAttr = true;
    }
    O.qCandOff.reserve((size_t)nQ + 1);
    O.qCandN.reserve((size_t)nQ + 1);
    if (O.candCap == 0) O.candCap = 1u << 14;
    if (O.poolCap == 0) O.poolCap = 1ull << 20;
    for (;;) {
        const unsigned candCap = std::min<unsigned>(O.candCap, C);
        O.candScratch.reserve((size_t)lookupGrid * candCap);
        O.poolChunk.reserve((size_t)O.poolCap);
        O.poolDist.reserve((size_t)O.poolCap);
        CK(cudaMemsetAsync(O.cursors.p, 0, 3 * sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(O.err.p, 0, sizeof(unsigned), st));
        CK(cudaMemsetAsync(O.cursors.p + 4, 0, sizeof(unsigned long long), st));
        ov_lookup_kernel<<<lookupGrid, OV_LTHREADS, lookupSmem, st>>>(I, O.qOff.p, O.qSeed.p, nQ, P.hitFraction, O.candScratch.p,
                                                                      candCap, O.qCandOff.p, O.qCandN.p, O.poolChunk.p,
                                                                      O.poolDist.p, O.cursors.p, O.poolCap, O.cursors.p + 4,
                                                                      O.err.p);
        CK(cudaGetLastError());
        const unsigned e = ov_fetch_err(O);
        ov_check_fatal(e);
        if (e & 4u) {
            O.candCap = std::min<unsigned long long>((unsigned long long)O.candCap * 4, 0x7fffffffu);
            continue;
        }
        if (e & 8u) {
            unsigned long long used = 0;
            CK(cudaMemcpy(&used, O.cursors.p, sizeof(used), cudaMemcpyDeviceToHost));
            O.poolCap = used + used / 8 + 1024;
            continue;
        }
        break;
    }
    unsigned long long nPool = 0, nPost = 0;
    CK(cudaMemcpy(&nPool, O.cursors.p, sizeof(nPool), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&nPost, O.cursors.p + 4, sizeof(nPost), cudaMemcpyDeviceToHost));
    R->candidates = (int64_t)nPool;
    R->posting_entries = (int64_t)nPost;
    mark(6);
    // ---- matchWorker ----
    const int W = 64;
    const int alignGrid = std::max(1, std::min(nQ, O.smCount * 4));
    const size_t threads = (size_t)alignGrid * W;
    O.oAPos.reserve(threads * OV_OPEN);
    O.oBPos.reserve(threads * OV_OPEN);
    O.oAGapIndex.reserve(threads * OV_OPEN);
    O.oLength.reserve(threads * OV_OPEN);
    O.oAGap.reserve(threads * OV_OPEN);
    O.oBGap.reserve(threads * OV_OPEN);
    O.oNode.reserve(threads * OV_OPEN);
    O.hitLen.reserve((size_t)nPool + 1);
    O.hitOff.reserve((size_t)nPool + 1);
    if (O.matchCap == 0) O.matchCap = 1ull << 22;
    for (;;) {
        O.nodes.reserve(threads * (size_t)O.nodeCap);
        O.nodePrev.reserve(threads * (size_t)O.nodeCap);
        O.matchPool.reserve((size_t)O.matchCap);
        OvAlignScratch A;
        A.oAPos = O.oAPos.p;
        A.oBPos = O.oBPos.p;
        A.oAGapIndex = O.oAGapIndex.p;
        A.oLength = O.oLength.p;
        A.oAGap = O.oAGap.p;
        A.oBGap = O.oBGap.p;
        A.oNode = O.oNode.p;
        A.nodes = O.nodes.p;
        A.nodePrev = O.nodePrev.p;
        A.nodeCap = O.nodeCap;
        CK(cudaMemsetAsync(O.cursors.p + 1, 0, 2 * sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(O.err.p, 0, sizeof(unsigned), st));
        ov_align_kernel<W><<<alignGrid, W, 0, st>>>(P, nQ, O.qOff.p, O.qPos.p, O.qSlot.p, O.qDistinct.p, O.qND.p, O.qCandOff.p,
                                                    O.qCandN.p, O.poolChunk.p, O.poolDist.p, O.chunks.p, O.rPos.p, O.rSeed.p, A,
                                                    O.hitLen.p, O.hitOff.p, O.matchPool.p, O.cursors.p + 1, O.matchCap,
                                                    O.cursors.p + 2, O.err.p);
        CK(cudaGetLastError());
        const unsigned e = ov_fetch_err(O);
        ov_check_fatal(e);
        if (e & 32u) {
            if (O.nodeCap >= (1 << 16)) throw std::runtime_error("overlap: chain node capacity exceeded");
            O.nodeCap *= 4;
            continue;
        }
        if (e & 64u) {
            unsigned long long used = 0;
            CK(cudaMemcpy(&used, O.cursors.p + 1, sizeof(used), cudaMemcpyDeviceToHost));
            O.matchCap = used + used / 8 + 1024;
            continue;
        }
        break;
    }
    mark(7);
    // ---- hits in delivery order ----
    O.qHits.reserve((size_t)nQ + 2);
    O.qHitOff.reserve((size_t)nQ + 2);
    O.qMatch.reserve((size_t)nQ + 2);
    O.qMatchOff.reserve((size_t)nQ + 2);
    CK(cudaMemsetAsync(O.qHits.p + nQ, 0, sizeof(unsigned), st));
    CK(cudaMemsetAsync(O.qMatch.p + nQ, 0, sizeof(unsigned long long), st));
    ov_hit_count_kernel<<<nQ, 128, 0, st>>>(nQ, O.qCandOff.p, O.qCandN.p, O.hitLen.p, O.qHits.p, O.qMatch.p);
    ov_scan_u32(O, O.qHits.p, O.qHitOff.p, nQ + 1);
    ov_scan_u64(O, O.qMatch.p, O.qMatchOff.p, nQ + 1);
    unsigned nHits = 0;
    unsigned long long nMatch = 0, pairs = 0;
    CK(cudaMemcpyAsync(&nHits, O.qHitOff.p + nQ, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&nMatch, O.qMatchOff.p + nQ, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&pairs, O.cursors.p + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    O.hits.reserve((size_t)nHits + 1);
    O.matches.reserve((size_t)nMatch + 1);
    ov_hit_write_kernel<<<nQ, 32, 0, st>>>(nQ, O.qCandOff.p, O.qCandN.p, O.poolChunk.p, O.hitLen.p, O.hitOff.p, O.matchPool.p,
                                           O.qHitOff.p, O.qMatchOff.p, O.hits.p, O.matches.p);
    CK(cudaGetLastError());
    R->num_hits = nHits;
    R->num_matches = (int64_t)nMatch;
    R->pairs = (int64_t)pairs;
    R->hits = (dp_overlap_hit*)malloc(((size_t)nHits + 1) * sizeof(dp_overlap_hit));
    R->matches = (uint16_t*)malloc(((size_t)nMatch + 1) * sizeof(uint16_t));
    if (!R->hits || !R->matches) {
        free(R->hits);
        free(R->matches);
        R->hits = nullptr;
        R->matches = nullptr;
        throw std::runtime_error("out of host memory");
    }
    static_assert(sizeof(dp_overlap_hit) == sizeof(OvHit), "hit record layout");
    if (nHits) CK(cudaMemcpyAsync(R->hits, O.hits.p, (size_t)nHits * sizeof(OvHit), cudaMemcpyDeviceToHost, st));
    if (nMatch) CK(cudaMemcpyAsync(R->matches, O.matches.p, (size_t)nMatch * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    mark(8);
    CK(cudaStreamSynchronize(st));
    auto el = [&](int a, int b) {
        float ms = 0;
        cudaEventElapsedTime(&ms, O.ev[a], O.ev[b]);
        return (double)ms;
    };
    R->ms_select = el(0, 1);
    R->ms_queries = el(1, 2);
    R->ms_scan = el(2, 3);
    R->ms_chunk = el(3, 4);
    R->ms_index = el(4, 5);
    R->ms_lookup = el(5, 6);
    R->ms_align = el(6, 7);
    R->ms_collect = el(7, 8);
    R->chunk_seeds = (int64_t)P2;
    R->seed_postings = (int64_t)P1;
    R->kernel_launches = 34;
    R->ms_total = now_ms() - t0;
}

}  // namespace

extern "C" {

int dp_overlapper_create(const uint8_t* bases_ascii, const int64_t* offsets, int64_t n_reads, int k,
                         const double* kmer_values, int overlap_size, int num_seeds, int seed_batch_size, int chunk_size,
                         int query_batch_size, double min_hits, int device, dp_overlapper** out) {
    API_TRY
    if (!bases_ascii || !offsets || !out || n_reads < 1) throw std::runtime_error("null or empty argument");
    if (k < 6 || k > 15) throw std::runtime_error("k must be in 6..15");
    if (n_reads > 0x7ffffff0ll) throw std::runtime_error("too many reads");
    if (overlap_size < 2 * k || overlap_size * 2 > OV_MAXSLICE) throw std::runtime_error("overlap_size must be in 2k..4096");
    if (overlap_size / 2 > 2 * OV_AMAX) throw std::runtime_error("overlap_size / 2 exceeds the reduced-query capacity");
    if (num_seeds < 1 || num_seeds > OV_MAXN) throw std::runtime_error("num_seeds must be in 1..256");
    if (seed_batch_size < 1 || seed_batch_size > (1 << 24)) throw std::runtime_error("bad seed_batch_size");
    if (chunk_size < 1 || chunk_size > 32000) throw std::runtime_error("chunk_size must be in 1..32000");
    if (query_batch_size < 1 || query_batch_size > (1 << 24)) throw std::runtime_error("bad query_batch_size");
    if (!(min_hits >= 0.0) || min_hits > 1.0) throw std::runtime_error("min_hits must be in 0..1");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev <= 0) throw std::runtime_error("no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= nDev) throw std::runtime_error("bad device");
    CK(cudaSetDevice(device));
    std::unique_ptr<dp_overlapper> O(new dp_overlapper());
    O->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    O->smCount = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&O->st, cudaStreamNonBlocking));
    for (auto& e : O->ev) CK(cudaEventCreate(&e));
    O->P.k = k;
    O->P.overlap = overlap_size;
    O->P.numSeeds = num_seeds;
    O->P.seedLimit = seed_batch_size;
    O->P.chunkSize = chunk_size;
    O->P.queryBatch = query_batch_size;
    O->P.redCap = overlap_size / 2;
    O->P.pad = 0;
    O->P.hitFraction = min_hits;
    ov_create(*O, bases_ascii, offsets, n_reads);
    if (kmer_values) {
        O->values.reserve((size_t)1 << (2 * k));
        CK(cudaMemcpy(O->values.p, kmer_values, sizeof(double) << (2 * k), cudaMemcpyHostToDevice));
        O->haveValues = true;
    }
    *out = O.release();
    API_CATCH
}

void dp_overlapper_destroy(dp_overlapper* o) {
    if (!o) return;
    cudaSetDevice(o->device);
    delete o;
}

int dp_overlapper_kmer_counts(dp_overlapper* o, uint64_t* counts) {
    API_TRY
    if (!o || !counts) throw std::runtime_error("null argument");
    CK(cudaSetDevice(o->device));
    const size_t nK = (size_t)1 << (2 * o->P.k);
    DBuf<unsigned long long> dC;
    dC.reserve(nK);
    CK(cudaMemcpyAsync(dC.p, counts, nK * sizeof(uint64_t), cudaMemcpyHostToDevice, o->st));
    ov_kmer_hist_kernel<<<o->smCount * 8, 256, 0, o->st>>>(o->words.p, o->readBase.p, o->readLen.p, o->nReads, o->P.k, dC.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(counts, dC.p, nK * sizeof(uint64_t), cudaMemcpyDeviceToHost, o->st));
    CK(cudaStreamSynchronize(o->st));
    API_CATCH
}

int dp_overlapper_set_values(dp_overlapper* o, const double* kmer_values) {
    API_TRY
    if (!o || !kmer_values) throw std::runtime_error("null argument");
    CK(cudaSetDevice(o->device));
    o->values.reserve((size_t)1 << (2 * o->P.k));
    CK(cudaMemcpy(o->values.p, kmer_values, sizeof(double) << (2 * o->P.k), cudaMemcpyHostToDevice));
    o->haveValues = true;
    API_CATCH
}

int dp_overlapper_round(dp_overlapper* o, const uint8_t* ignore, int64_t first_sequence, dp_overlap_round* out) {
    API_TRY
    if (!o || !out) throw std::runtime_error("null argument");
    CK(cudaSetDevice(o->device));
    ov_round(*o, ignore, first_sequence, out);
    API_CATCH
}

int dp_overlapper_seed_kmers(dp_overlapper* o, int64_t* kmers_out) {
    API_TRY
    if (!o || !kmers_out) throw std::runtime_error("null argument");
    CK(cudaSetDevice(o->device));
    std::vector<unsigned> h((size_t)o->S);
    if (o->S) CK(cudaMemcpy(h.data(), o->regKmer.p, (size_t)o->S * sizeof(unsigned), cudaMemcpyDeviceToHost));
    for (int i = 0; i < o->S; i++) kmers_out[i] = h[(size_t)i];
    API_CATCH
}

int dp_overlapper_queries(dp_overlapper* o, int64_t* meta, int64_t* seg_off, int64_t** segs) {
    API_TRY
    if (!o || !meta || !seg_off || !segs) throw std::runtime_error("null argument");
    CK(cudaSetDevice(o->device));
    const int nQ = o->nQueries;
    *segs = nullptr;
    std::vector<unsigned> qo((size_t)nQ + 1, 0);
    if (nQ) CK(cudaMemcpy(qo.data(), o->qOff.p, ((size_t)nQ + 1) * sizeof(unsigned), cudaMemcpyDeviceToHost));
    std::vector<int> hLen((size_t)o->nReads);
    CK(cudaMemcpy(hLen.data(), o->readLen.p, (size_t)o->nReads * sizeof(int), cudaMemcpyDeviceToHost));
    for (int q = 0; q < nQ; q++) {
        const OvSlice& s = o->hSlices[(size_t)(q >> 1)];
        const long long rlen = hLen[(size_t)s.read];
        meta[q * 6 + 0] = q >> 1;
        meta[q * 6 + 1] = s.read;
        meta[q * 6 + 2] = q & 1;
        meta[q * 6 + 3] = s.len;
        // the read is cached[id].SubSequence(0, Len()) (offset 0, inset 1: Q3); a slice of it is one more SubSequence
        const bool whole = s.start == 0 && s.len == rlen;
        meta[q * 6 + 4] = s.start;
        meta[q * 6 + 5] = whole ? 1 : 1 + rlen - (s.start + s.len - 1);
        seg_off[q] = 2ll * qo[(size_t)q] + q;
    }
    seg_off[nQ] = nQ ? 2ll * qo[(size_t)nQ] + nQ : 0;
    const long long total = seg_off[nQ];
    int64_t* h = (int64_t*)malloc((size_t)(total + 1) * sizeof(int64_t));
    if (!h) throw std::runtime_error("out of host memory");
    if (nQ) {
        DBuf<long long> d;
        d.reserve((size_t)total + 1);
        ov_export_query_segs_kernel<<<div_up((long long)nQ * 32, 128), 128, 0, o->st>>>(nQ, o->qOff.p, o->qPos.p, o->qSeed.p,
                                                                                       o->slices.p, o->regOfRank.p, o->P.k, d.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h, d.p, (size_t)total * sizeof(long long), cudaMemcpyDeviceToHost, o->st));
        CK(cudaStreamSynchronize(o->st));
    }
    *segs = h;
    API_CATCH
}

int dp_overlapper_chunks(dp_overlapper* o, const int32_t* ids, int64_t n, int64_t* meta, int64_t* seg_off, int64_t** segs) {
    API_TRY
    if (!o || !meta || !seg_off || !segs) throw std::runtime_error("null argument");
    CK(cudaSetDevice(o->device));
    *segs = nullptr;
    if (!ids) n = o->nChunks;
    std::vector<unsigned> hid((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        const long long id = ids ? ids[i] : i;
        if (id < 0 || id >= (long long)o->nChunks) throw std::runtime_error("chunk id out of range");
        hid[(size_t)i] = (unsigned)id;
    }
    DBuf<unsigned> dIds;
    DBuf<long long> dMeta, dSegOff, dSegs;
    dIds.reserve((size_t)n + 1);
    dMeta.reserve((size_t)n * 5 + 1);
    if (n) CK(cudaMemcpyAsync(dIds.p, hid.data(), (size_t)n * sizeof(unsigned), cudaMemcpyHostToDevice, o->st));
    if (n) {
        ov_export_chunk_meta_kernel<<<div_up(n, 256), 256, 0, o->st>>>(o->chunks.p, dIds.p, (int)n, dMeta.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(meta, dMeta.p, (size_t)n * 5 * sizeof(long long), cudaMemcpyDeviceToHost, o->st));
        CK(cudaStreamSynchronize(o->st));
    }
    long long total = 0;
    for (int64_t i = 0; i < n; i++) {
        seg_off[i] = total;
        total += 2 * meta[i * 5 + 4] + 1;
    }
    seg_off[n] = total;
    int64_t* h = (int64_t*)malloc((size_t)(total + 1) * sizeof(int64_t));
    if (!h) throw std::runtime_error("out of host memory");
    if (n) {
        dSegOff.reserve((size_t)n + 1);
        dSegs.reserve((size_t)total + 1);
        CK(cudaMemcpyAsync(dSegOff.p, seg_off, ((size_t)n + 1) * sizeof(long long), cudaMemcpyHostToDevice, o->st));
        ov_export_chunk_segs_kernel<<<div_up(n * 32, 128), 128, 0, o->st>>>(o->chunks.p, dIds.p, (int)n, o->rStart.p, o->rPos.p,
                                                                           o->rSeed.p, o->regOfRank.p, o->P.k, dSegOff.p, dSegs.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h, dSegs.p, (size_t)total * sizeof(long long), cudaMemcpyDeviceToHost, o->st));
        CK(cudaStreamSynchronize(o->st));
    }
    *segs = h;
    API_CATCH
}

}  // extern "C"
