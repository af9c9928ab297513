// downpore_b200 — shared declarations for the sm_100a kernels and the host driver.
//
// Data layout in HBM:
//   packed sequences : uint32 words, 16 bases per word, first base in the two most significant bits (the same bit
//                      order as the reference's packedSequence bytes read big-endian, sequence/sequence.go:42-53).
//   k-mer table      : uint2[4^k/32] = {32 seed flags, number of seeds before this word}; a seed's id is its rank
//                      among seed k-mers (seed ids are arbitrary labels in the reference: seeds/seeds.go:189).
//   seed -> chunks   : CSR (seedOff[S+1], seedChunks[]) of the DISTINCT chunk ids containing each seed, ascending
//                      (what SeedIndex.sequenceSets holds as bitsets, seeds/seeds.go:15,372-384).
//   seed -> postings : CSR (postOff[S+1], postChunk[], postPos[]) of EVERY occurrence of each seed, sorted by
//                      (chunk, scan position): what the chainer gathers for the query's seeds in a candidate chunk.
//   chunk -> seeds   : CSR (chunkOff[C+1], chunkPos[], chunkSeed[]) of every seed occurrence of each chunk in scan
//                      order (what SeedIndex.sequences[c].segments holds as gaps, seeds/seeds.go:33-50).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DP_WARP 32
#define DP_FULL 0xffffffffu

struct DpIndexDev {
    int k;
    int circular;
    int edge;                // query_size
    int maxWindow;           // longest window a query can have (2*edge)
    long long refLen;
    unsigned numSeeds;       // S
    unsigned numChunks;      // C
    unsigned maxChunkSeeds;  // longest chunk list
    const uint2* table;      // [4^k/32]
    const unsigned* seedOff;     // [S+1]
    const unsigned* seedChunks;  // [seedOff[S]]
    const unsigned* postOff;     // [S+1]  every occurrence of each seed, sorted by (chunk, scan position)
    const unsigned* postChunk;   // [postOff[S]]
    const int* postPos;          // [postOff[S]]
    const unsigned* filter;      // 2^filterBits bits: bit p = "some seed k-mer has the filterBits-bit prefix p" (0 = none)
    int filterBits;
    const unsigned* chunkOff;    // [C+1]
    const int* chunkPos;         // scan positions
    const unsigned* chunkSeed;   // seed ranks
    const long long* chunkOffset;  // SeedSequence.offset of the chunk
    const long long* chunkInset;   // SeedSequence.inset of the chunk (Q3: one too large for SubSequence chunks)
    const int* chunkScanLen;       // sum of the chunk's segments (bases actually scanned, Q2)
    // derived copy of the seed -> chunks runs for dp_lookup_mid_kernel (indexes of at most 16384 chunks; else null):
    // every run starts on a 16-byte block and is padded to whole blocks; a posting is stored ready to count,
    // (byte offset of the chunk's counter word) << 16 | (PRMT selector that puts a 1 into the chunk's byte of that word)
    const uint4* midSeed;  // [S] {first posting in seedChunks, run length, first 16-byte block in midPost, last chunk >> 6}
    const uint4* midPost;
};

// One performMapping() call (mapping/mapping.go:489-611): window [start, start+len) of read `read`.
struct DpWindow {
    int read;
    int start;
    int len;
    int whole;  // 1: the window is the un-sliced read (len <= 2*edge): raw packedSequence quirks apply (Q2, Q3)
};

// dp_mapping twin on the device (32 bytes)
struct DpMappingDev {
    long long start;
    long long end;
    int qOffset;
    int qInset;
    int ids;
    int rc;  // low byte = RC flag; the chain length rides in the upper bytes until the window is finalised
};

struct DpCounters {
    unsigned long long kmer_lookups;
    unsigned long long query_seeds;
    unsigned long long posting_runs;
    unsigned long long posting_entries;
    unsigned long long candidates;
    unsigned long long chain_cells;
    unsigned long long mappings;
    unsigned int overflow;  // DP_OV_* bits: a capacity of this launch was too small; the host reruns it with more room
    unsigned int pad;
};
// (none of these is a limit of the library: the host grows the flagged capacity and recomputes, see dp_api.cu)
#define DP_OV_RESULTS 1u  // mappings of one window before sort/dedupe (general chain kernel)
#define DP_OV_CHAINS 2u   // chains kept for one candidate (general chain kernel)
#define DP_OV_CANDS 4u    // candidate chunks of one window strand
#define DP_OV_OUTPOOL 8u  // the launch's pool of window results

__host__ __device__ inline unsigned dp_base_code(unsigned b) { return ((b >> 1) ^ ((b & 4) >> 2)) & 3; }

#ifdef __CUDACC__
__device__ __forceinline__ unsigned dp_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned dp_lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// k-mer starting at base p of a packed sequence (k <= 15): 16 bases from p in a 32-bit register, top-aligned
__device__ __forceinline__ unsigned dp_kmer_at(const unsigned* __restrict__ words, long long p, int k) {
    long long w = p >> 4;
    unsigned sh = ((unsigned)p & 15u) * 2u;
    unsigned hi = __ldg(words + w);
    unsigned lo = __ldg(words + w + 1);
    unsigned v = __funnelshift_l(lo, hi, sh);
    return v >> (32 - 2 * k);
}

// reverse complement of a k-mer held in the low 2k bits
__device__ __forceinline__ unsigned dp_revcomp(unsigned kmer, int k) {
    unsigned y = __brev(~kmer);
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    return y >> (32 - 2 * k);
}

// flag + rank in one 8-byte gather
__device__ __forceinline__ bool dp_seed_lookup(const uint2* __restrict__ table, unsigned kmer, unsigned* rank) {
    uint2 e = __ldg(table + (kmer >> 5));
    unsigned bit = 1u << (kmer & 31);
    *rank = e.y + __popc(e.x & (bit - 1));
    return (e.x & bit) != 0;
}
// prefix filter: a k-mer (low 2k bits) maps to its top `bits` bits (bits <= 2k)
__device__ __forceinline__ unsigned dp_filter_hash(unsigned kmer, int k, int bits) { return kmer >> (2 * k - bits); }
__device__ __forceinline__ bool dp_seed_flag(const uint2* __restrict__ table, unsigned kmer) {
    return (__ldg(&table[kmer >> 5].x) >> (kmer & 31)) & 1u;
}
#endif
