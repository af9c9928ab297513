// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17) of the `downpore map` hot path of jteutenberg/downpore.
// Every function cites the reference file:line it follows (paths relative to the
// reference checkout).  The three amd64 assembler scan routines and the three
// soft-union routines are emulated at register level, control flow included,
// so that their observable quirks are reproduced
// and not "fixed".
//
// Round-2 groundwork for the `overlap` path also lives here and is used by tests only: AddSeeds,
// SeedSequence.ReverseComplement, chunkWorker's seed-space chunking (seeds.cpp) and seedAligner.PairwiseAlignments
// (alignment.cpp); each is cross-checked against a second restatement in tests/test_oracle_overlap_*.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, load or call anything in this directory.
//
// PARITY PINNING: the reference ships known-answer tests only for the
// sequence/ and util/ layers (sequence/sequence_test.go, util/bitset_test.go);
// those are restated in tests/test_oracle_kat.py and pin this file's L0/L1
// layers.  seeds/ and mapping/ have no tests, fixtures or golden output in the
// reference and no Go toolchain exists in this environment, so for those layers
// this oracle is "parity unpinned": a literal restatement reviewed by hand.
// Go's sort.Sort (unstable, version dependent) is replaced by a stable sort
// (identical to Go >= 1.19 for n <= 12, where pdqsort uses insertion sort).
#pragma once
#include <cstdint>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>

namespace dpo {

typedef long long gint;  // Go `int` on amd64

// ----------------------------------------------------------------------------
// sequence/sequence.go : packedSequence (4 bases per byte, MSB first)
// ----------------------------------------------------------------------------
struct PackedSeq {
    // Go slice `data []byte` = (store, off, nbytes).  Sub-sequences share the
    // parent's store, exactly like Go slices, so asm over-reads past the slice
    // end see the parent's following bytes.  The store carries 16 zero pad bytes.
    std::shared_ptr<std::vector<uint8_t>> store;
    size_t off = 0;
    size_t nbytes = 0;
    gint id = 0;
    gint offset = 0;
    gint inset = 0;
    std::shared_ptr<std::string> name;
    gint length = 0;
    gint firstLen = 4;
    gint finalLen = 0;

    const uint8_t* data() const { return store->data() + off; }
    gint Len() const { return length; }
};

uint8_t base_code(uint8_t b);                                  // sequence.go:59,80
void packBytes(const uint8_t* seq, size_t n, uint8_t* data);   // asm_amd64.s:33-78
PackedSeq NewPackedSequence(gint id, const std::string& seq, std::shared_ptr<std::string> name);  // sequence.go:67-93
PackedSeq SubSequence(const PackedSeq& s, gint start, gint end);      // sequence.go:353-370
PackedSeq ReverseComplement(const PackedSeq& s);                      // sequence.go:179-198
PackedSeq Append(const PackedSeq& s, gint id, const PackedSeq& other);  // sequence.go:164-177
std::string String(const PackedSeq& s);                               // sequence.go:242-276
gint KmerAt(const PackedSeq& s, gint index, gint k);                  // sequence.go:440-442 + asm :3-30
gint NextKmer(const PackedSeq& s, gint current, gint mask, gint nextBaseIndex);  // sequence.go:447-453
gint CountKmers(const PackedSeq& s, gint upTo, gint k, const uint8_t* seeds);                 // sequence.go:329-331
gint CountKmersBetween(const PackedSeq& s, gint from, gint to, gint upTo, gint k, const uint8_t* seeds);  // :332-337
void WriteSegments(const PackedSeq& s, gint* segments, gint k, const uint8_t* seeds);         // sequence.go:338-340
std::vector<uint16_t> ShortKmers(const PackedSeq& s, gint k, bool collapse);                  // sequence.go:482-504

// asm emulations (sequence/asm_amd64.s). `avail` = bytes readable from `data`
// before the store's end; loads beyond read zeros (the reference would read
// unrelated heap memory there: undefined).
gint packedKmerAt(const uint8_t* data, size_t avail, gint offset, gint k);                  // :3-30
gint packedCountKmers(const uint8_t* data, gint nbytes, size_t avail, gint upTo, gint skipFront, gint skipBack, gint k, const uint8_t* seeds);  // :81-203
void packedWriteSegments(const uint8_t* data, gint nbytes, size_t avail, gint skipFront, gint skipBack, gint k, const uint8_t* seeds, gint* segments);  // :206-394

// byteSequence (sequence.go:33-40, 278-324, 429-446): the simple one-base-per-byte
// implementation, kept as an independent check of the asm emulation.
struct ByteSeq {
    std::vector<uint8_t> data;
    gint offset = 0, inset = 0;
};
ByteSeq NewByteSequence(const std::string& seq);
ByteSeq SubSequence(const ByteSeq& s, gint start, gint end);
ByteSeq ReverseComplement(const ByteSeq& s);
std::string String(const ByteSeq& s);
gint KmerAt(const ByteSeq& s, gint index, gint k);
gint NextKmer(const ByteSeq& s, gint current, gint mask, gint nextBaseIndex);
gint CountKmers(const ByteSeq& s, gint upTo, gint k, gint mask, const uint8_t* kmers);
gint CountKmersBetween(const ByteSeq& s, gint from, gint to, gint upTo, gint k, gint mask, const uint8_t* kmers);
void WriteSegments(const ByteSeq& s, gint* segments, gint k, gint mask, const uint8_t* seeds);
std::vector<uint16_t> ShortKmers(const ByteSeq& s, gint k, bool collapse);
gint KmerValue(const std::string& s);  // sequence.go:520-528

// ----------------------------------------------------------------------------
// util/bitset.go + util/asm_amd64.s
// ----------------------------------------------------------------------------
struct IntSet {
    std::vector<uint64_t> vs;
    uint64_t start = 1, end = 0, count = 0;
};
IntSet NewIntSet();                       // bitset.go:20-23
IntSet NewIntSetCapacity(gint capacity);  // bitset.go:25-28
bool Contains(const IntSet& s, uint64_t x);  // :65-72
void Add(IntSet& s, uint64_t x);             // :74-108
void Clear(IntSet& s);                       // :145-153
uint64_t CountIntersection(const IntSet& a, const IntSet& b);                  // :163-177
uint64_t CountIntersectionTo(const IntSet& a, const IntSet& b, gint maxCount);  // :179-195
uint64_t countIntersectionToAsm(const uint64_t* a, const uint64_t* b, gint n, gint maxCount);  // asm :14-117
void getSoftUnion4Asm(const uint64_t* vs, gint n, uint64_t out[4]);   // asm :121-193
void getSoftUnion8Asm(const uint64_t* vs, gint n, uint64_t out[4]);   // asm :196-314
void getSoftUnion16Asm(const uint64_t* vs, gint n, uint64_t out[4]);  // asm :317-509
std::vector<uint64_t> GetSharedIDs(const std::vector<const IntSet*>& sets, gint minCount, bool fast);  // bitset.go:308-411
uint64_t CountMembers(IntSet& s);  // :584-591
std::vector<uint64_t> AsUints(const IntSet& s);

// ----------------------------------------------------------------------------
// seeds/sequence.go (hot part) and seeds/seeds.go
// ----------------------------------------------------------------------------
struct SeedSequence {
    std::vector<gint> segments;  // gap, seed, gap, seed, ..., gap
    gint id = 0;
    gint length = 0;
    gint offset = 0;
    gint inset = 0;
    bool rc = false;
    gint GetNumSeeds() const { return (gint)segments.size() / 2; }   // sequence.go:1388
    gint GetSeed(gint i) const { return segments[i * 2 + 1]; }       // :1282
    gint Len() const { return length; }                              // :1384
};
struct SeedMatch {
    std::vector<gint> MatchA, MatchB;
    const SeedSequence* SeqA = nullptr;
    const SeedSequence* SeqB = nullptr;
};
gint GetSeedOffset(const SeedSequence& s, gint index, gint k);         // sequence.go:1239-1246
gint GetSeedOffsetFromEnd(const SeedSequence& s, gint index, gint k);  // sequence.go:1269-1276
bool Reduced(const SeedSequence& s, const IntSet& whitelist, gint k, gint minSeeds,
             SeedSequence* reduced, std::vector<gint>* index);          // sequence.go:85-123
std::vector<SeedMatch> Match(const SeedSequence& seq, const SeedSequence& query, const IntSet& querySet,
                             const IntSet& seqSet, gint minMatch, gint k, bool* nil_result);  // :361-394
std::vector<SeedMatch> DynamicMatch(const SeedSequence& seq, const SeedSequence& query, gint minMatch, gint k);  // :401-471 (no Reduced)
void GetBasesCovered(const SeedMatch& m, gint k, gint* countA, gint* countB);  // sequence.go:830-858
uint64_t ReverseComplementKmer(uint64_t seed, uint64_t k);             // sequence.go:125-132

struct Counters {  // work counters (SURVEY.md 8d canonical accounting)
    long long windows = 0;          // performMapping calls
    long long kmer_lookups = 0;     // k-mers visited by WriteSegments on query windows (both strands)
    long long query_seeds = 0;      // seeds found in query window-strands
    long long posting_runs = 0;     // included seed occurrences (sets given to GetSharedIDs)
    long long posting_entries = 0;  // sum of |D(s)| over included occurrences
    long long candidates = 0;       // chunks returned by Matches
    long long cand_pass = 0;        // candidates passing CountIntersectionTo
    long long chain_cells = 0;      // |reduced chunk| + |reduced query| summed over Match calls reaching dynamicMatch
    long long chains = 0;           // chains returned
    long long mappings = 0;         // mappings returned by Map
    long long sort_ties_unpinned = 0;  // sorts with n>12 and tied keys (Go order unknown)
    void add(const Counters& o);
};

struct SeedIndex {
    gint seedSize = 0;
    std::vector<uint8_t> kmers;          // []bool
    std::vector<SeedSequence> sequences;
    std::vector<IntSet> sequenceSets;    // seed -> set of chunk ids
    std::vector<IntSet> seedSets;        // chunk -> set of seeds
    std::vector<int32_t> kmerMap;
    std::vector<gint> seedMap;
    gint size = 0;
    // Memory-lean mode (references where the two bitset families would need tens of GB: #seeds x #chunks / 8 bytes
    // each). The index then keeps every chunk's segments as int32 and, per seed, the ascending list of the chunks
    // containing it; `sequenceSets` / `seedSets` stay empty and `sequences[c].segments` are empty. The IntSet a
    // query needs is materialised on demand by replaying the very Add() sequence that AddSequence / IndexSequences
    // would have run (SeedChunkSet / ChunkSeedSet below), so every downstream routine (GetSharedIDs, the soft-union
    // emulations, CountIntersectionTo, Reduced, Match) runs unchanged on identical inputs.
    bool lean = false;
    std::vector<std::vector<int32_t>> leanSegments;  // [chunk] gap, seed, gap, ..., gap
    std::vector<uint64_t> leanSeedOff;               // [size+1]
    std::vector<uint32_t> leanSeedChunks;            // distinct chunk ids per seed, ascending
};
uint64_t SeedChunkCount(const SeedIndex& g, gint seed);     // sequenceSets[seed].count
IntSet SeedChunkSet(const SeedIndex& g, gint seed);         // lean: sequenceSets[seed] rebuilt (seeds.go:372-384 order)
IntSet ChunkSeedSet(const SeedIndex& g, size_t chunk);      // lean: seedSets[chunk] rebuilt (seeds.go:272-290 order)
SeedSequence ChunkSequence(const SeedIndex& g, size_t chunk);  // lean: sequences[chunk] with its segments restored
void NewSeedIndex(SeedIndex& g, gint k);                                          // seeds.go:23-31
SeedSequence NewSeedSequence(const SeedIndex& g, const PackedSeq& seq, Counters* c);  // seeds.go:33-50
void AddSingleSeeds(SeedIndex& g, const PackedSeq& seq, gint seedRate, const double* ranks);  // seeds.go:160-200
void AddSeeds(SeedIndex& g, const PackedSeq& seq, gint minSeeds, const double* kmerRanks, const uint8_t* quality);  // seeds.go:62-156 (overlap path)
SeedSequence ReverseComplementSeq(const SeedSequence& s, gint k, const SeedIndex& g);  // seeds/sequence.go:134-159 (overlap path)
SeedSequence SubSequenceSeeds(const SeedSequence& s, gint start, gint end, gint length, gint offset, gint inset);  // :46-50
std::vector<SeedSequence> ChunkSeedSequence(const SeedSequence& s, gint chunkSize, gint minSeeds, gint overlap, gint k);  // overlap/overlap.go:253-318
// seeds/alignment.go:274-616 (overlap path): pool of pair states addressed by index (-1 = nil)
struct PairState {
    gint aPos = 0, bPos = 0, aGap = 0, bGap = 0, aGapIndex = 0, length = 0;
    int prev = -1;
    gint stackIndex = 0;
};
struct SeedAligner {
    std::vector<PairState> stackPool;
    std::vector<int> statesStack;
    gint nextState = 0;
    std::vector<gint> reduced, aMapping;
    std::vector<int> open, initials, results;
};
SeedAligner NewSeedAligner(gint maxLength);                                       // alignment.go:298-306
gint gapRangeMin(gint gap, gint k);                                               // :411-424
gint gapRangeMax(gint gap, gint k);
std::vector<SeedMatch> PairwiseAlignments(SeedAligner& al, const SeedSequence& a, const SeedSequence& b, const IntSet& aSet,
                                          const IntSet& bSet, gint minMatches, gint k);  // :426-616
void AddSequence(SeedIndex& g, SeedSequence&& seq);                               // seeds.go:272-290
void IndexSequences(SeedIndex& g);                                                // seeds.go:292-305,372-384
std::vector<uint64_t> Matches(const SeedIndex& g, const SeedSequence& query, double hitFraction, Counters* c);  // :335-353

// ----------------------------------------------------------------------------
// overlap/overlap.go + commands/overlap.go:115-160 — one round of `downpore overlap` up to the seed-match stream
// (overlap.cpp; canonical order = num_workers 1)
// ----------------------------------------------------------------------------
struct OverlapParams {
    gint overlapSize = 1000, k = 10, numSeeds = 15, seedBatchSize = 10000, chunkSize = 10000, queryBatchSize = 20000;
    double hitFraction = 0.25;
};
struct OverlapQuery {  // overlap.go:10-16
    gint ID = 0, SequenceID = 0;
    SeedSequence Query;
    bool rc = false;
};
struct OverlapHit {  // the *seeds.SeedMatch matchWorker sends: SeqA = the query, SeqB = index.sequences[target]
    gint queryID = 0;
    bool rc = false;
    gint target = 0;
    std::vector<gint> MatchA, MatchB;
};
struct OverlapRound {
    SeedIndex index;
    std::vector<OverlapQuery> queries;
    std::vector<OverlapHit> hits;
    gint numQuerySeqs = 0, nextFirstSequence = 0;
};
void OverlapRoundRun(const std::vector<PackedSeq>& reads, const std::vector<uint8_t>& ignore, gint firstSequence,
                     const double* values, const OverlapParams& P, OverlapRound& out);

// ----------------------------------------------------------------------------
// util/sequtil/kmers.go + commands/map.go:45-71
// ----------------------------------------------------------------------------
void KmerOccurrences(const PackedSeq& seq, gint k, std::vector<uint64_t>& counts);  // kmers.go:53-69 (accumulates)
std::vector<gint> TopOccurrencesTop(std::vector<uint64_t>& counts, gint k, gint topN);  // kmers.go:87-112 (2nd return)
std::vector<double> KmerValues(std::vector<uint64_t>& kmerCounts, gint k);          // map.go:46-71

// ----------------------------------------------------------------------------
// mapping/mapping.go
// ----------------------------------------------------------------------------
struct Mapping {
    gint queryLen = 0;  // stands in for Mapping.Query (only .Len() is ever used on it)
    gint Start = 0, End = 0, QueryOffset = 0, QueryInset = 0;
    bool RC = false;
    gint ids = 0;
};
struct Mapper {
    SeedIndex index;
    PackedSeq reference;
    gint edgeSize = 0;
    bool circular = false;
    std::string refName;
};
// NewMapper (mapping.go:67-109). Chunk ids = producer emission order (canonical choice for Q5).
// lean: -1 = decide by the size of the bitsets (DPO_LEAN_BYTES, default 4 GB), 0 / 1 = forced; threads: workers for
// the per-chunk NewSeedSequence calls (independent; results are stored in emission order).
void NewMapper(Mapper& m, const PackedSeq& reference, bool circular, gint k, const double* kmerValues,
               gint seedRate, gint edgeSize, gint chunkSize, int lean = -1, int threads = 1);
// Map (mapping.go:430-487). Returned mappings are in the slice order the reference returns.
std::vector<Mapping> Map(const Mapper& m, const PackedSeq& query, Counters* c);
std::vector<Mapping> performMappingPublic(const Mapper& m, const PackedSeq& query, Counters* c);  // mapping.go:489-611
void PairEndsPublic(gint refLen, bool circular, gint queryLen, const std::vector<Mapping>& hitsA, const std::vector<Mapping>& hitsB,
                    std::vector<Mapping>* remA, std::vector<Mapping>* remB, std::vector<Mapping>* matched, bool* matchedNil);  // :167-203
std::string AsString(const Mapper& m, const Mapping& mp, const std::string& qname);  // mapping.go:112-122

// ----------------------------------------------------------------------------
// sequence/seqio.go:188-267 parsing rules (first pass over a file, no cache)
// ----------------------------------------------------------------------------
struct FastaRecord {
    std::string name;
    std::string seq;
};
std::vector<FastaRecord> ReadFasta(const std::string& filename, gint minLength);
std::vector<FastaRecord> ParseFasta(const std::string& content, gint minLength);

}  // namespace dpo
