// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Plain C entry points so tests/ and bench.py's cpu_baseline leg can drive the oracle through ctypes.
#include <atomic>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "oracle.hpp"

using namespace dpo;

static thread_local std::string g_err;
#define DPO_TRY try {
#define DPO_CATCH(ret)                 \
    }                                  \
    catch (const std::exception& ex) { \
        g_err = ex.what();             \
        return ret;                    \
    }

extern "C" {

const char* dpo_last_error() { return g_err.c_str(); }

// ----- packed / byte sequences (KAT layer) ----------------------------------
void* dpo_packed_new(const char* ascii, long long n) {
    DPO_TRY return new PackedSeq(NewPackedSequence(0, std::string(ascii, (size_t)n), nullptr));
    DPO_CATCH(nullptr)
}
void dpo_packed_free(void* p) { delete (PackedSeq*)p; }
void* dpo_packed_sub(void* p, long long start, long long end) {
    DPO_TRY return new PackedSeq(SubSequence(*(PackedSeq*)p, start, end));
    DPO_CATCH(nullptr)
}
void* dpo_packed_rc(void* p) {
    DPO_TRY return new PackedSeq(ReverseComplement(*(PackedSeq*)p));
    DPO_CATCH(nullptr)
}
long long dpo_packed_len(void* p) { return ((PackedSeq*)p)->length; }
long long dpo_packed_nbytes(void* p) { return (long long)((PackedSeq*)p)->nbytes; }
void dpo_packed_bytes(void* p, unsigned char* out) { memcpy(out, ((PackedSeq*)p)->data(), ((PackedSeq*)p)->nbytes); }
void dpo_packed_fields(void* p, long long* out5) {
    PackedSeq* s = (PackedSeq*)p;
    out5[0] = s->offset;
    out5[1] = s->inset;
    out5[2] = s->firstLen;
    out5[3] = s->finalLen;
    out5[4] = s->length;
}
void dpo_packed_string(void* p, char* out) {
    std::string s = String(*(PackedSeq*)p);
    memcpy(out, s.data(), s.size());
}
long long dpo_packed_kmer_at(void* p, long long index, long long k) { return KmerAt(*(PackedSeq*)p, index, k); }
long long dpo_packed_next_kmer(void* p, long long cur, long long mask, long long idx) {
    return NextKmer(*(PackedSeq*)p, cur, mask, idx);
}
long long dpo_packed_count_kmers(void* p, long long upTo, long long k, const unsigned char* seeds) {
    return CountKmers(*(PackedSeq*)p, upTo, k, seeds);
}
long long dpo_packed_count_kmers_between(void* p, long long from, long long to, long long upTo, long long k,
                                         const unsigned char* seeds) {
    DPO_TRY return CountKmersBetween(*(PackedSeq*)p, from, to, upTo, k, seeds);
    DPO_CATCH(-1)
}
// segments must hold 2*(len+16)+1 entries; returns the number of entries written (2*hits+1)
long long dpo_packed_write_segments(void* p, long long k, const unsigned char* seeds, long long* segments) {
    PackedSeq* s = (PackedSeq*)p;
    size_t cap = (size_t)(2 * (s->length + 16) + 1);
    for (size_t i = 0; i < cap; i++) segments[i] = INT64_MIN;
    WriteSegments(*s, segments, k, seeds);
    size_t last = cap;
    while (last > 0 && segments[last - 1] == INT64_MIN) last--;
    return (long long)last;
}
long long dpo_packed_short_kmers(void* p, long long k, int collapse, unsigned short* out) {
    std::vector<uint16_t> v = ShortKmers(*(PackedSeq*)p, k, collapse != 0);
    memcpy(out, v.data(), v.size() * 2);
    return (long long)v.size();
}

void* dpo_byte_new(const char* ascii, long long n) { return new ByteSeq(NewByteSequence(std::string(ascii, (size_t)n))); }
void dpo_byte_free(void* p) { delete (ByteSeq*)p; }
void* dpo_byte_sub(void* p, long long start, long long end) { return new ByteSeq(SubSequence(*(ByteSeq*)p, start, end)); }
void* dpo_byte_rc(void* p) { return new ByteSeq(ReverseComplement(*(ByteSeq*)p)); }
long long dpo_byte_len(void* p) { return (long long)((ByteSeq*)p)->data.size(); }
void dpo_byte_fields(void* p, long long* out2) {
    out2[0] = ((ByteSeq*)p)->offset;
    out2[1] = ((ByteSeq*)p)->inset;
}
void dpo_byte_string(void* p, char* out) {
    std::string s = String(*(ByteSeq*)p);
    memcpy(out, s.data(), s.size());
}
long long dpo_byte_kmer_at(void* p, long long index, long long k) { return KmerAt(*(ByteSeq*)p, index, k); }
long long dpo_byte_next_kmer(void* p, long long cur, long long mask, long long idx) {
    return NextKmer(*(ByteSeq*)p, cur, mask, idx);
}
long long dpo_byte_count_kmers(void* p, long long upTo, long long k, long long mask, const unsigned char* seeds) {
    return CountKmers(*(ByteSeq*)p, upTo, k, mask, seeds);
}
long long dpo_byte_count_kmers_between(void* p, long long from, long long to, long long upTo, long long k,
                                       long long mask, const unsigned char* seeds) {
    return CountKmersBetween(*(ByteSeq*)p, from, to, upTo, k, mask, seeds);
}
long long dpo_byte_write_segments(void* p, long long k, long long mask, const unsigned char* seeds, long long* segments) {
    ByteSeq* s = (ByteSeq*)p;
    size_t cap = (size_t)(2 * ((long long)s->data.size() + 16) + 1);
    for (size_t i = 0; i < cap; i++) segments[i] = INT64_MIN;
    WriteSegments(*s, segments, k, mask, seeds);
    size_t last = cap;
    while (last > 0 && segments[last - 1] == INT64_MIN) last--;
    return (long long)last;
}
long long dpo_byte_short_kmers(void* p, long long k, int collapse, unsigned short* out) {
    std::vector<uint16_t> v = ShortKmers(*(ByteSeq*)p, k, collapse != 0);
    memcpy(out, v.data(), v.size() * 2);
    return (long long)v.size();
}
long long dpo_kmer_value(const char* s, long long n) { return KmerValue(std::string(s, (size_t)n)); }
void dpo_pack_bytes(const unsigned char* s, long long n, unsigned char* data) { packBytes(s, (size_t)n, data); }

// ----- bitsets ----------------------------------------------------------------
void* dpo_intset_new() { return new IntSet(NewIntSet()); }
void dpo_intset_free(void* p) { delete (IntSet*)p; }
void dpo_intset_add(void* p, unsigned long long x) { Add(*(IntSet*)p, x); }
int dpo_intset_contains(void* p, unsigned long long x) { return Contains(*(IntSet*)p, x) ? 1 : 0; }
unsigned long long dpo_intset_size(void* p) { return ((IntSet*)p)->count; }
unsigned long long dpo_intset_count_intersection(void* a, void* b) { return CountIntersection(*(IntSet*)a, *(IntSet*)b); }
long long dpo_intset_count_intersection_to(void* a, void* b, long long maxCount) {
    DPO_TRY return (long long)CountIntersectionTo(*(IntSet*)a, *(IntSet*)b, maxCount);
    DPO_CATCH(-1)
}
long long dpo_get_shared_ids(void** sets, long long n, long long minCount, int fast, unsigned long long* out,
                             long long cap) {
    DPO_TRY std::vector<const IntSet*> v;
    for (long long i = 0; i < n; i++) v.push_back((const IntSet*)sets[i]);
    std::vector<uint64_t> ids = GetSharedIDs(v, minCount, fast != 0);
    for (size_t i = 0; i < ids.size() && (long long)i < cap; i++) out[i] = ids[i];
    return (long long)ids.size();
    DPO_CATCH(-1)
}

// ----- overlap path groundwork: a bare SeedIndex and AddSeeds (seeds.go:62-156) ---------------------------------
void* dpo_seedindex_new(int k) {
    DPO_TRY SeedIndex* g = new SeedIndex();
    NewSeedIndex(*g, k);
    return g;
    DPO_CATCH(nullptr)
}
void dpo_seedindex_free(void* h) { delete (SeedIndex*)h; }
int dpo_seedindex_add_seeds(void* h, const char* ascii, long long n, long long minSeeds, const double* ranks,
                            const unsigned char* quality) {
    DPO_TRY PackedSeq seq = NewPackedSequence(0, std::string(ascii, (size_t)n), nullptr);
    AddSeeds(*(SeedIndex*)h, seq, minSeeds, ranks, quality);
    return 0;
    DPO_CATCH(-1)
}
// seed k-mers in registration order (seedMap)
long long dpo_seedindex_seeds(void* h, long long* out, long long cap) {
    const SeedIndex& g = *(const SeedIndex*)h;
    for (long long i = 0; i < g.size && i < cap; i++) out[i] = g.seedMap[(size_t)i];
    return g.size;
}

// NewSeedSequence (seeds.go:33-50) of a sequence against the bare index: segments -> out (returns their number, or the
// needed size if cap is too small); fields = {length, offset, inset}
long long dpo_seedindex_seed_sequence(void* h, const char* ascii, long long n, long long* out, long long cap, long long* fields) {
    DPO_TRY PackedSeq seq = NewPackedSequence(0, std::string(ascii, (size_t)n), nullptr);
    SeedSequence s = NewSeedSequence(*(SeedIndex*)h, seq, nullptr);
    for (size_t i = 0; i < s.segments.size() && (long long)i < cap; i++) out[i] = s.segments[i];
    if (fields) {
        fields[0] = s.length;
        fields[1] = s.offset;
        fields[2] = s.inset;
    }
    return (long long)s.segments.size();
    DPO_CATCH(-1)
}
// SeedSequence.ReverseComplement (seeds/sequence.go:134-159) on raw segments
long long dpo_seedindex_rc(void* h, const long long* segments, long long nseg, long long* out) {
    DPO_TRY SeedSequence s;
    s.segments.assign(segments, segments + nseg);
    const SeedIndex& g = *(const SeedIndex*)h;
    SeedSequence r = ReverseComplementSeq(s, g.seedSize, g);
    for (long long i = 0; i < nseg; i++) out[i] = r.segments[(size_t)i];
    return nseg;
    DPO_CATCH(-1)
}
// chunkWorker (overlap/overlap.go:253-318) for one seed sequence. out: per piece {nSegments, length, offset, inset,
// segments...}; returns the number of values written (or needed), pieces in *nPieces
long long dpo_chunk_seed_sequence(const long long* segments, long long nseg, long long length, long long chunkSize,
                                  long long minSeeds, long long overlap, long long k, long long* out, long long cap,
                                  long long* nPieces) {
    DPO_TRY SeedSequence s;
    s.segments.assign(segments, segments + nseg);
    s.length = length;
    std::vector<SeedSequence> pieces = ChunkSeedSequence(s, chunkSize, minSeeds, overlap, k);
    long long w = 0;
    for (const SeedSequence& p : pieces) {
        const long long vals[4] = {(long long)p.segments.size(), p.length, p.offset, p.inset};
        for (int i = 0; i < 4; i++, w++)
            if (w < cap) out[w] = vals[i];
        for (gint v : p.segments) {
            if (w < cap) out[w] = v;
            w++;
        }
    }
    if (nPieces) *nPieces = (long long)pieces.size();
    return w;
    DPO_CATCH(-1)
}

// seedAligner.PairwiseAlignments (seeds/alignment.go:426-616) as matchWorker calls it (overlap/overlap.go:346-363):
// aSet / bSet = the seeds of a / b; a fresh aligner of NewSeedAligner(maxLength). out: per match {length, MatchA...,
// MatchB...} in the order returned; returns values written (or needed), matches in *nMatches. -1 + error text when the
// reference would panic.
long long dpo_pairwise_alignments(const long long* aSeg, long long na, const long long* bSeg, long long nb,
                                  long long minMatches, long long k, long long maxLength, long long* out, long long cap,
                                  long long* nMatches) {
    DPO_TRY SeedSequence a, b;
    a.segments.assign(aSeg, aSeg + na);
    b.segments.assign(bSeg, bSeg + nb);
    IntSet aSet = NewIntSet(), bSet = NewIntSet();
    for (long long i = 1; i < na; i += 2) Add(aSet, (uint64_t)aSeg[i]);
    for (long long i = 1; i < nb; i += 2) Add(bSet, (uint64_t)bSeg[i]);
    SeedAligner al = NewSeedAligner(maxLength);
    std::vector<SeedMatch> ms = PairwiseAlignments(al, a, b, aSet, bSet, minMatches, k);
    long long w = 0;
    for (const SeedMatch& m : ms) {
        if (w < cap) out[w] = (long long)m.MatchA.size();
        w++;
        for (gint v : m.MatchA) {
            if (w < cap) out[w] = v;
            w++;
        }
        for (gint v : m.MatchB) {
            if (w < cap) out[w] = v;
            w++;
        }
    }
    if (nMatches) *nMatches = (long long)ms.size();
    return w;
    DPO_CATCH(-1)
}

// SeedSequence.Match (seeds/sequence.go:361-394) on explicit segment lists. mode 1: as performMapping calls it
// (mapping/mapping.go:518-552): querySet / seqSet = the seeds of the query / of seq, both sequences Reduced first;
// mode 0: dynamicMatch on the sequences as they are. out: per match {length, MatchA..., MatchB..., coveredA, coveredB}
// (GetBasesCovered, :830-858); returns values written (or needed), matches in *nMatches (-1: Match returned nil).
long long dpo_match(const long long* seqSeg, long long ns, const long long* querySeg, long long nq, long long minMatch,
                    long long k, int mode, long long* out, long long cap, long long* nMatches) {
    DPO_TRY SeedSequence seq, query;
    seq.segments.assign(seqSeg, seqSeg + ns);
    query.segments.assign(querySeg, querySeg + nq);
    std::vector<SeedMatch> ms;
    bool nil = false;
    if (mode == 1) {
        IntSet qSet = NewIntSet(), sSet = NewIntSet();
        for (long long i = 1; i < nq; i += 2) Add(qSet, (uint64_t)querySeg[i]);
        for (long long i = 1; i < ns; i += 2) Add(sSet, (uint64_t)seqSeg[i]);
        ms = Match(seq, query, qSet, sSet, minMatch, k, &nil);
    } else {
        ms = DynamicMatch(seq, query, minMatch, k);
    }
    long long w = 0;
    for (const SeedMatch& m : ms) {
        gint ca = 0, cb = 0;
        GetBasesCovered(m, k, &ca, &cb);
        if (w < cap) out[w] = (long long)m.MatchA.size();
        w++;
        for (gint v : m.MatchA) {
            if (w < cap) out[w] = v;
            w++;
        }
        for (gint v : m.MatchB) {
            if (w < cap) out[w] = v;
            w++;
        }
        if (w < cap) out[w] = ca;
        w++;
        if (w < cap) out[w] = cb;
        w++;
    }
    if (nMatches) *nMatches = nil ? -1 : (long long)ms.size();
    return w;
    DPO_CATCH(-1)
}

// SeedSequence.Reduced (seeds/sequence.go:85-123) with makeIndex: out = reduced segments, index = original positions;
// returns the number of segment values (-1: nil, fewer than minSeeds whitelisted seeds)
long long dpo_reduced(const long long* seg, long long n, const long long* whitelist, long long nw, long long k,
                      long long minSeeds, long long* out, long long* index) {
    DPO_TRY SeedSequence s, r;
    s.segments.assign(seg, seg + n);
    IntSet wl = NewIntSet();
    for (long long i = 0; i < nw; i++) Add(wl, (uint64_t)whitelist[i]);
    std::vector<gint> idx;
    if (!Reduced(s, wl, k, minSeeds, &r, &idx)) return -1;
    for (size_t i = 0; i < r.segments.size(); i++) out[i] = r.segments[i];
    for (size_t i = 0; i < idx.size(); i++) index[i] = idx[i];
    return (long long)r.segments.size();
    DPO_CATCH(-2)
}

// GetSeedOffset / GetSeedOffsetFromEnd (seeds/sequence.go:1239-1246, 1269-1276)
long long dpo_seed_offset(const long long* seg, long long n, long long index, long long k, int from_end) {
    SeedSequence s;
    s.segments.assign(seg, seg + n);
    return from_end ? GetSeedOffsetFromEnd(s, index, k) : GetSeedOffset(s, index, k);
}

// mapEnds' pairing step (mapping/mapping.go:167-203: removeDominated, matchPairs with isConsistent) on explicit hits.
// hits: rows of 6 {Start, End, QueryOffset, QueryInset, RC, ids}. out: remainingA rows, remainingB rows, matched rows;
// counts3 = their row counts (matched count -1 when matched == nil).
int dpo_pair_ends(long long refLen, int circular, long long queryLen, const long long* hitsA, long long nA,
                  const long long* hitsB, long long nB, long long* out, long long* counts3) {
    DPO_TRY auto load = [](const long long* h, long long n) {
        std::vector<Mapping> v((size_t)n);
        for (long long i = 0; i < n; i++) {
            v[(size_t)i].Start = h[6 * i];
            v[(size_t)i].End = h[6 * i + 1];
            v[(size_t)i].QueryOffset = h[6 * i + 2];
            v[(size_t)i].QueryInset = h[6 * i + 3];
            v[(size_t)i].RC = h[6 * i + 4] != 0;
            v[(size_t)i].ids = h[6 * i + 5];
        }
        return v;
    };
    std::vector<Mapping> ra, rb, mt;
    bool nil = true;
    PairEndsPublic(refLen, circular != 0, queryLen, load(hitsA, nA), load(hitsB, nB), &ra, &rb, &mt, &nil);
    long long w = 0;
    for (const std::vector<Mapping>* v : {&ra, &rb, &mt})
        for (const Mapping& mp : *v) {
            out[w++] = mp.Start;
            out[w++] = mp.End;
            out[w++] = mp.QueryOffset;
            out[w++] = mp.QueryInset;
            out[w++] = mp.RC ? 1 : 0;
            out[w++] = mp.ids;
        }
    counts3[0] = (long long)ra.size();
    counts3[1] = (long long)rb.size();
    counts3[2] = nil ? -1 : (long long)mt.size();
    return 0;
    DPO_CATCH(1)
}

// SeedIndex.Matches (seeds/seeds.go:335-353) against an index of explicit seed sequences: chunkSegs = the chunks' segment
// lists back to back, chunkOff[nChunks+1] their bounds; AddSequence + IndexSequences (seeds.go:272-305, 372-384) over
// numSeeds seed ids, then Matches(query, hitFraction). Returns the number of chunk ids written to out (ascending).
long long dpo_matches(const long long* chunkSegs, const long long* chunkOff, long long nChunks, long long numSeeds,
                      const long long* querySeg, long long nq, double hitFraction, long long* out, long long cap) {
    DPO_TRY SeedIndex g;
    NewSeedIndex(g, 5);
    g.size = numSeeds;
    g.sequenceSets.assign((size_t)numSeeds, NewIntSet());
    for (long long c = 0; c < nChunks; c++) {
        SeedSequence s;
        s.segments.assign(chunkSegs + chunkOff[c], chunkSegs + chunkOff[c + 1]);
        AddSequence(g, std::move(s));
    }
    IndexSequences(g);
    SeedSequence q;
    q.segments.assign(querySeg, querySeg + nq);
    std::vector<uint64_t> ids = Matches(g, q, hitFraction, nullptr);
    for (size_t i = 0; i < ids.size() && (long long)i < cap; i++) out[i] = (long long)ids[i];
    return (long long)ids.size();
    DPO_CATCH(-1)
}

// gapRange (seeds/alignment.go:411-424): out2 = {minGap, maxGap}
void dpo_gap_range(long long gap, long long k, long long* out2) {
    out2[0] = gapRangeMin(gap, k);
    out2[1] = gapRangeMax(gap, k);
}

// ----- k-mer statistics ---------------------------------------------------------
// values (4^k doubles) for a single-record reference, commands/map.go:45-71 with the canonical tie order
int dpo_kmer_values(const char* ref_ascii, long long n, int k, double* values_out) {
    DPO_TRY PackedSeq ref = NewPackedSequence(0, std::string(ref_ascii, (size_t)n), nullptr);
    std::vector<uint64_t> counts;
    KmerOccurrences(ref, k, counts);
    std::vector<double> v = KmerValues(counts, k);
    memcpy(values_out, v.data(), v.size() * sizeof(double));
    return 0;
    DPO_CATCH(1)
}
int dpo_kmer_counts(const char* ref_ascii, long long n, int k, unsigned long long* counts_out) {
    DPO_TRY PackedSeq ref = NewPackedSequence(0, std::string(ref_ascii, (size_t)n), nullptr);
    std::vector<uint64_t> counts;
    KmerOccurrences(ref, k, counts);
    memcpy(counts_out, counts.data(), counts.size() * sizeof(uint64_t));
    return 0;
    DPO_CATCH(1)
}

// accumulating batch form (getKmerValues of the overlap command counts every read): counts_io += occurrences
int dpo_kmer_counts_batch(const char* bases, const long long* offsets, long long nReads, int k, unsigned long long* counts_io) {
    DPO_TRY
    std::vector<uint64_t> counts(counts_io, counts_io + ((size_t)1 << (2 * k)));
    for (long long i = 0; i < nReads; i++) {
        if (offsets[i + 1] - offsets[i] < k) continue;
        PackedSeq r = NewPackedSequence(i, std::string(bases + offsets[i], (size_t)(offsets[i + 1] - offsets[i])), nullptr);
        KmerOccurrences(r, k, counts);
    }
    memcpy(counts_io, counts.data(), counts.size() * sizeof(uint64_t));
    return 0;
    DPO_CATCH(1)
}
int dpo_kmer_values_from_counts(unsigned long long* counts, int k, double* values_out) {
    DPO_TRY
    std::vector<uint64_t> c(counts, counts + ((size_t)1 << (2 * k)));
    std::vector<double> v = KmerValues(c, k);
    memcpy(values_out, v.data(), v.size() * sizeof(double));
    return 0;
    DPO_CATCH(1)
}

// ----- mapper -----------------------------------------------------------------
struct OMapper {
    Mapper m;
    Counters counters;
};

void* dpo_mapper_new(const char* ref_ascii, long long ref_len, int circular, int k, const double* kmer_values,
                     int seed_rate, int edge_size, int chunk_size) {
    DPO_TRY OMapper* om = new OMapper();
    PackedSeq ref = NewPackedSequence(0, std::string(ref_ascii, (size_t)ref_len), nullptr);
    NewMapper(om->m, ref, circular != 0, k, kmer_values, seed_rate, edge_size, chunk_size);
    om->m.refName = "ref";
    return om;
    DPO_CATCH(nullptr)
}
// lean: -1 auto / 0 / 1 (memory-lean index, oracle.hpp); threads: workers of the per-chunk seed scans
void* dpo_mapper_new_ex(const char* ref_ascii, long long ref_len, int circular, int k, const double* kmer_values,
                        int seed_rate, int edge_size, int chunk_size, int lean, int threads) {
    DPO_TRY OMapper* om = new OMapper();
    PackedSeq ref;
    {
        std::string text(ref_ascii, (size_t)ref_len);
        ref = NewPackedSequence(0, text, nullptr);
    }
    NewMapper(om->m, ref, circular != 0, k, kmer_values, seed_rate, edge_size, chunk_size, lean, threads);
    om->m.refName = "ref";
    return om;
    DPO_CATCH(nullptr)
}
int dpo_mapper_is_lean(void* p) { return ((OMapper*)p)->m.index.lean ? 1 : 0; }
void dpo_mapper_free(void* p) { delete (OMapper*)p; }
long long dpo_mapper_num_seeds(void* p) { return ((OMapper*)p)->m.index.size; }
long long dpo_mapper_num_chunks(void* p) { return (long long)((OMapper*)p)->m.index.sequences.size(); }
// seed k-mers in seed-id order
void dpo_mapper_seed_kmers(void* p, long long* out) {
    OMapper* om = (OMapper*)p;
    for (long long i = 0; i < om->m.index.size; i++) out[i] = om->m.index.seedMap[(size_t)i];
}
// chunk c: fields {offset, inset, length, nseeds}; segments (gap, kmer(not seed id), gap, ...)
long long dpo_mapper_chunk(void* p, long long c, long long* fields4, long long* segments, long long cap) {
    OMapper* om = (OMapper*)p;
    const SeedSequence s = ChunkSequence(om->m.index, (size_t)c);
    fields4[0] = s.offset;
    fields4[1] = s.inset;
    fields4[2] = s.length;
    fields4[3] = s.GetNumSeeds();
    if (segments) {
        for (size_t i = 0; i < s.segments.size() && (long long)i < cap; i++) {
            segments[i] = (i & 1) ? om->m.index.seedMap[(size_t)s.segments[i]] : s.segments[i];
        }
    }
    return (long long)s.segments.size();
}

// Stage dump for one window query (performMapping's inputs): the gapped-seed list of one strand.
// Seeds are reported as k-mer values (seed ids are an arbitrary labelling).
long long dpo_window_segments(void* p, const char* read_ascii, long long read_len, long long start, long long end,
                              int whole, int rc, long long* segments, long long cap, long long* fields3) {
    DPO_TRY OMapper* om = (OMapper*)p;
    PackedSeq read = NewPackedSequence(0, std::string(read_ascii, (size_t)read_len), nullptr);
    PackedSeq w = whole ? read : SubSequence(read, start, end);
    if (rc) w = ReverseComplement(w);
    SeedSequence s = NewSeedSequence(om->m.index, w, nullptr);
    for (size_t i = 0; i < s.segments.size() && (long long)i < cap; i++)
        segments[i] = (i & 1) ? om->m.index.seedMap[(size_t)s.segments[i]] : s.segments[i];
    fields3[0] = s.offset;
    fields3[1] = s.inset;
    fields3[2] = s.length;
    return (long long)s.segments.size();
    DPO_CATCH(-1)
}
// candidates of one window strand (SeedIndex.Matches)
long long dpo_window_candidates(void* p, const char* read_ascii, long long read_len, long long start, long long end,
                                int whole, int rc, long long* out, long long cap) {
    DPO_TRY OMapper* om = (OMapper*)p;
    PackedSeq read = NewPackedSequence(0, std::string(read_ascii, (size_t)read_len), nullptr);
    PackedSeq w = whole ? read : SubSequence(read, start, end);
    if (rc) w = ReverseComplement(w);
    SeedSequence s = NewSeedSequence(om->m.index, w, nullptr);
    std::vector<uint64_t> ids = Matches(om->m.index, s, 0.25, nullptr);
    for (size_t i = 0; i < ids.size() && (long long)i < cap; i++) out[i] = (long long)ids[i];
    return (long long)ids.size();
    DPO_CATCH(-1)
}
// performMapping of one window: rows of 6 {Start, End, QueryOffset, QueryInset, RC, ids}
long long dpo_window_mappings(void* p, const char* read_ascii, long long read_len, long long start, long long end,
                              int whole, long long* out, long long cap_rows) {
    DPO_TRY OMapper* om = (OMapper*)p;
    PackedSeq read = NewPackedSequence(0, std::string(read_ascii, (size_t)read_len), nullptr);
    PackedSeq w = whole ? read : SubSequence(read, start, end);
    std::vector<Mapping> r = performMappingPublic(om->m, w, nullptr);
    for (size_t i = 0; i < r.size() && (long long)i < cap_rows; i++) {
        out[i * 6 + 0] = r[i].Start;
        out[i * 6 + 1] = r[i].End;
        out[i * 6 + 2] = r[i].QueryOffset;
        out[i * 6 + 3] = r[i].QueryInset;
        out[i * 6 + 4] = r[i].RC ? 1 : 0;
        out[i * 6 + 5] = r[i].ids;
    }
    return (long long)r.size();
    DPO_CATCH(-1)
}

// Map a batch of reads with `threads` host threads. Results: out_offsets[n+1] and rows of 6 (as above) in a
// library-owned buffer returned through *rows_out (free with dpo_free). counters_out (11 long longs) optional.
int dpo_map_batch(void* p, long long n_reads, const char* bases, const long long* offsets, int threads,
                  long long** rows_out, long long* out_offsets, long long* counters_out) {
    DPO_TRY OMapper* om = (OMapper*)p;
    std::vector<std::vector<Mapping>> res((size_t)n_reads);
    if (threads < 1) threads = 1;
    std::vector<Counters> cs((size_t)threads);
    std::atomic<long long> next(0);
    std::vector<std::string> errs((size_t)threads);
    auto work = [&](int t) {
        try {
            for (;;) {
                long long i0 = next.fetch_add(64);
                if (i0 >= n_reads) break;
                long long i1 = std::min(n_reads, i0 + 64);
                for (long long i = i0; i < i1; i++) {
                    PackedSeq q = NewPackedSequence(i, std::string(bases + offsets[i], (size_t)(offsets[i + 1] - offsets[i])),
                                                    nullptr);
                    res[(size_t)i] = Map(om->m, q, &cs[(size_t)t]);
                }
            }
        } catch (const std::exception& ex) {
            errs[(size_t)t] = ex.what();
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    for (auto& e : errs)
        if (!e.empty()) throw std::runtime_error(e);
    long long total = 0;
    for (long long i = 0; i < n_reads; i++) {
        out_offsets[i] = total;
        total += (long long)res[(size_t)i].size();
    }
    out_offsets[n_reads] = total;
    long long* rows = (long long*)malloc(sizeof(long long) * 6 * (size_t)(total > 0 ? total : 1));
    long long r = 0;
    for (long long i = 0; i < n_reads; i++) {
        for (const Mapping& mp : res[(size_t)i]) {
            rows[r * 6 + 0] = mp.Start;
            rows[r * 6 + 1] = mp.End;
            rows[r * 6 + 2] = mp.QueryOffset;
            rows[r * 6 + 3] = mp.QueryInset;
            rows[r * 6 + 4] = mp.RC ? 1 : 0;
            rows[r * 6 + 5] = mp.ids;
            r++;
        }
    }
    *rows_out = rows;
    if (counters_out) {
        Counters tot;
        for (auto& c : cs) tot.add(c);
        om->counters.add(tot);
        counters_out[0] = tot.windows;
        counters_out[1] = tot.kmer_lookups;
        counters_out[2] = tot.query_seeds;
        counters_out[3] = tot.posting_runs;
        counters_out[4] = tot.posting_entries;
        counters_out[5] = tot.candidates;
        counters_out[6] = tot.cand_pass;
        counters_out[7] = tot.chain_cells;
        counters_out[8] = tot.chains;
        counters_out[9] = tot.mappings;
        counters_out[10] = tot.sort_ties_unpinned;
    }
    return 0;
    DPO_CATCH(1)
}
void dpo_free(void* p) { free(p); }


// ----- readFasta's first pass (seqio.cpp) -----------------------------------
// Returns a malloc'ed blob: for every record int64 name length, name bytes, int64 sequence length, sequence bytes;
// *n_records = records, *blob_bytes = size. nullptr on the reference's log.Fatal (invalid fastq).
unsigned char* dpo_parse_fasta(const char* content, long long n, long long minLength, long long* n_records, long long* blob_bytes) {
    DPO_TRY
    std::vector<FastaRecord> recs = ParseFasta(std::string(content, (size_t)n), minLength);
    size_t total = 0;
    for (auto& r : recs) total += 16 + r.name.size() + r.seq.size();
    unsigned char* blob = (unsigned char*)malloc(total ? total : 1);
    size_t o = 0;
    for (auto& r : recs) {
        long long a = (long long)r.name.size(), b = (long long)r.seq.size();
        memcpy(blob + o, &a, 8);
        memcpy(blob + o + 8, r.name.data(), r.name.size());
        o += 8 + r.name.size();
        memcpy(blob + o, &b, 8);
        memcpy(blob + o + 8, r.seq.data(), r.seq.size());
        o += 8 + r.seq.size();
    }
    *n_records = (long long)recs.size();
    *blob_bytes = (long long)total;
    return blob;
    DPO_CATCH(nullptr)
}

// ----- one round of `downpore overlap` (overlap.cpp) -------------------------
// params7 = {overlapSize, k, numSeeds, seedBatchSize, chunkSize, queryBatchSize}; returns a handle or nullptr.
void* dpo_overlap_round(const char* bases, const long long* offsets, long long nReads, const unsigned char* ignore,
                        long long firstSequence, const double* values, const long long* params6, double hitFraction) {
    DPO_TRY
    std::vector<PackedSeq> reads;
    reads.reserve((size_t)nReads);
    for (long long i = 0; i < nReads; i++)
        reads.push_back(NewPackedSequence(i, std::string(bases + offsets[i], (size_t)(offsets[i + 1] - offsets[i])), nullptr));
    std::vector<uint8_t> ign((size_t)nReads, 0);
    if (ignore) ign.assign(ignore, ignore + nReads);
    OverlapParams P;
    P.overlapSize = params6[0];
    P.k = params6[1];
    P.numSeeds = params6[2];
    P.seedBatchSize = params6[3];
    P.chunkSize = params6[4];
    P.queryBatchSize = params6[5];
    P.hitFraction = hitFraction;
    OverlapRound* R = new OverlapRound();
    try {
        OverlapRoundRun(reads, ign, firstSequence, values, P, *R);
    } catch (...) {
        delete R;
        throw;
    }
    return R;
    DPO_CATCH(nullptr)
}
void dpo_overlap_free(void* h) { delete (OverlapRound*)h; }
// what: 0 header {numSeeds, numQueries, numChunks, numHits, numQuerySeqs, nextFirstSequence}
//       1 seed k-mers in seed id order
//       2 queries:  per query {ID, SequenceID, rc, length, offset, inset, nSegments, segments...}
//       3 chunks:   per chunk {id, length, offset, inset, nSegments, segments...}
//       4 hits:     per hit {queryID, rc, target, n, MatchA..., MatchB...}
// Returns the number of int64 the section holds; fills out when cap suffices.
long long dpo_overlap_get(void* h, int what, long long* out, long long cap) {
    OverlapRound& R = *(OverlapRound*)h;
    std::vector<long long> v;
    if (what == 0) {
        v = {R.index.size, (long long)R.queries.size(), (long long)R.index.sequences.size(), (long long)R.hits.size(), R.numQuerySeqs,
             R.nextFirstSequence};
    } else if (what == 1) {
        for (gint i = 0; i < R.index.size; i++) v.push_back(R.index.seedMap[(size_t)i]);
    } else if (what == 2) {
        for (auto& q : R.queries) {
            v.insert(v.end(), {q.ID, q.SequenceID, q.rc ? 1 : 0, q.Query.length, q.Query.offset, q.Query.inset, (long long)q.Query.segments.size()});
            v.insert(v.end(), q.Query.segments.begin(), q.Query.segments.end());
        }
    } else if (what == 3) {
        for (auto& c : R.index.sequences) {
            v.insert(v.end(), {c.id, c.length, c.offset, c.inset, (long long)c.segments.size()});
            v.insert(v.end(), c.segments.begin(), c.segments.end());
        }
    } else if (what == 4) {
        for (auto& x : R.hits) {
            v.insert(v.end(), {x.queryID, x.rc ? 1 : 0, x.target, (long long)x.MatchA.size()});
            v.insert(v.end(), x.MatchA.begin(), x.MatchA.end());
            v.insert(v.end(), x.MatchB.begin(), x.MatchB.end());
        }
    }
    if ((long long)v.size() <= cap && out) memcpy(out, v.data(), v.size() * sizeof(long long));
    return (long long)v.size();
}
}  // extern "C"
