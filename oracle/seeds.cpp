// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Follows seeds/seeds.go and the hot part of seeds/sequence.go of the reference.
#include "oracle.hpp"

#include <deque>

#include <stdexcept>

namespace dpo {

void Counters::add(const Counters& o) {
    windows += o.windows;
    kmer_lookups += o.kmer_lookups;
    query_seeds += o.query_seeds;
    posting_runs += o.posting_runs;
    posting_entries += o.posting_entries;
    candidates += o.candidates;
    cand_pass += o.cand_pass;
    chain_cells += o.chain_cells;
    chains += o.chains;
    mappings += o.mappings;
    sort_ties_unpinned += o.sort_ties_unpinned;
}

// ---------------------------------------------------------------------------
// seeds/sequence.go
// ---------------------------------------------------------------------------
gint GetSeedOffset(const SeedSequence& s, gint index, gint k) {  // sequence.go:1239-1246
    index = index * 2 + 1;
    gint offset = s.segments[0];
    for (gint i = 2; i < index; i += 2) offset += s.segments[(size_t)i] + k;
    return offset;
}

gint GetSeedOffsetFromEnd(const SeedSequence& s, gint index, gint k) {  // sequence.go:1269-1276
    index = index * 2 + 1;
    gint offset = s.segments[s.segments.size() - 1];
    for (gint i = (gint)s.segments.size() - 3; i > index; i -= 2) offset += s.segments[(size_t)i] + k;
    return offset;
}

uint64_t ReverseComplementKmer(uint64_t seed, uint64_t k) {  // sequence.go:125-132
    uint64_t rc = 0;
    for (uint64_t j = 0; j < k; j++) {
        rc = (rc << 2) | ((seed ^ 3) & 3);
        seed = seed >> 2;
    }
    return rc;
}

// sequence.go:85-123. Returns false for the (nil, nil) result.
bool Reduced(const SeedSequence& s, const IntSet& whitelist, gint k, gint minSeeds, SeedSequence* reduced,
             std::vector<gint>* index) {
    gint count = 0;
    gint n = (gint)s.segments.size();
    gint prev = -1;
    for (gint i = 1; i < n; i += 2) {
        gint next = s.segments[(size_t)i];
        if (next != prev && Contains(whitelist, (uint64_t)next)) {
            count++;
            prev = next;
        }
    }
    if (count < minSeeds) return false;
    std::vector<gint> segs((size_t)(count * 2 + 1));
    gint offset = s.segments[0];
    index->assign((size_t)count, 0);
    prev = -1;
    gint j = 0;
    for (gint i = 1; i < n; i += 2) {
        gint seed = s.segments[(size_t)i];
        if (prev != seed && Contains(whitelist, (uint64_t)seed)) {
            segs[(size_t)j] = offset;
            segs[(size_t)j + 1] = seed;
            (*index)[(size_t)(j / 2)] = i / 2;
            j += 2;
            offset = s.segments[(size_t)i + 1];
            prev = seed;
        } else {
            offset += s.segments[(size_t)i + 1] + k;
        }
    }
    segs[(size_t)j] = offset;
    reduced->segments = std::move(segs);
    reduced->length = s.length;
    reduced->offset = s.offset;
    reduced->inset = s.inset;
    reduced->rc = s.rc;
    reduced->id = s.id;
    return true;
}

namespace {
// A Go slice of ints: shared backing array + length. nil <=> arr == nullptr.
struct Chain {
    std::shared_ptr<std::vector<gint>> arr;
    gint len = 0;
    bool nil() const { return !arr; }
    gint last() const { return (*arr)[(size_t)len - 1]; }
    std::vector<gint> values() const { return std::vector<gint>(arr->begin(), arr->begin() + len); }
};
Chain make_chain(gint first, gint cap) {  // make([]int, 1, cap); c[0] = first
    Chain c;
    c.arr = std::make_shared<std::vector<gint>>();
    c.arr->reserve((size_t)(cap > 1 ? cap : 1));
    c.arr->push_back(first);
    c.len = 1;
    return c;
}
Chain append(const Chain& c, gint v) {  // append within capacity writes in place at index len
    Chain r = c;
    if ((gint)r.arr->size() > r.len) (*r.arr)[(size_t)r.len] = v;
    else r.arr->push_back(v);
    r.len++;
    return r;
}

// sequence.go:476-576
void extendChain(const SeedSequence& a, const SeedSequence& b, std::vector<Chain>& chainsA, std::vector<Chain>& chainsB,
                 gint aIndex, gint bIndex, gint k, Chain* outA, Chain* outB) {
    Chain currentChainA = chainsA[(size_t)(aIndex / 2)];
    Chain currentChainB = chainsB[(size_t)(aIndex / 2)];
    const std::vector<gint>& as = a.segments;
    const std::vector<gint>& bs = b.segments;
    const gint alen = (gint)as.size(), blen = (gint)bs.size();
    gint offsetA = as[(size_t)aIndex + 1];
    gint offsetB = bs[(size_t)bIndex + 1];
    aIndex += 2;
    bIndex += 2;
    while (aIndex < alen && bIndex < blen) {
        gint aSeedIndex = aIndex / 2;
        gint minBOffset, maxBOffset;
        if (offsetA < 0) {
            minBOffset = -k;
            maxBOffset = 0;
        } else {
            minBOffset = (offsetA * 2) / 3 - k;
            maxBOffset = (offsetA * 3) / 2 + k;
        }
        while (maxBOffset < offsetB) {
            offsetA += as[(size_t)aIndex + 1] + k;
            aIndex += 2;
            if (aIndex >= alen) {
                *outA = currentChainA;
                *outB = currentChainB;
                return;
            }
            aSeedIndex = aIndex / 2;
            minBOffset = (offsetA * 2) / 3 - k;
            maxBOffset = (offsetA * 3) / 2 + k;
        }
        while (offsetB < minBOffset) {
            offsetB += bs[(size_t)bIndex + 1] + k;
            bIndex += 2;
            if (bIndex >= blen) {
                *outA = currentChainA;
                *outB = currentChainB;
                return;
            }
        }
        gint oldBIndex = bIndex;
        gint oldBOffset = offsetB;
        bool matched = false;
        gint seedA = as[(size_t)aIndex];
        while (offsetB <= maxBOffset) {
            if (seedA == bs[(size_t)bIndex]) {
                if (!chainsA[(size_t)aSeedIndex].nil()) {
                    if (bIndex / 2 == chainsB[(size_t)aSeedIndex].last() &&
                        chainsA[(size_t)aSeedIndex].len > currentChainA.len) {
                        *outA = currentChainA;
                        *outB = currentChainB;
                        return;
                    }
                }
                currentChainA = append(currentChainA, aSeedIndex);
                chainsA[(size_t)aSeedIndex] = currentChainA;
                currentChainB = append(currentChainB, bIndex / 2);
                chainsB[(size_t)aSeedIndex] = currentChainB;
                offsetA = as[(size_t)aIndex + 1];
                offsetB = bs[(size_t)bIndex + 1];
                aIndex += 2;
                bIndex += 2;
                matched = true;
                break;
            } else {
                offsetB += bs[(size_t)bIndex + 1] + k;
                bIndex += 2;
                if (bIndex >= blen) break;
            }
        }
        if (!matched) {
            offsetA += as[(size_t)aIndex + 1] + k;
            aIndex += 2;
            offsetB = oldBOffset;
            bIndex = oldBIndex;
        }
    }
    *outA = currentChainA;
    *outB = currentChainB;
}

// sequence.go:401-471. `seq` is the receiver (reduced chunk), `query` the reduced query.
std::vector<SeedMatch> dynamicMatch(const SeedSequence& seq, const SeedSequence& query, gint minMatch, gint k) {
    if (minMatch == 0) minMatch = 1;
    const std::vector<gint>& qs = query.segments;
    const std::vector<gint>& ss = seq.segments;
    const gint qlen = (gint)qs.size(), slen = (gint)ss.size();
    std::vector<Chain> chainsA((size_t)(qlen / 2)), chainsB((size_t)(qlen / 2));
    std::vector<SeedMatch> allGoodChains;
    for (gint qIndex = 1; qIndex < qlen - minMatch * 2 + 2; qIndex += 2) {
        if (qs[(size_t)qIndex - 1] < 0 && qIndex > 1 && qs[(size_t)qIndex + 1] < 0 &&
            qs[(size_t)qIndex] == qs[(size_t)qIndex - 2] && qs[(size_t)qIndex] == qs[(size_t)qIndex + 2]) {
            continue;
        }
        gint querySeedIndex = qIndex / 2;
        if (!chainsA[(size_t)querySeedIndex].nil()) continue;
        gint prevSeed = -1;
        for (gint i = 1; i < slen - minMatch * 2 + 2; i += 2) {
            gint nextSeed = ss[(size_t)i];
            if (nextSeed == qs[(size_t)qIndex] && nextSeed != prevSeed &&
                (chainsA[(size_t)querySeedIndex].nil() || chainsB[(size_t)querySeedIndex].last() != i / 2)) {
                chainsA[(size_t)querySeedIndex] = make_chain(querySeedIndex, (qlen - qIndex) / 2);
                chainsB[(size_t)querySeedIndex] = make_chain(i / 2, (qlen - qIndex) / 2);
                Chain chainA, chainB;
                extendChain(query, seq, chainsA, chainsB, qIndex, i, k, &chainA, &chainB);
                if (chainA.len >= minMatch) {
                    gint nextLength = (chainA.len * 2) / 3;
                    if (nextLength > minMatch) {
                        minMatch = nextLength;
                        for (gint j = (gint)allGoodChains.size() - 1; j >= 0; j--) {
                            if ((gint)allGoodChains[(size_t)j].MatchA.size() < nextLength) {
                                allGoodChains[(size_t)j] = allGoodChains[allGoodChains.size() - 1];
                                allGoodChains.pop_back();
                            }
                        }
                    }
                    SeedMatch sm;
                    sm.MatchA = chainA.values();
                    sm.MatchB = chainB.values();
                    sm.SeqA = &query;
                    sm.SeqB = &seq;
                    allGoodChains.push_back(std::move(sm));
                    gint remaining = 0;
                    for (const Chain& c : chainsA)
                        if (c.nil()) remaining++;
                    if (remaining < chainA.len) return allGoodChains;
                }
            }
            prevSeed = nextSeed;
        }
    }
    return allGoodChains;
}
}  // namespace

static thread_local Counters* tl_counters = nullptr;
struct CounterScope {
    Counters* prev;
    explicit CounterScope(Counters* c) : prev(tl_counters) { tl_counters = c; }
    ~CounterScope() { tl_counters = prev; }
};

// sequence.go:361-394
std::vector<SeedMatch> Match(const SeedSequence& seq, const SeedSequence& query, const IntSet& querySet,
                             const IntSet& seqSet, gint minMatch, gint k, bool* nil_result) {
    SeedSequence s, q;
    std::vector<gint> sIndex, qIndex;
    bool sOK = Reduced(seq, querySet, k, minMatch, &s, &sIndex);
    bool qOK = Reduced(query, seqSet, k, minMatch, &q, &qIndex);
    if (!sOK || !qOK) {
        *nil_result = true;
        return {};
    }
    if (tl_counters) tl_counters->chain_cells += s.GetNumSeeds() + q.GetNumSeeds();
    std::vector<SeedMatch> ms = dynamicMatch(s, q, minMatch, k);
    for (SeedMatch& m : ms) {
        for (gint& pos : m.MatchA) pos = qIndex[(size_t)pos];
        for (gint& pos : m.MatchB) pos = sIndex[(size_t)pos];
        m.SeqA = &query;
        m.SeqB = &seq;
    }
    *nil_result = ms.empty();
    return ms;
}

void set_match_counters(Counters* c) { tl_counters = c; }

// dynamicMatch on two sequences as they are (no Reduced): what the hand-worked vectors of tests/test_oracle_handworked.py pin
std::vector<SeedMatch> DynamicMatch(const SeedSequence& seq, const SeedSequence& query, gint minMatch, gint k) {
    std::vector<SeedMatch> ms = dynamicMatch(seq, query, minMatch, k);
    for (SeedMatch& m : ms) {
        m.SeqA = &query;
        m.SeqB = &seq;
    }
    return ms;
}

// sequence.go:830-858
void GetBasesCovered(const SeedMatch& m, gint k, gint* outA, gint* outB) {
    gint countA = (gint)m.MatchA.size() * k;
    gint countB = countA;
    gint prevA = m.MatchA[0];
    gint prevB = m.MatchB[0];
    const std::vector<gint>& sa = m.SeqA->segments;
    const std::vector<gint>& sb = m.SeqB->segments;
    for (size_t i = 0; i < m.MatchA.size(); i++) {
        gint s = m.MatchA[i];
        if (i == 0) continue;
        gint d1 = sa[(size_t)(prevA * 2 + 2)];
        gint d2 = sb[(size_t)(prevB * 2 + 2)];
        for (gint j = prevA + 2; j <= s; j++) d1 += sa[(size_t)(j * 2)] + k;
        gint s2 = m.MatchB[i];
        for (gint j = prevB + 2; j <= s2; j++) d2 += sb[(size_t)(j * 2)] + k;
        if (d1 < 0) countA += d1;
        if (d2 < 0) countB += d2;
        prevB = s2;
        prevA = s;
    }
    *outA = countA;
    *outB = countB;
}

// ---------------------------------------------------------------------------
// seeds/seeds.go
// ---------------------------------------------------------------------------
void NewSeedIndex(SeedIndex& g, gint k) {  // seeds.go:23-31
    size_t size = 1;
    for (gint j = k; j > 0; j--) size *= 4;
    g.kmers.assign(size, 0);
    g.kmerMap.assign(size, 0);
    g.seedSize = k;
    g.sequences.clear();
    g.sequenceSets.clear();
    g.seedSets.clear();
    g.seedMap.clear();
    g.size = 0;
}

SeedSequence NewSeedSequence(const SeedIndex& g, const PackedSeq& seq, Counters* c) {  // seeds.go:33-50
    gint k = g.seedSize;
    gint count = CountKmers(seq, seq.Len(), k, g.kmers.data());
    std::vector<gint> segments((size_t)(count * 2 + 1), 0);
    {
        // The reference writes straight into make([]int, count*2+1). Write into a sentinel-filled buffer large
        // enough for every visited k-mer instead, and check the writer produced exactly `count` pairs.
        const gint sentinel = INT64_MIN;
        std::vector<gint> big((size_t)(2 * (seq.Len() + 16) + 1), sentinel);
        WriteSegments(seq, big.data(), k, g.kmers.data());
        size_t lastw = big.size();
        while (lastw > 0 && big[lastw - 1] == sentinel) lastw--;
        if (lastw != (size_t)(count * 2 + 1)) throw std::runtime_error("oracle: CountKmers / WriteSegments disagree");
        for (size_t i = 0; i < segments.size(); i++) segments[i] = big[i];
        gint total = 0;
        for (size_t i = 0; i < segments.size(); i += 2) total += segments[i];
        total += count * k;
        if (c) c->kmer_lookups += total - k + 1;  // scanned bases - k + 1 = k-mers visited
    }
    for (size_t i = 1; i < segments.size(); i += 2) segments[i] = (gint)g.kmerMap[(size_t)segments[i]];
    if (c) c->query_seeds += count;
    SeedSequence s;
    s.segments = std::move(segments);
    s.length = seq.Len();
    s.id = seq.id;
    s.offset = seq.offset;
    s.inset = seq.inset;
    s.rc = false;
    return s;
}

static void register_seed(SeedIndex& g, gint kmer) {  // seeds.go:186-197 (body of the lock)
    if (!g.kmers[(size_t)kmer]) {
        g.kmers[(size_t)kmer] = 1;
        g.kmerMap[(size_t)kmer] = (int32_t)g.size;
        while ((gint)g.sequenceSets.size() <= g.size) {
            g.sequenceSets.push_back(NewIntSet());
            g.seedMap.push_back(-1);
        }
        g.seedMap[(size_t)g.size] = kmer;
        g.size++;
    }
}

void AddSingleSeeds(SeedIndex& g, const PackedSeq& seq, gint seedRate, const double* ranks) {  // seeds.go:160-200
    gint k = g.seedSize;
    gint mask = 0;
    for (gint i = 0; i < k; i++) mask = (mask << 2) | 3;
    for (gint i = 0; i < seq.Len() - seedRate; i += seedRate) {
        gint count = CountKmersBetween(seq, i, i + seedRate, 1, k, g.kmers.data());
        if (count == 0) {
            gint end = i + seedRate;
            gint kmer = KmerAt(seq, i, k);
            double bestValue = ranks[kmer];
            gint bestKmer = kmer;
            for (gint j = i + k; j < end; j++) {
                kmer = NextKmer(seq, kmer, mask, j);
                double value = ranks[kmer];
                if (value > bestValue) {
                    bestValue = value;
                    bestKmer = kmer;
                }
            }
            register_seed(g, bestKmer);
        }
    }
}

// seeds.go:62-156 — batch seed selection of the overlap path (round-2 groundwork: not used by `map`).
// `quality` is seq.Quality() (nil for FASTA). The reference computes CountKmers first and then discards the result
// (`count = 0`, seeds.go:76), so every call tops the index up by `minSeeds` k-mers: the best-valued k-mer of each
// k-block that contains no seed yet, blocks 3k apart. Quirks kept: the k-mer at position 0 is never a candidate; slots
// of topN that no block filled stay 0, so k-mer 0 (and its reverse complement) are registered as seeds.
void AddSeeds(SeedIndex& g, const PackedSeq& seq, gint minSeeds, const double* kmerRanks, const uint8_t* quality) {
    gint k = g.seedSize;
    gint mask = 0;
    for (gint i = 0; i < k; i++) mask = (mask << 2) | 3;
    gint count = 0;
    if (count < minSeeds) {
        std::vector<uint64_t> topN((size_t)(minSeeds - count), 0);
        std::vector<double> topNValues((size_t)(minSeeds - count), 0.0);
        gint kmer = KmerAt(seq, 0, k);
        gint nextIndex = k;
        while (nextIndex < seq.Len() - k) {
            bool reset = false;
            double bestValue = 0.0;
            gint bestSeed = 0;
            for (gint i = 0; nextIndex < seq.Len() && i < k; i++) {
                kmer = NextKmer(seq, kmer, mask, nextIndex);
                nextIndex++;
                if (g.kmers[(size_t)kmer]) {
                    reset = true;
                    break;
                }
                double value = kmerRanks[kmer];
                if (quality) value *= (double)quality[nextIndex - k / 2];
                if (value > bestValue) {
                    bestValue = value;
                    bestSeed = kmer;
                }
            }
            if (!reset) {
                size_t n = 0;
                for (; n < topNValues.size() && topNValues[n] < bestValue; n++) {
                    if (n > 0) {
                        topNValues[n - 1] = topNValues[n];
                        topN[n - 1] = topN[n];
                    }
                }
                if (n > 0) {
                    topNValues[n - 1] = bestValue;
                    topN[n - 1] = (uint64_t)bestSeed;
                }
            }
            nextIndex += k;
            if (nextIndex < seq.Len() - k) kmer = KmerAt(seq, nextIndex, k);
            nextIndex += k;
        }
        for (uint64_t km : topN) {  // the body of the lock (seeds.go:131-154)
            register_seed(g, (gint)km);
            register_seed(g, (gint)ReverseComplementKmer(km, (uint64_t)k));
        }
    }
}

// seeds/sequence.go:134-159 — the reverse-complement query of the overlap path: gaps reversed, every seed replaced by
// the seed id of its reverse-complement k-mer (kmerMap is 0 for a k-mer that is no seed: AddSeeds registers both
// strands, so that does not happen there). offset/inset stay those of the forward read.
SeedSequence ReverseComplementSeq(const SeedSequence& s, gint k, const SeedIndex& g) {
    size_t n = s.segments.size();
    SeedSequence r;
    r.segments.assign(n, 0);
    for (size_t i = 0; i < n; i++) {
        gint v = s.segments[i];
        if ((i & 1) == 0) {
            r.segments[n - 1 - i] = v;
        } else {
            uint64_t rc = ReverseComplementKmer((uint64_t)g.seedMap[(size_t)v], (uint64_t)k);
            r.segments[n - 1 - i] = (gint)g.kmerMap[(size_t)rc];
        }
    }
    r.id = s.id;
    r.length = s.length;
    r.offset = s.offset;
    r.inset = s.inset;
    r.rc = !s.rc;
    return r;
}

// seeds/sequence.go:46-50 (the Go slice aliases the parent's segments; nothing on this path writes through it)
SeedSequence SubSequenceSeeds(const SeedSequence& s, gint start, gint end, gint length, gint offset, gint inset) {
    SeedSequence r;
    r.segments.assign(s.segments.begin() + start * 2, s.segments.begin() + end * 2 + 3);
    r.length = length;
    r.offset = offset;
    r.inset = inset;
    r.rc = s.rc;
    r.id = s.id;
    return r;
}

static gint GetNextSeedOffset(const SeedSequence& s, gint index, gint k) {  // sequence.go:1278-1280
    return s.segments[(size_t)(index * 2 + 2)] + k;
}

// overlap/overlap.go:253-318 (chunkWorker, one sequence): the pieces handed to index.AddSequence, in order. A read's
// seed sequence is cut in SEED space: up to 100 seeds or chunkSize bases per piece, stepping back 5 seeds or overlap/2
// bases between pieces; a piece with fewer than minSeeds seeds is dropped; from 150 seeds before the end on, one last
// piece runs to the end.
std::vector<SeedSequence> ChunkSeedSequence(const SeedSequence& s, gint chunkSize, gint minSeeds, gint overlap, gint k) {
    std::vector<SeedSequence> out;
    gint numChunks = s.Len() / chunkSize + 1;
    if (numChunks == 1 || s.GetNumSeeds() < minSeeds * 3) {
        if (s.GetNumSeeds() >= minSeeds) out.push_back(s);
        return out;
    }
    gint prevSeedIndex = 0;
    gint totalOffset = GetSeedOffset(s, 0, k);
    gint lengthInBases = 0;
    for (;;) {
        gint seedCount = 0;
        if (prevSeedIndex >= s.GetNumSeeds() - 150) {
            if (prevSeedIndex == 0) {
                out.push_back(s);
            } else {
                gint newFirstGap = GetNextSeedOffset(s, prevSeedIndex - 1, k) - k;
                lengthInBases += GetSeedOffsetFromEnd(s, prevSeedIndex, k) + k + newFirstGap;
                out.push_back(SubSequenceSeeds(s, prevSeedIndex, s.GetNumSeeds() - 1, lengthInBases, totalOffset - newFirstGap, 0));
            }
            break;
        }
        for (; lengthInBases < chunkSize && seedCount < 100 && prevSeedIndex + seedCount < s.GetNumSeeds(); seedCount++)
            lengthInBases += GetNextSeedOffset(s, prevSeedIndex + seedCount, k);
        if (seedCount >= minSeeds) {
            gint newFirstGap = GetNextSeedOffset(s, prevSeedIndex - 1, k) - k;
            lengthInBases += newFirstGap;
            out.push_back(SubSequenceSeeds(s, prevSeedIndex, prevSeedIndex + seedCount - 1, lengthInBases, totalOffset - newFirstGap,
                                           s.length - totalOffset - lengthInBases + newFirstGap));
            totalOffset += lengthInBases - newFirstGap;
            lengthInBases = 0;
            prevSeedIndex += seedCount;
            if (prevSeedIndex >= s.GetNumSeeds()) break;
            for (seedCount = 0; seedCount < 5 && lengthInBases < overlap / 2 && prevSeedIndex > 0; seedCount++) {
                prevSeedIndex--;
                gint step = GetNextSeedOffset(s, prevSeedIndex, k);
                lengthInBases += step;
                totalOffset -= step;
            }
            lengthInBases = 0;
        } else {
            prevSeedIndex += seedCount;
            for (seedCount = 0; lengthInBases < overlap / 2 && prevSeedIndex > 0; seedCount++) {
                prevSeedIndex--;
                gint step = GetNextSeedOffset(s, prevSeedIndex, k);
                lengthInBases += step;
                totalOffset -= step;
            }
            lengthInBases = 0;
        }
    }
    return out;
}

// ---- memory-lean mode: the same sets, rebuilt per query from lists (see oracle.hpp) ----
static void IndexSequencesLean(SeedIndex& g) {
    const size_t S = (size_t)g.size, C = g.sequences.size();
    std::vector<uint32_t> last(S, 0xffffffffu);
    g.leanSeedOff.assign(S + 1, 0);
    for (size_t c = 0; c < C; c++) {
        const std::vector<int32_t>& seg = g.leanSegments[c];
        for (size_t j = 1; j < seg.size(); j += 2) {
            const size_t seed = (size_t)seg[j];
            if (last[seed] != (uint32_t)c) {
                last[seed] = (uint32_t)c;
                g.leanSeedOff[seed + 1]++;
            }
        }
    }
    for (size_t s2 = 0; s2 < S; s2++) g.leanSeedOff[s2 + 1] += g.leanSeedOff[s2];
    g.leanSeedChunks.assign((size_t)g.leanSeedOff[S], 0);
    std::vector<uint64_t> fill(g.leanSeedOff.begin(), g.leanSeedOff.end() - 1);
    std::fill(last.begin(), last.end(), 0xffffffffu);
    for (size_t c = 0; c < C; c++) {
        const std::vector<int32_t>& seg = g.leanSegments[c];
        for (size_t j = 1; j < seg.size(); j += 2) {
            const size_t seed = (size_t)seg[j];
            if (last[seed] != (uint32_t)c) {
                last[seed] = (uint32_t)c;
                g.leanSeedChunks[(size_t)fill[seed]++] = (uint32_t)c;
            }
        }
    }
    g.sequenceSets.clear();
    g.sequenceSets.shrink_to_fit();
}

uint64_t SeedChunkCount(const SeedIndex& g, gint seed) {
    if (!g.lean) return g.sequenceSets[(size_t)seed].count;
    return g.leanSeedOff[(size_t)seed + 1] - g.leanSeedOff[(size_t)seed];
}

IntSet SeedChunkSet(const SeedIndex& g, gint seed) {
    if (!g.lean) return g.sequenceSets[(size_t)seed];
    IntSet s = NewIntSet();  // register_seed (seeds.go:190)
    // IndexSequences walks the chunks in descending id order (seeds.go:372-384)
    for (uint64_t p = g.leanSeedOff[(size_t)seed + 1]; p > g.leanSeedOff[(size_t)seed]; p--) Add(s, (uint64_t)g.leanSeedChunks[(size_t)p - 1]);
    return s;
}

SeedSequence ChunkSequence(const SeedIndex& g, size_t chunk) {
    SeedSequence s = g.sequences[chunk];
    if (g.lean) s.segments.assign(g.leanSegments[chunk].begin(), g.leanSegments[chunk].end());
    return s;
}

IntSet ChunkSeedSet(const SeedIndex& g, size_t chunk) {
    if (!g.lean) return g.seedSets[chunk];
    const std::vector<int32_t>& seg = g.leanSegments[chunk];  // AddSequence (seeds.go:272-290)
    gint maxSeed = 0;
    for (size_t i = 1; i < seg.size(); i += 2)
        if (seg[i] > maxSeed) maxSeed = seg[i];
    IntSet seedSet = NewIntSetCapacity(maxSeed + 1);
    for (size_t i = 1; i < seg.size(); i += 2) Add(seedSet, (uint64_t)seg[i]);
    return seedSet;
}

void AddSequence(SeedIndex& g, SeedSequence&& seq) {  // seeds.go:272-290
    gint maxSeed = 0;
    for (size_t i = 1; i < seq.segments.size(); i += 2) {
        gint seed = seq.segments[i];
        if (seed > maxSeed) maxSeed = seed;
    }
    IntSet seedSet = NewIntSetCapacity(maxSeed + 1);
    for (size_t i = 1; i < seq.segments.size(); i += 2) Add(seedSet, (uint64_t)seq.segments[i]);
    g.sequences.push_back(std::move(seq));
    g.seedSets.push_back(std::move(seedSet));
}

// seeds.go:292-305 + 372-384. The seed-range partition over 4 goroutines covers every seed exactly once and
// each worker walks the chunks in descending id order; a single descending walk is equivalent.
void IndexSequences(SeedIndex& g) {
    if (g.lean) {
        IndexSequencesLean(g);
        return;
    }
    for (gint i = (gint)g.sequences.size() - 1; i >= 0; i--) {
        const SeedSequence& s = g.sequences[(size_t)i];
        for (size_t j = 1; j < s.segments.size(); j += 2) {
            gint seed = s.segments[j];
            Add(g.sequenceSets[(size_t)seed], (uint64_t)i);
        }
    }
}

std::vector<uint64_t> Matches(const SeedIndex& g, const SeedSequence& query, double hitFraction, Counters* c) {
    std::vector<const IntSet*> allSeedSets;  // seeds.go:335-353
    gint prevSeed = -1;
    uint64_t maxSeqs = (uint64_t)g.sequences.size();
    std::deque<IntSet> rebuilt;  // lean mode: the sets of this query, each distinct seed rebuilt once
    std::vector<std::pair<gint, const IntSet*>> have;
    for (size_t i = 1; i < query.segments.size(); i += 2) {
        gint seed = query.segments[i];
        if (seed != prevSeed && SeedChunkCount(g, seed) < maxSeqs) {
            const IntSet* adj = nullptr;
            if (!g.lean) {
                adj = &g.sequenceSets[(size_t)seed];
            } else {
                for (auto& h : have)
                    if (h.first == seed) adj = h.second;
                if (!adj) {
                    rebuilt.push_back(SeedChunkSet(g, seed));
                    adj = &rebuilt.back();
                    have.emplace_back(seed, adj);
                }
            }
            allSeedSets.push_back(adj);
            prevSeed = seed;
        }
    }
    if (allSeedSets.size() < 5) return {};
    if (c) {
        c->posting_runs += (long long)allSeedSets.size();
        for (const IntSet* s : allSeedSets) c->posting_entries += (long long)s->count;
    }
    gint minCount = (gint)(hitFraction * (double)allSeedSets.size() + 0.5);
    return GetSharedIDs(allSeedSets, minCount, true);
}

}  // namespace dpo
