// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Follows mapping/mapping.go, util/sequtil/kmers.go and commands/map.go:45-71 of the reference.
#include "oracle.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <deque>
#include <thread>
#include <stdexcept>

namespace dpo {

void set_match_counters(Counters* c);  // seeds.cpp

// ---------------------------------------------------------------------------
// util/sequtil/kmers.go
// ---------------------------------------------------------------------------
void KmerOccurrences(const PackedSeq& seq, gint k, std::vector<uint64_t>& counts) {  // kmers.go:53-69
    gint mask = 0;
    for (gint i = 0; i < k; i++) mask = (mask << 2) | 3;
    size_t size = (size_t)1 << (unsigned)(k * 2);
    if (counts.size() != size) counts.assign(size, 0);
    gint kmer = KmerAt(seq, 0, k);
    counts[(size_t)kmer]++;
    for (gint i = k; i < seq.Len(); i++) {
        kmer = NextKmer(seq, kmer, mask, i);
        counts[(size_t)kmer]++;
    }
}

// kmers.go:87-112 — second return value (`vs.ids[len-topN:]`). The forward/rc merge runs in place over ascending i,
// so a non-palindromic pair ends as 2*(c+c_rc). Go's unstable sort.Sort is replaced by the total order
// (count asc, id asc): Q10, parity unpinned, which is why the C ABI takes `values[]` as an input.
std::vector<gint> TopOccurrencesTop(std::vector<uint64_t>& counts, gint k, gint topN) {
    size_t n = counts.size();
    std::vector<gint> ids(n);
    for (size_t i = 0; i < n; i++) {
        uint64_t c = counts[i];
        ids[i] = (gint)i;
        uint64_t rc = ReverseComplementKmer((uint64_t)i, (uint64_t)k);
        c += counts[rc];
        counts[i] = c;
        counts[rc] = c;
    }
    std::stable_sort(ids.begin(), ids.end(), [&](gint a, gint b) { return counts[(size_t)a] < counts[(size_t)b]; });
    return std::vector<gint>(ids.end() - topN, ids.end());
}

// commands/map.go:46-71
std::vector<double> KmerValues(std::vector<uint64_t>& kmerCounts, gint k) {
    std::vector<double> values(kmerCounts.size());
    uint64_t tot = 0;
    for (uint64_t c : kmerCounts) tot += c;
    double tf = (double)tot;
    double targetFreq = 0.000005;
    for (size_t i = 0; i < kmerCounts.size(); i++) {
        uint64_t count = kmerCounts[i];
        double freq = (double)count / tf;
        if (count < 3) values[i] = 0;
        else if (freq <= targetFreq) values[i] = 1.0 - (targetFreq - freq);
        else values[i] = 1.0 - (freq - targetFreq);
    }
    std::vector<gint> top = TopOccurrencesTop(kmerCounts, k, (gint)kmerCounts.size() / 100);
    for (gint x : top) values[(size_t)x] = 0;
    values[0] = 0;
    return values;
}

// ---------------------------------------------------------------------------
// mapping/mapping.go
// ---------------------------------------------------------------------------
void NewMapper(Mapper& m, const PackedSeq& reference, bool circular, gint k, const double* kmerValues, gint seedRate,
               gint edgeSize, gint chunkSize, int lean, int threads) {  // mapping.go:67-109
    NewSeedIndex(m.index, k);
    m.reference = reference;
    m.edgeSize = edgeSize;
    m.circular = circular;
    AddSingleSeeds(m.index, reference, seedRate, kmerValues);
    // the chunks in the producer's emission order (mapping.go:80-95)
    std::vector<PackedSeq> chunks;
    for (gint j = 0; j < 10; j++) {
        gint start = j * chunkSize;
        gint step = chunkSize * 10 - edgeSize;
        for (gint i = start; i < reference.Len() - chunkSize / 2; i += step) {
            gint end = i + chunkSize;
            if (i >= reference.Len()) end = reference.Len();
            chunks.push_back(SubSequence(reference, i, end));
        }
    }
    if (circular) {
        chunks.push_back(Append(SubSequence(reference, reference.Len() - edgeSize, reference.Len()), 0,
                                SubSequence(reference, 0, edgeSize)));
    }
    // Memory-lean mode (oracle.hpp): decided by what the two bitset families would cost (#seeds x #chunks bits each)
    if (lean < 0) {
        const char* env = getenv("DPO_LEAN_BYTES");
        const double limit = env ? atof(env) : 4e9;
        lean = 2.0 * (double)m.index.size * (double)chunks.size() / 8.0 > limit ? 1 : 0;
    }
    m.index.lean = lean != 0;
    // NewSeedSequence per chunk only reads the index's k-mer tables: independent calls, stored in emission order
    // (the reference's goroutines deliver them in arrival order and ids follow, Q5: canonical = emission order)
    const size_t C = chunks.size();
    std::vector<SeedSequence> built(C);
    if (m.index.lean) m.index.leanSegments.resize(C);
    if (threads < 1) threads = 1;
    std::atomic<size_t> next(0);
    std::vector<std::string> errs((size_t)threads);
    auto work = [&](int t) {
        try {
            for (;;) {
                size_t c0 = next.fetch_add(16);
                if (c0 >= C) break;
                for (size_t c = c0; c < std::min(C, c0 + 16); c++) {
                    built[c] = NewSeedSequence(m.index, chunks[c], nullptr);
                    built[c].id = (gint)c;  // seq.SetID(ind)
                    if (m.index.lean) {     // (16 -> 4 bytes per value at once: the transient footprint stays small)
                        std::vector<int32_t> seg(built[c].segments.begin(), built[c].segments.end());
                        for (size_t z = 0; z < seg.size(); z++)
                            if ((gint)seg[z] != built[c].segments[z]) throw std::runtime_error("oracle: lean segment out of range");
                        m.index.leanSegments[c] = std::move(seg);
                        std::vector<gint>().swap(built[c].segments);
                    }
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
    for (size_t c = 0; c < C; c++) {
        if (m.index.lean) m.index.sequences.push_back(std::move(built[c]));  // (segments already in leanSegments[c])
        else AddSequence(m.index, std::move(built[c]));
    }
    IndexSequences(m.index);
}

std::string AsString(const Mapper& m, const Mapping& mp, const std::string& qname) {  // mapping.go:112-122
    const char* rc = mp.RC ? "-" : "+";
    gint mappedLength = mp.End - mp.Start;
    if (m.circular && mappedLength < 0) mappedLength = m.reference.Len() - mp.Start + mp.End;
    std::string s;
    s.reserve(128);
    s += qname;
    s += '\t';
    s += std::to_string(mp.queryLen);
    s += '\t';
    s += std::to_string(mp.QueryOffset);
    s += '\t';
    s += std::to_string(mp.queryLen - mp.QueryInset);
    s += '\t';
    s += rc;
    s += '\t';
    s += m.refName;
    s += '\t';
    s += std::to_string(m.reference.Len());
    s += '\t';
    s += std::to_string(mp.Start);
    s += '\t';
    s += std::to_string(mp.End);
    s += '\t';
    s += std::to_string(mp.ids);
    s += '\t';
    s += std::to_string(mappedLength);
    s += "\t255";
    return s;
}

namespace {

typedef std::vector<Mapping*> MList;  // a Go []*Mapping; no two live slices alias in the reference's flow

struct MapCtx {
    const Mapper& m;
    Counters* c;
    std::deque<Mapping> arena;
    MapCtx(const Mapper& mm, Counters* cc) : m(mm), c(cc) {}
    Mapping* make() {
        arena.emplace_back();
        arena.back().queryLen = -1;  // Query == nil until updateQuery
        return &arena.back();
    }
};

// sort.Sort replacement: stable; counts sorts whose Go result is version dependent (n > 12 with ties). H2.
template <class Key>
void go_sort(MList& v, Key key, Counters* c) {
    if (v.size() > 12 && c) {
        std::vector<gint> ks;
        for (Mapping* p : v) ks.push_back(key(p));
        std::sort(ks.begin(), ks.end());
        if (std::adjacent_find(ks.begin(), ks.end()) != ks.end()) c->sort_ties_unpinned++;
    }
    std::stable_sort(v.begin(), v.end(), [&](Mapping* a, Mapping* b) { return key(a) < key(b); });
}

void updateQuery(MList& ms, gint qlen) {  // mapping.go:124-128
    for (Mapping* mp : ms) mp->queryLen = qlen;
}

gint qlen_of(const Mapping* mp) {
    if (mp->queryLen < 0) throw std::runtime_error("oracle: nil Mapping.Query dereference (Go would panic)");
    return mp->queryLen;
}

bool isConsistent(MapCtx& x, const Mapping* left, const Mapping* right) {  // mapping.go:131-160
    if (left->RC != right->RC) return false;
    gint expectedDistance = right->QueryOffset - qlen_of(left) + left->QueryInset;
    gint distance;
    if (!left->RC) distance = right->Start - left->End;
    else distance = left->Start - right->End;
    if (x.m.circular && distance < -50) distance += x.m.reference.Len();
    if (distance < 50 && expectedDistance < 50 && distance > -50) return true;
    if (distance < 500) return (expectedDistance < (distance * 3) / 2 && expectedDistance > (distance * 2) / 3);
    if (distance > 5000) return (expectedDistance < (distance * 10) / 9 && expectedDistance > (distance * 9) / 10);
    // Go folds 3.0/2.0 and (10.0/9.0-3.0/2.0) exactly: 1.5 and the double nearest -7/18. No FMA on amd64 (v1).
    volatile double ratio = (double)(distance - 500) / 4500.0;
    volatile double prod = ratio * (-7.0 / 18.0);
    ratio = 1.5 + prod;
    volatile double a = (double)expectedDistance * ratio;
    volatile double b = (double)expectedDistance / ratio;
    return distance < (gint)a && distance > (gint)b;
}

// mapping.go:387-428. All call sites pass the same slice as `open` and `extended`.
MList removeDominated(MList open, gint queryLen, Counters* c) {
    if (open.empty()) return open;
    go_sort(open, [](Mapping* p) { return p->QueryOffset; }, c);
    const MList& extended = open;
    size_t j = 0;
    std::vector<char> toRemove(open.size(), 0);
    for (size_t i = 0; i < open.size(); i++) {
        Mapping* next = open[i];
        while (j < extended.size() && queryLen - extended[j]->QueryInset < next->QueryOffset) j++;
        if (j == extended.size()) return open;
        bool dominated = false;
        for (size_t kk = j; !dominated && kk < extended.size() && extended[kk]->QueryOffset < queryLen - next->QueryInset;
             kk++) {
            if (extended[kk]->ids * 4 > next->ids * 5) {
                gint start = next->QueryOffset;
                if (extended[kk]->QueryOffset > start) start = extended[kk]->QueryOffset;
                gint end = queryLen - next->QueryInset;
                if (extended[kk]->QueryInset > next->QueryInset) end = queryLen - extended[kk]->QueryInset;
                dominated = ((end - start) * 10 > (queryLen - next->QueryOffset - next->QueryInset) * 9);
            }
        }
        toRemove[i] = dominated;
    }
    gint last = (gint)open.size() - 1;
    for (gint i = last; i >= 0; i--) {
        if (toRemove[(size_t)i]) {
            open[(size_t)i] = open[(size_t)last];
            last--;
        }
    }
    open.resize((size_t)(last + 1));
    return open;
}

MList performMapping(MapCtx& x, const PackedSeq& query) {  // mapping.go:489-611
    const Mapper& m = x.m;
    const SeedIndex& index = m.index;
    gint k = index.seedSize;
    if (x.c) x.c->windows++;
    SeedSequence seedQuery = NewSeedSequence(index, query, x.c);
    SeedSequence rcQuery = NewSeedSequence(index, ReverseComplement(query), x.c);

    gint minMatches = seedQuery.GetNumSeeds() / 5;
    gint minRCMatches = rcQuery.GetNumSeeds() / 5;
    if (minMatches < 5) minMatches = 5;
    if (minRCMatches < 5) minRCMatches = 5;
    std::vector<uint64_t> matchingIndices = Matches(index, seedQuery, 0.25, x.c);
    std::vector<uint64_t> matchingRCIndices = Matches(index, rcQuery, 0.25, x.c);
    if (x.c) x.c->candidates += (long long)(matchingIndices.size() + matchingRCIndices.size());
    MList results;

    gint maxSeed = 0;
    for (gint i = 0; i < seedQuery.GetNumSeeds(); i++) {
        gint s = seedQuery.GetSeed(i);
        if (s > maxSeed) maxSeed = s;
    }
    IntSet seedSet = NewIntSetCapacity(maxSeed + 1);
    for (gint i = 0; i < seedQuery.GetNumSeeds(); i++) Add(seedSet, (uint64_t)seedQuery.GetSeed(i));
    set_match_counters(x.c);
    for (uint64_t idx : matchingIndices) {
        IntSet leanSet;
        if (index.lean) leanSet = ChunkSeedSet(index, (size_t)idx);
        const IntSet& matchSet = index.lean ? leanSet : index.seedSets[idx];
        if (CountIntersectionTo(matchSet, seedSet, minMatches) < (uint64_t)minMatches) continue;
        if (x.c) x.c->cand_pass++;
        SeedSequence leanSeq;
        if (index.lean) leanSeq = ChunkSequence(index, (size_t)idx);
        const SeedSequence& match = index.lean ? leanSeq : index.sequences[idx];
        bool isnil = false;
        std::vector<SeedMatch> seedMatches = Match(match, seedQuery, seedSet, matchSet, minMatches, k, &isnil);
        for (const SeedMatch& seedMatch : seedMatches) {
            if (x.c) x.c->chains++;
            gint start = match.offset + GetSeedOffset(match, seedMatch.MatchB[0], k);
            gint end = m.reference.Len() - match.inset - GetSeedOffsetFromEnd(match, seedMatch.MatchB.back(), k);
            if (m.circular && start > m.reference.Len()) start -= m.reference.Len();
            gint qOffset = GetSeedOffset(seedQuery, seedMatch.MatchA[0], k);
            gint qInset = GetSeedOffsetFromEnd(seedQuery, seedMatch.MatchA.back(), k);
            if (qOffset + qInset > (seedQuery.Len() * 2) / 3) continue;
            qOffset += seedQuery.offset;
            qInset += seedQuery.inset;
            gint cA, ids;
            GetBasesCovered(seedMatch, k, &cA, &ids);
            Mapping* mp = x.make();
            mp->Start = start;
            mp->End = end;
            mp->QueryOffset = qOffset;
            mp->QueryInset = qInset;
            mp->RC = false;
            mp->ids = ids;
            results.push_back(mp);
            gint limit = ((gint)seedMatch.MatchA.size() * 4) / 5;
            if (limit > minMatches) minMatches = limit;
            if (limit > minRCMatches) minRCMatches = limit;
        }
    }
    Clear(seedSet);
    for (gint i = 0; i < rcQuery.GetNumSeeds(); i++) Add(seedSet, (uint64_t)rcQuery.GetSeed(i));
    for (uint64_t idx : matchingRCIndices) {
        IntSet leanSet;
        if (index.lean) leanSet = ChunkSeedSet(index, (size_t)idx);
        const IntSet& matchSet = index.lean ? leanSet : index.seedSets[idx];
        if (CountIntersectionTo(matchSet, seedSet, minRCMatches) < (uint64_t)minRCMatches) continue;
        if (x.c) x.c->cand_pass++;
        SeedSequence leanSeq;
        if (index.lean) leanSeq = ChunkSequence(index, (size_t)idx);
        const SeedSequence& match = index.lean ? leanSeq : index.sequences[idx];
        bool isnil = false;
        std::vector<SeedMatch> seedMatches = Match(match, rcQuery, seedSet, matchSet, minRCMatches, k, &isnil);
        for (const SeedMatch& seedMatch : seedMatches) {
            if (x.c) x.c->chains++;
            gint start = match.offset + GetSeedOffset(match, seedMatch.MatchB[0], k);
            gint end = m.reference.Len() - match.inset - GetSeedOffsetFromEnd(match, seedMatch.MatchB.back(), k);
            if (m.circular && start > m.reference.Len()) start -= m.reference.Len();
            gint qInset = GetSeedOffset(rcQuery, seedMatch.MatchA[0], k);
            gint qOffset = GetSeedOffsetFromEnd(rcQuery, seedMatch.MatchA.back(), k);
            if (qOffset + qInset > (rcQuery.Len() * 2) / 3) continue;
            qInset += rcQuery.offset;
            qOffset += rcQuery.inset;
            gint cA, ids;
            GetBasesCovered(seedMatch, k, &cA, &ids);
            Mapping* mp = x.make();
            mp->Start = start;
            mp->End = end;
            mp->QueryOffset = qOffset;
            mp->QueryInset = qInset;
            mp->RC = true;
            mp->ids = ids;
            results.push_back(mp);
            gint limit = ((gint)seedMatch.MatchA.size() * 4) / 5;
            if (limit > minRCMatches) minRCMatches = limit;
        }
    }
    set_match_counters(nullptr);
    if (results.size() > 1) {
        go_sort(results, [](Mapping* p) { return p->Start; }, x.c);
        for (gint i = (gint)results.size() - 1; i > 0; i--) {
            Mapping* ra = results[(size_t)i - 1];
            Mapping* rb = results[(size_t)i];
            if (ra->RC == rb->RC && rb->Start < ra->End) {
                if (ra->End - ra->Start > rb->End - rb->Start) {
                    results[(size_t)i] = results[results.size() - 1];
                    results.pop_back();
                } else {
                    results[(size_t)i - 1] = results[(size_t)i];
                    results[(size_t)i] = results[results.size() - 1];
                    results.pop_back();
                }
            }
        }
    }
    return results;
}

struct Pairs {
    MList remainingA, remainingB, matched;
    bool matchedNil = true;
};

Pairs matchPairs(MapCtx& x, MList openA, MList openB) {  // mapping.go:174-203
    Pairs r;
    for (gint i = (gint)openA.size() - 1; i >= 0; i--) {
        Mapping* ra = openA[(size_t)i];
        for (gint j = (gint)openB.size() - 1; j >= 0; j--) {
            Mapping* rb = openB[(size_t)j];
            if (isConsistent(x, ra, rb)) {
                gint qOffset = ra->QueryOffset;
                gint qInset = rb->QueryInset;
                if (ra->RC) std::swap(ra, rb);
                Mapping* combined = x.make();
                combined->Start = ra->Start;
                combined->End = rb->End;
                combined->queryLen = ra->queryLen;
                combined->QueryOffset = qOffset;
                combined->QueryInset = qInset;
                combined->RC = ra->RC;
                combined->ids = ra->ids + rb->ids;
                r.matchedNil = false;
                r.matched.push_back(combined);
                if (ra->RC) std::swap(ra, rb);
                openA[(size_t)i] = openA[openA.size() - 1];
                openA.pop_back();
                openB[(size_t)j] = openB[openB.size() - 1];
                openB.pop_back();
                break;
            }
        }
    }
    r.remainingA = std::move(openA);
    r.remainingB = std::move(openB);
    return r;
}

void append(MList& a, const MList& b) { a.insert(a.end(), b.begin(), b.end()); }

PackedSeq sub(const PackedSeq& q, gint start, gint end) {
    if (start < 0 || start >= end) throw std::runtime_error("oracle: bad query window (Go would panic or misbehave)");
    return SubSequence(q, start, end);
}

Pairs mapEnds(MapCtx& x, const PackedSeq& query) {  // mapping.go:164-172
    gint e = x.m.edgeSize;
    MList openA = performMapping(x, sub(query, 0, e));
    MList openB = performMapping(x, sub(query, query.Len() - e, query.Len()));
    openA = removeDominated(openA, query.Len(), x.c);
    openB = removeDominated(openB, query.Len(), x.c);
    updateQuery(openA, query.Len());
    updateQuery(openB, query.Len());
    return matchPairs(x, openA, openB);
}

// mapping.go:207-288. openA/openB are only read as lists; the Mappings they point to are updated in place.
void findSplitPoint(MapCtx& x, const PackedSeq& query, const MList& openA, const MList& openB, gint left, gint right) {
    gint e = x.m.edgeSize;
    while (right - left >= e) {
        gint start = (right + left - e) / 2;
        gint end = start + e;
        MList mid = performMapping(x, sub(query, start, end));
        gint newLeft = left;
        gint newRight = right;
        gint afterA = 0;
        gint afterB = 0;
        for (Mapping* mm : mid) {
            mm->queryLen = query.Len();
            for (Mapping* ma : openA) {
                if (isConsistent(x, ma, mm)) {
                    ma->QueryInset = mm->QueryInset;
                    ma->ids += mm->ids;
                    if (ma->RC) ma->Start = mm->Start;
                    else ma->End = mm->End;
                    gint midMatched = query.Len() - mm->QueryInset - mm->QueryOffset;
                    if (midMatched > afterA) afterA = midMatched;
                    if (query.Len() - mm->QueryInset > newLeft) newLeft = query.Len() - mm->QueryInset;
                    break;
                }
            }
            if (afterA < (e * 2) / 3) {
                for (Mapping* mb : openB) {
                    if (isConsistent(x, mm, mb)) {
                        mb->QueryOffset = mm->QueryOffset;
                        mb->ids += mm->ids;
                        if (mb->RC) mb->End = mm->End;
                        else mb->Start = mm->Start;
                        gint midMatched = query.Len() - mm->QueryInset - mm->QueryOffset;
                        if (midMatched > afterB) afterB = midMatched;
                        if (mm->QueryOffset < newRight) newRight = mm->QueryOffset;
                        break;
                    }
                }
            }
        }
        if (afterA > 0 && afterB > 0) {
            MList empty;
            if (newLeft - left > e * 2) findSplitPoint(x, query, openA, empty, newLeft - e * 2, newLeft - e);
            if (right - newRight > e * 2) findSplitPoint(x, query, empty, openB, newRight + e, newRight + e * 2);
            return;
        }
        if (afterA == 0 && afterB == 0) {
            MList empty;
            if (!openA.empty()) findSplitPoint(x, query, openA, empty, left, start);
            if (!openB.empty()) findSplitPoint(x, query, empty, openB, end, right);
            return;
        }
        left = newLeft;
        right = newRight;
    }
}

struct Next {
    MList newA, newB, matched;
    bool matchedNil = true;
};

Next mapNext(MapCtx& x, const PackedSeq& query, MList openA, MList openB) {  // mapping.go:305-383
    gint e = x.m.edgeSize;
    gint qlen = query.Len();
    Next out;
    MList newA, newB;
    if (qlen < e * 4) {
        newA = performMapping(x, sub(query, e, qlen - e));
        newA = removeDominated(newA, qlen, x.c);
        updateQuery(newA, qlen);
        Pairs p = matchPairs(x, openA, newA);
        openA = p.remainingA;
        newA = p.remainingB;
        if (!p.matchedNil) {
            openA = newA;  // append(newA, extended...)
            append(openA, p.matched);
        } else {
            append(openA, newA);
        }
        Pairs p2 = matchPairs(x, openA, openB);
        if (p2.matchedNil) {
            out.newA = p2.remainingA;
            out.newB = p2.remainingB;
            out.matchedNil = true;
            return out;
        }
        out.matched = p2.matched;
        out.matchedNil = false;
        return out;  // newA[:0], newB[:0], matched
    }
    // 1.
    newA = performMapping(x, sub(query, e, e * 2));
    newA = removeDominated(newA, qlen, x.c);
    updateQuery(newA, qlen);
    {
        Pairs p = matchPairs(x, openA, newA);
        openA = p.remainingA;
        newA = p.remainingB;
        append(openA, newA);
        if (!p.matchedNil) append(openA, p.matched);
    }
    newB = performMapping(x, sub(query, qlen - e * 2, qlen - e));
    newB = removeDominated(newB, qlen, x.c);
    updateQuery(newB, qlen);
    {
        Pairs p = matchPairs(x, newB, openB);  // openB, newB, extended = matchPairs(newB, openB)
        openB = p.remainingA;
        newB = p.remainingB;
        append(openB, newB);
        if (!p.matchedNil) append(openB, p.matched);
    }
    Pairs pm = matchPairs(x, openA, openB);
    newA = pm.remainingA;
    newB = pm.remainingB;
    MList matched = pm.matched;
    bool matchedNil = pm.matchedNil;
    // 2.
    if (matchedNil) {
        if (qlen > e * 5) {
            openA = performMapping(x, sub(query, e * 2, e * 3));
            openA = removeDominated(openA, qlen, x.c);
            updateQuery(openA, qlen);
            Pairs p = matchPairs(x, newA, openA);  // openA, newA, extended = matchPairs(newA, openA)
            openA = p.remainingA;
            newA = p.remainingB;
            if (!p.matchedNil) append(openA, p.matched);
            append(openA, newA);
        }
        if (qlen > e * 6) {
            openB = performMapping(x, sub(query, qlen - e * 3, qlen - e * 2));
            openB = removeDominated(openB, qlen, x.c);
            updateQuery(openB, qlen);
            Pairs p = matchPairs(x, openB, newB);
            openB = p.remainingA;
            newB = p.remainingB;
            if (!p.matchedNil) append(openB, p.matched);
            append(openB, newB);
        } else {
            openB = newB;
        }
        if (qlen > e * 5) {
            Pairs p = matchPairs(x, openA, openB);
            newA = p.remainingA;
            newB = p.remainingB;
            matched = p.matched;
            matchedNil = p.matchedNil;
        }
    }
    out.newA = newA;
    out.newB = newB;
    out.matched = matched;
    out.matchedNil = matchedNil;
    return out;
}

}  // namespace

// mapEnds' pairing step (mapping.go:167-171) on explicit hit lists: removeDominated on both, updateQuery, matchPairs.
// For the hand-worked vectors of tests/test_oracle_handworked.py (only refLen and circular of the mapper are used).
void PairEndsPublic(gint refLen, bool circular, gint queryLen, const std::vector<Mapping>& hitsA,
                    const std::vector<Mapping>& hitsB, std::vector<Mapping>* remA, std::vector<Mapping>* remB,
                    std::vector<Mapping>* matched, bool* matchedNil) {
    Mapper m;
    m.circular = circular;
    m.reference.length = refLen;
    MapCtx x(m, nullptr);
    MList openA, openB;
    for (const Mapping& h : hitsA) {
        Mapping* p = x.make();
        *p = h;
        p->queryLen = -1;
        openA.push_back(p);
    }
    for (const Mapping& h : hitsB) {
        Mapping* p = x.make();
        *p = h;
        p->queryLen = -1;
        openB.push_back(p);
    }
    openA = removeDominated(openA, queryLen, nullptr);
    openB = removeDominated(openB, queryLen, nullptr);
    updateQuery(openA, queryLen);
    updateQuery(openB, queryLen);
    Pairs r = matchPairs(x, openA, openB);
    for (Mapping* p : r.remainingA) remA->push_back(*p);
    for (Mapping* p : r.remainingB) remB->push_back(*p);
    for (Mapping* p : r.matched) matched->push_back(*p);
    *matchedNil = r.matchedNil;
}

std::vector<Mapping> performMappingPublic(const Mapper& m, const PackedSeq& query, Counters* c) {
    MapCtx x(m, c);
    MList r = performMapping(x, query);
    std::vector<Mapping> out;
    for (Mapping* p : r) out.push_back(*p);
    return out;
}

std::vector<Mapping> Map(const Mapper& m, const PackedSeq& query, Counters* c) {  // mapping.go:430-487
    MapCtx x(m, c);
    gint e = m.edgeSize;
    gint qlen = query.Len();
    MList results;
    if (qlen <= e * 2) {
        results = performMapping(x, query);
        results = removeDominated(results, qlen, c);
        updateQuery(results, qlen);
    } else {
        Pairs ends = mapEnds(x, query);
        MList openA = ends.remainingA, openB = ends.remainingB;
        if (!ends.matchedNil) {
            results = ends.matched;
        } else if (qlen < e * 3) {
            results = openA;
            append(results, openB);
        } else {
            Next nx = mapNext(x, query, openA, openB);
            openA = nx.newA;
            openB = nx.newB;
            if (!nx.matchedNil) {
                results = nx.matched;
            } else {
                gint left = e * 2;
                gint right = qlen - e * 2;
                for (Mapping* a : openA)
                    if (a->QueryInset > left) left = a->QueryInset;
                left = qlen - right;  // Q8: overwrites the value computed above
                for (Mapping* b : openB)
                    if (b->QueryOffset < right) right = b->QueryOffset;
                findSplitPoint(x, query, openA, openB, left, right);
                gint size = qlen - e;
                for (gint i = (gint)openA.size() - 1; i >= 0; i--) {
                    if (openA[(size_t)i]->QueryInset >= size) {
                        openA[(size_t)i] = openA[openA.size() - 1];
                        openA.pop_back();
                    }
                }
                for (gint i = (gint)openB.size() - 1; i >= 0; i--) {
                    if (openB[(size_t)i]->QueryOffset >= size) {
                        openB[(size_t)i] = openB[openB.size() - 1];
                        openB.pop_back();
                    }
                }
                results = openA;
                append(results, openB);
            }
        }
    }
    std::vector<Mapping> out;
    out.reserve(results.size());
    for (Mapping* p : results) out.push_back(*p);
    if (c) c->mappings += (long long)out.size();
    return out;
}

}  // namespace dpo
