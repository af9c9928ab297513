// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// One round of `downpore overlap` (commands/overlap.go:115-160) up to and including the stream of seed matches that
// overlapper.FindOverlaps delivers: PrepareQueries (overlap/overlap.go:157-215, QueryEdges: getEdges :56-94),
// AddSequences (:218-251, chunkWorker :253-318) and matchWorker (:346-387), glued from the pieces restated in
// seeds.cpp / alignment.cpp. Where the reference leaves the order to the goroutine scheduler the canonical choice is
// num_workers = 1: reads in file order (AddSeeds sees the k-mer table as the previous slice left it, index.Size() is
// up to date at every `>= seedLimit` test), chunks numbered in emission order, queries in slice order with the forward
// query in front of its reverse complement, candidates in ascending chunk order.
// The sequence set is the cached one (himem, the default): every read reaches the round as
// cached[id].SubSequence(0, Len()) (sequence/seqio.go:118-125), never as the raw packed sequence.
#include "oracle.hpp"

#include <stdexcept>

namespace dpo {

void OverlapRoundRun(const std::vector<PackedSeq>& reads, const std::vector<uint8_t>& ignore, gint firstSequence,
                     const double* values, const OverlapParams& P, OverlapRound& out) {
    out = OverlapRound();
    SeedIndex& g = out.index;
    NewSeedIndex(g, P.k);
    const gint nReads = (gint)reads.size();
    // ---- PrepareQueries(numSeeds, seedBatchSize, values, GetNSequencesFrom(firstSequence, queryBatchSize), QueryEdges) ----
    std::vector<PackedSeq> cached;
    {
        gint sent = 0;
        for (gint id = firstSequence; id < nReads && sent < P.queryBatchSize; id++) {  // seqio.go:118-125
            if (ignore[(size_t)id]) continue;
            sent++;
            if (g.size >= P.seedBatchSize) break;  // overlap.go:59 (the rest of the channel is drained unseen)
            PackedSeq s = SubSequence(reads[(size_t)id], 0, reads[(size_t)id].Len());
            s.id = id;
            if (s.Len() < P.overlapSize * 2) {
                AddSeeds(g, s, P.numSeeds, values, nullptr);
                cached.push_back(s);
            } else {
                PackedSeq s1 = SubSequence(s, 0, P.overlapSize);
                PackedSeq s2 = SubSequence(s, s.Len() - P.overlapSize, s.Len());
                AddSeeds(g, s1, P.numSeeds, values, nullptr);
                AddSeeds(g, s2, P.numSeeds, values, nullptr);
                cached.push_back(s1);
                cached.push_back(s2);
            }
        }
    }
    gint queryID = 0;
    for (const PackedSeq& s : cached) {  // overlap.go:175-205
        OverlapQuery q;
        q.ID = queryID;
        q.SequenceID = s.id;
        q.Query = NewSeedSequence(g, s, nullptr);
        q.rc = false;
        OverlapQuery r;
        r.ID = queryID;
        r.SequenceID = s.id;
        r.Query = ReverseComplementSeq(q.Query, P.k, g);
        r.rc = true;
        queryID++;
        out.queries.push_back(std::move(q));
        out.queries.push_back(std::move(r));
    }
    out.numQuerySeqs = 0;
    out.nextFirstSequence = firstSequence;
    if (out.queries.empty()) return;  // commands/overlap.go:132-134: the command ends here
    out.nextFirstSequence = out.queries.back().SequenceID + 1;  // :137-145
    for (const OverlapQuery& q : out.queries) {
        if (q.ID >= out.numQuerySeqs) out.numQuerySeqs = q.ID + 1;
        if (q.SequenceID >= out.nextFirstSequence) out.nextFirstSequence = q.SequenceID + 1;
    }
    // ---- AddSequences(GetSequences()) ----
    for (gint id = 0; id < nReads; id++) {
        if (ignore[(size_t)id]) continue;
        PackedSeq s = SubSequence(reads[(size_t)id], 0, reads[(size_t)id].Len());
        s.id = id;
        SeedSequence ss = NewSeedSequence(g, s, nullptr);
        for (SeedSequence& piece : ChunkSeedSequence(ss, P.chunkSize, P.numSeeds, P.overlapSize, P.k)) AddSequence(g, std::move(piece));
    }
    IndexSequences(g);
    // ---- FindOverlaps: matchWorker over the queries in order ----
    SeedAligner aligner = NewSeedAligner(P.overlapSize / 2);
    for (const OverlapQuery& q : out.queries) {
        IntSet seedSet = NewIntSet();
        for (gint i = 0; i < q.Query.GetNumSeeds(); i++) Add(seedSet, (uint64_t)q.Query.GetSeed(i));
        std::vector<uint64_t> matches = Matches(g, q.Query, P.hitFraction, nullptr);
        gint minMatches = (gint)(P.hitFraction * (double)q.Query.GetNumSeeds() + 0.5);
        for (uint64_t match : matches) {
            const IntSet& matchSet = g.seedSets[(size_t)match];
            if (CountIntersectionTo(matchSet, seedSet, minMatches) < (uint64_t)minMatches) continue;
            const SeedSequence& m = g.sequences[(size_t)match];
            std::vector<SeedMatch> sMatches = PairwiseAlignments(aligner, q.Query, m, seedSet, matchSet, minMatches, P.k);
            if (sMatches.empty()) continue;  // nil
            const SeedMatch* best = nullptr;
            gint bestCount = 0;  // never updated (overlap.go:369-372): the last alignment with a non-zero cover wins
            for (const SeedMatch& sm : sMatches) {
                gint ca, cb;
                GetBasesCovered(sm, P.k, &ca, &cb);
                if (cb > bestCount) best = &sm;
            }
            if (!best) throw std::runtime_error("oracle: Go would panic: nil best match (overlap.go:374)");
            OverlapHit h;
            h.queryID = q.ID;
            h.rc = q.rc;
            h.target = (gint)match;
            h.MatchA = best->MatchA;
            h.MatchB = best->MatchB;
            out.hits.push_back(std::move(h));
            if ((gint)best->MatchA.size() * 2 > minMatches * 3) minMatches = ((gint)best->MatchA.size() * 2) / 3;
        }
    }
}

}  // namespace dpo
