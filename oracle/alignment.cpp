// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  Round-2 groundwork for the overlap path: a literal restatement
// of seeds/alignment.go:274-616 (seedAligner, PairwiseAlignments).  Go pointers into the state pool become indices;
// every slice access that Go would bounds-check throws "Go would panic" here.  The code is transcribed statement by
// statement, including the two removeOpenState calls whose arguments sit in the wrong slots (alignment.go:490,497
// against the signature at :390) and the `break searchMatch` that ends the scan of the open list after one extension.
#include "oracle.hpp"

#include <stdexcept>

namespace dpo {

namespace {
[[noreturn]] void go_panic(const char* what) { throw std::runtime_error(std::string("oracle: Go would panic: ") + what); }
}  // namespace

SeedAligner NewSeedAligner(gint maxLength) {  // alignment.go:298-306
    SeedAligner sa;
    sa.stackPool.assign(10000, PairState());
    sa.statesStack.assign(10000, -1);
    sa.reduced.assign((size_t)maxLength, 0);
    sa.open.assign(500, -1);
    sa.initials.assign((size_t)maxLength, -1);
    sa.results.assign(500, -1);
    sa.nextState = (gint)sa.statesStack.size() - 1;
    sa.aMapping.assign(sa.reduced.size() / 2, 0);
    for (size_t i = 0; i < sa.stackPool.size(); i++) sa.statesStack[i] = (int)i;
    return sa;
}

static int popState(SeedAligner& al) {  // :308-313
    if (al.nextState < 0 || al.nextState >= (gint)al.statesStack.size()) go_panic("statesStack index");
    int s = al.statesStack[(size_t)al.nextState];
    al.stackPool[(size_t)s].stackIndex = al.nextState;
    al.nextState--;
    return s;
}

static void pushState(SeedAligner& al, int s) {  // :315-324
    gint n = al.nextState + 1;
    if (n < 0 || n >= (gint)al.statesStack.size()) go_panic("statesStack index");
    int top = al.statesStack[(size_t)n];
    al.statesStack[(size_t)n] = s;
    gint si = al.stackPool[(size_t)s].stackIndex;
    if (si < 0 || si >= (gint)al.statesStack.size()) go_panic("statesStack index");
    al.statesStack[(size_t)si] = top;
    al.stackPool[(size_t)top].stackIndex = si;
    al.stackPool[(size_t)s].stackIndex = n;
    al.nextState = n;
}

static SeedMatch extractMatch(const SeedAligner& al, int s, const std::vector<gint>& aMapping, gint aMapLen) {  // :326-335
    gint len = al.stackPool[(size_t)s].length;
    SeedMatch m;
    m.MatchA.assign((size_t)len, 0);
    m.MatchB.assign((size_t)len, 0);
    while (s != -1) {
        const PairState& st = al.stackPool[(size_t)s];
        if (st.length - 1 < 0 || st.length - 1 >= len) go_panic("extractMatch index");
        if (st.aPos / 2 < 0 || st.aPos / 2 >= aMapLen) go_panic("aMapping index");
        m.MatchA[(size_t)(st.length - 1)] = aMapping[(size_t)(st.aPos / 2)];
        m.MatchB[(size_t)(st.length - 1)] = st.bPos / 2;
        s = st.prev;
    }
    return m;
}

// :341-388. Returns startSize; aLen through *aLenOut (aRed = reduced[:aLen*2+1], aMapping = aMapping[:aLen]).
static gint prepareInitial(SeedAligner& al, const std::vector<gint>& aSegments, const IntSet& bSet, gint minMatches, gint k,
                           gint* aLenOut) {
    gint maxAIndex = (gint)aSegments.size() - minMatches * 2 + 1;
    gint aLen = 0;
    gint offset = -k;
    gint startSize = 0;
    std::vector<gint>& aRed = al.reduced;
    std::vector<gint>& aMapping = al.aMapping;
    gint prevSeed = -1;
    const gint n = (gint)aSegments.size();
    for (gint i = 1; i < n; i += 2) {
        gint aSeed = aSegments[(size_t)i];
        if (!Contains(bSet, (uint64_t)aSeed)) {
            offset += aSegments[(size_t)(i - 1)] + k;
            maxAIndex--;
            continue;
        }
        if (aSeed == prevSeed && (i >= n - 2 || aSegments[(size_t)(i + 2)] == prevSeed)) {
            offset += aSegments[(size_t)(i - 1)] + k;
            maxAIndex--;
            continue;
        }
        prevSeed = aSeed;
        offset += aSegments[(size_t)(i - 1)] + k;
        if (aLen * 2 + 1 >= (gint)aRed.size()) go_panic("reduced index");
        aRed[(size_t)(aLen * 2)] = offset;
        aRed[(size_t)(aLen * 2 + 1)] = aSeed;
        if (aLen >= (gint)aMapping.size()) go_panic("aMapping index");
        aMapping[(size_t)aLen] = i / 2;
        offset = -k;
        if (aLen <= maxAIndex) {
            int state = popState(al);
            PairState& st = al.stackPool[(size_t)state];
            st.aPos = aLen * 2 + 1;
            st.length = 0;
            st.prev = -1;
            if (aLen >= (gint)al.initials.size()) go_panic("initials index");
            al.initials[(size_t)aLen] = state;
            startSize++;
        }
        aLen++;
    }
    if (aLen * 2 >= (gint)aRed.size()) go_panic("reduced index");
    aRed[(size_t)(aLen * 2)] = 0;
    while (startSize > 0 && al.stackPool[(size_t)al.initials[(size_t)(startSize - 1)]].aPos > maxAIndex) {
        startSize--;
        pushState(al, al.initials[(size_t)startSize]);
    }
    *aLenOut = aLen;
    return startSize;
}

// :390-409 — (openSize, resultsSize, minMatches) are returned through the references, in that order
static void removeOpenState(SeedAligner& al, gint index, gint minMatches, gint openSize, gint resultsSize, gint& outOpen,
                            gint& outResults, gint& outMin) {
    if (index < 0 || index >= (gint)al.open.size() || openSize - 1 < 0 || openSize - 1 >= (gint)al.open.size())
        go_panic("open index");
    int s = al.open[(size_t)index];
    al.open[(size_t)index] = al.open[(size_t)(openSize - 1)];
    openSize--;
    if (s == -1) go_panic("nil state");
    if (al.stackPool[(size_t)s].length >= minMatches) {
        gint len = al.stackPool[(size_t)s].length;
        if ((len * 2) / 3 > minMatches) minMatches = (len * 2) / 3;
        if (resultsSize < 0 || resultsSize >= (gint)al.results.size()) go_panic("results index");
        al.results[(size_t)resultsSize] = s;
        resultsSize++;
    } else {
        while (s != -1) {
            pushState(al, s);
            s = al.stackPool[(size_t)s].prev;
        }
    }
    outOpen = openSize;
    outResults = resultsSize;
    outMin = minMatches;
}

gint gapRangeMin(gint gap, gint k) {  // :411-424, first result
    gint minGap = (gap * 2) / 3 - k;
    gint maxGap = (gap * 3) / 2 + k + 1;
    if (minGap < 0) minGap = -k;
    else if (maxGap < 20) minGap = 0;
    return minGap;
}
gint gapRangeMax(gint gap, gint k) {  // :411-424, second result
    gint minGap = (gap * 2) / 3 - k;
    gint maxGap = (gap * 3) / 2 + k + 1;
    if (minGap < 0) {
        if (maxGap < 0) maxGap = 0;
    } else if (maxGap < 20) {
        maxGap = 20;
    }
    return maxGap;
}

std::vector<SeedMatch> PairwiseAlignments(SeedAligner& al, const SeedSequence& a, const SeedSequence& b, const IntSet& aSet,
                                          const IntSet& bSet, gint minMatches, gint k) {  // :426-616
    const std::vector<gint>& aSegments = a.segments;
    const std::vector<gint>& bSegments = b.segments;
    if (minMatches == 0) minMatches = 1;
    al.nextState = (gint)al.statesStack.size() - 1;  // reset(), :337-339
    gint aLen = 0;
    gint initialSize = prepareInitial(al, aSegments, bSet, minMatches, k, &aLen);
    const std::vector<gint>& aRed = al.reduced;
    const gint aRedLen = aLen * 2 + 1;  // len(aRed)
    auto AR = [&](gint i) -> gint {
        if (i < 0 || i >= aRedLen) go_panic("aRed index");
        return aRed[(size_t)i];
    };
    gint openSize = 0, resultsSize = 0;
    const gint bLen = (gint)bSegments.size();
    gint maxBIndex = bLen - minMatches * 2 + 1;
    gint bOffset = 0;
    gint prevSeed = -1;
    auto OPEN = [&](gint i) -> int& {
        if (i < 0 || i >= (gint)al.open.size()) go_panic("open index");
        return al.open[(size_t)i];
    };
    for (gint bIndex = 1; bIndex < bLen; bIndex += 2) {
        gint bSeed = bSegments[(size_t)bIndex];
        if (!Contains(aSet, (uint64_t)bSeed)) {
            bOffset += bSegments[(size_t)(bIndex + 1)] + k;
            continue;
        }
        if (bSeed == prevSeed && (bIndex >= bLen - 2 || bSegments[(size_t)(bIndex + 2)] == prevSeed)) {
            bOffset += bSegments[(size_t)(bIndex + 1)] + k;
            continue;
        }
        prevSeed = bSeed;
        gint found = -1, prevFound = -1;
        bool leftSearch = false;  // `break searchMatch`
        for (gint i = openSize - 1; i >= 0 && !leftSearch; i--) {
            int s = OPEN(i);
            if (s == -1) go_panic("nil state");
            PairState* S = &al.stackPool[(size_t)s];
            S->bGap += bOffset;
            gint minGap = gapRangeMin(S->bGap, k), maxGap = gapRangeMax(S->bGap, k);
            bool ended = false;
            while (S->aGap < minGap) {
                if (S->aGapIndex >= aRedLen) {
                    ended = true;
                    removeOpenState(al, i, minMatches, openSize, resultsSize, openSize, resultsSize, minMatches);
                    leftSearch = true;
                    break;
                }
                S->aGap += AR(S->aGapIndex + 1) + k;
                S->aGapIndex += 2;
            }
            if (leftSearch) break;
            if (!ended) {
                if (S->aGap <= maxGap) {
                    gint g = S->aGap;
                    for (gint j = S->aGapIndex; j < aRedLen && g <= maxGap; j += 2) {
                        if (AR(j) == bSeed) {
                            if (found != -1 && prevFound > i && prevFound < openSize) {
                                int s2 = OPEN(prevFound);
                                if (s2 == -1) go_panic("nil state");
                                const PairState& S2 = al.stackPool[(size_t)s2];
                                if (S->aPos == S2.aPos && S->bPos == S2.bPos) {
                                    gint d1, d2;
                                    if (S->length < S2.length) {
                                        // openSize, _, _ = removeOpenState(i, openSize, resultsSize, s.length+1)
                                        removeOpenState(al, i, openSize, resultsSize, S->length + 1, openSize, d1, d2);
                                        if (prevFound == openSize - 1) prevFound = i;
                                        leftSearch = true;
                                        break;
                                    } else {
                                        removeOpenState(al, prevFound, openSize, resultsSize, S2.length + 1, openSize, d1, d2);
                                    }
                                }
                            }
                            found = j;
                            prevFound = i;
                            int ns = popState(al);
                            S = &al.stackPool[(size_t)s];
                            PairState& NS = al.stackPool[(size_t)ns];
                            NS.prev = s;
                            NS.aPos = j;
                            NS.bPos = bIndex;
                            NS.aGapIndex = j + 2;
                            NS.aGap = AR(j + 1);
                            NS.bGap = bSegments[(size_t)(bIndex + 1)];
                            NS.length = S->length + 1;
                            OPEN(i) = ns;
                            if ((NS.length * 2) / 3 > minMatches) {
                                minMatches = (NS.length * 2) / 3;
                                maxBIndex = bLen - minMatches * 2 + 1;
                            }
                            leftSearch = true;
                            break;
                        }
                        g += AR(j + 1) + k;
                    }
                    if (leftSearch) break;
                }
                if (S->length + (bLen - bIndex) < minMatches) {
                    removeOpenState(al, i, minMatches, openSize, resultsSize, openSize, resultsSize, minMatches);
                } else {
                    S->bGap += bSegments[(size_t)(bIndex + 1)] + k;
                }
            }
        }
        bOffset = 0;
        if (bIndex <= maxBIndex) {
            for (gint i = 0; i < initialSize; i++) {
                int s = al.initials[(size_t)i];
                gint aPos = al.stackPool[(size_t)s].aPos;
                if (aPos != found && AR(aPos) == bSeed) {
                    if (found != -1) {
                        for (gint j = 0; j < openSize; j++) {
                            const PairState& O = al.stackPool[(size_t)OPEN(j)];
                            if (O.bPos == bIndex && O.aPos == aPos) {
                                found = aPos;
                                break;
                            }
                        }
                    }
                    if (found == aPos || openSize >= (gint)al.open.size()) continue;
                    int ns = popState(al);
                    PairState& NS = al.stackPool[(size_t)ns];
                    NS.aPos = aPos;
                    NS.bPos = bIndex;
                    NS.aGapIndex = aPos + 2;
                    NS.aGap = AR(aPos + 1);
                    NS.bGap = bSegments[(size_t)(bIndex + 1)];
                    NS.length = 1;
                    NS.prev = -1;
                    OPEN(openSize) = ns;
                    openSize++;
                }
            }
        }
    }
    for (gint i = 0; i < openSize; i++) {
        int s = OPEN(i);
        if (al.stackPool[(size_t)s].length >= minMatches) {
            if (resultsSize >= (gint)al.results.size()) go_panic("results index");
            al.results[(size_t)resultsSize] = s;
            resultsSize++;
        }
    }
    std::vector<SeedMatch> matches;
    for (gint i = resultsSize - 1; i >= 0; i--) {
        SeedMatch r = extractMatch(al, al.results[(size_t)i], al.aMapping, aLen);
        r.SeqA = &a;
        r.SeqB = &b;
        matches.push_back(std::move(r));
    }
    return matches;
}

}  // namespace dpo
