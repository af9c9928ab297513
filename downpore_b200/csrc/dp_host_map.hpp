// downpore_b200 — host side of Mapper.Map (mapping/mapping.go:124-487): the per-read strategy that decides which
// windows of a read are queried and how the window hits are paired. The window queries themselves
// (performMapping, mapping.go:489-611) run on the GPU in batched rounds.
//
// Replay scheme: run_map() executes the reference's control flow for one read against a cache of window results.
// When it needs a window that has not been computed yet it records the request and unwinds; the driver computes all
// requested windows of all reads in one GPU round and replays the read. Map() only ever mutates its own copies of the
// hits, so replaying from the start is exact. Windows the flow is certain to need together are requested together.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "dp_common.cuh"

namespace dph {

struct Hit {  // mapping.Mapping (mapping.go:11-20)
    long long start, end;
    long long qOffset, qInset;
    long long ids;
    bool rc;
};

struct WinRef {  // one computed window of a read
    int start, len, whole;
    int n;
    const DpMappingDev* maps;
};

struct Params {
    long long refLen;
    int edge;
    bool circular;
};

typedef std::vector<int> List;  // []*Mapping as indices into the per-read pool

class ReadMapper {
   public:
    ReadMapper(const Params& p) : P(p) {}

    // Returns true when the read is finished (results filled); false when windows were requested.
    bool run(int readIndex, long long readLen, const WinRef* cached, int nCached, std::vector<DpWindow>& requests,
             std::vector<Hit>& results) {
        read_ = readIndex;
        qlen_ = readLen;
        cached_ = cached;
        nCached_ = nCached;
        req_ = &requests;
        missing_ = false;
        pool_.clear();
        List res;
        bool ok = map(res);
        if (!ok || missing_) return false;
        results.clear();
        for (int i : res) results.push_back(pool_[i]);
        return true;
    }

   private:
    const Params& P;
    int read_ = 0;
    long long qlen_ = 0;
    const WinRef* cached_ = nullptr;
    int nCached_ = 0;
    std::vector<DpWindow>* req_ = nullptr;
    bool missing_ = false;
    std::vector<Hit> pool_;

    // performMapping(query.SubSequence(start,end)) or performMapping(query): cache lookup or request
    bool perform(long long start, long long end, bool whole, List& out) {
        int len = (int)(end - start);
        for (int i = 0; i < nCached_; i++) {
            const WinRef& w = cached_[i];
            if (w.start == (int)start && w.len == len && w.whole == (int)whole) {
                out.clear();
                for (int j = 0; j < w.n; j++) {
                    const DpMappingDev& m = w.maps[j];
                    Hit h;
                    h.start = m.start;
                    h.end = m.end;
                    h.qOffset = m.qOffset;
                    h.qInset = m.qInset;
                    h.ids = m.ids;
                    h.rc = (m.rc & 0xff) != 0;
                    pool_.push_back(h);
                    out.push_back((int)pool_.size() - 1);
                }
                return true;
            }
        }
        DpWindow wd;
        wd.read = read_;
        wd.start = (int)start;
        wd.len = len;
        wd.whole = whole ? 1 : 0;
        req_->push_back(wd);
        missing_ = true;
        return false;
    }

    bool isConsistent(const Hit& left, const Hit& right) const {  // mapping.go:131-160
        if (left.rc != right.rc) return false;
        long long expectedDistance = right.qOffset - qlen_ + left.qInset;
        long long distance = !left.rc ? right.start - left.end : left.start - right.end;
        if (P.circular && distance < -50) distance += P.refLen;
        if (distance < 50 && expectedDistance < 50 && distance > -50) return true;
        if (distance < 500) return expectedDistance < (distance * 3) / 2 && expectedDistance > (distance * 2) / 3;
        if (distance > 5000) return expectedDistance < (distance * 10) / 9 && expectedDistance > (distance * 9) / 10;
        // 3.0/2.0 + ratio*(10.0/9.0-3.0/2.0): Go folds the constants exactly (1.5, nearest double to -7/18) and
        // rounds the product and the sum separately on amd64 (no FMA) — keep the two roundings.
        volatile double ratio = (double)(distance - 500) / 4500.0;
        volatile double prod = ratio * (-7.0 / 18.0);
        ratio = 1.5 + prod;
        volatile double a = (double)expectedDistance * ratio;
        volatile double b = (double)expectedDistance / ratio;
        return distance < (long long)a && distance > (long long)b;
    }

    // mapping.go:387-428 with open == extended (every call site)
    void removeDominated(List& open) {
        if (open.empty()) return;
        std::stable_sort(open.begin(), open.end(), [&](int a, int b) { return pool_[a].qOffset < pool_[b].qOffset; });
        const size_t n = open.size();
        size_t j = 0;
        std::vector<char> toRemove(n, 0);
        for (size_t i = 0; i < n; i++) {
            const Hit& next = pool_[open[i]];
            while (j < n && qlen_ - pool_[open[j]].qInset < next.qOffset) j++;
            if (j == n) return;
            bool dominated = false;
            for (size_t k = j; !dominated && k < n && pool_[open[k]].qOffset < qlen_ - next.qInset; k++) {
                const Hit& ex = pool_[open[k]];
                if (ex.ids * 4 > next.ids * 5) {
                    long long start = std::max(next.qOffset, ex.qOffset);
                    long long end = qlen_ - next.qInset;
                    if (ex.qInset > next.qInset) end = qlen_ - ex.qInset;
                    dominated = (end - start) * 10 > (qlen_ - next.qOffset - next.qInset) * 9;
                }
            }
            toRemove[i] = dominated;
        }
        long long last = (long long)n - 1;
        for (long long i = last; i >= 0; i--) {
            if (toRemove[(size_t)i]) {
                open[(size_t)i] = open[(size_t)last];
                last--;
            }
        }
        open.resize((size_t)(last + 1));
    }

    // mapping.go:174-203; returns true when `matched` is non-nil
    bool matchPairs(List& openA, List& openB, List& matched) {
        bool any = false;
        matched.clear();
        for (long long i = (long long)openA.size() - 1; i >= 0; i--) {
            for (long long j = (long long)openB.size() - 1; j >= 0; j--) {
                const Hit ra = pool_[openA[(size_t)i]];
                const Hit rb = pool_[openB[(size_t)j]];
                if (isConsistent(ra, rb)) {
                    const Hit& first = ra.rc ? rb : ra;   // `if ra.RC { ra, rb = rb, ra }`
                    const Hit& second = ra.rc ? ra : rb;
                    Hit c;
                    c.start = first.start;
                    c.end = second.end;
                    c.qOffset = ra.qOffset;
                    c.qInset = rb.qInset;
                    c.rc = first.rc;
                    c.ids = ra.ids + rb.ids;
                    pool_.push_back(c);
                    matched.push_back((int)pool_.size() - 1);
                    any = true;
                    openA[(size_t)i] = openA.back();
                    openA.pop_back();
                    openB[(size_t)j] = openB.back();
                    openB.pop_back();
                    break;
                }
            }
        }
        return any;
    }

    static void append(List& a, const List& b) { a.insert(a.end(), b.begin(), b.end()); }

    // mapping.go:207-288
    bool findSplitPoint(const List& openA, const List& openB, long long left, long long right) {
        const long long e = P.edge;
        while (right - left >= e) {
            long long start = (right + left - e) / 2;
            long long end = start + e;
            List mid;
            if (!perform(start, end, false, mid)) return false;
            long long newLeft = left, newRight = right, afterA = 0, afterB = 0;
            for (int mi : mid) {
                const Hit mm = pool_[mi];
                for (int ai : openA) {
                    Hit& ma = pool_[ai];
                    if (isConsistent(ma, mm)) {
                        ma.qInset = mm.qInset;
                        ma.ids += mm.ids;
                        if (ma.rc) ma.start = mm.start;
                        else ma.end = mm.end;
                        long long midMatched = qlen_ - mm.qInset - mm.qOffset;
                        if (midMatched > afterA) afterA = midMatched;
                        if (qlen_ - mm.qInset > newLeft) newLeft = qlen_ - mm.qInset;
                        break;
                    }
                }
                if (afterA < (e * 2) / 3) {
                    for (int bi : openB) {
                        Hit& mb = pool_[bi];
                        if (isConsistent(mm, mb)) {
                            mb.qOffset = mm.qOffset;
                            mb.ids += mm.ids;
                            if (mb.rc) mb.end = mm.end;
                            else mb.start = mm.start;
                            long long midMatched = qlen_ - mm.qInset - mm.qOffset;
                            if (midMatched > afterB) afterB = midMatched;
                            if (mm.qOffset < newRight) newRight = mm.qOffset;
                            break;
                        }
                    }
                }
            }
            if (afterA > 0 && afterB > 0) {
                List empty;
                if (newLeft - left > e * 2)
                    if (!findSplitPoint(openA, empty, newLeft - e * 2, newLeft - e)) return false;
                if (right - newRight > e * 2)
                    if (!findSplitPoint(empty, openB, newRight + e, newRight + e * 2)) return false;
                return true;
            }
            if (afterA == 0 && afterB == 0) {
                List empty;
                if (!openA.empty())
                    if (!findSplitPoint(openA, empty, left, start)) return false;
                if (!openB.empty())
                    if (!findSplitPoint(empty, openB, end, right)) return false;
                return true;
            }
            left = newLeft;
            right = newRight;
        }
        return true;
    }

    // mapping.go:305-383; returns false on a missing window. matchedAny <=> matched != nil
    bool mapNext(List& openA, List& openB, List& outA, List& outB, List& matched, bool& matchedAny) {
        const long long e = P.edge;
        List newA, newB, extended;
        if (qlen_ < e * 4) {
            if (!perform(e, qlen_ - e, false, newA)) return false;
            removeDominated(newA);
            bool ext = matchPairs(openA, newA, extended);
            if (ext) {
                openA = newA;
                append(openA, extended);
            } else {
                append(openA, newA);
            }
            matchedAny = matchPairs(openA, openB, matched);
            if (!matchedAny) {
                outA = openA;
                outB = openB;
            } else {
                outA.clear();
                outB.clear();
            }
            return true;
        }
        // both second-step windows are always needed: request them together
        bool okA = perform(e, e * 2, false, newA);
        bool okB = perform(qlen_ - e * 2, qlen_ - e, false, newB);
        if (!okA || !okB) return false;
        removeDominated(newA);
        {
            bool ext = matchPairs(openA, newA, extended);
            append(openA, newA);
            if (ext) append(openA, extended);
        }
        removeDominated(newB);
        {
            // openB, newB, extended = matchPairs(newB, openB)
            bool ext = matchPairs(newB, openB, extended);
            List remB = newB, remOld = openB;
            openB = remB;
            newB = remOld;
            append(openB, newB);
            if (ext) append(openB, extended);
        }
        matchedAny = matchPairs(openA, openB, matched);
        newA = openA;
        newB = openB;
        if (!matchedAny) {
            List o3A, o3B;
            bool need3A = qlen_ > e * 5, need3B = qlen_ > e * 6;
            bool ok1 = true, ok2 = true;
            if (need3A) ok1 = perform(e * 2, e * 3, false, o3A);
            if (need3B) ok2 = perform(qlen_ - e * 3, qlen_ - e * 2, false, o3B);
            if (!ok1 || !ok2) return false;
            if (need3A) {
                openA = o3A;
                removeDominated(openA);
                // openA, newA, extended = matchPairs(newA, openA)
                bool ext = matchPairs(newA, openA, extended);
                List remNew = newA, remOpen = openA;
                openA = remNew;
                newA = remOpen;
                if (ext) append(openA, extended);
                append(openA, newA);
            }
            if (need3B) {
                openB = o3B;
                removeDominated(openB);
                bool ext = matchPairs(openB, newB, extended);
                if (ext) append(openB, extended);
                append(openB, newB);
            } else {
                openB = newB;
            }
            if (need3A) {
                matchedAny = matchPairs(openA, openB, matched);
                newA = openA;
                newB = openB;
            }
        }
        outA = newA;
        outB = newB;
        return true;
    }

    bool map(List& results) {  // mapping.go:430-487
        const long long e = P.edge;
        if (qlen_ <= e * 2) {
            if (!perform(0, qlen_, true, results)) return false;
            removeDominated(results);
            return true;
        }
        List openA, openB, matched;
        bool okA = perform(0, e, false, openA);
        bool okB = perform(qlen_ - e, qlen_, false, openB);
        if (!okA || !okB) return false;
        removeDominated(openA);
        removeDominated(openB);
        bool any = matchPairs(openA, openB, matched);
        if (any) {
            results = matched;
            return true;
        }
        if (qlen_ < e * 3) {
            results = openA;
            append(results, openB);
            return true;
        }
        List nA, nB;
        if (!mapNext(openA, openB, nA, nB, matched, any)) return false;
        openA = nA;
        openB = nB;
        if (any) {
            results = matched;
            return true;
        }
        long long left = qlen_ - (qlen_ - e * 2);  // Q8: `left = query.Len() - right` with right = len - 2e
        long long right = qlen_ - e * 2;
        for (int bi : openB)
            if (pool_[bi].qOffset < right) right = pool_[bi].qOffset;
        if (!findSplitPoint(openA, openB, left, right)) return false;
        long long size = qlen_ - e;
        for (long long i = (long long)openA.size() - 1; i >= 0; i--) {
            if (pool_[openA[(size_t)i]].qInset >= size) {
                openA[(size_t)i] = openA.back();
                openA.pop_back();
            }
        }
        for (long long i = (long long)openB.size() - 1; i >= 0; i--) {
            if (pool_[openB[(size_t)i]].qOffset >= size) {
                openB[(size_t)i] = openB.back();
                openB.pop_back();
            }
        }
        results = openA;
        append(results, openB);
        return true;
    }
};

}  // namespace dph
