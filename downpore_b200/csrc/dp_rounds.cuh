// downpore_b200 — Mapper.Map's rounds after the first, on the device (mapping/mapping.go:430-487 Map, :305-383 mapNext,
// :207-288 findSplitPoint, :174-203 matchPairs, :387-428 removeDominated, :131-160 isConsistent).
//
// Map() is a sequential strategy per read: which window is queried next depends on the hits of the windows queried so
// far. One THREAD per unresolved read runs that strategy against the read's window cache in HBM. When it asks for a
// window that has not been computed yet it appends the request to the next launch's window list and stops; the window
// kernels (extract, lookup, chain) then run for all requested windows of all reads at once, a collect kernel files
// their hits in the cache, and the strategy runs again FROM THE START — Map() only ever changes its own copies of the
// hits, so a replay against a larger cache takes the same path further. A read is done when a replay reaches the end of
// Map(); its records go to a result pool that the host splices into the sub-batch's block. Between rounds the host reads
// one small counter block (requests, overflow flags): no per-read work is left on the CPU.
//
// The lists of the strategy ([]*Mapping in the reference) are index lists into a per-thread pool of hits; both live in
// a per-thread scratch area in HBM (hitCap hits, DP_RL_LISTS lists of listCap entries). A read that outgrows either,
// or a cache that fills up, raises DP_OV_ROUNDS: the host grows the capacities and recomputes (never a wrong answer).
#pragma once
#include "dp_common.cuh"
#include "dp_finish.cuh"

#define DP_OV_ROUNDS 16u  // scratch / cache of the later rounds of Map()
#define DP_RL_LISTS 16
#define DP_RL_STACK 48    // pending findSplitPoint calls of one read

enum { DP_RC_ENT = 0, DP_RC_REC, DP_RC_REQ, DP_RC_RES, DP_RC_OVF, DP_RC_OPEN, DP_RC_LEN_LO, DP_RC_LEN_HI, DP_RC_N };

struct DpRoundsDev {
    // unresolved reads ("slots") of the sub-batch
    const DpUnresolved* unres;
    int nSlots;
    const int* readLen;
    // window cache: per slot a linked list of entries {start, len, whole, next}, {n, first record}
    int* head;
    int4* ent;
    int2* ent2;
    DpMappingDev* cacheMaps;
    unsigned entCap, cacheCap;
    // the next launch's windows
    DpWindow* wins;
    int* winSlot;
    unsigned winCap;
    // results
    unsigned char* done;
    int* resN;
    unsigned* resOff;
    DpMappingDev* resMaps;
    unsigned resCap;
    // per-thread scratch
    DpHit* hits;
    int* lists;
    int hitCap, listCap;
    unsigned* cur;  // DP_RC_* counters
    // mapper parameters
    int edge, circular;
    long long refLen;
};

// files `n` hits (outMaps[off..]) of window {start, len, whole} of `slot` in the cache
__device__ __forceinline__ void dp_rounds_file(const DpRoundsDev& R, int slot, int start, int len, int whole, int n,
                                               const DpMappingDev* __restrict__ src) {
    const unsigned e = atomicAdd(R.cur + DP_RC_ENT, 1u);
    const unsigned o = atomicAdd(R.cur + DP_RC_REC, (unsigned)n);
    if (e >= R.entCap || (unsigned long long)o + (unsigned)n > R.cacheCap) {
        atomicOr(R.cur + DP_RC_OVF, DP_OV_ROUNDS);
        return;
    }
    for (int i = 0; i < n; i++) R.cacheMaps[o + i] = src[i];
    R.ent2[e] = make_int2(n, (int)o);
    __threadfence();  // (entries are only read by later kernels; the fence keeps the list consistent for a debugger's sake)
    const int next = atomicExch(R.head + slot, (int)e);
    R.ent[e] = make_int4(start, len, whole, next);
}

// round-0 windows of the unresolved reads -> cache (thread per slot)
__global__ void dp_rounds_seed_kernel(DpRoundsDev R, int minLen, const int* __restrict__ outN,
                                      const unsigned* __restrict__ outOff, const DpMappingDev* __restrict__ outMaps) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= R.nSlots) return;
    const int r = R.unres[slot].read;
    const int qlen = R.readLen[r];
    R.done[slot] = 0;
    R.resN[slot] = 0;
    R.resOff[slot] = 0;
    if (qlen < minLen) return;
    const int e = R.edge;
    if (qlen <= 2 * e) {
        dp_rounds_file(R, slot, 0, qlen, 1, outN[2 * r], outMaps + outOff[2 * r]);
    } else {
        // (filed in reverse so that the list reads first window, second window)
        dp_rounds_file(R, slot, qlen - e, e, 0, outN[2 * r + 1], outMaps + outOff[2 * r + 1]);
        dp_rounds_file(R, slot, 0, e, 0, outN[2 * r], outMaps + outOff[2 * r]);
    }
}

// the windows of the launch that just finished -> cache (thread per window)
__global__ void dp_rounds_collect_kernel(DpRoundsDev R, int nWin, const int* __restrict__ outN,
                                         const unsigned* __restrict__ outOff, const DpMappingDev* __restrict__ outMaps,
                                         unsigned long long outCap) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nWin) return;
    const DpWindow win = R.wins[w];
    const int n = outN[w];
    const unsigned o = outOff[w];
    if ((unsigned long long)o + (unsigned)n > outCap) return;  // (the launch flagged its own overflow)
    dp_rounds_file(R, R.winSlot[w], win.start, win.len, win.whole, n, outMaps + o);
}

struct DpRList {
    int* v;
    int n;
};

struct DpRoundsCtx {
    const DpRoundsDev& R;
    int slot, read;
    long long qlen;
    DpHit* pool;
    int poolN;
    DpRList L[DP_RL_LISTS];
    bool missing, ovf;

    __device__ DpRoundsCtx(const DpRoundsDev& r) : R(r) {}

    __device__ void push(DpRList& l, int x) {
        if (l.n >= R.listCap) {
            ovf = true;
            return;
        }
        l.v[l.n++] = x;
    }
    __device__ int add_hit(const DpHit& h) {
        if (poolN >= R.hitCap) {
            ovf = true;
            return 0;
        }
        pool[poolN] = h;
        return poolN++;
    }
    __device__ void copy(DpRList& a, const DpRList& b) {
        a.n = b.n;
        for (int i = 0; i < b.n; i++) a.v[i] = b.v[i];
    }
    __device__ void append(DpRList& a, const DpRList& b) {
        for (int i = 0; i < b.n; i++) push(a, b.v[i]);
    }

    // performMapping(query.SubSequence(start, end)) / performMapping(query): the cached hits, or a request
    __device__ bool perform(long long start, long long end, bool whole, DpRList& out) {
        const int len = (int)(end - start);
        out.n = 0;
        for (int ei = R.head[slot]; ei >= 0;) {
            const int4 en = R.ent[ei];
            if (en.x == (int)start && en.y == len && en.z == (int)whole) {
                const int2 e2 = R.ent2[ei];
                for (int j = 0; j < e2.x; j++) push(out, add_hit(dp_load_hit(R.cacheMaps[(unsigned)e2.y + j])));
                return true;
            }
            ei = en.w;
        }
        const unsigned q = atomicAdd(R.cur + DP_RC_REQ, 1u);
        if (q < R.winCap) {
            DpWindow wd;
            wd.read = read;
            wd.start = (int)start;
            wd.len = len;
            wd.whole = whole ? 1 : 0;
            R.wins[q] = wd;
            R.winSlot[q] = slot;
            atomicAdd(reinterpret_cast<unsigned long long*>(R.cur + DP_RC_LEN_LO), (unsigned long long)len);
        } else {
            ovf = true;
        }
        missing = true;
        return false;
    }

    __device__ bool consistent(const DpHit& a, const DpHit& b) const {
        return dp_is_consistent(a, b, qlen, R.circular != 0, R.refLen);
    }

    // removeDominated(open, open, queryLen) (every call site passes the same list twice)
    __device__ void remove_dominated(DpRList& open, DpRList& flags) {
        const int n = open.n;
        if (n == 0) return;
        for (int i = 1; i < n; i++) {  // stable insertion sort by QueryOffset
            const int x = open.v[i];
            const int key = pool[x].qOffset;
            int j = i;
            while (j > 0 && key < pool[open.v[j - 1]].qOffset) {
                open.v[j] = open.v[j - 1];
                j--;
            }
            open.v[j] = x;
        }
        int j = 0;
        for (int i = 0; i < n; i++) {
            const DpHit next = pool[open.v[i]];
            while (j < n && qlen - pool[open.v[j]].qInset < next.qOffset) j++;
            if (j == n) return;  // mapping.go:399-401: the sorted list goes back unfiltered
            bool dominated = false;
            for (int k = j; !dominated && k < n && pool[open.v[k]].qOffset < qlen - next.qInset; k++) {
                const DpHit ex = pool[open.v[k]];
                if ((long long)ex.ids * 4 > (long long)next.ids * 5) {
                    long long s = next.qOffset > ex.qOffset ? next.qOffset : ex.qOffset;
                    long long e = qlen - next.qInset;
                    if (ex.qInset > next.qInset) e = qlen - ex.qInset;
                    dominated = (e - s) * 10 > (qlen - next.qOffset - next.qInset) * 9;
                }
            }
            flags.v[i] = dominated ? 1 : 0;
        }
        int last = n - 1;
        for (int i = last; i >= 0; i--) {
            if (flags.v[i]) {
                open.v[i] = open.v[last];
                last--;
            }
        }
        open.n = last + 1;
    }

    // matchPairs (mapping.go:174-203); true <=> matched != nil
    __device__ bool match_pairs(DpRList& openA, DpRList& openB, DpRList& matched) {
        bool any = false;
        matched.n = 0;
        for (int i = openA.n - 1; i >= 0; i--) {
            for (int j = openB.n - 1; j >= 0; j--) {
                const DpHit ra = pool[openA.v[i]];
                const DpHit rb = pool[openB.v[j]];
                if (consistent(ra, rb)) {
                    const DpHit& first = ra.rc ? rb : ra;  // `if ra.RC { ra, rb = rb, ra }`
                    const DpHit& second = ra.rc ? ra : rb;
                    DpHit c;
                    c.start = first.start;
                    c.end = second.end;
                    c.qOffset = ra.qOffset;
                    c.qInset = rb.qInset;
                    c.rc = first.rc;
                    c.ids = ra.ids + rb.ids;
                    push(matched, add_hit(c));
                    any = true;
                    openA.v[i] = openA.v[openA.n - 1];
                    openA.n--;
                    openB.v[j] = openB.v[openB.n - 1];
                    openB.n--;
                    break;
                }
            }
        }
        return any;
    }

    // findSplitPoint (mapping.go:207-288). Its recursive calls are "left side with openA only, then right side with
    // openB only, then return": a stack of pending calls visits them in the same order.
    __device__ bool find_split_point(const DpRList& openA0, const DpRList& openB0, long long left0, long long right0,
                                     DpRList& mid) {
        const long long e = R.edge;
        struct Call {
            long long left, right;
            int useA, useB;
        } stack[DP_RL_STACK];
        int sp = 0;
        stack[sp++] = {left0, right0, 1, 1};
        while (sp > 0) {
            const Call c = stack[--sp];
            long long left = c.left, right = c.right;
            const int nA = c.useA ? openA0.n : 0, nB = c.useB ? openB0.n : 0;
            while (right - left >= e) {
                const long long start = (right + left - e) / 2;
                const long long end = start + e;
                if (!perform(start, end, false, mid)) return false;
                long long newLeft = left, newRight = right, afterA = 0, afterB = 0;
                for (int t = 0; t < mid.n; t++) {
                    const DpHit mm = pool[mid.v[t]];
                    for (int a = 0; a < nA; a++) {
                        DpHit& ma = pool[openA0.v[a]];
                        if (consistent(ma, mm)) {
                            ma.qInset = mm.qInset;
                            ma.ids += mm.ids;
                            if (ma.rc) ma.start = mm.start;
                            else ma.end = mm.end;
                            const long long midMatched = qlen - mm.qInset - mm.qOffset;
                            if (midMatched > afterA) afterA = midMatched;
                            if (qlen - mm.qInset > newLeft) newLeft = qlen - mm.qInset;
                            break;
                        }
                    }
                    if (afterA < (e * 2) / 3) {
                        for (int b = 0; b < nB; b++) {
                            DpHit& mb = pool[openB0.v[b]];
                            if (consistent(mm, mb)) {
                                mb.qOffset = mm.qOffset;
                                mb.ids += mm.ids;
                                if (mb.rc) mb.end = mm.end;
                                else mb.start = mm.start;
                                const long long midMatched = qlen - mm.qInset - mm.qOffset;
                                if (midMatched > afterB) afterB = midMatched;
                                if (mm.qOffset < newRight) newRight = mm.qOffset;
                                break;
                            }
                        }
                    }
                }
                if (afterA > 0 && afterB > 0) {
                    if (sp + 2 > DP_RL_STACK) {
                        ovf = true;
                        return false;
                    }
                    // (pushed in reverse: the left call runs first)
                    if (right - newRight > e * 2) stack[sp++] = {newRight + e, newRight + e * 2, 0, c.useB};
                    if (newLeft - left > e * 2) stack[sp++] = {newLeft - e * 2, newLeft - e, c.useA, 0};
                    break;
                }
                if (afterA == 0 && afterB == 0) {
                    if (sp + 2 > DP_RL_STACK) {
                        ovf = true;
                        return false;
                    }
                    if (nB > 0) stack[sp++] = {end, right, 0, 1};
                    if (nA > 0) stack[sp++] = {left, start, 1, 0};
                    break;
                }
                left = newLeft;
                right = newRight;
            }
        }
        return true;
    }

    enum { OPENA = 0, OPENB, MATCHED, OUTA, OUTB, NEWA, NEWB, EXT, T1, T2, O3A, O3B, MID, RES, FLAGS };

    // mapNext (mapping.go:305-383); false on a missing window
    __device__ bool map_next(DpRList& openA, DpRList& openB, DpRList& outA, DpRList& outB, DpRList& matched, bool& matchedAny) {
        const long long e = R.edge;
        DpRList &newA = L[NEWA], &newB = L[NEWB], &extended = L[EXT], &flags = L[FLAGS];
        if (qlen < e * 4) {
            if (!perform(e, qlen - e, false, newA)) return false;
            remove_dominated(newA, flags);
            const bool ext = match_pairs(openA, newA, extended);
            if (ext) {
                copy(openA, newA);
                append(openA, extended);
            } else {
                append(openA, newA);
            }
            matchedAny = match_pairs(openA, openB, matched);
            if (!matchedAny) {
                copy(outA, openA);
                copy(outB, openB);
            } else {
                outA.n = 0;
                outB.n = 0;
            }
            return true;
        }
        // both second-step windows are always needed: request them together
        const bool okA = perform(e, e * 2, false, newA);
        const bool okB = perform(qlen - e * 2, qlen - e, false, newB);
        if (!okA || !okB) return false;
        remove_dominated(newA, flags);
        {
            const bool ext = match_pairs(openA, newA, extended);
            append(openA, newA);
            if (ext) append(openA, extended);
        }
        remove_dominated(newB, flags);
        {
            // openB, newB, extended = matchPairs(newB, openB)
            const bool ext = match_pairs(newB, openB, extended);
            copy(L[T1], newB);
            copy(L[T2], openB);
            copy(openB, L[T1]);
            copy(newB, L[T2]);
            append(openB, newB);
            if (ext) append(openB, extended);
        }
        matchedAny = match_pairs(openA, openB, matched);
        copy(newA, openA);
        copy(newB, openB);
        if (!matchedAny) {
            DpRList &o3A = L[O3A], &o3B = L[O3B];
            const bool need3A = qlen > e * 5, need3B = qlen > e * 6;
            bool ok1 = true, ok2 = true;
            if (need3A) ok1 = perform(e * 2, e * 3, false, o3A);
            if (need3B) ok2 = perform(qlen - e * 3, qlen - e * 2, false, o3B);
            if (!ok1 || !ok2) return false;
            if (need3A) {
                copy(openA, o3A);
                remove_dominated(openA, flags);
                // openA, newA, extended = matchPairs(newA, openA)
                const bool ext = match_pairs(newA, openA, extended);
                copy(L[T1], newA);
                copy(L[T2], openA);
                copy(openA, L[T1]);
                copy(newA, L[T2]);
                if (ext) append(openA, extended);
                append(openA, newA);
            }
            if (need3B) {
                copy(openB, o3B);
                remove_dominated(openB, flags);
                const bool ext = match_pairs(openB, newB, extended);
                if (ext) append(openB, extended);
                append(openB, newB);
            } else {
                copy(openB, newB);
            }
            if (need3A) {
                matchedAny = match_pairs(openA, openB, matched);
                copy(newA, openA);
                copy(newB, openB);
            }
        }
        copy(outA, newA);
        copy(outB, newB);
        return true;
    }

    // Map (mapping.go:430-487); false on a missing window. Results in L[RES].
    __device__ bool map() {
        const long long e = R.edge;
        DpRList &results = L[RES], &openA = L[OPENA], &openB = L[OPENB], &matched = L[MATCHED], &flags = L[FLAGS];
        if (qlen <= e * 2) {
            if (!perform(0, qlen, true, results)) return false;
            remove_dominated(results, flags);
            return true;
        }
        const bool okA = perform(0, e, false, openA);
        const bool okB = perform(qlen - e, qlen, false, openB);
        if (!okA || !okB) return false;
        remove_dominated(openA, flags);
        remove_dominated(openB, flags);
        bool any = match_pairs(openA, openB, matched);
        if (any) {
            copy(results, matched);
            return true;
        }
        if (qlen < e * 3) {
            copy(results, openA);
            append(results, openB);
            return true;
        }
        if (!map_next(openA, openB, L[OUTA], L[OUTB], matched, any)) return false;
        copy(openA, L[OUTA]);
        copy(openB, L[OUTB]);
        if (any) {
            copy(results, matched);
            return true;
        }
        const long long left = qlen - (qlen - e * 2);  // Q8: `left = query.Len() - right` with right = len - 2e
        long long right = qlen - e * 2;
        for (int i = 0; i < openB.n; i++)
            if (pool[openB.v[i]].qOffset < right) right = pool[openB.v[i]].qOffset;
        if (!find_split_point(openA, openB, left, right, L[MID])) return false;
        const long long size = qlen - e;
        for (int i = openA.n - 1; i >= 0; i--) {
            if (pool[openA.v[i]].qInset >= size) {
                openA.v[i] = openA.v[openA.n - 1];
                openA.n--;
            }
        }
        for (int i = openB.n - 1; i >= 0; i--) {
            if (pool[openB.v[i]].qOffset >= size) {
                openB.v[i] = openB.v[openB.n - 1];
                openB.n--;
            }
        }
        copy(results, openA);
        append(results, openB);
        return true;
    }
};

// One replay of Map() per open slot (threads stride over the slots; scratch is per thread).
__global__ void __launch_bounds__(64) dp_rounds_replay_kernel(DpRoundsDev R, int minLen) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nT = gridDim.x * blockDim.x;
    for (int slot = t; slot < R.nSlots; slot += nT) {
        if (R.done[slot]) continue;
        DpRoundsCtx C(R);
        C.slot = slot;
        C.read = R.unres[slot].read;
        C.qlen = R.readLen[C.read];
        C.pool = R.hits + (size_t)t * R.hitCap;
        C.poolN = 0;
        for (int l = 0; l < DP_RL_LISTS; l++) {
            C.L[l].v = R.lists + ((size_t)t * DP_RL_LISTS + l) * R.listCap;
            C.L[l].n = 0;
        }
        C.missing = false;
        C.ovf = false;
        bool ok = true;
        if (C.qlen >= minLen) ok = C.map();  // (shorter reads have no mapping: the reference's scans over-read them)
        if (C.ovf) {
            atomicOr(R.cur + DP_RC_OVF, DP_OV_ROUNDS);
            continue;
        }
        if (!ok || C.missing) {
            atomicAdd(R.cur + DP_RC_OPEN, 1u);
            continue;
        }
        const DpRList& res = C.L[DpRoundsCtx::RES];
        const unsigned o = atomicAdd(R.cur + DP_RC_RES, (unsigned)res.n);
        if ((unsigned long long)o + (unsigned)res.n > R.resCap) {
            atomicOr(R.cur + DP_RC_OVF, DP_OV_ROUNDS);
            continue;
        }
        for (int i = 0; i < res.n; i++) R.resMaps[o + i] = dp_store_hit(C.pool[res.v[i]]);
        R.resN[slot] = res.n;
        R.resOff[slot] = o;
        R.done[slot] = 1;
    }
}
