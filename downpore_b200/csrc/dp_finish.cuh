// downpore_b200 — device side of Mapper.Map's first step (mapping/mapping.go:430-445, mapEnds :164-172,
// removeDominated :387-428, matchPairs :174-203, isConsistent :131-160) plus the small table kernels around it.
//
// Round 0 of every read is fixed: the whole read if len <= 2*edge, else its first and last `edge` bases. After the
// window kernels have produced the hits of those windows, one thread per read finishes the read here if Map() would
// return at this point (short read; end-to-end pair found; or len < 3*edge). Everything else is flagged unresolved
// and continues with the later rounds of Map() (dp_rounds.cuh) — a fraction of a percent of reads on ONT-like data.
#pragma once
#include "dp_common.cuh"

#define DP_FIN_CAP 8  // hits per window handled here; more -> the general strategy kernel (dp_rounds.cuh)

// read table: lengths and packed-word demand from the (sub-batch relative) byte offsets
__global__ void dp_read_table_kernel(const long long* __restrict__ seqOff, long long n, int* __restrict__ readLen,
                                     long long* __restrict__ wordsNeeded) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        long long len = seqOff[i + 1] - seqOff[i];
        readLen[i] = (int)len;
        wordsNeeded[i] = (len + 15) / 16 + 1;
    }
    if (i == n) wordsNeeded[n] = 0;
}

// round-0 windows: two slots per read (the second stays empty for short reads)
__global__ void dp_round0_windows_kernel(const int* __restrict__ readLen, long long n, int edge, int minLen,
                                         DpWindow* __restrict__ wins) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int len = readLen[i];
    DpWindow a, b;
    a.read = b.read = (int)i;
    a.start = b.start = 0;
    a.len = b.len = 0;
    a.whole = b.whole = 0;
    if (len >= minLen) {
        if (len <= 2 * edge) {
            a.len = len;
            a.whole = 1;
        } else {
            a.len = edge;
            b.start = len - edge;
            b.len = edge;
        }
    }
    wins[2 * i] = a;
    wins[2 * i + 1] = b;
}

struct DpHit {
    long long start, end;
    int qOffset, qInset, ids, rc;
};

__device__ __forceinline__ bool dp_is_consistent(const DpHit& left, const DpHit& right, long long qlen, bool circular,
                                                 long long refLen) {
    if (left.rc != right.rc) return false;
    long long expectedDistance = (long long)right.qOffset - qlen + left.qInset;
    long long distance = !left.rc ? right.start - left.end : left.start - right.end;
    if (circular && distance < -50) distance += refLen;
    if (distance < 50 && expectedDistance < 50 && distance > -50) return true;
    if (distance < 500) return expectedDistance < (distance * 3) / 2 && expectedDistance > (distance * 2) / 3;
    if (distance > 5000) return expectedDistance < (distance * 10) / 9 && expectedDistance > (distance * 9) / 10;
    // Go: ratio = 3.0/2.0 + ratio*(10.0/9.0-3.0/2.0), constants folded exactly, no FMA on amd64: round each step
    double ratio = __ddiv_rn((double)(distance - 500), 4500.0);
    ratio = __dadd_rn(1.5, __dmul_rn(ratio, -7.0 / 18.0));
    double a = __dmul_rn((double)expectedDistance, ratio);
    double b = __ddiv_rn((double)expectedDistance, ratio);
    return distance < (long long)a && distance > (long long)b;
}

// removeDominated(open, open, qlen) on a list of at most DP_FIN_CAP hits; returns the new length
__device__ int dp_remove_dominated(DpHit* open, int n, long long qlen) {
    if (n == 0) return 0;
    for (int i = 1; i < n; i++) {  // stable insertion sort by QueryOffset (= Go's sort.Sort for n <= 12)
        DpHit x = open[i];
        int j = i;
        while (j > 0 && x.qOffset < open[j - 1].qOffset) {
            open[j] = open[j - 1];
            j--;
        }
        open[j] = x;
    }
    bool toRemove[DP_FIN_CAP];
    int j = 0;
    for (int i = 0; i < n; i++) {
        const DpHit& next = open[i];
        while (j < n && qlen - open[j].qInset < next.qOffset) j++;
        if (j == n) return n;  // mapping.go:399-401: returns the (sorted) list unfiltered
        bool dominated = false;
        for (int k = j; !dominated && k < n && open[k].qOffset < qlen - next.qInset; k++) {
            if ((long long)open[k].ids * 4 > (long long)next.ids * 5) {
                long long s = next.qOffset;
                if (open[k].qOffset > s) s = open[k].qOffset;
                long long e = qlen - next.qInset;
                if (open[k].qInset > next.qInset) e = qlen - open[k].qInset;
                dominated = (e - s) * 10 > (qlen - next.qOffset - next.qInset) * 9;
            }
        }
        toRemove[i] = dominated;
    }
    int last = n - 1;
    for (int i = last; i >= 0; i--) {
        if (toRemove[i]) {
            open[i] = open[last];
            last--;
        }
    }
    return last + 1;
}

__device__ __forceinline__ DpHit dp_load_hit(const DpMappingDev& m) {
    DpHit h;
    h.start = m.start;
    h.end = m.end;
    h.qOffset = m.qOffset;
    h.qInset = m.qInset;
    h.ids = m.ids;
    h.rc = m.rc & 0xff;
    return h;
}
__device__ __forceinline__ DpMappingDev dp_store_hit(const DpHit& h) {
    DpMappingDev m;
    m.start = h.start;
    m.end = h.end;
    m.qOffset = h.qOffset;
    m.qInset = h.qInset;
    m.ids = h.ids;
    m.rc = h.rc;
    return m;
}

// Map()'s first decision for one read, in two passes around a device-wide exclusive scan so that the records land in
// READ ORDER (the host then takes a sub-batch's results as one block instead of visiting every read):
//   WRITE = false: cnt[r] = records the read delivers — its final mappings if Map() returns at this point, else none
//                  (the later rounds deliver them, dp_rounds.cuh);
//   WRITE = true : the same decision again (cheaper than parking the records), records written at off[r] (the scan of
//                  cnt); reads that are not finished are appended to `unres` as {read, nA, nB}.
// finMaps may live in page-locked host memory mapped into the device address space: consecutive reads write consecutive
// records, so the posted writes coalesce and the sub-batch needs no copy call for its records. When the records do not
// fit finCapacity nothing is written (the host grows the buffer and repeats this pass).
struct DpUnresolved {
    int read, nA, nB;
};

template <bool WRITE>
__global__ void __launch_bounds__(128) dp_finish_round0_kernel(DpIndexDev I, const int* __restrict__ readLen,
                                                               long long n, int minLen,
                                                               const int* __restrict__ outN,
                                                               const unsigned* __restrict__ outOff,
                                                               const DpMappingDev* __restrict__ outMaps,
                                                               int* __restrict__ cnt, const unsigned* __restrict__ off,
                                                               DpMappingDev* __restrict__ finMaps,
                                                               unsigned long long finCapacity,
                                                               DpUnresolved* __restrict__ unres, int* __restrict__ nUnres,
                                                               int unresCap, const DpCounters* __restrict__ ctr,
                                                               DpCounters* __restrict__ hostCtr) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // the launch's work counters and overflow flags travel with the results (every kernel that updates them has
    // finished: same stream)
    if (WRITE && r == 0 && hostCtr) *hostCtr = *ctr;
    if (!WRITE && r == n) cnt[n] = 0;
    if (r >= n) return;
    const long long qlen = readLen[r];
    const int e = I.edge;
    DpHit A[DP_FIN_CAP], B[DP_FIN_CAP], M[DP_FIN_CAP];
    int nOut = 0;
    const DpHit* outList = A;
    const DpHit* outList2 = B;
    int nOut2 = 0;
    bool done = true;
    int nA = 0, nB = 0;
    if (qlen >= minLen) {
        nA = outN[2 * r];
        nB = outN[2 * r + 1];
        if (nA > DP_FIN_CAP || nB > DP_FIN_CAP) {
            done = false;
        } else {
            for (int i = 0; i < nA; i++) A[i] = dp_load_hit(outMaps[outOff[2 * r] + i]);
            for (int i = 0; i < nB; i++) B[i] = dp_load_hit(outMaps[outOff[2 * r + 1] + i]);
            if (qlen <= 2 * e) {
                nOut = dp_remove_dominated(A, nA, qlen);  // mapping.go:433-436
            } else {
                int mA = dp_remove_dominated(A, nA, qlen);
                int mB = dp_remove_dominated(B, nB, qlen);
                // matchPairs (mapping.go:174-203)
                int nM = 0;
                for (int i = mA - 1; i >= 0; i--) {
                    for (int j = mB - 1; j >= 0; j--) {
                        if (dp_is_consistent(A[i], B[j], qlen, I.circular != 0, I.refLen)) {
                            const DpHit& first = A[i].rc ? B[j] : A[i];
                            const DpHit& second = A[i].rc ? A[i] : B[j];
                            DpHit c;
                            c.start = first.start;
                            c.end = second.end;
                            c.qOffset = A[i].qOffset;
                            c.qInset = B[j].qInset;
                            c.rc = first.rc;
                            c.ids = A[i].ids + B[j].ids;
                            M[nM++] = c;
                            A[i] = A[mA - 1];
                            mA--;
                            B[j] = B[mB - 1];
                            mB--;
                            break;
                        }
                    }
                }
                if (nM > 0) {
                    outList = M;
                    nOut = nM;
                } else if (qlen < 3ll * e) {  // mapping.go:444-445: append(openA, openB...)
                    nOut = mA;
                    nOut2 = mB;
                } else {
                    done = false;
                }
            }
        }
    }
    if (!WRITE) {
        cnt[r] = done ? nOut + nOut2 : 0;
        return;
    }
    if (!done) {  // the later rounds of Map() take over (the window results stay on the device)
        const int slot = atomicAdd(nUnres, 1);
        if (slot < unresCap) {
            DpUnresolved u;
            u.read = (int)r;
            u.nA = nA;
            u.nB = nB;
            unres[slot] = u;
        }
    }
    if ((unsigned long long)off[n] > finCapacity) return;
    const unsigned base = off[r];
    if (done) {
        for (int i = 0; i < nOut; i++) finMaps[base + i] = dp_store_hit(outList[i]);
        for (int i = 0; i < nOut2; i++) finMaps[base + nOut + i] = dp_store_hit(outList2[i]);
    }
}
