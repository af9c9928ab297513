// downpore_b200 — the input and output side of `downpore map` on the device:
//   * record splitting: the first-pass rules of the reference's reader, readFasta (sequence/seqio.go:188-267), applied to
//     a FASTA/FASTQ file image in device-visible memory -> a table of (name span, sequence span) records; the sequences
//     are then mapped where they lie (dp_mapper_map_batch_spans), nothing is copied per read on the host;
//   * PAF formatting: Mapper.AsString (mapping/mapping.go:112-122) for every mapping of a batch -> one text block.
#pragma once
#include "dp_common.cuh"

struct DpRecordDev {  // = dp_record of the C ABI
    long long nameStart, nameLen, seqStart, seqLen;
};

// ---------------------------------------------------------------------------------------------------------------
// Lines. A line starts at byte 0 and after every '\n' that is not the last byte. Each thread owns DP_IO_TILE bytes of
// the 16-byte-aligned span around the buffer: pass 1 counts the line starts of its tile, a device scan turns the counts
// into offsets, pass 2 writes the positions.
// ---------------------------------------------------------------------------------------------------------------
#define DP_IO_TILE 64

template <bool WRITE>
__global__ void __launch_bounds__(256) dp_line_starts_kernel(const unsigned char* __restrict__ buf, long long n,
                                                             long long nTiles, int* __restrict__ cnt,
                                                             const int* __restrict__ off, long long* __restrict__ lineStart) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nTiles) return;
    if (t == nTiles) {  // sentinel entry of the scan
        if (!WRITE) cnt[nTiles] = 0;
        return;
    }
    const unsigned mis = (unsigned)((unsigned long long)buf & 15ull);
    const uint4* blk = reinterpret_cast<const uint4*>(buf - mis) + t * (DP_IO_TILE / 16);
    const long long first = t * DP_IO_TILE - (long long)mis;  // buffer position of the tile's first byte
    int c = 0;
    int o = WRITE ? off[t] : 0;
#pragma unroll
    for (int b = 0; b < DP_IO_TILE / 16; b++) {
        const long long p0 = first + 16 * b;
        if (p0 >= n || p0 + 16 <= 0) continue;
        const uint4 v = __ldg(blk + b);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const long long p = p0 + j;  // a '\n' at p starts a line at p + 1
            const bool nl = ((w[j >> 2] >> (8 * (j & 3))) & 0xffu) == (unsigned)'\n';
            if (nl && p >= 0 && p + 1 < n) {
                if (WRITE) lineStart[1 + o + c] = p + 1;
                c++;
            }
        }
    }
    if (!WRITE) cnt[t] = c;
    if (WRITE && t == 0) lineStart[0] = 0;
}

// first byte of every line -> 0 name line, 1 sequence line ('A'..'T', seqio.go:209), 2 '@' (fastq comment), 3 '+'
__global__ void dp_line_class_kernel(const unsigned char* __restrict__ buf, const long long* __restrict__ lineStart,
                                     long long nLines, unsigned char* __restrict__ cls) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nLines) return;
    const unsigned char c = buf[lineStart[i]];
    cls[i] = (c >= 'A' && c <= 'T') ? 1 : (c == '@' ? 2 : (c == '+' ? 3 : 0));
}

// The reader's line automaton (seqio.go:189-263) over the line classes; the only sequential step, run by one warp: the
// lanes stage 4096 classes at a time in shared memory, lane 0 walks them. Every sequence line it visits becomes a
// candidate record (sequence line, line of the name in force); once a '@' line has been seen (isFastq) a sequence line
// is followed by its '+' line — anything else there is the reference's log.Fatal — and a quality line, both skipped
// whatever they start with. A piece of a file that is not its last (`final` = 0) stops where the '+' line is not in the
// piece yet; the host cuts the piece at the last name line (res[2]) and hands the rest over with the next piece.
// res[0] = candidates, res[1] = 1 on an invalid fastq record, res[2] = the last name line, res[3] = isFastq before it.
__global__ void __launch_bounds__(32) dp_record_walk_kernel(const unsigned char* __restrict__ cls, long long nLines,
                                                            bool endsWithNewline, bool final, bool fastqIn,
                                                            long long* __restrict__ candSeq,
                                                            long long* __restrict__ candName, long long* __restrict__ res) {
    __shared__ unsigned char sh[4096 + 16];
    const unsigned lane = threadIdx.x;
    long long nCand = 0, i = 1, lastName = 0;
    bool isFastq = fastqIn || (nLines > 0 && cls[0] == 2), bad = false, stop = false;
    bool fqAtName = fastqIn;
    // (a file whose first line has no newline yields nothing: seqio.go:192-197)
    while (i < nLines && !bad && !stop) {
        const long long base = i;
        const long long m = min((long long)4096, nLines - base);
        for (long long j = lane; j < m; j += 32) sh[j] = cls[base + j];
        __syncwarp();
        if (lane == 0) {
            // walk while the line and (for fastq) its '+' line are staged
            while (i < nLines && i - base < m - 1 + (base + m == nLines ? 1 : 0)) {
                const unsigned char c = sh[i - base];
                if (c == 1) {
                    if (isFastq) {
                        // the '+' line must exist, end in '\n' (ReadBytes err == nil) and start with '+'
                        const bool have = i + 1 < nLines && (i + 1 < nLines - 1 || endsWithNewline);
                        if (!have && !final) {
                            stop = true;
                            break;
                        }
                        candSeq[nCand] = i;
                        candName[nCand] = lastName;
                        nCand++;
                        if (!have || sh[i + 1 - base] != 3) {
                            bad = true;
                            break;
                        }
                        i += 3;
                    } else {
                        candSeq[nCand] = i;
                        candName[nCand] = lastName;
                        nCand++;
                        i += 1;
                    }
                } else {
                    fqAtName = isFastq;
                    if (c == 2) isFastq = true;
                    lastName = i;
                    i += 1;
                }
            }
        }
        i = __shfl_sync(DP_FULL, i, 0);
        bad = __shfl_sync(DP_FULL, (int)bad, 0) != 0;
        stop = __shfl_sync(DP_FULL, (int)stop, 0) != 0;
        __syncwarp();
    }
    if (lane == 0) {
        res[0] = nCand;
        res[1] = bad ? 1 : 0;
        res[2] = lastName;
        res[3] = fqAtName ? 1 : 0;
    }
}

__device__ __forceinline__ bool dp_is_space(unsigned char c) {  // strings.TrimSpace, ASCII subset
    return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f';
}

// A candidate is kept iff its line, newline included, has at least min_length bytes (seqio.go:210); the sequence is the
// line minus its LAST BYTE (also when the file ends without a newline, :217), the name the line after its first byte,
// TrimSpace'd (:216). WRITE = false: keep flags for the scan; WRITE = true: the records at their scanned positions.
template <bool WRITE>
__global__ void dp_record_finish_kernel(const unsigned char* __restrict__ buf, long long n,
                                        const long long* __restrict__ lineStart, long long nLines,
                                        const long long* __restrict__ candSeq, const long long* __restrict__ candName,
                                        long long nCand, long long minLength, long long dropName,
                                        int* __restrict__ keep, const int* __restrict__ off,
                                        DpRecordDev* __restrict__ out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (!WRITE && r == nCand) keep[nCand] = 0;
    if (r >= nCand) return;
    const long long s = candSeq[r];
    const long long s0 = lineStart[s], s1 = s + 1 < nLines ? lineStart[s + 1] : n;
    // (dropName >= 0: a piece that is not the file's last keeps nothing after its last name line)
    const bool kept = s1 - s0 >= minLength && candName[r] != dropName;
    if (!WRITE) {
        keep[r] = kept ? 1 : 0;
        return;
    }
    if (!kept) return;
    const long long m = candName[r];
    long long a = lineStart[m] + 1, b = m + 1 < nLines ? lineStart[m + 1] : n;
    if (a > b) a = b;
    while (a < b && dp_is_space(buf[a])) a++;
    while (b > a && dp_is_space(buf[b - 1])) b--;
    DpRecordDev rec;
    rec.nameStart = a;
    rec.nameLen = b - a;
    rec.seqStart = s0;
    rec.seqLen = s1 - s0 - 1;
    out[off[r]] = rec;
}

// ---------------------------------------------------------------------------------------------------------------
// PAF lines (mapping/mapping.go:112-122):
//   name \t qlen \t qOffset \t qlen-qInset \t +|- \t refName \t refLen \t start \t end \t ids \t mappedLength \t 255 \n
// One thread per mapping; pass 1 the line length, device scan, pass 2 the characters.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dp_dec_len(long long v) {
    int n = v < 0 ? 1 : 0;
    unsigned long long u = v < 0 ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
    do {
        n++;
        u /= 10;
    } while (u);
    return n;
}
__device__ __forceinline__ char* dp_dec_put(char* p, long long v) {
    const int n = dp_dec_len(v);
    unsigned long long u = v < 0 ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
    if (v < 0) p[0] = '-';
    for (int i = n - 1; i >= (v < 0 ? 1 : 0); i--) {
        p[i] = (char)('0' + (int)(u % 10));
        u /= 10;
    }
    return p + n;
}

struct DpPafParams {
    long long refLen;
    int circular;
    int refNameLen;
    char refName[256];
};

template <bool WRITE>
__global__ void dp_paf_kernel(DpPafParams P, const DpMappingDev* __restrict__ maps, long long nMaps,
                              const long long* __restrict__ outOff, long long nReads,
                              const unsigned char* __restrict__ names, const DpRecordDev* __restrict__ recs,
                              int* __restrict__ lineLen, const long long* __restrict__ lineOff,
                              char* __restrict__ text) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (!WRITE && j == nMaps) lineLen[nMaps] = 0;
    if (j >= nMaps) return;
    // the read of mapping j: the last i with outOff[i] <= j
    long long lo = 0, hi = nReads;
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (outOff[mid] <= j) lo = mid;
        else hi = mid;
    }
    const long long i = lo;
    const DpMappingDev m = maps[j];
    const DpRecordDev rec = recs[i];
    const long long qlen = rec.seqLen;
    long long mapped = m.end - m.start;
    if (P.circular && mapped < 0) mapped = P.refLen - m.start + m.end;
    const long long f[9] = {qlen, (long long)m.qOffset, qlen - m.qInset, 0, P.refLen, m.start, m.end, (long long)m.ids, mapped};
    if (!WRITE) {
        long long len = rec.nameLen + 1 + 2 + P.refNameLen + 1 + 4;  // name\t, +\t, refName\t, 255\n
        for (int x = 0; x < 9; x++)
            if (x != 3) len += dp_dec_len(f[x]) + 1;
        lineLen[j] = (int)len;
        return;
    }
    char* p = text + lineOff[j];
    const unsigned char* nm = names + rec.nameStart;
    for (long long x = 0; x < rec.nameLen; x++) *p++ = (char)nm[x];
    *p++ = '\t';
    for (int x = 0; x < 9; x++) {
        if (x == 3) {
            *p++ = (m.rc & 0xff) ? '-' : '+';
            *p++ = '\t';
            for (int y = 0; y < P.refNameLen; y++) *p++ = P.refName[y];
            *p++ = '\t';
        } else {
            p = dp_dec_put(p, f[x]);
            *p++ = '\t';
        }
    }
    *p++ = '2';
    *p++ = '5';
    *p++ = '5';
    *p++ = '\n';
}
