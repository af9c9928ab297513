"""ctypes access to the CPU oracle (oracle/liboracle.so).

ORACLE = TEST INFRASTRUCTURE ONLY. Import this from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
only; nothing under downpore_b200/ may import it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

c_ll = ctypes.c_longlong
c_vp = ctypes.c_void_p


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < newest:
        subprocess.check_call(["make", "-C", _HERE, "-j4", "all"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = ctypes.CDLL(_LIB)
    sig = {
        "dpo_last_error": (ctypes.c_char_p, []),
        "dpo_packed_new": (c_vp, [c_vp, c_ll]),
        "dpo_packed_free": (None, [c_vp]),
        "dpo_packed_sub": (c_vp, [c_vp, c_ll, c_ll]),
        "dpo_packed_rc": (c_vp, [c_vp]),
        "dpo_packed_len": (c_ll, [c_vp]),
        "dpo_packed_nbytes": (c_ll, [c_vp]),
        "dpo_packed_bytes": (None, [c_vp, c_vp]),
        "dpo_packed_fields": (None, [c_vp, c_vp]),
        "dpo_packed_string": (None, [c_vp, c_vp]),
        "dpo_packed_kmer_at": (c_ll, [c_vp, c_ll, c_ll]),
        "dpo_packed_next_kmer": (c_ll, [c_vp, c_ll, c_ll, c_ll]),
        "dpo_packed_count_kmers": (c_ll, [c_vp, c_ll, c_ll, c_vp]),
        "dpo_packed_count_kmers_between": (c_ll, [c_vp, c_ll, c_ll, c_ll, c_ll, c_vp]),
        "dpo_packed_write_segments": (c_ll, [c_vp, c_ll, c_vp, c_vp]),
        "dpo_packed_short_kmers": (c_ll, [c_vp, c_ll, ctypes.c_int, c_vp]),
        "dpo_byte_new": (c_vp, [c_vp, c_ll]),
        "dpo_byte_free": (None, [c_vp]),
        "dpo_byte_sub": (c_vp, [c_vp, c_ll, c_ll]),
        "dpo_byte_rc": (c_vp, [c_vp]),
        "dpo_byte_len": (c_ll, [c_vp]),
        "dpo_byte_fields": (None, [c_vp, c_vp]),
        "dpo_byte_string": (None, [c_vp, c_vp]),
        "dpo_byte_kmer_at": (c_ll, [c_vp, c_ll, c_ll]),
        "dpo_byte_next_kmer": (c_ll, [c_vp, c_ll, c_ll, c_ll]),
        "dpo_byte_count_kmers": (c_ll, [c_vp, c_ll, c_ll, c_ll, c_vp]),
        "dpo_byte_count_kmers_between": (c_ll, [c_vp, c_ll, c_ll, c_ll, c_ll, c_ll, c_vp]),
        "dpo_byte_write_segments": (c_ll, [c_vp, c_ll, c_ll, c_vp, c_vp]),
        "dpo_byte_short_kmers": (c_ll, [c_vp, c_ll, ctypes.c_int, c_vp]),
        "dpo_kmer_value": (c_ll, [c_vp, c_ll]),
        "dpo_pack_bytes": (None, [c_vp, c_ll, c_vp]),
        "dpo_intset_new": (c_vp, []),
        "dpo_intset_free": (None, [c_vp]),
        "dpo_intset_add": (None, [c_vp, ctypes.c_ulonglong]),
        "dpo_intset_contains": (ctypes.c_int, [c_vp, ctypes.c_ulonglong]),
        "dpo_intset_size": (ctypes.c_ulonglong, [c_vp]),
        "dpo_intset_count_intersection": (ctypes.c_ulonglong, [c_vp, c_vp]),
        "dpo_intset_count_intersection_to": (c_ll, [c_vp, c_vp, c_ll]),
        "dpo_get_shared_ids": (c_ll, [c_vp, c_ll, c_ll, ctypes.c_int, c_vp, c_ll]),
        "dpo_kmer_values": (ctypes.c_int, [c_vp, c_ll, ctypes.c_int, c_vp]),
        "dpo_kmer_counts": (ctypes.c_int, [c_vp, c_ll, ctypes.c_int, c_vp]),
        "dpo_mapper_new": (c_vp, [c_vp, c_ll, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
        "dpo_mapper_new_ex": (c_vp, [c_vp, c_ll, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                              ctypes.c_int, ctypes.c_int]),
        "dpo_mapper_is_lean": (ctypes.c_int, [c_vp]),
        "dpo_mapper_free": (None, [c_vp]),
        "dpo_mapper_num_seeds": (c_ll, [c_vp]),
        "dpo_mapper_num_chunks": (c_ll, [c_vp]),
        "dpo_mapper_seed_kmers": (None, [c_vp, c_vp]),
        "dpo_mapper_chunk": (c_ll, [c_vp, c_ll, c_vp, c_vp, c_ll]),
        "dpo_window_segments": (c_ll, [c_vp, c_vp, c_ll, c_ll, c_ll, ctypes.c_int, ctypes.c_int, c_vp, c_ll, c_vp]),
        "dpo_window_candidates": (c_ll, [c_vp, c_vp, c_ll, c_ll, c_ll, ctypes.c_int, ctypes.c_int, c_vp, c_ll]),
        "dpo_window_mappings": (c_ll, [c_vp, c_vp, c_ll, c_ll, c_ll, ctypes.c_int, c_vp, c_ll]),
        "dpo_map_batch": (ctypes.c_int, [c_vp, c_ll, c_vp, c_vp, ctypes.c_int, c_vp, c_vp, c_vp]),
        "dpo_free": (None, [c_vp]),
        "dpo_kmer_counts_batch": (ctypes.c_int, [c_vp, c_vp, c_ll, ctypes.c_int, c_vp]),
        "dpo_kmer_values_from_counts": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp]),
        "dpo_overlap_round": (c_vp, [c_vp, c_vp, c_ll, c_vp, c_ll, c_vp, c_vp, ctypes.c_double]),
        "dpo_overlap_free": (None, [c_vp]),
        "dpo_overlap_get": (c_ll, [c_vp, ctypes.c_int, c_vp, c_ll]),
        "dpo_parse_fasta": (c_vp, [c_vp, c_ll, c_ll, ctypes.POINTER(c_ll), ctypes.POINTER(c_ll)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _err():
    return lib().dpo_last_error().decode()


def _u8(x):
    if isinstance(x, (bytes, bytearray, str)):
        x = x.encode() if isinstance(x, str) else x
        return np.frombuffer(bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


class Packed:
    """packedSequence (sequence/sequence.go:43-53)."""

    def __init__(self, ascii_or_handle, _own=True):
        L = lib()
        if isinstance(ascii_or_handle, int):
            self.h = ascii_or_handle
        else:
            a = _u8(ascii_or_handle)
            self.h = L.dpo_packed_new(a.ctypes.data, a.size)
        if not self.h:
            raise RuntimeError(_err())

    def __del__(self):
        if getattr(self, "h", None):
            lib().dpo_packed_free(self.h)
            self.h = None

    def sub(self, start, end):
        h = lib().dpo_packed_sub(self.h, start, end)
        if not h:
            raise RuntimeError(_err())
        return Packed(h)

    def rc(self):
        return Packed(lib().dpo_packed_rc(self.h))

    def __len__(self):
        return lib().dpo_packed_len(self.h)

    def bytes(self):
        out = np.empty(lib().dpo_packed_nbytes(self.h), dtype=np.uint8)
        lib().dpo_packed_bytes(self.h, out.ctypes.data)
        return out

    def fields(self):
        f = np.zeros(5, dtype=np.int64)
        lib().dpo_packed_fields(self.h, f.ctypes.data)
        return dict(offset=int(f[0]), inset=int(f[1]), firstLen=int(f[2]), finalLen=int(f[3]), length=int(f[4]))

    def string(self):
        out = np.empty(len(self), dtype=np.uint8)
        lib().dpo_packed_string(self.h, out.ctypes.data)
        return out.tobytes().decode()

    def kmer_at(self, i, k):
        return lib().dpo_packed_kmer_at(self.h, i, k)

    def next_kmer(self, cur, mask, idx):
        return lib().dpo_packed_next_kmer(self.h, cur, mask, idx)

    def count_kmers(self, up_to, k, seeds):
        return lib().dpo_packed_count_kmers(self.h, up_to, k, seeds.ctypes.data)

    def count_kmers_between(self, frm, to, up_to, k, seeds):
        r = lib().dpo_packed_count_kmers_between(self.h, frm, to, up_to, k, seeds.ctypes.data)
        if r < 0:
            raise RuntimeError(_err())
        return r

    def write_segments(self, k, seeds):
        seg = np.empty(2 * (len(self) + 16) + 1, dtype=np.int64)
        n = lib().dpo_packed_write_segments(self.h, k, seeds.ctypes.data, seg.ctypes.data)
        return seg[:n].copy()

    def short_kmers(self, k, collapse):
        out = np.empty(len(self) + 1, dtype=np.uint16)
        n = lib().dpo_packed_short_kmers(self.h, k, int(collapse), out.ctypes.data)
        return out[:n].copy()


class Byte:
    """byteSequence (sequence/sequence.go:33-40): independent check of the packed asm emulation."""

    def __init__(self, ascii_or_handle):
        L = lib()
        if isinstance(ascii_or_handle, int):
            self.h = ascii_or_handle
        else:
            a = _u8(ascii_or_handle)
            self.h = L.dpo_byte_new(a.ctypes.data, a.size)

    def __del__(self):
        if getattr(self, "h", None):
            lib().dpo_byte_free(self.h)
            self.h = None

    def sub(self, start, end):
        return Byte(lib().dpo_byte_sub(self.h, start, end))

    def rc(self):
        return Byte(lib().dpo_byte_rc(self.h))

    def __len__(self):
        return lib().dpo_byte_len(self.h)

    def fields(self):
        f = np.zeros(2, dtype=np.int64)
        lib().dpo_byte_fields(self.h, f.ctypes.data)
        return dict(offset=int(f[0]), inset=int(f[1]))

    def string(self):
        out = np.empty(len(self), dtype=np.uint8)
        lib().dpo_byte_string(self.h, out.ctypes.data)
        return out.tobytes().decode()

    def kmer_at(self, i, k):
        return lib().dpo_byte_kmer_at(self.h, i, k)

    def next_kmer(self, cur, mask, idx):
        return lib().dpo_byte_next_kmer(self.h, cur, mask, idx)

    def count_kmers(self, up_to, k, mask, seeds):
        return lib().dpo_byte_count_kmers(self.h, up_to, k, mask, seeds.ctypes.data)

    def count_kmers_between(self, frm, to, up_to, k, mask, seeds):
        return lib().dpo_byte_count_kmers_between(self.h, frm, to, up_to, k, mask, seeds.ctypes.data)

    def write_segments(self, k, mask, seeds):
        seg = np.empty(2 * (len(self) + 16) + 1, dtype=np.int64)
        n = lib().dpo_byte_write_segments(self.h, k, mask, seeds.ctypes.data, seg.ctypes.data)
        return seg[:n].copy()

    def short_kmers(self, k, collapse):
        out = np.empty(len(self) + 1, dtype=np.uint16)
        n = lib().dpo_byte_short_kmers(self.h, k, int(collapse), out.ctypes.data)
        return out[:n].copy()


def kmer_value(s):
    a = _u8(s)
    return lib().dpo_kmer_value(a.ctypes.data, a.size)


def pack_bytes(s, out_len):
    a = _u8(s)
    out = np.zeros(out_len, dtype=np.uint8)
    lib().dpo_pack_bytes(a.ctypes.data, a.size, out.ctypes.data)
    return out


class IntSet:
    def __init__(self):
        self.h = lib().dpo_intset_new()

    def __del__(self):
        if getattr(self, "h", None):
            lib().dpo_intset_free(self.h)
            self.h = None

    def add(self, x):
        lib().dpo_intset_add(self.h, x)

    def contains(self, x):
        return bool(lib().dpo_intset_contains(self.h, x))

    def size(self):
        return lib().dpo_intset_size(self.h)

    def count_intersection(self, other):
        return lib().dpo_intset_count_intersection(self.h, other.h)

    def count_intersection_to(self, other, max_count):
        r = lib().dpo_intset_count_intersection_to(self.h, other.h, max_count)
        if r < 0:
            raise RuntimeError(_err())
        return r


def get_shared_ids(sets, min_count, fast):
    arr = (c_vp * len(sets))(*[s.h for s in sets])
    cap = 1 << 20
    out = np.empty(cap, dtype=np.uint64)
    n = lib().dpo_get_shared_ids(arr, len(sets), min_count, int(fast), out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError(_err())
    return out[:n].copy()


class SeedIndex:
    """A bare seeds.SeedIndex with AddSeeds (seeds/seeds.go:62-156): groundwork for the overlap path, not used by map."""

    def __init__(self, k):
        lib().dpo_seedindex_new.restype = c_vp
        self.h = lib().dpo_seedindex_new(int(k))
        self.k = int(k)
        if not self.h:
            raise RuntimeError(_err())

    def __del__(self):
        try:
            lib().dpo_seedindex_free(c_vp(self.h))
        except Exception:
            pass

    def add_seeds(self, seq, min_seeds, ranks, quality=None):
        a = _u8(seq)
        ranks = np.ascontiguousarray(ranks, dtype=np.float64)
        assert ranks.size == 4 ** self.k
        q = None if quality is None else np.ascontiguousarray(quality, dtype=np.uint8)
        assert q is None or q.size == a.size
        rc = lib().dpo_seedindex_add_seeds(c_vp(self.h), a.ctypes.data_as(ctypes.c_char_p), ctypes.c_longlong(a.size),
                                           ctypes.c_longlong(int(min_seeds)), ranks.ctypes.data_as(c_vp),
                                           None if q is None else q.ctypes.data_as(c_vp))
        if rc != 0:
            raise RuntimeError(_err())

    def seed_sequence(self, seq):
        """NewSeedSequence (seeds.go:33-50): (segments [gap, seed, gap, ..., gap], length, offset, inset)."""
        a = _u8(seq)
        lib().dpo_seedindex_seed_sequence.restype = ctypes.c_longlong
        cap = 2 * a.size + 3
        out = np.empty(cap, dtype=np.int64)
        f = np.zeros(3, dtype=np.int64)
        n = lib().dpo_seedindex_seed_sequence(c_vp(self.h), a.ctypes.data_as(ctypes.c_char_p), ctypes.c_longlong(a.size),
                                              out.ctypes.data_as(c_vp), ctypes.c_longlong(cap), f.ctypes.data_as(c_vp))
        if n < 0:
            raise RuntimeError(_err())
        return out[: int(n)].copy(), int(f[0]), int(f[1]), int(f[2])

    def reverse_complement(self, segments):
        """SeedSequence.ReverseComplement (seeds/sequence.go:134-159) on raw segments."""
        seg = np.ascontiguousarray(segments, dtype=np.int64)
        out = np.empty_like(seg)
        lib().dpo_seedindex_rc.restype = ctypes.c_longlong
        if lib().dpo_seedindex_rc(c_vp(self.h), seg.ctypes.data_as(c_vp), ctypes.c_longlong(seg.size), out.ctypes.data_as(c_vp)) < 0:
            raise RuntimeError(_err())
        return out

    def seeds(self):
        lib().dpo_seedindex_seeds.restype = ctypes.c_longlong
        n = lib().dpo_seedindex_seeds(c_vp(self.h), None, ctypes.c_longlong(0))
        out = np.empty(max(int(n), 1), dtype=np.int64)
        lib().dpo_seedindex_seeds(c_vp(self.h), out.ctypes.data_as(c_vp), ctypes.c_longlong(int(n)))
        return out[: int(n)].copy()


def chunk_seed_sequence(segments, length, chunk_size, min_seeds, overlap, k):
    """chunkWorker (overlap/overlap.go:253-318) for one seed sequence: [(segments, length, offset, inset), ...]."""
    seg = np.ascontiguousarray(segments, dtype=np.int64)
    lib().dpo_chunk_seed_sequence.restype = ctypes.c_longlong
    cap = 8 * seg.size + 64
    out = np.empty(cap, dtype=np.int64)
    npieces = ctypes.c_longlong(0)
    w = lib().dpo_chunk_seed_sequence(seg.ctypes.data_as(c_vp), ctypes.c_longlong(seg.size), ctypes.c_longlong(int(length)),
                                      ctypes.c_longlong(int(chunk_size)), ctypes.c_longlong(int(min_seeds)),
                                      ctypes.c_longlong(int(overlap)), ctypes.c_longlong(int(k)), out.ctypes.data_as(c_vp),
                                      ctypes.c_longlong(cap), ctypes.byref(npieces))
    if w < 0:
        raise RuntimeError(_err())
    assert w <= cap
    pieces, at = [], 0
    for _ in range(npieces.value):
        ns, ln, off, ins = (int(x) for x in out[at:at + 4])
        pieces.append((out[at + 4:at + 4 + ns].copy(), ln, off, ins))
        at += 4 + ns
    return pieces


def pairwise_alignments(a_segments, b_segments, min_matches, k, max_length=500):
    """seedAligner.PairwiseAlignments (seeds/alignment.go:426-616) with aSet/bSet = the seeds of a/b:
    [(MatchA, MatchB), ...] in the order returned. Raises RuntimeError where the reference would panic."""
    a = np.ascontiguousarray(a_segments, dtype=np.int64)
    b = np.ascontiguousarray(b_segments, dtype=np.int64)
    lib().dpo_pairwise_alignments.restype = ctypes.c_longlong
    cap = 1 << 20
    out = np.empty(cap, dtype=np.int64)
    nm = ctypes.c_longlong(0)
    w = lib().dpo_pairwise_alignments(a.ctypes.data_as(c_vp), ctypes.c_longlong(a.size), b.ctypes.data_as(c_vp),
                                      ctypes.c_longlong(b.size), ctypes.c_longlong(int(min_matches)), ctypes.c_longlong(int(k)),
                                      ctypes.c_longlong(int(max_length)), out.ctypes.data_as(c_vp), ctypes.c_longlong(cap),
                                      ctypes.byref(nm))
    if w < 0:
        raise RuntimeError(_err())
    res, at = [], 0
    for _ in range(nm.value):
        n = int(out[at])
        res.append((out[at + 1:at + 1 + n].copy(), out[at + 1 + n:at + 1 + 2 * n].copy()))
        at += 1 + 2 * n
    return res


def match(seq_segments, query_segments, min_match, k, reduced=True):
    """SeedSequence.Match (seeds/sequence.go:361-394) of `seq` against `query` on explicit segment lists. reduced=True: as
    performMapping calls it (querySet / seqSet = the seeds of the query / of seq); False: dynamicMatch on the sequences as
    they are. Returns None when Match returns nil, else [(MatchA, MatchB, coveredA, coveredB), ...] (GetBasesCovered)."""
    sq = np.ascontiguousarray(seq_segments, dtype=np.int64)
    q = np.ascontiguousarray(query_segments, dtype=np.int64)
    lib().dpo_match.restype = ctypes.c_longlong
    cap = 1 << 16
    out = np.empty(cap, dtype=np.int64)
    nm = ctypes.c_longlong(0)
    w = lib().dpo_match(sq.ctypes.data_as(c_vp), ctypes.c_longlong(sq.size), q.ctypes.data_as(c_vp), ctypes.c_longlong(q.size),
                        ctypes.c_longlong(int(min_match)), ctypes.c_longlong(int(k)), ctypes.c_int(1 if reduced else 0),
                        out.ctypes.data_as(c_vp), ctypes.c_longlong(cap), ctypes.byref(nm))
    if w < 0:
        raise RuntimeError(_err())
    if nm.value < 0:
        return None
    res, at = [], 0
    for _ in range(nm.value):
        n = int(out[at])
        res.append((list(out[at + 1:at + 1 + n]), list(out[at + 1 + n:at + 1 + 2 * n]), int(out[at + 1 + 2 * n]),
                    int(out[at + 2 + 2 * n])))
        at += 3 + 2 * n
    return res


def reduced(segments, whitelist, k, min_seeds):
    """SeedSequence.Reduced (seeds/sequence.go:85-123): (segments, index) or None."""
    sg = np.ascontiguousarray(segments, dtype=np.int64)
    wl = np.ascontiguousarray(whitelist, dtype=np.int64)
    out = np.empty(sg.size, dtype=np.int64)
    idx = np.empty(sg.size, dtype=np.int64)
    lib().dpo_reduced.restype = ctypes.c_longlong
    n = lib().dpo_reduced(sg.ctypes.data_as(c_vp), ctypes.c_longlong(sg.size), wl.ctypes.data_as(c_vp),
                          ctypes.c_longlong(wl.size), ctypes.c_longlong(int(k)), ctypes.c_longlong(int(min_seeds)),
                          out.ctypes.data_as(c_vp), idx.ctypes.data_as(c_vp))
    if n == -1:
        return None
    if n < 0:
        raise RuntimeError(_err())
    return list(out[:n]), list(idx[:n // 2])


def seed_offset(segments, index, k, from_end=False):
    """GetSeedOffset / GetSeedOffsetFromEnd (seeds/sequence.go:1239-1246, 1269-1276)."""
    sg = np.ascontiguousarray(segments, dtype=np.int64)
    lib().dpo_seed_offset.restype = ctypes.c_longlong
    return int(lib().dpo_seed_offset(sg.ctypes.data_as(c_vp), ctypes.c_longlong(sg.size), ctypes.c_longlong(int(index)),
                                     ctypes.c_longlong(int(k)), ctypes.c_int(1 if from_end else 0)))


def pair_ends(ref_len, circular, query_len, hits_a, hits_b):
    """mapEnds' pairing step (mapping/mapping.go:167-203) on explicit hits, rows {Start, End, QueryOffset, QueryInset, RC,
    ids}: (remainingA, remainingB, matched) as lists of rows; matched is None where the reference returns nil."""
    a = np.ascontiguousarray(hits_a, dtype=np.int64).reshape(-1, 6)
    b = np.ascontiguousarray(hits_b, dtype=np.int64).reshape(-1, 6)
    out = np.zeros(6 * (len(a) + len(b)) + 6, dtype=np.int64)
    cnt = np.zeros(3, dtype=np.int64)
    if lib().dpo_pair_ends(ctypes.c_longlong(int(ref_len)), ctypes.c_int(1 if circular else 0), ctypes.c_longlong(int(query_len)),
                           a.ctypes.data_as(c_vp), ctypes.c_longlong(len(a)), b.ctypes.data_as(c_vp), ctypes.c_longlong(len(b)),
                           out.ctypes.data_as(c_vp), cnt.ctypes.data_as(c_vp)):
        raise RuntimeError(_err())
    rows = out.reshape(-1, 6)
    na, nb, nm = int(cnt[0]), int(cnt[1]), int(cnt[2])
    ra = [list(map(int, r)) for r in rows[:na]]
    rb = [list(map(int, r)) for r in rows[na:na + nb]]
    return ra, rb, (None if nm < 0 else [list(map(int, r)) for r in rows[na + nb:na + nb + nm]])


def matches(chunk_seeds, num_seeds, query_seeds, hit_fraction):
    """SeedIndex.Matches (seeds/seeds.go:335-353) of a query (its seeds in order) against chunks given as seed lists (gaps are
    irrelevant to it and set to 1): the candidate chunk ids."""
    def seg(seeds):
        out = [1]
        for x in seeds:
            out += [int(x), 1]
        return out
    segs, off = [], [0]
    for c in chunk_seeds:
        segs += seg(c)
        off.append(len(segs))
    cs = np.ascontiguousarray(segs, dtype=np.int64)
    co = np.ascontiguousarray(off, dtype=np.int64)
    q = np.ascontiguousarray(seg(query_seeds), dtype=np.int64)
    out = np.zeros(max(1, len(chunk_seeds)), dtype=np.int64)
    lib().dpo_matches.restype = ctypes.c_longlong
    n = lib().dpo_matches(cs.ctypes.data_as(c_vp), co.ctypes.data_as(c_vp), ctypes.c_longlong(len(chunk_seeds)),
                          ctypes.c_longlong(int(num_seeds)), q.ctypes.data_as(c_vp), ctypes.c_longlong(q.size),
                          ctypes.c_double(float(hit_fraction)), out.ctypes.data_as(c_vp), ctypes.c_longlong(out.size))
    if n < 0:
        raise RuntimeError(_err())
    return [int(x) for x in out[:n]]


def gap_range(gap, k):
    """gapRange (seeds/alignment.go:411-424): (minGap, maxGap)."""
    out = np.zeros(2, dtype=np.int64)
    lib().dpo_gap_range(ctypes.c_longlong(int(gap)), ctypes.c_longlong(int(k)), out.ctypes.data_as(c_vp))
    return int(out[0]), int(out[1])


def kmer_values(ref, k):
    """values[] of commands/map.go:45-71 for a single-record reference (canonical tie order, Q10)."""
    a = _u8(ref)
    out = np.empty(4 ** k, dtype=np.float64)
    if lib().dpo_kmer_values(a.ctypes.data, a.size, k, out.ctypes.data):
        raise RuntimeError(_err())
    return out


def kmer_counts(ref, k):
    a = _u8(ref)
    out = np.empty(4 ** k, dtype=np.uint64)
    if lib().dpo_kmer_counts(a.ctypes.data, a.size, k, out.ctypes.data):
        raise RuntimeError(_err())
    return out


COUNTER_NAMES = ["windows", "kmer_lookups", "query_seeds", "posting_runs", "posting_entries", "candidates",
                 "cand_pass", "chain_cells", "chains", "mappings", "sort_ties_unpinned"]


class Mapper:
    """mapping.Mapper (mapping/mapping.go:22-26) over the oracle."""

    def __init__(self, ref, values, circular=True, k=11, seed_rate=40, edge_size=1000, chunk_size=10000, lean=None,
                 threads=None):
        """lean: None = decided by the size the index bitsets would have, True/False = forced (memory-lean index:
        the bitsets are rebuilt per query from lists, everything downstream runs unchanged; oracle.hpp).
        threads: workers of the per-chunk seed scans of the index build (default: all cores)."""
        self.ref = _u8(ref)
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        assert self.values.size == 4 ** k
        self.k = k
        self.circular = circular
        self.h = lib().dpo_mapper_new_ex(self.ref.ctypes.data, self.ref.size, int(circular), k, self.values.ctypes.data,
                                         seed_rate, edge_size, chunk_size, -1 if lean is None else int(bool(lean)),
                                         threads or (os.cpu_count() or 1))
        if not self.h:
            raise RuntimeError(_err())

    @property
    def lean(self):
        return bool(lib().dpo_mapper_is_lean(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().dpo_mapper_free(self.h)
            self.h = None

    @property
    def num_seeds(self):
        return lib().dpo_mapper_num_seeds(self.h)

    @property
    def num_chunks(self):
        return lib().dpo_mapper_num_chunks(self.h)

    def seed_kmers(self):
        out = np.empty(self.num_seeds, dtype=np.int64)
        lib().dpo_mapper_seed_kmers(self.h, out.ctypes.data)
        return out

    def chunk(self, c):
        f = np.zeros(4, dtype=np.int64)
        n = lib().dpo_mapper_chunk(self.h, c, f.ctypes.data, None, 0)
        seg = np.empty(n, dtype=np.int64)
        lib().dpo_mapper_chunk(self.h, c, f.ctypes.data, seg.ctypes.data, n)
        return dict(offset=int(f[0]), inset=int(f[1]), length=int(f[2]), nseeds=int(f[3]), segments=seg)

    def window_segments(self, read, start=0, end=0, whole=False, rc=False):
        a = _u8(read)
        seg = np.empty(2 * (a.size + 16) + 1, dtype=np.int64)
        f = np.zeros(3, dtype=np.int64)
        n = lib().dpo_window_segments(self.h, a.ctypes.data, a.size, start, end, int(whole), int(rc), seg.ctypes.data,
                                      seg.size, f.ctypes.data)
        if n < 0:
            raise RuntimeError(_err())
        return seg[:n].copy(), dict(offset=int(f[0]), inset=int(f[1]), length=int(f[2]))

    def window_candidates(self, read, start=0, end=0, whole=False, rc=False):
        a = _u8(read)
        out = np.empty(max(16, self.num_chunks), dtype=np.int64)
        n = lib().dpo_window_candidates(self.h, a.ctypes.data, a.size, start, end, int(whole), int(rc), out.ctypes.data,
                                        out.size)
        if n < 0:
            raise RuntimeError(_err())
        return out[:n].copy()

    def window_mappings(self, read, start=0, end=0, whole=False):
        a = _u8(read)
        cap = 4096
        out = np.empty((cap, 6), dtype=np.int64)
        n = lib().dpo_window_mappings(self.h, a.ctypes.data, a.size, start, end, int(whole), out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError(_err())
        return out[:n].copy()

    def map_batch(self, bases, offsets, threads=1):
        """-> (rows[int64, M x 6] = Start, End, QueryOffset, QueryInset, RC, ids; out_offsets[n+1]; counters dict)"""
        bases = _u8(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        out_off = np.empty(n + 1, dtype=np.int64)
        ctr = np.zeros(len(COUNTER_NAMES), dtype=np.int64)
        rows_p = ctypes.POINTER(c_ll)()
        rc = lib().dpo_map_batch(self.h, n, bases.ctypes.data, offsets.ctypes.data, threads, ctypes.byref(rows_p),
                                 out_off.ctypes.data, ctr.ctypes.data)
        if rc:
            raise RuntimeError(_err())
        total = int(out_off[n])
        rows = np.ctypeslib.as_array(rows_p, shape=(max(total, 1) * 6,))[: total * 6].reshape(total, 6).copy()
        lib().dpo_free(rows_p)
        return rows, out_off, dict(zip(COUNTER_NAMES, (int(x) for x in ctr)))


OVERLAP_DEFAULTS = dict(overlap_size=1000, k=10, num_seeds=15, seed_batch_size=10000, chunk_size=10000,
                        query_batch_size=20000, min_hits=0.25)


def overlap_values(bases, offsets, k):
    """getKmerValues (commands/overlap.go:41-95) without a seed_values file: the formula and the TopOccurrences cut of
    `map` (kmer_values), counted over ALL reads."""
    bases = _u8(bases)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    counts = np.zeros(4 ** k, dtype=np.uint64)
    if lib().dpo_kmer_counts_batch(bases.ctypes.data_as(c_vp), offsets.ctypes.data_as(c_vp), offsets.size - 1, k,
                                   counts.ctypes.data_as(c_vp)) != 0:
        raise RuntimeError(_err())
    return kmer_values_from_counts(counts, k)


def kmer_values_from_counts(counts, k):
    counts = np.ascontiguousarray(counts, dtype=np.uint64).copy()
    out = np.zeros(4 ** k, dtype=np.float64)
    if lib().dpo_kmer_values_from_counts(counts.ctypes.data_as(c_vp), k, out.ctypes.data_as(c_vp)) != 0:
        raise RuntimeError(_err())
    return out


class OverlapRound:
    """One round of `downpore overlap` up to the seed-match stream (oracle/overlap.cpp)."""

    def __init__(self, bases, offsets, values, first_sequence=0, ignore=None, **params):
        p = dict(OVERLAP_DEFAULTS)
        p.update(params)
        self.params = p
        bases = _u8(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        values = np.ascontiguousarray(values, dtype=np.float64)
        assert values.size == 4 ** p["k"]
        ign = None if ignore is None else np.ascontiguousarray(ignore, dtype=np.uint8)
        p6 = np.array([p["overlap_size"], p["k"], p["num_seeds"], p["seed_batch_size"], p["chunk_size"], p["query_batch_size"]],
                      dtype=np.int64)
        self.h = lib().dpo_overlap_round(bases.ctypes.data_as(c_vp), offsets.ctypes.data_as(c_vp), n,
                                         None if ign is None else ign.ctypes.data_as(c_vp), first_sequence,
                                         values.ctypes.data_as(c_vp), p6.ctypes.data_as(c_vp), float(p["min_hits"]))
        if not self.h:
            raise RuntimeError(_err())
        hd = self._get(0)
        self.num_seeds, self.num_queries, self.num_chunks, self.num_hits, self.num_query_seqs, self.next_first_sequence = (int(x) for x in hd)
        self.seed_kmers = self._get(1)
        v, at, self.queries = self._get(2), 0, []
        for _ in range(self.num_queries):
            qid, sid, rc, ln, off, ins, ns = (int(x) for x in v[at:at + 7])
            self.queries.append(dict(id=qid, sequence_id=sid, rc=bool(rc), length=ln, offset=off, inset=ins, segments=v[at + 7:at + 7 + ns].copy()))
            at += 7 + ns
        v, at, self.chunks = self._get(3), 0, []
        for _ in range(self.num_chunks):
            rid, ln, off, ins, ns = (int(x) for x in v[at:at + 5])
            self.chunks.append(dict(read=rid, length=ln, offset=off, inset=ins, segments=v[at + 5:at + 5 + ns].copy()))
            at += 5 + ns
        v, at, self.hits = self._get(4), 0, []
        for _ in range(self.num_hits):
            qid, rc, tgt, m = (int(x) for x in v[at:at + 4])
            self.hits.append(dict(query_id=qid, rc=bool(rc), target=tgt, match_a=v[at + 4:at + 4 + m].copy(), match_b=v[at + 4 + m:at + 4 + 2 * m].copy()))
            at += 4 + 2 * m
        lib().dpo_overlap_free(c_vp(self.h))
        self.h = None

    def _get(self, what):
        n = lib().dpo_overlap_get(c_vp(self.h), what, None, 0)
        out = np.zeros(max(int(n), 1), dtype=np.int64)
        lib().dpo_overlap_get(c_vp(self.h), what, out.ctypes.data_as(c_vp), int(n))
        return out[:int(n)]


def parse_fasta(content, min_length=0):
    """readFasta's first pass (sequence/seqio.go:188-267) -> [(name bytes, sequence bytes)]; raises on the reference's
    log.Fatal (invalid fastq)."""
    content = bytes(content)
    n, nb = c_ll(), c_ll()
    p = lib().dpo_parse_fasta(content, len(content), min_length, ctypes.byref(n), ctypes.byref(nb))
    if not p:
        raise RuntimeError(_err())
    blob = ctypes.string_at(p, nb.value)
    lib().dpo_free(p)
    out, o = [], 0
    for _ in range(n.value):
        a = int.from_bytes(blob[o:o + 8], "little")
        name = blob[o + 8:o + 8 + a]
        o += 8 + a
        b = int.from_bytes(blob[o:o + 8], "little")
        out.append((name, blob[o + 8:o + 8 + b]))
        o += 8 + b
    return out


def paf_lines(rows, out_off, names, lengths, ref_name, ref_len, circular):
    """AsString (mapping/mapping.go:112-122) over map_batch output."""
    lines = []
    for i in range(len(out_off) - 1):
        for r in rows[out_off[i]:out_off[i + 1]]:
            start, end, qoff, qin, rc, ids = (int(v) for v in r)
            ml = end - start
            if circular and ml < 0:
                ml = ref_len - start + end
            lines.append("%s\t%d\t%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t255" % (
                names[i], lengths[i], qoff, lengths[i] - qin, "-" if rc else "+", ref_name, ref_len, start, end, ids, ml))
    return lines
