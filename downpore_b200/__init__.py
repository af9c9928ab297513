"""downpore_b200 — B200-native (sm_100a) implementation of the `downpore map` hot path.

Python mirror of the reference's Go interface for this path (mapping/mapping.go:22-26, :67):

    mapper = NewMapper(reference, circular, k, kmer_values, seed_rate, edge_size, chunk_size)
    mappings, offsets = mapper.map_batch(bases, read_offsets)      # Mapper.Map for every read
    line = mapper.as_string(mapping, query_name, query_len)        # Mapper.AsString

Everything runs through the C ABI of include/downpore_b200.h (libdownpore_b200.so, built from csrc/ by
__graft_entry__.build() or `make -C downpore_b200/csrc`). There is no CPU fallback: importing works anywhere, but any
call fails loudly if the CUDA library is missing or no GPU is usable.
"""
import ctypes
import os
import subprocess
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdownpore_b200.so")
CSRC = os.path.join(_HERE, "csrc")
HOST_PATH = os.path.join(_HERE, "..", "tools", "dp_map")  # the `downpore map` host over the C ABI (tools/dp_map.cpp)

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p

MAPPING_DTYPE = np.dtype([("start", "<i8"), ("end", "<i8"), ("q_offset", "<i4"), ("q_inset", "<i4"), ("ids", "<i4"),
                          ("rc", "u1"), ("pad", "u1", (3,))], align=False)
assert MAPPING_DTYPE.itemsize == 32


class Stats(ctypes.Structure):
    _fields_ = [("ms_total", ctypes.c_double), ("ms_pack", ctypes.c_double), ("ms_extract", ctypes.c_double),
                ("ms_lookup", ctypes.c_double), ("ms_chain", ctypes.c_double), ("ms_host_logic", ctypes.c_double),
                ("ms_h2d", ctypes.c_double), ("rounds", c_i64), ("windows", c_i64), ("kmer_lookups", c_i64),
                ("query_seeds", c_i64), ("posting_runs", c_i64), ("posting_entries", c_i64), ("candidates", c_i64),
                ("chain_cells", c_i64), ("mappings", c_i64), ("kernel_launches", c_i64), ("bases", c_i64),
                ("h2d_bytes", c_i64), ("ms_reduce", ctypes.c_double), ("ms_finish", ctypes.c_double),
                ("retries", c_i64), ("short_reads", c_i64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class OverlapRoundStruct(ctypes.Structure):
    _fields_ = [("num_seeds", c_i64), ("num_queries", c_i64), ("num_query_seqs", c_i64), ("next_first_sequence", c_i64),
                ("num_chunks", c_i64), ("num_hits", c_i64), ("num_matches", c_i64), ("hits", c_vp), ("matches", c_vp),
                ("read_seeds", c_i64), ("chunk_seeds", c_i64), ("seed_postings", c_i64), ("posting_entries", c_i64), ("candidates", c_i64),
                ("pairs", c_i64), ("kernel_launches", c_i64), ("ms_total", ctypes.c_double),
                ("ms_select", ctypes.c_double), ("ms_queries", ctypes.c_double), ("ms_scan", ctypes.c_double),
                ("ms_chunk", ctypes.c_double), ("ms_index", ctypes.c_double), ("ms_lookup", ctypes.c_double),
                ("ms_align", ctypes.c_double), ("ms_collect", ctypes.c_double)]


OVERLAP_HIT_DTYPE = np.dtype([("query_id", "<i4"), ("rc", "<i4"), ("target", "<i4"), ("n", "<i4"), ("at", "<i8")])
assert OVERLAP_HIT_DTYPE.itemsize == 24


def build(force=False):
    """Compile csrc/ for sm_100a into libdownpore_b200.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp"))]
    srcs.append(os.path.join(_HERE, "..", "include", "downpore_b200.h"))
    srcs.append(os.path.join(_HERE, "..", "tools", "dp_map.cpp"))
    newest = max(os.path.getmtime(s) for s in srcs)
    stale = [p for p in (LIB_PATH, HOST_PATH) if not os.path.exists(p) or os.path.getmtime(p) < newest]
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "all"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None

_SIGNATURES = {
    "dp_last_error": (ctypes.c_char_p, []),
    "dp_version": (ctypes.c_char_p, []),
    "dp_free": (None, [c_vp]),
    "dp_probe_gather_gbs": (ctypes.c_int, [ctypes.c_int, c_i64, ctypes.POINTER(ctypes.c_double)]),
    "dp_host_alloc": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_size_t]),
    "dp_host_free": (None, [c_vp]),
    "dp_mapper_create": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "dp_mapper_destroy": (None, [c_vp]),
    "dp_mapper_map_batch_packed": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)]),
    "dp_mapper_index_image_size": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "dp_mapper_index_export": (ctypes.c_int, [c_vp, c_vp, c_i64]),
    "dp_mapper_create_from_index": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "dp_mapper_map_batch": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)]),
    "dp_mapper_map_batch_device": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, ctypes.POINTER(c_vp),
                                                  ctypes.POINTER(c_vp)]),
    "dp_mapper_paf_line": (ctypes.c_int, [c_vp, c_vp, ctypes.c_char_p, c_i64, ctypes.c_char_p, c_vp, ctypes.c_int]),
    "dp_mapper_get_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(Stats)]),
    "dp_mapper_index_info": (ctypes.c_int, [c_vp, c_vp]),
    "dp_mapper_params": (ctypes.c_int, [c_vp, c_vp]),
    "dp_mapper_seed_kmers": (ctypes.c_int, [c_vp, c_vp]),
    "dp_mapper_chunk": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dp_mapper_probe_window": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, ctypes.c_int, c_vp, c_vp, c_vp, c_i64,
                                              c_vp, c_vp, c_i64, c_vp, c_vp, c_i64]),
    "dp_pack": (ctypes.c_int, [c_vp, c_i64, c_vp, ctypes.c_int]),
    "dp_kmer_counts": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, c_vp, ctypes.c_int]),
    "dp_split_records": (ctypes.c_int, [c_vp, c_i64, c_i64, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                        ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "dp_mapper_map_batch_spans": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)]),
    "dp_mapper_paf_block": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, ctypes.c_char_p, ctypes.POINTER(c_vp),
                                           ctypes.POINTER(c_i64)]),
    "dp_device_alloc": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_size_t, ctypes.c_int]),
    "dp_device_free": (None, [c_vp, ctypes.c_int]),
    "dp_device_copy": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_int]),
    "dp_overlapper_create": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "dp_overlapper_kmer_counts": (ctypes.c_int, [c_vp, c_vp]),
    "dp_overlapper_set_values": (ctypes.c_int, [c_vp, c_vp]),
    "dp_overlapper_round": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "dp_overlapper_queries": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "dp_overlapper_chunks": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "dp_overlapper_seed_kmers": (ctypes.c_int, [c_vp, c_vp]),
    "dp_overlapper_destroy": (None, [c_vp]),
}


def exported_symbols():
    """Names declared in include/downpore_b200.h."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("downpore_b200: %s is missing — run __graft_entry__.build() (there is no CPU fallback)"
                               % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class DownporeError(RuntimeError):
    pass


def _check(rc):
    if rc:
        raise DownporeError(lib().dp_last_error().decode())


def _u8(x):
    if isinstance(x, str):
        x = x.encode()
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def _adopt(ptr, nbytes, dtype):
    buf = (ctypes.c_char * nbytes).from_address(ptr.value)
    weakref.finalize(buf, lib().dp_free, ptr.value)
    return np.frombuffer(buf, dtype=dtype)


def _adopt_or_empty(addr, nbytes, dtype):
    """malloc'ed output of the library (address as an int or None) -> numpy array that frees it."""
    if not addr:
        return np.zeros(0, dtype=dtype)
    if nbytes <= 0:
        lib().dp_free(addr)
        return np.zeros(0, dtype=dtype)
    return _adopt(c_vp(addr), nbytes, dtype)


def pack(ascii_seq, device=0):
    """sequence.NewPackedSequence (sequence/sequence.go:67-93) through the device pack kernel -> bytes."""
    a = _u8(ascii_seq)
    out = np.zeros((a.size + 3) // 4, dtype=np.uint8)
    _check(lib().dp_pack(a.ctypes.data, a.size, out.ctypes.data, device))
    return out


def pack_batch(bases, offsets):
    """Host-side stand-in for what the Go reader hands over: every read of an ASCII batch as sequence.packedSequence
    bytes (sequence/sequence.go:67-93: 4 bases per byte, first base in the top bits, tail byte zero-padded), back to back.
    Returns (packed uint8, byte_offsets int64[n+1], lengths int64[n]). numpy only (test and bench input preparation)."""
    bases = _u8(bases)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    lengths = np.diff(offsets)
    nbytes = (lengths + 3) // 4
    byte_off = np.zeros(offsets.size, dtype=np.int64)
    np.cumsum(nbytes, out=byte_off[1:])
    code = (((bases >> 1) ^ ((bases & 4) >> 2)) & 3).astype(np.uint8)  # the pack kernel's base code (asm_amd64.s:33-78)
    out = np.zeros(int(byte_off[-1]), dtype=np.uint8)
    if lengths.size and np.all(lengths == lengths[0]) and lengths[0] % 4 == 0 and offsets[0] == 0:
        c = code[: offsets[-1]].reshape(-1, 4)
        out[:] = (c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]
    else:
        for i in range(lengths.size):
            c = np.zeros(int(nbytes[i]) * 4, dtype=np.uint8)
            c[: lengths[i]] = code[offsets[i]: offsets[i + 1]]
            c = c.reshape(-1, 4)
            out[byte_off[i]: byte_off[i + 1]] = (c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]
    return out, byte_off, lengths.astype(np.int64)


def kmer_counts(ascii_seq, k, counts=None, device=0):
    """sequtil.KmerOccurrences (util/sequtil/kmers.go:34-69) for one record, accumulated into `counts`."""
    a = _u8(ascii_seq)
    if counts is None:
        counts = np.zeros(4 ** k, dtype=np.uint64)
    _check(lib().dp_kmer_counts(a.ctypes.data, a.size, k, counts.ctypes.data, device))
    return counts


def host_alloc(nbytes):
    """dp_host_alloc: page-locked, device-mapped host memory as a uint8 numpy array (freed with the array)."""
    p = c_vp()
    _check(lib().dp_host_alloc(ctypes.byref(p), nbytes))
    buf = (ctypes.c_ubyte * nbytes).from_address(p.value)
    weakref.finalize(buf, lib().dp_host_free, p.value)
    return np.frombuffer(buf, dtype=np.uint8)


RECORD_DTYPE = np.dtype([("name_start", "<i8"), ("name_len", "<i8"), ("seq_start", "<i8"), ("seq_len", "<i8")])


class DeviceImage:
    """A file image (or a piece of one) uploaded once into device memory (dp_device_alloc / dp_device_copy)."""

    def __init__(self, data, device=0):
        a = _u8(data)
        self.nbytes, self.device = int(a.size), device
        p = c_vp()
        _check(lib().dp_device_alloc(ctypes.byref(p), max(self.nbytes, 1), device))
        self.ptr = p.value
        self._fin = weakref.finalize(self, lib().dp_device_free, p.value, device)
        _check(lib().dp_device_copy(self.ptr, a.ctypes.data, self.nbytes, device))


def _image_ptr(image):
    if isinstance(image, DeviceImage):
        return image.ptr, image.nbytes, None
    if isinstance(image, tuple):  # (raw pointer, bytes)
        return int(image[0]), int(image[1]), None
    a = _u8(image)
    return a.ctypes.data, int(a.size), a


def split_records(image, min_length=0, final=True, is_fastq=0, device=0):
    """readFasta's record rules (sequence/seqio.go:188-267) on the device: the (name, sequence) spans of a FASTA/FASTQ file
    image. `image`: bytes / numpy uint8 / DeviceImage / (pointer, nbytes). Returns (records[RECORD_DTYPE], consumed,
    is_fastq): with final=False only the records in front of the piece's last name line, `consumed` = where it starts."""
    ptr, n, keep = _image_ptr(image)
    rec_p, nrec, cons, fq = c_vp(), c_i64(), c_i64(), ctypes.c_int(int(is_fastq))
    _check(lib().dp_split_records(ptr, n, int(min_length), int(bool(final)), ctypes.byref(fq), device, ctypes.byref(rec_p),
                                  ctypes.byref(nrec), ctypes.byref(cons)))
    k = int(nrec.value)
    if k:
        recs = _adopt(rec_p, k * RECORD_DTYPE.itemsize, RECORD_DTYPE)
    else:
        lib().dp_free(rec_p)
        recs = np.zeros(0, dtype=RECORD_DTYPE)
    return recs, int(cons.value), int(fq.value)


def probe_gather_gbs(table_bytes=8 << 30, device=0):
    """Random 32 B-sector gather bandwidth (GB/s of sector traffic) over a table >> L2: the HBM gather roofline."""
    out = ctypes.c_double()
    _check(lib().dp_probe_gather_gbs(device, table_bytes, ctypes.byref(out)))
    return out.value


def kmer_values(counts, k):
    """values[] of commands/map.go:46-71 from a k-mer histogram.

    Host-side numpy: in the drop-in this stays in the Go host (the top-1 % cut depends on Go's sort.Sort tie order,
    SURVEY Q10). Tie order here: (count, k-mer id) ascending, the same canonical order the oracle uses.
    """
    counts = np.asarray(counts, dtype=np.uint64).copy()
    n = counts.size
    tot = float(counts.sum(dtype=np.uint64))
    freq = counts.astype(np.float64) / tot
    target = 0.000005
    values = np.where(freq <= target, 1.0 - (target - freq), 1.0 - (freq - target))
    values[counts < 3] = 0.0
    # TopOccurrences (util/sequtil/kmers.go:87-112): in-place forward/rc merge over ascending ids
    ids = np.arange(n, dtype=np.int64)
    rc = np.zeros(n, dtype=np.int64)
    x = ids.copy()
    for _ in range(k):
        rc = (rc << 2) | ((x ^ 3) & 3)
        x >>= 2
    merged = counts + counts[rc]
    merged = np.where(rc == ids, merged, 2 * merged)  # a non-palindromic pair is summed twice by the in-place loop
    order = np.argsort(merged, kind="stable")
    top = order[n - n // 100:]
    values[top] = 0.0
    values[0] = 0.0
    return values


class Mapper:
    """mapping.Mapper (mapping/mapping.go:22-26) bound to one CUDA device."""

    def __init__(self, reference, kmer_values, circular=True, k=11, seed_rate=40, edge_size=1000, chunk_size=10000,
                 device=0, ref_name="ref"):
        self._h = None
        ref = _u8(reference)
        vals = np.ascontiguousarray(kmer_values, dtype=np.float64)
        if vals.size != 4 ** k:
            raise ValueError("kmer_values must hold 4^k doubles")
        h = c_vp()
        _check(lib().dp_mapper_create(ref.ctypes.data, ref.size, int(bool(circular)), k, vals.ctypes.data, seed_rate,
                                      edge_size, chunk_size, device, ctypes.byref(h)))
        self._h = h
        self.k = k
        self.circular = bool(circular)
        self.ref_len = int(ref.size)
        self.ref_name = ref_name
        self.edge_size = edge_size
        self.device = device

    # ---- index image: replicate over GPUs / persist (dp_mapper_index_export, dp_mapper_create_from_index) ----
    @classmethod
    def from_index(cls, image_ptr, nbytes, device=0, ref_name="ref"):
        """Open a mapper on `device` from an index image at address `image_ptr` (device or host memory)."""
        h = c_vp()
        _check(lib().dp_mapper_create_from_index(image_ptr, nbytes, device, ctypes.byref(h)))
        self = cls.__new__(cls)
        self._h = h
        self.device = device
        self.ref_name = ref_name
        info = np.zeros(8, dtype=np.int64)
        _check(lib().dp_mapper_params(h, info.ctypes.data))
        self.k, self.circular, self.ref_len, self.edge_size = int(info[0]), bool(info[1]), int(info[2]), int(info[3])
        return self

    def index_image_size(self):
        n = c_i64()
        _check(lib().dp_mapper_index_image_size(self._h, ctypes.byref(n)))
        return int(n.value)

    def export_index(self, image_ptr, nbytes):
        """Write the index image to `image_ptr` (device or host memory, at least index_image_size() bytes)."""
        _check(lib().dp_mapper_index_export(self._h, image_ptr, nbytes))

    def save_index(self, path):
        """On-disk index: the image bytes, as written by export_index into host memory."""
        n = self.index_image_size()
        buf = np.empty(n, dtype=np.uint8)
        self.export_index(buf.ctypes.data, n)
        buf.tofile(path)

    @classmethod
    def load_index(cls, path, device=0, ref_name="ref"):
        buf = np.fromfile(path, dtype=np.uint8)
        return cls.from_index(buf.ctypes.data, buf.size, device=device, ref_name=ref_name)

    def close(self):
        if self._h:
            lib().dp_mapper_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _collect(self, n, out_p, off_p):
        """Wraps the two malloc'ed result buffers as numpy arrays without copying; dp_free runs when the arrays die."""
        off = _adopt(off_p, (n + 1) * 8, np.int64)
        total = int(off[n])
        if total:
            maps = _adopt(out_p, total * MAPPING_DTYPE.itemsize, MAPPING_DTYPE)
        else:
            lib().dp_free(out_p)
            maps = np.zeros(0, dtype=MAPPING_DTYPE)
        return maps, off

    def map_batch(self, bases, offsets):
        """Mapper.Map over a batch. bases: concatenated ASCII (numpy uint8 or a pinned torch tensor's numpy view);
        offsets: n+1 int64. Returns (mappings[MAPPING_DTYPE], out_offsets[n+1])."""
        bases = _u8(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        out_p, off_p = c_vp(), c_vp()
        _check(lib().dp_mapper_map_batch(self._h, n, bases.ctypes.data, offsets.ctypes.data, ctypes.byref(out_p),
                                         ctypes.byref(off_p)))
        return self._collect(n, out_p, off_p)

    def map_batch_packed(self, packed, byte_offsets, lengths):
        """Mapper.Map over reads that are packed already (sequence.packedSequence bytes, see pack_batch). `packed`: numpy
        uint8 array, or an int (raw host or device pointer, e.g. a pinned or CUDA torch tensor's data_ptr())."""
        byte_offsets = np.ascontiguousarray(byte_offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        n = lengths.size
        if not isinstance(packed, int):
            packed = _u8(packed)
            self._keep = packed
            packed = packed.ctypes.data
        out_p, off_p = c_vp(), c_vp()
        _check(lib().dp_mapper_map_batch_packed(self._h, n, packed, byte_offsets.ctypes.data, lengths.ctypes.data,
                                                ctypes.byref(out_p), ctypes.byref(off_p)))
        return self._collect(n, out_p, off_p)

    def map_batch_spans(self, image, records):
        """Mapper.Map over reads mapped where they lie in a file image (records of split_records)."""
        ptr, _, keep = _image_ptr(image)
        records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        n = records.size
        out_p, off_p = c_vp(), c_vp()
        _check(lib().dp_mapper_map_batch_spans(self._h, n, ptr, records.ctypes.data, ctypes.byref(out_p),
                                               ctypes.byref(off_p)))
        return self._collect(n, out_p, off_p)

    def paf_block(self, image, records, maps, out_off):
        """Mapper.AsString for every mapping of a batch, formatted on the device -> bytes (lines end in a newline)."""
        ptr, _, keep = _image_ptr(image)
        records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        maps = np.ascontiguousarray(maps, dtype=MAPPING_DTYPE)
        out_off = np.ascontiguousarray(out_off, dtype=np.int64)
        txt_p, nb = c_vp(), c_i64()
        _check(lib().dp_mapper_paf_block(self._h, records.size, ptr, records.ctypes.data, maps.ctypes.data,
                                         out_off.ctypes.data, self.ref_name.encode(), ctypes.byref(txt_p),
                                         ctypes.byref(nb)))
        try:
            return ctypes.string_at(txt_p.value, nb.value)
        finally:
            lib().dp_free(txt_p)

    def map_batch_ptr(self, host_ptr, offsets):
        """Same as map_batch for a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        out_p, off_p = c_vp(), c_vp()
        _check(lib().dp_mapper_map_batch(self._h, n, host_ptr, offsets.ctypes.data, ctypes.byref(out_p),
                                         ctypes.byref(off_p)))
        return self._collect(n, out_p, off_p)

    def map_batch_device(self, device_ptr, offsets):
        """Same, with the ASCII reads already resident on this mapper's device (device_ptr: int)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        out_p, off_p = c_vp(), c_vp()
        _check(lib().dp_mapper_map_batch_device(self._h, n, device_ptr, offsets.ctypes.data, ctypes.byref(out_p),
                                                ctypes.byref(off_p)))
        return self._collect(n, out_p, off_p)

    def as_string(self, mapping, query_name, query_len):
        """Mapper.AsString (mapping/mapping.go:112-122): one PAF line."""
        rec = np.zeros(1, dtype=MAPPING_DTYPE)
        rec[0] = mapping
        buf = ctypes.create_string_buffer(1024)
        n = lib().dp_mapper_paf_line(self._h, rec.ctypes.data, query_name.encode(), query_len, self.ref_name.encode(),
                                     buf, 1024)
        if n < 0:
            raise DownporeError("PAF line too long")
        return buf.value.decode()

    def paf_lines(self, maps, out_off, names, lengths):
        lines = []
        for i in range(len(out_off) - 1):
            for j in range(out_off[i], out_off[i + 1]):
                lines.append(self.as_string(maps[j], names[i], int(lengths[i])))
        return lines

    def stats(self):
        s = Stats()
        _check(lib().dp_mapper_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    def index_info(self):
        out = np.zeros(5, dtype=np.int64)
        _check(lib().dp_mapper_index_info(self._h, out.ctypes.data))
        return dict(num_seeds=int(out[0]), num_chunks=int(out[1]), chunk_postings=int(out[2]),
                    seed_postings=int(out[3]), index_bytes=int(out[4]))

    # ---- test probes ----
    def seed_kmers(self):
        out = np.zeros(self.index_info()["num_seeds"], dtype=np.int64)
        _check(lib().dp_mapper_seed_kmers(self._h, out.ctypes.data))
        return out

    def chunk(self, c):
        f = np.zeros(4, dtype=np.int64)
        _check(lib().dp_mapper_chunk(self._h, c, f.ctypes.data, None, None))
        n = int(f[3])
        pos = np.zeros(max(n, 1), dtype=np.int32)
        kmer = np.zeros(max(n, 1), dtype=np.int64)
        _check(lib().dp_mapper_chunk(self._h, c, f.ctypes.data, pos.ctypes.data, kmer.ctypes.data))
        return dict(offset=int(f[0]), inset=int(f[1]), length=int(f[2]), nseeds=n, pos=pos[:n], kmer=kmer[:n])

    def probe_window(self, read, start=0, end=0, whole=False):
        """One performMapping call: per-strand seeds, candidates, and the window's mappings."""
        a = _u8(read)
        cap = 2 * (a.size + 8)
        ns = np.zeros(2, dtype=np.int32)
        spos = np.zeros(cap, dtype=np.int32)
        skmer = np.zeros(cap, dtype=np.int64)
        nc = np.zeros(2, dtype=np.int32)
        ccap = 4096
        cand = np.zeros(ccap, dtype=np.int32)
        nm = np.zeros(1, dtype=np.int32)
        maps = np.zeros(256, dtype=MAPPING_DTYPE)
        _check(lib().dp_mapper_probe_window(self._h, a.ctypes.data, a.size, start, end, int(whole), ns.ctypes.data,
                                            spos.ctypes.data, skmer.ctypes.data, cap, nc.ctypes.data, cand.ctypes.data,
                                            ccap, nm.ctypes.data, maps.ctypes.data, 256))
        f, r = int(ns[0]), int(ns[1])
        return dict(seeds=[(spos[:f].copy(), skmer[:f].copy()), (spos[f:f + r].copy(), skmer[f:f + r].copy())],
                    candidates=[cand[:nc[0]].copy(), cand[nc[0]:nc[0] + nc[1]].copy()], mappings=maps[:nm[0]].copy())


def NewMapper(reference, circular, k, kmer_values, seed_rate, edge_size, chunk_size, num_workers=0, device=0):
    """mapping.NewMapper (mapping/mapping.go:67). num_workers is accepted for signature parity and ignored: the
    goroutine pool is replaced by batched GPU rounds."""
    return Mapper(reference, kmer_values, circular=circular, k=k, seed_rate=seed_rate, edge_size=edge_size,
                  chunk_size=chunk_size, device=device)


def replicate_index(mapper, src=0, device=0, group=None, ref_name="ref"):
    """Index replication for read-sharded multi-GPU mapping: the rank `src` holds a built `mapper`; its
    index image is broadcast ONCE over the process group (NCCL over NVLink when the group is an NCCL group) and every
    other rank opens a mapper from it. Returns this rank's mapper. No other collective exists on the map path."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
    size = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        size[0] = mapper.index_image_size()
    dist.broadcast(size, src=src, group=group)
    n = int(size.item())
    image = torch.empty(n, dtype=torch.uint8, device=dev)
    if rank == src:
        mapper.export_index(image.data_ptr(), n)
    dist.broadcast(image, src=src, group=group)
    if rank == src:
        return mapper
    return Mapper.from_index(image.data_ptr(), n, device=device, ref_name=ref_name)


# ---------------------------------------------------------------------------------------------------------------------
# `downpore overlap`: one round up to the seed-match stream (overlap.Overlapper, overlap/overlap.go:24-29)
# ---------------------------------------------------------------------------------------------------------------------
OVERLAP_DEFAULTS = dict(overlap_size=1000, k=10, num_seeds=15, seed_batch_size=10000, chunk_size=10000,
                        query_batch_size=20000, min_hits=0.25)  # commands/overlap.go:26-27


class OverlapRound:
    """Result of Overlapper.round(): counters, timings, and the hits (a structured array + the MatchA/MatchB pool)."""

    def __init__(self, st, hits, matches):
        for name, _ in OverlapRoundStruct._fields_:
            if name not in ("hits", "matches"):
                setattr(self, name, getattr(st, name))
        self.hits = hits
        self.matches = matches

    def hit(self, i):
        """(query_id, rc, target, MatchA, MatchB) of hit i."""
        h = self.hits[i]
        at, n = int(h["at"]), int(h["n"])
        return int(h["query_id"]), bool(h["rc"]), int(h["target"]), self.matches[at:at + n], self.matches[at + n:at + 2 * n]


class Overlapper:
    """The sequence set of `downpore overlap` (himem) on one GPU plus overlap.Overlapper's three steps as one round:
    PrepareQueries, AddSequences, FindOverlaps (commands/overlap.go:115-160)."""

    def __init__(self, bases, offsets, kmer_values=None, device=0, **params):
        p = dict(OVERLAP_DEFAULTS)
        p.update(params)
        self.params = p
        self.k = p["k"]
        bases = _u8(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.n_reads = offsets.size - 1
        self.device = device
        vals = None
        if kmer_values is not None:
            vals = np.ascontiguousarray(kmer_values, dtype=np.float64)
            if vals.size != 4 ** self.k:
                raise ValueError("kmer_values must have 4^k entries")
        h = c_vp()
        _check(lib().dp_overlapper_create(bases.ctypes.data_as(c_vp), offsets.ctypes.data_as(c_vp), self.n_reads, self.k,
                                          None if vals is None else vals.ctypes.data_as(c_vp), p["overlap_size"],
                                          p["num_seeds"], p["seed_batch_size"], p["chunk_size"], p["query_batch_size"],
                                          float(p["min_hits"]), device, ctypes.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, lib().dp_overlapper_destroy, h)

    def close(self):
        if self._h is not None:
            self._fin()
            self._h = None

    def kmer_counts(self):
        """sequtil.KmerOccurrences over every read (commands/overlap.go:43)."""
        counts = np.zeros(4 ** self.k, dtype=np.uint64)
        _check(lib().dp_overlapper_kmer_counts(self._h, counts.ctypes.data_as(c_vp)))
        return counts

    def set_values(self, kmer_values):
        vals = np.ascontiguousarray(kmer_values, dtype=np.float64)
        if vals.size != 4 ** self.k:
            raise ValueError("kmer_values must have 4^k entries")
        _check(lib().dp_overlapper_set_values(self._h, vals.ctypes.data_as(c_vp)))

    def round(self, first_sequence=0, ignore=None):
        st = OverlapRoundStruct()
        ign = None
        if ignore is not None:
            ign = np.ascontiguousarray(ignore, dtype=np.uint8)
            if ign.size != self.n_reads:
                raise ValueError("ignore must have one flag per read")
        _check(lib().dp_overlapper_round(self._h, None if ign is None else ign.ctypes.data_as(c_vp), int(first_sequence),
                                         ctypes.byref(st)))
        hits = _adopt_or_empty(st.hits, st.num_hits * OVERLAP_HIT_DTYPE.itemsize, OVERLAP_HIT_DTYPE)
        matches = _adopt_or_empty(st.matches, st.num_matches * 2, np.uint16)
        self._last = st
        return OverlapRound(st, hits, matches)

    def seed_kmers(self):
        out = np.zeros(max(int(self._last.num_seeds), 1), dtype=np.int64)
        _check(lib().dp_overlapper_seed_kmers(self._h, out.ctypes.data_as(c_vp)))
        return out[:int(self._last.num_seeds)]

    def _segments(self, fn, n, width, *head):
        meta = np.zeros((max(n, 1), width), dtype=np.int64)
        seg_off = np.zeros(n + 1, dtype=np.int64)
        segs = c_vp()
        _check(fn(self._h, *head, meta.ctypes.data_as(c_vp), seg_off.ctypes.data_as(c_vp), ctypes.byref(segs)))
        flat = _adopt_or_empty(segs.value, int(seg_off[n]) * 8, np.int64)
        return meta[:n], seg_off, flat

    def queries(self):
        """[{id, sequence_id, rc, length, offset, inset, segments}] of the last round."""
        n = int(self._last.num_queries)
        meta, so, flat = self._segments(lib().dp_overlapper_queries, n, 6)
        return [dict(id=int(m[0]), sequence_id=int(m[1]), rc=bool(m[2]), length=int(m[3]), offset=int(m[4]), inset=int(m[5]),
                     segments=flat[so[i]:so[i + 1]]) for i, m in enumerate(meta)]

    def chunks(self, ids=None):
        """[{read, length, offset, inset, segments}] of the requested chunks (all by default) of the last round."""
        if ids is None:
            n, head = int(self._last.num_chunks), (None, 0)
        else:
            ida = np.ascontiguousarray(ids, dtype=np.int32)
            n, head = ida.size, (ida.ctypes.data_as(c_vp), ida.size)
        meta, so, flat = self._segments(lib().dp_overlapper_chunks, n, 5, *head)
        return [dict(read=int(m[0]), length=int(m[1]), offset=int(m[2]), inset=int(m[3]), segments=flat[so[i]:so[i + 1]])
                for i, m in enumerate(meta)]
