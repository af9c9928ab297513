"""ctypes wrapper over tools/synth.cpp: seeded synthetic references and ONT-like reads."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libdpsynth.so")


def build(force=False):
    src = os.path.join(_HERE, "synth.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", _LIB, src])
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.dps_reference.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
        _lib.dps_reference.restype = None
        _lib.dps_reads.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64,
                                   ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.dps_reads.restype = None
        _lib.dps_make_repeats.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int, ctypes.c_double,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_double]
        _lib.dps_make_repeats.restype = None
    return _lib


def _threads():
    return max(1, min(32, os.cpu_count() or 1))


def reference(seed, length):
    """uint8 numpy array of `length` ASCII bases (uniform i.i.d. ACGT)."""
    out = np.empty(length, dtype=np.uint8)
    _load().dps_reference(seed, length, out.ctypes.data, _threads())
    return out


def reference_rep(seed, length, families=50, frac=0.10, min_len=300, max_len=6000, max_div=0.15):
    """The `rep` variant of reference(seed, length): about `frac` of its bases replaced by copies of
    `families` random repeat units of min_len..max_len bases at 0..max_div substitution divergence, either strand."""
    out = reference(seed, length)
    _load().dps_make_repeats(out.ctypes.data, length, seed, families, frac, min_len, max_len, max_div)
    return out


def reads(ref, seed, n, read_len, circular=True, first_index=0, p_sub=0.04, p_ins=0.03, p_del=0.03, out=None,
          with_truth=False):
    """(n*read_len) uint8 array of concatenated fixed-length reads; offsets are i*read_len."""
    ref = np.ascontiguousarray(ref, dtype=np.uint8)
    if out is None:
        out = np.empty(n * read_len, dtype=np.uint8)
    truth = np.empty((n, 2), dtype=np.int64) if with_truth else None
    _load().dps_reads(ref.ctypes.data, ref.size, int(circular), seed, first_index, n, read_len, p_sub, p_ins, p_del,
                      out.ctypes.data, truth.ctypes.data if with_truth else None, _threads())
    if with_truth:
        return out, truth
    return out


def write_fasta(path, names, seqs):
    with open(path, "wb") as f:
        for name, s in zip(names, seqs):
            f.write(b">" + name.encode() + b"\n")
            f.write(bytes(s) if not isinstance(s, (bytes, bytearray)) else s)
            f.write(b"\n")
