"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols
include/downpore_b200.h declares. No compute call is made here (there is no GPU on the CPU tier and no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import downpore_b200 as dp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "downpore_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dp_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    path = dp.build()
    assert os.path.exists(path)
    names = header_symbols()
    assert names == dp.exported_symbols()
    L = ctypes.CDLL(path)
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (dp_[a-z_0-9]+)", out)))
    assert exported == names


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", dp.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    dp.build()
    with pytest.raises(dp.DownporeError):
        dp.pack("ACGT")
    vals = np.zeros(4 ** 11)
    with pytest.raises(dp.DownporeError):
        dp.Mapper(np.frombuffer(b"ACGT" * 1000, dtype=np.uint8), vals)


def test_mapping_record_layout():
    assert dp.MAPPING_DTYPE.itemsize == 32
    assert dp.MAPPING_DTYPE.fields["q_offset"][1] == 16 and dp.MAPPING_DTYPE.fields["rc"][1] == 28


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under downpore_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "downpore_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text, f


def test_kmer_values_matches_oracle():
    from oracle import pyoracle as po
    from tools import synth
    ref = synth.reference(3, 400_000)
    for k in (9, 11):
        counts = po.kmer_counts(ref, k)
        assert np.array_equal(dp.kmer_values(counts, k), po.kmer_values(ref, k))
