"""The oracle's memory-lean index mode (oracle.hpp) against its plain mode: same seeds, same chunks, same candidates,
same mapping records and the same work counters, on references where both fit. The lean mode is what lets the oracle
check the CUDA path on the 3.1 Gb reference of BASELINE config 4 (tests/test_gpu_scale_parity.py)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tools import synth


def kmer_values(ref, k):
    import downpore_b200 as dp  # host-side numpy formula only (commands/map.go:46-71); no device call
    return dp.kmer_values(po.kmer_counts(ref, k), k)


@pytest.mark.parametrize("circular,n_ref,k", [(True, 400_000, 11), (False, 1_300_000, 11), (False, 700_000, 13)])
def test_lean_equals_plain(circular, n_ref, k):
    ref = synth.reference(31, n_ref)
    vals = kmer_values(ref, k)
    a = po.Mapper(ref, vals, circular=circular, k=k, lean=False, threads=1)
    b = po.Mapper(ref, vals, circular=circular, k=k, lean=True, threads=3)
    assert not a.lean and b.lean
    assert a.num_seeds == b.num_seeds and a.num_chunks == b.num_chunks
    assert np.array_equal(a.seed_kmers(), b.seed_kmers())
    for c in range(a.num_chunks):
        x, y = a.chunk(c), b.chunk(c)
        assert (x["offset"], x["inset"], x["length"], x["nseeds"]) == (y["offset"], y["inset"], y["length"], y["nseeds"])
        assert np.array_equal(x["segments"], y["segments"])
    n, L = 300, 7000
    rd = synth.reads(ref, 5, n, L, circular=circular)
    offs = np.arange(n + 1, dtype=np.int64) * L
    ra = a.map_batch(rd, offs, threads=2)
    rb = b.map_batch(rd, offs, threads=2)
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]) and ra[2] == rb[2]
    r = rd[:L]
    for strand in (False, True):
        assert np.array_equal(a.window_candidates(r, 0, 1000, False, strand), b.window_candidates(r, 0, 1000, False, strand))


def test_lean_on_a_repeat_reference():
    """Tandem copies of a 13 kb unit with a dozen substitutions each: 40 candidates per window, repeated seeds inside a
    query, seeds present in most chunks."""
    rng = np.random.default_rng(5)
    unit = synth.reference(77, 13_000)
    copies = []
    for _ in range(40):
        u = unit.copy()
        u[rng.integers(0, len(u), size=12)] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=12)]
        copies.append(u)
    ref = np.concatenate(copies)
    vals = kmer_values(ref, 11)
    a = po.Mapper(ref, vals, circular=False, lean=False)
    b = po.Mapper(ref, vals, circular=False, lean=True)
    n, L = 40, 4000
    rd = synth.reads(ref, 6, n, L, circular=False)
    offs = np.arange(n + 1, dtype=np.int64) * L
    ra = a.map_batch(rd, offs, threads=2)
    rb = b.map_batch(rd, offs, threads=2)
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]) and ra[2] == rb[2]
    assert ra[2]["mappings"] > 20 * n
