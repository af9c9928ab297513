"""GPU parity at the sizes BASELINE.json names (run with -m gpu on a B200): the CUDA path, through the C ABI, against
the CPU oracle on the FULL references of configs 3 and 4, on the repeat-rich `rep` variant and on a
tandem-repeat reference that overflows every default device capacity. Bit-exact records and work counters.

The oracle indexes the 3.1 Gb reference in its memory-lean mode (oracle.hpp: the bitsets of the reference's SeedIndex,
#seeds x #chunks bits twice, are rebuilt per query from lists; every routine downstream runs unchanged);
tests/test_oracle_lean.py proves that mode equal to the plain one where both fit."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402

import downpore_b200 as dp  # noqa: E402

CORES = os.cpu_count() or 4
COUNTERS = ("windows", "kmer_lookups", "query_seeds", "posting_runs", "posting_entries", "candidates", "chain_cells",
            "mappings")


def rows_of(maps):
    if len(maps) == 0:
        return np.zeros((0, 6), dtype=np.int64)
    return np.stack([maps["start"], maps["end"], maps["q_offset"], maps["q_inset"], maps["rc"], maps["ids"]],
                    axis=1).astype(np.int64)


class _Env:
    def __init__(self, env):
        self.env = env

    def __enter__(self):
        self.old = {k2: os.environ.get(k2) for k2 in self.env}
        os.environ.update(self.env)

    def __exit__(self, *a):
        for k2, v in self.old.items():
            if v is None:
                os.environ.pop(k2, None)
            else:
                os.environ[k2] = v


def check_against_oracle(ref, k, n, L, read_seed, circular=False, lean=None, expect_retries=None, what=""):
    vals = dp.kmer_values(dp.kmer_counts(ref, k), k)
    gm = dp.Mapper(ref, vals, circular=circular, k=k)
    om = po.Mapper(ref, vals, circular=circular, k=k, lean=lean, threads=CORES)
    if lean is not None:
        assert om.lean == lean
    info = gm.index_info()
    assert info["num_seeds"] == om.num_seeds and info["num_chunks"] == om.num_chunks, what
    assert np.array_equal(np.sort(om.seed_kmers()), gm.seed_kmers()), what
    for c in (0, om.num_chunks // 2, om.num_chunks - 1):  # (every chunk is compared on the smaller references)
        oc, gc = om.chunk(c), gm.chunk(c)
        seg = oc["segments"]
        pos = np.cumsum(seg[0::2][:-1]) + k * np.arange(len(seg) // 2)
        assert np.array_equal(pos, gc["pos"]) and np.array_equal(seg[1::2], gc["kmer"]), (what, c)
    rd = synth.reads(ref, read_seed, n, L, circular=circular)
    offs = np.arange(n + 1, dtype=np.int64) * L
    orow, ooff, octr = om.map_batch(rd, offs, threads=CORES)
    gmaps, goff = gm.map_batch(rd, offs)
    st = gm.stats()
    gm.close()
    assert np.array_equal(ooff, goff), what
    assert np.array_equal(orow, rows_of(gmaps)), what
    for key in COUNTERS:
        assert st[key] == octr[key], (what, key)
    if expect_retries is not None:
        assert (st["retries"] > 0) == expect_retries, (what, st["retries"])
    return st, octr


def test_config3_full_reference():
    """BASELINE config 3: the full 64 Mb linear reference (6 465 chunks), 20 000 x 20 kb reads."""
    ref = synth.reference(3, 64_000_000)
    st, octr = check_against_oracle(ref, 11, 20_000, 20_000, 13, what="config3")
    assert octr["mappings"] >= 19_900


def test_config3_rep_variant():
    """The `rep` variant of config 3's reference: 10 % of the bases are copies of 50 repeat families of
    300-6000 b at 0-15 % divergence (~40 copies each): windows inside a repeat see dozens of candidates and chains."""
    ref = synth.reference_rep(3, 64_000_000)
    st, octr = check_against_oracle(ref, 11, 20_000, 20_000, 13, what="config3-rep")
    assert octr["candidates"] > 2 * octr["windows"]  # the repeats are felt


@pytest.mark.parametrize("k", [11, 13])
def test_config4_reference(k):
    """BASELINE config 4's reference: 3.1 Gb linear, 313 131 chunks, k = 11 (default) and k = 13; 2 000 x 15 kb reads.
    The oracle runs in its memory-lean mode (its plain bitsets would need 2 x 34 GB at k = 11)."""
    ref = synth.reference(4, 3_100_000_000)
    st, octr = check_against_oracle(ref, k, 2_000, 15_000, 14, lean=True, what="config4 k=%d" % k)
    assert octr["mappings"] >= 1_990


def tandem_reference(copies=24, unit_len=12_000):
    rng = np.random.default_rng(5)
    unit = synth.reference(77, unit_len)
    out = []
    for _ in range(copies):
        u = unit.copy()
        pos = rng.integers(0, len(u), size=12)  # a dozen substitutions per copy
        u[pos] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=12)]
        out.append(u)
    return np.concatenate(out)


def test_tandem_repeats_return_the_oracles_answer():
    """A reference made of 24 copies of one 12 kb unit: every read window hits every copy — more window results than
    the default pool of a launch holds. The reference keeps every hit (mapping/mapping.go:504-552); so does the device
    path: the launch is flagged, the capacity grown, the sub-batch recomputed (dp_stats.retries)."""
    ref = tandem_reference()
    n, L = 40, 6000
    st, octr = check_against_oracle(ref, 11, n, L, 5, expect_retries=True, what="tandem")
    assert octr["mappings"] > 16 * n


@pytest.mark.parametrize("env", [
    {"DP_CAP_OUT": "1"},                                            # pool of window results
    {"DP_CAP_RESULTS": "1", "DP_CHAIN_FAST": "0"},                  # mappings per window (general chain kernel)
    {"DP_CAP_CHAINS": "1", "DP_CHAIN_FAST": "0"},                   # chains per candidate
    {"DP_CAP_CANDS": "1"},                                          # candidates per window strand
    {"DP_CAP_CANDS": "1", "DP_CAND_BUDGET_MB": "0"},                # ... with the range cut down to single reads
    {"DP_CAP_OUT": "1", "DP_CAP_RESULTS": "1", "DP_CAP_CHAINS": "1", "DP_CAP_CANDS": "2"},
])
def test_every_capacity_has_an_exact_way_out(env):
    """Tiny starting capacities on a 64-copy tandem reference (65 candidates and 60 chains per window) and on an ordinary one: every overflow path (flag ->
    fourfold growth -> recompute; candidate lists over the memory budget -> the range is cut first) gives the oracle's
    records and counters."""
    for ref, n, L, seed, retries in ((tandem_reference(64, 13_000), 60, 4_000, 6, True),
                                     (synth.reference(9, 300_000), 300, 5_000, 7, None)):
        with _Env(env):
            check_against_oracle(ref, 11, n, L, seed, expect_retries=retries, what=str(env))
