"""GPU parity tests of the `overlap` path (BASELINE config 5): one round of commands/overlap.go:115-160 up to
the seed-match stream — PrepareQueries (seed selection AddSeeds, queries and their reverse complements), AddSequences
(seed sequences of all reads, seed-space chunks, index) and FindOverlaps (Matches, CountIntersectionTo,
PairwiseAlignments, best match, threshold escalation) — through the C ABI (dp_overlapper_*) against the oracle's literal
restatement (oracle/overlap.cpp, num_workers = 1 order). Integer / index work: the bar is bit-exact at every stage."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402

import downpore_b200 as dp  # noqa: E402


def read_set(ref, seed, lengths, **kw):
    """Concatenated simulated reads of the given lengths (one synth.reads call per distinct length)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    offs = np.zeros(lengths.size + 1, dtype=np.int64)
    np.cumsum(lengths, out=offs[1:])
    bases = np.empty(int(offs[-1]), dtype=np.uint8)
    for i, L in enumerate(lengths):
        bases[offs[i]:offs[i + 1]] = synth.reads(ref, seed, 1, int(L), first_index=i, **kw)
    return bases, offs


def compare_round(bases, offs, values, first_sequence=0, ignore=None, **params):
    o = po.OverlapRound(bases, offs, values, first_sequence=first_sequence, ignore=ignore, **params)
    g = dp.Overlapper(bases, offs, values, **params)
    r = g.round(first_sequence=first_sequence, ignore=ignore)
    assert (r.num_seeds, r.num_queries, r.num_query_seqs, r.next_first_sequence) == \
        (o.num_seeds, o.num_queries, o.num_query_seqs, o.next_first_sequence)
    assert np.array_equal(g.seed_kmers(), o.seed_kmers), "AddSeeds: registration order differs"
    if o.num_queries == 0:
        g.close()
        return o, r
    gq = g.queries()
    assert len(gq) == len(o.queries)
    for a, b in zip(gq, o.queries):
        assert (a["id"], a["sequence_id"], a["rc"], a["length"], a["offset"], a["inset"]) == \
            (b["id"], b["sequence_id"], b["rc"], b["length"], b["offset"], b["inset"])
        assert np.array_equal(a["segments"], b["segments"]), "query %d rc=%d: segments differ" % (a["id"], a["rc"])
    assert r.num_chunks == o.num_chunks
    gc = g.chunks()
    for i, (a, b) in enumerate(zip(gc, o.chunks)):
        assert (a["read"], a["length"], a["offset"], a["inset"]) == (b["read"], b["length"], b["offset"], b["inset"]), i
        assert np.array_equal(a["segments"], b["segments"]), "chunk %d: segments differ" % i
    ghits = [r.hit(i) for i in range(int(r.num_hits))]
    ohits = [(h["query_id"], h["rc"], h["target"], h["match_a"], h["match_b"]) for h in o.hits]
    assert len(ghits) == len(ohits), "hits: %d on the device, %d in the oracle" % (len(ghits), len(ohits))
    for i, (a, b) in enumerate(zip(ghits, ohits)):
        assert a[:3] == b[:3], "hit %d: %r on the device, %r in the oracle" % (i, a[:3], b[:3])
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]), "hit %d: alignment differs" % i
    # a subset of chunks fetched by id equals the full dump
    if r.num_chunks > 3:
        ids = [int(r.num_chunks) - 1, 0, int(r.num_chunks) // 2]
        for a, i in zip(g.chunks(ids), ids):
            assert np.array_equal(a["segments"], o.chunks[i]["segments"]) and a["read"] == o.chunks[i]["read"]
    g.close()
    return o, r


@pytest.fixture(scope="module")
def small():
    ref = synth.reference(1, 200_000)
    n, L = 400, 8000
    rd = synth.reads(ref, 15, n, L, circular=True)
    offs = np.arange(n + 1, dtype=np.int64) * L
    vals = po.overlap_values(rd, offs, 10)
    return rd, offs, vals


def test_round_defaults(small):
    rd, offs, vals = small
    o, r = compare_round(rd, offs, vals)
    assert o.num_hits > 500 and r.pairs >= r.num_hits


def test_kmer_counts_and_values(small):
    rd, offs, vals = small
    g = dp.Overlapper(rd, offs, None)
    counts = g.kmer_counts()
    gv = dp.kmer_values(counts, 10)
    assert np.array_equal(gv, vals)
    g.set_values(gv)
    r = g.round()
    assert r.num_queries > 0
    g.close()


def test_later_round_with_ignored_reads(small):
    rd, offs, vals = small
    n = offs.size - 1
    rng = np.random.default_rng(5)
    ignore = (rng.random(n) < 0.2).astype(np.uint8)
    o, r = compare_round(rd, offs, vals, first_sequence=167, ignore=ignore)
    assert o.num_hits > 100
    # the last round: nothing left to query
    compare_round(rd, offs, vals, first_sequence=n)
    allign = np.ones(n, dtype=np.uint8)
    compare_round(rd, offs, vals, first_sequence=0, ignore=allign)


@pytest.mark.parametrize("params", [
    dict(seed_batch_size=2000),
    dict(seed_batch_size=40000, query_batch_size=150),
    dict(num_seeds=30, min_hits=0.5),
    dict(num_seeds=8, min_hits=0.1),
    dict(overlap_size=600, chunk_size=3000),
    dict(overlap_size=2000, num_seeds=40),
])
def test_round_parameters(small, params):
    rd, offs, vals = small
    compare_round(rd, offs, vals, **params)


@pytest.mark.parametrize("k", [8, 9, 11, 12])
def test_round_other_k(k):
    ref = synth.reference(2, 120_000)
    n, L = 200, 6000
    rd = synth.reads(ref, 25, n, L, circular=True)
    offs = np.arange(n + 1, dtype=np.int64) * L
    vals = po.overlap_values(rd, offs, k)
    compare_round(rd, offs, vals, k=k)


def test_mixed_read_lengths():
    """Reads below 2 x overlap_size (one slice), below chunk_size (one chunk), and up to 45 kb (many seed-space chunks,
    the 150-seeds-before-the-end rule), lengths of every residue mod 4."""
    ref = synth.reference(3, 300_000)
    rng = np.random.default_rng(11)
    lengths = np.concatenate([rng.integers(1000, 2000, 40), rng.integers(2000, 9999, 80), rng.integers(10000, 45000, 60),
                              [1000, 1001, 1002, 1003, 1999, 2000, 2001, 9999, 10000, 10001]])
    rng.shuffle(lengths)
    bases, offs = read_set(ref, 31, lengths)
    vals = po.overlap_values(bases, offs, 10)
    o, r = compare_round(bases, offs, vals)
    assert o.num_chunks > len(lengths)
    compare_round(bases, offs, vals, seed_batch_size=60000)


def test_low_error_reads_long_chains():
    """Nearly error-free reads: long chains, frequent threshold escalation inside a query's candidate list, the clamped
    soft-union levels and the level-16 under-count (many included seeds per query)."""
    ref = synth.reference(4, 60_000)
    n, L = 300, 7000
    rd = synth.reads(ref, 41, n, L, circular=True, p_sub=0.005, p_ins=0.002, p_del=0.002)
    offs = np.arange(n + 1, dtype=np.int64) * L
    vals = po.overlap_values(rd, offs, 10)
    o, r = compare_round(rd, offs, vals)
    assert o.num_hits > 2000
    compare_round(rd, offs, vals, num_seeds=60, seed_batch_size=30000, min_hits=0.3)


def test_repeat_rich_reads():
    """Reads from a reference full of repeat families and tandem copies: seeds that recur inside a query and a chunk,
    many chain starts per seed, same-as-neighbour repeats in prepareInitial."""
    ref = synth.reference_rep(5, 150_000, families=20, frac=0.4, min_len=200, max_len=2000, max_div=0.05)
    unit = synth.reference(6, 37)
    ref[50_000:50_000 + 37 * 200] = np.tile(unit, 200)
    n, L = 300, 9000
    rd = synth.reads(ref, 51, n, L, circular=True, p_sub=0.02, p_ins=0.01, p_del=0.01)
    offs = np.arange(n + 1, dtype=np.int64) * L
    vals = po.overlap_values(rd, offs, 10)
    compare_round(rd, offs, vals)
    compare_round(rd, offs, vals, num_seeds=25, min_hits=0.15)


def test_lowercase_and_n_bases(small):
    rd, offs, vals = small
    rd = rd[:offs[120]].copy()
    offs = offs[:121]
    rng = np.random.default_rng(3)
    idx = rng.integers(0, rd.size, rd.size // 50)
    rd[idx] = ord("N")
    idx = rng.integers(0, rd.size, rd.size // 10)
    rd[idx] = np.char.lower(rd[idx].view("S1")).view(np.uint8)
    v = po.overlap_values(rd, offs, 10)
    compare_round(rd, offs, v)
