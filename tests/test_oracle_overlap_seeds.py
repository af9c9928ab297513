"""Round-2 groundwork (the overlap path): the oracle's literal restatement of SeedIndex.AddSeeds
(seeds/seeds.go:62-156) against an independent restatement over plain strings (no packed sequences, no asm emulation:
k-mers are computed from the text). The reference has no test of its own here."""
import numpy as np
import pytest

from oracle import pyoracle as po

CODE = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}


def kmer_at(text, i, k):
    v = 0
    for b in text[i:i + k]:
        v = (v << 2) | CODE[b]
    return v


def revcomp_kmer(x, k):
    r = 0
    for _ in range(k):
        r = (r << 2) | ((x ^ 3) & 3)
        x >>= 2
    return r


def add_seeds_spec(is_seed, order, text, k, min_seeds, ranks, quality=None):
    """One AddSeeds call. is_seed: set of seed k-mers (updated); order: registration order (appended to)."""
    L = len(text)
    top = [0] * min_seeds          # ascending by value; slot 0 is the worst kept
    top_v = [0.0] * min_seeds
    nxt = k                        # the k-mer ENDING at index nxt is the next candidate: starts at nxt - k + 1
    while nxt < L - k:
        reset, best_v, best = False, 0.0, 0
        i = 0
        while nxt < L and i < k:
            km = kmer_at(text, nxt - k + 1, k)
            nxt += 1
            if km in is_seed:
                reset = True
                break
            v = float(ranks[km])
            if quality is not None:
                v *= float(quality[nxt - k // 2])
            if v > best_v:
                best_v, best = v, km
            i += 1
        if not reset:
            n = 0
            while n < min_seeds and top_v[n] < best_v:
                if n > 0:
                    top_v[n - 1], top[n - 1] = top_v[n], top[n]
                n += 1
            if n > 0:
                top_v[n - 1], top[n - 1] = best_v, best
        nxt += 2 * k               # skip k, re-anchor, skip k
    for km in top:
        for x in (km, revcomp_kmer(km, k)):
            if x not in is_seed:
                is_seed.add(x)
                order.append(x)


@pytest.mark.parametrize("k,with_quality", [(6, False), (8, False), (10, True), (7, True)])
def test_add_seeds_matches_the_text_restatement(k, with_quality):
    rng = np.random.default_rng(1000 + k)
    ranks = rng.random(4 ** k)
    ranks[rng.random(4 ** k) < 0.3] = 0.0          # count < 3 -> value 0 in getKmerValues
    ranks[0] = 0.0
    g = po.SeedIndex(k)
    is_seed, order = set(), []
    for n_seq in range(30):
        L = int(rng.integers(3 * k + 2, 1500))
        text = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L))
        q = rng.integers(1, 60, size=L).astype(np.uint8) if with_quality else None
        min_seeds = int(rng.integers(1, 16))
        g.add_seeds(text, min_seeds, ranks, q)
        add_seeds_spec(is_seed, order, text, k, min_seeds, ranks, q)
        assert list(g.seeds()) == order, (k, n_seq, L, min_seeds)
    assert len(order) > 20


def test_unfilled_slots_register_kmer_zero():
    """A sequence with fewer candidate blocks than num_seeds leaves zeros in topN: k-mer 0 and its reverse complement
    (TTT...) become seeds (seeds.go:84-88,131-154) — a quirk the GPU path will have to keep."""
    k = 6
    ranks = np.full(4 ** k, 0.5)
    g = po.SeedIndex(k)
    g.add_seeds(b"ACGTTGCAAGGCTTAACCGGATATCGCGAT", 15, ranks)
    seeds = list(g.seeds())
    assert 0 in seeds and (4 ** k - 1) in seeds
