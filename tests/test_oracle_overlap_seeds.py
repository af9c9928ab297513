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


# ---- SeedSequence.ReverseComplement (seeds/sequence.go:134-159) -------------------------------------------------

def revcomp_text(text):
    comp = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A")}
    return bytes(comp[b] for b in reversed(text))


@pytest.mark.parametrize("k", [6, 9])
def test_reverse_complement_of_a_seed_sequence_is_the_seed_sequence_of_the_reverse_complement(k):
    """With a seed set closed under reverse complement (AddSeeds registers both strands) the seed sequence of the
    reverse-complemented TEXT is the ReverseComplement of the forward seed sequence; and rc(rc(s)) = s. (Lengths with
    len % 4 != 0: raw packed sequences whose length is a multiple of four lose k-mers at the end, Q2.)"""
    rng = np.random.default_rng(7 + k)
    ranks = rng.random(4 ** k)
    g = po.SeedIndex(k)
    texts = []
    for _ in range(12):
        L = int(rng.integers(300, 2000))
        L += 1 if L % 4 == 0 else 0
        t = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L))
        texts.append(t)
        g.add_seeds(t, 12, ranks)
    for t in texts:
        fwd, length, _, _ = g.seed_sequence(t)
        assert fwd[0::2].sum() + k * (len(fwd) // 2) == length == len(t)      # gaps + seeds tile the read
        rc = g.reverse_complement(fwd)
        assert np.array_equal(g.reverse_complement(rc), fwd)
        rc_text, _, _, _ = g.seed_sequence(revcomp_text(t))
        assert np.array_equal(rc, rc_text)
        assert len(fwd) // 2 >= 12


# ---- chunkWorker (overlap/overlap.go:253-318) ------------------------------------------------------------------

def chunk_spec(seg, length, chunk_size, min_seeds, overlap, k):
    """Independent transcription: returns [(first seed, last seed, length, offset, inset)]."""
    ns = len(seg) // 2
    nso = lambda i: seg[i * 2 + 2] + k                         # GetNextSeedOffset
    def from_end(i):                                           # GetSeedOffsetFromEnd
        off = seg[-1]
        j = len(seg) - 3
        while j > 2 * i + 1:
            off += seg[j] + k
            j -= 2
        return off
    out = []
    if length // chunk_size + 1 == 1 or ns < 3 * min_seeds:
        return [(0, ns - 1, length, 0, 0)] if ns >= min_seeds else []
    prev, total, lib_ = 0, seg[0], 0
    while True:
        if prev >= ns - 150:
            if prev == 0:
                out.append((0, ns - 1, length, 0, 0))
            else:
                gap = nso(prev - 1) - k
                lib_ += from_end(prev) + k + gap
                out.append((prev, ns - 1, lib_, total - gap, 0))
            break
        cnt = 0
        while lib_ < chunk_size and cnt < 100 and prev + cnt < ns:
            lib_ += nso(prev + cnt)
            cnt += 1
        if cnt >= min_seeds:
            gap = nso(prev - 1) - k
            lib_ += gap
            out.append((prev, prev + cnt - 1, lib_, total - gap, length - total - lib_ + gap))
            total += lib_ - gap
            lib_ = 0
            prev += cnt
            if prev >= ns:
                break
            back = 0
            while back < 5 and lib_ < overlap // 2 and prev > 0:
                prev -= 1
                lib_ += nso(prev)
                total -= nso(prev)
                back += 1
            lib_ = 0
        else:
            prev += cnt
            while lib_ < overlap // 2 and prev > 0:
                prev -= 1
                lib_ += nso(prev)
                total -= nso(prev)
            lib_ = 0
    return out


@pytest.mark.parametrize("seed", range(6))
def test_seed_space_chunking(seed):
    rng = np.random.default_rng(300 + seed)
    k = 10
    for _ in range(40):
        ns = int(rng.integers(1, 900))
        density = float(rng.choice([15.0, 40.0, 120.0, 400.0]))          # mean gap between seeds
        gaps = rng.geometric(1.0 / density, size=ns + 1).astype(np.int64) - 1
        seg = np.empty(2 * ns + 1, dtype=np.int64)
        seg[0::2] = gaps
        seg[1::2] = rng.integers(0, 5000, size=ns)
        length = int(gaps.sum() + k * ns)
        chunk_size = int(rng.choice([2000, 10000]))
        min_seeds = int(rng.choice([5, 15]))
        overlap = int(rng.choice([400, 1000]))
        got = po.chunk_seed_sequence(seg, length, chunk_size, min_seeds, overlap, k)
        want = chunk_spec([int(x) for x in seg], length, chunk_size, min_seeds, overlap, k)
        assert len(got) == len(want), (seed, ns, density, chunk_size, min_seeds, overlap)
        whole = len(want) == 1 and want[0][:2] == (0, ns - 1) and want[0][3:] == (0, 0)
        for (gseg, glen, goff, gins), (a, b, ln, off, ins) in zip(got, want):
            assert np.array_equal(gseg, seg[2 * a: 2 * b + 3])
            if whole:
                assert glen == length
                continue
            # (no tiling invariant is asserted: after a dropped piece the reference's running offset steps back over
            # bases it has already counted, and its offsets stop tiling the read — both transcriptions reproduce that)
            assert (glen, goff, gins) == (ln, off, ins)
