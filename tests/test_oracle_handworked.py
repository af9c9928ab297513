"""Hand-worked vectors for the layers of the oracle the reference has no tests for (seeds/sequence.go: Reduced,
GetSeedOffset, dynamicMatch / extendChain, Match, GetBasesCovered). Every expected value below was derived on paper from
the Go source, step by step, BEFORE the oracle was run on the case; the derivations are kept as comments so that a
reader can check them against /root/reference/seeds/sequence.go without running anything. They are the human-checked
pin under the otherwise 'parity unpinned' chaining layer.

A seed sequence is its `segments` list: gap, seed, gap, seed, ..., gap (seeds/sequence.go:17-30); a negative gap means
the next seed overlaps the previous one. k = 5 throughout."""
from oracle import pyoracle as po

K = 5


def test_reduced_by_hand():
    """Reduced (seeds/sequence.go:85-123), whitelist {7, 9}, on seeds 7, 9, 9, 4, 7 at positions 3, 18, 25, 35, 48.

    First pass (:88-95) counts whitelisted seeds that differ from the previous KEPT seed: 7 (kept), 9 (kept), 9 (same as
    prev: skipped), 4 (not whitelisted; prev stays 9), 7 (differs from 9: kept) -> count 3.
    Second pass (:104-121): offset starts at segments[0] = 3.
      seed 7 kept: segs[0..1] = 3, 7; index[0] = 0; offset = gap after it = 10
      seed 9 kept: segs[2..3] = 10, 9; index[1] = 1; offset = 2
      seed 9 dropped: offset += next gap + k = 2 + 5 + 5 = 12
      seed 4 dropped: offset += 8 + 5 -> 25
      seed 7 kept: segs[4..5] = 25, 7; index[2] = 4; offset = trailing gap 1
      segs[6] = 1
    Check by positions: 3, 3+5+10 = 18, 18+5+25 = 48 — the positions of the kept seeds in the original."""
    s = [3, 7, 10, 9, 2, 9, 5, 4, 8, 7, 1]
    assert po.reduced(s, [7, 9], K, 1) == ([3, 7, 10, 9, 25, 7, 1], [0, 1, 4])
    # fewer than minSeeds whitelisted seeds -> nil (:96-98)
    assert po.reduced(s, [7, 9], K, 4) is None
    # nothing whitelisted differs from everything -> count 0 < 1 -> nil
    assert po.reduced(s, [5], K, 1) is None


def test_seed_offsets_by_hand():
    """GetSeedOffset(2) (:1239-1246): segments[0] + (segments[2] + k) + (segments[4] + k) = 3 + 15 + 7 = 25.
    GetSeedOffsetFromEnd(2) (:1269-1276): segments[10] + (segments[8] + k) + (segments[6] + k) = 1 + 13 + 10 = 24.
    Check: the sequence spans 48 + 5 + 1 = 54 bases, seed 2 ends at 25 + 5 = 30, 54 - 30 = 24."""
    s = [3, 7, 10, 9, 2, 9, 5, 4, 8, 7, 1]
    assert po.seed_offset(s, 2, K) == 25
    assert po.seed_offset(s, 2, K, from_end=True) == 24
    assert po.seed_offset(s, 0, K) == 3 and po.seed_offset(s, 4, K, from_end=True) == 1


# query a: seeds 11 12 13 14 15 16 17 18 (seed indices 0..7)
A = [2, 11, 10, 12, 10, 13, -2, 14, 10, 15, 10, 16, 10, 17, 10, 18, 3]
# chunk b: seeds 11 13 14 15 16 17 18 (seed indices 0..6); 12 is missing, a long gap (40) sits between 14 and 15
B = [7, 11, 25, 13, -2, 14, 40, 15, 10, 16, 10, 17, 12, 18, 4]


def test_dynamic_match_by_hand():
    """b.dynamicMatch(a, minMatch = 2, k = 5) (seeds/sequence.go:401-471, extendChain :476-576).

    Outer loop: qIndex in 1, 3, .., 13 (qIndex < len(a) - 2*minMatch + 2 = 15); inner i in 1, .., 11 (i < 13).

    qIndex = 1 (a0 = 11): b0 = 11 matches at i = 1, no chain yet -> chainsA[0] = [0], chainsB[0] = [0]; extendChain(1, 1):
      offsetA = a[2] = 10, offsetB = b[2] = 25, aIndex = 3, bIndex = 3
      a1 (12): window for offsetA 10 is [10*2/3 - 5, 10*3/2 + 5] = [1, 20]; 20 < 25, so a moves on (:497-506):
          offsetA = 10 + a[4] + 5 = 25, aIndex = 5 (a2 = 13), window [25*2/3 - 5, 25*3/2 + 5] = [11, 42]
          scan (:520): offsetB 25 <= 42 and b[3] = 13 == 13 -> HIT: chain A [0, 2], B [0, 1]
          offsetA = a[6] = -2, offsetB = b[4] = -2, aIndex = 7, bIndex = 5
      a3 (14): offsetA < 0 -> window [-5, 0] (:489-491); b[5] = 14 at offset -2 -> HIT: A [0, 2, 3], B [0, 1, 2]
          offsetA = a[8] = 10, offsetB = b[6] = 40, aIndex = 9, bIndex = 7
      a4 (15): window [1, 20], 20 < 40 -> a moves on: offsetA = 10 + a[10] + 5 = 25, aIndex = 11 (a5 = 16),
          window [11, 42]; scan from (bIndex 7, offsetB 40): b[7] = 15 != 16 -> offsetB = 40 + b[8] + 5 = 55 > 42: no match
          (:566-572) offsetA = 25 + a[12] + 5 = 40, aIndex = 13, b back to (7, 40)
      a6 (17): window [40*2/3 - 5, 40*3/2 + 5] = [21, 65]; scan: 15 at 40, 16 at 55, then 55 + b[10] + 5 = 70 > 65: no
          match; offsetA = 40 + a[14] + 5 = 55, aIndex = 15, b back to (7, 40)
      a7 (18): window [55*2/3 - 5, 55*3/2 + 5] = [31, 87]; scan: 15 at 40, 16 at 55, 17 at 70, then 70 + b[12] + 5 = 87:
          87 <= 87 and b[13] = 18 -> HIT: A [0, 2, 3, 7], B [0, 1, 2, 6]; aIndex = 17 = len(a): done
      chain of 4 >= minMatch; nextLength = 4*2/3 = 2, not above minMatch (:438-449); kept. chainsA is set for a0, a2, a3,
      a7: remaining = 4, not < 4 (:457): go on.
    qIndex = 3 (12): not in b. qIndex = 5, 7 (13, 14): chainsA set -> skipped (:415-417).
    qIndex = 9 (a4 = 15): b3 = 15 at i = 7 -> chain [4] / [3]; extendChain(9, 7): offsets 10 / 10
      a5 (16) = b4 at offset 10 in [1, 20] -> HIT [4, 5] / [3, 4]; a6 (17) = b5 likewise -> [4, 5, 6] / [3, 4, 5];
      offsetA = 10, offsetB = b[12] = 12
      a7 (18) = b6 at 12 in [1, 20]: a chain to a7 exists already, it ends on the same b seed (6) and is longer (4 > 3)
      -> 'they have a better chain already' (:538-540): returns [4, 5, 6] / [3, 4, 5]
      3 >= 2, nextLength 2: kept. remaining = 1 (only a1) < 3 -> return both chains (:457-462).

    GetBasesCovered (:830-858), first chain: 4 seeds * 5 = 20 on both sides; between a2 and a3 the gap is -2 (overlap) in
    both sequences: 20 - 2 = 18, 18; every other distance is positive. Second chain: 15, 15."""
    got = po.match(B, A, 2, K, reduced=False)
    assert got == [([0, 2, 3, 7], [0, 1, 2, 6], 18, 18), ([4, 5, 6], [3, 4, 5], 15, 15)]


def test_match_by_hand():
    """b.Match(a, seeds of a, seeds of b, minMatch = 2, k = 5) as performMapping calls it (mapping/mapping.go:518-521).

    Reduced first (:366-372): every seed of b is a seed of a -> b unchanged. a loses seed 12 (not in b): the gap in front
    of 13 becomes 10 + (10 + 5) = 25: a' = [2, 11, 25, 13, -2, 14, 10, 15, 10, 16, 10, 17, 10, 18, 3], index [0, 2, 3, 4, 5,
    6, 7]. dynamicMatch(b, a'): a'0 = 11 = b0; offsets 25 / 25, window [11, 42]: a'1 (13) = b1 HIT; a'2 (14) = b2 at -2
    HIT; then exactly as above (15 is out of reach, 16 and 17 are not found, 18 is found at offset 87): chain [0, 1, 2, 6] /
    [0, 1, 2, 6]. a' has 7 seeds, 4 are chained: remaining = 3 < 4 -> return at once (:457-462): the second chain of the
    un-reduced run is never started. Mapped back through the index (:379-388): MatchA = [0, 2, 3, 7]."""
    got = po.match(B, A, 2, K, reduced=True)
    assert got == [([0, 2, 3, 7], [0, 1, 2, 6], 18, 18)]
    # minMatch 8: a' has only 7 whitelisted seeds -> Reduced returns nil -> Match returns nil (:373-375)
    assert po.match(B, A, 8, K, reduced=True) is None


def test_mapping_coordinates_by_hand():
    """The coordinates performMapping derives from the chain above (mapping/mapping.go:522-536) for a chunk b with
    offset 1000 in a reference of 5000 bases, inset = 5000 - 1000 - len(b):
      len(b) = 7 + 7*5 + (25 - 2 + 40 + 10 + 10 + 12) + 4 = 141 -> inset 3859
      start = offset + GetSeedOffset(MatchB[0] = 0) = 1000 + 7
      end   = refLen - inset - GetSeedOffsetFromEnd(MatchB[-1] = 6) = 5000 - 3859 - 4 = 1137
      qOffset = GetSeedOffset(a, 0) = 2; qInset = GetSeedOffsetFromEnd(a, 7) = 3
    (the chain spans b from its first seed at 7 to the end of its last seed at 141 - 4 = 137: 1000 + 137 = 1137)."""
    assert po.seed_offset(B, 0, K) == 7 and po.seed_offset(B, 6, K, from_end=True) == 4
    assert 1000 + po.seed_offset(B, 0, K) == 1007 and 5000 - 3859 - po.seed_offset(B, 6, K, from_end=True) == 1137
    assert po.seed_offset(A, 0, K) == 2 and po.seed_offset(A, 7, K, from_end=True) == 3


def test_map_ends_pairing_by_hand():
    """Map()'s first decision on explicit window hits (mapping/mapping.go:164-203: removeDominated :387-428, matchPairs
    :174-203, isConsistent :131-160). Query of 10000 bases, linear reference; rows are {Start, End, QueryOffset, QueryInset,
    RC, ids}.

    openA = A1 {5000, 5900, 50, 9050, fwd, 300}, A2 {70000, 70400, 100, 9500, fwd, 100}.
    removeDominated sorts by QueryOffset: A1, A2. A1 is dominated by nothing (300*4 > 300*5 and 100*4 > 300*5 are both
    false). A2 against A1: 300*4 = 1200 > 100*5 = 500; overlap in the query: start = max(100, 50) = 100,
    end = 10000 - 9500 = 500 (A1's inset 9050 is not larger than A2's 9500); (500 - 100)*10 = 4000 >
    (10000 - 100 - 9500)*9 = 3600 -> dominated, removed. openA = [A1].
    openB = B1 {13950, 14900, 9040, 30, fwd, 280}, B2 {40000, 40900, 9100, 20, rc, 250}: 250*4 = 1000 > 280*5 and
    280*4 = 1120 > 250*5 = 1250 are both false -> both stay (sorted: B1, B2).
    matchPairs walks openB from the back: A1/B2 differ in strand. A1/B1: expectedDistance = 9040 - 10000 + 9050 = 8090,
    distance = 13950 - 5900 = 8050 > 5000 -> needs 8090 < 8050*10/9 = 8944 and 8090 > 8050*9/10 = 7245: consistent.
    combined = {A1.Start, B1.End, A1.QueryOffset, B1.QueryInset, fwd, 300 + 280}; A1 and B1 leave their lists."""
    a1, a2 = [5000, 5900, 50, 9050, 0, 300], [70000, 70400, 100, 9500, 0, 100]
    b1, b2 = [13950, 14900, 9040, 30, 0, 280], [40000, 40900, 9100, 20, 1, 250]
    ra, rb, matched = po.pair_ends(100000, False, 10000, [a2, a1], [b2, b1])
    assert ra == [] and rb == [b2] and matched == [[5000, 14900, 50, 30, 0, 580]]


def test_is_consistent_middle_band_by_hand():
    """isConsistent between 500 and 5000 bases apart (mapping/mapping.go:155-159): distance = 4750 - 2000 = 2750,
    ratio = (2750 - 500)/4500 = 0.5 -> 3/2 + 0.5*(10/9 - 3/2) = 1.30555..
    expectedDistance 2500 (= 3500 - 10000 + 9000): 2500*1.30555 = 3263.9 -> 3263 > 2750 and 2500/1.30555 = 1914.9 -> 1914
    < 2750: consistent, the pair is merged. expectedDistance 2100: 2100*1.30555 = 2741.7 -> 2741, 2750 < 2741 fails: the
    hits stay apart and matched is nil."""
    a = [1000, 2000, 0, 9000, 0, 100]
    ra, rb, matched = po.pair_ends(100000, False, 10000, [a], [[4750, 5700, 3500, 5600, 0, 120]])
    assert ra == [] and rb == [] and matched == [[1000, 5700, 0, 5600, 0, 220]]
    b = [4750, 5700, 3100, 5600, 0, 120]
    ra, rb, matched = po.pair_ends(100000, False, 10000, [a], [b])
    assert ra == [a] and rb == [b] and matched is None
    # circular reference: a pair across the join. distance = 150 - 99900 = -99750 < -50 -> + refLen = 250 < 500:
    # needs expectedDistance < 250*3/2 = 375 and > 250*2/3 = 166; expectedDistance = 1200 - 10000 + 9000 = 200
    a = [99000, 99900, 0, 9000, 0, 100]
    b = [150, 1100, 1200, 7800, 0, 90]
    assert po.pair_ends(100000, True, 10000, [a], [b])[2] == [[99000, 1100, 0, 7800, 0, 190]]
    assert po.pair_ends(100000, False, 10000, [a], [b])[2] is None


def _uniform_segments(gaps_after, lead, trail, first_seed=100):
    """segments of a seed sequence with distinct seeds first_seed, first_seed+1, ..: lead, s0, gaps_after[0], s1, .., trail"""
    seg = [lead]
    for i, g in enumerate(gaps_after + [trail]):
        seg += [first_seed + i, g]
    return seg


def test_seed_space_chunking_by_hand():
    """chunkWorker (overlap/overlap.go:253-318), k = 10, chunk_size 3000 bases, overlap 1000, on 400 seeds that sit 20 bases
    apart (gap 10), 5 bases of lead, 7 of trail: length 5 + 400*10 + 399*10 + 7 = 8002, numChunks = 8002/3000 + 1 = 3.

    GetNextSeedOffset(i) = gap after seed i + k = 20 (17 after the last seed, 15 for i = -1: the lead + k).
    Piece 1: prevSeedIndex 0, totalOffset = GetSeedOffset(0) = 5. The count loop stops at 100 seeds (2000 bases < 3000):
      newFirstGap = 15 - 10 = 5, length 2005, SubSequence(0, 99, 2005, offset 5 - 5 = 0, inset 8002 - 5 - 2005 + 5 = 5997);
      totalOffset = 5 + 2000 = 2005, prevSeedIndex = 100, then back 5 seeds (5 * 20 = 100 < overlap/2 = 500, the count of
      5 stops it): prevSeedIndex 95, totalOffset 1905 (= 5 + 95*20: the position of seed 95).
    Piece 2: (95, 194): newFirstGap 10, length 2010, offset 1895, inset 8002 - 1905 - 2010 + 10 = 4097; on to 190 / 3805.
    Piece 3: (190, 289): length 2010, offset 3795, inset 2197; on to 285 / 5705.
    285 >= 400 - 150: the rest in one piece (:268-277): length = GetSeedOffsetFromEnd(285) + k + newFirstGap =
      (8002 - (5 + 285*20 + 10)) + 10 + 10 = 2287 + 20 = 2307, SubSequence(285, 399, 2307, 5705 - 10 = 5695, 0).
    Every piece satisfies offset + length + inset = 8002. A piece's segments are s.segments[2*start : 2*end + 3]."""
    s = _uniform_segments([10] * 399, 5, 7)
    got = po.chunk_seed_sequence(s, 8002, 3000, 15, 1000, 10)
    want = [(0, 99, 2005, 0, 5997), (95, 194, 2010, 1895, 4097), (190, 289, 2010, 3795, 2197), (285, 399, 2307, 5695, 0)]
    assert [(ln, off, ins) for _, ln, off, ins in got] == [w[2:] for w in want]
    for (seg, _, _, _), (a, b, _, _, _) in zip(got, want):
        assert list(seg) == s[2 * a: 2 * b + 3]
    # a sequence of one chunk_size or with fewer than 3 * min_seeds seeds goes in whole (:259-263) - or not at all
    short = _uniform_segments([10] * 39, 5, 7)
    assert [(ln, off, ins) for _, ln, off, ins in po.chunk_seed_sequence(short, 802, 3000, 15, 1000, 10)] == [(802, 0, 0)]
    assert po.chunk_seed_sequence(_uniform_segments([10] * 9, 5, 7), 202, 3000, 15, 1000, 10) == []


def test_seed_space_chunking_with_a_dropped_piece_by_hand():
    """The same walk over 510 seeds with a sparse stretch: seeds 0..199 are 20 bases apart, the ten gaps after seeds
    199..208 are 490 (500 bases per step), seeds 209..509 are 20 apart again; lead 5, trail 7. pos[199] = 3985,
    pos[209] = 8985, pos[509] = 14985, length 15002. min_seeds = 16.

    Pieces 1 and 2 as before: (0, 99, 2005, 0, 12997), (95, 194, 2010, 1895, 11097); on to seed 190, totalOffset 3805.
    From seed 190 the count loop adds 9 * 20 = 180, then 500 per seed: 680, 1180, 1680, 2180, 2680, 3180 >= 3000 after 15
    seeds (190..204) — fewer than min_seeds: the piece is dropped (:303-314). prevSeedIndex = 205. The walk back by
    overlap/2 that follows does NOT happen: unlike the branch that keeps a piece, this one does not reset lengthInBases
    before its loop, so `lengthInBases < overlap/2` reads 3180 < 500. (My first derivation reset it, walked one seed back
    and expected offset 2815 for the next piece; the restatement said 3315 — and the source agrees with the restatement.)
    totalOffset stays 3805 although seed 205 sits at 6985: from here on offset and inset no longer tile the read.
    From seed 205: 4 * 500 = 2000, + 20 (the short gap after 209) = 2020, + 49 * 20 = 3000 after 54 seeds (205..258):
      newFirstGap = GetNextSeedOffset(204) - k = 490, length 3490, offset 3805 - 490 = 3315,
      inset 15002 - 3805 - 3490 + 490 = 8197; totalOffset = 3805 + 3000 = 6805, back five seeds of 20: seed 254, 6705.
    (254, 353): length 2010, offset 6695, inset 15002 - 6705 - 2010 + 10 = 6297; on to 349 / 8605.
    (349, 448): length 2010, offset 8595, inset 4397; on to 444 / 10505.
    444 >= 510 - 150: the rest: GetSeedOffsetFromEnd(444) = 15002 - (8985 + 235*20 + 10) = 1307, length 1327,
      SubSequence(444, 509, 1327, 10495, 0)."""
    gaps = [10] * 199 + [490] * 10 + [10] * 300
    s = _uniform_segments(gaps, 5, 7)
    assert len(s) == 2 * 510 + 1
    got = po.chunk_seed_sequence(s, 15002, 3000, 16, 1000, 10)
    want = [(0, 99, 2005, 0, 12997), (95, 194, 2010, 1895, 11097), (205, 258, 3490, 3315, 8197), (254, 353, 2010, 6695, 6297),
            (349, 448, 2010, 8595, 4397), (444, 509, 1327, 10495, 0)]
    assert [(ln, off, ins) for _, ln, off, ins in got] == [w[2:] for w in want]
    for (seg, _, _, _), (a, b, _, _, _) in zip(got, want):
        assert list(seg) == s[2 * a: 2 * b + 3]
    # with min_seeds = 15 the sparse piece is kept: (190, 204) of 3180 + 10 bases at offset 3795
    got15 = po.chunk_seed_sequence(s, 15002, 3000, 15, 1000, 10)
    assert [(ln, off, ins) for _, ln, off, ins in got15][2] == (3190, 3795, 15002 - 3805 - 3190 + 10)


def test_gap_range_by_hand():
    """gapRange (seeds/alignment.go:411-424): minGap = gap*2/3 - k, maxGap = gap*3/2 + k + 1 (Go's integer division truncates
    towards zero); a negative minGap becomes -k (and a negative maxGap 0); otherwise a maxGap below 20 becomes 20 with
    minGap 0.
      gap 100, k 10: 66 - 10 = 56, 150 + 11 = 161
      gap   6, k 10: 4 - 10 = -6 -> -10; 9 + 11 = 20 stays
      gap -30, k 10: -20 - 10 = -30 -> -10; -45 + 11 = -34 -> 0
      gap  16, k 10: 10 - 10 = 0 (not negative); 24 + 11 = 35 >= 20 stays
      gap   6, k  4: 4 - 4 = 0; 9 + 5 = 14 < 20 -> (0, 20)"""
    assert po.gap_range(100, 10) == (56, 161)
    assert po.gap_range(6, 10) == (-10, 20)
    assert po.gap_range(-30, 10) == (-10, 0)
    assert po.gap_range(16, 10) == (0, 35)
    assert po.gap_range(6, 4) == (0, 20)


def _kmer_id(s):
    """A = 0, C = 1, G = 2, T = 3, first base most significant (sequence/sequence.go:520-528 KmerValue)."""
    v = 0
    for c in s:
        v = v * 4 + "ACGT".index(c)
    return v


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def test_add_seeds_by_hand():
    """AddSeeds (seeds/seeds.go:62-156), k = 5, on the 62 bases below with an empty index.

    The walk (:83-127): kmer = KmerAt(0), nextIndex = 5; a block rolls NextKmer k times, i.e. examines the k-mers that START
    at p0+1 .. p0+5; after a block nextIndex += k, kmer = KmerAt(nextIndex), nextIndex += k: the next block starts 3k = 15
    further on. With 62 bases (loop while nextIndex < 57) the blocks examine starts 1-5, 16-20, 31-35, 46-50.
    Values: start 3 (TCACA) 5.0; starts 17 (GCTCA) and 18 (CTCAC) 7.0 each; start 48 (GAGAG) 2.0; everything else 0.
      block 1: best = TCACA 5.0      block 2: GCTCA 7.0 (`value > bestValue` is strict: the first of the two wins)
      block 3: nothing above 0.0: bestValue stays 0 and the insertion below places nothing      block 4: GAGAG 2.0
    topN insertion (:108-120; slot 0 is the bottom, a new value passes every smaller one), minSeeds = 3:
      5.0 -> [0, 0, 5];  7.0 -> [0, 5, 7];  2.0 passes only the 0 -> [2, 5, 7]  =  GAGAG, TCACA, GCTCA
    Registration (:131-154) in that order, each k-mer followed by its reverse complement:
      GAGAG, CTCTC, TCACA, TGTGA, GCTCA, TGAGC.
    With minSeeds = 4 the bottom slot stays unfilled and holds k-mer 0: AAAAA and TTTTT are registered FIRST."""
    s1 = "GGATCACAGTCTACACTGCTCACTCCAACCCCGGCCCCTGAGTCCGAGGAGAGGGTGCTTCA"
    assert len(s1) == 62 and (s1[3:8], s1[17:22], s1[18:23], s1[48:53]) == ("TCACA", "GCTCA", "CTCAC", "GAGAG")
    ranks = [0.0] * 4 ** 5
    ranks[_kmer_id("TCACA")] = 5.0
    ranks[_kmer_id("GCTCA")] = 7.0
    ranks[_kmer_id("CTCAC")] = 7.0
    ranks[_kmer_id("GAGAG")] = 2.0
    want = ["GAGAG", "CTCTC", "TCACA", "TGTGA", "GCTCA", "TGAGC"]
    assert [_rc(x) for x in want[0::2]] == want[1::2]
    g = po.SeedIndex(5)
    g.add_seeds(s1.encode(), 3, ranks)
    assert list(g.seeds()) == [_kmer_id(x) for x in want]
    g4 = po.SeedIndex(5)
    g4.add_seeds(s1.encode(), 4, ranks)
    assert list(g4.seeds()) == [_kmer_id(x) for x in ["AAAAA", "TTTTT"] + want]

    # A second sequence against the index of the first call: a block that meets a seed is abandoned (:91-94).
    #   s2 holds GCTCA (a seed now) at start 3: block 1 examines starts 1 (GTGCT, value 100: discarded with the block), 2, and
    #   3 -> reset at nextIndex = 8; nextIndex += 5 -> KmerAt(13) -> nextIndex = 18: block 2 examines starts 14-18, where
    #   start 15 (TAGCC) has value 9.0; block 3 (starts 29-33) has nothing. minSeeds = 2: topN = [0, TAGCC] -> AAAAA, TTTTT,
    #   TAGCC, GGCTA are appended.
    s2 = "CGTGCTCAACCGTCGTAGCCATGCTGCTTCATTGCAGGTT"
    assert len(s2) == 40 and s2[3:8] == "GCTCA" and s2[15:20] == "TAGCC" and s2[1:6] == "GTGCT"
    ranks[_kmer_id("GTGCT")] = 100.0
    ranks[_kmer_id("TAGCC")] = 9.0
    g.add_seeds(s2.encode(), 2, ranks)
    assert list(g.seeds()) == [_kmer_id(x) for x in want + ["AAAAA", "TTTTT", "TAGCC", "GGCTA"]]


def test_matches_by_hand():
    """SeedIndex.Matches (seeds/seeds.go:335-353) over six chunks and ten seeds:
        c0 {0, 1, 2, 3}   c1 {0, 4, 5}   c2 {0, 1, 2, 6, 7}   c3 {0, 8}   c4 {0, 9, 3}   c5 {0}
    Query seeds in order 0, 1, 1, 2, 3, 9, 6, 7. The sets handed to GetSharedIDs (:340-346): seed 0 sits in all six
    chunks (Size() == len(sequences): left out); the second 1 repeats its predecessor (left out); 1, 2, 3, 9, 6, 7 go in:
    six sets >= 5, minCount = int(0.25 * 6 + 0.5) = 2. Chunks per set: 1 {c0, c2}, 2 {c0, c2}, 3 {c0, c4}, 9 {c4}, 6 {c2},
    7 {c2} -> c0 counts 3, c2 counts 4, c4 counts 2, the rest 0: ids with at least two sets = 0, 2, 4 (all ids sit in one
    64-bit word: no set is dropped on the way, and level 2 of the four-level soft union is exact, util/bitset.go:376-387)."""
    chunks = [[0, 1, 2, 3], [0, 4, 5], [0, 1, 2, 6, 7], [0, 8], [0, 9, 3], [0]]
    assert po.matches(chunks, 10, [0, 1, 1, 2, 3, 9, 6, 7], 0.25) == [0, 2, 4]
    # hitFraction 0.5: minCount = int(3.5) = 3 -> c0 (3) and c2 (4)
    assert po.matches(chunks, 10, [0, 1, 1, 2, 3, 9, 6, 7], 0.5) == [0, 2]
    # four usable seeds only (1, 2, 3, 9): 'not many usable seeds in the query' -> nothing (:347-349)
    assert po.matches(chunks, 10, [0, 1, 2, 3, 9], 0.25) == []
    # a repeat that is NOT adjacent counts twice (:342 only compares with the previous included seed):
    # 1, 2, 1, 3, 9, 6 -> six sets, minCount 2: c0 = 1 + 1 + 1 + 1 = 4, c2 = 1 + 1 + 1 + 1 = 4, c4 = 2
    assert po.matches(chunks, 10, [1, 2, 1, 3, 9, 6], 0.25) == [0, 2, 4]
    # ... which matters at the threshold: 4, 5, 4, 5, 8 -> five sets (minCount int(1.75) = 1 -> level 1: any chunk hit):
    # c1 (4, 5, 4, 5) and c3 (8)
    assert po.matches(chunks, 10, [4, 5, 4, 5, 8], 0.25) == [1, 3]
