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
