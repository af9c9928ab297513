"""Round-2 groundwork (overlap path): the oracle's index-based restatement of seedAligner.PairwiseAlignments
(seeds/alignment.go:274-616) against a second, object-based transcription of the same Go text, on random pairs of seed
sequences — plus the behaviour one expects on clean inputs. The reference has no test here; two transcriptions that
agree guard against slips, not against a shared misreading."""
import numpy as np
import pytest

from oracle import pyoracle as po


class GoPanic(Exception):
    pass


class St:
    __slots__ = ("aPos", "bPos", "aGap", "bGap", "aGapIndex", "length", "prev")


def gap_range(gap, k):
    tdiv = lambda a, b: int(a / b) if a * b < 0 else a // b      # Go truncates toward zero
    mn, mx = tdiv(gap * 2, 3) - k, tdiv(gap * 3, 2) + k + 1
    if mn < 0:
        mn = -k
        if mx < 0:
            mx = 0
    elif mx < 20:
        mx, mn = 20, 0
    return mn, mx


def pairwise_spec(aseg, bseg, min_matches, k, max_length=500):
    aset, bset = set(aseg[1::2]), set(bseg[1::2])
    if min_matches == 0:
        min_matches = 1
    # prepareInitial
    max_a = len(aseg) - min_matches * 2 + 1
    ared, amap, initials = [], [], []
    offset, prev = -k, -1
    for i in range(1, len(aseg), 2):
        sd = aseg[i]
        if sd not in bset or (sd == prev and (i >= len(aseg) - 2 or aseg[i + 2] == prev)):
            offset += aseg[i - 1] + k
            max_a -= 1
            continue
        prev = sd
        offset += aseg[i - 1] + k
        if len(ared) + 2 > max_length - 1 or len(amap) >= max_length // 2:
            raise GoPanic()
        ared += [offset, sd]
        amap.append(i // 2)
        offset = -k
        if len(amap) - 1 <= max_a:
            s = St()
            s.aPos, s.length, s.prev = (len(amap) - 1) * 2 + 1, 0, None
            initials.append(s)
    ared.append(0)
    while initials and initials[-1].aPos > max_a:
        initials.pop()
    open_ = [None] * 500
    results = [None] * 500
    osz = rsz = 0

    def remove_open(index, mm, osz_, rsz_):
        if not (0 <= index < 500 and 0 <= osz_ - 1 < 500):
            raise GoPanic()
        s = open_[index]
        open_[index] = open_[osz_ - 1]
        osz_ -= 1
        if s is None:
            raise GoPanic()
        if s.length >= mm:
            if (s.length * 2) // 3 > mm:
                mm = (s.length * 2) // 3
            if not 0 <= rsz_ < 500:
                raise GoPanic()
            results[rsz_] = s
            rsz_ += 1
        return osz_, rsz_, mm

    blen = len(bseg)
    max_b = blen - min_matches * 2 + 1
    boff, prev = 0, -1
    for bi in range(1, blen, 2):
        bs = bseg[bi]
        if bs not in aset or (bs == prev and (bi >= blen - 2 or bseg[bi + 2] == prev)):
            boff += bseg[bi + 1] + k
            continue
        prev = bs
        found = prev_found = -1
        i = osz - 1
        while i >= 0:
            s = open_[i]
            if s is None:
                raise GoPanic()
            s.bGap += boff
            mn, mx = gap_range(s.bGap, k)
            left = False
            while s.aGap < mn:
                if s.aGapIndex >= len(ared):
                    osz, rsz, min_matches = remove_open(i, min_matches, osz, rsz)
                    left = True
                    break
                if s.aGapIndex + 1 >= len(ared):
                    raise GoPanic()
                s.aGap += ared[s.aGapIndex + 1] + k
                s.aGapIndex += 2
            if left:
                break
            matched = False
            if s.aGap <= mx:
                g, j = s.aGap, s.aGapIndex
                while j < len(ared) and g <= mx:
                    if ared[j] == bs:
                        if found != -1 and i < prev_found < osz:
                            s2 = open_[prev_found]
                            if s2 is None:
                                raise GoPanic()
                            if s.aPos == s2.aPos and s.bPos == s2.bPos:
                                if s.length < s2.length:
                                    osz, _, _ = remove_open(i, osz, rsz, s.length + 1)     # (sic: argument order)
                                    matched = True
                                    break
                                osz, _, _ = remove_open(prev_found, osz, rsz, s2.length + 1)  # (sic)
                        found, prev_found = j, i
                        ns = St()
                        ns.prev, ns.aPos, ns.bPos, ns.aGapIndex = s, j, bi, j + 2
                        if j + 1 >= len(ared):
                            raise GoPanic()
                        ns.aGap, ns.bGap, ns.length = ared[j + 1], bseg[bi + 1], s.length + 1
                        if not 0 <= i < 500:
                            raise GoPanic()
                        open_[i] = ns
                        if (ns.length * 2) // 3 > min_matches:
                            min_matches = (ns.length * 2) // 3
                            max_b = blen - min_matches * 2 + 1
                        matched = True
                        break
                    if j + 1 >= len(ared):
                        raise GoPanic()
                    g += ared[j + 1] + k
                    j += 2
            if matched:
                break
            if s.length + (blen - bi) < min_matches:
                osz, rsz, min_matches = remove_open(i, min_matches, osz, rsz)
            else:
                s.bGap += bseg[bi + 1] + k
            i -= 1
        boff = 0
        if bi <= max_b:
            for s in initials:
                ap = s.aPos
                if ap != found and ared[ap] == bs:
                    if found != -1:
                        for j in range(osz):
                            if open_[j].bPos == bi and open_[j].aPos == ap:
                                found = ap
                                break
                    if found == ap or osz >= 500:
                        continue
                    ns = St()
                    ns.aPos, ns.bPos, ns.aGapIndex, ns.aGap = ap, bi, ap + 2, ared[ap + 1]
                    ns.bGap, ns.length, ns.prev = bseg[bi + 1], 1, None
                    open_[osz] = ns
                    osz += 1
    for i in range(osz):
        if open_[i].length >= min_matches:
            if rsz >= 500:
                raise GoPanic()
            results[rsz] = open_[i]
            rsz += 1
    out = []
    for i in range(rsz - 1, -1, -1):
        s = results[i]
        ma, mb = [0] * s.length, [0] * s.length
        while s is not None:
            ma[s.length - 1], mb[s.length - 1] = amap[s.aPos // 2], s.bPos // 2
            s = s.prev
        out.append((ma, mb))
    return out


def random_pair(rng, k):
    """b = a noisy copy of a stretch of a (seeds dropped, inserted, gaps jittered), embedded in unrelated seeds."""
    na = int(rng.integers(10, 120))
    a = np.empty(2 * na + 1, dtype=np.int64)
    a[0::2] = rng.integers(0, 120, size=na + 1)
    a[1::2] = rng.integers(0, 60 if rng.random() < 0.5 else 4000, size=na)     # small alphabets repeat seeds
    lo = int(rng.integers(0, na // 2))
    hi = int(rng.integers(lo + 3, na + 1))
    seeds, gaps = [], []
    for j in range(lo, hi):
        if rng.random() < 0.15:
            continue                                                            # seed lost
        if rng.random() < 0.1:
            seeds.append(int(rng.integers(0, 4000)))                            # spurious seed
            gaps.append(int(rng.integers(0, 60)))
        seeds.append(int(a[2 * j + 1]))
        gaps.append(max(0, int(a[2 * j] * rng.uniform(0.8, 1.25))))
    pre = int(rng.integers(0, 15))
    seeds = [int(x) for x in rng.integers(0, 4000, size=pre)] + seeds + [int(x) for x in rng.integers(0, 4000, size=pre)]
    gaps = [int(x) for x in rng.integers(0, 120, size=pre)] + gaps + [int(x) for x in rng.integers(0, 120, size=pre)]
    b = np.empty(2 * len(seeds) + 1, dtype=np.int64)
    b[1::2] = seeds
    b[0:-1:2] = gaps
    b[-1] = int(rng.integers(0, 100))
    return a, b


@pytest.mark.parametrize("seed", range(10))
def test_two_transcriptions_agree(seed):
    rng = np.random.default_rng(900 + seed)
    k = 10
    agree = hits = panics = 0
    for _ in range(150):
        a, b = random_pair(rng, k)
        mm = int(rng.integers(0, 12))
        try:
            want = pairwise_spec([int(x) for x in a], [int(x) for x in b], mm, k)
        except GoPanic:
            want = "panic"
        try:
            got = [(list(map(int, ma)), list(map(int, mb))) for ma, mb in po.pairwise_alignments(a, b, mm, k)]
        except RuntimeError as ex:
            assert "Go would panic" in str(ex)
            got = "panic"
        assert got == want, (seed, mm, a.tolist(), b.tolist())
        agree += 1
        hits += got != "panic" and len(got) > 0
        panics += got == "panic"
    assert hits > 30 and panics < agree // 4


def test_clean_inputs_behave():
    k = 10
    ns = 40
    a = np.empty(2 * ns + 1, dtype=np.int64)
    a[0::2] = 35
    a[1::2] = np.arange(500, 500 + ns)
    (ma, mb), = po.pairwise_alignments(a, a, 10, k)                  # identical: one chain through every seed
    assert list(ma) == list(range(ns)) and list(mb) == list(range(ns))
    b = a[20:61].copy()                                              # seeds 10..29 of a
    (ma, mb), = po.pairwise_alignments(a, b, 10, k)
    assert list(ma) == list(range(10, 30)) and list(mb) == list(range(20))
    assert po.pairwise_alignments(a, b, 25, k) == []                 # not enough seeds for the threshold
    c = b.copy()
    c[1::2] += 5000                                                  # no shared seed
    assert po.pairwise_alignments(a, c, 3, k) == []
