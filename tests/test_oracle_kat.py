"""Known-answer tests of the reference, restated against the oracle.

sequence/sequence_test.go (Test1..Test10) and util/bitset_test.go (Test1, Test2): these are the only golden
vectors the reference holds for the map path; they pin the oracle's sequence/ and util/ layers.
"""
import numpy as np
import pytest

from oracle import pyoracle as po

SEQ = "GGGAAGTGACTGCCTTAAAATGAGGGTTACCCCTTTTAGTTGACAAGACGCTTGCGGCTATTATGGCTAG"  # sequence_test.go:7-9


def kmer_set(s, k):  # sequence_test.go:11-40
    ks = np.zeros(4 ** k, dtype=np.uint8)
    count = 0
    for i in range(4, len(s) - k - 1, 5):
        ks[po.kmer_value(s[i:i + k])] = 1
        ks[po.kmer_value(s[i + 1:i + 1 + k])] = 1
        count += 2
    for sub in (s[0:k], s[1:k + 1], s[len(s) - k:]):
        x = po.kmer_value(sub)
        if not ks[x]:
            ks[x] = 1
            count += 1
    return ks, count


def mask_of(k):
    return (1 << (2 * k)) - 1


def test1_lengths():  # sequence_test.go:42-56
    for i in range(0, 5):
        s = SEQ[:len(SEQ) - i]
        assert len(po.Byte(s)) == len(po.Packed(s)) == len(s)


def test2_string():  # sequence_test.go:58-72
    for i in range(0, 5):
        s = SEQ[:len(SEQ) - i]
        assert po.Byte(s).string() == s
        assert po.Packed(s).string() == s


def test3_reverse_complement():  # sequence_test.go:74-81 (the reference only checks the full, len%4==2, string)
    assert po.Byte(SEQ).rc().string() == po.Packed(SEQ).rc().string()
    for i in (0, 1, 3):  # len % 4 != 0
        s = SEQ[:len(SEQ) - i]
        assert po.Byte(s).rc().string() == po.Packed(s).rc().string()
        assert po.Packed(s).rc().rc().string() == s


def test3b_reverse_complement_q2():
    """Q2: a raw sequence with len%4==0 has finalLen=0, so its RC has firstLen=0 and String() skips byte 0
    (sequence.go:244-253): the first four bases are lost and the tail is padded with 'A'."""
    s = SEQ[:68]
    true_rc = po.Byte(s).rc().string()
    assert po.Packed(s).rc().string() == true_rc[4:] + "AAAA"
    f = po.Packed(s).rc().fields()
    assert (f["firstLen"], f["finalLen"]) == (0, 4)


def test4_subsequence():  # sequence_test.go:82-98
    s1, s2 = po.Byte(SEQ), po.Packed(SEQ)
    for i in range(15, 20):
        assert s1.sub(i - 15, i).string() == s2.sub(i - 15, i).string() == SEQ[i - 15:i]
        assert s1.sub(i, i + 30).string() == s2.sub(i, i + 30).string() == SEQ[i:i + 30]
    for i in range(0, 5):  # beyond the reference: every alignment of both ends, and the RC of a window
        for j in range(15, 20):
            a, b = s1.sub(i, j), s2.sub(i, j)
            assert a.string() == b.string() == SEQ[i:j]
            assert len(a) == len(b)
            assert a.rc().string() == b.rc().string()


def test5_kmer_at():  # sequence_test.go:99-110
    s1, s2 = po.Byte(SEQ), po.Packed(SEQ)
    for i in range(len(SEQ) - 6):
        assert s1.kmer_at(i, 6) == s2.kmer_at(i, 6) == po.kmer_value(SEQ[i:i + 6])


def test6_count_kmers():  # sequence_test.go:112-154 (k=17 needs a 16 GiB table there; k=12 stands in for it here)
    ks, count = kmer_set(SEQ, 6)
    s1, s2 = po.Byte(SEQ), po.Packed(SEQ)
    assert s1.count_kmers(100, 6, mask_of(6), ks) == s2.count_kmers(100, 6, ks) == count
    assert s1.count_kmers(7, 6, mask_of(6), ks) >= 7
    assert s2.count_kmers(7, 6, ks) >= 7
    sub = SEQ[7:len(SEQ) - 7]
    s1, s2 = s1.sub(7, len(SEQ) - 7), s2.sub(7, len(SEQ) - 7)
    for k in (8, 12):
        ks, count = kmer_set(sub, k)
        assert s1.count_kmers(100, k, mask_of(k), ks) == s2.count_kmers(100, k, ks) == count


def test7_iterate_kmers():  # sequence_test.go:155-175
    s1, s2 = po.Byte(SEQ), po.Packed(SEQ)
    m = mask_of(6)
    for i in range(len(SEQ) - 6):
        k1, k2 = s1.kmer_at(i, 6), s2.kmer_at(i, 6)
        assert k1 == k2
        assert s1.next_kmer(k1, m, i + 6) == s2.next_kmer(k2, m, i + 6) == po.kmer_value(SEQ[i + 1:i + 7])


def test8_segments():  # sequence_test.go:176-209
    ks, count = kmer_set(SEQ, 6)
    s1, s2 = po.Byte(SEQ), po.Packed(SEQ)
    a, b = s1.write_segments(6, mask_of(6), ks), s2.write_segments(6, ks)
    assert len(a) == 2 * count + 1
    assert np.array_equal(a, b)
    a = s1.sub(2, len(SEQ) - 2).write_segments(6, mask_of(6), ks)
    b = s2.sub(2, len(SEQ) - 2).write_segments(6, ks)
    assert np.array_equal(a, b)


def test9_packing():  # sequence_test.go:211-233 — the one golden byte vector: "CGGT" -> 0x6B
    d = po.pack_bytes("CGGT", 2)
    assert d[0] == 0x6B and d[1] == 0
    s = "CGGT" * 5
    d = po.pack_bytes(s, len(s) // 4 + 1)
    assert all(x == 0x6B for x in d[:-1]) and d[-1] == 0
    assert list(po.Packed(s).bytes()) == [0x6B] * 5


def test10_short_kmers():  # sequence_test.go:235-264
    s1, s2 = po.Byte(SEQ), po.Packed(SEQ)
    assert np.array_equal(s1.short_kmers(6, False), s2.short_kmers(6, False))
    assert np.array_equal(s1.short_kmers(3, True), s2.short_kmers(3, True))


def test_bitset1_count_intersection():  # util/bitset_test.go:7-37
    a, b = po.IntSet(), po.IntSet()
    count = 0
    for i in range(1001, 3000, 5):
        a.add(i)
    for j in range(101, 2013, 3):
        b.add(j)
        if a.contains(j):
            count += 1
    assert count > 0
    assert a.count_intersection(b) == count and b.count_intersection(a) == count
    assert b.count_intersection_to(a, count + 10) == count
    assert a.count_intersection_to(b, count + 10) == count


def test_bitset2_shared_ids():  # util/bitset_test.go:38-161
    sets = [po.IntSet() for _ in range(20)]
    counts = [0] * 500
    c = {16: 0, 8: 0, 4: 0, 2: 0}
    for i in range(500):
        if i % 7 == 0:
            counts[i] = 16
        elif i % 5 == 0:
            counts[i] = 8
        elif i % 3 == 0:
            counts[i] = 4
        elif i % 2 == 0:
            counts[i] = 2
        if counts[i]:
            c[counts[i]] += 1
        for j in range(counts[i]):
            sets[j].add(i)
    assert (c[16], c[8], c[4], c[2]) == (72, 85, 114, 114)
    for fast in (False, True):
        for min_count, expect, floor in ((16, c[16], 16), (15, c[16], 16), (8, c[8] + c[16], 8),
                                         (4, c[8] + c[16] + c[4], 4), (2, c[8] + c[16] + c[4] + c[2], 2)):
            ids = po.get_shared_ids(sets, min_count, fast)
            assert len(ids) == expect, (min_count, fast)
            assert all(counts[int(i)] >= floor for i in ids)
            assert list(ids) == sorted(ids)


# ---- beyond the reference's tests: byteSequence as an independent check of the asm emulation --------------------

@pytest.mark.parametrize("seed", range(6))
def test_packed_scan_matches_byte_scan_on_windows(seed):
    """SubSequence windows visit every k-mer exactly once; raw sequences with len%4==0 lose 4 bases (Q2)."""
    rng = np.random.default_rng(seed)
    L = int(rng.integers(200, 400))
    s = "".join("ACGT"[c] for c in rng.integers(0, 4, L))
    k = int(rng.choice([5, 8, 11, 13]))
    ks = (rng.random(4 ** k) < 0.2).astype(np.uint8)
    p, b = po.Packed(s), po.Byte(s)
    for _ in range(20):
        i = int(rng.integers(0, L // 2))
        j = int(rng.integers(i + 4 * k, L + 1))
        pw, bw = p.sub(i, j), b.sub(i, j)
        assert np.array_equal(pw.write_segments(k, ks), bw.write_segments(k, mask_of(k), ks))
        assert np.array_equal(pw.rc().write_segments(k, ks), bw.rc().write_segments(k, mask_of(k), ks))
        assert pw.fields()["offset"] == bw.fields()["offset"] == i
        assert pw.fields()["inset"] == bw.fields()["inset"] + 1  # Q3
    raw = p.write_segments(k, ks)
    ref = b.write_segments(k, mask_of(k), ks)
    if L % 4:
        assert np.array_equal(raw, ref)
    else:
        cut = b.sub(0, L - 4).write_segments(k, mask_of(k), ks)
        assert np.array_equal(raw, cut)
