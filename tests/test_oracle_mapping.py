"""CPU tests of the oracle's seeds/ and mapping/ layers (the reference has no tests there: SURVEY.md section 4).

They pin the oracle against (a) hand-checkable truths on synthetic data — reads map back to where they were drawn
from, with the inclusive-end convention of Q3 — and (b) the committed golden fixtures, so an accidental change to the
oracle shows up as a diff."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as po
from tools import synth

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def mapper_300k():
    ref = synth.reference(2, 300_000)
    vals = po.kmer_values(ref, 11)
    return ref, po.Mapper(ref, vals, circular=True)


def test_reads_map_to_their_origin(mapper_300k):
    ref, om = mapper_300k
    n, L = 200, 8000
    rd, truth = synth.reads(ref, 5, n, L, circular=True, with_truth=True)
    offs = np.arange(n + 1, dtype=np.int64) * L
    rows, out_off, ctr = om.map_batch(rd, offs, threads=4)
    mapped = 0
    for i in range(n):
        r = rows[out_off[i]:out_off[i + 1]]
        if len(r) == 1:
            mapped += 1
            start, end, qoff, qin, rc, ids = (int(v) for v in r[0])
            assert rc == truth[i, 1]
            # reported start = template start + the unmapped lead-in of the read (query offset on '+', query inset
            # on '-'), up to the indel drift accumulated over that lead-in
            lead = qin if rc else qoff
            d = (start - (truth[i, 0] + lead)) % len(ref)
            d = min(d, len(ref) - d)
            assert d < 100 + 0.15 * lead, (i, start, lead, truth[i])
    assert mapped >= 0.95 * n
    assert ctr["sort_ties_unpinned"] == 0


def test_error_free_read_coordinates_q3(mapper_300k):
    """An exact copy of ref[a:b) must map to [first seed .. last seed end], query span likewise; window hits carry
    inclusive end coordinates (Q3), so End - Start and query end - query start agree."""
    ref, om = mapper_300k
    a, b = 50_000, 58_000
    read = ref[a:b].copy()
    rows, out_off, _ = om.map_batch(read, np.array([0, len(read)], dtype=np.int64))
    assert len(rows) == 1
    start, end, qoff, qin, rc, ids = (int(v) for v in rows[0])
    assert rc == 0
    assert start - a == qoff                      # same offset on both sides
    assert (len(read) - qin) - qoff == end - start  # identical spans: both ends use the inclusive convention
    assert 0 <= qoff < 100 and 0 < qin < 100
    assert ids > 0


def test_reverse_strand_and_origin_spanning(mapper_300k):
    ref, om = mapper_300k
    L = len(ref)
    comp = np.frombuffer(b"TGCA", dtype=np.uint8)
    code = ((ref >> 1) ^ ((ref & 4) >> 2)) & 3
    fwd = np.concatenate([ref[L - 3000:], ref[:3000]])
    rcr = comp[code[10_000:16_000]][::-1].copy()
    bases = np.concatenate([fwd, rcr])
    rows, out_off, _ = om.map_batch(bases, np.array([0, 6000, 12000], dtype=np.int64))
    assert out_off.tolist() == [0, 1, 2]
    assert rows[0][4] == 0 and rows[0][0] >= L - 3000 and rows[0][1] <= 3000  # wraps the circular join
    assert rows[1][4] == 1 and 10_000 <= rows[1][0] and rows[1][1] <= 16_000


@pytest.mark.parametrize("name", ["small_circular", "small_linear"])
def test_golden_fixture(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ref, circular, reads = make_golden.case_inputs(name)
    vals = po.kmer_values(ref, 11)
    om = po.Mapper(ref, vals, circular=circular)
    assert om.num_seeds == int(g["num_seeds"]) and om.num_chunks == int(g["num_chunks"])
    bases, offs = make_golden.concat(reads)
    rows, out_off, ctr = om.map_batch(bases, offs, threads=2)
    assert np.array_equal(out_off, g["out_off"])
    assert np.array_equal(rows, g["rows"])
    assert [ctr[k] for k in po.COUNTER_NAMES] == g["counters"].tolist()


def test_chunk_layout_and_q13(mapper_300k):
    """mapping.go:79-95: ten interleaved groups of chunk_size chunks stepping by 10*chunk-edge, plus the join chunk."""
    ref, om = mapper_300k
    L = len(ref)
    starts = []
    for j in range(10):
        i = j * 10000
        while i < L - 5000:
            starts.append(i)
            i += 10 * 10000 - 1000
    assert om.num_chunks == len(starts) + 1
    for c, s in enumerate(starts):
        ch = om.chunk(c)
        end = min(s + 10000, L)
        assert (ch["offset"], ch["length"], ch["inset"]) == (s, end - s, L - end + 1)
    j = om.chunk(om.num_chunks - 1)
    assert (j["offset"], j["inset"], j["length"]) == (L - 1000, L - 1000 + 1, 2000)
    # Q2: the 2000-base join chunk is a raw packedSequence with len%4==0: four bases are never scanned
    seg = j["segments"]
    assert int(seg[0::2].sum()) + 11 * (len(seg) // 2) == 2000 - 4


def test_oracle_cli_paf(tmp_path):
    ref = synth.reference(9, 150_000)
    rd = synth.reads(ref, 10, 20, 5000, circular=True)
    synth.write_fasta(tmp_path / "ref.fa", ["ref desc"], [ref])
    synth.write_fasta(tmp_path / "reads.fa", ["read%d" % i for i in range(20)], [rd[i * 5000:(i + 1) * 5000] for i in range(20)])
    exe = os.path.join(os.path.dirname(po.__file__), "oracle_map")
    po.build()
    out = subprocess.run([exe, "-i", str(tmp_path / "reads.fa"), "-r", str(tmp_path / "ref.fa"), "-n", "2"],
                         capture_output=True, text=True, check=True)
    lines = out.stdout.strip().split("\n")
    vals = po.kmer_values(ref, 11)
    om = po.Mapper(ref, vals, circular=True)
    rows, out_off, _ = om.map_batch(rd, np.arange(21, dtype=np.int64) * 5000)
    want = po.paf_lines(rows, out_off, ["read%d" % i for i in range(20)], [5000] * 20, "ref desc", len(ref), True)
    assert lines == want
    assert "Unmapped:" in out.stderr
