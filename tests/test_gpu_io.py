"""GPU parity tests of the input and output side: device-side record splitting against the oracle's
restatement of readFasta (sequence/seqio.go:188-267), reads mapped where they lie in the file image, and device-side PAF
formatting against Mapper.AsString (mapping/mapping.go:112-122). Byte work: the bar is bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402

import downpore_b200 as dp  # noqa: E402

K = 11


def spans_as_pairs(image, recs):
    image = bytes(image)
    return [(image[r["name_start"]:r["name_start"] + r["name_len"]], image[r["seq_start"]:r["seq_start"] + r["seq_len"]])
            for r in recs]


def random_file(rng, n, fastq, newline_at_end=True, odd=False):
    """A FASTA / FASTQ text with the things the reader's rules trip over: names with blanks and tabs around them, empty
    lines, sequences below the minimum, quality lines that start like sequences, names or comments."""
    out = []
    for i in range(n):
        L = int(rng.integers(1, 400))
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L))
        name = b"read_%d  \t" % i if i % 3 == 0 else b" r%d extra fields" % i
        out.append((b"@" if fastq else b">") + name + b"\n")
        if odd and i % 7 == 3:
            out.append(b"\n")  # an empty line: a name line with an empty name
        if odd and i % 11 == 5 and not fastq:
            out.append(b">second header without a sequence\n")
        out.append(seq + b"\n")
        if odd and not fastq and i % 5 == 1:  # a second sequence line under the same name: a record of its own
            out.append(bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(rng.integers(1, 300)))) + b"\n")
        if fastq:
            out.append(b"+\n")
            q = bytes(rng.integers(33, 75, L).astype(np.uint8))
            if odd and i % 4 == 0:
                q = b"ACGT"[i % 4:i % 4 + 1] + q[1:]  # looks like a sequence line
            if odd and i % 4 == 1:
                q = b"@" + q[1:]  # looks like a fastq name line
            if odd and i % 4 == 2:
                q = b"+" + q[1:]
            out.append(q + b"\n")
    text = b"".join(out)
    return text if newline_at_end else text[:-1]


@pytest.mark.parametrize("fastq", [False, True])
@pytest.mark.parametrize("newline_at_end", [True, False])
@pytest.mark.parametrize("min_length", [0, 120])
def test_split_records_matches_the_reader(fastq, newline_at_end, min_length):
    rng = np.random.default_rng(5 + fastq + 2 * newline_at_end)
    text = random_file(rng, 700, fastq, newline_at_end, odd=True)
    want = po.parse_fasta(text, min_length)
    recs, consumed, _ = dp.split_records(text, min_length)
    assert consumed == len(text)
    assert spans_as_pairs(text, recs) == want
    # the same image resident on the device, at an odd address (the line scan reads aligned 16-byte blocks)
    img = dp.DeviceImage(b"xyz" + text)
    recs2, _, _ = dp.split_records((img.ptr + 3, len(text)), min_length)
    assert np.array_equal(recs2, recs)


def test_split_records_edge_files():
    for text in (b"", b">only a header", b">only a header\n", b"ACGT\n", b">a\nACGT", b">a\nACGT\n", b"\n\n\n", b">a\n\nACGT\n>b\n",
                 b">a\nACGT\n>b\nTTTT\nGG\n>c", b"@a\nACGT\n+\nIIII\n", b"@a\nACGT\n+\nIIII", b"@a\nACGT\n+\n", b">a\r\nACGT\r\n",
                 b">a\nNNNN\nacgt\nACGU\n", b">a\n" + b"A" * 100000 + b"\n" + b">b\n" + b"C" * 70001):
        for ml in (0, 5):
            want = po.parse_fasta(text, ml)
            recs, consumed, _ = dp.split_records(text, ml)
            assert spans_as_pairs(text, recs) == want, (text[:40], ml)
    # the reference's log.Fatal: a sequence line of a fastq file without its '+' line
    for text in (b"@a\nACGT\nIIII\n", b"@a\nACGT\n+", b"@a\nACGT\n", b">a\nACGT\n@b\nACGT\n>c\n"):
        with pytest.raises(RuntimeError):
            po.parse_fasta(text, 0)
        with pytest.raises(dp.DownporeError, match="Invalid fastq"):
            dp.split_records(text, 0)


@pytest.mark.parametrize("fastq", [False, True])
def test_split_records_in_pieces(fastq):
    """A file handed over piece by piece (final=False: records in front of the piece's last name line, the rest rides at
    the head of the next piece) yields the records of the whole file."""
    rng = np.random.default_rng(17 + fastq)
    text = random_file(rng, 500, fastq, True, odd=True)
    want = po.parse_fasta(text, 50)
    got, pos, fq = [], 0, 0
    piece = 6000
    while True:
        end = min(len(text), pos + piece)
        final = end == len(text)
        chunk = text[pos:end]
        recs, consumed, fq = dp.split_records(chunk, 50, final=final, is_fastq=fq)
        got += spans_as_pairs(chunk, recs)
        if final:
            break
        assert consumed > 0
        pos += consumed
    assert got == want


@pytest.fixture(scope="module")
def mapper_pair():
    ref = synth.reference(3, 400_000)
    vals = po.kmer_values(ref, K)
    om = po.Mapper(ref, vals, circular=True, k=K)
    gm = dp.Mapper(ref, vals, circular=True, k=K)
    return ref, om, gm


@pytest.mark.parametrize("fastq", [False, True])
def test_file_image_to_paf_text(mapper_pair, fastq):
    """The whole host-free route: file image -> records (device) -> Map() on the reads where they lie -> PAF text
    (device) equals the oracle's reader + Map + AsString on the same file; pageable, page-locked and device images."""
    import torch
    ref, om, gm = mapper_pair
    rng = np.random.default_rng(23)
    reads = []
    for L, n in ((3000, 120), (999, 5), (2000, 5), (2001, 5), (5000, 40), (40, 3), (7001, 10)):
        rd = synth.reads(ref, 200 + L, n, L, circular=True)
        reads += [rd[i * L:(i + 1) * L] for i in range(n)]
    a = synth.reads(ref, 301, 6, 2500, circular=True)
    b = synth.reads(ref, 302, 6, 3500, circular=True)
    reads += [np.concatenate([a[i * 2500:(i + 1) * 2500], b[i * 3500:(i + 1) * 3500]]) for i in range(6)]  # chimeras
    order = rng.permutation(len(reads))
    lines = []
    for j, i in enumerate(order):
        r = reads[i].tobytes()
        name = b"read/%d len=%d" % (j, len(r)) if j % 2 else b"r%d" % j
        lines.append((b"@" if fastq else b">") + name + b"\n" + r + b"\n")
        if fastq:
            lines.append(b"+\n" + b"I" * len(r) + b"\n")
    text = b"".join(lines)
    min_length = 100
    parsed = po.parse_fasta(text, min_length)
    assert len(parsed) == len(reads) - 3
    bases = np.frombuffer(b"".join(s for _, s in parsed), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum([len(s) for _, s in parsed])]).astype(np.int64)
    orow, ooff, _ = om.map_batch(bases, offs, threads=4)
    want = "".join(l + "\n" for l in po.paf_lines(orow, ooff, [n.decode() for n, _ in parsed], np.diff(offs), "ref", len(ref), True))
    assert want.count("\n") >= len(parsed) - 8
    pinned = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).pin_memory()
    for image in (text, (pinned.data_ptr(), len(text)), dp.DeviceImage(text)):
        recs, _, _ = dp.split_records(image, min_length)
        assert spans_as_pairs(text, recs) == parsed
        maps, off = gm.map_batch_spans(image, recs)
        assert np.array_equal(off, ooff)
        got = gm.paf_block(image, recs, maps, off)
        assert got.decode() == want
    # and the per-line entry agrees with the block
    one = gm.as_string(maps[0], parsed[0][0].decode(), len(parsed[0][1]))
    assert want.startswith(one + "\n") or int(off[1]) == 0
