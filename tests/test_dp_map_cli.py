"""`dp_map` — the C++ host that stands where the reference's `downpore map` command stands (commands/map.go:33-116,
downpore.go:34-51, sequence/seqio.go:188-267) — against the oracle's CLI twin on the same files.

CPU tier: the host builds, keeps the reference's flag surface, and fails loudly without a GPU.
GPU tier: byte-identical PAF output and stderr counters for FASTA and FASTQ inputs with the parsing edge cases the
reference's reader has (min_length on the raw line, names with spaces, last line without newline, lower case / N)."""
import os
import subprocess

import numpy as np
import pytest

import downpore_b200 as dp
from oracle import pyoracle as po
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_MAP = os.path.join(ROOT, "oracle", "oracle_map")


def host():
    dp.build()
    assert os.path.exists(dp.HOST_PATH)
    return dp.HOST_PATH


def write_inputs(tmp, fastq, circular):
    ref = synth.reference(21, 250_000)
    extra = synth.reference(22, 30_000)  # second reference record: counted, not indexed (Q14)
    ref_path = os.path.join(tmp, "ref.fasta")
    with open(ref_path, "wb") as f:
        f.write(b">chrA some description\n" + bytes(ref) + b"\n>chrB\n" + bytes(extra) + b"\n")
    reads = []
    rd = synth.reads(ref, 31, 60, 6000, circular=circular)
    reads += [bytes(rd[i * 6000:(i + 1) * 6000]) for i in range(60)]
    for L in (300, 499, 500, 501, 1000, 1999, 2000, 2001, 2996, 3000, 4003):
        x = synth.reads(ref, 200 + L, 2, L, circular=circular)
        reads += [bytes(x[:L]), bytes(x[L:2 * L])]
    low = bytearray(reads[3].lower())
    low[100:110] = b"N" * 10
    low[0:1] = b"A"  # a sequence line must start in 'A'..'T' to be seen as one
    reads.append(bytes(low))
    a = synth.reads(ref, 41, 1, 4000, circular=circular)
    b = synth.reads(ref, 42, 1, 5000, circular=circular)
    reads.append(bytes(a) + bytes(b))  # chimera: exercises the later Map() rounds
    path = os.path.join(tmp, "reads.fastq" if fastq else "reads.fasta")
    with open(path, "wb") as f:
        for i, s in enumerate(reads):
            name = b"read%d  strand=? len=%d " % (i, len(s))
            if fastq:
                f.write(b"@" + name + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
            else:
                last = i == len(reads) - 1
                f.write(b">" + name + b"\n" + s + (b"" if last else b"\n"))  # the last line has no newline
    return ref_path, path


def run(cmd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, env=e)


def test_host_builds_and_keeps_the_flag_surface():
    h = host()
    r = run([h, "-bogus", "1"])
    assert r.returncode != 0 and b"Unrecognised argument:bogus" in r.stderr
    r = run([h, "-k", "x1", "-i", "a", "-r", "b"])
    assert r.returncode != 0 and b"Invalid integer argument value:x1" in r.stderr
    r = run([h])
    assert r.returncode != 0 and b"usage" in r.stderr


def test_host_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ref_path, reads_path = write_inputs(str(tmp_path), False, True)
    r = run([host(), "-input", reads_path, "-reference", ref_path])
    assert r.returncode != 0 and r.stdout == b"" and b"dp_kmer_counts" in r.stderr


def counters(stderr):
    keep = [ln for ln in stderr.decode().splitlines()
            if ln.split(":")[0] in ("Uniquely mapped", "Multiple mappings", "total", "Unmapped") or ln.startswith("K-mer")]
    return keep


@pytest.mark.gpu
@pytest.mark.parametrize("fastq,circular", [(False, True), (True, False)])
def test_cli_output_identical_to_oracle(tmp_path, fastq, circular):
    po.build()
    ref_path, reads_path = write_inputs(str(tmp_path), fastq, circular)
    args = ["-input", reads_path, "-r", ref_path, "-ci", "true" if circular else "false", "--num_workers", "4"]
    want = run([ORACLE_MAP] + args)
    assert want.returncode == 0, want.stderr
    # small batches: several dp_mapper_map_batch calls and batch-order reassembly are exercised
    got = run([host()] + args, env={"DOWNPORE_BATCH": "16", "DOWNPORE_BATCH_BYTES": str(1 << 22)})
    assert got.returncode == 0, got.stderr
    assert want.stdout.count(b"\n") > 60
    assert got.stdout == want.stdout
    assert counters(got.stderr) == counters(want.stderr)
    # non-default parameters through the aliases
    args2 = ["-i", reads_path, "-r", ref_path, "-ci", "true" if circular else "false", "-k", "10", "-q", "800",
             "-m", "1000", "-ch", "8000", "-s", "36"]
    want = run([ORACLE_MAP] + args2)
    got = run([host()] + args2)
    assert want.returncode == 0 and got.returncode == 0, (want.stderr, got.stderr)
    assert got.stdout == want.stdout and counters(got.stderr) == counters(want.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("fastq,circular", [(False, True), (True, False)])
def test_cli_device_io_identical_to_oracle(tmp_path, fastq, circular):
    """DOWNPORE_DEVICE_IO=1: the file goes to the GPU as it is; records are split, reads mapped where they lie and PAF
    lines formatted on the device (dp_split_records, dp_mapper_map_batch_spans, dp_mapper_paf_block). Small pieces: the
    piece cutting, the carried fastq state and the piece-order output are exercised."""
    po.build()
    ref_path, reads_path = write_inputs(str(tmp_path), fastq, circular)
    args = ["-input", reads_path, "-r", ref_path, "-ci", "true" if circular else "false"]
    want = run([ORACLE_MAP] + args)
    assert want.returncode == 0, want.stderr
    for piece in (1 << 30, 40000, 9000):
        got = run([host()] + args, env={"DOWNPORE_DEVICE_IO": "1", "DOWNPORE_PIECE_BYTES": str(piece)})
        assert got.returncode == 0, got.stderr
        assert got.stdout == want.stdout, piece
        assert counters(got.stderr) == counters(want.stderr)


@pytest.mark.gpu
def test_cli_on_two_gpus_in_one_process(tmp_path):
    """DOWNPORE_GPUS=0,1: one mapping thread per GPU inside one process, the index built on the first GPU and opened on the
    second from its image (dp_mapper_index_export -> dp_mapper_create_from_index, the peer-copy route INTEGRATION.md
    recommends to a Go host); output identical to the single-GPU run and to the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    po.build()
    ref_path, reads_path = write_inputs(str(tmp_path), False, True)
    args = ["-input", reads_path, "-r", ref_path, "-ci", "true"]
    want = run([ORACLE_MAP] + args)
    one = run([host()] + args, env={"DOWNPORE_BATCH": "8", "DOWNPORE_GPUS": "0"})
    two = run([host()] + args, env={"DOWNPORE_BATCH": "8", "DOWNPORE_GPUS": "0,1"})
    assert want.returncode == 0 and one.returncode == 0 and two.returncode == 0, (one.stderr, two.stderr)
    assert two.stdout == one.stdout == want.stdout
    assert counters(two.stderr) == counters(want.stderr)
