"""Generates tests/golden/*.npz from the oracle (the reference ships no mapping fixtures: SURVEY.md 8c).

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by tools/synth.py, so the fixtures only hold the expected outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case_inputs(name):
    """Deterministic inputs of a golden case: (ref, circular, reads list)."""
    if name == "small_circular":
        ref = synth.reference(21, 120_000)
        circular = True
    elif name == "small_linear":
        ref = synth.reference(22, 150_000)
        circular = False
    else:
        raise KeyError(name)
    reads = []
    rd = synth.reads(ref, 31, 96, 6000, circular=circular)
    reads += [rd[i * 6000:(i + 1) * 6000] for i in range(96)]
    for L in (499, 500, 1000, 1500, 1996, 2000, 2001, 2500, 2999, 3000, 3600, 4000, 4001, 5100, 6100, 9000):
        x = synth.reads(ref, 1000 + L, 3, L, circular=circular)
        reads += [x[i * L:(i + 1) * L] for i in range(3)]
    a = synth.reads(ref, 41, 6, 4500, circular=circular)
    b = synth.reads(ref, 42, 6, 5500, circular=circular)
    for i in range(6):  # two-segment chimeras (findSplitPoint)
        reads.append(np.concatenate([a[i * 4500:(i + 1) * 4500], b[i * 5500:(i + 1) * 5500]]))
    # error-free reads, one spanning the origin
    L = len(ref)
    reads.append(ref[1000:9000].copy())
    reads.append(np.concatenate([ref[L - 3000:], ref[:3000]]))
    lower = ref[20000:26000].copy()
    lower[::7] = np.frombuffer(b"acgtn", dtype=np.uint8)[np.arange(len(lower[::7])) % 5]
    reads.append(lower)  # lowercase / N bases
    return ref, circular, reads


def concat(reads):
    bases = np.concatenate(reads)
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    return bases, offs


def main():
    for name in ("small_circular", "small_linear"):
        ref, circular, reads = case_inputs(name)
        vals = po.kmer_values(ref, 11)
        om = po.Mapper(ref, vals, circular=circular)
        bases, offs = concat(reads)
        rows, out_off, ctr = om.map_batch(bases, offs, threads=4)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rows=rows, out_off=out_off,
                            num_seeds=om.num_seeds, num_chunks=om.num_chunks,
                            counters=np.array([ctr[k] for k in po.COUNTER_NAMES], dtype=np.int64))
        print(name, "reads", len(reads), "mappings", len(rows), ctr)


if __name__ == "__main__":
    main()
