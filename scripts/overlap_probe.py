"""Stage timings of dp_overlapper_round on a synthetic read set (BASELINE config 5 shape): python scripts/overlap_probe.py
[n_reads] [read_len] [rounds] [oracle_reads]."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import downpore_b200 as dp  # noqa: E402
from tools import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
oracle_reads = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ref = synth.reference(1, 4_600_000)
t = time.time()
rd = synth.reads(ref, 15, n, L, circular=True)
offs = np.arange(n + 1, dtype=np.int64) * L
print("reads generated in %.1f s" % (time.time() - t), flush=True)
t = time.time()
g = dp.Overlapper(rd, offs, None)
print("create (pack %d reads): %.2f s" % (n, time.time() - t), flush=True)
t = time.time()
counts = g.kmer_counts()
t1 = time.time()
vals = dp.kmer_values(counts, 10)
g.set_values(vals)
print("kmer counts %.2f s, values %.2f s" % (t1 - t, time.time() - t1), flush=True)
first = 0
for r in range(rounds):
    t = time.time()
    res = g.round(first_sequence=first)
    wall = time.time() - t
    d = {k: getattr(res, k) for k in ("num_seeds", "num_queries", "num_chunks", "num_hits", "read_seeds", "chunk_seeds",
                                      "seed_postings", "posting_entries", "candidates", "pairs", "next_first_sequence")}
    ms = {k: round(getattr(res, k), 3) for k in ("ms_total", "ms_select", "ms_queries", "ms_scan", "ms_chunk", "ms_index",
                                                 "ms_lookup", "ms_align", "ms_collect")}
    print(json.dumps(dict(round=r, wall_s=round(wall, 4), **d, **ms)), flush=True)
    first = res.next_first_sequence
if oracle_reads:
    from oracle import pyoracle as po
    m = oracle_reads
    t = time.time()
    ov = po.overlap_values(rd[:offs[m]], offs[:m + 1], 10)
    t1 = time.time()
    o = po.OverlapRound(rd[:offs[m]], offs[:m + 1], ov)
    print("oracle on %d reads: values %.2f s, one round %.2f s, %d hits" % (m, t1 - t, time.time() - t1, o.num_hits))
