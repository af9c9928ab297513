"""single-lane stage times of config-2 sub-batches (A/B of library builds)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DP_LANES"] = "1"; os.environ["DP_RAMP"] = "0"
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
n, L = 262144, 10000
ref = synth.reference(1, 4_600_000)
vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
gm = dp.Mapper(ref, vals, circular=True)
rd = synth.reads(ref, 12, n, L); offs = np.arange(n + 1, dtype=np.int64) * L
d = torch.from_numpy(rd).cuda()
best = {}
for it in range(5):
    gm.map_batch_device(d.data_ptr(), offs); st = gm.stats()
    for k in ("ms_pack", "ms_extract", "ms_lookup", "ms_reduce", "ms_chain"):
        best[k] = min(best.get(k, 1e9), st[k])
print(sys.argv[1] if len(sys.argv) > 1 else "", {k: round(v, 3) for k, v in best.items()})
