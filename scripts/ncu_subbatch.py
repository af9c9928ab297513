"""One 65 536-read sub-batch of config 2 on a single lane, device-resident and packed-pinned: the process profiled by ncu for
profiles/r2*_ncu_summary.txt (DP_LANES=1 DP_RAMP=0 so that every kernel runs alone on one whole sub-batch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DP_LANES", "1"); os.environ.setdefault("DP_RAMP", "0")
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
n, L = 65536, 10000
ref = synth.reference(1, 4_600_000)
vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
gm = dp.Mapper(ref, vals, circular=True)
rd = synth.reads(ref, 12, n, L, circular=True); offs = np.arange(n + 1, dtype=np.int64) * L
d = torch.from_numpy(rd).cuda()
pk, boff, lens = dp.pack_batch(rd, offs)
pkp = torch.from_numpy(pk).pin_memory()
gm.map_batch_device(d.data_ptr(), offs)
torch.cuda.synchronize()
torch.cuda.profiler.start() if hasattr(torch.cuda, "profiler") else None
gm.map_batch_device(d.data_ptr(), offs)
gm.map_batch_packed(pkp.data_ptr(), boff, lens)
torch.cuda.synchronize()
print(gm.stats())
