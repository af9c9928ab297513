"""Config 2 device-resident step time under a given host-thread setup (DP_LANES / DP_SYNC from the environment, cores via
taskset): the N = 8 box leaves a rank 4 vCPUs, which one GPU under `taskset -c 0-3` reproduces.
usage: taskset -c 0-3 env DP_SYNC=block DP_LANES=6 python scripts/host_starved.py [n_reads]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = 10000
os.sched_setaffinity(0, os.sched_getaffinity(0))
aff = sorted(os.sched_getaffinity(0))
ref = synth.reference(1, 4_600_000)
vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
gm = dp.Mapper(ref, vals, circular=True)
rd = synth.reads(ref, 12, n, L); offs = np.arange(n + 1, dtype=np.int64) * L
d = torch.from_numpy(rd).cuda()
ts = []
for it in range(9):
    torch.cuda.synchronize(); t = time.time()
    maps, off = gm.map_batch_device(d.data_ptr(), offs)
    ts.append((time.time() - t) * 1e3)
    del maps, off
ts = sorted(ts[2:])
print("cores %d lanes %s sync %s: step ms min %.2f median %.2f max %.2f -> %.0f Gbp/s (median)" % (
    len(aff), os.environ.get("DP_LANES", "6"), os.environ.get("DP_SYNC", "auto"), ts[0], ts[len(ts) // 2], ts[-1],
    n * L / ts[len(ts) // 2] / 1e6), flush=True)
