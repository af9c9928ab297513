"""Per-call times of the first 14 calls of a fresh mapper (device-resident, then packed pinned): nothing may be allocated
in the middle of a call after the first one or two."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
n, L = 1000000, 10000
ref = synth.reference(1, 4_600_000)
vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
pinned = torch.empty(n * L, dtype=torch.uint8).pin_memory()
synth.reads(ref, 12, n, L, circular=True, out=pinned.numpy()); offs = np.arange(n + 1, dtype=np.int64) * L
pk, boff, lens = dp.pack_batch(pinned.numpy(), offs)
pkp = torch.from_numpy(pk).pin_memory()
d = pinned.cuda()
gm = dp.Mapper(ref, vals, circular=True)
for mode in ("device", "packed", "device"):
    ts = []
    for it in range(14):
        t = time.time()
        if mode == "device":
            gm.map_batch_device(d.data_ptr(), offs)
        else:
            gm.map_batch_packed(pkp.data_ptr(), boff, lens)
        ts.append((time.time() - t) * 1e3)
    print(mode, " ".join("%.1f" % x for x in ts), flush=True)
