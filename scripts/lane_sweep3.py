"""As lane_sweep2.py, but every setting gets a fresh mapper (settings read when lanes are created)."""
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
for v in [""] + sys.argv[1:]:
    env = dict(kv.split("=") for kv in v.split()) if v else {}
    os.environ.update(env)
    gm = dp.Mapper(ref, vals, circular=True)
    res = []
    for mode in ("device", "packed"):
        ts = []
        for it in range(9):
            t = time.time()
            if mode == "device":
                gm.map_batch_device(d.data_ptr(), offs)
            else:
                gm.map_batch_packed(pkp.data_ptr(), boff, lens)
            ts.append((time.time() - t) * 1e3)
        ts = sorted(ts[3:])
        res.append("%s median %.2f min %.2f" % (mode, ts[len(ts) // 2], ts[0]))
    print("[%s] %s" % (v, " | ".join(res)), flush=True)
    gm.close()
    for k in env:
        del os.environ[k]
