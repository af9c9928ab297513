import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
n, L = int(os.environ.get("N_READS", 60000)), 20000
ref = synth.reference(3, 64_000_000)
vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
rd = synth.reads(ref, 13, n, L, circular=False); offs = np.arange(n + 1, dtype=np.int64) * L
d = torch.from_numpy(rd).cuda()
os.environ["DP_HOST_PROFILE"] = "1"
for v in sys.argv[1:] or [""]:
    env = dict(kv.split("=") for kv in v.split()) if v else {}
    os.environ.update(env)
    gm = dp.Mapper(ref, vals, circular=False)
    for it in range(3):
        t = time.time(); gm.map_batch_device(d.data_ptr(), offs); dt = time.time() - t
        st = gm.stats()
        print("[%s] call %d: %.1f ms rounds %d windows %d launches %d retries %d" % (v, it, dt * 1e3, st["rounds"], st["windows"], st["kernel_launches"], st["retries"]), flush=True)
    gm.close()
    for k in env:
        del os.environ[k]
