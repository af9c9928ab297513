import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
n, L = 1000000, 10000
ref = synth.reference(1, 4_600_000)
vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
rd = synth.reads(ref, 12, n, L, circular=True); offs = np.arange(n + 1, dtype=np.int64) * L
d = torch.from_numpy(rd).cuda()
for v in sys.argv[1:]:
    env = dict(kv.split("=") for kv in v.split()) if v else {}
    os.environ.update(env)
    gm = dp.Mapper(ref, vals, circular=True)
    for it in range(4):
        gm.map_batch_device(d.data_ptr(), offs)
    os.environ["DP_TRACE"] = "1"
    sys.stderr.write("=== trace [%s]\n" % v); sys.stderr.flush()
    gm.map_batch_device(d.data_ptr(), offs)
    del os.environ["DP_TRACE"]
    gm.close()
    for k in env:
        del os.environ[k]
