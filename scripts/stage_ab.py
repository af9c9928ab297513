"""Single-lane stage times (best of 5) of one workload under several environment settings: A/B of kernel routes.
usage: stage_ab.py config2|config3 [n_reads] -- 'ENV=V ENV2=V' 'ENV=V' ...   (each quoted group is one variant)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DP_LANES"] = "1"; os.environ["DP_RAMP"] = "0"
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
cfg = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "--" else 262144
variants = sys.argv[sys.argv.index("--") + 1:] if "--" in sys.argv else [""]
if cfg == "config3":
    ref_len, L, rs, qs, circ = 64_000_000, 20000, 3, 13, False
else:
    ref_len, L, rs, qs, circ = 4_600_000, 10000, 1, 12, True
ref = synth.reference(rs, ref_len)
vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
rd = synth.reads(ref, qs, n, L, circular=circ); offs = np.arange(n + 1, dtype=np.int64) * L
d = torch.from_numpy(rd).cuda()
base = None
for v in variants:
    env = dict(kv.split("=") for kv in v.split()) if v else {}
    os.environ.update(env)
    gm = dp.Mapper(ref, vals, circular=circ)
    best = {}
    for it in range(5):
        maps, off = gm.map_batch_device(d.data_ptr(), offs); st = gm.stats()
        for k in ("ms_pack", "ms_extract", "ms_lookup", "ms_reduce", "ms_chain", "ms_finish", "ms_total"):
            best[k] = min(best.get(k, 1e9), st[k])
    rows = np.stack([maps[f] for f in ("start", "end", "q_offset", "q_inset", "rc", "ids")], axis=1)
    same = None if base is None else bool(np.array_equal(base[0], rows) and np.array_equal(base[1], off))
    if base is None:
        base = (rows.copy(), off.copy())
    print(cfg, n, "[%s]" % v, {k: round(x, 3) for k, x in best.items()}, "same_as_first", same, flush=True)
    gm.close()
    for k in env:
        del os.environ[k]
