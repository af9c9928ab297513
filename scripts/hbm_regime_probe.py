"""dp_lookup_block_kernel on a 1 Gb reference (the bench's roofline_hbm_regime) under a few settings; prints ms_lookup."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import downpore_b200 as dp
from tools import synth
peak = 6550.4
for env in [{}] + [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:]]:
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    r = bench.hbm_regime(dp, synth, 0, peak, "probe", None)
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    print(env, "ms_lookup %.3f frac %.3f actual %.1f GB/s Gbp/s %.2f mapped %.3f" % (r["ms_lookup"], r["frac"], r["achieved_actual_bytes"], r["Gbp_per_s"], r["mapped_fraction"]), flush=True)
