"""Device timeline of one dp_mapper_map_batch_packed call on pinned host reads (config-2 shape): DP_TRACE=1 lines on stderr,
one per sub-batch. usage: trace_probe.py [n_reads] ['ENV=V ENV2=V' ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L = 10000
ref = synth.reference(1, 4_600_000)
vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
gm = dp.Mapper(ref, vals, circular=True)
pinned = torch.empty(n * L, dtype=torch.uint8).pin_memory()
synth.reads(ref, 12, n, L, circular=True, out=pinned.numpy()); offs = np.arange(n + 1, dtype=np.int64) * L
pk, boff, lens = dp.pack_batch(pinned.numpy(), offs)
pkp = torch.from_numpy(pk).pin_memory()
d = pinned.cuda()
for it in range(3):
    t = time.time(); gm.map_batch_device(d.data_ptr(), offs); print("device %.2f ms" % ((time.time() - t) * 1e3), flush=True)
for v in [""] + sys.argv[2:]:
    env = dict(kv.split("=") for kv in v.split()) if v else {}
    os.environ.update(env)
    ts = []
    for it in range(6):
        t = time.time(); gm.map_batch_packed(pkp.data_ptr(), boff, lens); ts.append((time.time() - t) * 1e3)
    st = gm.stats()
    print("packed [%s] ms %s  h2d %.0f MB  pack %.1f host %.1f" % (v, " ".join("%.1f" % x for x in ts), st["h2d_bytes"] / 1e6, st["ms_pack"], st["ms_host_logic"]), flush=True)
    if not v:
        os.environ["DP_TRACE"] = "1"; os.environ["DP_HOST_PROFILE"] = "1"
        sys.stderr.write("=== trace [default]\n"); sys.stderr.flush()
        gm.map_batch_packed(pkp.data_ptr(), boff, lens)
        sys.stderr.write("=== trace [device-resident reads]\n"); sys.stderr.flush()
        gm.map_batch_device(d.data_ptr(), offs)
        del os.environ["DP_TRACE"]; del os.environ["DP_HOST_PROFILE"]
    for k in env:
        del os.environ[k]
