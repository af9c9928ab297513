"""Pinned-host (e2e) step times of one workload with the host profile of the lanes. usage: e2e_probe.py config2|config3 [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
cfg = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
if cfg == "config3":
    ref_len, L, rs, qs, circ = 64_000_000, 20000, 3, 13, False
else:
    ref_len, L, rs, qs, circ = 4_600_000, 10000, 1, 12, True
ref = synth.reference(rs, ref_len)
vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
gm = dp.Mapper(ref, vals, circular=circ)
pinned = torch.empty(n * L, dtype=torch.uint8).pin_memory()
synth.reads(ref, qs, n, L, circular=circ, out=pinned.numpy()); offs = np.arange(n + 1, dtype=np.int64) * L
d = pinned.cuda()
for it in range(3):
    t = time.time(); gm.map_batch_device(d.data_ptr(), offs); print("device %.2f ms" % ((time.time() - t) * 1e3), flush=True)
for it in range(6):
    t = time.time(); gm.map_batch_ptr(pinned.data_ptr(), offs); dt = time.time() - t
    st = gm.stats(); print("pinned %.2f ms  h2d %.0f MB  pack %.1f" % (dt * 1e3, st["h2d_bytes"] / 1e6, st["ms_pack"]), flush=True)
# packed entry (dp_mapper_map_batch_packed) under pull variants given as extra arguments: 'ENV=V ENV2=V' ...
pk, boff, lens = dp.pack_batch(pinned.numpy(), offs)
pkp = torch.from_numpy(pk).pin_memory()
for v in sys.argv[3:] or [""]:
    env = dict(kv.split("=") for kv in v.split()) if v else {}
    os.environ.update(env)
    ts = []
    for it in range(7):
        t = time.time(); gm.map_batch_packed(pkp.data_ptr(), boff, lens); ts.append((time.time() - t) * 1e3)
    st = gm.stats()
    print("packed [%s] ms %s  h2d %.0f MB  pack %.1f" % (v, " ".join("%.1f" % x for x in ts), st["h2d_bytes"] / 1e6, st["ms_pack"]), flush=True)
    for k in env:
        del os.environ[k]
