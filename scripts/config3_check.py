"""BASELINE config 3 shape: 64 Mb linear reference, 20 kb reads; parity on a sample + timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools import synth
from oracle import pyoracle as po
import downpore_b200 as dp
K = 11
ref_len = int(os.environ.get('REF_LEN', 64_000_000)); n = int(os.environ.get('N_READS', 200000)); L = 20000
ns = int(os.environ.get('N_SAMPLE', 3000))
t = time.time(); ref = synth.reference(3, ref_len); print('ref', round(time.time() - t, 1))
t = time.time(); vals = dp.kmer_values(dp.kmer_counts(ref, K), K); print('values', round(time.time() - t, 1))
t = time.time(); gm = dp.Mapper(ref, vals, circular=False); print('gpu index', round(time.time() - t, 2), gm.index_info())
rd = synth.reads(ref, 13, n, L, circular=False); offs = np.arange(n + 1, dtype=np.int64) * L
pinned = torch.from_numpy(rd).pin_memory()
for it in range(3):
    t = time.time(); maps, off = gm.map_batch_ptr(pinned.data_ptr(), offs); dt = time.time() - t
    st = gm.stats()
    print('e2e pinned: %.1f ms %.1f Gbp/s mapped %.3f' % (dt * 1e3, n * L / dt / 1e9, float((np.diff(off) > 0).mean())), {k: round(v, 2) if isinstance(v, float) else v for k, v in st.items()})
if ns:
    t = time.time(); om = po.Mapper(ref, vals, circular=False); print('oracle index', round(time.time() - t, 1))
    t = time.time(); orow, ooff, octr = om.map_batch(rd[:ns * L], offs[:ns + 1], threads=os.cpu_count()); dt = time.time() - t
    print('oracle %d reads %.2f s -> %.3f Gbp/s' % (ns, dt, ns * L / dt / 1e9), octr)
    g = np.stack([maps['start'], maps['end'], maps['q_offset'], maps['q_inset'], maps['rc'], maps['ids']], axis=1).astype(np.int64)[:int(off[ns])]
    print('PARITY', np.array_equal(ooff, off[:ns + 1]) and np.array_equal(orow, g))
