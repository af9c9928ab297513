"""BASELINE config 4 shape (human-scale linear reference, 15 kb reads): does the index build, what do the stages cost.
REF_LEN defaults to 1 Gb (3.1e9 for the full config); no oracle at this size (its bitsets need ~10-100 GB)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools import synth
import downpore_b200 as dp
K = int(os.environ.get('K', 11))
ref_len = int(float(os.environ.get('REF_LEN', 1e9))); n = int(os.environ.get('N_READS', 20000)); L = 15000
t = time.time(); ref = synth.reference(4, ref_len); print('ref', round(time.time() - t, 1), flush=True)
t = time.time(); counts = dp.kmer_counts(ref, K); print('counts', round(time.time() - t, 1), flush=True)
t = time.time(); vals = dp.kmer_values(counts, K); print('values', round(time.time() - t, 1), flush=True)
t = time.time(); gm = dp.Mapper(ref, vals, circular=False, k=K); print('gpu index', round(time.time() - t, 2), gm.index_info(), flush=True)
print('gpu mem used GB', round((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9, 1))
rd, truth = synth.reads(ref, 14, n, L, circular=False, with_truth=True); offs = np.arange(n + 1, dtype=np.int64) * L
pinned = torch.from_numpy(rd).pin_memory()
for it in range(3):
    t = time.time(); maps, off = gm.map_batch_ptr(pinned.data_ptr(), offs); dt = time.time() - t
    st = gm.stats()
    print('e2e pinned: %.1f ms %.2f Gbp/s mapped %.3f' % (dt * 1e3, n * L / dt / 1e9, float((np.diff(off) > 0).mean())),
          {k: round(v, 2) if isinstance(v, float) else v for k, v in st.items()}, flush=True)
# plausibility without an oracle: uniquely mapped reads land where they were drawn from
ok = tot = 0
for i in range(n):
    if off[i + 1] - off[i] == 1:
        m = maps[off[i]]
        tot += 1
        lead = int(m['q_inset'] if m['rc'] else m['q_offset'])
        if abs(int(m['start']) - (int(truth[i, 0]) + lead)) < 200 + 0.15 * lead and int(m['rc']) == int(truth[i, 1]):
            ok += 1
print('uniquely mapped %d of %d, at the true locus %d' % (tot, n, ok))
sec = 32.0 * st['posting_runs'] + 4.0 * st['posting_entries']
print('lookup: %.1f ms, posting bytes %.2f GB -> %.1f GB/s' % (st['ms_lookup'], sec / 1e9, sec / 1e9 / (st['ms_lookup'] * 1e-3)))
