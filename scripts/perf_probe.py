"""Quick performance probe: config-1/2 shaped workload, prints the per-stage stats of the C-ABI call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tools import synth
import downpore_b200 as dp

def main():
    n = int(os.environ.get('N_READS', 100000)); L = int(os.environ.get('READ_LEN', 10000))
    ref_len = int(os.environ.get('REF_LEN', 4_600_000))
    t = time.time(); ref = synth.reference(1, ref_len); print('ref gen', round(time.time() - t, 2))
    t = time.time(); counts = dp.kmer_counts(ref, 11); print('gpu kmer counts', round(time.time() - t, 3))
    t = time.time(); vals = dp.kmer_values(counts, 11); print('values (numpy host)', round(time.time() - t, 2))
    t = time.time(); gm = dp.Mapper(ref, vals, circular=True); print('index build', round(time.time() - t, 3), gm.index_info())
    t = time.time(); gm2 = dp.Mapper(ref, vals, circular=True); print('index build (warm)', round(time.time() - t, 3)); gm2.close()
    t = time.time(); rd = synth.reads(ref, 12, n, L); print('reads gen', round(time.time() - t, 2), 'cores', os.cpu_count())
    offs = np.arange(n + 1, dtype=np.int64) * L
    d = torch.from_numpy(rd).cuda()
    pinned = torch.from_numpy(rd).pin_memory()
    for it in range(3):
        torch.cuda.synchronize(); t = time.time()
        maps, off = gm.map_batch_device(d.data_ptr(), offs)
        torch.cuda.synchronize(); dt = time.time() - t
        st = gm.stats()
        print('device-resident: %.1f ms -> %.2f Gbp/s; mapped reads %d/%d' % (dt * 1e3, n * L / dt / 1e9, int((np.diff(off) > 0).sum()), n))
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
    for it in range(2):
        t = time.time(); maps, off = gm.map_batch_ptr(pinned.data_ptr(), offs); dt = time.time() - t
        print('pinned host: %.1f ms -> %.2f Gbp/s' % (dt * 1e3, n * L / dt / 1e9), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in gm.stats().items() if k.startswith('ms_') or k in ('h2d_bytes', 'rounds')})
    t = time.time(); maps, off = gm.map_batch(rd, offs); dt = time.time() - t
    print('pageable host: %.1f ms -> %.2f Gbp/s' % (dt * 1e3, n * L / dt / 1e9), 'h2d ms', round(gm.stats()['ms_h2d'], 1))

if __name__ == '__main__':
    main()
