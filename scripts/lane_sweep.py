"""Config-2 shaped workload; sweeps DP_LANES for the device-resident and the pinned-host entry points."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tools import synth
import downpore_b200 as dp


def main():
    n = int(os.environ.get('N_READS', 1000000)); L = 10000
    ref = synth.reference(1, 4_600_000)
    vals = dp.kmer_values(dp.kmer_counts(ref, 11), 11)
    gm = dp.Mapper(ref, vals, circular=True)
    if os.environ.get('USE_DP_ALLOC'):
        hostbuf = dp.host_alloc(n * L)
        synth.reads(ref, 12, n, L, out=hostbuf)
        pinned = torch.from_numpy(hostbuf)
        d = torch.empty(n * L, dtype=torch.uint8, device='cuda')
        torch.cuda.current_stream().synchronize()
        import ctypes
        cudart = ctypes.CDLL('libcudart.so')
        cudart.cudaMemcpy(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(pinned.data_ptr()), ctypes.c_size_t(n * L), 1)
    else:
        pinned = torch.empty(n * L, dtype=torch.uint8).pin_memory()
        synth.reads(ref, 12, n, L, out=pinned.numpy())
        d = pinned.cuda()
    offs = np.arange(n + 1, dtype=np.int64) * L
    keys = ('ms_total', 'ms_pack', 'ms_extract', 'ms_lookup', 'ms_chain', 'ms_host_logic', 'rounds')
    for lanes in [int(x) for x in os.environ.get('LANES', '2,3,4,6').split(',')]:
        os.environ['DP_LANES'] = str(lanes)
        for mode in ('device', 'host'):
            best = 1e9
            for it in range(4):
                torch.cuda.synchronize(); t = time.time()
                if mode == 'device':
                    maps, off = gm.map_batch_device(d.data_ptr(), offs)
                else:
                    maps, off = gm.map_batch_ptr(pinned.data_ptr(), offs)
                torch.cuda.synchronize(); dt = time.time() - t
                if it:
                    best = min(best, dt)
            st = gm.stats()
            print('lanes %d %-6s best %.1f ms -> %.1f Gbp/s' % (lanes, mode, best * 1e3, n * L / best / 1e9),
                  {k: round(st[k], 1) for k in keys}, flush=True)


if __name__ == '__main__':
    main()
