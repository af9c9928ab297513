"""Config-2 shaped workload through dp_mapper_map_batch on pinned host reads under env settings given as
KEY=V,KEY=V arguments (one run each); prints the mean of 4 steps after 2 warm-ups and the stage sums."""
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
    pinned = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    synth.reads(ref, 12, n, L, out=pinned.numpy())
    d = pinned.cuda() if os.environ.get('WITH_DEVICE') else None
    offs = np.arange(n + 1, dtype=np.int64) * L
    keys = ('ms_pack', 'ms_extract', 'ms_lookup', 'ms_chain', 'ms_reduce', 'ms_host_logic', 'rounds')
    for arg in [''] + sys.argv[1:]:
        env = dict(kv.split('=') for kv in arg.split(',')) if arg else {}
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        gm = dp.Mapper(ref, vals, circular=True)
        for mode in (('host', 'device') if d is not None else ('host',)):
            ts = []
            for it in range(6):
                torch.cuda.synchronize(); t = time.time()
                if mode == 'host':
                    gm.map_batch_ptr(pinned.data_ptr(), offs)
                else:
                    gm.map_batch_device(d.data_ptr(), offs)
                torch.cuda.synchronize(); ts.append(time.time() - t)
            st = gm.stats()
            m = sum(ts[2:]) / 4
            print('%-40s %-6s %.1f ms -> %.1f Gbp/s' % (arg or 'default', mode, m * 1e3, n * L / m / 1e9),
                  {k: round(st[k], 1) for k in keys}, flush=True)
        gm.close()
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v


if __name__ == '__main__':
    main()
