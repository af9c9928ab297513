"""First-contact GPU check: compares every stage of the CUDA path with the oracle and prints the first differences."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
from oracle import pyoracle as po
import downpore_b200 as dp

def main():
    print(dp.lib().dp_version().decode())
    # pack KAT
    print('pack CGGT ->', hex(int(dp.pack('CGGT')[0])))
    rng = np.random.default_rng(0)
    for L in (5, 16, 17, 63, 64, 65, 1000, 100003):
        s = ''.join('ACGTNacgtn'[c] for c in rng.integers(0, 10, L))
        a = dp.pack(s); b = po.Packed(s).bytes()
        print('pack', L, 'ok' if np.array_equal(a, b) else 'DIFF')
    ref_len = int(os.environ.get('REF_LEN', 300000))
    ref = synth.reference(2, ref_len)
    k = 11
    t = time.time(); c_g = dp.kmer_counts(ref, k); print('gpu counts', time.time() - t)
    c_o = po.kmer_counts(ref, k)
    print('counts equal', np.array_equal(c_g, c_o))
    vals = dp.kmer_values(c_o, k)
    for circular in (True, False):
        t = time.time(); om = po.Mapper(ref, vals, circular=circular); print('oracle mapper', time.time() - t)
        t = time.time(); gm = dp.Mapper(ref, vals, circular=circular); print('gpu mapper', time.time() - t)
        info = gm.index_info(); print(info, 'oracle seeds', om.num_seeds, 'chunks', om.num_chunks)
        so = np.sort(om.seed_kmers()); sg = gm.seed_kmers()
        print('seed sets equal', np.array_equal(so, sg))
        bad = 0
        for c in range(om.num_chunks):
            oc = om.chunk(c); gc = gm.chunk(c)
            seg = oc['segments']; gaps = seg[0::2]; kms = seg[1::2]
            pos = np.cumsum(gaps[:-1]) + k * np.arange(len(kms))
            ok = (oc['offset'], oc['inset'], oc['length'], oc['nseeds']) == (gc['offset'], gc['inset'], gc['length'], gc['nseeds']) \
                and np.array_equal(pos, gc['pos']) and np.array_equal(kms, gc['kmer'])
            if not ok:
                bad += 1
                if bad < 3: print('chunk diff', c, {x: oc[x] for x in ('offset','inset','length','nseeds')}, {x: gc[x] for x in ('offset','inset','length','nseeds')})
        print('chunks differing', bad)
        # reads
        n, rl = 300, 5000
        rd = synth.reads(ref, 7, n, rl, circular=circular)
        reads = [rd[i*rl:(i+1)*rl] for i in range(n)]
        # extra shapes: short, len%4==0, chimeric
        extra = []
        for L in (500, 999, 1000, 1500, 2000, 1996, 2001, 2400, 2999, 3000, 3500, 4000, 4100, 6001, 7000):
            x = synth.reads(ref, 100 + L, 4, L, circular=circular)
            extra += [x[i*L:(i+1)*L] for i in range(4)]
        a = synth.reads(ref, 55, 8, 4000, circular=circular); b = synth.reads(ref, 56, 8, 5000, circular=circular)
        for i in range(8):
            extra.append(np.concatenate([a[i*4000:(i+1)*4000], b[i*5000:(i+1)*5000]]))
        reads += extra
        # window probes
        nb = 0
        for r in reads[:40] + extra:
            L = len(r)
            wins = [(0, L, True)] if L <= 2000 else [(0, 1000, False), (L - 1000, L, False), (1000, 2000, False)]
            for (s, e, whole) in wins:
                g = gm.probe_window(r, s, e, whole)
                for strand in (0, 1):
                    seg, f = om.window_segments(r, s, e, whole, bool(strand))
                    gaps = seg[0::2]; kms = seg[1::2]
                    pos = np.cumsum(gaps[:-1]) + k * np.arange(len(kms))
                    gp, gk = g['seeds'][strand]
                    if not (np.array_equal(pos, gp) and np.array_equal(kms, gk)):
                        nb += 1
                        if nb < 4: print('SEED DIFF', L, s, e, whole, strand, len(pos), len(gp), pos[:5], gp[:5])
                    oc = om.window_candidates(r, s, e, whole, bool(strand))
                    if not np.array_equal(oc, g['candidates'][strand]):
                        nb += 1
                        if nb < 8: print('CAND DIFF', L, s, e, whole, strand, oc, g['candidates'][strand])
                omap = om.window_mappings(r, s, e, whole)
                gmap = g['mappings']
                grows = np.array([[m['start'], m['end'], m['q_offset'], m['q_inset'], m['rc'], m['ids']] for m in gmap], dtype=np.int64).reshape(-1, 6)
                if not np.array_equal(omap, grows):
                    nb += 1
                    if nb < 12: print('MAP DIFF', L, s, e, whole, '\n', omap, '\n', grows)
        print('window probe diffs', nb)
        bases = np.concatenate(reads); offs = np.zeros(len(reads) + 1, dtype=np.int64); offs[1:] = np.cumsum([len(r) for r in reads])
        t = time.time(); orow, ooff, octr = om.map_batch(bases, offs, threads=8); print('oracle map', time.time() - t, octr)
        t = time.time(); gmaps, goff = gm.map_batch(bases, offs); print('gpu map', time.time() - t)
        print(gm.stats())
        grow = np.stack([gmaps['start'], gmaps['end'], gmaps['q_offset'], gmaps['q_inset'], gmaps['rc'], gmaps['ids']], axis=1).astype(np.int64) if len(gmaps) else np.zeros((0, 6), np.int64)
        print('offsets equal', np.array_equal(ooff, goff), 'rows equal', orow.shape == grow.shape and np.array_equal(orow, grow), len(orow), len(grow))
        if not np.array_equal(ooff, goff) or not np.array_equal(orow, grow):
            shown = 0
            for i in range(len(reads)):
                a = orow[ooff[i]:ooff[i+1]]; b = grow[goff[i]:goff[i+1]]
                if not np.array_equal(a, b):
                    print('read', i, 'len', len(reads[i]), '\n oracle', a.tolist(), '\n gpu   ', b.tolist())
                    shown += 1
                    if shown > 5: break
        del gm

if __name__ == '__main__':
    main()
