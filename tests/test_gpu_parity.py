"""GPU parity tests (run with -m gpu on a B200): every stage of the CUDA path, called through the C ABI, against the
CPU oracle on the same seeded inputs. Integer work: the bar is bit-exact."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402

import downpore_b200 as dp  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
K = 11


def rows_of(maps):
    if len(maps) == 0:
        return np.zeros((0, 6), dtype=np.int64)
    return np.stack([maps["start"], maps["end"], maps["q_offset"], maps["q_inset"], maps["rc"], maps["ids"]],
                    axis=1).astype(np.int64)


def positions(seg, k=K):
    gaps, kms = seg[0::2], seg[1::2]
    return np.cumsum(gaps[:-1]) + k * np.arange(len(kms)), kms


def mixed_reads(ref, circular, seed=7, n=200, rl=5000):
    rd = synth.reads(ref, seed, n, rl, circular=circular)
    reads = [rd[i * rl:(i + 1) * rl] for i in range(n)]
    for L in (499, 500, 999, 1000, 1500, 1996, 2000, 2001, 2400, 2999, 3000, 3500, 4000, 4100, 5001, 6001, 7000):
        x = synth.reads(ref, 100 + L, 3, L, circular=circular)
        reads += [x[i * L:(i + 1) * L] for i in range(3)]
    a = synth.reads(ref, 55, 8, 4000, circular=circular)
    b = synth.reads(ref, 56, 8, 5000, circular=circular)
    for i in range(8):
        reads.append(np.concatenate([a[i * 4000:(i + 1) * 4000], b[i * 5000:(i + 1) * 5000]]))
    return reads


# ---------------------------------------------------------------------------------------------------------------
def test_pack_kat_and_random():
    """sequence_test.go:211-233 golden vector through the device pack kernel, then random sequences vs the oracle."""
    assert dp.pack("CGGT")[0] == 0x6B
    assert list(dp.pack("CGGT" * 5)) == [0x6B] * 5
    rng = np.random.default_rng(0)
    for L in (1, 3, 4, 5, 15, 16, 17, 31, 63, 64, 65, 70, 1000, 16384, 16385, 100003):
        s = "".join("ACGTNacgtnRY"[c] for c in rng.integers(0, 12, L))
        assert np.array_equal(dp.pack(s), po.Packed(s).bytes()), L


def test_kmer_counts():
    ref = synth.reference(4, 500_000)
    for k in (7, 11, 13):
        assert np.array_equal(dp.kmer_counts(ref, k), po.kmer_counts(ref, k)), k


@pytest.fixture(scope="module", params=[(True, 300_000, 2), (False, 300_000, 2), (True, 1_203_457, 5)])
def pair(request):
    circular, n, seed = request.param
    ref = synth.reference(seed, n)
    vals = dp.kmer_values(dp.kmer_counts(ref, K), K)
    om = po.Mapper(ref, vals, circular=circular)
    gm = dp.Mapper(ref, vals, circular=circular)
    yield ref, circular, om, gm
    gm.close()


def test_index_matches_oracle(pair):
    """AddSingleSeeds seed set (Q4), chunk layout (Q5, Q13) and every chunk's gapped-seed list (Q2)."""
    ref, circular, om, gm = pair
    info = gm.index_info()
    assert info["num_seeds"] == om.num_seeds and info["num_chunks"] == om.num_chunks
    assert np.array_equal(np.sort(om.seed_kmers()), gm.seed_kmers())
    for c in range(om.num_chunks):
        oc, gc = om.chunk(c), gm.chunk(c)
        assert (oc["offset"], oc["inset"], oc["length"], oc["nseeds"]) == (gc["offset"], gc["inset"], gc["length"], gc["nseeds"])
        pos, kms = positions(oc["segments"])
        assert np.array_equal(pos, gc["pos"]) and np.array_equal(kms, gc["kmer"]), c


def test_window_stages(pair):
    """performMapping stage by stage: seeds (both strands), candidate chunks, window mappings."""
    ref, circular, om, gm = pair
    reads = mixed_reads(ref, circular, n=24)
    for r in reads:
        L = len(r)
        wins = [(0, L, True)] if L <= 2000 else [(0, 1000, False), (L - 1000, L, False), (1000, min(2000, L - 1000), False)]
        for (s, e, whole) in wins:
            if e - s < 100:
                continue
            g = gm.probe_window(r, s, e, whole)
            for strand in (0, 1):
                seg, _ = om.window_segments(r, s, e, whole, bool(strand))
                pos, kms = positions(seg)
                assert np.array_equal(pos, g["seeds"][strand][0]), (L, s, e, strand)
                assert np.array_equal(kms, g["seeds"][strand][1]), (L, s, e, strand)
                assert np.array_equal(om.window_candidates(r, s, e, whole, bool(strand)), g["candidates"][strand])
            assert np.array_equal(om.window_mappings(r, s, e, whole), rows_of(g["mappings"])), (L, s, e)


def test_map_batch_matches_oracle(pair):
    ref, circular, om, gm = pair
    reads = mixed_reads(ref, circular)
    bases, offs = make_golden.concat(reads)
    orow, ooff, octr = om.map_batch(bases, offs, threads=4)
    gmaps, goff = gm.map_batch(bases, offs)
    assert np.array_equal(ooff, goff)
    assert np.array_equal(orow, rows_of(gmaps))
    st = gm.stats()
    for key in ("windows", "kmer_lookups", "query_seeds", "posting_runs", "posting_entries", "candidates",
                "chain_cells", "mappings"):
        assert st[key] == octr[key], key
    # PAF lines through the C ABI formatter == AsString restated in python over the oracle rows
    names = ["read%d" % i for i in range(len(reads))]
    lens = [len(r) for r in reads]
    assert gm.paf_lines(gmaps, goff, names, lens) == po.paf_lines(orow, ooff, names, lens, "ref", len(ref), circular)


@pytest.mark.parametrize("name", ["small_circular", "small_linear"])
def test_golden_fixture(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ref, circular, reads = make_golden.case_inputs(name)
    vals = dp.kmer_values(dp.kmer_counts(ref, K), K)
    gm = dp.Mapper(ref, vals, circular=circular)
    info = gm.index_info()
    assert info["num_seeds"] == int(g["num_seeds"]) and info["num_chunks"] == int(g["num_chunks"])
    bases, offs = make_golden.concat(reads)
    gmaps, goff = gm.map_batch(bases, offs)
    assert np.array_equal(goff, g["out_off"])
    assert np.array_equal(rows_of(gmaps), g["rows"])


def test_device_resident_entry_point_and_idempotence(pair):
    import torch
    ref, circular, om, gm = pair
    n, L = 500, 4000
    rd = synth.reads(ref, 77, n, L, circular=circular)
    offs = np.arange(n + 1, dtype=np.int64) * L
    host_maps, host_off = gm.map_batch(rd, offs)
    d = torch.from_numpy(rd).cuda()
    dev_maps, dev_off = gm.map_batch_device(d.data_ptr(), offs)
    assert np.array_equal(host_off, dev_off) and np.array_equal(rows_of(host_maps), rows_of(dev_maps))
    pinned = torch.from_numpy(rd).pin_memory()
    pin_maps, pin_off = gm.map_batch_ptr(pinned.data_ptr(), offs)
    assert np.array_equal(host_off, pin_off) and np.array_equal(rows_of(host_maps), rows_of(pin_maps))
    # batch composition must not matter: mapping the two halves separately gives the same groups
    a_maps, a_off = gm.map_batch(rd[: (n // 2) * L], offs[: n // 2 + 1])
    b_maps, b_off = gm.map_batch(rd[(n // 2) * L:], offs[n // 2:] - offs[n // 2])
    assert np.array_equal(rows_of(host_maps), np.concatenate([rows_of(a_maps), rows_of(b_maps)]))


@pytest.mark.parametrize("env", [{}, {"DP_PULL_CTAS": "3"}, {"DP_PULL_TMA": "0"}])
def test_pinned_host_reads_on_mixed_reads(pair, env):
    """dp_mapper_map_batch on page-locked reads: only the queried windows cross the link, as TMA bulk copies into an HBM
    staging buffer (neighbouring windows of consecutive reads share one copy) or, DP_PULL_TMA=0, as zero-copy loads.
    Ragged, short (whole-read windows, len % 4 == 0) and chimeric reads, shifted by one byte so that no read starts
    on a 16-byte boundary by construction."""
    import torch
    ref, circular, om, gm = pair
    reads = mixed_reads(ref, circular, seed=41, n=333, rl=4999)
    bases = np.concatenate([np.frombuffer(b"A", dtype=np.uint8)] + reads)
    offs = (np.concatenate([[0], np.cumsum([len(r) for r in reads])]) + 1).astype(np.int64)
    orow, ooff, _ = om.map_batch(bases[1:], offs - 1, threads=4)
    pinned = torch.from_numpy(bases).pin_memory()
    with _Env(env):
        maps, off = gm.map_batch_ptr(pinned.data_ptr(), offs)
        st = gm.stats()
    assert np.array_equal(off, ooff) and np.array_equal(rows_of(maps), orow)
    assert 0 < st["h2d_bytes"] < len(bases)  # windows only


def test_packed_reads_entry_point(pair):
    """dp_mapper_map_batch_packed: the reads handed over as the reference's own packedSequence bytes
    (sequence/sequence.go:67-93) map exactly like their ASCII — pageable, page-locked (zero-copy pull of the queried
    bytes) and device-resident buffers; ragged, short, chimeric, empty, tiny, all-N and lowercase reads; a leading pad
    byte so that reads start at every byte alignment; the packed bytes themselves equal the oracle's NewPackedSequence."""
    import torch
    ref, circular, om, gm = pair
    reads = mixed_reads(ref, circular, seed=43, n=150, rl=4999)
    good = synth.reads(ref, 92, 2, 3000, circular=circular)
    reads += [np.zeros(0, dtype=np.uint8), np.frombuffer(b"ACGTACGTACGTACGTACGTAC", dtype=np.uint8),
              np.frombuffer(b"N" * 2500, dtype=np.uint8), np.char.lower(good[:3000].view("S1")).view(np.uint8), good[3000:]]
    bases = np.concatenate(reads)
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    orow, ooff, _ = om.map_batch(bases, offs, threads=4)
    packed, boff, lens = dp.pack_batch(bases, offs)
    for i in (0, 3, len(reads) - 1, len(reads) - 2):
        assert bytes(packed[boff[i]:boff[i + 1]]) == bytes(po.Packed(reads[i].tobytes()).bytes()) or len(reads[i]) == 0
    packed = np.concatenate([np.zeros(1, dtype=np.uint8), packed, np.zeros(16, dtype=np.uint8)])
    boff = boff + 1
    for mode in ("pageable", "pinned", "device"):
        if mode == "pageable":
            maps, off = gm.map_batch_packed(packed, boff, lens)
        elif mode == "pinned":
            t = torch.from_numpy(packed).pin_memory()
            for env in ({"DP_PULL_TMA": "0"}, {"DP_PULL_CTAS": "3"}, {}):  # zero-copy loads; TMA pull (few / default CTAs)
                with _Env(env):
                    maps, off = gm.map_batch_packed(t.data_ptr(), boff, lens)
                assert np.array_equal(off, ooff) and np.array_equal(rows_of(maps), orow), (mode, env)
            assert 0 < gm.stats()["h2d_bytes"] < len(packed)  # windows only
        else:
            t = torch.from_numpy(packed).cuda()
            maps, off = gm.map_batch_packed(t.data_ptr(), boff, lens)
        assert np.array_equal(off, ooff) and np.array_equal(rows_of(maps), orow), mode


def chimera_reads(ref, circular, seed=81, n=120):
    """A read set where Map() rarely returns after its first decision: two- and three-segment chimeras of every size class
    (mapNext's short and long paths, findSplitPoint with one- and two-sided matches), junk inserts (no match at all in the
    middle: both recursive calls), reads with a noisy end, plus ordinary reads."""
    rng = np.random.default_rng(seed)
    out = []
    pieces = {L: synth.reads(ref, seed + L, n, L, circular=circular) for L in (1200, 1700, 2600, 3400, 5200, 8000)}

    def piece(L, i):
        return pieces[L][(i % n) * L:((i % n) + 1) * L]

    sizes = list(pieces)
    for i in range(n):
        a, b = sizes[rng.integers(len(sizes))], sizes[rng.integers(len(sizes))]
        kind = i % 6
        if kind == 0:
            out.append(np.concatenate([piece(a, i), piece(b, i + 7)]))
        elif kind == 1:
            out.append(np.concatenate([piece(a, i), piece(b, i + 3), piece(sizes[rng.integers(len(sizes))], i + 11)]))
        elif kind == 2:  # junk in the middle
            junk = rng.integers(0, 4, size=int(rng.integers(300, 2500))).astype(np.uint8)
            out.append(np.concatenate([piece(a, i), np.frombuffer(b"ACGT", dtype=np.uint8)[junk], piece(b, i + 5)]))
        elif kind == 3:  # noisy front end
            junk = rng.integers(0, 4, size=int(rng.integers(200, 1400))).astype(np.uint8)
            out.append(np.concatenate([np.frombuffer(b"ACGT", dtype=np.uint8)[junk], piece(b, i)]))
        elif kind == 4:  # noisy back end
            junk = rng.integers(0, 4, size=int(rng.integers(200, 1400))).astype(np.uint8)
            out.append(np.concatenate([piece(a, i), np.frombuffer(b"ACGT", dtype=np.uint8)[junk]]))
        else:
            out.append(piece(a, i))
    return out


def test_later_rounds_of_map_on_the_device(pair):
    """Map()'s rounds after the first (mapping/mapping.go:305-383 mapNext, :207-288 findSplitPoint) run on the device for
    the reads the first decision leaves open: a chimera-heavy read set (five reads in six are chimeric or have a junk end)
    maps exactly like the oracle — also from scratch capacities so small that the strategy kernel has to hand the
    sub-batch back for a rerun with more room, and in several small sub-batches."""
    ref, circular, om, gm = pair
    reads = chimera_reads(ref, circular)
    bases, offs = make_golden.concat(reads)
    orow, ooff, octr = om.map_batch(bases, offs, threads=4)
    for env in ({}, {"DP_ROUNDS_HITS": "3", "DP_ROUNDS_LIST": "2"}, {"DP_ROUNDS_CACHE": "1", "DP_SUB_READS": "32"},
                {"DP_SUB_READS": "16", "DP_LANES": "2"}):
        with _Env(env):
            maps, off = gm.map_batch(bases, offs)
        assert np.array_equal(off, ooff), env
        assert np.array_equal(rows_of(maps), orow), env
        st = gm.stats()
        assert st["rounds"] > 2 or env.get("DP_SUB_READS"), st
        if not env:
            for key in ("windows", "kmer_lookups", "query_seeds", "posting_runs", "posting_entries", "candidates",
                        "chain_cells", "mappings"):
                assert st[key] == octr[key], key
            assert st["windows"] > 2.5 * len(reads)  # most reads went past their two round-0 windows


def test_later_rounds_next_to_round0(pair):
    """The later rounds of a sub-batch run on the lane's rounds workspace next to round 0 of its next sub-batch when the
    reads are resident on the device (and, with DP_ROUNDS_DEFER=1, when they are pulled out of page-locked memory): many
    small sub-batches per lane (the sub-batch's tables change hands every time), chimera-heavy reads (every sub-batch has
    open reads), also from scratch capacities so small that the rounds hand the sub-batch back to be mapped again in
    line, and with a result array that outgrows its reservation."""
    import torch
    ref, circular, om, gm = pair
    reads = chimera_reads(ref, circular, seed=83, n=150)
    bases, offs = make_golden.concat(reads)
    orow, ooff, octr = om.map_batch(bases, offs, threads=4)
    dev = torch.from_numpy(bases).cuda()
    pinned = torch.from_numpy(bases).pin_memory()
    for env in ({"DP_SUB_READS": "16"}, {"DP_SUB_READS": "8", "DP_LANES": "2"}, {"DP_SUB_READS": "16", "DP_ROUNDS_HITS": "3"},
                {"DP_SUB_READS": "32", "DP_ROUNDS_CACHE": "1", "DP_ROUNDS_LIST": "2"}, {"DP_SUB_READS": "16", "DP_RESULT_CAP": "30"},
                {"DP_SUB_READS": "16", "DP_ROUNDS_DEFER": "0"}):
        with _Env(env):
            for _ in range(2):
                maps, off = gm.map_batch_device(dev.data_ptr(), offs)
                assert np.array_equal(off, ooff) and np.array_equal(rows_of(maps), orow), ("device", env)
            if env == {"DP_SUB_READS": "16"}:  # the work counters do not depend on where the rounds run. (Against the oracle
                # `windows` can be one short on this read set: findSplitPoint may ask for a window Map() has queried
                # before — the reference maps it again, the window cache answers it.)
                st = gm.stats()
                with _Env({"DP_ROUNDS_DEFER": "0"}):
                    gm.map_batch_device(dev.data_ptr(), offs)
                    st_line = gm.stats()
                keys = ("windows", "kmer_lookups", "query_seeds", "posting_runs", "posting_entries", "candidates",
                        "chain_cells", "mappings", "rounds")
                assert [st[k2] for k2 in keys[:-1]] == [st_line[k2] for k2 in keys[:-1]], (
                    {k2: (st[k2], st_line[k2], octr.get(k2)) for k2 in keys})
        with _Env(dict(env, DP_ROUNDS_DEFER=env.get("DP_ROUNDS_DEFER", "1"))):
            maps, off = gm.map_batch_ptr(pinned.data_ptr(), offs)
            assert np.array_equal(off, ooff) and np.array_equal(rows_of(maps), orow), ("pinned", env)


def test_result_delivery_in_pieces(pair):
    """A batch cut into many sub-batches (six lanes finishing them out of order) is delivered in read order, also when the
    result outgrows the room reserved up front and the rest is placed after the lanes have finished."""
    ref, circular, om, gm = pair
    reads = mixed_reads(ref, circular, seed=71, n=260, rl=3000)
    bases, offs = make_golden.concat(reads)
    orow, ooff, _ = om.map_batch(bases, offs, threads=4)
    for env in ({"DP_SUB_READS": "16"}, {"DP_SUB_READS": "16", "DP_RESULT_CAP": "40"}, {"DP_SUB_READS": "7", "DP_LANES": "3"},
                {"DP_SUB_READS": "64", "DP_RESULT_CAP": "1"}):
        with _Env(env):
            for _ in range(3):
                maps, off = gm.map_batch(bases, offs)
                assert np.array_equal(off, ooff) and np.array_equal(rows_of(maps), orow), env


def test_concurrent_callers_of_one_mapper(pair):
    """The reference runs num_workers goroutines against one Mapper (commands/map.go:84-86): several threads may call
    dp_mapper_map_batch on one mapper at once (each call works on its own lanes) and get what a lone caller gets."""
    import threading
    ref, circular, om, gm = pair
    batches = []
    for t in range(4):
        reads = mixed_reads(ref, circular, seed=60 + t, n=120, rl=4000 + 500 * t)
        bases = np.concatenate(reads)
        offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
        batches.append((bases, offs))
    alone = [gm.map_batch(b, o) for b, o in batches]
    got = [None] * len(batches)
    errs = []

    def call(i):
        try:
            for _ in range(3):
                got[i] = gm.map_batch(*batches[i])
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    th = [threading.Thread(target=call, args=(i,)) for i in range(len(batches))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for (m1, o1), (m2, o2) in zip(alone, got):
        assert np.array_equal(o1, o2) and np.array_equal(rows_of(m1), rows_of(m2))


def test_edge_inputs(pair):
    """Empty batch; empty, tiny (< k + 12 bases: no mapping), all-N and lowercase reads between ordinary ones; the same
    through pageable, pinned and device-resident entry points."""
    import torch
    ref, circular, om, gm = pair
    maps, off = gm.map_batch(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64))
    assert len(maps) == 0 and list(off) == [0]
    good = synth.reads(ref, 91, 6, 3000, circular=circular)
    reads = [good[0:3000], np.zeros(0, dtype=np.uint8), np.frombuffer(b"ACGTACGTACGTACGTACGTAC", dtype=np.uint8),
             np.frombuffer(b"N" * 2500, dtype=np.uint8), np.char.lower(good[3000:6000].view("S1")).view(np.uint8),
             good[6000:9000], np.frombuffer(b"A", dtype=np.uint8), np.zeros(0, dtype=np.uint8), good[9000:12000],
             np.frombuffer(b"acgtn" * 200, dtype=np.uint8), good[12000:18000]]
    bases = np.concatenate(reads)
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    orow, ooff, _ = om.map_batch(bases, offs, threads=2)
    for mode in ("pageable", "pinned", "device"):
        if mode == "pageable":
            maps, off = gm.map_batch(bases, offs)
        elif mode == "pinned":
            t = torch.from_numpy(bases).pin_memory()
            maps, off = gm.map_batch_ptr(t.data_ptr(), offs)
        else:
            t = torch.from_numpy(bases).cuda()
            maps, off = gm.map_batch_device(t.data_ptr(), offs)
        assert np.array_equal(off, ooff) and np.array_equal(rows_of(maps), orow), mode
    assert off[2] == off[1] and off[3] == off[2]  # the empty and the tiny read have no mapping
    assert off[5] > off[4]                        # lowercase maps like uppercase


def test_other_parameters():
    """k, seed_rate, query_size and chunk_size other than the defaults (commands/map.go:19-21)."""
    ref = synth.reference(8, 400_000)
    for (k, rate, edge, chunk) in ((9, 30, 600, 5000), (13, 48, 1000, 10000), (11, 41, 801, 7001)):
        vals = dp.kmer_values(dp.kmer_counts(ref, k), k)
        om = po.Mapper(ref, vals, circular=True, k=k, seed_rate=rate, edge_size=edge, chunk_size=chunk)
        gm = dp.Mapper(ref, vals, circular=True, k=k, seed_rate=rate, edge_size=edge, chunk_size=chunk)
        assert np.array_equal(np.sort(om.seed_kmers()), gm.seed_kmers()), (k, rate)
        rd = synth.reads(ref, 3, 150, 5000)
        offs = np.arange(151, dtype=np.int64) * 5000
        orow, ooff, _ = om.map_batch(rd, offs, threads=4)
        gmaps, goff = gm.map_batch(rd, offs)
        assert np.array_equal(ooff, goff) and np.array_equal(orow, rows_of(gmaps)), (k, rate, edge, chunk)
        gm.close()


def test_config1_full_parity():
    """BASELINE config 1: 4.6 Mb circular reference, 10k x 10 kb reads, byte-identical mapping records."""
    ref = synth.reference(1, 4_600_000)
    vals = dp.kmer_values(dp.kmer_counts(ref, K), K)
    om = po.Mapper(ref, vals, circular=True)
    gm = dp.Mapper(ref, vals, circular=True)
    assert np.array_equal(np.sort(om.seed_kmers()), gm.seed_kmers())
    n, L = 10_000, 10_000
    rd = synth.reads(ref, 11, n, L)
    offs = np.arange(n + 1, dtype=np.int64) * L
    orow, ooff, octr = om.map_batch(rd, offs, threads=os.cpu_count() or 4)
    gmaps, goff = gm.map_batch(rd, offs)
    assert np.array_equal(ooff, goff)
    assert np.array_equal(orow, rows_of(gmaps))
    assert gm.stats()["posting_entries"] == octr["posting_entries"]
    gm.close()


def test_config2_scale_properties():
    """BASELINE config 2's shape at a fifth of its size (200k x 10 kb reads = 2 Gbp, several sub-batches on several
    lanes), checked through properties that need no oracle: every uniquely mapped read lies at the locus and strand it
    was simulated from; the call is idempotent; the pinned-host entry point (TMA pull) and the device-resident one
    return identical records; a 5 000-read slice agrees with the oracle."""
    import torch
    ref = synth.reference(1, 4_600_000)
    vals = dp.kmer_values(dp.kmer_counts(ref, K), K)
    gm = dp.Mapper(ref, vals, circular=True)
    n, L = 200_000, 10_000
    pinned = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    _, truth = synth.reads(ref, 12, n, L, out=pinned.numpy(), with_truth=True)
    offs = np.arange(n + 1, dtype=np.int64) * L
    maps, off = gm.map_batch_ptr(pinned.data_ptr(), offs)
    rows = rows_of(maps)
    maps2, off2 = gm.map_batch_ptr(pinned.data_ptr(), offs)
    assert np.array_equal(off, off2) and np.array_equal(rows, rows_of(maps2))
    d = pinned.cuda()
    maps3, off3 = gm.map_batch_device(d.data_ptr(), offs)
    assert np.array_equal(off, off3) and np.array_equal(rows, rows_of(maps3))
    per_read = np.diff(off)
    assert (per_read > 0).mean() > 0.999
    uniq = np.nonzero(per_read == 1)[0]
    r = rows[off[uniq]]
    lead = np.where(r[:, 4] != 0, r[:, 3], r[:, 2])            # unmapped query bases before the reference start
    delta = (r[:, 0] - (truth[uniq, 0] + lead)) % len(ref)
    delta = np.minimum(delta, len(ref) - delta)
    assert (r[:, 4] == truth[uniq, 1]).all()
    near = delta < 200 + 0.15 * lead
    assert near.mean() > 0.999, (float(near.mean()), int(delta[~near].max()))  # (reads across the origin may differ)
    om = po.Mapper(ref, vals, circular=True)
    ns = 5000
    orow, ooff, _ = om.map_batch(pinned.numpy()[: ns * L], offs[: ns + 1], threads=os.cpu_count() or 4)
    assert np.array_equal(ooff, off[: ns + 1]) and np.array_equal(orow, rows[: off[ns]])
    gm.close()


class _Env:
    def __init__(self, env):
        self.env = env

    def __enter__(self):
        self.old = {k2: os.environ.get(k2) for k2 in self.env}
        os.environ.update(self.env)

    def __exit__(self, *a):
        for k2, v in self.old.items():
            if v is None:
                os.environ.pop(k2, None)
            else:
                os.environ[k2] = v


@pytest.fixture(scope="module")
def large_ref():
    """16.4 Mb linear reference, 20 kb reads (BASELINE config 3 in miniature): ~70 seeds per window strand, so the
    level clamp (Q11) and the level-16 under-count (Q6) are live; 1657 chunks."""
    ref = synth.reference(3, 16_400_000)
    vals = dp.kmer_values(dp.kmer_counts(ref, K), K)
    om = po.Mapper(ref, vals, circular=False)
    n, L = 1500, 20_000
    rd = synth.reads(ref, 13, n, L, circular=False)
    offs = np.arange(n + 1, dtype=np.int64) * L
    orow, ooff, octr = om.map_batch(rd, offs, threads=os.cpu_count() or 4)
    return ref, vals, om, rd, offs, orow, ooff, octr


@pytest.mark.parametrize("env", [
    {},                                                        # lean register kernel + the general warp kernel for what it defers
    {"DP_LOOKUP_MID": "1"},                                    # warp per window strand, 16-byte item loads (the config-3 route)
    {"DP_LOOKUP_MID": "1", "DP_CAP_CANDS": "1"},               # ... with the candidate-capacity retry
    {"DP_LOOKUP_SMALL": "0"},                                  # general warp kernel only, shared-memory counters
    {"DP_LOOKUP_SMEM_CHUNKS": "100"},                          # warp kernel, counters in global memory
    {"DP_LOOKUP_BLOCK": "1"},                                  # CTA per window strand, one counter per chunk, 32-posting items
    {"DP_LOOKUP_BLOCK": "1", "DP_LOOKUP_SEG": "128"},          # 128-posting items (indexes with long runs)
    {"DP_LOOKUP_BLOCK": "1", "DP_LOOKUP_SEG": "128", "DP_LOOKUP_GSHIFT": "3"},  # a counter per 8 chunks + exact recount
    {"DP_LOOKUP_BLOCK": "1", "DP_LOOKUP_GSHIFT": "5", "DP_LOOKUP_GLIST": "2", "DP_LOOKUP_GBATCH": "3"},  # spilled group list
])
def test_lookup_kernels_on_a_large_reference(large_ref, env):
    """Every route through the index lookup gives the oracle's records and counters."""
    ref, vals, om, rd, offs, orow, ooff, octr = large_ref
    with _Env(env):
        gm = dp.Mapper(ref, vals, circular=False)
        assert gm.index_info()["num_chunks"] == om.num_chunks > 1536
        if not env:
            assert np.array_equal(np.sort(om.seed_kmers()), gm.seed_kmers())
        gmaps, goff = gm.map_batch(rd, offs)
        st = gm.stats()
        gm.close()
    assert np.array_equal(ooff, goff)
    assert np.array_equal(orow, rows_of(gmaps))
    for key in ("windows", "posting_runs", "posting_entries", "candidates", "chain_cells"):
        assert st[key] == octr[key], key


@pytest.mark.parametrize("env", [{"DP_LOOKUP_MID": "1"}, {"DP_LOOKUP_BLOCK": "1"}, {"DP_LOOKUP_BLOCK": "1", "DP_LOOKUP_GSHIFT": "1", "DP_LOOKUP_SEG": "128"},
                                 {"DP_LOOKUP_BLOCK": "1", "DP_LOOKUP_GSHIFT": "3", "DP_LOOKUP_GLIST": "1", "DP_LOOKUP_GBATCH": "1"}])
def test_block_lookup_on_mixed_reads(env):
    """The CTA-per-window-strand lookup on a small circular reference with short, chimeric and whole-read windows
    (few seeds per strand: every threshold level below 13 occurs), also with coarse counters (many groups pass the
    low thresholds: the group list spills and is recounted one group at a time)."""
    ref = synth.reference(12, 500_000)
    vals = dp.kmer_values(dp.kmer_counts(ref, K), K)
    om = po.Mapper(ref, vals, circular=True)
    reads = mixed_reads(ref, True, seed=29, n=300, rl=7000)
    bases = np.concatenate(reads)
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    orow, ooff, octr = om.map_batch(bases, offs, threads=4)
    with _Env({"DP_LOOKUP_SMALL": "0"}):
        gm = dp.Mapper(ref, vals, circular=True)
        maps0, off0 = gm.map_batch(bases, offs)
        st0 = gm.stats()  # general warp-per-window-strand route only
        gm.close()
    assert np.array_equal(off0, ooff) and np.array_equal(rows_of(maps0), orow)
    with _Env(env):
        gm = dp.Mapper(ref, vals, circular=True)
        maps, off = gm.map_batch(bases, offs)
        st = gm.stats()
        gm.close()
    assert np.array_equal(off, ooff) and np.array_equal(rows_of(maps), orow)
    # (the oracle's counters can be larger here: Map() re-queries a window for a read across the circular join, the
    # host replay serves the repeat from its cache)
    for key in ("posting_runs", "posting_entries", "candidates"):
        assert st[key] == st0[key] <= octr[key], key
    assert st["kernel_launches"] > st0["kernel_launches"]  # the CTA route launches the deferral kernel as well


def test_index_image_round_trip(tmp_path):
    """dp_mapper_index_export / dp_mapper_create_from_index: a mapper opened from an image (host memory, a file, or a
    device buffer — what the NCCL broadcast delivers) maps exactly like the mapper that was built from the reference."""
    import torch
    ref = synth.reference(8, 400_000)
    vals = dp.kmer_values(dp.kmer_counts(ref, K), K)
    gm = dp.Mapper(ref, vals, circular=True)
    reads = mixed_reads(ref, True, seed=17, n=150, rl=6000)
    bases = np.concatenate(reads)
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    want_maps, want_off = gm.map_batch(bases, offs)
    n = gm.index_image_size()
    assert n > 1 << 20
    # (a) host image
    host = np.empty(n, dtype=np.uint8)
    gm.export_index(host.ctypes.data, n)
    assert bytes(host[:8]) == b"DPB200IX"
    m2 = dp.Mapper.from_index(host.ctypes.data, n)
    # (b) file
    path = os.path.join(str(tmp_path), "ref.dpix")
    gm.save_index(path)
    m3 = dp.Mapper.load_index(path)
    # (c) device buffer
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    gm.export_index(dev.data_ptr(), n)
    m4 = dp.Mapper.from_index(dev.data_ptr(), n)
    del dev
    for m in (m2, m3, m4):
        assert m.index_info()["num_seeds"] == gm.index_info()["num_seeds"]
        assert (m.k, m.circular, m.ref_len, m.edge_size) == (gm.k, gm.circular, gm.ref_len, gm.edge_size)
        maps, off = m.map_batch(bases, offs)
        assert np.array_equal(off, want_off) and np.array_equal(rows_of(maps), rows_of(want_maps))
        c = m.chunk(3)
        assert np.array_equal(c["pos"], gm.chunk(3)["pos"]) and np.array_equal(c["kmer"], gm.chunk(3)["kmer"])
        m.close()
    # a truncated or foreign image is refused
    with pytest.raises(dp.DownporeError):
        dp.Mapper.from_index(host.ctypes.data, n // 2)
    bad = host.copy()
    bad[:8] = 0
    with pytest.raises(dp.DownporeError):
        dp.Mapper.from_index(bad.ctypes.data, n)
    # header fields that do not fit the layout or the parameter ranges, and a payload whose offsets do not end where the
    # header says, are refused too (the image is also the on-disk index format)
    hdr = np.frombuffer(host[:64].tobytes(), dtype=np.int32).copy()
    for word, value in ((3, 3), (5, 2), (6, 7), (8, 40), (12, int(hdr[12]) + 1), (13, int(hdr[13]) + 1)):  # k, seedRate, edge, filterBits, numSeeds, numChunks
        bad = host.copy()
        bad[:64].view(np.int32)[word] = value
        with pytest.raises(dp.DownporeError):
            dp.Mapper.from_index(bad.ctypes.data, n)
    bad = host.copy()
    off_seedoff = int(np.frombuffer(host[:512].tobytes(), dtype=np.uint64)[88 // 8 + 2])  # DpImageHeader.off[IX_SEEDOFF]
    n_seeds = int(hdr[12])
    bad[off_seedoff + 4 * n_seeds: off_seedoff + 4 * n_seeds + 4].view(np.uint32)[0] += 1
    with pytest.raises(dp.DownporeError):
        dp.Mapper.from_index(bad.ctypes.data, n)
    gm.close()


@pytest.mark.parametrize("env", [{"DP_CHAIN_FAST": "0"}, {}, {"DP_FAST_POOL_WORDS": "3000"}, {"DP_FAST_POOL_WORDS": "1"}])
def test_chain_paths_agree_with_oracle(env):
    """The chaining stage has a fast path (dp_reduce_kernel + dp_chain_thread_kernel) and an exact general kernel that
    takes over whatever the fast path hands back. All routes — general only, fast, fast with a starved list pool (most
    windows handed back), fast with no pool at all — must give the oracle's records."""
    ref = synth.reference(12, 500_000)
    vals = dp.kmer_values(dp.kmer_counts(ref, K), K)
    om = po.Mapper(ref, vals, circular=True)
    reads = mixed_reads(ref, True, seed=23, n=400, rl=7000)
    bases = np.concatenate(reads)
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    orow, ooff, _ = om.map_batch(bases, offs, threads=4)
    old = {k2: os.environ.get(k2) for k2 in ("DP_CHAIN_FAST", "DP_FAST_POOL_WORDS")}
    try:
        os.environ.update(env)
        gm = dp.Mapper(ref, vals, circular=True)
        maps, off = gm.map_batch(bases, offs)
        st = gm.stats()
        gm.close()
    finally:
        for k2, v in old.items():
            if v is None:
                os.environ.pop(k2, None)
            else:
                os.environ[k2] = v
    assert np.array_equal(off, ooff) and np.array_equal(rows_of(maps), orow)
    assert st["mappings"] == len(orow)
