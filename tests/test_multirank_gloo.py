"""world_size-2 gloo test (CPU) of bench.py's multi-rank plumbing: read sharding, rank-local read generation,
max/sum reductions. The data path itself has no collective (reads shard embarrassingly)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import bench
    from oracle import pyoracle as po
    from tools import synth
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    n_total, L = 64, 3000
    lo, hi = bench.shard_bounds(n_total, rank, world)
    ref = synth.reference(5, 200_000)
    # every rank generates only its shard; the union must equal the single-process read set
    mine = synth.reads(ref, 9, hi - lo, L, first_index=lo)
    vals = po.kmer_values(ref, 11)
    om = po.Mapper(ref, vals, circular=True)
    rows, off, _ = om.map_batch(mine, np.arange(hi - lo + 1, dtype=np.int64) * L, threads=1)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), reads=mine, rows=rows, off=off, lo=lo, hi=hi)
    t = bench.max_over_ranks(float(rank + 1), world, "cpu")
    s = bench.sum_over_ranks(float(hi - lo), world, "cpu")
    assert t == float(world) and s == float(n_total)
    bench.barrier(world)
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import pyoracle as po
    from tools import synth
    ref = synth.reference(5, 200_000)
    n_total, L = 64, 3000
    full = synth.reads(ref, 9, n_total, L)
    vals = po.kmer_values(ref, 11)
    om = po.Mapper(ref, vals, circular=True)
    rows, off, _ = om.map_batch(full, np.arange(n_total + 1, dtype=np.int64) * L, threads=2)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(p["lo"]) for p in parts] == [0, 32] and [int(p["hi"]) for p in parts] == [32, 64]
    assert np.array_equal(np.concatenate([p["reads"] for p in parts]), full)
    assert np.array_equal(np.concatenate([p["rows"] for p in parts]), rows)  # sharding does not change any record


def test_shard_bounds_cover_everything():
    import bench
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            edges = [bench.shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1


# ---------------------------------------------------------------------------------------------------------------
# replicate_index (the one exchange step of the multi-GPU path): rank 0's index image is broadcast once;
# on the CPU tier the image is a stand-in byte string and Mapper.from_index is intercepted (no GPU here).
class _FakeMapper:
    def __init__(self, blob):
        self.blob = blob

    def index_image_size(self):
        return len(self.blob)

    def export_index(self, ptr, n):
        import ctypes
        assert n == len(self.blob)
        ctypes.memmove(ptr, self.blob, n)


def _replicate_worker(rank, world, port, out_dir):
    import ctypes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import downpore_b200 as dp
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    blob = bytes(np.random.default_rng(3).integers(0, 256, 100_003, dtype=np.uint8))
    got = {}

    def fake_from_index(image_ptr, nbytes, device=0, ref_name="ref"):
        got["bytes"] = ctypes.string_at(image_ptr, nbytes)
        return "opened-from-image"

    dp.Mapper.from_index = staticmethod(fake_from_index)
    src = _FakeMapper(blob) if rank == 0 else None
    m = dp.replicate_index(src, src=0, device=0)
    if rank == 0:
        assert m is src
    else:
        assert m == "opened-from-image" and got["bytes"] == blob
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_replicate_index_broadcasts_the_image(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_replicate_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
