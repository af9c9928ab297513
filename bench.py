#!/usr/bin/env python
"""bench.py — `downpore map` hot path on B200 (BASELINE.json metric: mapped Gbp/s, device-timed, + roofline).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU port of the reference (oracle)

A "step" = one pass of the hot path (pack -> seed extract -> index lookup -> chain -> Map() pairing) over one batch of
synthetic reads. Workload at every N: BASELINE config 2 per GPU — a 4.6 Mb uniform circular reference and
`--reads` (default 1M) simulated 10 kb ONT-like reads (4 % sub / 3 % ins / 3 % del) PER GPU (weak scaling: reads are
sharded, the index is rebuilt identically on every GPU, no data-path collective).
  value : whole-job Gbp/s (all submitted bases of all ranks / max-over-ranks time) with the ASCII reads already
          resident in HBM (dp_mapper_map_batch_device);
  e2e   : the same through dp_mapper_map_batch with pinned HOST buffers, H2D copies and the D2H of the mapping
          records inside the timed region.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` captures:
# profiles/r2ad_ncu_summary.txt (one 65536-read sub-batch of config 2 on a single lane, the code of this round; lookup =
# dp_lookup_small_kernel 33.2 MB + the deferral pass of dp_lookup_kernel 8.1 MB; chain = dp_chain_thread_kernel; finish = its
# count and write passes) and profiles/r1h_ncu_lookup_block_3.1Gb.txt (80000 window strands against the 3.1 Gb reference of
# config 4)
TRAFFIC_PER_LAUNCH = {"pack": 169.8e6, "extract": 50.0e6, "lookup": 41.3e6, "reduce": 93.9e6, "chain": 80.7e6,
                      "finish": 11.9e6, "lookup_block": 56.87e9}
# smsp__issue_active.avg.pct_of_peak_sustained_active of the same captures (config 2 is L2-resident: SURVEY 8d asks for
# issue utilisation beside the HBM fraction there); lookup = dp_lookup_small_kernel
ISSUE_ACTIVE_PCT = {"pack": 56.3, "extract": 68.5, "lookup": 68.0, "reduce": 54.0, "chain": 38.0, "finish": 15.6,
                    "lookup_block": 48.2}

K = 11
EDGE = 1000
# BASELINE.json configs (SURVEY 8d seeds). `reads` = reads per GPU per step (weak scaling: reads are sharded).
WORKLOADS = {
    "config2": dict(name="BASELINE config 2", ref_len=4_600_000, read_len=10_000, ref_seed=1, read_seed=12, circular=True,
                    reads=1_000_000),
    "config3": dict(name="BASELINE config 3", ref_len=64_000_000, read_len=20_000, ref_seed=3, read_seed=13, circular=False,
                    reads=250_000),
    "config4": dict(name="BASELINE config 4", ref_len=3_100_000_000, read_len=15_000, ref_seed=4, read_seed=14,
                    circular=False, reads=500_000),
}
REF_LEN, READ_LEN, REF_SEED, READ_SEED, CIRCULAR, WORKLOAD = 4_600_000, 10_000, 1, 12, True, "BASELINE config 2"


def select_workload(name):
    """--workload config2 (default: the configuration the metric is quoted on) | config3 (64 Mb linear reference,
    20 kb reads). Same contract and JSON line either way; the default run also measures config 3 (and config 4 on eight
    GPUs) and reports them in `configs`."""
    global REF_LEN, READ_LEN, REF_SEED, READ_SEED, CIRCULAR, WORKLOAD
    w = WORKLOADS[name]
    REF_LEN, READ_LEN, REF_SEED, READ_SEED, CIRCULAR, WORKLOAD = (w["ref_len"], w["read_len"], w["ref_seed"], w["read_seed"],
                                                                  w["circular"], w["name"])


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_T0 = time.time()


def progress(msg):
    """Phase marks on stderr (seconds since start): where a run spent its time, should a box be slow."""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write("[bench %7.1f s] %s\n" % (time.time() - _T0, msg))
        sys.stderr.flush()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shard_bounds(n_total, rank, world):
    """Contiguous shard [lo, hi) of n_total read indices for `rank` (reads are independent)."""
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi


def config_dict(args, world):
    return {"workload": "%s: synthetic %.1f Mb %s reference, %d simulated %d kb ONT-like reads "
                        "(4%% sub, 3%% ins, 3%% del) per GPU" % (WORKLOAD, REF_LEN / 1e6, "circular" if CIRCULAR else "linear",
                                                               args.reads, READ_LEN // 1000),
            "reads_per_gpu": args.reads, "read_len": READ_LEN, "ref_len": REF_LEN, "k": K, "seed_rate": 40,
            "query_size": EDGE, "chunk_size": 10000, "circular": CIRCULAR,
            "parallelism": "reads sharded over %d GPU(s), index replicated (%s), no data-path collective" % (
                world, "built on rank 0, one NCCL broadcast of its image" if world > 1 and args.index == "broadcast"
                else "built per GPU"),
            "l2_policy": "inputs larger than L2 (%.1f GB ASCII per GPU per step; every step streams all of it)" % (
                args.reads * READ_LEN / 1e9)}


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the CPU port of the reference's map path (oracle/), all host threads, bounded sample.
    Under torchrun only rank 0 works; the other ranks exit without joining a process group."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    from tools import synth
    cores = os.cpu_count() or 1
    ref = synth.reference(REF_SEED, REF_LEN)
    vals = po.kmer_values(ref, K)
    om = po.Mapper(ref, vals, circular=CIRCULAR)
    n = args.ref_sample
    rd = synth.reads(ref, READ_SEED, n, READ_LEN, circular=CIRCULAR)
    offs = np.arange(n + 1, dtype=np.int64) * READ_LEN
    for _ in range(min(args.warmup, 1)):
        om.map_batch(rd[: 2000 * READ_LEN], offs[:2001], threads=cores)
    t0 = time.time()
    for _ in range(args.steps):
        om.map_batch(rd, offs, threads=cores)
    dt = (time.time() - t0) / args.steps
    gbps = n * READ_LEN / dt / 1e9
    sample = ("%d of the workload's %d reads timed per step (read set seed %d, indices 0..%d); a rate, so comparable "
              "with the GPU arm's whole workload" % (n, args.reads, READ_SEED, n - 1))
    line = {"impl": "reference", "metric": "mapped Gbp/s", "value": gbps, "unit": "Gbp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": config_dict(args, max(world, 1)), "sample_reads_per_step": n,
            "cpu_baseline": {"value": gbps, "unit": "Gbp/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "C++ restatement of the reference (no Go toolchain in this image), mapping phase only"},
            "e2e": {"value": gbps, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
def hbm_regime(dp, synth, device, peak, peak_src, gather_gbs, ref_len=3_100_000_000, n=20_000, L=15_000):
    """dp_lookup_block_kernel on BASELINE config 4's human-scale reference (index far larger than L2): posting bytes per
    second against the HBM peak."""
    import torch
    ref = synth.reference(4, ref_len)
    t0 = time.time()
    vals = dp.kmer_values(dp.kmer_counts(ref, K, device=device), K)
    gm = dp.Mapper(ref, vals, circular=False, device=device)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    info = gm.index_info()
    pinned = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    synth.reads(ref, 14, n, L, circular=False, out=pinned.numpy())
    del ref
    offs = np.arange(n + 1, dtype=np.int64) * L
    d_reads = pinned.to(torch.device("cuda", device))
    os.environ["DP_LANES"] = "1"   # kernels alone on the device: the CUDA-event brackets are per-kernel times
    os.environ["DP_RAMP"] = "0"
    try:
        for _ in range(3):
            maps, off = gm.map_batch_device(d_reads.data_ptr(), offs)
        t0 = time.time()
        maps, off = gm.map_batch_device(d_reads.data_ptr(), offs)
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = gm.stats()
    finally:
        del os.environ["DP_LANES"], os.environ["DP_RAMP"]
    gm.close()
    canon = 16.0 * st["posting_runs"] + 8.0 * st["posting_entries"]     # SURVEY 8d accounting (8 B per posting)
    actual = 8.0 * st["posting_runs"] + 4.0 * st["posting_entries"]      # what the kernel reads (4 B per posting)
    ms = st["ms_lookup"]
    return {"workload": "synthetic %.1f Gb linear reference, %d x %d b reads, device-resident, one lane" % (ref_len / 1e9, n, L),
            "kernel": "dp_lookup_block_kernel", "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
            "achieved": canon / (ms * 1e-3) / 1e9, "frac": canon / (ms * 1e-3) / 1e9 / peak,
            "achieved_actual_bytes": actual / (ms * 1e-3) / 1e9, "frac_actual_bytes": actual / (ms * 1e-3) / 1e9 / peak,
            "traffic": TRAFFIC_PER_LAUNCH.get("lookup_block"),
            "ms_lookup": ms, "posting_runs": st["posting_runs"], "posting_entries": st["posting_entries"],
            "window_strands": 2 * st["windows"], "hbm_gather_ceiling_GBs": gather_gbs,
            "Gbp_per_s": n * L / dt / 1e9, "mapped_fraction": float((np.diff(off) > 0).mean()),
            "stage_ms": {k2: st[k2] for k2 in ("ms_pack", "ms_extract", "ms_lookup", "ms_chain")},
            "index": dict(info, build_s=t_build),
            "note": "achieved = SURVEY 8d algorithmic bytes (16 B per included run + 8 B per posting) / CUDA-event time of "
                    "the lookup kernels of one map call; the kernel's posting array holds 4 B per posting, so the bytes "
                    "it really moves are achieved_actual_bytes (ncu dram bytes per launch in `traffic`)"}


def bind_to_gpu_numa_node(local):
    """The e2e path pulls the queried windows over PCIe out of this rank's pinned buffer: keep the rank's threads (and so
    the pages they first touch) on the NUMA node its GPU hangs off. Returns (node, original affinity) or (None, None)."""
    if os.environ.get("DP_BENCH_NUMA", "1") == "0":
        return None, None
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(path).read().strip())
        if node < 0:
            return None, None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        before = os.sched_getaffinity(0)
        mine = before & cpus
        if len(mine) < 2 or mine == before:
            return (node if mine == before else None), None
        os.sched_setaffinity(0, mine)
        return node, before
    except Exception:
        return None, None


def measure(dp, synth, torch, args, rank, world, local, dev, steps, warmup, want_iso=True):
    """One workload (the module-level REF_LEN/READ_LEN/... selected by select_workload) on this rank's GPU: builds or
    receives the index, maps `args.reads` reads per step device-resident (`value`) and from pinned host memory (`e2e`),
    and times every kernel family on a single lane. Returns a dict of rank-local and reduced results."""
    progress("workload %s: reference" % WORKLOAD)
    ref = synth.reference(REF_SEED, REF_LEN)
    t_c0 = time.time()
    t_counts = t_index = t_repl = 0.0
    if world > 1 and args.index == "broadcast":
        # the index is built once (rank 0) and replicated by ONE NCCL broadcast of its image over NVLink
        gm = None
        if rank == 0:
            vals = dp.kmer_values(dp.kmer_counts(ref, K, device=local), K)
            t_counts = time.time() - t_c0
            t1 = time.time()
            gm = dp.Mapper(ref, vals, circular=CIRCULAR, device=local)
            torch.cuda.synchronize()
            t_index = time.time() - t1
        barrier(world)
        t2 = time.time()
        gm = dp.replicate_index(gm, src=0, device=local)
        torch.cuda.synchronize()
        barrier(world)
        t_repl = time.time() - t2
    else:
        vals = dp.kmer_values(dp.kmer_counts(ref, K, device=local), K)
        t_counts = time.time() - t_c0
        t1 = time.time()
        gm = dp.Mapper(ref, vals, circular=CIRCULAR, device=local)
        torch.cuda.synchronize()
        t_index = time.time() - t1
    info = gm.index_info()
    progress("index built; generating reads")

    n = args.reads
    first = rank * n  # weak scaling: every rank maps its own n reads of the same read set
    pinned = torch.empty(n * READ_LEN, dtype=torch.uint8).pin_memory()
    host = pinned.numpy()
    synth.reads(ref, READ_SEED, n, READ_LEN, circular=CIRCULAR, first_index=first, out=host)
    offs = np.arange(n + 1, dtype=np.int64) * READ_LEN
    d_reads = pinned.to(dev, non_blocking=False)
    torch.cuda.synchronize()
    bases_rank = n * READ_LEN

    progress("reads on the device; timing")

    def step_device():
        return gm.map_batch_device(d_reads.data_ptr(), offs)

    def step_host():
        return gm.map_batch_ptr(pinned.data_ptr(), offs)

    # ---- device-resident: `value` ----
    # (nvidia-smi is started in front of the warm-up steps: its start-up — NVML initialisation over every GPU of the box —
    # disturbs running work for a few hundred ms, which would cover the whole timed region; it then samples every 100 ms
    # through the warm-up and the timed steps)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        step_device()
    agg = None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.time()
    ev0.record()
    step_walls = []
    for _ in range(steps):
        ts = time.time()
        maps, off = step_device()   # returns after the library has synchronised its streams (results are on the host)
        step_walls.append(round((time.time() - ts) * 1e3, 2))
        st = gm.stats()
        if agg is None:
            agg = {k: 0 for k in st}
        for k2, v in st.items():
            agg[k2] += v
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    wall_dev = time.time() - t0
    clocks = sampler.stop() if rank == 0 else None
    # device clock (CUDA events bracketing the K steps) — the wall clock is reported beside it
    dt_dev = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, dev)
    wall_dev = max_over_ranks(wall_dev, world, dev)
    mapped_frac = float((np.diff(off) > 0).mean())
    d2h_bytes = int(len(maps) * 32 + (n + 1) * 8)

    progress("value timed; e2e (ASCII entry)")
    # ---- end to end from pinned host memory: `e2e` ----
    for _ in range(min(warmup, 2)):
        step_host()
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.time()
    ev0.record()
    for _ in range(steps):
        step_host()
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    wall_e2e = max_over_ranks(time.time() - t0, world, dev)
    dt_e2e = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, dev)
    st_e2e = gm.stats()

    # ---- end to end with the reads in the form the reference's reader hands to Mapper.Map: packedSequence bytes
    #      (sequence/seqio.go:158,219; 4 bases per byte) in pinned host memory -> dp_mapper_map_batch_packed. The packing
    #      is input preparation (done here on the GPU with torch, outside the timed region), as the ASCII batch is. ----
    progress("e2e (packed entry)")
    assert READ_LEN % 4 == 0
    pk_pinned = torch.empty(n * READ_LEN // 4, dtype=torch.uint8).pin_memory()
    piece = 1 << 28
    for o in range(0, n * READ_LEN, piece):
        b = d_reads[o: o + piece]
        c = (((b >> 1) ^ ((b & 4) >> 2)) & 3).view(-1, 4)
        pk_pinned[o // 4: (o + b.numel()) // 4].copy_((c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3])
        del b, c
    torch.cuda.synchronize()
    pk_off = np.arange(n + 1, dtype=np.int64) * (READ_LEN // 4)
    pk_len = np.full(n, READ_LEN, dtype=np.int64)

    def step_packed():
        return gm.map_batch_packed(pk_pinned.data_ptr(), pk_off, pk_len)

    for _ in range(min(warmup, 2)):
        pmaps, poff = step_packed()
    packed_same = bool(np.array_equal(poff, off) and pmaps.tobytes() == maps.tobytes())
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.time()
    ev0.record()
    for _ in range(steps):
        step_packed()
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    wall_pk = max_over_ranks(time.time() - t0, world, dev)
    dt_pk = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, dev)
    st_pk = gm.stats()
    del pk_pinned

    # ---- per-kernel durations for the roofline: extra untimed-for-`value` steps on a single lane, so that the
    #      CUDA-event brackets on the launching stream see each kernel alone (with several lanes the brackets of
    #      concurrently running kernels overlap and each reads long) ----
    iso = None
    if want_iso:
        progress("single-lane stage times")
        os.environ["DP_LANES"] = "1"
        os.environ["DP_ROUNDS_DEFER"] = "0"  # (the later rounds in line: on their own stream they would overlap the brackets)
        step_device()
        step_device()
        iso = gm.stats()
        del os.environ["DP_LANES"]
        del os.environ["DP_ROUNDS_DEFER"]
    progress("workload %s measured" % WORKLOAD)

    total_bases = sum_over_ranks(float(bases_rank), world, dev)
    S = steps
    res = {"value": total_bases * steps / dt_dev / 1e9, "e2e": total_bases * steps / dt_e2e / 1e9,
           "e2e_packed": total_bases * steps / dt_pk / 1e9, "e2e_packed_ms_per_step": dt_pk / steps * 1e3,
           "e2e_packed_wall_ms_per_step": wall_pk / steps * 1e3,
           "e2e_packed_h2d_bytes_per_step": int(st_pk["h2d_bytes"] + 2 * (n + 1) * 8),
           "e2e_packed_host_buffer_bytes_per_step": int(bases_rank // 4), "e2e_packed_equals_ascii": packed_same,
           "ms_per_step": dt_dev / steps * 1e3, "wall_ms_per_step": wall_dev / steps * 1e3, "step_wall_ms": step_walls,
           "e2e_ms_per_step": dt_e2e / steps * 1e3, "e2e_wall_ms_per_step": wall_e2e / steps * 1e3,
           "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] + (n + 1) * 8), "d2h_bytes_per_step": d2h_bytes,
           "host_buffer_bytes_per_step": int(bases_rank), "mapped_fraction": mapped_frac, "bases_per_step": total_bases,
           "clocks": clocks, "gpu_launches": int(agg["kernel_launches"]), "steps": steps, "warmup": warmup,
           "index": dict(info, build_s=t_index, replicate_s=t_repl, kmer_count_and_values_s=t_counts,
                         replication=("one NCCL broadcast of the index image from rank 0" if world > 1 and
                                      args.index == "broadcast" else "built on every rank")),
           "stats_per_step": {k2: (v / S) for k2, v in agg.items()}}
    # ---- roofline of every kernel family (algorithmic bytes: SURVEY.md 8d canonical accounting) ----
    if iso is not None:
        peak, peak_src = measured_peaks()
        ws = 2 * agg["windows"] / S            # window strands per step
        bytes_pack = 1.25 * EDGE * agg["windows"] / S  # only the queried windows are packed
        bytes_extract = ws * ((EDGE + 3) // 4) + 4.0 * agg["kmer_lookups"] / S
        bytes_lookup = 16.0 * agg["posting_runs"] / S + 8.0 * agg["posting_entries"] / S
        bytes_lookup_actual = 16.0 * agg["posting_runs"] / S + 4.0 * agg["posting_entries"] / S  # 16 B gather per run, 4 B per posting
        bytes_chain = 8.0 * agg["chain_cells"] / S
        # (iso["ms_chain"] includes Map()'s first decision, dp_finish_round0_kernel: a kernel of its own, listed as such)
        ms_finish = iso.get("ms_finish", 0.0)
        bytes_finish = 32.0 * agg["mappings"] / S
        kern = {}
        for name, b2, ms in (("pack", bytes_pack, iso["ms_pack"]), ("extract", bytes_extract, iso["ms_extract"]),
                             ("lookup", bytes_lookup, iso["ms_lookup"]), ("reduce", bytes_chain, iso["ms_reduce"]),
                             ("chain", bytes_chain, iso["ms_chain"] - ms_finish), ("finish", bytes_finish, ms_finish)):
            ach = b2 / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            kern[name] = {"ms_per_step": ms, "algorithmic_bytes_per_step": b2, "achieved_GBs": ach, "frac": ach / peak}
        ms_l = iso["ms_lookup"]
        kern["lookup"]["actual_bytes_per_step"] = bytes_lookup_actual
        kern["lookup"]["frac_actual_bytes"] = bytes_lookup_actual / (ms_l * 1e-3) / 1e9 / peak if ms_l > 0 else 0.0
        res["kernels"] = kern
        res["peak"], res["peak_src"] = peak, peak_src
        res["sector_bytes_lookup"] = 32.0 * agg["posting_runs"] / S + 4.0 * agg["posting_entries"] / S
    res["_maps"], res["_off"], res["_host"], res["_offs"], res["_ref"], res["_vals"] = maps, off, host, offs, ref, (
        vals if not (world > 1 and args.index == "broadcast" and rank != 0) else None)
    gm.close()
    del d_reads, pinned
    torch.cuda.empty_cache()
    return res


def cpu_sample(po, res, ns, cores, all_cores=None):
    """The oracle port on the first `ns` reads of the rank's batch: CPU baseline and parity check in one. It gets every
    core the process may use (`all_cores`: the affinity before the rank was bound to its GPU's NUMA node); the binding
    is restored afterwards, so that the next workload's pinned buffers are first touched on the GPU's node again."""
    bound = os.sched_getaffinity(0)
    if all_cores:
        os.sched_setaffinity(0, all_cores)
    try:
        om = po.Mapper(res["_ref"], res["_vals"], circular=CIRCULAR)
        host, offs, maps, off = res["_host"], res["_offs"], res["_maps"], res["_off"]
        t0 = time.time()
        orow, ooff, _ = om.map_batch(host[: ns * READ_LEN], offs[: ns + 1], threads=cores)
        dt = time.time() - t0
    finally:
        if all_cores:
            os.sched_setaffinity(0, bound)
    grow = np.stack([maps["start"], maps["end"], maps["q_offset"], maps["q_inset"], maps["rc"], maps["ids"]],
                    axis=1).astype(np.int64)[: int(off[ns])]
    parity = bool(np.array_equal(ooff, off[: ns + 1]) and np.array_equal(orow, grow))
    return {"value": ns * READ_LEN / dt / 1e9, "unit": "Gbp/s", "cores": cores, "kind": "port",
            "sample": "first %d reads of the step's batch, mapping phase only" % ns, "parity_with_gpu_on_sample": parity}


def e2e_dict(r):
    """The contract's e2e object. `value` = the reference-facing call a Go host makes, dp_mapper_map_batch_packed: its
    reader already holds every read as packedSequence bytes (sequence/seqio.go:158,219), so those are the host buffers;
    the ASCII entry point (dp_mapper_map_batch, one byte per base over the link) is reported beside it."""
    return {"value": r["e2e_packed"], "unit": "Gbp/s", "h2d_bytes_per_step": r["e2e_packed_h2d_bytes_per_step"],
            "d2h_bytes_per_step": r["d2h_bytes_per_step"], "ms_per_step": r["e2e_packed_ms_per_step"],
            "wall_ms_per_step": r["e2e_packed_wall_ms_per_step"],
            "host_buffer_bytes_per_step": r["e2e_packed_host_buffer_bytes_per_step"],
            "entry": "dp_mapper_map_batch_packed: pinned host buffer of the reads as the reference's packedSequence bytes "
                     "(4 bases per byte)",
            "results_equal_ascii_entry": r["e2e_packed_equals_ascii"],
            "ascii_value": r["e2e"], "ascii_ms_per_step": r["e2e_ms_per_step"],
            "ascii_h2d_bytes_per_step": r["h2d_bytes_per_step"],
            "ascii_host_buffer_bytes_per_step": r["host_buffer_bytes_per_step"],
            "note": "per GPU. The reads stay in the caller's pinned buffer; only the queried windows cross PCIe (a TMA pull "
                    "kernel for either entry: cp.async.bulk host -> shared memory -> HBM staging; a quarter of the bytes "
                    "for packed reads), so h2d bytes < host buffer bytes; the mapping records are written in read order "
                    "into HBM and copied as one block per sub-batch into the call's result array. ascii_* = "
                    "dp_mapper_map_batch on pinned ASCII reads (the input form the CPU arm is given)"}


LOOKUP_KERNEL = {"config2": "dp_lookup_small_kernel + dp_lookup_kernel (window strands with more than 32 seeds)",
                 "config3": "dp_lookup_mid_kernel", "config4": "dp_lookup_block_kernel"}


def side_workload(dp, synth, torch, po, args, name, rank, world, local, dev, cores, parity_reads, all_cores=None):
    """A further BASELINE config measured in the same run; returns the summary stored under line['configs'][name]."""
    keep = (REF_LEN, READ_LEN, REF_SEED, READ_SEED, CIRCULAR, WORKLOAD, args.reads)
    select_workload(name)
    args.reads = WORKLOADS[name]["reads"] if args.side_reads <= 0 else args.side_reads
    try:
        steps, warmup = max(1, min(args.steps, 5)), max(3, min(args.warmup, 3))
        r = measure(dp, synth, torch, args, rank, world, local, dev, steps, warmup)
        k = r["kernels"]
        out = {"workload": config_dict(args, world)["workload"], "reads_per_gpu": args.reads, "n_gpus": world,
               "steps": steps, "warmup": warmup, "value": r["value"], "unit": "Gbp/s", "ms_per_step": r["ms_per_step"],
               "e2e": e2e_dict(r),
               "mapped_fraction": r["mapped_fraction"], "index": r["index"], "gpu_launches": r["gpu_launches"],
               "roofline": {"kernel": LOOKUP_KERNEL[name], "bound": "hbm" if name == "config4" else "l2/issue",
                            "achieved": k["lookup"]["achieved_GBs"], "peak": r["peak"], "unit": "GB/s",
                            "frac": k["lookup"]["frac"], "frac_actual_bytes": k["lookup"]["frac_actual_bytes"],
                            "ms_lookup_per_step_single_lane": k["lookup"]["ms_per_step"], "kernels": k},
               "stats_per_step": r["stats_per_step"]}
        if rank == 0 and parity_reads > 0 and not args.no_cpu_baseline and r["_vals"] is not None:
            progress("%s: cpu baseline (oracle index + %d reads)" % (name, min(parity_reads, args.reads)))
            out["cpu_baseline"] = cpu_sample(po, r, min(parity_reads, args.reads), cores, all_cores)
            progress("%s: cpu baseline done" % name)
        return out
    finally:
        (globals()["REF_LEN"], globals()["READ_LEN"], globals()["REF_SEED"], globals()["READ_SEED"], globals()["CIRCULAR"],
         globals()["WORKLOAD"], args.reads) = keep


def overlap_workload(dp, synth, po, args, device, cores, peak):
    """BASELINE config 5 (`downpore overlap`): reads simulated from reference 1 (seed 15), overlap defaults
    (commands/overlap.go:26-27). A step = one ROUND of commands/overlap.go:115-160 up to the seed-match stream
    (PrepareQueries, AddSequences over the whole read set, FindOverlaps) through dp_overlapper_round: the read set stays
    on the device between rounds (the reference's himem cache), the round's ignore flags go in and its hits come out
    every step. Gbp/s counts the bases of the read set, every one of which AddSequences visits every round."""
    n, L = args.overlap_reads, 10_000
    ref = synth.reference(1, 4_600_000)
    rd = synth.reads(ref, 15, n, L, circular=True)
    offs = np.arange(n + 1, dtype=np.int64) * L
    t0 = time.time()
    g = dp.Overlapper(rd, offs, None, device=device)
    create_s = time.time() - t0
    t0 = time.time()
    vals = dp.kmer_values(g.kmer_counts(), 10)
    g.set_values(vals)
    values_s = time.time() - t0
    ignore = np.zeros(n, dtype=np.uint8)
    first, warm, steps = 0, 2, max(3, min(args.steps, 10))
    res, wall, stage = [], [], {}
    for i in range(warm + steps):
        t0 = time.time()
        r = g.round(first_sequence=first, ignore=ignore)
        dt = time.time() - t0
        first = r.next_first_sequence
        if i >= warm:
            res.append(r)
            wall.append(dt)
            for key in ("ms_select", "ms_queries", "ms_scan", "ms_chunk", "ms_index", "ms_lookup", "ms_align", "ms_collect"):
                stage[key] = stage.get(key, 0.0) + getattr(r, key) / steps
    g.close()
    ms = 1e3 * sum(wall) / steps
    dev_ms = sum(stage.values())
    last = res[-1]
    # the scan kernel: packed reads in, (position, seed) pairs + a sentinel per read out
    scan_bytes = n * L / 4 + 8.0 * (last.read_seeds + n)
    out = {"workload": "BASELINE config 5: %d simulated 10 kb ONT-like reads of the 4.6 Mb reference (%.0fx coverage), "
                       "`downpore overlap` defaults (k 10, 15 seeds per query slice, 10000 seeds and 20000 queries per "
                       "round); one step = one round up to the seed-match stream" % (n, n * L / 4.6e6),
           "reads": n, "read_len": L, "n_gpus": 1, "steps": steps, "warmup": warm,
           "value": n * L / (dev_ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": dev_ms,
           "e2e": {"value": n * L / (ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": ms,
                   "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": int(last.num_hits * 24 + last.num_matches * 2),
                   "entry": "dp_overlapper_round: ignore flags in, hit records and MatchA/MatchB lists out; wall clock of "
                            "the call"},
           "rounds_per_s": 1e3 / ms, "queries_per_round": int(last.num_queries), "hits_per_round": int(last.num_hits),
           "chunks": int(last.num_chunks), "read_seeds": int(last.read_seeds), "seed_postings": int(last.seed_postings),
           "posting_entries": int(last.posting_entries), "candidates": int(last.candidates), "pairs": int(last.pairs),
           "stage_ms": stage, "create_s": create_s, "kmer_values_s": values_s, "gpu_launches": int(last.kernel_launches),
           "roofline": {"kernel": "ov_scan_kernel (AddSequences: NewSeedSequence over every read)", "bound": "shared memory / issue",
                        "achieved": scan_bytes / (stage["ms_scan"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": scan_bytes / (stage["ms_scan"] * 1e-3) / 1e9 / peak,
                        "note": "algorithmic bytes = packed reads (1/4 B per base) + 8 B per seed occurrence and per read; "
                                "the kernel's limiter is the seed-flag lookup in shared memory (one per base), see "
                                "profiles/"},
           "note": "rounds of a run are sequential (each round's ignore flags come from the host's consensus step over "
                   "the previous round's hits); a full run over this read set is about reads / 167 rounds"}
    if po is not None and args.overlap_cpu_reads > 0:
        m = min(args.overlap_cpu_reads, n)
        ov = po.overlap_values(rd[: m * L], offs[: m + 1], 10)
        t0 = time.time()
        progress("config5: oracle round on %d reads" % m)
        o = po.OverlapRound(rd[: m * L], offs[: m + 1], ov)
        dt = time.time() - t0
        g2 = dp.Overlapper(rd[: m * L], offs[: m + 1], ov, device=device)
        r2 = g2.round()
        same = (int(r2.num_hits) == o.num_hits and int(r2.num_chunks) == o.num_chunks and all(
            (a[0], a[1], a[2]) == (h["query_id"], h["rc"], h["target"]) and np.array_equal(a[3], h["match_a"])
            and np.array_equal(a[4], h["match_b"]) for a, h in zip((r2.hit(i) for i in range(int(r2.num_hits))), o.hits)))
        g2.close()
        out["cpu_baseline"] = {"value": m * L / dt / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "port",
                               "sample": "one round over the first %d reads (the oracle's round is single-threaded: the "
                                         "canonical num_workers = 1 order)" % m,
                               "ms_per_round": 1e3 * dt, "gpu_ms_per_round_same_sample": r2.ms_total,
                               "parity_with_gpu_on_sample": bool(same)}
    return out


def run_ours(args):
    import torch
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa_node, affinity_before = bind_to_gpu_numa_node(local)
    import downpore_b200 as dp
    from tools import synth
    po = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import pyoracle as po   # CPU baseline / parity sample only (never on the measured path)

    r = measure(dp, synth, torch, args, rank, world, local, dev, args.steps, args.warmup)
    kern, peak, peak_src = r["kernels"], r["peak"], r["peak_src"]
    gather_gbs = None
    if rank == 0:
        try:
            gather_gbs = dp.probe_gather_gbs(8 << 30, device=local)
        except Exception:
            gather_gbs = None
    # The roofline object is about the index-lookup kernels (the gather kernel BASELINE.json's north star names; by the
    # committed ncu launch list, profiles/, the lookup family and dp_extract_kernel have the largest shares, 24 % each);
    # every family is listed in `kernels`.
    dominant = "lookup"
    l2_resident = args.workload in ("config2", "config3")
    roofline = {"kernel": LOOKUP_KERNEL[args.workload], "bound": "hbm", "achieved": kern[dominant]["achieved_GBs"],
                "peak": peak, "unit": "GB/s", "frac": kern[dominant]["frac"],
                "traffic": TRAFFIC_PER_LAUNCH.get(dominant) if args.workload == "config2" else None,
                "peak_source": peak_src,
                "frac_actual_bytes": kern[dominant]["frac_actual_bytes"],
                "regime": ("l2/issue: the index (7.3 MB) and the k-mer table (1 MB) are L2-resident, ncu dram throughput "
                           "< 2 % for every kernel of this workload; the HBM fraction says how far the kernel is from "
                           "being HBM-bound, issue_active_frac_ncu how busy it keeps the SMs" if l2_resident else "hbm"),
                "issue_active_frac_ncu": (ISSUE_ACTIVE_PCT[dominant] / 100.0) if args.workload == "config2" else None,
                "note": "achieved = SURVEY 8d algorithmic bytes of one step (16 B per included run + 8 B per posting) / "
                        "CUDA-event time of the lookup kernels on their launching stream over one step run on a single "
                        "lane (kernels not overlapping; the timed `value` steps run 6 lanes). frac_actual_bytes counts "
                        "the 4 B per posting the index really holds. traffic = ncu dram bytes per launch of a 65536-read "
                        "sub-batch (profiles/). hbm_regime_* = the same kernel family where it IS bound by HBM: the 3.1 Gb "
                        "reference of BASELINE config 4",
                "hbm_gather_ceiling_GBs": gather_gbs,
                "lookup_sector_GBs": (r["sector_bytes_lookup"] / (kern["lookup"]["ms_per_step"] * 1e-3) / 1e9
                                      if kern["lookup"]["ms_per_step"] > 0 else None),
                "issue_active_pct_ncu": ISSUE_ACTIVE_PCT if args.workload == "config2" else None,
                "kernels": kern}
    for name, kv in kern.items():  # flat copies: the driver's parser keeps scalars of this object
        roofline["ms_%s" % name] = kv["ms_per_step"]
        roofline["frac_%s" % name] = kv["frac"]

    n = args.reads
    line = {"metric": "mapped Gbp/s", "value": r["value"], "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "wall_ms_per_step": r["wall_ms_per_step"], "step_wall_ms_rank0": r["step_wall_ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": config_dict(args, world), "host_numa_node_rank0": numa_node,
            "e2e": e2e_dict(r),
            "gpu_launches": r["gpu_launches"],
            "clocks": r["clocks"], "roofline": roofline,
            "mapped_fraction": r["mapped_fraction"], "bases_per_step": r["bases_per_step"],
            "bases_note": "Gbp/s counts ALL submitted bases; Map() touches 2-12 windows of 1000 bases per read",
            "index": r["index"], "stats_per_step": r["stats_per_step"]}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on a bounded sample ----
    cores = 1
    if rank == 0 and not args.no_cpu_baseline:
        cores = len(affinity_before or os.sched_getaffinity(0)) or 1
        if world == 1:
            progress("cpu baseline (oracle) on %d reads" % min(args.cpu_sample, n))
            line["cpu_baseline"] = cpu_sample(po, r, min(args.cpu_sample, n), cores, affinity_before)
            progress("cpu baseline done")
    del r

    # ---- the other BASELINE configs in the same run: config 3 at every N; config 4 (3.1 Gb reference, index built on
    #      rank 0 and replicated by one broadcast, reads sharded) on eight GPUs ----
    configs = {}
    if args.workload == "config2" and not args.no_side_configs:
        try:
            configs["config3"] = side_workload(dp, synth, torch, po, args, "config3", rank, world, local, dev, cores,
                                               2000 if world == 1 else 500, affinity_before)
        except Exception as ex:  # never lose the headline line to an auxiliary measurement
            configs["config3"] = {"error": str(ex)}
        if world >= 8 or args.config4:
            try:
                configs["config4"] = side_workload(dp, synth, torch, po, args, "config4", rank, world, local, dev, cores, 0,
                                                   affinity_before)
            except Exception as ex:
                configs["config4"] = {"error": str(ex)}
    if rank == 0 and world == 1 and args.workload == "config2" and not args.no_side_configs and args.overlap_reads > 0:
        try:
            configs["config5"] = overlap_workload(dp, synth, po, args, local, cores, peak)
        except Exception as ex:
            configs["config5"] = {"error": str(ex)}
    if configs:
        line["configs"] = configs
        for cname, c in configs.items():  # flat copies where the driver's parser keeps them
            if "error" in c:
                roofline["%s_error" % cname] = c["error"]
                continue
            if cname == "config5":
                roofline["config5_value_Gbps"] = c["value"]
                roofline["config5_e2e_Gbps"] = c["e2e"]["value"]
                roofline["config5_ms_per_round"] = c["ms_per_step"]
                roofline["config5_reads"] = c["reads"]
                roofline["config5_scan_frac"] = c["roofline"]["frac"]
                if "cpu_baseline" in c:
                    roofline["config5_cpu_Gbps"] = c["cpu_baseline"]["value"]
                    roofline["config5_parity_on_sample"] = c["cpu_baseline"]["parity_with_gpu_on_sample"]
                continue
            roofline["%s_value_Gbps" % cname] = c["value"]
            roofline["%s_e2e_Gbps" % cname] = c["e2e"]["value"]
            roofline["%s_e2e_ascii_Gbps" % cname] = c["e2e"]["ascii_value"]
            roofline["%s_ms_per_step" % cname] = c["ms_per_step"]
            roofline["%s_reads_per_gpu" % cname] = c["reads_per_gpu"]
            roofline["%s_lookup_ms_single_lane" % cname] = c["roofline"]["ms_lookup_per_step_single_lane"]
            roofline["%s_lookup_frac" % cname] = c["roofline"]["frac"]
            roofline["%s_lookup_frac_actual_bytes" % cname] = c["roofline"]["frac_actual_bytes"]
            roofline["%s_replicate_s" % cname] = c["index"]["replicate_s"]
            roofline["%s_index_bytes" % cname] = c["index"]["index_bytes"]
            if "cpu_baseline" in c:
                roofline["%s_cpu_Gbps" % cname] = c["cpu_baseline"]["value"]
                roofline["%s_parity_on_sample" % cname] = c["cpu_baseline"]["parity_with_gpu_on_sample"]

    # ---- the index-lookup kernel where it IS bound by HBM (rank 0, N=1): the 3.1 Gb reference of BASELINE config 4
    #      (313k chunks, 677M postings, 15 GB index >> L2), 15 kb reads, single lane ----
    if rank == 0 and world == 1 and not args.no_hbm_regime:
        try:
            progress("lookup on the 3.1 Gb reference (hbm regime)")
            h = hbm_regime(dp, synth, local, peak, peak_src, gather_gbs)
            progress("hbm regime done")
            line["roofline_hbm_regime"] = h
            for k2 in ("frac", "frac_actual_bytes", "achieved", "achieved_actual_bytes", "ms_lookup", "Gbp_per_s", "kernel",
                       "workload", "traffic"):
                roofline["hbm_regime_%s" % k2] = h[k2]
        except Exception as ex:  # never lose the headline line to the auxiliary measurement
            line["roofline_hbm_regime"] = {"error": str(ex)}
    if rank == 0:
        emit(line)


def main():
    # Keep stdout clean for the one JSON line: native libraries (NCCL prints its version banner) write to fd 1.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=int(os.environ.get("DP_BENCH_READS", 1_000_000)),
                    help="reads per GPU per step (BASELINE config 2: 1M)")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4"],
                    help="config2 = the configuration the metric is quoted on (default); config3 = 64 Mb linear reference, "
                         "20 kb reads (use --reads 500000: 10 GB of ASCII per GPU per step)")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="reads timed on the CPU oracle for cpu_baseline")
    ap.add_argument("--ref-sample", type=int, default=100_000, help="reads per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true", help="skip BASELINE configs 3 (every N) and 4 (N = 8) after the headline workload")
    ap.add_argument("--config4", action="store_true", help="also run BASELINE config 4 (3.1 Gb reference) below eight GPUs")
    ap.add_argument("--side-reads", type=int, default=0, help="reads per GPU per step of the side configs (default: their own)")
    ap.add_argument("--overlap-reads", type=int, default=int(os.environ.get("DP_BENCH_OVERLAP_READS", 500_000)),
                    help="reads of BASELINE config 5 (`overlap`, N = 1 only; 0 skips it)")
    ap.add_argument("--overlap-cpu-reads", type=int, default=20_000, help="reads of the oracle's overlap round (cpu_baseline)")
    ap.add_argument("--no-hbm-regime", action="store_true", help="skip the lookup measurement on the 3.1 Gb reference (BASELINE config 4)")
    ap.add_argument("--index", default="broadcast", choices=["broadcast", "rebuild"],
                    help="N>1: replicate rank 0's index by one NCCL broadcast (default) or rebuild it on every rank")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.impl != "reference":
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
