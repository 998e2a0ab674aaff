#!/usr/bin/env python
"""bench.py — sampled edges/s through the GraphSAGE mini-batch train step (sample -> SAGE fwd -> CE -> bwd
-> [all-reduce] -> Adam) on a synthetic ogbn-products-shaped graph (BASELINE.json metric / configs[3]).

    python bench.py --gpus 1 --steps K --warmup W            # our arm (B200, libngnn_b200.so)
    torchrun ... bench.py --gpus N --steps K --warmup W      # data parallel, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle) on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

PREWARM_STEPS = int(os.environ.get("NGNN_BENCH_PREWARM", "300"))   # untimed, part of setup (~0.2 s)
REF_MAX_STEPS, REF_MAX_WARMUP = 40, 2        # --impl reference: ~3 s per CPU step
METRIC = "sampled_edges_per_sec"
UNIT = "edges/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="products", choices=["products", "arxiv", "pubmed", "cora", "computers"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debug only; default = full shape)")
    ap.add_argument("--law", default="powerlaw", choices=["powerlaw", "uniform"])
    ap.add_argument("--cpu-steps", type=int, default=3, help="steps of the bounded cpu_baseline sample (0 = skip)")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--ncu-range", action="store_true", help="cudaProfilerStart/Stop around the first timed region")
    ap.add_argument("--autograd", action="store_true", help="time the per-kernel autograd variant instead of the fused step")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks during the timed region
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, read through NVML from the timing loop itself every few
    steps (a polling thread costs the host-bound loop a GIL hand-off per sample: measured 5-10 % on `value`)."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self._reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self._reasons_fn(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit and name != "gpu_idle":
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        self.sample()

    def stop(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------- problem setup
def build_problem(args, device):
    from noise_gnn_b200.synthetic import make_dataset
    data, sh, train_idx = make_dataset(args.workload, seed=1232, law=args.law, device=device, noise_type="sym",
                                       noise_rate=0.3, scale=args.scale)
    return data, sh, train_idx


def agg_l1_bytes(touched, n_dst, e, F):
    """ALGORITHMIC bytes of the layer-1 aggregation launch: every distinct table row read once,
    indices once, mean + root outputs written once."""
    rows = torch.unique(touched).numel()
    return 4 * F * rows + 4 * e + 4 * (n_dst + 1) + 4 * n_dst + 2 * 4 * F * n_dst


# --------------------------------------------------------------------------- our arm
def _claim_stdout():
    """Everything libraries print to stdout (the NCCL version banner, warnings) goes to stderr; the returned fd is the
    real stdout, for the ONE JSON line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_stdout_fd: int, obj) -> None:
    sys.stdout.flush()
    os.write(real_stdout_fd, (json.dumps(obj) + "\n").encode())


def run_ours(args):
    import torch.distributed as dist

    real_stdout = _claim_stdout()

    from noise_gnn_b200 import NeighborLoader, SAGE, _lib, ops
    from noise_gnn_b200.train import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a B200: there is no CPU fallback")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner (NCCL_DEBUG >= VERSION) to stdout
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    assert lib.ngnn_device_supported() == 1, "libngnn_b200.so is sm_100a only"
    for kv in filter(None, os.environ.get("NGNN_TUNING", "").split(",")):      # A/B switches, e.g. NGNN_TUNING=7:0
        k, v = kv.split(":")
        _lib.call("ngnn_set_tuning", int(k), int(v))

    t_setup = time.time()
    data, sh, train_idx = build_problem(args, device)
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size,
                            shuffle=True, seed=1232, rank=rank, world_size=world)
    torch.manual_seed(1232)
    model = SAGE(sh.features, sh.hidden, sh.classes, sh.layers, dropout=sh.dropout).to(device)
    model.train()
    trainer = Trainer(model, lr=1e-3, world_size=world)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    K, W = args.steps, args.warmup
    nb = loader.num_batches_global

    def batches(start_epoch=0):
        """Endless stream of batches: epoch after epoch from `start_epoch` (W + K may exceed one epoch, e.g. 49 steps per
        rank at 8 GPUs); the same start epoch replays the same blocks."""
        loader.epoch = start_epoch
        while True:
            for b in loader:
                yield b

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    def sum_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=device)
            dist.all_reduce(t)
            return float(t)
        return float(v)

    step_fn = trainer.train_step_autograd if args.autograd else trainer.train_step

    # ---- untimed pre-warm (part of setup): grows the caching allocator's pools, loads every kernel variant, lets the
    #      clocks ramp.  The W warm-up steps the contract asks for still run before each timed region.
    loader.seeds_on_device = True
    it = batches(0)
    for _ in range(PREWARM_STEPS):
        step_fn(next(it))
    it.close()
    H = len(sh.fanouts)
    cap_n, cap_e = loader.max_nodes, loader.max_edges
    torch.cuda.synchronize()

    # ---- timed region 1: `value` — every input (graph, features, labels, the epoch's seed order) resident in HBM ----
    loader.seeds_on_device = True
    it = batches(0)
    for _ in range(W):
        step_fn(next(it))
    trainer.reset_stats()
    agg_events = []
    ops.timers = {"agg_l1": agg_events}
    clocks = ClockSampler(local_rank)
    barrier()
    if not args.autograd:
        _lib.call("ngnn_probe_enable", K)
    if args.ncu_range:
        torch.cuda.profiler.start()
    launches0 = lib.ngnn_launch_count()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    edges = 0
    len_done = 0
    pending_stats = None
    for _ in range(K):
        batch = next(it)
        step_fn(batch)
        # the loop logs its loss / accuracy every step like the reference's (pipeline.py:164-165): an asynchronous
        # device->host copy resolved one step later.  It also paces the host: without it the host runs many steps ahead
        # and the far-future blocks' sampling competes with the current step for the GPU (measured slower, 1 and 8 GPUs)
        handle = trainer.read_stats_async()
        if pending_stats is not None:
            trainer.resolve_stats(pending_stats)
        pending_stats = handle
        edges += batch.num_edges
        len_done += 1
        if len_done == K // 3 or len_done == (2 * K) // 3:
            clocks.sample()                                # GPU busy with the steps just enqueued; two NVML reads per run
    trainer.resolve_stats(pending_stats)
    ev1.record()
    barrier()
    if args.ncu_range:
        torch.cuda.profiler.stop()
    clock_info = clocks.stop()
    launches = lib.ngnn_launch_count() - launches0
    ops.timers = None
    it.close()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    edges_total = sum_over_ranks(edges)
    value = edges_total / (ms_total * 1e-3)

    # ---- roofline of the layer-1 aggregation: CUDA events recorded by the library around that launch, inside the
    #      timed region above (ngnn_probe_*), or by ops._timed in the autograd variant
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    if args.autograd:
        agg_ms = [a.elapsed_time(b) for a, b in agg_events]
    else:
        import ctypes
        buf, cnt = (ctypes.c_float * K)(), ctypes.c_int32(0)
        _lib.call("ngnn_probe_read", buf, K, ctypes.byref(cnt))
        agg_ms = list(buf[:cnt.value])
        _lib.call("ngnn_probe_enable", 0)
    # algorithmic bytes of each timed layer-1 launch: the sampler is a pure function of (seed, epoch, batch), so the same
    # blocks are sampled again here, outside the timed region, and their distinct table rows counted
    loader.seeds_on_device = True
    it = batches(0)
    for _ in range(W):
        next(it)
    agg_bytes = []
    for _ in range(K):
        blk = next(it).block
        n_dst, e1, _ = SAGE.layer_extents(blk, sh.layers)[0]
        agg_bytes.append(agg_l1_bytes(torch.cat([blk.col_global[:e1], blk.n_id[:n_dst]]), n_dst, e1, sh.features))
    it.close()
    achieved = (sum(agg_bytes) / len(agg_bytes)) / (sum(agg_ms) / len(agg_ms) * 1e-3) / 1e9 if agg_ms else None
    roofline = {"kernel": "k_agg_fwd_pipe<1,6,true> (K-AGG layer 1: mean of sampled in-neighbours + root gather from the resident table)",
                "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on a products layer-1 block, one
                # ncu --set full capture (profiles/r01_ncu_agg_pipe_summary.txt): 216.9 MB + 11.4 MB
                "traffic": 228.3e6,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "avg_launch_us": 1e3 * sum(agg_ms) / len(agg_ms) if agg_ms else None,
                "algorithmic_bytes_per_launch": sum(agg_bytes) / len(agg_bytes) if agg_bytes else None,
                "share_of_step": (sum(agg_ms) / ms_total) if agg_ms else None,
                "timing": "CUDA events on the launching stream around this kernel's launch, every step of the timed region"}

    # ---- timed region 2: `e2e` — public API with HOST seed buffers; loss / accuracy read back every step ----
    loader.seeds_on_device = False
    it = batches(0)
    for _ in range(W):
        step_fn(next(it))
    barrier()
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    edges2 = 0
    t_wall = time.perf_counter()
    pending_stats = None
    for _ in range(K):
        batch = next(it)                                   # pinned H2D of the seed ids (inside the iterator)
        step_fn(batch)
        # D2H of this step's loss / accuracy (float(loss) / int(correct) of reference pipeline.py:164-165), every step;
        # the copy is asynchronous and read one step later so logging does not drain the GPU
        handle = trainer.read_stats_async()
        if pending_stats is not None:
            loss, correct = trainer.resolve_stats(pending_stats)
        pending_stats = handle
        edges2 += batch.num_edges
    loss, correct = trainer.resolve_stats(pending_stats)
    e1_.record()
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    it.close()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1_), wall_ms))
    e2e_value = sum_over_ranks(edges2) / (e2e_ms * 1e-3)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sh.batch_size * 8,
           "d2h_bytes_per_step": 8 + 4 * 2 * (H + 1), "ms_per_step": e2e_ms / K,
           "note": "graph + feature table uploaded once (resident); per step: seed ids H2D (pinned), block extents D2H, loss/correct "
                   "D2H (async copy, resolved one step later; the last one before the closing timestamp)"}

    # ---- per-kernel-class breakdown (untimed extra pass through the autograd variant: one FFI call per kernel class) ----
    breakdown = None
    # (single process only: the autograd variant all-reduces in its optimizer step, so running it on rank 0 alone
    #  would leave the other ranks' collectives unmatched)
    if not args.no_breakdown and rank == 0 and world == 1:
        it = batches(0)
        for _ in range(W):
            trainer.train_step_autograd(next(it))
        ops.timers, ops.timers_open = {}, True
        nbk = min(10, K)
        for _ in range(nbk):
            trainer.train_step_autograd(next(it))
        torch.cuda.synchronize()
        breakdown = {k: round(1e3 * sum(a.elapsed_time(b) for a, b in v) / nbk, 2)
                     for k, v in sorted(ops.timers.items())}          # us per step
        ops.timers, ops.timers_open = None, False
        it.close()

    # ---- CPU baseline on the host cores (rank 0, N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and args.cpu_steps > 0:
        order = loader.epoch_permutation(0)
        cpu_baseline = run_cpu_steps(loader, data, sh, [loader.batch_seeds(order, i) for i in range(args.cpu_steps + 1)], warmup=1)

    if rank == 0:
        steps_per_epoch = -(-nb // world)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}-shaped graph ({data.num_nodes} nodes, {data.num_edges} directed edges, "
                                   f"F={sh.features}, C={sh.classes}), SAGE L={sh.layers} hidden={sh.hidden} fan-out={list(sh.fanouts)} "
                                   f"bs={sh.batch_size}/GPU, dropout={sh.dropout}, Adam lr=1e-3",
                       "degree_law": args.law, "scale": args.scale, "seed": 1232,
                       "step": "sample block -> SAGE fwd (trimmed to the rows the seed outputs depend on, exact) -> CE -> bwd -> "
                               "allreduce(N>1) -> Adam",
                       "l2": "inputs_larger_than_l2 (0.98 GB feature table + 0.5 GB CSC, a fresh random block every step)",
                       "parallelism": f"dp{world}",
                       "prewarm": f"{PREWARM_STEPS} untimed steps during setup (allocator pools, kernel variants, power state), "
                                  f"then the {W} warm-up steps before each timed region"},
            "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "epoch_time_s": steps_per_epoch * ms_total / K * 1e-3, "steps_per_epoch": steps_per_epoch,
            "avg_block": {"edges": edges_total / (K * world)}, "setup_s": setup_s, "kernel_us_per_step": breakdown,
        }
        _emit(real_stdout, out)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU oracle arm
def run_cpu_steps(loader, data, sh, seed_batches, warmup=1, threads=None):
    """Times the oracle (the reference's PyG CPU op sequence + sequential sampler with a prefetch thread, like
    NeighborLoader(num_workers=1)) on the host cores for len(seed_batches)-warmup steps of the same workload."""
    import numpy as np

    from oracle import sage_oracle, sampler as oracle_sampler, structure

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    colptr, row = loader.colptr.cpu().numpy(), loader.row.cpu().numpy()
    x, y, yhn = data.x.cpu(), data.y.cpu(), data.yhn.cpu()
    cs = oracle_sampler.CSampler(colptr, row)
    torch.manual_seed(1232)
    ref = sage_oracle.SAGERef(sh.features, sh.hidden, sh.classes, sh.layers, dropout=sh.dropout)
    ref.train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    fan = list(sh.fanouts)

    def produce(i):           # what the reference's single loader worker does per batch
        blk = cs.sample(seed_batches[i].numpy(), fan, seed=1232, epoch=0, batch_idx=i)
        n_id = torch.from_numpy(blk.n_id.astype(np.int64))
        ei = torch.from_numpy(structure.csr_to_coo(blk.rowptr, blk.col))
        return x[n_id], ei, y[n_id], yhn[n_id], blk.e

    results = {}

    def worker():
        for i in range(len(seed_batches)):
            results[i] = produce(i)

    th = threading.Thread(target=worker, daemon=True)
    t0 = None
    edges = 0
    th.start()
    for i in range(len(seed_batches)):
        if i == warmup:
            t0 = time.perf_counter()
        while i not in results:
            time.sleep(0.0005)
        xb, ei, yb, yhnb, e = results.pop(i)
        sage_oracle.train_step(ref, opt, xb, ei, yb, yhnb, len(seed_batches[i]))
        if i >= warmup:
            edges += e
    dt = time.perf_counter() - t0
    th.join()
    steps = len(seed_batches) - warmup
    cpu_model = None
    try:
        with open("/proc/cpuinfo") as f:
            cpu_model = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), None)
    except OSError:
        pass
    return {"value": edges / dt, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model,
            "torch_threads": torch.get_num_threads(),
            "sample": f"{steps} train steps (bs {len(seed_batches[0])}, fan-out {fan}, whole sampled block per layer as the "
                      f"reference computes) after {warmup} warm-up, sampler in a prefetch thread (num_workers=1)",
            "ms_per_step": 1e3 * dt / steps, "cpu_count": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  PyG is not installable here
    (SURVEY §8c), so this is the oracle port (kind 'port') with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    real_stdout = _claim_stdout()
    from noise_gnn_b200 import NeighborLoader
    if not torch.cuda.is_available():
        raise SystemExit("the synthetic graph is generated and CSC-sorted on the GPU for both arms; no GPU found")
    device = torch.device("cuda", 0)
    data, sh, train_idx = build_problem(args, device)
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size,
                            shuffle=True, seed=1232)
    order = loader.epoch_permutation(0)
    # One CPU step of this workload takes ~3 s on the box's 16 cores: the run is bounded to a few minutes by timing at most
    # REF_MAX_STEPS of the requested K steps (and REF_MAX_WARMUP of the W warm-ups); the metric is a rate, `steps` says
    # how many were timed.
    K, W = min(args.steps, REF_MAX_STEPS), min(args.warmup, REF_MAX_WARMUP)
    seeds = [loader.batch_seeds(order, i % loader.num_batches_global) for i in range(K + W)]
    res = run_cpu_steps(loader, data, sh, seeds, warmup=W)
    if (K, W) != (args.steps, args.warmup):
        res["sample"] += f"; bounded from the requested --steps {args.steps} --warmup {args.warmup}"
    out = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K,
           "warmup": W, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.workload}-shaped graph ({data.num_nodes} nodes, {data.num_edges} directed edges, "
                                  f"F={sh.features}, C={sh.classes}), SAGE L={sh.layers} hidden={sh.hidden} fan-out={list(sh.fanouts)} "
                                  f"bs={sh.batch_size}, dropout={sh.dropout}, Adam lr=1e-3", "seed": 1232,
                      "degree_law": args.law, "scale": args.scale},
           "cpu_baseline": res,
           "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(real_stdout, out)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
