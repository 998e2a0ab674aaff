#!/usr/bin/env python
"""bench.py — sampled edges/s through the GraphSAGE mini-batch train step (sample -> SAGE fwd -> CE -> bwd
-> [all-reduce] -> Adam) on a synthetic ogbn-products-shaped graph (BASELINE.json metric / configs[3]).

    python bench.py --gpus 1 --steps K --warmup W            # our arm (B200, libngnn_b200.so)
    torchrun ... bench.py --gpus N --steps K --warmup W      # data parallel, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle) on the host cores
    python bench.py --impl dropin --steps K --warmup W       # the reference's own loop body on the drop-in modules (GPU)
    python bench.py --workload arxiv|pubmed|cora             # the other BASELINE configs
    python bench.py --workload computers --sweep             # BASELINE configs[4]: K-AGG sweep, F x fan-out

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

PREWARM_STEPS = int(os.environ.get("NGNN_BENCH_PREWARM", "200"))   # untimed, part of setup: graph capture, allocator, clocks
REF_MAX_STEPS = 20                            # --impl reference: ~3 s per CPU step on the box's 16 cores
METRIC = "sampled_edges_per_sec"
UNIT = "edges/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--repeats", type=int, default=7, help="the (warmup + steps) region is measured this many times; the median is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "dropin"])
    ap.add_argument("--workload", default="products", choices=["products", "arxiv", "pubmed", "cora", "computers"])
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[4]: the K-AGG sweep (with --workload computers)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debug only; default = full shape)")
    ap.add_argument("--law", default="powerlaw", choices=["powerlaw", "uniform"])
    ap.add_argument("--cpu-steps", type=int, default=3, help="steps of the bounded cpu_baseline sample (0 = skip)")
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly (A/B against the replayed step)")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks during the timed region
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, read through NVML from the timing loop itself (the GPU is busy
    with the steps already enqueued)."""
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

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------- helpers
def workload_string(args, data, sh, per_gpu=True):
    return (f"{args.workload}-shaped graph ({data.num_nodes} nodes, {data.num_edges} directed edges, F={sh.features}, "
            f"C={sh.classes}), SAGE L={sh.layers} hidden={sh.hidden} fan-out={list(sh.fanouts)} bs={sh.batch_size}"
            f"{'/GPU' if per_gpu else ''}, dropout={sh.dropout}, Adam lr=1e-3")


def agg_l1_bytes(touched, n_dst, e, F):
    """ALGORITHMIC bytes of the layer-1 aggregation launch: every distinct table row read once,
    indices once, mean + root outputs written once."""
    rows = torch.unique(touched).numel()
    return 4 * F * rows + 4 * e + 4 * (n_dst + 1) + 4 * n_dst + 2 * 4 * F * n_dst


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


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            return next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), None)
    except OSError:
        return None


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist

    real_stdout = _claim_stdout()

    from noise_gnn_b200 import NeighborLoader, SAGE, _lib
    from noise_gnn_b200.loader import BlockSlot
    from noise_gnn_b200.synthetic import make_dataset
    from noise_gnn_b200.train import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a B200: there is no CPU fallback")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    assert lib.ngnn_device_supported() == 1, "libngnn_b200.so is sm_100a only"
    for kv in filter(None, os.environ.get("NGNN_TUNING", "").split(",")):      # A/B switches, e.g. NGNN_TUNING=7:0
        k, v = kv.split(":")
        _lib.call("ngnn_set_tuning", int(k), int(v))

    t_setup = time.time()
    data, sh, train_idx = make_dataset(args.workload, seed=1232, law=args.law, device=device, noise_type="sym", noise_rate=0.3,
                                       scale=args.scale)
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size,
                            shuffle=True, seed=1232, rank=rank, world_size=world)
    torch.manual_seed(1232)
    model = SAGE(sh.features, sh.hidden, sh.classes, sh.layers, dropout=sh.dropout).to(device)
    model.train()
    trainer = Trainer(model, lr=1e-3, world_size=world, rank=rank, use_graph=not args.no_graph)
    torch.cuda.synchronize()

    K, W, R = args.steps, max(args.warmup, 0), max(args.repeats, 1)
    H, L = len(sh.fanouts), sh.layers
    spe = len(loader)

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

    # ---- untimed pre-warm (part of setup): captures the step's CUDA graphs, grows the allocator pools, ramps the clocks
    trainer.run_steps(loader, PREWARM_STEPS, start_epoch=0, seeds_resident=True, log_every_step=False)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup
    clocks = ClockSampler(local_rank)
    pos = {"next": PREWARM_STEPS}                    # schedule position: every region trains on blocks nobody has seen yet

    def region(seeds_resident, log_every_step, steps=K, sample_clocks=False):
        """W untimed warm-up steps, then exactly `steps` timed steps bracketed by barrier + synchronize; device time (events)."""
        start = pos["next"]
        pos["next"] += W + steps
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c = {}

        def hook(j):
            if j == W - 1:
                barrier()
                c["l0"], c["r0"] = lib.ngnn_launch_count(), trainer.replayed_launches
                ev0.record()
            elif sample_clocks and j in (W + steps // 3, W + (2 * steps) // 3):
                clocks.sample()                                # the GPU is busy with the steps just enqueued
            if j == W + steps - 1:
                ev1.record()
                c["l1"], c["r1"] = lib.ngnn_launch_count(), trainer.replayed_launches
        if W == 0:
            barrier()
            c["l0"], c["r0"] = lib.ngnn_launch_count(), trainer.replayed_launches
            ev0.record()
        trainer.run_steps(loader, W + steps, start_epoch=start // spe, start_step=start % spe, seeds_resident=seeds_resident,
                          log_every_step=log_every_step, on_step=hook)
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        launches = (c["l1"] - c["l0"]) + (c["r1"] - c["r0"])
        return ms, start + W, launches

    # the sampler is a pure function of (seed, epoch, batch): the blocks of a timed region are sampled again afterwards,
    # outside it, to count their edges (and the distinct table rows their layer-1 aggregation reads)
    scratch = BlockSlot(loader, 0)

    tbytes = []

    def region_blocks(first, steps, want_bytes=False):
        edges, nbytes = [], []
        for j in range(first, first + steps):
            epoch, i = j // spe, j % spe
            g = loader.sharder.global_batch_index(i)
            seeds = loader.batch_seeds(loader.epoch_permutation(epoch), g)
            loader.launch_sample(scratch, seeds.pin_memory(), seeds.numel(), epoch, g, transposes=0)
            c = scratch.counts.tolist()
            edges.append(c[2 * H + 1])
            if want_bytes:
                n_dst, e1 = c[min(L - 1, H)], c[H + 1 + min(L, H)]
                nbytes.append(agg_l1_bytes(torch.cat([scratch.colg[:e1], scratch.n_id[:n_dst]]), n_dst, e1, sh.features))
                if L >= 2:
                    # K-AGG-T into layer 1's output rows: n_src = those rows, n_dst2 = layer 2's destination rows, e2 = its edges
                    n_src, n_dst2, e2, Fh = n_dst, c[min(L - 2, H)], c[H + 1 + min(L - 1, H)], sh.hidden
                    survey = 4 * Fh * n_dst2 + 4 * e2 + 4 * (n_src + 1) + 4 * Fh * n_src          # SURVEY §8(d), literal
                    fused = survey + 4 * Fh * n_src + 4 * Fh * n_dst2                               # + gate rows read + root-gradient rows read
                    tbytes.append((fused, survey))
        return edges, nbytes

    # ---- `value`: every input (graph, features, labels, the epochs' seed orders) resident in HBM ----
    vals, launches_v = [], 0
    for r in range(R):
        ms, first, launches_v = region(seeds_resident=True, log_every_step=False, sample_clocks=(r == R // 2))
        e_total = sum_over_ranks(sum(region_blocks(first, K)[0]))
        vals.append((e_total / (ms * 1e-3), ms, e_total))
    vals.sort()
    value, ms_total, edges_total = vals[len(vals) // 2]

    # ---- `e2e`: the public call (Trainer.run_steps = the loop of PipelineCO.train) with HOST seed buffers: per step the
    #      seed ids travel H2D from pinned memory and the loss / accuracy accumulators D2H (read back every step) ----
    e2es = []
    for r in range(R):
        ms, first, _ = region(seeds_resident=False, log_every_step=True)
        e_tot = sum_over_ranks(sum(region_blocks(first, K)[0]))
        e2es.append((e_tot / (ms * 1e-3), ms))
    e2es.sort()
    e2e_value, e2e_ms = e2es[len(e2es) // 2]
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sh.batch_size * 8, "d2h_bytes_per_step": 8,
           "ms_per_step": e2e_ms / K, "min": e2es[0][0], "max": e2es[-1][0], "repeats": R,
           "note": "graph + feature table uploaded once (resident); per step: seed ids H2D (pinned host memory), 32 bytes of "
                   "control words as kernel arguments, loss/correct accumulators D2H (asynchronous copy into pinned memory, "
                   "all resolved before the closing timestamp's synchronize)"}

    # ---- roofline of the layer-1 aggregation: the same steps issued EAGERLY (same kernels, same streams, not replayed)
    #      with a CUDA-event pair recorded by the library around that launch in every step (ngnn_probe_*) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    import ctypes
    Kp = min(K, 16)
    trainer.use_graph = False
    start = pos["next"]
    pos["next"] += W + Kp
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def hold_gpu():
        # Eager issue is host-bound (~0.45 ms of launches per step): an event recorded by a host that is NOT ahead of the GPU
        # would time the host's gap between two enqueues, not the kernel.  A spin kernel holds the GPU for ~25 ms while the
        # host queues all Kp steps; everything between the two events of a pair is then GPU-side.
        torch.cuda._sleep(50_000_000)

    def probe_hook(j):
        if j == W - 1:
            barrier()
            _lib.call("ngnn_probe_enable", Kp)
            hold_gpu()
            ev0.record()
        if j == W + Kp - 1:
            ev1.record()
    if W == 0:
        _lib.call("ngnn_probe_enable", Kp)
        hold_gpu()
        ev0.record()
    trainer.run_steps(loader, W + Kp, start_epoch=start // spe, start_step=start % spe, seeds_resident=True,
                      log_every_step=False, on_step=probe_hook)
    barrier()
    eager_ms = ev0.elapsed_time(ev1)
    trainer.use_graph = not args.no_graph
    buf, cnt = (ctypes.c_float * Kp)(), ctypes.c_int32(0)
    _lib.call("ngnn_probe_read", buf, Kp, ctypes.byref(cnt))
    agg_ms = list(buf[:cnt.value])
    _lib.call("ngnn_probe_read_device_clock", buf, Kp, ctypes.byref(cnt))
    agg_ms_dev = [v for v in buf[:cnt.value] if v > 0]
    _lib.call("ngnn_probe_read_agg_t", buf, Kp, ctypes.byref(cnt))
    aggt_ms = list(buf[:cnt.value])
    _lib.call("ngnn_probe_enable", 0)
    _, agg_bytes = region_blocks(start + W, Kp, want_bytes=True)
    k_agg_t = None
    if aggt_ms and tbytes:
        t_us = 1e3 * sum(aggt_ms) / len(aggt_ms)
        fused, survey = sum(b[0] for b in tbytes) / len(tbytes), sum(b[1] for b in tbytes) / len(tbytes)
        k_agg_t = {"kernel": "K-AGG-T into layer 1's output rows (transpose-sum of dmean + root-gradient rows, fused ReLU/dropout gate)",
                   "bound": "hbm", "avg_launch_us": t_us, "achieved": fused / (t_us * 1e-6) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                   "frac": fused / (t_us * 1e-6) / 1e9 / peak_gbs, "algorithmic_bytes_per_launch": fused,
                   "frac_survey_formula": survey / (t_us * 1e-6) / 1e9 / peak_gbs, "survey_bytes_per_launch": survey,
                   "note": "same eagerly issued steps and event-pair timing as the K-AGG figure; algorithmic bytes = SURVEY §8(d)'s "
                           "K-AGG-T formula (dmean rows read, indices, extents, dX rows written) + what the fused epilogue must also "
                           "read: the gate rows (the saved layer output) and the root-gradient rows; frac_survey_formula uses the "
                           "literal formula alone"}
    achieved = (sum(agg_bytes) / len(agg_bytes)) / (sum(agg_ms) / len(agg_ms) * 1e-3) / 1e9 if agg_ms else None
    roofline = {"kernel": "K-AGG layer 1 (mean of sampled in-neighbours + root gather from the resident table)",
                "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, ONE ncu --set full capture of a step of this workload
                # (profiles/r02_ncu_agg_pipe_raw.csv: 211.7 MB + 43.8 MB); a static figure, not re-measured by this run
                "traffic": 255.5e6 if args.workload == "products" and args.scale == 1.0 else None,
                "traffic_source": "static: profiles/r02_ncu_agg_pipe_raw.csv (ncu --set full, one launch)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "avg_launch_us": 1e3 * sum(agg_ms) / len(agg_ms) if agg_ms else None,
                # the same launches by the device's own clock (first CTA start -> last CTA end, %globaltimer): what the event
                # pair adds on top is the two records and the launch latency
                "avg_launch_us_device_clock": 1e3 * sum(agg_ms_dev) / len(agg_ms_dev) if agg_ms_dev else None,
                "frac_device_clock": (sum(agg_bytes) / len(agg_bytes)) / (sum(agg_ms_dev) / len(agg_ms_dev) * 1e-3) / 1e9 / peak_gbs
                if agg_ms_dev else None,
                "algorithmic_bytes_per_launch": sum(agg_bytes) / len(agg_bytes) if agg_bytes else None,
                "share_of_step": (sum(agg_ms) / len(agg_ms)) / (ms_total / K) if agg_ms else None,
                "k_agg_t": k_agg_t,
                "timing": f"CUDA events on the launching stream around this kernel's launch in {Kp} eagerly issued steps, queued "
                          f"while a spin kernel holds the GPU so that the host is ahead ({eager_ms / Kp:.3f} ms/step vs "
                          f"{ms_total / K:.3f} replayed); share_of_step is against the replayed step"}

    # ---- CPU baseline + parity check on the host cores (rank 0, N = 1 only) ----
    cpu_baseline = parity = None
    if rank == 0 and world == 1 and args.cpu_steps > 0:
        order = loader.epoch_permutation(0)
        colptr, row = loader.colptr.cpu().numpy(), loader.row.cpu().numpy()
        cpu_baseline = run_cpu_steps(colptr, row, data.x.cpu(), data.y.cpu(), data.yhn.cpu(), sh,
                                     [loader.batch_seeds(order, i) for i in range(args.cpu_steps + 1)], warmup=1)
        parity = parity_check(loader, data, sh, device, colptr, row)

    if rank == 0:
        steps_per_epoch = -(-loader.num_batches_global // world)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "repeats": R, "value_min": vals[0][0], "value_max": vals[-1][0],
            "config": {"workload": workload_string(args, data, sh), "degree_law": args.law, "scale": args.scale, "seed": 1232,
                       "step": "sample block (+ its CSC transposes) -> SAGE fwd (trimmed to the rows the seed outputs depend on, "
                               "exact) -> CE -> bwd -> allreduce(N>1) -> Adam; one CUDA-graph replay per step"
                               + (" [--no-graph: issued eagerly]" if args.no_graph else ""),
                       "l2": "inputs_larger_than_l2 (0.98 GB feature table + 0.5 GB CSC, a fresh random block every step)",
                       "parallelism": f"dp{world}",
                       "timing": f"median of {R} regions of {W} warm-up + {K} timed steps, CUDA events, max over ranks",
                       "prewarm": f"{PREWARM_STEPS} untimed steps during setup (graph capture, allocator pools, power state)"},
            "clocks": clocks.result(), "e2e": e2e, "gpu_launches": int(launches_v), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity_check": parity,
            "epoch_time_s": steps_per_epoch * ms_total / K * 1e-3, "steps_per_epoch": steps_per_epoch,
            "avg_block": {"edges": edges_total / (K * world)}, "setup_s": setup_s,
            "graph_replays_per_region": K if not args.no_graph else 0,
        }
        _emit(real_stdout, out)
    if world > 1:
        # the captured steps hold NCCL work: drop them before the group goes; a watchdog makes sure the process ends either way
        t = threading.Timer(60.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        trainer.release_graphs()
        dist.barrier()
        dist.destroy_process_group()


def parity_check(loader, data, sh, device, colptr, row):
    """One step of the timed configuration (full graph, bs 512, the block of epoch 0 / batch 0), dropout off, through the
    fused step, against the fp64 oracle network on the sequential C sampler's block: block bit-exact, loss, and the worst
    parameter-gradient error (max |d| / max |g|) on identical ReLU gates (see oracle/sage_oracle.py)."""
    import numpy as np

    from noise_gnn_b200 import SAGE
    from noise_gnn_b200.train import Trainer
    from oracle import sage_oracle, sampler as oracle_sampler, structure
    t0 = time.time()
    torch.manual_seed(7)
    ref = sage_oracle.SAGERef(sh.features, sh.hidden, sh.classes, sh.layers, dropout=0.0, dtype=torch.float64)
    net = SAGE(sh.features, sh.hidden, sh.classes, sh.layers, dropout=0.0).to(device)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    net.train()
    tr = Trainer(net, lr=1e-3, use_graph=False)
    loader.transpose_hops = max(loader.transpose_hops, min(sh.layers - 1, len(sh.fanouts)))
    seeds = loader.batch_seeds(loader.epoch_permutation(0), 0)
    batch = loader.sample(seeds, epoch=0, batch_idx=0)
    want = oracle_sampler.sample_block(colptr, row, seeds.numpy(), list(sh.fanouts), seed=loader.seed, epoch=0, batch_idx=0)
    block_exact = bool(np.array_equal(batch.block.n_id.cpu().numpy(), want.n_id) and
                       np.array_equal(batch.block.rowptr.cpu().numpy(), want.rowptr) and
                       np.array_equal(batch.block.col.cpu().numpy(), want.col))
    tr.forward_backward(batch)
    loss, _ = tr.read_stats()
    with torch.no_grad():
        _, hidden = net.forward_batch(batch, return_hidden=True)
    relu_masks = [(h > 0).cpu() for h in hidden[:-1]]
    n_id = torch.from_numpy(want.n_id.astype(np.int64))
    ei = torch.from_numpy(structure.csr_to_coo(want.rowptr, want.col))
    bs = seeds.numel()
    out = ref(data.x.cpu()[n_id].double(), ei, relu_masks=relu_masks)[:bs]
    loss_ref = torch.nn.functional.cross_entropy(out, data.yhn.cpu()[n_id][:bs].view(-1))
    loss_ref.backward()
    err = max(float((p.grad.double().cpu() - q.grad).abs().max() / q.grad.abs().max().clamp(min=1e-30))
              for (_, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()))
    return {"block_bit_exact": block_exact, "loss": loss, "loss_oracle": float(loss_ref.detach()),
            "loss_rel_err": abs(loss - float(loss_ref.detach())) / max(abs(float(loss_ref.detach())), 1e-30),
            "max_grad_rel_err": err, "tolerance": 2e-5, "ok": bool(block_exact and err < 2e-5),
            "what": "timed configuration, block (epoch 0, batch 0), dropout off, fused step vs fp64 oracle on the C sampler's block",
            "seconds": round(time.time() - t0, 1)}


# --------------------------------------------------------------------------- CPU oracle arm
def run_cpu_steps(colptr, row, x, y, yhn, sh, seed_batches, warmup=1, threads=None):
    """Times the oracle (the reference's PyG CPU op sequence + sequential sampler with a prefetch thread, like
    NeighborLoader(num_workers=1)) on the host cores for len(seed_batches)-warmup steps of the same workload."""
    import numpy as np

    from oracle import sage_oracle, sampler as oracle_sampler, structure

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
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
    return {"value": edges / dt, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": _cpu_model(),
            "torch_threads": torch.get_num_threads(),
            "sample": f"{steps} train steps (bs {len(seed_batches[0])}, fan-out {fan}, whole sampled block per layer as the "
                      f"reference computes) after {warmup} warm-up, sampler in a prefetch thread (num_workers=1)",
            "ms_per_step": 1e3 * dt / steps, "cpu_count": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  PyG is not installable here, so
    this is the oracle port (kind 'port') with all host threads.  Nothing of libngnn_b200.so is loaded in this process:
    the synthetic graph is generated with torch on the CPU and put in CSC order by the numpy oracle."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    real_stdout = _claim_stdout()
    import numpy as np

    from noise_gnn_b200.sharding import SeedSharder
    from noise_gnn_b200.synthetic import make_dataset
    from oracle import structure
    t0 = time.time()
    data, sh, train_idx = make_dataset(args.workload, seed=1232, law=args.law, device="cpu", noise_type="sym", noise_rate=0.3,
                                       scale=args.scale)
    colptr, row, _ = structure.coo_to_csr(data.edge_index[0].numpy(), data.edge_index[1].numpy(), data.num_nodes)
    colptr, row = colptr.astype(np.int32), row.astype(np.int32)
    sharder = SeedSharder(train_idx, sh.batch_size, True, 1232)
    order = sharder.epoch_permutation(0)
    setup_s = time.time() - t0
    # One CPU step of this workload takes ~3 s on the box's 16 cores: the run is bounded to a few minutes by timing at most
    # REF_MAX_STEPS of the requested K steps; the metric is a rate, `steps` says how many were timed.
    K, W = min(args.steps, REF_MAX_STEPS), args.warmup
    seeds = [sharder.batch_seeds(order, i % sharder.num_batches_global) for i in range(K + W)]
    res = run_cpu_steps(colptr, row, data.x, data.y, data.yhn, sh, seeds, warmup=W)
    if K != args.steps:
        res["sample"] += f"; bounded from the requested --steps {args.steps}"
    out = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K,
           "warmup": W, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_string(args, data, sh), "seed": 1232, "degree_law": args.law, "scale": args.scale,
                      "graph": "generated with torch's CPU generator (same law, shape and seed as the GPU arm's; a different draw)"},
           "cpu_baseline": res, "setup_s": setup_s,
           "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(real_stdout, out)


# --------------------------------------------------------------------------- drop-in arm
def run_dropin(args):
    """--impl dropin: the loop body of PipelineCO.train (reference src/pipeline.py:152-169) VERBATIM — model(batch.x,
    batch.edge_index)[:batch_size], F.cross_entropy, float(loss), loss.backward(), torch.optim.Adam.step() — with the
    reference's imports resolved by compat/torch_geometric to the drop-in NeighborLoader / SAGEConv: every layer on the
    whole sampled block like the reference, autograd and torch's optimizer in charge, one kernel call per op."""
    import torch.nn.functional as F
    real_stdout = _claim_stdout()
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    from torch_geometric.loader import NeighborLoader          # the reference's own import lines
    from noise_gnn_b200 import SAGE
    from noise_gnn_b200.synthetic import make_dataset
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    data, sh, train_idx = make_dataset(args.workload, seed=1232, law=args.law, device=device, noise_type="sym", noise_rate=0.3,
                                       scale=args.scale)
    data.yhn = data.yhn.view(-1, 1)
    train_loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size,
                                  shuffle=True, num_workers=1, persistent_workers=True)
    torch.manual_seed(1232)
    model = SAGE(sh.features, sh.hidden, sh.classes, sh.layers, dropout=sh.dropout).to(device)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train()
    K, W = args.steps, args.warmup
    it = iter(train_loader)

    def next_batch():
        nonlocal it
        try:
            return next(it)
        except StopIteration:
            it = iter(train_loader)
            return next(it)

    def step(batch):
        batch = batch.to(device)
        out = model(batch.x, batch.edge_index)[:batch.batch_size]
        y = batch.y[:batch.batch_size].squeeze()
        yhn = batch.yhn[:batch.batch_size].squeeze()
        loss = F.cross_entropy(out, yhn)
        total = float(loss)
        correct = int(out.argmax(dim=-1).eq(y).sum())
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return total, correct, batch.num_edges

    for _ in range(W + 20):
        step(next_batch())
    vals = []
    for _ in range(max(args.repeats, 1)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        edges = 0
        for _ in range(K):
            edges += step(next_batch())[2]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        vals.append((edges / dt, dt))
    vals.sort()
    v, dt = vals[len(vals) // 2]
    out = {"impl": "dropin", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "repeats": len(vals), "value_min": vals[0][0], "value_max": vals[-1][0],
           "config": {"workload": workload_string(args, data, sh, per_gpu=False), "degree_law": args.law, "scale": args.scale,
                      "step": "reference loop body verbatim on the drop-in modules: every layer on the WHOLE block (untrimmed), "
                              "x[n_id] materialised, int64 edge_index exported, autograd, two host reads per step, torch Adam",
                      "timing": "host wall clock around K steps (the loop synchronises every step through float(loss))"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": sh.batch_size * 8, "d2h_bytes_per_step": 12}}
    _emit(real_stdout, out)


def run_sweep(args):
    """BASELINE configs[4]: K-AGG / K-AGG-T sweep over feature width x fan-out on the Computers-shaped graph
    (profiles/agg_sweep_c5.py does the work and writes the table); one summary JSON line."""
    real_stdout = _claim_stdout()
    out_path = os.path.join(ROOT, "gpurun_out", "agg_sweep_c5.json")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "agg_sweep_c5.py"), "--out", out_path],
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stderr[-2000:])
    res = json.load(open(out_path))
    rows = [x for x in res["results"] if x["case"].startswith("computers")]
    fr = [x["fwd_frac"] for x in rows]
    br = [x["bwd_frac"] for x in rows if "bwd_frac" in x]
    prod = [x for x in res["results"] if x["case"].startswith("products")]
    out = {"metric": "agg_hbm_fraction_of_roofline", "value": statistics.median(fr), "unit": "fraction of measured HBM peak",
           "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "computers-shaped graph (13752 nodes), K-AGG fwd / K-AGG-T bwd, F in {64..1433} x fan-out in {5..25}, "
                                  "L2 flushed before every launch"},
           "roofline": {"bound": "hbm", "peak": res["peak_gbs"], "unit": "GB/s", "fwd_frac_median": statistics.median(fr),
                        "fwd_frac_min": min(fr), "fwd_frac_max": max(fr), "bwd_frac_median": statistics.median(br) if br else None,
                        "bwd_frac_min": min(br) if br else None, "bwd_frac_max": max(br) if br else None,
                        "products_layer1_point": prod[0] if prod else None},
           "table": rows}
    _emit(real_stdout, out)


if __name__ == "__main__":
    a = parse_args()
    if a.sweep:
        run_sweep(a)
    elif a.impl == "reference":
        run_reference(a)
    elif a.impl == "dropin":
        run_dropin(a)
    else:
        run_ours(a)
