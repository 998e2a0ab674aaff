"""Microbenchmark of the layer-1 K-AGG launch on a products-shaped graph (run under gpurun; optionally under ncu).

    python profiles/prof_agg.py [--sweep] [--reps N]

Times ngnn_sage_agg_fwd (mean of sampled in-neighbours + fused root gather from the resident 0.98 GB table) with
CUDA events, L2 flushed before every launch, and prints achieved GB/s against the compulsory-bytes formula."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import NeighborLoader, SAGE, _lib, ops  # noqa: E402
from noise_gnn_b200.synthetic import make_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sweep", action="store_true")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()

dev = torch.device("cuda", 0)
data, sh, train_idx = make_dataset("products", device=dev, scale=args.scale)
loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size, shuffle=True)
batches = []
for b in loader:
    batches.append(b)
    if len(batches) >= args.reps:
        break
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)   # 256 MB, READ before each launch: evicts L2 with clean lines


def run(tag):
    ms, by = [], []
    for b in batches:
        blk = b.block
        n_dst, e1, _ = SAGE.layer_extents(blk, sh.layers)[0]
        rows = torch.unique(torch.cat([blk.col_global[:e1], blk.n_id[:n_dst]])).numel()
        by.append(4 * sh.features * rows + 4 * e1 + 4 * (n_dst + 1) + 4 * n_dst + 2 * 4 * sh.features * n_dst)
        flush.sum()          # a write-flush would leave 126 MB of dirty lines whose write-back the kernel then pays for
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.agg_fwd(blk.rowptr, blk.col_global, loader.x, n_dst, root_idx=blk.n_id)
        z.record()
        z.synchronize()
        ms.append(a.elapsed_time(z))
    ms_s = sorted(ms)
    med = ms_s[len(ms_s) // 2]
    gbs = (sum(by) / len(by)) / (med * 1e-3) / 1e9
    print(f"{tag:28s} median {med * 1e3:7.1f} us  min {ms_s[0] * 1e3:7.1f} us  {gbs:7.1f} GB/s  {gbs / peak:5.3f} of measured peak "
          f"(n_dst {n_dst}, e {e1}, bytes {by[-1] / 1e6:.1f} MB)", flush=True)


run("warm-up")
for unroll in (8, 6, 5, 4, 3):
    _lib.call("ngnn_set_tuning", 0, unroll)
    run(f"pipelined persistent unroll={unroll}")
_lib.call("ngnn_set_tuning", 0, 0)
if args.sweep:
    _lib.call("ngnn_set_tuning", 3, 0)
    for group in (32, 16, 8):
        for threads in (128, 256):
            for unroll in (2, 4, 8):
                _lib.call("ngnn_set_tuning", 0, unroll)
                _lib.call("ngnn_set_tuning", 1, threads)
                _lib.call("ngnn_set_tuning", 2, group)
                run(f"lanes/row={group} threads={threads} unroll={unroll}")
    _lib.call("ngnn_set_tuning", 0, 0)
    _lib.call("ngnn_set_tuning", 1, 256)
    _lib.call("ngnn_set_tuning", 2, 32)
    _lib.call("ngnn_set_tuning", 3, 1)
run("default")
# (round 1 also timed a variant with the neighbour rows staged through shared memory by the async copy engines — UBLKCP /
#  LDGSTS, 75-83 us against 50 — which was removed in round 2; profiles/r01_kagg_experiments.txt keeps its numbers)

# ---- K-AGG-T on the layer-2 backward shape of the same blocks: dY_1 = gate(h_1) * (A^T dmean_2 + droot_2)
F2 = sh.hidden
ms, by = [], []
for b in batches:
    blk = b.block
    ext = SAGE.layer_extents(blk, sh.layers)
    n_dst2, e2, n_src2 = ext[1]
    colptr_t, row_t = blk.transpose(e2, n_src2)
    dmean = torch.randn((n_dst2, F2), device=dev)
    droot = torch.randn((n_dst2, F2), device=dev)
    h1 = torch.randn((n_src2, F2), device=dev)
    dx = torch.empty((n_src2, F2), device=dev)
    flush.sum()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.agg_bwd(colptr_t, row_t, dmean, n_src2, dx_root=droot, n_root=n_dst2, act_ref=h1, act_scale=2.0, out=dx)
    z.record()
    z.synchronize()
    ms.append(a.elapsed_time(z))
    by.append(4 * F2 * (2 * n_dst2 + 2 * n_src2) + 4 * e2 + 4 * (n_src2 + 1))
ms.sort()
med = ms[len(ms) // 2]
gbs = (sum(by) / len(by)) / (med * 1e-3) / 1e9
print(f"K-AGG-T layer-2 backward (n_src {n_src2}, e {e2}, F {F2}): median {med * 1e3:7.1f} us  {gbs:7.1f} GB/s  {gbs / peak:5.3f} of measured peak "
      f"(bytes {by[-1] / 1e6:.1f} MB: dmean + droot + gate read, dX written)")
