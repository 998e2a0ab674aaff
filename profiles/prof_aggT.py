"""Microbenchmark of K-AGG-T (transpose segment sum + root rows + ReLU/dropout gate) on the layer-2 backward of a products block
(~77 k output rows x 256, ~84 k transposed edges) and of the layer-1 K-AGG, L2 flushed before every launch.
    python profiles/prof_aggT.py [--reps 5]
Prints achieved GB/s of the algorithmic bytes against the measured HBM peak for each tuning variant."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import NeighborLoader, _lib, ops  # noqa: E402
from noise_gnn_b200.synthetic import make_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--tune", action="append", default=[], help="key=value for ngnn_set_tuning, repeatable")
args = ap.parse_args()
dev = torch.device("cuda", 0)
for kv in args.tune:
    k_, v_ = kv.split("=")
    _lib.call("ngnn_set_tuning", int(k_), int(v_))
data, sh, train_idx = make_dataset("products", device=dev)
loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size, shuffle=True)
loader.transpose_hops = 2
batches = []
for b in loader:
    batches.append(b)
    if len(batches) >= args.reps:
        break
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
F = sh.hidden


def run(tag):
    ms, by = [], []
    for b in batches:
        blk = b.block
        n_dst, n_src, e = blk.hop_nodes[1], blk.hop_nodes[2], blk.hop_edges[2]
        ct, rt = blk._t[(e, n_src)]
        dmean = torch.randn(n_dst, F, device=dev)
        droot = torch.randn(n_dst, F, device=dev)
        h = torch.randn(n_src, F, device=dev)
        out = torch.empty(n_src, F, device=dev)
        # read gate rows + write dX rows for every source row, read the dmean rows the edges name (each distinct row once) and
        # the root rows, indices and extents once
        by.append(4 * F * n_src * 2 + 4 * F * n_dst * 2 + 4 * e + 4 * (n_src + 1))
        flush.sum()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.agg_bwd(ct, rt, dmean, n_src, dx_root=droot, n_root=n_dst, act_ref=h, act_scale=2.0, out=out)
        z.record()
        z.synchronize()
        ms.append(a.elapsed_time(z))
    ms.sort()
    med = ms[len(ms) // 2]
    gbs = (sum(by) / len(by)) / (med * 1e-3) / 1e9
    print(f"{tag:44s} median {med * 1e3:7.1f} us  min {ms[0] * 1e3:7.1f} us  {gbs:7.1f} GB/s  {gbs / peak:5.3f} of measured peak  "
          f"({by[-1] / 1e6:.1f} MB)", flush=True)


run("warm-up")
for l1 in (0, 1):
    _lib.call("ngnn_set_tuning", 13, l1)
    for wide in (0, 1, 2):
        for threads in (128, 256):
            _lib.call("ngnn_set_tuning", 9, wide)
            _lib.call("ngnn_set_tuning", 1, threads)
            run(f"K-AGG-T variant={wide} threads={threads} l1={l1}")
_lib.call("ngnn_set_tuning", 9, 0)
_lib.call("ngnn_set_tuning", 1, 256)
_lib.call("ngnn_set_tuning", 13, 0)
