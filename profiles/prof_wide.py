"""K-AGG / K-AGG-T on wide rows (F = 767, 1024, 1433) of the Computers-shaped graph, fan-out 10: padded (128-bit kernels) against
unpadded (scalar kernels) rows and CTA sizes — the part of the C5 sweep that moved between rounds.
    python profiles/prof_wide.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import NeighborLoader, _lib, ops  # noqa: E402
from noise_gnn_b200.synthetic import make_dataset  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)


def timed(fn, reps=9):
    ms = []
    for _ in range(reps):
        flush.sum()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); z.record(); z.synchronize()
        ms.append(a.elapsed_time(z))
    ms.sort()
    return ms[len(ms) // 2] * 1e3


data, sh, _ = make_dataset("computers", device=dev)
N = data.num_nodes
loader = NeighborLoader(data, input_nodes=None, num_neighbors=[10], batch_size=N, shuffle=False)
blk = next(iter(loader)).block
n, e = blk.n_rows, blk.e
ct, rt = blk.transpose(e, n)
deg = (ct[1:] - ct[:-1])
print("transposed row lengths: max", int(deg.max()), "p99", int(torch.quantile(deg.float(), 0.99)), "mean", float(deg.float().mean()), flush=True)
for F in (256, 512, 1024):
    x = torch.zeros((n, F), device=dev).normal_(); out = torch.zeros((n, F), device=dev); dm = x.clone(); dx = out.clone()
    for lr in (1024, 128, 64, 48, 32):
        _lib.call("ngnn_set_tuning", 11, lr)
        tf = timed(lambda: ops.agg_fwd(blk.rowptr, blk.col, x, n, out=out))
        tb = timed(lambda: ops.agg_bwd(ct, rt, dm, n, out=dx))
        print(f"F={F:5d} long_row={lr:5d}  fwd {tf:7.1f} us  bwd {tb:7.1f} us", flush=True)
_lib.call("ngnn_set_tuning", 11, 64)
for F in ():
    for pad in (True, False):
        ld = (F + 3) // 4 * 4 if pad else F
        mk = lambda: torch.zeros((n, ld), device=dev)[:, :F] if pad else torch.zeros((n, F), device=dev)
        x, out, dm, dx = mk().normal_(), mk(), mk().normal_(), mk()
        for threads in (128, 256, 512):
            _lib.call("ngnn_set_tuning", 1, threads)
            tf = timed(lambda: ops.agg_fwd(blk.rowptr, blk.col, x, n, out=out))
            tb = timed(lambda: ops.agg_bwd(ct, rt, dm, n, out=dx))
            print(f"F={F:5d} {'padded  ' if pad else 'unpadded'} threads={threads:3d}  fwd {tf:7.1f} us  bwd {tb:7.1f} us", flush=True)
_lib.call("ngnn_set_tuning", 1, 256)
