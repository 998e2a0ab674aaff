"""K-AGG-T (and the generic K-AGG) on mid-width rows (F = 128, 256) of the Computers-shaped graph: lanes per row, rows in flight,
L1 allocation and the hub-row threshold.
    python profiles/prof_mid.py"""
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


def tune(**kw):
    for k, v in kw.items():
        _lib.call("ngnn_set_tuning", int(k[1:]), v)


data, sh, _ = make_dataset("computers", device=dev)
N = data.num_nodes
for fan in (10, 25):
    loader = NeighborLoader(data, input_nodes=None, num_neighbors=[fan], batch_size=N, shuffle=False)
    blk = next(iter(loader)).block
    n, e = blk.n_rows, blk.e
    ct, rt = blk.transpose(e, n)
    for F in (64, 128, 256):
        x, out, dm, dx = (torch.randn(n, F, device=dev) for _ in range(4))
        for pipe in (1, 0):
            for wide in ((0, 1, 2) if (F == 256 and pipe == 0) else (0,)):
                tune(k3=pipe, k9=wide)
                tf = timed(lambda: ops.agg_fwd(blk.rowptr, blk.col, x, n, out=out))
                tb = timed(lambda: ops.agg_bwd(ct, rt, dm, n, out=dx))
                print(f"fan={fan:2d} F={F} pipe={pipe} variant={wide}  fwd {tf:7.1f} us  bwd {tb:7.1f} us", flush=True)
tune(k3=1, k9=0)
