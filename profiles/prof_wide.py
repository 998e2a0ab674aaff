"""K-AGG / K-AGG-T on wide rows (F = 512 .. 1433) of the Computers-shaped graph: the row's warp walking its column chunks
(ngnn_set_tuning(12, 0)) against the chunks spread over grid.y (1..5 = vectors per lane x neighbour rows in flight).
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
names = {0: "walk <32,8,1>/<32,4,2>", 2: "split <32,4,2>", 6: "split <32,4,2> L1", 7: "split <32,4,1>", 8: "split <32,4,4>",
         9: "scalar", 10: "split <32,4,2> T128", 11: "split <32,4,2> T512", 12: "walk <32,4,2> L1", 13: "walk <32,8,1> L1"}
for fan in (10, 25):
    loader = NeighborLoader(data, input_nodes=None, num_neighbors=[fan], batch_size=N, shuffle=False)
    blk = next(iter(loader)).block
    n, e = blk.n_rows, blk.e
    ct, rt = blk.transpose(e, n)
    for F in (512, 767, 1024, 1433):
        ld = (F + 3) // 4 * 4
        mk = lambda: torch.zeros((n, ld), device=dev)[:, :F]
        x, out, dm, dx = mk().normal_(), mk(), mk().normal_(), mk()
        for v in names:
            _lib.call("ngnn_set_tuning", 12, v)
            tf = timed(lambda: ops.agg_fwd(blk.rowptr, blk.col, x, n, out=out))
            tb = timed(lambda: ops.agg_bwd(ct, rt, dm, n, out=dx))
            print(f"fan={fan:2d} F={F:5d} {names[v]:24s} fwd {tf:7.1f} us  bwd {tb:7.1f} us", flush=True)
_lib.call("ngnn_set_tuning", 12, 0)
