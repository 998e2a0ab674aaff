"""BASELINE.json configs[4]: SAGE aggregation microbench sweep on an Amazon-Computers-shaped graph
(13,752 nodes, 491,722 directed edges): feature width F in {64..1433} x fan-out in {5..25}, forward (K-AGG, CSR
segment mean) and backward (K-AGG-T, CSC transpose segment sum) separately, HBM GB/s against the measured peak.

    python profiles/agg_sweep_c5.py [--reps 7] [--out gpurun_out/agg_sweep_c5.json]

Block = every node as a destination with min(deg, fanout) sampled in-neighbours (SURVEY §8d "C5 sweep"), built by
the GPU sampler.  Algorithmic bytes: fwd 4F*u + 4e + 4(n+1) + 4F*n ; bwd 4F*n_dst + 4e + 4(u+1) + 4F*u
(u = distinct source rows).  The table (13,752 x F fp32 = 3.5 .. 79 MB) fits L2, so L2 is flushed (256 MB read)
before every launch and the figure is first-touch HBM traffic; `gather_gbs` is the no-reuse model 4F*e + ...
One products-shaped point (table 0.98 GB >> L2) is added so the roofline claim is tested where HBM really limits.
"""
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
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "agg_sweep_c5.json"))
ap.add_argument("--no-products", action="store_true")
ap.add_argument("--tune", action="append", default=[], help="key=value for ngnn_set_tuning, repeatable")
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
for kv in args.tune:
    k_, v_ = kv.split("=")
    _lib.call("ngnn_set_tuning", int(k_), int(v_))
pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk_path))["hbm_gbs"] if os.path.exists(pk_path) else 6650.0
flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)      # 256 MB, read before each launch (clean eviction)


def timed(fn):
    ms = []
    for _ in range(args.reps):
        flush.sum()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        z.record()
        z.synchronize()
        ms.append(a.elapsed_time(z))
    ms.sort()
    return ms[len(ms) // 2]


def one_block(name, fanout, scale=1.0, all_nodes=True, batch=None):
    data, sh, train_idx = make_dataset(name, device=dev, scale=scale)
    N = data.num_nodes
    if all_nodes:
        loader = NeighborLoader(data, input_nodes=None, num_neighbors=[fanout], batch_size=N, shuffle=False)
    else:
        loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(fanout), batch_size=batch, shuffle=True)
    b = next(iter(loader))
    return loader, b, sh


results = []


def padded(rows, F):
    """[rows, F] fp32 with the row stride rounded up to 4 floats — what the loader's table and the step arena use, so that
    widths like 767 / 1433 take the 128-bit kernels (the pad columns are never looked at)."""
    return torch.zeros((rows, (F + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :F]



def measure(tag, blk, n_dst, e, x, col, root_idx=None):
    """fwd on (rowptr, col) reading rows of x; bwd on the transpose writing n_src rows."""
    F = x.size(1)
    n_src = x.size(0) if root_idx is None and col is blk.col else None
    u = torch.unique(col[:e]).numel()
    out = padded(n_dst, F)
    t_f = timed(lambda: ops.agg_fwd(blk.rowptr, col, x, n_dst, out=out))
    by_f = 4 * F * u + 4 * e + 4 * (n_dst + 1) + 4 * F * n_dst
    ga_f = 4 * F * e + 4 * e + 4 * (n_dst + 1) + 4 * F * n_dst
    rec = {"case": tag, "F": F, "n_dst": n_dst, "e": e, "distinct_src": u,
           "fwd_us": round(t_f * 1e3, 2), "fwd_gbs": round(by_f / t_f / 1e6, 1), "fwd_frac": round(by_f / t_f / 1e6 / peak, 3),
           "fwd_gather_gbs": round(ga_f / t_f / 1e6, 1), "fwd_bytes": by_f}
    if n_src is not None:          # local-id block: the backward is defined (sources are block rows)
        colptr_t, row_t = blk.transpose(e, n_src)
        dmean = padded(n_dst, F).normal_()
        dx = padded(n_src, F)
        t_b = timed(lambda: ops.agg_bwd(colptr_t, row_t, dmean, n_src, out=dx))
        by_b = 4 * F * n_dst + 4 * e + 4 * (n_src + 1) + 4 * F * n_src
        rec.update({"bwd_us": round(t_b * 1e3, 2), "bwd_gbs": round(by_b / t_b / 1e6, 1),
                    "bwd_frac": round(by_b / t_b / 1e6 / peak, 3), "bwd_bytes": by_b})
    results.append(rec)
    print(json.dumps(rec), flush=True)


for fanout in (5, 10, 15, 20, 25):
    loader, b, sh = one_block("computers", fanout)
    blk = b.block
    n, e = blk.n_rows, blk.e
    for F in (64, 128, 256, 512, 767, 1024, 1433):
        x = padded(n, F).normal_()
        measure(f"computers fanout={fanout}", blk, n, e, x, blk.col)
        del x
    del loader, b

if not args.no_products:
    # non-L2-resident point: layer-1 block of the products workload, rows gathered from the 0.98 GB table by global id
    from noise_gnn_b200 import SAGE
    loader, b, sh = one_block("products", (15, 10, 5), all_nodes=False, batch=512)
    blk = b.block
    n_dst, e1, _ = SAGE.layer_extents(blk, sh.layers)[0]
    measure("products layer-1 block (table 0.98 GB)", blk, n_dst, e1, loader.x, blk.col_global)

with open(args.out, "w") as f:
    json.dump({"peak_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if os.path.exists(pk_path) else "fallback",
               "l2": "flushed by a 256 MB read before every launch", "reps": args.reps, "results": results}, f, indent=1)

print(f"\n{'case':42s} {'F':>5s} {'e':>8s} | {'fwd us':>8s} {'GB/s':>7s} {'frac':>5s} {'gather':>7s} | {'bwd us':>8s} {'GB/s':>7s} {'frac':>5s}")
for r in results:
    print(f"{r['case']:42s} {r['F']:5d} {r['e']:8d} | {r['fwd_us']:8.1f} {r['fwd_gbs']:7.0f} {r['fwd_frac']:5.2f} {r['fwd_gather_gbs']:7.0f} | "
          + (f"{r['bwd_us']:8.1f} {r['bwd_gbs']:7.0f} {r['bwd_frac']:5.2f}" if "bwd_us" in r else ""))
