"""Times (and, under ncu, profiles) the dense kernels on the products layer-1 / layer-2 shapes.  Run under gpurun.

    python profiles/prof_gemm.py [--reps 5] [--which fwd|wgrad|dgrad|all]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--which", default="all")
ap.add_argument("--drop", type=float, default=0.5)
args = ap.parse_args()
dev = torch.device("cuda", 0)
flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)


def timed(fn, reps):
    ms = []
    for _ in range(reps):
        flush.sum()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); z.record(); z.synchronize()
        ms.append(a.elapsed_time(z))
    ms.sort()
    return ms[len(ms) // 2] * 1e3


for (n, F, O) in ((77000, 100, 256), (7700, 256, 256), (512, 256, 47)):
    a_l, a_r = torch.randn(n, F, device=dev), torch.randn(n, F, device=dev)
    w_l, w_r, b = torch.randn(O, F, device=dev) / F ** 0.5, torch.randn(O, F, device=dev) / F ** 0.5, torch.randn(O, device=dev)
    dy = torch.randn(n, O, device=dev)
    rowptr = torch.arange(n + 1, dtype=torch.int32, device=dev)
    flops = 2.0 * n * 2 * F * O
    if args.which in ("fwd", "all"):
        for ts in (0, 1):
            _lib.call("ngnn_set_tuning", 6, ts)
            t0 = timed(lambda: ops.gemm_fwd(a_l, a_r, w_l, w_r, b, n, act=1), args.reps)
            t1 = timed(lambda: ops.gemm_fwd(a_l, a_r, w_l, w_r, b, n, act=1, drop_p=args.drop, seed=1, offset=2), args.reps)
            print(f"fwd   n={n:6d} F={F:4d} O={O:4d} A-in-{'TMEM' if ts else 'smem'}: {t0:7.1f} us ({flops / t0 / 1e6:6.1f} TF/s fp32-equiv)   "
                  f"with dropout {args.drop}: {t1:7.1f} us", flush=True)
        _lib.call("ngnn_set_tuning", 6, 1)
    if args.which in ("dgrad", "all"):
        t = timed(lambda: ops.dgrad(dy, w_l, w_r, rowptr, n), args.reps)
        print(f"dgrad n={n:6d} F={F:4d} O={O:4d}: {t:7.1f} us ({flops / t / 1e6:6.1f} TF/s fp32-equiv)", flush=True)
    if args.which in ("wgrad", "all"):
        t = timed(lambda: ops.wgrad(dy, a_l, a_r, n, F), args.reps)
        print(f"wgrad n={n:6d} F={F:4d} O={O:4d}: {t:7.1f} us ({flops / t / 1e6:6.1f} TF/s fp32-equiv) (incl. reduce + db)", flush=True)
