"""Pipeline trace of the tcgen05 K-GEMM (CTA 0): clock64 per K-block event.  Run under gpurun."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
import itertools
for (bn, ts), (n, F, O) in itertools.product(((128, 1),), ((77000, 100, 256),)):
    _lib.call('ngnn_set_tuning', 4, bn)
    _lib.call('ngnn_set_tuning', 6, ts)
    print('BN_max', bn, 'A-in-TMEM' if ts else 'A-in-smem')
    a_l, a_r = torch.randn(n, F, device=dev), torch.randn(n, F, device=dev)
    w_l, w_r, b = torch.randn(O, F, device=dev), torch.randn(O, F, device=dev), torch.randn(O, device=dev)
    for _ in range(3):
        ops.gemm_fwd(a_l, a_r, w_l, w_r, b, n, act=1)
    tr = torch.zeros(8 + 8 * 64 + 4 * 16, dtype=torch.int64, device=dev)
    _lib.call("ngnn_debug_set_trace", ops._ptr(tr))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.gemm_fwd(a_l, a_r, w_l, w_r, b, n, act=1)
    e.record()
    torch.cuda.synchronize()
    _lib.call("ngnn_debug_set_trace", None)
    t = tr.cpu().tolist()
    t0 = t[0]
    kb = 2 * ((F + 31) // 32)
    print(f"n={n} F={F} O={O}: kernel+prep {s.elapsed_time(e) * 1e3:.1f} us; CTA0: tmem_full at {t[1] - t0} cyc, epilogue done at {t[3] - t0} cyc, {kb} K-blocks")
    print("  kb: producer_go  raw_landed  conv_done  mma_start  mma_issued  committed  wait1_done  loop_top   (cycles since CTA start)")
    for k in range(min(24, kb * 3)):
        r = t[8 + 8 * k: 8 + 8 * k + 8]
        print(f"  {k:2d}: " + "  ".join(f"{v - t0:9d}" for v in r))
