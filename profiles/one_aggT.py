"""One K-AGG-T launch on the layer-2 backward of a products block (for an ncu --set full capture).
    ncu --set full --clock-control none --import-source on -k regex:k_seg_reduce -s 1 -c 1 -o gpurun_out/aggT python profiles/one_aggT.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import NeighborLoader, _lib, ops  # noqa: E402
from noise_gnn_b200.synthetic import make_dataset  # noqa: E402

for kv in sys.argv[1:]:
    k_, v_ = kv.split("=")
    _lib.call("ngnn_set_tuning", int(k_), int(v_))
dev = torch.device("cuda", 0)
data, sh, train_idx = make_dataset("products", device=dev)
loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size, shuffle=True)
loader.transpose_hops = 2
b = next(iter(loader))
blk = b.block
F = sh.hidden
n_dst, n_src, e = blk.hop_nodes[1], blk.hop_nodes[2], blk.hop_edges[2]
ct, rt = blk._t[(e, n_src)]
dmean, droot = torch.randn(n_dst, F, device=dev), torch.randn(n_dst, F, device=dev)
h, out = torch.randn(n_src, F, device=dev), torch.empty(n_src, F, device=dev)
flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
deg = (ct[1:] - ct[:-1]).float()
print("rows", n_src, "edges", e, "deg mean", float(deg.mean()), "max", int(deg.max()), "zero rows", int((deg == 0).sum()), flush=True)
for _ in range(3):
    flush.sum()
    ops.agg_bwd(ct, rt, dmean, n_src, dx_root=droot, n_root=n_dst, act_ref=h, act_scale=2.0, out=out)
torch.cuda.synchronize()
