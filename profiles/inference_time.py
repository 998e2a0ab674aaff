"""Layer-wise inference (SAGE.inference / test_ogb, reference sage.py:42-58, pipeline.py:175-197) on the
products-shaped graph: every node, batch 4096, the training fan-outs, 3 layers; activations resident on the GPU."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import NeighborLoader, SAGE  # noqa: E402
from noise_gnn_b200.synthetic import make_dataset  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
data, sh, train_idx = make_dataset("products", device=dev)
sub = NeighborLoader(data, input_nodes=None, num_neighbors=list(sh.fanouts), batch_size=4096, shuffle=False)   # pipeline.py:85-92
model = SAGE(sh.features, sh.hidden, sh.classes, sh.layers, dropout=sh.dropout).to(dev).eval()
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.inference(data.x, sub, dev, return_cpu=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"inference pass {rep}: {dt:.3f} s for {data.num_nodes} nodes x {sh.layers} layers ({len(sub)} batches per layer, "
          f"{sh.layers * data.num_nodes / dt / 1e6:.1f} M node-layers/s), out {tuple(out.shape)}", flush=True)
