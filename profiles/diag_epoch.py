"""Diagnostic driver for Trainer.train_epoch on a small graph: runs the epoch loop in one of several modes and prints progress, so that
a failing mode / step can be told apart (each mode is run in its own process by the caller).
    python profiles/diag_epoch.py MODE [steps]     MODE in eager_sync | eager | graph | graph_noaux
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import NeighborLoader, SAGE, _lib  # noqa: E402
from noise_gnn_b200.synthetic import make_dataset  # noqa: E402
from noise_gnn_b200.train import Trainer  # noqa: E402

mode = sys.argv[1]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
data, sh, train_idx = make_dataset("arxiv", scale=0.05, device="cpu", noise_type="sym", noise_rate=0.3)
loader = NeighborLoader(data, input_nodes=train_idx[:32 * 9 + 5], num_neighbors=[10, 5], batch_size=32, shuffle=True, seed=1232)
torch.manual_seed(1232)
net = SAGE(sh.features, 64, sh.classes, 3, dropout=0.0).to(dev)
net.train()
if mode == "graph_noaux":
    _lib.call("ngnn_set_step_overlap", 0)
tr = Trainer(net, lr=1e-3, use_graph=mode.startswith("graph"))
if mode == "eager_sync":
    orig = tr._enqueue_pair

    def wrapped(gs, k, a, b):
        orig(gs, k, a, b)
        torch.cuda.synchronize()
        print("  step ok", flush=True)
    tr._enqueue_pair = wrapped
for ep in range(3):
    loss, correct, log = tr.train_epoch(loader, epoch=ep)
    torch.cuda.synchronize()
    print(mode, "epoch", ep, "loss", loss, "correct", correct, [round(float(v), 5) for v in log[:, 0]], flush=True)
print(mode, "DONE", flush=True)
