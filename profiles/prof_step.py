"""Short eager run of the train step for `ncu` launch lists / single-kernel captures (profiles/r02_*):
    python profiles/prof_step.py [--steps 4] [--workload products] [--graph]
Issues the same calls as bench.py's steps (sample next block on the side stream || fused step || Adam) without replaying a
graph by default, so every kernel appears as its own launch."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import NeighborLoader, SAGE  # noqa: E402
from noise_gnn_b200.synthetic import make_dataset  # noqa: E402
from noise_gnn_b200.train import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--workload", default="products")
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
data, sh, train_idx = make_dataset(a.workload, seed=1232, device=dev, noise_type="sym", noise_rate=0.3)
loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size, shuffle=True, seed=1232)
torch.manual_seed(1232)
model = SAGE(sh.features, sh.hidden, sh.classes, sh.layers, dropout=sh.dropout).to(dev)
model.train()
tr = Trainer(model, lr=1e-3, use_graph=a.graph)
tr.run_steps(loader, a.steps, start_epoch=0, seeds_resident=True, log_every_step=False)
torch.cuda.synchronize()
print("done", a.steps, "steps")
