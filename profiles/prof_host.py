import os, sys, cProfile, pstats
sys.path.insert(0, "/root/repo")
import torch
from noise_gnn_b200 import NeighborLoader, SAGE, _lib
from noise_gnn_b200.synthetic import make_dataset
from noise_gnn_b200.train import Trainer
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
data, sh, train_idx = make_dataset("products", device=dev, noise_type="sym", noise_rate=0.3)
loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size, shuffle=True, seed=1232, seeds_on_device=True)
model = SAGE(sh.features, sh.hidden, sh.classes, sh.layers, dropout=sh.dropout).to(dev); model.train()
trainer = Trainer(model, lr=1e-3)
it = iter(loader)
for _ in range(10): trainer.train_step(next(it))
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): trainer.train_step(next(it))
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
