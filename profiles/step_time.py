"""Plain step timing of the products workload (no probes, no clock sampler): ms/step by configuration."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from noise_gnn_b200 import NeighborLoader, SAGE, _lib  # noqa: E402
from noise_gnn_b200.synthetic import make_dataset  # noqa: E402
from noise_gnn_b200.train import Trainer  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
data, sh, train_idx = make_dataset("products", device=dev, noise_type="sym", noise_rate=0.3)
loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size, shuffle=True, seed=1232)
torch.manual_seed(1232)
model = SAGE(sh.features, sh.hidden, sh.classes, sh.layers, dropout=sh.dropout).to(dev)
model.train()
trainer = Trainer(model, lr=1e-3)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 60


def run(tag, on_device, probe=False):
    loader.seeds_on_device = on_device
    loader.epoch = 0
    it = iter(loader)
    for _ in range(5):
        trainer.train_step(next(it))
    torch.cuda.synchronize()
    if probe:
        _lib.call("ngnn_probe_enable", K)
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    host = 0.0
    for _ in range(K):
        b = next(it)
        h0 = time.perf_counter()
        trainer.train_step(b)
        host += time.perf_counter() - h0
    z.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    if probe:
        _lib.call("ngnn_probe_enable", 0)
    print(f"{tag:44s} {a.elapsed_time(z) / K:7.3f} ms/step (device)  host issue {1e3 * t_issue / K:6.3f} ms/step  "
          f"of which train_step {1e3 * host / K:6.3f}", flush=True)
    del it


run("seeds on device", True)
run("seeds on device (again)", True)
run("seeds on device + probe events", True, probe=True)
run("host seeds", False)
_lib.call("ngnn_set_step_overlap", 0)
run("seeds on device, single stream", True)
_lib.call("ngnn_set_step_overlap", 1)
_lib.call("ngnn_set_tuning", 6, 0)
run("seeds on device, SS GEMMs", True)
_lib.call("ngnn_set_tuning", 6, 1)

# ---- what bench.py adds around the loop
sys.path.insert(0, ROOT)
import bench  # noqa: E402

cs = bench.ClockSampler(0)
cs.start()
run("seeds on device + NVML clock sampler thread", True)
print(cs.stop())
keep = []
_orig = trainer.train_step
def step_and_cat(b):
    _orig(b)
    n_dst, e1, _ = SAGE.layer_extents(b.block, sh.layers)[0]
    keep.append(torch.cat([b.block.col_global[:e1], b.block.n_id[:n_dst]]))
trainer.train_step = step_and_cat
run("seeds on device + torch.cat of touched ids", True)
