"""Boundary proof (run with -m gpu): the reference's OWN loop bodies, through the ``compat/torch_geometric`` import shim,
on the B200 kernels, against the CPU oracle on identical blocks.

/root/reference does not exist on the GPU box and its sources may not be copied, so the two loop bodies are restated
here line for line with their citations (they are a dozen lines each); the network is the reference's unmodified
``src/models/layers/sage.py`` when that file is present (builder container with a GPU), else
``noise_gnn_b200.SAGE``, whose ``__init__`` / ``forward`` mirror it (tests/test_compat_shim.py loads the real file on
the same drop-in classes on the CPU)."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sage_oracle, sampler

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SAGE = "/root/reference/src/models/layers/sage.py"


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp(min=1e-30))


@pytest.fixture()
def shim(cuda_device):
    saved = {k: v for k, v in sys.modules.items() if k == "torch_geometric" or k.startswith("torch_geometric.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    yield cuda_device
    sys.path.remove(os.path.join(ROOT, "compat"))
    for k in [k for k in sys.modules if k == "torch_geometric" or k.startswith("torch_geometric.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def _network_class():
    if os.path.exists(REF_SAGE):                       # the unmodified reference module, PyG imports resolved by the shim
        spec = importlib.util.spec_from_file_location("ref_sage_gpu", REF_SAGE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.SAGE
    import noise_gnn_b200
    return noise_gnn_b200.SAGE


def reference_train_epoch(train_loader, model, optimizer, device, max_steps):
    """Loop body of PipelineCO.train, reference src/pipeline.py:152-169 (compare_loss == 'normal'), verbatim."""
    model.train()
    total_loss = total_correct = 0
    losses = []
    for step, batch in enumerate(train_loader):
        if step >= max_steps:
            break
        batch = batch.to(device)
        out = model(batch.x, batch.edge_index)[:batch.batch_size]
        y = batch.y[:batch.batch_size].squeeze()
        yhn = batch.yhn[:batch.batch_size].squeeze()
        loss = F.cross_entropy(out, yhn)
        total_loss += float(loss)
        total_correct += int(out.argmax(dim=-1).eq(y).sum())
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        losses.append(float(loss))
    return losses, total_correct


def reference_inference(model, x_all, subgraph_loader, device):
    """SAGE.inference, reference src/models/layers/sage.py:42-58, verbatim (x_all on the HOST, ids from the batch)."""
    for i in range(model.num_layers):
        xs = []
        for batch in subgraph_loader:
            x = x_all[batch.n_id].to(device)
            edge_index = batch.edge_index.to(device)
            x = model.convs[i](x, edge_index)
            x = x[:batch.batch_size]
            if i != model.num_layers - 1:
                x = x.relu()
            xs.append(x.cpu())
        x_all = torch.cat(xs, dim=0)
    return x_all


def _problem(fan, bs, scale=0.03):
    from noise_gnn_b200.synthetic import make_dataset
    from torch_geometric.data import Data
    d, sh, train_idx = make_dataset("arxiv", scale=scale, device="cpu", noise_type="sym", noise_rate=0.3)
    data = Data(x=d.x, edge_index=d.edge_index, y=d.y)
    data.yhn = d.yhn.view(-1, 1)                         # assigned after loading, like reference src/pipeline.py:72
    return data, sh, train_idx


def test_reference_train_loop_body_through_the_shim(shim):
    dev = shim
    from torch_geometric.loader import NeighborLoader       # the reference's import (src/pipeline.py:6)
    fan, bs, steps = [10, 5], 64, 4
    data, sh, train_idx = _problem(fan, bs)
    # constructed exactly as reference src/pipeline.py:75-83
    train_loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=fan, batch_size=bs, shuffle=True,
                                  num_workers=1, persistent_workers=True)
    torch.manual_seed(1232)
    ref = sage_oracle.SAGERef(sh.features, 64, sh.classes, 3, dropout=0.0)
    net = _network_class()(sh.features, 64, sh.classes, 3, dropout=0.0).to(dev)
    net.load_state_dict(ref.state_dict())
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)                    # reference src/models/model.py:67-69
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3)
    losses, correct = reference_train_epoch(train_loader, net, opt, dev, steps)
    # oracle: the same loop on the CPU over the blocks the sequential C sampler emits for the same keys
    cs = sampler.CSampler(train_loader.colptr.cpu().numpy(), train_loader.row.cpu().numpy())
    order = train_loader.epoch_permutation(0)
    from oracle import structure
    losses_ref, correct_ref = [], 0
    ref.train()
    for b in range(steps):
        blk = cs.sample(train_loader.batch_seeds(order, b).numpy(), fan, seed=train_loader.seed, epoch=0, batch_idx=b)
        n_id = torch.from_numpy(blk.n_id.astype(np.int64))
        ei = torch.from_numpy(structure.csr_to_coo(blk.rowptr, blk.col))
        l, c = sage_oracle.train_step(ref, opt_ref, data.x[n_id], ei, data.y[n_id], data.yhn[n_id], len(n_id[:bs]))
        losses_ref.append(l); correct_ref += c
    assert np.allclose(losses, losses_ref, rtol=2e-4, atol=1e-5), (losses, losses_ref)
    assert abs(correct - correct_ref) <= 1
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(p, q) < 1e-3, k
    assert len(train_loader) == -(-len(train_idx) // bs)                # reference src/pipeline.py:170


def test_reference_inference_loop_through_the_shim(shim):
    """x_all lives on the HOST as in the reference (self.data.x, src/pipeline.py:179) and is indexed by the batch's n_id."""
    dev = shim
    from torch_geometric.loader import NeighborLoader
    fan = [10, 5]
    data, sh, train_idx = _problem(fan, 64, scale=0.01)
    sub = NeighborLoader(data, input_nodes=None, num_neighbors=fan, batch_size=256, num_workers=1, persistent_workers=True)
    torch.manual_seed(3)
    ref = sage_oracle.SAGERef(sh.features, 64, sh.classes, 3, dropout=0.5, dtype=torch.float64)
    net = _network_class()(sh.features, 64, sh.classes, 3, dropout=0.5).to(dev)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    net.eval(); ref.eval()
    with torch.no_grad():
        out = reference_inference(net, data.x, sub, dev)
        # oracle: identical blocks (the loader's epoch counter advanced once per layer: 0, 1, 2)
        x_all = data.x.double()
        for i in range(3):
            sub.epoch = i
            xs = []
            for batch in sub:
                h = ref.convs[i](x_all[batch.n_id.cpu()], batch.edge_index.cpu())[: batch.batch_size]
                xs.append(h.relu() if i != 2 else h)
            x_all = torch.cat(xs)
    assert out.device.type == "cpu" and out.shape == x_all.shape
    assert rel_err(out, x_all) < 1e-5
