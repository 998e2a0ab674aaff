"""GPU parity of the whole step (ngnn_sage_step, one C-ABI call) and of the prefetching loader against the oracle."""
import numpy as np
import pytest
import torch

from oracle import philox, sage_oracle, sampler

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp(min=1e-30))


@pytest.fixture(scope="module")
def dev(cuda_device):
    return cuda_device


def _setup(dev, L, fan, dropout, bs=64, hidden=64, name="arxiv", scale=0.02):
    from noise_gnn_b200 import NeighborLoader, SAGE
    from noise_gnn_b200.synthetic import make_dataset
    data, sh, train_idx = make_dataset(name, scale=scale, device="cpu", noise_type="sym", noise_rate=0.3)
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=fan, batch_size=bs, shuffle=True, seed=1232)
    torch.manual_seed(1232)
    ref = sage_oracle.SAGERef(sh.features, hidden, sh.classes, L, dropout=dropout, dtype=torch.float64)
    net = SAGE(sh.features, hidden, sh.classes, L, dropout=dropout).to(dev)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    return data, sh, loader, ref, net


@pytest.mark.parametrize("L,fan", [(3, [15, 10, 5]), (3, [10, 5]), (2, [10, 5]), (2, [7]), (2, [4, 4, 4]), (4, [5, 5])])
def test_fused_step_matches_oracle_and_autograd(dev, L, fan):
    from noise_gnn_b200.train import Trainer
    data, sh, loader, ref, net = _setup(dev, L, fan, dropout=0.0)
    trainer = Trainer(net, lr=1e-3)
    batch = next(iter(loader))
    bs = batch.batch_size
    tgt, y = batch.yhn[:bs].view(-1).cpu(), batch.y[:bs].view(-1).cpu()
    out_ref = ref(batch.x.cpu().double(), batch.edge_index.cpu())[:bs]
    loss_ref = torch.nn.functional.cross_entropy(out_ref, tgt)
    loss_ref.backward()
    logits = trainer.forward_backward(batch, want_logits=True)
    loss, correct = trainer.read_stats()
    assert rel_err(logits, out_ref) < 1e-5
    assert abs(loss - float(loss_ref)) < 1e-5 * max(1.0, float(loss_ref))
    assert correct == int((out_ref.argmax(-1) == y).sum())
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, q.grad) < 2e-5, k
    fused = trainer.buckets.grad.clone()
    # the per-kernel autograd variant runs the same kernels in the same order: identical gradients bit for bit
    net.zero_grad(set_to_none=False)
    out2 = net.forward_batch(batch)
    torch.nn.functional.cross_entropy(out2, tgt.to(dev)).backward()
    trainer.buckets.rebind_grads()
    assert rel_err(trainer.buckets.grad, fused) < 1e-6
    # eval mode / no-grad forward gives the same logits and leaves the gradients alone
    logits_eval = trainer.forward_backward(batch, train=False, want_logits=True)
    assert torch.equal(logits_eval, logits)


def test_fused_step_dropout_uses_the_philox_oracle_masks(dev):
    from noise_gnn_b200.train import Trainer
    L, fan, p = 3, [10, 5, 5], 0.5
    data, sh, loader, ref, net = _setup(dev, L, fan, dropout=p)
    net.train(); ref.train()
    trainer = Trainer(net, lr=1e-3)
    batch = next(iter(loader))
    bs, n = batch.batch_size, batch.num_nodes
    step = trainer.steps + 1
    masks = [torch.from_numpy(philox.dropout_keep_mask(n, 64, p, seed=net.drop_seed, offset=step * L + i)) for i in range(L - 1)]
    tgt = batch.yhn[:bs].view(-1).cpu()
    out_ref = ref(batch.x.cpu().double(), batch.edge_index.cpu(), dropout_masks=masks)[:bs]
    loss_ref = torch.nn.functional.cross_entropy(out_ref, tgt)
    loss_ref.backward()
    logits = trainer.forward_backward(batch, want_logits=True)
    assert rel_err(logits, out_ref) < 1e-5
    for (k, q), (_, r) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(q.grad, r.grad) < 2e-5, k


def test_training_loop_tracks_the_oracle_for_several_steps(dev):
    """Sample -> step -> Adam for 5 steps: losses and final parameters follow the CPU oracle trained on the same blocks."""
    from noise_gnn_b200.train import Trainer
    L, fan = 3, [10, 5]
    data, sh, loader, ref, net = _setup(dev, L, fan, dropout=0.0, bs=32)
    ref = ref.float()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    trainer = Trainer(net, lr=1e-3)
    cs = sampler.CSampler(loader.colptr.cpu().numpy(), loader.row.cpu().numpy())
    order = loader.epoch_permutation(0)
    losses, losses_ref = [], []
    for b, batch in enumerate(loader):
        if b >= 5:
            break
        blk = cs.sample(loader.batch_seeds(order, b).numpy(), fan, seed=1232, epoch=0, batch_idx=b)
        assert np.array_equal(batch.block.col.cpu().numpy(), blk.col)          # prefetching loader == oracle block
        assert np.array_equal(batch.block.n_id.cpu().numpy(), blk.n_id)
        trainer.reset_stats()
        trainer.train_step(batch)
        losses.append(trainer.read_stats()[0])
        l_ref, _ = sage_oracle.train_step(ref, opt, batch.x.cpu(), batch.edge_index.cpu(), batch.y.cpu(), batch.yhn.cpu(),
                                          batch.batch_size)
        losses_ref.append(l_ref)
    assert np.allclose(losses, losses_ref, rtol=2e-4, atol=1e-5), (losses, losses_ref)
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(p, q) < 1e-3, k


def test_inference_matches_oracle_layerwise(dev):
    from noise_gnn_b200 import NeighborLoader
    L, fan = 3, [10, 5]
    data, sh, loader, ref, net = _setup(dev, L, fan, dropout=0.5, scale=0.01)
    net.eval(); ref.eval()
    sub = NeighborLoader(data, input_nodes=None, num_neighbors=fan, batch_size=256, seed=7)
    out = net.inference(data.x, sub, dev)
    # oracle on the identical blocks: re-iterate the same epoch key
    sub.epoch = 0
    x_all = data.x.double()
    for i in range(L):
        sub.epoch = i       # inference() consumed one epoch per layer: 0, 1, 2
        xs = []
        for batch in sub:
            x = x_all[batch.n_id.cpu()]
            h = ref.convs[i](x, batch.edge_index.cpu())[: batch.batch_size]
            xs.append(h.relu() if i != L - 1 else h)
        x_all = torch.cat(xs)
    assert out.shape == x_all.shape
    assert rel_err(out, x_all) < 1e-5


@pytest.mark.parametrize("name,scale,hidden", [("cora", 1.0, 512), ("pubmed", 0.25, 256), ("arxiv", 0.02, 256),
                                               ("products", 0.002, 256), ("computers", 0.1, 256)])
def test_fused_step_on_every_baseline_config_shape(dev, name, scale, hidden):
    """BASELINE.json configs C1-C5 (feature widths 1433 / 500 / 128 / 100 / 767, the YAML layer counts and fan-outs):
    one fused step against the untrimmed fp64 oracle on the identical block.  F = 1433 / 767 exercise the scalar
    aggregation and SIMT GEMM kernels (rows that TMA cannot address), the others the tcgen05 path."""
    from noise_gnn_b200 import NeighborLoader, SAGE
    from noise_gnn_b200.synthetic import SHAPES, make_dataset
    from noise_gnn_b200.train import Trainer
    sh = SHAPES[name]
    data, _, train_idx = make_dataset(name, scale=scale, device="cpu", noise_type="sym", noise_rate=0.3)
    bs = min(sh.batch_size, 128, len(train_idx))
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=bs, shuffle=True)
    torch.manual_seed(5)
    ref = sage_oracle.SAGERef(sh.features, hidden, sh.classes, sh.layers, dropout=0.0, dtype=torch.float64)
    net = SAGE(sh.features, hidden, sh.classes, sh.layers, dropout=0.0).to(dev)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    trainer = Trainer(net)
    batch = next(iter(loader))
    tgt = batch.yhn[: batch.batch_size].view(-1).cpu()
    out_ref = ref(batch.x.cpu().double(), batch.edge_index.cpu())[: batch.batch_size]
    torch.nn.functional.cross_entropy(out_ref, tgt).backward()
    logits = trainer.forward_backward(batch, want_logits=True)
    assert rel_err(logits, out_ref) < 1e-5
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, q.grad) < 2e-5, (name, k)


# ------------------------------------------------------------------ co-teaching (reference pipeline.py:95-142, losses.py:10-49)
@pytest.mark.parametrize("bs,C,forget", [(300, 7, 0.2), (512, 47, 0.5), (64, 3, 0.0), (1000, 40, 0.35)])
def test_ct_loss_matches_oracle(dev, bs, C, forget):
    """The drop-in CTLoss (device-side ranking) against the line-by-line restatement of the reference's CTLoss:
    both exchanged losses, both pure ratios, the four index lists (bit-exact), and the gradients w.r.t. both logits."""
    from noise_gnn_b200 import CTLoss
    from oracle import ct_oracle
    g = torch.Generator().manual_seed(bs + C)
    y1 = torch.randn(bs, C, generator=g, dtype=torch.float64)
    y2 = torch.randn(bs, C, generator=g, dtype=torch.float64)
    # fp32-representable inputs so both sides rank the very same numbers
    y1, y2 = y1.float().double().requires_grad_(True), y2.float().double().requires_grad_(True)
    yn = torch.randint(0, C, (bs,), generator=g)
    N = 5 * bs
    ind = torch.randperm(N, generator=g)[:bs + 50]                  # batch.n_id: seeds first, more nodes after
    noise_or_not = torch.rand(N, generator=g) < 0.7
    want = ct_oracle.ct_loss(y1, y2, yn, forget, ind, noise_or_not)
    (want[0] * 1.5 + want[1] * 0.5).backward()
    a1 = y1.detach().float().to(dev).requires_grad_(True)
    a2 = y2.detach().float().to(dev).requires_grad_(True)
    got = CTLoss(dev)(a1, a2, yn.to(dev), forget, ind.to(dev), noise_or_not.to(dev))
    (got[0] * 1.5 + got[1] * 0.5).backward()
    # fp32 vs fp64 per-sample losses can order near-ties differently; compare the selections as sets when they differ in order
    for k in (4, 5, 6, 7):
        assert sorted(got[k].cpu().tolist()) == sorted(want[k].tolist()), k
    assert rel_err(got[0].detach(), want[0].detach()) < 1e-5 and rel_err(got[1].detach(), want[1].detach()) < 1e-5
    assert abs(float(got[2]) - float(want[2])) < 1e-6 and abs(float(got[3]) - float(want[3])) < 1e-6
    assert rel_err(a1.grad, y1.grad) < 1e-5 and rel_err(a2.grad, y2.grad) < 1e-5


def test_coteaching_step_matches_oracle(dev):
    """CoTeachingTrainer.train_step (2 x ngnn_sage_forward, ngnn_ct_loss, 2 x ngnn_sage_backward, 2 x Adam) against two
    fp64 oracle networks + the oracle CTLoss on the identical sampled block: both exchanged losses and every parameter
    gradient of both networks."""
    from noise_gnn_b200 import NeighborLoader, SAGE
    from noise_gnn_b200.synthetic import make_dataset
    from noise_gnn_b200.train import CoTeachingTrainer
    from oracle import ct_oracle
    data, sh, train_idx = make_dataset("arxiv", scale=0.02, device="cpu", noise_type="sym", noise_rate=0.3)
    data.clean = (data.yhn.view(-1) == data.y.view(-1))
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=[10, 5], batch_size=128, shuffle=True, seed=1232)
    batch = next(iter(loader))
    bs = batch.batch_size
    refs, nets = [], []
    for s in (1, 2):
        torch.manual_seed(s)
        ref = sage_oracle.SAGERef(sh.features, 64, sh.classes, 3, dropout=0.0, dtype=torch.float64)
        net = SAGE(sh.features, 64, sh.classes, 3, dropout=0.0).to(dev)
        net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
        net.train()
        refs.append(ref); nets.append(net)
    x_cpu, ei_cpu = batch.x.cpu().double(), batch.edge_index.cpu()
    yhn = batch.yhn[:bs].view(-1).cpu()
    o1, o2 = refs[0](x_cpu, ei_cpu)[:bs], refs[1](x_cpu, ei_cpu)[:bs]
    forget = 0.25
    want = ct_oracle.ct_loss(o1, o2, yhn, forget, batch.n_id.cpu(), data.clean)
    want[0].backward()
    want[1].backward()
    ct = CoTeachingTrainer(nets[0], nets[1], lr=1e-3)
    before = [t.buckets.param.clone() for t in (ct.t1, ct.t2)]
    ct.train_step(batch, forget, clean_attr="clean")
    st = ct.read_stats()
    assert abs(st[0] - float(want[0])) < 1e-5 * max(1.0, abs(float(want[0])))
    assert abs(st[1] - float(want[1])) < 1e-5 * max(1.0, abs(float(want[1])))
    assert abs(st[4] - float(want[2])) < 1e-6 and abs(st[5] - float(want[3])) < 1e-6
    y = batch.y[:bs].view(-1).cpu()
    assert int(st[2]) == int((o1.argmax(-1) == y).sum()) and int(st[3]) == int((o2.argmax(-1) == y).sum())
    for net, ref in zip(nets, refs):
        for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            assert rel_err(p.grad, q.grad) < 2e-5, k
    for t, b in zip((ct.t1, ct.t2), before):
        assert not torch.equal(t.buckets.param, b)


def test_hot_rows_first_table_addresses_the_same_rows(dev):
    """The loader's second, in-degree-sorted copy of the feature table and the per-block table indices: x_hot[remap[g]]
    is row g, (col_table, n_table) are remap of (col_global, n_id), the hot prefix holds the highest-degree nodes, and
    the fused step gives the same loss / gradients with and without it (bitwise: same rows, same order)."""
    from noise_gnn_b200 import NeighborLoader, SAGE
    from noise_gnn_b200.synthetic import make_dataset
    from noise_gnn_b200.train import Trainer
    data, sh, train_idx = make_dataset("arxiv", scale=0.05, device="cpu", noise_type="sym", noise_rate=0.3)
    kw = dict(input_nodes=train_idx, num_neighbors=[10, 5], batch_size=128, shuffle=True, seed=1232)
    hot = NeighborLoader(data, hot_feature_bytes=64 * 1024, **kw)
    plain = NeighborLoader(data, hot_feature_bytes=0, **kw)
    assert plain.x_hot is None and hot.x_hot is not None and 0 < hot.hot_rows < data.num_nodes
    remap = hot.remap.long()
    assert torch.equal(torch.sort(remap).values, torch.arange(data.num_nodes, device=remap.device))
    assert torch.equal(hot.x_hot[remap], hot.x)
    deg = (hot.colptr[1:] - hot.colptr[:-1]).long()
    inv = torch.empty_like(remap); inv[remap] = torch.arange(data.num_nodes, device=remap.device)
    assert bool((deg[inv][:-1] >= deg[inv][1:]).all())                   # table order = descending in-degree
    bh, bp = next(iter(hot)), next(iter(plain))
    assert torch.equal(bh.block.col_global, bp.block.col_global) and torch.equal(bh.block.n_id, bp.block.n_id)
    assert torch.equal(bh.block.col_table.long(), remap[bh.block.col_global.long()])
    assert torch.equal(bh.block.n_table.long(), remap[bh.block.n_id.long()])
    assert bp.block.col_table is None
    grads, losses = [], []
    for batch in (bh, bp):
        torch.manual_seed(7)
        net = SAGE(sh.features, 64, sh.classes, 3, dropout=0.0).to(dev)
        tr = Trainer(net, lr=1e-3)
        tr.forward_backward(batch)
        grads.append(tr.buckets.grad.clone()); losses.append(tr.read_stats()[0])
    assert losses[0] == losses[1] and torch.equal(grads[0], grads[1])


def test_ctloss_two_backwards_two_optimizers_like_the_reference(dev):
    """The reference's train_ct (src/pipeline.py:127-133) runs ``loss_1.backward(); optimizer1.step()`` and THEN
    ``loss_2.backward(); optimizer2.step()`` on the two losses of one CTLoss call: each loss must carry its own autograd
    node (a shared node would have freed its saved tensors at the first backward)."""
    from noise_gnn_b200 import CTLoss
    from oracle import ct_oracle
    g = torch.Generator().manual_seed(3)
    bs, C, forget = 200, 7, 0.3
    w1 = torch.randn(16, C, generator=g).double().requires_grad_(True)
    w2 = torch.randn(16, C, generator=g).double().requires_grad_(True)
    x = torch.randn(bs, 16, generator=g).double()
    yn = torch.randint(0, C, (bs,), generator=g)
    ind = torch.arange(bs)
    clean = torch.rand(bs, generator=g) < 0.7
    want = ct_oracle.ct_loss(x @ w1, x @ w2, yn, forget, ind, clean)
    want[0].backward(); want[1].backward()
    a1 = w1.detach().float().to(dev).requires_grad_(True)
    a2 = w2.detach().float().to(dev).requires_grad_(True)
    opt1, opt2 = torch.optim.SGD([a1], lr=0.1), torch.optim.SGD([a2], lr=0.1)
    xd = x.float().to(dev)
    got = CTLoss(dev)(xd @ a1, xd @ a2, yn.to(dev), forget, ind.to(dev), clean.to(dev))
    opt1.zero_grad(); got[0].backward(); opt1.step()
    g1 = a1.grad.clone()
    opt2.zero_grad(); got[1].backward(); opt2.step()              # second backward through the same CTLoss call
    assert rel_err(g1, w1.grad) < 1e-5 and rel_err(a2.grad, w2.grad) < 1e-5
    assert rel_err(got[0].detach(), want[0].detach()) < 1e-5 and rel_err(got[1].detach(), want[1].detach()) < 1e-5


def test_inference_rejects_loaders_that_do_not_walk_all_nodes_in_order(dev):
    """Layer outputs are indexed by global node id by the next layer (reference sage.py:50): a subset or shuffled loader
    would silently misalign rows, so it is refused."""
    from noise_gnn_b200 import NeighborLoader
    data, sh, loader, ref, net = _setup(dev, 3, [10, 5], dropout=0.0, scale=0.01)
    net.eval()
    with pytest.raises(ValueError, match="all nodes"):
        net.inference(data.x, NeighborLoader(data, input_nodes=torch.arange(100), num_neighbors=[10, 5], batch_size=64), dev)
    with pytest.raises(ValueError, match="unshuffled"):
        net.inference(data.x, NeighborLoader(data, input_nodes=None, num_neighbors=[10, 5], batch_size=64, shuffle=True), dev)
