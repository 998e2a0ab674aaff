"""GPU tests (run with -m gpu) of the replayed step: device-side extents, the sampler-built transposes and the CUDA-graph
epoch loop (Trainer.train_epoch) against the host-extent path and the CPU oracle on identical blocks."""
import numpy as np
import pytest
import torch

from oracle import sage_oracle, sampler, structure

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp(min=1e-30))


@pytest.fixture(scope="module")
def dev(cuda_device):
    return cuda_device


def _problem(dev, fan, bs, dropout, L=3, hidden=64, scale=0.05, seeds=None, name="arxiv"):
    from noise_gnn_b200 import NeighborLoader, SAGE
    from noise_gnn_b200.synthetic import make_dataset
    data, sh, train_idx = make_dataset(name, scale=scale, device="cpu", noise_type="sym", noise_rate=0.3)
    if seeds is not None:
        train_idx = train_idx[:seeds]
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=fan, batch_size=bs, shuffle=True, seed=1232)
    torch.manual_seed(1232)
    ref = sage_oracle.SAGERef(sh.features, hidden, sh.classes, L, dropout=dropout, dtype=torch.float64)
    net = SAGE(sh.features, hidden, sh.classes, L, dropout=dropout).to(dev)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    return data, sh, loader, ref, net


@pytest.mark.parametrize("fan,T", [([15, 10, 5], 2), ([10, 5], 2), ([4, 4, 4, 4], 4), ([25], 1), ([40, 3], 2)])
def test_sampler_transposes_match_the_stable_oracle_transpose(dev, fan, T):
    """ngnn_sample_block_ex's CSC transposes of the hop prefixes (histogram -> scan -> scatter -> per-row sort) against
    oracle/structure.csr_transpose (stable argsort) of the same prefix of the oracle's block; the block itself bit-exact."""
    data, sh, loader, _, _ = _problem(dev, fan, 128, 0.0)
    loader.transpose_hops = T
    cs = sampler.CSampler(loader.colptr.cpu().numpy(), loader.row.cpu().numpy())
    order = loader.epoch_permutation(0)
    for b, batch in enumerate(loader):
        if b >= 3:
            break
        want = cs.sample(loader.batch_seeds(order, b).numpy(), fan, seed=1232, epoch=0, batch_idx=b)
        blk = batch.block
        assert np.array_equal(blk.rowptr.cpu().numpy(), want.rowptr) and np.array_equal(blk.col.cpu().numpy(), want.col)
        assert np.array_equal(blk.n_id.cpu().numpy(), want.n_id)
        assert len(blk._t) == T
        for h in range(1, T + 1):
            e_lim, n_cols = int(want.edge_counts[h]), int(want.node_counts[h])
            ct, rt = blk._t[(e_lim, n_cols)]
            oct_, ort, _ = structure.csr_transpose(want.rowptr, want.col, want.n, e_lim)
            assert np.array_equal(ct.cpu().numpy(), oct_[: n_cols + 1]), (b, h)
            assert int(oct_[n_cols]) == e_lim
            assert np.array_equal(rt.cpu().numpy(), ort), (b, h)


def test_sampler_transposes_with_hub_rows(dev):
    """A star-shaped graph: one source feeds every destination, so its transposed row has thousands of entries (the CTA
    bitonic sort and, beyond 4096 entries, the in-place heapsort)."""
    from noise_gnn_b200 import Data, NeighborLoader
    n = 9000
    hub_src = torch.zeros(n - 1, dtype=torch.long)
    others = torch.arange(1, n)
    g = torch.Generator().manual_seed(0)
    extra_src = torch.randint(1, n, (3 * n,), generator=g)
    extra_dst = torch.randint(1, n, (3 * n,), generator=g)
    ei = torch.stack([torch.cat([hub_src, extra_src]), torch.cat([others, extra_dst])])
    data = Data(x=torch.randn(n, 8), edge_index=ei, y=torch.zeros(n, dtype=torch.long))
    for bs in (2000, 6000):                                   # hub row of ~2000 (bitonic) and ~6000 (heapsort) entries
        loader = NeighborLoader(data, input_nodes=torch.arange(1, bs + 1), num_neighbors=[4, 2], batch_size=bs, shuffle=False)
        loader.transpose_hops = 2
        batch = next(iter(loader))
        blk = batch.block
        rowptr, col = blk.rowptr.cpu().numpy(), blk.col.cpu().numpy()
        for h in (1, 2):
            e_lim, n_cols = blk.hop_edges[h], blk.hop_nodes[h]
            ct, rt = blk._t[(e_lim, n_cols)]
            oct_, ort, _ = structure.csr_transpose(rowptr, col, blk.n_rows, e_lim)
            assert np.array_equal(ct.cpu().numpy(), oct_[: n_cols + 1]) and np.array_equal(rt.cpu().numpy(), ort)
        lens = (blk._t[(blk.hop_edges[1], blk.hop_nodes[1])][0][1:] - blk._t[(blk.hop_edges[1], blk.hop_nodes[1])][0][:-1])
        assert int(lens.max()) > (4096 if bs == 6000 else 1000)


@pytest.mark.parametrize("L,fan,dropout,name", [(3, [15, 10, 5], 0.0, "arxiv"), (3, [10, 5], 0.5, "arxiv"), (2, [10, 5], 0.0, "arxiv"),
                                                 (2, [7], 0.5, "arxiv"), (4, [5, 5], 0.0, "arxiv"),
                                                 (2, [10, 5], 0.5, "cora"), (2, [10, 5], 0.0, "computers")])
def test_device_extent_step_equals_host_extent_step(dev, L, fan, dropout, name):
    """ngnn_sage_step with the extents left on the device (worst-case launches, counts read by the kernels, sampler-built
    transposes, dropout offset from the control words) against the same step with host extents: loss and every gradient."""
    import ctypes
    from noise_gnn_b200 import _lib, ops
    from noise_gnn_b200.train import Trainer
    # (cora: F = 1433, computers: F = 767 — contractions too long for one tensor-memory accumulation chain, so layer 1 takes the
    #  SIMT GEMMs, which read the device-side extents as well)
    data, sh, loader, ref, net = _problem(dev, fan, 64, dropout, L=L, name=name, scale=0.05 if name == "arxiv" else 0.5)
    net.train()
    tr = Trainer(net, lr=1e-3)
    loader.transpose_hops = min(L - 1, len(fan))
    batch = next(iter(loader))
    tr.forward_backward(batch)                                  # host extents; dropout offset = steps * L = L
    loss_h, corr_h = tr.read_stats()
    g_host = tr.buckets.grad.clone()
    # the same block through the device-extent path
    gs = tr._graph_state(loader, "yhn", "y")
    slot = gs["slots"][0]
    seeds = loader.batch_seeds(loader.epoch_permutation(0), 0)
    _lib.call("ngnn_step_ctl_set", ops._ptr(slot.ctl), 0, 0, 1 * L, 1.0, ops._stream())
    loader.launch_sample(slot, seeds.pin_memory(), len(seeds), 0, 0, use_ctl=True, transposes=gs["T"])
    assert torch.equal(slot.counts.cpu(), batch._slot.counts.cpu())
    tr.reset_stats()
    tr.buckets.grad.zero_()
    tr._enqueue_agg1(gs, 0, len(seeds))                       # layer 1's aggregation is its own call on this path
    bd = tr._slot_desc(gs, 0, len(seeds))
    ms, arena, table = gs["ms"], gs["arena"], gs["table"]
    _lib.call("ngnn_sage_step", ctypes.byref(ms), ops._ptr(tr.buckets.param), ops._ptr(tr.buckets.grad), ctypes.byref(bd),
              tr._max_nodes, tr._max_edges, ops._ptr(table), table.stride(0), ops._ptr(gs["tgt"]), ops._ptr(gs["lab"]),
              tr._drop_seed(), 0, ops._ptr(tr.stats), None, 0, ops._ptr(arena), arena.numel(), ops._stream())
    loss_d, corr_d = tr.read_stats()
    assert abs(loss_d - loss_h) < 1e-6 * max(1.0, abs(loss_h)) and corr_d == corr_h
    assert rel_err(tr.buckets.grad, g_host) < 1e-6          # (the weight-gradient slices differ, so not bit for bit)


@pytest.mark.parametrize("dropout", [0.0, 0.5])
def test_train_epoch_on_the_captured_step_tracks_eager_and_oracle(dev, dropout):
    """Trainer.train_epoch (two eager steps, then CUDA-graph replays, then an eager ragged tail) against (a) the eager
    batch-by-batch loop on an identically initialised model — same blocks, same dropout streams — and (b), without dropout,
    the CPU oracle trained on the blocks of the sequential C sampler."""
    from noise_gnn_b200.train import Trainer
    fan, bs, L = [10, 5], 32, 3
    n_seeds = 32 * 9 + 5                                       # 10 steps, the last one ragged (5 seeds)
    data, sh, loader, ref, net = _problem(dev, fan, bs, dropout, L=L, seeds=n_seeds)
    _, _, loader2, _, net2 = _problem(dev, fan, bs, dropout, L=L, seeds=n_seeds)
    net2.drop_seed = net.drop_seed
    net.train(); net2.train()
    tr = Trainer(net, lr=1e-3)
    mean_loss, correct, log = tr.train_epoch(loader, epoch=0)
    assert tr._gs["graphs"][0] is not None and tr._gs["graphs"][1] is not None     # the replayed path ran
    steps = len(loader)
    assert steps == 10 and log.shape == (steps, 2)
    losses = np.diff(np.concatenate([[0.0], log[:, 0].numpy()]))
    # (a) eager loop
    tr2 = Trainer(net2, lr=1e-3, use_graph=False)
    loader2.epoch = 0
    eager = []
    for batch in loader2:
        tr2.reset_stats()
        tr2.train_step(batch)
        eager.append(tr2.read_stats()[0])
    assert np.allclose(losses, eager, rtol=2e-5, atol=1e-6), (np.abs(losses - np.array(eager)) / np.array(eager)).tolist()
    assert abs(mean_loss - float(np.mean(eager))) < 1e-5
    # (Adam moves a parameter whose gradient is zero to rounding by +-lr per step whatever its size, so the two runs' parameters
    #  agree to a few lr, not to rounding; the losses above are the sensitive check)
    assert rel_err(tr.buckets.param, tr2.buckets.param) < 5e-2
    # a second epoch replays the captured graphs on new blocks
    m2, _, _ = tr.train_epoch(loader, epoch=1)
    loader2.epoch = 1
    tr2.reset_stats()
    for batch in loader2:
        tr2.train_step(batch)
    assert abs(m2 - tr2.read_stats()[0] / steps) < 1e-4
    if dropout == 0.0:
        # (b) the CPU oracle on the C sampler's blocks of epoch 0
        torch.manual_seed(1232)
        ref = ref.float()
        opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
        cs = sampler.CSampler(loader.colptr.cpu().numpy(), loader.row.cpu().numpy())
        order = loader.epoch_permutation(0)
        want = []
        ref.train()
        for b in range(steps):
            sd = loader.batch_seeds(order, b)
            blk = cs.sample(sd.numpy(), fan, seed=1232, epoch=0, batch_idx=b)
            n_id = torch.from_numpy(blk.n_id.astype(np.int64))
            ei = torch.from_numpy(structure.csr_to_coo(blk.rowptr, blk.col))
            l, _ = sage_oracle.train_step(ref, opt, data.x[n_id], ei, data.y[n_id], data.yhn[n_id], len(sd))
            want.append(l)
        assert np.allclose(losses, want, rtol=2e-4, atol=1e-5), (losses, want)


def test_train_epoch_max_steps_and_resident_seeds(dev):
    from noise_gnn_b200.train import Trainer
    data, sh, loader, ref, net = _problem(dev, [10, 5], 32, 0.0, seeds=32 * 12)
    _, _, _, _, net2 = _problem(dev, [10, 5], 32, 0.0, seeds=32 * 12)
    net.train(); net2.train()
    a = Trainer(net, lr=1e-3)
    b = Trainer(net2, lr=1e-3)
    la, ca, _ = a.train_epoch(loader, epoch=0, max_steps=7)
    lb, cb, _ = b.train_epoch(loader, epoch=0, max_steps=7, seeds_resident=True, log_every_step=False)
    assert la == lb and ca == cb and torch.equal(a.buckets.param, b.buckets.param)


def test_full_batch_config_replays_across_epochs(dev):
    """The reference's full-batch configs (pubmed: 60 train seeds, cora: 140, batch_size 512): one batch per epoch, shorter than
    batch_size.  Every round is 'full' (all the seeds), so run_steps replays the captured step across epoch boundaries; the
    losses equal the eager batch-by-batch loop's."""
    from noise_gnn_b200.train import Trainer
    fan, L, epochs = [10, 5], 3, 8
    data, sh, loader, ref, net = _problem(dev, fan, 512, 0.0, L=L, name="pubmed", scale=1.0)
    _, _, loader2, _, net2 = _problem(dev, fan, 512, 0.0, L=L, name="pubmed", scale=1.0)
    assert len(loader) == 1 and loader.sharder.full_len == 60
    net.train(); net2.train()
    tr = Trainer(net, lr=1e-3)
    total, correct, log = tr.run_steps(loader, epochs, start_epoch=0)
    assert tr.graph_replays == epochs - 3                      # two eager steps, then replays; the last step has no next block
    losses = np.diff(np.concatenate([[0.0], log[:, 0].numpy()]))
    tr2 = Trainer(net2, lr=1e-3, use_graph=False)
    eager = []
    for ep in range(epochs):
        loader2.epoch = ep
        for batch in loader2:
            assert batch.batch_size == 60
            tr2.reset_stats()
            tr2.train_step(batch)
            eager.append(tr2.read_stats()[0])
    assert np.allclose(losses, eager, rtol=2e-5, atol=1e-6), (losses, eager)
