"""GPU parity (run with -m gpu) of the SAGEPL extras: fused adding_noise (reference
src/models/layers/sagePL.py:41-49) forward + backward, the SAGEPL network with layer-1 data gradients, shuffle_pos
(src/utils/augmentation.py:88-102) bit-exact against its Philox twin, and sampling WITH replacement
(src/pipeline_contrast.py:249-279) through the fused step."""
import numpy as np
import pytest
import torch

from oracle import sage_oracle, sagepl_oracle, sampler

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    assert got.shape == want.shape
    return float((got - want).abs().max() / want.abs().max().clamp(min=1e-30))


@pytest.fixture(scope="module")
def dev(cuda_device):
    return cuda_device


@pytest.mark.parametrize("n,N,F,with_ids", [(500, 3000, 100, True), (257, 257, 128, False), (1000, 5000, 1433, True), (64, 64, 7, False)])
def test_adding_noise_forward_backward(dev, n, N, F, with_ids):
    from noise_gnn_b200 import ops
    g = torch.Generator().manual_seed(n + F)
    x64 = torch.randn(n, F, generator=g, dtype=torch.float64)
    x64[3, :5] = 0.0                                             # sign(0) = 0
    noise64 = torch.randn(N, F, generator=g, dtype=torch.float64)
    noise64[7] = 0.0                                             # a zero row: F.normalize divides by eps
    n_id = torch.randperm(N, generator=g)[:n] if with_ids else None
    if with_ids:
        n_id[0] = 7
    gout = torch.randn(n, F, generator=g, dtype=torch.float64)
    xr, nr = x64.clone().requires_grad_(True), noise64.clone().requires_grad_(True)
    want = sagepl_oracle.adding_noise(xr, nr, 0.3, n_id)
    (want * gout).sum().backward()
    xg = x64.float().to(dev).requires_grad_(True)
    ng = noise64.float().to(dev).requires_grad_(True)
    idx = None if n_id is None else n_id.int().to(dev)
    got = ops.NoiseAddFunction.apply(xg, ng, idx, 0.3, n_id is None)
    (got * gout.float().to(dev)).sum().backward()
    assert rel_err(got, want) < 1e-6
    assert rel_err(xg.grad, xr.grad) < 1e-6
    live = torch.ones(N, dtype=torch.bool)
    live[7] = False                                              # the zero row's gradient is rate/eps * g: compare it separately
    assert rel_err(ng.grad.cpu()[live], nr.grad[live]) < 1e-5
    if with_ids:
        assert rel_err(ng.grad.cpu()[7], nr.grad[7]) < 1e-5
        untouched = torch.ones(N, dtype=torch.bool); untouched[n_id] = False
        assert float(ng.grad.cpu()[untouched].abs().max()) == 0.0


@pytest.mark.parametrize("n,F,prob", [(300, 100, 0.1), (50, 1433, 0.3), (1000, 128, 0.5), (10, 16, 0.0), (7, 2048, 1.0), (33, 100, 0.01)])
def test_shuffle_rows_bit_exact_vs_the_philox_twin(dev, n, F, prob):
    from noise_gnn_b200 import ops
    g = torch.Generator().manual_seed(F)
    x = torch.randn(n, F, generator=g)
    k = int(F * prob)
    got = ops.shuffle_rows(x.to(dev), k, seed=1232, offset=5).cpu().numpy()
    want = sagepl_oracle.shuffle_rows(x.numpy(), k, seed=1232, offset=5)
    assert np.array_equal(got, want)
    # validity (the law of augmentation.py:88-102): every row is a permutation of itself that moves at most k positions
    assert np.array_equal(np.sort(got, axis=1), np.sort(x.numpy(), axis=1))
    assert int((got != x.numpy()).sum(axis=1).max()) <= k
    if k > 8 and n >= 50:
        assert (got != x.numpy()).any()
        other = ops.shuffle_rows(x.to(dev), k, seed=1232, offset=6).cpu().numpy()
        assert (other != got).any()                              # a new call offset gives new draws


def test_shuffle_pos_drop_in(dev):
    from noise_gnn_b200 import shuffle_pos
    x = torch.randn(200, 100)
    a = shuffle_pos(x, device=dev, prob=0.1)
    b = shuffle_pos(x.to(dev), device=dev, prob=0.1)
    assert a.is_cuda and a.shape == x.shape and not a.requires_grad
    assert torch.equal(torch.sort(a.cpu(), dim=1).values, torch.sort(x, dim=1).values)
    assert int((a.cpu() != x).sum(1).max()) <= 10 and not torch.equal(a, b)


def test_sagepl_network_forward_backward_vs_oracle(dev):
    """The six outputs of SAGEPL.forward and the gradients of every parameter — INCLUDING the noise table, which needs the
    data gradient of layer 1 — on an identical sampled block (reference loop: src/pipeline_test.py:123-125)."""
    from noise_gnn_b200 import NeighborLoader, SAGEPL
    from noise_gnn_b200.synthetic import make_dataset
    data, sh, train_idx = make_dataset("arxiv", scale=0.02, device="cpu")
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=[10, 5], batch_size=64, shuffle=True)
    batch = next(iter(loader))
    N = data.num_nodes
    torch.manual_seed(3)
    ref = sagepl_oracle.SAGEPLRef(sh.features, 64, sh.classes, 3, N, dropout=0.0, dtype=torch.float64)
    net = SAGEPL(sh.features, 64, sh.classes, 3, N, dropout=0.0).to(dev)
    assert sorted(net.state_dict()) == sorted(ref.state_dict())
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    x_cpu, ei_cpu, nid_cpu = batch.x.cpu().double(), batch.edge_index.cpu(), batch.n_id.cpu()
    want = ref(x_cpu, ei_cpu, noise_rate=0.2, n_id=nid_cpu)
    got = net(batch.x, batch.edge_index, noise_rate=0.2, n_id=batch.n_id)
    for a, b in zip(got, want):
        assert rel_err(a, b) < 1e-5
    bs = batch.batch_size
    tgt = batch.y[:bs].view(-1).cpu()
    ce = torch.nn.functional.cross_entropy
    (ce(want[2][:bs], tgt) + ce(want[5][:bs], tgt) + want[3].pow(2).mean()).backward()
    (ce(got[2][:bs], tgt.to(dev)) + ce(got[5][:bs], tgt.to(dev)) + got[3].pow(2).mean()).backward()
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, q.grad) < 2e-5, k
    assert float(net.noise.grad.abs().max()) > 0


def test_fused_step_with_replacement_sampling(dev):
    """replace=True (src/pipeline_contrast.py:249-279): duplicate edges in a row are counted each time by the mean and
    by the transposed backward; block bit-exact vs the C oracle, fused step vs the untrimmed fp64 oracle."""
    from noise_gnn_b200 import NeighborLoader, SAGE
    from noise_gnn_b200.synthetic import make_dataset
    from noise_gnn_b200.train import Trainer
    data, sh, train_idx = make_dataset("arxiv", scale=0.02, device="cpu", noise_type="sym", noise_rate=0.3)
    fan = [6, 4]
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=fan, batch_size=64, shuffle=True, replace=True, seed=1232)
    torch.manual_seed(1)
    ref = sage_oracle.SAGERef(sh.features, 64, sh.classes, 3, dropout=0.0, dtype=torch.float64)
    net = SAGE(sh.features, 64, sh.classes, 3, dropout=0.0).to(dev)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    tr = Trainer(net)
    tr.forward_backward(next(iter(loader)))                    # sets the loader's transposes for the next blocks
    loader.epoch = 0
    batch = next(iter(loader))
    want = sampler.sample_block(loader.colptr.cpu().numpy(), loader.row.cpu().numpy(),
                                loader.batch_seeds(loader.epoch_permutation(0), 0).numpy(), fan, True, seed=1232, epoch=0, batch_idx=0)
    assert np.array_equal(batch.block.col.cpu().numpy(), want.col) and np.array_equal(batch.block.rowptr.cpu().numpy(), want.rowptr)
    ei = batch.edge_index.cpu()
    assert int((ei[:, 1:] == ei[:, :-1]).all(0).sum()) > 0       # the block really holds duplicated edges
    bs = batch.batch_size
    out_ref = ref(batch.x.cpu().double(), ei)[:bs]
    torch.nn.functional.cross_entropy(out_ref, batch.yhn[:bs].view(-1).cpu()).backward()
    tr.reset_stats()
    logits = tr.forward_backward(batch, want_logits=True)
    assert rel_err(logits, out_ref) < 1e-5
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, q.grad) < 2e-5, k
