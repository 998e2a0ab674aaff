"""GPU parity tests (run with -m gpu on a B200): every CUDA kernel, through the C ABI, against the CPU oracle
on identical seeded inputs.  Integer / index outputs are compared bit-exactly; fp32 outputs against an fp64
oracle at the north-star tolerance (1e-5 relative to the tensor's scale)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import philox, sage_oracle, sampler, structure

pytestmark = pytest.mark.gpu

GOLD = json.loads((Path(__file__).parent / "golden" / "sage_known_answer.json").read_text())
RTOL = 1e-5   # BASELINE.json north_star: fp32 forward/backward within 1e-5 relative


def rel_err(got: torch.Tensor, want: torch.Tensor) -> float:
    want = want.detach().double().cpu()
    got = got.detach().double().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.numel() == 0:
        return 0.0
    return float((got - want).abs().max().detach() / want.abs().max().clamp(min=1e-30).detach())


@pytest.fixture(scope="module")
def dev(cuda_device):
    from noise_gnn_b200 import _lib
    assert _lib.load().ngnn_device_supported() == 1, "these kernels are sm_100a only"
    return cuda_device


def random_coo(n, e, seed, sort=False):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    if sort:
        ei = ei[:, torch.argsort(ei[1], stable=True)]
    return ei


# ------------------------------------------------------------------ structure (bit exact)
@pytest.mark.parametrize("n,e", [(6, 9), (1, 0), (50, 1), (1000, 20000), (70000, 300000)])
def test_coo_to_csr_and_transpose_bit_exact(dev, n, e):
    from noise_gnn_b200 import ops
    ei = random_coo(n, e, seed=n + e)
    if n == 6:
        ei = torch.tensor([GOLD["src"], GOLD["dst"]])
    blk = ops.coo_to_csr(ei.to(dev), n)
    rowptr, col, perm = structure.coo_to_csr(ei[0].numpy(), ei[1].numpy(), n)
    assert np.array_equal(blk.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(blk.col.cpu().numpy(), col)
    assert np.array_equal(blk.perm.cpu().numpy(), perm)
    for e_lim in {e, e // 2}:
        ct, rt, pt = ops.csr_transpose(blk.rowptr, blk.col, n, e_lim, n)
        oct_, ort, opt = structure.csr_transpose(rowptr, col, n, e_lim)
        assert np.array_equal(ct.cpu().numpy(), oct_)
        assert np.array_equal(rt.cpu().numpy(), ort)
        assert np.array_equal(pt.cpu().numpy(), opt)
    coo = ops.csr_to_coo(blk.rowptr, blk.col, n, e)
    assert np.array_equal(coo.cpu().numpy(), structure.csr_to_coo(rowptr, col))


def test_gather_rows_bit_exact(dev):
    from noise_gnn_b200 import ops
    for F in (1, 7, 100, 128, 1433):
        t = torch.randn(300, F)
        idx = torch.randint(0, 300, (1000,), dtype=torch.int32)
        out = ops.gather_rows(t.to(dev), idx.to(dev))
        assert torch.equal(out.cpu(), t[idx.long()])


# ------------------------------------------------------------------ K-AGG
@pytest.mark.parametrize("F", [4, 32, 64, 100, 128, 256, 500, 512, 1024, 1100, 3, 767, 1433])
def test_agg_fwd_matches_oracle(dev, F):
    from noise_gnn_b200 import ops
    n, e = 700, 6000
    ei = random_coo(n, e, seed=F)
    ei[1, ei[1] == 5] = 6                      # node 5: zero in-degree
    ei[:, 10] = ei[:, 11]                      # a duplicated edge
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(F + 1))
    want = sage_oracle.mean_aggregate(x.double(), ei)
    blk = ops.coo_to_csr(ei.to(dev), n)
    got = ops.agg_fwd(blk.rowptr, blk.col, x.to(dev), n)
    assert rel_err(got, want) < RTOL
    assert float(got[5].abs().max()) == 0.0
    # prefix of rows only (trimmed layer) + fused root gather
    root_idx = torch.randint(0, n, (n,), dtype=torch.int32)
    m2, r2 = ops.agg_fwd(blk.rowptr, blk.col, x.to(dev), 123, root_idx=root_idx.to(dev))
    assert torch.equal(m2, got[:123])
    assert torch.equal(r2.cpu(), x[root_idx[:123].long()])


def test_agg_fwd_long_rows_and_determinism(dev):
    from noise_gnn_b200 import ops
    n, F = 64, 100
    g = torch.Generator().manual_seed(77)
    ei = torch.stack([torch.randint(0, n, (5000,), generator=g), torch.zeros(5000, dtype=torch.long)])   # one row of degree 5000
    x = torch.randn(n, F, generator=g)
    blk = ops.coo_to_csr(ei.to(dev), n)
    a = ops.agg_fwd(blk.rowptr, blk.col, x.to(dev), n)
    b = ops.agg_fwd(blk.rowptr, blk.col, x.to(dev), n)
    assert torch.equal(a, b), "aggregation is not bitwise reproducible"       # atomic-free, fixed summation order
    # a 5000-term fp32 sum: compare with the error the fp32 CPU oracle itself makes against fp64
    want = sage_oracle.mean_aggregate(x.double(), ei)
    err_ref = rel_err(sage_oracle.mean_aggregate(x, ei), want)
    assert rel_err(a, want) < max(RTOL, 4 * err_ref), (rel_err(a, want), err_ref)


@pytest.mark.parametrize("F", [64, 100, 256, 47])
def test_agg_bwd_matches_oracle(dev, F):
    from noise_gnn_b200 import ops
    n, e, n_dst = 600, 5000, 200
    ei = random_coo(n, e, seed=3 * F, sort=True)
    ei = ei[:, ei[1] < n_dst]                  # destinations are a prefix, sources anywhere
    e = ei.size(1)
    blk = ops.coo_to_csr(ei.to(dev), n)
    dmean = torch.randn(n_dst, F)
    droot = torch.randn(n_dst, F)
    h = torch.randn(n, F)
    # oracle: dx[j] = sum_{(j->i)} dmean[i] (+ droot[j] for j < n_dst), gated by h > 0 with scale 2
    want = torch.zeros(n, F, dtype=torch.float64).index_add_(0, ei[0], dmean.double()[ei[1]])
    want[:n_dst] += droot.double()
    want_gated = torch.where(h > 0, want * 2.0, torch.zeros_like(want))
    ct, rt, _ = ops.csr_transpose(blk.rowptr, blk.col, n, e, n)
    got = ops.agg_bwd(ct, rt, dmean.to(dev), n, dx_root=droot.to(dev), n_root=n_dst)
    assert rel_err(got, want) < RTOL
    got2 = ops.agg_bwd(ct, rt, dmean.to(dev), n, dx_root=droot.to(dev), n_root=n_dst, act_ref=h.to(dev), act_scale=2.0)
    assert rel_err(got2, want_gated) < RTOL


@pytest.mark.parametrize("F", [64, 100, 256, 512, 1024])
def test_agg_hub_rows_are_split_across_the_cta(dev, F):
    """Rows far longer than kLongRow (power-law hubs: un-sampled convolutions, transposed blocks) — the generic kernel's
    CTA-cooperative path: forward with the GCN bias (sum), backward with add rows and gate; against the oracle within the
    fp32 summation-order tolerance, and bitwise reproducible run to run."""
    from noise_gnn_b200 import ops
    n = 700
    g = torch.Generator().manual_seed(F)
    hubs = {3: 5000, 4: 1025, 250: 2500, 699: 3333}        # several per CTA (rows 3, 4) and the last row
    deg = torch.randint(0, 12, (n,), generator=g)
    for r, d in hubs.items():
        deg[r] = d
    dst = torch.repeat_interleave(torch.arange(n), deg)
    src = torch.randint(0, n, (dst.numel(),), generator=g)
    ei = torch.stack([src, dst])
    x = torch.randn(n, F, generator=g)
    bias = torch.randn(F, generator=g)
    blk = ops.coo_to_csr(ei.to(dev), n)
    # forward, sum + bias (GCN propagation; always the generic kernel)
    msg = x.double().index_select(0, src)
    want = torch.zeros(n, F, dtype=torch.float64).index_add_(0, dst, msg) + bias.double()
    ref32 = torch.zeros(n, F).index_add_(0, dst, x.index_select(0, src)) + bias
    tol = max(RTOL, 4 * rel_err(ref32, want))
    a = ops.gcn_agg_fwd(blk.rowptr, blk.col, x.to(dev), n, bias=bias.to(dev))
    b = ops.gcn_agg_fwd(blk.rowptr, blk.col, x.to(dev), n, bias=bias.to(dev))
    assert torch.equal(a, b), "not bitwise reproducible"
    assert rel_err(a, want) < tol
    # backward form on the same structure: add rows on a prefix + gate
    add = torch.randn(300, F, generator=g)
    h = torch.randn(n, F, generator=g)
    want_b = torch.zeros(n, F, dtype=torch.float64).index_add_(0, dst, msg)
    want_b[:300] += add.double()
    want_b = torch.where(h > 0, want_b * 2.0, torch.zeros_like(want_b))
    c = ops.agg_bwd(blk.rowptr, blk.col, x.to(dev), n, dx_root=add.to(dev), n_root=300, act_ref=h.to(dev), act_scale=2.0)
    d = ops.agg_bwd(blk.rowptr, blk.col, x.to(dev), n, dx_root=add.to(dev), n_root=300, act_ref=h.to(dev), act_scale=2.0)
    assert torch.equal(c, d)
    assert rel_err(c, want_b) < tol


# ------------------------------------------------------------------ K-GEMM / K-DGRAD / K-WGRAD
GEMM_SHAPES = [(1, 4, 3), (130, 100, 256), (700, 256, 47), (257, 128, 40), (300, 1433, 7), (90, 767, 10),
               (1000, 500, 3), (513, 256, 256), (64, 512, 7), (300, 64, 512), (4000, 100, 256)]


@pytest.fixture(params=["auto", "ss", "simt"])
def gemm_path(request, dev):
    """Runs a GEMM test three times: automatic dispatch (tcgen05 where the operands are TMA-addressable; A operand in
    tensor memory), tcgen05 with both operands in shared memory ("ss"), and forced SIMT."""
    from noise_gnn_b200 import _lib
    _lib.call("ngnn_set_gemm_path", 1 if request.param == "simt" else 0)
    _lib.call("ngnn_set_tuning", 6, 0 if request.param == "ss" else 1)
    yield "auto" if request.param == "ss" else request.param
    _lib.call("ngnn_set_gemm_path", 0)
    _lib.call("ngnn_set_tuning", 6, 1)


@pytest.mark.parametrize("n,F,O", GEMM_SHAPES)
def test_gemm_fwd_matches_oracle(dev, gemm_path, n, F, O):
    from noise_gnn_b200 import ops
    g = torch.Generator().manual_seed(n + F + O)
    a_l, a_r = torch.randn(n, F, generator=g), torch.randn(n + 5, F, generator=g)
    w_l, w_r = torch.randn(O, F, generator=g) / F ** 0.5, torch.randn(O, F, generator=g) / F ** 0.5
    b = torch.randn(O, generator=g)
    want = a_l.double() @ w_l.double().T + a_r[:n].double() @ w_r.double().T + b.double()
    got, path = ops.gemm_fwd(a_l.to(dev), a_r.to(dev), w_l.to(dev), w_r.to(dev), b.to(dev), n, return_path=True)
    assert path == (1 if (gemm_path == "auto" and F % 4 == 0) else 0)     # tcgen05 path is the one that ran
    assert rel_err(got, want) < RTOL
    got_relu = ops.gemm_fwd(a_l.to(dev), a_r.to(dev), w_l.to(dev), w_r.to(dev), b.to(dev), n, act=1)
    assert rel_err(got_relu, want.clamp(min=0)) < RTOL
    # root-only / no-bias variants
    got_r = ops.gemm_fwd(None, a_r.to(dev), None, w_r.to(dev), None, n)
    assert rel_err(got_r, a_r[:n].double() @ w_r.double().T) < RTOL


def test_gemm_fused_dropout_mask_is_the_philox_oracle_mask(dev, gemm_path):
    from noise_gnn_b200 import ops
    n, F, O, p = 300, 64, 100, 0.5
    a = torch.randn(n, F); w = torch.randn(O, F) / 8; b = torch.randn(O)
    want = (a.double() @ w.double().T + b.double()).clamp(min=0)
    keep = torch.from_numpy(philox.dropout_keep_mask(n, O, p, seed=1232, offset=77))
    want = torch.where(keep, want / (1 - p), torch.zeros_like(want))
    got = ops.gemm_fwd(a.to(dev), None, w.to(dev), None, b.to(dev), n, act=1, drop_p=p, seed=1232, offset=77)
    assert torch.equal((got != 0).cpu(), (want != 0))           # identical mask (given relu support)
    assert rel_err(got, want) < RTOL


@pytest.mark.parametrize("n,F,O", GEMM_SHAPES)
def test_dgrad_and_wgrad_match_oracle(dev, gemm_path, n, F, O):
    from noise_gnn_b200 import ops
    g = torch.Generator().manual_seed(7 * n + F + O)
    dy = torch.randn(n, O, generator=g)
    a_l, a_r = torch.randn(n, F, generator=g), torch.randn(n, F, generator=g)
    w_l, w_r = torch.randn(O, F, generator=g), torch.randn(O, F, generator=g)
    deg = torch.randint(0, 6, (n,), generator=g)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)]).int()
    dmean, droot = ops.dgrad(dy.to(dev), w_l.to(dev), w_r.to(dev), rowptr.to(dev), n)
    inv = 1.0 / deg.clamp(min=1).double()
    assert rel_err(dmean, (dy.double() @ w_l.double()) * inv[:, None]) < RTOL
    assert rel_err(droot, dy.double() @ w_r.double()) < RTOL
    dw_l, dw_r, db = ops.wgrad(dy.to(dev), a_l.to(dev), a_r.to(dev), n, F)
    assert rel_err(dw_l, dy.double().T @ a_l.double()) < RTOL
    assert rel_err(dw_r, dy.double().T @ a_r.double()) < RTOL
    assert rel_err(db, dy.double().sum(0)) < RTOL
    # accumulate adds into the outputs; repeat is bitwise reproducible (fixed split order)
    dw2, _, db2 = ops.wgrad(dy.to(dev), a_l.to(dev), None, n, F, dw_l=dw_l.clone(), db=db.clone(), accumulate=True,
                            want_r=False)
    assert rel_err(dw2, 2 * (dy.double().T @ a_l.double())) < RTOL
    dw3, _, _ = ops.wgrad(dy.to(dev), a_l.to(dev), a_r.to(dev), n, F)
    assert torch.equal(dw3, dw_l)


def test_wgrad_long_reduction(dev):
    from noise_gnn_b200 import ops
    n, F, O = 90000, 100, 47
    g = torch.Generator().manual_seed(5)
    dy, a = torch.randn(n, O, generator=g), torch.randn(n, F, generator=g)
    dw_l, _, db = ops.wgrad(dy.to(dev), a.to(dev), None, n, F, want_r=False)
    assert rel_err(dw_l, dy.double().T @ a.double()) < RTOL
    assert rel_err(db, dy.double().sum(0)) < RTOL


# ------------------------------------------------------------------ SAGEConv module (drop-in) vs oracle
def test_sageconv_exact_known_answer(dev):
    from noise_gnn_b200 import SAGEConv
    g = GOLD
    conv = SAGEConv(3, 2).to(dev)
    with torch.no_grad():
        conv.lin_l.weight.copy_(torch.tensor(g["w_l"])); conv.lin_l.bias.copy_(torch.tensor(g["b_l"]))
        conv.lin_r.weight.copy_(torch.tensor(g["w_r"]))
    x = torch.tensor(g["x"], device=dev, requires_grad=True)
    out = conv(x, torch.tensor([g["src"], g["dst"]], device=dev))
    assert rel_err(out, torch.tensor(g["out"])) < 1e-6
    (out * torch.tensor(g["grad_out"], device=dev)).sum().backward()
    assert rel_err(conv.lin_l.weight.grad, torch.tensor(g["d_w_l"])) < 1e-6
    assert rel_err(conv.lin_r.weight.grad, torch.tensor(g["d_w_r"])) < 1e-6
    assert rel_err(conv.lin_l.bias.grad, torch.tensor(g["d_b_l"])) < 1e-6
    assert rel_err(x.grad, torch.tensor(g["d_x"])) < 1e-6


@pytest.mark.parametrize("n,e,F,O,sort", [(500, 4000, 100, 256, True), (500, 4000, 256, 47, False),
                                          (300, 2500, 1433, 7, False), (800, 0, 64, 16, True),
                                          (1, 3, 8, 4, True), (400, 3000, 767, 10, True)])
def test_sageconv_forward_backward_vs_oracle(dev, n, e, F, O, sort):
    from noise_gnn_b200 import SAGEConv
    torch.manual_seed(n + F)
    ei = random_coo(n, e, seed=e + 1, sort=sort)
    ref = sage_oracle.SAGEConvRef(F, O, dtype=torch.float64)
    conv = SAGEConv(F, O).to(dev)
    conv.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    x64 = torch.randn(n, F, dtype=torch.float64, requires_grad=True)
    x = x64.detach().float().to(dev).requires_grad_(True)
    gout = torch.randn(n, O, dtype=torch.float64)
    o_ref = ref(x64, ei)
    (o_ref * gout).sum().backward()
    o = conv(x, ei.to(dev))
    (o * gout.float().to(dev)).sum().backward()
    assert rel_err(o, o_ref) < RTOL
    assert rel_err(conv.lin_l.weight.grad, ref.lin_l.weight.grad) < RTOL
    assert rel_err(conv.lin_r.weight.grad, ref.lin_r.weight.grad) < RTOL
    assert rel_err(conv.lin_l.bias.grad, ref.lin_l.bias.grad) < RTOL
    assert rel_err(x.grad, x64.grad) < RTOL


# ------------------------------------------------------------------ GCN path (reference convolution.py:7-53, module 'gcn')
@pytest.mark.parametrize("n,e,F,O,sort", [(500, 4000, 100, 256, True), (500, 4000, 256, 47, False),
                                          (300, 2500, 1433, 7, False), (800, 0, 64, 16, True),
                                          (1, 3, 8, 4, True), (400, 3000, 767, 10, True), (2000, 30000, 128, 40, False)])
def test_gcnconv_forward_backward_vs_oracle(dev, n, e, F, O, sort):
    """GCNConv(normalize=False): out = A_sum (x W^T) + b — unsorted COO, duplicate edges, zero in-degree rows,
    widths the tcgen05 path takes (F % 4 == 0) and the ones it cannot (1433, 767)."""
    from noise_gnn_b200 import GCNConv
    torch.manual_seed(n + F)
    ei = random_coo(n, e, seed=e + 1, sort=sort)
    if e > 20:
        ei[:, 10] = ei[:, 11]                      # a duplicated edge
    ref = sage_oracle.GCNConvRef(F, O, dtype=torch.float64)
    with torch.no_grad():
        ref.bias.normal_()                         # PyG initialises the bias to zero; make it visible to the test
    conv = GCNConv(F, O, normalize=False).to(dev)
    assert sorted(conv.state_dict()) == sorted(ref.state_dict()) == ["bias", "lin.weight"]
    conv.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    x64 = torch.randn(n, F, dtype=torch.float64, requires_grad=True)
    x = x64.detach().float().to(dev).requires_grad_(True)
    gout = torch.randn(n, O, dtype=torch.float64)
    o_ref = ref(x64, ei)
    (o_ref * gout).sum().backward()
    o = conv(x, ei.to(dev))
    (o * gout.float().to(dev)).sum().backward()
    assert rel_err(o, o_ref) < RTOL
    assert rel_err(conv.lin.weight.grad, ref.lin.weight.grad) < RTOL
    assert rel_err(conv.bias.grad, ref.bias.grad) < RTOL
    assert rel_err(x.grad, x64.grad) < RTOL


def test_simplegcn_network_on_a_sampled_block_vs_oracle(dev):
    """The reference's SimpleGCN (3 x GCNConv, relu between) on an identical sampled block: seed-row logits, loss and
    every parameter gradient against the fp64 oracle network."""
    from noise_gnn_b200 import NeighborLoader, SimpleGCN
    from noise_gnn_b200.synthetic import make_dataset
    data, sh, train_idx = make_dataset("arxiv", scale=0.02, device="cpu")
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=[10, 5], batch_size=64, shuffle=True)
    batch = next(iter(loader))
    torch.manual_seed(1232)
    ref = sage_oracle.SimpleGCNRef(sh.features, 64, sh.classes, 3, dropout=0.0, dtype=torch.float64)
    net = SimpleGCN(sh.features, 64, sh.classes, 3, dropout=0.0).to(dev)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    bs = batch.batch_size
    tgt = batch.yhn[:bs].view(-1).cpu()
    loss_ref = torch.nn.functional.cross_entropy(ref(batch.x.cpu().double(), batch.edge_index.cpu())[:bs], tgt)
    loss_ref.backward()
    out = net(batch.x, batch.edge_index)[:bs]
    loss = torch.nn.functional.cross_entropy(out, tgt.to(dev))
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) < 1e-5 * max(1.0, abs(float(loss_ref.detach())))
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, q.grad) < 2e-5, k


def test_sage_network_untrimmed_and_trimmed_vs_oracle(dev):
    """Reference-exact mode and the trimmed fused mode both reproduce the oracle network's seed rows and
    parameter gradients on an identical sampled block (dropout off: RNG streams are not comparable)."""
    from noise_gnn_b200 import NeighborLoader, SAGE
    from noise_gnn_b200.synthetic import make_dataset
    data, sh, train_idx = make_dataset("arxiv", scale=0.02, device="cpu")
    for L, fan in ((3, [15, 10, 5]), (3, [10, 5]), (2, [10, 5])):
        loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=fan, batch_size=64, shuffle=True)
        batch = next(iter(loader))
        torch.manual_seed(1232)
        ref = sage_oracle.SAGERef(sh.features, 64, sh.classes, L, dropout=0.0, dtype=torch.float64)
        net = SAGE(sh.features, 64, sh.classes, L, dropout=0.0).to(dev)
        net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
        x_cpu, ei_cpu = batch.x.cpu().double(), batch.edge_index.cpu()
        bs = batch.batch_size
        tgt = batch.yhn[:bs].view(-1).cpu()
        loss_ref = torch.nn.functional.cross_entropy(ref(x_cpu, ei_cpu)[:bs], tgt)
        loss_ref.backward()
        grads_ref = {k: p.grad.clone() for k, p in ref.named_parameters()}
        for mode in ("untrimmed", "trimmed"):
            net.zero_grad()
            out = net(batch.x, batch.edge_index)[:bs] if mode == "untrimmed" else net.forward_batch(batch)
            loss = torch.nn.functional.cross_entropy(out, tgt.to(dev))
            loss.backward()
            assert abs(float(loss) - float(loss_ref)) < 1e-5 * max(1.0, abs(float(loss_ref))), mode
            for k, p in net.named_parameters():
                assert rel_err(p.grad, grads_ref[k]) < 2e-5, (mode, L, fan, k)


# ------------------------------------------------------------------ sampler: bit-exact vs the C oracle + validity
@pytest.mark.parametrize("fan,replace", [([15, 10, 5], False), ([10, 5], False), ([25], False), ([3, 3, 3, 3], False),
                                         ([4, 4], True), ([32, 2], False), ([40], False), ([35, 3], True)])
def test_sampler_bit_exact_vs_oracle(dev, fan, replace):
    from noise_gnn_b200 import NeighborLoader
    from noise_gnn_b200.synthetic import make_dataset
    from oracle.validity import check_block_validity
    data, sh, train_idx = make_dataset("arxiv", scale=0.05, device="cpu")
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=fan, batch_size=128, shuffle=True,
                            replace=replace, return_e_id=True, seed=1232)
    colptr, row = loader.colptr.cpu().numpy(), loader.row.cpu().numpy()
    oc, orow, operm = structure.coo_to_csr(data.edge_index[0].numpy(), data.edge_index[1].numpy(), data.num_nodes)
    assert np.array_equal(colptr, oc) and np.array_equal(row, orow)
    cs = sampler.CSampler(colptr, row)
    order = loader.epoch_permutation(0)
    for b, batch in enumerate(loader):
        if b >= 3:
            break
        seeds = loader.batch_seeds(order, b).numpy()
        want = cs.sample(seeds, fan, replace, seed=1232, epoch=0, batch_idx=b)
        blk = batch.block
        assert blk.hop_nodes == want.node_counts.tolist() and blk.hop_edges == want.edge_counts.tolist()
        assert np.array_equal(blk.n_id.cpu().numpy(), want.n_id)
        assert np.array_equal(blk.rowptr.cpu().numpy(), want.rowptr)
        assert np.array_equal(blk.col.cpu().numpy(), want.col)
        assert np.array_equal(blk.col_global.cpu().numpy(), want.col_global)
        assert np.array_equal(batch._e_pos.cpu().numpy(), want.e_pos)
        check_block_validity(want, colptr, row, seeds, fan, replace)
        # PyG-facing views of the same block
        assert torch.equal(batch.x.cpu(), data.x[torch.from_numpy(want.n_id).long()])
        assert torch.equal(batch.y.cpu(), data.y[torch.from_numpy(want.n_id).long()])
        assert np.array_equal(batch.edge_index.cpu().numpy(), structure.csr_to_coo(want.rowptr, want.col))
        assert np.array_equal(batch.e_id.cpu().numpy(), operm[want.e_pos])
        assert bool((batch.edge_index[1][1:] >= batch.edge_index[1][:-1]).all())
    # the scratch maps are restored: a second pass over the same epoch key gives the same blocks
    loader.epoch = 0
    again = next(iter(loader))
    first = cs.sample(loader.batch_seeds(order, 0).numpy(), fan, replace, seed=1232, epoch=0, batch_idx=0)
    assert np.array_equal(again.block.col_global.cpu().numpy(), first.col_global)


def test_sampler_dp_sharding_is_rank_agnostic(dev):
    """Rank r of R sees global batches r, r+R, ...: the union over ranks equals the single-rank epoch."""
    from noise_gnn_b200 import NeighborLoader
    from noise_gnn_b200.synthetic import make_dataset
    data, sh, train_idx = make_dataset("pubmed", scale=0.2, device="cpu")
    mk = lambda r, R: NeighborLoader(data, input_nodes=torch.arange(300), num_neighbors=[5, 5], batch_size=32,
                                     shuffle=True, rank=r, world_size=R)
    single = [b.block.col_global.cpu() for b in mk(0, 1)]
    shards = [[b.block.col_global.cpu() for b in mk(r, 2)] for r in range(2)]
    assert len(single) == 10 and len(shards[0]) == len(shards[1]) == 5
    for g, blk in enumerate(single):
        assert torch.equal(blk, shards[g % 2][g // 2])


# ------------------------------------------------------------------ loss / optimizer / whole train step
def test_ce_and_adam_match_torch(dev):
    from noise_gnn_b200 import ops
    bs, C = 512, 47
    logits = torch.randn(bs, C) * 3
    tgt, y = torch.randint(0, C, (bs,)), torch.randint(0, C, (bs,))
    l64 = logits.double().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(l64, tgt)
    loss.backward()
    stats, dl = ops.ce_fwd_bwd(logits.to(dev), tgt.to(dev), y.to(dev))
    assert abs(float(stats[0]) - float(loss)) < 1e-5 * float(loss)
    assert int(stats[1]) == int((logits.argmax(-1) == y).sum())
    assert rel_err(dl, l64.grad) < RTOL
    # Adam: 5 steps against torch.optim.Adam in fp64
    p64 = torch.randn(1000, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([p64], lr=1e-3)
    p = p64.detach().float().to(dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    for it in range(5):
        g = torch.randn(1000, dtype=torch.float64)
        p64.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.float().to(dev), m, v, step, lr=1e-3)
    assert int(step) == 5
    assert rel_err(p, p64.detach()) < RTOL
