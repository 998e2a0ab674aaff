"""GPU parity at HEADLINE scale (run with -m gpu on a B200).

The hot kernels are persistent: `k_tc_gemm` strides its tiles by the grid (one CTA per SM) and the K-AGG pipeline rotates its
three register stages once per extra row a warp owns.  The small-shape tests never make those loops iterate, so these
tests run the shapes BASELINE.json configs[3] (products, bs 512, fan-out [15,10,5]) produces — several waves of tiles per
CTA, several rows per warp — against the fp64 oracle, and one full-scale fused train step on the products-shaped graph
against the untrimmed fp64 oracle network on the identical block.

Error measures: `rel_err` = max |got - want| / max |want| (scale-relative, the bar the north star states) AND
`dot_err` = max |got - want| / (|a| . |w| + floor), the element-wise forward-error bound of a dot product, so that a
wrong small entry cannot hide behind a large one."""
import numpy as np
import pytest
import torch

from oracle import philox, sage_oracle, sampler

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    return float((got - want).abs().max() / want.abs().max().clamp(min=1e-30))


def dot_err(got, want, bound):
    """max over elements of |got - want| / bound, bound = sum_k |a_ik| |w_jk| (+ |bias|): the natural element-wise scale."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float(((got - want).abs() / bound.clamp(min=1e-30)).max())


@pytest.fixture(scope="module")
def dev(cuda_device):
    from noise_gnn_b200 import _lib
    assert _lib.load().ngnn_device_supported() == 1
    return cuda_device


@pytest.fixture(params=["auto", "ss", "simt"])
def gemm_path(request, dev):
    from noise_gnn_b200 import _lib
    _lib.call("ngnn_set_gemm_path", 1 if request.param == "simt" else 0)
    _lib.call("ngnn_set_tuning", 6, 0 if request.param == "ss" else 1)
    yield request.param
    _lib.call("ngnn_set_gemm_path", 0)
    _lib.call("ngnn_set_tuning", 6, 1)


# 602 x 2 = 1204 tiles (8.1 per CTA), 157 x 2 = 314 tiles, 602 tiles: the persistent loop, both TMEM accumulator
# buffers, the smem ring wrapping across tile boundaries
BIG_GEMMS = [(77056, 100, 256), (20000, 256, 256), (77056, 256, 47)]


@pytest.mark.parametrize("n,F,O", BIG_GEMMS)
def test_gemm_fwd_many_tiles_per_cta(dev, gemm_path, n, F, O):
    from noise_gnn_b200 import ops
    g = torch.Generator().manual_seed(n + F + O)
    a_l, a_r = torch.randn(n, F, generator=g), torch.randn(n, F, generator=g)
    w_l, w_r = torch.randn(O, F, generator=g) / F ** 0.5, torch.randn(O, F, generator=g) / F ** 0.5
    b = torch.randn(O, generator=g)
    want = a_l.double() @ w_l.double().T + a_r.double() @ w_r.double().T + b.double()
    bound = a_l.double().abs() @ w_l.double().abs().T + a_r.double().abs() @ w_r.double().abs().T + b.double().abs()
    d = [t.to(dev) for t in (a_l, a_r, w_l, w_r, b)]
    got, path = ops.gemm_fwd(*d, n, return_path=True)
    assert path == (0 if gemm_path == "simt" else 1)
    assert rel_err(got, want) < RTOL
    assert dot_err(got, want, bound) < RTOL
    # fused ReLU + Philox dropout on every tile: the mask is the numpy oracle's mask, the kept values are exact
    p = 0.5
    got_d = ops.gemm_fwd(*d, n, act=1, drop_p=p, seed=1232, offset=9)
    keep = torch.from_numpy(philox.dropout_keep_mask(n, O, p, seed=1232, offset=9))
    want_d = torch.where(keep, want.clamp(min=0) / (1 - p), torch.zeros_like(want))
    # (a pre-activation within rounding distance of zero may take either side of the ReLU: excluded from the mask check)
    clear = want.abs() > 1e-5 * bound
    assert torch.equal((got_d != 0).cpu()[clear], (want_d != 0)[clear])
    assert dot_err(got_d, want_d, bound / (1 - p)) < RTOL
    # bitwise reproducible
    assert torch.equal(got, ops.gemm_fwd(*d, n))


def _pad4(t):
    """The same values in storage whose row stride is a multiple of 4 floats (what the loader's feature table and the
    step's activation arena use): TMA needs 16-byte row strides."""
    n, F = t.shape
    buf = torch.zeros(n, (F + 3) // 4 * 4, dtype=t.dtype, device=t.device)
    buf[:, :F] = t
    return buf[:, :F]


# F % 4 != 0 and contractions longer than one tensor-memory accumulation chain (F = 1433 is cora, 767 computers, 500 pubmed):
# with padded row strides these run on the tensor cores, a long contraction as several chained launches
LONG_K = [(2708, 1433, 7), (3000, 767, 10), (5000, 1433, 256), (4000, 500, 3), (1000, 2047, 64), (700, 641, 512)]


@pytest.mark.parametrize("n,F,O", LONG_K)
def test_gemm_fwd_long_or_ragged_k_on_tensor_cores(dev, n, F, O):
    from noise_gnn_b200 import ops
    g = torch.Generator().manual_seed(n + F + O)
    a_l, a_r = torch.randn(n, F, generator=g), torch.randn(n, F, generator=g)
    w_l, w_r = torch.randn(O, F, generator=g) / F ** 0.5, torch.randn(O, F, generator=g) / F ** 0.5
    b = torch.randn(O, generator=g)
    want = a_l.double() @ w_l.double().T + a_r.double() @ w_r.double().T + b.double()
    bound = a_l.double().abs() @ w_l.double().abs().T + a_r.double().abs() @ w_r.double().abs().T + b.double().abs()
    d = [_pad4(a_l.to(dev)), _pad4(a_r.to(dev)), w_l.to(dev), w_r.to(dev), b.to(dev)]
    got, path = ops.gemm_fwd(*d, n, return_path=True)
    assert path == 1
    assert rel_err(got, want) < RTOL and dot_err(got, want, bound) < RTOL
    p = 0.5
    got_d, path = ops.gemm_fwd(*d, n, act=1, drop_p=p, seed=1232, offset=9, return_path=True)
    assert path == 1
    keep = torch.from_numpy(philox.dropout_keep_mask(n, O, p, seed=1232, offset=9))
    want_d = torch.where(keep, want.clamp(min=0) / (1 - p), torch.zeros_like(want))
    clear = want.abs() > 1e-5 * bound
    assert torch.equal((got_d != 0).cpu()[clear], (want_d != 0)[clear])
    assert dot_err(got_d, want_d, bound / (1 - p)) < RTOL
    # root-only (one operand), no bias
    got_r, path = ops.gemm_fwd(None, d[1], None, d[3], None, n, return_path=True)
    assert path == 1 and rel_err(got_r, a_r.double() @ w_r.double().T) < RTOL
    assert torch.equal(got, ops.gemm_fwd(*d, n))


@pytest.mark.parametrize("n,F,O", BIG_GEMMS)
def test_dgrad_wgrad_many_tiles_per_cta(dev, gemm_path, n, F, O):
    from noise_gnn_b200 import ops
    g = torch.Generator().manual_seed(3 * n + F + O)
    dy = torch.randn(n, O, generator=g)
    a_l, a_r = torch.randn(n, F, generator=g), torch.randn(n, F, generator=g)
    w_l, w_r = torch.randn(O, F, generator=g), torch.randn(O, F, generator=g)
    deg = torch.randint(0, 6, (n,), generator=g)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)]).int()
    dmean, droot = ops.dgrad(dy.to(dev), w_l.to(dev), w_r.to(dev), rowptr.to(dev), n)
    inv = 1.0 / deg.clamp(min=1).double()
    want_m = (dy.double() @ w_l.double()) * inv[:, None]
    bound_m = (dy.double().abs() @ w_l.double().abs()) * inv[:, None]
    assert rel_err(dmean, want_m) < RTOL and dot_err(dmean, want_m, bound_m) < RTOL
    want_r = dy.double() @ w_r.double()
    assert rel_err(droot, want_r) < RTOL and dot_err(droot, want_r, dy.double().abs() @ w_r.double().abs()) < RTOL
    dw_l, dw_r, db = ops.wgrad(dy.to(dev), a_l.to(dev), a_r.to(dev), n, F)
    for got, a in ((dw_l, a_l), (dw_r, a_r)):
        want = dy.double().T @ a.double()
        assert rel_err(got, want) < RTOL
        assert dot_err(got, want, dy.double().abs().T @ a.double().abs()) < RTOL
    assert rel_err(db, dy.double().sum(0)) < RTOL
    dw2, _, _ = ops.wgrad(dy.to(dev), a_l.to(dev), a_r.to(dev), n, F)
    assert torch.equal(dw2, dw_l)                                   # fixed split-K order: bitwise reproducible


@pytest.mark.parametrize("F,n_dst,table_rows", [(100, 77000, 1_200_000), (256, 66000, 300_000), (128, 150_000, 1_000_000)])
def test_agg_fwd_table_gather_many_rows_per_warp(dev, F, n_dst, table_rows):
    """K-AGG with the fused root gather from a table that does not fit L2, >= 60 k destination rows: every warp of the
    persistent pipelined kernel owns several rows (the 3-stage register rotation runs), degrees 0..25 plus a few long rows."""
    from noise_gnn_b200 import ops
    g = torch.Generator().manual_seed(F + n_dst)
    deg = torch.randint(0, 12, (n_dst,), generator=g)
    deg[5], deg[6], deg[n_dst - 1], deg[n_dst // 2] = 0, 33, 70, 700
    dst = torch.repeat_interleave(torch.arange(n_dst), deg)
    src = torch.randint(0, table_rows, (dst.numel(),), generator=g)
    table = torch.randn(table_rows, F, generator=g)
    root_idx = torch.randint(0, table_rows, (n_dst,), generator=g, dtype=torch.int32)
    want = sage_oracle.mean_aggregate(table.double(), torch.stack([src, dst]), n_dst)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)]).int()
    mean, root = ops.agg_fwd(rowptr.to(dev), src.int().to(dev), table.to(dev), n_dst, root_idx=root_idx.to(dev))
    assert rel_err(mean, want) < RTOL
    absmean = sage_oracle.mean_aggregate(table.double().abs(), torch.stack([src, dst]), n_dst)
    assert float(((mean.double().cpu() - want).abs() / absmean.clamp(min=1e-30)).max()) < RTOL
    assert torch.equal(root.cpu(), table[root_idx.long()])
    assert float(mean[5].abs().max()) == 0.0
    # without the root gather (layers >= 2) and bitwise repeatability
    mean2 = ops.agg_fwd(rowptr.to(dev), src.int().to(dev), table.to(dev), n_dst)
    assert torch.equal(mean2, mean)


def test_agg_bwd_many_rows(dev):
    """K-AGG-T at the layer-2 backward shape of a products block: ~77 k source rows x 256, most with one transposed
    neighbour, a few hubs, add rows on a prefix, gate."""
    from noise_gnn_b200 import ops
    F, n_src, n_dst = 256, 77000, 7600
    g = torch.Generator().manual_seed(11)
    e = 84000
    src = torch.randint(0, n_src, (e,), generator=g)
    src[:3000] = 17                                              # a hub: 3000 transposed neighbours (> kLongRow)
    dst = torch.sort(torch.randint(0, n_dst, (e,), generator=g)).values
    ei = torch.stack([src, dst])
    blk = ops.coo_to_csr(ei.to(dev), n_src)
    dmean, droot, h = torch.randn(n_dst, F, generator=g), torch.randn(n_dst, F, generator=g), torch.randn(n_src, F, generator=g)
    want = torch.zeros(n_src, F, dtype=torch.float64).index_add_(0, src, dmean.double()[dst])
    want[:n_dst] += droot.double()
    want = torch.where(h > 0, want * 2.0, torch.zeros_like(want))
    ct, rt, _ = ops.csr_transpose(blk.rowptr, blk.col, n_src, e, n_src)
    got = ops.agg_bwd(ct, rt, dmean.to(dev), n_src, dx_root=droot.to(dev), n_root=n_dst, act_ref=h.to(dev), act_scale=2.0)
    assert rel_err(got, want) < RTOL
    assert torch.equal(got, ops.agg_bwd(ct, rt, dmean.to(dev), n_src, dx_root=droot.to(dev), n_root=n_dst, act_ref=h.to(dev),
                                        act_scale=2.0))


@pytest.fixture(scope="module")
def products_full(dev):
    """BASELINE.json configs[3] at full scale: 2,449,029 nodes / 123.7 M directed edges / F 100, generated on the GPU."""
    from noise_gnn_b200 import NeighborLoader
    from noise_gnn_b200.synthetic import make_dataset
    data, sh, train_idx = make_dataset("products", seed=1232, law="powerlaw", device=dev, noise_type="sym", noise_rate=0.3)
    loader = NeighborLoader(data, input_nodes=train_idx, num_neighbors=list(sh.fanouts), batch_size=sh.batch_size,
                            shuffle=True, seed=1232)
    return data, sh, loader


@pytest.mark.parametrize("dropout", [0.0, 0.5])
def test_full_scale_products_step_matches_oracle(dev, products_full, dropout):
    """One fused train step (sample -> trimmed SAGE fwd -> CE -> bwd) on the full products-shaped graph, bs 512, fan-out
    [15,10,5], hidden 256, against the UNtrimmed fp64 oracle network on the identical block: sampler bit-exact vs the C
    oracle, loss, seed-row logits, every parameter gradient; with dropout the oracle applies the numpy Philox masks."""
    from noise_gnn_b200 import SAGE
    from noise_gnn_b200.train import Trainer
    data, sh, loader = products_full
    L = sh.layers
    torch.manual_seed(1232)
    ref = sage_oracle.SAGERef(sh.features, sh.hidden, sh.classes, L, dropout=dropout, dtype=torch.float64)
    net = SAGE(sh.features, sh.hidden, sh.classes, L, dropout=dropout).to(dev)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    net.train(); ref.train()
    trainer = Trainer(net, lr=1e-3)
    loader.epoch = 0
    batch = next(iter(loader))
    blk = batch.block
    assert blk.hop_nodes[2] > 148 * 128 * 2            # layer 1: several 128-row tiles per CTA
    # sampler: bit-exact against the sequential C oracle on the full graph
    seeds = loader.batch_seeds(loader.epoch_permutation(0), 0).numpy()
    want = sampler.sample_block(loader.colptr.cpu().numpy(), loader.row.cpu().numpy(), seeds, list(sh.fanouts), seed=1232,
                                epoch=0, batch_idx=0)
    assert np.array_equal(blk.n_id.cpu().numpy(), want.n_id)
    assert np.array_equal(blk.rowptr.cpu().numpy(), want.rowptr)
    assert np.array_equal(blk.col.cpu().numpy(), want.col)
    bs, n = batch.batch_size, batch.num_nodes
    masks = None
    if dropout > 0:
        # rows beyond the last trimmed extent never reach the seed rows: their mask is irrelevant (ones)
        step = trainer.steps + 1
        masks = []
        for i in range(L - 1):
            rows = blk.hop_nodes[min(L - 1 - i, len(blk.hop_nodes) - 1)]
            m = np.ones((n, sh.hidden), dtype=bool)
            m[:rows] = philox.dropout_keep_mask(rows, sh.hidden, dropout, seed=net.drop_seed, offset=step * L + i)
            masks.append(torch.from_numpy(m))
    logits = trainer.forward_backward(batch, want_logits=True)
    loss, correct = trainer.read_stats()
    # The hidden activations of the same kernels (per-layer variant, same dropout stream): their signs are the ReLU gates the
    # backward used.  The fp64 oracle takes those gates (oracle/sage_oracle.py::SAGERef.forward), so that the comparison
    # measures arithmetic, not which side of zero a 1e-7 pre-activation fell on.
    net._drop_calls = trainer.steps - 1
    with torch.no_grad():
        _, hidden = net.forward_batch(batch, return_hidden=True)
    relu_masks, flips = [], 0
    for i in range(L - 1):
        m = hidden[i] > 0
        if masks is not None:
            m = m | ~masks[i][: m.size(0)].to(dev)              # a dropped element's gate is irrelevant
        relu_masks.append(m.cpu())
    tgt = batch.yhn[:bs].view(-1).cpu()
    x_cpu, ei_cpu = batch.x.cpu().double(), batch.edge_index.cpu()
    out_ref = ref(x_cpu, ei_cpu, dropout_masks=masks, relu_masks=relu_masks)[:bs]
    loss_ref = torch.nn.functional.cross_entropy(out_ref, tgt)
    loss_ref.backward()
    assert rel_err(logits, out_ref) < RTOL
    assert abs(loss - float(loss_ref)) < 1e-5 * max(1.0, float(loss_ref))
    assert correct == int((out_ref.argmax(-1) == batch.y[:bs].view(-1).cpu()).sum())
    errs = {k: rel_err(p.grad, q.grad) for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters())}
    assert max(errs.values()) < 2e-5, errs
    # how many gates differ from the oracle's own sign, and what they cost: the free-running oracle (its own ReLU) may be
    # further away, but only through those few elements
    ref.zero_grad()
    with torch.no_grad():
        h = x_cpu
        for i in range(L - 1):
            z = ref.convs[i](h, ei_cpu)
            own = z[: relu_masks[i].size(0)] > 0
            keep = masks[i][: own.size(0)] if masks is not None else torch.ones_like(own)
            flips += int(((own != relu_masks[i]) & keep).sum())
            h = torch.where(z > 0, z, torch.zeros_like(z))
            if masks is not None:
                h = h * masks[i][: h.size(0)].to(h.dtype) / (1.0 - dropout)
    n_gates = sum(int(m.numel()) for m in relu_masks)
    assert flips <= max(64, n_gates // 100_000), (flips, n_gates)   # a handful of 2e7 gates sit within rounding of zero


@pytest.mark.parametrize("F,n_src,n_dst", [(256, 77000, 7600), (100, 50000, 9000), (200, 33, 7), (132, 4100, 4100)])
def test_agg_bwd_staged_gate_kernel(dev, F, n_src, n_dst):
    """K-AGG-T of the training backward (gate + root rows; the gate rows staged through shared memory by bulk copies) at
    the products layer-2 shape: mostly one transposed entry per row, some empty rows, a few hub rows handed to the whole CTA,
    a ragged last tile; against the fp64 formula and bit for bit against the generic kernel."""
    from noise_gnn_b200 import _lib, ops
    g = torch.Generator().manual_seed(F + n_src)
    deg = torch.randint(0, 3, (n_src,), generator=g)
    deg[torch.randint(0, n_src, (max(n_src // 2000, 2),), generator=g)] = 300          # hubs (> the 64-entry threshold)
    deg[5] = 17                                                                       # a second index window, not a hub
    colptr = torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)])
    e = int(colptr[-1])
    row_t = torch.randint(0, n_dst, (e,), generator=g)
    ld = (F + 3) // 4 * 4
    pad = lambda t: _pad4(t) if ld != F else t
    dmean, droot = torch.randn(n_dst, F, generator=g), torch.randn(n_dst, F, generator=g)
    h = torch.randn(n_src, F, generator=g)
    h[h.abs() < 0.3] = 0.0                                                            # exact zeros gate to zero
    want = torch.zeros(n_src, F, dtype=torch.float64)
    want.index_add_(0, torch.repeat_interleave(torch.arange(n_src), deg), dmean.double()[row_t])
    want[:n_dst] += droot.double()
    want = torch.where(h > 0, want * 2.0, torch.zeros_like(want))
    d = dict(ct=colptr.int().to(dev), rt=row_t.int().to(dev), dm=pad(dmean.to(dev)), dr=pad(droot.to(dev)), h=pad(h.to(dev)))
    outs = []
    for staged in (2, 0):                                     # 2 = the staged kernel whatever the row count
        _lib.call("ngnn_set_tuning", 14, staged)
        out = pad(torch.full((n_src, F), float("nan"), device=dev))
        ops.agg_bwd(d["ct"], d["rt"], d["dm"], n_src, dx_root=d["dr"], n_root=n_dst, act_ref=d["h"], act_scale=2.0, out=out)
        outs.append(out.clone())
    _lib.call("ngnn_set_tuning", 14, 1)
    assert rel_err(outs[0], want) < RTOL
    if F > 128:
        assert torch.equal(outs[0], outs[1])                  # same summation order as the generic kernel's half-warp form
    else:
        assert rel_err(outs[0], outs[1]) < 1e-6               # (narrower rows: the generic kernel cuts hub rows into 8 slices, not 4)
    # without root rows
    _lib.call("ngnn_set_tuning", 14, 2)
    out = ops.agg_bwd(d["ct"], d["rt"], d["dm"], n_src, act_ref=d["h"], act_scale=1.0)
    _lib.call("ngnn_set_tuning", 14, 1)
    want2 = torch.zeros(n_src, F, dtype=torch.float64)
    want2.index_add_(0, torch.repeat_interleave(torch.arange(n_src), deg), dmean.double()[row_t])
    assert rel_err(out, torch.where(h > 0, want2, torch.zeros_like(want2))) < RTOL
