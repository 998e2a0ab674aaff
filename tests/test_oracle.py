"""CPU tests: the oracle against published / exact / independent answers (no GPU needed)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import philox, sage_oracle, sampler, structure
from oracle.validity import check_block_validity

GOLD = json.loads((Path(__file__).parent / "golden" / "sage_known_answer.json").read_text())


def test_philox_random123_known_answers():
    # kat_vectors of Random123 (philox4x32, 10 rounds)
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    import ctypes
    lib = sampler._load()
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == want
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
        lib.ngnn_oracle_philox(c, k, o)
        assert tuple(o) == want


def test_sageconv_oracle_matches_exact_known_answer():
    g = GOLD
    conv = sage_oracle.SAGEConvRef(3, 2, dtype=torch.float64)
    with torch.no_grad():
        conv.lin_l.weight.copy_(torch.tensor(g["w_l"])); conv.lin_l.bias.copy_(torch.tensor(g["b_l"]))
        conv.lin_r.weight.copy_(torch.tensor(g["w_r"]))
    x = torch.tensor(g["x"], dtype=torch.float64, requires_grad=True)
    ei = torch.tensor([g["src"], g["dst"]])
    out = conv(x, ei)
    assert torch.allclose(out, torch.tensor(g["out"], dtype=torch.float64), rtol=0, atol=1e-12)
    assert torch.allclose(sage_oracle.mean_aggregate(x, ei), torch.tensor(g["mean"], dtype=torch.float64), atol=1e-12)
    (out * torch.tensor(g["grad_out"], dtype=torch.float64)).sum().backward()
    for got, key in ((conv.lin_l.weight.grad, "d_w_l"), (conv.lin_r.weight.grad, "d_w_r"),
                     (conv.lin_l.bias.grad, "d_b_l"), (x.grad, "d_x")):
        assert torch.allclose(got, torch.tensor(g[key], dtype=torch.float64), rtol=0, atol=1e-12), key


def test_sageconv_oracle_fp32_vs_fp64_calibrates_tolerance():
    torch.manual_seed(0)
    n, e, F, O = 500, 3000, 100, 47
    x = torch.randn(n, F, dtype=torch.float64)
    ei = torch.randint(0, n, (2, e))
    c64 = sage_oracle.SAGEConvRef(F, O, dtype=torch.float64)
    c32 = sage_oracle.SAGEConvRef(F, O)
    c32.load_state_dict({k: v.float() for k, v in c64.state_dict().items()})
    o64, o32 = c64(x, ei), c32(x.float(), ei)
    rel = (o32.double() - o64).abs().max() / o64.abs().max()
    assert rel < 1e-5   # the north-star tolerance is achievable in plain fp32


def test_structure_oracle_known_answer():
    src = np.array([1, 3, 1, 2, 4, 0, 2, 5, 3]); dst = np.array([0, 1, 0, 2, 2, 3, 4, 4, 0])
    rowptr, col, perm = structure.coo_to_csr(src, dst, 6)
    assert rowptr.tolist() == [0, 3, 4, 6, 7, 9, 9]
    assert col.tolist() == [1, 1, 3, 3, 2, 4, 0, 2, 5]
    assert perm.tolist() == [0, 2, 8, 1, 3, 4, 5, 6, 7]
    colptr_t, row_t, perm_t = structure.csr_transpose(rowptr, col, 6)
    assert colptr_t.tolist() == [0, 1, 3, 5, 7, 8, 9]
    assert row_t.tolist() == [3, 0, 0, 2, 4, 0, 1, 2, 4]
    assert np.array_equal(structure.csr_to_coo(rowptr, col), np.stack([src[perm], dst[perm]]))
    # empty graph
    rp, c, p = structure.coo_to_csr(np.zeros(0, int), np.zeros(0, int), 4)
    assert rp.tolist() == [0] * 5 and len(c) == 0


def _random_csc(n, m, seed):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, m); dst = rng.integers(0, n, m)
    rowptr, col, _ = structure.coo_to_csr(src, dst, n)
    return rowptr, col


@pytest.mark.parametrize("replace", [False, True])
def test_c_sampler_matches_python_twin(replace):
    colptr, row = _random_csc(300, 4000, 1)
    seeds = np.random.default_rng(2).permutation(300)[:17]
    for fan in ([3], [15, 10, 5], [25, 1], [2, 2, 2, 2]):
        a = sampler.sample_block(colptr, row, seeds, fan, replace, seed=1232, epoch=3, batch_idx=7)
        b = sampler.sample_block_py(colptr, row, seeds, fan, replace, seed=1232, epoch=3, batch_idx=7)
        for f in ("n_id", "rowptr", "col", "col_global", "e_pos", "node_counts", "edge_counts"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), (fan, f)


def test_sampler_oracle_validity_and_determinism():
    colptr, row = _random_csc(500, 9000, 5)
    seeds = np.random.default_rng(6).permutation(500)[:32]
    for fan, rep in (([15, 10, 5], False), ([4, 4], True), ([30], False)):
        blk = sampler.sample_block(colptr, row, seeds, fan, rep, seed=1232, epoch=0, batch_idx=0)
        check_block_validity(blk, colptr, row, seeds, fan, rep)
        again = sampler.sample_block(colptr, row, seeds, fan, rep, seed=1232, epoch=0, batch_idx=0)
        assert np.array_equal(blk.col_global, again.col_global)
        other = sampler.sample_block(colptr, row, seeds, fan, rep, seed=1232, epoch=1, batch_idx=0)
        if max(fan) < 30:   # fan-out 30 exceeds most degrees here => take-all, nothing random
            assert not np.array_equal(blk.e_pos, other.e_pos)


def test_sampler_oracle_positions_are_uniform():
    # chi-square on the sampled positions of one high-degree node over many (epoch, batch) keys
    d, k = 40, 10
    colptr = np.array([0, d] + [d] * d, dtype=np.int32)
    row = np.arange(1, d + 1, dtype=np.int32)
    s = sampler.CSampler(colptr, row)
    hist = np.zeros(d)
    trials = 4000
    for t in range(trials):
        blk = s.sample(np.array([0]), [k], False, seed=1232, epoch=t, batch_idx=t // 7)
        hist[blk.e_pos] += 1
    expect = trials * k / d
    chi2 = ((hist - expect) ** 2 / expect).sum()
    assert chi2 < 80.0   # 39 dof: mean 39, 99.99th percentile ~ 78


def test_dropout_mask_rate():
    m = philox.dropout_keep_mask(512, 256, 0.5, seed=1232, offset=9)
    assert abs(m.mean() - 0.5) < 0.01
    assert philox.dropout_keep_mask(8, 7, 0.0, 1, 1).all()


def test_gcnconv_oracle_hand_computed():
    """GCNConv(normalize=False) on a 4-node graph worked by hand: sum over in-neighbours of x W^T, + bias; node 3 has
    no in-edge (=> bias only); edge (0 -> 1) is duplicated and counts twice; the self loop (2 -> 2) is an ordinary edge."""
    import torch
    from oracle import sage_oracle
    conv = sage_oracle.GCNConvRef(2, 2, dtype=torch.float64)
    with torch.no_grad():
        conv.lin.weight.copy_(torch.tensor([[1.0, 2.0], [0.5, -1.0]], dtype=torch.float64))
        conv.bias.copy_(torch.tensor([0.25, -0.5], dtype=torch.float64))
    x = torch.tensor([[1.0, 0.0], [0.0, 1.0], [2.0, 1.0], [3.0, 3.0]], dtype=torch.float64)
    ei = torch.tensor([[0, 0, 2, 1, 3], [1, 1, 2, 0, 0]])
    z = [[1.0, 0.5], [2.0, -1.0], [4.0, 0.0], [9.0, -1.5]]                 # x W^T
    want = [[z[1][0] + z[3][0] + 0.25, z[1][1] + z[3][1] - 0.5],           # node 0 <- 1, 3
            [2 * z[0][0] + 0.25, 2 * z[0][1] - 0.5],                       # node 1 <- 0 twice
            [z[2][0] + 0.25, z[2][1] - 0.5],                               # node 2 <- itself
            [0.25, -0.5]]                                                  # node 3: no in-edge
    assert torch.equal(conv(x, ei), torch.tensor(want, dtype=torch.float64))


def test_ct_oracle_hand_computed_exchange():
    """CTLoss restatement (reference src/utils/losses.py:19-49) on a 4-sample, 2-class case worked by hand: each network
    is trained on the samples its PEER ranks easiest; forget_rate 0.5 keeps int(0.5 * 4) = 2 of them."""
    import math

    import torch
    from oracle import ct_oracle
    # logits chosen so that per-sample CE is log(1 + exp(-m)) with margins m (target class always 0)
    m1 = [3.0, -1.0, 2.0, 0.0]      # network 1: easiest samples are 0 and 2
    m2 = [-2.0, 4.0, 1.0, 0.5]      # network 2: easiest samples are 1 and 2
    y1 = torch.tensor([[m, 0.0] for m in m1], dtype=torch.float64)
    y2 = torch.tensor([[m, 0.0] for m in m2], dtype=torch.float64)
    yn = torch.zeros(4, dtype=torch.long)
    ind = torch.tensor([7, 5, 3, 1, 0])                         # batch.n_id (seeds first)
    clean = torch.tensor([1, 0, 0, 1, 0, 1, 0, 0], dtype=torch.float64)   # noise_or_not by global id
    l1, l2, p1, p2, u1, u2, n1, n2 = ct_oracle.ct_loss(y1, y2, yn, 0.5, ind, clean)
    ce = lambda m: math.log1p(math.exp(-m))
    assert u1.tolist() == [0, 2] and n1.tolist() == [3, 1]      # ascending loss under network 1: 0, 2, 3, 1
    assert u2.tolist() == [1, 2] and n2.tolist() == [3, 0]      # under network 2: 1, 2, 3, 0
    assert abs(float(l1) - (ce(m1[1]) + ce(m1[2])) / 2) < 1e-12   # network 1 on network 2's selection {1, 2}
    assert abs(float(l2) - (ce(m2[0]) + ce(m2[2])) / 2) < 1e-12   # network 2 on network 1's selection {0, 2}
    assert float(p1) == float(clean[7] + clean[3]) / 2 == 0.5     # global ids of samples 0 and 2: one clean, one not
    assert float(p2) == float(clean[5] + clean[3]) / 2 == 1.0     # global ids of samples 1 and 2: both clean


def test_ctloss_autograd_nodes_are_independent():
    """CPU check of the autograd plumbing behind the drop-in CTLoss (no kernel call): two losses with precomputed
    gradients backpropagate one after the other, as the reference's train_ct does (src/pipeline.py:127-133)."""
    import torch
    from noise_gnn_b200.losses import _ScaledGrad
    y1 = torch.randn(5, 3, requires_grad=True)
    y2 = torch.randn(5, 3, requires_grad=True)
    d1, d2 = torch.randn(5, 3), torch.randn(5, 3)
    l1 = _ScaledGrad.apply(y1 * 1.0, torch.tensor(1.5), d1)
    l2 = _ScaledGrad.apply(y2 * 1.0, torch.tensor(2.5), d2)
    l1.backward()
    (l2 * 2).backward()
    assert torch.equal(y1.grad, d1) and torch.equal(y2.grad, d2 * 2) and float(l1) == 1.5


def test_shuffle_rows_twin_follows_the_reference_law():
    """oracle/sagepl_oracle.shuffle_rows (the bit-exact target of ngnn_shuffle_rows): per row, at most k = int(F * prob)
    positions change and the row stays a permutation of itself (reference src/utils/augmentation.py:88-102)."""
    import numpy as np
    from oracle import sagepl_oracle
    rng = np.random.default_rng(0)
    x = rng.standard_normal((40, 100)).astype(np.float32)
    for k in (0, 1, 10, 50, 100):
        out = sagepl_oracle.shuffle_rows(x, k, seed=1232, offset=3)
        assert np.array_equal(np.sort(out, 1), np.sort(x, 1))
        assert int((out != x).sum(1).max()) <= k
    a = sagepl_oracle.shuffle_rows(x, 10, seed=1232, offset=3)
    assert np.array_equal(a, sagepl_oracle.shuffle_rows(x, 10, seed=1232, offset=3))
    assert not np.array_equal(a, sagepl_oracle.shuffle_rows(x, 10, seed=1232, offset=4))
    assert 5 < (a != x).sum(1).mean() <= 10          # ~k(1 - 1/k) positions actually move


def test_adding_noise_oracle_matches_the_reference_expression():
    import torch
    from oracle import sagepl_oracle
    x, noise = torch.randn(6, 5), torch.randn(9, 5)
    n_id = torch.tensor([8, 0, 3, 4, 1, 2])
    want = x + torch.nn.functional.normalize(noise[n_id]) * 0.1
    assert torch.allclose(sagepl_oracle.adding_noise(x, noise, 0.1, n_id), want)
    x9 = torch.randn(9, 5)
    want = x9 + torch.sign(x9) * torch.nn.functional.normalize(noise) * 0.1
    assert torch.allclose(sagepl_oracle.adding_noise(x9, noise, 0.1, None), want)
