"""CPU tests of the multi-GPU path's host logic (world_size 2, gloo): seed sharding and the gradient all-reduce.

The CUDA kernels cannot run here, so each rank computes its block and gradient with the CPU oracle; what is under
test is the product's own sharding (noise_gnn_b200.sharding) and exchange (noise_gnn_b200.dp) code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from noise_gnn_b200.sharding import SeedSharder


def test_sharder_deals_global_batches_round_robin():
    nodes = torch.arange(1000, 1103)          # 103 seeds, bs 10 -> 11 global batches
    single = SeedSharder(nodes, 10, shuffle=True, seed=1232)
    assert single.num_batches_global == 11 and len(single) == 11
    order = single.epoch_permutation(3)
    assert sorted(order.tolist()) == nodes.tolist()
    assert torch.equal(order, SeedSharder(nodes, 10, True, 1232, rank=1, world_size=4).epoch_permutation(3))   # rank-agnostic
    assert not torch.equal(order, single.epoch_permutation(4))
    for R in (2, 3, 4, 8):
        shards = [SeedSharder(nodes, 10, True, 1232, rank=r, world_size=R) for r in range(R)]
        steps = len(shards[0])
        assert all(len(s) == steps for s in shards) and steps == -(-11 // R)     # equal step counts on every rank
        seen = [s.global_batch_index(i) for i in range(steps) for s in shards]
        assert seen[:11] == list(range(11))                                       # r, r+R, r+2R, ... covers the epoch in order
        assert all(g < 11 for g in seen)                                          # the padded round wraps around
    last = single.batch_seeds(order, 10)
    assert len(last) == 3                                                         # partial last batch, as PyG's loader
    assert len(SeedSharder(nodes, 10, True, drop_last=True)) == 10
    with pytest.raises(ValueError):
        SeedSharder(nodes, 10, True, rank=2, world_size=2)
    # full-batch configs (fewer seeds than batch_size): the one batch is "full", a genuinely short last batch is not
    fb = SeedSharder(torch.arange(60), 512, True)
    assert len(fb) == 1 and fb.full_len == 60 and fb.round_is_full(0)
    assert single.full_len == 10 and single.round_is_full(9) and not single.round_is_full(10)
    two = [SeedSharder(nodes, 10, True, 1232, rank=r, world_size=2) for r in range(2)]
    assert all(s.round_is_full(4) for s in two) and not any(s.round_is_full(5) for s in two)    # round 5 = batch 10 (3 seeds) + filler


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from noise_gnn_b200 import dp
    from noise_gnn_b200.synthetic import make_dataset
    from oracle import sage_oracle, sampler, structure
    data, sh, train_idx = make_dataset("pubmed", scale=0.05, device="cpu")
    colptr, row, _ = structure.coo_to_csr(data.edge_index[0].numpy(), data.edge_index[1].numpy(), data.num_nodes)
    cs = sampler.CSampler(colptr, row)
    fan, bs = [5, 5], 8
    seeds_all = torch.arange(64)
    torch.manual_seed(7)
    model = sage_oracle.SAGERef(sh.features, 16, sh.classes, 2, dropout=0.0, dtype=torch.float64)

    def grad_of_batch(g, order, sharder):
        blk = cs.sample(sharder.batch_seeds(order, g).numpy(), fan, seed=1232, epoch=0, batch_idx=g)
        n_id = torch.from_numpy(blk.n_id.astype(np.int64))
        ei = torch.from_numpy(structure.csr_to_coo(blk.rowptr, blk.col))
        model.zero_grad()
        out = model(data.x[n_id].double(), ei)[:bs]
        torch.nn.functional.cross_entropy(out, data.y[n_id][:bs].view(-1)).backward()
        return torch.cat([p.grad.view(-1) for p in model.parameters()]).clone()

    sharder = SeedSharder(seeds_all, bs, shuffle=True, seed=1232, rank=rank, world_size=world)
    order = sharder.epoch_permutation(0)
    g_mine = sharder.global_batch_index(0)
    bucket = grad_of_batch(g_mine, order, sharder)
    scale = dp.allreduce_mean_(bucket, world_size=world)
    reduced = bucket * scale
    # single-process reference: the mean of the gradients of global batches 0..world-1
    solo = SeedSharder(seeds_all, bs, shuffle=True, seed=1232)
    want = sum(grad_of_batch(g, order, solo) for g in range(world)) / world
    torch.save({"reduced": reduced, "want": want, "g": g_mine, "scale": scale}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert [r["g"] for r in res] == [0, 1]
    assert all(abs(r["scale"] - 0.5) < 1e-12 for r in res)
    assert torch.equal(res[0]["reduced"], res[1]["reduced"])                  # every rank holds the same averaged gradient
    assert torch.allclose(res[0]["reduced"], res[0]["want"], rtol=1e-12, atol=1e-14)


def test_ragged_and_padded_rounds_average_over_the_real_seeds():
    """The last round of an epoch can hold a short batch and, on some ranks, a wrap-around filler: with the sharder's
    loss_scale as the weight of each rank's MEAN, the all-reduced average (sum / R) is the mean over the round's real seeds
    (pad the last round, mask padded seeds out of the loss).  Pure host arithmetic, no process group needed."""
    n, bs = 10 * 16 + 5, 16                       # 11 global batches, the last with 5 seeds
    vals = torch.randn(n, dtype=torch.float64)    # a per-seed quantity (stands in for the per-seed loss gradient)
    nodes = torch.arange(n)
    for R in (1, 2, 3, 4, 8):
        shards = [SeedSharder(nodes, bs, True, 1232, rank=r, world_size=R) for r in range(R)]
        order = shards[0].epoch_permutation(0)
        steps = len(shards[0])
        for i in range(steps):
            real = [r for r in range(R) if i * R + r < shards[0].num_batches_global]
            assert [s.is_padded(i) for s in shards] == [r not in real for r in range(R)]
            seeds = [s.batch_seeds(order, s.global_batch_index(i)) for s in shards]
            means = [vals[sd].mean() for sd in seeds]
            got = sum(s.loss_scale(i) * m for s, m in zip(shards, means)) / R
            want = torch.cat([seeds[r] for r in real])
            assert abs(float(got) - float(vals[want].mean())) < 1e-12, (R, i)
            assert abs(sum(s.loss_scale(i) for s in shards) - R) < 1e-9
        # every real batch is trained exactly once per epoch
        trained = sorted(s.global_batch_index(i) for s in shards for i in range(steps) if not s.is_padded(i))
        assert trained == list(range(shards[0].num_batches_global))
