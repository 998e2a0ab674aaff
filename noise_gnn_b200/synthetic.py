"""Synthetic graphs with the dataset shapes the reference trains on, plus its label-noise law.

The reference loads ogbn-products / ogbn-arxiv / PubMed / Cora / Amazon-Computers from disk
(src/utils/load_utils.py:14-51) and corrupts the labels with ``flip_label`` (src/utils/noise.py:6-61).
There is no network here, so the benchmark and tests use random graphs of the same node / edge /
feature / class counts (BASELINE.json ``configs``), seeded with the value every reference YAML carries
(``seed: 1232``).  Pure torch tensor ops (plumbing, not the product); runs on CPU or GPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from .loader import Data


@dataclass(frozen=True)
class Shape:
    nodes: int
    undirected_edges: int      # stored as both directions => 2x directed entries
    features: int
    classes: int
    train_seeds: int
    feature_law: str           # 'normal' | 'sparse_rownorm' | 'bernoulli'
    hidden: int
    layers: int
    fanouts: tuple
    batch_size: int
    dropout: float


# shapes: BASELINE.json configs / SURVEY §8; model keys from the reference YAMLs
SHAPES = {
    # config/config_cora.yml (hidden 512, L 2, [10,5], bs 512)
    "cora": Shape(2_708, 5_278, 1_433, 7, 140, "bernoulli", 512, 2, (10, 5), 512, 0.5),
    # config/config_pubmed.yml (hidden 256, L 3, [10,5], full batch = 60 seeds, sym noise 0.3)
    "pubmed": Shape(19_717, 44_324, 500, 3, 60, "sparse_rownorm", 256, 3, (10, 5), 60, 0.5),
    # config/config_arxiv.yml (hidden 256, L 3, bs 512); fan-out [15,10] per BASELINE.json (YAML: [10,5])
    "arxiv": Shape(169_343, 1_166_243, 128, 40, 90_941, "normal", 256, 3, (15, 10), 512, 0.2),
    # config/config_products.yml (hidden 256, L 3, [15,10,5], bs 512, 196,615 train seeds)
    "products": Shape(2_449_029, 61_859_140, 100, 47, 196_615, "normal", 256, 3, (15, 10, 5), 512, 0.5),
    # Amazon-Computers shape for the aggregation sweep (load_utils.py:43-47)
    "computers": Shape(13_752, 245_861, 767, 10, 300, "normal", 256, 2, (10, 5), 512, 0.5),
}


def _endpoints(n_nodes: int, count: int, law: str, gen: torch.Generator, device) -> torch.Tensor:
    if law == "uniform":
        return torch.randint(0, n_nodes, (count,), generator=gen, device=device, dtype=torch.int64)
    # Chung-Lu style: P(node i) ~ (i + i0)^-alpha via inverse CDF, then ids are shuffled by the caller
    alpha, i0 = 0.6, 50.0
    w = (torch.arange(n_nodes, device=device, dtype=torch.float64) + i0).pow(-alpha)
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1]
    u = torch.rand(count, generator=gen, device=device, dtype=torch.float64)
    return torch.searchsorted(cdf, u).clamp_(max=n_nodes - 1)


def random_undirected_graph(n_nodes: int, n_undirected: int, law: str = "powerlaw", seed: int = 1232,
                            device="cpu") -> torch.Tensor:
    """Simple undirected graph (no self loops, no multi-edges) as int64 COO [2, 2*m] with both directions."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    max_pairs = n_nodes * (n_nodes - 1) // 2
    if n_undirected > max_pairs:
        raise ValueError("more edges requested than a simple graph can hold")
    keys = torch.empty(0, dtype=torch.int64, device=device)
    while keys.numel() < n_undirected:
        need = n_undirected - keys.numel()
        count = int(need * 1.15) + 1024
        u = _endpoints(n_nodes, count, law, gen, device)
        v = _endpoints(n_nodes, count, law, gen, device)
        ok = u != v
        lo, hi = torch.minimum(u, v)[ok], torch.maximum(u, v)[ok]
        keys = torch.unique(torch.cat([keys, lo * n_nodes + hi]))
    if keys.numel() > n_undirected:   # drop a random surplus (torch.unique sorts, so never truncate a prefix)
        keep = torch.randperm(keys.numel(), generator=gen, device=device)[:n_undirected]
        keys = keys[keep]
    relabel = torch.randperm(n_nodes, generator=gen, device=device)   # decorrelate degree from node id
    a, b = relabel[keys // n_nodes], relabel[keys % n_nodes]
    order = torch.randperm(keys.numel(), generator=gen, device=device)
    a, b = a[order], b[order]
    return torch.stack([torch.cat([a, b]), torch.cat([b, a])])


def make_features(n_nodes: int, n_feat: int, law: str, gen: torch.Generator, device) -> torch.Tensor:
    if law == "normal":
        return torch.randn(n_nodes, n_feat, generator=gen, device=device, dtype=torch.float32)
    if law == "bernoulli":       # bag-of-words 0/1 (Cora-like density)
        return (torch.rand(n_nodes, n_feat, generator=gen, device=device) < 0.013).float()
    if law == "sparse_rownorm":  # PubMed TF-IDF after T.NormalizeFeatures (load_utils.py:35): rows sum to 1
        x = torch.rand(n_nodes, n_feat, generator=gen, device=device)
        x = x * (torch.rand(n_nodes, n_feat, generator=gen, device=device) < 0.10)
        return (x / x.sum(1, keepdim=True).clamp(min=1e-12)).float()
    raise ValueError(law)


def noise_matrix(n_classes: int, noise_type: str, prob: float, gen: torch.Generator | None = None) -> torch.Tensor:
    """Row-stochastic class-transition matrix, the law of reference src/utils/noise.py:11-50."""
    eye = torch.eye(n_classes, dtype=torch.float64)
    if noise_type == "sym":
        return (1 - prob) * eye + (1 - eye) * (prob / (n_classes - 1))
    if noise_type == "next_pair":
        return (1 - prob) * eye + prob * torch.roll(eye, shifts=1, dims=1)
    if noise_type == "rand_pair":
        p1 = torch.randperm(n_classes, generator=gen)
        p2 = torch.randperm(n_classes, generator=gen)
        m = (1 - prob) * eye
        m[p1, p2] += prob
        return m
    raise ValueError(f"unknown noise type {noise_type!r}")


def flip_label(labels: torch.Tensor, n_classes: int, noise_type: str = "sym", prob: float = 0.3, seed: int = 1232):
    """Vectorised restatement of reference flip_label (src/utils/noise.py:6-61): each label is redrawn from
    its row of the transition matrix.  Returns (noisy_labels [N] int64, noise_mat [C,C] float64)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    mat = noise_matrix(n_classes, noise_type, prob, gen)
    y = labels.view(-1).cpu().long()
    rows = mat[y].clamp(min=0)
    noisy = torch.multinomial(rows, 1, generator=gen).view(-1)
    return noisy.to(labels.device), mat


def make_dataset(name: str, seed: int = 1232, law: str = "powerlaw", device="cpu", noise_type: str | None = None,
                 noise_rate: float = 0.3, scale: float = 1.0):
    """Synthetic stand-in for reference load_network (src/utils/load_utils.py:14-51) + flip_label.

    Returns (data, shape, train_idx).  `scale` < 1 shrinks node/edge/seed counts proportionally (tests)."""
    sh = SHAPES[name]
    n = max(int(sh.nodes * scale), 8)
    m = max(int(sh.undirected_edges * scale), 8)
    device = torch.device(device)
    ei = random_undirected_graph(n, m, law=law, seed=seed, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1)
    x = make_features(n, sh.features, sh.feature_law, gen, device)
    y = torch.randint(0, sh.classes, (n,), generator=gen, device=device, dtype=torch.int64)
    n_train = max(min(int(math.ceil(sh.train_seeds * scale)), n), 1)
    train_idx = torch.randperm(n, generator=gen, device=device)[:n_train].cpu()
    data = Data(x=x, edge_index=ei, y=y.view(-1, 1) if name in ("arxiv", "products") else y)
    if noise_type is not None:
        data.yhn, data.noise_mat = flip_label(y, sh.classes, noise_type, noise_rate, seed=seed + 2)
    else:
        data.yhn = y.clone()
    return data, sh, train_idx
