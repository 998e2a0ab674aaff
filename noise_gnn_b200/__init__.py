"""noise_gnn_b200 — B200-native GraphSAGE mini-batch hot path of hhilsber/noise-GNN.

(The task names the package ``noise-gnn_b200``; a hyphen is not a legal Python identifier, hence the
underscore.)  Public surface, mirroring the two third-party entry points the reference's hot path calls:

* ``SAGEConv``        drop-in for ``torch_geometric.nn.SAGEConv``            (reference src/models/layers/sage.py:4,16-19,34)
* ``NeighborLoader``  drop-in for ``torch_geometric.loader.NeighborLoader``  (reference src/pipeline.py:6,75-92,152)
* ``Data``            minimal ``torch_geometric.data.Data`` bag
* ``SAGE``            the reference's network module (src/models/layers/sage.py:6-78) with a trimmed fused mode
* ``SAGEPL`` / ``shuffle_pos``  the noise-injecting SAGE variant (src/models/layers/sagePL.py) and the feature-shuffling
                      augmentation (src/utils/augmentation.py:88-102): fused gather-normalize-add kernel, device-side shuffle
* ``GCNConv`` / ``SimpleGCN``  the reference's ``module: 'gcn'`` path (src/models/layers/convolution.py:7-53,
                      ``GCNConv(normalize=False)``): same kernels, sum instead of mean, linear before aggregation
* ``ops``             tensor-level wrappers over the C ABI in include/ngnn_b200.h (libngnn_b200.so)

Everything computes in hand-written sm_100a CUDA behind a C ABI; there is no CPU or library fallback —
importing works anywhere, calling without the built library or without a B200 raises.
"""
from . import _build, _lib  # noqa: F401


def build(force: bool = False, verbose: bool = False):
    """Compile csrc/*.cu for sm_100a into noise_gnn_b200/libngnn_b200.so."""
    return _build.build(force=force, verbose=verbose)


def __getattr__(name):
    # torch-dependent modules are imported lazily so `import noise_gnn_b200` stays cheap
    if name in ("SAGEConv", "GCNConv"):
        from . import conv
        return getattr(conv, name)
    if name == "SimpleGCN":
        from .gcn import SimpleGCN
        return SimpleGCN
    if name in ("NeighborLoader", "Data", "Batch"):
        from . import loader
        return getattr(loader, name)
    if name == "SAGE":
        from .sage import SAGE
        return SAGE
    if name in ("SAGEPL", "shuffle_pos"):
        from . import sagepl
        return getattr(sagepl, name)
    if name == "CTLoss":
        from .losses import CTLoss
        return CTLoss
    if name in ("ops", "conv", "loader", "sage", "sagepl", "gcn", "losses", "synthetic", "train", "dp"):
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)


__all__ = ["SAGEConv", "GCNConv", "NeighborLoader", "Data", "Batch", "SAGE", "SAGEPL", "shuffle_pos", "SimpleGCN", "CTLoss", "build"]
