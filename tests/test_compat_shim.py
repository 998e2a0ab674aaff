"""The reference's own network module (src/models/layers/sage.py, unmodified) resolves its PyG imports to the
B200-native drop-ins through compat/torch_geometric.  /root/reference only exists in the builder container."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SAGE = "/root/reference/src/models/layers/sage.py"


@pytest.fixture()
def shim_on_path():
    saved = {k: v for k, v in sys.modules.items() if k == "torch_geometric" or k.startswith("torch_geometric.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    yield
    sys.path.remove(os.path.join(ROOT, "compat"))
    for k in [k for k in sys.modules if k == "torch_geometric" or k.startswith("torch_geometric.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_shim_exposes_only_the_hot_path(shim_on_path):
    import torch_geometric
    from torch_geometric.data import Data
    from torch_geometric.loader import NeighborLoader
    from torch_geometric.nn import GCNConv, SAGEConv
    import noise_gnn_b200
    assert SAGEConv is noise_gnn_b200.SAGEConv and NeighborLoader is noise_gnn_b200.NeighborLoader and Data is noise_gnn_b200.Data
    assert GCNConv is noise_gnn_b200.GCNConv
    with pytest.raises(NotImplementedError):
        GCNConv(4, 4)                               # normalize=True (PyG's default) is not what the reference builds
    assert sorted(GCNConv(4, 4, normalize=False).state_dict()) == ["bias", "lin.weight"]
    from torch_geometric.datasets import Planetoid
    with pytest.raises(NotImplementedError):
        Planetoid(root="x", name="pubmed")


@pytest.mark.skipif(not os.path.exists(REF_SAGE), reason="reference tree only exists in the builder container")
def test_reference_sage_module_builds_on_the_drop_in(shim_on_path):
    spec = importlib.util.spec_from_file_location("ref_sage", REF_SAGE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import noise_gnn_b200
    net = mod.SAGE(100, 256, 47, 3, dropout=0.5)
    assert all(isinstance(c, noise_gnn_b200.SAGEConv) for c in net.convs)
    ours = noise_gnn_b200.SAGE(100, 256, 47, 3, dropout=0.5)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    ours.load_state_dict(net.state_dict())      # state_dicts interchange (PyG parameter names)
    net.reset_parameters()


REF_GCN = "/root/reference/src/models/layers/convolution.py"


@pytest.mark.skipif(not os.path.exists(REF_GCN), reason="reference tree only exists in the builder container")
def test_reference_simplegcn_module_builds_on_the_drop_in(shim_on_path):
    """reference src/models/layers/convolution.py (unmodified) constructs on the drop-in GCNConv, with the same
    state_dict layout as noise_gnn_b200.SimpleGCN."""
    spec = importlib.util.spec_from_file_location("ref_gcn", REF_GCN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import noise_gnn_b200
    net = mod.SimpleGCN(128, 256, 40, 3, dropout=0.5)
    assert all(isinstance(c, noise_gnn_b200.GCNConv) for c in net.convs)
    ours = noise_gnn_b200.SimpleGCN(128, 256, 40, 3, dropout=0.5)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    ours.load_state_dict(net.state_dict())
    net.reset_parameters()
