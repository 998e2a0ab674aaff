"""CPU tests: the C-ABI library builds for sm_100a, loads, and exports every symbol include/ngnn_b200.h declares."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "ngnn_b200.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"\b(ngnn_[a-z0-9_]+)\s*\(", HEADER)))


def test_library_builds_and_exports_every_declared_symbol():
    from noise_gnn_b200 import _build, _lib
    path = _build.build()
    assert path.exists()
    lib = ctypes.CDLL(str(path))
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ngnn_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), "python signatures and header disagree"


def test_version_and_error_string_without_gpu():
    from noise_gnn_b200 import _lib
    lib = _lib.load()
    assert lib.ngnn_version() >= 100
    # argument validation happens before any CUDA call, so it works without a device
    rc = lib.ngnn_sage_agg_fwd(None, None, None, 0, -1, 4, None, 0, None, None, 0, None)
    assert rc == -1
    assert "negative" in _lib.last_error()
    with pytest.raises(_lib.NgnnError):
        _lib.call("ngnn_sage_gemm_fwd", None, 0, None, 0, None, None, None, 4, 4, 4, 7, 0.0, 0, 0, None, 0, None, None, 0, None)


def test_library_contains_sm100a_code_only():
    from noise_gnn_b200 import _build
    out = subprocess.run(["cuobjdump", "-lelf", str(_build.build())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_cpu_tensors_are_rejected_loudly():
    import torch
    from noise_gnn_b200 import SAGEConv
    conv = SAGEConv(4, 3)
    assert sorted(conv.state_dict()) == ["lin_l.bias", "lin_l.weight", "lin_r.weight"]
    assert conv.lin_l.weight.shape == (3, 4) and conv.lin_r.weight.shape == (3, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv(torch.randn(5, 4), torch.zeros(2, 3, dtype=torch.long))


def test_product_path_never_imports_oracle():
    for f in (ROOT / "noise_gnn_b200").rglob("*.py"):
        txt = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
    for f in (ROOT / "noise_gnn_b200" / "csrc").glob("*.cu*"):
        includes = [l for l in f.read_text().splitlines() if l.lstrip().startswith("#include")]
        assert not any("oracle" in l for l in includes), f


@pytest.mark.parametrize("bs,fan,N", [(512, [15, 10, 5], 2_449_029), (512, [15, 10], 169_343), (60, [10, 5], 19_717),
                                      (512, [10, 5], 2_708), (7, [3, 3, 3, 3], 50), (4096, [25], 13_752)])
def test_host_capacity_math_matches_the_library(bs, fan, N):
    """The block capacities the trainer sizes its arena with (noise_gnn_b200.train.hop_capacities, pure Python) against the
    sampler's own worst case (ngnn_sample_capacity — host arithmetic, no CUDA call): the arena must cover every block the
    sampler can emit."""
    from noise_gnn_b200 import _lib
    from noise_gnn_b200.train import hop_capacities
    lib = _lib.load()
    fan_c = (ctypes.c_int32 * len(fan))(*fan)
    mn, me = ctypes.c_int64(), ctypes.c_int64()
    assert lib.ngnn_sample_capacity(bs, fan_c, len(fan), N, ctypes.byref(mn), ctypes.byref(me)) == 0
    nodes, edges = hop_capacities(bs, fan, N)
    assert len(nodes) == len(edges) == len(fan) + 1 and nodes[0] == bs and edges[0] == 0
    assert all(a <= b for a, b in zip(nodes, nodes[1:])) and all(a <= b for a, b in zip(edges, edges[1:]))
    assert nodes[-1] >= mn.value and edges[-1] >= me.value
    # per-hop prefixes of the same law: asking the library for h hops must never exceed the h-th cumulative capacity
    for h in range(1, len(fan) + 1):
        assert lib.ngnn_sample_capacity(bs, fan_c, h, N, ctypes.byref(mn), ctypes.byref(me)) == 0
        assert nodes[h] >= mn.value and edges[h] >= me.value


@pytest.mark.parametrize("L,F,Hd,C", [(3, 100, 256, 47), (2, 1433, 512, 7), (3, 500, 256, 3), (3, 128, 256, 40), (4, 8, 16, 4)])
def test_flat_bucket_layout_and_arena_size_without_gpu(L, F, Hd, C):
    """ngnn_sage_num_params = the parameter count of the reference-shaped SAGE stack (lin_l.weight, lin_l.bias,
    lin_r.weight per layer), and the step arena grows with the declared block capacity (host arithmetic)."""
    import torch
    from noise_gnn_b200 import SAGE, _lib
    from noise_gnn_b200.train import hop_capacities
    lib = _lib.load()
    lib.ngnn_sage_num_params.restype = ctypes.c_int64
    ms = _lib.SageModel(L, F, Hd, C, 0.5, 1)
    net = SAGE(F, Hd, C, L, dropout=0.5)
    assert lib.ngnn_sage_num_params(ctypes.byref(ms)) == sum(p.numel() for p in net.parameters())
    assert [k for k, _ in net.named_parameters()][:3] == ["convs.0.lin_l.weight", "convs.0.lin_l.bias", "convs.0.lin_r.weight"]
    sizes = []
    for bs in (64, 512):
        nodes, edges = hop_capacities(bs, [10, 5], 100_000)
        mn = (ctypes.c_int64 * 3)(*nodes)
        me = (ctypes.c_int64 * 3)(*edges)
        sizes.append(lib.ngnn_sage_step_workspace_bytes(ctypes.byref(ms), 2, mn, me))
    assert 0 < sizes[0] < sizes[1]
