"""Tensor-level wrappers over the C ABI (include/ngnn_b200.h) and the SAGEConv autograd function.

torch is used for device memory, streams and autograd bookkeeping only; every computation below is a
call into libngnn_b200.so on ``torch.cuda.current_stream()``.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib

_I32 = torch.int32
_F32 = torch.float32


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (raw handle: the Python Stream object costs ~14 us,
    and every ABI call needs one)."""
    if _raw_stream is not None and _cur_device is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("noise_gnn_b200 kernels run on CUDA tensors only (there is no CPU fallback)")


def _rows(t: torch.Tensor, name: str) -> torch.Tensor:
    """fp32 2-D tensor with unit inner stride (row-major with a leading dimension)."""
    if t.dtype != _F32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"{name}: expected a 2-D tensor, got {tuple(t.shape)}")
    if t.size(1) > 1 and t.stride(1) != 1 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.size(0) > 1 else max(t.size(1), 1)


class _Workspaces:
    """Grow-only scratch buffers, one per (device, stream, tag), so stream order protects reuse."""

    def __init__(self):
        self._bufs = {}

    def get(self, nbytes: int, device: torch.device, tag: str) -> torch.Tensor:
        key = (device.index, _stream(), tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


workspaces = _Workspaces()

# Optional CUDA-event timing of individual ABI calls (bench.py): `timers` maps a tag to a list of
# (start, end) events; only tags already present are recorded unless `timers_open` is set.
timers = None
timers_open = False


class _timed:
    def __init__(self, tag: str):
        self.on = timers is not None and (timers_open or tag in timers)
        self.tag = tag

    def __enter__(self):
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.on:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            timers.setdefault(self.tag, []).append((self.a, b))


# ----------------------------------------------------------------------------- structure
class Block:
    """A message-flow block in CSR by destination (int32), with lazily built transposes.

    rowptr [n_rows+1], col [e] local source ids.  ``hop_nodes`` / ``hop_edges`` are the cumulative
    per-hop counts of a sampled block (host ints) when it came from the sampler; ``col_global`` and
    ``n_id`` let layer 1 aggregate straight from the resident feature table.
    """

    def __init__(self, rowptr, col, n_rows: int, e: int, n_cols: Optional[int] = None, perm=None,
                 hop_nodes=None, hop_edges=None, col_global=None, n_id=None):
        self.rowptr, self.col, self.perm = rowptr, col, perm
        self.n_rows, self.e = int(n_rows), int(e)
        self.n_cols = int(n_cols if n_cols is not None else n_rows)
        self.hop_nodes, self.hop_edges = hop_nodes, hop_edges
        self.col_global, self.n_id = col_global, n_id
        self.col_table = self.n_table = None      # table rows of col_global / n_id when the loader keeps a hot-rows-first table
        self._t = {}

    def transpose(self, e_limit: Optional[int] = None, n_cols: Optional[int] = None):
        """(colptr_t, row_t) of the first e_limit edges, columns 0..n_cols-1 (CSC by source)."""
        e_limit = self.e if e_limit is None else int(e_limit)
        n_cols = self.n_cols if n_cols is None else int(n_cols)
        key = (e_limit, n_cols)
        if key not in self._t:
            self._t[key] = csr_transpose(self.rowptr, self.col, self.n_rows, e_limit, n_cols)[:2]
        return self._t[key]


def coo_to_csr(edge_index: torch.Tensor, n_rows: int) -> Block:
    """Stable sort by destination of a PyG COO edge_index (int64 [2,e], any order, duplicates kept)."""
    _check_cuda(edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must be an int64 tensor of shape [2, e]")
    ei = edge_index.contiguous()
    e = ei.size(1)
    dev = ei.device
    rowptr = torch.empty(n_rows + 1, dtype=_I32, device=dev)
    col = torch.empty(e, dtype=_I32, device=dev)
    perm = torch.empty(e, dtype=_I32, device=dev)
    L = _lib.load()
    nbytes = L.ngnn_coo_to_csr_workspace_bytes(e, n_rows)
    ws = workspaces.get(nbytes, dev, "sort")
    with _timed("coo_to_csr"):
        _lib.call("ngnn_coo_to_csr", _ptr(ei[0]), _ptr(ei[1]), e, n_rows, _ptr(rowptr), _ptr(col), _ptr(perm),
                  _ptr(ws), ws.numel(), _stream())
    return Block(rowptr, col, n_rows, e, perm=perm)


def csr_transpose(rowptr, col, n_rows: int, e_limit: int, n_cols: int):
    _check_cuda(rowptr, col)
    dev = rowptr.device
    colptr_t = torch.empty(n_cols + 1, dtype=_I32, device=dev)
    row_t = torch.empty(e_limit, dtype=_I32, device=dev)
    perm_t = torch.empty(e_limit, dtype=_I32, device=dev)
    L = _lib.load()
    nbytes = L.ngnn_csr_transpose_workspace_bytes(e_limit, n_cols)
    ws = workspaces.get(nbytes, dev, "sort")
    with _timed("transpose"):
        _lib.call("ngnn_csr_transpose", _ptr(rowptr), _ptr(col), n_rows, e_limit, n_cols, _ptr(colptr_t), _ptr(row_t),
                  _ptr(perm_t), _ptr(ws), ws.numel(), _stream())
    return colptr_t, row_t, perm_t


def csr_to_coo(rowptr, col, n_rows: int, e: int) -> torch.Tensor:
    _check_cuda(rowptr, col)
    ei = torch.empty((2, e), dtype=torch.int64, device=rowptr.device)
    _lib.call("ngnn_csr_to_coo", _ptr(rowptr), _ptr(col), n_rows, e, _ptr(ei), _stream())
    return ei


def gather_rows(table: torch.Tensor, idx: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
    _check_cuda(table, idx)
    table = _rows(table, "table")
    n = idx.numel() if n is None else n
    out = torch.empty((n, table.size(1)), dtype=_F32, device=table.device)
    with _timed("gather_rows"):
        _lib.call("ngnn_gather_rows", _ptr(table), _ld(table), _ptr(idx), n, table.size(1), _ptr(out), _ld(out), _stream())
    return out


# ----------------------------------------------------------------------------- kernels
def agg_fwd(rowptr, col, x, n_dst: int, root_idx=None, out=None, root_out=None, tag=None):
    """mean[i] = 1/max(deg,1) * sum x[col[p]]; optional fused root gather root[i] = x[root_idx[i]]."""
    _check_cuda(rowptr, col, x)
    x = _rows(x, "x")
    F_ = x.size(1)
    mean = out if out is not None else torch.empty((n_dst, F_), dtype=_F32, device=x.device)
    root = None
    if root_idx is not None:
        root = root_out if root_out is not None else torch.empty((n_dst, F_), dtype=_F32, device=x.device)
    with _timed(tag or "agg_fwd"):
        _lib.call("ngnn_sage_agg_fwd", _ptr(rowptr), _ptr(col), _ptr(x), _ld(x), n_dst, F_, _ptr(mean), _ld(mean),
                  _ptr(root_idx), _ptr(root), _ld(root) if root is not None else 0, _stream())
    return (mean, root) if root_idx is not None else mean


def gcn_agg_fwd(rowptr, col, z, n_dst: int, bias=None, out=None, tag=None):
    """out[i] = sum_{p} z[col[p]] + bias  (GCNConv(normalize=False) propagation: sum, duplicates counted, no self loops)."""
    _check_cuda(rowptr, col, z, bias)
    z = _rows(z, "z")
    O = z.size(1)
    y = out if out is not None else torch.empty((n_dst, O), dtype=_F32, device=z.device)
    with _timed(tag or "gcn_agg_fwd"):
        _lib.call("ngnn_gcn_agg_fwd", _ptr(rowptr), _ptr(col), _ptr(z), _ld(z), n_dst, O,
                  _ptr(bias.contiguous() if bias is not None else None), _ptr(y), _ld(y), _stream())
    return y


def agg_bwd(colptr_t, row_t, dmean_scaled, n_src: int, dx_root=None, n_root: int = 0, act_ref=None,
            act_scale: float = 1.0, out=None, tag=None):
    _check_cuda(colptr_t, row_t, dmean_scaled)
    dmean_scaled = _rows(dmean_scaled, "dmean_scaled")
    F_ = dmean_scaled.size(1)
    dx = out if out is not None else torch.empty((n_src, F_), dtype=_F32, device=dmean_scaled.device)
    if dx_root is not None:
        dx_root = _rows(dx_root, "dx_root")
    if act_ref is not None:
        act_ref = _rows(act_ref, "act_ref")
    with _timed(tag or "agg_bwd"):
        _lib.call("ngnn_sage_agg_bwd", _ptr(colptr_t), _ptr(row_t), _ptr(dmean_scaled), _ld(dmean_scaled), n_src, F_,
                  _ptr(dx_root), _ld(dx_root) if dx_root is not None else 0, n_root,
                  _ptr(act_ref), _ld(act_ref) if act_ref is not None else 0, float(act_scale), _ptr(dx), _ld(dx), _stream())
    return dx


def gemm_fwd(a_l, a_r, w_l, w_r, bias, n: int, act: int = 0, drop_p: float = 0.0, seed: int = 0, offset: int = 0,
             out=None, return_path: bool = False, tag=None):
    """out = drop(act(a_l @ w_l.T + a_r @ w_r.T + bias)) over the first n rows."""
    ref = a_l if a_l is not None else a_r
    _check_cuda(ref, w_l, w_r, bias)
    if a_l is not None:
        a_l = _rows(a_l, "a_l")
    if a_r is not None:
        a_r = _rows(a_r, "a_r")
    w = w_l if w_l is not None else w_r
    O, F_ = w.shape
    w_l = None if w_l is None else w_l.contiguous()
    w_r = None if w_r is None else w_r.contiguous()
    y = out if out is not None else torch.empty((n, O), dtype=_F32, device=ref.device)
    path = ctypes.c_int32(0)
    ws = workspaces.get(_lib.load().ngnn_sage_gemm_workspace_bytes(F_, O), ref.device, "gemm")
    with _timed(tag or "gemm_fwd"):
        _lib.call("ngnn_sage_gemm_fwd", _ptr(a_l), _ld(a_l) if a_l is not None else 0, _ptr(a_r),
                  _ld(a_r) if a_r is not None else 0, _ptr(w_l), _ptr(w_r), _ptr(bias), n, F_, O, int(act), float(drop_p),
                  int(seed) & (2**64 - 1), int(offset) & (2**64 - 1), _ptr(y), _ld(y), ctypes.byref(path),
                  _ptr(ws), ws.numel(), _stream())
    return (y, path.value) if return_path else y


def dgrad(dy, w_l, w_r, rowptr, n: int, want_mean: bool = True, want_root: bool = True, tag=None):
    _check_cuda(dy)
    dy = _rows(dy, "dy")
    w = w_l if w_l is not None else w_r
    O, F_ = w.shape
    dev = dy.device
    dmean = torch.empty((n, F_), dtype=_F32, device=dev) if want_mean else None
    droot = torch.empty((n, F_), dtype=_F32, device=dev) if want_root else None
    ws = workspaces.get(_lib.load().ngnn_sage_dgrad_workspace_bytes(F_, O), dev, "dgrad")
    with _timed(tag or "dgrad"):
        _lib.call("ngnn_sage_dgrad", _ptr(dy), _ld(dy), _ptr(w_l.contiguous() if w_l is not None else None),
                  _ptr(w_r.contiguous() if w_r is not None else None), _ptr(rowptr), n, F_, O,
                  _ptr(dmean), _ld(dmean) if dmean is not None else 0, _ptr(droot), _ld(droot) if droot is not None else 0,
                  _ptr(ws), ws.numel(), _stream())
    return dmean, droot


def wgrad(dy, a_l, a_r, n: int, F_: int, dw_l=None, dw_r=None, db=None, accumulate: bool = False,
          want_l: bool = True, want_r: bool = True, want_b: bool = True, tag=None):
    _check_cuda(dy)
    dy = _rows(dy, "dy")
    O = dy.size(1)
    dev = dy.device
    if a_l is not None:
        a_l = _rows(a_l, "a_l")
    if a_r is not None:
        a_r = _rows(a_r, "a_r")
    if dw_l is None and want_l and a_l is not None:
        dw_l = torch.empty((O, F_), dtype=_F32, device=dev)
    if dw_r is None and want_r and a_r is not None:
        dw_r = torch.empty((O, F_), dtype=_F32, device=dev)
    if db is None and want_b:
        db = torch.empty((O,), dtype=_F32, device=dev)
    L = _lib.load()
    nbytes = L.ngnn_sage_wgrad_workspace_bytes(n, F_, O)
    ws = workspaces.get(nbytes, dev, "wgrad")
    with _timed(tag or "wgrad"):
        _lib.call("ngnn_sage_wgrad", _ptr(dy), _ld(dy), _ptr(a_l), _ld(a_l) if a_l is not None else 0, _ptr(a_r),
                  _ld(a_r) if a_r is not None else 0, n, F_, O, _ptr(dw_l), _ptr(dw_r), _ptr(db), int(accumulate),
                  _ptr(ws), ws.numel(), _stream())
    return dw_l, dw_r, db


def act_bwd(dh, h, scale: float = 1.0):
    _check_cuda(dh, h)
    dh, h = _rows(dh, "dh"), _rows(h, "h")
    dz = torch.empty_like(dh, memory_format=torch.contiguous_format)
    with _timed("act_bwd"):
        _lib.call("ngnn_act_bwd", _ptr(dh), _ld(dh), _ptr(h), _ld(h), dh.size(0), dh.size(1), float(scale), _ptr(dz),
                  _ld(dz), _stream())
    return dz


def ce_fwd_bwd(logits, target, y_true=None, grad_scale: float = 1.0, stats=None, want_grad: bool = True):
    """Mean CE over the rows of `logits`; stats[0] += loss, stats[1] += #correct; returns (stats, dlogits)."""
    _check_cuda(logits, target)
    logits = _rows(logits, "logits")
    bs, C = logits.shape
    if target.dtype != torch.int64:
        target = target.long()
    target = target.contiguous()
    if y_true is not None:
        y_true = y_true.long().contiguous() if y_true.dtype != torch.int64 else y_true.contiguous()
    if stats is None:
        stats = torch.zeros(2, dtype=_F32, device=logits.device)
    dlogits = torch.empty((bs, C), dtype=_F32, device=logits.device) if want_grad else None
    scratch = torch.empty(2 * bs, dtype=_F32, device=logits.device)
    with _timed("ce"):
        _lib.call("ngnn_ce_fwd_bwd", _ptr(logits), _ld(logits), _ptr(target), _ptr(y_true), bs, C, float(grad_scale),
                  _ptr(stats), _ptr(dlogits), _ld(dlogits) if dlogits is not None else 0, _ptr(scratch), _stream())
    return stats, dlogits


def ct_loss(logits1, logits2, target, num_remember: int, y_true=None, row_ids=None, clean_mask=None, stats=None,
            want_grad: bool = True, want_order: bool = False):
    """Co-teaching loss (reference src/utils/losses.py:10-49) on the device.  Returns (stats[6], dlogits1, dlogits2,
    order1, order2): stats += [loss_1, loss_2, correct_1, correct_2, pure_ratio_1, pure_ratio_2]."""
    _check_cuda(logits1, logits2, target)
    logits1, logits2 = _rows(logits1, "logits1"), _rows(logits2, "logits2")
    bs, C = logits1.shape
    dev = logits1.device
    target = target.contiguous()
    if target.dtype != torch.int64:
        target = target.long()
    if y_true is not None:
        y_true = y_true.long().contiguous() if y_true.dtype != torch.int64 else y_true.contiguous()
    if clean_mask is not None:
        clean_mask = clean_mask.to(torch.uint8).contiguous()
    if stats is None:
        stats = torch.zeros(6, dtype=_F32, device=dev)
    d1 = torch.empty((bs, C), dtype=_F32, device=dev) if want_grad else None
    d2 = torch.empty((bs, C), dtype=_F32, device=dev) if want_grad else None
    o1 = torch.empty(bs, dtype=_I32, device=dev) if want_order else None
    o2 = torch.empty(bs, dtype=_I32, device=dev) if want_order else None
    scratch = torch.empty(12 * max(bs, 1), dtype=_F32, device=dev)
    with _timed("ct_loss"):
        _lib.call("ngnn_ct_loss", _ptr(logits1), _ld(logits1), _ptr(logits2), _ld(logits2), _ptr(target), _ptr(y_true),
                  _ptr(row_ids), _ptr(clean_mask), bs, C, int(num_remember), _ptr(stats), _ptr(d1), _ld(d1) if d1 is not None else 0,
                  _ptr(d2), _ld(d2) if d2 is not None else 0, _ptr(o1), _ptr(o2), _ptr(scratch), _stream())
    return stats, d1, d2, o1, o2


def adam_step(param, grad, exp_avg, exp_avg_sq, step_dev, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
              grad_scale=1.0, advance_step=True):
    _check_cuda(param, grad, exp_avg, exp_avg_sq, step_dev)
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not t.is_contiguous() or t.dtype != _F32:
            raise ValueError("adam_step expects contiguous float32 flat buckets")
    with _timed("adam"):
        _lib.call("ngnn_adam_step", _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(lr),
                  float(betas[0]), float(betas[1]), float(eps), float(weight_decay), float(grad_scale), _ptr(step_dev),
                  (2 if step_dev.numel() >= 2 else 1) if advance_step else 0, _stream())


# ----------------------------------------------------------------------------- SAGEPL extras
class NoiseAddFunction(torch.autograd.Function):
    """noisy_x = x + s * rate * normalize(noise[idx]) (reference sagePL.py:41-49), one fused pass each way."""

    @staticmethod
    def forward(ctx, x, noise, idx, rate: float, use_sign: bool):
        x, noise = _rows(x, "x"), _rows(noise, "noise")
        n, F_ = x.shape
        out = torch.empty((n, F_), dtype=_F32, device=x.device)
        _lib.call("ngnn_noise_add_fwd", _ptr(x), _ld(x), _ptr(noise), _ld(noise), _ptr(idx), n, F_, float(rate), int(use_sign),
                  _ptr(out), _ld(out), _stream())
        ctx.save_for_backward(x if use_sign else None, noise, idx)
        ctx.rate, ctx.use_sign = float(rate), bool(use_sign)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, noise, idx = ctx.saved_tensors
        dout = _rows(dout, "dout")
        n, F_ = dout.shape
        dnoise = None
        if ctx.needs_input_grad[1]:
            dnoise = torch.zeros_like(noise)           # dense, like the reference's index_select backward
            _lib.call("ngnn_noise_add_bwd", _ptr(dout), _ld(dout), _ptr(x), _ld(x) if x is not None else 0, _ptr(noise), _ld(noise),
                      _ptr(idx), n, F_, ctx.rate, int(ctx.use_sign), _ptr(dnoise), _ld(dnoise), _stream())
        return (dout if ctx.needs_input_grad[0] else None), dnoise, None, None, None


def shuffle_rows(x: torch.Tensor, k: int, seed: int = 1232, offset: int = 0) -> torch.Tensor:
    """Per row, k distinct random positions get their values permuted among themselves (shuffle_pos on the device)."""
    _check_cuda(x)
    x = _rows(x, "x")
    out = torch.empty_like(x, memory_format=torch.contiguous_format)
    _lib.call("ngnn_shuffle_rows", _ptr(x), _ld(x), x.size(0), x.size(1), int(k), int(seed) & (2**64 - 1), int(offset) & (2**64 - 1),
              _ptr(out), _ld(out), _stream())
    return out


# ----------------------------------------------------------------------------- SAGEConv autograd
class SAGEConvFunction(torch.autograd.Function):
    """out[:n_dst] = lin_l(mean_{j->i} x_j) + lin_r(x_i), on a CSR block; backward is atomic-free."""

    @staticmethod
    def forward(ctx, x, w_l, b_l, w_r, block: Block, n_dst: int, e_limit: int):
        x = _rows(x, "x")
        mean = agg_fwd(block.rowptr, block.col, x, n_dst)
        out = gemm_fwd(mean, x if w_r is not None else None, w_l, w_r, b_l, n_dst)
        ctx.save_for_backward(x, w_l, w_r, mean)
        ctx.block, ctx.n_dst, ctx.e_limit = block, n_dst, e_limit
        ctx.has_bias = b_l is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w_l, w_r, mean = ctx.saved_tensors
        block, n_dst, e_limit = ctx.block, ctx.n_dst, ctx.e_limit
        need_x, need_wl, need_b, need_wr = ctx.needs_input_grad[:4]
        dy = _rows(dy, "dy")
        F_ = x.size(1)
        dw_l = dw_r = db = dx = None
        if need_wl or need_wr or (need_b and ctx.has_bias):
            dw_l, dw_r, db = wgrad(dy, mean if need_wl else None, x if (need_wr and w_r is not None) else None, n_dst,
                                   F_, want_l=need_wl, want_r=need_wr and w_r is not None,
                                   want_b=need_b and ctx.has_bias)
        if need_x:
            dmean, droot = dgrad(dy, w_l, w_r, block.rowptr, n_dst, want_root=w_r is not None)
            colptr_t, row_t = block.transpose(e_limit, x.size(0))
            dx = agg_bwd(colptr_t, row_t, dmean, x.size(0), dx_root=droot, n_root=n_dst)
        return dx, dw_l, db, dw_r, None, None, None


# ----------------------------------------------------------------------------- GCNConv autograd
class GCNConvFunction(torch.autograd.Function):
    """out = A_sum (x W^T) + b on a CSR block (PyG GCNConv(normalize=False): linear first, then sum over in-neighbours).

    forward:  z = x W^T  (K-GEMM, single operand)  ->  out = segment-sum(z) + b   (K-AGG without the 1/deg scale)
    backward: db = colsum(dOut);  dZ = A^T dOut (K-AGG-T);  dW = dZ^T x (K-WGRAD);  dX = dZ W (K-DGRAD)."""

    @staticmethod
    def forward(ctx, x, weight, bias, block: Block):
        x = _rows(x, "x")
        n = x.size(0)
        z = gemm_fwd(None, x, None, weight, None, n)
        out = gcn_agg_fwd(block.rowptr, block.col, z, n, bias=bias)
        ctx.save_for_backward(x, weight)
        ctx.block, ctx.has_bias = block, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight = ctx.saved_tensors
        block = ctx.block
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        dout = _rows(dout, "dout")
        n, F_ = x.size(0), x.size(1)
        dx = dw = db = None
        if need_b and ctx.has_bias:
            _, _, db = wgrad(dout, None, None, n, F_, want_l=False, want_r=False, want_b=True)
        if need_x or need_w:
            colptr_t, row_t = block.transpose(block.e, n)
            dz = agg_bwd(colptr_t, row_t, dout, n)
            if need_w:
                _, dw, _ = wgrad(dz, None, x, n, F_, want_l=False, want_r=True, want_b=False)
            if need_x:
                _, dx = dgrad(dz, None, weight, None, n, want_mean=False, want_root=True)
        return dx, dw, db, None
