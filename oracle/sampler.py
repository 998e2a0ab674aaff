"""Fan-out neighbour sampler oracle — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``sample_block`` binds ``sampler_oracle.c`` (the sequential restatement of what the reference gets
from ``NeighborLoader`` -> pyg-lib ``neighbor_sample``; reference src/pipeline.py:75-83,152);
``sample_block_py`` is an independent pure-Python twin for tiny graphs used to pin the C code.
Both use the Philox stream documented in sampler_oracle.c, so they are also the bit-exact target
of the CUDA sampler.
"""
from __future__ import annotations

import ctypes
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from . import philox

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libngnn_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """gcc -O2 the plain-C oracle into oracle/_build/ (git-ignored)."""
    src = _DIR / "sampler_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", str(_SO), str(src)], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
        _lib.ngnn_oracle_sample_block.restype = ctypes.c_int
    return _lib


@dataclass
class Block:
    n_id: np.ndarray        # int32 [n] global ids, seeds first
    rowptr: np.ndarray      # int32 [n+1] CSR by destination over all local nodes
    col: np.ndarray         # int32 [e] local source ids
    col_global: np.ndarray  # int32 [e] global source ids
    e_pos: np.ndarray       # int32 [e] position in the CSC row[] array
    node_counts: np.ndarray  # int32 [H+1] cumulative nodes (node_counts[0] = bs)
    edge_counts: np.ndarray  # int32 [H+1] cumulative edges (edge_counts[0] = 0)

    @property
    def n(self) -> int:
        return int(self.node_counts[-1])

    @property
    def e(self) -> int:
        return int(self.edge_counts[-1])


def capacity(bs: int, fanouts, N: int):
    fr, nodes, edges = bs, bs, 0
    for f in fanouts:
        e = fr * f
        edges += e
        fr = min(e, N)
        nodes += fr
    return min(nodes, N + bs), edges


class CSampler:
    """Holds the N-sized relabel scratch so repeated calls do not re-allocate (used by the CPU baseline)."""

    def __init__(self, colptr: np.ndarray, row: np.ndarray):
        self.colptr = np.ascontiguousarray(colptr, dtype=np.int32)
        self.row = np.ascontiguousarray(row, dtype=np.int32)
        self.N = len(self.colptr) - 1
        self.local_of = np.full(self.N, -1, dtype=np.int32)
        self.lib = _load()

    def sample(self, seeds, fanouts, replace=False, seed=0, epoch=0, batch_idx=0) -> Block:
        seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        fan = np.ascontiguousarray(fanouts, dtype=np.int32)
        bs, H = len(seeds), len(fan)
        cap_n, cap_e = capacity(bs, fan.tolist(), self.N)
        n_id = np.empty(cap_n, np.int32)
        rowptr = np.empty(cap_n + 1, np.int32)
        col = np.empty(max(cap_e, 1), np.int32)
        colg = np.empty(max(cap_e, 1), np.int32)
        epos = np.empty(max(cap_e, 1), np.int32)
        counts = np.zeros(2 * (H + 1), np.int32)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = self.lib.ngnn_oracle_sample_block(
            P(self.colptr), P(self.row), ctypes.c_int64(self.N), P(seeds), ctypes.c_int32(bs), P(fan),
            ctypes.c_int32(H), ctypes.c_int32(int(replace)), ctypes.c_uint64(seed), ctypes.c_uint32(epoch),
            ctypes.c_uint32(batch_idx), P(n_id), P(rowptr), P(col), P(colg), P(epos), P(counts),
            ctypes.c_int64(cap_n), ctypes.c_int64(cap_e), P(self.local_of))
        if rc != 0:
            raise RuntimeError(f"ngnn_oracle_sample_block failed: {rc}")
        n, e = int(counts[H]), int(counts[2 * H + 1])
        return Block(n_id[:n].copy(), rowptr[:n + 1].copy(), col[:e].copy(), colg[:e].copy(), epos[:e].copy(),
                     counts[:H + 1].copy(), counts[H + 1:].copy())


def sample_block(colptr, row, seeds, fanouts, replace=False, seed=0, epoch=0, batch_idx=0) -> Block:
    return CSampler(colptr, row).sample(seeds, fanouts, replace, seed, epoch, batch_idx)


def _word(v, h, j, batch_idx, epoch, seed):
    w = philox.philox4x32_10(v, (h << 16) | (j >> 2), batch_idx, epoch, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return int(w[j & 3])


def sample_block_py(colptr, row, seeds, fanouts, replace=False, seed=0, epoch=0, batch_idx=0) -> Block:
    """Pure-Python twin of sampler_oracle.c (dict relabel map, python loops) — tiny graphs only."""
    colptr = np.asarray(colptr)
    row = np.asarray(row)
    local = {}
    n_id, col, colg, epos = [], [], [], []
    for s in seeds:
        local[int(s)] = len(n_id)
        n_id.append(int(s))
    rowptr = [0]
    node_counts, edge_counts = [len(n_id)], [0]
    lo, hi = 0, len(n_id)
    for h, fanout in enumerate(fanouts):
        for i in range(lo, hi):
            v = n_id[i]
            beg, d = int(colptr[v]), int(colptr[v + 1] - colptr[v])
            if replace:
                k = fanout if d > 0 else 0
                pos = [(_word(v, h, j, batch_idx, epoch, seed) * d) >> 32 for j in range(k)]
            elif d <= fanout:
                pos = list(range(d))
            else:
                pos = []
                for j in range(fanout):
                    jj = d - fanout + j
                    t = (_word(v, h, j, batch_idx, epoch, seed) * (jj + 1)) >> 32
                    pos.append(jj if t in pos else t)
            for p in pos:
                g = int(row[beg + p])
                if g not in local:
                    local[g] = len(n_id)
                    n_id.append(g)
                col.append(local[g])
                colg.append(g)
                epos.append(beg + p)
            rowptr.append(len(col))
        lo, hi = hi, len(n_id)
        node_counts.append(len(n_id))
        edge_counts.append(len(col))
    rowptr.extend([len(col)] * (len(n_id) - lo))
    i32 = lambda a: np.asarray(a, dtype=np.int32)
    return Block(i32(n_id), i32(rowptr), i32(col), i32(colg), i32(epos), i32(node_counts), i32(edge_counts))
