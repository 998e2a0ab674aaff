"""Block-structure oracle (numpy) — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates what PyG does around ``SAGEConv`` / ``NeighborLoader`` with the COO ``edge_index`` the
reference passes in (reference src/models/layers/sage.py:34; unsorted COO from
src/utils/augmentation.py:82-86): a STABLE sort by destination gives CSR, a stable sort of the CSR
order by source gives the transpose used by the backward.  These are the bit-exact targets of
``ngnn_coo_to_csr`` / ``ngnn_csr_transpose``.
"""
from __future__ import annotations

import numpy as np


def coo_to_csr(src: np.ndarray, dst: np.ndarray, n_rows: int):
    """perm = argsort_stable(dst); col = src[perm]; rowptr[i] = #edges with dst < i."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    perm = np.argsort(dst, kind="stable")
    col = src[perm]
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr.astype(np.int32), col.astype(np.int32), perm.astype(np.int32)


def csr_rows(rowptr: np.ndarray) -> np.ndarray:
    """Destination row of every CSR position."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    return np.repeat(np.arange(len(rowptr) - 1, dtype=np.int64), np.diff(rowptr))


def csr_transpose(rowptr: np.ndarray, col: np.ndarray, n_cols: int, e_limit: int | None = None):
    """Transpose of the first e_limit CSR edges: perm_t = argsort_stable(col[:e]); row_t = rows[perm_t]."""
    col = np.asarray(col, dtype=np.int64)
    e = len(col) if e_limit is None else e_limit
    rows = csr_rows(rowptr)[:e]
    perm_t = np.argsort(col[:e], kind="stable")
    row_t = rows[perm_t]
    colptr_t = np.zeros(n_cols + 1, dtype=np.int64)
    np.add.at(colptr_t, col[:e] + 1, 1)
    colptr_t = np.cumsum(colptr_t)
    return colptr_t.astype(np.int32), row_t.astype(np.int32), perm_t.astype(np.int32)


def csr_to_coo(rowptr: np.ndarray, col: np.ndarray) -> np.ndarray:
    """PyG edge_index [2, e]: row 0 = source, row 1 = destination."""
    return np.stack([np.asarray(col, dtype=np.int64), csr_rows(rowptr)])
