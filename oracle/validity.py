"""Sampler validity checks (north star: every sampled node is a true neighbour, fan-out caps respected, no
duplicates) — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import numpy as np


def check_block_validity(blk, colptr, row, seeds, fanouts, replace=False):
    """North-star validity: true neighbours, fan-out caps, no duplicate positions, seeds first, unique n_id."""
    n, e = blk.n, blk.e
    assert np.array_equal(blk.n_id[:len(seeds)], np.asarray(seeds, dtype=np.int32))
    assert len(np.unique(blk.n_id)) == n
    assert blk.rowptr[0] == 0 and blk.rowptr[-1] == e and np.all(np.diff(blk.rowptr) >= 0)
    assert np.array_equal(blk.n_id[blk.col], blk.col_global)
    assert np.array_equal(row[blk.e_pos], blk.col_global)
    deg_g = np.diff(colptr)
    for h, fan in enumerate(fanouts):
        lo = 0 if h == 0 else blk.node_counts[h - 1]
        hi = blk.node_counts[h]
        for i in range(lo, hi):
            v = blk.n_id[i]
            seg = blk.e_pos[blk.rowptr[i]:blk.rowptr[i + 1]]
            d = deg_g[v]
            want = (fan if d > 0 else 0) if replace else min(d, fan)
            assert len(seg) == want
            assert np.all((seg >= colptr[v]) & (seg < colptr[v + 1]))         # true in-neighbours of v
            if not replace:
                assert len(np.unique(seg)) == len(seg)                         # distinct positions
                if d <= fan:
                    assert np.array_equal(seg, np.arange(colptr[v], colptr[v + 1]))   # take-all keeps stored order
    assert np.all(np.diff(blk.rowptr)[blk.node_counts[len(fanouts) - 1]:] == 0)   # last hop not expanded
    # first-seen order: new ids appear in increasing order along the edge list
    first = {}
    for p, c in enumerate(blk.col):
        first.setdefault(int(c), p)
    new_ids = [c for c in sorted(first, key=first.get) if c >= len(seeds)]
    assert new_ids == sorted(new_ids)
