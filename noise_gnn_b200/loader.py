"""Data / NeighborLoader — drop-in for ``torch_geometric.loader.NeighborLoader`` as the reference uses it.

Reference call sites: constructed at src/pipeline.py:75-92 (train loader: ``input_nodes=train idx``,
``num_neighbors=config['nbr_neighbors']``, ``batch_size``, ``shuffle=True``, ``num_workers=1``,
``persistent_workers=True``; eval loader: ``input_nodes`` = all splits / None, batch 4092/4096),
re-created per run at :211-219, iterated at :152 and src/models/layers/sage.py:49, ``len(loader)`` at
:170; batch attributes used: ``.to(device)``, ``.x``, ``.edge_index``, ``.y``, ``.yhn``, ``.n_id``,
``.batch_size`` (:153-157, :118).

B200-first: instead of CPU worker processes that sample, slice ``x[n_id]`` and ship ~180 MB per batch
over PCIe, the graph (CSC, int32), the feature table and every ``[N, ...]`` node
attribute are uploaded ONCE and stay resident in HBM; each batch is sampled on the GPU by
``ngnn_sample_block`` (counter-based Philox, sync-free on the device) and comes back already on the
device, so ``batch.to(device)`` is a no-op.  The only per-step host->device traffic is the seed ids.
``num_workers`` / ``persistent_workers`` are accepted and ignored (there are no worker processes).
"""
from __future__ import annotations

import ctypes
import math
import weakref
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from .sharding import SeedSharder


class Data:
    """Minimal PyG ``Data``-like attribute bag: x, edge_index, y plus ad-hoc node attributes
    (the reference attaches ``data.yhn`` after loading, src/pipeline.py:72,208)."""

    def __init__(self, x=None, edge_index=None, y=None, **kwargs):
        self.x, self.edge_index, self.y = x, edge_index, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        if self.x is not None:
            return int(self.x.size(0))
        if "_num_nodes" in self.__dict__:
            return int(self.__dict__["_num_nodes"])
        return int(self.edge_index.max()) + 1

    @num_nodes.setter
    def num_nodes(self, v):
        self.__dict__["_num_nodes"] = int(v)

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.size(1))

    @property
    def num_features(self) -> int:
        return int(self.x.size(1))

    def keys(self):
        return [k for k, v in self.__dict__.items() if not k.startswith("_") and v is not None]

    def to(self, device, **kw):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, **kw))
        return self

    def __repr__(self):
        parts = [f"{k}={list(v.shape) if torch.is_tensor(v) else v}" for k, v in ((k, getattr(self, k)) for k in self.keys())]
        return f"Data({', '.join(parts)})"


class _DeviceIndex(torch.Tensor):
    """``Batch.n_id``: an int64 index tensor on the GPU that may also index a HOST tensor.

    The reference's layer-wise inference does ``x_all[batch.n_id].to(device)`` with ``x_all`` on the CPU
    (src/models/layers/sage.py:50) — with PyG the batch is still on the CPU at that point.  Our batches are born on the
    device, and torch refuses to index a CPU tensor with a CUDA index; this subclass makes exactly that one expression
    work (the ids are copied to the host for it — a synchronising D2H, the price of running the reference's loop
    unmodified; ``noise_gnn_b200.SAGE.inference`` keeps everything on the GPU instead).  Every other operation behaves
    like, and returns, a plain tensor."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        with torch._C.DisableTorchFunctionSubclass():
            if (func is torch.Tensor.__getitem__ and len(args) == 2 and isinstance(args[1], cls)
                    and isinstance(args[0], torch.Tensor) and args[0].device != args[1].device):
                return args[0][args[1].as_subclass(torch.Tensor).to(args[0].device)]
            args = tuple(a.as_subclass(torch.Tensor) if isinstance(a, cls) else a for a in args)
            return func(*args, **kwargs)


class Batch:
    """One sampled message-flow block, resident on the device.

    ``x``, ``edge_index``, ``e_id`` and the node attributes are materialised lazily on first access
    (``x`` by a row gather from the resident feature table, ``edge_index`` as PyG's int64 COO with the
    CSR block attached so SAGEConv does not re-sort it)."""

    def __init__(self, loader: "NeighborLoader", block: ops.Block, n_id32: torch.Tensor, e_pos: Optional[torch.Tensor],
                 batch_size: int, input_id: torch.Tensor):
        self._loader, self.block = loader, block
        self._n_id32, self._e_pos = n_id32, e_pos
        self.batch_size = int(batch_size)
        self.input_id = input_id
        self.num_nodes, self.num_edges = block.n_rows, block.e
        hn, he = block.hop_nodes, block.hop_edges
        self.num_sampled_nodes = [hn[0]] + [hn[i + 1] - hn[i] for i in range(len(hn) - 1)]
        self.num_sampled_edges = [he[i + 1] - he[i] for i in range(len(he) - 1)]
        self._cache = {}

    @property
    def device(self):
        return self._n_id32.device

    def to(self, device, *args, **kwargs):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("a noise_gnn_b200 Batch lives on the GPU that sampled it; .to(cpu) is not supported "
                               "(copy the individual tensors you need)")
        return self

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        cache = self.__dict__["_cache"]
        if name in cache:
            return cache[name]
        loader = self.__dict__["_loader"]
        if name == "x":
            v = ops.gather_rows(loader.x, self._n_id32, self.num_nodes)
        elif name == "edge_index":
            v = ops.csr_to_coo(self.block.rowptr, self.block.col, self.block.n_rows, self.block.e)
            v._ngnn_block = self.block
        elif name == "n_id":
            v = self._n_id32.long().as_subclass(_DeviceIndex)
        elif name == "e_id":
            if self._e_pos is None:
                raise AttributeError("e_id was not requested (NeighborLoader(..., return_e_id=True))")
            v = loader.csc_perm.index_select(0, self._e_pos.long()).long()
        elif name in loader.node_attrs:
            v = loader.node_attrs[name].index_select(0, self.n_id)
        else:
            raise AttributeError(name)
        cache[name] = v
        return v

    def __repr__(self):
        return (f"Batch(num_nodes={self.num_nodes}, num_edges={self.num_edges}, batch_size={self.batch_size}, "
                f"hops={self.num_sampled_nodes})")


class BlockSlot:
    """Device buffers of ONE sampled block at the loader's worst-case capacity: allocated once and reused, so that sampling
    is a fixed launch sequence over fixed addresses (no allocation per batch, capturable in a CUDA graph).  ``trans`` holds
    the CSC transposes of the hop prefixes 1..T the backward reads (built by the sampler itself)."""

    def __init__(self, loader: "NeighborLoader", transposes: int):
        dev, H = loader.device, len(loader._fan_cap)
        i32 = dict(dtype=torch.int32, device=dev)
        self.n_id = torch.empty(loader.max_nodes, **i32)
        self.rowptr = torch.empty(loader.max_nodes + 1, **i32)
        self.col = torch.empty(max(loader.max_edges, 1), **i32)
        self.colg = torch.empty(max(loader.max_edges, 1), **i32)
        self.epos = torch.empty(max(loader.max_edges, 1), **i32) if loader.return_e_id else None
        self.edst = torch.empty(max(loader.max_edges, 1), **i32)          # local destination id of every sampled edge
        self.counts = torch.zeros(2 * (H + 1), **i32)
        self.colt = torch.empty(max(loader.max_edges, 1), **i32) if loader.remap is not None else None
        self.nt = torch.empty(loader.max_nodes, **i32) if loader.remap is not None else None
        self.seeds = torch.zeros(loader.batch_size, dtype=torch.int64, device=dev)
        self.ctl = torch.zeros(8, **i32)                       # ngnn_step_ctl_t
        self.ctl_next = torch.zeros(8, **i32)                  # the control words of the NEXT block of this slot, staged ahead
        self.trans = []
        self.ensure_transposes(loader, transposes)
        self.owner = None                                      # weakref to the Batch handed out on these buffers
        self.pending = False                                   # sampled into, Batch not created yet
        self.host_counts = torch.empty(2 * (H + 1), dtype=torch.int32, pin_memory=True)

    def ensure_transposes(self, loader, transposes: int):
        i32 = dict(dtype=torch.int32, device=loader.device)
        while len(self.trans) < transposes:
            b = len(self.trans) + 1
            self.trans.append((torch.empty(loader.cap_nodes[b] + 1, **i32), torch.empty(max(loader.cap_edges[b], 1), **i32)))
        T = len(self.trans)
        self._colptr_t = (ctypes.c_void_p * max(T, 1))(*[t[0].data_ptr() for t in self.trans])
        self._row_t = (ctypes.c_void_p * max(T, 1))(*[t[1].data_ptr() for t in self.trans])


class NeighborLoader:
    def __init__(self, data, num_neighbors: Sequence[int], input_nodes=None, batch_size: int = 1,
                 shuffle: bool = False, replace: bool = False, num_workers: int = 0,
                 persistent_workers: bool = False, drop_last: bool = False, device=None, seed: int = 1232,
                 rank: int = 0, world_size: int = 1, return_e_id: bool = False, seeds_on_device: bool = False,
                 hot_feature_bytes: int = 0, **kwargs):
        unsupported = {k: v for k, v in kwargs.items() if k in ("disjoint", "temporal_strategy", "time_attr",
                       "weight_attr", "subgraph_type", "transform", "filter_per_worker") and v not in (None, False, "directional")}
        if unsupported:
            raise NotImplementedError(f"NeighborLoader options not used by the reference: {sorted(unsupported)}")
        if not torch.cuda.is_available():
            raise RuntimeError("noise_gnn_b200.NeighborLoader samples on the GPU; no CUDA device is available "
                               "(there is no CPU fallback)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.num_neighbors = [int(f) for f in num_neighbors]
        if any(f < 1 for f in self.num_neighbors):
            raise NotImplementedError("num_neighbors entries must be >= 1 (-1 = all neighbours is not used by the reference)")
        self.batch_size, self.shuffle, self.replace, self.drop_last = int(batch_size), bool(shuffle), bool(replace), bool(drop_last)
        self.seed, self.rank, self.world_size = int(seed), int(rank), int(world_size)
        self.return_e_id = return_e_id
        self.seeds_on_device = bool(seeds_on_device)   # keep each epoch's seed order in HBM (no per-step H2D at all)
        self.transpose_hops = 0     # hop prefixes 1..k whose CSC transpose the iterator builds on its side stream (Trainer sets it)
        self.epoch = 0
        self.data = data
        N = data.num_nodes
        self.num_nodes = N

        with torch.cuda.device(self.device):
            # --- resident graph: CSC by destination (stable => in-neighbours keep stored order) ---
            ei = data.edge_index.to(self.device)
            blk = ops.coo_to_csr(ei, N)
            self.colptr, self.row, self.csc_perm = blk.rowptr, blk.col, blk.perm
            self.num_edges = blk.e
            del ei
            # --- resident feature table and node attributes (anything shaped [N, ...]) ---
            self.x = None
            self.node_attrs = {}
            for k in data.keys():
                v = getattr(data, k)
                if not torch.is_tensor(v) or k == "edge_index" or v.dim() == 0 or v.size(0) != N:
                    continue
                if k == "x":
                    # Row stride rounded up to 64 B (16 floats): HBM is fetched in 64-byte units, so a 400-byte row at
                    # stride 400 straddles 7-8 of them (measured +20 % DRAM read traffic on the layer-1 gather) while
                    # at stride 448 it is exactly 7.  The tensor keeps its [N, F] shape; only stride(0) changes.
                    F_ = int(v.size(1)) if v.dim() == 2 else 1
                    ldp = (F_ + 15) // 16 * 16
                    buf = torch.zeros((N, ldp), dtype=torch.float32, device=self.device)   # pad columns stay zero
                    self.x = buf[:, :F_] if v.dim() == 2 else buf
                    self.x.copy_(v.reshape(N, -1).to(self.device, dtype=torch.float32))
                else:
                    self.node_attrs[k] = v.to(self.device)
            # --- hot-rows-first copy of the feature table for the fused step's layer-1 gather ---
            # On a power-law graph a few percent of the nodes receive a large share of the sampled edges.  The table
            # is stored a second time sorted by in-degree (descending) so that the hot rows are the first
            # `hot_rows` table rows: K-AGG gathers those with L2 evict_last priority and everything else with
            # evict_first, so the hot set stays L2-resident from block to block instead of being re-fetched from HBM.
            # remap[global id] = table row; blocks carry (col_table, n_table) = remap of (col_global, n_id).
            # OFF by default (hot_feature_bytes = 0): measured on the products workload it takes 3 us off the layer-1
            # aggregation (55.5 -> 52.4 us, roofline 0.55 -> 0.58) but adds one launch and two allocations per batch
            # to a loop that is host-bound, a net loss for the step; every row then keeps the uniform evict_last hint.
            self.x_hot, self.remap, self.hot_rows = None, None, 0
            if self.x is not None and self.x.dim() == 2 and hot_feature_bytes > 0 and N > 1:
                deg = (self.colptr[1:] - self.colptr[:-1]).to(torch.int64)
                order = torch.argsort(deg, descending=True, stable=True)
                self.remap = torch.empty(N, dtype=torch.int32, device=self.device)
                self.remap[order] = torch.arange(N, dtype=torch.int32, device=self.device)
                ldp = self.x.stride(0)
                buf = torch.zeros((N, ldp), dtype=torch.float32, device=self.device)   # pad columns stay zero
                self.x_hot = buf[:, :self.x.size(1)]
                self.x_hot.copy_(self.x.index_select(0, order))
                self.hot_rows = int(min(N, hot_feature_bytes // (ldp * 4)))
                del deg, order
            # --- seeds ---
            if input_nodes is None:
                nodes = torch.arange(N, dtype=torch.int64)
            else:
                nodes = torch.as_tensor(input_nodes).cpu()
                if nodes.dtype == torch.bool:
                    nodes = nodes.nonzero().view(-1)
                nodes = nodes.to(torch.int64).view(-1)
            if nodes.numel() and (int(nodes.min()) < 0 or int(nodes.max()) >= N):
                raise ValueError(f"input_nodes must lie in [0, {N}): got ids in [{int(nodes.min())}, {int(nodes.max())}]")
            self.input_nodes = nodes
            self.sharder = SeedSharder(nodes, self.batch_size, self.shuffle, self.seed, self.rank, self.world_size,
                                       self.drop_last)
            # --- sampler capacities and workspace ---
            L = _lib.load()
            self._fan = (ctypes.c_int32 * len(self.num_neighbors))(*self.num_neighbors)
            self._fan_cap = (ctypes.c_int32 * len(self.num_neighbors))(*self.num_neighbors)
            mn, me = ctypes.c_int64(), ctypes.c_int64()
            _lib.call("ngnn_sample_capacity", self.batch_size, self._fan, len(self.num_neighbors), N,
                      ctypes.byref(mn), ctypes.byref(me))
            self.max_nodes, self.max_edges = mn.value, me.value
            # cumulative worst-case nodes / edges after each hop (capacities of the transposed prefixes, arena sizing)
            self.cap_nodes, self.cap_edges = [self.batch_size], [0]
            for h in range(1, len(self.num_neighbors) + 1):
                _lib.call("ngnn_sample_capacity", self.batch_size, self._fan, h, N, ctypes.byref(mn), ctypes.byref(me))
                self.cap_nodes.append(mn.value)
                self.cap_edges.append(me.value)
            self._slots = []
            nbytes = L.ngnn_sample_workspace_bytes(N, self.batch_size, self._fan, len(self.num_neighbors))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.call("ngnn_sample_workspace_init", ops._ptr(self._ws), self._ws.numel(), N, ops._stream())

    def set_num_neighbors(self, num_neighbors: Sequence[int]) -> None:
        """Change the fan-outs between epochs.  Only prefixes / smaller fan-outs than the loader was built for are
        accepted: the sampler workspace and block capacities were sized at construction."""
        fan = [int(f) for f in num_neighbors]
        cap = [int(self._fan_cap[i]) for i in range(len(self._fan_cap))]
        if not fan or len(fan) > len(cap) or any(f < 1 or f > c for f, c in zip(fan, cap)):
            raise ValueError(f"num_neighbors {fan} exceeds the capacity this loader was built with ({cap})")
        self.num_neighbors = fan
        self._fan = (ctypes.c_int32 * len(fan))(*fan)

    # ------------------------------------------------------------------ batching
    @property
    def num_batches_global(self) -> int:
        return self.sharder.num_batches_global

    def __len__(self) -> int:
        return len(self.sharder)

    def epoch_permutation(self, epoch: int) -> torch.Tensor:
        """Rank-agnostic seed order for an epoch: a pure function of (seed, epoch)."""
        return self.sharder.epoch_permutation(epoch)

    def batch_seeds(self, order: torch.Tensor, global_batch_idx: int) -> torch.Tensor:
        return self.sharder.batch_seeds(order, global_batch_idx)

    def label_array(self, name: str) -> torch.Tensor:
        """A node attribute as a contiguous int64 [N] device array indexed by global id (for the fused step)."""
        cache = self.__dict__.setdefault("_label_cache", {})
        src = self.node_attrs[name]
        hit = cache.get(name)
        if hit is None or hit[0] is not src:
            cache[name] = (src, src.view(-1).to(torch.int64).contiguous())
        return cache[name][1]

    # ------------------------------------------------------------------ sampling
    def _acquire_slot(self) -> "BlockSlot":
        """A BlockSlot no live Batch refers to (grows the pool when every slot is still held by the consumer)."""
        for slot in self._slots:
            if not slot.pending and (slot.owner is None or slot.owner() is None):
                slot.ensure_transposes(self, self.transpose_hops)
                slot.pending = True
                return slot
        slot = BlockSlot(self, self.transpose_hops)
        slot.pending = True
        self._slots.append(slot)
        return slot

    def fixed_slots(self, count: int, transposes: int):
        """`count` dedicated BlockSlots for a captured step (noise_gnn_b200.train.Trainer): their addresses never change."""
        cache = self.__dict__.setdefault("_fixed_slots", [])
        while len(cache) < count:
            cache.append(BlockSlot(self, transposes))
        for slot in cache:
            slot.ensure_transposes(self, transposes)
        return cache[:count]

    def launch_sample(self, slot: "BlockSlot", seeds: Optional[torch.Tensor], bs: int, epoch: int, batch_idx: int,
                      use_ctl: bool = False, transposes: Optional[int] = None):
        """Enqueue one ngnn_sample_block_ex into `slot` on the CURRENT stream (no host sync, no allocation).  `seeds`: int64
        host (pinned) or device tensor copied into the slot's seed buffer, or None when the caller already filled it.
        use_ctl: the RNG key is read from the slot's device-side control words (set with ngnn_step_ctl_set) instead of
        (epoch, batch_idx) — what a captured launch sequence needs."""
        H = len(self.num_neighbors)
        if bs <= 0:
            raise ValueError("cannot sample an empty seed batch")
        if bs > self.batch_size:
            raise ValueError(f"{bs} seeds exceed the loader's batch_size {self.batch_size}")
        if seeds is not None:
            slot.seeds[:bs].copy_(seeds.view(-1)[:bs], non_blocking=True)
        T = len(slot.trans) if transposes is None else min(int(transposes), len(slot.trans))
        T = min(T, H)
        with ops._timed("sample"):
            _lib.call("ngnn_sample_block_ex", ops._ptr(self.colptr), ops._ptr(self.row), self.num_nodes, ops._ptr(slot.seeds), bs,
                      self._fan, H, int(self.replace), self.seed & (2**64 - 1), epoch & 0xFFFFFFFF, batch_idx & 0xFFFFFFFF,
                      ops._ptr(slot.ctl) if use_ctl else None, ops._ptr(slot.n_id), ops._ptr(slot.rowptr), ops._ptr(slot.col),
                      ops._ptr(slot.colg), ops._ptr(slot.epos), ops._ptr(slot.edst), ops._ptr(slot.counts), T, slot._colptr_t,
                      slot._row_t,
                      ops._ptr(self._ws), self._ws.numel(), ops._stream())
        if self.remap is not None:                        # table rows of the block (same stream, no host round trip)
            _lib.call("ngnn_block_table_index", ops._ptr(self.remap), ops._ptr(slot.colg), ops._ptr(slot.n_id), ops._ptr(slot.counts),
                      H, self.max_nodes, self.max_edges, ops._ptr(slot.colt), ops._ptr(slot.nt), ops._stream())
        return T

    def _launch_sample(self, seeds: torch.Tensor, epoch: int, batch_idx: int):
        """Sample into a free slot on the CURRENT stream and start the (asynchronous) read of its extents."""
        if not seeds.is_cuda and not seeds.is_pinned():
            seeds = seeds.pin_memory()
        seeds = seeds.to(torch.int64)
        bs = seeds.numel()
        slot = self._acquire_slot()
        T = self.launch_sample(slot, seeds, bs, epoch, batch_idx)
        slot.host_counts.copy_(slot.counts, non_blocking=True)      # the one device->host read per batch: block extents
        return dict(slot=slot, bs=bs, T=T)

    def _finish_sample(self, pend) -> Batch:
        H = len(self.num_neighbors)
        slot = pend["slot"]
        c = slot.host_counts.tolist()
        hop_nodes, hop_edges = c[:H + 1], c[H + 1:2 * (H + 1)]
        n, e = hop_nodes[-1], hop_edges[-1]
        block = ops.Block(slot.rowptr[:n + 1], slot.col[:e], n, e, hop_nodes=hop_nodes, hop_edges=hop_edges,
                          col_global=slot.colg[:e], n_id=slot.n_id[:n])
        block.counts = slot.counts
        for b in range(1, pend["T"] + 1):                 # transposes built by the sampler
            block._t[(hop_edges[b], hop_nodes[b])] = (slot.trans[b - 1][0][:hop_nodes[b] + 1], slot.trans[b - 1][1][:hop_edges[b]])
        if slot.colt is not None:
            block.col_table, block.n_table = slot.colt[:e], slot.nt[:n]
        batch = Batch(self, block, slot.n_id[:n], None if slot.epos is None else slot.epos[:e], pend["bs"], slot.seeds[:pend["bs"]])
        batch._slot = slot
        slot.owner = weakref.ref(batch)
        slot.pending = False
        return batch

    def sample(self, seeds: torch.Tensor, epoch: int = 0, batch_idx: int = 0) -> Batch:
        """Sample one block for explicit seeds (host or device int64) on the current stream."""
        with torch.cuda.device(self.device):
            side = self.__dict__.get("_side")
            if side is not None:                          # the sampler workspace is shared with the iterator's side stream
                torch.cuda.current_stream().wait_stream(side)
            pend = self._launch_sample(torch.as_tensor(seeds), epoch, batch_idx)
            torch.cuda.current_stream().synchronize()
            return self._finish_sample(pend)

    def __iter__(self):
        """Yields device-resident batches.  Sampling (block + the backward's CSC transposes, one fixed launch sequence) runs
        up to two blocks ahead on a side stream into pooled buffers; the host blocks only on the extents of the batch it is
        about to hand out (sampled long before).  A batch's buffers return to the pool when the consumer drops it."""
        epoch = self.epoch
        self.epoch += 1
        order = self.epoch_permutation(epoch)
        steps = len(self)
        with torch.cuda.device(self.device):
            if self.__dict__.get("_side") is None:
                self._side = torch.cuda.Stream(device=self.device)
            side = self._side
            side.wait_stream(torch.cuda.current_stream())          # the resident graph was built on the caller's stream
            order = order.to(self.device) if self.seeds_on_device else order.pin_memory()
            pending = {}
            state = {"next": 0}

            def advance(upto):
                """Launch sampling for every step < upto that has not been launched yet (idempotent)."""
                while state["next"] < min(upto, steps):
                    k = state["next"]
                    g = self.sharder.global_batch_index(k)
                    # buffers freed by the consumer may still be read by work it enqueued: order the reuse after it
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        pend = self._launch_sample(self.batch_seeds(order, g), epoch, g)
                        ev = torch.cuda.Event()
                        ev.record(side)
                    pending[k] = (pend, ev)
                    state["next"] = k + 1

            advance(2)
            for i in range(steps):
                advance(i + 1)
                pend, ev = pending.pop(i)
                ev.synchronize()                                   # host needs the block extents of batch i
                batch = self._finish_sample(pend)
                torch.cuda.current_stream().wait_event(ev)
                del pend
                # Trainer.train_step calls this right after it has enqueued the step, so the next sampling is queued
                # while the GPU is busy; a plain consumer gets the same effect when the generator resumes
                batch._prefetch = lambda upto=i + 3: advance(upto)
                yield batch
                del batch
                advance(i + 3)
