"""Train step of the hot loop — the body of ``PipelineCO.train`` (reference src/pipeline.py:152-169) on the
B200-native path:

    batch -> ngnn_sage_step (ONE C-ABI call: trimmed fused SAGE forward, softmax-CE + accuracy count on the seed
             rows, backward straight into the flat gradient bucket)
          -> [DP: one NCCL all-reduce of the flat gradient bucket] -> ngnn_adam_step (fused Adam).

Parameters, gradients and Adam state live in flat fp32 buckets (the model's ``Parameter``s are views into them,
so ``state_dict`` keeps PyG's names); loss / accuracy accumulate in a device tensor and are read by the caller
when it wants them (the reference reads them every step, src/pipeline.py:164-165).  Nothing in the step
synchronises with the host.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib, dp, ops


class FlatBuckets:
    """Re-homes a module's parameters (and their .grad) into contiguous flat fp32 buffers, in
    ``module.parameters()`` order — for SAGE: per layer lin_l.weight, lin_l.bias, lin_r.weight, which is
    the layout ngnn_sage_step expects."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev = params[0].device
        total = sum(p.numel() for p in params)
        self.param = torch.empty(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            self.param[off:off + n].copy_(p.data.view(-1))
            p.data = self.param[off:off + n].view_as(p)
            p.grad = self.grad[off:off + n].view_as(p)
            off += n
        self.params = params
        self.numel = total

    def rebind_grads(self):
        """autograd may replace p.grad when it was None; keep them pointing into the bucket."""
        off = 0
        for p in self.params:
            n = p.numel()
            view = self.grad[off:off + n].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view
            off += n


def hop_capacities(batch_size: int, fanouts, num_nodes: int):
    """Cumulative worst-case nodes / edges after each hop (the extents ngnn_sample_block can produce)."""
    fr, nodes, edges = batch_size, [batch_size], [0]
    for f in fanouts:
        e = fr * f
        edges.append(edges[-1] + e)
        fr = min(e, num_nodes)
        nodes.append(min(nodes[-1] + fr, num_nodes + batch_size))
    return nodes, edges


class Trainer:
    """Owns the flat buckets, the step arena and the fused optimizer for one SAGE network."""

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 process_group=None, world_size: int = 1, rank: int = 0, use_graph: bool = True):
        self.model = model
        self.rank, self.use_graph = int(rank), bool(use_graph)
        self._gs = None                      # captured-step state (see run_steps)
        self.graph_replays = 0               # steps issued as a graph replay, and the kernel launches those replays contained
        self.replayed_launches = 0
        self.buckets = FlatBuckets(model)
        dev = self.buckets.param.device
        self.exp_avg = torch.zeros_like(self.buckets.param)
        self.exp_avg_sq = torch.zeros_like(self.buckets.param)
        self.step_dev = torch.zeros(2, dtype=torch.int64, device=dev)    # [completed steps, ticket word of the Adam kernel]
        self.stats = torch.zeros(2, dtype=torch.float32, device=dev)     # [sum of step losses, #correct]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.process_group, self.world_size = process_group, world_size
        self.steps = 0
        self._arena = None
        self._arena_key = None
        convs = model.convs
        self._cfg = dict(num_layers=len(convs), in_dim=convs[0].in_channels,
                         hidden_dim=convs[0].out_channels if len(convs) > 1 else convs[0].out_channels,
                         out_dim=convs[-1].out_channels)
        n_expected = sum(2 * c.in_channels * c.out_channels + c.out_channels for c in convs)
        if n_expected != self.buckets.numel:
            raise ValueError("the fused step expects a plain SAGE stack (lin_l.weight, lin_l.bias, lin_r.weight per layer)")

    def reset_stats(self):
        self.stats.zero_()

    def read_stats(self):
        """(sum of per-step mean losses, number of correct seed predictions) — one device->host read."""
        s = self.stats.tolist()
        return s[0], int(s[1])

    def read_stats_async(self):
        """Starts the device->host copy of (loss sum, correct count) as of the work enqueued so far and returns a
        handle for `resolve_stats`; lets a training loop log every step without draining the GPU each time."""
        ring = self.__dict__.setdefault("_stats_ring", [torch.empty(2, dtype=torch.float32, pin_memory=True) for _ in range(8)])
        i = self.__dict__.get("_stats_i", 0)
        self._stats_i = i + 1
        slot = ring[i % len(ring)]
        slot.copy_(self.stats, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return slot, ev

    @staticmethod
    def resolve_stats(handle):
        slot, ev = handle
        ev.synchronize()
        s = slot.tolist()
        return s[0], int(s[1])

    # ------------------------------------------------------------------ fused step
    def _model_struct(self, training: bool):
        m = self.model
        return _lib.SageModel(self._cfg["num_layers"], self._cfg["in_dim"], self._cfg["hidden_dim"], self._cfg["out_dim"],
                              float(m.dropout), int(training))

    @staticmethod
    def _table(loader, blk, bd):
        """Feature table the layer-1 gather reads: the loader's hot-rows-first copy when the block carries its table
        rows (col_table / n_table), else the table in node-id order."""
        if loader.x_hot is not None and blk.col_table is not None and blk.n_table is not None:
            bd.col_table, bd.n_table, bd.hot_rows = blk.col_table.data_ptr(), blk.n_table.data_ptr(), loader.hot_rows
            return loader.x_hot
        return loader.x

    def _ensure_arena(self, loader, ms):
        key = (loader.batch_size, tuple(loader.num_neighbors), loader.num_nodes)
        if self._arena_key != key:
            H = len(loader.num_neighbors)
            nodes, edges = list(loader.cap_nodes[:H + 1]), list(loader.cap_edges[:H + 1])   # the sampler's own worst case
            self._max_nodes = (ctypes.c_int64 * (H + 1))(*nodes)
            self._max_edges = (ctypes.c_int64 * (H + 1))(*edges)
            nbytes = _lib.load().ngnn_sage_step_workspace_bytes(ctypes.byref(ms), H, self._max_nodes, self._max_edges)
            if nbytes == 0:
                raise RuntimeError("ngnn_sage_step_workspace_bytes rejected the model / loader configuration")
            self._arena = torch.empty(nbytes, dtype=torch.uint8, device=self.buckets.param.device)
            self._arena_key = key
        return self._arena

    def forward_backward(self, batch, target_attr: Optional[str] = "yhn", label_attr: Optional[str] = "y",
                         train: bool = True, want_logits: bool = False):
        """One ngnn_sage_step call on a Batch of our NeighborLoader; gradients land in the flat bucket."""
        loader, blk = batch._loader, batch.block
        ms = self._model_struct(training=train and self.model.training)
        arena = self._ensure_arena(loader, ms)
        H = len(blk.hop_nodes) - 1
        hop_n = (ctypes.c_int32 * (H + 1))(*blk.hop_nodes)
        hop_e = (ctypes.c_int32 * (H + 1))(*blk.hop_edges)
        bd = _lib.BlockDesc(blk.rowptr.data_ptr(), blk.col.data_ptr(), blk.col_global.data_ptr(), blk.n_id.data_ptr(), H,
                            hop_n, hop_e)
        if train:
            need = min(self._cfg["num_layers"] - 1, H)
            loader.transpose_hops = max(loader.transpose_hops, need)      # later batches arrive with them prebuilt
            for b in range(1, min(need, 7) + 1):
                pre = blk._t.get((blk.hop_edges[b], blk.hop_nodes[b]))
                if pre is not None:
                    bd.colptr_t[b], bd.row_t[b] = pre[0].data_ptr(), pre[1].data_ptr()
        tgt = loader.label_array(target_attr) if target_attr else None
        lab = loader.label_array(label_attr) if label_attr else None
        table = self._table(loader, blk, bd)
        logits = None
        if want_logits:
            logits = torch.empty((batch.batch_size, self._cfg["out_dim"]), dtype=torch.float32, device=arena.device)
        self.steps += 1
        with ops._timed("sage_step"):
            _lib.call("ngnn_sage_step", ctypes.byref(ms), ops._ptr(self.buckets.param),
                      ops._ptr(self.buckets.grad) if train else None, ctypes.byref(bd), self._max_nodes, self._max_edges,
                      ops._ptr(table), table.stride(0), ops._ptr(tgt), ops._ptr(lab), self._drop_seed(),
                      self.steps * self._cfg["num_layers"], ops._ptr(self.stats), ops._ptr(logits),
                      logits.stride(0) if logits is not None else 0, ops._ptr(arena), arena.numel(), ops._stream())
        return logits

    def _drop_seed(self) -> int:
        """Key of the fused dropout stream: the model's seed, decorrelated across data-parallel ranks."""
        return (int(self.model.drop_seed) + 0x9E3779B97F4A7C15 * self.rank) & (2**64 - 1)

    # ------------------------------------------------------------------ the epoch loop on a captured step
    def _graph_state(self, loader, target_attr, label_attr):
        """Everything a replayed step addresses — two block slots, the arena, the label arrays, the side stream — keyed on
        what the captured launch sequence bakes in."""
        tgt = loader.label_array(target_attr)
        lab = loader.label_array(label_attr) if label_attr else None
        ms = self._model_struct(training=self.model.training)
        arena = self._ensure_arena(loader, ms)
        table = loader.x_hot if loader.x_hot is not None else loader.x
        key = (id(loader), loader.batch_size, tuple(loader.num_neighbors), tgt.data_ptr(), lab.data_ptr() if lab is not None else 0,
               bool(self.model.training), float(self.model.dropout), arena.data_ptr(), table.data_ptr(), self.world_size)
        gs = self._gs
        if gs is None or gs["key"] != key:
            H = len(loader.num_neighbors)
            T = min(self._cfg["num_layers"] - 1, H, 4)
            gs = dict(key=key, loader=loader, tgt=tgt, lab=lab, ms=ms, arena=arena, table=table, H=H, T=T,
                      capture_stream=torch.cuda.Stream(device=arena.device),
                      slots=loader.fixed_slots(2, T), side=torch.cuda.Stream(device=arena.device), graphs=[None, None],
                      graph_launches=[0, 0], stage=torch.cuda.Stream(device=arena.device), staged=[None, None], freed=[None, None])
            self._gs = gs
        return gs

    def release_graphs(self):
        """Drop the captured steps (they are re-captured on demand).  Call before ``dist.destroy_process_group()``: a CUDA graph
        that holds a captured NCCL all-reduce keeps the communicator busy, and tearing the group down under it hangs."""
        if self._gs is not None:
            self._gs["graphs"] = [None, None]
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def _slot_desc(self, gs, k: int, bs: int):
        """ngnn_block_t over slot k with DEVICE-side extents (counts), the sampler-built transposes and the control words."""
        slot, loader = gs["slots"][k], gs["loader"]
        bd = _lib.BlockDesc(slot.rowptr.data_ptr(), slot.col.data_ptr(), slot.colg.data_ptr(), slot.n_id.data_ptr(), gs["H"], None, None)
        for b in range(1, gs["T"] + 1):
            bd.colptr_t[b], bd.row_t[b] = slot.trans[b - 1][0].data_ptr(), slot.trans[b - 1][1].data_ptr()
        if loader.x_hot is not None:
            bd.col_table, bd.n_table, bd.hot_rows = slot.colt.data_ptr(), slot.nt.data_ptr(), loader.hot_rows
        bd.counts, bd.batch_size, bd.ctl = slot.counts.data_ptr(), int(bs), slot.ctl.data_ptr()
        bd.agg1_buffer = k + 1                       # layer 1's aggregation of the block in slot k lives in arena copy k
        return bd

    def _enqueue_agg1(self, gs, k: int, bs: int):
        """Layer 1's aggregation of the block in slot k into arena copy k (on the current stream)."""
        bd = self._slot_desc(gs, k, bs)
        ms, arena, table = gs["ms"], gs["arena"], gs["table"]
        _lib.call("ngnn_sage_agg1", ctypes.byref(ms), ctypes.byref(bd), self._max_nodes, self._max_edges, ops._ptr(table),
                  table.stride(0), k, ops._ptr(arena), arena.numel(), ops._stream())

    def _enqueue_pair(self, gs, k: int, bs_cur: int, bs_next: int):
        """layer-1 aggregation of slot k, then [rest of the step on slot k]  ||  [sample the next block -> slot 1-k] on the side
        stream, then all-reduce + Adam.  The aggregation is the one kernel that needs the whole HBM bandwidth, so it runs with
        the GPU to itself (measured: 49 us alone, 70 us beside the sampler, 83 us beside the GEMMs, whose shared-memory
        traffic shares the L1 data path with its gathers); the sampler — latency-bound, light — runs under the GEMMs.
        The same call sequence runs eagerly (first steps, ragged tail) and under graph capture."""
        loader, side = gs["loader"], gs["side"]
        cur = torch.cuda.current_stream()
        ms, arena, table = gs["ms"], gs["arena"], gs["table"]
        # the weight pack (depends on the parameters only) beside the aggregation (depends on the block only)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            _lib.call("ngnn_sage_prep_weights", ctypes.byref(ms), ops._ptr(self.buckets.param), gs["H"], self._max_nodes,
                      self._max_edges, ops._ptr(arena), arena.numel(), ops._stream())
        self._enqueue_agg1(gs, k, bs_cur)      # HBM-bound, alone on the GPU: the sampler is forked only behind it
        cur.wait_stream(side)
        if bs_next > 0:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                nxt = gs["slots"][1 - k]
                nxt.ctl.copy_(nxt.ctl_next, non_blocking=True)      # the staged control words become the slot's live ones
                loader.launch_sample(nxt, None, bs_next, 0, 0, use_ctl=True, transposes=gs["T"])
        bd = self._slot_desc(gs, k, bs_cur)
        bd.weights_prepared = 1
        _lib.call("ngnn_sage_step", ctypes.byref(ms), ops._ptr(self.buckets.param), ops._ptr(self.buckets.grad), ctypes.byref(bd),
                  self._max_nodes, self._max_edges, ops._ptr(table), table.stride(0), ops._ptr(gs["tgt"]), ops._ptr(gs["lab"]),
                  self._drop_seed(), 0, ops._ptr(self.stats), None, 0, ops._ptr(arena), arena.numel(), ops._stream())
        self.optimizer_step()
        if bs_next > 0:
            cur.wait_stream(side)

    def run_steps(self, loader, n_steps: int, start_epoch: int = 0, start_step: int = 0, target_attr: str = "yhn",
                  label_attr: Optional[str] = "y", seeds_resident: bool = False, log_every_step: bool = True, on_step=None):
        """`n_steps` consecutive train steps of this rank's schedule, starting at local step `start_step` of epoch
        `start_epoch` and continuing across epoch boundaries (each epoch has its own rank-agnostic seed permutation).

        Every step is the SAME launch sequence over fixed buffers — sample the next block (side stream) while the current
        one runs forward / loss / backward / [all-reduce] / Adam — with all block extents left on the device, so it is
        captured once per buffer parity in a CUDA graph and replayed: per step the host issues, on a staging stream and one
        step ahead, the next block's seed ids (H2D from pinned memory, or a device copy when ``seeds_resident``) and one
        control-word kernel, then one graph launch and, with ``log_every_step``, an asynchronous read-back of the loss /
        accuracy accumulators (the reference reads them every step, src/pipeline.py:164-165).  The first two steps and
        ragged batches run the same calls eagerly.
        ``on_step(j)`` is called after step j has been enqueued.

        Returns (sum of step losses, correct seed predictions, per-step cumulative (loss, correct) snapshots or None)."""
        dev = self.buckets.param.device
        n_steps = int(n_steps)
        if n_steps <= 0:
            return 0.0, 0, None
        with torch.cuda.device(dev):
            gs = self._graph_state(loader, target_attr, label_attr)
            sh = loader.sharder
            spe = len(loader)                                   # steps per epoch on this rank
            L = self._cfg["num_layers"]
            bs_full = sh.full_len
            steps0 = self.steps                                 # the step that consumes schedule entry j is number steps0 + j + 1
            cur = torch.cuda.current_stream()
            side = loader.__dict__.get("_side")
            if side is not None:                                # the sampler workspace may still be in use by the loader's own iterator
                cur.wait_stream(side)
            orders = {}

            def order_of(epoch):
                if epoch not in orders:
                    o = loader.epoch_permutation(epoch)
                    orders[epoch] = o.to(dev) if seeds_resident else o.pin_memory()
                return orders[epoch]

            stage = gs["stage"]
            stage.wait_stream(cur)                              # whatever used the slots before this call has been enqueued on `cur`
            gs["staged"], gs["freed"] = [None, None], [None, None]

            def stage_block(j, k):
                """Seed ids + control words of schedule entry j into slot k's staging fields, on the staging stream: they are
                only read by the sampler of the pair that runs one step later, so the copies run under the current step
                instead of between two graph launches.  Returns the batch length."""
                epoch, i = start_epoch + (start_step + j) // spe, (start_step + j) % spe
                g = sh.global_batch_index(i)
                bs, w = sh.batch_len(g), sh.loss_scale(i)
                slot = gs["slots"][k]
                if gs["freed"][k] is not None:
                    stage.wait_event(gs["freed"][k])            # the pair that last read this slot's staging fields is done
                with torch.cuda.stream(stage):
                    slot.seeds[:bs].copy_(loader.batch_seeds(order_of(epoch), g), non_blocking=True)
                    _lib.call("ngnn_step_ctl_set", ops._ptr(slot.ctl_next), epoch & 0xFFFFFFFF, g & 0xFFFFFFFF, (steps0 + j + 1) * L,
                              float(w), ops._stream())
                    ev = torch.cuda.Event()
                    ev.record()
                gs["staged"][k] = ev
                return bs

            def mark_freed(k):
                ev = torch.cuda.Event()
                ev.record()
                gs["freed"][k] = ev

            lib = _lib.load()
            log = torch.empty((n_steps, 2), dtype=torch.float32, pin_memory=True) if log_every_step else None
            self.stats.zero_()
            # prologue: the first block is sampled eagerly into slot 0
            bs_cur = stage_block(0, 0)
            cur.wait_event(gs["staged"][0])
            gs["slots"][0].ctl.copy_(gs["slots"][0].ctl_next, non_blocking=True)
            loader.launch_sample(gs["slots"][0], None, bs_cur, 0, 0, use_ctl=True, transposes=gs["T"])
            mark_freed(0)
            for j in range(n_steps):
                k = j & 1
                bs_next = stage_block(j + 1, 1 - k) if j + 1 < n_steps else 0
                if bs_next > 0:
                    cur.wait_event(gs["staged"][1 - k])
                self.steps += 1
                # replay only when this round and the next are full on EVERY rank (all ranks then agree on graph vs eager)
                i_cur, i_next = (start_step + j) % spe, (start_step + j + 1) % spe
                full = bs_next > 0 and sh.round_is_full(i_cur) and sh.round_is_full(i_next)
                if self.use_graph and full and j >= 2:
                    if gs["graphs"][k] is None:
                        c0 = lib.ngnn_launch_count()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=gs["capture_stream"]):
                            self._enqueue_pair(gs, k, bs_full, bs_full)
                        gs["graphs"][k] = g
                        gs["graph_launches"][k] = int(lib.ngnn_launch_count() - c0)
                    gs["graphs"][k].replay()
                    self.graph_replays += 1
                    self.replayed_launches += gs["graph_launches"][k]
                else:
                    self._enqueue_pair(gs, k, bs_cur, bs_next)
                if bs_next > 0:
                    mark_freed(1 - k)                           # this pair's sampler was the reader of slot 1-k's staging fields
                if log is not None:
                    log[j].copy_(self.stats, non_blocking=True)
                if on_step is not None:
                    on_step(j)
                bs_cur = bs_next
            total = self.stats.tolist()                         # one synchronising read at the end
            return total[0], int(total[1]), log

    def train_epoch(self, loader, target_attr: str = "yhn", label_attr: Optional[str] = "y", epoch: Optional[int] = None,
                    max_steps: Optional[int] = None, seeds_resident: bool = False, log_every_step: bool = True):
        """One pass of ``PipelineCO.train`` (reference src/pipeline.py:144-173) over this rank's share of the epoch, on the
        captured step (see run_steps).  Returns (train_loss = mean step loss as the reference logs it, correct seed
        predictions, per-step cumulative (loss, correct) snapshots or None)."""
        if epoch is None:
            epoch = loader.epoch
            loader.epoch += 1
        steps = len(loader) if max_steps is None else min(int(max_steps), len(loader))
        if steps <= 0:
            return 0.0, 0, None
        loss_sum, correct, log = self.run_steps(loader, steps, start_epoch=epoch, target_attr=target_attr, label_attr=label_attr,
                                                seeds_resident=seeds_resident, log_every_step=log_every_step)
        return loss_sum / steps, correct, log

    def optimizer_step(self):
        bk = self.buckets
        # DP: one all-reduce of the flat bucket; the 1/world averaging is folded into the Adam kernel
        scale = dp.allreduce_mean_(bk.grad, self.process_group, self.world_size)
        ops.adam_step(bk.param, bk.grad, self.exp_avg, self.exp_avg_sq, self.step_dev, lr=self.lr, betas=self.betas,
                      eps=self.eps, weight_decay=self.weight_decay, grad_scale=scale)

    def train_step(self, batch, target_attr: str = "yhn", label_attr: Optional[str] = "y", want_logits: bool = False):
        logits = self.forward_backward(batch, target_attr, label_attr, train=True, want_logits=want_logits)
        self.optimizer_step()
        prefetch = getattr(batch, "_prefetch", None)
        if prefetch is not None:          # queue the next block's sampling now, while the GPU is busy with this step
            prefetch()
        return logits

    # ------------------------------------------------------------------ split step (losses that couple several networks)
    def _block_desc(self, batch):
        blk = batch.block
        H = len(blk.hop_nodes) - 1
        hop_n = (ctypes.c_int32 * (H + 1))(*blk.hop_nodes)
        hop_e = (ctypes.c_int32 * (H + 1))(*blk.hop_edges)
        bd = _lib.BlockDesc(blk.rowptr.data_ptr(), blk.col.data_ptr(), blk.col_global.data_ptr(), blk.n_id.data_ptr(), H,
                            hop_n, hop_e)
        need = min(self._cfg["num_layers"] - 1, H)
        batch._loader.transpose_hops = max(batch._loader.transpose_hops, need)
        for b in range(1, min(need, 7) + 1):
            pre = blk._t.get((blk.hop_edges[b], blk.hop_nodes[b]))
            if pre is not None:
                bd.colptr_t[b], bd.row_t[b] = pre[0].data_ptr(), pre[1].data_ptr()
        return bd, (hop_n, hop_e)

    def forward_only(self, batch):
        """Training-mode forward (ngnn_sage_forward): seed-row logits; activations stay in the arena for backward_from."""
        loader = batch._loader
        ms = self._model_struct(training=self.model.training)
        arena = self._ensure_arena(loader, ms)
        bd, keep = self._block_desc(batch)
        logits = torch.empty((batch.batch_size, self._cfg["out_dim"]), dtype=torch.float32, device=arena.device)
        table = self._table(loader, batch.block, bd)
        self.steps += 1
        _lib.call("ngnn_sage_forward", ctypes.byref(ms), ops._ptr(self.buckets.param), ctypes.byref(bd), self._max_nodes,
                  self._max_edges, ops._ptr(table), table.stride(0), self._drop_seed(),
                  self.steps * self._cfg["num_layers"], ops._ptr(logits), logits.stride(0), ops._ptr(arena), arena.numel(),
                  ops._stream())
        return logits

    def backward_from(self, batch, dlogits):
        """Backward of the preceding forward_only on the same batch from d(loss)/d(logits) (ngnn_sage_backward)."""
        loader = batch._loader
        ms = self._model_struct(training=self.model.training)
        arena = self._ensure_arena(loader, ms)
        bd, keep = self._block_desc(batch)
        dlogits = ops._rows(dlogits, "dlogits")
        table = self._table(loader, batch.block, bd)
        _lib.call("ngnn_sage_backward", ctypes.byref(ms), ops._ptr(self.buckets.param), ops._ptr(self.buckets.grad),
                  ctypes.byref(bd), self._max_nodes, self._max_edges, ops._ptr(table), table.stride(0),
                  ops._ptr(dlogits), ops._ld(dlogits), ops._ptr(arena), arena.numel(), ops._stream())

    # ------------------------------------------------------------------ autograd variant (same kernels, ~80 FFI calls)
    def train_step_autograd(self, batch, target_attr: str = "yhn", label_attr: Optional[str] = "y"):
        model, bk = self.model, self.buckets
        bs = batch.batch_size
        logits = model.forward_batch(batch)
        target = getattr(batch, target_attr)[:bs].view(-1)
        y_true = getattr(batch, label_attr)[:bs].view(-1) if label_attr else None
        _, dlogits = ops.ce_fwd_bwd(logits.detach(), target, y_true, stats=self.stats)
        bk.grad.zero_()                       # optimizer.zero_grad()
        logits.backward(dlogits)              # loss.backward(): wgrad / dgrad / transpose-sum kernels
        bk.rebind_grads()
        self.optimizer_step()
        return logits


class CoTeachingTrainer:
    """Loop body of ``PipelineCO.train_ct`` (reference src/pipeline.py:111-133) for two peer SAGE networks on the fused
    path: two ngnn_sage_forward calls on the same block, ONE ngnn_ct_loss (device-side small-loss selection and exchange,
    reference src/utils/losses.py:19-49 without its two host argsorts), two ngnn_sage_backward calls, two Adam steps.
    Nothing in the step synchronises with the host; ``stats`` accumulates
    [loss_1, loss_2, correct_1, correct_2, pure_ratio_1, pure_ratio_2] sums for the epoch logging."""

    def __init__(self, model1, model2, lr: float = 1e-3, **adam):
        self.t1, self.t2 = Trainer(model1, lr=lr, **adam), Trainer(model2, lr=lr, **adam)
        self.stats = torch.zeros(6, dtype=torch.float32, device=self.t1.buckets.param.device)

    def reset_stats(self):
        self.stats.zero_()

    def read_stats(self):
        return self.stats.tolist()

    def train_step(self, batch, forget_rate: float, target_attr: str = "yhn", label_attr: Optional[str] = "y",
                   clean_attr: Optional[str] = None):
        loader, blk = batch._loader, batch.block
        bs = batch.batch_size
        out1 = self.t1.forward_only(batch)
        out2 = self.t2.forward_only(batch)
        num_remember = int((1 - forget_rate) * bs)                       # reference losses.py:28-29
        tgt = loader.label_array(target_attr)
        lab = loader.label_array(label_attr) if label_attr else None
        clean = loader.node_attrs[clean_attr].view(-1) if clean_attr else None
        _, d1, d2, _, _ = ops.ct_loss(out1, out2, tgt, num_remember, y_true=lab, row_ids=blk.n_id, clean_mask=clean,
                                      stats=self.stats)
        self.t1.backward_from(batch, d1)
        self.t1.optimizer_step()
        self.t2.backward_from(batch, d2)
        self.t2.optimizer_step()
        prefetch = getattr(batch, "_prefetch", None)
        if prefetch is not None:
            prefetch()
        return out1, out2
