"""Data-parallel plumbing over torch.distributed (NCCL on B200 / NVLink 5; gloo for CPU tests).

Replaces the Horovod path of the reference (utils/distributed.py):
  * `all_reduce_and_rescale_tensors(grads, 1.0)` (:16-43; call sites train_vqa.py:193-199,
    pretrain.py:302-308) — copy every grad into one flat buffer, `hvd.allreduce_` (Horovod 0.16.4
    default = AVERAGE over ranks), copy back — becomes `GradientReducer`: the encoder-layer
    gradients already live in ONE flat arena (`UniterModel.grad_arena()`, parameters' .grad are
    views of it), so they are all-reduced in place with no copy-in / copy-out; the few remaining
    parameters (embeddings, pooler, task head) go through one small flat bucket.
  * `broadcast_tensors(params, 0)` (:100-148; train_vqa.py:147) becomes `broadcast_parameters`.
One process per GPU; the path shards by samples only (pure data parallelism).
"""
import torch
import torch.distributed as dist

from .model import UniterModel


def _avg_all_reduce(t, async_op=False, group=None):
    """Mean over ranks.  NCCL reduces with AVG directly; gloo (CPU tests) sums then divides."""
    if dist.get_backend() == "nccl":
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op, group=group)
    w = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=False)
    t.div_(dist.get_world_size())
    return None


def _capped_nccl_group(max_ctas):
    """A second NCCL communicator whose collectives use at most `max_ctas` CTAs, so that the
    overlapped all-reduce and the persistent GEMM / row kernels (which leave exactly that many SMs
    free, ub200_set_sm_reserve) do not fight over SMs.  None if this torch build cannot cap it."""
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(max_ctas)
        opts.config.min_ctas = min(int(max_ctas), 4)
        return dist.new_group(backend="nccl", pg_options=opts)
    except Exception:     # older torch / NCCL without per-communicator config
        return None


def broadcast_parameters(model, root=0):
    """Rank `root`'s parameters and buffers -> every rank (startup only)."""
    works = []
    for t in list(model.parameters()) + list(model.buffers()):
        works.append(dist.broadcast(t.data, src=root, async_op=True))
    for w in works:
        w.wait()


class GradientReducer:
    """Average gradients over ranks after backward: arena in place + one bucket for the rest."""

    def __init__(self, model, overlap_chunks=4, sm_reserve=0):
        """`overlap_chunks` > 1: the encoder arena is all-reduced in that many layer groups while
        the backward of the earlier layers still runs.  `sm_reserve` > 0: during that overlap the
        library's persistent kernels leave `sm_reserve` SMs to the collective, and the collective
        runs on a communicator capped to the same number of CTAs (must be called by all ranks).
        Measured on 2 x B200 (C2, profiles/r01_scale2_variants.json): no overlap 5.40 ms/step,
        4 chunks 5.13, 4 chunks + reserve 8 / 16: 5.22 / 5.24 — NCCL's CTAs co-reside with the
        persistent CTAs well enough that giving up SMs costs more than it saves, hence default 0."""
        self.model = model
        self.sm_reserve = int(sm_reserve)
        self._group = None
        self._lib = None
        if self.sm_reserve > 0 and overlap_chunks > 1 and dist.is_initialized() and dist.get_backend() == "nccl":
            self._group = _capped_nccl_group(self.sm_reserve)
            from . import _lib
            self._lib = _lib.load()
        self.encoders = [m for m in model.modules() if isinstance(m, UniterModel)]
        self._others = None
        self._flat = None
        self.overlap_chunks = overlap_chunks
        self._pending = []
        self._reduced = set()
        self._comm_stream = None
        self._reserved = False

    # ---- overlap: called by _EncoderStack.backward after the kernels of layers [lo, hi) are enqueued
    def _on_chunk(self, enc, lo, hi):
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        ev = torch.cuda.Event()
        ev.record()
        if self._lib is not None and not self._reserved:
            self._lib.ub200_set_sm_reserve(self.sm_reserve)   # launches after this point
            self._reserved = True
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ev)
            self._pending.append(_avg_all_reduce(enc.arena_slice(lo, hi), async_op=True, group=self._group))
        self._reduced.add(id(enc))

    def _arena_param_ids(self):
        ids = set()
        for enc in self.encoders:
            if enc._arena is None:
                enc._build_arena()
            ids.update(id(p) for p, _ in enc._arena["views"])
        return ids

    def reduce(self):
        works = []
        arena_ids = self._arena_param_ids()
        works += self._pending
        self._pending = []
        for enc in self.encoders:
            if id(enc) not in self._reduced:      # not already reduced chunk-wise during backward
                works.append(_avg_all_reduce(enc.grad_arena(), async_op=True, group=self._group))
        self._reduced = set()
        # parameters outside the arena (tied weights appear once: parameters() de-duplicates)
        others = [p for p in self.model.parameters() if p.grad is not None and id(p) not in arena_ids]
        if others:
            grads = [p.grad for p in others]
            n = sum(g.numel() for g in grads)
            if self._flat is None or self._flat.numel() != n or self._flat.dtype != grads[0].dtype:
                self._flat = torch.empty(n, device=grads[0].device, dtype=grads[0].dtype)
            views = list(self._flat.split([g.numel() for g in grads]))
            torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
            w = _avg_all_reduce(self._flat, async_op=True)
            if w is not None:
                w.wait()
            torch._foreach_copy_([g.view(-1) for g in grads], views)
        for w in works:
            if w is not None:
                w.wait()

    def backward_and_reduce(self, loss):
        """loss.backward() with the encoder gradients all-reduced chunk by chunk while the rest of
        the backward is still running (replaces the non-overlapped Horovod call of
        train_vqa.py:193-199), then the remaining parameters."""
        overlap = self.overlap_chunks > 1 and dist.get_backend() == "nccl"
        if overlap:
            for enc in self.encoders:
                enc._bwd_chunk_hook = self._on_chunk
                enc._bwd_chunks = self.overlap_chunks
        try:
            loss.backward()
        finally:
            for enc in self.encoders:
                enc._bwd_chunk_hook = None
            if self._reserved:
                self._lib.ub200_set_sm_reserve(0)
                self._reserved = False
        self.reduce()
