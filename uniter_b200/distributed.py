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


def _avg_all_reduce(t, async_op=False):
    """Mean over ranks.  NCCL reduces with AVG directly; gloo (CPU tests) sums then divides."""
    if dist.get_backend() == "nccl":
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
    w = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=False)
    t.div_(dist.get_world_size())
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

    def __init__(self, model):
        self.model = model
        self.encoders = [m for m in model.modules() if isinstance(m, UniterModel)]
        self._others = None
        self._flat = None

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
        for enc in self.encoders:
            works.append(_avg_all_reduce(enc.grad_arena(), async_op=True))
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
        loss.backward()
        self.reduce()
