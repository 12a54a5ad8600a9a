"""CPU: host logic of the whole-model gradient arena (uniter_b200.arena) — layout, the claim /
begin_step write protocols and the folding of autograd-allocated gradients.  (The kernels that
write into the views are exercised by the GPU tests; this is the bookkeeping they rely on.)"""
import torch

from tests import util


def _model():
    from uniter_b200.heads import UniterForPretraining
    torch.manual_seed(0)
    return UniterForPretraining(util.tiny_config(), 64, 11)


def test_arena_layout_covers_every_parameter_once_in_completion_order():
    from uniter_b200.arena import GradArena
    mod = _model()
    arena = GradArena.attach(mod)
    assert GradArena.attach(mod) is arena                       # idempotent
    params = {id(p): (n, p) for n, p in mod.named_parameters()}
    spans = []
    for pid, (n, p) in params.items():
        v = arena.view(p)
        assert v.shape == p.shape and v.dtype == p.dtype and p._ub_grad_view is v
        off = arena._offsets[pid]
        assert off % 8 == 0 or n.endswith(("key.weight", "value.weight", "key.bias", "value.bias")) or \
            "LayerNorm" in n or "layer_norm" in n or "bias" in n or "embeddings" in n or "pos_linear" in n, n
        assert v.data_ptr() == arena.flat.data_ptr() + off * arena.flat.element_size()
        spans.append((off, off + p.numel(), n))
    spans.sort()
    for (a0, a1, an), (b0, b1, bn) in zip(spans, spans[1:]):
        assert a1 <= b0, (an, bn)                                # no overlap
    # tied parameters appear once (decoder <-> word embeddings, feat_regress <-> img_linear)
    assert mod.cls.predictions.decoder.weight is mod.uniter.embeddings.word_embeddings.weight
    assert len(params) == len(list(mod.parameters()))
    seg = arena.segments
    order = [seg["head"], seg["enc0.pooler"], seg["enc0.layers"], seg["enc0.front"]]
    assert order[0][0] == 0 and all(a[1] == b[0] for a, b in zip(order, order[1:])) and order[-1][1] == arena.numel
    # q / k / v gradients of a layer are contiguous (the fused QKV wgrad writes them as one [3H, H] block)
    att = mod.uniter.encoder.layer[1].attention.self
    H = att.query.weight.size(0)
    assert arena._offsets[id(att.key.weight)] == arena._offsets[id(att.query.weight)] + H * H
    assert arena._offsets[id(att.value.weight)] == arena._offsets[id(att.query.weight)] + 2 * H * H
    # the layer slices the reducer ships are contiguous and tile the "layers" segment
    lo = mod.uniter.arena_slice(0, 1).data_ptr()
    hi = mod.uniter.arena_slice(1, 2).data_ptr()
    per = arena._enc_plans[0]["per_layer"]
    assert hi - lo == per * arena.flat.element_size() and seg["enc0.layers"][1] - seg["enc0.layers"][0] == 2 * per


def test_claim_protocols_and_foreign_gradients():
    from uniter_b200.arena import GradArena
    mod = _model()
    arena = GradArena.attach(mod)
    w = mod.uniter.encoder.layer[0].intermediate.dense.weight
    word = mod.uniter.embeddings.word_embeddings.weight
    # implicit protocol: .grad None = fresh -> overwrite; attached view = accumulate
    assert w.grad is None and arena.claim([w]) is False and w.grad.data_ptr() == arena.view(w).data_ptr()
    assert arena.claim([w]) is True
    mod.zero_grad(set_to_none=True)
    assert arena.claim([w, word]) is False
    # a group with one live member accumulates; its fresh members are zeroed first
    mod.zero_grad(set_to_none=True)
    arena.view(word).fill_(3.0)                                  # stale data from an earlier step
    arena.claim([w])
    arena.view(w).fill_(1.0)
    assert arena.claim([w, word]) is True
    assert float(arena.view(word).abs().sum()) == 0.0 and float(arena.view(w)[0, 0]) == 1.0
    # a gradient tensor autograd allocated itself is folded into the view
    mod.zero_grad(set_to_none=True)
    word.grad = torch.full_like(word, 2.0)
    assert arena.claim([word]) is True and word.grad.data_ptr() == arena.view(word).data_ptr()
    assert float(arena.view(word)[5, 5]) == 2.0
    b = mod.itm_output.bias
    b.grad = torch.ones_like(b)
    assert arena.fold_foreign() == 1 and b.grad.data_ptr() == arena.view(b).data_ptr()
    # explicit protocol: managed parameters become fresh, the autograd-managed slices are zeroed,
    # every .grad is attached (fixed addresses for CUDA-graph capture)
    arena.mark_managed([w, word])
    arena.view(b).fill_(7.0)
    arena.view(w).fill_(5.0)
    arena.begin_step()
    assert float(arena.view(b).abs().sum()) == 0.0               # unmanaged: zeroed
    assert float(arena.view(w)[0, 0]) == 5.0                     # managed: left for its first writer
    assert all(p.grad is not None and p.grad.data_ptr() == arena.view(p).data_ptr() for p in mod.parameters())
    assert arena.claim([w]) is False and arena.claim([w]) is True
    arena.view(word).fill_(9.0)                                  # never claimed in this step: stale
    arena.finish_step()
    assert float(arena.view(word).abs().sum()) == 0.0 and float(arena.view(w)[0, 0]) == 5.0
    arena.begin_step(accumulate=True)
    assert arena.claim([word]) is True                           # accumulation window: nothing is fresh
    arena.begin_step(zero_all=True)
    assert float(arena.flat.abs().sum()) == 0.0 and arena.claim([w]) is True
    arena.end_step_mode()
