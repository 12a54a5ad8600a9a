"""GPU: the host-free step (uniter_b200.graphed.GraphedStep) — CUDA-graph replay per token bucket.

Properties checked (all through the C ABI, tiny MLM model):
  * a replayed step equals the eager step on the same batch (dropout off): the dummy sequence that
    pads the token count to the bucket, the padded masked-token list and the static gradient arena
    change nothing — the loss BIT FOR BIT (the forward has no atomics), every gradient to within
    the run-to-run noise of the fp32 / bf16 atomics of the backward (split-K decoder dgrad, column
    sums, embedding scatter), i.e. orders of magnitude below the parity tolerance;
  * two different batches of one bucket replay ONE graph and each matches its eager step;
  * with dropout on, replays of the same batch draw different masks (device-side stream offset),
    and the backward of a replay regenerates the masks of its own forward (gradient check against
    a finite difference is not possible with dropout; instead: identical counter => identical step);
  * gradient accumulation (accumulate=True) adds into the arena.
"""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _model(dtype=torch.bfloat16, p_drop=0.0):
    from uniter_b200.heads import UniterForMLM
    torch.manual_seed(0)
    mod = UniterForMLM(util.tiny_config(), 64)
    mod.load_state_dict(util.head_state(mod, seed=9), strict=False)
    mod = mod.to("cuda", dtype).train()
    for m in mod.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = p_drop
    return mod


def _batch(seed, tl=None, nb=None):
    from uniter_b200.synth import pad_mlm_index, synth_batch
    if tl is None:
        b = synth_batch(6, 5, 12, 3, 9, seed=seed, img_dim=64, vocab_size=2000, mlm_prob=0.3)
    else:
        b = synth_batch(len(tl), 0, 0, 0, 0, seed=seed, img_dim=64, vocab_size=2000, mlm_prob=0.3,
                        txt_lens=tl, num_bbs=nb)
    b = pad_mlm_index(b, 16)
    lens = [a + c for a, c in zip(b["txt_lens"], b["num_bbs"])]
    return {k: v for k, v in b.items() if torch.is_tensor(v)}, lens


def _loss_fn(mod):
    return lambda b: (mod(b).sum() * b["mlm_inv_n"]).squeeze()


def _eager(mod, hb, lens):
    from uniter_b200.model import register_lengths
    b = {k: v.cuda() for k, v in hb.items()}
    register_lengths(b["attn_masks"], lens, prefix=True)
    mod.zero_grad(set_to_none=True)
    loss = _loss_fn(mod)(b)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach().clone(), {n: p.grad.detach().clone() for n, p in mod.named_parameters()
                                   if p.grad is not None}


def test_graph_replay_equals_eager_bit_for_bit_and_buckets_share_a_graph():
    from uniter_b200.graphed import GraphedStep
    mod = _model()
    # same shapes (Lt, Li, L, B) but different valid lengths -> different T, same 64-token bucket
    tl_a, nb_a = [12, 5, 7, 9, 12, 6], [9, 3, 4, 9, 5, 7]      # T = 88
    tl_b, nb_b = [12, 8, 5, 9, 11, 6], [9, 4, 6, 3, 5, 8]      # T = 86
    ha, la = _batch(31, tl_a, nb_a)
    hb, lb = _batch(32, tl_b, nb_b)
    assert {k: v.shape for k, v in ha.items()} == {k: v.shape for k, v in hb.items()}
    step = GraphedStep(mod, _loss_fn(mod), token_bucket=64)
    for hbatch, lens in ((ha, la), (hb, lb), (ha, la)):
        ref_loss, ref_g = _eager(mod, hbatch, lens)
        loss = step(hbatch, lens)
        torch.cuda.synchronize()
        assert torch.equal(loss, ref_loss), (loss.item(), ref_loss.item())
        got = {n: p.grad for n, p in mod.named_parameters() if p.grad is not None}
        assert set(ref_g) <= set(got)
        for n, g in ref_g.items():
            d = (got[n].float() - g.float()).norm().item()
            assert d <= 4e-3 * g.float().norm().item() + 1e-6, (n, d, g.float().norm().item())
    assert step.captures == 1 and len(step.buckets) == 1
    bk = next(iter(step.buckets.values()))
    assert bk.T_pad == 128 and bk.launches > 20
    # gradients live in the arena at fixed addresses
    q = mod.uniter.encoder.layer[0].attention.self.query.weight
    assert q.grad.data_ptr() == step.arena.view(q).data_ptr()


def test_graph_replays_draw_fresh_dropout_masks_and_accumulate():
    from uniter_b200.graphed import GraphedStep
    mod = _model(p_drop=0.1)
    hb, lens = _batch(41)
    step = GraphedStep(mod, _loss_fn(mod), token_bucket=64)
    l1 = step(hb, lens).clone()
    g1 = mod.uniter.encoder.layer[1].output.dense.weight.grad.clone()
    l2 = step(hb, lens).clone()
    g2 = mod.uniter.encoder.layer[1].output.dense.weight.grad.clone()
    torch.cuda.synchronize()
    assert torch.isfinite(l1) and torch.isfinite(l2)
    assert not torch.equal(l1, l2) and not torch.equal(g1, g2)        # different masks per replay
    # same device counter => the very same step (forward and backward masks are a pure function of it)
    c = step.rng_counter.clone()
    step.rng_counter.copy_(c - 64)
    l2b = step(hb, lens).clone()
    torch.cuda.synchronize()
    assert torch.equal(l2b, l2)
    g2b = mod.uniter.encoder.layer[1].output.dense.weight.grad.float()
    assert ((g2b - g2.float()).norm() / g2.float().norm()).item() < 4e-3
    # accumulation: a second graph (accumulate=True) adds to what the first left in the arena
    mod2 = _model(p_drop=0.0)
    step2 = GraphedStep(mod2, _loss_fn(mod2), token_bucket=64)
    step2(hb, lens)
    w = mod2.uniter.encoder.layer[0].intermediate.dense.weight
    e = mod2.uniter.embeddings.word_embeddings.weight
    gw, ge = w.grad.float().clone(), e.grad.float().clone()
    step2(hb, lens, accumulate=True)
    torch.cuda.synchronize()
    assert step2.captures == 2
    assert ((w.grad.float() - 2 * gw).norm() / (2 * gw).norm()).item() < 2e-2
    assert ((e.grad.float() - 2 * ge).norm() / (2 * ge).norm()).item() < 2e-2
