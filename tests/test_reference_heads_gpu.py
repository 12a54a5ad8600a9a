"""GPU, SURVEY.md §8c G6: the reference's OWN, UNMODIFIED task heads (model/vqa.py,
model/pretrain.py, model/itm.py staged in oracle/_ref by oracle/make_ref.py) running on top of the
drop-in `uniter_b200.UniterModel` — the INTEGRATION.md recipe (`model.<head>.UniterModel = ours`)
exercised end to end on the device, forward and backward.

Logits are checked against the goldens the same reference heads produced over the reference
encoder on CPU (tests/golden/heads_tiny.npz), north-star tolerance 1e-2 in fp16; gradients of the
reference head's own loss against the CPU oracle.
"""
import pytest
import torch

from oracle import encoder_oracle as orc
from oracle import ref_loader
from tests import util

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="reference sources not staged")]


class _swap:
    """with _swap(rvqa, rpre): the reference modules construct OUR UniterModel."""

    def __init__(self, *mods):
        self.mods = mods

    def __enter__(self):
        from uniter_b200.model import UniterModel
        self.saved = [m.UniterModel for m in self.mods]
        for m in self.mods:
            m.UniterModel = UniterModel

    def __exit__(self, *a):
        for m, s in zip(self.mods, self.saved):
            m.UniterModel = s


def _tiny_ref_config(rm):
    c = util.TINY
    return rm.UniterConfig(c["vocab_size"], hidden_size=c["hidden_size"],
                           num_hidden_layers=c["num_hidden_layers"],
                           num_attention_heads=c["num_attention_heads"],
                           intermediate_size=c["intermediate_size"],
                           max_position_embeddings=c["max_position_embeddings"],
                           type_vocab_size=c["type_vocab_size"])


def _set_dropout_zero(model):
    """utils/misc.set_dropout (utils/misc.py:57-63) with p = 0: train mode made deterministic."""
    for _, module in model.named_modules():
        if isinstance(module, torch.nn.Dropout):
            module.p = 0.0


def _tensors(batch):
    return {k: v.cuda() for k, v in batch.items() if torch.is_tensor(v)}


def test_unmodified_reference_vqa_head_over_drop_in_encoder():
    from uniter_b200.model import UniterModel
    from uniter_b200.synth import seeded_state
    rm, rvqa = ref_loader.load("model.model", "model.vqa")
    g = util.load_golden("heads_tiny")
    with _swap(rvqa):
        vqa = rvqa.UniterForVisualQuestionAnswering(_tiny_ref_config(rm), 64, 17)
    assert isinstance(vqa.uniter, UniterModel)
    st = seeded_state({k: tuple(v.shape) for k, v in vqa.state_dict().items()}, seed=3)
    vqa.load_state_dict(st, strict=True)
    vqa = vqa.cuda().half().eval()
    batch = util.heads_batch()
    b = _tensors(batch)
    b["targets"] = torch.rand(3, 17, generator=torch.Generator().manual_seed(5)).cuda().half()
    logits = vqa(b, compute_loss=False)
    err = (logits.float().cpu() - torch.from_numpy(g["vqa_logits"])).abs().max().item()
    assert err <= 1e-2, err
    # backward of the reference's own loss (model/vqa.py:46-49) through the CUDA encoder
    loss = vqa(b, compute_loss=True)
    (loss.float().mean() * 256.0).backward()
    rs = {k: v.half().float().requires_grad_(True) for k, v in st.items()}
    enc = {k[len("uniter."):]: v for k, v in rs.items() if k.startswith("uniter.")}
    seq = orc.uniter_forward(enc, 2, 2, batch["input_ids"], batch["position_ids"],
                             batch["img_feat"].half().float(), batch["img_pos_feat"].half().float(),
                             batch["attn_masks"], batch["gather_index"], output_all_encoded_layers=False)
    ref_logits = orc.vqa_head(rs, orc.pooler(enc, seq))
    torch.nn.functional.binary_cross_entropy_with_logits(
        ref_logits, b["targets"].float().cpu(), reduction="none").mean().backward()
    params = dict(vqa.named_parameters())
    for name in ("uniter.encoder.layer.1.intermediate.dense.weight",
                 "uniter.encoder.layer.0.attention.self.value.weight", "uniter.pooler.dense.weight",
                 "uniter.img_embeddings.img_linear.weight", "vqa_output.0.weight"):
        got = params[name].grad.float().cpu() / 256.0
        want = rs[name].grad
        rel = ((got - want).norm() / (want.norm() + 1e-12)).item()
        assert rel <= 4e-2, (name, rel)


def test_unmodified_reference_pretraining_heads_over_drop_in_encoder():
    """UniterForPretraining.forward(batch, task) for mlm / itm / mrfr / mrc: the reference's own
    forward_* code (model/pretrain.py:107-229) over the CUDA encoder; MLM / ITM logits against the
    reference goldens, MRFR / MRC against the CPU oracle through the same reference heads on CPU."""
    from uniter_b200.synth import seeded_state
    rm, rpre = ref_loader.load("model.model", "model.pretrain")
    g = util.load_golden("heads_tiny")
    cfg = _tiny_ref_config(rm)
    with _swap(rpre):
        pre = rpre.UniterForPretraining(cfg, 64, 11)
    assert pre.cls.predictions.decoder.weight is pre.uniter.embeddings.word_embeddings.weight
    assert pre.feat_regress.weight is pre.uniter.img_embeddings.img_linear.weight
    st = seeded_state({k: tuple(v.shape) for k, v in pre.state_dict().items()}, seed=4)
    pre.load_state_dict(st, strict=True)
    cpu_ref = rpre.UniterForPretraining(cfg, 64, 11)          # the reference over its own encoder
    cpu_ref.load_state_dict({k: v.half().float() for k, v in st.items()}, strict=True)
    cpu_ref.eval()
    pre = pre.cuda().half().eval()
    batch = util.heads_batch()
    b = _tensors(batch)
    with torch.no_grad():
        scores = pre(b, task="mlm", compute_loss=False)
    err = (scores.float().cpu() - torch.from_numpy(g["mlm_scores"])).abs().max().item()
    assert err <= 1e-2, ("mlm", err)
    bi = dict(b)
    bi["targets"] = torch.tensor([1, 0, 1]).cuda()
    bi["ot_inputs"] = None
    with torch.no_grad():
        itm, _ = pre(bi, task="itm", compute_loss=False)
    err = (itm.float().cpu() - torch.from_numpy(g["itm_scores"])).abs().max().item()
    assert err <= 1e-2, ("itm", err)
    # MRFR / MRC: masked regions (model/pretrain.py:135-154, :201-229)
    gen = torch.Generator().manual_seed(8)
    img_masks = torch.rand(batch["img_feat"].shape[:2], generator=gen) < 0.4
    for i, nb in enumerate(batch["num_bbs"]):
        img_masks[i, nb:] = False
    img_masks[0, 0] = True
    Lt = batch["input_ids"].size(1)
    img_mask_tgt = torch.zeros_like(batch["attn_masks"], dtype=torch.bool)
    for i, tl in enumerate(batch["txt_lens"]):
        nb = batch["num_bbs"][i]
        img_mask_tgt[i, tl:tl + nb] = img_masks[i, :nb]
    n = int(img_masks.sum())
    cb = {k: v for k, v in batch.items() if torch.is_tensor(v)}
    cb["img_feat"] = cb["img_feat"].half().float()
    cb["img_pos_feat"] = cb["img_pos_feat"].half().float()
    extra = {"img_masks": img_masks, "img_mask_tgt": img_mask_tgt,
             "feat_targets": batch["img_feat"][img_masks].half().float(),
             "label_targets": torch.softmax(torch.randn(n, 11, generator=gen), -1)}
    for task in ("mrfr", "mrc"):
        with torch.no_grad():
            want = cpu_ref(dict(cb, **extra), task=task, compute_loss=False)
            dev_extra = {k: (v.cuda().half() if v.is_floating_point() else v.cuda()) for k, v in extra.items()}
            got = pre(dict(b, **dev_extra), task=task, compute_loss=False)
        err = (got.float().cpu() - want).abs().max().item()
        assert got.shape == want.shape and err <= 1e-2, (task, err)


def test_unmodified_reference_hard_negative_itm_over_drop_in_encoder():
    """model/itm.py:57-147 (UniterForImageTextRetrievalHardNeg): no-grad eval scoring of all pairs,
    top-k hard negatives, train-mode forward + backward on the selected rows — the reference's own
    class driving the CUDA encoder through model.train()/eval() toggles inside one step."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_goldens
    from uniter_b200.synth import seeded_state
    rm, ritm = ref_loader.load("model.model", "model.itm")
    c = dict(util.TINY, img_dim=16)
    cfg = _tiny_ref_config(rm)
    with _swap(ritm):
        mod = ritm.UniterForImageTextRetrievalHardNeg(cfg, 16, hard_size=3)
    st = seeded_state({k: tuple(v.shape) for k, v in mod.state_dict().items()}, seed=6)
    mod.load_state_dict(st, strict=True)
    mod = mod.cuda().half().train()
    _set_dropout_zero(mod)
    compared = 0
    for sf in ("t", "i"):
        batch, _ = make_goldens.hardneg_inputs(sf, seed=77)
        b = {k: v.cuda() for k, v in batch.items()}
        picked = {}
        orig = mod._get_hard_batch
        mod._get_hard_batch = lambda bt, sc, sfrom, _o=orig: picked.setdefault("gpu", _o(bt, sc, sfrom))
        mod.zero_grad(set_to_none=True)
        loss = mod(b, sample_from=sf, compute_loss=True)
        mod._get_hard_batch = orig
        assert loss.shape[0] == 1 and torch.isfinite(loss.float()).all()
        loss.float().mean().backward()
        gw = mod.uniter.encoder.layer[0].intermediate.dense.weight.grad
        assert gw is not None and torch.isfinite(gw.float()).all()
        # the same step through the reference class over the reference encoder (CPU fp32)
        ref = ritm.UniterForImageTextRetrievalHardNeg(cfg, 16, hard_size=3)
        ref.load_state_dict({k: v.half().float() for k, v in st.items()}, strict=True)
        ref.train()
        _set_dropout_zero(ref)
        cbatch, _ = make_goldens.hardneg_inputs(sf, seed=77)
        cbatch["img_feat"] = cbatch["img_feat"].half().float()
        cbatch["img_pos_feat"] = cbatch["img_pos_feat"].half().float()
        rorig = ref._get_hard_batch
        ref._get_hard_batch = lambda bt, sc, sfrom, _o=rorig: picked.setdefault("cpu", _o(bt, sc, sfrom))
        rloss = ref(cbatch, sample_from=sf, compute_loss=True)
        key = "img_feat" if sf == "t" else "input_ids"
        g_rows, c_rows = picked["gpu"][key].float().cpu(), picked["cpu"][key].float()
        if g_rows.shape == c_rows.shape and torch.equal(g_rows.half(), c_rows.half()):
            # same hard negatives mined (top-k over 16-bit scores can legitimately differ on near ties)
            assert (loss.float().cpu() - rloss.detach()).abs().max().item() <= 1e-2, sf
            compared += 1
    assert compared >= 1


@pytest.mark.parametrize("use_index", [True, False])
def test_library_pretraining_heads_match_the_reference_model(use_index):
    """OUR UniterForPretraining (every head on libub200: LibTransform / LibLinear / fused MLM head /
    library pooler) against the UNMODIFIED reference UniterForPretraining over the reference encoder
    on CPU fp32 (weights rounded to fp16): logits of mlm / mrfr / mrc / itm within 1e-2 (north
    star), per-element losses, and gradients of a multi-task loss for the head parameters and both
    tied weights (decoder <-> word embeddings, feat_regress.weight <-> img_linear.weight)."""
    from uniter_b200.heads import UniterForPretraining
    from uniter_b200.synth import seeded_state, synth_batch, synth_mrm
    rm, rpre = ref_loader.load("model.model", "model.pretrain")
    cfg = _tiny_ref_config(rm)
    ref = rpre.UniterForPretraining(cfg, 64, 11)
    st = seeded_state({k: tuple(v.shape) for k, v in ref.state_dict().items()}, seed=4)
    ref.load_state_dict({k: v.half().float() for k, v in st.items()}, strict=True)
    ref.eval()
    mod = UniterForPretraining(util.tiny_config(), 64, 11)
    missing = mod.load_state_dict(st, strict=True)
    mod = mod.cuda().half().eval()
    base = synth_batch(5, 5, 9, 4, 8, seed=17, img_dim=64, vocab_size=2000, mlm_prob=0.3)
    mb = synth_mrm(base, mask_prob=0.3, label_dim=11, seed=3)
    keys = [k for k, v in mb.items() if torch.is_tensor(v)]
    if not use_index:           # the reference's own boolean-mask row selection
        keys = [k for k in keys if k not in ("mlm_index", "mlm_targets", "mrm_index", "mrm_valid", "mrm_inv_n")]
    cb = {k: (mb[k].half().float() if mb[k].is_floating_point() else mb[k]) for k in keys}
    db = {k: mb[k].cuda() for k in keys}
    cb["targets"] = torch.tensor([1, 0, 1, 1, 0])
    db["targets"] = cb["targets"].cuda()
    cb["ot_inputs"] = None
    plain_c = dict(cb, img_feat=base["img_feat"].half().float())      # mlm / itm see unmasked regions
    plain_d = dict(db, img_feat=base["img_feat"].cuda())
    total_c, total_d = 0.0, 0.0
    for task in ("mlm", "mrfr", "mrc", "mrc-kl", "itm"):
        bc, bd = (plain_c, plain_d) if task in ("mlm", "itm") else (cb, db)
        with torch.no_grad():
            want = ref(bc, task=task, compute_loss=False)
            got = mod(bd, task=task, compute_loss=False)
        want = want[0] if isinstance(want, tuple) else want
        got = got[0] if isinstance(got, tuple) else got
        assert got.shape == want.shape, (task, got.shape, want.shape)
        err = (got.float().cpu() - want).abs().max().item()
        assert err <= 1e-2, (task, err)
        lw = ref(bc, task=task, compute_loss=True)
        lg = mod(bd, task=task, compute_loss=True)
        lw = lw[0] if isinstance(lw, tuple) else lw
        lg = lg[0] if isinstance(lg, tuple) else lg
        assert lg.shape == lw.shape, (task, lg.shape, lw.shape)
        assert (lg.float().cpu() - lw.detach()).abs().max().item() <= 3e-2, task
        total_c = total_c + lw.float().mean()
        total_d = total_d + lg.float().mean()
    (total_d * 64.0).backward()
    total_c.backward()
    gp = dict(mod.named_parameters())
    rp = dict(ref.named_parameters())
    for name in ("feat_regress.net.0.weight", "feat_regress.net.2.weight", "feat_regress.bias",
                 "region_classifier.net.0.weight", "region_classifier.net.3.weight",
                 "region_classifier.net.3.bias", "itm_output.weight", "itm_output.bias",
                 "uniter.pooler.dense.weight", "uniter.pooler.dense.bias",
                 "cls.predictions.transform.dense.weight", "cls.predictions.bias",
                 "uniter.embeddings.word_embeddings.weight", "uniter.img_embeddings.img_linear.weight",
                 "uniter.img_embeddings.mask_embedding.weight",
                 "uniter.encoder.layer.1.output.dense.weight"):
        got = gp[name].grad.float().cpu() / 64.0
        want = rp[name].grad
        rel = ((got - want).norm() / (want.norm() + 1e-12)).item()
        assert rel <= 4e-2, (name, rel)
