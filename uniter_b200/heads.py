"""Callers of the hot path: the reference's task heads on top of the drop-in `UniterModel`.

Parameter names and module trees follow the reference so its checkpoints load
(`uniter.*`, `cls.predictions.*`, `vqa_output.*`, `itm_output.*`, `rank_output.*`):

* ``UniterForMLM`` — the MLM branch of ``UniterForPretraining`` (model/pretrain.py:50-60,
  107-133).  SURVEY.md §8f-1: the head runs on libub200 — masked rows are gathered straight from
  the PACKED encoder output, ``BertPredictionHeadTransform`` (model/layer.py:188-203) is the
  tcgen05 GEMM with the bias+GELU epilogue + the LayerNorm kernel, the tied decoder
  (model/layer.py:206-222) is the same GEMM over the un-padded [V, H] embedding table
  (``n_valid``), and the cross-entropy is one fused kernel per direction; the decoder's dgrad
  (few output tiles, K = V = 28996) runs split-K.
* ``UniterForVisualQuestionAnswering`` (model/vqa.py:16-52) and ``UniterForImageTextRetrieval``
  (model/itm.py:14-59): pooler + small classifier, plain torch on top of the encoder — they are
  here so the parity tests can check head-level logits against the reference's goldens.
"""
from collections import defaultdict

import torch
from torch import nn
from torch.nn import functional as F

from . import ops
from .model import LibLinear, UniterModel, UniterPreTrainedModel, gather_packed_rows


class GELU(nn.Module):
    """model/layer.py:40-42 (erf form)."""

    def forward(self, x):
        return F.gelu(x)


def gelu(x):
    return F.gelu(x)  # erf form, == model/layer.py:31-37


class BertPredictionHeadTransform(nn.Module):
    """model/layer.py:188-203 (parameter container; the fused head reads the weights by pointer)."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=1e-12)


class BertLMPredictionHead(nn.Module):
    """model/layer.py:206-222 — decoder weight tied to the word embeddings."""

    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(bert_model_embedding_weights.size(1),
                                 bert_model_embedding_weights.size(0), bias=False)
        self.decoder.weight = bert_model_embedding_weights
        self.bias = nn.Parameter(torch.zeros(bert_model_embedding_weights.size(0)))


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.predictions = BertLMPredictionHead(config, bert_model_embedding_weights)


_PAD_LOGIT = -30000.0   # finite in fp16 and bf16; exp(pad - max) underflows to exactly 0


class _MlmHead(torch.autograd.Function):
    """scores = decoder(LayerNorm(gelu(dense(h)))) + bias ; loss = cross_entropy(scores, targets)
    on libub200.  Returns the per-row loss (fp32) or, with ``want_scores``, the [n, V] scores.
    Rows whose target is outside [0, V) (padding of a fixed-size index list) get loss 0 and
    contribute no gradient.  The head's parameters are read by pointer and their gradients are
    written straight into the gradient arena (the tied decoder / word-embedding gradient is one
    buffer that the decoder wgrad writes first and the embedding scatter adds to later)."""

    @staticmethod
    def forward(ctx, h, module, targets, want_scores):
        p = module.cls.predictions
        dense_w, dense_b = p.transform.dense.weight, p.transform.dense.bias
        ln_g, ln_b = p.transform.LayerNorm.weight, p.transform.LayerNorm.bias
        word_w, dec_bias = p.decoder.weight, p.bias
        n, H = h.shape
        V = word_w.size(0)
        Vp = (V + 7) // 8 * 8
        dtype = h.dtype
        t, pre = ops.gemm(h, dense_w, bias=dense_b, gelu=True)        # model/layer.py:199-200
        z = ops.layernorm_fwd(t, ln_g, ln_b)                         # :201
        if Vp != V:
            bias_p = torch.full((Vp,), _PAD_LOGIT, device=h.device, dtype=dtype)
            bias_p[:V] = dec_bias
        else:
            bias_p = dec_bias
        logits = torch.empty(n, Vp, device=h.device, dtype=dtype)
        ops.gemm(z, word_w, bias=bias_p, out=logits, n_valid=V if Vp != V else 0)   # :220-221
        if want_scores:           # validation path (compute_loss=False): forward only
            scores = logits[:, :V]
            ctx.mark_non_differentiable(scores)
            return scores
        targets = targets.contiguous()
        loss, lse = ops.ce_fwd(logits, targets, V)                    # model/pretrain.py:122-125
        ctx.save_for_backward(h, pre, t, z, logits, lse, targets)
        ctx.V = V
        ctx.module = module
        return loss

    @staticmethod
    def backward(ctx, dloss):
        from .arena import GradArena
        h, pre, t, z, logits, lse, targets = ctx.saved_tensors
        module = ctx.module
        p = module.cls.predictions
        dense_w, dense_b = p.transform.dense.weight, p.transform.dense.bias
        ln_g, ln_b = p.transform.LayerNorm.weight, p.transform.LayerNorm.bias
        word_w, dec_bias = p.decoder.weight, p.bias
        arena = GradArena.for_params(module, dense_w)
        arena.mark_managed([dense_w, dense_b, ln_g, ln_b, dec_bias])
        V = ctx.V
        dtype = h.dtype
        # the scores are dead after this point: the gradient overwrites them
        dlog = ops.ce_bwd_(logits, targets, lse, dloss.contiguous().float(), V)
        dlv = dlog[:, :V]                                             # [n, V], row pitch Vp
        # dz = dlog W_dec: 2 x (H / 128) output tiles, K = V -> split-K over the SMs (fp32 atomics)
        dz = ops.cvt_from_f32(ops.gemm(dlv, word_w, b_major=1, k_splits=-1), dtype)
        acc = arena.claim([word_w])
        ops.gemm(dlv, z, a_major=1, b_major=1, out=arena.view(word_w), accumulate=acc)   # [V, H] = dlog^T z
        dt, _, dg, db, _ = ops.layernorm_bwd(dz, t, ln_g, want_dbias=False)
        dpre = ops.dgelu_mul(dt, pre)
        dh = ops.gemm(dpre, dense_w, b_major=1)
        acc = arena.claim([dense_w, dense_b, ln_g, ln_b, dec_bias])
        ops.gemm(dpre, h, a_major=1, b_major=1, out=arena.view(dense_w), accumulate=acc)
        ops.cvt_from_f32(ops.colsum(dpre), dtype, out=arena.view(dense_b), accumulate=acc)
        ops.cvt_from_f32(dg, dtype, out=arena.view(ln_g), accumulate=acc)
        ops.cvt_from_f32(db, dtype, out=arena.view(ln_b), accumulate=acc)
        ops.cvt_from_f32(ops.colsum(dlog)[:V], dtype, out=arena.view(dec_bias), accumulate=acc)
        return dh, None, None, None


class UniterForMLM(UniterPreTrainedModel):
    """The MLM branch of UniterForPretraining (model/pretrain.py:50-60, 107-133)."""

    def __init__(self, config, img_dim):
        super().__init__(config)
        self.uniter = UniterModel(config, img_dim)
        self.cls = BertOnlyMLMHead(config, self.uniter.embeddings.word_embeddings.weight)
        self.apply(self.init_weights)

    def forward(self, batch, compute_loss=True):
        """Per-masked-token loss [n] (fp32), or the scores [n, V].  With loader-provided
        `mlm_index` (flat positions b * L + j; entries equal to B * L and targets of -1 are PADDING
        of a fixed-size list: zero rows, zero loss, no gradient) nothing here depends on data on
        the device, so the step can be captured in a CUDA graph."""
        batch = defaultdict(lambda: None, batch)
        input_ids = batch["input_ids"]
        packed, meta = self.uniter.encode_packed(
            input_ids, batch["position_ids"], batch["img_feat"], batch["img_pos_feat"],
            batch["attn_masks"], batch["gather_index"], output_all_encoded_layers=False,
            txt_type_ids=batch["txt_type_ids"])
        L = meta["L"]
        if batch["mlm_index"] is not None:
            # loader-provided flat positions (b * L + j) of the masked tokens: static shapes, no sync
            flat, targets = batch["mlm_index"], batch["mlm_targets"]
        else:
            # model/pretrain.py:115-118,129-133: text part only, rows where txt_labels != -1
            txt_labels = batch["txt_labels"]
            pos = (txt_labels != -1).nonzero(as_tuple=False)           # device sync, like the reference
            flat = pos[:, 0] * L + pos[:, 1]
            targets = txt_labels[pos[:, 0], pos[:, 1]]
        if flat.numel() == 0:
            V = self.uniter.config.vocab_size
            return packed.new_zeros(0, dtype=torch.float32) if compute_loss else packed.new_zeros(0, V)
        rows = meta["unpack_ext"][flat]                                # packed row of each masked token
        masked_output = gather_packed_rows(packed, rows)               # [n, H]
        return _MlmHead.apply(masked_output, self, targets, not compute_loss)


class LibTransform(torch.autograd.Function):
    """z = LayerNorm(gelu(h W^T + b)) — BertPredictionHeadTransform (model/layer.py:188-203) and the
    `net.0 / net.1 / net.2` prefix of RegionFeatureRegression / RegionClassification
    (model/pretrain.py:19-47) — on libub200: GEMM with the bias + GELU epilogue, LayerNorm kernels,
    dGELU kernel, dgrad / wgrad GEMMs.  Gradients of parameters that live in a gradient arena are
    written there directly."""

    @staticmethod
    def forward(ctx, h, dense_w, dense_b, ln_g, ln_b):
        h = h.contiguous()
        t, pre = ops.gemm(h, dense_w, bias=dense_b, gelu=True)
        z = ops.layernorm_fwd(t, ln_g, ln_b)
        ctx.save_for_backward(h, pre, t)
        ctx.params = (dense_w, dense_b, ln_g, ln_b)
        return z

    @staticmethod
    def backward(ctx, dz):
        h, pre, t = ctx.saved_tensors
        dense_w, dense_b, ln_g, ln_b = ctx.params
        dtype = h.dtype
        dt, _, dg, db, _ = ops.layernorm_bwd(dz.contiguous(), t, ln_g, want_dbias=False)
        dpre = ops.dgelu_mul(dt, pre)
        dh = ops.gemm(dpre, dense_w, b_major=1) if ctx.needs_input_grad[0] else None
        arena = getattr(dense_w, "_ub_arena", None)
        if arena is not None and arena._still_valid() and all(id(q) in arena._views for q in ctx.params):
            arena.mark_managed(ctx.params)
            acc = arena.claim(list(ctx.params))
            ops.gemm(dpre, h, a_major=1, b_major=1, out=arena.view(dense_w), accumulate=acc)
            ops.cvt_from_f32(ops.colsum(dpre), dtype, out=arena.view(dense_b), accumulate=acc)
            ops.cvt_from_f32(dg, dtype, out=arena.view(ln_g), accumulate=acc)
            ops.cvt_from_f32(db, dtype, out=arena.view(ln_b), accumulate=acc)
            return dh, None, None, None, None
        return (dh, ops.gemm(dpre, h, a_major=1, b_major=1), ops.cvt_from_f32(ops.colsum(dpre), dtype),
                ops.cvt_from_f32(dg, dtype), ops.cvt_from_f32(db, dtype))


class RegionFeatureRegression(nn.Module):
    """model/pretrain.py:19-32 (MRFR head): LN(gelu(dense(h))) @ img_linear.weight + bias — the
    output projection is TIED to the image embedding's input projection, read as a [K, N] operand."""

    def __init__(self, hidden_size, feat_dim, img_linear_weight):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), GELU(),
                                 nn.LayerNorm(hidden_size, eps=1e-12))
        self.weight = img_linear_weight
        self.bias = nn.Parameter(torch.zeros(feat_dim))

    def forward(self, input_):
        hidden = LibTransform.apply(input_, self.net[0].weight, self.net[0].bias, self.net[2].weight,
                                    self.net[2].bias)
        return LibLinear.apply(hidden, self.weight, self.bias, True, False)


class RegionClassification(nn.Module):
    """model/pretrain.py:35-47 (MRC / MRC-kl head)."""

    def __init__(self, hidden_size, label_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), GELU(),
                                 nn.LayerNorm(hidden_size, eps=1e-12), nn.Linear(hidden_size, label_dim))

    def forward(self, input_):
        hidden = LibTransform.apply(input_, self.net[0].weight, self.net[0].bias, self.net[2].weight,
                                    self.net[2].bias)
        return LibLinear.apply(hidden, self.net[3].weight, self.net[3].bias, False, False)


def _masked_rows(packed, meta, mask2d, index):
    """Rows of the packed encoder output selected by a [B, L] boolean mask (model/pretrain.py:129-133,
    `_compute_masked_hidden`) — or by a loader-provided flat index list (b * L + j; entries equal to
    B * L are padding of a fixed-size list -> zero rows), which needs no device read."""
    if index is None:
        pos = mask2d.nonzero(as_tuple=False)                   # device sync, like the reference
        index = pos[:, 0] * meta["L"] + pos[:, 1]
    return gather_packed_rows(packed, meta["unpack_ext"][index])


class UniterForPretraining(UniterPreTrainedModel):
    """model/pretrain.py:50-229 with every head on libub200: MLM (fused head + cross-entropy), MRFR,
    MRC / MRC-kl (LibTransform + LibLinear over the masked REGION rows gathered straight from the
    packed encoder output) and ITM (library pooler + LibLinear).  Same parameter names, weight tying
    (cls.predictions.decoder <-> word_embeddings, feat_regress.weight <-> img_linear.weight) and
    forward(batch, task, compute_loss) contract; the OT term of ITM (model/ot.py) is outside the hot
    path and not implemented (ot_inputs must be None).

    Fixed-shape (CUDA-graph friendly) variants of the reference's data-dependent row selections are
    taken from the batch when the loader provides them: `mlm_index` / `mlm_targets`, `mrm_index`."""

    def __init__(self, config, img_dim, img_label_dim):
        super().__init__(config)
        self.uniter = UniterModel(config, img_dim)
        self.cls = BertOnlyMLMHead(config, self.uniter.embeddings.word_embeddings.weight)
        self.feat_regress = RegionFeatureRegression(config.hidden_size, img_dim,
                                                    self.uniter.img_embeddings.img_linear.weight)
        self.region_classifier = RegionClassification(config.hidden_size, img_label_dim)
        self.itm_output = nn.Linear(config.hidden_size, 2)
        self.apply(self.init_weights)

    def _encode(self, batch, img_masks=None):
        return self.uniter.encode_packed(
            batch["input_ids"], batch["position_ids"], batch["img_feat"], batch["img_pos_feat"],
            batch["attn_masks"], batch["gather_index"], output_all_encoded_layers=False,
            img_masks=img_masks, txt_type_ids=batch["txt_type_ids"])

    def forward(self, batch, task, compute_loss=True):
        batch = defaultdict(lambda: None, batch)
        if task == "mlm":
            return self.forward_mlm(batch, compute_loss)
        if task == "mrfr":
            return self.forward_mrfr(batch, compute_loss)
        if task == "itm":
            return self.forward_itm(batch, compute_loss)
        if task.startswith("mrc"):
            return self.forward_mrc(batch, task, compute_loss)
        raise ValueError("invalid task")

    def forward_mlm(self, batch, compute_loss=True):                      # model/pretrain.py:107-127
        packed, meta = self._encode(batch)
        L = meta["L"]
        if batch["mlm_index"] is not None:
            flat, targets = batch["mlm_index"], batch["mlm_targets"]
        else:
            txt_labels = batch["txt_labels"]
            pos = (txt_labels != -1).nonzero(as_tuple=False)
            flat = pos[:, 0] * L + pos[:, 1]
            targets = txt_labels[pos[:, 0], pos[:, 1]]
        masked_output = gather_packed_rows(packed, meta["unpack_ext"][flat])
        return _MlmHead.apply(masked_output, self, targets, not compute_loss)

    def forward_mrfr(self, batch, compute_loss=True):                     # model/pretrain.py:135-154
        packed, meta = self._encode(batch, img_masks=batch["img_masks"])
        masked_output = _masked_rows(packed, meta, batch["img_mask_tgt"], batch["mrm_index"])
        prediction_feat = self.feat_regress(masked_output)
        if compute_loss:
            return F.mse_loss(prediction_feat, batch["feat_targets"].to(prediction_feat.dtype),
                              reduction="none")
        return prediction_feat

    def forward_itm(self, batch, compute_loss=True):                      # model/pretrain.py:156-199
        if batch["ot_inputs"] is not None:
            raise NotImplementedError("the OT / WRA term (model/ot.py) is outside the hot path")
        packed, meta = self._encode(batch)
        B, L = meta["n_batch"], meta["L"]
        cls_rows = meta["unpack_ext"][::L][:B]          # packed row of position (b, 0): the [CLS] token
        pooled = self.uniter.pooler(gather_packed_rows(packed, cls_rows.contiguous()))
        itm_scores = LibLinear.apply(pooled, self.itm_output.weight, self.itm_output.bias, False, False)
        if compute_loss:
            return F.cross_entropy(itm_scores.float(), batch["targets"], reduction="none"), None
        return itm_scores, None

    def forward_mrc(self, batch, task, compute_loss=True):                # model/pretrain.py:201-229
        packed, meta = self._encode(batch, img_masks=batch["img_masks"])
        masked_output = _masked_rows(packed, meta, batch["img_mask_tgt"], batch["mrm_index"])
        prediction_soft_label = self.region_classifier(masked_output)
        if not compute_loss:
            return prediction_soft_label
        label_targets = batch["label_targets"]
        if "kl" in task:
            logp = F.log_softmax(prediction_soft_label.float(), dim=-1)
            return F.kl_div(logp, label_targets.float(), reduction="none")
        label_targets = torch.max(label_targets[:, 1:], dim=-1)[1] + 1   # background is never a target
        return F.cross_entropy(prediction_soft_label.float(), label_targets, ignore_index=0, reduction="none")


class UniterForVisualQuestionAnswering(UniterPreTrainedModel):
    """model/vqa.py:16-52."""

    def __init__(self, config, img_dim, num_answer):
        super().__init__(config)
        self.uniter = UniterModel(config, img_dim)
        self.vqa_output = nn.Sequential(
            nn.Linear(config.hidden_size, config.hidden_size * 2),
            GELU(),
            nn.LayerNorm(config.hidden_size * 2, eps=1e-12),
            nn.Linear(config.hidden_size * 2, num_answer))
        self.apply(self.init_weights)

    def forward(self, batch, compute_loss=True):
        batch = defaultdict(lambda: None, batch)
        # pooler(sequence_output) (model/vqa.py:36-43) from the packed encoder output: only the [CLS]
        # rows are gathered, the padded [B, L, H] tensor is never materialised
        packed, meta = self.uniter.encode_packed(
            batch["input_ids"], batch["position_ids"], batch["img_feat"], batch["img_pos_feat"],
            batch["attn_masks"], batch["gather_index"], output_all_encoded_layers=False)
        B, L = meta["n_batch"], meta["L"]
        cls_rows = meta["unpack_ext"][::L][:B].contiguous()
        pooled_output = self.uniter.pooler(gather_packed_rows(packed, cls_rows))
        answer_scores = self.vqa_output(pooled_output)
        if compute_loss:
            return F.binary_cross_entropy_with_logits(answer_scores, batch["targets"], reduction="none")
        return answer_scores


class UniterForImageTextRetrieval(UniterPreTrainedModel):
    """model/itm.py:14-59 (ITM classifier of model/pretrain.py:159-176 shares ``itm_output``)."""

    def __init__(self, config, img_dim, margin=0.2):
        super().__init__(config)
        self.uniter = UniterModel(config, img_dim)
        self.itm_output = nn.Linear(config.hidden_size, 2)
        self.rank_output = nn.Linear(config.hidden_size, 1)
        self.margin = margin
        self.apply(self.init_weights)

    def init_output(self):
        """need to be called after from pretrained (model/itm.py:26-29)"""
        self.rank_output.weight.data = self.itm_output.weight.data[1:, :]
        self.rank_output.bias.data = self.itm_output.bias.data[1:]

    def pooled(self, batch):
        """pooler(sequence_output) (model/itm.py:36-42) without materialising the padded [B, L, H]
        tensor: the [CLS] rows are gathered straight from the packed encoder output."""
        batch = defaultdict(lambda: None, batch)
        packed, meta = self.uniter.encode_packed(
            batch["input_ids"], batch["position_ids"], batch["img_feat"], batch["img_pos_feat"],
            batch["attn_masks"], batch["gather_index"], output_all_encoded_layers=False)
        B, L = meta["n_batch"], meta["L"]
        cls_rows = meta["unpack_ext"][::L][:B].contiguous()      # packed row of position (b, 0)
        return self.uniter.pooler(gather_packed_rows(packed, cls_rows))

    def itm_scores(self, batch):
        """model/pretrain.py:163-164: itm_output(pooler(sequence_output))."""
        return LibLinear.apply(self.pooled(batch), self.itm_output.weight, self.itm_output.bias, False, False)

    def forward(self, batch, compute_loss=True):
        rank_scores = LibLinear.apply(self.pooled(batch), self.rank_output.weight, self.rank_output.bias,
                                      False, False)
        if compute_loss:
            scores = torch.sigmoid(rank_scores).contiguous().view(-1, batch["sample_size"])
            pos, neg = scores[:, :1], scores[:, 1:]
            return torch.clamp(self.margin + neg - pos, 0)
        return rank_scores


class UniterForImageTextRetrievalHardNeg(UniterForImageTextRetrieval):
    """model/itm.py:57-147 — in-batch hard-negative mining: score all N candidate pairs of one
    text (sample_from='t') or one image ('i') without gradients in eval mode, keep the positive
    (row 0) and the `hard_size` best-scoring negatives, and train on those."""

    def __init__(self, config, img_dim, margin=0.2, hard_size=16):
        super().__init__(config, img_dim, margin)
        self.hard_size = hard_size

    def forward(self, batch, sample_from="t", compute_loss=True):
        n_pairs = batch["attn_masks"].size(0)
        if sample_from == "t":                       # one text shared by every pair
            if batch["input_ids"].size(0) == 1:
                batch["input_ids"] = batch["input_ids"].expand(n_pairs, -1)
        elif sample_from == "i":                     # one image shared by every pair
            for key in ("img_feat", "img_pos_feat"):
                if batch[key].size(0) == 1:
                    batch[key] = batch[key].expand(n_pairs, -1, -1)
        else:
            raise ValueError()
        if not (self.training and compute_loss):
            return super().forward(batch, compute_loss)
        with torch.no_grad():
            self.eval()
            scores = super().forward(batch, compute_loss=False)
            hard_batch = self._get_hard_batch(batch, scores, sample_from)
            self.train()
        return super().forward(hard_batch, compute_loss=True)

    def _get_hard_batch(self, batch, scores, sample_from="t"):
        """Row selection of model/itm.py:92-147 (pure index logic, pinned bit-exactly against the
        reference in tests/test_heads_optim_cpu.py)."""
        batch = defaultdict(lambda: None, batch)
        k = self.hard_size
        neg = scores.squeeze(-1)[1:].topk(k, sorted=False)[1] + 1          # positive is row 0
        rows = torch.cat([neg.new_zeros(1), neg])
        masks = batch["attn_masks"].index_select(0, rows)
        gather = batch["gather_index"].index_select(0, rows)
        pos = batch["position_ids"]
        if pos.size(0) != 1:
            pos = pos[:k + 1]
        ids, feat, box = batch["input_ids"], batch["img_feat"], batch["img_pos_feat"]
        # host-known lengths of the candidate pairs (our collates provide them): the lengths of the
        # mined rows are then known after ONE small read of the top-k indices, and the train forward
        # packs without reading the device again (the reference syncs here too, model/itm.py:113)
        lens_all = None
        if batch["txt_lens"] is not None and batch["num_bbs"] is not None:
            lens_all = [a + b for a, b in zip(batch["txt_lens"], batch["num_bbs"])]
            rows_h = rows.tolist()
            lens_sel = [lens_all[r] for r in rows_h]
        if sample_from == "t":
            longest = max(lens_sel) if lens_all is not None else masks.sum(dim=1).max().item()   # cut to minimum padding
            n_img = longest - ids.size(1)
            masks, gather = masks[:, :longest], gather[:, :longest]
            feat = feat.index_select(0, rows)[:, :n_img, :]
            box = box.index_select(0, rows)[:, :n_img, :]
            ids = ids[:k + 1]
        elif sample_from == "i":
            ids = ids.index_select(0, rows)
            feat, box = feat[:k + 1], box[:k + 1]
        else:
            raise ValueError()
        if lens_all is not None:
            from .model import register_lengths
            masks = masks.contiguous()
            register_lengths(masks, lens_sel, prefix=True)
        return {"sample_size": k + 1, "input_ids": ids, "position_ids": pos, "img_feat": feat,
                "img_pos_feat": box, "attn_masks": masks, "gather_index": gather}
