"""Seeded synthetic batches and weights in the reference's batch-dict schema.

The reference's collate functions (data/vqa.py:44-71, data/mlm.py:96-136, data/itm.py:282-369)
emit padded tensors plus a `gather_index`; BASELINE configs use synthetic data of that shape.
Everything here is CPU torch with an explicit Generator, so the same call yields the same
batch in the golden generator, the tests, smoke() and bench.py.
"""
import hashlib

import torch


def get_gather_index(txt_lens, num_bbs, batch_size, max_len, out_size):
    """Compaction index with the semantics of data/data.py:271-279: slot j of sample i reads
    text row j for j < tl, image row (j - tl) (stored after the max_len text rows) for
    tl <= j < tl + nbb, and itself (a padding row) otherwise."""
    gi = torch.arange(out_size, dtype=torch.long).repeat(batch_size, 1)
    for i in range(batch_size):
        tl, nbb = int(txt_lens[i]), int(num_bbs[i])
        gi[i, tl:tl + nbb] = max_len + torch.arange(nbb, dtype=torch.long)
    return gi


def synth_batch(batch_size, tl_lo, tl_hi, nbb_lo, nbb_hi, seed, img_dim=2048, vocab_size=28996,
                txt_lens=None, num_bbs=None, mlm_prob=0.0):
    """One padded batch dict.  Lengths ~ U{lo..hi} unless given explicitly."""
    g = torch.Generator().manual_seed(seed)
    if txt_lens is None:
        txt_lens = torch.randint(tl_lo, tl_hi + 1, (batch_size,), generator=g).tolist()
    if num_bbs is None:
        num_bbs = torch.randint(nbb_lo, nbb_hi + 1, (batch_size,), generator=g).tolist()
    Lt, Li = max(txt_lens), max(num_bbs)
    L = max(t + n for t, n in zip(txt_lens, num_bbs))
    input_ids = torch.zeros(batch_size, Lt, dtype=torch.long)
    img_feat = torch.zeros(batch_size, Li, img_dim)
    img_pos_feat = torch.zeros(batch_size, Li, 7)
    attn_masks = torch.zeros(batch_size, L, dtype=torch.long)
    txt_labels = torch.full((batch_size, Lt), -1, dtype=torch.long)
    lo_id = min(1000, vocab_size - 1)
    for i, (tl, nbb) in enumerate(zip(txt_lens, num_bbs)):
        ids = torch.randint(lo_id, vocab_size, (tl,), generator=g)
        ids[0] = 101 % vocab_size    # [CLS]
        ids[-1] = 102 % vocab_size   # [SEP]
        input_ids[i, :tl] = ids
        img_feat[i, :nbb] = torch.randn(nbb, img_dim, generator=g)
        xy = torch.rand(nbb, 4, generator=g)
        x1 = torch.minimum(xy[:, 0], xy[:, 2]); x2 = torch.maximum(xy[:, 0], xy[:, 2])
        y1 = torch.minimum(xy[:, 1], xy[:, 3]); y2 = torch.maximum(xy[:, 1], xy[:, 3])
        w, h = x2 - x1, y2 - y1
        img_pos_feat[i, :nbb] = torch.stack([x1, y1, x2, y2, w, h, w * h], 1)
        attn_masks[i, :tl + nbb] = 1
        if mlm_prob > 0 and tl > 2:
            m = torch.rand(tl, generator=g) < mlm_prob
            m[0] = False; m[-1] = False
            if not m.any():
                m[1] = True
            txt_labels[i, :tl][m] = ids[m]
    batch = {
        "input_ids": input_ids,
        "position_ids": torch.arange(Lt, dtype=torch.long).unsqueeze(0),
        "img_feat": img_feat,
        "img_pos_feat": img_pos_feat,
        "attn_masks": attn_masks,
        "gather_index": get_gather_index(txt_lens, num_bbs, batch_size, Lt, L),
        "txt_lens": txt_lens,
        "num_bbs": num_bbs,
    }
    if mlm_prob > 0:
        batch["txt_labels"] = txt_labels
        # host-side compaction of the masked positions (flat index b * L + j into the encoder's
        # [B, L] output): lets the MLM head gather its rows with a static shape instead of the
        # boolean-mask indexing of model/pretrain.py:129-133, which costs a device sync per step
        pos = (txt_labels != -1).nonzero(as_tuple=False)
        batch["mlm_index"] = (pos[:, 0] * L + pos[:, 1]).contiguous()
        batch["mlm_targets"] = txt_labels[pos[:, 0], pos[:, 1]].contiguous()
    return batch


def seeded_state(shapes, seed=0, perturb=True):
    """Deterministic fp32 weights for a {key: shape} schema, independent of module construction
    order: each tensor is drawn from its own Generator seeded by sha1(key) ^ seed.
    Linear / Embedding weights ~ N(0, 0.02) (model/model.py:136-141); with `perturb`, biases
    ~ N(0, 0.02) and LayerNorm weights ~ 1 + N(0, 0.1) so that bias / affine bugs cannot hide
    behind the reference's zero / one initialisation."""
    state = {}
    for key in sorted(shapes):
        h = int(hashlib.sha1(key.encode()).hexdigest()[:8], 16)
        g = torch.Generator().manual_seed((h ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        shape = tuple(shapes[key])
        is_ln = "LayerNorm" in key or "layer_norm" in key
        if key.endswith("bias"):
            t = torch.randn(shape, generator=g) * 0.02 if perturb else torch.zeros(shape)
        elif is_ln:
            t = 1.0 + (torch.randn(shape, generator=g) * 0.1 if perturb else torch.zeros(shape))
        else:
            t = torch.randn(shape, generator=g) * 0.02
        state[key] = t
    return state


def uniter_state_shapes(hidden, layers, inter, vocab, max_pos, type_vocab, img_dim):
    """State-dict schema of the reference UniterModel (verified against
    model/model.py:217-304 and model/layer.py:53-185 by tests/golden/make_goldens.py)."""
    H, I = hidden, inter
    s = {
        "embeddings.word_embeddings.weight": (vocab, H),
        "embeddings.position_embeddings.weight": (max_pos, H),
        "embeddings.token_type_embeddings.weight": (type_vocab, H),
        "embeddings.LayerNorm.weight": (H,), "embeddings.LayerNorm.bias": (H,),
        "img_embeddings.img_linear.weight": (H, img_dim), "img_embeddings.img_linear.bias": (H,),
        "img_embeddings.img_layer_norm.weight": (H,), "img_embeddings.img_layer_norm.bias": (H,),
        "img_embeddings.pos_layer_norm.weight": (H,), "img_embeddings.pos_layer_norm.bias": (H,),
        "img_embeddings.pos_linear.weight": (H, 7), "img_embeddings.pos_linear.bias": (H,),
        "img_embeddings.mask_embedding.weight": (2, img_dim),
        "img_embeddings.LayerNorm.weight": (H,), "img_embeddings.LayerNorm.bias": (H,),
        "pooler.dense.weight": (H, H), "pooler.dense.bias": (H,),
    }
    for i in range(layers):
        p = "encoder.layer.%d." % i
        for n in ("query", "key", "value"):
            s[p + "attention.self.%s.weight" % n] = (H, H)
            s[p + "attention.self.%s.bias" % n] = (H,)
        s[p + "attention.output.dense.weight"] = (H, H)
        s[p + "attention.output.dense.bias"] = (H,)
        s[p + "attention.output.LayerNorm.weight"] = (H,)
        s[p + "attention.output.LayerNorm.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (I, H)
        s[p + "intermediate.dense.bias"] = (I,)
        s[p + "output.dense.weight"] = (H, I)
        s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,)
        s[p + "output.LayerNorm.bias"] = (H,)
    return s


def pad_mlm_index(batch, multiple=64):
    """Fixed-size masked-token lists for CUDA-graph replay: `mlm_index` is padded to a multiple of
    `multiple` with B * L ("no row": UniterForMLM maps it to a zero row) and `mlm_targets` with -1
    (ignored by the cross-entropy: loss 0, no gradient); `mlm_inv_n` = 1 / number of real masked
    tokens, so that `loss.sum() * mlm_inv_n` is the reference's `loss.mean()`
    (model/pretrain.py:122-127 + pretrain.py:297)."""
    idx, tgt = batch["mlm_index"], batch["mlm_targets"]
    n = idx.numel()
    n_pad = max((n + multiple - 1) // multiple * multiple, multiple)
    B, L = batch["attn_masks"].shape
    out = dict(batch)
    out["mlm_index"] = torch.cat([idx, torch.full((n_pad - n,), B * L, dtype=idx.dtype)])
    out["mlm_targets"] = torch.cat([tgt, torch.full((n_pad - n,), -1, dtype=tgt.dtype)])
    out["mlm_inv_n"] = torch.tensor([1.0 / max(n, 1)], dtype=torch.float32)
    return out


def synth_mrm(batch, mask_prob=0.15, label_dim=1601, seed=0, pad_multiple=0):
    """MRM inputs on top of a synth_batch, with the semantics of data/mrm.py: per-region Bernoulli
    mask with at least one masked region per sample (:14-20), `img_mask_tgt` over the joint [B, L]
    layout (:23-26), `feat_targets` = the original features of the masked regions (:29-34), the
    masked regions' input features zeroed (:37-40), and soft labels `label_targets` [n, label_dim]
    (data/mrm.py:157-163: the detector's class distribution; here softmax of N(0, 1)).
    `mrm_index` = flat positions b * L + j of the masked regions (host-side compaction: what
    `_compute_masked_hidden`, model/pretrain.py:129-133, extracts with a device-side boolean mask);
    with `pad_multiple` it is padded with B * L and the targets with zero rows, and `mrm_valid` [n]
    marks the real rows."""
    g = torch.Generator().manual_seed(seed)
    img_feat = batch["img_feat"]
    B, Li, D = img_feat.shape
    L = batch["attn_masks"].size(1)
    img_masks = torch.zeros(B, Li, dtype=torch.bool)
    img_mask_tgt = torch.zeros(B, L, dtype=torch.bool)
    for i, (tl, nbb) in enumerate(zip(batch["txt_lens"], batch["num_bbs"])):
        m = torch.rand(nbb, generator=g) < mask_prob
        if not m.any():
            m[int(torch.randint(0, nbb, (1,), generator=g))] = True
        img_masks[i, :nbb] = m
        img_mask_tgt[i, tl:tl + nbb] = m
    feat_targets = img_feat[img_masks].contiguous()
    n = feat_targets.size(0)
    out = dict(batch)
    out["img_feat"] = img_feat.masked_fill(img_masks.unsqueeze(-1), 0)
    out["img_masks"] = img_masks
    out["img_mask_tgt"] = img_mask_tgt
    out["feat_targets"] = feat_targets
    out["label_targets"] = torch.softmax(torch.randn(n, label_dim, generator=g), dim=-1)
    pos = img_mask_tgt.nonzero(as_tuple=False)
    idx = (pos[:, 0] * L + pos[:, 1]).contiguous()
    valid = torch.ones(n, dtype=torch.float32)
    if pad_multiple:
        n_pad = max((n + pad_multiple - 1) // pad_multiple * pad_multiple, pad_multiple)
        idx = torch.cat([idx, torch.full((n_pad - n,), B * L, dtype=idx.dtype)])
        out["feat_targets"] = torch.cat([feat_targets, torch.zeros(n_pad - n, D)])
        out["label_targets"] = torch.cat([out["label_targets"],
                                          torch.full((n_pad - n, label_dim), 1.0 / label_dim)])
        valid = torch.cat([valid, torch.zeros(n_pad - n)])
    out["mrm_index"] = idx
    out["mrm_valid"] = valid
    out["mrm_inv_n"] = torch.tensor([1.0 / max(n, 1)], dtype=torch.float32)
    return out
