"""Callers of the hot path used by the benchmark / smoke configs: the MLM pre-training head.

The reference's heads are out of scope to rewrite: they are kept as plain
torch modules here with the reference's parameter names so a `UniterForPretraining` checkpoint
loads (`uniter.*`, `cls.predictions.*`), and they consume the drop-in `UniterModel` exactly as
model/pretrain.py:107-133 does (text slice -> masked rows only -> transform -> tied decoder).
Kernels for this head are row (f)-1 of SURVEY.md §8 ("next").
"""
import torch
from torch import nn
from torch.nn import functional as F

from .model import UniterModel, UniterPreTrainedModel


def gelu(x):
    return F.gelu(x)  # erf form, == model/layer.py:31-37


class BertPredictionHeadTransform(nn.Module):
    """model/layer.py:188-203."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=1e-12)

    def forward(self, hidden_states):
        return self.LayerNorm(gelu(self.dense(hidden_states)))


class BertLMPredictionHead(nn.Module):
    """model/layer.py:206-222 — decoder weight tied to the word embeddings."""

    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(bert_model_embedding_weights.size(1),
                                 bert_model_embedding_weights.size(0), bias=False)
        self.decoder.weight = bert_model_embedding_weights
        self.bias = nn.Parameter(torch.zeros(bert_model_embedding_weights.size(0)))

    def forward(self, hidden_states):
        return self.decoder(self.transform(hidden_states)) + self.bias


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.predictions = BertLMPredictionHead(config, bert_model_embedding_weights)

    def forward(self, sequence_output):
        return self.predictions(sequence_output)


class UniterForMLM(UniterPreTrainedModel):
    """The MLM branch of UniterForPretraining (model/pretrain.py:50-60, 107-133)."""

    def __init__(self, config, img_dim):
        super().__init__(config)
        self.uniter = UniterModel(config, img_dim)
        self.cls = BertOnlyMLMHead(config, self.uniter.embeddings.word_embeddings.weight)
        self.apply(self.init_weights)

    def forward(self, batch, compute_loss=True):
        input_ids = batch["input_ids"]
        txt_labels = batch["txt_labels"]
        sequence_output = self.uniter(input_ids, batch["position_ids"], batch["img_feat"],
                                      batch["img_pos_feat"], batch["attn_masks"],
                                      batch["gather_index"], output_all_encoded_layers=False)
        if "mlm_index" in batch:
            # loader-provided flat positions of the masked tokens: static-shape gather, no sync
            H = sequence_output.size(-1)
            masked_output = sequence_output.reshape(-1, H).index_select(0, batch["mlm_index"])
            targets = batch["mlm_targets"]
        else:
            sequence_output = sequence_output[:, :input_ids.size(1), :]
            mask = (txt_labels != -1)
            masked_output = sequence_output[mask.unsqueeze(-1).expand_as(sequence_output)] \
                .contiguous().view(-1, sequence_output.size(-1))
            targets = txt_labels[mask]
        prediction_scores = self.cls(masked_output)
        if compute_loss:
            return F.cross_entropy(prediction_scores.float(), targets, reduction="none")
        return prediction_scores
