"""Text decoder -- parameter containers mirroring ``model/modules/roberta.py`` (post-LN RoBERTa turned into a causal
decoder with vision cross-attention and adaptors).  state_dict keys / shapes are identical to the reference's
``RobertaForCausalLMModified``; forward / backward / generate run in ``prismer_b200.engine``.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn

from .utils import Adaptor, LayerNorm


class RobertaConfig(SimpleNamespace):
    """Plain-attribute stand-in for ``transformers.RobertaConfig`` (only the fields the path reads)."""

    @classmethod
    def from_dict(cls, d):
        d = dict(d)
        d.setdefault("vision_hidden_size", d["hidden_size"])
        d.setdefault("layer_norm_eps", 1e-5)
        return cls(**d)


class RobertaEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size,
                                                padding_idx=config.pad_token_id)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.padding_idx = config.pad_token_id


class RobertaSelfAttention(nn.Module):
    def __init__(self, config, is_cross_attention=False):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        kv_in = config.vision_hidden_size if is_cross_attention else config.hidden_size
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(kv_in, config.hidden_size)
        self.value = nn.Linear(kv_in, config.hidden_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class RobertaSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class RobertaAttention(nn.Module):
    def __init__(self, config, is_cross_attention=False):
        super().__init__()
        self.self = RobertaSelfAttention(config, is_cross_attention)
        self.output = RobertaSelfOutput(config)


class RobertaIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class RobertaOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class RobertaLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = RobertaAttention(config)
        self.intermediate = RobertaIntermediate(config)
        self.output = RobertaOutput(config)


class RobertaEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([nn.ModuleList([RobertaLayer(config), RobertaAttention(config, is_cross_attention=True),
                                                   Adaptor(config.hidden_size, norm_late=True)])
                                    for _ in range(config.num_hidden_layers)])
        self.output_layer = RobertaLayer(config)


class RobertaModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = RobertaEmbeddings(config)
        self.encoder = RobertaEncoder(config)


class RobertaLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.layer_norm = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.decoder.bias = self.bias  # roberta.py:417-419


class CausalLMOutput(SimpleNamespace):
    pass


class RobertaForCausalLMModified(nn.Module):
    """Same call surface as ``roberta.py:337-406``: ``forward(input_ids, attention_mask, encoder_hidden_states, labels)``
    -> object with ``.loss`` ([B] per-sample sums) and ``.logits``; ``generate(...)`` (greedy / beam, no sampling)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.roberta = RobertaModel(config)
        self.lm_head = RobertaLMHead(config)
        self._init_weights()
        # weight tying (roberta.py:352-356): one Parameter object under both names
        self.lm_head.decoder.weight = self.roberta.embeddings.word_embeddings.weight

    def _init_weights(self):
        """roberta.py:248-260"""
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(0.0, std)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.Embedding):
                m.weight.data.normal_(0.0, std)
                if m.padding_idx is not None:
                    m.weight.data[m.padding_idx].zero_()
            elif isinstance(m, nn.LayerNorm):
                m.bias.data.zero_()
                m.weight.data.fill_(1.0)

    def get_output_embeddings(self):
        return self.lm_head.decoder

    def forward(self, input_ids=None, attention_mask=None, encoder_hidden_states=None, labels=None, weights=None,
                return_dict=True, **_):
        from .. import engine
        return engine.decoder_apply(self, input_ids, attention_mask, encoder_hidden_states, labels, weights)

    def prepare_inputs_for_generation(self, input_ids, attention_mask=None, encoder_hidden_states=None, **kw):
        if attention_mask is None:
            attention_mask = input_ids.new_ones(input_ids.shape)
        return {"input_ids": input_ids, "attention_mask": attention_mask, "encoder_hidden_states": encoder_hidden_states}

    @torch.no_grad()
    def generate(self, input_ids=None, encoder_hidden_states=None, attention_mask=None, num_beams=1, max_length=20,
                 min_length=0, length_penalty=1.0, do_sample=False, **_):
        from .. import generation
        if do_sample:
            raise NotImplementedError("sampling is not part of the reference path")
        if num_beams == 1:
            return generation.greedy(self, input_ids, encoder_hidden_states, attention_mask, max_length, min_length)
        return generation.beam_search(self, input_ids, encoder_hidden_states, attention_mask, num_beams, max_length,
                                      min_length, length_penalty)


ROBERTA_PRETRAINED_MODEL_ARCHIVE_LIST = ["roberta-base", "roberta-large"]


def load_decoder(name: str, config: RobertaConfig) -> RobertaForCausalLMModified:
    """Same signature as ``roberta.py:433``; no network: random init, pretrained weights via ``load_state_dict``."""
    return RobertaForCausalLMModified(config)
