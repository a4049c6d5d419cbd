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


def _linear(n_in: int, n_out: int) -> nn.Linear:
    return nn.Linear(n_in, n_out)


def _norm(cfg) -> LayerNorm:
    return LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


def _container(**children) -> nn.Module:
    """A bare nn.Module whose only job is to hold named children (keeps the reference's state_dict keys)."""
    m = nn.Module()
    for name, child in children.items():
        setattr(m, name, child)
    return m


class RobertaEmbeddings(nn.Module):
    """word / position / token-type tables + LayerNorm (keys of roberta.py:48-64); gathered by ``prismer_embed_fwd``."""

    def __init__(self, cfg):
        super().__init__()
        H, pad = cfg.hidden_size, cfg.pad_token_id
        self.word_embeddings = nn.Embedding(cfg.vocab_size, H, padding_idx=pad)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, H, padding_idx=pad)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab_size, H)
        self.LayerNorm = _norm(cfg)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(cfg.max_position_embeddings).expand((1, -1)))
        self.padding_idx = pad


class RobertaSelfAttention(nn.Module):
    """q/k/v projections; K/V read ``vision_hidden_size`` features when this is the cross-attention (roberta.py:79-93)."""

    def __init__(self, cfg, is_cross_attention=False):
        super().__init__()
        H = cfg.hidden_size
        src = cfg.vision_hidden_size if is_cross_attention else H
        self.num_attention_heads = cfg.num_attention_heads
        self.query, self.key, self.value = _linear(H, H), _linear(src, H), _linear(src, H)
        self.dropout = nn.Dropout(cfg.attention_probs_dropout_prob)


class RobertaSelfOutput(nn.Module):
    def __init__(self, cfg, n_in=None):
        super().__init__()
        self.dense = _linear(n_in or cfg.hidden_size, cfg.hidden_size)
        self.LayerNorm = _norm(cfg)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)


class RobertaAttention(nn.Module):
    def __init__(self, cfg, is_cross_attention=False):
        super().__init__()
        setattr(self, "self", RobertaSelfAttention(cfg, is_cross_attention))
        self.output = RobertaSelfOutput(cfg)


class RobertaIntermediate(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = _linear(cfg.hidden_size, cfg.intermediate_size)


class RobertaOutput(RobertaSelfOutput):
    def __init__(self, cfg):
        super().__init__(cfg, n_in=cfg.intermediate_size)


class RobertaLayer(nn.Module):
    """self-attention block + MLP (post-LN); the engine runs them as ``mode='attention'`` / ``mode='mlp'`` of roberta.py:186-199."""

    def __init__(self, cfg):
        super().__init__()
        self.attention, self.intermediate, self.output = RobertaAttention(cfg), RobertaIntermediate(cfg), RobertaOutput(cfg)


class RobertaEncoder(nn.Module):
    """num_hidden_layers x [RobertaLayer, cross RobertaAttention, norm-late Adaptor] + a cross-attention-free output_layer."""

    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        triple = lambda: nn.ModuleList([RobertaLayer(cfg), RobertaAttention(cfg, True), Adaptor(cfg.hidden_size, norm_late=True)])
        self.layer = nn.ModuleList([triple() for _ in range(cfg.num_hidden_layers)])
        self.output_layer = RobertaLayer(cfg)


class RobertaModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.config, self.embeddings, self.encoder = cfg, RobertaEmbeddings(cfg), RobertaEncoder(cfg)


class RobertaLMHead(nn.Module):
    """dense -> erf-GELU -> LayerNorm -> tied decoder (+ bias); decoder.bias aliases ``bias`` (roberta.py:417-419)."""

    def __init__(self, cfg):
        super().__init__()
        self.dense, self.layer_norm = _linear(cfg.hidden_size, cfg.hidden_size), _norm(cfg)
        self.decoder = _linear(cfg.hidden_size, cfg.vocab_size)
        self.bias = nn.Parameter(torch.zeros(cfg.vocab_size))
        self.decoder.bias = self.bias


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
                return_dict=True, encoder_repeat: int = 1, **_):
        """``encoder_repeat`` = k (extension, inference only): ``input_ids`` carries k consecutive rows per row of
        ``encoder_hidden_states`` -- what the reference obtains by ``tile``-ing the encoder states k times (prismer_caption.py:94-96)."""
        from .. import engine
        return engine.decoder_apply(self, input_ids, attention_mask, encoder_hidden_states, labels, weights, enc_repeat=encoder_repeat)

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
