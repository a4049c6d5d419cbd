"""Tokenizer plumbing.  ``RobertaTokenizer.from_pretrained`` (model/prismer.py:32) needs vocab files; there is no network
here, so when they are not available locally a deterministic stand-in with the same call surface is used (ids in,
ids out; <s>=0, <pad>=1, </s>=2).  It hashes whitespace-separated words into the RoBERTa id range -- enough to drive the
string API (`forward(experts, caption=[...])`, `.generate()` -> list[str]) on synthetic data."""
from __future__ import annotations

import zlib
from types import SimpleNamespace
from typing import List, Union

import torch


class BatchEncoding(SimpleNamespace):
    def to(self, device):
        return BatchEncoding(input_ids=self.input_ids.to(device), attention_mask=self.attention_mask.to(device))


class HashTokenizer:
    bos_token_id, pad_token_id, eos_token_id = 0, 1, 2

    def __init__(self, vocab_size: int = 50265):
        self.vocab_size = vocab_size

    def _word(self, w: str) -> int:
        if w == "<s>":
            return 0
        if w == "</s>":
            return 2
        return 3 + zlib.crc32(w.encode()) % (self.vocab_size - 3)

    def _encode(self, text: str, add_special_tokens=True) -> List[int]:
        text = text.replace("<s>", " <s> ").replace("</s>", " </s> ")
        ids = [self._word(w) for w in text.split()]
        return ([0] + ids + [2]) if add_special_tokens else ids

    def __call__(self, text: Union[str, List[str]], padding=False, truncation=False, max_length=None, return_tensors=None,
                 add_special_tokens=True):
        single = isinstance(text, str)
        seqs = [self._encode(t, add_special_tokens) for t in ([text] if single else text)]
        if truncation and max_length:
            seqs = [s[:max_length - 1] + [s[-1]] if len(s) > max_length else s for s in seqs]
        if return_tensors is None:
            if single:
                return SimpleNamespace(input_ids=seqs[0], attention_mask=[1] * len(seqs[0]))
            return SimpleNamespace(input_ids=seqs, attention_mask=[[1] * len(s) for s in seqs])
        L = max(len(s) for s in seqs)
        ids = torch.full((len(seqs), L), self.pad_token_id, dtype=torch.long)
        att = torch.zeros((len(seqs), L), dtype=torch.long)
        for i, s in enumerate(seqs):
            ids[i, :len(s)] = torch.tensor(s)
            att[i, :len(s)] = 1
        return BatchEncoding(input_ids=ids, attention_mask=att)

    def decode(self, ids, skip_special_tokens=True) -> str:
        ids = ids.tolist() if hasattr(ids, "tolist") else list(ids)
        return " ".join(f"tok{int(i)}" for i in ids if not (skip_special_tokens and int(i) in (0, 1, 2)))


def build_tokenizer(model_name: str, vocab_size: int):
    try:
        from transformers import RobertaTokenizer
        tok = RobertaTokenizer.from_pretrained(model_name, local_files_only=True)
        if tok.vocab_size < vocab_size - 16 or tok.pad_token_id != 1:   # transformers >= 5 returns an empty tokenizer
            raise RuntimeError("no local vocabulary")
        return tok
    except Exception:
        return HashTokenizer(vocab_size)
