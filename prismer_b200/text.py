"""Tokenisation + label masking of the task wrappers as pure host functions: the reference does this
inside ``forward`` on every step (prismer_caption.py:21-26, prismer_vqa.py:18-33); here the model calls the same functions, and a
data loader may call them in its workers and hand ``input_ids / attention_mask / labels`` tensors to ``forward`` instead of strings."""
from __future__ import annotations

from typing import List, Tuple

import torch


def caption_inputs(tokenizer, caption: List[str], prefix: str = "") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]:
    """prismer_caption.py:21-26 -> (input_ids, attention_mask, labels, prompt_length): pads and the prefix tokens are not scored."""
    tok = tokenizer(caption, padding="longest", truncation=True, max_length=30, return_tensors="pt")
    prompt_length = len(tokenizer(prefix).input_ids) - 1 if len(prefix) > 0 else 0          # drop </s>
    labels = tok.input_ids.masked_fill(tok.input_ids == tokenizer.pad_token_id, -100)
    if prompt_length:
        labels[:, :prompt_length] = -100
    return tok.input_ids, tok.attention_mask, labels, prompt_length


def vqa_question(tokenizer, question: List[str]):
    """prismer_vqa.py:18-20: ``<s>Question`` (capitalised), at most 35 tokens, right-padded, no </s>."""
    return tokenizer(["<s>" + q.capitalize() for q in question], padding="longest", truncation=True, max_length=35,
                     add_special_tokens=False, return_tensors="pt")


def vqa_answers(tokenizer, answer: List[str]):
    """prismer_vqa.py:26-27 / :69-70: `` Answer</s>`` (capitalised), right-padded."""
    return tokenizer([" " + a.capitalize() + "</s>" for a in answer], padding="longest", return_tensors="pt", add_special_tokens=False)


def vqa_inputs(tokenizer, question: List[str], answer: List[str]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """prismer_vqa.py:18-33 -> (input_ids, attention_mask, labels): question ++ answer, only the answer tokens are scored."""
    q, a = vqa_question(tokenizer, question), vqa_answers(tokenizer, answer)
    input_ids = torch.cat([q.input_ids, a.input_ids], dim=1).long()
    attention_mask = torch.cat([q.attention_mask, a.attention_mask], dim=1)
    labels = input_ids.masked_fill(input_ids == tokenizer.pad_token_id, -100)
    labels[:, :-a.input_ids.shape[1]] = -100
    return input_ids, attention_mask, labels
