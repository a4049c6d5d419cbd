"""``PrismerCaption`` -- the call surface of ``model/prismer_caption.py:14-112`` on the sm_100a engine."""
import torch

from . import engine, text
from .prismer import Prismer


class PrismerCaption(Prismer):
    def forward(self, experts, caption=None, answer=None, train=True, prefix="", inference="generate", k_test=32,
                input_ids=None, attention_mask=None, prompt_length=None, num_beams=3):
        """Reference signature; additionally accepts pre-tokenised ``input_ids/attention_mask`` (+ ``prompt_length``) so
        the data loader can own tokenisation."""
        device = experts["rgb"].device
        if train:
            if input_ids is None:                                                             # strings: tokenise here, as the reference
                input_ids, attention_mask, labels, _ = text.caption_inputs(self.tokenizer, caption, prefix)
                input_ids, attention_mask, labels = input_ids.to(device), attention_mask.to(device), labels.to(device)
            else:                                                                             # tensors from the data loader (N3)
                labels = input_ids.masked_fill(input_ids == self.tokenizer.pad_token_id, -100)    # prismer_caption.py:22
                if prompt_length:
                    labels[:, :prompt_length] = -100                                          # prismer_caption.py:24-26
            return engine.train_loss(self, experts, input_ids, attention_mask, labels)

        if inference == "generate":
            if input_ids is None:
                tok = self.tokenizer([prefix] * experts["rgb"].size(0), padding="longest", return_tensors="pt").to(device)
                input_ids, attention_mask = tok.input_ids[:, :-1], tok.attention_mask[:, :-1]   # drop </s>
            enc = self.expert_encoder(experts).transpose(0, 1)                                  # 'l b d -> b l d'
            outputs = self.text_decoder.generate(input_ids=input_ids.contiguous(), encoder_hidden_states=enc,
                                                 attention_mask=attention_mask, num_beams=num_beams, max_length=20, min_length=8)
            if caption is None and prefix is None:
                return outputs
            captions = []
            for output in outputs:
                decoded = self.tokenizer.decode(output, skip_special_tokens=True)
                space_idx = 1 if len(prefix) > 0 else 0
                captions.append(decoded[len(prefix) + space_idx:])
            return captions

        if inference == "rank":
            answer_tok = self.tokenizer([" " + a.lower() + "</s>" for a in answer], padding="longest", return_tensors="pt",
                                        add_special_tokens=False).to(device)
            ptok = self.tokenizer([prefix] * experts["rgb"].size(0), padding="longest", return_tensors="pt").to(device)
            return rank(self, experts, ptok.input_ids[:, :-1], ptok.attention_mask[:, :-1], answer_tok.input_ids,
                        answer_tok.attention_mask, k_test)
        raise ValueError(inference)


def tile(x, dim, n_tile):
    """prismer_caption.py:115-121: repeat every entry n_tile times along ``dim`` (== repeat_interleave)."""
    return x.repeat_interleave(n_tile, dim=dim)


@torch.no_grad()
def rank(model, experts, start_ids, start_mask, answer_ids, answer_mask, k_test):
    """inference == 'rank' (prismer_caption.py:59-112 / prismer_vqa.py:64-113).  Integer selection steps (index_select,
    top-k, argmax) run on fp32 device values; returns LongTensor[B] of answer indices."""
    from .modules.roberta import CausalLMOutput  # noqa: F401
    pad = model.tokenizer.pad_token_id
    enc = model.expert_encoder(experts).transpose(0, 1)
    start_output = model.text_decoder(start_ids.contiguous(), attention_mask=start_mask, encoder_hidden_states=enc)
    logits = start_output.logits[:, -1, :]
    prob_first = torch.softmax(logits, dim=1).index_select(dim=1, index=answer_ids[:, 0])
    _, topk_ids = prob_first.topk(k_test, dim=1)
    a_ids = torch.cat([answer_ids.index_select(0, t) for t in topk_ids], dim=0)
    a_att = torch.cat([answer_mask.index_select(0, t) for t in topk_ids], dim=0)
    input_ids = torch.cat([tile(start_ids, 0, k_test), a_ids], dim=1).long()
    attention_masks = torch.cat([tile(start_mask, 0, k_test), a_att], dim=1)
    targets = input_ids.masked_fill(input_ids == pad, -100)
    targets[:, :-answer_ids.shape[1]] = -100
    # the reference tiles the encoder states k_test times (prismer_caption.py:94-96); here the k candidates of an image share its visual
    # K/V (projected once per image) through the attention kernel's kv_div: same arithmetic, 1/k of the K/V projection work and memory
    out = model.text_decoder(input_ids, attention_mask=attention_masks, encoder_hidden_states=enc, labels=targets, encoder_repeat=k_test)
    log_probs_sum = (-out.loss / torch.sum(targets != -100, dim=-1)).view(-1, k_test)
    max_topk_ids = log_probs_sum.argmax(dim=1)
    return topk_ids[max_topk_ids >= 0, max_topk_ids]
