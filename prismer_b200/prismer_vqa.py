"""``PrismerVQA`` -- the call surface of ``model/prismer_vqa.py:15-113`` on the sm_100a engine."""
import torch

from . import engine
from .prismer import Prismer
from .prismer_caption import rank


class PrismerVQA(Prismer):
    def forward(self, experts, question, answer=None, weights=None, train=True, inference="rank", k_test=128):
        device = experts["rgb"].device
        question = ["<s>" + q.capitalize() for q in question]
        q = self.tokenizer(question, padding="longest", truncation=True, max_length=35, add_special_tokens=False,
                           return_tensors="pt").to(device)
        if train:
            a = self.tokenizer([" " + x.capitalize() + "</s>" for x in answer], padding="longest", return_tensors="pt",
                               add_special_tokens=False).to(device)
            input_ids = torch.cat([q.input_ids, a.input_ids], dim=1).long()
            attention_mask = torch.cat([q.attention_mask, a.attention_mask], dim=1)
            targets = input_ids.masked_fill(input_ids == self.tokenizer.pad_token_id, -100)
            targets[:, :-a.input_ids.shape[1]] = -100                                   # prismer_vqa.py:32-33
            w = weights.to(device=device, dtype=torch.float32).contiguous() if weights is not None else None
            return engine.train_loss(self, experts, input_ids, attention_mask, targets, w)   # mean(weights * loss)
        if inference == "generate":
            enc = self.expert_encoder(experts).transpose(0, 1)
            outputs = self.text_decoder.generate(input_ids=q.input_ids, encoder_hidden_states=enc, attention_mask=q.attention_mask,
                                                 max_length=q.input_ids.shape[1] + 10, min_length=q.input_ids.shape[1] + 2,
                                                 num_beams=3, length_penalty=-1)
            return [self.tokenizer.decode(outputs[i, q.input_ids.shape[1]:], skip_special_tokens=True).lower().strip()
                    for i in range(len(outputs))]
        if inference == "rank":
            a = self.tokenizer([" " + x.capitalize() + "</s>" for x in answer], padding="longest", return_tensors="pt",
                               add_special_tokens=False).to(device)
            return rank(self, experts, q.input_ids, q.attention_mask, a.input_ids, a.attention_mask, k_test)
        raise ValueError(inference)
