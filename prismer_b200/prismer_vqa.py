"""``PrismerVQA`` -- the call surface of ``model/prismer_vqa.py:15-113`` on the sm_100a engine."""
import torch

from . import engine, text
from .prismer import Prismer
from .prismer_caption import rank


class PrismerVQA(Prismer):
    def forward(self, experts, question=None, answer=None, weights=None, train=True, inference="rank", k_test=128,
                input_ids=None, attention_mask=None, labels=None):
        """Reference signature; ``train=True`` additionally accepts pre-tokenised ``input_ids / attention_mask / labels``
        (``text.vqa_inputs`` run by the data loader) instead of the strings."""
        device = experts["rgb"].device
        if train:
            if input_ids is None:
                input_ids, attention_mask, labels = text.vqa_inputs(self.tokenizer, question, answer)    # prismer_vqa.py:18-33
            input_ids, attention_mask, labels = input_ids.to(device), attention_mask.to(device), labels.to(device)
            w = weights.to(device=device, dtype=torch.float32).contiguous() if weights is not None else None
            return engine.train_loss(self, experts, input_ids, attention_mask, labels, w)               # mean(weights * loss)
        q = text.vqa_question(self.tokenizer, question).to(device)
        if inference == "generate":
            enc = self.expert_encoder(experts).transpose(0, 1)
            outputs = self.text_decoder.generate(input_ids=q.input_ids, encoder_hidden_states=enc, attention_mask=q.attention_mask,
                                                 max_length=q.input_ids.shape[1] + 10, min_length=q.input_ids.shape[1] + 2,
                                                 num_beams=3, length_penalty=-1)
            return [self.tokenizer.decode(outputs[i, q.input_ids.shape[1]:], skip_special_tokens=True).lower().strip()
                    for i in range(len(outputs))]
        if inference == "rank":
            a = text.vqa_answers(self.tokenizer, answer).to(device)
            return rank(self, experts, q.input_ids, q.attention_mask, a.input_ids, a.attention_mask, k_test)
        raise ValueError(inference)
