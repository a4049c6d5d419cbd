"""Golden vectors for GREEDY decoding of right-padded prompts (the VQA ``generate`` path with ``num_beams=1``,
prismer_vqa.py:45-57; BASELINE.json's "greedy captions/sec" decode mode): the UNMODIFIED reference decoder
(model/modules/roberta.py ``RobertaForCausalLMModified`` through ``oracle/reference_shim.py``) under ``transformers``
``generate(num_beams=1, do_sample=False, attention_mask=<prompt mask>)``.

What this pins (round-1 VERDICT weak #1a): with a right-padded prompt HF takes ``logits[:, -1]`` -- the PAD position of a
short row -- for the first generated token, position ids skip pads (roberta.py:38-45) and the generated tokens attend to
every non-pad prompt token.  A padded row therefore does NOT decode like the same row unpadded; the product must match THIS.

TEST INFRASTRUCTURE: needs /root/reference; run here, never on the GPU box.

    python oracle/gen_golden_greedy.py   ->  tests/golden/prismer_tiny_greedy.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_shim  # noqa: E402
from oracle.gen_golden import GOLD, TINY_DEC  # noqa: E402
from oracle.gen_golden_beam import decoder_state  # noqa: E402
from tests.helpers import GREEDY_CASES, beam_case_inputs  # noqa: E402


def main():
    ns = reference_shim.load()
    out = {}
    for c in GREEDY_CASES:
        dec = decoder_state(ns, c["boost"])
        ids, mask, enc = beam_case_inputs(c)
        with torch.no_grad():
            g = dec.generate(input_ids=ids, encoder_hidden_states=enc, attention_mask=mask, num_beams=1, do_sample=False,
                             max_length=c["T0"] + c["max_add"], min_length=c["T0"] + c["min_add"])
        out[c["name"] + ".ids"] = g.numpy()
        out[c["name"] + ".prompt"] = ids.numpy()
        out[c["name"] + ".mask"] = mask.numpy()
        print(c["name"], tuple(g.shape))
        print(g.numpy())
    np.savez_compressed(os.path.join(GOLD, "prismer_tiny_greedy.npz"), **out)


if __name__ == "__main__":
    main()
