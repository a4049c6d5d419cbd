"""Golden vectors for beam search: run the UNMODIFIED reference decoder
(model/modules/roberta.py ``RobertaForCausalLMModified`` through ``oracle/reference_shim.py``) under
``transformers`` ``generate(num_beams=3, ...)`` exactly as the reference drives it

  * caption: prismer_caption.py:42-50   (num_beams=3, max_length=20, min_length=8)
  * VQA    : prismer_vqa.py:45-57       (num_beams=3, max_length=T0+10, min_length=T0+2, length_penalty=-1,
                                         right-padded questions + attention_mask)

on a seeded tiny decoder and seeded encoder states, and commit ids + sequence scores to
``tests/golden/prismer_tiny_beam.npz``.  With random weights eos never wins inside 20 tokens, so each case adds a
constant to the LM-head eos bias (``eos_boost``) -- that makes hypotheses finish at different lengths and exercises
the finished-beam pool, the MinLength processor, the length penalty and the early-stop heuristic.

TEST INFRASTRUCTURE: needs /root/reference; run here, never on the GPU box.  The beam-search algorithm itself is
third-party (``transformers``; the reference pins ~=4.26.1, this container has 5.5.0 -- SURVEY.md section 8c): these
vectors pin the 5.5.0 behaviour, the only one that can run here.

    python oracle/gen_golden_beam.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_shim  # noqa: E402
from oracle.gen_golden import GOLD, TINY_DEC  # noqa: E402
from prismer_b200 import synthetic  # noqa: E402

from tests.helpers import BEAM_CASES as CASES, beam_case_inputs as case_inputs, beam_decoder_state  # noqa: E402


def decoder_state(ns, boost):
    dec = ns.build_decoder(TINY_DEC)
    dec.load_state_dict(beam_decoder_state(dec.state_dict(), boost))
    dec.lm_head.decoder.bias = dec.lm_head.bias          # keep the tie the reference sets up (roberta.py:417-419)
    return dec.eval()


def main():
    ns = reference_shim.load()
    out = {}
    for c in CASES:
        dec = decoder_state(ns, c["boost"])
        ids, mask, enc = case_inputs(c)
        with torch.no_grad():
            g = dec.generate(input_ids=ids, encoder_hidden_states=enc, attention_mask=mask, num_beams=c["nb"],
                             max_length=c["T0"] + c["max_add"], min_length=c["T0"] + c["min_add"], length_penalty=c["lp"],
                             return_dict_in_generate=True, output_scores=True)
        out[c["name"] + ".ids"] = g.sequences.numpy()
        out[c["name"] + ".scores"] = g.sequences_scores.numpy()
        out[c["name"] + ".prompt"] = ids.numpy()
        out[c["name"] + ".mask"] = mask.numpy()
        print(c["name"], g.sequences.shape, g.sequences_scores.numpy().round(4))
        print(g.sequences.numpy())
    np.savez_compressed(os.path.join(GOLD, "prismer_tiny_beam.npz"), **out)


if __name__ == "__main__":
    main()
