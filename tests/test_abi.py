"""CPU: the C-ABI library loads and exports every symbol include/prismer_sm100.h declares (no compute without a GPU),
and the ctypes signatures agree with the header's arity."""
import os
import re

from prismer_b200 import _C, _C_decl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _C.lib()
    hdr = open(os.path.join(ROOT, "include", "prismer_sm100.h")).read()
    decls = re.findall(r"\nint (prismer_\w+)\(([^;]*?)\);", hdr, flags=re.S)
    assert len(decls) >= 30
    for name, args in decls:
        assert hasattr(lib, name), f"{name} not exported"
        n = len([a for a in args.split(",") if a.strip() and a.strip() != "void"])
        if name in _C_decl.SIGNATURES:
            assert n == len(_C_decl.SIGNATURES[name]), (name, n, len(_C_decl.SIGNATURES[name]))
    assert lib.prismer_abi_version() == 1


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from prismer_b200 import ops
    with pytest.raises(_C.PrismerError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_state_dict_layout_is_the_reference_layout():
    """Key set / shapes of the BASE model equal SURVEY.md section 8b (spot checks) and the tied LM head is one tensor."""
    from prismer_b200 import modeling
    from tests.helpers import TINY_DEC
    sd = modeling.template_state_dict(256, 2, 16, 64, ["depth", "seg_coco", "obj_detection"], TINY_DEC)
    for k in ["expert_encoder.positional_embedding", "expert_encoder.instance_embedding", "expert_encoder.conv1.rgb.weight",
              "expert_encoder.conv1.depth.1.weight", "expert_encoder.conv1.seg.11.running_var", "expert_encoder.conv1.obj_detection.13.weight",
              "expert_encoder.transformer.resblocks.1.0.attn.in_proj_weight", "expert_encoder.transformer.resblocks.0.1.adaptor.down_proj.bias",
              "expert_encoder.resampler.latents", "expert_encoder.resampler.perceiver_blocks.3.ln_ff.weight", "expert_encoder.ln_post.bias",
              "text_decoder.roberta.embeddings.position_ids", "text_decoder.roberta.encoder.layer.1.1.self.key.weight",
              "text_decoder.roberta.encoder.layer.0.2.adaptor_ln.weight", "text_decoder.roberta.encoder.output_layer.output.LayerNorm.bias",
              "text_decoder.lm_head.decoder.weight", "text_decoder.lm_head.decoder.bias", "text_decoder.lm_head.bias"]:
        assert k in sd, k
    dec = modeling.build_decoder(TINY_DEC)
    assert dec.lm_head.decoder.weight is dec.roberta.embeddings.word_embeddings.weight
    assert dec.lm_head.decoder.bias is dec.lm_head.bias
