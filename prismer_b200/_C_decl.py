"""argtypes / restype declarations for the remaining C-ABI entry points (kept next to the header)."""
from ctypes import c_float, c_int, c_longlong, c_uint32, c_ulonglong, c_void_p

P, I, L, F, U32, U64 = c_void_p, c_int, c_longlong, c_float, c_uint32, c_ulonglong

SIGNATURES = {
    "prismer_layernorm_fwd": [P, L, P, P, P, L, P, P, I, I, F, P],
    "prismer_layernorm_bwd": [P, L, P, L, P, P, P, P, L, P, L, P, L, P, P, I, I, F, P, U32, P],
    "prismer_colsum": [P, L, P, I, I, P],
    "prismer_act_bwd": [P, P, P, L, I, P],
    "prismer_cast_f32_bf16": [P, P, L, P],
    "prismer_dropout": [P, P, L, F, P, U32, P],
    "prismer_adamw_step": [P, P, P, P, P, L, F, F, F, F, F, I, F, P],
    "prismer_assemble_tokens": [P, P, P, P, P, P, L, L, I, I, I, I, I, I, I, P],
    "prismer_assemble_tokens_bwd": [P, L, L, P, P, P, P, I, I, I, I, I, I, I, P],
    "prismer_id_presence": [P, L, P, P],
    "prismer_pos_grad": [P, L, L, I, I, I, I, I, P, P],
    "prismer_broadcast_rows": [P, P, L, L, I, I, I, P],
    "prismer_reduce_batch": [P, L, L, I, I, I, P, P],
    "prismer_copy_rows": [P, L, P, L, L, I, I, P],
    "prismer_embed_fwd": [P, P, P, P, P, P, I, I, I, I, I, P],
    "prismer_embed_bwd": [P, P, P, P, P, P, I, I, I, P],
    "prismer_ce_loss_fwd": [P, L, P, P, P, P, P, P, I, I, I, F, P],
    "prismer_ce_loss_bwd": [P, L, P, P, P, P, P, L, I, I, I, F, P],
    "prismer_argmax": [P, L, I, I, I, I, P, P],
    "prismer_patchify": [P, P, I, I, I, I, I, P],
    "prismer_resample_bilinear": [P, P, I, I, I, I, I, I, P],
    "prismer_expand_labels": [P, P, L, P, I, I, L, I, P],
    "prismer_label_resample": [P, P, L, P, I, I, I, I, I, I, P],
    "prismer_im2col_first": [P, I, L, L, L, L, P, I, I, I, I, I, I, I, I, I, P],
    "prismer_im2col_nhwc": [P, P, P, P, I, I, I, I, I, I, I, I, P],
    "prismer_bn_stats": [P, P, L, I, P, P, P, P, P, P, P, P, F, F, I, P],
    "prismer_bn_relu_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, P],
    "prismer_bn_relu_bwd_eval": [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, P],
    "prismer_conv_weight_pack": [P, P, I, I, I, I, P],
    "prismer_conv_weight_unpack_grad": [P, P, I, I, I, I, P],
    "prismer_cast_pad": [P, P, L, I, I, P],
    "prismer_unpad_add": [P, P, L, I, I, P],
    "prismer_set_attention_path": [I],
    "prismer_skinny_linear": [P, L, P, L, P, P, L, P, L, I, I, I, I, I, P],
    "prismer_decode_attention": [P, L, P, P, L, L, I, P, P, L, P, P, L, L, P, I, P, L, I, I, I, F, P],
}


def declare(lib):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
