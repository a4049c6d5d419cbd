"""Thin tensor-level wrappers over the C-ABI (pointer + dims marshalling only; all arithmetic is in the CUDA library)."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import _C
from ._C import GemmArgs

GEMM_PROFILE = None   # set to a list to record (M, N, K, start_event, end_event) per GEMM launch (bench.py roofline pass)

ACT = {"none": 0, None: 0, "quickgelu": 1, "gelu": 2, "sqrelu": 3, "relu": 4}
BF16, F32 = torch.bfloat16, torch.float32


def check(rc, what=""):
    _C.check(rc, what)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _C.PrismerError("prismer_b200 ops need CUDA tensors (there is no CPU fallback)")


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), f"need a 2-D row-major view, got {tuple(t.shape)} / {t.stride()}"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gemm(a: torch.Tensor, b: torch.Tensor, *, trans_a: bool = False, trans_b: bool = False,
         bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act=0,
         aux_out: Optional[torch.Tensor] = None, aux_in: Optional[torch.Tensor] = None, act_grad=0,
         out: Optional[torch.Tensor] = None, out_dtype=BF16, accumulate: bool = False, alpha: float = 1.0,
         drop_p: float = 0.0, seed: Optional[torch.Tensor] = None, rng_stream: int = 0, force_bn: int = 0,
         max_ctas: int = 0, force_splits: int = 0) -> torch.Tensor:
    """C[M,N] = epilogue(alpha * op(A) . op(B)^T); A is [M,K] (or [K,M] if trans_a), B is [N,K] (or [K,N] if trans_b)."""
    _req_cuda(a, b)
    assert a.dtype == BF16 and b.dtype == BF16
    if trans_a:
        K, M = a.shape
    else:
        M, K = a.shape
    if trans_b:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, f"contraction mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N)
    act = ACT.get(act, act)
    act_grad = ACT.get(act_grad, act_grad)
    args = GemmArgs()
    args.A, args.B, args.C = a.data_ptr(), b.data_ptr(), out.data_ptr()
    args.M, args.N, args.K = M, N, K
    args.lda, args.ldb, args.ldc = _ld(a), _ld(b), _ld(out)
    args.transA, args.transB = int(trans_a), int(trans_b)
    if bias is not None:
        assert bias.dtype == F32 and bias.numel() == N and bias.is_contiguous()
        args.bias = bias.data_ptr()
    if residual is not None:
        assert residual.dtype == BF16 and residual.shape == (M, N)
        args.residual, args.ldr = residual.data_ptr(), _ld(residual)
    aux = aux_out if aux_out is not None else aux_in
    if aux is not None:
        assert aux.dtype == BF16 and aux.shape == (M, N)
        args.ldaux = _ld(aux)
        args.aux_out, args.aux_in = _p(aux_out), _p(aux_in)
    args.act, args.act_grad = act, act_grad
    args.out_fp32 = int(out.dtype == F32)
    args.accumulate = int(accumulate)
    args.alpha = alpha
    args.drop_p = drop_p
    if drop_p > 0:
        assert seed is not None and seed.dtype == torch.int64
        _req_cuda(seed)
        args.seed = seed.data_ptr()
    args.rng_stream = rng_stream
    args.force_bn, args.max_ctas, args.force_splits = force_bn, max_ctas, force_splits
    fn = _C.lib().prismer_gemm_bf16
    if GEMM_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(ctypes.byref(args), _stream()), "gemm_bf16")
        e1.record()
        GEMM_PROFILE.append((M, N, K, e0, e1))
        return out
    check(fn(ctypes.byref(args), _stream()), "gemm_bf16")
    return out


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, save_stats: bool = True,
                  out: Optional[torch.Tensor] = None):
    """x: [..., D] bf16 (rows contiguous along D).  Returns (y, mean, rstd)."""
    _req_cuda(x)
    D = x.shape[-1]
    x2 = x.reshape(-1, D)
    rows = x2.shape[0]
    y = torch.empty((rows, D), dtype=BF16, device=x.device) if out is None else out
    mean = rstd = None
    if save_stats:
        stats = torch.empty((2, rows), dtype=F32, device=x.device)
        mean, rstd = stats[0], stats[1]
    check(_C.lib().prismer_layernorm_fwd(x2.data_ptr(), _ld(x2), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ld(y),
                                         _p(mean), _p(rstd), rows, D, eps, _stream()), "layernorm_fwd")
    return (y.view(x.shape) if out is None else y), mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, *, dres=None, dgamma=None, dbeta=None, need_dx=True, dz=False,
                  drop_p: float = 0.0, seed=None, rng_stream: int = 0):
    D = x.shape[-1]
    dy2, x2 = dy.reshape(-1, D), x.reshape(-1, D)
    rows = x2.shape[0]
    dx = torch.empty_like(x2) if need_dx else None
    dzt = torch.empty_like(x2) if dz else None
    dres2 = dres.reshape(-1, D) if dres is not None else None
    check(_C.lib().prismer_layernorm_bwd(dy2.data_ptr(), _ld(dy2), x2.data_ptr(), _ld(x2), mean.data_ptr(), rstd.data_ptr(),
                                         gamma.data_ptr(), _p(dres2), _ld(dres2) if dres2 is not None else 0,
                                         _p(dx), _ld(dx) if dx is not None else 0, _p(dzt), _ld(dzt) if dzt is not None else 0,
                                         _p(dgamma), _p(dbeta), rows, D, drop_p, _p(seed), rng_stream, _stream()),
          "layernorm_bwd")
    return (dx.view(x.shape) if dx is not None else None), (dzt.view(x.shape) if dzt is not None else None)


def _bhv(t: torch.Tensor, H: int, d: int):
    """[B, L, H*d]-shaped view (last dim contiguous) -> (ptr, batch stride, row stride)."""
    assert t.dim() == 3 and t.stride(2) == 1 and t.shape[2] == H * d and t.dtype == BF16, (t.shape, t.stride())
    return t.data_ptr(), t.stride(0), t.stride(1)


def set_attention_path(legacy: bool):
    """True: every attention call runs the mma.sync kernels; False (default): the tcgen05 / TMEM kernels where they apply."""
    check(_C.lib().prismer_set_attention_path(int(bool(legacy))), "set_attention_path")


def attention_fwd(q, k, v, heads: int, *, causal=False, key_mask=None, drop_p=0.0, seed=None, rng_stream=0,
                  need_lse=True, scale=None, out=None, kv_div=1):
    """q: [B,Lq,H*d]; k, v: [B / kv_div, Lk, H*d] (views into packed projections are fine; with ``kv_div`` > 1 every group of kv_div
    consecutive query batch rows reads the same keys / values).  Returns (o [B,Lq,H*d], lse [B,H,Lq])."""
    _req_cuda(q, k, v)
    B, Lq, HD = q.shape
    Lk = k.shape[1]
    d = HD // heads
    o = torch.empty((B, Lq, HD), dtype=BF16, device=q.device) if out is None else out
    lse = torch.empty((B, heads, Lq), dtype=F32, device=q.device) if need_lse else None
    a = _C.AttnArgs()
    a.q, a.q_bs, a.q_rs = _bhv(q, heads, d)
    a.k, a.k_bs, a.k_rs = _bhv(k, heads, d)
    a.v, a.v_bs, a.v_rs = _bhv(v, heads, d)
    a.o, a.o_bs, a.o_rs = _bhv(o, heads, d)
    a.lse = _p(lse)
    if key_mask is not None:
        assert key_mask.dtype == torch.int64 and key_mask.shape == (B, Lk) and key_mask.is_contiguous()
        a.key_mask = key_mask.data_ptr()
    a.B, a.H, a.Lq, a.Lk, a.d, a.causal = B, heads, Lq, Lk, d, int(causal)
    a.scale = scale if scale is not None else d ** -0.5
    a.drop_p = drop_p
    if drop_p > 0:
        a.seed = seed.data_ptr()
    a.rng_stream = rng_stream
    a.kv_div = kv_div
    assert k.shape[0] * kv_div == B, "keys / values batch must be the query batch divided by kv_div"
    check(_C.lib().prismer_attention_fwd(ctypes.byref(a), _stream()), "attention_fwd")
    return o, lse


def attention_bwd(dout, q, k, v, o, lse, heads: int, *, causal=False, key_mask=None, drop_p=0.0, seed=None,
                  rng_stream=0, scale=None, dq=None, dk=None, dv=None):
    B, Lq, HD = q.shape
    Lk = k.shape[1]
    d = HD // heads
    dq = torch.empty((B, Lq, HD), dtype=BF16, device=q.device) if dq is None else dq
    dk = torch.empty((B, Lk, HD), dtype=BF16, device=q.device) if dk is None else dk
    dv = torch.empty((B, Lk, HD), dtype=BF16, device=q.device) if dv is None else dv
    delta = torch.empty((B, heads, Lq), dtype=F32, device=q.device)
    a = _C.AttnArgs()
    a.q, a.q_bs, a.q_rs = _bhv(q, heads, d)
    a.k, a.k_bs, a.k_rs = _bhv(k, heads, d)
    a.v, a.v_bs, a.v_rs = _bhv(v, heads, d)
    a.o, a.o_bs, a.o_rs = _bhv(o, heads, d)
    a.dout, a.do_bs, a.do_rs = _bhv(dout, heads, d)
    a.dq, a.dq_bs, a.dq_rs = _bhv(dq, heads, d)
    a.dk, a.dk_bs, a.dk_rs = _bhv(dk, heads, d)
    a.dv, a.dv_bs, a.dv_rs = _bhv(dv, heads, d)
    a.lse, a.delta = lse.data_ptr(), delta.data_ptr()
    if key_mask is not None:
        a.key_mask = key_mask.data_ptr()
    a.B, a.H, a.Lq, a.Lk, a.d, a.causal = B, heads, Lq, Lk, d, int(causal)
    a.scale = scale if scale is not None else d ** -0.5
    a.drop_p = drop_p
    if drop_p > 0:
        a.seed = seed.data_ptr()
    a.rng_stream = rng_stream
    check(_C.lib().prismer_attention_bwd(ctypes.byref(a), _stream()), "attention_bwd")
    return dq, dk, dv


# ------------------------------------------------------------------------------------------------ small kernels
def colsum(x2d: torch.Tensor, out: torch.Tensor):
    """out[N] (fp32) += column sums of x2d [M,N] bf16."""
    M, N = x2d.shape
    check(_C.lib().prismer_colsum(x2d.data_ptr(), _ld(x2d), out.data_ptr(), M, N, _stream()), "colsum")


def act_bwd(dy, z, act):
    dz = torch.empty_like(dy)
    check(_C.lib().prismer_act_bwd(dy.data_ptr(), z.data_ptr(), dz.data_ptr(), dy.numel(), ACT.get(act, act), _stream()), "act_bwd")
    return dz


def dropout(x, p, seed, rng_stream):
    y = torch.empty_like(x)
    check(_C.lib().prismer_dropout(x.data_ptr(), y.data_ptr(), x.numel(), p, seed.data_ptr(), rng_stream, _stream()), "dropout")
    return y


def cast_bf16(src: torch.Tensor, dst: torch.Tensor):
    assert src.dtype == F32 and dst.dtype == BF16 and src.numel() == dst.numel()
    check(_C.lib().prismer_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "cast")


def adamw_step(p, g, m, v, p16, lr, beta1, beta2, eps, wd, step, grad_scale=1.0):
    check(_C.lib().prismer_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _p(p16), p.numel(), lr, beta1, beta2,
                                      eps, wd, step, grad_scale, _stream()), "adamw")


def assemble_tokens(src, pos16, dst, dst_bs, dst_rs, B, n_tok, D, gh, gw, inst=None, table=None, inst_emb16=None):
    Hi, Wi = (inst.shape[-2], inst.shape[-1]) if inst is not None else (0, 0)
    check(_C.lib().prismer_assemble_tokens(src.data_ptr(), pos16.data_ptr(), _p(inst), _p(table), _p(inst_emb16), dst.data_ptr(),
                                           dst_bs, dst_rs, B, n_tok, D, gh, gw, Hi, Wi, _stream()), "assemble_tokens")


def id_presence(inst, out=None):
    flags = torch.empty(256, dtype=torch.int32, device=inst.device) if out is None else out
    check(_C.lib().prismer_id_presence(inst.data_ptr(), inst.numel(), flags.data_ptr(), _stream()), "id_presence")
    return flags


def assemble_tokens_bwd(ddst, ddst_bs, ddst_rs, dsrc, B, n_tok, D, gh, gw, inst=None, table=None, dinst_emb=None):
    Hi, Wi = (inst.shape[-2], inst.shape[-1]) if inst is not None else (0, 0)
    check(_C.lib().prismer_assemble_tokens_bwd(ddst.data_ptr(), ddst_bs, ddst_rs, dsrc.data_ptr(), _p(inst), _p(table), _p(dinst_emb), B,
                                               n_tok, D, gh, gw, Hi, Wi, _stream()), "assemble_tokens_bwd")


def pos_grad(dtok, bs, rs, B, n_tok, D, n_slots, slot_stride, dpos):
    check(_C.lib().prismer_pos_grad(dtok.data_ptr(), bs, rs, B, n_tok, D, n_slots, slot_stride, dpos.data_ptr(), _stream()), "pos_grad")


def broadcast_rows(src16, dst, dst_bs, dst_rs, B, n, D):
    check(_C.lib().prismer_broadcast_rows(src16.data_ptr(), dst.data_ptr(), dst_bs, dst_rs, B, n, D, _stream()), "broadcast_rows")


def reduce_batch(d, bs, rs, B, n, D, out):
    check(_C.lib().prismer_reduce_batch(d.data_ptr(), bs, rs, B, n, D, out.data_ptr(), _stream()), "reduce_batch")


def copy_rows(src2d, dst2d, add=False):
    rows, D = src2d.shape
    check(_C.lib().prismer_copy_rows(src2d.data_ptr(), _ld(src2d), dst2d.data_ptr(), _ld(dst2d), rows, D, int(add), _stream()), "copy_rows")


def embed_fwd(ids, word16, pos16, type16, pad_id, past_len=0):
    B, T = ids.shape
    H = word16.shape[1]
    out = torch.empty((B * T, H), dtype=BF16, device=ids.device)
    pos_ids = torch.empty((B * T,), dtype=torch.int32, device=ids.device)
    check(_C.lib().prismer_embed_fwd(ids.data_ptr(), word16.data_ptr(), pos16.data_ptr(), type16.data_ptr(), out.data_ptr(),
                                     pos_ids.data_ptr(), B, T, H, pad_id, past_len, _stream()), "embed_fwd")
    return out, pos_ids


def embed_bwd(de, ids, pos_ids, dword, dpos, dtype_, pad_id):
    rows, H = de.shape
    check(_C.lib().prismer_embed_bwd(de.data_ptr(), ids.data_ptr(), pos_ids.data_ptr(), _p(dword), _p(dpos), _p(dtype_), rows, H,
                                     pad_id, _stream()), "embed_bwd")


def ce_loss_fwd(logits, labels, V, weights=None, smoothing=0.1):
    """logits fp32 [B*T, ld>=V]; labels int64 [B,T].  Returns (mean_loss[1], sample_loss[B], row_lse[B*T])."""
    B, T = labels.shape
    dev = logits.device
    row_loss = torch.empty(B * T, dtype=F32, device=dev)
    row_lse = torch.empty(B * T, dtype=F32, device=dev)
    sample = torch.empty(B, dtype=F32, device=dev)
    mean = torch.empty(1, dtype=F32, device=dev)
    check(_C.lib().prismer_ce_loss_fwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), _p(weights), row_loss.data_ptr(),
                                       row_lse.data_ptr(), sample.data_ptr(), mean.data_ptr(), B, T, V, smoothing, _stream()), "ce_fwd")
    return mean, sample, row_lse


def ce_loss_bwd(logits, labels, row_lse, V, weights=None, gscale=None, smoothing=0.1):
    B, T = labels.shape
    ldo = logits.stride(0)
    d = torch.empty((B * T, ldo), dtype=BF16, device=logits.device)
    check(_C.lib().prismer_ce_loss_bwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), row_lse.data_ptr(), _p(weights),
                                       _p(gscale), d.data_ptr(), ldo, B, T, V, smoothing, _stream()), "ce_bwd")
    return d


def argmax(logits, V, suppress_eos=False, eos=2):
    rows = logits.shape[0]
    out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    check(_C.lib().prismer_argmax(logits.data_ptr(), logits.stride(0), rows, V, int(suppress_eos), eos, out.data_ptr(), _stream()), "argmax")
    return out


# ------------------------------------------------------------------------------------------------ conv stems
def patchify(x, p, Kpad):
    B, C, R, _ = x.shape
    g = R // p
    out = torch.empty((B * g * g, Kpad), dtype=BF16, device=x.device)
    check(_C.lib().prismer_patchify(x.data_ptr(), out.data_ptr(), B, C, R, p, Kpad, _stream()), "patchify")
    return out


def resample_bilinear(x, Ho, Wo):
    B, C, Hi, Wi = x.shape
    out = torch.empty((B, Ho, Wo, C), dtype=BF16, device=x.device)
    check(_C.lib().prismer_resample_bilinear(x.data_ptr(), out.data_ptr(), B, C, Hi, Wi, Ho, Wo, _stream()), "resample")
    return out


def _table_stride(table, B):
    """table: fp32 [256, C] (shared) or [B, 256, C] (per image) -> element stride between images."""
    assert table.dtype == F32 and table.is_contiguous() and table.shape[-2] == 256
    if table.dim() == 2:
        return 0
    assert table.shape[0] == B
    return table.shape[1] * table.shape[2]


def expand_labels(u8, table):
    """uint8 [B, Cin, H, W] through fp32 table [256, C] / [B, 256, C] -> fp32 NCHW [B, Cin*C, H, W] (dataset/utils.py:117-160)."""
    assert u8.dtype == torch.uint8 and u8.is_contiguous() and u8.dim() == 4
    B, Cin, H, W = u8.shape
    C = table.shape[-1]
    out = torch.empty((B, Cin * C, H, W), dtype=F32, device=u8.device)
    check(_C.lib().prismer_expand_labels(u8.data_ptr(), table.data_ptr(), _table_stride(table, B), out.data_ptr(), B, Cin, H * W, C,
                                         _stream()), "expand_labels")
    return out


def label_resample(u8, table, Ho, Wo):
    """uint8 [B, 1, H, W] label map + table -> bf16 NHWC [B, Ho, Wo, C]: in-painting fused with UpsamplingBilinear2d (vit.py:89)."""
    assert u8.dtype == torch.uint8 and u8.is_contiguous() and u8.dim() == 4 and u8.shape[1] == 1
    B, _, Hi, Wi = u8.shape
    C = table.shape[-1]
    out = torch.empty((B, Ho, Wo, C), dtype=BF16, device=u8.device)
    check(_C.lib().prismer_label_resample(u8.data_ptr(), table.data_ptr(), _table_stride(table, B), out.data_ptr(), B, C, Hi, Wi, Ho, Wo,
                                          _stream()), "label_resample")
    return out


def im2col_first(x, nhwc_bf16: bool, B, Cin, H, W, ksz, stride, Kpad):
    Ho, Wo = (H + 2 * (ksz // 2) - ksz) // stride + 1, (W + 2 * (ksz // 2) - ksz) // stride + 1
    out = torch.empty((B * Ho * Wo, Kpad), dtype=BF16, device=x.device)
    if nhwc_bf16:
        sb, sc, sy, sx = H * W * Cin, 1, W * Cin, Cin
    else:
        sb, sc, sy, sx = Cin * H * W, H * W, W, 1
    check(_C.lib().prismer_im2col_first(x.data_ptr(), int(nhwc_bf16), sb, sc, sy, sx, out.data_ptr(), B, Cin, H, W, ksz, stride, Ho, Wo,
                                        Kpad, _stream()), "im2col_first")
    return out, Ho, Wo


def im2col_nhwc(x, B, H, W, C, ksz, stride, scale=None, shift=None):
    Ho, Wo = (H + 2 * (ksz // 2) - ksz) // stride + 1, (W + 2 * (ksz // 2) - ksz) // stride + 1
    out = torch.empty((B * Ho * Wo, ksz * ksz * C), dtype=BF16, device=x.device)
    check(_C.lib().prismer_im2col_nhwc(x.data_ptr(), _p(scale), _p(shift), out.data_ptr(), B, H, W, C, ksz, stride, Ho, Wo, _stream()),
          "im2col_nhwc")
    return out, Ho, Wo


def bn_stats(y2d, bn, training: bool):
    """Returns per-channel fp32 (scale, shift, mean, rstd) for BatchNorm2d module ``bn`` over y2d [M, C] (NHWC rows)."""
    M, C = y2d.shape
    st = torch.empty((6, C), dtype=F32, device=y2d.device)
    check(_C.lib().prismer_bn_stats(y2d.data_ptr(), st[4:6].data_ptr(), M, C, bn.weight.data_ptr(), bn.bias.data_ptr(),
                                    bn.running_mean.data_ptr(), bn.running_var.data_ptr(), st[0].data_ptr(), st[1].data_ptr(),
                                    st[2].data_ptr(), st[3].data_ptr(), bn.eps, bn.momentum, int(training), _stream()), "bn_stats")
    return st[0], st[1], st[2], st[3]


def bn_relu_bwd(dAcol, y2d, scale, shift, mean, rstd, gamma, dgamma, dbeta, B, H, W, C, ksz, stride, Ho, Wo, training=True):
    """``training=False``: the BatchNorm ran on its running statistics (eval()) -> gradient without the batch-mean terms."""
    M = B * H * W
    dn = torch.empty((M, C), dtype=BF16, device=y2d.device)
    dy = torch.empty((M, C), dtype=BF16, device=y2d.device)
    red = torch.empty((2, C), dtype=F32, device=y2d.device)
    fn = _C.lib().prismer_bn_relu_bwd if training else _C.lib().prismer_bn_relu_bwd_eval
    check(fn(dAcol.data_ptr(), y2d.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                       rstd.data_ptr(), gamma.data_ptr(), dn.data_ptr(), dy.data_ptr(), red.data_ptr(), _p(dgamma),
                                       _p(dbeta), B, H, W, C, ksz, stride, Ho, Wo, _stream()), "bn_relu_bwd")
    return dy


def conv_weight_pack(w, Kpad, out=None):
    Cout, Cin, k, _ = w.shape
    if out is None or out.shape != (Cout, Kpad):
        out = torch.empty((Cout, Kpad), dtype=BF16, device=w.device)
    check(_C.lib().prismer_conv_weight_pack(w.data_ptr(), out.data_ptr(), Cout, Cin, k, Kpad, _stream()), "conv_weight_pack")
    return out


def conv_weight_unpack_grad(dwp, grad):
    Cout, Cin, k, _ = grad.shape
    check(_C.lib().prismer_conv_weight_unpack_grad(dwp.data_ptr(), grad.data_ptr(), Cout, Cin, k, dwp.shape[1], _stream()), "conv_unpack")


def cast_pad(src2d, Cpad, out=None):
    R, C = src2d.shape
    if out is None or out.shape != (R, Cpad):
        out = torch.empty((R, Cpad), dtype=BF16, device=src2d.device)
    check(_C.lib().prismer_cast_pad(src2d.data_ptr(), out.data_ptr(), R, C, Cpad, _stream()), "cast_pad")
    return out


def unpad_add(src2d, dst2d):
    R, C = dst2d.shape
    check(_C.lib().prismer_unpad_add(src2d.data_ptr(), dst2d.data_ptr(), R, C, src2d.shape[1], _stream()), "unpad_add")


# ------------------------------------------------------------------------------------------------ EXPERIMENTAL (round-2 candidate)
