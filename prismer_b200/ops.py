"""Thin tensor-level wrappers over the C-ABI (pointer + dims marshalling only; all arithmetic is in the CUDA library)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _C
from ._C import GemmArgs, check

ACT = {"none": 0, None: 0, "quickgelu": 1, "gelu": 2, "sqrelu": 3, "relu": 4}
BF16, F32 = torch.bfloat16, torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _C.PrismerError("prismer_b200 ops need CUDA tensors (there is no CPU fallback)")


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, f"need a 2-D row-major view, got {tuple(t.shape)} / {t.stride()}"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gemm(a: torch.Tensor, b: torch.Tensor, *, trans_a: bool = False, trans_b: bool = False,
         bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act=0,
         aux_out: Optional[torch.Tensor] = None, aux_in: Optional[torch.Tensor] = None, act_grad=0,
         out: Optional[torch.Tensor] = None, out_dtype=BF16, accumulate: bool = False, alpha: float = 1.0,
         drop_p: float = 0.0, seed: Optional[torch.Tensor] = None, rng_stream: int = 0, force_bn: int = 0,
         max_ctas: int = 0) -> torch.Tensor:
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
        assert seed is not None and seed.dtype == torch.int64 and seed.is_cuda
        args.seed = seed.data_ptr()
    args.rng_stream = rng_stream
    args.force_bn, args.max_ctas = force_bn, max_ctas
    check(_C.lib().prismer_gemm_bf16(ctypes.byref(args), _stream()), "gemm_bf16")
    return out


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, save_stats: bool = True):
    """x: [..., D] bf16 (rows contiguous along D).  Returns (y, mean, rstd)."""
    _req_cuda(x)
    D = x.shape[-1]
    x2 = x.reshape(-1, D)
    rows = x2.shape[0]
    y = torch.empty_like(x2)
    mean = rstd = None
    if save_stats:
        stats = torch.empty((2, rows), dtype=F32, device=x.device)
        mean, rstd = stats[0], stats[1]
    check(_C.lib().prismer_layernorm_fwd(x2.data_ptr(), _ld(x2), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ld(y),
                                         _p(mean), _p(rstd), rows, D, eps, _stream()), "layernorm_fwd")
    return y.view(x.shape), mean, rstd


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
