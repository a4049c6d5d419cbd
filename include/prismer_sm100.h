/* libprismer_sm100.so -- C ABI of the B200-native Prismer hot path (sm_100a only).
 *
 * Conventions:
 *   - every function is enqueue-only on the passed stream, allocates nothing, never throws / exits;
 *   - returns 0 on success, a negative PRISMER_ERR_* code otherwise (cudaGetLastError folded in);
 *   - the caller (PyTorch's caching allocator on the Python side) owns every buffer;
 *   - activations / compute weights are bf16, row-major, 16-byte aligned; statistics, losses, gradients of
 *     parameters, master weights and optimizer state are fp32; token ids are int64 as in the reference.
 *
 * Each entry point cites the reference interface (file:line under NVlabs/prismer @ 4f27ab3) that it replaces.
 */
#ifndef PRISMER_SM100_H_
#define PRISMER_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define PRISMER_ABI_VERSION 1

/* activation codes */
#define PRISMER_ACT_NONE 0
#define PRISMER_ACT_QUICKGELU 1 /* model/modules/utils.py:23-25 */
#define PRISMER_ACT_GELU 2      /* exact erf GELU, model/modules/roberta.py:164,423 */
#define PRISMER_ACT_SQRELU 3    /* model/modules/utils.py:28-30 */
#define PRISMER_ACT_RELU 4      /* nn.ReLU in the conv stems, model/modules/vit.py:91-118 */

int prismer_abi_version(void);
/* 0 when the current device is sm_100 (B200); PRISMER_ERR_ARCH otherwise. */
int prismer_check_device(void);

/* ---------------------------------------------------------------------------------------------------------
 * GEMM:  C[M,N] = epilogue(alpha * op(A) . op(B)^T)       (tcgen05.mma + TMA + TMEM, persistent)
 *   transA = 0: A is [M,K] row-major (lda)      transA = 1: A is stored [K,M] row-major (lda)   (MN-major)
 *   transB = 0: B is [N,K] row-major (ldb)      transB = 1: B is stored [K,N] row-major (ldb)   (MN-major)
 *   epilogue order: +bias[N] -> (aux_out <- pre-activation) -> act | * act'(aux_in) -> dropout -> +residual -> store
 * Replaces: nn.Linear / nn.MultiheadAttention in/out projections (vit.py:41-47,52-53; resampler.py:18-24;
 * utils.py:52-56; roberta.py:84-91,98-104,133,163,176,415,418,422-425) and the stem convolutions expressed as
 * im2col GEMMs (vit.py:86-120), forward (x.W^T), dgrad (dy.W) and wgrad (dy^T.x).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct PrismerGemmArgs {
  const void* A;
  const void* B;
  void* C;
  int M, N, K;
  long long lda, ldb, ldc;
  int transA, transB;
  const float* bias;       /* fp32 [N] or NULL */
  const void* residual;    /* bf16 [M,N] (ldr) or NULL */
  long long ldr;
  void* aux_out;           /* bf16 [M,N] (ldaux): receives the pre-activation (training) or NULL */
  const void* aux_in;      /* bf16 [M,N] (ldaux): pre-activation consumed when act_grad != 0 */
  long long ldaux;
  int act;                 /* PRISMER_ACT_* applied in the forward epilogue */
  int act_grad;            /* PRISMER_ACT_*: multiply the result by act'(aux_in) (dgrad through the activation) */
  int out_fp32;            /* 0: C is bf16, 1: C is fp32 */
  int accumulate;          /* fp32 output only: C += result */
  float alpha;
  float drop_p;            /* dropout probability on the branch output, 0 = off (roberta.py:134,177) */
  const unsigned long long* seed; /* device pointer to the 64-bit Philox key (graph-replay safe) */
  uint32_t rng_stream;     /* distinguishes dropout call sites */
  int force_bn;            /* 0 = heuristic, else 64 / 128 / 256 */
  int max_ctas;            /* 0 = number of SMs */
} PrismerGemmArgs;

int prismer_gemm_bf16(const PrismerGemmArgs* args, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * LayerNorm (fp32 statistics, bf16 in/out), eps as given.  mean/rstd (fp32 [rows]) may be NULL in inference.
 * Replaces model/modules/utils.py:14-19 (LayerNorm.forward) at: vit.py:57,59,169,171; utils.py:62,64;
 * resampler.py:34-35; roberta.py:74,139,182,424 -- and its autograd backward.
 * bwd: dx = LN'(dy) [+ dres]; optional dz = dropout_mask(LN'(dy))/(1-p) (same Philox stream as the forward GEMM epilogue
 * that applied the dropout, roberta.py:134-140); dgamma/dbeta are ACCUMULATED (+=) into fp32 [D] (NULL = frozen).
 * --------------------------------------------------------------------------------------------------------- */
int prismer_layernorm_fwd(const void* x, long long ldx, const float* gamma, const float* beta, void* y, long long ldy,
                          float* mean, float* rstd, int rows, int D, float eps, cudaStream_t stream);
int prismer_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* mean,
                          const float* rstd, const float* gamma, const void* dres, long long lddres, void* dx,
                          long long lddx, void* dz, long long lddz, float* dgamma, float* dbeta, int rows, int D,
                          float drop_p, const unsigned long long* seed, uint32_t rng_stream, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PRISMER_SM100_H_ */
