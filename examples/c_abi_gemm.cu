// The C ABI without PyTorch: y = quickgelu(x . W^T + b) through prismer_gemm_bf16 and a LayerNorm through prismer_layernorm_fwd on
// plain cudaMalloc'd buffers, checked against a host loop.  This is what a non-Python consumer of include/prismer_sm100.h links.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -I include examples/c_abi_gemm.cu -o /tmp/c_abi_gemm \
//        -L prismer_b200 -lprismer_sm100 -Xlinker -rpath -Xlinker $PWD/prismer_b200 && /tmp/c_abi_gemm
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "prismer_sm100.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
  const int M = 300, N = 200, K = 136;          // ragged on purpose: M, N tails and a K tail
  printf("prismer_abi_version = %d\n", prismer_abi_version());
  std::vector<__nv_bfloat16> hx(M * K), hw(N * K);
  std::vector<float> hb(N), ref(M * N);
  srand(1);
  for (auto& v : hx) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
  for (auto& v : hw) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
  for (auto& v : hb) v = (rand() % 2001 - 1000) / 1000.0f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += __bfloat162float(hx[m * K + k]) * __bfloat162float(hw[n * K + k]);
      const float z = acc + hb[n];
      ref[m * N + n] = z / (1.0f + expf(-1.702f * z));                       // QuickGELU, model/modules/utils.py:25
    }
  __nv_bfloat16 *dx, *dw, *dy;
  float* db;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2)); CK(cudaMalloc(&dy, (size_t)M * N * 2)); CK(cudaMalloc(&db, N * 4));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), N * 4, cudaMemcpyHostToDevice));
  cudaStream_t s;
  CK(cudaStreamCreate(&s));

  PrismerGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.A = dx; a.B = dw; a.C = dy; a.M = M; a.N = N; a.K = K; a.lda = K; a.ldb = K; a.ldc = N;
  a.bias = db; a.act = PRISMER_ACT_QUICKGELU; a.alpha = 1.0f;
  int rc = prismer_gemm_bf16(&a, s);
  if (rc) { printf("prismer_gemm_bf16 failed: %d\n", rc); return 1; }
  CK(cudaStreamSynchronize(s));
  std::vector<__nv_bfloat16> hy((size_t)M * N);
  CK(cudaMemcpy(hy.data(), dy, hy.size() * 2, cudaMemcpyDeviceToHost));
  double num = 0, den = 0;
  for (size_t i = 0; i < hy.size(); ++i) { const double d = __bfloat162float(hy[i]) - ref[i]; num += d * d; den += (double)ref[i] * ref[i]; }
  const double gemm_err = sqrt(num / den);
  printf("gemm + bias + quickgelu: rel-L2 vs host loop %.2e\n", gemm_err);

  // misaligned leading dimension must be refused with an error code, not a crash
  a.lda = K + 1;
  const int rc_bad = prismer_gemm_bf16(&a, s);
  printf("misaligned lda -> rc %d (expected %d)\n", rc_bad, PRISMER_ERR_ALIGN);

  // LayerNorm over the GEMM output rows (N = 200 is a multiple of 8)
  std::vector<float> hg(N, 1.0f), hbeta(N, 0.0f);
  float *dg, *dbeta, *dmean, *drstd;
  __nv_bfloat16* dz;
  CK(cudaMalloc(&dg, N * 4)); CK(cudaMalloc(&dbeta, N * 4)); CK(cudaMalloc(&dmean, M * 4)); CK(cudaMalloc(&drstd, M * 4)); CK(cudaMalloc(&dz, (size_t)M * N * 2));
  CK(cudaMemcpy(dg, hg.data(), N * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbeta, hbeta.data(), N * 4, cudaMemcpyHostToDevice));
  rc = prismer_layernorm_fwd(dy, N, dg, dbeta, dz, N, dmean, drstd, M, N, 1e-5f, s);
  if (rc) { printf("prismer_layernorm_fwd failed: %d\n", rc); return 1; }
  CK(cudaStreamSynchronize(s));
  std::vector<__nv_bfloat16> hz((size_t)M * N);
  CK(cudaMemcpy(hz.data(), dz, hz.size() * 2, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (int m = 0; m < M; ++m) {
    double mu = 0, var = 0;
    for (int n = 0; n < N; ++n) mu += __bfloat162float(hy[m * N + n]);
    mu /= N;
    for (int n = 0; n < N; ++n) { const double d = __bfloat162float(hy[m * N + n]) - mu; var += d * d; }
    var /= N;
    for (int n = 0; n < N; ++n) {
      const double want = (__bfloat162float(hy[m * N + n]) - mu) / sqrt(var + 1e-5);
      const double got = __bfloat162float(hz[m * N + n]);
      const double e = fabs(got - bf((float)want));
      if (e > worst) worst = e;
    }
  }
  printf("layernorm: max |diff| vs host loop %.3e\n", worst);
  const bool ok = gemm_err < 1e-2 && rc_bad == PRISMER_ERR_ALIGN && worst < 4e-2;
  printf(ok ? "OK\n" : "FAILED\n");
  return ok ? 0 : 1;
}
