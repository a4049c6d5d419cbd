// Shared device helpers: activations, Philox RNG for dropout, vector load/store, reductions.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#ifndef PRISMER_OK      // also published by include/prismer_sm100.h (identical values)
#define PRISMER_OK 0
#define PRISMER_ERR_SHAPE -1
#define PRISMER_ERR_ALIGN -2
#define PRISMER_ERR_ARCH -3
#define PRISMER_ERR_CUDA -4
#define PRISMER_ERR_DRIVER -5
#endif

// activation codes used across the C-ABI
enum : int { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_GELU = 2, ACT_SQRELU = 3, ACT_RELU = 4 };

// erf with |abs error| < 1.2e-7 (Abramowitz & Stegun 7.1.26-style rational in t = 1/(1+p|x|), one exp + one rcp):
// far below the bf16 output grid, ~4x cheaper than erff() in the GEMM epilogues.
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.0f - p * t * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case ACT_QUICKGELU: return x * sigmoid_fast(1.702f * x);                     // utils.py:25  x*sigmoid(1.702x)
    case ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));   // exact erf GELU (roberta.py:164)
    case ACT_SQRELU: { float r = fmaxf(x, 0.f); return r * r; }                  // utils.py:30
    case ACT_RELU: return fmaxf(x, 0.f);
    default: return x;
  }
}
// derivative wrt the pre-activation z
__device__ __forceinline__ float act_bwd(int act, float z) {
  switch (act) {
    case ACT_QUICKGELU: { float s = sigmoid_fast(1.702f * z); return s * (1.0f + 1.702f * z * (1.0f - s)); }
    case ACT_GELU: {
      float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752f));
      float pdf = 0.3989422804014327f * __expf(-0.5f * z * z);
      return cdf + z * pdf;
    }
    case ACT_SQRELU: return z > 0.f ? 2.0f * z : 0.f;
    case ACT_RELU: return z > 0.f ? 1.0f : 0.f;
    default: return 1.0f;
  }
}
// 32-wide epilogue helpers: the activation code is warp-uniform, so dispatch once per 32-element chunk and keep the inner
// loops branch-free.
__device__ __forceinline__ void act_fwd32(int act, float* v) {
  switch (act) {
    case ACT_QUICKGELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] * sigmoid_fast(1.702f * v[j]);
      break;
    case ACT_GELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + fast_erf(v[j] * 0.70710678118654752f));
      break;
    case ACT_SQRELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float r = fmaxf(v[j], 0.f); v[j] = r * r; }
      break;
    case ACT_RELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    default: break;
  }
}
__device__ __forceinline__ void act_bwd8(int act, float* v, const float* z) {
  switch (act) {
    case ACT_QUICKGELU:
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float s = sigmoid_fast(1.702f * z[t]); v[t] *= s * (1.0f + 1.702f * z[t] * (1.0f - s)); }
      break;
    case ACT_GELU:
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float cdf = 0.5f * (1.0f + fast_erf(z[t] * 0.70710678118654752f));
        v[t] *= cdf + z[t] * 0.3989422804014327f * __expf(-0.5f * z[t] * z[t]);
      }
      break;
    case ACT_SQRELU:
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] *= z[t] > 0.f ? 2.0f * z[t] : 0.f;
      break;
    case ACT_RELU:
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] *= z[t] > 0.f ? 1.0f : 0.f;
      break;
    default: break;
  }
}

// ------------------------------------------------------------------ Philox4x32-10 (counter based; same mask in fwd/bwd)
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0(static_cast<uint32_t>(seed)), k1(static_cast<uint32_t>(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// keep-decisions for 8 consecutive elements starting at element index 8*g of stream `stream`:
// bit j of the result = element 8*g+j is KEPT.  thr16 = round(p * 65536).
__device__ __forceinline__ uint32_t dropout_keep8(const Philox& ph, uint64_t g, uint32_t stream, uint32_t thr16) {
  uint4 r = ph(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), stream, 0x5052534Du);
  uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m |= ((w[j] & 0xFFFFu) >= thr16 ? 1u : 0u) << (2 * j);
    m |= ((w[j] >> 16) >= thr16 ? 1u : 0u) << (2 * j + 1);
  }
  return m;
}
__device__ __forceinline__ bool dropout_keep1(const Philox& ph, uint64_t idx, uint32_t stream, uint32_t thr16) {
  return (dropout_keep8(ph, idx >> 3, stream, thr16) >> (idx & 7)) & 1u;
}

// ------------------------------------------------------------------ small utilities
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8 bf16 moved as ONE 128-bit access (a struct of four bfloat162 compiles to four 32-bit accesses: 4x the LSU work and,
// for row-strided epilogue stores, 4x the partial-sector writes).
typedef uint4 bf16x8;

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
  const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  return make_uint4(pack_bf162(f[0], f[1]), pack_bf162(f[2], f[3]), pack_bf162(f[4], f[5]), pack_bf162(f[6], f[7]));
}

// ------------------------------------------------------------------ programmatic dependent launch (opt-in build flag)
// -DPRISMER_PDL (PRISMER_PDL=1 python -m prismer_b200.build): the hot kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and start with griddepcontrol.launch_dependents + griddepcontrol.wait, so
// the next kernel's launch / prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail;
// every global read or write of such a kernel comes after the wait.  Without the flag (the default, the hardware-validated
// build) both macros are empty and pdl_launch is a plain <<<>>> launch.  Round-2 experiment; see NOTES_NEXT_ROUND.md.
#ifdef PRISMER_PDL
#define PDL_GRID_SYNC() do { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)
#else
#define PDL_GRID_SYNC() ((void)0)
#endif
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline void pdl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
#ifdef PRISMER_PDL
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
#else
  kern<<<grid, block, smem, stream>>>(static_cast<KArgs>(args)...);
#endif
}
#endif

static inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? PRISMER_OK : PRISMER_ERR_CUDA; }
#define LAUNCH_CHECK() cuda_status(cudaGetLastError())
