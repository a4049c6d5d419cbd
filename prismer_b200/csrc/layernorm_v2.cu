// Register-lean LayerNorm backward: the kernel behind prismer_layernorm_bwd for D <= 1024 (every LayerNorm of Prismer-BASE / LARGE).
// Validated on B200 in round 2 against the round-1 kernel of layernorm.cu (dx / dz equal up to one bf16 ulp, dgamma / dbeta up to fp32
// summation order) and adopted: 102 calls of a BASE step 2.81 ms -> 1.89 ms.  The round-1 ln_bwd_kernel<3> (D = 768) needs 146
// registers -> ONE 256-thread block per SM (8 warps) and ~1 TB/s, because every lane carries 48 dgamma/dbeta accumulators plus the
// 48 unpacked xhat / g*dy values across the row reduction.  Here
//   * the inputs stay PACKED in registers (x and dy as uint4: 2 x VPL registers x 4) and are unpacked again for the output pass,
//   * dgamma / dbeta are accumulated in each warp's private shared-memory slice (the [warps][2][D] buffer the validated kernel
//     already allocates for its final reduction) with conflict-free 128-bit read-modify-writes -- and not at all when the
//     LayerNorm is frozen (template flag), which is 24 of the 36 ViT sites under freeze_vision.
// Same arguments and arithmetic as the round-1 kernel (which remains the path for 1024 < D <= 2048).
#include "common.cuh"
#include "prismer_sm100.h"

namespace {

constexpr int kWarps = 8;

template <int VPL, bool HAS_DG>
__global__ void __launch_bounds__(kWarps * 32, HAS_DG ? 2 : 3)
ln_bwd2_kernel(const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ x, long long ldx,
               const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
               const bf16* __restrict__ dres, long long lddres, bf16* __restrict__ dx, long long lddx, bf16* __restrict__ dz,
               long long lddz, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int D, float drop_p, uint32_t thr16,
               const unsigned long long* seed, uint32_t rng_stream) {
  extern __shared__ float red[];  // HAS_DG: [kWarps][2][D]
  PDL_GRID_SYNC();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  float* rg = red + warp * 2 * D;
  float* rb = rg + D;
  if constexpr (HAS_DG) {
    for (int c = lane; c < 2 * D; c += 32) rg[c] = 0.f;
    __syncwarp();
  }
  const bool has_drop = dz != nullptr && drop_p > 0.f;
  const Philox philox(has_drop ? *seed : 0ull);
  const float drop_scale = has_drop ? 1.0f / (1.0f - drop_p) : 1.0f;

  for (long long row = static_cast<long long>(blockIdx.x) * kWarps + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * kWarps) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    uint4 xr[VPL], dr[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        xr[i] = *reinterpret_cast<const uint4*>(x + row * ldx + vi * 8);
        dr[i] = *reinterpret_cast<const uint4*>(dy + row * lddy + vi * 8);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float xv[8], dyv[8];
        unpack8(xr[i], xv);
        unpack8(dr[i], dyv);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float pg[8], pb[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float xh = (xv[t] - mean) * rstd;
          const float gy = g[t] * dyv[t];
          s1 += gy;
          s2 += gy * xh;
          pg[t] = dyv[t] * xh;
          pb[t] = dyv[t];
        }
        if constexpr (HAS_DG) {     // this warp's private accumulators: lane <-> 32 consecutive bytes, 128-bit accesses
          float4* ag = reinterpret_cast<float4*>(rg + vi * 8);
          float4* ab = reinterpret_cast<float4*>(rb + vi * 8);
          float4 a0 = ag[0], a1 = ag[1], b0 = ab[0], b1 = ab[1];
          a0.x += pg[0]; a0.y += pg[1]; a0.z += pg[2]; a0.w += pg[3]; a1.x += pg[4]; a1.y += pg[5]; a1.z += pg[6]; a1.w += pg[7];
          b0.x += pb[0]; b0.y += pb[1]; b0.z += pb[2]; b0.w += pb[3]; b1.x += pb[4]; b1.y += pb[5]; b1.z += pb[6]; b1.w += pb[7];
          ag[0] = a0; ag[1] = a1; ab[0] = b0; ab[1] = b1;
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float xv[8], dyv[8], o[8];
        unpack8(xr[i], xv);
        unpack8(dr[i], dyv);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float xh = (xv[t] - mean) * rstd;
          const float gy = g[t] * dyv[t];
          o[t] = rstd * (gy - s1 - xh * s2);
        }
        if (dz) {
          float z[8];
          if (has_drop) {
            const uint32_t keep = dropout_keep8(philox, (static_cast<unsigned long long>(row) * D + vi * 8) >> 3, rng_stream, thr16);
#pragma unroll
            for (int t = 0; t < 8; ++t) z[t] = ((keep >> t) & 1u) ? o[t] * drop_scale : 0.f;
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) z[t] = o[t];
          }
          *reinterpret_cast<bf16x8*>(dz + row * lddz + vi * 8) = pack8(z);
        }
        if (dres) {
          float r[8];
          unpack8(*reinterpret_cast<const bf16x8*>(dres + row * lddres + vi * 8), r);
#pragma unroll
          for (int t = 0; t < 8; ++t) o[t] += r[t];
        }
        if (dx) *reinterpret_cast<bf16x8*>(dx + row * lddx + vi * 8) = pack8(o);
      }
    }
  }
  if constexpr (HAS_DG) {
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) { a += red[w * 2 * D + c]; b += red[w * 2 * D + D + c]; }
      atomicAdd(dgamma + c, a);
      atomicAdd(dbeta + c, b);
    }
  }
}

template <int VPL, bool HAS_DG>
int launch_ln2(const void* dy, long long lddy, const void* x, long long ldx, const float* mean, const float* rstd, const float* gamma,
               const void* dres, long long lddres, void* dx, long long lddx, void* dz, long long lddz, float* dgamma, float* dbeta,
               int rows, int D, float drop_p, const unsigned long long* seed, uint32_t rng_stream, cudaStream_t stream) {
  const size_t smem = HAS_DG ? sizeof(float) * kWarps * 2 * D : 0;
  static size_t configured = 0;
  if (smem > 48 * 1024 && configured < smem) {
    if (cudaFuncSetAttribute(ln_bwd2_kernel<VPL, HAS_DG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return PRISMER_ERR_CUDA;
    configured = smem;
  }
  int grid = (rows + kWarps - 1) / kWarps;
  const int cap = 148 * (HAS_DG ? 2 : 3) * 2;                  // two waves of the resident blocks
  if (grid > cap) grid = cap;
  const uint32_t thr16 = static_cast<uint32_t>(drop_p * 65536.0f + 0.5f);
  pdl_launch(ln_bwd2_kernel<VPL, HAS_DG>, dim3(grid), dim3(kWarps * 32), smem, stream,
      reinterpret_cast<const bf16*>(dy), lddy, reinterpret_cast<const bf16*>(x), ldx, mean, rstd, gamma,
      reinterpret_cast<const bf16*>(dres), lddres, reinterpret_cast<bf16*>(dx), lddx, reinterpret_cast<bf16*>(dz), lddz, dgamma, dbeta, rows,
      D, drop_p, thr16, seed, rng_stream);
  return LAUNCH_CHECK();
}

}  // namespace

int prismer_ln_bwd_lean(const void* dy, long long lddy, const void* x, long long ldx, const float* mean,
                                        const float* rstd, const float* gamma, const void* dres, long long lddres, void* dx,
                                        long long lddx, void* dz, long long lddz, float* dgamma, float* dbeta, int rows, int D,
                                        float drop_p, const unsigned long long* seed, uint32_t rng_stream, cudaStream_t stream) {
  const int vpl = (D / 8 + 31) / 32;      // arguments validated by the caller (prismer_layernorm_bwd)
#define LN2(V)                                                                                                                   \
  return dgamma ? launch_ln2<V, true>(dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dz, lddz, dgamma, dbeta, rows, D, \
                                      drop_p, seed, rng_stream, stream)                                                          \
                : launch_ln2<V, false>(dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dz, lddz, dgamma, dbeta, rows, D, \
                                       drop_p, seed, rng_stream, stream)
  switch (vpl) {
    case 1: LN2(1);
    case 2: LN2(2);
    case 3: LN2(3);
    default: LN2(4);
  }
#undef LN2
}
