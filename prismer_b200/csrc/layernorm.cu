// LayerNorm forward / backward (fp32 statistics, bf16 I/O): one warp per row, 128-bit loads, warp-shuffle reductions.
// HBM-bound: algorithmic bytes = rows*D*2 (read) + rows*D*2 (write) forward.
// Replaces model/modules/utils.py:14-19 (LayerNorm.forward: fp32 layer_norm, eps 1e-5, cast back) at every site listed in
// SURVEY.md section 2.3 row K5, and its autograd backward.
#include "common.cuh"
#include "prismer_sm100.h"

namespace {

constexpr int kWarpsPerBlock = 8;

template <int VPL>  // 16-byte vectors (8 bf16) per lane; D <= VPL*256
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_fwd_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
              bf16* __restrict__ y, long long ldy, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
              int D, float eps) {
  PDL_GRID_SYNC();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  for (long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * kWarpsPerBlock) {
    const bf16* xr = x + row * ldx;
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        unpack8(*reinterpret_cast<const bf16x8*>(xr + vi * 8), v[i]);
#pragma unroll
        for (int t = 0; t < 8; ++t) s += v[i][t];
      }
    }
    const float mean = warp_sum(s) / D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int t = 0; t < 8; ++t) { const float d = v[i][t] - mean; ss += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / D + eps);
    if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
    bf16* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float o[8];
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = (v[i][t] - mean) * rstd * g[t] + b[t];
        *reinterpret_cast<bf16x8*>(yr + vi * 8) = pack8(o);
      }
    }
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) [+ dres];   dgamma += sum_rows dy*xhat; dbeta += sum_rows dy
// optional second output dz = dropout_mask(dx) * scale  (gradient of the branch that went through dropout before the
// residual add; roberta.py:134-140,177-183).
template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_bwd_kernel(const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ x, long long ldx,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
              const bf16* __restrict__ dres, long long lddres, bf16* __restrict__ dx, long long lddx,
              bf16* __restrict__ dz, long long lddz, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows,
              int D, float drop_p, uint32_t thr16, const unsigned long long* seed, uint32_t rng_stream) {
  extern __shared__ float red[];  // [kWarpsPerBlock][2][D] only when dgamma != nullptr
  PDL_GRID_SYNC();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  float dg[VPL][8], db[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int t = 0; t < 8; ++t) { dg[i][t] = 0.f; db[i][t] = 0.f; }
  const bool has_drop = dz != nullptr && drop_p > 0.f;
  const Philox philox(has_drop ? *seed : 0ull);
  const float drop_scale = has_drop ? 1.0f / (1.0f - drop_p) : 1.0f;

  for (long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * kWarpsPerBlock) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[VPL][8], gy[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float xv[8], dyv[8];
        unpack8(*reinterpret_cast<const bf16x8*>(x + row * ldx + vi * 8), xv);
        unpack8(*reinterpret_cast<const bf16x8*>(dy + row * lddy + vi * 8), dyv);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          xh[i][t] = (xv[t] - mean) * rstd;
          gy[i][t] = g[t] * dyv[t];
          s1 += gy[i][t];
          s2 += gy[i][t] * xh[i][t];
          dg[i][t] += dyv[t] * xh[i][t];
          db[i][t] += dyv[t];
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = rstd * (gy[i][t] - s1 - xh[i][t] * s2);
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
  if (dgamma) {
    float* rg = red + warp * 2 * D;
    float* rb = rg + D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
#pragma unroll
        for (int t = 0; t < 8; ++t) { rg[vi * 8 + t] = dg[i][t]; rb[vi * 8 + t] = db[i][t]; }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) { a += red[w * 2 * D + c]; b += red[w * 2 * D + D + c]; }
      atomicAdd(dgamma + c, a);
      atomicAdd(dbeta + c, b);
    }
  }
}

int grid_for(int rows) {
  int blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int cap = 148 * 8;
  return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
}

}  // namespace

extern "C" int prismer_layernorm_fwd(const void* x, long long ldx, const float* gamma, const float* beta, void* y,
                                     long long ldy, float* mean, float* rstd, int rows, int D, float eps,
                                     cudaStream_t stream) {
  if (rows <= 0) return PRISMER_OK;
  if (D <= 0 || (D % 8) || D > 2048 || (ldx % 8) || (ldy % 8)) return PRISMER_ERR_SHAPE;
  const int vpl = (D / 8 + 31) / 32;
  const bf16* xp = reinterpret_cast<const bf16*>(x);
  bf16* yp = reinterpret_cast<bf16*>(y);
  const int grid = grid_for(rows), block = kWarpsPerBlock * 32;
#define LN_FWD(V) pdl_launch(ln_fwd_kernel<V>, dim3(grid), dim3(block), 0, stream, xp, ldx, gamma, beta, yp, ldy, mean, rstd, rows, D, eps)
  switch (vpl) {
    case 1: LN_FWD(1); break;
    case 2: LN_FWD(2); break;
    case 3: LN_FWD(3); break;
    case 4: LN_FWD(4); break;
    case 5: LN_FWD(5); break;
    case 6: LN_FWD(6); break;
    case 7: LN_FWD(7); break;
    default: LN_FWD(8); break;
  }
#undef LN_FWD
  return LAUNCH_CHECK();
}

// register-lean kernel for D <= 1024 (layernorm_v2.cu)
int prismer_ln_bwd_lean(const void* dy, long long lddy, const void* x, long long ldx, const float* mean, const float* rstd,
                        const float* gamma, const void* dres, long long lddres, void* dx, long long lddx, void* dz, long long lddz,
                        float* dgamma, float* dbeta, int rows, int D, float drop_p, const unsigned long long* seed, uint32_t rng_stream,
                        cudaStream_t stream);

extern "C" int prismer_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* mean,
                                     const float* rstd, const float* gamma, const void* dres, long long lddres, void* dx,
                                     long long lddx, void* dz, long long lddz, float* dgamma, float* dbeta, int rows,
                                     int D, float drop_p, const unsigned long long* seed, uint32_t rng_stream,
                                     cudaStream_t stream) {
  if (rows <= 0) return PRISMER_OK;
  if (D <= 0 || (D % 8) || D > 2048 || (ldx % 8) || (lddy % 8) || (lddx % 8)) return PRISMER_ERR_SHAPE;
  if ((dgamma == nullptr) != (dbeta == nullptr)) return PRISMER_ERR_SHAPE;
  if (dz && drop_p > 0.f && !seed) return PRISMER_ERR_SHAPE;
  if (D <= 1024)
    return prismer_ln_bwd_lean(dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dz, lddz, dgamma, dbeta, rows, D, drop_p, seed,
                               rng_stream, stream);
  const int vpl = (D / 8 + 31) / 32;
  int grid = grid_for(rows);
  if (dgamma && grid > 148 * 4) grid = 148 * 4;  // fewer, fatter blocks -> fewer atomics
  const int block = kWarpsPerBlock * 32;
  const size_t smem = dgamma ? sizeof(float) * kWarpsPerBlock * 2 * D : 0;
  const uint32_t thr16 = static_cast<uint32_t>(drop_p * 65536.0f + 0.5f);
#define LN_BWD(V)                                                                                                    \
  do {                                                                                                               \
    static int configured_##V = 0;                                                                                   \
    if (smem > 48 * 1024 && configured_##V < (int)smem) {                                                            \
      cudaFuncSetAttribute(ln_bwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                \
      configured_##V = (int)smem;                                                                                    \
    }                                                                                                                \
    pdl_launch(ln_bwd_kernel<V>, dim3(grid), dim3(block), smem, stream,                                              \
        reinterpret_cast<const bf16*>(dy), lddy, reinterpret_cast<const bf16*>(x), ldx, mean, rstd, gamma,           \
        reinterpret_cast<const bf16*>(dres), lddres, reinterpret_cast<bf16*>(dx), lddx, reinterpret_cast<bf16*>(dz), \
        lddz, dgamma, dbeta, rows, D, drop_p, thr16, seed, rng_stream);                                              \
  } while (0)
  switch (vpl) {
    case 1: LN_BWD(1); break;
    case 2: LN_BWD(2); break;
    case 3: LN_BWD(3); break;
    case 4: LN_BWD(4); break;
    case 5: LN_BWD(5); break;
    case 6: LN_BWD(6); break;
    case 7: LN_BWD(7); break;
    default: LN_BWD(8); break;
  }
#undef LN_BWD
  return LAUNCH_CHECK();
}
