// Small HBM-bound kernels: bias-gradient column sums, activation backward, casts, fused AdamW, token assembly.
#include "common.cuh"
#include "prismer_sm100.h"

namespace {

// ------------------------------------------------------------------ colsum: out[N] += sum_rows x[M,N]   (bias gradients)
// block = 256 threads = 32 column-vectors (8 bf16 each -> 256 columns) x 8 row lanes
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ out,
                                                     int M, int N, int rows_per_block) {
  __shared__ float red[8][256];
  PDL_GRID_SYNC();
  const int cv = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col0 = blockIdx.x * 256 + cv * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col0 < N) {
    const bool full = col0 + 8 <= N;
    for (int r = r0 + rl; r < r1; r += 8) {
      const bf16* p = x + static_cast<long long>(r) * ldx + col0;
      if (full) {
        float f[8]; unpack8(*reinterpret_cast<const bf16x8*>(p), f);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] += f[t];
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) if (col0 + t < N) acc[t] += __bfloat162float(p[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) red[rl][cv * 8 + t] = acc[t];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][c];
  if (blockIdx.x * 256 + c < N) atomicAdd(out + blockIdx.x * 256 + c, s);
}

// ------------------------------------------------------------------ dz = dy * act'(z)
__global__ void act_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ z, bf16* __restrict__ dz, long long n8,
                               int act) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float a[8], b[8];
    unpack8(reinterpret_cast<const bf16x8*>(dy)[i], a);
    unpack8(reinterpret_cast<const bf16x8*>(z)[i], b);
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] *= act_bwd(act, b[t]);
    reinterpret_cast<bf16x8*>(dz)[i] = pack8(a);
  }
}

// ------------------------------------------------------------------ y = dropout(x): element idx keyed Philox mask
// (embedding dropout roberta.py:75; the same call with dy gives its backward)
__global__ void dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8, float scale, uint32_t thr16,
                               const unsigned long long* __restrict__ seed, uint32_t stream) {
  const Philox ph(*seed);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float a[8];
    unpack8(reinterpret_cast<const bf16x8*>(x)[i], a);
    const uint32_t keep = dropout_keep8(ph, static_cast<unsigned long long>(i), stream, thr16);
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] = ((keep >> t) & 1u) ? a[t] * scale : 0.f;
    reinterpret_cast<bf16x8*>(y)[i] = pack8(a);
  }
}

// ------------------------------------------------------------------ fp32 -> bf16 cast (compute copies of the weights)
__global__ void cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = __float2bfloat16(src[(n4 << 2) + threadIdx.x]);
}

// ------------------------------------------------------------------ fused AdamW over a flat fp32 buffer
// torch.optim.AdamW semantics (train_caption.py:111-112): p *= 1 - lr*wd; m,v update; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// grad_scale folds the data-parallel average (1/world) into the update; also refreshes the bf16 compute copy.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, bf16* __restrict__ p16, long long n, float lr, float beta1,
                                                    float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  constexpr int U = 4;   // independent 128-bit streams per thread: 16 loads in flight before the first use
  for (long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 pv[U], gv[U], mv[U], vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        pv[u] = reinterpret_cast<const float4*>(p)[i]; gv[u] = reinterpret_cast<const float4*>(g)[i];
        mv[u] = reinterpret_cast<const float4*>(m)[i]; vv[u] = reinterpret_cast<const float4*>(v)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n4) continue;
      float pp[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w}, gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
      float mm[4] = {mv[u].x, mv[u].y, mv[u].z, mv[u].w}, vvv[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float gr = gg[t] * grad_scale;
        pp[t] *= (1.0f - lr * wd);
        mm[t] = beta1 * mm[t] + (1.0f - beta1) * gr;
        vvv[t] = beta2 * vvv[t] + (1.0f - beta2) * gr * gr;
        const float denom = sqrtf(vvv[t]) / bc2_sqrt + eps;
        pp[t] -= (lr / bc1) * (mm[t] / denom);
      }
      reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
      reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
      reinterpret_cast<float4*>(v)[i] = make_float4(vvv[0], vvv[1], vvv[2], vvv[3]);
      if (p16) {
        uint2 o; o.x = pack_bf162(pp[0], pp[1]); o.y = pack_bf162(pp[2], pp[3]);
        reinterpret_cast<uint2*>(p16)[i] = o;
      }
    }
  }
}

// ------------------------------------------------------------------ token assembly (vit.py:141-159)
// dst[b, n, :] = src[b*n_tok + n, :] + pos[n, :] (+ inst_emb[table[inst(b, nearest(n))], :])
// inst: int64 [B, 1, Hi, Wi] full-resolution instance map; nearest resize to gh x gw as F.interpolate(mode='nearest').
__global__ void assemble_kernel(const bf16* __restrict__ src, const bf16* __restrict__ pos, const long long* __restrict__ inst,
                                const int* __restrict__ table, const bf16* __restrict__ inst_emb, bf16* __restrict__ dst,
                                long long dst_bs, long long dst_rs, int B, int n_tok, int D, int gh, int gw, int Hi, int Wi) {
  const int vpr = D >> 3;
  const long long total = static_cast<long long>(B) * n_tok * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % vpr);
    const long long rn = i / vpr;
    const int n = static_cast<int>(rn % n_tok), b = static_cast<int>(rn / n_tok);
    float a[8], pz[8];
    unpack8(*reinterpret_cast<const bf16x8*>(src + rn * D + c * 8), a);
    unpack8(*reinterpret_cast<const bf16x8*>(pos + static_cast<long long>(n) * D + c * 8), pz);
    if (inst) {
      const int y = n / gw, x = n % gw;
      // nearest (legacy): src = floor(dst * in/out)
      const int sy = min(static_cast<int>(floorf(y * (static_cast<float>(Hi) / gh))), Hi - 1);
      const int sx = min(static_cast<int>(floorf(x * (static_cast<float>(Wi) / gw))), Wi - 1);
      const long long id = inst[(static_cast<long long>(b) * Hi + sy) * Wi + sx];
      const int row = table[id & 255];
      if (row >= 0) {
        float e[8];
        unpack8(*reinterpret_cast<const bf16x8*>(inst_emb + static_cast<long long>(row) * D + c * 8), e);
#pragma unroll
        for (int t = 0; t < 8; ++t) a[t] += e[t];
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] += pz[t];
    *reinterpret_cast<bf16x8*>(dst + b * dst_bs + static_cast<long long>(n) * dst_rs + c * 8) = pack8(a);
  }
}

// backward of the assembly for one modality: dsrc[b*n_tok+n,:] = ddst[b,n,:]   (a strided row copy)
__global__ void assemble_bwd_kernel(const bf16* __restrict__ ddst, long long ddst_bs, long long ddst_rs, bf16* __restrict__ dsrc, int B,
                                    int n_tok, int D) {
  const unsigned vpr = D >> 3;
  const unsigned total = static_cast<unsigned>(B) * n_tok * vpr;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c = i % vpr, rn = i / vpr, n = rn % n_tok, b = rn / n_tok;
    *reinterpret_cast<bf16x8*>(dsrc + static_cast<size_t>(rn) * D + c * 8) =
        *reinterpret_cast<const bf16x8*>(ddst + b * ddst_bs + static_cast<long long>(n) * ddst_rs + c * 8);
  }
}

// dinst_emb[table[inst(b, nearest(n))], :] += ddst[b, n, :]   (vit.py:147-150: the instance embedding added to the object-detection tokens).
// B * n_tok token rows scatter into the few table rows a batch uses (<= 128): one global atomic per element (the first version) puts
// hundreds of same-address atomics in a row.  Here a block owns IE_CH channels and a slice of the tokens, accumulates its
// [IE_ROWS x IE_CH] slab in shared memory and flushes only the non-zero entries.
constexpr int IE_ROWS = 128, IE_CH = 32, IE_THREADS = 256;
__global__ void __launch_bounds__(IE_THREADS) inst_emb_grad_kernel(const bf16* __restrict__ ddst, long long ddst_bs, long long ddst_rs,
                                                                   const long long* __restrict__ inst, const int* __restrict__ table,
                                                                   float* __restrict__ dinst_emb, int B, int n_tok, int D, int gh, int gw,
                                                                   int Hi, int Wi) {
  __shared__ float slab[IE_ROWS * IE_CH];
  for (int i = threadIdx.x; i < IE_ROWS * IE_CH; i += IE_THREADS) slab[i] = 0.f;
  __syncthreads();
  const int c0 = blockIdx.x * IE_CH;                       // first channel of this block
  const int cvi = threadIdx.x & 3, lane = threadIdx.x >> 2; // 4 channel vectors x 64 token lanes
  const int ntok = B * n_tok;
  const bool live = c0 + cvi * 8 < D;
  for (int tk = blockIdx.y * (IE_THREADS / 4) + lane; tk < ntok; tk += gridDim.y * (IE_THREADS / 4)) {
    const int n = tk % n_tok, b = tk / n_tok;
    const int y = n / gw, x = n % gw;
    // nearest (legacy): src = floor(dst * in/out), as in assemble_kernel
    const int sy = min(static_cast<int>(floorf(y * (static_cast<float>(Hi) / gh))), Hi - 1);
    const int sx = min(static_cast<int>(floorf(x * (static_cast<float>(Wi) / gw))), Wi - 1);
    const long long id = inst[(static_cast<long long>(b) * Hi + sy) * Wi + sx];
    const int row = table[id & 255];
    if (row < 0 || !live) continue;
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(ddst + b * ddst_bs + static_cast<long long>(n) * ddst_rs + c0 + cvi * 8), f);
    if (row < IE_ROWS) {
#pragma unroll
      for (int t = 0; t < 8; ++t) atomicAdd(&slab[row * IE_CH + cvi * 8 + t], f[t]);
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) atomicAdd(dinst_emb + static_cast<long long>(row) * D + c0 + cvi * 8 + t, f[t]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IE_ROWS * IE_CH; i += IE_THREADS) {
    const float v = slab[i];
    const int row = i / IE_CH, ch = c0 + i % IE_CH;
    if (v != 0.f && ch < D) atomicAdd(dinst_emb + static_cast<long long>(row) * D + ch, v);
  }
}

// dpos[n, :] += sum_b sum_{slots} dtok[b, slot_off[s] + n, :]   (positional embedding shared by all modalities)
__global__ void pos_grad_kernel(const bf16* __restrict__ dtok, long long bs, long long rs, int B, int n_tok, int D, int n_slots,
                                int slot_stride, float* __restrict__ dpos) {
  const int vpr = D >> 3;
  const int total = n_tok * vpr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % vpr, n = i / vpr;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < B; ++b)
      for (int s = 0; s < n_slots; ++s) {
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(dtok + b * bs + (static_cast<long long>(s) * slot_stride + n) * rs + c * 8), f);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] += f[t];
      }
#pragma unroll
    for (int t = 0; t < 8; ++t) dpos[static_cast<long long>(n) * D + c * 8 + t] += acc[t];
  }
}

// rows [B, n, D] <- broadcast of a [n, D] table (resampler latents, resampler.py:47)
__global__ void broadcast_rows_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long long dst_bs, long long dst_rs,
                                      int B, int n, int D) {
  const int vpr = D >> 3;
  const long long total = static_cast<long long>(B) * n * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % vpr);
    const long long rn = i / vpr;
    const int r = static_cast<int>(rn % n), b = static_cast<int>(rn / n);
    *reinterpret_cast<bf16x8*>(dst + b * dst_bs + static_cast<long long>(r) * dst_rs + c * 8) =
        *reinterpret_cast<const bf16x8*>(src + static_cast<long long>(r) * D + c * 8);
  }
}

// dsrc[n, :] (fp32, +=) = sum_b d[b, n, :]
__global__ void reduce_batch_kernel(const bf16* __restrict__ d, long long bs, long long rs, int B, int n, int D,
                                    float* __restrict__ out) {
  const int vpr = D >> 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * vpr; i += gridDim.x * blockDim.x) {
    const int c = i % vpr, r = i / vpr;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < B; ++b) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(d + b * bs + static_cast<long long>(r) * rs + c * 8), f);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] += f[t];
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) out[static_cast<long long>(r) * D + c * 8 + t] += acc[t];
  }
}

// strided 2-D copy of bf16 rows (16-byte vectors): dst[r, :] = src[r, :]
__global__ void copy_rows_kernel(const bf16* __restrict__ src, long long lds, bf16* __restrict__ dst, long long ldd,
                                 long long rows, int D, int add) {
  const int vpr = D >> 3;
  const long long total = rows * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % vpr);
    const long long r = i / vpr;
    bf16x8 v = *reinterpret_cast<const bf16x8*>(src + r * lds + c * 8);
    if (add) {
      float a[8], b[8];
      unpack8(v, a);
      unpack8(*reinterpret_cast<const bf16x8*>(dst + r * ldd + c * 8), b);
#pragma unroll
      for (int t = 0; t < 8; ++t) a[t] += b[t];
      v = pack8(a);
    }
    *reinterpret_cast<bf16x8*>(dst + r * ldd + c * 8) = v;
  }
}

// flags[id & 255] = 1 for every id present in the instance map (device side of ``instance.unique()``, vit.py:144)
__global__ void id_presence_kernel(const long long* __restrict__ ids, long long n, int* __restrict__ flags) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int id = static_cast<int>(ids[i] & 255);
    if (flags[id] == 0) flags[id] = 1;
  }
}

inline int blocks_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = 148 * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int prismer_colsum(const void* x, long long ldx, float* out, int M, int N, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return PRISMER_OK;
  if (ldx % 8) return PRISMER_ERR_ALIGN;
  const int gx = (N + 255) / 256;
  int gy = (148 * 4 + gx - 1) / gx;
  int rpb = (M + gy - 1) / gy;
  rpb = ((rpb + 7) / 8) * 8;
  gy = (M + rpb - 1) / rpb;
  pdl_launch(colsum_kernel, dim3(gx, gy), dim3(256), 0, stream, reinterpret_cast<const bf16*>(x), ldx, out, M, N, rpb);
  return LAUNCH_CHECK();
}

extern "C" int prismer_act_bwd(const void* dy, const void* z, void* dz, long long n, int act, cudaStream_t stream) {
  if (n <= 0) return PRISMER_OK;
  if (n % 8) return PRISMER_ERR_SHAPE;
  act_bwd_kernel<<<blocks_for(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(z),
                                                            reinterpret_cast<bf16*>(dz), n / 8, act);
  return LAUNCH_CHECK();
}

extern "C" int prismer_cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
  if (n <= 0) return PRISMER_OK;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 7)) return PRISMER_ERR_ALIGN;
  cast_kernel<<<blocks_for(n / 4 + 1, 256), 256, 0, stream>>>(src, reinterpret_cast<bf16*>(dst), n);
  return LAUNCH_CHECK();
}

extern "C" int prismer_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr,
                                  float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                  cudaStream_t stream) {
  if (n <= 0) return PRISMER_OK;
  if (n % 4 || step < 1) return PRISMER_ERR_SHAPE;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<blocks_for(n / 4, 256), 256, 0, stream>>>(p, g, m, v, reinterpret_cast<bf16*>(p_bf16), n, lr, beta1, beta2,
                                                          eps, weight_decay, bc1, sqrtf(bc2), grad_scale);
  return LAUNCH_CHECK();
}

extern "C" int prismer_assemble_tokens(const void* src, const void* pos, const void* inst, const int* table,
                                       const void* inst_emb, void* dst, long long dst_bs, long long dst_rs, int B, int n_tok,
                                       int D, int gh, int gw, int Hi, int Wi, cudaStream_t stream) {
  if (D % 8 || gh * gw != n_tok) return PRISMER_ERR_SHAPE;
  const long long total = static_cast<long long>(B) * n_tok * (D / 8);
  assemble_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(src), reinterpret_cast<const bf16*>(pos), reinterpret_cast<const long long*>(inst), table,
      reinterpret_cast<const bf16*>(inst_emb), reinterpret_cast<bf16*>(dst), dst_bs, dst_rs, B, n_tok, D, gh, gw, Hi, Wi);
  return LAUNCH_CHECK();
}

extern "C" int prismer_assemble_tokens_bwd(const void* ddst, long long ddst_bs, long long ddst_rs, void* dsrc, const void* inst,
                                           const int* table,
                                           float* dinst_emb, int B, int n_tok, int D, int gh, int gw, int Hi, int Wi,
                                           cudaStream_t stream) {
  if (D % 8 || gh * gw != n_tok) return PRISMER_ERR_SHAPE;
  const long long total = static_cast<long long>(B) * n_tok * (D / 8);
  if (total >= (1ll << 31)) return PRISMER_ERR_SHAPE;
  assemble_bwd_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(ddst), ddst_bs, ddst_rs,
                                                                  reinterpret_cast<bf16*>(dsrc), B, n_tok, D);
  if (inst && dinst_emb) {
    const int ntok = B * n_tok;
    int ysplit = (ntok + 64 * 16 - 1) / (64 * 16);          // >= 16 tokens per thread before another slice of the tokens pays
    ysplit = ysplit < 1 ? 1 : (ysplit > 16 ? 16 : ysplit);
    inst_emb_grad_kernel<<<dim3((D + IE_CH - 1) / IE_CH, ysplit), IE_THREADS, 0, stream>>>(
        reinterpret_cast<const bf16*>(ddst), ddst_bs, ddst_rs, reinterpret_cast<const long long*>(inst), table, dinst_emb, B, n_tok, D, gh,
        gw, Hi, Wi);
  }
  return LAUNCH_CHECK();
}

extern "C" int prismer_pos_grad(const void* dtok, long long bs, long long rs, int B, int n_tok, int D, int n_slots, int slot_stride,
                                float* dpos, cudaStream_t stream) {
  if (D % 8) return PRISMER_ERR_SHAPE;
  pos_grad_kernel<<<blocks_for(static_cast<long long>(n_tok) * (D / 8), 128), 128, 0, stream>>>(
      reinterpret_cast<const bf16*>(dtok), bs, rs, B, n_tok, D, n_slots, slot_stride, dpos);
  return LAUNCH_CHECK();
}

extern "C" int prismer_broadcast_rows(const void* src, void* dst, long long dst_bs, long long dst_rs, int B, int n, int D,
                                      cudaStream_t stream) {
  if (D % 8) return PRISMER_ERR_SHAPE;
  broadcast_rows_kernel<<<blocks_for(static_cast<long long>(B) * n * (D / 8), 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(src), reinterpret_cast<bf16*>(dst), dst_bs, dst_rs, B, n, D);
  return LAUNCH_CHECK();
}

extern "C" int prismer_reduce_batch(const void* d, long long bs, long long rs, int B, int n, int D, float* out,
                                    cudaStream_t stream) {
  if (D % 8) return PRISMER_ERR_SHAPE;
  reduce_batch_kernel<<<blocks_for(static_cast<long long>(n) * (D / 8), 128), 128, 0, stream>>>(
      reinterpret_cast<const bf16*>(d), bs, rs, B, n, D, out);
  return LAUNCH_CHECK();
}

extern "C" int prismer_copy_rows(const void* src, long long lds, void* dst, long long ldd, long long rows, int D, int add,
                                 cudaStream_t stream) {
  if (rows <= 0) return PRISMER_OK;
  if (D % 8 || lds % 8 || ldd % 8) return PRISMER_ERR_SHAPE;
  copy_rows_kernel<<<blocks_for(rows * (D / 8), 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(src), lds,
                                                                       reinterpret_cast<bf16*>(dst), ldd, rows, D, add);
  return LAUNCH_CHECK();
}

extern "C" int prismer_id_presence(const void* ids, long long n, int* flags, cudaStream_t stream) {
  if (cudaMemsetAsync(flags, 0, 256 * sizeof(int), stream) != cudaSuccess) return PRISMER_ERR_CUDA;
  if (n <= 0) return PRISMER_OK;
  id_presence_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(reinterpret_cast<const long long*>(ids), n, flags);
  return LAUNCH_CHECK();
}

extern "C" int prismer_dropout(const void* x, void* y, long long n, float p, const unsigned long long* seed, uint32_t rng_stream,
                               cudaStream_t stream) {
  if (n <= 0) return PRISMER_OK;
  if (n % 8 || !seed || p < 0.f || p >= 1.f) return PRISMER_ERR_SHAPE;
  dropout_kernel<<<blocks_for(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), n / 8,
                                                            1.0f / (1.0f - p), static_cast<uint32_t>(p * 65536.0f + 0.5f), seed,
                                                            rng_stream);
  return LAUNCH_CHECK();
}
