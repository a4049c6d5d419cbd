// Decoder input embeddings and the label-smoothed LM loss (HBM-bound gather / row-reduction kernels).
//   embed      : roberta.py:38-45,66-72  pos_id = cumsum(id != pad) * (id != pad) + pad;  e = word[id] + type[0] + pos[pos_id]
//   embed_bwd  : scatter-add into the fp32 gradient tables (padding_idx rows receive no gradient, roberta.py:51,64)
//   ce_fwd/bwd : roberta.py:381-387  shift; CrossEntropyLoss(reduction='none', label_smoothing=0.1, ignore_index=-100);
//                per-sample sum over T; batch mean (prismer_caption.py:33) with optional per-sample weights (prismer_vqa.py:40)
//   argmax     : greedy decoding step (logits[:, -1] -> MinLength processor -> argmax; lowest index wins ties like torch.argmax)
#include "common.cuh"
#include "prismer_sm100.h"

namespace {

__global__ void __launch_bounds__(128) embed_fwd_kernel(const long long* __restrict__ ids, const bf16* __restrict__ word,
                                                        const bf16* __restrict__ pos, const bf16* __restrict__ type,
                                                        bf16* __restrict__ out, int* __restrict__ pos_ids_out, int B, int T,
                                                        int H, int pad_id, int past_len) {
  // one warp per token
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * T) return;
  const int b = warp / T, t = warp % T;
  int cnt = 0;
  for (int j = lane; j <= t; j += 32) cnt += ids[static_cast<long long>(b) * T + j] != pad_id;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const long long id = ids[static_cast<long long>(b) * T + t];
  const int pid = (id != pad_id) ? cnt + past_len + pad_id : pad_id;
  if (lane == 0 && pos_ids_out) pos_ids_out[warp] = pid;
  for (int c = lane; c < (H >> 3); c += 32) {
    float w[8], p[8], ty[8];
    unpack8(*reinterpret_cast<const bf16x8*>(word + id * H + c * 8), w);
    unpack8(*reinterpret_cast<const bf16x8*>(pos + static_cast<long long>(pid) * H + c * 8), p);
    unpack8(*reinterpret_cast<const bf16x8*>(type + c * 8), ty);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = (w[k] + ty[k]) + p[k];
    *reinterpret_cast<bf16x8*>(out + static_cast<long long>(warp) * H + c * 8) = pack8(w);
  }
}

__global__ void __launch_bounds__(128) embed_bwd_kernel(const bf16* __restrict__ de, const long long* __restrict__ ids,
                                                        const int* __restrict__ pos_ids, float* __restrict__ dword,
                                                        float* __restrict__ dpos, float* __restrict__ dtype, int rows, int H,
                                                        int pad_id) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long id = ids[warp];
  const int pid = pos_ids[warp];
  for (int c = lane; c < (H >> 3); c += 32) {
    float g[8];
    unpack8(*reinterpret_cast<const bf16x8*>(de + static_cast<long long>(warp) * H + c * 8), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (dword && id != pad_id) atomicAdd(dword + id * H + c * 8 + k, g[k]);
      if (dpos && pid != pad_id) atomicAdd(dpos + static_cast<long long>(pid) * H + c * 8 + k, g[k]);
      if (dtype) atomicAdd(dtype + c * 8 + k, g[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- cross entropy
constexpr int CE_THREADS = 512;

__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : (is_max ? -INFINITY : 0.f);
  if (w == 0) {
    r = is_max ? warp_max(r) : warp_sum(r);
    if (lane == 0) sh[0] = r;
  }
  __syncthreads();
  return sh[0];
}

// one block per (b, t) row of logits; label = labels[b, t+1] (shift), rows t == T-1 are skipped.
// ONE pass over the row (the rows of a [960, 50265] fp32 buffer do not all fit in L2, a second pass re-reads most of them from HBM):
// every thread keeps a running (max, sum of exp relative to that max, plain sum); VEC: 128-bit loads when the rows are 16-byte aligned.
template <bool VEC>
__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const float* __restrict__ logits, long long ld,
                                                            const long long* __restrict__ labels, float* __restrict__ row_loss,
                                                            float* __restrict__ row_lse, int T, int V, float smoothing) {
  __shared__ float sh[32];
  const int r = blockIdx.x, t = r % T;
  const float* x = logits + static_cast<long long>(r) * ld;
  const long long lab = (t < T - 1) ? labels[r + 1] : -100;
  float mx = -INFINITY, se = 0.f, sx = 0.f;
  if (VEC) {
    const int nv = V >> 2;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      const float m2 = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      se = se * expf(mx - m2) + expf(v.x - m2) + expf(v.y - m2) + expf(v.z - m2) + expf(v.w - m2);
      sx += (v.x + v.y) + (v.z + v.w);
      mx = m2;
    }
    const int i = (nv << 2) + threadIdx.x;                       // <= 3 tail elements
    if (i < V) {
      const float v = x[i], m2 = fmaxf(mx, v);
      se = se * expf(mx - m2) + expf(v - m2);
      sx += v;
      mx = m2;
    }
  } else {
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      const float v = x[i], m2 = fmaxf(mx, v);
      se = se * expf(mx - m2) + expf(v - m2);
      sx += v;
      mx = m2;
    }
  }
  const float bm = block_reduce(mx, sh, true);
  se = block_reduce(mx == -INFINITY ? 0.f : se * expf(mx - bm), sh, false);
  sx = block_reduce(sx, sh, false);
  if (threadIdx.x == 0) {
    const float lse = bm + logf(se);
    row_lse[r] = lse;
    float loss = 0.f;
    if (lab >= 0) loss = (1.0f - smoothing) * (lse - x[lab]) + smoothing * (lse - sx / V);
    row_loss[r] = loss;
  }
}

// per-sample sums (roberta.py:387) and the weighted batch mean (prismer_caption.py:33 / prismer_vqa.py:40-41)
__global__ void ce_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ weights,
                                 float* __restrict__ sample_loss, float* __restrict__ mean_loss, int B, int T) {
  __shared__ float sh[32];
  float tot = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += row_loss[b * T + t];
    sample_loss[b] = s;
    tot += (weights ? weights[b] : 1.0f) * s;
  }
  tot = block_reduce(tot, sh, false);
  if (threadIdx.x == 0) *mean_loss = tot / B;
}

// dlogits[b,t,:] = g_b * (softmax - (1-eps)*onehot - eps/V), g_b = gscale * weight_b / B ; zero for ignored rows; bf16 out,
// padded columns [V, ldo) are zeroed so the buffer can feed the dgrad / wgrad GEMMs directly.  VEC: 4 columns per thread and
// iteration (128-bit loads, 64-bit stores) when ld and ldo are multiples of 4 and both buffers 16 / 8-byte aligned.
template <bool VEC>
__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(const float* __restrict__ logits, long long ld,
                                                            const long long* __restrict__ labels, const float* __restrict__ row_lse,
                                                            const float* __restrict__ weights, const float* __restrict__ gscale,
                                                            bf16* __restrict__ dlogits, long long ldo, int B, int T, int V,
                                                            float smoothing) {
  const int r = blockIdx.x, t = r % T, b = r / T;
  const float* x = logits + static_cast<long long>(r) * ld;
  bf16* d = dlogits + static_cast<long long>(r) * ldo;
  const long long lab = (t < T - 1) ? labels[r + 1] : -100;
  const int n = static_cast<int>(ldo);
  if (lab < 0) {
    if (VEC) {
      for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) reinterpret_cast<uint2*>(d)[i] = make_uint2(0u, 0u);
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = __float2bfloat16(0.f);
    }
    return;
  }
  const float g = (gscale ? *gscale : 1.0f) * (weights ? weights[b] : 1.0f) / B;
  const float lse = row_lse[r];
  const float off = smoothing / V;
  const int ilab = static_cast<int>(lab);
  if (VEC) {
    for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) {
      const int c = i << 2;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (c + 3 < V) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(x) + i);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (expf(v[k] - lse) - off - (c + k == ilab ? 1.0f - smoothing : 0.f)) * g;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c + k < V) v[k] = (expf(x[c + k] - lse) - off - (c + k == ilab ? 1.0f - smoothing : 0.f)) * g;
      }
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
      reinterpret_cast<uint2*>(d)[i] = make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float v = 0.f;
      if (i < V) v = (expf(x[i] - lse) - off - (i == ilab ? 1.0f - smoothing : 0.f)) * g;
      d[i] = __float2bfloat16(v);
    }
  }
}

// greedy step: out[b] = argmax_v(logits[b*row_stride_rows + ...]) with eos suppressed when suppress_eos != 0
__global__ void __launch_bounds__(CE_THREADS) argmax_kernel(const float* __restrict__ logits, long long ld, int V,
                                                            int suppress_eos, int eos, long long* __restrict__ out) {
  __shared__ float sv[CE_THREADS / 32];
  __shared__ int si[CE_THREADS / 32];
  const float* x = logits + static_cast<long long>(blockIdx.x) * ld;
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float v = x[i];
    if (suppress_eos && i == eos) v = -INFINITY;
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < CE_THREADS / 32; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    out[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
  }
}

}  // namespace

extern "C" int prismer_embed_fwd(const void* ids, const void* word, const void* pos, const void* type, void* out, int* pos_ids,
                                 int B, int T, int H, int pad_id, int past_len, cudaStream_t stream) {
  if (B <= 0 || T <= 0) return PRISMER_OK;
  if (H % 8) return PRISMER_ERR_SHAPE;
  const int warps = B * T;
  embed_fwd_kernel<<<(warps + 3) / 4, 128, 0, stream>>>(reinterpret_cast<const long long*>(ids), reinterpret_cast<const bf16*>(word),
                                                      reinterpret_cast<const bf16*>(pos), reinterpret_cast<const bf16*>(type),
                                                      reinterpret_cast<bf16*>(out), pos_ids, B, T, H, pad_id, past_len);
  return LAUNCH_CHECK();
}

extern "C" int prismer_embed_bwd(const void* de, const void* ids, const int* pos_ids, float* dword, float* dpos, float* dtype,
                                 int rows, int H, int pad_id, cudaStream_t stream) {
  if (rows <= 0) return PRISMER_OK;
  if (H % 8) return PRISMER_ERR_SHAPE;
  embed_bwd_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(reinterpret_cast<const bf16*>(de), reinterpret_cast<const long long*>(ids),
                                                     pos_ids, dword, dpos, dtype, rows, H, pad_id);
  return LAUNCH_CHECK();
}

extern "C" int prismer_ce_loss_fwd(const float* logits, long long ld, const void* labels, const float* weights, float* row_loss,
                                   float* row_lse, float* sample_loss, float* mean_loss, int B, int T, int V, float smoothing,
                                   cudaStream_t stream) {
  if (B <= 0 || T <= 0 || V <= 0) return PRISMER_ERR_SHAPE;
  const bool vec = ld % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  if (vec)
    ce_fwd_kernel<true><<<B * T, CE_THREADS, 0, stream>>>(logits, ld, reinterpret_cast<const long long*>(labels), row_loss, row_lse, T, V,
                                                        smoothing);
  else
    ce_fwd_kernel<false><<<B * T, CE_THREADS, 0, stream>>>(logits, ld, reinterpret_cast<const long long*>(labels), row_loss, row_lse, T, V,
                                                         smoothing);
  ce_reduce_kernel<<<1, 256, 0, stream>>>(row_loss, weights, sample_loss, mean_loss, B, T);
  return LAUNCH_CHECK();
}

extern "C" int prismer_ce_loss_bwd(const float* logits, long long ld, const void* labels, const float* row_lse,
                                   const float* weights, const float* gscale, void* dlogits, long long ldo, int B, int T, int V,
                                   float smoothing, cudaStream_t stream) {
  if (B <= 0 || T <= 0 || V <= 0 || ldo < V) return PRISMER_ERR_SHAPE;
  const bool vec = ld % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 7) == 0;
  if (vec)
    ce_bwd_kernel<true><<<B * T, CE_THREADS, 0, stream>>>(logits, ld, reinterpret_cast<const long long*>(labels), row_lse, weights, gscale,
                                                        reinterpret_cast<bf16*>(dlogits), ldo, B, T, V, smoothing);
  else
    ce_bwd_kernel<false><<<B * T, CE_THREADS, 0, stream>>>(logits, ld, reinterpret_cast<const long long*>(labels), row_lse, weights, gscale,
                                                         reinterpret_cast<bf16*>(dlogits), ldo, B, T, V, smoothing);
  return LAUNCH_CHECK();
}

extern "C" int prismer_argmax(const float* logits, long long ld, int rows, int V, int suppress_eos, int eos, void* out,
                              cudaStream_t stream) {
  if (rows <= 0) return PRISMER_OK;
  argmax_kernel<<<rows, CE_THREADS, 0, stream>>>(logits, ld, V, suppress_eos, eos, reinterpret_cast<long long*>(out));
  return LAUNCH_CHECK();
}
