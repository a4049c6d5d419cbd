// Conv-stem support kernels (all HBM-bound; the contractions themselves run on the tcgen05 GEMM):
//   patchify / im2col (k in {1,3}, fused BatchNorm-affine + ReLU of the previous layer on load), bilinear resample of the
//   fp32 NCHW expert channel stack to bf16 NHWC, BatchNorm batch statistics + running-stat update, the BatchNorm/ReLU
//   backward fused with col2im, and the conv-weight layout transforms.
// Reference: model/modules/vit.py:86-120 (nn.Conv2d / nn.UpsamplingBilinear2d / nn.BatchNorm2d / nn.ReLU stems).
// Activations are NHWC bf16 [B*H*W, C]; GEMM K order is (kh, kw, c) so every gather is a 16-byte vector access.
#include "common.cuh"
#include "prismer_sm100.h"

namespace {

inline int blocks_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = 148 * 32;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

// grid for the (lanes x channel-vector) stem kernels: enough blocks to cover ``units`` (pixels / slots) at ``per_thread`` units per
// thread and iteration, capped at 4 resident-block waves of the 148 SMs (grid-stride loops take the rest)
inline int stem_grid(long long units, int C, int per_thread) {
  const int lanes = 384 / (C / 8);
  long long b = (units + static_cast<long long>(lanes) * per_thread - 1) / (static_cast<long long>(lanes) * per_thread);
  const long long cap = 148 * 4 * 2;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------------- rgb patchify (vit.py:86)
// x fp32 NCHW [B,3,R,R] -> out bf16 [B*g*g, Kpad], K order (c, kh, kw) == nn.Conv2d weight.flatten(1); pad columns zero.
__global__ void patchify_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B, int Cin, int R, int p, int g, int K,
                                int Kpad) {
  const int kv = Kpad >> 3;
  const long long total = static_cast<long long>(B) * g * g * kv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kc = static_cast<int>(i % kv);
    const long long row = i / kv;
    const int px = static_cast<int>(row % g), py = static_cast<int>((row / g) % g), b = static_cast<int>(row / (g * g));
    float v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int k = kc * 8 + t;
      float val = 0.f;
      if (k < K) {
        const int c = k / (p * p), kh = (k / p) % p, kw = k % p;
        val = x[((static_cast<long long>(b) * Cin + c) * R + py * p + kh) * R + px * p + kw];
      }
      v[t] = val;
    }
    *reinterpret_cast<bf16x8*>(out + row * Kpad + kc * 8) = pack8(v);
  }
}

// ---------------------------------------------------------------------------------------------- bilinear resample
// nn.UpsamplingBilinear2d == bilinear, align_corners=True (vit.py:89,106).  fp32 NCHW [B,C,Hi,Wi] -> bf16 NHWC [B,Ho,Wo,C].
// One block per (b, output row): stages the two needed input rows for a chunk of 16 channels in smem with full-row
// coalesced 128-bit loads (the "channel stack" read that bounds this kernel), then writes 32-byte channel runs.
constexpr int RS_CH = 16;
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B, int C, int Hi,
                                                       int Wi, int Ho, int Wo) {
  extern __shared__ float srow[];  // [2][RS_CH][Wi]
  const int y = blockIdx.x % Ho, b = blockIdx.x / Ho;
  const float sy = Ho > 1 ? static_cast<float>(Hi - 1) / (Ho - 1) : 0.f;
  const float sx = Wo > 1 ? static_cast<float>(Wi - 1) / (Wo - 1) : 0.f;
  const float fy = y * sy;
  const int y0 = min(static_cast<int>(fy), Hi - 1), y1 = min(y0 + 1, Hi - 1);
  const float wy = fy - y0;
  for (int c0 = 0; c0 < C; c0 += RS_CH) {
    const int nc = min(RS_CH, C - c0);
    __syncthreads();
    const int vec_per_row = Wi >> 2;  // Wi % 4 == 0 enforced on host
    for (int i = threadIdx.x; i < 2 * nc * vec_per_row; i += blockDim.x) {
      const int v = i % vec_per_row, c = (i / vec_per_row) % nc, r = i / (vec_per_row * nc);
      const float4 val = *reinterpret_cast<const float4*>(
          x + ((static_cast<long long>(b) * C + c0 + c) * Hi + (r ? y1 : y0)) * Wi + v * 4);
      *reinterpret_cast<float4*>(srow + (r * RS_CH + c) * Wi + v * 4) = val;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Wo * nc; i += blockDim.x) {
      const int c = i % nc, xo = i / nc;
      const float fx = xo * sx;
      const int x0 = min(static_cast<int>(fx), Wi - 1), x1 = min(x0 + 1, Wi - 1);
      const float wx = fx - x0;
      const float* r0 = srow + c * Wi;
      const float* r1 = srow + (RS_CH + c) * Wi;
      const float top = r0[x0] + (r0[x1] - r0[x0]) * wx;
      const float bot = r1[x0] + (r1[x1] - r1[x0]) * wx;
      out[((static_cast<long long>(b) * Ho + y) * Wo + xo) * C + c0 + c] = __float2bfloat16(top + (bot - top) * wy);
    }
  }
}

// ---------------------------------------------------------------------------------------------- im2col
// First layer: generic strided fp32 / bf16 input with few channels (1 or 3); K = k*k*Cin padded to Kpad; per-element gather.
// The (c, kh, kw) decomposition of the <= 64 K indices comes from a small shared-memory table instead of three divisions per element;
// index arithmetic is 32-bit (the input is a few MB and stays in L2, the kernel is bound by its 13-26 MB of output).
__global__ void __launch_bounds__(256) im2col_first_kernel(const void* __restrict__ in, int in_is_bf16, long long sb, long long sc,
                                                           long long sy, long long sx, bf16* __restrict__ out, int B, int Cin, int H, int W,
                                                           int ksz, int stride, int Ho, int Wo, int Kpad) {
  __shared__ int lut[64];                       // k -> c | kh << 8 | kw << 16, or -1 for the zero padding columns
  const int pad = ksz == 3 ? 1 : 0;
  const int K = ksz * ksz * Cin;
  const unsigned kv = Kpad >> 3;
  for (int k = threadIdx.x; k < 64; k += blockDim.x)
    lut[k] = (k < K) ? ((k % Cin) | ((k / (Cin * ksz)) << 8) | (((k / Cin) % ksz) << 16)) : -1;
  __syncthreads();
  const unsigned total = static_cast<unsigned>(B) * Ho * Wo * kv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned kc = i % kv, row = i / kv;
    const unsigned xo = row % Wo, t2 = row / Wo, yo = t2 % Ho, b = t2 / Ho;
    const int y0 = static_cast<int>(yo) * stride - pad, x0 = static_cast<int>(xo) * stride - pad;
    const long long base = b * sb;
    float v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int e = lut[kc * 8 + t];
      float val = 0.f;
      if (e >= 0) {
        const int c = e & 255, yi = y0 + ((e >> 8) & 255), xi = x0 + (e >> 16);
        if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
          const long long off = base + c * sc + yi * sy + xi * sx;
          val = in_is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(in)[off]) : __ldg(reinterpret_cast<const float*>(in) + off);
        }
      }
      v[t] = val;
    }
    *reinterpret_cast<bf16x8*>(out + static_cast<size_t>(row) * Kpad + kc * 8) = pack8(v);
  }
}

// Thread layout shared by the NHWC stem kernels below: a block of ST_THREADS threads is (lanes x cv) with cv = C / 8 channel vectors, so
// every thread keeps ONE channel vector for its whole life (per-channel coefficients and partial sums live in registers) and a warp
// touches runs of consecutive 16-byte vectors.  384 = lcm-friendly for every stem width (cv = 8, 12, 24, 48, 96 / 16, 32, 64, 128).
// All index arithmetic is 32-bit unsigned (pixel / slot counts are < 2^31; byte offsets are widened at the last multiply): the
// previous versions spent their time in 64-bit divisions and per-element scalar loads of the coefficients, not on HBM.
constexpr int ST_THREADS = 384;

__device__ __forceinline__ void load8f(const float* __restrict__ p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// NHWC bf16 input [B,H,W,C] (C % 8 == 0); optional per-channel affine + ReLU (= BatchNorm + ReLU of the producer layer);
// out [B*Ho*Wo, k*k*C], K order (kh, kw, c).  Zero padding applies to the *post-ReLU* activation (conv pads its input).
// A "slot" is one (output pixel, tap) pair = C contiguous output elements at out + slot * C; two slots per thread and iteration.
template <int KSZ>
__global__ void __launch_bounds__(ST_THREADS) im2col_nhwc_kernel(const bf16* __restrict__ in, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, bf16* __restrict__ out, int H, int W,
                                                                 int C, int stride, int Ho, int Wo, unsigned nslots) {
  constexpr unsigned TAPS = KSZ * KSZ;
  constexpr int PAD = KSZ == 3 ? 1 : 0;
  const unsigned cv = C >> 3;
  const unsigned lanes = ST_THREADS / cv, cvi = threadIdx.x % cv, lane = threadIdx.x / cv;
  if (lane >= lanes) return;
  const bool affine = scale != nullptr;
  float sc[8], sh[8];
  if (affine) { load8f(scale + cvi * 8, sc); load8f(shift + cvi * 8, sh); }
  const unsigned step = gridDim.x * lanes;
  for (unsigned s0 = blockIdx.x * lanes + lane; s0 < nslots; s0 += 2 * step) {
    bf16x8 raw[2];
    bool inside[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned slot = s0 + u * step;
      inside[u] = false;
      raw[u] = make_uint4(0, 0, 0, 0);
      if (slot < nslots) {
        const unsigned tap = slot % TAPS, row = slot / TAPS;
        const unsigned xo = row % Wo, t = row / Wo, yo = t % Ho, b = t / Ho;
        const int yi = static_cast<int>(yo) * stride - PAD + static_cast<int>(tap / KSZ);
        const int xi = static_cast<int>(xo) * stride - PAD + static_cast<int>(tap % KSZ);
        if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
          inside[u] = true;
          raw[u] = *reinterpret_cast<const bf16x8*>(in + (static_cast<size_t>(b * H + yi) * W + xi) * C + cvi * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned slot = s0 + u * step;
      if (slot < nslots) {
        if (affine && inside[u]) {
          float v[8];
          unpack8(raw[u], v);
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = fmaxf(fmaf(v[t], sc[t], sh[t]), 0.f);
          raw[u] = pack8(v);
        }
        *reinterpret_cast<bf16x8*>(out + static_cast<size_t>(slot) * C + cvi * 8) = raw[u];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- BatchNorm statistics
// y bf16 [M, C]: acc[0][c] += sum, acc[1][c] += sum of squares (acc must be zeroed by the caller).  Same (lanes x channel-vector)
// thread layout as the kernels above; four independent 16-byte row loads in flight per thread.
__global__ void __launch_bounds__(ST_THREADS) bn_stats_kernel(const bf16* __restrict__ y, float* __restrict__ acc, unsigned M, int C) {
  const unsigned cv = C >> 3;
  const unsigned lanes = ST_THREADS / cv, cvi = threadIdx.x % cv, lane = threadIdx.x / cv;
  __shared__ float red[2][ST_THREADS * 8];   // [sum | sumsq][lane * C + channel]; lanes * C <= 3072
  if (lane < lanes) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned step = gridDim.x * lanes;
    for (unsigned r0 = blockIdx.x * lanes + lane; r0 < M; r0 += 4 * step) {
      bf16x8 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned r = r0 + u * step;
        raw[u] = r < M ? *reinterpret_cast<const bf16x8*>(y + static_cast<size_t>(r) * C + cvi * 8) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int t = 0; t < 8; ++t) { s[t] += f[t]; q[t] = fmaf(f[t], f[t], q[t]); }
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) { red[0][lane * C + cvi * 8 + t] = s[t]; red[1][lane * C + cvi * 8 + t] = q[t]; }
  }
  __syncthreads();
  // one atomic per channel per block
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const int which = c / C, ch = c % C;
    float a = 0.f;
    for (unsigned l = 0; l < lanes; ++l) a += red[which][l * C + ch];
    atomicAdd(acc + which * C + ch, a);
  }
}

// finalize: batch mean / biased var -> (scale, shift, mean, rstd); running stats (momentum 0.1, unbiased var) updated in place
__global__ void bn_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long M,
                                   int C, float eps, float momentum, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    mean = acc[c] / M;
    var = fmaxf(acc[C + c] / M - mean * mean, 0.f);
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (static_cast<float>(M) / fmaxf(static_cast<float>(M - 1), 1.f));
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (mean_out) { mean_out[c] = mean; rstd_out[c] = rstd; }
}

// ---------------------------------------------------------------------------------------------- BN + ReLU backward (1/2)
// da[pix, c] = col2im(dAcol) for the consumer conv (KSZ, STRIDE; consumer output grid Ho x Wo); n = y*scale + shift;
// dn = (n > 0) ? da : 0; red[0][c] += dn; red[1][c] += dn * xhat  (xhat = (y - mean) * rstd).   Writes dn (bf16).
// Which (kh, kw) taps of which consumer outputs saw input pixel (yi, xi):  yo * STRIDE - 1 + kh == yi.
//   STRIDE 2: yi even -> kh = 1, yo = yi / 2;  yi odd -> kh = 0, yo = (yi + 1) / 2 and kh = 2, yo = (yi - 1) / 2   (same in x): <= 4 taps
//   STRIDE 1: kh = 0, 1, 2 -> yo = yi + 1 - kh: <= 9 taps.      KSZ 1: the pixel itself.
// Every 16-byte piece of dAcol is read exactly once over the grid; all tap loads of a pixel are issued before the first is consumed.
template <int KSZ, int STRIDE>
__global__ void __launch_bounds__(ST_THREADS, (KSZ == 3 && STRIDE == 1) ? 1 : 2) bn_relu_bwd_gather_kernel(const bf16* __restrict__ dAcol, const bf16* __restrict__ y,
                                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                        bf16* __restrict__ dn, float* __restrict__ red, int H, int W, int C,
                                                                        int Ho, int Wo, unsigned npix) {
  constexpr int NT = KSZ == 1 ? 1 : (STRIDE == 2 ? 4 : 9);
  const unsigned cv = C >> 3;
  const unsigned lanes = ST_THREADS / cv, cvi = threadIdx.x % cv, lane = threadIdx.x / cv;
  const size_t ldr = static_cast<size_t>(KSZ * KSZ) * C;        // dAcol row length
  float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    float sc[8], sh[8], mu[8];
    load8f(scale + cvi * 8, sc); load8f(shift + cvi * 8, sh); load8f(mean + cvi * 8, mu);
    const unsigned step = gridDim.x * lanes;
    for (unsigned pix = blockIdx.x * lanes + lane; pix < npix; pix += step) {
      const unsigned xi = pix % W, t = pix / W, yi = t % H, b = t / H;
      bf16x8 tv[NT];
      const bf16x8 yraw = *reinterpret_cast<const bf16x8*>(y + static_cast<size_t>(pix) * C + cvi * 8);
      if constexpr (KSZ == 1) {
        tv[0] = *reinterpret_cast<const bf16x8*>(dAcol + static_cast<size_t>(pix) * C + cvi * 8);
      } else if constexpr (STRIDE == 2) {
        const int yodd = yi & 1, xodd = xi & 1;
        const int kh[2] = {yodd ? 0 : 1, 2}, kw[2] = {xodd ? 0 : 1, 2};
        const int yo[2] = {static_cast<int>(yi + 1 - kh[0]) >> 1, (static_cast<int>(yi) - 1) >> 1};
        const int xo[2] = {static_cast<int>(xi + 1 - kw[0]) >> 1, (static_cast<int>(xi) - 1) >> 1};
        const bool yok[2] = {yo[0] < Ho, yodd != 0}, xok[2] = {xo[0] < Wo, xodd != 0};
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tv[a * 2 + c] = make_uint4(0, 0, 0, 0);
            if (yok[a] && xok[c])
              tv[a * 2 + c] = *reinterpret_cast<const bf16x8*>(dAcol + (static_cast<size_t>(b * Ho + yo[a]) * Wo + xo[c]) * ldr +
                                                               (kh[a] * 3 + kw[c]) * C + cvi * 8);
          }
      } else {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int yo = static_cast<int>(yi) + 1 - kh, xo = static_cast<int>(xi) + 1 - kw;
            tv[kh * 3 + kw] = make_uint4(0, 0, 0, 0);
            if (yo >= 0 && yo < Ho && xo >= 0 && xo < Wo)
              tv[kh * 3 + kw] = *reinterpret_cast<const bf16x8*>(dAcol + (static_cast<size_t>(b * Ho + yo) * Wo + xo) * ldr + (kh * 3 + kw) * C +
                                                                 cvi * 8);
          }
      }
      float da[8] = {0, 0, 0, 0, 0, 0, 0, 0}, yv[8];
#pragma unroll
      for (int tp = 0; tp < NT; ++tp) {
        float f[8];
        unpack8(tv[tp], f);
#pragma unroll
        for (int t8 = 0; t8 < 8; ++t8) da[t8] += f[t8];
      }
      unpack8(yraw, yv);
#pragma unroll
      for (int t8 = 0; t8 < 8; ++t8) {
        const float g = fmaf(yv[t8], sc[t8], sh[t8]) > 0.f ? da[t8] : 0.f;
        da[t8] = g;
        s0[t8] += g;
        s1[t8] = fmaf(g, yv[t8] - mu[t8], s1[t8]);
      }
      *reinterpret_cast<bf16x8*>(dn + static_cast<size_t>(pix) * C + cvi * 8) = pack8(da);
    }
  }
  __shared__ float sred[2][ST_THREADS * 8];                    // [sum | sum * (y - mean)][lane * C + channel]; lanes * C <= 3072
  if (lane < lanes) {
#pragma unroll
    for (int t8 = 0; t8 < 8; ++t8) { sred[0][lane * C + cvi * 8 + t8] = s0[t8]; sred[1][lane * C + cvi * 8 + t8] = s1[t8]; }
  }
  __syncthreads();
  // one atomic per channel and block; the second sum becomes sum(dn * xhat) by the channel's rstd
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const int which = c / C, ch = c % C;
    float a = 0.f;
    for (unsigned l = 0; l < lanes; ++l) a += sred[which][l * C + ch];
    if (which) a *= rstd[ch];
    atomicAdd(red + which * C + ch, a);
  }
}

// BN backward (2/2): dy = gamma*rstd * (dn - mean(dn) - xhat * mean(dn*xhat));  dgamma += sum dn*xhat;  dbeta += sum dn.
// Per thread (one channel vector): dy = a * dn + k1 * (y - mean) + k0 with a = gamma*rstd, k1 = -a*rstd*mean(dn*xhat), k0 = -a*mean(dn).
// EVAL (running statistics are constants, nn.BatchNorm2d in eval()): dy = a * dn.
template <bool EVAL>
__global__ void __launch_bounds__(ST_THREADS) bn_bwd_apply_kernel(const bf16* __restrict__ dn, const bf16* __restrict__ y,
                                                                  const float* __restrict__ red, const float* __restrict__ gamma,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  bf16* __restrict__ dy, float* __restrict__ dgamma,
                                                                  float* __restrict__ dbeta, unsigned M, int C) {
  const unsigned cv = C >> 3;
  const unsigned lanes = ST_THREADS / cv, cvi = threadIdx.x % cv, lane = threadIdx.x / cv;
  if (lane < lanes) {
    float a[8], k1[8], k0[8], mu[8];
    {
      float g[8], rs[8], r0[8], r1[8];
      load8f(gamma + cvi * 8, g); load8f(rstd + cvi * 8, rs);
      if (!EVAL) { load8f(red + cvi * 8, r0); load8f(red + C + cvi * 8, r1); load8f(mean + cvi * 8, mu); }
      const float invM = 1.0f / static_cast<float>(M);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        a[t] = g[t] * rs[t];
        if (!EVAL) { k1[t] = -a[t] * rs[t] * (r1[t] * invM); k0[t] = -a[t] * (r0[t] * invM); }
      }
    }
    const unsigned step = gridDim.x * lanes;
    for (unsigned r0 = blockIdx.x * lanes + lane; r0 < M; r0 += 2 * step) {
      bf16x8 graw[2], yraw[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const unsigned r = r0 + u * step;
        if (r < M) {
          graw[u] = *reinterpret_cast<const bf16x8*>(dn + static_cast<size_t>(r) * C + cvi * 8);
          if (!EVAL) yraw[u] = *reinterpret_cast<const bf16x8*>(y + static_cast<size_t>(r) * C + cvi * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const unsigned r = r0 + u * step;
        if (r < M) {
          float g[8], yv[8];
          unpack8(graw[u], g);
          if (!EVAL) {
            unpack8(yraw[u], yv);
#pragma unroll
            for (int t = 0; t < 8; ++t) g[t] = fmaf(a[t], g[t], fmaf(k1[t], yv[t] - mu[t], k0[t]));
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) g[t] *= a[t];
          }
          *reinterpret_cast<bf16x8*>(dy + static_cast<size_t>(r) * C + cvi * 8) = pack8(g);
        }
      }
    }
  }
  if (blockIdx.x == 0 && dgamma) {
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) { dgamma[ch] += red[C + ch]; dbeta[ch] += red[ch]; }
  }
}

// ---------------------------------------------------------------------------------------------- conv weight layouts
// w fp32 [Cout, Cin, k, k] (reference layout) -> bf16 [Cout, Kpad] with K order (kh, kw, c), zero padded
__global__ void conv_weight_pack_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int ksz, int Kpad) {
  const int K = Cin * ksz * ksz;
  const long long total = static_cast<long long>(Cout) * Kpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Kpad), co = static_cast<int>(i / Kpad);
    float v = 0.f;
    if (k < K) {
      const int c = k % Cin, tap = k / Cin;
      v = w[(static_cast<long long>(co) * Cin + c) * ksz * ksz + tap];
    }
    out[i] = __float2bfloat16(v);
  }
}
// dw_packed fp32 [Cout, Kpad] (K order (kh,kw,c)) -> grad fp32 [Cout, Cin, k, k] +=
__global__ void conv_weight_unpack_grad_kernel(const float* __restrict__ dwp, float* __restrict__ grad, int Cout, int Cin, int ksz,
                                               int Kpad) {
  const int K = Cin * ksz * ksz;
  const long long total = static_cast<long long>(Cout) * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i % K), co = static_cast<int>(i / K);
    const int tap = r % (ksz * ksz), c = r / (ksz * ksz);
    grad[i] += dwp[static_cast<long long>(co) * Kpad + tap * Cin + c];
  }
}
// generic fp32 [R, C] -> bf16 [R, Cpad] (zero padded) and its gradient counterpart
__global__ void cast_pad_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long R, int C, int Cpad) {
  const long long total = R * Cpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    dst[i] = __float2bfloat16(c < C ? src[(i / Cpad) * C + c] : 0.f);
  }
}
__global__ void unpad_add_kernel(const float* __restrict__ src, float* __restrict__ dst, long long R, int C, int Cpad) {
  const long long total = R * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] += src[(i / C) * Cpad + (i % C)];
}

}  // namespace

extern "C" int prismer_patchify(const float* x, void* out, int B, int Cin, int R, int p, int Kpad, cudaStream_t stream) {
  // nn.Conv2d(kernel = stride = p, no padding) ignores the R % p trailing pixels (ViT-L/14 at 480 px: 34 x 34 patches of 476 pixels)
  if (R < p || Kpad % 8 || Kpad < Cin * p * p) return PRISMER_ERR_SHAPE;
  const int g = R / p;
  patchify_kernel<<<blocks_for(static_cast<long long>(B) * g * g * (Kpad / 8), 256), 256, 0, stream>>>(
      x, reinterpret_cast<bf16*>(out), B, Cin, R, p, g, Cin * p * p, Kpad);
  return LAUNCH_CHECK();
}

extern "C" int prismer_resample_bilinear(const float* x, void* out, int B, int C, int Hi, int Wi, int Ho, int Wo,
                                         cudaStream_t stream) {
  if (Wi % 4 || (reinterpret_cast<uintptr_t>(x) & 15)) return PRISMER_ERR_ALIGN;
  const size_t smem = sizeof(float) * 2 * RS_CH * Wi;
  static size_t configured = 0;
  if (smem > 48 * 1024 && configured < smem) {
    if (cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return PRISMER_ERR_CUDA;
    configured = smem;
  }
  resample_kernel<<<B * Ho, 256, smem, stream>>>(x, reinterpret_cast<bf16*>(out), B, C, Hi, Wi, Ho, Wo);
  return LAUNCH_CHECK();
}

extern "C" int prismer_im2col_first(const void* in, int in_is_bf16, long long sb, long long sc, long long sy, long long sx,
                                    void* out, int B, int Cin, int H, int W, int ksz, int stride, int Ho, int Wo, int Kpad,
                                    cudaStream_t stream) {
  if (Kpad % 8 || Kpad > 64 || Kpad < ksz * ksz * Cin || (ksz != 1 && ksz != 3)) return PRISMER_ERR_SHAPE;
  const long long total = static_cast<long long>(B) * Ho * Wo * (Kpad / 8);
  if (total <= 0 || total >= (1ll << 31)) return PRISMER_ERR_SHAPE;
  im2col_first_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(in, in_is_bf16, sb, sc, sy, sx, reinterpret_cast<bf16*>(out), B, Cin, H, W,
                                                                 ksz, stride, Ho, Wo, Kpad);
  return LAUNCH_CHECK();
}

extern "C" int prismer_im2col_nhwc(const void* in, const float* scale, const float* shift, void* out, int B, int H, int W, int C,
                                   int ksz, int stride, int Ho, int Wo, cudaStream_t stream) {
  if (C % 8 || C / 8 > ST_THREADS || (ksz != 1 && ksz != 3) || ((scale == nullptr) != (shift == nullptr))) return PRISMER_ERR_SHAPE;
  const long long nslots = static_cast<long long>(B) * Ho * Wo * ksz * ksz;
  if (nslots <= 0 || nslots >= (1ll << 31) || static_cast<long long>(B) * H * W >= (1ll << 31)) return PRISMER_ERR_SHAPE;
  const int grid = stem_grid(nslots, C, 2);
  if (ksz == 3)
    im2col_nhwc_kernel<3><<<grid, ST_THREADS, 0, stream>>>(reinterpret_cast<const bf16*>(in), scale, shift, reinterpret_cast<bf16*>(out), H, W,
                                                         C, stride, Ho, Wo, static_cast<unsigned>(nslots));
  else
    im2col_nhwc_kernel<1><<<grid, ST_THREADS, 0, stream>>>(reinterpret_cast<const bf16*>(in), scale, shift, reinterpret_cast<bf16*>(out), H, W,
                                                         C, stride, Ho, Wo, static_cast<unsigned>(nslots));
  return LAUNCH_CHECK();
}

extern "C" int prismer_bn_stats(const void* y, float* acc, long long M, int C, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* rstd,
                                float eps, float momentum, int training, cudaStream_t stream) {
  if (C % 8 || C / 8 > ST_THREADS || M <= 0 || M >= (1ll << 31)) return PRISMER_ERR_SHAPE;
  if (training) {
    if (cudaMemsetAsync(acc, 0, sizeof(float) * 2 * C, stream) != cudaSuccess) return PRISMER_ERR_CUDA;
    bn_stats_kernel<<<stem_grid(M, C, 4), ST_THREADS, 0, stream>>>(reinterpret_cast<const bf16*>(y), acc, static_cast<unsigned>(M), C);
  }
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(acc, gamma, beta, running_mean, running_var, scale, shift, mean, rstd, M, C,
                                                         eps, momentum, training);
  return LAUNCH_CHECK();
}

namespace {
template <bool EVAL>
int bn_relu_bwd_launch(const void* dAcol, const void* y, const float* scale, const float* shift, const float* mean, const float* rstd,
                       const float* gamma, void* dn_scratch, void* dy, float* red, float* dgamma, float* dbeta, int B, int H, int W, int C,
                       int ksz, int stride, int Ho, int Wo, cudaStream_t stream) {
  if (C % 8 || C / 8 > ST_THREADS || (ksz != 1 && ksz != 3) || (ksz == 3 && stride != 1 && stride != 2) || (ksz == 1 && stride != 1))
    return PRISMER_ERR_SHAPE;
  const long long M = static_cast<long long>(B) * H * W;
  if (M <= 0 || M >= (1ll << 31) || static_cast<long long>(B) * Ho * Wo >= (1ll << 31)) return PRISMER_ERR_SHAPE;
  if (cudaMemsetAsync(red, 0, sizeof(float) * 2 * C, stream) != cudaSuccess) return PRISMER_ERR_CUDA;
  const bf16* dA = reinterpret_cast<const bf16*>(dAcol);
  const bf16* yy = reinterpret_cast<const bf16*>(y);
  bf16* dn = reinterpret_cast<bf16*>(dn_scratch);
  const unsigned npix = static_cast<unsigned>(M);
  const int grid = stem_grid(M, C, 1);
  if (ksz == 1)
    bn_relu_bwd_gather_kernel<1, 1><<<grid, ST_THREADS, 0, stream>>>(dA, yy, scale, shift, mean, rstd, dn, red, H, W, C, Ho, Wo, npix);
  else if (stride == 2)
    bn_relu_bwd_gather_kernel<3, 2><<<grid, ST_THREADS, 0, stream>>>(dA, yy, scale, shift, mean, rstd, dn, red, H, W, C, Ho, Wo, npix);
  else
    bn_relu_bwd_gather_kernel<3, 1><<<grid, ST_THREADS, 0, stream>>>(dA, yy, scale, shift, mean, rstd, dn, red, H, W, C, Ho, Wo, npix);
  bn_bwd_apply_kernel<EVAL><<<stem_grid(M, C, 2), ST_THREADS, 0, stream>>>(dn, yy, red, gamma, mean, rstd, reinterpret_cast<bf16*>(dy), dgamma,
                                                                         dbeta, npix, C);
  return LAUNCH_CHECK();
}
}  // namespace

extern "C" int prismer_bn_relu_bwd(const void* dAcol, const void* y, const float* scale, const float* shift, const float* mean,
                                   const float* rstd, const float* gamma, void* dn_scratch, void* dy, float* red, float* dgamma,
                                   float* dbeta, int B, int H, int W, int C, int ksz, int stride, int Ho, int Wo,
                                   cudaStream_t stream) {
  return bn_relu_bwd_launch<false>(dAcol, y, scale, shift, mean, rstd, gamma, dn_scratch, dy, red, dgamma, dbeta, B, H, W, C, ksz, stride, Ho,
                                   Wo, stream);
}

// Same as prismer_bn_relu_bwd for a BatchNorm that ran on its running statistics (module in eval()): nn.BatchNorm2d's eval-mode
// gradient has no batch-mean terms.
extern "C" int prismer_bn_relu_bwd_eval(const void* dAcol, const void* y, const float* scale, const float* shift, const float* mean,
                                        const float* rstd, const float* gamma, void* dn_scratch, void* dy, float* red, float* dgamma,
                                        float* dbeta, int B, int H, int W, int C, int ksz, int stride, int Ho, int Wo,
                                        cudaStream_t stream) {
  return bn_relu_bwd_launch<true>(dAcol, y, scale, shift, mean, rstd, gamma, dn_scratch, dy, red, dgamma, dbeta, B, H, W, C, ksz, stride, Ho,
                                  Wo, stream);
}

extern "C" int prismer_conv_weight_pack(const float* w, void* out, int Cout, int Cin, int ksz, int Kpad, cudaStream_t stream) {
  if (Kpad % 8 || Kpad < Cin * ksz * ksz) return PRISMER_ERR_SHAPE;
  conv_weight_pack_kernel<<<blocks_for(static_cast<long long>(Cout) * Kpad, 256), 256, 0, stream>>>(w, reinterpret_cast<bf16*>(out),
                                                                                                  Cout, Cin, ksz, Kpad);
  return LAUNCH_CHECK();
}

extern "C" int prismer_conv_weight_unpack_grad(const float* dwp, float* grad, int Cout, int Cin, int ksz, int Kpad,
                                               cudaStream_t stream) {
  conv_weight_unpack_grad_kernel<<<blocks_for(static_cast<long long>(Cout) * Cin * ksz * ksz, 256), 256, 0, stream>>>(dwp, grad, Cout,
                                                                                                                    Cin, ksz, Kpad);
  return LAUNCH_CHECK();
}

extern "C" int prismer_cast_pad(const float* src, void* dst, long long R, int C, int Cpad, cudaStream_t stream) {
  if (Cpad < C) return PRISMER_ERR_SHAPE;
  cast_pad_kernel<<<blocks_for(R * Cpad, 256), 256, 0, stream>>>(src, reinterpret_cast<bf16*>(dst), R, C, Cpad);
  return LAUNCH_CHECK();
}

extern "C" int prismer_unpad_add(const float* src, float* dst, long long R, int C, int Cpad, cudaStream_t stream) {
  if (Cpad < C) return PRISMER_ERR_SHAPE;
  unpad_add_kernel<<<blocks_for(R * C, 256), 256, 0, stream>>>(src, dst, R, C, Cpad);
  return LAUNCH_CHECK();
}
