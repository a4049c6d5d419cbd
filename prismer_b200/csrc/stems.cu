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
__global__ void im2col_first_kernel(const void* __restrict__ in, int in_is_bf16, long long sb, long long sc, long long sy,
                                    long long sx, bf16* __restrict__ out, int B, int Cin, int H, int W, int ksz, int stride,
                                    int Ho, int Wo, int Kpad) {
  const int pad = ksz == 3 ? 1 : 0;
  const int K = ksz * ksz * Cin, kv = Kpad >> 3;
  const long long total = static_cast<long long>(B) * Ho * Wo * kv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kc = static_cast<int>(i % kv);
    const long long row = i / kv;
    const int xo = static_cast<int>(row % Wo), yo = static_cast<int>((row / Wo) % Ho), b = static_cast<int>(row / (static_cast<long long>(Wo) * Ho));
    float v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int k = kc * 8 + t;
      float val = 0.f;
      if (k < K) {
        const int c = k % Cin, kw = (k / Cin) % ksz, kh = k / (Cin * ksz);
        const int yi = yo * stride - pad + kh, xi = xo * stride - pad + kw;
        if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
          const long long off = b * sb + c * sc + yi * sy + xi * sx;
          val = in_is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(in)[off]) : reinterpret_cast<const float*>(in)[off];
        }
      }
      v[t] = val;
    }
    *reinterpret_cast<bf16x8*>(out + row * Kpad + kc * 8) = pack8(v);
  }
}

// NHWC bf16 input [B,H,W,C] (C % 8 == 0); optional per-channel affine + ReLU (= BatchNorm + ReLU of the producer layer);
// out [B*Ho*Wo, k*k*C], K order (kh, kw, c).  Zero padding applies to the *post-ReLU* activation (conv pads its input).
__global__ void im2col_nhwc_kernel(const bf16* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ shift,
                                   bf16* __restrict__ out, int B, int H, int W, int C, int ksz, int stride, int Ho, int Wo) {
  const int pad = ksz == 3 ? 1 : 0;
  const int cv = C >> 3, taps = ksz * ksz;
  const long long total = static_cast<long long>(B) * Ho * Wo * taps * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    const int tap = static_cast<int>((i / cv) % taps);
    const long long row = i / (static_cast<long long>(cv) * taps);
    const int xo = static_cast<int>(row % Wo), yo = static_cast<int>((row / Wo) % Ho), b = static_cast<int>(row / (static_cast<long long>(Wo) * Ho));
    const int kh = tap / ksz, kw = tap % ksz;
    const int yi = yo * stride - pad + kh, xi = xo * stride - pad + kw;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
      unpack8(*reinterpret_cast<const bf16x8*>(in + ((static_cast<long long>(b) * H + yi) * W + xi) * C + c * 8), v);
      if (scale) {
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = fmaxf(v[t] * __ldg(scale + c * 8 + t) + __ldg(shift + c * 8 + t), 0.f);
      }
    }
    *reinterpret_cast<bf16x8*>(out + row * (static_cast<long long>(taps) * C) + tap * C + c * 8) = pack8(v);
  }
}

// ---------------------------------------------------------------------------------------------- BatchNorm statistics
// y bf16 [M, C]: acc[0][c] += sum, acc[1][c] += sum of squares (acc must be zeroed by the caller)
__global__ void __launch_bounds__(256) bn_stats_kernel(const bf16* __restrict__ y, float* __restrict__ acc, long long M, int C,
                                                       int rows_per_block) {
  // thread -> (channel vector cvi, row lane); blockDim.x = 256
  const int cv = C >> 3;
  const int lanes = 256 / cv > 0 ? 256 / cv : 1;       // row lanes per block when cv <= 256
  const int cvi = threadIdx.x % cv, rl = threadIdx.x / cv;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  __shared__ float red[2][2048];   // [sum | sumsq][lane * C + channel]; lanes * C <= 2048
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes) {
    for (long long r = r0 + rl; r < r1; r += lanes) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(y + r * C + cvi * 8), f);
#pragma unroll
      for (int t = 0; t < 8; ++t) { s[t] += f[t]; q[t] += f[t] * f[t]; }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) { red[0][rl * C + cvi * 8 + t] = s[t]; red[1][rl * C + cvi * 8 + t] = q[t]; }
  }
  __syncthreads();
  // one atomic per channel per block (instead of one per thread): 20x fewer same-address atomics
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const int which = c / C, ch = c % C;
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += red[which][l * C + ch];
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
// da[pix, c] = col2im(dAcol) for the consumer conv (ksz, stride; consumer output grid Ho x Wo); n = y*scale + shift;
// dn = (n > 0) ? da : 0; red[0][c] += dn; red[1][c] += dn * xhat  (xhat = (y - mean) * rstd).   Writes dn (bf16).
__global__ void __launch_bounds__(256) bn_relu_bwd_gather_kernel(const bf16* __restrict__ dAcol, const bf16* __restrict__ y,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 bf16* __restrict__ dn, float* __restrict__ red, int B, int H, int W,
                                                                 int C, int ksz, int stride, int Ho, int Wo) {
  const int pad = ksz == 3 ? 1 : 0;
  const int cv = C >> 3, taps = ksz * ksz;
  const long long npix = static_cast<long long>(B) * H * W;
  // each thread owns one channel vector and strides over pixels, so the per-channel partial sums stay in registers
  const int cvi = threadIdx.x % cv;
  const int lanes = blockDim.x / cv;
  const int rl = threadIdx.x / cv;
  float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes) {
    float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { sc[t] = scale[cvi * 8 + t]; sh[t] = shift[cvi * 8 + t]; mu[t] = mean[cvi * 8 + t]; rs[t] = rstd[cvi * 8 + t]; }
    for (long long pix = static_cast<long long>(blockIdx.x) * lanes + rl; pix < npix; pix += static_cast<long long>(gridDim.x) * lanes) {
      const int xi = static_cast<int>(pix % W), yi = static_cast<int>((pix / W) % H), b = static_cast<int>(pix / (static_cast<long long>(W) * H));
      float da[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      // all (<= 9) tap loads are issued before any is consumed: 9 independent 128-bit loads in flight per thread
      bf16x8 taps_v[9];
      bool taps_ok[9];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const int kh = tp / 3, kw = tp % 3;
        bool ok = tp < taps && kh < ksz && kw < ksz;
        const int ty = yi + pad - (ksz == 3 ? kh : 0), tx = xi + pad - (ksz == 3 ? kw : 0);
        ok = ok && (ksz == 3 || tp == 0) && ty >= 0 && tx >= 0 && (ty % stride) == 0 && (tx % stride) == 0;
        const int yo = ty / stride, xo = tx / stride;
        ok = ok && yo < Ho && xo < Wo;
        taps_ok[tp] = ok;
        taps_v[tp] = make_uint4(0, 0, 0, 0);
        if (ok) {
          const int tapi = ksz == 3 ? tp : 0;
          taps_v[tp] = *reinterpret_cast<const bf16x8*>(dAcol + ((static_cast<long long>(b) * Ho + yo) * Wo + xo) * (static_cast<long long>(taps) * C) +
                                                        tapi * C + cvi * 8);
        }
      }
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        if (taps_ok[tp]) {
          float f[8];
          unpack8(taps_v[tp], f);
#pragma unroll
          for (int t = 0; t < 8; ++t) da[t] += f[t];
        }
      }
      float yv[8];
      unpack8(*reinterpret_cast<const bf16x8*>(y + pix * C + cvi * 8), yv);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float n = yv[t] * sc[t] + sh[t];
        const float g = n > 0.f ? da[t] : 0.f;
        da[t] = g;
        s0[t] += g;
        s1[t] += g * (yv[t] - mu[t]) * rs[t];
      }
      *reinterpret_cast<bf16x8*>(dn + pix * C + cvi * 8) = pack8(da);
    }
  }
  __shared__ float sred[2][2048];
  if (rl < lanes) {
#pragma unroll
    for (int t = 0; t < 8; ++t) { sred[0][rl * C + cvi * 8 + t] = s0[t]; sred[1][rl * C + cvi * 8 + t] = s1[t]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const int which = c / C, ch = c % C;
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += sred[which][l * C + ch];
    atomicAdd(red + which * C + ch, a);
  }
}

// BN backward (2/2): dy = gamma*rstd * (dn - mean(dn) - xhat * mean(dn*xhat));  dgamma += sum dn*xhat;  dbeta += sum dn
__global__ void bn_bwd_apply_kernel(const bf16* __restrict__ dn, const bf16* __restrict__ y, const float* __restrict__ red,
                                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                                    bf16* __restrict__ dy, float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C) {
  const int cv = C >> 3;
  const long long total = M * cv;
  const float invM = 1.0f / static_cast<float>(M);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    float g[8], yv[8];
    unpack8(*reinterpret_cast<const bf16x8*>(dn + i * 8), g);
    unpack8(*reinterpret_cast<const bf16x8*>(y + i * 8), yv);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int ch = c * 8 + t;
      const float xh = (yv[t] - mean[ch]) * rstd[ch];
      g[t] = gamma[ch] * rstd[ch] * (g[t] - red[ch] * invM - xh * red[C + ch] * invM);
    }
    *reinterpret_cast<bf16x8*>(dy + i * 8) = pack8(g);
  }
  if (blockIdx.x == 0 && dgamma) {
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) { dgamma[ch] += red[C + ch]; dbeta[ch] += red[ch]; }
  }
}

// eval-mode BatchNorm (running statistics are constants): dy = gamma*rstd * dn;  dgamma += sum dn*xhat;  dbeta += sum dn
__global__ void bn_bwd_apply_eval_kernel(const bf16* __restrict__ dn, const float* __restrict__ red, const float* __restrict__ gamma,
                                         const float* __restrict__ rstd, bf16* __restrict__ dy, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta, long long M, int C) {
  const int cv = C >> 3;
  const long long total = M * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    float g[8];
    unpack8(*reinterpret_cast<const bf16x8*>(dn + i * 8), g);
#pragma unroll
    for (int t = 0; t < 8; ++t) g[t] *= gamma[c * 8 + t] * rstd[c * 8 + t];
    *reinterpret_cast<bf16x8*>(dy + i * 8) = pack8(g);
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
  if (Kpad % 8 || Kpad < ksz * ksz * Cin || (ksz != 1 && ksz != 3)) return PRISMER_ERR_SHAPE;
  im2col_first_kernel<<<blocks_for(static_cast<long long>(B) * Ho * Wo * (Kpad / 8), 256), 256, 0, stream>>>(
      in, in_is_bf16, sb, sc, sy, sx, reinterpret_cast<bf16*>(out), B, Cin, H, W, ksz, stride, Ho, Wo, Kpad);
  return LAUNCH_CHECK();
}

extern "C" int prismer_im2col_nhwc(const void* in, const float* scale, const float* shift, void* out, int B, int H, int W, int C,
                                   int ksz, int stride, int Ho, int Wo, cudaStream_t stream) {
  if (C % 8 || (ksz != 1 && ksz != 3) || ((scale == nullptr) != (shift == nullptr))) return PRISMER_ERR_SHAPE;
  im2col_nhwc_kernel<<<blocks_for(static_cast<long long>(B) * Ho * Wo * ksz * ksz * (C / 8), 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(in), scale, shift, reinterpret_cast<bf16*>(out), B, H, W, C, ksz, stride, Ho, Wo);
  return LAUNCH_CHECK();
}

extern "C" int prismer_bn_stats(const void* y, float* acc, long long M, int C, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* rstd,
                                float eps, float momentum, int training, cudaStream_t stream) {
  if (C % 8 || C / 8 > 256) return PRISMER_ERR_SHAPE;
  if (training) {
    if (cudaMemsetAsync(acc, 0, sizeof(float) * 2 * C, stream) != cudaSuccess) return PRISMER_ERR_CUDA;
    const int lanes = 256 / (C / 8);
    long long rpb = (M + 148 * 4 - 1) / (148 * 4);
    rpb = ((rpb + lanes - 1) / lanes) * lanes;
    if (rpb < lanes) rpb = lanes;
    const int grid = static_cast<int>((M + rpb - 1) / rpb);
    bn_stats_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(y), acc, M, C, static_cast<int>(rpb));
  }
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(acc, gamma, beta, running_mean, running_var, scale, shift, mean, rstd, M, C,
                                                         eps, momentum, training);
  return LAUNCH_CHECK();
}

extern "C" int prismer_bn_relu_bwd(const void* dAcol, const void* y, const float* scale, const float* shift, const float* mean,
                                   const float* rstd, const float* gamma, void* dn_scratch, void* dy, float* red, float* dgamma,
                                   float* dbeta, int B, int H, int W, int C, int ksz, int stride, int Ho, int Wo,
                                   cudaStream_t stream) {
  if (C % 8 || C / 8 > 256 || (ksz != 1 && ksz != 3)) return PRISMER_ERR_SHAPE;
  if (cudaMemsetAsync(red, 0, sizeof(float) * 2 * C, stream) != cudaSuccess) return PRISMER_ERR_CUDA;
  const long long M = static_cast<long long>(B) * H * W;
  const int lanes = 256 / (C / 8);
  long long blocks = (M + lanes - 1) / lanes;
  if (blocks > 148 * 8) blocks = 148 * 8;
  bn_relu_bwd_gather_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(dAcol), reinterpret_cast<const bf16*>(y), scale, shift, mean, rstd,
      reinterpret_cast<bf16*>(dn_scratch), red, B, H, W, C, ksz, stride, Ho, Wo);
  bn_bwd_apply_kernel<<<blocks_for(M * (C / 8), 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dn_scratch),
                                                                       reinterpret_cast<const bf16*>(y), red, gamma, mean, rstd,
                                                                       reinterpret_cast<bf16*>(dy), dgamma, dbeta, M, C);
  return LAUNCH_CHECK();
}

// Same as prismer_bn_relu_bwd for a BatchNorm that ran on its running statistics (module in eval()): nn.BatchNorm2d's eval-mode
// gradient has no batch-mean terms.
extern "C" int prismer_bn_relu_bwd_eval(const void* dAcol, const void* y, const float* scale, const float* shift, const float* mean,
                                        const float* rstd, const float* gamma, void* dn_scratch, void* dy, float* red, float* dgamma,
                                        float* dbeta, int B, int H, int W, int C, int ksz, int stride, int Ho, int Wo,
                                        cudaStream_t stream) {
  if (C % 8 || C / 8 > 256 || (ksz != 1 && ksz != 3)) return PRISMER_ERR_SHAPE;
  if (cudaMemsetAsync(red, 0, sizeof(float) * 2 * C, stream) != cudaSuccess) return PRISMER_ERR_CUDA;
  const long long M = static_cast<long long>(B) * H * W;
  const int lanes = 256 / (C / 8);
  long long blocks = (M + lanes - 1) / lanes;
  if (blocks > 148 * 8) blocks = 148 * 8;
  bn_relu_bwd_gather_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(dAcol), reinterpret_cast<const bf16*>(y), scale, shift, mean, rstd,
      reinterpret_cast<bf16*>(dn_scratch), red, B, H, W, C, ksz, stride, Ho, Wo);
  bn_bwd_apply_eval_kernel<<<blocks_for(M * (C / 8), 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dn_scratch), red, gamma, rstd,
                                                                            reinterpret_cast<bf16*>(dy), dgamma, dbeta, M, C);
  return LAUNCH_CHECK();
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
