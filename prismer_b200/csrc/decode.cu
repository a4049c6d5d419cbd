// KV-cached single-token decode step (replaces the reference's cache-less re-forward of the whole prefix,
// roberta.py:401-406 driven from prismer_caption.py:45-50): the kernels a decoder layer needs when every sequence contributes ONE new
// token (M = batch [* beams] rows, typically 32).  At that size every product is a weight-streaming, launch-latency-bound "skinny"
// GEMM -- 348 MB of bf16 weights per step for Prismer-BASE against 11 GFLOP -- so these are small mma.sync kernels with no TMA /
// TMEM / mbarrier prologue (a 148-CTA persistent tcgen05 GEMM spends longer setting up than this whole product takes):
//
//   skinny_linear : y[M,N] = act(x[M,K] . W[N,K]^T + bias) (+ residual);  32 rows x 16 columns per CTA, the 4 warps split K, weights
//                   go global -> mma B fragments directly (32-byte sectors fully used), x is staged once in shared memory.
//                   (Two ways of folding the post-LayerNorm into this kernel were built, validated and measured on B200, and dropped:
//                   "the last CTA normalises the completed rows" cost ~20 us per launch in its single-CTA tail; "every consumer CTA
//                   normalises the staged rows on load" cost 19.3 us instead of 12.3 us per launch, 37.4 ms per 19-pass decode against
//                   30.7 ms with a separate 3 us ln_fwd launch on the 32 rows -- so the LayerNorm stays a kernel of its own.)
//   decode_attn   : one query per (batch, head) over a K/V cache (self-attention: the new token's k / v are appended to the cache in the
//                   same kernel) or over the projected visual tokens (cross-attention); one warp per (batch, head).
#include "common.cuh"
#include "prismer_sm100.h"

#include <cstdlib>

namespace {

constexpr int kRows = 32;        // rows per CTA (two m16 tiles)
constexpr int kKc = 1024;        // K chunk staged in shared memory
constexpr int kPad = 8;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool ok) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

struct SkinnyParams {
  const bf16* x; long long ldx;
  const bf16* w; long long ldw;
  const float* bias;
  const bf16* residual; long long ldr;
  void* out; long long ldo; int out_fp32;
  int M, N, K, act;
  int kc_len;                    // K chunk staged in shared memory per round trip (multiple of 16, <= kKc)
};

// Every global byte this CTA needs (its 16 x Kc weight slice and the 32 x Kc activation rows) is requested up front with cp.async --
// ONE memory round trip per K chunk instead of a dependent load per k-step (the first version of this kernel was latency-bound:
// 134 ms per 19-token decode).  The products then run from shared memory with ldmatrix + mma.sync; the 4 warps split the k-steps.
// COLS = 16 (body layers: N <= 3072, one wave of CTAs) or 64 (tied LM head, N = 50265: with 16 columns the 3142 CTAs re-read the 48 KB of
// x from L2 150 MB worth per launch -- more than the 77 MB of weights they stream).  The arithmetic per output element is the same.
template <int COLS>
__global__ void __launch_bounds__(128) skinny_linear_kernel(SkinnyParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * COLS, m0 = blockIdx.y * kRows;
  const int mrows = min(kRows, p.M - m0), ncols = min(COLS, p.N - n0);
  const int kc_max = min(p.K, p.kc_len), ld = kc_max + kPad;
  bf16* sx = reinterpret_cast<bf16*>(smem_raw);
  bf16* sw = sx + kRows * ld;
#ifdef PRISMER_PDL
  // Programmatic dependent launch: the weights do not depend on the previous kernel, so this CTA's 16 x K weight slice is pulled
  // towards L2 while the producer of `x` is still running; everything that reads activations comes after the grid dependency wait.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  {
    const int lines = (p.K * 2 + 127) >> 7;                       // 128-byte lines per weight row
    for (int i = threadIdx.x; i < ncols * lines; i += 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.w + static_cast<long long>(n0 + i / lines) * p.ldw + (i % lines) * 64));
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  float acc[2][COLS / 8][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < COLS / 8; ++b) acc[a][b][0] = acc[a][b][1] = acc[a][b][2] = acc[a][b][3] = 0.f;

  for (int kc = 0; kc < p.K; kc += p.kc_len) {
    const int kn = min(p.kc_len, p.K - kc);                     // multiple of 16 (checked on the host)
    const int vpr = kn >> 3;                               // 16-byte vectors per row
    if (kc > 0) __syncthreads();
    for (int i = threadIdx.x; i < kRows * vpr; i += 128) {
      const int r = i / vpr, c = i % vpr;
      const bool ok = r < mrows;
      cp_async16(sx + r * ld + c * 8, p.x + static_cast<long long>(ok ? m0 + r : m0) * p.ldx + kc + c * 8, ok);
    }
    for (int i = threadIdx.x; i < COLS * vpr; i += 128) {
      const int r = i / vpr, c = i % vpr;
      const bool ok = r < ncols;
      cp_async16(sw + r * ld + c * 8, p.w + static_cast<long long>(ok ? n0 + r : n0) * p.ldw + kc + c * 8, ok);
    }
    cp_async_wait_all();
    __syncthreads();
    const int steps = kn >> 4;
    for (int s = warp; s < steps; s += 4) {
      uint32_t a0[4], a1[4], bw[4];
      ldsm_x4(a0, sx + (lane & 15) * ld + s * 16 + (lane >> 4) * 8);
      ldsm_x4(a1, sx + (16 + (lane & 15)) * ld + s * 16 + (lane >> 4) * 8);
      // B operand ([n][k] tile, k contiguous): bw[0], bw[1] = n8 tile 0; bw[2], bw[3] = n8 tile 1 of every 16-column group
#pragma unroll
      for (int g = 0; g < COLS / 16; ++g) {
        ldsm_x4(bw, sw + (g * 16 + (lane & 7) + (lane >> 4) * 8) * ld + s * 16 + ((lane >> 3) & 1) * 8);
        mma16816(acc[0][2 * g], a0, bw[0], bw[1]);
        mma16816(acc[0][2 * g + 1], a0, bw[2], bw[3]);
        mma16816(acc[1][2 * g], a1, bw[0], bw[1]);
        mma16816(acc[1][2 * g + 1], a1, bw[2], bw[3]);
      }
    }
  }
  // cross-warp reduction of the split-K partials through shared memory: red[warp][row][col]
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < COLS / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = mt * 16 + (lane >> 2) + (e >> 1) * 8, c = nt * 8 + 2 * (lane & 3) + (e & 1);
        red[(warp * kRows + r) * COLS + c] = acc[mt][nt][e];
      }
  __syncthreads();
  for (int o = threadIdx.x; o < kRows * COLS; o += 128) {
    const int r = o / COLS, c = o % COLS;
    const int m = m0 + r, n = n0 + c;
    if (r >= mrows || n >= p.N) continue;
    float v = red[o] + red[kRows * COLS + o] + red[2 * kRows * COLS + o] + red[3 * kRows * COLS + o];
    if (p.bias) v += p.bias[n];
    v = act_fwd(p.act, v);
    if (p.residual) v += __bfloat162float(p.residual[static_cast<long long>(m) * p.ldr + n]);
    if (p.out_fp32) reinterpret_cast<float*>(p.out)[static_cast<long long>(m) * p.ldo + n] = v;
    else reinterpret_cast<bf16*>(p.out)[static_cast<long long>(m) * p.ldo + n] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------ single-query attention
struct DecAttnParams {
  const bf16* q; long long q_bs;                   // [B, H*64] new-token queries
  const bf16* k; const bf16* v; long long kv_bs, kv_rs;   // keys / values: base + b*bs + j*rs + h*64, j < len
  int len;
  const bf16* k_new; const bf16* v_new; long long new_bs;  // self-attention: the new token's k / v (key index `len`), appended to ...
  bf16* k_cache; bf16* v_cache; long long c_bs, c_rs;      // ... the cache at row `len`
  const long long* key_mask; int mask_ld;          // [B, >= len+1] 1 = attend (prompt padding), or null
  bf16* o; long long o_bs;
  int B, H;
  float scale;
};

// One CTA (4 warps) per (batch, head), head dim 64, up to 321 keys.  All K and V rows of the head are requested with cp.async in one
// round trip (16-byte chunks XOR-swizzled by the key index: conflict-free for both access patterns below), then
//   scores : thread <-> key (dot product over the 8 chunks of its row), block-wide max / sum through shared memory,
//   P.V    : warp w takes keys j = w, w+4, ..., lane <-> channels 2*lane, 2*lane+1; the four partial sums meet in shared memory.
// Probabilities are rounded to bf16 before P.V like the tensor-core kernels (scores and accumulation in fp32).
constexpr int kMaxKeys = 328;

__global__ void __launch_bounds__(128) decode_attn_kernel(DecAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PDL_GRID_SYNC();
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = p.len + (p.k_new ? 1 : 0);
  uint8_t* sK = smem_raw;                             // [total][128 B] swizzled
  uint8_t* sV = sK + total * 128;
  float* sP = reinterpret_cast<float*>(sV + total * 128);      // [kMaxKeys] scores -> probabilities
  float* sRed = sP + kMaxKeys;                                 // [8] block reductions, then [4][64] partial outputs
  auto key_ptr = [&](int j) -> const bf16* {
    return j < p.len ? p.k + b * p.kv_bs + static_cast<long long>(j) * p.kv_rs + h * 64 : p.k_new + b * p.new_bs + h * 64;
  };
  auto val_ptr = [&](int j) -> const bf16* {
    return j < p.len ? p.v + b * p.kv_bs + static_cast<long long>(j) * p.kv_rs + h * 64 : p.v_new + b * p.new_bs + h * 64;
  };
  for (int i = tid; i < total * 8; i += 128) {
    const int j = i >> 3, c = i & 7;
    cp_async16(sK + j * 128 + ((c ^ (j & 7)) << 4), key_ptr(j) + c * 8, true);
    cp_async16(sV + j * 128 + ((c ^ (j & 7)) << 4), val_ptr(j) + c * 8, true);
  }
  float qv[64];                                        // every thread keeps the whole query (8 x 16 B broadcast loads)
  {
    const uint4* qp = reinterpret_cast<const uint4*>(p.q + b * p.q_bs + h * 64);
#pragma unroll
    for (int c = 0; c < 8; ++c) unpack8(qp[c], qv + 8 * c);
  }
  cp_async_wait_all();
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < total; j += 128) {
    float sc = -INFINITY;
    if (!(p.key_mask && p.key_mask[static_cast<long long>(b) * p.mask_ld + j] == 0)) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float kv[8];
        unpack8(*reinterpret_cast<const uint4*>(sK + j * 128 + ((c ^ (j & 7)) << 4)), kv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(qv[8 * c + e], kv[e], acc);
      }
      sc = acc * p.scale;
    }
    sP[j] = sc;
    mx = fmaxf(mx, sc);
  }
  mx = warp_max(mx);
  if (lane == 0) sRed[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(sRed[0], sRed[1]), fmaxf(sRed[2], sRed[3]));
  const float msafe = mx == -INFINITY ? 0.f : mx;
  float l = 0.f;
  for (int j = tid; j < total; j += 128) {
    const float sc = sP[j];
    const float e = sc == -INFINITY ? 0.f : __expf(sc - msafe);
    l += e;
    sP[j] = __bfloat162float(__float2bfloat16(e));
  }
  l = warp_sum(l);
  if (lane == 0) sRed[4 + warp] = l;
  __syncthreads();
  l = sRed[4] + sRed[5] + sRed[6] + sRed[7];
  float o0 = 0.f, o1 = 0.f;
  for (int j = warp; j < total; j += 4) {
    const float pj = sP[j];
    const int c = lane >> 2;                             // 16-byte chunk holding channels 2*lane, 2*lane+1
    const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(sV + j * 128 + ((c ^ (j & 7)) << 4) + (lane & 3) * 4);
    o0 = fmaf(pj, __bfloat162float(vv.x), o0);
    o1 = fmaf(pj, __bfloat162float(vv.y), o1);
  }
  __syncthreads();                                       // sRed[0..7] consumed by everyone before it is reused
  float* part = sRed;                                    // [4][64]
  part[warp * 64 + 2 * lane] = o0;
  part[warp * 64 + 2 * lane + 1] = o1;
  __syncthreads();
  if (tid < 64) {
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    const float v = (part[tid] + part[64 + tid] + part[128 + tid] + part[192 + tid]) * inv;
    p.o[b * p.o_bs + h * 64 + tid] = __float2bfloat16(v);
  }
  if (p.k_new && tid >= 64) {   // append the new token's k / v to the cache (row `len`): 32 threads each
    const int t = tid - 64, which = t >> 5, c = t & 31;
    const bf16* src = (which ? p.v_new : p.k_new) + b * p.new_bs + h * 64 + 2 * c;
    bf16* dst = (which ? p.v_cache : p.k_cache) + b * p.c_bs + static_cast<long long>(p.len) * p.c_rs + h * 64 + 2 * c;
    *reinterpret_cast<__nv_bfloat162*>(dst) = *reinterpret_cast<const __nv_bfloat162*>(src);
  }
}

}  // namespace

extern "C" int prismer_skinny_linear(const void* x, long long ldx, const void* w, long long ldw, const float* bias, const void* residual,
                                     long long ldr, void* out, long long ldo, int out_fp32, int M, int N, int K, int act, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 16) || (ldx % 8) || (ldw % 8)) return PRISMER_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return PRISMER_ERR_ALIGN;     // 16-byte cp.async rows
  SkinnyParams p;
  p.x = reinterpret_cast<const bf16*>(x); p.ldx = ldx; p.w = reinterpret_cast<const bf16*>(w); p.ldw = ldw; p.bias = bias;
  p.residual = reinterpret_cast<const bf16*>(residual); p.ldr = ldr; p.out = out; p.ldo = ldo; p.out_fp32 = out_fp32;
  p.M = M; p.N = N; p.K = K; p.act = act;
  // 64 columns per CTA for vocabulary-sized N (the tied LM head), 16 otherwise
  static const bool wide_ok = [] { const char* e = getenv("PRISMER_SKINNY_WIDE"); return !(e && e[0] == '0'); }();
  const int cols = (wide_ok && N >= 8192) ? 64 : 16;
  // K chunk per shared-memory round trip: the whole K (<= 1024) when the grid fits one wave -- a single memory round trip per CTA --,
  // 256 when there are many more CTAs than SMs (LM head): several CTAs per SM hide each other's latency
  const long long ctas = static_cast<long long>((N + cols - 1) / cols) * ((M + kRows - 1) / kRows);
  int kc = K < kKc ? K : kKc;
  if (ctas > 3 * 148 && kc > 256) kc = 256;
  p.kc_len = kc;
  size_t smem = static_cast<size_t>(kRows + cols) * (kc + kPad) * 2;
  const size_t red = static_cast<size_t>(4) * kRows * cols * 4;
  if (smem < red) smem = red;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(skinny_linear_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (kRows + 16) * (kKc + kPad) * 2) != cudaSuccess ||
        cudaFuncSetAttribute(skinny_linear_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (kRows + 64) * (kKc + kPad) * 2) != cudaSuccess)
      return PRISMER_ERR_CUDA;
    configured = true;
  }
  dim3 grid((N + cols - 1) / cols, (M + kRows - 1) / kRows);
  if (cols == 64) pdl_launch(skinny_linear_kernel<64>, grid, dim3(128), smem, stream, p);
  else pdl_launch(skinny_linear_kernel<16>, grid, dim3(128), smem, stream, p);
  return LAUNCH_CHECK();
}

extern "C" int prismer_decode_attention(const void* q, long long q_bs, const void* k, const void* v, long long kv_bs, long long kv_rs,
                                        int len, const void* k_new, const void* v_new, long long new_bs, void* k_cache, void* v_cache,
                                        long long c_bs, long long c_rs, const void* key_mask, int mask_ld, void* o, long long o_bs, int B,
                                        int H, int d, float scale, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || d != 64 || len < 0 || len + (k_new ? 1 : 0) > 320 || len + (k_new ? 1 : 0) <= 0) return PRISMER_ERR_SHAPE;
  if ((q_bs % 8) || (kv_bs % 8) || (kv_rs % 8) || (o_bs % 2)) return PRISMER_ERR_ALIGN;
  if ((k_new != nullptr) != (v_new != nullptr) || (k_new && (!k_cache || !v_cache))) return PRISMER_ERR_SHAPE;
  DecAttnParams p;
  p.q = reinterpret_cast<const bf16*>(q); p.q_bs = q_bs;
  p.k = reinterpret_cast<const bf16*>(k); p.v = reinterpret_cast<const bf16*>(v); p.kv_bs = kv_bs; p.kv_rs = kv_rs; p.len = len;
  p.k_new = reinterpret_cast<const bf16*>(k_new); p.v_new = reinterpret_cast<const bf16*>(v_new); p.new_bs = new_bs;
  p.k_cache = reinterpret_cast<bf16*>(k_cache); p.v_cache = reinterpret_cast<bf16*>(v_cache); p.c_bs = c_bs; p.c_rs = c_rs;
  p.key_mask = reinterpret_cast<const long long*>(key_mask); p.mask_ld = mask_ld;
  p.o = reinterpret_cast<bf16*>(o); p.o_bs = o_bs; p.B = B; p.H = H; p.scale = scale;
  const int total = len + (k_new ? 1 : 0);
  const size_t smem = static_cast<size_t>(total) * 256 + (kMaxKeys + 256) * 4;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 321 * 256 + (kMaxKeys + 256) * 4) != cudaSuccess)
      return PRISMER_ERR_CUDA;
    configured = true;
  }
  pdl_launch(decode_attn_kernel, dim3(B * H), dim3(128), smem, stream, p);
  return LAUNCH_CHECK();
}
