// Fused multi-head attention, forward and backward, flash style (scores never touch HBM).
//   forward : O = dropout(softmax(scale * Q K^T + mask)) V, LSE saved for the backward
//   backward: dQ kernel (CTA per query tile, also emits delta = rowsum(dO*O)), then dK/dV kernel (CTA per key tile)
// Covers the four attention shapes of the hot path:
//   ViT self-attention (vit.py:52-53), resampler cross-attention over cat(latents, experts) (resampler.py:30-31),
//   decoder causal self-attention with padding mask + dropout and vision cross-attention (roberta.py:95-126).
// Masking semantics: the reference adds finfo.min and clamps (roberta.py:113-115); masked probabilities are exactly 0
// as long as a row has one unmasked key, which causal masking guarantees -> identical to -inf masking used here.
// Tensor cores via mma.sync m16n8k16 (bf16 in, fp32 accumulate): tiles are tiny (T<=45, S<=1240, d<=128); the
// attention share of the step is ~5% of FLOPs.
#include "common.cuh"
#include "prismer_sm100.h"

#include <mutex>
#include <unordered_set>

namespace {

constexpr int TQ = 64;      // query rows per CTA (4 warps x 16)
constexpr int TK = 64;      // keys per chunk
constexpr int NWARP = 4;
constexpr int PAD = 8;      // bf16 padding per smem row (keeps 16 B alignment, breaks ldmatrix bank conflicts)

struct AttnParams {
  const bf16* q; const bf16* k; const bf16* v; bf16* o;
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;  // batch / row strides in elements; head h at offset h*d
  float* lse;                 // [B, H, Lq]
  const long long* kmask;     // [B, Lk] key padding mask (1 = attend) or null
  int B, H, Lq, Lk;
  int causal;
  float scale;
  float drop_p; uint32_t thr16; float drop_scale; const unsigned long long* seed; uint32_t rng_stream;
  // backward
  const bf16* dout; long long do_bs, do_rs;
  bf16* dq; long long dq_bs, dq_rs;
  bf16* dk; long long dk_bs, dk_rs;
  bf16* dv; long long dv_bs, dv_rs;
  float* delta;               // [B, H, Lq]
  int kv_div;                 // forward: K/V batch index = b / kv_div (>= 1)
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Cooperative load of a [rows<=64, D] tile (row stride rs) into padded smem; rows >= nvalid are zero-filled.
template <int D>
__device__ __forceinline__ void load_tile(bf16* s, const bf16* g, long long rs, int nvalid) {
  constexpr int VPR = D / 8;
  for (int i = threadIdx.x; i < 64 * VPR; i += NWARP * 32) {
    const int r = i / VPR, c = i % VPR;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < nvalid) val = *reinterpret_cast<const uint4*>(g + static_cast<long long>(r) * rs + c * 8);
    *reinterpret_cast<uint4*>(s + r * (D + PAD) + c * 8) = val;
  }
}

// Asynchronous variant (cp.async.cg, 16 B, zero-fill for rows >= nvalid): the copy of chunk i+1 overlaps the math of chunk i.
template <int D>
__device__ __forceinline__ void load_tile_async(bf16* s, const bf16* g, long long rs, int nvalid) {
  constexpr int VPR = D / 8;
  for (int i = threadIdx.x; i < 64 * VPR; i += NWARP * 32) {
    const int r = i / VPR, c = i % VPR;
    const bool ok = r < nvalid;
    const bf16* src = ok ? g + static_cast<long long>(r) * rs + c * 8 : g;
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(s + r * (D + PAD) + c * 8));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A-operand fragments of a 16-row slab (rows r0..r0+15) of a [64, D] smem tile, for k-step kk (16 columns).
template <int D>
__device__ __forceinline__ void frag_a(uint32_t (&a)[4], const bf16* s, int r0, int kk, int lane) {
  ldsm_x4(a, s + (r0 + (lane & 15)) * (D + PAD) + kk * 16 + (lane >> 4) * 8);
}
// B-operand (k contiguous in smem: tile is [n][k]) for two adjacent n-tiles (16 n) at k-step kk:
// r[0],r[1] = b0,b1 of n-tile 2j ; r[2],r[3] = b0,b1 of n-tile 2j+1
template <int D>
__device__ __forceinline__ void frag_b_nk(uint32_t (&r)[4], const bf16* s, int n0, int kk, int lane) {
  ldsm_x4(r, s + (n0 + (lane & 7) + (lane >> 4) * 8) * (D + PAD) + kk * 16 + ((lane >> 3) & 1) * 8);
}
// B-operand (n contiguous in smem: tile is [k][n]) for k rows k0..k0+15 and two adjacent n-tiles starting at column n0
template <int D>
__device__ __forceinline__ void frag_b_kn(uint32_t (&r)[4], const bf16* s, int k0, int n0, int lane) {
  ldsm_x4_t(r, s + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * (D + PAD) + n0 + (lane >> 4) * 8);
}

__device__ __forceinline__ bool key_ok(const AttnParams& p, int b, int qi, int kj) {
  if (kj >= p.Lk) return false;
  if (p.causal && kj > qi) return false;
  if (p.kmask && p.kmask[static_cast<long long>(b) * p.Lk + kj] == 0) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(NWARP * 32) attn_fwd_kernel(AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PDL_GRID_SYNC();
  constexpr int TILE = 64 * (D + PAD);
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sKV = sQ + TILE;                        // [2 stages][K | V]
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qt * TQ;
  const bf16* qg = p.q + b * p.q_bs + static_cast<long long>(q0) * p.q_rs + h * D;
  int kend = p.Lk;
  if (p.causal) kend = min(p.Lk, q0 + TQ);
  const bf16* kbase = p.k + (b / p.kv_div) * p.k_bs + h * D;
  const bf16* vbase = p.v + (b / p.kv_div) * p.v_bs + h * D;
  // prologue: chunk 0 in flight while Q is staged
  load_tile_async<D>(sKV, kbase, p.k_rs, min(TK, p.Lk));
  load_tile_async<D>(sKV + TILE, vbase, p.v_rs, min(TK, p.Lk));
  cp_async_commit();
  load_tile<D>(sQ, qg, p.q_rs, min(TQ, p.Lq - q0));
  __syncthreads();
  uint32_t qa[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) frag_a<D>(qa[kk], sQ, warp * 16, kk, lane);

  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  const int row[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  const bool has_drop = p.drop_p > 0.f;
  const Philox philox(has_drop ? *p.seed : 0ull);
  const unsigned long long kgroups = (p.Lk + 7) >> 3;
  const bool wact = q0 + warp * 16 < p.Lq;      // warps whose 16 query rows are all padding skip the math

  int stage = 0;
  for (int k0 = 0; k0 < kend; k0 += TK, stage ^= 1) {
    const bf16* sK = sKV + stage * 2 * TILE;
    const bf16* sV = sK + TILE;
    if (k0 + TK < kend) {                       // prefetch the next chunk into the other stage
      bf16* nK = sKV + (stage ^ 1) * 2 * TILE;
      load_tile_async<D>(nK, kbase + static_cast<long long>(k0 + TK) * p.k_rs, p.k_rs, min(TK, p.Lk - k0 - TK));
      load_tile_async<D>(nK + TILE, vbase + static_cast<long long>(k0 + TK) * p.v_rs, p.v_rs, min(TK, p.Lk - k0 - TK));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    int kvalid = min(TK, p.Lk - k0);
    if (p.causal) kvalid = min(kvalid, q0 + warp * 16 + 16 - k0);
    if (wact && kvalid > 0) {
    float s[TK / 8][4];
#pragma unroll
    for (int j = 0; j < TK / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
      for (int j2 = 0; j2 < TK / 16; ++j2) {
        if (j2 * 16 >= kvalid) continue;          // warp-uniform: key columns past Lk (or past the causal frontier)
        uint32_t bk[4];
        frag_b_nk<D>(bk, sK, j2 * 16, kk, lane);
        mma16816(s[2 * j2], qa[kk], bk[0], bk[1]);
        mma16816(s[2 * j2 + 1], qa[kk], bk[2], bk[3]);
      }
    }
    // scale + mask, running max
    float mx[2] = {m[0], m[1]};
    const bool need_mask = p.causal || p.kmask != nullptr || k0 + TK > p.Lk;   // warp-uniform fast path for interior chunks
    if (need_mask) {
#pragma unroll
      for (int j = 0; j < TK / 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const int kj = k0 + j * 8 + 2 * t + (e & 1);
          const float val = key_ok(p, b, row[r], kj) ? s[j][e] * p.scale : -INFINITY;
          s[j][e] = val;
          mx[r] = fmaxf(mx[r], val);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < TK / 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s[j][e] *= p.scale;
          mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], msafe[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      msafe[r] = mx[r] == -INFINITY ? 0.f : mx[r];
      corr[r] = __expf(m[r] - msafe[r]);  // m = -inf -> 0
      m[r] = mx[r];
      l[r] *= corr[r];
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < TK / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = __expf(s[j][e] - msafe[e >> 1]);
        s[j][e] = pv;
        ls[e >> 1] += pv;
      }
    }
    l[0] += ls[0]; l[1] += ls[1];   // per-thread partial; reduced across the quad at the end
    if (has_drop) {
#pragma unroll
      for (int j = 0; j < TK / 8; ++j) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const unsigned long long grow = (static_cast<unsigned long long>(b) * p.H + h) * p.Lq + row[r];
          const uint32_t keep = dropout_keep8(philox, grow * kgroups + ((k0 >> 3) + j), p.rng_stream, p.thr16);
          s[j][2 * r] = ((keep >> (2 * t)) & 1u) ? s[j][2 * r] * p.drop_scale : 0.f;
          s[j][2 * r + 1] = ((keep >> (2 * t + 1)) & 1u) ? s[j][2 * r + 1] * p.drop_scale : 0.f;
        }
      }
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < TK / 16; ++kk) {
      if (kk * 16 >= kvalid) continue;
      uint32_t pa[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                        pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
      for (int n2 = 0; n2 < D / 16; ++n2) {
        uint32_t bv[4];
        frag_b_kn<D>(bv, sV, kk * 16, n2 * 16, lane);
        mma16816(o[2 * n2], pa, bv[0], bv[1]);
        mma16816(o[2 * n2 + 1], pa, bv[2], bv[3]);
      }
    }
    }  // wact
    __syncthreads();   // every warp is done with this stage before the next iteration's prefetch may overwrite it
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (row[r] < p.Lq) {
      const float inv = l[r] > 0.f ? 1.0f / l[r] : 0.f;
      bf16* og = p.o + b * p.o_bs + static_cast<long long>(row[r]) * p.o_rs + h * D;
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        __nv_bfloat162 v2 = __floats2bfloat162_rn(o[i][2 * r] * inv, o[i][2 * r + 1] * inv);
        *reinterpret_cast<__nv_bfloat162*>(og + i * 8 + 2 * t) = v2;
      }
      if (p.lse && t == 0) p.lse[(static_cast<long long>(b) * p.H + h) * p.Lq + row[r]] = m[r] + __logf(l[r]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: dQ (+delta)
template <int D>
__global__ void __launch_bounds__(NWARP * 32) attn_bwd_dq_kernel(AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PDL_GRID_SYNC();
  constexpr int TILE = 64 * (D + PAD);
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sdO = sQ + TILE;
  bf16* sKV = sdO + TILE;                       // [2 stages][K | V]
  bf16* sO = sKV + 2 * TILE;                    // O is staged in stage 1's K slot during the prologue
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qt * TQ;
  const int nq = min(TQ, p.Lq - q0);
  int kend = p.Lk;
  if (p.causal) kend = min(p.Lk, q0 + TQ);
  const bf16* kbase = p.k + b * p.k_bs + h * D;
  const bf16* vbase = p.v + b * p.v_bs + h * D;
  load_tile_async<D>(sKV, kbase, p.k_rs, min(TK, p.Lk));
  load_tile_async<D>(sKV + TILE, vbase, p.v_rs, min(TK, p.Lk));
  cp_async_commit();
  load_tile<D>(sQ, p.q + b * p.q_bs + static_cast<long long>(q0) * p.q_rs + h * D, p.q_rs, nq);
  load_tile<D>(sdO, p.dout + b * p.do_bs + static_cast<long long>(q0) * p.do_rs + h * D, p.do_rs, nq);
  load_tile<D>(sO, p.o + b * p.o_bs + static_cast<long long>(q0) * p.o_rs + h * D, p.o_rs, nq);
  __syncthreads();
  // delta = rowsum(dO * O): each warp its 16 rows, 2 lanes per row
  const long long bh = static_cast<long long>(b) * p.H + h;
  {
    const int r = warp * 16 + (lane >> 1);
    float acc = 0.f;
    for (int c = (lane & 1) * (D / 2); c < ((lane & 1) + 1) * (D / 2); ++c)
      acc += __bfloat162float(sdO[r * (D + PAD) + c]) * __bfloat162float(sO[r * (D + PAD) + c]);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if ((lane & 1) == 0 && q0 + r < p.Lq) p.delta[bh * p.Lq + q0 + r] = acc;
  }
  const int row[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  float lse[2], dl[2];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const bool ok = row[r] < p.Lq;
    lse[r] = ok ? p.lse[bh * p.Lq + row[r]] : 0.f;
    dl[r] = ok ? p.delta[bh * p.Lq + row[r]] : 0.f;   // written above by this warp's lanes (same CTA, after sync)
  }
  float dq[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const bool has_drop = p.drop_p > 0.f;
  const Philox philox(has_drop ? *p.seed : 0ull);
  const unsigned long long kgroups = (p.Lk + 7) >> 3;
  const bool wact = q0 + warp * 16 < p.Lq;

  int stage = 0;
  for (int k0 = 0; k0 < kend; k0 += TK, stage ^= 1) {
    const bf16* sK = sKV + stage * 2 * TILE;
    const bf16* sV = sK + TILE;
    if (k0 + TK < kend) {
      bf16* nK = sKV + (stage ^ 1) * 2 * TILE;
      load_tile_async<D>(nK, kbase + static_cast<long long>(k0 + TK) * p.k_rs, p.k_rs, min(TK, p.Lk - k0 - TK));
      load_tile_async<D>(nK + TILE, vbase + static_cast<long long>(k0 + TK) * p.v_rs, p.v_rs, min(TK, p.Lk - k0 - TK));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    int kvalid = min(TK, p.Lk - k0);
    if (p.causal) kvalid = min(kvalid, q0 + warp * 16 + 16 - k0);
    if (wact && kvalid > 0) {
    float s[TK / 8][4], dp[TK / 8][4];
#pragma unroll
    for (int j = 0; j < TK / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t qa[4], doa[4];
      frag_a<D>(qa, sQ, warp * 16, kk, lane);
      frag_a<D>(doa, sdO, warp * 16, kk, lane);
#pragma unroll
      for (int j2 = 0; j2 < TK / 16; ++j2) {
        if (j2 * 16 >= kvalid) continue;
        uint32_t bk[4], bv[4];
        frag_b_nk<D>(bk, sK, j2 * 16, kk, lane);
        mma16816(s[2 * j2], qa, bk[0], bk[1]);
        mma16816(s[2 * j2 + 1], qa, bk[2], bk[3]);
        frag_b_nk<D>(bv, sV, j2 * 16, kk, lane);
        mma16816(dp[2 * j2], doa, bv[0], bv[1]);
        mma16816(dp[2 * j2 + 1], doa, bv[2], bv[3]);
      }
    }
    // dS = P * (dP_masked - delta) * scale
#pragma unroll
    for (int j = 0; j < TK / 8; ++j) {
      uint32_t keep[2] = {0xffffffffu, 0xffffffffu};
      if (has_drop) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
          keep[r] = dropout_keep8(philox, (bh * p.Lq + row[r]) * kgroups + ((k0 >> 3) + j), p.rng_stream, p.thr16);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int kj = k0 + j * 8 + 2 * t + (e & 1);
        float pv = 0.f;
        if (row[r] < p.Lq && key_ok(p, b, row[r], kj)) pv = __expf(s[j][e] * p.scale - lse[r]);
        float dpv = dp[j][e];
        if (has_drop) dpv = ((keep[r] >> (2 * t + (e & 1))) & 1u) ? dpv * p.drop_scale : 0.f;
        s[j][e] = pv * (dpv - dl[r]) * p.scale;
      }
    }
    // dQ += dS K   (B operand: K tile is [key k][d n] -> n contiguous -> transposed ldmatrix)
#pragma unroll
    for (int kk = 0; kk < TK / 16; ++kk) {
      if (kk * 16 >= kvalid) continue;
      uint32_t pa[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                        pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
      for (int n2 = 0; n2 < D / 16; ++n2) {
        uint32_t bk[4];
        frag_b_kn<D>(bk, sK, kk * 16, n2 * 16, lane);
        mma16816(dq[2 * n2], pa, bk[0], bk[1]);
        mma16816(dq[2 * n2 + 1], pa, bk[2], bk[3]);
      }
    }
    }  // wact
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (row[r] < p.Lq) {
      bf16* dg = p.dq + b * p.dq_bs + static_cast<long long>(row[r]) * p.dq_rs + h * D;
#pragma unroll
      for (int i = 0; i < D / 8; ++i)
        *reinterpret_cast<__nv_bfloat162*>(dg + i * 8 + 2 * t) = __floats2bfloat162_rn(dq[i][2 * r], dq[i][2 * r + 1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
template <int D>
__global__ void __launch_bounds__(NWARP * 32, (D <= 64) ? 3 : 1) attn_bwd_dkv_kernel(AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PDL_GRID_SYNC();
  constexpr int TILE = 64 * (D + PAD);
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);
  bf16* sV = sK + TILE;
  bf16* sQdO = sV + TILE;                       // [2 stages][Q | dO]
  float* sStat = reinterpret_cast<float*>(sQdO + 4 * TILE);   // [2 stages][lse(64) | delta(64)]
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int k0 = kt * TK;
  const int nk = min(TK, p.Lk - k0);
  const long long bh = static_cast<long long>(b) * p.H + h;
  load_tile<D>(sK, p.k + b * p.k_bs + static_cast<long long>(k0) * p.k_rs + h * D, p.k_rs, nk);
  load_tile<D>(sV, p.v + b * p.v_bs + static_cast<long long>(k0) * p.v_rs + h * D, p.v_rs, nk);
  __syncthreads();
  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
  const int key[2] = {k0 + warp * 16 + g, k0 + warp * 16 + g + 8};
  bool kvalid[2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
    kvalid[r] = key[r] < p.Lk && !(p.kmask && p.kmask[static_cast<long long>(b) * p.Lk + key[r]] == 0);
  const bool has_drop = p.drop_p > 0.f;
  const Philox philox(has_drop ? *p.seed : 0ull);
  const unsigned long long kgroups = (p.Lk + 7) >> 3;
  const bool kact = k0 + warp * 16 < p.Lk;     // warps whose 16 keys are all padding skip the math

  const int qstart = p.causal ? (k0 / TQ) * TQ : 0;
  auto prefetch = [&](int q0, int st) {
    const int nq = min(TQ, p.Lq - q0);
    bf16* dQ = sQdO + st * 2 * TILE;
    load_tile_async<D>(dQ, p.q + b * p.q_bs + static_cast<long long>(q0) * p.q_rs + h * D, p.q_rs, nq);
    load_tile_async<D>(dQ + TILE, p.dout + b * p.do_bs + static_cast<long long>(q0) * p.do_rs + h * D, p.do_rs, nq);
    cp_async_commit();
    if (threadIdx.x < 64) {
      const bool ok = threadIdx.x < nq;
      sStat[st * 128 + threadIdx.x] = ok ? p.lse[bh * p.Lq + q0 + threadIdx.x] : 0.f;
      sStat[st * 128 + 64 + threadIdx.x] = ok ? p.delta[bh * p.Lq + q0 + threadIdx.x] : 0.f;
    }
  };
  prefetch(qstart, 0);
  int stage = 0;
  for (int q0 = qstart; q0 < p.Lq; q0 += TQ, stage ^= 1) {
    const bf16* sQ = sQdO + stage * 2 * TILE;
    const bf16* sdO = sQ + TILE;
    const float* sLse = sStat + stage * 128;
    const float* sDl = sLse + 64;
    if (q0 + TQ < p.Lq) {
      prefetch(q0 + TQ, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    // S^T = K Q^T (16 keys x 64 queries per warp), dP^T = V dO^T
    const int nqv = min(TQ, p.Lq - q0);
    if (kact) {
    float s[TQ / 8][4], dp[TQ / 8][4];
#pragma unroll
    for (int j = 0; j < TQ / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t ka[4], va[4];
      frag_a<D>(ka, sK, warp * 16, kk, lane);
      frag_a<D>(va, sV, warp * 16, kk, lane);
#pragma unroll
      for (int j2 = 0; j2 < TQ / 16; ++j2) {
        if (j2 * 16 >= nqv) continue;               // query columns past Lq
        uint32_t bq[4], bo[4];
        frag_b_nk<D>(bq, sQ, j2 * 16, kk, lane);
        mma16816(s[2 * j2], ka, bq[0], bq[1]);
        mma16816(s[2 * j2 + 1], ka, bq[2], bq[3]);
        frag_b_nk<D>(bo, sdO, j2 * 16, kk, lane);
        mma16816(dp[2 * j2], va, bo[0], bo[1]);
        mma16816(dp[2 * j2 + 1], va, bo[2], bo[3]);
      }
    }
    // P^T (dropped for dV), dS^T
    float pd[TQ / 8][4];
#pragma unroll
    for (int j = 0; j < TQ / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;                       // key index within the pair
        const int qi_l = j * 8 + 2 * t + (e & 1);   // local query index
        const int qi = q0 + qi_l;
        float pv = 0.f;
        if (qi < p.Lq && kvalid[r] && !(p.causal && key[r] > qi)) pv = __expf(s[j][e] * p.scale - sLse[qi_l]);
        float keepf = 1.f;
        if (has_drop && pv != 0.f) {
          const uint32_t keep = dropout_keep8(philox, (bh * p.Lq + qi) * kgroups + (key[r] >> 3), p.rng_stream, p.thr16);
          keepf = ((keep >> (key[r] & 7)) & 1u) ? p.drop_scale : 0.f;
        }
        pd[j][e] = pv * keepf;
        s[j][e] = pv * (dp[j][e] * keepf - sDl[qi_l]) * p.scale;
      }
    }
    // dV += Pd^T dO ; dK += dS^T Q    (B operands [k = query][n = d]: n contiguous -> transposed ldmatrix)
#pragma unroll
    for (int kk = 0; kk < TQ / 16; ++kk) {
      if (kk * 16 >= nqv) continue;
      uint32_t pa[4] = {pack2(pd[2 * kk][0], pd[2 * kk][1]), pack2(pd[2 * kk][2], pd[2 * kk][3]),
                        pack2(pd[2 * kk + 1][0], pd[2 * kk + 1][1]), pack2(pd[2 * kk + 1][2], pd[2 * kk + 1][3])};
      uint32_t sa[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                        pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
      for (int n2 = 0; n2 < D / 16; ++n2) {
        uint32_t bo[4], bq[4];
        frag_b_kn<D>(bo, sdO, kk * 16, n2 * 16, lane);
        mma16816(dv[2 * n2], pa, bo[0], bo[1]);
        mma16816(dv[2 * n2 + 1], pa, bo[2], bo[3]);
        frag_b_kn<D>(bq, sQ, kk * 16, n2 * 16, lane);
        mma16816(dk[2 * n2], sa, bq[0], bq[1]);
        mma16816(dk[2 * n2 + 1], sa, bq[2], bq[3]);
      }
    }
    }  // kact
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (key[r] < p.Lk) {
      bf16* kg = p.dk + b * p.dk_bs + static_cast<long long>(key[r]) * p.dk_rs + h * D;
      bf16* vg = p.dv + b * p.dv_bs + static_cast<long long>(key[r]) * p.dv_rs + h * D;
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        *reinterpret_cast<__nv_bfloat162*>(kg + i * 8 + 2 * t) = __floats2bfloat162_rn(dk[i][2 * r], dk[i][2 * r + 1]);
        *reinterpret_cast<__nv_bfloat162*>(vg + i * 8 + 2 * t) = __floats2bfloat162_rn(dv[i][2 * r], dv[i][2 * r + 1]);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------ backward, Lq <= 64 fused
// One CTA per (b, h) when all queries fit one tile (resampler: 64 latents; decoder: T <= 45): S, P, dP, dS are computed ONCE
// per key chunk; dQ accumulates in registers over the chunks, and since every query is in this CTA, dK/dV of a chunk are
// complete after that chunk and are written straight out (5 matrix products per chunk instead of 7 over two kernels, half the
// loads, half the dropout RNG, one launch).  P^T / dS^T reach the second pair of products through shared memory.
template <int D>
__device__ __forceinline__ void frag_a_t(uint32_t (&a)[4], const bf16* s, int m0, int k0, int lane, int ld) {
  // A[m][k] fragment (16x16) from a tile stored [k][m] (m contiguous): transposed ldmatrix
  ldsm_x4_t(a, s + (k0 + (lane & 7) + ((lane >> 4) & 1) * 8) * ld + m0 + ((lane >> 3) & 1) * 8);
}

template <int D>
__global__ void __launch_bounds__(NWARP * 32) attn_bwd_fused_kernel(AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PDL_GRID_SYNC();
  constexpr int TILE = 64 * (D + PAD);
  constexpr int PLD = TK + PAD;
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sdO = sQ + TILE;
  bf16* sKV = sdO + TILE;                       // [2 stages][K | V]
  bf16* sP = sKV + 4 * TILE;                    // [64 q][TK keys] dropped probabilities
  bf16* sdS = sP + 64 * PLD;                    // [64 q][TK keys] score gradients
  float* sDelta = reinterpret_cast<float*>(sdS + 64 * PLD);
  bf16* sO = sKV + 2 * TILE;                    // prologue only (stage 1's K slot)
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nq = p.Lq;                          // <= 64
  const long long bh = static_cast<long long>(b) * p.H + h;
  const bf16* kbase = p.k + b * p.k_bs + h * D;
  const bf16* vbase = p.v + b * p.v_bs + h * D;
  load_tile_async<D>(sKV, kbase, p.k_rs, min(TK, p.Lk));
  load_tile_async<D>(sKV + TILE, vbase, p.v_rs, min(TK, p.Lk));
  cp_async_commit();
  load_tile<D>(sQ, p.q + b * p.q_bs + h * D, p.q_rs, nq);
  load_tile<D>(sdO, p.dout + b * p.do_bs + h * D, p.do_rs, nq);
  load_tile<D>(sO, p.o + b * p.o_bs + h * D, p.o_rs, nq);
  __syncthreads();
  {
    const int r = warp * 16 + (lane >> 1);
    float acc = 0.f;
    for (int c = (lane & 1) * (D / 2); c < ((lane & 1) + 1) * (D / 2); ++c)
      acc += __bfloat162float(sdO[r * (D + PAD) + c]) * __bfloat162float(sO[r * (D + PAD) + c]);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if ((lane & 1) == 0) sDelta[r] = acc;
  }
  __syncthreads();
  const int row[2] = {warp * 16 + g, warp * 16 + g + 8};
  float lse[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lse[r] = row[r] < nq ? p.lse[bh * p.Lq + row[r]] : 0.f;
    dl[r] = sDelta[row[r]];
  }
  float dq[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const bool has_drop = p.drop_p > 0.f;
  const Philox philox(has_drop ? *p.seed : 0ull);
  const unsigned long long kgroups = (p.Lk + 7) >> 3;
  const bool wact = warp * 16 < nq;
  const int kend = p.causal ? min(p.Lk, TQ) : p.Lk;

  int stage = 0;
  for (int k0 = 0; k0 < kend; k0 += TK, stage ^= 1) {
    const bf16* sK = sKV + stage * 2 * TILE;
    const bf16* sV = sK + TILE;
    if (k0 + TK < kend) {
      bf16* nK = sKV + (stage ^ 1) * 2 * TILE;
      load_tile_async<D>(nK, kbase + static_cast<long long>(k0 + TK) * p.k_rs, p.k_rs, min(TK, p.Lk - k0 - TK));
      load_tile_async<D>(nK + TILE, vbase + static_cast<long long>(k0 + TK) * p.v_rs, p.v_rs, min(TK, p.Lk - k0 - TK));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int kvalid = min(TK, p.Lk - k0);
    // ---- phase A: this warp's 16 query rows x the chunk's keys
    {
      float s[TK / 8][4], dp[TK / 8][4];
#pragma unroll
      for (int j = 0; j < TK / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
      if (wact) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          uint32_t qa[4], doa[4];
          frag_a<D>(qa, sQ, warp * 16, kk, lane);
          frag_a<D>(doa, sdO, warp * 16, kk, lane);
#pragma unroll
          for (int j2 = 0; j2 < TK / 16; ++j2) {
            if (j2 * 16 >= kvalid) continue;
            uint32_t bk[4], bv[4];
            frag_b_nk<D>(bk, sK, j2 * 16, kk, lane);
            mma16816(s[2 * j2], qa, bk[0], bk[1]);
            mma16816(s[2 * j2 + 1], qa, bk[2], bk[3]);
            frag_b_nk<D>(bv, sV, j2 * 16, kk, lane);
            mma16816(dp[2 * j2], doa, bv[0], bv[1]);
            mma16816(dp[2 * j2 + 1], doa, bv[2], bv[3]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < TK / 8; ++j) {
        uint32_t keep[2] = {0xffffffffu, 0xffffffffu};
        if (has_drop && wact) {
#pragma unroll
          for (int r = 0; r < 2; ++r)
            keep[r] = dropout_keep8(philox, (bh * p.Lq + row[r]) * kgroups + ((k0 >> 3) + j), p.rng_stream, p.thr16);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float pdv[2], dsv[2];
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const int e = 2 * r + c2;
            const int kj = k0 + j * 8 + 2 * t + c2;
            float pv = 0.f;
            if (wact && row[r] < nq && key_ok(p, b, row[r], kj)) pv = __expf(s[j][e] * p.scale - lse[r]);
            float keepf = 1.f;
            if (has_drop) keepf = ((keep[r] >> (2 * t + c2)) & 1u) ? p.drop_scale : 0.f;
            pdv[c2] = pv * keepf;
            dsv[c2] = pv * (dp[j][e] * keepf - dl[r]) * p.scale;
            s[j][e] = dsv[c2];
          }
          *reinterpret_cast<__nv_bfloat162*>(sP + row[r] * PLD + j * 8 + 2 * t) = __floats2bfloat162_rn(pdv[0], pdv[1]);
          *reinterpret_cast<__nv_bfloat162*>(sdS + row[r] * PLD + j * 8 + 2 * t) = __floats2bfloat162_rn(dsv[0], dsv[1]);
        }
      }
      if (wact) {   // dQ += dS K
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk) {
          if (kk * 16 >= kvalid) continue;
          uint32_t pa[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                            pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
          for (int n2 = 0; n2 < D / 16; ++n2) {
            uint32_t bk[4];
            frag_b_kn<D>(bk, sK, kk * 16, n2 * 16, lane);
            mma16816(dq[2 * n2], pa, bk[0], bk[1]);
            mma16816(dq[2 * n2 + 1], pa, bk[2], bk[3]);
          }
        }
      }
    }
    __syncthreads();
    // ---- phase B: this warp's 16 keys of the chunk: dV = P^T dO, dK = dS^T Q (reduction over all queries of the CTA)
    if (warp * 16 < kvalid) {
      float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < TQ / 16; ++kk) {
        if (kk * 16 >= nq) continue;
        uint32_t pa[4], sa[4];
        frag_a_t<D>(pa, sP, warp * 16, kk * 16, lane, PLD);
        frag_a_t<D>(sa, sdS, warp * 16, kk * 16, lane, PLD);
#pragma unroll
        for (int n2 = 0; n2 < D / 16; ++n2) {
          uint32_t bo[4], bq[4];
          frag_b_kn<D>(bo, sdO, kk * 16, n2 * 16, lane);
          mma16816(dv[2 * n2], pa, bo[0], bo[1]);
          mma16816(dv[2 * n2 + 1], pa, bo[2], bo[3]);
          frag_b_kn<D>(bq, sQ, kk * 16, n2 * 16, lane);
          mma16816(dk[2 * n2], sa, bq[0], bq[1]);
          mma16816(dk[2 * n2 + 1], sa, bq[2], bq[3]);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int key = k0 + warp * 16 + g + 8 * r;
        if (key < p.Lk) {
          bf16* kg = p.dk + b * p.dk_bs + static_cast<long long>(key) * p.dk_rs + h * D;
          bf16* vg = p.dv + b * p.dv_bs + static_cast<long long>(key) * p.dv_rs + h * D;
#pragma unroll
          for (int i = 0; i < D / 8; ++i) {
            *reinterpret_cast<__nv_bfloat162*>(kg + i * 8 + 2 * t) = __floats2bfloat162_rn(dk[i][2 * r], dk[i][2 * r + 1]);
            *reinterpret_cast<__nv_bfloat162*>(vg + i * 8 + 2 * t) = __floats2bfloat162_rn(dv[i][2 * r], dv[i][2 * r + 1]);
          }
        }
      }
    }
    __syncthreads();
  }
  // keys of chunks that were never visited (causal: none beyond the tile) need zero gradients
  if (p.causal && kend < p.Lk) {
    for (int key = kend + warp; key < p.Lk; key += NWARP)
      for (int c = lane; c < D; c += 32) {
        p.dk[b * p.dk_bs + static_cast<long long>(key) * p.dk_rs + h * D + c] = __float2bfloat16(0.f);
        p.dv[b * p.dv_bs + static_cast<long long>(key) * p.dv_rs + h * D + c] = __float2bfloat16(0.f);
      }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (row[r] < nq) {
      bf16* dg = p.dq + b * p.dq_bs + static_cast<long long>(row[r]) * p.dq_rs + h * D;
#pragma unroll
      for (int i = 0; i < D / 8; ++i)
        *reinterpret_cast<__nv_bfloat162*>(dg + i * 8 + 2 * t) = __floats2bfloat162_rn(dq[i][2 * r], dq[i][2 * r + 1]);
    }
  }
}

template <int D> int smem_fused() { return 6 * 64 * (D + PAD) * 2 + 2 * 64 * (TK + PAD) * 2 + 64 * 4; }

template <int D> int smem_fwd() { return 5 * 64 * (D + PAD) * 2; }
template <int D> int smem_dq() { return 6 * 64 * (D + PAD) * 2; }
template <int D> int smem_dkv() { return 6 * 64 * (D + PAD) * 2 + 4 * 64 * 4; }

template <typename K>
int set_smem(K kern, int bytes) {
  // opt in to > 48 KiB dynamic shared memory once per kernel (keeps stream capture free of attribute calls)
  static std::mutex mu;
  static std::unordered_set<const void*> done;
  if (bytes <= 48 * 1024) return PRISMER_OK;
  std::lock_guard<std::mutex> lock(mu);
  const void* key = reinterpret_cast<const void*>(kern);
  if (done.count(key)) return PRISMER_OK;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return PRISMER_ERR_CUDA;
  done.insert(key);
  return PRISMER_OK;
}

int fill(AttnParams& p, const PrismerAttnArgs* a) {
  if (!a || a->B <= 0 || a->H <= 0 || a->Lq <= 0 || a->Lk <= 0) return PRISMER_ERR_SHAPE;
  if (a->d != 32 && a->d != 64 && a->d != 96 && a->d != 128) return PRISMER_ERR_SHAPE;
  const long long strides[] = {a->q_bs, a->q_rs, a->k_bs, a->k_rs, a->v_bs, a->v_rs, a->o_bs, a->o_rs};
  for (long long s : strides) if (s % 8) return PRISMER_ERR_ALIGN;
  p.q = (const bf16*)a->q; p.k = (const bf16*)a->k; p.v = (const bf16*)a->v; p.o = (bf16*)a->o;
  p.q_bs = a->q_bs; p.q_rs = a->q_rs; p.k_bs = a->k_bs; p.k_rs = a->k_rs; p.v_bs = a->v_bs; p.v_rs = a->v_rs;
  p.o_bs = a->o_bs; p.o_rs = a->o_rs;
  p.lse = a->lse; p.kmask = (const long long*)a->key_mask;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk; p.causal = a->causal; p.scale = a->scale;
  p.drop_p = a->drop_p; p.thr16 = static_cast<uint32_t>(a->drop_p * 65536.0f + 0.5f);
  p.drop_scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  p.seed = a->seed; p.rng_stream = a->rng_stream;
  if (a->drop_p > 0.f && !a->seed) return PRISMER_ERR_SHAPE;
  p.dout = (const bf16*)a->dout; p.do_bs = a->do_bs; p.do_rs = a->do_rs;
  p.dq = (bf16*)a->dq; p.dq_bs = a->dq_bs; p.dq_rs = a->dq_rs;
  p.dk = (bf16*)a->dk; p.dk_bs = a->dk_bs; p.dk_rs = a->dk_rs;
  p.dv = (bf16*)a->dv; p.dv_bs = a->dv_bs; p.dv_rs = a->dv_rs;
  p.delta = a->delta;
  p.kv_div = a->kv_div > 1 ? a->kv_div : 1;
  if (p.kv_div > 1 && (p.kmask || a->B % p.kv_div)) return PRISMER_ERR_SHAPE;     // key masks are per query batch row
  return PRISMER_OK;
}

}  // namespace

// tcgen05 / TMEM kernels for the encoder's self-attention shapes (attention_sm100.cu): 1 = handled, 0 = shape not covered
int attn_sm100_try_fwd(const PrismerAttnArgs* a, cudaStream_t stream, int* rc_out);
int attn_sm100_try_bwd(const PrismerAttnArgs* a, cudaStream_t stream, int* rc_out);

extern "C" int prismer_attention_fwd(const PrismerAttnArgs* a, cudaStream_t stream) {
  AttnParams p;
  int rc = fill(p, a);
  if (rc) return rc;
  if (attn_sm100_try_fwd(a, stream, &rc)) return rc;
  dim3 grid((p.Lq + TQ - 1) / TQ, p.H, p.B);
#define FWD(D_)                                                                     \
  { rc = set_smem(attn_fwd_kernel<D_>, smem_fwd<D_>()); if (rc) return rc;          \
    pdl_launch(attn_fwd_kernel<D_>, grid, dim3(NWARP * 32), static_cast<size_t>(smem_fwd<D_>()), stream, p); }
  switch (a->d) { case 32: FWD(32) break; case 64: FWD(64) break; case 96: FWD(96) break; default: FWD(128) break; }
#undef FWD
  return LAUNCH_CHECK();
}

extern "C" int prismer_attention_bwd(const PrismerAttnArgs* a, cudaStream_t stream) {
  AttnParams p;
  int rc = fill(p, a);
  if (rc) return rc;
  if (!p.dout || !p.dq || !p.dk || !p.dv || !p.delta || !p.lse || p.kv_div > 1) return PRISMER_ERR_SHAPE;
  const long long strides[] = {a->do_bs, a->do_rs, a->dq_bs, a->dq_rs, a->dk_bs, a->dk_rs, a->dv_bs, a->dv_rs};
  for (long long s : strides) if (s % 8) return PRISMER_ERR_ALIGN;
  if (attn_sm100_try_bwd(a, stream, &rc)) return rc;
  if (p.Lq <= TQ) {   // all queries in one tile: fused single-kernel backward
    dim3 gf(1, p.H, p.B);
#define BWDF(D_)                                                                                 \
  { rc = set_smem(attn_bwd_fused_kernel<D_>, smem_fused<D_>()); if (rc) return rc;               \
    pdl_launch(attn_bwd_fused_kernel<D_>, gf, dim3(NWARP * 32), static_cast<size_t>(smem_fused<D_>()), stream, p); }
    switch (a->d) { case 32: BWDF(32) break; case 64: BWDF(64) break; case 96: BWDF(96) break; default: BWDF(128) break; }
#undef BWDF
    return LAUNCH_CHECK();
  }
  dim3 gq((p.Lq + TQ - 1) / TQ, p.H, p.B), gk((p.Lk + TK - 1) / TK, p.H, p.B);
#define BWD(D_)                                                                                  \
  { rc = set_smem(attn_bwd_dq_kernel<D_>, smem_dq<D_>()); if (rc) return rc;                     \
    rc = set_smem(attn_bwd_dkv_kernel<D_>, smem_dkv<D_>()); if (rc) return rc;                   \
    pdl_launch(attn_bwd_dq_kernel<D_>, gq, dim3(NWARP * 32), static_cast<size_t>(smem_dq<D_>()), stream, p);                        \
    pdl_launch(attn_bwd_dkv_kernel<D_>, gk, dim3(NWARP * 32), static_cast<size_t>(smem_dkv<D_>()), stream, p); }
  switch (a->d) { case 32: BWD(32) break; case 64: BWD(64) break; case 96: BWD(96) break; default: BWD(128) break; }
#undef BWD
  return LAUNCH_CHECK();
}
