// Blackwell-native multi-head attention for the encoder's self-attention (nn.MultiheadAttention core, vit.py:52-53), head dim 64,
// no mask / dropout, Lq, Lk <= 320 (Prismer-BASE @224: S = 196 + 64 = 260; Prismer-LARGE @224: S = 256 + 64 = 320):
// forward and a SINGLE-PASS backward, one CTA per (batch, head), every matrix product on tcgen05.mma with TMEM accumulators.
//
//   forward : S_i = Q_i K^T for ALL keys lands in TMEM (<= 320 fp32 columns) -> softmax warps read it with tcgen05.ld (row max,
//             exp2, row sum), write P (bf16) into 128B-swizzled shared-memory atoms -> O_i = P V (tcgen05.mma, A = P from smem)
//             -> O / l and the log-sum-exp go to global.  The S product of tile i+1 overlaps the O epilogue of tile i.
//   backward: Q, K, V, dO of the head stay in shared memory (TMA, 16-row boxes, rows >= L zero-filled by the TMA unit).  Key tiles j
//             outer, query tiles i inner; per (j, i) and 64-key half h the tensor core produces S^h and dP^h into a ring of three
//             64-column TMEM slots, the softmax warps turn them into P^h and dS^h (bf16, shared memory), and the second-stage
//             products  dV_j += P^T dO_i,  dK_j += dS^T Q_i,  dQ_i += dS K_j  accumulate in TMEM (dQ for all three query tiles is
//             resident: 192 columns; dK_j / dV_j: 128 columns).  S is computed ONCE per tile (the mma.sync kernels it replaces
//             recomputed it in separate dQ and dK/dV passes) and nothing but Q/K/V/dO/O in and dQ/dK/dV out touches HBM.
//
// Every operand tile is "[rows] x 128 bytes, 128B-swizzled", which is both a K-major and an MN-major UMMA operand -- the same
// P / dS / Q / K / V / dO bytes feed products that contract over the 64 columns and products that contract over the rows.
// Replaces (for these shapes) attn_fwd_kernel / attn_bwd_dq_kernel / attn_bwd_dkv_kernel of attention.cu (mma.sync m16n8k16).
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "prismer_sm100.h"

#include <cstdio>
#include <mutex>

namespace {

constexpr int kMaxL = 320;                 // rows of Q / K that fit the shared-memory and TMEM plan
constexpr int kEpiWarps = 8;                // two groups of four (a warp may only touch TMEM lane quarter warp_id % 4)
constexpr int kThreads = 32 * (2 + kEpiWarps);   // warp 0: TMA, warp 1: MMA issue + TMEM owner, warps 2..9: softmax / epilogue
constexpr int kAtomBytes = 128 * 128;      // one [128 rows x 64 bf16] swizzled atom of P / dS
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct TcParams {
  int B, H, Sq, Sk;
  float scale, sl2;                         // softmax scale, scale * log2(e)
  float* lse;                               // [B, H, Sq] natural-log LSE of the scaled scores (written by fwd, read by bwd)
  bf16* o; long long o_bs, o_rs;            // forward output / backward input (delta = rowsum(dO * O))
  const bf16* dout; long long do_bs, do_rs;
  bf16* dq; long long dq_bs, dq_rs;
  bf16* dk; long long dk_bs, dk_rs;
  bf16* dv; long long dv_bs, dv_rs;
  int pos[4];                               // per tensor map: packed coordinate positions (see coords())
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Bounded mbarrier wait: the two kernels below are static schedules of ~10 barriers shared by three warp roles; a protocol slip would
// otherwise spin forever and take the GPU with it.  After ~2^22 failed try_waits (each suspends for the hardware time limit; a
// healthy wait completes in microseconds) the CTA reports which barrier starved and traps -> the launch fails with an error instead.
__device__ __noinline__ void attn_deadlock(int tag, uint32_t parity) {
  printf("attention_sm100: barrier wait timed out (tag %d parity %u block %d thread %d)\n", tag, parity, blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity, int tag) {
  const uint32_t addr = ptx::smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (spin > (1u << 22)) attn_deadlock(tag, parity);
  }
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// The 4-D map's outer dimensions (head, row, batch) are ordered by increasing byte stride on the host; `code` packs the position
// (1..3) of each: bits [0,2) head, [2,4) row, [4,6) batch.
__device__ __forceinline__ void load_rows(void* dst, const CUtensorMap* map, uint64_t* bar, int code, int h, int row, int b) {
  int c[4] = {0, 0, 0, 0};
  c[code & 3] = h; c[(code >> 2) & 3] = row; c[(code >> 4) & 3] = b;
  tma_load_4d(dst, map, bar, 0, c[1], c[2], c[3]);
}

__device__ __forceinline__ uint64_t desc_k(uint32_t addr) { return ptx::make_smem_desc_sw128(addr, 16, 1024); }            // contraction along the 64 columns
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo) { return ptx::make_smem_desc_sw128(addr, lbo, 1024); }  // contraction along rows

// one thread's 32 consecutive bf16 (64 B) of row r into a swizzled atom: 16-byte chunks c0..c0+3 of the row
__device__ __forceinline__ void store_row32(uint8_t* atom, int r, int chunk0, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(atom + r * 128 + (((chunk0 + j) ^ (r & 7)) << 4)) = pack8(v + 8 * j);
}

__device__ __forceinline__ void store_global64(bf16* dst, const float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) reinterpret_cast<uint4*>(dst)[j] = pack8(v + 8 * j);
}

__device__ __forceinline__ int ceil16(int x) { return (x + 15) & ~15; }

// the eight softmax warps only (TMA / MMA warps never join): named barrier 1
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void store_global32(bf16* dst, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(dst)[j] = pack8(v + 8 * j);
}

// ------------------------------------------------------------------------------------------------ forward
// TMEM: S at columns [0, SkP), O at [384, 448).  smem: Q | K | V | P atoms (ceil(Sk/64) x 16 KB) | barriers.
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Sq = p.Sq, Sk = p.Sk, SqP = ceil16(Sq), SkP = ceil16(Sk);
  const int nTq = (Sq + 127) >> 7, nKA = (Sk + 63) >> 6;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SqP * 128;
  uint8_t* sV = sK + SkP * 128;
  uint8_t* sP = sV + SkP * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + nKA * kAtomBytes);
  uint64_t *bar_qk = bars, *bar_v = bars + 1, *s_full = bars + 2, *s_free = bars + 3, *o_full = bars + 4, *o_free = bars + 5,
           *p_free = bars + 6, *p_ready = bars + 7;          // p_ready[0..4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* sStat = reinterpret_cast<float*>(bars + 16);      // [2 groups][128 rows] row max, then [2][128] row sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmQ); ptx::prefetch_tensormap(&tmK); ptx::prefetch_tensormap(&tmV);
    ptx::mbar_init(bar_qk, 1); ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(s_full, 1); ptx::mbar_init(s_free, kEpiWarps);
    ptx::mbar_init(o_full, 1); ptx::mbar_init(o_free, kEpiWarps); ptx::mbar_init(p_free, 1);
    for (int a = 0; a < 5; ++a) ptx::mbar_init(&p_ready[a], kEpiWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  PDL_GRID_SYNC();   // (PDL build) the set-up above overlapped the previous kernel's tail; global memory is touched below

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar_qk, (SqP + SkP) * 128);
      for (int r = 0; r < SqP; r += 16) load_rows(sQ + r * 128, &tmQ, bar_qk, p.pos[0], h, r, b);
      for (int r = 0; r < SkP; r += 16) load_rows(sK + r * 128, &tmK, bar_qk, p.pos[1], h, r, b);
      ptx::mbar_arrive_expect_tx(bar_v, SkP * 128);
      for (int r = 0; r < SkP; r += 16) load_rows(sV + r * 128, &tmV, bar_v, p.pos[2], h, r, b);
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV), aP = ptx::smem_u32(sP);
    const int n0 = SkP < 256 ? SkP : 256, n1 = SkP - n0;
    const uint32_t id_pv = ptx::make_idesc_bf16(128, 64, 0, 1);
    wait_bar(bar_qk, 0, 1);
    for (int i = 0; i < nTq; ++i) {
      if (i > 0) wait_bar(s_free, (i - 1) & 1, 2);
      ptx::tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tmem, desc_k(aQ + i * 128 * 128 + kk * 32), desc_k(aK + kk * 32), ptx::make_idesc_bf16(128, n0, 0, 0), kk > 0);
        if (n1 > 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ptx::umma_f16(tmem + 256, desc_k(aQ + i * 128 * 128 + kk * 32), desc_k(aK + 256 * 128 + kk * 32),
                          ptx::make_idesc_bf16(128, n1, 0, 0), kk > 0);
        }
        ptx::umma_commit(s_full);
      }
      __syncwarp();
      if (i == 0) wait_bar(bar_v, 0, 3);
      else wait_bar(o_free, (i - 1) & 1, 4);
      for (int a = 0; a < nKA; ++a) {
        wait_bar(&p_ready[a], i & 1, 5);
        ptx::tc_fence_after();
        if (lane == 0) {
          const int steps = ceil16(min(64, Sk - 64 * a)) >> 4;
          for (int s = 0; s < steps; ++s)
            ptx::umma_f16(tmem + 384, desc_k(aP + a * kAtomBytes + s * 32), desc_mn(aV + (64 * a + 16 * s) * 128, 8192), id_pv,
                          (a > 0 || s > 0) ? 1u : 0u);
        }
        __syncwarp();
      }
      if (lane == 0) { ptx::umma_commit(o_full); ptx::umma_commit(p_free); }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue warps (thread <-> query row)
    // Two groups of four warps share every tile: group g owns the 32-column chunks c = g, g+2, ... of S (= half g of every 64-key
    // P atom) and columns [32g, 32g+32) of O; row max / row sum are combined through shared memory.
    const int q = warp & 3, grp = (warp - 2) >> 2, r = q * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const int nch = (Sk + 31) >> 5;
    const long long bh = static_cast<long long>(b) * p.H + h;
    for (int i = 0; i < nTq; ++i) {
      const int g = i * 128 + r;
      const bool wvalid = i * 128 + q * 32 < Sq, rv = g < Sq;
      wait_bar(s_full, i & 1, 6);
      ptx::tc_fence_after();
      float mx = -INFINITY;
      if (wvalid) {
        for (int c = grp; c < nch; c += 2) {
          uint32_t raw[32];
          ptx::tmem_ld_32x32(trow + c * 32, raw);
          ptx::tmem_ld_wait();
          if (c * 32 + 32 <= Sk) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(raw[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (c * 32 + j < Sk) mx = fmaxf(mx, __uint_as_float(raw[j]));
          }
        }
      }
      sStat[grp * 128 + r] = mx;
      epi_sync();
      const float m2 = fmaxf(sStat[r], sStat[128 + r]) * p.sl2;     // scale > 0: max commutes with the scaling
      float l = 0.f;
      if (i > 0) wait_bar(p_free, (i - 1) & 1, 7);     // P V of the previous tile has finished reading the P atoms
      for (int a = 0; a < nKA; ++a) {
        const int c = 2 * a + grp;
        if (wvalid && c < nch) {
          uint32_t raw[32];
          ptx::tmem_ld_32x32(trow + c * 32, raw);
          ptx::tmem_ld_wait();
          float v[32];
          const bool full = c * 32 + 32 <= Sk;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e = ex2(fmaf(__uint_as_float(raw[j]), p.sl2, -m2));
            if (!full && c * 32 + j >= Sk) e = 0.f;
            v[j] = e; l += e;
          }
          store_row32(sP + a * kAtomBytes, r, grp * 4, v);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_ready[a]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_free);
      sStat[256 + grp * 128 + r] = l;
      epi_sync();
      l = sStat[256 + r] + sStat[384 + r];
      wait_bar(o_full, i & 1, 8);
      ptx::tc_fence_after();
      if (wvalid) {
        uint32_t raw[32];
        float v[32];
        const float inv = 1.0f / l;
        ptx::tmem_ld_32x32(trow + 384 + 32 * grp, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * inv;
        if (rv) {
          store_global32(p.o + b * p.o_bs + static_cast<long long>(g) * p.o_rs + h * 64 + 32 * grp, v);
          if (p.lse && grp == 0) p.lse[bh * Sq + g] = (m2 + __log2f(l)) * kLn2;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_free);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// TMEM columns: dQ_i at 64*i (i < 3), dK_j at 192, dV_j at 256, S / dP ring slots at 320 + 64*s (s < 3).
// smem: Q | K | V | dO (ceil16(L) rows x 128 B each) | P (2 atoms) | dS (2 atoms) | barriers.
struct BwdBars {
  uint64_t qk, vdo, slot_full[3], slot_free[3], p_ready, p_free, ds_ready, ds_free, dkv_full, dkv_free;
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Sq = p.Sq, Sk = p.Sk, SqP = ceil16(Sq), SkP = ceil16(Sk);
  const int nTq = (Sq + 127) >> 7, nTk = (Sk + 127) >> 7, T = nTq * nTk;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SqP * 128;
  uint8_t* sV = sK + SkP * 128;
  uint8_t* sdO = sV + SkP * 128;
  uint8_t* sP = sdO + SqP * 128;
  uint8_t* sdS = sP + 2 * kAtomBytes;
  BwdBars* bar = reinterpret_cast<BwdBars*>(sdS + 2 * kAtomBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmQ); ptx::prefetch_tensormap(&tmK); ptx::prefetch_tensormap(&tmV); ptx::prefetch_tensormap(&tmdO);
    ptx::mbar_init(&bar->qk, 1); ptx::mbar_init(&bar->vdo, 1);
    for (int s = 0; s < 3; ++s) { ptx::mbar_init(&bar->slot_full[s], 1); ptx::mbar_init(&bar->slot_free[s], 4); }
    ptx::mbar_init(&bar->p_ready, kEpiWarps); ptx::mbar_init(&bar->p_free, 1);
    ptx::mbar_init(&bar->ds_ready, kEpiWarps); ptx::mbar_init(&bar->ds_free, 1);
    ptx::mbar_init(&bar->dkv_full, 1); ptx::mbar_init(&bar->dkv_free, kEpiWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(&bar->tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bar->tmem_slot;
  PDL_GRID_SYNC();

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bar->qk, (SqP + SkP) * 128);
      for (int r = 0; r < SqP; r += 16) load_rows(sQ + r * 128, &tmQ, &bar->qk, p.pos[0], h, r, b);
      for (int r = 0; r < SkP; r += 16) load_rows(sK + r * 128, &tmK, &bar->qk, p.pos[1], h, r, b);
      ptx::mbar_arrive_expect_tx(&bar->vdo, (SqP + SkP) * 128);
      for (int r = 0; r < SkP; r += 16) load_rows(sV + r * 128, &tmV, &bar->vdo, p.pos[2], h, r, b);
      for (int r = 0; r < SqP; r += 16) load_rows(sdO + r * 128, &tmdO, &bar->vdo, p.pos[3], h, r, b);
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (static schedule, mirrored by the epilogue warps)
    const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV), adO = ptx::smem_u32(sdO),
                   aP = ptx::smem_u32(sP), adS = ptx::smem_u32(sdS);
    const uint32_t id_kk = ptx::make_idesc_bf16(128, 64, 0, 0), id_mm = ptx::make_idesc_bf16(128, 64, 1, 1),
                   id_km = ptx::make_idesc_bf16(128, 64, 0, 1);
    int k = 0;                                                   // ring item counter
    auto item = [&](uint32_t a_rows, uint32_t b_rows) {          // D[slot] = A[128 rows] . B[64 rows]^T over the 64 columns
      const int slot = k % 3, n = k / 3;
      if (n > 0) wait_bar(&bar->slot_free[slot], (n - 1) & 1, 9);
      ptx::tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tmem + 320 + 64 * slot, desc_k(a_rows + kk * 32), desc_k(b_rows + kk * 32), id_kk, kk > 0);
        ptx::umma_commit(&bar->slot_full[slot]);
      }
      __syncwarp();
      ++k;
    };
    auto stage_dk_dq = [&](int t) {                              // tile t's dS is in shared memory: dK_j += dS^T Q_i, dQ_i += dS K_j
      const int j = t / nTq, i = t % nTq;
      wait_bar(&bar->ds_ready, t & 1, 10);
      ptx::tc_fence_after();
      if (lane == 0) {
        const int qs = ceil16(min(128, Sq - 128 * i)) >> 4, ks = ceil16(min(128, Sk - 128 * j)) >> 4;
        for (int s = 0; s < qs; ++s)
          ptx::umma_f16(tmem + 192, desc_mn(adS + s * 2048, kAtomBytes), desc_mn(aQ + (128 * i + 16 * s) * 128, 8192), id_mm,
                        (i > 0 || s > 0) ? 1u : 0u);
        for (int s = 0; s < ks; ++s)
          ptx::umma_f16(tmem + 64 * i, desc_k(adS + (s >> 2) * kAtomBytes + (s & 3) * 32), desc_mn(aK + (128 * j + 16 * s) * 128, 8192),
                        id_km, (j > 0 || s > 0) ? 1u : 0u);
        ptx::umma_commit(&bar->ds_free);
        if (i == nTq - 1) ptx::umma_commit(&bar->dkv_full);
      }
      __syncwarp();
    };
    wait_bar(&bar->qk, 0, 11);
    wait_bar(&bar->vdo, 0, 12);
    for (int t = 0; t < T; ++t) {
      const int j = t / nTq, i = t % nTq;
      const uint32_t qi = aQ + i * 128 * 128, doi = adO + i * 128 * 128;
      item(qi, aK + (128 * j) * 128);                            // S^0
      item(doi, aV + (128 * j) * 128);                           // dP^0
      if (t > 0) stage_dk_dq(t - 1);
      if (128 * j + 64 < Sk) {
        item(qi, aK + (128 * j + 64) * 128);                     // S^1
        item(doi, aV + (128 * j + 64) * 128);                    // dP^1
      }
      wait_bar(&bar->p_ready, t & 1, 13);                      // dV_j += P^T dO_i
      if (i == 0 && j > 0) wait_bar(&bar->dkv_free, (j - 1) & 1, 14);
      ptx::tc_fence_after();
      if (lane == 0) {
        const int qs = ceil16(min(128, Sq - 128 * i)) >> 4;
        for (int s = 0; s < qs; ++s)
          ptx::umma_f16(tmem + 256, desc_mn(aP + s * 2048, kAtomBytes), desc_mn(adO + (128 * i + 16 * s) * 128, 8192), id_mm,
                        (i > 0 || s > 0) ? 1u : 0u);
        ptx::umma_commit(&bar->p_free);
      }
      __syncwarp();
    }
    stage_dk_dq(T - 1);
  } else {
    // ---------------------------------------------------------------- softmax-backward / epilogue warps (thread <-> tile row)
    // Two groups of four warps: group g consumes the ring items of key half g of every tile (S^g then dP^g, issued by the MMA warp in
    // the order S^0 dP^0 S^1 dP^1), writes atom g of P / dS, reads out dK (g = 0) or dV (g = 1) and columns [32g, 32g+32) of dQ.
    const int q = warp & 3, grp = (warp - 2) >> 2, r = q * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const long long bh = static_cast<long long>(b) * p.H + h;
    // per-row statistics of the (up to three) query tiles: log2-domain LSE and delta = rowsum(dO * O)
    float lse2[3] = {0.f, 0.f, 0.f}, dlt[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int g = i * 128 + r;
      if (g < Sq) {
        lse2[i] = p.lse[bh * Sq + g] * kLog2e;
        const uint4* po = reinterpret_cast<const uint4*>(p.o + b * p.o_bs + static_cast<long long>(g) * p.o_rs + h * 64);
        const uint4* pd = reinterpret_cast<const uint4*>(p.dout + b * p.do_bs + static_cast<long long>(g) * p.do_rs + h * 64);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float x[8], y[8];
          unpack8(po[c], x); unpack8(pd[c], y);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc = fmaf(x[e], y[e], acc);
        }
        dlt[i] = acc;
      }
    }
    float pr[64];                                                // probabilities of the current (tile, half), fp32
    auto readout32 = [&](uint32_t col, bf16* dst, bool ok) {
      uint32_t raw[32];
      ptx::tmem_ld_32x32(trow + col, raw);
      ptx::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(raw[c]);
      if (ok) store_global32(dst, v);
    };
    auto readout_dkv = [&](int j) {                              // group 0: dK_j, group 1: dV_j
      wait_bar(&bar->dkv_full, j & 1, 16);
      ptx::tc_fence_after();
      const int g = 128 * j + r;
      if (128 * j + 32 * q < Sk) {
        bf16* dst = grp == 0 ? p.dk + b * p.dk_bs + static_cast<long long>(g) * p.dk_rs + h * 64
                             : p.dv + b * p.dv_bs + static_cast<long long>(g) * p.dv_rs + h * 64;
        readout32(192 + 64 * grp, dst, g < Sk);
        readout32(192 + 64 * grp + 32, dst + 32, g < Sk);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->dkv_free);
    };
    int kbase = 0;                                               // ring index of this tile's first item
    for (int t = 0; t < T; ++t) {
      const int j = t / nTq, i = t % nTq;
      const int g = 128 * i + r;
      const bool wvalid = 128 * i + 32 * q < Sq, rv = g < Sq;
      const float l2 = i == 0 ? lse2[0] : (i == 1 ? lse2[1] : lse2[2]);
      const float dl = i == 0 ? dlt[0] : (i == 1 ? dlt[1] : dlt[2]);
      const int nh = (128 * j + 64 < Sk) ? 2 : 1;
      const bool mine = grp < nh;                                // this group's key half exists in this tile
      if (t > 0) wait_bar(&bar->p_free, (t - 1) & 1, 17);        // dV of the previous tile has read the P atoms
      if (mine) {
        // ---- S^g -> P^g
        const int k = kbase + 2 * grp, slot = k % 3;
        wait_bar(&bar->slot_full[slot], (k / 3) & 1, 15);
        ptx::tc_fence_after();
        const int k0 = 128 * j + 64 * grp;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (wvalid) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(trow + 320 + 64 * slot + 32 * half, raw);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float e = ex2(fmaf(__uint_as_float(raw[c]), p.sl2, -l2));
              pr[32 * half + c] = (rv && k0 + 32 * half + c < Sk) ? e : 0.f;
            }
            store_row32(sP + grp * kAtomBytes, r, 4 * half, pr + 32 * half);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar->slot_free[slot]);  // the accumulator is in registers: the slot may be refilled
        ptx::fence_proxy_async();
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->p_ready);
      if (t > 0) wait_bar(&bar->ds_free, (t - 1) & 1, 18);       // dK / dQ of the previous tile have read the dS atoms
      if (mine) {
        // ---- dP^g -> dS^g = P * (dP - delta) * scale
        const int k = kbase + 2 * grp + 1, slot = k % 3;
        wait_bar(&bar->slot_full[slot], (k / 3) & 1, 15);
        ptx::tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (wvalid) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(trow + 320 + 64 * slot + 32 * half, raw);
            ptx::tmem_ld_wait();
            float ds[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float pv = pr[32 * half + c];
              ds[c] = pv == 0.f ? 0.f : pv * (__uint_as_float(raw[c]) - dl) * p.scale;
            }
            store_row32(sdS + grp * kAtomBytes, r, 4 * half, ds);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar->slot_free[slot]);
        ptx::fence_proxy_async();
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->ds_ready);
      if (i == 0 && j > 0) readout_dkv(j - 1);                   // dK / dV of key tile j-1 are complete (the issuer committed dkv_full)
      kbase += 2 * nh;
    }
    readout_dkv(nTk - 1);                                        // its commit also covers every dQ product
    for (int i = 0; i < nTq; ++i) {
      const int g = 128 * i + r;
      if (128 * i + 32 * q < Sq)
        readout32(64 * i + 32 * grp, p.dq + b * p.dq_bs + static_cast<long long>(g) * p.dq_rs + h * 64 + 32 * grp, g < Sq);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

// 4-D bf16 map over  base + b*bs + row*rs + h*64 + c  (elements): innermost the 64 head channels, the three outer dimensions
// (head, row, batch) ordered by increasing stride; box = 64 channels x 16 rows of one (batch, head).  Returns the packed positions.
int make_map(CUtensorMap* map, const void* base, long long bs, long long rs, int B, int H, int L, int* code) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return PRISMER_ERR_DRIVER;
  struct Dim { unsigned long long size, stride; unsigned box; int what; };   // what: 0 head, 1 row, 2 batch
  Dim d[3] = {{static_cast<unsigned long long>(H), 128ull, 1u, 0},
              {static_cast<unsigned long long>(L), static_cast<unsigned long long>(rs) * 2, 16u, 1},
              {static_cast<unsigned long long>(B), static_cast<unsigned long long>(bs) * 2, 1u, 2}};
  for (int a = 0; a < 3; ++a)
    for (int c = a + 1; c < 3; ++c)
      if (d[c].stride < d[a].stride) { Dim t = d[a]; d[a] = d[c]; d[c] = t; }
  cuuint64_t dims[4] = {64, d[0].size, d[1].size, d[2].size};
  cuuint64_t strides[3] = {d[0].stride, d[1].stride, d[2].stride};
  cuuint32_t box[4] = {64, d[0].box, d[1].box, d[2].box};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int a = 0; a < 3; ++a)
    if (dims[a + 1] > 1 && (strides[a] % 16 || strides[a] == 0)) return PRISMER_ERR_ALIGN;
  for (int a = 0; a < 3; ++a) if (strides[a] == 0) strides[a] = 16;      // size-1 dimension: any legal stride
  int cd = 0;
  for (int a = 0; a < 3; ++a) cd |= (a + 1) << (2 * d[a].what);
  *code = cd;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PRISMER_OK : PRISMER_ERR_DRIVER;
}

int up16(int x) { return (x + 15) & ~15; }
int fwd_smem(int Sq, int Sk) { return (up16(Sq) + 2 * up16(Sk)) * 128 + ((Sk + 63) / 64) * kAtomBytes + 128 + 2048 + 1024; }
int bwd_smem(int Sq, int Sk) { return 2 * (up16(Sq) + up16(Sk)) * 128 + 4 * kAtomBytes + 256 + 1024; }

int g_force_legacy = 0;

bool supported(const PrismerAttnArgs* a) {
  return !g_force_legacy && a->kv_div <= 1 && a->d == 64 && !a->causal && !a->key_mask && a->drop_p == 0.f && a->Lq >= 64 && a->Lq <= kMaxL &&
         a->Lk >= 16 && a->Lk <= kMaxL;
}

void fill(TcParams& p, const PrismerAttnArgs* a) {
  p.B = a->B; p.H = a->H; p.Sq = a->Lq; p.Sk = a->Lk;
  p.scale = a->scale; p.sl2 = a->scale * kLog2e;
  p.lse = a->lse;
  p.o = reinterpret_cast<bf16*>(a->o); p.o_bs = a->o_bs; p.o_rs = a->o_rs;
  p.dout = reinterpret_cast<const bf16*>(a->dout); p.do_bs = a->do_bs; p.do_rs = a->do_rs;
  p.dq = reinterpret_cast<bf16*>(a->dq); p.dq_bs = a->dq_bs; p.dq_rs = a->dq_rs;
  p.dk = reinterpret_cast<bf16*>(a->dk); p.dk_bs = a->dk_bs; p.dk_rs = a->dk_rs;
  p.dv = reinterpret_cast<bf16*>(a->dv); p.dv_bs = a->dv_bs; p.dv_rs = a->dv_rs;
}

bool aligned16(const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; }

}  // namespace

// 1 = handled (rc in *rc_out), 0 = shape not covered by the tcgen05 kernels (caller falls through to the mma.sync kernels)
int attn_sm100_try_fwd(const PrismerAttnArgs* a, cudaStream_t stream, int* rc_out) {
  if (!supported(a) || a->scale <= 0.f) return 0;
  if (!aligned16(a->q) || !aligned16(a->k) || !aligned16(a->v) || !aligned16(a->o)) return 0;
  TcParams p;
  fill(p, a);
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_map(&tq, a->q, a->q_bs, a->q_rs, a->B, a->H, a->Lq, &p.pos[0])) ||
      (rc = make_map(&tk, a->k, a->k_bs, a->k_rs, a->B, a->H, a->Lk, &p.pos[1])) ||
      (rc = make_map(&tv, a->v, a->v_bs, a->v_rs, a->B, a->H, a->Lk, &p.pos[2]))) {
    *rc_out = rc;
    return 1;
  }
  p.pos[3] = 0;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(kMaxL, kMaxL)) != cudaSuccess) {
      *rc_out = PRISMER_ERR_CUDA;
      return 1;
    }
    configured = true;
  }
  pdl_launch(attn_fwd_tc_kernel, dim3(a->B * a->H), dim3(kThreads), static_cast<size_t>(fwd_smem(a->Lq, a->Lk)), stream, tq, tk, tv, p);
  *rc_out = LAUNCH_CHECK();
  return 1;
}

int attn_sm100_try_bwd(const PrismerAttnArgs* a, cudaStream_t stream, int* rc_out) {
  if (!supported(a) || a->scale <= 0.f) return 0;
  if (!aligned16(a->q) || !aligned16(a->k) || !aligned16(a->v) || !aligned16(a->o) || !aligned16(a->dout) || !aligned16(a->dq) ||
      !aligned16(a->dk) || !aligned16(a->dv))
    return 0;
  TcParams p;
  fill(p, a);
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_map(&tq, a->q, a->q_bs, a->q_rs, a->B, a->H, a->Lq, &p.pos[0])) ||
      (rc = make_map(&tk, a->k, a->k_bs, a->k_rs, a->B, a->H, a->Lk, &p.pos[1])) ||
      (rc = make_map(&tv, a->v, a->v_bs, a->v_rs, a->B, a->H, a->Lk, &p.pos[2])) ||
      (rc = make_map(&tdo, a->dout, a->do_bs, a->do_rs, a->B, a->H, a->Lq, &p.pos[3]))) {
    *rc_out = rc;
    return 1;
  }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(kMaxL, kMaxL)) != cudaSuccess) {
      *rc_out = PRISMER_ERR_CUDA;
      return 1;
    }
    configured = true;
  }
  pdl_launch(attn_bwd_tc_kernel, dim3(a->B * a->H), dim3(kThreads), static_cast<size_t>(bwd_smem(a->Lq, a->Lk)), stream, tq, tk, tv, tdo, p);
  *rc_out = LAUNCH_CHECK();
  return 1;
}

extern "C" int prismer_set_attention_path(int mode) {
  if (mode != 0 && mode != 1) return PRISMER_ERR_SHAPE;
  g_force_legacy = mode;
  return PRISMER_OK;
}
