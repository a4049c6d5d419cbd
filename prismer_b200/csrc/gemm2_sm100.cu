// EXPERIMENTAL (round-2 candidate; compiled and exported as prismer_gemm_bf16_2cta, NOT on the default path, not yet run on
// hardware): the persistent warp-specialised GEMM of gemm_sm100.cu on CTA PAIRS -- tcgen05.mma.cta_group::2.
//
//   cluster (2,1,1) = two SMs of one TPC work on one 256 x BN output tile:
//     * each CTA TMA-loads ITS 128 rows of A and ITS BN/2 rows of B per k-block (cp.async.bulk.tensor ... cta_group::2, completing on
//       the leader's mbarrier), so the L2->SM operand traffic per output element -- the measured limiter of the single-CTA
//       kernel (profiles/ncu_r1_summary.md: lts 35 %, tensor pipe 56 %) -- drops by a third at BN = 256 (32 KB instead of 48 KB
//       per 128 x 256 x 64 MMA block) and the smem ring gets 6 stages instead of 4;
//     * the leader CTA's warp 1 issues ONE tcgen05.mma (M = 256) per UMMA_K that reads both CTAs' shared memory and writes both
//       CTAs' TMEM; tcgen05.commit ... multicast::cluster releases the smem stage / publishes the accumulator in both CTAs;
//     * both CTAs run the unchanged fused epilogue on their own 128 x BN half; the peer's epilogue warps arrive remotely on the
//       leader's "accumulator drained" barrier.
// The epilogue is a verbatim copy of gemm_sm100.cu's (kept separate so that the hardware-validated translation unit stays
// byte-identical until this one has been run; to be merged into one template afterwards).  No split-K here.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "prismer_sm100.h"

#include <mutex>

namespace {

constexpr int BM = 128;          // rows per CTA; UMMA_M = 256 over the CTA pair
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int kNumEpiWarps = 8;     // two warps per TMEM lane quarter, each takes every other 32-column chunk
constexpr int kThreads = 32 * (2 + kNumEpiWarps);

template <int BN> struct Cfg {
  static constexpr int kABytes = BM * BK * 2;           // this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * BK * 2;     // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 6 : 8;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (power of two >= 32 for BN in {64,128,256})
  static constexpr int kStagingBytes = kNumEpiWarps * 4096;   // per-epilogue-warp 32x32 staging tile (coalesced stores)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kStagingBytes;
};

// ---------------------------------------------------------------- cta_group::2 PTX (local to this experimental file)
namespace ptx2 {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {      // executed by the same warp id in BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols) : "memory");
}
// shared::cluster address of the object at `p`'s offset in the pair's leader CTA (rank 0)
__device__ __forceinline__ uint32_t leader_addr(const void* p) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(ptx::smem_u32(p)), "r"(0u));
  return r;
}
// 2-D tiled load into THIS CTA's smem, completing transaction bytes on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread retired) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_both(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   ptx::smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// arrive on the barrier at this smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(ptx::smem_u32(bar)),
      "r"(cta)
      : "memory");
}
}  // namespace ptx2

struct EpiParams {
  void* C; long long ldc;
  const float* bias;
  const bf16* residual; long long ldr;
  bf16* aux_out; const bf16* aux_in; long long ldaux;
  int act;          // activation applied to (acc + bias)
  int act_grad;     // != 0: out = acc * act'(aux_in)   (dgrad through an activation)
  int out_fp32;     // C is fp32 (else bf16)
  int accumulate;   // fp32 out only: C += result
  float alpha;      // result scale (applied to the accumulator first)
  float drop_p; uint32_t drop_thr16; float drop_scale; const unsigned long long* seed; uint32_t rng_stream;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                  EpiParams ep) {
  constexpr int splits = 1;
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full = empty_bar + C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* staging = smem + C::kStages * C::kStageBytes + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx2::cluster_ctarank();           // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_m = (M + 2 * BM - 1) / (2 * BM), num_n = (N + BN - 1) / BN;   // 256-row tiles over the pair
  const int num_tiles = num_m * num_n;
  const int num_k = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    // the leader's "accumulator drained" barrier collects the epilogue warps of BOTH CTAs
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 2 * kNumEpiWarps); }
    ptx::fence_barrier_init();
  }
  __syncwarp();                                           // warp 0 reconverges: the cluster barrier instructions are .aligned
  ptx2::cluster_sync();                                   // both CTAs' barriers exist before any remote arrive / TMA completion
  if (warp == 1) ptx2::tmem_alloc<C::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / num_n) * (2 * BM) + static_cast<int>(rank) * BM;          // this CTA's 128 rows of A
        const int n0 = (tile % num_n) * BN + static_cast<int>(rank) * (BN / 2);          // this CTA's half of the B tile
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);        // own copy: released by the leader's multicast commit
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          const uint32_t lfull = ptx2::leader_addr(&full_bar[stage]);
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);   // both CTAs' bytes land on the leader
          const int k0 = kb * BK;
          if constexpr (!A_MN) {
            ptx2::tma_load_2d(sa, &tmA, lfull, k0, m0);                       // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)                                  // box {64 m, 64 k} per 64-wide MN atom
              ptx2::tma_load_2d(sa + j * (BK * 128), &tmA, lfull, m0 + 64 * j, k0);
          }
          if constexpr (!B_MN) {
            ptx2::tma_load_2d(sb, &tmB, lfull, k0, n0);                       // box {64 k, BN/2 n}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j)
              ptx2::tma_load_2d(sb + j * (BK * 128), &tmB, lfull, n0 + 64 * j, k0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int kb0 = 0, kb1 = num_k;
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // the epilogues of BOTH CTAs have drained this accumulator stage
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            // K-major: 8-row groups 1024 B apart (SBO), advance 32 B per UMMA_K inside the 128 B swizzle span.
            // MN-major: 64-element MN atoms BK*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO), advance 16 rows.
            const uint64_t da = A_MN ? ptx::make_smem_desc_sw128(sa + kk * (UMMA_K * 128), BK * 128, 1024)
                                     : ptx::make_smem_desc_sw128(sa + kk * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_MN ? ptx::make_smem_desc_sw128(sb + kk * (UMMA_K * 128), BK * 128, 1024)
                                     : ptx::make_smem_desc_sw128(sb + kk * (UMMA_K * 2), 16, 1024);
            ptx2::umma_f16(tmem_d, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          ptx2::umma_commit_both(&empty_bar[stage]);               // smem slot free in both CTAs once these MMAs retire
          if (kb == kb1 - 1) ptx2::umma_commit_both(&tmem_full[acc]);  // accumulator halves ready for both epilogues
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ epilogue warps (both CTAs, each on its 128 x BN half)
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;            // which half of the 32-column chunks this warp handles
    int acc = 0; uint32_t acc_phase = 0;
    const bool has_drop = ep.drop_p > 0.f;
    unsigned long long seed = 0;
    if (has_drop) seed = *ep.seed;
    const Philox philox(seed);
    uint8_t* stg = staging + (warp - 2) * 4096;
    // Coalesced stores of a 32x32 chunk: every lane parks its row in a swizzled (bank-conflict-free) smem tile, then the warp
    // writes it back with each instruction covering whole rows segments (bf16: 8 rows x 64 B, fp32: 4 rows x 128 B) -- full
    // 32-byte sectors instead of 32 row-strided 16-byte pieces per instruction.
    auto store_bf16_staged = [&](bf16* blk, long long ld, const float* v, int rows_valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = pack8(v + 8 * j);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = (lane >> 2) + 8 * i, ch = lane & 3;
        const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
        if (rr < rows_valid) *reinterpret_cast<uint4*>(blk + rr * ld + ch * 8) = val;
      }
      __syncwarp();
    };
    auto store_f32_staged = [&](float* blk, long long ld, const float* v, int rows_valid, int mode /*0 store, 1 accumulate, 2 atomic*/) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = (lane >> 3) + 4 * i, ch = lane & 7;
        float4 val = *reinterpret_cast<const float4*>(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
        if (rr < rows_valid) {
          float4* dst = reinterpret_cast<float4*>(blk + rr * ld + ch * 4);
          if (mode == 2) {
            atomicAdd(dst, val);
          } else {
            if (mode == 1) { const float4 p = *dst; val.x += p.x; val.y += p.y; val.z += p.z; val.w += p.w; }
            *dst = val;
          }
        }
      }
      __syncwarp();
    };
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile / num_n) * (2 * BM) + static_cast<int>(rank) * BM, n0 = (tile % num_n) * BN;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      const int rows_valid = min(32, M - (m0 + q * 32));   // rows of this warp's 32-row slab that exist (may be <= 0)
#pragma unroll 1
      for (int c = chalf * 32; c < BN; c += 64) {
        const int col0 = n0 + c;
        const bool full = (col0 + 32 <= N);                  // warp-uniform
        // the bf16 side input of this chunk (saved pre-activation for act_grad, else the residual) is requested BEFORE the TMEM
        // load is waited for, so its global latency overlaps the accumulator read
        const bf16* side = ep.act_grad ? ep.aux_in : ep.residual;
        const bool side_pre = side != nullptr && full && row_ok;
        uint4 pre[4];
        if (side_pre) {
          const uint4* sp = reinterpret_cast<const uint4*>(side + static_cast<long long>(row) * (ep.act_grad ? ep.ldaux : ep.ldr) + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) pre[j] = sp[j];
        }
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c, raw);
        ptx::tmem_ld_wait();
        if (col0 >= N || rows_valid <= 0) continue;          // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * ep.alpha;
        if (row_ok) {
          // ---- bias
          if (ep.bias) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            } else {
              _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < N) { v[j] += __ldg(ep.bias + col0 + j); }
            }
          }
          // ---- aux: save pre-activation / multiply by the activation derivative
          if (ep.aux_out && !full) {
            bf16* ap = ep.aux_out + static_cast<long long>(row) * ep.ldaux + col0;
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < N) { ap[j] = __float2bfloat16(v[j]); }
          }
        }
        if (ep.aux_out && full)      // pre-activation saved for the backward (all lanes participate in the staged store)
          store_bf16_staged(ep.aux_out + static_cast<long long>(m0 + q * 32) * ep.ldaux + col0, ep.ldaux, v, rows_valid);
        if (row_ok) {
          if (ep.act_grad) {
            const bf16* ap = ep.aux_in + static_cast<long long>(row) * ep.ldaux + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float z[8]; unpack8(pre[j], z);
                act_bwd8(ep.act_grad, v + 8 * j, z);
              }
            } else {
              _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < N) { v[j] *= act_bwd(ep.act_grad, __bfloat162float(ap[j])); }
            }
          } else if (ep.act) {
            act_fwd32(ep.act, v);
          }
          // ---- dropout on the branch output (before the residual add): roberta.py:138,181
          if (has_drop) {
            const unsigned long long e0 = static_cast<unsigned long long>(row) * N + col0;  // N % 8 == 0 enforced on host
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t keep = dropout_keep8(philox, (e0 + j) >> 3, ep.rng_stream, ep.drop_thr16);
#pragma unroll
              for (int t = 0; t < 8; ++t) v[j + t] = ((keep >> t) & 1u) ? v[j + t] * ep.drop_scale : 0.f;
            }
          }
          // ---- residual
          if (ep.residual) {
            const bf16* rp = ep.residual + static_cast<long long>(row) * ep.ldr + col0;
            if (full && !ep.act_grad) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float r[8]; unpack8(pre[j], r);
#pragma unroll
                for (int t = 0; t < 8; ++t) v[8 * j + t] += r[t];
              }
            } else if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float r[8]; unpack8(*reinterpret_cast<const bf16x8*>(rp + j), r);
#pragma unroll
                for (int t = 0; t < 8; ++t) v[j + t] += r[t];
              }
            } else {
              _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < N) { v[j] += __bfloat162float(rp[j]); }
            }
          }
        }
        // ---- store
        if (full) {
          if (ep.out_fp32)
            store_f32_staged(reinterpret_cast<float*>(ep.C) + static_cast<long long>(m0 + q * 32) * ep.ldc + col0, ep.ldc, v, rows_valid,
                             splits > 1 ? 2 : (ep.accumulate ? 1 : 0));
          else
            store_bf16_staged(reinterpret_cast<bf16*>(ep.C) + static_cast<long long>(m0 + q * 32) * ep.ldc + col0, ep.ldc, v, rows_valid);
        } else if (row_ok) {                                  // N tail: per-element path
          if (ep.out_fp32) {
            float* cp = reinterpret_cast<float*>(ep.C) + static_cast<long long>(row) * ep.ldc + col0;
            if (splits > 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < N) atomicAdd(cp + j, v[j]);
            } else {
              _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < N) { cp[j] = ep.accumulate ? cp[j] + v[j] : v[j]; }
            }
          } else {
            bf16* cp = reinterpret_cast<bf16*>(ep.C) + static_cast<long long>(row) * ep.ldc + col0;
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < N) { cp[j] = __float2bfloat16(v[j]); }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx2::mbar_arrive_cluster(&tmem_empty[acc], 0);       // always on the leader's barrier
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  // neither CTA may leave (or free TMEM) while the other can still read its shared memory or signal its barriers
  ptx::tc_fence_before();
  __syncwarp();                                           // producer / MMA lanes rejoin their warps (.aligned barrier below)
  ptx2::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx2::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode2() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// 2-D bf16 map over a row-major [rows, cols] matrix with row stride `ld` elements; box = {box_cols, box_rows}.
int make_map_2d2(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                int box_rows) {
  PFN_encodeTiled enc = get_encode2();
  if (!enc) return PRISMER_ERR_DRIVER;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PRISMER_OK : PRISMER_ERR_DRIVER;
}

template <int BN, bool A_MN, bool B_MN>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, int max_pairs, cudaStream_t stream) {
  auto kern = gemm2_bf16_kernel<BN, A_MN, B_MN>;
  static int max_clusters = 0;  // per-instantiation; benign race (idempotent)
  if (max_clusters == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes) != cudaSuccess) return PRISMER_ERR_CUDA;
    // how many CTA pairs can be co-resident (GPCs with an odd number of SMs leave one unpaired)
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(2 * 74); q.blockDim = dim3(kThreads); q.dynamicSmemBytes = Cfg<BN>::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    q.attrs = at; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 64; }
    max_clusters = n;
  }
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  int pairs = tiles < max_clusters ? tiles : max_clusters;
  if (max_pairs > 0 && pairs > max_pairs) pairs = max_pairs;
  kern<<<dim3(2 * pairs), dim3(kThreads), Cfg<BN>::kSmemBytes, stream>>>(ta, tb, M, N, K, ep);   // __cluster_dims__(2,1,1)
  return LAUNCH_CHECK();
}

}  // namespace

// Same argument block as prismer_gemm_bf16 (include/prismer_sm100.h); split-K requests (force_splits > 1) are refused, the N tile is
// 256 unless force_bn == 128.  max_ctas counts CTA PAIRS here.
extern "C" int prismer_gemm_bf16_2cta(const PrismerGemmArgs* a, cudaStream_t stream) {
  if (!a || a->M <= 0 || a->N <= 0 || a->K <= 0) return PRISMER_ERR_SHAPE;
  if ((a->lda % 8) || (a->ldb % 8)) return PRISMER_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15) ||
      (reinterpret_cast<uintptr_t>(a->C) & 15))
    return PRISMER_ERR_ALIGN;
  if (a->drop_p > 0.f && ((a->N % 8) || !a->seed)) return PRISMER_ERR_SHAPE;
  if (a->accumulate && !a->out_fp32) return PRISMER_ERR_SHAPE;
  if (a->force_splits > 1) return PRISMER_ERR_SHAPE;
  const int celt = a->out_fp32 ? 4 : 8;
  if (a->ldc % celt) return PRISMER_ERR_ALIGN;
  if (a->residual && (a->ldr % 8)) return PRISMER_ERR_ALIGN;
  if ((a->aux_out || a->aux_in) && (a->ldaux % 8)) return PRISMER_ERR_ALIGN;
  const int bn = a->force_bn == 128 ? 128 : 256;

  CUtensorMap ta, tb;
  int rc;
  if (!a->transA) rc = make_map_2d2(&ta, a->A, a->M, a->K, a->lda, BK, BM);       // A[M,K]: box {64 k, 128 m} per CTA
  else rc = make_map_2d2(&ta, a->A, a->K, a->M, a->lda, 64, BK);                   // A^T stored as [K,M]
  if (rc) return rc;
  if (!a->transB) rc = make_map_2d2(&tb, a->B, a->N, a->K, a->ldb, BK, bn / 2);   // B[N,K]: box {64 k, BN/2 n} per CTA
  else rc = make_map_2d2(&tb, a->B, a->K, a->N, a->ldb, 64, BK);                   // B^T stored as [K,N]
  if (rc) return rc;

  EpiParams ep;
  ep.C = a->C; ep.ldc = a->ldc;
  ep.bias = a->bias;
  ep.residual = reinterpret_cast<const bf16*>(a->residual); ep.ldr = a->ldr;
  ep.aux_out = reinterpret_cast<bf16*>(a->aux_out);
  ep.aux_in = reinterpret_cast<const bf16*>(a->aux_in);
  ep.ldaux = a->ldaux;
  ep.act = a->act; ep.act_grad = a->act_grad;
  ep.out_fp32 = a->out_fp32; ep.accumulate = a->accumulate;
  ep.alpha = a->alpha;
  ep.drop_p = a->drop_p;
  ep.drop_thr16 = static_cast<uint32_t>(a->drop_p * 65536.0f + 0.5f);
  ep.drop_scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  ep.seed = a->seed; ep.rng_stream = a->rng_stream;
  if (ep.act_grad && !ep.aux_in) return PRISMER_ERR_SHAPE;

#define DISPATCH2(BN_)                                                                                          \
  if (!a->transA && !a->transB) return launch2<BN_, false, false>(ta, tb, a->M, a->N, a->K, ep, a->max_ctas, stream); \
  if (!a->transA && a->transB) return launch2<BN_, false, true>(ta, tb, a->M, a->N, a->K, ep, a->max_ctas, stream);   \
  if (a->transA && !a->transB) return launch2<BN_, true, false>(ta, tb, a->M, a->N, a->K, ep, a->max_ctas, stream);   \
  return launch2<BN_, true, true>(ta, tb, a->M, a->N, a->K, ep, a->max_ctas, stream);
  if (bn == 256) { DISPATCH2(256) }
  DISPATCH2(128)
#undef DISPATCH2
}
