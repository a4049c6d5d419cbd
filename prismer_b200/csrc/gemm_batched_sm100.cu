// EXPERIMENTAL (round-2 candidate, compiled but NOT on the default path and not yet validated on hardware):
// batched variant of the persistent tcgen05 GEMM of gemm_sm100.cu for attention expressed as matrix products over
// (batch, head) problems -- S = scale*Q.K^T (or P = exp(scale*Q.K^T - lse) straight from the epilogue), O = P.V, dV = P^T.dO,
// dS = P*(dO.V^T - delta)*scale, dQ = dS.K, dK = dS^T.Q -- with the
// score / probability matrices of the ViT shape (384 x 260 x 260 bf16 = 52 MB per layer) staying L2-resident.
//
//   C_i[M,N] = epilogue(alpha * op(A_i) . op(B_i)^T),  i = (bo, bi) in [0, batch_outer) x [0, batch_inner)
//   operand X_i = X + bo * X_bs_outer + bi * X_bs_inner  (elements);  rows of every problem are bounded by ITS OWN M / N / K
//   (4-D TMA tensor maps {inner, rows, batch_inner, batch_outer}: out-of-range rows of a problem are zero-filled and never
//   bleed into the neighbouring problem).
//
// Same pipeline as the validated kernel: warp 0 TMA producer, warp 1 tcgen05.mma issuer (2 TMEM accumulator stages),
// 8 epilogue warps with swizzled smem-staged coalesced stores.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "prismer_sm100.h"

#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int kNumEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kNumEpiWarps);

template <int BN> struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kStagingBytes = kNumEpiWarps * 4096;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kStagingBytes;
};

struct BatchedEpi {
  bf16* C; long long ldc, c_bo, c_bi;
  const bf16* aux; long long ldaux, aux_bo, aux_bi;   // mode 1: probabilities P (same logical layout as C)
  const float* rowvec; long long rowvec_bs;            // mode 1: delta, fp32 [batch, M]
  int mode;                                            // 0: C = alpha*acc     1: C = aux * (acc - rowvec[row]) * alpha
                                                       // 2: C = exp(alpha*acc - rowvec[row])  (probabilities from saved LSE)
  float alpha;
  int batch_inner;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_batched_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                         int batch, BatchedEpi ep) {
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
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int tiles_per_problem = num_m * num_n;
  const int num_tiles = tiles_per_problem * batch;
  const int num_k = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], kNumEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int prob = tile / tiles_per_problem, t = tile % tiles_per_problem;
        const int bo = prob / ep.batch_inner, bi = prob % ep.batch_inner;
        const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
          const int k0 = kb * BK;
          if constexpr (!A_MN) {
            tma_load_4d(sa, &tmA, &full_bar[stage], k0, m0, bi, bo);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_4d(sa + j * (BK * 128), &tmA, &full_bar[stage], m0 + 64 * j, k0, bi, bo);
          }
          if constexpr (!B_MN) {
            tma_load_4d(sb, &tmB, &full_bar[stage], k0, n0, bi, bo);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_4d(sb + j * (BK * 128), &tmB, &full_bar[stage], n0 + 64 * j, k0, bi, bo);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint64_t da = A_MN ? ptx::make_smem_desc_sw128(sa + kk * (UMMA_K * 128), BK * 128, 1024)
                                     : ptx::make_smem_desc_sw128(sa + kk * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_MN ? ptx::make_smem_desc_sw128(sb + kk * (UMMA_K * 128), BK * 128, 1024)
                                     : ptx::make_smem_desc_sw128(sb + kk * (UMMA_K * 2), 16, 1024);
            ptx::umma_f16(tmem_d, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) ptx::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    uint8_t* stg = staging + (warp - 2) * 4096;
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int prob = tile / tiles_per_problem, t = tile % tiles_per_problem;
      const int bo = prob / ep.batch_inner, bi = prob % ep.batch_inner;
      const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
      bf16* Cp = ep.C + bo * ep.c_bo + bi * ep.c_bi;
      const bf16* Xp = ep.aux ? ep.aux + bo * ep.aux_bo + bi * ep.aux_bi : nullptr;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      const int rows_valid = min(32, M - (m0 + q * 32));
      float rv = 0.f;
      if (ep.mode != 0 && row_ok) rv = ep.rowvec[static_cast<long long>(prob) * ep.rowvec_bs + row];
#pragma unroll 1
      for (int c = chalf * 32; c < BN; c += 64) {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c, raw);
        ptx::tmem_ld_wait();
        const int col0 = n0 + c;
        if (col0 >= N || rows_valid <= 0) continue;
        const bool full = (col0 + 32 <= N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (row_ok) {
          if (ep.mode == 1) {
            const bf16* ap = Xp + static_cast<long long>(row) * ep.ldaux + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float pz[8]; unpack8(*reinterpret_cast<const bf16x8*>(ap + j), pz);
#pragma unroll
                for (int t2 = 0; t2 < 8; ++t2) v[j + t2] = pz[t2] * (v[j + t2] - rv) * ep.alpha;
              }
            } else {
              _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < N) { v[j] = __bfloat162float(ap[j]) * (v[j] - rv) * ep.alpha; }
            }
          } else if (ep.mode == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __expf(v[j] * ep.alpha - rv);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= ep.alpha;
          }
        }
        if (full) {
          store_bf16_staged(Cp + static_cast<long long>(m0 + q * 32) * ep.ldc + col0, ep.ldc, v, rows_valid);
        } else if (row_ok) {
          bf16* cp = Cp + static_cast<long long>(row) * ep.ldc + col0;
          _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < N) { cp[j] = __float2bfloat16(v[j]); }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_b() {
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

// 4-D bf16 map {inner (contiguous), rows, batch_inner, batch_outer}; box = {box_inner, box_rows, 1, 1}
int make_map_4d(CUtensorMap* map, const void* ptr, long long inner, long long rows, long long ld, int bi, long long bs_inner,
                int bo, long long bs_outer, int box_inner, int box_rows) {
  PFN_encodeTiled enc = get_encode_b();
  if (!enc) return PRISMER_ERR_DRIVER;
  if ((ld % 8) || (bs_inner % 8) || (bs_outer % 8)) return PRISMER_ERR_ALIGN;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(bi),
                        static_cast<cuuint64_t>(bo)};
  // a batch dimension of extent 1 still needs a non-zero stride
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(bs_inner > 0 ? bs_inner : ld) * 2,
                           static_cast<cuuint64_t>(bs_outer > 0 ? bs_outer : ld) * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PRISMER_OK : PRISMER_ERR_DRIVER;
}

template <int BN, bool A_MN, bool B_MN>
int launch_b(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, int batch, const BatchedEpi& ep, cudaStream_t stream) {
  auto kern = gemm_bf16_batched_kernel<BN, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes) != cudaSuccess) return PRISMER_ERR_CUDA;
    configured = true;
  }
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const long long tiles = static_cast<long long>((M + BM - 1) / BM) * ((N + BN - 1) / BN) * batch;
  const int grid = static_cast<int>(tiles < sms ? tiles : sms);
  kern<<<grid, kThreads, Cfg<BN>::kSmemBytes, stream>>>(ta, tb, M, N, K, batch, ep);
  return LAUNCH_CHECK();
}

}  // namespace

extern "C" int prismer_gemm_bf16_batched(const PrismerBatchedGemmArgs* a, cudaStream_t stream) {
  if (!a || a->M <= 0 || a->N <= 0 || a->K <= 0 || a->batch_outer <= 0 || a->batch_inner <= 0) return PRISMER_ERR_SHAPE;
  if ((a->lda % 8) || (a->ldb % 8) || (a->ldc % 8)) return PRISMER_ERR_ALIGN;
  if (a->mode == 1 && (!a->aux || !a->rowvec || (a->ldaux % 8))) return PRISMER_ERR_SHAPE;
  if (a->mode == 2 && !a->rowvec) return PRISMER_ERR_SHAPE;
  if (a->mode < 0 || a->mode > 2) return PRISMER_ERR_SHAPE;
  // N tile: least padded columns (attention scores, N = 260: 5 x 64 = 320 instead of 2 x 256 = 512); ties -> the larger tile
  int bn = 256;
  {
    long long best = -1;
    for (int cand : {256, 128, 64}) {
      const long long padded = static_cast<long long>((a->N + cand - 1) / cand) * cand;
      if (best < 0 || padded < best) { best = padded; bn = cand; }
    }
  }
  if (a->force_bn == 64 || a->force_bn == 128 || a->force_bn == 256) bn = a->force_bn;
  CUtensorMap ta, tb;
  int rc;
  if (!a->transA) rc = make_map_4d(&ta, a->A, a->K, a->M, a->lda, a->batch_inner, a->a_bs_inner, a->batch_outer, a->a_bs_outer, BK, BM);
  else rc = make_map_4d(&ta, a->A, a->M, a->K, a->lda, a->batch_inner, a->a_bs_inner, a->batch_outer, a->a_bs_outer, 64, BK);
  if (rc) return rc;
  if (!a->transB) rc = make_map_4d(&tb, a->B, a->K, a->N, a->ldb, a->batch_inner, a->b_bs_inner, a->batch_outer, a->b_bs_outer, BK, bn);
  else rc = make_map_4d(&tb, a->B, a->N, a->K, a->ldb, a->batch_inner, a->b_bs_inner, a->batch_outer, a->b_bs_outer, 64, BK);
  if (rc) return rc;
  BatchedEpi ep;
  ep.C = reinterpret_cast<bf16*>(a->C); ep.ldc = a->ldc; ep.c_bo = a->c_bs_outer; ep.c_bi = a->c_bs_inner;
  ep.aux = reinterpret_cast<const bf16*>(a->aux); ep.ldaux = a->ldaux; ep.aux_bo = a->aux_bs_outer; ep.aux_bi = a->aux_bs_inner;
  ep.rowvec = a->rowvec; ep.rowvec_bs = a->rowvec_bs;
  ep.mode = a->mode; ep.alpha = a->alpha; ep.batch_inner = a->batch_inner;
  const int batch = a->batch_outer * a->batch_inner;
#define DISPATCH_B(BN_)                                                                                              \
  if (!a->transA && !a->transB) return launch_b<BN_, false, false>(ta, tb, a->M, a->N, a->K, batch, ep, stream);     \
  if (!a->transA && a->transB) return launch_b<BN_, false, true>(ta, tb, a->M, a->N, a->K, batch, ep, stream);       \
  if (a->transA && !a->transB) return launch_b<BN_, true, false>(ta, tb, a->M, a->N, a->K, batch, ep, stream);       \
  return launch_b<BN_, true, true>(ta, tb, a->M, a->N, a->K, batch, ep, stream);
  if (bn == 256) { DISPATCH_B(256) }
  if (bn == 128) { DISPATCH_B(128) }
  DISPATCH_B(64)
#undef DISPATCH_B
}
