// EXPERIMENTAL companions of gemm_batched_sm100.cu (round-2 candidate, not on the default path, not yet validated on hardware):
// row softmax over L2-resident score matrices and the delta = rowsum(dO * O) reduction of the attention backward.
#include "common.cuh"
#include "prismer_sm100.h"

namespace {

// in place: P[r, :Lk] = softmax(S[r, :Lk]) (fp32 math), P[r, Lk:ld] = 0.   One warp per row.
__global__ void __launch_bounds__(256) softmax_rows_kernel(bf16* __restrict__ s, long long rows, int Lk, int ld) {
  const int lane = threadIdx.x & 31;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows;
       r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    bf16* p = s + r * ld;
    float mx = -INFINITY;
    for (int c = lane; c < Lk; c += 32) mx = fmaxf(mx, __bfloat162float(p[c]));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < Lk; c += 32) sum += __expf(__bfloat162float(p[c]) - mx);
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int c = lane; c < ld; c += 32) p[c] = __float2bfloat16(c < Lk ? __expf(__bfloat162float(p[c]) - mx) * inv : 0.f);
  }
}

// delta[(b*H + h) * Lq + q] = sum_d dO[q,b,h,d] * O[q,b,h,d];  tensors addressed as base + b*bs + q*rs + h*d
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ o, long long bs,
                                                         long long rs, float* __restrict__ delta, int B, int H, int Lq, int d) {
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(B) * H * Lq;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; i < total;
       i += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const int q = static_cast<int>(i % Lq);
    const int h = static_cast<int>((i / Lq) % H);
    const int b = static_cast<int>(i / (static_cast<long long>(Lq) * H));
    const long long off = b * bs + q * rs + static_cast<long long>(h) * d;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc += __bfloat162float(dout[off + c]) * __bfloat162float(o[off + c]);
    acc = warp_sum(acc);
    if (lane == 0) delta[i] = acc;
  }
}

}  // namespace

extern "C" int prismer_softmax_rows(void* s, long long rows, int Lk, int ld, cudaStream_t stream) {
  if (rows <= 0) return PRISMER_OK;
  if (Lk <= 0 || ld < Lk) return PRISMER_ERR_SHAPE;
  long long blocks = (rows * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  softmax_rows_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(reinterpret_cast<bf16*>(s), rows, Lk, ld);
  return LAUNCH_CHECK();
}

extern "C" int prismer_attn_delta(const void* dout, const void* o, long long bs, long long rs, float* delta, int B, int H, int Lq,
                                  int d, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * H * Lq;
  if (total <= 0) return PRISMER_OK;
  long long blocks = (total * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  attn_delta_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dout), reinterpret_cast<const bf16*>(o), bs,
                                                                rs, delta, B, H, Lq, d);
  return LAUNCH_CHECK();
}
