// Conversion of fp32 / bf16 activations and fp32 weights into bf16 hi/lo plane tensors (planes.cuh).
// Pure streaming kernels: 16-byte vector loads and stores, grid sized to a multiple of the SM count.
#include "planes.cuh"

namespace dkd {
namespace {

template <typename T>
__global__ void tokens_to_planes_kernel(const T* __restrict__ src, int64_t B, int T_tok, int off, int n_tok, int D, int P,
                                        __nv_bfloat16* __restrict__ dst) {
  const int vec_per_row = D / 8;
  const int64_t M = B * n_tok;
  const int64_t total = M * vec_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int64_t m = idx / vec_per_row;
    const int c = (int)(idx - m * vec_per_row) * 8;
    const int64_t b = m / n_tok;
    const int64_t srow = b * T_tok + off + (m - b * n_tok);
    float v[8];
    if constexpr (sizeof(T) == 4) {
      Vec<float, 4>::load(reinterpret_cast<const float*>(src) + srow * D + c, *reinterpret_cast<float(*)[4]>(&v[0]));
      Vec<float, 4>::load(reinterpret_cast<const float*>(src) + srow * D + c + 4, *reinterpret_cast<float(*)[4]>(&v[4]));
    } else {
      Vec<__nv_bfloat16, 8>::load(reinterpret_cast<const __nv_bfloat16*>(src) + srow * D + c, v);
    }
    Vec<__nv_bfloat16, 8>::store(dst + m * D + c, v);
    if (P == 2) {
      float lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) lo[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
      Vec<__nv_bfloat16, 8>::store(dst + M * D + m * D + c, lo);
    }
  }
}

__global__ void weight_to_planes_kernel(const float* __restrict__ W, int N, int K, int P, __nv_bfloat16* __restrict__ Wp,
                                        __nv_bfloat16* __restrict__ Wt) {
  const int total = N * K;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx / K, k = idx - n * K;
    const float x = W[idx];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    if (Wp) { Wp[idx] = hi; if (P == 2) Wp[total + idx] = lo; }
    if (Wt) { Wt[k * N + n] = hi; if (P == 2) Wt[total + k * N + n] = lo; }
  }
}

__global__ void fill_ones_tile_kernel(__nv_bfloat16* ones) {
  for (int idx = threadIdx.x; idx < 2 * 64 * 64; idx += blockDim.x)
    ones[idx] = __float2bfloat16_rn((idx < 64 * 64 && (idx & 63) == 0) ? 1.f : 0.f);
}

}  // namespace

int launch_tokens_to_planes(const void* src, int dtype, int64_t B, int T, int off, int n_tok, int D, int P,
                            __nv_bfloat16* dst, cudaStream_t st) {
  DKD_REQUIRE(D % 8 == 0, DKD_E_SHAPE, "tokens_to_planes: D %% 8 != 0");
  DKD_REQUIRE((((uintptr_t)src) & 15) == 0, DKD_E_ALIGN, "tokens_to_planes: source must be 16-byte aligned");
  const int64_t total = B * n_tok * (D / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  if (dtype == DKD_F32)
    tokens_to_planes_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float*>(src), B, T, off, n_tok, D, P, dst);
  else
    tokens_to_planes_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), B, T, off, n_tok, D, P, dst);
  return check_launch("tokens_to_planes");
}

int launch_weight_to_planes(const float* W, int N, int K, int P, __nv_bfloat16* Wp, __nv_bfloat16* Wt, cudaStream_t st) {
  const int total = N * K;
  int blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  weight_to_planes_kernel<<<blocks, 256, 0, st>>>(W, N, K, P, Wp, Wt);
  return check_launch("weight_to_planes");
}

int launch_fill_ones_tile(__nv_bfloat16* ones, cudaStream_t st) {
  fill_ones_tile_kernel<<<1, 256, 0, st>>>(ones);
  return check_launch("fill_ones_tile");
}

}  // namespace dkd
