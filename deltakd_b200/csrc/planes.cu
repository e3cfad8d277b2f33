// Conversion of fp32 / bf16 activations and fp32 weights into bf16 hi/lo plane tensors (planes.cuh).
// Pure streaming kernels: 16-byte vector loads and stores, grid sized to a multiple of the SM count.
#include "planes.cuh"

namespace dkd {
namespace {

template <typename T>
__global__ void tokens_to_planes_kernel(const T* __restrict__ src, int64_t B, int T_tok, int off, int n_tok, int D, int P,
                                        const float* __restrict__ drop_mask, __nv_bfloat16* __restrict__ dst) {
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
    if (drop_mask != nullptr && __ldg(drop_mask + m) != 0.f) {  // masked token: its row is never used downstream
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    } else if constexpr (sizeof(T) == 4) {
      Vec<float, 4>::load(reinterpret_cast<const float*>(src) + srow * D + c, *reinterpret_cast<float(*)[4]>(&v[0]));
      Vec<float, 4>::load(reinterpret_cast<const float*>(src) + srow * D + c + 4, *reinterpret_cast<float(*)[4]>(&v[4]));
    } else {
      Vec<__nv_bfloat16, 8>::load(reinterpret_cast<const __nv_bfloat16*>(src) + srow * D + c, v);
    }
    Vec<__nv_bfloat16, 8>::store(dst + m * D + c, v);
    if (P >= 2) {
      float lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) lo[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
      Vec<__nv_bfloat16, 8>::store(dst + M * D + m * D + c, lo);
      if (P >= 3) {  // third plane: all 24 mantissa bits (fp32-exact operands for 6-term products)
#pragma unroll
        for (int j = 0; j < 8; ++j) lo[j] = lo[j] - __bfloat162float(__float2bfloat16_rn(lo[j]));
        Vec<__nv_bfloat16, 8>::store(dst + 2 * M * D + m * D + c, lo);
      }
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

// conv weight W[co][ci][3][3] fp32 -> forward planes Wc[P][co][tap*C + ci] and dgrad planes Wd[P][ci][(8-tap)*C + co]
__global__ void conv_weight_to_planes_kernel(const float* __restrict__ W, int C, int P, __nv_bfloat16* __restrict__ Wc,
                                             __nv_bfloat16* __restrict__ Wd) {
  const int total = C * C * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int tap = idx % 9, ci = (idx / 9) % C, co = idx / (9 * C);
    const float x = W[idx];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    const int64_t fc = (int64_t)co * 9 * C + tap * C + ci;
    const int64_t fd = (int64_t)ci * 9 * C + (8 - tap) * C + co;
    Wc[fc] = hi; Wd[fd] = hi;
    if (P == 2) { Wc[total + fc] = lo; Wd[total + fd] = lo; }
  }
}

// dWt[tap][co][ci] fp32 -> dW[co][ci][tap] (PyTorch conv weight layout)
__global__ void conv_wgrad_transpose_kernel(const float* __restrict__ dWt, int C, float* __restrict__ dW) {
  const int total = C * C * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int tap = idx % 9, ci = (idx / 9) % C, co = idx / (9 * C);
    dW[idx] = dWt[((int64_t)tap * C + co) * C + ci];
  }
}

// column sums of plane tensors [P][M][N] (hi + lo), optionally split by a 0/1 row mask:
//   out_keep[c] += sum_{rows with mask==0 (or all rows if mask==null)} x[r,c];  out_masked[c] += sum_{mask!=0} x[r,c]
__global__ void colsum_planes_kernel(const __nv_bfloat16* __restrict__ X, int64_t M, int N, int P, const float* __restrict__ mask,
                                     float* __restrict__ out_keep, float* __restrict__ out_masked) {
  const int c = threadIdx.x * 2;  // blockDim.x == N/2
  float k0 = 0.f, k1 = 0.f, m0 = 0.f, m1 = 0.f;
  for (int64_t r = blockIdx.x; r < M; r += gridDim.x) {
    float a = 0.f, b = 0.f;
    for (int pl = 0; pl < P; ++pl) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(X + (int64_t)pl * M * N + r * N + c);
      a += __low2float(h); b += __high2float(h);
    }
    if (mask != nullptr && __ldg(mask + r) != 0.f) { m0 += a; m1 += b; } else { k0 += a; k1 += b; }
  }
  if (out_keep) { atomicAdd(out_keep + c, k0); atomicAdd(out_keep + c + 1, k1); }
  if (out_masked) { atomicAdd(out_masked + c, m0); atomicAdd(out_masked + c + 1, m1); }
}

__global__ void fold_partials_kernel(const double* __restrict__ partials, int n, float scale, float* __restrict__ loss) {
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) a += partials[i];
  a = warp_sum(a);
  if (threadIdx.x == 0) *loss += (float)(a * (double)scale);
}

}  // namespace

int launch_fold_partials(const double* partials, int n, float scale, float* loss, cudaStream_t st) {
  fold_partials_kernel<<<1, 32, 0, st>>>(partials, n, scale, loss);
  return check_launch("fold_partials");
}

int launch_tokens_to_planes(const void* src, int dtype, int64_t B, int T, int off, int n_tok, int D, int P,
                            const float* drop_mask, __nv_bfloat16* dst, cudaStream_t st) {
  DKD_REQUIRE(D % 8 == 0, DKD_E_SHAPE, "tokens_to_planes: D %% 8 != 0");
  DKD_REQUIRE((((uintptr_t)src) & 15) == 0, DKD_E_ALIGN, "tokens_to_planes: source must be 16-byte aligned");
  const int64_t total = B * n_tok * (D / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  if (dtype == DKD_F32)
    tokens_to_planes_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float*>(src), B, T, off, n_tok, D, P, drop_mask, dst);
  else
    tokens_to_planes_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), B, T, off, n_tok, D, P, drop_mask, dst);
  return check_launch("tokens_to_planes");
}

int launch_weight_to_planes(const float* W, int N, int K, int P, __nv_bfloat16* Wp, __nv_bfloat16* Wt, cudaStream_t st) {
  const int total = N * K;
  int blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  weight_to_planes_kernel<<<blocks, 256, 0, st>>>(W, N, K, P, Wp, Wt);
  return check_launch("weight_to_planes");
}

int launch_conv_weight_to_planes(const float* W, int C, int P, __nv_bfloat16* Wc, __nv_bfloat16* Wd, cudaStream_t st) {
  conv_weight_to_planes_kernel<<<kNumSMs * 4, 256, 0, st>>>(W, C, P, Wc, Wd);
  return check_launch("conv_weight_to_planes");
}

int launch_conv_wgrad_transpose(const float* dWt, int C, float* dW, cudaStream_t st) {
  conv_wgrad_transpose_kernel<<<kNumSMs * 4, 256, 0, st>>>(dWt, C, dW);
  return check_launch("conv_wgrad_transpose");
}

int launch_colsum_planes(const __nv_bfloat16* X, int64_t M, int N, int P, const float* mask, float* out_keep, float* out_masked,
                         cudaStream_t st) {
  DKD_REQUIRE(N % 2 == 0 && N / 2 <= 1024, DKD_E_SHAPE, "colsum_planes: N");
  if (out_keep) cudaMemsetAsync(out_keep, 0, (size_t)N * sizeof(float), st);
  if (out_masked) cudaMemsetAsync(out_masked, 0, (size_t)N * sizeof(float), st);
  int64_t blocks = M < (int64_t)kNumSMs * 8 ? M : (int64_t)kNumSMs * 8;
  colsum_planes_kernel<<<(unsigned)blocks, N / 2, 0, st>>>(X, M, N, P, mask, out_keep, out_masked);
  return check_launch("colsum_planes");
}

int launch_fill_ones_tile(__nv_bfloat16* ones, cudaStream_t st) {
  fill_ones_tile_kernel<<<1, 256, 0, st>>>(ones);
  return check_launch("fill_ones_tile");
}

}  // namespace dkd
