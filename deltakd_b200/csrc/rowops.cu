// Row-wise normalisation and column reductions of the token streams that FEED the loss path
// (the callers either side of the path).  The reference builds its DeiT student / teacher with
// timm 0.9.12 (model/models.py:59-74); every transformer block there is  x + attn(LayerNorm(x)),  x + mlp(LayerNorm(x))
// with nn.LayerNorm(eps=1e-6) and nn.Linear biases.  On the [B*197, 192] student stream ATen's LayerNorm backward
// (GammaBetaBackward: 378 us per call) and the bias-gradient `sum(0)` (81 us per call) are 38 % of the whole
// distillation step on a B200 — they are plain HBM-bound passes:
//
//   dkd_layernorm_fwd : y = (x - mean) * rstd * gamma + beta              read x, write y (+ mean, rstd)
//   dkd_layernorm_bwd : dx, dgamma, dbeta from dy, x, mean, rstd, gamma    read dy + x, write dx
//   dkd_colsum        : out[n] = sum_m a[m, n]   (bias gradient)           read a
//
// One warp per row, the row held in registers (two-pass mean / variance, exact); each lane owns fixed columns, so the
// column sums (dgamma, dbeta, colsum) accumulate in registers across the warp's rows, are combined per CTA in shared
// memory and folded over the CTAs in a fixed order by a second small launch (deterministic).
// HBM-bound: algorithmic bytes = M * D * (sizeof x + sizeof y) forward, M * D * (sizeof dy + 2 sizeof x) backward.
#include "common.cuh"

namespace dkd {
namespace {

constexpr int kRowWarps = 8;                 // warps (rows in flight) per CTA
constexpr int kRowThreads = 32 * kRowWarps;
constexpr int kMaxRowCtas = kNumSMs * 4;     // persistent-style grid: CTAs stride over the rows

struct LnParams {
  const void* x;
  const void* dy;
  const void* gamma;
  const void* beta;
  void* y;
  void* dx;
  float* mean;
  float* rstd;
  float* partial;   // [gridDim.x][2][D]  (backward)
  int64_t M;
  int D;
  float eps;
};

// chunk c (4 consecutive columns) of this lane: c = lane + 32 * i, i < NCH; valid when 4 * c < D
template <typename XT, typename PT, typename YT, int NCH>
__global__ void __launch_bounds__(kRowThreads) layernorm_fwd_kernel(LnParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D, nchunks = D >> 2;
  float g[NCH][4], b[NCH][4];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      Vec<PT, 4>::load(reinterpret_cast<const PT*>(p.gamma) + 4 * c, g[i]);
      if (p.beta) Vec<PT, 4>::load(reinterpret_cast<const PT*>(p.beta) + 4 * c, b[i]);
      else { b[i][0] = b[i][1] = b[i][2] = b[i][3] = 0.f; }
    }
  }
  const float invD = 1.f / (float)D;
  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + warp; row < p.M; row += (int64_t)gridDim.x * kRowWarps) {
    const XT* xr = reinterpret_cast<const XT*>(p.x) + row * D;
    float v[NCH][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        Vec<XT, 4>::load(xr + 4 * c, v[i]);
        s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
      }
    }
    const float mean = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[i][j] -= mean; q = fmaf(v[i][j], v[i][j], q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invD + p.eps);
    YT* yr = reinterpret_cast<YT*>(p.y) + row * D;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf(v[i][j] * rstd, g[i][j], b[i][j]);
        Vec<YT, 4>::store(yr + 4 * c, o);
      }
    }
    if (lane == 0 && p.mean) { p.mean[row] = mean; p.rstd[row] = rstd; }
  }
}

template <typename GT, typename XT, typename PT, int NCH>
__global__ void __launch_bounds__(kRowThreads, NCH <= 2 ? 3 : 2) layernorm_bwd_kernel(LnParams p) {
  extern __shared__ float s_part[];   // [kRowWarps][2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D, nchunks = D >> 2;
  float g[NCH][4], dg[NCH][4], db[NCH][4];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) Vec<PT, 4>::load(reinterpret_cast<const PT*>(p.gamma) + 4 * c, g[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j) dg[i][j] = db[i][j] = 0.f;
  }
  const float invD = 1.f / (float)D;
  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + warp; row < p.M; row += (int64_t)gridDim.x * kRowWarps) {
    const XT* xr = reinterpret_cast<const XT*>(p.x) + row * D;
    const GT* gr = reinterpret_cast<const GT*>(p.dy) + row * D;
    const float mean = __ldg(p.mean + row), rstd = __ldg(p.rstd + row);
    float xh[NCH][4], gy[NCH][4];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float dyv[4];
        Vec<XT, 4>::load(xr + 4 * c, xh[i]);
        Vec<GT, 4>::load(gr + 4 * c, dyv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xh[i][j] = (xh[i][j] - mean) * rstd;
          dg[i][j] = fmaf(dyv[j], xh[i][j], dg[i][j]);
          db[i][j] += dyv[j];
          gy[i][j] = dyv[j] * g[i][j];
          c1 += gy[i][j];
          c2 = fmaf(gy[i][j], xh[i][j], c2);
        }
      }
    }
    c1 = warp_sum(c1) * invD;
    c2 = warp_sum(c2) * invD;
    if (p.dx) {
      XT* dr = reinterpret_cast<XT*>(p.dx) + row * D;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = rstd * (gy[i][j] - c1 - xh[i][j] * c2);
          Vec<XT, 4>::store(dr + 4 * c, o);
        }
      }
    }
  }
  // per-CTA column partials: warps -> shared memory -> fixed-order sum -> partial[blockIdx.x]
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s_part[(warp * 2 + 0) * D + 4 * c + j] = dg[i][j];
        s_part[(warp * 2 + 1) * D + 4 * c + j] = db[i][j];
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * D; k += kRowThreads) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) a += s_part[w * 2 * D + k];
    p.partial[(size_t)blockIdx.x * 2 * D + k] = a;
  }
}

// out[k] = sum_{c < n} partial[c][k], k < width.  32 columns x 32 row lanes per CTA: lane q sums partials q, q+32, ...
// in 4 independent chains, the 32 lanes are combined through shared memory — all in a fixed order (deterministic).
// (One thread per column walking all n partials is a chain of n/4 dependent L2 round trips: 23 us for n = 592.)
// out_a gets columns [0, split), out_b columns [split, width) (dgamma | dbeta); either may be null.
constexpr int kFoldCols = 32, kFoldLanes = 32;
__global__ void __launch_bounds__(kFoldCols * kFoldLanes) fold_columns_kernel(const float* partial, int n, int width, int split,
                                                                              float* out_a, float* out_b) {
  __shared__ float s_f[kFoldLanes][kFoldCols];
  const int col = threadIdx.x % kFoldCols, q = threadIdx.x / kFoldCols;
  const int k = blockIdx.x * kFoldCols + col;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (k < width) {
    int c = q;
    for (; c + 3 * kFoldLanes < n; c += 4 * kFoldLanes) {
      a0 += partial[(size_t)(c + 0 * kFoldLanes) * width + k];
      a1 += partial[(size_t)(c + 1 * kFoldLanes) * width + k];
      a2 += partial[(size_t)(c + 2 * kFoldLanes) * width + k];
      a3 += partial[(size_t)(c + 3 * kFoldLanes) * width + k];
    }
    for (; c < n; c += kFoldLanes) a0 += partial[(size_t)c * width + k];
  }
  s_f[q][col] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (q == 0 && k < width) {
    float r = 0.f;
#pragma unroll
    for (int j = 0; j < kFoldLanes; ++j) r += s_f[j][col];
    if (k < split) { if (out_a) out_a[k] = r; }
    else if (out_b) out_b[k - split] = r;
  }
}

// a [M, N] -> per-CTA column partials.  Thread t owns the 8-column group t % G of the rows t / G + RL * i.
template <typename T>
__global__ void __launch_bounds__(kRowThreads) colsum_kernel(const T* a, int64_t M, int N, int64_t rows_per_cta, float* partial) {
  extern __shared__ float s_cs[];   // [RL][N]
  const int G = N >> 3, RL = kRowThreads / G;
  const int cg = threadIdx.x % G, rl = threadIdx.x / G;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  if (rl < RL) {
    int64_t r = r0 + rl;
    // two rows in flight per thread
    for (; r + RL < r1; r += 2 * RL) {
      float u[8], w[8];
      if constexpr (sizeof(T) == 2) { Vec<T, 8>::load(a + r * N + 8 * cg, u); Vec<T, 8>::load(a + (r + RL) * N + 8 * cg, w); }
      else {
        Vec<T, 4>::load(a + r * N + 8 * cg, *reinterpret_cast<float(*)[4]>(&u[0]));
        Vec<T, 4>::load(a + r * N + 8 * cg + 4, *reinterpret_cast<float(*)[4]>(&u[4]));
        Vec<T, 4>::load(a + (r + RL) * N + 8 * cg, *reinterpret_cast<float(*)[4]>(&w[0]));
        Vec<T, 4>::load(a + (r + RL) * N + 8 * cg + 4, *reinterpret_cast<float(*)[4]>(&w[4]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += u[j] + w[j];
    }
    for (; r < r1; r += RL) {
      float u[8];
      if constexpr (sizeof(T) == 2) Vec<T, 8>::load(a + r * N + 8 * cg, u);
      else {
        Vec<T, 4>::load(a + r * N + 8 * cg, *reinterpret_cast<float(*)[4]>(&u[0]));
        Vec<T, 4>::load(a + r * N + 8 * cg + 4, *reinterpret_cast<float(*)[4]>(&u[4]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += u[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_cs[rl * N + 8 * cg + j] = acc[j];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < N; k += kRowThreads) {
    float s = 0.f;
    for (int q = 0; q < RL; ++q) s += s_cs[q * N + k];
    partial[(size_t)blockIdx.x * N + k] = s;
  }
}

// Head-layout copy of attention tensors: dst[b, h, n, 0:hd] = src[b, h, n, 0:hd], each side with its own (b, h, n) element
// strides, hd contiguous.  Runs are walked in (b, n, h) order, 16 bytes per thread: with a token-major side ([B, N, H, hd])
// that side is fully coalesced and the other moves whole 128-byte head rows.
struct HeadCopyParams {
  const char* src;
  char* dst;
  int64_t sb, sh, sn, db, dh, dn;   // BYTE strides
  int64_t runs;                     // B * N * H
  int H, N, vec_per_run;            // 16-byte vectors per run (hd * elt / 16)
};
__global__ void __launch_bounds__(256) head_copy_kernel(HeadCopyParams p) {
  const int64_t total = p.runs * p.vec_per_run;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t run = i / p.vec_per_run;
    const int v = (int)(i - run * p.vec_per_run);
    const int h = (int)(run % p.H);
    const int64_t bn = run / p.H;
    const int n = (int)(bn % p.N);
    const int64_t b = bn / p.N;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(p.src + b * p.sb + h * p.sh + n * p.sn) + v);
    *(reinterpret_cast<uint4*>(p.dst + b * p.db + h * p.dh + n * p.dn) + v) = w;
  }
}

int row_grid(int64_t M) {
  const int64_t ctas = (M + kRowWarps - 1) / kRowWarps;
  return (int)(ctas < kMaxRowCtas ? ctas : kMaxRowCtas);
}

template <typename XT, typename PT, typename YT>
int launch_ln_fwd(const LnParams& p, cudaStream_t st) {
  const int grid = row_grid(p.M), nch = (p.D / 4 + 31) / 32;
  switch (nch) {
    case 1: layernorm_fwd_kernel<XT, PT, YT, 1><<<grid, kRowThreads, 0, st>>>(p); break;
    case 2: layernorm_fwd_kernel<XT, PT, YT, 2><<<grid, kRowThreads, 0, st>>>(p); break;
    case 3: layernorm_fwd_kernel<XT, PT, YT, 3><<<grid, kRowThreads, 0, st>>>(p); break;
    case 4: layernorm_fwd_kernel<XT, PT, YT, 4><<<grid, kRowThreads, 0, st>>>(p); break;
    case 5: case 6: layernorm_fwd_kernel<XT, PT, YT, 6><<<grid, kRowThreads, 0, st>>>(p); break;
    default: layernorm_fwd_kernel<XT, PT, YT, 8><<<grid, kRowThreads, 0, st>>>(p); break;
  }
  return check_launch("dkd_layernorm_fwd");
}
template <typename XT, typename PT>
int launch_ln_fwd_y(const LnParams& p, int y_dtype, cudaStream_t st) {
  return y_dtype == DKD_F32 ? launch_ln_fwd<XT, PT, float>(p, st) : launch_ln_fwd<XT, PT, __nv_bfloat16>(p, st);
}

template <typename GT, typename XT, typename PT>
int launch_ln_bwd(const LnParams& p, int grid, cudaStream_t st) {
  const int nch = (p.D / 4 + 31) / 32;
  const size_t smem = (size_t)kRowWarps * 2 * p.D * sizeof(float);
  switch (nch) {
    case 1: layernorm_bwd_kernel<GT, XT, PT, 1><<<grid, kRowThreads, smem, st>>>(p); break;
    case 2: layernorm_bwd_kernel<GT, XT, PT, 2><<<grid, kRowThreads, smem, st>>>(p); break;
    case 3: layernorm_bwd_kernel<GT, XT, PT, 3><<<grid, kRowThreads, smem, st>>>(p); break;
    default: layernorm_bwd_kernel<GT, XT, PT, 4><<<grid, kRowThreads, smem, st>>>(p); break;
  }
  return check_launch("dkd_layernorm_bwd");
}
template <typename GT, typename XT>
int launch_ln_bwd_p(const LnParams& p, int p_dtype, int grid, cudaStream_t st) {
  return p_dtype == DKD_F32 ? launch_ln_bwd<GT, XT, float>(p, grid, st) : launch_ln_bwd<GT, XT, __nv_bfloat16>(p, grid, st);
}

bool dtype_ok(int d) { return d == DKD_F32 || d == DKD_BF16; }
bool aligned16(const void* q) { return (((uintptr_t)q) & 15) == 0; }

}  // namespace
}  // namespace dkd

extern "C" {

int dkd_layernorm_fwd(const void* x, const void* gamma, const void* beta, int64_t M, int D, int x_dtype, int p_dtype, int y_dtype,
                      float eps, void* y, float* mean, float* rstd, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  DKD_REQUIRE(dtype_ok(x_dtype) && dtype_ok(p_dtype) && dtype_ok(y_dtype), DKD_E_DTYPE, "dkd_layernorm_fwd: dtype");
  DKD_REQUIRE(M >= 0 && D > 0 && D % 4 == 0 && D <= 1024, DKD_E_SHAPE, "dkd_layernorm_fwd: D=%d must be a multiple of 4, <= 1024", D);
  DKD_REQUIRE(x && gamma && y && ((mean == nullptr) == (rstd == nullptr)), DKD_E_SHAPE, "dkd_layernorm_fwd: null pointer");
  DKD_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), DKD_E_ALIGN, "dkd_layernorm_fwd: 16-byte alignment");
  if (M == 0) return DKD_OK;
  LnParams p{};
  p.x = x; p.gamma = gamma; p.beta = beta; p.y = y; p.mean = mean; p.rstd = rstd; p.M = M; p.D = D; p.eps = eps;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_dtype == DKD_F32) return p_dtype == DKD_F32 ? launch_ln_fwd_y<float, float>(p, y_dtype, st) : launch_ln_fwd_y<float, __nv_bfloat16>(p, y_dtype, st);
  return p_dtype == DKD_F32 ? launch_ln_fwd_y<__nv_bfloat16, float>(p, y_dtype, st) : launch_ln_fwd_y<__nv_bfloat16, __nv_bfloat16>(p, y_dtype, st);
}

size_t dkd_layernorm_bwd_workspace_bytes(int64_t M, int D) {
  return (size_t)dkd::row_grid(M > 0 ? M : 1) * 2 * (size_t)(D > 0 ? D : 0) * sizeof(float) + 256;
}

int dkd_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, int64_t M, int D,
                      int dy_dtype, int x_dtype, int p_dtype, void* dx, float* dgamma, float* dbeta, void* workspace,
                      size_t workspace_bytes, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  DKD_REQUIRE(dtype_ok(x_dtype) && dtype_ok(p_dtype) && dtype_ok(dy_dtype), DKD_E_DTYPE, "dkd_layernorm_bwd: dtype");
  DKD_REQUIRE(M > 0 && D > 0 && D % 4 == 0 && D <= 512, DKD_E_SHAPE, "dkd_layernorm_bwd: D=%d must be a multiple of 4, <= 512", D);
  DKD_REQUIRE(dy && x && gamma && mean && rstd && workspace, DKD_E_SHAPE, "dkd_layernorm_bwd: null pointer");
  DKD_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(gamma) && aligned16(workspace), DKD_E_ALIGN,
              "dkd_layernorm_bwd: 16-byte alignment");
  DKD_REQUIRE(workspace_bytes >= dkd_layernorm_bwd_workspace_bytes(M, D), DKD_E_WORKSPACE, "dkd_layernorm_bwd: workspace too small");
  LnParams p{};
  p.x = x; p.dy = dy; p.gamma = gamma; p.dx = dx; p.mean = const_cast<float*>(mean); p.rstd = const_cast<float*>(rstd);
  p.partial = reinterpret_cast<float*>(workspace); p.M = M; p.D = D;
  const int grid = row_grid(M);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dy_dtype == DKD_F32) rc = x_dtype == DKD_F32 ? launch_ln_bwd_p<float, float>(p, p_dtype, grid, st) : launch_ln_bwd_p<float, __nv_bfloat16>(p, p_dtype, grid, st);
  else rc = x_dtype == DKD_F32 ? launch_ln_bwd_p<__nv_bfloat16, float>(p, p_dtype, grid, st) : launch_ln_bwd_p<__nv_bfloat16, __nv_bfloat16>(p, p_dtype, grid, st);
  if (rc != DKD_OK) return rc;
  if (dgamma || dbeta) {
    fold_columns_kernel<<<(2 * D + kFoldCols - 1) / kFoldCols, kFoldCols * kFoldLanes, 0, st>>>(p.partial, grid, 2 * D, D, dgamma, dbeta);
    rc = check_launch("dkd_layernorm_bwd: fold");
  }
  return rc;
}

static int colsum_grid(int64_t M, int N, int64_t* rows_per_cta) {
  const int RL = dkd::kRowThreads / (N / 8);
  int64_t rows = (M + dkd::kMaxRowCtas - 1) / dkd::kMaxRowCtas;
  const int64_t min_rows = 4 * (int64_t)RL;   // at least 4 rows per thread
  if (rows < min_rows) rows = min_rows;
  *rows_per_cta = rows;
  return (int)((M + rows - 1) / rows);
}

int dkd_head_copy(const void* src, void* dst, int64_t B, int H, int N, int hd, int elt_bytes, int64_t src_b, int64_t src_h,
                  int64_t src_n, int64_t dst_b, int64_t dst_h, int64_t dst_n, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  DKD_REQUIRE(B >= 0 && H > 0 && N > 0 && hd > 0 && (elt_bytes == 2 || elt_bytes == 4), DKD_E_SHAPE, "dkd_head_copy: bad shape");
  DKD_REQUIRE(((int64_t)hd * elt_bytes) % 16 == 0, DKD_E_SHAPE, "dkd_head_copy: head rows must be multiples of 16 bytes");
  DKD_REQUIRE(src && dst, DKD_E_SHAPE, "dkd_head_copy: null pointer");
  const int64_t strides[6] = {src_b, src_h, src_n, dst_b, dst_h, dst_n};
  for (int i = 0; i < 6; ++i)
    DKD_REQUIRE(strides[i] >= 0 && (strides[i] * elt_bytes) % 16 == 0, DKD_E_ALIGN, "dkd_head_copy: strides must be multiples of 16 bytes");
  DKD_REQUIRE(aligned16(src) && aligned16(dst), DKD_E_ALIGN, "dkd_head_copy: 16-byte alignment");
  if (B == 0) return DKD_OK;
  HeadCopyParams p;
  p.src = reinterpret_cast<const char*>(src); p.dst = reinterpret_cast<char*>(dst);
  p.sb = src_b * elt_bytes; p.sh = src_h * elt_bytes; p.sn = src_n * elt_bytes;
  p.db = dst_b * elt_bytes; p.dh = dst_h * elt_bytes; p.dn = dst_n * elt_bytes;
  p.runs = B * N * H; p.H = H; p.N = N; p.vec_per_run = hd * elt_bytes / 16;
  const int64_t total = p.runs * p.vec_per_run;
  const int64_t want = (total + 256 * 4 - 1) / (256 * 4);          // ~4 vectors per thread
  const int grid = (int)(want < 1 ? 1 : want > kNumSMs * 16 ? kNumSMs * 16 : want);
  head_copy_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("dkd_head_copy");
}

size_t dkd_colsum_workspace_bytes(int64_t M, int N) {
  if (M <= 0 || N <= 0 || N % 8 != 0 || N / 8 > dkd::kRowThreads) return 256;
  int64_t rows;
  return (size_t)colsum_grid(M, N, &rows) * (size_t)N * sizeof(float) + 256;
}

int dkd_colsum(const void* a, int64_t M, int N, int dtype, float* out, void* workspace, size_t workspace_bytes, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  DKD_REQUIRE(dtype_ok(dtype), DKD_E_DTYPE, "dkd_colsum: dtype %d", dtype);
  DKD_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && N / 8 <= kRowThreads, DKD_E_SHAPE, "dkd_colsum: N=%d must be a multiple of 8, <= %d", N, 8 * kRowThreads);
  DKD_REQUIRE(a && out && workspace, DKD_E_SHAPE, "dkd_colsum: null pointer");
  DKD_REQUIRE(aligned16(a) && aligned16(workspace), DKD_E_ALIGN, "dkd_colsum: 16-byte alignment");
  DKD_REQUIRE(workspace_bytes >= dkd_colsum_workspace_bytes(M, N), DKD_E_WORKSPACE, "dkd_colsum: workspace too small");
  int64_t rows;
  const int grid = colsum_grid(M, N, &rows);
  const int RL = kRowThreads / (N / 8);
  const size_t smem = (size_t)RL * N * sizeof(float);
  float* partial = reinterpret_cast<float*>(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DKD_F32) colsum_kernel<float><<<grid, kRowThreads, smem, st>>>(reinterpret_cast<const float*>(a), M, N, rows, partial);
  else colsum_kernel<__nv_bfloat16><<<grid, kRowThreads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(a), M, N, rows, partial);
  rc = check_launch("dkd_colsum");
  if (rc != DKD_OK) return rc;
  fold_columns_kernel<<<(N + kFoldCols - 1) / kFoldCols, kFoldCols * kFoldLanes, 0, st>>>(partial, grid, N, N, out, nullptr);
  return check_launch("dkd_colsum: fold");
}

}  // extern "C"
