// Layer-wise hidden-state matching: loss += scale * || s[:,off:] W^T + b - t[:,off:] ||^2 with all gradients.
// Reference: curkd_loss early / mid (model/loss.py:376-393) and the ViTKD mimicking term (loss.py:277-289)
// — per layer `mse_loss(Linear(192->384)(student[:,1:]), teacher[:,2:], reduction='sum')` and its autograd
// backward (grad of the student feature, of the Linear weight and of its bias).
//
// Launch sequence per layer (all on `stream`, no host sync):
//   1. tokens_to_planes   s[B,Ts,Ds] (patch tokens only) -> bf16 planes S[P][M][Ds]      (P = 1 bf16 | 2 bf16x3)
//   2. weight_to_planes   W[Dt,Ds] -> Wp[P][Dt][Ds] and its transpose Wt[P][Ds][Dt]
//   3. gemm_tn + ResidualMse epilogue : Y = S W^T (tcgen05, TMEM) ; d = Y + b - t ; loss partial ; G = 2*scale*d
//                                       written as bf16 planes G[P][M][Dt]            (teacher read in place)
//   4. gemm_tn + StoreRows epilogue   : g_s[:,off:,:] = G W    (rows scattered into [B,Ts,Ds], CLS rows zeroed)
//   5. gemm_nt (split-K, MN-major)    : g_W += G^T S, g_b += G^T 1 (ones-column trick)
//   6. fold_partials                  : loss += sum of the per-CTA partials, fixed order (deterministic)
#include "gemm_nt.cuh"
#include "gemm_tn.cuh"
#include "planes.cuh"

namespace dkd {
namespace {

// ---------------------------------------------------------------------------------- epilogues
struct ResidualMseParams {
  const void* t;          // teacher [B, Tt, N] (fp32 or bf16), rows b*Tt + t_off + i
  const float* bias;      // [N] or null
  __nv_bfloat16* G;       // planes [P][M][N]
  double* partials;       // [gridDim.x]
  int64_t M;
  int N, n_tok, Tt, t_off, planes;
  float gscale;           // G = gscale * d
  int t_is_bf16;
};

template <class Cfg>
struct ResidualMseEpi {
  using Params = ResidualMseParams;
  struct State { float acc; };
  static __device__ __forceinline__ void init(const Params&, State& st) { st.acc = 0.f; }

  static __device__ __forceinline__ void tile(const Params& p, State& st, int m0, int n0, int row_in_tile, uint32_t t_acc) {
    const int64_t m = (int64_t)m0 + row_in_tile;
    const bool live = m < p.M;
    const int64_t b = live ? m / p.n_tok : 0;
    const int64_t trow = b * p.Tt + p.t_off + (live ? m - b * p.n_tok : 0);
#pragma unroll 1
    for (int c0 = 0; c0 < Cfg::BN; c0 += 32) {
      float v[32];
      sm100::tmem_ld32(t_acc + c0, v);
      float tv[32];
      if (live) {
        if (p.t_is_bf16) {
          const __nv_bfloat16* tp = reinterpret_cast<const __nv_bfloat16*>(p.t) + trow * p.N + n0 + c0;
#pragma unroll
          for (int j = 0; j < 4; ++j) Vec<__nv_bfloat16, 8>::load(tp + 8 * j, *reinterpret_cast<float(*)[8]>(&tv[8 * j]));
        } else {
          const float* tp = reinterpret_cast<const float*>(p.t) + trow * p.N + n0 + c0;
#pragma unroll
          for (int j = 0; j < 8; ++j) Vec<float, 4>::load(tp + 4 * j, *reinterpret_cast<float(*)[4]>(&tv[4 * j]));
        }
      }
      sm100::tmem_ld_wait();
      if (live) {
        float hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = v[j] + (p.bias ? __ldg(p.bias + n0 + c0 + j) : 0.f) - tv[j];
          st.acc = fmaf(d, d, st.acc);
          const float g = p.gscale * d;
          hi[j] = g;
          lo[j] = g - __bfloat162float(__float2bfloat16_rn(g));
        }
        __nv_bfloat16* gp = p.G + m * p.N + n0 + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) Vec<__nv_bfloat16, 8>::store(gp + 8 * j, *reinterpret_cast<float(*)[8]>(&hi[8 * j]));
        if (p.planes == 2) {
          __nv_bfloat16* gl = gp + p.M * p.N;
#pragma unroll
          for (int j = 0; j < 4; ++j) Vec<__nv_bfloat16, 8>::store(gl + 8 * j, *reinterpret_cast<float(*)[8]>(&lo[8 * j]));
        }
      }
    }
  }

  static __device__ __forceinline__ void finish(const Params& p, State& st, int tid) {
    __shared__ double s_part[4];
    double a = (double)st.acc;
    a = warp_sum(a);
    if ((tid & 31) == 0) s_part[tid >> 5] = a;
    asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
    if (tid == 0) p.partials[blockIdx.x] = s_part[0] + s_part[1] + s_part[2] + s_part[3];
  }
};

struct StoreRowsParams {
  void* out;              // [B, T_out, N_total] fp32 or bf16; rows b*T_out + off + i
  int64_t M;
  int N_total, n_tok, T_out, off;
  int out_is_bf16;
};

template <class Cfg>
struct StoreRowsEpi {
  using Params = StoreRowsParams;
  struct State {};
  static __device__ __forceinline__ void init(const Params&, State&) {}
  static __device__ __forceinline__ void tile(const Params& p, State&, int m0, int n0, int row_in_tile, uint32_t t_acc) {
    const int64_t m = (int64_t)m0 + row_in_tile;
    const bool live = m < p.M;
    const int64_t b = live ? m / p.n_tok : 0;
    const int64_t i = live ? m - b * p.n_tok : 0;
    const int64_t orow = b * p.T_out + p.off + i;
#pragma unroll 1
    for (int c0 = 0; c0 < Cfg::BN; c0 += 32) {
      float v[32];
      sm100::tmem_ld32(t_acc + c0, v);
      sm100::tmem_ld_wait();
      if (!live) continue;
      float z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0.f;
      if (p.out_is_bf16) {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.N_total + n0 + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) Vec<__nv_bfloat16, 8>::store(op + 8 * j, *reinterpret_cast<float(*)[8]>(&v[8 * j]));
        if (i == 0)  // the special-token rows in front of this sample's patches get zero gradient
          for (int r = 1; r <= p.off; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) Vec<__nv_bfloat16, 8>::store(op - (int64_t)r * p.N_total + 8 * j, *reinterpret_cast<float(*)[8]>(&z[8 * j]));
      } else {
        float* op = reinterpret_cast<float*>(p.out) + orow * p.N_total + n0 + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) Vec<float, 4>::store(op + 4 * j, *reinterpret_cast<float(*)[4]>(&v[4 * j]));
        if (i == 0)
          for (int r = 1; r <= p.off; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) Vec<float, 4>::store(op - (int64_t)r * p.N_total + 4 * j, *reinterpret_cast<float(*)[4]>(&z[4 * j]));
      }
    }
  }
  static __device__ __forceinline__ void finish(const Params&, State&, int) {}
};

__global__ void fold_partials_kernel(const double* __restrict__ partials, int n, float scale, float* __restrict__ loss) {
  // one warp, fixed order: deterministic
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) a += partials[i];
  a = warp_sum(a);
  if (threadIdx.x == 0) *loss += (float)(a * (double)scale);
}

using FwdCfg = GemmCfg<192, 1, 4, 2>;   // Y tile 128 x 192, 4-stage ring, 2 TMEM accumulators
using DgradCfg = GemmCfg<192, 1, 4, 2>; // g_s tile 128 x 192 (K = 384)
using WgradCfg = GemmNtCfg<3, true, 1, 4>;  // g_W tile 128(n) x 192(k) + ones column (g_b)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Workspace {
  __nv_bfloat16 *S, *Wp, *Wt, *G, *ones;
  double* partials;
  size_t bytes;
};
Workspace carve(void* base, int64_t M, int Ds, int Dt, int P) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 1024); return reinterpret_cast<char*>(base) + o; };
  w.S = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * M * Ds * 2));
  w.G = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * M * Dt * 2));
  w.Wp = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * Dt * Ds * 2));
  w.Wt = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * Dt * Ds * 2));
  w.ones = reinterpret_cast<__nv_bfloat16*>(take((size_t)2 * 64 * 64 * 2));
  w.partials = reinterpret_cast<double*>(take((size_t)kNumSMs * sizeof(double)));
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace dkd

extern "C" {

size_t dkd_align_mse_workspace_bytes(int64_t B, int n_tok, int Ds, int Dt, int precision) {
  return dkd::carve(nullptr, B * n_tok, Ds, Dt, precision == DKD_PREC_BF16X3 ? 2 : 1).bytes;
}

int dkd_align_mse_fwdbwd(const void* s, const void* t, const float* W, const float* bias, int64_t B, int Ts, int s_off,
                         int Tt, int t_off, int n_tok, int Ds, int Dt, int dtype, int precision, float scale, void* g_s,
                         float* g_W, float* g_b, float* loss, void* workspace, size_t workspace_bytes, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  DKD_REQUIRE(dtype == DKD_F32 || dtype == DKD_BF16, DKD_E_DTYPE, "dkd_align_mse_fwdbwd: dtype %d", dtype);
  DKD_REQUIRE(precision == DKD_PREC_BF16 || precision == DKD_PREC_BF16X3, DKD_E_UNSUPPORTED, "dkd_align_mse_fwdbwd: precision %d", precision);
  DKD_REQUIRE(B > 0 && n_tok > 0 && s_off >= 0 && t_off >= 0 && Ts >= s_off + n_tok && Tt >= t_off + n_tok, DKD_E_SHAPE,
              "dkd_align_mse_fwdbwd: bad token geometry");
  DKD_REQUIRE(Ds == 192 && Dt == 384, DKD_E_SHAPE, "dkd_align_mse_fwdbwd: built for DeiT-Tiny -> DeiT-Small widths (192 -> 384), got %d -> %d", Ds, Dt);
  DKD_REQUIRE(s && t && W && loss && workspace, DKD_E_SHAPE, "dkd_align_mse_fwdbwd: null pointer");
  DKD_REQUIRE((((uintptr_t)workspace) & 1023) == 0, DKD_E_ALIGN, "dkd_align_mse_fwdbwd: workspace must be 1024-byte aligned");
  const int P = precision == DKD_PREC_BF16X3 ? 2 : 1;
  const int64_t M = B * n_tok;
  DKD_REQUIRE(M < (1ll << 31) - 256, DKD_E_SHAPE, "dkd_align_mse_fwdbwd: too many rows");
  Workspace ws = carve(workspace, M, Ds, Dt, P);
  DKD_REQUIRE(workspace_bytes >= ws.bytes, DKD_E_WORKSPACE, "dkd_align_mse_fwdbwd: workspace %zu < %zu", workspace_bytes, ws.bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool want_grads = g_s != nullptr || g_W != nullptr || g_b != nullptr;

  // 1-2. operand planes
  rc = launch_tokens_to_planes(s, dtype, B, Ts, s_off, n_tok, Ds, P, ws.S, st);
  if (rc != DKD_OK) return rc;
  rc = launch_weight_to_planes(W, Dt, Ds, P, ws.Wp, want_grads ? ws.Wt : nullptr, st);
  if (rc != DKD_OK) return rc;

  // 3. forward GEMM + residual epilogue
  {
    using Cfg = FwdCfg;
    using L = PlaneLoader<Cfg>;
    using E = ResidualMseEpi<Cfg>;
    GemmParams<L, E> p;
    rc = make_plane_tmap(&p.ld.tmA, ws.S, P, M, Ds, Ds, M * Ds, Cfg::BM, "align_mse S");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmB, ws.Wp, P, Dt, Ds, Ds, (int64_t)Dt * Ds, Cfg::BN, "align_mse W");
    if (rc != DKD_OK) return rc;
    p.ld.k_blocks = Ds / 64;
    p.ld.nterms = P == 2 ? 3 : 1;
    p.ep.t = t; p.ep.bias = bias; p.ep.G = ws.G; p.ep.partials = ws.partials;
    p.ep.M = M; p.ep.N = Dt; p.ep.n_tok = n_tok; p.ep.Tt = Tt; p.ep.t_off = t_off; p.ep.planes = P;
    p.ep.gscale = 2.f * scale; p.ep.t_is_bf16 = dtype == DKD_BF16;
    p.m_tiles = (int)((M + Cfg::BM - 1) / Cfg::BM);
    p.n_tiles = Dt / Cfg::BN;
    const int grid = min(kNumSMs, p.m_tiles * p.n_tiles);
    auto kern = gemm_tn_kernel<Cfg, L, E>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_align_mse_fwdbwd: forward GEMM");
    if (rc != DKD_OK) return rc;
    fold_partials_kernel<<<1, 32, 0, st>>>(ws.partials, grid, scale, loss);
    rc = check_launch("dkd_align_mse_fwdbwd: fold");
    if (rc != DKD_OK) return rc;
  }
  if (!want_grads) return DKD_OK;

  // 4. g_s = G W  (contract over Dt)
  if (g_s) {
    using Cfg = DgradCfg;
    using L = PlaneLoader<Cfg>;
    using E = StoreRowsEpi<Cfg>;
    GemmParams<L, E> p;
    rc = make_plane_tmap(&p.ld.tmA, ws.G, P, M, Dt, Dt, M * Dt, Cfg::BM, "align_mse G");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmB, ws.Wt, P, Ds, Dt, Dt, (int64_t)Dt * Ds, Cfg::BN, "align_mse W^T");
    if (rc != DKD_OK) return rc;
    p.ld.k_blocks = Dt / 64;
    p.ld.nterms = P == 2 ? 3 : 1;
    p.ep.out = g_s; p.ep.M = M; p.ep.N_total = Ds; p.ep.n_tok = n_tok; p.ep.T_out = Ts; p.ep.off = s_off;
    p.ep.out_is_bf16 = dtype == DKD_BF16;
    p.m_tiles = (int)((M + Cfg::BM - 1) / Cfg::BM);
    p.n_tiles = Ds / Cfg::BN;
    const int grid = min(kNumSMs, p.m_tiles * p.n_tiles);
    auto kern = gemm_tn_kernel<Cfg, L, E>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_align_mse_fwdbwd: dgrad GEMM");
    if (rc != DKD_OK) return rc;
  }

  // 5. g_W = G^T S, g_b = G^T 1
  if (g_W || g_b) {
    using Cfg = WgradCfg;
    DKD_REQUIRE(g_W != nullptr, DKD_E_UNSUPPORTED, "dkd_align_mse_fwdbwd: g_b without g_W is not supported");
    GemmNtParams p;
    rc = make_plane_tmap(&p.tmA, ws.G, P, M, Dt, Dt, M * Dt, 64, "align_mse G^T");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.tmB, ws.S, P, M, Ds, Ds, M * Ds, 64, "align_mse S (wgrad)");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.tmOnes, ws.ones, 2, 64, 64, 64, 64 * 64, 64, "ones tile");
    if (rc != DKD_OK) return rc;
    launch_fill_ones_tile(ws.ones, st);
    cudaMemsetAsync(g_W, 0, (size_t)Dt * Ds * sizeof(float), st);
    if (g_b) cudaMemsetAsync(g_b, 0, (size_t)Dt * sizeof(float), st);
    p.D = g_W; p.Dcol = g_b; p.ldd = Ds;
    p.na_tiles = Dt / 128;
    p.total_row_blocks = (int)((M + 63) / 64);
    int splits = max(1, kNumSMs / p.na_tiles);
    p.row_blocks_per_split = (p.total_row_blocks + splits - 1) / splits;
    p.splits = (p.total_row_blocks + p.row_blocks_per_split - 1) / p.row_blocks_per_split;
    p.nterms = P == 2 ? 3 : 1;
    p.b_col0 = 0;
    p.alpha = 1.f;
    const int grid = min(kNumSMs, p.na_tiles * p.splits);
    auto kern = gemm_nt_kernel<Cfg>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_align_mse_fwdbwd: wgrad GEMM");
    if (rc != DKD_OK) return rc;
  }
  return DKD_OK;
}

}  // extern "C"
