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
#include "epilogues.cuh"
#include "gemm_nt.cuh"
#include "planes.cuh"

namespace dkd {
namespace {

using FwdCfg = GemmCfg<192, 1, 4, 2>;   // Y tile 128 x 192, 4-stage ring, 2 TMEM accumulators
using DgradCfg = GemmCfg<192, 1, 4, 2>; // g_s tile 128 x 192 (K = 384)
using WgradCfg = GemmNtCfg<3, true, 208, 0, 4>;  // g_W tile 128(n) x 192(k) + ones column (g_b)

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
  rc = launch_tokens_to_planes(s, dtype, B, Ts, s_off, n_tok, Ds, P, nullptr, ws.S, st);
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
    rc = launch_fold_partials(ws.partials, grid, scale, loss, st);
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
    p.ep.out = g_s; p.ep.drop_mask = nullptr; p.ep.bias = nullptr; p.ep.alpha = 1.f; p.ep.M = M; p.ep.N_total = Ds; p.ep.n_tok = n_tok; p.ep.T_out = Ts; p.ep.off = s_off;
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
    using L = NtPlainLoader<Cfg>;
    DKD_REQUIRE(g_W != nullptr, DKD_E_UNSUPPORTED, "dkd_align_mse_fwdbwd: g_b without g_W is not supported");
    GemmNtParamsT<Cfg, L> p;
    rc = make_plane_tmap(&p.ld.tmA, ws.G, P, M, Dt, Dt, M * Dt, Cfg::KROWS, "align_mse G^T");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmB, ws.S, P, M, Ds, Ds, M * Ds, Cfg::KROWS, "align_mse S (wgrad)");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmOnes, ws.ones, 2, 64, 64, 64, 64 * 64, Cfg::KROWS, "ones tile");
    if (rc != DKD_OK) return rc;
    rc = launch_fill_ones_tile(ws.ones, st);
    if (rc != DKD_OK) return rc;
    cudaMemsetAsync(g_W, 0, (size_t)Dt * Ds * sizeof(float), st);
    if (g_b) cudaMemsetAsync(g_b, 0, (size_t)Dt * sizeof(float), st);
    p.ep.D = g_W; p.ep.Dcol = g_b; p.ep.ldd = Ds; p.ep.alpha = 1.f;
    p.ld.ldd = Ds;
    p.ld.na_tiles = Dt / 128;
    p.ld.total_row_blocks = (int)((M + Cfg::KROWS - 1) / Cfg::KROWS);
    nt_make_splits(p.ld.total_row_blocks, kNumSMs / p.ld.na_tiles, &p.ld.splits, &p.ld.row_blocks_per_split);
    p.ld.b_col0 = 0;
    p.nterms = P == 2 ? 3 : 1;
    const int grid = min(kNumSMs, p.ld.na_tiles * p.ld.splits);
    auto kern = gemm_nt_kernel<Cfg, L>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_align_mse_fwdbwd: wgrad GEMM");
    if (rc != DKD_OK) return rc;
  }
  return DKD_OK;
}

}  // extern "C"
