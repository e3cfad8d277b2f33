// WassKD 'l1' term: sorted-L1 (1-D Wasserstein) distance per (sample, channel) over the token axis.
// Reference: model/loss.py:187-199 —
//     a = align_wasskd[i](student[i][:, 1:]);  sort(a, dim=1), sort(teacher[i][:, 2:], dim=1);  mean|diff|
// and its autograd backward: g_a[b, pi(r), d] = sign(sorted_a[b,r,d] - sorted_t[b,r,d]) / numel, with pi the
// student sort permutation.
//
//   1-2. operand planes (S, W) as in align_mse.cu
//   3.   a = S W^T + b  (gemm_tn, fp32 rows to scratch)
//   4.   sort kernel: one CTA per (sample, 32-channel group); the [n_tok x 32] student and teacher tiles are
//        staged in shared memory with coalesced 128-byte rows; each warp sorts whole columns with a 256-wide
//        bitonic network held in registers (8 keys per lane, cross-lane steps by warp shuffle; the student
//        carries its token index as payload, ties ordered by index); |diff| is block-reduced and sign(diff)
//        is scattered back through shared memory so the +-1 gradient plane is written with full rows.
//   5-6. g_s = c * G W, g_W = c * G^T S, g_b = c * G^T 1  (tcgen05; c = scale folded into the epilogues)
// The sort is on-chip (shared memory + registers): HBM traffic is the a/t tile reads and the +-1 plane write.
#include "epilogues.cuh"
#include "gemm_nt.cuh"
#include "planes.cuh"

namespace dkd {
namespace {

constexpr int kSortThreads = 256;
constexpr int kCh = 32;        // channels per CTA
constexpr int kMaxTok = 256;   // bitonic width

struct KV { float v; int i; };
__device__ __forceinline__ bool kv_less(const KV& a, const KV& b) { return a.v < b.v || (a.v == b.v && a.i < b.i); }

// 256 (key, index) pairs per warp: element e = r*32 + lane lives in slot r of lane.  Ascending on exit.
template <bool WITH_INDEX>
__device__ __forceinline__ void warp_bitonic_256(float (&key)[8], int (&idx)[8], int lane) {
#pragma unroll
  for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int pr = r ^ jr;
          if (pr > r) {
            const int e = r * 32 + lane;
            const bool up = (e & k) == 0;
            KV a{key[r], idx[r]}, b{key[pr], idx[pr]};
            const bool swap = up ? kv_less(b, a) : kv_less(a, b);
            if (swap) {
              key[r] = b.v; key[pr] = a.v;
              if (WITH_INDEX) { idx[r] = b.i; idx[pr] = a.i; }
            }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int e = r * 32 + lane;
          const bool up = (e & k) == 0;
          const bool lower = (lane & j) == 0;       // this lane holds the lower-index element of the pair
          KV mine{key[r], idx[r]}, other;
          other.v = __shfl_xor_sync(0xffffffffu, key[r], j);
          other.i = WITH_INDEX ? __shfl_xor_sync(0xffffffffu, idx[r], j) : 0;
          if (!WITH_INDEX) { mine.i = lower ? 0 : 1; other.i = lower ? 1 : 0; }  // stable tie-break by position
          const bool keep_min = (lower == up);
          const bool take_other = keep_min ? kv_less(other, mine) : kv_less(mine, other);
          if (take_other) { key[r] = other.v; if (WITH_INDEX) idx[r] = other.i; }
        }
      }
    }
  }
}

struct SortParams {
  const float* a;        // [M, N] fp32 aligned student (scratch)
  const void* t;         // teacher [B, Tt, N]
  __nv_bfloat16* G;      // [P][M][N] : plane 0 = sign(diff) in {-1,0,+1}, plane 1 (if any) = 0
  double* partials;      // [gridDim]
  int64_t M;
  int N, n_tok, Tt, t_off, planes, t_is_bf16, write_grad;
};

__global__ void __launch_bounds__(kSortThreads) wass_sort_kernel(SortParams p) {
  extern __shared__ float smem_f[];
  float (*sa)[kCh + 1] = reinterpret_cast<float (*)[kCh + 1]>(smem_f);                          // [n_tok][33]
  float (*st)[kCh + 1] = reinterpret_cast<float (*)[kCh + 1]>(smem_f + (size_t)p.n_tok * (kCh + 1));
  __shared__ float red[kSortThreads / 32];
  const int groups = p.N / kCh;
  const int b = blockIdx.x / groups, c0 = (blockIdx.x % groups) * kCh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // coalesced tile loads: 32 consecutive channels (128 B) per token row
  for (int tok = warp; tok < p.n_tok; tok += kSortThreads / 32) {
    const int64_t toff = ((int64_t)b * p.Tt + p.t_off + tok) * p.N + c0 + lane;
    sa[tok][lane] = p.a[((int64_t)b * p.n_tok + tok) * p.N + c0 + lane];
    st[tok][lane] = p.t_is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.t)[toff]) : reinterpret_cast<const float*>(p.t)[toff];
  }
  __syncthreads();

  float local = 0.f;
  for (int col = warp; col < kCh; col += kSortThreads / 32) {
    float ka[8], kt[8];
    int ia[8], it_[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = r * 32 + lane;
      const bool real = e < p.n_tok;          // slots beyond n_tok are +inf padding that sorts to the end
      ka[r] = real ? sa[e][col] : INFINITY; ia[r] = e;
      kt[r] = real ? st[e][col] : INFINITY; it_[r] = 0;
    }
    warp_bitonic_256<true>(ka, ia, lane);
    warp_bitonic_256<false>(kt, it_, lane);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = r * 32 + lane;
      if (e < p.n_tok) {
        const float d = ka[r] - kt[r];
        local += fabsf(d);
        // gradient of |d| w.r.t. the student element that landed at sorted position e
        sa[ia[r]][col] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      }
    }
  }
  local = warp_sum(local);
  if (lane == 0) red[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kSortThreads / 32; ++w) s += (double)red[w];
    p.partials[blockIdx.x] = s;
  }
  if (!p.write_grad) return;
  // +-1 gradient plane, 64-byte rows per CTA (32 bf16 channels)
  for (int tok = warp; tok < p.n_tok; tok += kSortThreads / 32) {
    const int64_t off = ((int64_t)b * p.n_tok + tok) * p.N + c0 + lane;
    p.G[off] = __float2bfloat16_rn(sa[tok][lane]);
    if (p.planes == 2) p.G[p.M * p.N + off] = __float2bfloat16_rn(0.f);
  }
}

using FwdCfg = GemmCfg<192, 1, 4, 2>;
using DgradCfg = GemmCfg<192, 1, 4, 2>;
using WgradCfg = GemmNtCfg<3, true, 208, 0, 4>;

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Workspace {
  __nv_bfloat16 *S, *Wp, *Wt, *G, *ones;
  float* A;
  double* partials;
  size_t bytes;
};
Workspace carve(void* base, int64_t B, int64_t M, int Ds, int Dt, int P) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 1024); return reinterpret_cast<char*>(base) + o; };
  w.S = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * M * Ds * 2));
  w.G = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * M * Dt * 2));
  w.A = reinterpret_cast<float*>(take((size_t)M * Dt * 4));
  w.Wp = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * Dt * Ds * 2));
  w.Wt = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * Dt * Ds * 2));
  w.ones = reinterpret_cast<__nv_bfloat16*>(take((size_t)2 * 64 * 64 * 2));
  w.partials = reinterpret_cast<double*>(take((size_t)B * (Dt / kCh) * sizeof(double)));
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace dkd

extern "C" {

size_t dkd_wass_l1_workspace_bytes(int64_t B, int n_tok, int Ds, int Dt, int precision) {
  return dkd::carve(nullptr, B, B * n_tok, Ds, Dt, precision == DKD_PREC_BF16X3 ? 2 : 1).bytes;
}

int dkd_wass_l1_fwdbwd(const void* s, const void* t, const float* W, const float* bias, int64_t B, int Ts, int s_off, int Tt,
                       int t_off, int n_tok, int Ds, int Dt, int dtype, int precision, float scale, void* g_s, float* g_W,
                       float* g_b, float* loss, void* workspace, size_t workspace_bytes, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  const char* fn = "dkd_wass_l1_fwdbwd";
  DKD_REQUIRE(dtype == DKD_F32 || dtype == DKD_BF16, DKD_E_DTYPE, "%s: dtype %d", fn, dtype);
  DKD_REQUIRE(precision == DKD_PREC_BF16 || precision == DKD_PREC_BF16X3, DKD_E_UNSUPPORTED, "%s: precision %d", fn, precision);
  DKD_REQUIRE(B > 0 && n_tok > 0 && n_tok <= kMaxTok && s_off >= 0 && t_off >= 0 && Ts >= s_off + n_tok && Tt >= t_off + n_tok,
              DKD_E_SHAPE, "%s: bad token geometry (n_tok <= %d)", fn, kMaxTok);
  DKD_REQUIRE(Ds == 192 && Dt == 384, DKD_E_SHAPE, "%s: built for widths 192 -> 384, got %d -> %d", fn, Ds, Dt);
  DKD_REQUIRE(s && t && W && loss && workspace, DKD_E_SHAPE, "%s: null pointer", fn);
  DKD_REQUIRE((((uintptr_t)workspace) & 1023) == 0, DKD_E_ALIGN, "%s: workspace must be 1024-byte aligned", fn);
  const int P = precision == DKD_PREC_BF16X3 ? 2 : 1;
  const int64_t M = B * n_tok;
  DKD_REQUIRE(M < (1ll << 31) - 256 && B * (Dt / kCh) < (1ll << 31), DKD_E_SHAPE, "%s: too many rows", fn);
  Workspace ws = carve(workspace, B, M, Ds, Dt, P);
  DKD_REQUIRE(workspace_bytes >= ws.bytes, DKD_E_WORKSPACE, "%s: workspace %zu < %zu", fn, workspace_bytes, ws.bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool want_grads = g_s || g_W || g_b;
  const int nterms = P == 2 ? 3 : 1;

  rc = launch_tokens_to_planes(s, dtype, B, Ts, s_off, n_tok, Ds, P, nullptr, ws.S, st);
  if (rc != DKD_OK) return rc;
  rc = launch_weight_to_planes(W, Dt, Ds, P, ws.Wp, want_grads ? ws.Wt : nullptr, st);
  if (rc != DKD_OK) return rc;

  {  // a = S W^T + b, fp32 rows
    using Cfg = FwdCfg;
    using L = PlaneLoader<Cfg>;
    using E = StoreRowsEpi<Cfg>;
    GemmParams<L, E> p;
    rc = make_plane_tmap(&p.ld.tmA, ws.S, P, M, Ds, Ds, M * Ds, Cfg::BM, "wass S");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmB, ws.Wp, P, Dt, Ds, Ds, (int64_t)Dt * Ds, Cfg::BN, "wass W");
    if (rc != DKD_OK) return rc;
    p.ld.k_blocks = Ds / 64; p.ld.nterms = nterms;
    p.ep.out = ws.A; p.ep.drop_mask = nullptr; p.ep.bias = bias; p.ep.alpha = 1.f;
    p.ep.M = M; p.ep.N_total = Dt; p.ep.n_tok = (int)M; p.ep.T_out = (int)M; p.ep.off = 0; p.ep.out_is_bf16 = 0;
    p.m_tiles = (int)((M + Cfg::BM - 1) / Cfg::BM); p.n_tiles = Dt / Cfg::BN;
    const int grid = min(kNumSMs, p.m_tiles * p.n_tiles);
    auto kern = gemm_tn_kernel<Cfg, L, E>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_wass_l1_fwdbwd: align GEMM");
    if (rc != DKD_OK) return rc;
  }
  const int sort_grid = (int)(B * (Dt / kCh));
  {
    SortParams sp;
    sp.a = ws.A; sp.t = t; sp.G = ws.G; sp.partials = ws.partials; sp.M = M; sp.N = Dt; sp.n_tok = n_tok; sp.Tt = Tt;
    sp.t_off = t_off; sp.planes = P; sp.t_is_bf16 = dtype == DKD_BF16; sp.write_grad = want_grads;
    const size_t sort_smem = (size_t)2 * n_tok * (kCh + 1) * sizeof(float);
    cudaFuncSetAttribute(wass_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem);
    wass_sort_kernel<<<sort_grid, kSortThreads, sort_smem, st>>>(sp);
    rc = check_launch("dkd_wass_l1_fwdbwd: sort");
    if (rc != DKD_OK) return rc;
    rc = launch_fold_partials(ws.partials, sort_grid, scale, loss, st);
    if (rc != DKD_OK) return rc;
  }
  if (!want_grads) return DKD_OK;

  if (g_s) {  // g_s = scale * G W
    using Cfg = DgradCfg;
    using L = PlaneLoader<Cfg>;
    using E = StoreRowsEpi<Cfg>;
    GemmParams<L, E> p;
    rc = make_plane_tmap(&p.ld.tmA, ws.G, P, M, Dt, Dt, M * Dt, Cfg::BM, "wass G");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmB, ws.Wt, P, Ds, Dt, Dt, (int64_t)Dt * Ds, Cfg::BN, "wass W^T");
    if (rc != DKD_OK) return rc;
    p.ld.k_blocks = Dt / 64; p.ld.nterms = nterms;
    p.ep.out = g_s; p.ep.drop_mask = nullptr; p.ep.bias = nullptr; p.ep.alpha = scale;
    p.ep.M = M; p.ep.N_total = Ds; p.ep.n_tok = n_tok; p.ep.T_out = Ts; p.ep.off = s_off; p.ep.out_is_bf16 = dtype == DKD_BF16;
    p.m_tiles = (int)((M + Cfg::BM - 1) / Cfg::BM); p.n_tiles = Ds / Cfg::BN;
    const int grid = min(kNumSMs, p.m_tiles * p.n_tiles);
    auto kern = gemm_tn_kernel<Cfg, L, E>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_wass_l1_fwdbwd: dgrad GEMM");
    if (rc != DKD_OK) return rc;
  }
  if (g_W || g_b) {
    using Cfg = WgradCfg;
    using L = NtPlainLoader<Cfg>;
    DKD_REQUIRE(g_W != nullptr, DKD_E_UNSUPPORTED, "%s: g_b without g_W is not supported", fn);
    GemmNtParamsT<Cfg, L> p;
    rc = make_plane_tmap(&p.ld.tmA, ws.G, P, M, Dt, Dt, M * Dt, Cfg::KROWS, "wass G^T");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmB, ws.S, P, M, Ds, Ds, M * Ds, Cfg::KROWS, "wass S (wgrad)");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmOnes, ws.ones, 2, 64, 64, 64, 64 * 64, Cfg::KROWS, "ones tile");
    if (rc != DKD_OK) return rc;
    rc = launch_fill_ones_tile(ws.ones, st);
    if (rc != DKD_OK) return rc;
    cudaMemsetAsync(g_W, 0, (size_t)Dt * Ds * sizeof(float), st);
    if (g_b) cudaMemsetAsync(g_b, 0, (size_t)Dt * sizeof(float), st);
    p.ep.D = g_W; p.ep.Dcol = g_b; p.ep.ldd = Ds; p.ep.alpha = scale;
    p.ld.ldd = Ds; p.ld.na_tiles = Dt / 128; p.ld.b_col0 = 0;
    p.ld.total_row_blocks = (int)((M + Cfg::KROWS - 1) / Cfg::KROWS);
    nt_make_splits(p.ld.total_row_blocks, kNumSMs / p.ld.na_tiles, &p.ld.splits, &p.ld.row_blocks_per_split);
    p.nterms = nterms;
    const int grid = min(kNumSMs, p.ld.na_tiles * p.ld.splits);
    auto kern = gemm_nt_kernel<Cfg, L>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_wass_l1_fwdbwd: wgrad GEMM");
    if (rc != DKD_OK) return rc;
  }
  return DKD_OK;
}

}  // extern "C"
