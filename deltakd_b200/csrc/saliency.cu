// Saliency scores for saliency-MGD mask selection (no gradient flows through them).
// Reference: saliency_masking (model/misc.py:38-165) with SimpleAttention / SimpleCrossAttention
// (model/models.py:14-56).  The score's ascending order picks the kept tokens (dkd_mask_rank).
//
//   method 1 (misc.py:62-70, models.py:46-56): self-attention over the patch tokens, 8 heads x 48;
//       score_i = mean_h softmax_j(q_i . k_j * 48^-1/2)[i]  — only the DIAGONAL of each head's map is used,
//       so the maps are never materialised: the qk projection runs on tcgen05 and writes q, k as bf16 planes; a second
//       tcgen05 GEMM per (sample, head, query half) leaves the 128 x 196 logits in tensor memory, where the epilogue
//       reduces every row to its softmax diagonal (online softmax in registers, MUFU ex2).
//   method 2 (misc.py:88-116): query = CLS token only, keys = [CLS] + patches; score = head-mean of that row,
//       patches only.      method 3 (misc.py:135-148, models.py:24-35): CLS -> patches cross attention.
//       With one query per sample the key projection is folded into the query:
//           q_h . (W_k,h x_j + b_k,h) = (W_k,h^T q_h) . x_j + q_h . b_k,h
//       so each sample costs two 384x384 mat-vecs and one pass over its tokens (HBM-bound: the teacher tokens
//       are read once), instead of a [B*197, 384] x [384, 384] key GEMM.
#include "epilogues.cuh"
#include "planes.cuh"

namespace dkd {
namespace {

constexpr int kHeadDim = 48;
constexpr int kMaxKeys = 256;

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return Elt<T>::ld(p); }

// ---------------------------------------------------------------- methods 2 / 3: one query per sample
struct ClsParams {
  const void* xq; int64_t xq_stride;   // query token of sample b at xq + b*xq_stride (elements)
  const void* xk; int64_t xk_stride;   // key token j of sample b at xk + b*xk_stride + j*D
  const float *Wq, *bq, *Wk, *bk;      // [D, D], [D]
  float* score;                        // [B, n_keys]
  int n_keys, D, H, query_is_key;
  float scale;
};

template <typename T>
__global__ void __launch_bounds__(256) cls_score_kernel(ClsParams p) {
  extern __shared__ float sm[];
  const int D = p.D, H = p.H;
  float* xq = sm;                 // [D]
  float* q = xq + D;              // [D]
  float* u = q + D;               // [H][D]
  float* cst = u + H * D;         // [H]
  float* logit = cst + 8;         // [H][n_keys + 1]
  float* hsum = logit + H * (p.n_keys + 1);  // [H] max, [H] 1/sum
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nk = p.n_keys + (p.query_is_key ? 1 : 0);
  const T* xqp = reinterpret_cast<const T*>(p.xq) + b * p.xq_stride;
  const T* xkp = reinterpret_cast<const T*>(p.xk) + b * p.xk_stride;

  for (int k = tid; k < D; k += 256) xq[k] = ldf(xqp + k);
  __syncthreads();
  // q = Wq xq + bq : one warp per output row, lanes stride the contraction (coalesced weight rows)
  for (int c = warp; c < D; c += 8) {
    const float* w = p.Wq + (size_t)c * D;
    float a = 0.f;
    for (int k = lane; k < D; k += 32) a = fmaf(__ldg(w + k), xq[k], a);
    a = warp_sum(a);
    if (lane == 0) q[c] = a + (p.bq ? __ldg(p.bq + c) : 0.f);
  }
  __syncthreads();
  // u[h][k] = scale * sum_d Wk[h*hd + d][k] * q[h*hd + d] ;  cst[h] = scale * q_h . bk_h
  for (int idx = tid; idx < H * D; idx += 256) {
    const int h = idx / D, k = idx - h * D;
    float a = 0.f;
#pragma unroll 4
    for (int d = 0; d < kHeadDim; ++d) a = fmaf(__ldg(p.Wk + (size_t)(h * kHeadDim + d) * D + k), q[h * kHeadDim + d], a);
    u[idx] = a * p.scale;
  }
  if (tid < H) {
    float a = 0.f;
    for (int d = 0; d < kHeadDim; ++d) a = fmaf(q[tid * kHeadDim + d], p.bk ? __ldg(p.bk + tid * kHeadDim + d) : 0.f, a);
    cst[tid] = a * p.scale;
  }
  __syncthreads();
  // logits: one warp per key token (key 0 = the query token itself when query_is_key)
  for (int j = warp; j < nk; j += 8) {
    const T* x = p.query_is_key ? (j == 0 ? xqp : xkp + (size_t)(j - 1) * D) : xkp + (size_t)j * D;
    float acc[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) acc[h] = 0.f;
    for (int k = lane; k < D; k += 32) {
      const float xv = ldf(x + k);
#pragma unroll
      for (int h = 0; h < 8; ++h) acc[h] = fmaf(u[h * D + k], xv, acc[h]);
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) acc[h] = warp_sum(acc[h]);
    if (lane < 8) {
      float v = acc[0];
#pragma unroll
      for (int h = 1; h < 8; ++h) v = lane == h ? acc[h] : v;
      logit[lane * (p.n_keys + 1) + j] = v + cst[lane];
    }
  }
  __syncthreads();
  // per-head softmax statistics: warp h
  if (warp < H) {
    const float* l = logit + warp * (p.n_keys + 1);
    float m = -INFINITY;
    for (int j = lane; j < nk; j += 32) m = fmaxf(m, l[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < nk; j += 32) s += expf(l[j] - m);
    s = warp_sum(s);
    if (lane == 0) { hsum[warp] = m; hsum[H + warp] = s; }
  }
  __syncthreads();
  const int skip = p.query_is_key ? 1 : 0;
  for (int i = tid; i < p.n_keys; i += 256) {
    float a = 0.f;
    for (int h = 0; h < H; ++h) a += expf(logit[h * (p.n_keys + 1) + i + skip] - hsum[h]) / hsum[H + h];
    p.score[(size_t)b * p.n_keys + i] = a / (float)H;
  }
}

// ---------------------------------------------------------------- method 1: diagonal of the self-attention maps
// Two tcgen05 GEMMs, no map materialised:
//   1. qk = X W^T + b, written by the epilogue as THREE bf16 planes (hi, mid, lo = all 24 mantissa bits) of [M, 2D];
//      the q half carries 48^-1/2 log2(e), so the logits below are in base 2
//   2. per (sample, head): S = q_h k_h^T on the tensor core — M tile 128 query rows, N = 208 key rows, K = 48: the
//      64-channel box that starts at the head's first channel is loaded and the MMAs stop after 3 K steps; six plane
//      products (fp32-exact operands: the scores rank tokens, a 1e-5 logit error flips masks).  The epilogue keeps the
//      accumulator row in registers chunk by chunk: online softmax over the 196 real keys and the diagonal entry,
//      p_ii = 2^(s_ii - m) / l, stored per head; a last small kernel averages the heads in a fixed order.
// The SIMT version of step 2 (one CTA per sample, K_h in shared memory, one query row per thread) took 954 us of the
// 5.85 ms saliency-MGD call at B = 512: 15 GFLOP of fp32 FMAs fed from shared memory.
constexpr int kQkPlanes = 3;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

using QkCfg = GemmCfg<192, 1, 4, 2>;

struct QkPlanesParams {
  __nv_bfloat16* planes;   // [3][M][2D]
  const float* bias;       // [2D] or null
  int64_t M;
  int D;
  float qscale;            // applied to the q half (columns < D)
};
struct QkPlanesEpi {
  using Params = QkPlanesParams;
  struct State {};
  static __device__ __forceinline__ void init(const Params&, State&) {}
  static __device__ __forceinline__ void tile(const Params& p, State&, int m0, int n0, int row_in_tile, uint32_t t_acc) {
    const int64_t m = (int64_t)m0 + row_in_tile;
    const bool live = m < p.M;
    const float sc = n0 < p.D ? p.qscale : 1.f;      // BN = 192 divides D = 384: a tile is all q or all k
    const int64_t plane = p.M * 2 * p.D;
#pragma unroll 1
    for (int c0 = 0; c0 < QkCfg::BN; c0 += 32) {
      float v[32];
      sm100::tmem_ld32(t_acc + c0, v);
      sm100::tmem_ld_wait();
      if (!live) continue;
      float bs[32];
      ldg_vec32(p.bias ? p.bias + n0 + c0 : nullptr, bs);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (v[j] + bs[j]) * sc;
      __nv_bfloat16* op = p.planes + m * 2 * p.D + n0 + c0;
#pragma unroll
      for (int pl = 0; pl < kQkPlanes; ++pl) {
        stg256(op + pl * plane, *reinterpret_cast<float(*)[16]>(&v[0]));
        stg256(op + pl * plane + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
        if (pl + 1 < kQkPlanes) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] -= __bfloat162float(__float2bfloat16_rn(v[j]));
        }
      }
    }
  }
  static __device__ __forceinline__ void finish(const Params&, State&, int) {}
};

// tile index mt = ((b * H + h) * 2 + half): query rows [128 half, 128 half + 128) of sample b, head h; one N tile
using ScoreCfg = GemmCfg<208, 1, 4, 2>;
struct ScoreLoaderParams {
  CUtensorMap tmQ, tmK;   // 4-D {2D, n_tok, B, planes}; Q: box {64, 128, 1, 1}, K: box {64, 208, 1, 1}
  int H, D;
};
struct ScoreLoader {
  using Params = ScoreLoaderParams;
  static constexpr int K_STEPS = kHeadDim / 16;   // 3 of the 4 K steps of a 64-channel stage
  static constexpr uint32_t TX_BYTES = ScoreCfg::STAGE_BYTES;
  static __device__ __forceinline__ int num_k_iters(const Params&) { return 6; }
  static __device__ __forceinline__ void prefetch(const Params& p) { sm100::tma_prefetch_desc(&p.tmQ); sm100::tma_prefetch_desc(&p.tmK); }
  static __device__ __forceinline__ void issue(const Params& p, int kit, int mt, int, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    int pa, pb;
    term_planes(kit, 6, pa, pb);
    const int half = mt & 1, bh = mt >> 1, h = bh % p.H, b = bh / p.H;
    sm100::tma_load_4d(sA, &p.tmQ, bar, h * kHeadDim, half * 128, b, pa);
    sm100::tma_load_4d(sB, &p.tmK, bar, p.D + h * kHeadDim, 0, b, pb);
  }
};
struct DiagEpiParams {
  float* head_score;   // [B][H][n_tok]
  int n_tok, H;
};
struct DiagEpi {
  using Params = DiagEpiParams;
  struct State {};
  static __device__ __forceinline__ void init(const Params&, State&) {}
  static __device__ __forceinline__ void tile(const Params& p, State&, int m0, int, int row_in_tile, uint32_t t_acc) {
    const int mt = m0 >> 7;
    const int half = mt & 1, bh = mt >> 1;
    const int i = half * 128 + row_in_tile;         // query token
    float m = -INFINITY, l = 0.f, sii = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 192; c0 += 32) {
      float v[32];
      sm100::tmem_ld32(t_acc + c0, v);
      sm100::tmem_ld_wait();
      float c[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
      for (int j = 4; j < 32; j += 4) { c[0] = fmaxf(c[0], v[j]); c[1] = fmaxf(c[1], v[j + 1]); c[2] = fmaxf(c[2], v[j + 2]); c[3] = fmaxf(c[3], v[j + 3]); }
      const float mn = fmaxf(m, fmaxf(fmaxf(c[0], c[1]), fmaxf(c[2], c[3])));
      float s0 = l * ex2(m - mn), s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        s0 += ex2(v[j] - mn); s1 += ex2(v[j + 1] - mn); s2 += ex2(v[j + 2] - mn); s3 += ex2(v[j + 3] - mn);
      }
      l = (s0 + s1) + (s2 + s3); m = mn;
      if ((i >> 5) == (c0 >> 5)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) sii = (i & 31) == j ? v[j] : sii;
      }
    }
    {
      float v[32];
      sm100::tmem_ld32(t_acc + 176, v);   // columns 176..207: keys 192..195 are v[16..19]
      sm100::tmem_ld_wait();
      const float mn = fmaxf(m, fmaxf(fmaxf(v[16], v[17]), fmaxf(v[18], v[19])));
      l = l * ex2(m - mn) + ((ex2(v[16] - mn) + ex2(v[17] - mn)) + (ex2(v[18] - mn) + ex2(v[19] - mn)));
      m = mn;
      if (i >= 192) sii = i == 192 ? v[16] : i == 193 ? v[17] : i == 194 ? v[18] : v[19];
    }
    if (i < p.n_tok) p.head_score[(size_t)bh * p.n_tok + i] = ex2(sii - m) / l;
  }
  static __device__ __forceinline__ void finish(const Params&, State&, int) {}
};

__global__ void __launch_bounds__(256) head_mean_kernel(const float* __restrict__ hs, float* __restrict__ score, int64_t total, int n_tok, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t b = idx / n_tok, i = idx - b * n_tok;
  float t = 0.f;
  for (int h = 0; h < H; ++h) t += hs[((size_t)b * H + h) * n_tok + i];
  score[idx] = t / (float)H;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Workspace {
  __nv_bfloat16 *X, *Wp, *QK;
  float* HS;
  size_t bytes;
};
Workspace carve(void* base, int64_t B, int n_tok, int D, int P, int H) {
  Workspace w;
  const int64_t M = B * n_tok;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 1024); return reinterpret_cast<char*>(base) + o; };
  w.X = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * M * D * 2));
  w.Wp = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * 2 * D * D * 2));
  w.QK = reinterpret_cast<__nv_bfloat16*>(take((size_t)kQkPlanes * M * 2 * D * 2));
  w.HS = reinterpret_cast<float*>(take((size_t)B * H * n_tok * 4));
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace dkd

extern "C" {

int dkd_saliency_cls_score(const void* xq, int64_t xq_stride, const void* xk, int64_t xk_stride, int n_keys, int64_t B, int D,
                           int dtype, const float* Wq, const float* bq, const float* Wk, const float* bk, int num_heads,
                           int query_is_key, float* score, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  const char* fn = "dkd_saliency_cls_score";
  DKD_REQUIRE(dtype == DKD_F32 || dtype == DKD_BF16, DKD_E_DTYPE, "%s: dtype %d", fn, dtype);
  DKD_REQUIRE(B >= 0 && B < (1ll << 31) && n_keys > 0 && n_keys < kMaxKeys, DKD_E_SHAPE, "%s: need 0 < n_keys < %d", fn, kMaxKeys);
  DKD_REQUIRE(num_heads == 8 && D == num_heads * kHeadDim, DKD_E_SHAPE, "%s: built for 8 heads x 48 (D = 384), got %d heads, D = %d", fn,
              num_heads, D);
  if (B == 0) return DKD_OK;
  DKD_REQUIRE(xq && xk && Wq && Wk && score, DKD_E_SHAPE, "%s: null pointer", fn);
  ClsParams p;
  p.xq = xq; p.xq_stride = xq_stride; p.xk = xk; p.xk_stride = xk_stride; p.Wq = Wq; p.bq = bq; p.Wk = Wk; p.bk = bk;
  p.score = score; p.n_keys = n_keys; p.D = D; p.H = num_heads; p.query_is_key = query_is_key ? 1 : 0;
  p.scale = 1.0f / sqrtf((float)kHeadDim);
  const size_t smem = (size_t)(2 * D + num_heads * D + 8 + num_heads * (n_keys + 1) + 2 * num_heads) * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DKD_F32) {
    cudaFuncSetAttribute(cls_score_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cls_score_kernel<float><<<(unsigned)B, 256, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(cls_score_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cls_score_kernel<__nv_bfloat16><<<(unsigned)B, 256, smem, st>>>(p);
  }
  return check_launch(fn);
}

size_t dkd_saliency_selfdiag_workspace_bytes(int64_t B, int n_tok, int D, int precision) {
  return dkd::carve(nullptr, B, n_tok, D, precision == DKD_PREC_BF16X3 ? 2 : 1, 8).bytes;
}

int dkd_saliency_selfdiag_score(const void* x, int64_t B, int T, int off, int n_tok, int D, int dtype, const float* qk_w,
                                const float* qk_b, int num_heads, int precision, float* score, void* workspace,
                                size_t workspace_bytes, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  const char* fn = "dkd_saliency_selfdiag_score";
  DKD_REQUIRE(dtype == DKD_F32 || dtype == DKD_BF16, DKD_E_DTYPE, "%s: dtype %d", fn, dtype);
  DKD_REQUIRE(precision == DKD_PREC_BF16 || precision == DKD_PREC_BF16X3, DKD_E_UNSUPPORTED, "%s: precision %d", fn, precision);
  DKD_REQUIRE(B >= 0 && n_tok == 196 && off >= 0 && T >= off + n_tok, DKD_E_SHAPE,
              "%s: built for 196 patch tokens per sample (two 128-row query tiles, 208-column key tile), got %d", fn, n_tok);
  DKD_REQUIRE(num_heads == 8 && D == num_heads * kHeadDim, DKD_E_SHAPE, "%s: built for 8 heads x 48 (D = 384), got %d heads, D = %d", fn,
              num_heads, D);
  if (B == 0) return DKD_OK;
  DKD_REQUIRE(x && qk_w && score && workspace, DKD_E_SHAPE, "%s: null pointer", fn);
  DKD_REQUIRE((((uintptr_t)workspace) & 1023) == 0, DKD_E_ALIGN, "%s: workspace must be 1024-byte aligned", fn);
  const int P = precision == DKD_PREC_BF16X3 ? 2 : 1;
  const int64_t M = B * n_tok;
  DKD_REQUIRE(M < (1ll << 31) - 256, DKD_E_SHAPE, "%s: too many rows", fn);
  DKD_REQUIRE(B * num_heads * 2 < (1ll << 31), DKD_E_SHAPE, "%s: too many (sample, head) tiles", fn);
  Workspace ws = carve(workspace, B, n_tok, D, P, num_heads);
  DKD_REQUIRE(workspace_bytes >= ws.bytes, DKD_E_WORKSPACE, "%s: workspace %zu < %zu", fn, workspace_bytes, ws.bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  rc = launch_tokens_to_planes(x, dtype, B, T, off, n_tok, D, P, nullptr, ws.X, st);
  if (rc != DKD_OK) return rc;
  rc = launch_weight_to_planes(qk_w, 2 * D, D, P, ws.Wp, nullptr, st);
  if (rc != DKD_OK) return rc;
  {  // qk = X W^T + b  as three bf16 planes, q pre-scaled
    using Cfg = QkCfg;
    using L = PlaneLoader<Cfg>;
    using E = QkPlanesEpi;
    GemmParams<L, E> p;
    rc = make_plane_tmap(&p.ld.tmA, ws.X, P, M, D, D, M * D, Cfg::BM, "saliency X");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmB, ws.Wp, P, 2 * D, D, D, (int64_t)2 * D * D, Cfg::BN, "saliency qk weight");
    if (rc != DKD_OK) return rc;
    p.ld.k_blocks = D / 64; p.ld.nterms = P == 2 ? 3 : 1;
    p.ep.planes = ws.QK; p.ep.bias = qk_b; p.ep.M = M; p.ep.D = D;
    p.ep.qscale = 1.4426950408889634f / sqrtf((float)kHeadDim);
    p.m_tiles = (int)((M + Cfg::BM - 1) / Cfg::BM); p.n_tiles = 2 * D / Cfg::BN;
    const int grid = min(kNumSMs, p.m_tiles * p.n_tiles);
    auto kern = gemm_tn_kernel<Cfg, L, E>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_saliency_selfdiag_score: qk GEMM");
    if (rc != DKD_OK) return rc;
  }
  {  // per (sample, head): softmax diagonal of q_h k_h^T
    using Cfg = ScoreCfg;
    GemmParams<ScoreLoader, DiagEpi> p;
    const uint64_t dims[4] = {(uint64_t)(2 * D), (uint64_t)n_tok, (uint64_t)B, (uint64_t)kQkPlanes};
    const uint64_t strides[3] = {(uint64_t)2 * D * 2, (uint64_t)n_tok * 2 * D * 2, (uint64_t)M * 2 * D * 2};
    const uint32_t box_q[4] = {64, 128, 1, 1}, box_k[4] = {64, 208, 1, 1};
    rc = make_tmap_bf16(&p.ld.tmQ, ws.QK, 4, dims, strides, box_q, "saliency q planes");
    if (rc != DKD_OK) return rc;
    rc = make_tmap_bf16(&p.ld.tmK, ws.QK, 4, dims, strides, box_k, "saliency k planes");
    if (rc != DKD_OK) return rc;
    p.ld.H = num_heads; p.ld.D = D;
    p.ep.head_score = ws.HS; p.ep.n_tok = n_tok; p.ep.H = num_heads;
    p.m_tiles = (int)(B * num_heads * 2); p.n_tiles = 1;
    const int grid = min(kNumSMs, p.m_tiles);
    auto kern = gemm_tn_kernel<Cfg, ScoreLoader, DiagEpi>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_saliency_selfdiag_score: score GEMM");
    if (rc != DKD_OK) return rc;
  }
  head_mean_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(ws.HS, score, M, n_tok, num_heads);
  return check_launch(fn);
}

}  // extern "C"
