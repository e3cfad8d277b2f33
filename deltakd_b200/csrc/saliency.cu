// Saliency scores for saliency-MGD mask selection (no gradient flows through them).
// Reference: saliency_masking (model/misc.py:38-165) with SimpleAttention / SimpleCrossAttention
// (model/models.py:14-56).  The score's ascending order picks the kept tokens (dkd_mask_rank).
//
//   method 1 (misc.py:62-70, models.py:46-56): self-attention over the patch tokens, 8 heads x 48;
//       score_i = mean_h softmax_j(q_i . k_j * 48^-1/2)[i]  — only the DIAGONAL of each head's map is used,
//       so the maps are never materialised: the qk projection runs on tcgen05 (gemm_tn, fp32 rows to scratch)
//       and one CTA per sample walks the heads with K_h in shared memory, one query row per thread, online
//       softmax in registers.
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
struct DiagParams {
  const float* qk;   // [B*n_tok, 2*D] fp32: q = cols [0, D), k = cols [D, 2D)
  float* score;      // [B, n_tok]
  int n_tok, D, H;
  float scale;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256) selfdiag_score_kernel(DiagParams p) {
  extern __shared__ float sm[];
  float* ks = sm;  // [n_tok][48]
  const int b = blockIdx.x, i = threadIdx.x;
  const int ld = 2 * p.D;
  const float* base = p.qk + (size_t)b * p.n_tok * ld;
  const bool live = i < p.n_tok;
  float total = 0.f;
  for (int h = 0; h < p.H; ++h) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < p.n_tok * (kHeadDim / 4); idx += blockDim.x) {
      const int j = idx / (kHeadDim / 4), c = (idx - j * (kHeadDim / 4)) * 4;
      *reinterpret_cast<float4*>(ks + j * kHeadDim + c) =
          *reinterpret_cast<const float4*>(base + (size_t)j * ld + p.D + h * kHeadDim + c);
    }
    __syncthreads();
    if (!live) continue;
    float q[kHeadDim];
#pragma unroll
    for (int c = 0; c < kHeadDim; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(base + (size_t)i * ld + h * kHeadDim + c);
      const float sc = p.scale * 1.4426950408889634f;   // logits in base 2
      q[c] = v.x * sc; q[c + 1] = v.y * sc; q[c + 2] = v.z * sc; q[c + 3] = v.w * sc;
    }
    // online softmax in base 2 over groups of 4 keys: one rescale per group (branch-free), MUFU ex2 (relative error 2^-22,
    // far inside the 1e-5 score gate); the logits carry log2(e) through the pre-scaled query
    float m = -INFINITY, l = 0.f, sii = 0.f;
    auto dot = [&](int j) {
      const float4* kr = reinterpret_cast<const float4*>(ks + j * kHeadDim);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int c = 0; c < kHeadDim / 4; ++c) {
        const float4 kv = kr[c];
        a0 = fmaf(q[4 * c], kv.x, a0); a1 = fmaf(q[4 * c + 1], kv.y, a1);
        a2 = fmaf(q[4 * c + 2], kv.z, a2); a3 = fmaf(q[4 * c + 3], kv.w, a3);
      }
      return (a0 + a1) + (a2 + a3);
    };
    int j = 0;
    for (; j + 4 <= p.n_tok; j += 4) {
      const float s0 = dot(j), s1 = dot(j + 1), s2 = dot(j + 2), s3 = dot(j + 3);
      sii = i == j ? s0 : i == j + 1 ? s1 : i == j + 2 ? s2 : i == j + 3 ? s3 : sii;
      const float mn = fmaxf(fmaxf(m, fmaxf(s0, s1)), fmaxf(s2, s3));
      l = l * ex2(m - mn) + ((ex2(s0 - mn) + ex2(s1 - mn)) + (ex2(s2 - mn) + ex2(s3 - mn)));
      m = mn;
    }
    for (; j < p.n_tok; ++j) {
      const float s0 = dot(j);
      sii = i == j ? s0 : sii;
      const float mn = fmaxf(m, s0);
      l = l * ex2(m - mn) + ex2(s0 - mn);
      m = mn;
    }
    total += ex2(sii - m) / l;
  }
  if (live) p.score[(size_t)b * p.n_tok + i] = total / (float)p.H;
}

using QkCfg = GemmCfg<192, 1, 4, 2>;

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Workspace {
  __nv_bfloat16 *X, *Wp;
  float* QK;
  size_t bytes;
};
Workspace carve(void* base, int64_t M, int D, int P) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 1024); return reinterpret_cast<char*>(base) + o; };
  w.X = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * M * D * 2));
  w.Wp = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * 2 * D * D * 2));
  w.QK = reinterpret_cast<float*>(take((size_t)M * 2 * D * 4));
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
  return dkd::carve(nullptr, B * n_tok, D, precision == DKD_PREC_BF16X3 ? 2 : 1).bytes;
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
  DKD_REQUIRE(B >= 0 && n_tok > 0 && n_tok <= kMaxKeys && off >= 0 && T >= off + n_tok, DKD_E_SHAPE, "%s: bad token geometry (n_tok <= %d)", fn,
              kMaxKeys);
  DKD_REQUIRE(num_heads == 8 && D == num_heads * kHeadDim, DKD_E_SHAPE, "%s: built for 8 heads x 48 (D = 384), got %d heads, D = %d", fn,
              num_heads, D);
  if (B == 0) return DKD_OK;
  DKD_REQUIRE(x && qk_w && score && workspace, DKD_E_SHAPE, "%s: null pointer", fn);
  DKD_REQUIRE((((uintptr_t)workspace) & 1023) == 0, DKD_E_ALIGN, "%s: workspace must be 1024-byte aligned", fn);
  const int P = precision == DKD_PREC_BF16X3 ? 2 : 1;
  const int64_t M = B * n_tok;
  DKD_REQUIRE(M < (1ll << 31) - 256, DKD_E_SHAPE, "%s: too many rows", fn);
  Workspace ws = carve(workspace, M, D, P);
  DKD_REQUIRE(workspace_bytes >= ws.bytes, DKD_E_WORKSPACE, "%s: workspace %zu < %zu", fn, workspace_bytes, ws.bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  rc = launch_tokens_to_planes(x, dtype, B, T, off, n_tok, D, P, nullptr, ws.X, st);
  if (rc != DKD_OK) return rc;
  rc = launch_weight_to_planes(qk_w, 2 * D, D, P, ws.Wp, nullptr, st);
  if (rc != DKD_OK) return rc;
  {  // qk = X W^T + b  (fp32 rows)
    using Cfg = QkCfg;
    using L = PlaneLoader<Cfg>;
    using E = StoreRowsEpi<Cfg>;
    GemmParams<L, E> p;
    rc = make_plane_tmap(&p.ld.tmA, ws.X, P, M, D, D, M * D, Cfg::BM, "saliency X");
    if (rc != DKD_OK) return rc;
    rc = make_plane_tmap(&p.ld.tmB, ws.Wp, P, 2 * D, D, D, (int64_t)2 * D * D, Cfg::BN, "saliency qk weight");
    if (rc != DKD_OK) return rc;
    p.ld.k_blocks = D / 64; p.ld.nterms = P == 2 ? 3 : 1;
    p.ep.out = ws.QK; p.ep.drop_mask = nullptr; p.ep.bias = qk_b; p.ep.alpha = 1.f;
    p.ep.M = M; p.ep.N_total = 2 * D; p.ep.n_tok = (int)M; p.ep.T_out = (int)M; p.ep.off = 0; p.ep.out_is_bf16 = 0;
    p.m_tiles = (int)((M + Cfg::BM - 1) / Cfg::BM); p.n_tiles = 2 * D / Cfg::BN;
    const int grid = min(kNumSMs, p.m_tiles * p.n_tiles);
    auto kern = gemm_tn_kernel<Cfg, L, E>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_saliency_selfdiag_score: qk GEMM");
    if (rc != DKD_OK) return rc;
  }
  DiagParams dp;
  dp.qk = ws.QK; dp.score = score; dp.n_tok = n_tok; dp.D = D; dp.H = num_heads; dp.scale = 1.0f / sqrtf((float)kHeadDim);
  const size_t smem = (size_t)n_tok * kHeadDim * sizeof(float);
  const int threads = (n_tok + 31) / 32 * 32;
  cudaFuncSetAttribute(selfdiag_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  selfdiag_score_kernel<<<(unsigned)B, threads, smem, st>>>(dp);
  return check_launch(fn);
}

}  // extern "C"
