// LRKD: low-rank projection matching.
// Reference: lrkd branch (model/loss.py:80-103) + lrkd_loss (:314-330): per layer pair (student 0,1,-1 / teacher 0,1,11)
//     T = teacher[:, 2:].reshape(M, Dt) ;  U, S, _ = svd(T) ;  A = U[:, :k] diag(S[:k])       ( = T V_k )
//     s' = Linear(Ds -> k)(student[:, 1:]).reshape(M, k) ;  loss += coef * mean((A - s')^2)
// and its autograd backward (to the student feature and the Linear).
//
// The tall SVD is replaced by the eigen-decomposition of the Dt x Dt Gram matrix (A = T V_k needs only V_k):
//   1. teacher -> bf16 planes (hi, mid, lo = all 24 mantissa bits) + exact fp64 column sums of squares
//   2. Gram  G = T^T T  on tcgen05 (gemm_nt, split over rows, 6 plane products = fp32-exact operands); each split
//      stores its fp32 partial tile, a second kernel adds the partials in fp64 in a fixed order (deterministic),
//      symmetrises and puts the exact diagonal in
//   3. eigenvectors of G: one-sided (Hestenes) Jacobi in fp64, all layers batched in ONE cooperative launch, two-level
//      ordering: 12 CTAs per matrix each sweep a pair of 16-column groups in shared memory (31 inner rounds), one grid
//      barrier per group round (23 per sweep); stops when every rotated pair is orthogonal to 1e-10
//   4. select: column norms = eigenvalues, rank them (descending, ties by index), V_k = top-k normalised columns;
//      builds the fused operand [W' | -V_k] (bf16 hi/lo planes) and the transposed head for the backward
//   5. d = [s | T] [W' | -V_k]^T + b' = s' - A  in ONE tcgen05 GEMM (K = Ds + Dt), epilogue: loss partial and
//      g_s' = 2 coef/(M k) d as planes ;  6. g_s = g_s' W' ;  7. g_W' = g_s'^T s, g_b' = g_s'^T 1
// Column signs of singular vectors are arbitrary (LAPACK's differ between fp32 and fp64): V_k and S_k
// can be returned so that callers / tests align signs; here each v is normalised to a positive largest component.
#include <cooperative_groups.h>

#include "epilogues.cuh"
#include "gemm_nt.cuh"
#include "planes.cuh"

namespace dkd {
namespace {

constexpr int kN = 384;              // Dt: order of the Gram matrix
// Columns per group G: a CTA holds 2G columns and runs one warp per column pair (G warps).  G = 8 (24 CTAs per matrix)
// is the default — with fewer pairs per SM the fp64 pipe of each SM is less contended and a round is ~20 % shorter
// than with G = 16 (G = 4 measured slower again: 95 grid barriers per sweep); G = 16 (12 CTAs per matrix) is used when
// 24 * n_layers CTAs would not be co-resident.
template <int G> struct JacobiCfg {
  static constexpr int CTAS = kN / G / 2;
  static constexpr int THREADS = 32 * G;
  static constexpr size_t SMEM = (size_t)2 * G * kN * sizeof(double);
};
constexpr int kMaxSweeps = 14;
constexpr double kJacobiTol = 1e-10;   // pairs with |cos| below this are not rotated
// Jacobi converges quadratically: a sweep that SAW no |cos| above 1e-6 leaves the columns orthogonal to ~1e-12, so it
// is the last one (no separate verification sweep).
constexpr double kJacobiStop = 1e-6;
constexpr int kMaxLayers = 8;

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- 1. teacher planes + exact column sums of squares
template <typename T>
__global__ void __launch_bounds__(192) teacher_planes_kernel(const T* __restrict__ src, int64_t B, int T_tok, int off, int n_tok,
                                                             int P, __nv_bfloat16* __restrict__ dst, double* __restrict__ sumsq_part) {
  // 48 threads cover one 384-wide row (8 channels each); 4 rows per block iteration; a thread keeps its channel group
  const int cg = (threadIdx.x % 48) * 8, rl = threadIdx.x / 48;
  const int64_t M = B * n_tok;
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0;
  for (int64_t m = (int64_t)blockIdx.x * 4 + rl; m < M; m += (int64_t)gridDim.x * 4) {
    const int64_t b = m / n_tok;
    const int64_t srow = b * T_tok + off + (m - b * n_tok);
    float v[8];
    if constexpr (sizeof(T) == 4) {
      Vec<float, 4>::load(reinterpret_cast<const float*>(src) + srow * kN + cg, *reinterpret_cast<float(*)[4]>(&v[0]));
      Vec<float, 4>::load(reinterpret_cast<const float*>(src) + srow * kN + cg + 4, *reinterpret_cast<float(*)[4]>(&v[4]));
    } else {
      Vec<__nv_bfloat16, 8>::load(reinterpret_cast<const __nv_bfloat16*>(src) + srow * kN + cg, v);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fma((double)v[j], (double)v[j], acc[j]);
    float r[8];
    Vec<__nv_bfloat16, 8>::store(dst + m * kN + cg, v);
    if (P >= 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
      Vec<__nv_bfloat16, 8>::store(dst + M * kN + m * kN + cg, r);
    }
    if (P >= 3) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = r[j] - __bfloat162float(__float2bfloat16_rn(r[j]));
      Vec<__nv_bfloat16, 8>::store(dst + 2 * M * kN + m * kN + cg, r);
    }
  }
  __shared__ double red[4][kN];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < kN; c += 192)
    sumsq_part[(size_t)blockIdx.x * kN + c] = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
}

// ---------------------------------------------------------------- 2. Gram: per-split partial tiles -> fp64, symmetric
using GramCfg = GemmNtCfg<6, false, 256, 128, 3, 64>;   // 128 x 384 fp32 tile in TMEM, 64 rows per stage

struct NtGramLoader : NtPlainLoader<GramCfg> {
  static __device__ __forceinline__ void decode(const Params& p, int item, Item& it) {
    const int tile = item % p.na_tiles, split = item / p.na_tiles;
    it.rb0 = split * p.row_blocks_per_split;
    it.rb1 = min(it.rb0 + p.row_blocks_per_split, p.total_row_blocks);
    it.a_col0 = tile * 128;
    it.b_col0 = 0;
    it.d_off = ((int64_t)split * kN + tile * 128) * kN;   // partial[split][tile*128 ..][0 .. 384)
    it.dcol_off = -1;
    it.aux = 0;
  }
};

__global__ void gram_reduce_kernel(const float* __restrict__ part, int splits, double* __restrict__ gsum) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kN * kN) return;
  double a = 0.0;
  for (int s = 0; s < splits; ++s) a += (double)part[(size_t)s * kN * kN + idx];
  gsum[idx] = a;
}

__global__ void gram_symmetrize_kernel(const double* __restrict__ gsum, const double* __restrict__ sumsq_part, int nparts,
                                       double* __restrict__ W) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kN * kN) return;
  const int i = idx / kN, j = idx - i * kN;
  double v;
  if (i == j) {
    v = 0.0;
    for (int s = 0; s < nparts; ++s) v += sumsq_part[(size_t)s * kN + i];
  } else {
    v = 0.5 * (gsum[idx] + gsum[j * kN + i]);
  }
  W[idx] = v;
}

// ---------------------------------------------------------------- 3. batched one-sided Jacobi (cooperative launch)
struct JacobiParams {
  double* W;                       // [L][n][n], column j contiguous at j*n (symmetric start)
  unsigned* bar;                   // [L] barrier counters, zeroed before launch
  unsigned long long* offmax;      // [L][2] max |cos| of the sweep (bits of a non-negative double), zeroed
  int* sweeps_out;                 // [L]
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void layer_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < target) {
      if (clock64() - t0 > 4000000000ll) {
        printf("dkd: LRKD Jacobi barrier timed out (block %d,%d)\n", blockIdx.x, blockIdx.y);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Two-level (block) one-sided Jacobi.  The 384 columns form 24 groups of 16; in an outer round every CTA owns a pair
// of groups (round-robin tournament over the groups: 23 outer rounds per sweep), copies its 32 columns (96 KB of fp64)
// into shared memory, rotates the 16 x 16 cross pairs there (16 inner rounds, 16 warps = 16 disjoint column pairs per
// inner round, __syncthreads between rounds; the pairs inside a group are done once per sweep, in outer round 0) and
// writes them back.  One grid barrier per OUTER round: 23 per sweep instead of the 383 of a flat ordering — the
// solver is bound by the latency of its sequential rounds, not by arithmetic.
template <int G>
__global__ void __launch_bounds__(32 * G) jacobi_kernel(JacobiParams p) {
  constexpr int n = kN, PER = n / 32;          // 12 elements per lane
  constexpr int NG = n / G;                    // column groups
  constexpr int kJacobiThreads = 32 * G;
  constexpr int LC = 2 * G;                    // 32 columns per CTA
  extern __shared__ double scol[];             // [LC][n]
  const int layer = blockIdx.y;
  double* W = p.W + (size_t)layer * n * n;
  unsigned* bar = p.bar + layer;
  unsigned long long* offmax = p.offmax + 2 * layer;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // 16 warps
  // columns whose norm is below 1e-13 * trace(G) belong to the null space (rank-deficient teacher): never rotated
  __shared__ double s_tiny;
  __shared__ double s_off[kJacobiThreads / 32];
  if (warp == 0) {
    double tr = 0.0;
    for (int i = lane; i < n; i += 32) tr += W[(size_t)i * n + i];
    tr = warp_sum_d(tr);
    if (lane == 0) s_tiny = (1e-13 * tr) * (1e-13 * tr);
  }
  __syncthreads();
  const double tiny = s_tiny;
  unsigned epoch = 0;
  int sweep = 0;
  for (; sweep < kMaxSweeps; ++sweep) {
    double local_off = 0.0;
    for (int r = 0; r < NG - 1; ++r) {
      // round-robin tournament over groups: position 0 is fixed, the others rotate; this CTA takes (pos c, pos NG-1-c)
      const int pi = blockIdx.x, pj = NG - 1 - blockIdx.x;
      const int gp = pi == 0 ? 0 : 1 + (pi - 1 + r) % (NG - 1);
      const int gq = 1 + (pj - 1 + r) % (NG - 1);
      for (int idx = threadIdx.x; idx < LC * n; idx += kJacobiThreads) {
        const int c = idx / n, i = idx - c * n;
        const int gc = (c < G ? gp * G : gq * G - G) + c;
        scol[idx] = __ldcg(W + (size_t)gc * n + i);
      }
      __syncthreads();
      // inner rounds: in outer round 0 first the pairs INSIDE each of the two groups (15 rounds, 8 pairs per group),
      // then — every outer round — the 16 x 16 cross pairs (16 rounds of 16 disjoint pairs).  Per sweep every column
      // pair is rotated exactly once: 23 * 16 + 15 = 383 sequential rounds, as in a flat ordering, but 360 of the
      // 383 synchronisations are __syncthreads instead of grid barriers.
      const int n_inner = (r == 0 ? G - 1 : 0) + G;
      for (int ir = 0; ir < n_inner; ++ir) {
        int cp, cq;
        if (r == 0 && ir < G - 1) {            // within-group tournament: first half of the warps group P, second half group Q
          const int base = (warp / (G / 2)) * G, w8 = warp % (G / 2);
          const int qi = w8, qj = G - 1 - w8;
          cp = base + (qi == 0 ? 0 : 1 + (qi - 1 + ir) % (G - 1));
          cq = base + 1 + (qj - 1 + ir) % (G - 1);
        } else {
          const int k = ir - (r == 0 ? G - 1 : 0);
          cp = warp;
          cq = G + ((warp + k) & (G - 1));
        }
        double* a = scol + cp * n;
        double* b = scol + cq * n;
        double x[PER], y[PER];
        double aa = 0.0, bb = 0.0, ab = 0.0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { x[k] = a[lane + 32 * k]; y[k] = b[lane + 32 * k]; }
#pragma unroll
        for (int k = 0; k < PER; ++k) { aa = fma(x[k], x[k], aa); bb = fma(y[k], y[k], bb); ab = fma(x[k], y[k], ab); }
        aa = warp_sum_d(aa); bb = warp_sum_d(bb); ab = warp_sum_d(ab);
        const double prod = aa * bb;
        const double cosv = (prod > 0.0 && (aa > tiny || bb > tiny)) ? fabs(ab) * rsqrt(prod) : 0.0;
        local_off = fmax(local_off, cosv);
        if (cosv > kJacobiTol) {
          // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (bb - aa) / (2 ab), written with one division:
          // t = sign(d h) |h| / (|d| + sqrt(d^2 + h^2)), d = bb - aa, h = 2 ab ;  c = 1/sqrt(1 + t^2), s = c t
          const double d = bb - aa, h = 2.0 * ab;
          const double t = ((d >= 0.0) == (h >= 0.0) ? fabs(h) : -fabs(h)) / (fabs(d) + sqrt(fma(d, d, h * h)));
          const double c = rsqrt(fma(t, t, 1.0)), sn = c * t;
#pragma unroll
          for (int k = 0; k < PER; ++k) {
            a[lane + 32 * k] = c * x[k] - sn * y[k];
            b[lane + 32 * k] = sn * x[k] + c * y[k];
          }
        }
        __syncthreads();
      }
      for (int idx = threadIdx.x; idx < LC * n; idx += kJacobiThreads) {
        const int c = idx / n, i = idx - c * n;
        const int gc = (c < G ? gp * G : gq * G - G) + c;
        W[(size_t)gc * n + i] = scol[idx];
      }
      if (r == NG - 2) {   // end of the sweep: publish this CTA's largest |cos|
        if (lane == 0) s_off[warp] = local_off;
        __syncthreads();
        if (threadIdx.x == 0) {
          double m = 0.0;
          for (int w = 0; w < kJacobiThreads / 32; ++w) m = fmax(m, s_off[w]);
          atomicMax(offmax + (sweep & 1), (unsigned long long)__double_as_longlong(m));
        }
      }
      ++epoch;
      layer_barrier(bar, epoch * gridDim.x);
      if (r == 0 && blockIdx.x == 0 && threadIdx.x == 0) offmax[(sweep + 1) & 1] = 0ull;  // buffer of the next sweep
    }
    const double off = __longlong_as_double((long long)__ldcg(offmax + (sweep & 1)));
    if (off <= kJacobiStop) { ++sweep; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.sweeps_out) p.sweeps_out[layer] = sweep;
}

// ---------------------------------------------------------------- 4. select top-k, build fused operands
struct SelectParams {
  const double* W;                 // [L][n][n] after Jacobi: column j = lambda_j v_j
  const float* head_w[kMaxLayers];  // W' [k, Ds]
  const float* head_b[kMaxLayers];  // b' [k] or null
  float* Vk_out[kMaxLayers];        // [k, n] or null
  float* S_out[kMaxLayers];         // [k] or null
  __nv_bfloat16* Wcat;             // [L][P][BN][Ds + n]   K-major operand [W' | -V_k], rows >= k zero
  __nv_bfloat16* Wt;               // [L][P][Ds][BN]       W'^T (dgrad operand), cols >= k zero
  float* bias_pad;                 // [L][BN]
  int k, BN, Ds, P;
};

__global__ void __launch_bounds__(kN) select_kernel(SelectParams p) {
  constexpr int n = kN;
  __shared__ double lam[n];
  __shared__ int order[n];         // order[r] = column with rank r
  __shared__ float sgn[n];
  const int layer = blockIdx.x, j = threadIdx.x;
  const double* W = p.W + (size_t)layer * n * n;
  {
    const double* col = W + (size_t)j * n;
    double a = 0.0, big = 0.0;
    for (int i = 0; i < n; ++i) {
      const double v = col[i];
      a = fma(v, v, a);
      if (fabs(v) > fabs(big)) big = v;
    }
    lam[j] = sqrt(a);
    sgn[j] = big < 0.0 ? -1.f : 1.f;   // sign convention: largest-magnitude component positive
  }
  __syncthreads();
  {
    const double v = lam[j];
    int rank = 0;
    for (int i = 0; i < n; ++i) rank += (lam[i] > v) || (lam[i] == v && i < j);
    order[rank] = j;
  }
  __syncthreads();
  const int K = p.Ds + n;
  __nv_bfloat16* wc = p.Wcat + (size_t)layer * p.P * p.BN * K;
  __nv_bfloat16* wt = p.Wt + (size_t)layer * p.P * p.Ds * p.BN;
  const float* hw = p.head_w[layer];
  const float* hb = p.head_b[layer];
  float* vk = p.Vk_out[layer];
  float* so = p.S_out[layer];
  auto put = [&](__nv_bfloat16* base, size_t plane_stride, size_t off, float x) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    base[off] = hi;
    if (p.P == 2) base[plane_stride + off] = __float2bfloat16_rn(x - __bfloat162float(hi));
  };
  for (int r = 0; r < p.BN; ++r) {
    const bool real = r < p.k;
    // V part: thread j = channel
    float v = 0.f;
    if (real) {
      const int col = order[r];
      v = lam[col] > 0.0 ? (float)(W[(size_t)col * n + j] / lam[col]) * sgn[col] : 0.f;
      if (vk) vk[(size_t)r * n + j] = v;
      if (so && j == 0) so[r] = (float)sqrt(lam[col]);
    }
    put(wc, (size_t)p.BN * K, (size_t)r * K + p.Ds + j, -v);
    if (j < p.Ds) {
      const float w = real ? hw[(size_t)r * p.Ds + j] : 0.f;
      put(wc, (size_t)p.BN * K, (size_t)r * K + j, w);
      put(wt, (size_t)p.Ds * p.BN, (size_t)j * p.BN + r, w);
    }
    if (j == 0) p.bias_pad[(size_t)layer * p.BN + r] = (real && hb) ? hb[r] : 0.f;
  }
}

// ---------------------------------------------------------------- 5. residual GEMM over the concatenated K = Ds + Dt
struct ConcatLoaderParams {
  CUtensorMap tmS, tmT, tmW;
  int s_blocks;   // Ds / 64
  int k_blocks;   // (Ds + Dt) / 64
  int nterms;
};
template <class Cfg>
struct ConcatLoader {
  using Params = ConcatLoaderParams;
  static constexpr uint32_t TX_BYTES = Cfg::STAGE_BYTES;
  static __device__ __forceinline__ int num_k_iters(const Params& p) { return p.k_blocks * p.nterms; }
  static __device__ __forceinline__ void prefetch(const Params& p) {
    sm100::tma_prefetch_desc(&p.tmS);
    sm100::tma_prefetch_desc(&p.tmT);
    sm100::tma_prefetch_desc(&p.tmW);
  }
  static __device__ __forceinline__ void issue(const Params& p, int kit, int mt, int nt, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    const int term = kit / p.k_blocks, kb = kit - term * p.k_blocks;
    int pa, pb;
    term_planes(term, p.nterms, pa, pb);
    if (kb < p.s_blocks) sm100::tma_load_3d(sA, &p.tmS, bar, kb * 64, mt * Cfg::BM, pa);
    else sm100::tma_load_3d(sA, &p.tmT, bar, (kb - p.s_blocks) * 64, mt * Cfg::BM, pa);
    sm100::tma_load_3d(sB, &p.tmW, bar, kb * 64, nt * Cfg::BN, pb);
  }
};

using DgradCfg = GemmCfg<192, 1, 4, 2>;
using WgradCfg = GemmNtCfg<3, true, 208, 0, 4>;

struct Workspace {
  __nv_bfloat16 *Tp[kMaxLayers], *Sp, *Gp, *Wcat, *Wt, *ones;
  float *part, *bias_pad;
  double *gsum, *W, *sumsq, *partials;
  unsigned* bar;
  unsigned long long* offmax;
  size_t bytes;
};
constexpr int kSumsqBlocks = 296;

Workspace carve(void* base, int L, int64_t M, int Ds, int BN, int PT, int P) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 1024); return reinterpret_cast<char*>(base) + o; };
  for (int l = 0; l < kMaxLayers; ++l) w.Tp[l] = l < L ? reinterpret_cast<__nv_bfloat16*>(take((size_t)PT * M * kN * 2)) : nullptr;
  w.Sp = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * M * Ds * 2));
  w.Gp = reinterpret_cast<__nv_bfloat16*>(take((size_t)P * M * BN * 2));
  w.Wcat = reinterpret_cast<__nv_bfloat16*>(take((size_t)L * P * BN * (Ds + kN) * 2));
  w.Wt = reinterpret_cast<__nv_bfloat16*>(take((size_t)L * P * Ds * BN * 2));
  w.ones = reinterpret_cast<__nv_bfloat16*>(take((size_t)2 * 64 * 64 * 2));
  w.part = reinterpret_cast<float*>(take((size_t)kNumSMs * kN * kN * 4 / 3 + (size_t)kN * kN * 4));
  w.bias_pad = reinterpret_cast<float*>(take((size_t)L * BN * 4));
  w.gsum = reinterpret_cast<double*>(take((size_t)kN * kN * 8));
  w.W = reinterpret_cast<double*>(take((size_t)L * kN * kN * 8));
  w.sumsq = reinterpret_cast<double*>(take((size_t)kSumsqBlocks * kN * 8));
  w.partials = reinterpret_cast<double*>(take((size_t)kNumSMs * 8));
  w.bar = reinterpret_cast<unsigned*>(take((size_t)kMaxLayers * 4));
  w.offmax = reinterpret_cast<unsigned long long*>(take((size_t)kMaxLayers * 2 * 8));
  w.bytes = off;
  return w;
}

int bn_for_rank(int rank) { return rank <= 64 ? 64 : 128; }

template <int BN>
int run_residual(const Workspace& ws, int layer, int64_t M, int Ds, int P, float gscale, cudaStream_t st, int* grid_out) {
  using Cfg = GemmCfg<BN, 1, 4, 2>;
  using L = ConcatLoader<Cfg>;
  using E = ResidualMseEpi<Cfg>;
  GemmParams<L, E> p;
  const int K = Ds + kN;
  int rc = make_plane_tmap(&p.ld.tmS, ws.Sp, P, M, Ds, Ds, M * Ds, Cfg::BM, "lrkd S");
  if (rc != DKD_OK) return rc;
  rc = make_plane_tmap(&p.ld.tmT, ws.Tp[layer], P, M, kN, kN, M * kN, Cfg::BM, "lrkd T");
  if (rc != DKD_OK) return rc;
  rc = make_plane_tmap(&p.ld.tmW, ws.Wcat + (size_t)layer * P * BN * K, P, BN, K, K, (int64_t)BN * K, Cfg::BN, "lrkd [W'|-V]");
  if (rc != DKD_OK) return rc;
  p.ld.s_blocks = Ds / 64; p.ld.k_blocks = K / 64; p.ld.nterms = P == 2 ? 3 : 1;
  p.ep.t = nullptr; p.ep.bias = ws.bias_pad + (size_t)layer * BN; p.ep.G = ws.Gp; p.ep.partials = ws.partials;
  p.ep.M = M; p.ep.N = BN; p.ep.n_tok = 1; p.ep.Tt = 1; p.ep.t_off = 0; p.ep.planes = P; p.ep.gscale = gscale; p.ep.t_is_bf16 = 0;
  p.m_tiles = (int)((M + Cfg::BM - 1) / Cfg::BM);
  p.n_tiles = 1;
  const int grid = min(kNumSMs, p.m_tiles);
  auto kern = gemm_tn_kernel<Cfg, L, E>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
  *grid_out = grid;
  return check_launch("dkd_lrkd_fwdbwd: residual GEMM");
}

}  // namespace
}  // namespace dkd

extern "C" {

size_t dkd_lrkd_workspace_bytes(int n_layers, int64_t B, int n_tok, int Ds, int Dt, int rank, int dtype, int precision) {
  using namespace dkd;
  if (n_layers < 1 || n_layers > kMaxLayers || Dt != kN) return 0;
  const int P = precision == DKD_PREC_BF16X3 ? 2 : 1;
  const int PT = dtype == DKD_F32 ? 3 : P;
  return carve(nullptr, n_layers, B * n_tok, Ds, bn_for_rank(rank), PT, P).bytes;
}

int dkd_lrkd_fwdbwd(int n_layers, const void* const* s, const void* const* t, const float* const* W, const float* const* bias,
                    const float* coef, int64_t B, int Ts, int s_off, int Tt, int t_off, int n_tok, int Ds, int Dt, int rank,
                    int dtype, int precision, void* const* g_s, float* const* g_W, float* const* g_b, float* loss,
                    float* const* Vk_out, float* const* S_out, int* sweeps_out, void* workspace, size_t workspace_bytes,
                    dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  const char* fn = "dkd_lrkd_fwdbwd";
  DKD_REQUIRE(dtype == DKD_F32 || dtype == DKD_BF16, DKD_E_DTYPE, "%s: dtype %d", fn, dtype);
  DKD_REQUIRE(precision == DKD_PREC_BF16 || precision == DKD_PREC_BF16X3, DKD_E_UNSUPPORTED, "%s: precision %d", fn, precision);
  DKD_REQUIRE(n_layers >= 1 && n_layers <= kMaxLayers, DKD_E_SHAPE, "%s: 1..%d layers", fn, kMaxLayers);
  DKD_REQUIRE(B > 0 && n_tok > 0 && s_off >= 0 && t_off >= 0 && Ts >= s_off + n_tok && Tt >= t_off + n_tok, DKD_E_SHAPE,
              "%s: bad token geometry", fn);
  DKD_REQUIRE(Ds == 192 && Dt == kN, DKD_E_SHAPE, "%s: built for widths 192 -> 384, got %d -> %d", fn, Ds, Dt);
  DKD_REQUIRE(rank >= 1 && rank <= 128, DKD_E_SHAPE, "%s: rank %d outside [1, 128]", fn, rank);
  DKD_REQUIRE(s && t && W && coef && loss && workspace, DKD_E_SHAPE, "%s: null pointer", fn);
  for (int l = 0; l < n_layers; ++l) {
    DKD_REQUIRE(s[l] && t[l] && W[l], DKD_E_SHAPE, "%s: null pointer in layer %d", fn, l);
    DKD_REQUIRE(g_s == nullptr || (((uintptr_t)g_s[l]) & 31) == 0, DKD_E_ALIGN, "%s: g_s[%d] must be 32-byte aligned", fn, l);
  }
  DKD_REQUIRE((((uintptr_t)workspace) & 1023) == 0, DKD_E_ALIGN, "%s: workspace must be 1024-byte aligned", fn);
  const int P = precision == DKD_PREC_BF16X3 ? 2 : 1;
  const int PT = dtype == DKD_F32 ? 3 : P;
  const int BN = bn_for_rank(rank);
  const int64_t M = B * n_tok;
  DKD_REQUIRE(M < (1ll << 31) - 256, DKD_E_SHAPE, "%s: too many rows", fn);
  DKD_REQUIRE(M >= rank, DKD_E_SHAPE, "%s: rank %d exceeds the %lld teacher rows", fn, rank, (long long)M);
  Workspace ws = carve(workspace, n_layers, M, Ds, BN, PT, P);
  DKD_REQUIRE(workspace_bytes >= ws.bytes, DKD_E_WORKSPACE, "%s: workspace %zu < %zu", fn, workspace_bytes, ws.bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  bool want_grads = false;
  for (int l = 0; l < n_layers; ++l)
    want_grads = want_grads || (g_s && g_s[l]) || (g_W && g_W[l]) || (g_b && g_b[l]);

  cudaMemsetAsync(ws.bar, 0, (size_t)kMaxLayers * 4, st);
  cudaMemsetAsync(ws.offmax, 0, (size_t)kMaxLayers * 2 * 8, st);

  // ---- 1-2. per layer: teacher planes, Gram partials, fp64 reduction
  for (int l = 0; l < n_layers; ++l) {
    if (dtype == DKD_F32)
      teacher_planes_kernel<float><<<kSumsqBlocks, 192, 0, st>>>(reinterpret_cast<const float*>(t[l]), B, Tt, t_off, n_tok, PT, ws.Tp[l], ws.sumsq);
    else
      teacher_planes_kernel<__nv_bfloat16><<<kSumsqBlocks, 192, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(t[l]), B, Tt, t_off, n_tok, PT, ws.Tp[l], ws.sumsq);
    rc = check_launch("dkd_lrkd_fwdbwd: teacher planes");
    if (rc != DKD_OK) return rc;
    using Cfg = GramCfg;
    using L = NtGramLoader;
    GemmNtParamsT<Cfg, L> p;
    rc = make_plane_tmap(&p.ld.tmA, ws.Tp[l], PT, M, kN, kN, M * kN, Cfg::KROWS, "lrkd Gram A");
    if (rc != DKD_OK) return rc;
    p.ld.tmB = p.ld.tmA;
    p.ld.tmOnes = p.ld.tmA;
    p.ld.na_tiles = kN / 128;
    p.ld.total_row_blocks = (int)((M + Cfg::KROWS - 1) / Cfg::KROWS);
    nt_make_splits(p.ld.total_row_blocks, kNumSMs / p.ld.na_tiles, &p.ld.splits, &p.ld.row_blocks_per_split);
    p.ld.b_col0 = 0; p.ld.ldd = kN;
    p.ep.D = ws.part; p.ep.Dcol = nullptr; p.ep.ldd = kN; p.ep.alpha = 1.f; p.ep.store = 1; p.ep.rows_valid = 128;
    p.nterms = dtype == DKD_F32 ? 6 : 1;
    const int grid = min(kNumSMs, p.ld.na_tiles * p.ld.splits);
    auto kern = gemm_nt_kernel<Cfg, L>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    rc = check_launch("dkd_lrkd_fwdbwd: Gram GEMM");
    if (rc != DKD_OK) return rc;
    gram_reduce_kernel<<<(kN * kN + 255) / 256, 256, 0, st>>>(ws.part, p.ld.splits, ws.gsum);
    rc = check_launch("dkd_lrkd_fwdbwd: Gram reduce");
    if (rc != DKD_OK) return rc;
    gram_symmetrize_kernel<<<(kN * kN + 255) / 256, 256, 0, st>>>(ws.gsum, ws.sumsq, kSumsqBlocks, ws.W + (size_t)l * kN * kN);
    rc = check_launch("dkd_lrkd_fwdbwd: Gram symmetrize");
    if (rc != DKD_OK) return rc;
  }

  // ---- 3. eigenvectors: all layers in one cooperative launch
  {
    JacobiParams jp;
    jp.W = ws.W; jp.bar = ws.bar; jp.offmax = ws.offmax; jp.sweeps_out = sweeps_out;
    void* kargs[] = {&jp};
    cudaError_t e;
    if (JacobiCfg<8>::CTAS * n_layers <= kNumSMs) {
      cudaFuncSetAttribute(jacobi_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JacobiCfg<8>::SMEM);
      e = cudaLaunchCooperativeKernel((const void*)jacobi_kernel<8>, dim3(JacobiCfg<8>::CTAS, n_layers), dim3(JacobiCfg<8>::THREADS), kargs,
                                      JacobiCfg<8>::SMEM, st);
    } else {
      cudaFuncSetAttribute(jacobi_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JacobiCfg<16>::SMEM);
      e = cudaLaunchCooperativeKernel((const void*)jacobi_kernel<16>, dim3(JacobiCfg<16>::CTAS, n_layers), dim3(JacobiCfg<16>::THREADS),
                                      kargs, JacobiCfg<16>::SMEM, st);
    }
    if (e != cudaSuccess) {
      set_error("%s: cooperative launch of the Jacobi eigensolver failed: %s", fn, cudaGetErrorString(e));
      cudaGetLastError();
      return DKD_E_LAUNCH;
    }
    rc = check_launch("dkd_lrkd_fwdbwd: Jacobi");
    if (rc != DKD_OK) return rc;
  }

  // ---- 4. top-k selection and fused operands
  {
    SelectParams sp{};
    sp.W = ws.W;
    for (int l = 0; l < n_layers; ++l) {
      sp.head_w[l] = W[l];
      sp.head_b[l] = bias ? bias[l] : nullptr;
      sp.Vk_out[l] = Vk_out ? Vk_out[l] : nullptr;
      sp.S_out[l] = S_out ? S_out[l] : nullptr;
    }
    sp.Wcat = ws.Wcat; sp.Wt = ws.Wt; sp.bias_pad = ws.bias_pad; sp.k = rank; sp.BN = BN; sp.Ds = Ds; sp.P = P;
    select_kernel<<<n_layers, kN, 0, st>>>(sp);
    rc = check_launch("dkd_lrkd_fwdbwd: select");
    if (rc != DKD_OK) return rc;
  }

  // ---- 5-7. per layer: residual GEMM (+ loss), dgrad, wgrad
  for (int l = 0; l < n_layers; ++l) {
    rc = launch_tokens_to_planes(s[l], dtype, B, Ts, s_off, n_tok, Ds, P, nullptr, ws.Sp, st);
    if (rc != DKD_OK) return rc;
    const float c = coef[l] / ((float)M * (float)rank);
    int grid = 0;
    rc = BN == 64 ? run_residual<64>(ws, l, M, Ds, P, 2.f * c, st, &grid) : run_residual<128>(ws, l, M, Ds, P, 2.f * c, st, &grid);
    if (rc != DKD_OK) return rc;
    rc = launch_fold_partials(ws.partials, grid, c, loss, st);
    if (rc != DKD_OK) return rc;
    if (!want_grads) continue;
    void* gs = g_s ? g_s[l] : nullptr;
    float* gW = g_W ? g_W[l] : nullptr;
    float* gb = g_b ? g_b[l] : nullptr;
    if (gs) {
      using Cfg = DgradCfg;
      using L = PlaneLoader<Cfg>;
      using E = StoreRowsEpi<Cfg>;
      GemmParams<L, E> p;
      rc = make_plane_tmap(&p.ld.tmA, ws.Gp, P, M, BN, BN, M * BN, Cfg::BM, "lrkd G");
      if (rc != DKD_OK) return rc;
      rc = make_plane_tmap(&p.ld.tmB, ws.Wt + (size_t)l * P * Ds * BN, P, Ds, BN, BN, (int64_t)Ds * BN, Cfg::BN, "lrkd W'^T");
      if (rc != DKD_OK) return rc;
      p.ld.k_blocks = BN / 64; p.ld.nterms = P == 2 ? 3 : 1;
      p.ep.out = gs; p.ep.drop_mask = nullptr; p.ep.bias = nullptr; p.ep.alpha = 1.f; p.ep.M = M; p.ep.N_total = Ds; p.ep.n_tok = n_tok;
      p.ep.T_out = Ts; p.ep.off = s_off; p.ep.out_is_bf16 = dtype == DKD_BF16;
      p.m_tiles = (int)((M + Cfg::BM - 1) / Cfg::BM); p.n_tiles = Ds / Cfg::BN;
      const int g2 = min(kNumSMs, p.m_tiles * p.n_tiles);
      auto kern = gemm_tn_kernel<Cfg, L, E>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      kern<<<g2, Cfg::THREADS, Cfg::SMEM, st>>>(p);
      rc = check_launch("dkd_lrkd_fwdbwd: dgrad GEMM");
      if (rc != DKD_OK) return rc;
    }
    if (gW || gb) {
      using Cfg = WgradCfg;
      using L = NtPlainLoader<Cfg>;
      DKD_REQUIRE(gW != nullptr, DKD_E_UNSUPPORTED, "%s: g_b without g_W is not supported", fn);
      GemmNtParamsT<Cfg, L> p;
      rc = make_plane_tmap(&p.ld.tmA, ws.Gp, P, M, BN, BN, M * BN, Cfg::KROWS, "lrkd G^T");
      if (rc != DKD_OK) return rc;
      rc = make_plane_tmap(&p.ld.tmB, ws.Sp, P, M, Ds, Ds, M * Ds, Cfg::KROWS, "lrkd S (wgrad)");
      if (rc != DKD_OK) return rc;
      rc = make_plane_tmap(&p.ld.tmOnes, ws.ones, 2, 64, 64, 64, 64 * 64, Cfg::KROWS, "ones tile");
      if (rc != DKD_OK) return rc;
      rc = launch_fill_ones_tile(ws.ones, st);
      if (rc != DKD_OK) return rc;
      cudaMemsetAsync(gW, 0, (size_t)rank * Ds * sizeof(float), st);
      if (gb) cudaMemsetAsync(gb, 0, (size_t)rank * sizeof(float), st);
      p.ep.D = gW; p.ep.Dcol = gb; p.ep.ldd = Ds; p.ep.alpha = 1.f; p.ep.store = 0; p.ep.rows_valid = rank;
      p.ld.ldd = Ds; p.ld.na_tiles = 1; p.ld.b_col0 = 0;
      p.ld.total_row_blocks = (int)((M + Cfg::KROWS - 1) / Cfg::KROWS);
      nt_make_splits(p.ld.total_row_blocks, kNumSMs, &p.ld.splits, &p.ld.row_blocks_per_split);
      p.nterms = P == 2 ? 3 : 1;
      const int g3 = min(kNumSMs, p.ld.splits);
      auto kern = gemm_nt_kernel<Cfg, L>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      kern<<<g3, Cfg::THREADS, Cfg::SMEM, st>>>(p);
      rc = check_launch("dkd_lrkd_fwdbwd: wgrad GEMM");
      if (rc != DKD_OK) return rc;
    }
  }
  return DKD_OK;
}

}  // extern "C"
