// Epilogue policies for gemm_tn_kernel (thread = accumulator row; TMEM read in 32-column chunks).
#pragma once
#include "gemm_tn.cuh"

namespace dkd {

// 32 consecutive activations (fp32 or bf16 storage) -> fp32 registers
__device__ __forceinline__ void load_act32(const void* base, int64_t elem_off, int is_bf16, float (&x)[32]) {
  if (is_bf16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + elem_off;
    ldg256(p, *reinterpret_cast<float(*)[16]>(&x[0]));
    ldg256(p + 16, *reinterpret_cast<float(*)[16]>(&x[16]));
  } else {
    const float* p = reinterpret_cast<const float*>(base) + elem_off;
#pragma unroll
    for (int j = 0; j < 4; ++j) ldg256(p + 8 * j, *reinterpret_cast<float(*)[8]>(&x[8 * j]));
  }
}

// per-CTA fp64 partial of the epilogue threads' fp32 accumulators -> partials[blockIdx.x]
template <int EPI_WARPS = 4>
__device__ __forceinline__ void epilogue_block_partial(float acc, int tid, double* partials) {
  __shared__ double s_part[8];
  double a = warp_sum((double)acc);
  if ((tid & 31) == 0) s_part[tid >> 5] = a;
  asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");  // the epilogue warps only
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < EPI_WARPS; ++w) t += s_part[w];
    partials[blockIdx.x] = t;
  }
}

// *loss += scale * sum(partials[0..n))  — one warp, fixed order (deterministic)
int launch_fold_partials(const double* partials, int n, float scale, float* loss, cudaStream_t st);

struct ResidualMseParams {
  const void* t;          // teacher [B, Tt, N] (fp32 or bf16), rows b*Tt + t_off + i; null -> no subtraction
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

  // teacher values of one 32-column chunk of this thread's row (zeros when there is no teacher / the row is dead)
  static __device__ __forceinline__ void load_teacher(const Params& p, bool live, int64_t trow, int col, float (&tv)[32]) {
    if (p.t == nullptr || !live) {
#pragma unroll
      for (int j = 0; j < 32; ++j) tv[j] = 0.f;
    } else if (p.t_is_bf16) {
      const __nv_bfloat16* tp = reinterpret_cast<const __nv_bfloat16*>(p.t) + trow * p.N + col;
      ldg256(tp, *reinterpret_cast<float(*)[16]>(&tv[0]));
      ldg256(tp + 16, *reinterpret_cast<float(*)[16]>(&tv[16]));
    } else {
      const float* tp = reinterpret_cast<const float*>(p.t) + trow * p.N + col;
#pragma unroll
      for (int j = 0; j < 4; ++j) ldg256(tp + 8 * j, *reinterpret_cast<float(*)[8]>(&tv[8 * j]));
    }
  }
  static __device__ __forceinline__ void chunk(const Params& p, State& st, bool live, int64_t m, int col, uint32_t t_addr, const float (&tv)[32]) {
    float v[32];
    sm100::tmem_ld32(t_addr, v);
    sm100::tmem_ld_wait();
    if (live) {
      float hi[32], lo[32], bs[32];
      ldg_vec32(p.bias ? p.bias + col : nullptr, bs);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = v[j] + bs[j] - tv[j];
        st.acc = fmaf(d, d, st.acc);
        const float g = p.gscale * d;
        hi[j] = g;
        lo[j] = g - __bfloat162float(__float2bfloat16_rn(g));
      }
      __nv_bfloat16* gp = p.G + m * p.N + col;
      stg256(gp, *reinterpret_cast<float(*)[16]>(&hi[0]));
      stg256(gp + 16, *reinterpret_cast<float(*)[16]>(&hi[16]));
      if (p.planes == 2) {
        __nv_bfloat16* gl = gp + p.M * p.N;
        stg256(gl, *reinterpret_cast<float(*)[16]>(&lo[0]));
        stg256(gl + 16, *reinterpret_cast<float(*)[16]>(&lo[16]));
      }
    }
  }

  // The teacher chunk of the NEXT step is requested before the current chunk is processed (two register buffers):
  // each DRAM round trip overlaps a chunk of TMEM reads, arithmetic and stores.  With two epilogue groups, group g
  // takes chunks g, g+2, ...
  static __device__ __forceinline__ void tile(const Params& p, State& st, int m0, int n0, int row_in_tile, uint32_t t_acc, int group = 0,
                                              int groups = 1) {
    const int64_t m = (int64_t)m0 + row_in_tile;
    const bool live = m < p.M;
    const int64_t b = live ? m / p.n_tok : 0;
    const int64_t trow = b * p.Tt + p.t_off + (live ? m - b * p.n_tok : 0);
    const int step = 32 * groups;
    float ta[32], tb[32];
    load_teacher(p, live, trow, n0 + group * 32, ta);
#pragma unroll 1
    for (int c0 = group * 32; c0 < Cfg::BN; c0 += 2 * step) {
      if (c0 + step < Cfg::BN) load_teacher(p, live, trow, n0 + c0 + step, tb);
      chunk(p, st, live, m, n0 + c0, t_acc + c0, ta);
      if (c0 + step < Cfg::BN) {
        if (c0 + 2 * step < Cfg::BN) load_teacher(p, live, trow, n0 + c0 + 2 * step, ta);
        chunk(p, st, live, m, n0 + c0 + step, t_acc + c0 + step, tb);
      }
    }
  }

  static __device__ __forceinline__ void finish(const Params& p, State& st, int tid) {
    epilogue_block_partial<Cfg::EPI_WARPS>(st.acc, tid, p.partials);
  }
};

struct StoreRowsParams {
  void* out;              // [B, T_out, N_total] fp32 or bf16; rows b*T_out + off + i
  const float* drop_mask; // [M] 0/1 or null: rows with mask != 0 are stored as zeros
  const float* bias;      // [N_total] added after scaling, or null
  float alpha;            // out = alpha * acc (+ bias)
  int64_t M;
  int N_total, n_tok, T_out, off;
  int out_is_bf16;
};

template <class Cfg>
struct StoreRowsEpi {
  using Params = StoreRowsParams;
  struct State {};
  static __device__ __forceinline__ void init(const Params&, State&) {}
  static __device__ __forceinline__ void tile(const Params& p, State&, int m0, int n0, int row_in_tile, uint32_t t_acc, int group = 0,
                                              int groups = 1) {
    const int64_t m = (int64_t)m0 + row_in_tile;
    const bool live = m < p.M;
    const int64_t b = live ? m / p.n_tok : 0;
    const int64_t i = live ? m - b * p.n_tok : 0;
    const int64_t orow = b * p.T_out + p.off + i;
#pragma unroll 1
    for (int c0 = group * 32; c0 < Cfg::BN; c0 += 32 * groups) {
      float v[32];
      sm100::tmem_ld32(t_acc + c0, v);
      sm100::tmem_ld_wait();
      if (!live) continue;
      if (p.drop_mask != nullptr && __ldg(p.drop_mask + m) != 0.f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      } else {
        float bs[32];
        ldg_vec32(p.bias ? p.bias + n0 + c0 : nullptr, bs);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] * p.alpha + bs[j];
      }
      float z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0.f;
      if (p.out_is_bf16) {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.N_total + n0 + c0;
        stg256(op, *reinterpret_cast<float(*)[16]>(&v[0]));
        stg256(op + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
        if (i == 0)  // the special-token rows in front of this sample's patches get zero gradient
          for (int r = 1; r <= p.off; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) Vec<__nv_bfloat16, 8>::store(op - (int64_t)r * p.N_total + 8 * j, *reinterpret_cast<float(*)[8]>(&z[8 * j]));
      } else {
        float* op = reinterpret_cast<float*>(p.out) + orow * p.N_total + n0 + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) stg256(op + 8 * j, *reinterpret_cast<float(*)[8]>(&v[8 * j]));
        if (i == 0)
          for (int r = 1; r <= p.off; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) Vec<float, 4>::store(op - (int64_t)r * p.N_total + 4 * j, *reinterpret_cast<float(*)[4]>(&z[4 * j]));
      }
    }
  }
  static __device__ __forceinline__ void finish(const Params&, State&, int) {}
};


}  // namespace dkd
