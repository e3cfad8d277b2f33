// Fused forward + dgrad of the hidden-state matching loss (CurKD early / mid, ViTKD mimicking; model/loss.py:376-393,
// 277-289) for one Linear(192 -> 384) head:
//     Y = S W^T + b ;  d = Y - t ;  loss += sum d^2 ;  G = gscale * d (bf16 planes, kept for the weight-gradient GEMM) ;
//     g_s = G W
// in ONE persistent tcgen05 kernel per layer.  Per 128-row tile the S planes stay in shared memory, W streams through a
// ring in 64-row chunks of the teacher width (Dt), and the SAME shared-memory chunk is used twice: K-major as the B
// operand of the forward MMA (Y_c = S W_c^T, K = Ds) and MN-major as the B operand of the dgrad MMA
// (g_s += d_c W_c, K = the chunk's 64 teacher channels).  The residual chunk d_c goes TMEM -> registers (teacher
// subtracted, loss accumulated) -> a 128-byte-swizzled K-major shared-memory tile -> A operand of dgrad (and source of the G store):
// the G planes are never re-read from HBM for dgrad and the g_s accumulator (128 x 192 fp32) lives in TMEM across the
// six chunks.  Compared with the three-launch form (forward, dgrad, wgrad) this removes the dgrad kernel's 154 MB (fp32
// mode) read of G per layer and one pass over the S planes.
//
// Warp roles (352 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (two groups of four: group g
// owns columns [32g, 32g+32) of every 64-column chunk and alternate 32-column pieces of g_s), warp 10 sends every
// finished d tile to the G planes with ONE tensor store per plane (cp.async.bulk.tensor from the swizzled operand tile):
// the epilogue threads issue no global stores for G (a thread-per-row store touches 32 cache lines per instruction — the
// three-launch form's forward kernel is bound by exactly that), and the chunk's critical path is only
// TMEM load -> subtract -> shared-memory store -> fence -> arrive.  The MMA thread issues forward and dgrad chunks in
// whichever order their inputs become ready (W ring slot loaded / d tile written).
// TMEM (512 columns): Y chunk double-buffered (2 x 64) + g_s double-buffered (2 x 192).
// Shared memory, P = 2 planes (fp32 parity mode): S 96 KB + W ring 2 x 48 KB + d tile 32 KB = 224 KB;
//                P = 1 (bf16): S 48 KB + W ring 4 x 24 KB + d tile 16 KB = 160 KB.
#pragma once
#include <stdlib.h>
#include "epilogues.cuh"

namespace dkd {

template <int P_>
struct AlignFusedCfg {
  static constexpr int P = P_;                        // bf16 planes per operand (1: bf16, 2: bf16x3)
  static constexpr int TERMS = P_ == 2 ? 3 : 1;
  static constexpr int DS = 192, DT = 384, CH = 64;   // student width, teacher width, teacher channels per chunk
  static constexpr int NCH = DT / CH;                 // 6 chunks per tile
  static constexpr int KB = DS / 64;                  // 3 K blocks of the forward GEMM
  static constexpr int WST = P_ == 2 ? 2 : 4;         // W ring stages
  static constexpr int S_PLANE = KB * 128 * 128;      // 48 KB: [KB][128 rows][128 B]
  static constexpr int W_PLANE = KB * CH * 128;       // 24 KB: [KB][64 rows][128 B]
  static constexpr int D_PLANE = 128 * 128;           // 16 KB: [128 rows][64 k]
  static constexpr int S_BYTES = P_ * S_PLANE, W_STAGE = P_ * W_PLANE, D_BYTES = P_ * D_PLANE;
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32;   // + the G-store warp
  static constexpr int NBARS = 2 + 2 * WST + 4 + 2 + 4;
  static constexpr size_t SMEM = (size_t)S_BYTES + (size_t)WST * W_STAGE + D_BYTES + 1024 + 256;
  static constexpr uint32_t TM_Y = 0, TM_GS = 128;    // TMEM column offsets
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

struct AlignFusedParams {
  CUtensorMap tmS;        // S planes [P][M][192], box {64, 128, 1}
  CUtensorMap tmW;        // W planes [P][384][192], box {64, 64, 1}
  CUtensorMap tmG;        // G planes [P][M][384], box {64, 128, 1} (stores)
  const void* t;          // teacher [B, Tt, 384]
  const float* bias;      // [384] or null
  __nv_bfloat16* G;       // planes [P][M][384]
  void* g_s;              // [B, Ts, 192] dtype, or null
  double* partials;       // [gridDim.x]
  int64_t M;
  int n_tok, Tt, t_off, Ts, s_off;
  int t_is_bf16, out_is_bf16;
  float gscale;           // G = gscale * d
  float gs_alpha;         // g_s = gs_alpha * (G W)
  int m_tiles;
  int debug_skip;         // timing experiments only (DKD_FUSED_DEBUG bit mask; results are then WRONG): 1 no G store, 2 no teacher
                          // loads, 4 no g_s stores, 8 no dgrad MMAs, 16 no forward MMAs
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) align_fused_kernel(const __grid_constant__ AlignFusedParams p) {
  using namespace sm100;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sS = smem;
  uint8_t* sW = sS + Cfg::S_BYTES;
  uint8_t* sD = sW + (size_t)Cfg::WST * Cfg::W_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + Cfg::D_BYTES);
  uint64_t* s_full = bars;
  uint64_t* s_empty = bars + 1;
  uint64_t* w_full = bars + 2;
  uint64_t* w_empty = w_full + Cfg::WST;
  uint64_t* y_full = w_empty + Cfg::WST;      // [2]
  uint64_t* y_empty = y_full + 2;             // [2]
  uint64_t* d_full = y_empty + 2;
  uint64_t* d_empty = d_full + 1;
  uint64_t* gs_full = d_empty + 1;            // [2]
  uint64_t* gs_empty = gs_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gs_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    mbar_init(s_full, 1); mbar_init(s_empty, 1);
    for (int s = 0; s < Cfg::WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&y_full[a], 1); mbar_init(&y_empty[a], Cfg::EPI_WARPS);
      mbar_init(&gs_full[a], 1); mbar_init(&gs_empty[a], Cfg::EPI_WARPS);
    }
    mbar_init(d_full, Cfg::EPI_WARPS); mbar_init(d_empty, 2);   // d tile free = dgrad MMAs retired + tensor store has read it
    fence_barrier_init();
    tma_prefetch_desc(&p.tmS);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmG);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      uint32_t ti = 0, wi = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++ti) {
        mbar_wait(s_empty, (ti & 1) ^ 1);
        mbar_expect_tx(s_full, Cfg::S_BYTES);
#pragma unroll
        for (int pl = 0; pl < Cfg::P; ++pl)
#pragma unroll
          for (int kb = 0; kb < Cfg::KB; ++kb)
            tma_load_3d(sS + (size_t)pl * Cfg::S_PLANE + (size_t)kb * 128 * 128, &p.tmS, s_full, kb * 64, tile * 128, pl);
        for (int c = 0; c < Cfg::NCH; ++c, ++wi) {
          const int ws = wi % Cfg::WST;
          mbar_wait(&w_empty[ws], ((wi / Cfg::WST) & 1) ^ 1);
          mbar_expect_tx(&w_full[ws], Cfg::W_STAGE);
          uint8_t* dst = sW + (size_t)ws * Cfg::W_STAGE;
#pragma unroll
          for (int pl = 0; pl < Cfg::P; ++pl)
#pragma unroll
            for (int kb = 0; kb < Cfg::KB; ++kb)
              tma_load_3d(dst + (size_t)pl * Cfg::W_PLANE + (size_t)kb * Cfg::CH * 128, &p.tmW, &w_full[ws], kb * 64, c * Cfg::CH, pl);
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc_f = make_idesc_bf16(128, Cfg::CH, MAJOR_K, MAJOR_K);     // Y_c[128, 64] = S[128, 192] W_c[64, 192]^T
      constexpr uint32_t idesc_d = make_idesc_bf16(128, Cfg::DS, MAJOR_K, MAJOR_MN);    // g_s[128, 192] += d_c[128, 64] W_c[64, 192]
      uint32_t ti = 0, ci = 0;     // tiles / chunks processed by this CTA
      // dgrad of chunk `cd` (global chunk counter) of the current tile into g_s buffer `gsb`; `first` = chunk 0 of its tile
      auto dgrad = [&](uint32_t cd, uint32_t gsb, bool first) {
        const int ws = cd % Cfg::WST;
        tc_fence_after();
        const uint32_t d_addr = smem_u32(sD), w_addr = smem_u32(sW + (size_t)ws * Cfg::W_STAGE);
        const uint32_t t_gs = tmem_base + Cfg::TM_GS + gsb * Cfg::DS;
#pragma unroll
        for (int term = 0; term < Cfg::TERMS; ++term) {
          const uint32_t a_pl = d_addr + (Cfg::P == 2 && term == 1 ? Cfg::D_PLANE : 0);
          const uint32_t b_pl = w_addr + (Cfg::P == 2 && term == 0 ? Cfg::W_PLANE : 0);
          if (p.debug_skip & 8) continue;
#pragma unroll
          for (int k = 0; k < Cfg::CH / 16; ++k)
            umma_bf16(t_gs, kmajor_desc(a_pl + k * 32), mnmajor_desc(b_pl + k * 2048, Cfg::CH * 128), idesc_d,
                      (first && term == 0 && k == 0) ? 0u : 1u);
        }
        umma_commit(d_empty);          // the d tile may be rewritten ...
        umma_commit(&w_empty[ws]);     // ... and the W chunk's ring slot refilled once these MMAs retire
      };
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++ti) {
        const uint32_t gsb = ti & 1;
        mbar_wait(s_full, ti & 1);
        mbar_wait(&gs_empty[gsb], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        // Forward chunk f needs its W ring slot loaded and a free Y buffer; dgrad chunk d needs the d tile of chunk d from
        // the epilogue.  Whichever is ready goes first (forward preferred: it feeds the epilogue): with two ring slots the
        // W chunk c+2 can only be fetched once dgrad(c) has retired, and a fixed order would expose that fetch every chunk.
        int f = 0, d = 0;
        const uint32_t c0 = ci;                      // global index of this tile's chunk 0
        const long long t_start = clock64();
        while (d < Cfg::NCH) {
          bool did = false;
          if (f < Cfg::NCH) {
            const uint32_t cf = c0 + f;
            const int ws = cf % Cfg::WST;
            const uint32_t yb = cf & 1;
            if (mbar_try_wait(&w_full[ws], (cf / Cfg::WST) & 1) && mbar_try_wait(&y_empty[yb], ((cf >> 1) & 1) ^ 1)) {
              tc_fence_after();
              const uint32_t s_addr = smem_u32(sS), w_addr = smem_u32(sW + (size_t)ws * Cfg::W_STAGE);
              const uint32_t t_y = tmem_base + Cfg::TM_Y + yb * Cfg::CH;
#pragma unroll
              for (int term = 0; term < Cfg::TERMS; ++term) {
                const uint32_t a_pl = s_addr + (Cfg::P == 2 && term == 1 ? Cfg::S_PLANE : 0);
                const uint32_t b_pl = w_addr + (Cfg::P == 2 && term == 0 ? Cfg::W_PLANE : 0);
#pragma unroll
                for (int kb = 0; kb < Cfg::KB; ++kb)
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (!(p.debug_skip & 16)) umma_bf16(t_y, kmajor_desc(a_pl + kb * 128 * 128 + k * 32), kmajor_desc(b_pl + kb * Cfg::CH * 128 + k * 32), idesc_f,
                              (term | kb | k) != 0 ? 1u : 0u);
              }
              umma_commit(&y_full[yb]);
              if (f == Cfg::NCH - 1) umma_commit(s_empty);    // last forward MMA of the tile: the S planes may be replaced
              ++f;
              did = true;
            }
          }
          if (!did && d < f && mbar_try_wait(d_full, (c0 + d) & 1)) {
            dgrad(c0 + d, gsb, d == 0);
            ++d;
            did = true;
          }
          if (!did && clock64() - t_start > 4000000000ll) {
            printf("dkd: fused align MMA issue timed out (block %d, f %d, d %d)\n", blockIdx.x, f, d);
            __trap();
          }
        }
        ci += Cfg::NCH;
        umma_commit(&gs_full[gsb]);
      }
    }
  } else if (warp == 2 + Cfg::EPI_WARPS) {
    // ============================== G store warp ==============================
    // every finished d tile (= this chunk of G as bf16 planes, already in the swizzled box layout) goes to HBM with one
    // tensor store per plane; rows past M are clipped by the tensor map
    if (lane == 0) {
      uint32_t ci = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        for (int c = 0; c < Cfg::NCH; ++c, ++ci) {
          mbar_wait(d_full, ci & 1);
          if (!(p.debug_skip & 1)) {
#pragma unroll
            for (int pl = 0; pl < Cfg::P; ++pl) tma_store_3d(&p.tmG, sD + (size_t)pl * Cfg::D_PLANE, c * Cfg::CH, tile * 128, pl);
            tma_store_commit();
            tma_store_wait_read<0>();       // the store has read the tile
          }
          mbar_arrive(d_empty);
        }
      }
      tma_store_wait<0>();
    }
  } else {
    // ============================== epilogue ==================================
    const int quad = warp & 3;                 // TMEM lane quadrant of this warp
    const int group = (warp - 2) >> 2;         // 0 | 1
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    float acc = 0.f;
    float ta[32], tb[32];
    uint32_t ti = 0, ci = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++ti) {
      const int64_t m = (int64_t)tile * 128 + row_in_tile;
      const bool live = m < p.M;
      const int64_t b = live ? m / p.n_tok : 0;
      const int64_t tok = live ? m - b * p.n_tok : 0;
      const int64_t trow = b * p.Tt + p.t_off + tok;
      // teacher values: two register buffers used alternately (no copies: a copy would wait for the load it was meant to
      // hide); chunk c+1's values are requested before chunk c is processed, chunk 0 of the NEXT tile before this tile's
      // g_s rows are stored
      auto load_t = [&](int64_t trow_, bool live_, int col, float (&x)[32]) {
        if (!live_ || (p.debug_skip & 2)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = 0.f;
        } else {
          load_act32(p.t, trow_ * Cfg::DT + col, p.t_is_bf16, x);
        }
      };
      auto chunk = [&](int c, const float (&tv)[32]) {
        const uint32_t yb = ci & 1;
        const int col = c * Cfg::CH + group * 32;           // teacher / G column of this thread's 32 values
        float bs[32];
        ldg_vec32(p.bias ? p.bias + col : nullptr, bs);
        mbar_wait(&y_full[yb], (ci >> 1) & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem_base + lane_base + Cfg::TM_Y + yb * Cfg::CH + group * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&y_empty[yb]);           // the accumulator chunk is in registers: the next forward may overwrite it
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float d0 = v[j] + bs[j] - tv[j], d1 = v[j + 1] + bs[j + 1] - tv[j + 1];
          if (live) { acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); }
          const float g0 = live ? p.gscale * d0 : 0.f, g1 = live ? p.gscale * d1 : 0.f;
          const __nv_bfloat162 h = __floats2bfloat162_rn(g0, g1);
          hi[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          if (Cfg::P == 2) {
            const __nv_bfloat162 l = __floats2bfloat162_rn(g0 - __low2float(h), g1 - __high2float(h));
            lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&l);
          }
        }
        // d_c into the K-major operand tile — dgrad's A operand AND the source of the G-plane tensor store
        // (128-byte swizzle: 16-byte chunk j of row r sits at chunk j ^ (r & 7))
        mbar_wait(d_empty, (ci & 1) ^ 1);
        {
          uint8_t* drow = sD + row_in_tile * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int phys = (group * 4 + q) ^ (row_in_tile & 7);
            *reinterpret_cast<uint4*>(drow + phys * 16) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
            if (Cfg::P == 2)
              *reinterpret_cast<uint4*>(drow + Cfg::D_PLANE + phys * 16) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
          }
        }
        fence_proxy_async();        // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive(d_full);
        ++ci;
      };
      if (ti == 0) load_t(trow, live, group * 32, ta);      // later tiles: requested at the end of the previous tile
      // the next tile's row (for its chunk-0 prefetch)
      const int64_t m_n = (int64_t)(tile + (int)gridDim.x) * 128 + row_in_tile;
      const bool live_n = tile + (int)gridDim.x < p.m_tiles && m_n < p.M;
      const int64_t b_n = live_n ? m_n / p.n_tok : 0;
      const int64_t trow_n = b_n * p.Tt + p.t_off + (live_n ? m_n - b_n * p.n_tok : 0);
      static_assert(Cfg::NCH % 2 == 0, "chunks are processed in pairs");
#pragma unroll 1
      for (int c = 0; c < Cfg::NCH; c += 2) {
        load_t(trow, live, (c + 1) * Cfg::CH + group * 32, tb);
        chunk(c, ta);
        if (c + 2 < Cfg::NCH) load_t(trow, live, (c + 2) * Cfg::CH + group * 32, ta);
        else load_t(trow_n, live_n, group * 32, ta);
        chunk(c + 1, tb);
      }
      // ---- g_s rows of this tile: 6 pieces of 32 columns, alternating between the two groups
      const uint32_t gsb = ti & 1;
      mbar_wait(&gs_full[gsb], (ti >> 1) & 1);
      tc_fence_after();
      const int64_t orow = b * p.Ts + p.s_off + tok;
#pragma unroll 1
      for (int c0 = group * 32; c0 < Cfg::DS; c0 += 64) {
        float v[32];
        tmem_ld32(tmem_base + lane_base + Cfg::TM_GS + gsb * Cfg::DS + c0, v);
        tmem_ld_wait();
        if (!live || p.g_s == nullptr || (p.debug_skip & 4)) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.gs_alpha;
        float z[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = 0.f;
        if (p.out_is_bf16) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.g_s) + orow * Cfg::DS + c0;
          stg256(op, *reinterpret_cast<float(*)[16]>(&v[0]));
          stg256(op + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
          if (tok == 0)      // the special-token rows in front of this sample's patches get zero gradient
            for (int r = 1; r <= p.s_off; ++r) {
              stg256(op - (int64_t)r * Cfg::DS, *reinterpret_cast<float(*)[16]>(&z[0]));
              stg256(op - (int64_t)r * Cfg::DS + 16, *reinterpret_cast<float(*)[16]>(&z[16]));
            }
        } else {
          float* op = reinterpret_cast<float*>(p.g_s) + orow * Cfg::DS + c0;
#pragma unroll
          for (int j = 0; j < 4; ++j) stg256(op + 8 * j, *reinterpret_cast<float(*)[8]>(&v[8 * j]));
          if (tok == 0)
            for (int r = 1; r <= p.s_off; ++r)
#pragma unroll
              for (int j = 0; j < 4; ++j) stg256(op - (int64_t)r * Cfg::DS + 8 * j, *reinterpret_cast<float(*)[8]>(&z[8 * j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&gs_empty[gsb]);
    }
    epilogue_block_partial<Cfg::EPI_WARPS>(acc, threadIdx.x - 64, p.partials);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// host: forward + dgrad of one layer.  `grid_out` = number of loss partials written.
template <int P>
inline int align_fused_fwd_dgrad_t(const __nv_bfloat16* S, const __nv_bfloat16* Wp, const float* bias, const void* t, int t_is_bf16, int Tt,
                                   int t_off, int n_tok, __nv_bfloat16* G, double* partials, float gscale, void* g_s, int Ts, int s_off,
                                   int out_is_bf16, float gs_alpha, int64_t M, cudaStream_t st, int* grid_out, const char* what) {
  using Cfg = AlignFusedCfg<P>;
  AlignFusedParams p;
  int rc = make_plane_tmap(&p.tmS, S, P, M, Cfg::DS, Cfg::DS, M * Cfg::DS, 128, what);
  if (rc != DKD_OK) return rc;
  rc = make_plane_tmap(&p.tmW, Wp, P, Cfg::DT, Cfg::DS, Cfg::DS, (int64_t)Cfg::DT * Cfg::DS, Cfg::CH, what);
  if (rc != DKD_OK) return rc;
  rc = make_plane_tmap(&p.tmG, G, P, M, Cfg::DT, Cfg::DT, M * Cfg::DT, 128, what);
  if (rc != DKD_OK) return rc;
  p.t = t; p.bias = bias; p.G = G; p.g_s = g_s; p.partials = partials; p.M = M; p.n_tok = n_tok; p.Tt = Tt; p.t_off = t_off;
  p.Ts = Ts; p.s_off = s_off; p.t_is_bf16 = t_is_bf16; p.out_is_bf16 = out_is_bf16; p.gscale = gscale; p.gs_alpha = gs_alpha;
  p.m_tiles = (int)((M + 127) / 128);
  { const char* e = getenv("DKD_FUSED_DEBUG"); p.debug_skip = e ? atoi(e) : 0; }
  const int grid = p.m_tiles < kNumSMs ? p.m_tiles : kNumSMs;
  auto kern = align_fused_kernel<Cfg>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
  *grid_out = grid;
  return check_launch(what);
}

}  // namespace dkd
