// Split-K "contract over rows" tcgen05 GEMM for sm_100a:   D[NA, NB] += sum_r A[r, NA] * B[r, NB]
// (weight gradients and Gram-type products).  Both operands are row-major bf16 plane tensors whose
// contraction index is the ROW, so they are consumed as MN-major UMMA operands: a TMA box
// {64 cols, 64 rows} (128-byte swizzle) is one 64(mn) x 64(k) operand block.
//
// Work item = (NA tile of 128 columns, row split).  Each CTA loops over its items; per item it
// accumulates over the split's rows in TMEM and the epilogue adds the 128 x NB tile into fp32 D with
// red.global.add (D is zeroed by the host wrapper).  Warp roles as in gemm_tn.cuh.
//
// The B operand may carry one extra box from a constant "ones" tile (column 0 = 1): its first 16
// columns extend N by 16 and column NB_DATA of the result is then sum_r A[r, :] (the bias gradient).
#pragma once
#include "umma.cuh"

namespace dkd {

template <int NB_BOXES_, bool ONES_, int NI_, int STAGES_>
struct GemmNtCfg {
  static constexpr int NA = 128;                      // UMMA M
  static constexpr int NB_BOXES = NB_BOXES_;          // 64-column boxes of B data
  static constexpr bool ONES = ONES_;
  static constexpr int NB_DATA = 64 * NB_BOXES_;
  static constexpr int NB = NB_DATA + (ONES_ ? 16 : 0);   // MMA N (all instructions together)
  static constexpr int NI = NI_;
  static constexpr int N_INSTR = NB / NI_;
  static constexpr int BOX_BYTES = 64 * 128;          // 64 rows x 128 B
  static constexpr int A_BYTES = 2 * BOX_BYTES;
  static constexpr int B_BYTES = (NB_BOXES_ + (ONES_ ? 1 : 0)) * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = STAGES_;
  static constexpr int TMEM_COLS = NB <= 32 ? 32 : NB <= 64 ? 64 : NB <= 128 ? 128 : NB <= 256 ? 256 : 512;
  static constexpr int THREADS = 192;
  static constexpr size_t SMEM = (size_t)STAGES_ * STAGE_BYTES + 1024 + 256;
  static_assert(N_INSTR % 16 == 0 && N_INSTR <= 256 && (N_INSTR % 64 == 0 || NI_ == 1), "UMMA N split must fall on box boundaries");
  static_assert(NB <= 512 && SMEM <= 227 * 1024, "resources");
};

struct GemmNtParams {
  CUtensorMap tmA, tmB, tmOnes;  // 3-D {cols, rows, planes}, box {64, 64, 1}
  float* D;                      // [NA_total, ldd] fp32, += via red.add
  float* Dcol;                   // [NA_total] fp32 (+= column NB_DATA), or null
  int ldd;
  int na_tiles;                  // NA_total / 128
  int splits;                    // row splits
  int row_blocks_per_split;      // 64-row blocks per split
  int total_row_blocks;
  int nterms;                    // 1 (bf16) or 3 (bf16x3)
  int b_col0;                    // first B column (64-aligned)
  float alpha;                   // scale applied in the epilogue
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) gemm_nt_kernel(const __grid_constant__ GemmNtParams p) {
  using namespace sm100;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* acc_full = empty + Cfg::STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.na_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto item_range = [&](int item, int& tile, int& rb0, int& rb1) {
    tile = item % p.na_tiles;
    const int split = item / p.na_tiles;
    rb0 = split * p.row_blocks_per_split;
    rb1 = min(rb0 + p.row_blocks_per_split, p.total_row_blocks);
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int tile, rb0, rb1;
        item_range(item, tile, rb0, rb1);
        for (int term = 0; term < p.nterms; ++term) {
          const int pa = term == 2 ? 1 : 0, pb = term == 1 ? 1 : 0;
          for (int rb = rb0; rb < rb1; ++rb) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
            uint8_t* a = sA + (size_t)s * Cfg::A_BYTES;
            uint8_t* b = sB + (size_t)s * Cfg::B_BYTES;
            tma_load_3d(a, &p.tmA, &full[s], tile * 128, rb * 64, pa);
            tma_load_3d(a + Cfg::BOX_BYTES, &p.tmA, &full[s], tile * 128 + 64, rb * 64, pa);
#pragma unroll
            for (int i = 0; i < Cfg::NB_BOXES; ++i)
              tma_load_3d(b + (size_t)i * Cfg::BOX_BYTES, &p.tmB, &full[s], p.b_col0 + i * 64, rb * 64, pb);
            if constexpr (Cfg::ONES)  // plane 0 of the ones tile is {1,0,0,...}; the lo plane contributes nothing
              tma_load_3d(b + (size_t)Cfg::NB_BOXES * Cfg::BOX_BYTES, &p.tmOnes, &full[s], 0, 0, pb);
            if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, Cfg::N_INSTR, MAJOR_MN, MAJOR_MN);
      int s = 0; uint32_t ph = 0; uint32_t aph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int tile, rb0, rb1;
        item_range(item, tile, rb0, rb1);
        const int nk = (rb1 - rb0) * p.nterms;
        if (nk <= 0) continue;  // (the host never creates empty splits)
        mbar_wait(acc_empty, aph ^ 1);
        tc_fence_after();
        for (int kit = 0; kit < nk; ++kit) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + (size_t)s * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)s * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 64 rows = 4 x UMMA_K; 16 k-rows = 2 groups of 8 rows = 2048 B
            const uint64_t da = mnmajor_desc(a_addr + k * 2048, Cfg::BOX_BYTES);
#pragma unroll
            for (int ni = 0; ni < Cfg::NI; ++ni) {
              const uint64_t db = mnmajor_desc(b_addr + ni * (Cfg::N_INSTR / 64) * Cfg::BOX_BYTES + k * 2048, Cfg::BOX_BYTES);
              umma_bf16(tmem_base + ni * Cfg::N_INSTR, da, db, idesc, (kit | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty[s]);
          if (kit == nk - 1) umma_commit(acc_full);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        aph ^= 1;
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t aph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int tile, rb0, rb1;
      item_range(item, tile, rb0, rb1);
      if (rb1 <= rb0) continue;
      {
        mbar_wait(acc_full, aph);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16);
        float* drow = p.D + (size_t)(tile * 128 + row) * p.ldd;
#pragma unroll 1
        for (int c0 = 0; c0 < Cfg::NB_DATA; c0 += 32) {
          float v[32];
          tmem_ld32(t_acc + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(drow + c0 + j, v[j] * p.alpha);
        }
        if constexpr (Cfg::ONES) {
          float v[32];
          tmem_ld32(t_acc + Cfg::NB_DATA - 16, v);  // columns NB_DATA-16 .. NB_DATA+15 (stay inside the allocation)
          tmem_ld_wait();
          if (p.Dcol) atomicAdd(p.Dcol + tile * 128 + row, v[16] * p.alpha);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      aph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dkd
