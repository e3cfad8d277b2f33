// Persistent, warp-specialised tcgen05 GEMM skeleton for sm_100a:   D[M,N] = sum_k A[M,k] * B[N,k]
// (both operands K-major bf16, fp32 accumulation in TMEM).
//
//   warp 0      : TMA producer (one elected lane) — fills a STAGES-deep ring of {A 128x64, B BNx64} tiles
//   warp 1      : TMEM allocation + tcgen05.mma issue (one elected lane), tcgen05.commit frees ring slots
//   warps 2..5  : epilogue — tcgen05.ld of the accumulator (thread = row), policy-defined math and stores
//   (warps 6..9 : second epilogue group when Cfg::EPI_WARPS == 8)
//
// TMEM holds ACC accumulator stages of BN fp32 columns so the epilogue of tile i overlaps the MMAs
// of tile i+1.  The fp32-parity mode ("bf16x3") is expressed by the Loader as extra K iterations
// over (A plane, B plane) pairs: hi*hi + hi*lo + lo*hi accumulate into the same TMEM tile.
//
// Policies:
//   Loader : Params (tensor maps, extents); num_k_iters(p); issue(p, kit, m_tile, n_tile, sA, sB, bar)
//   Epi    : Params; State; init(state); tile(p, state, m0, n0, lane_row, tmem_acc); finish(p, state)
#pragma once
#include <type_traits>

#include "umma.cuh"

namespace dkd {

// PLANES_ = 2: a ring stage holds BOTH bf16 planes (hi, lo) of the A and B tiles of one K block and the three
// bf16x3 products (hi*lo, lo*hi, hi*hi) are issued from it — every operand byte crosses L2 -> shared memory once
// instead of 1.5 times (the plain scheme replays the K loop per product and re-fetches the hi planes).
// EPI_WARPS_ = 8: two epilogue warps per TMEM lane quadrant (two per scheduler); group g of the two takes the 32-column
// chunks g, g+2, ... of a tile.  The epilogue is a chain of TMEM reads, DRAM round trips and stores; with a single
// epilogue warp per scheduler nothing hides those latencies.  Policies used this way take (group, groups) in tile().
// CLUSTER_ > 1: the kernel runs as thread-block clusters of CLUSTER_ CTAs that walk their tiles in lockstep; loaders that
// support it fetch the operand every CTA of the cluster needs (the weights of a convolution) ONCE per cluster and multicast
// it — the L2 -> SM operand feed (~55 B/clk/SM measured) is what bounds the K = 3456 convolution GEMMs, not the tensor pipe.
template <int BN_, int NI_, int STAGES_, int ACC_, int TILE_M_ = 128, int PLANES_ = 1, int EPI_WARPS_ = 4, int CLUSTER_ = 1>
struct GemmCfg {
  static constexpr int CLUSTER = CLUSTER_;
  static_assert(CLUSTER_ >= 1 && CLUSTER_ <= 8, "portable cluster size");
  static constexpr int PLANES = PLANES_;
  static constexpr int EPI_WARPS = EPI_WARPS_;
  static constexpr int EPI_GROUPS = EPI_WARPS_ / 4;
  static_assert(EPI_WARPS_ == 4 || EPI_WARPS_ == 8, "4 or 8 epilogue warps");
  static constexpr int BM = 128;       // UMMA M (cta_group::1)
  static constexpr int TILE_M = TILE_M_;  // rows of the tile that carry data (126 = 9 image rows for the conv loader)
  static constexpr int BN = BN_;       // tile N
  static constexpr int NI = NI_;       // MMA instructions per K step (N split)
  static constexpr int N_INSTR = BN_ / NI_;
  static constexpr int BK = 64;        // one 128-byte swizzle row of bf16
  static constexpr int STAGES = STAGES_;
  static constexpr int ACC = ACC_;
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = PLANES_ * (A_BYTES + B_BYTES);
  static constexpr int TMEM_COLS_USED = ACC * BN;
  static constexpr int TMEM_COLS = TMEM_COLS_USED <= 32 ? 32 : TMEM_COLS_USED <= 64 ? 64 : TMEM_COLS_USED <= 128 ? 128 : TMEM_COLS_USED <= 256 ? 256 : 512;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS_;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(N_INSTR % 16 == 0 && N_INSTR >= 16 && N_INSTR <= 256, "UMMA N");
  static_assert(TMEM_COLS_USED <= 512, "TMEM columns");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

// 16-element K steps issued per 64-wide ring stage: 4, unless the Loader says fewer (`static constexpr int K_STEPS`) — a
// per-head attention operand is 48 channels wide: the 64-wide box is loaded, the MMAs stop after 3 steps.
template <class L, class = void> struct loader_ksteps { static constexpr int value = 4; };
template <class L> struct loader_ksteps<L, std::void_t<decltype(L::K_STEPS)>> { static constexpr int value = L::K_STEPS; };

template <class Loader, class Epi>
struct GemmParams {
  typename Loader::Params ld;
  typename Epi::Params ep;
  int m_tiles, n_tiles;
  // PLANES == 2 only: the A operand's lo plane is identically zero (e.g. a +-1 gradient plane): its loads and the
  // (A lo, B hi) product are skipped — two products and a third less A traffic.  The loader honours `ld.a_lo_zero`.
  int a_lo_zero = 0;
};

template <class Cfg, class Loader, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS, 1) gemm_tn_kernel(const __grid_constant__ GemmParams<Loader, Epi> p) {
  using namespace sm100;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  constexpr int A_STAGE = Cfg::PLANES * Cfg::A_BYTES, B_STAGE = Cfg::PLANES * Cfg::B_BYTES;
  uint8_t* sB = smem + (size_t)Cfg::STAGES * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;
  uint64_t* acc_empty = acc_full + Cfg::ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + Cfg::ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int num_k = Loader::num_k_iters(p.ld);
  // Clusters walk tiles in lockstep: every CTA of a cluster runs the same number of tile iterations (a CTA whose tile
  // index is past the end runs a phantom tile: its loads are zero-filled, its epilogue stores nothing).
  const int crank = Cfg::CLUSTER > 1 ? (int)cluster_ctarank() : 0;
  const int tile0 = (int)blockIdx.x - crank;           // first tile of this CTA's cluster
  constexpr uint16_t kClusterMask = (uint16_t)((1u << Cfg::CLUSTER) - 1u);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], Cfg::CLUSTER); }
    for (int a = 0; a < Cfg::ACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], Cfg::EPI_WARPS); }
    fence_barrier_init();
    Loader::prefetch(p.ld);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if constexpr (Cfg::CLUSTER > 1) cluster_sync();   // every CTA's barriers exist before a peer multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int base = tile0; base < num_tiles; base += gridDim.x) {
        const int tile = base + crank;
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        for (int kit = 0; kit < num_k; ++kit) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Loader::TX_BYTES - ((Cfg::PLANES == 2 && p.a_lo_zero) ? Cfg::A_BYTES : 0));
          if constexpr (Cfg::CLUSTER > 1) Loader::issue_cluster(p.ld, kit, mt, nt, sA + (size_t)s * A_STAGE, sB + (size_t)s * B_STAGE, &full[s], crank);
          else Loader::issue(p.ld, kit, mt, nt, sA + (size_t)s * A_STAGE, sB + (size_t)s * B_STAGE, &full[s]);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // The whole warp walks the loop (warp-uniform addresses stay on the uniform datapath) and one elected lane issues:
    // under `if (lane == 0)` every descriptor is a per-thread value that ptxas moves to uniform registers with an
    // ELECT / R2UR / BRA.U.ANY loop — ~13 instructions per MMA instead of 1-3.
    {
      constexpr uint32_t idesc = make_idesc_bf16(Cfg::BM, Cfg::N_INSTR, MAJOR_K, MAJOR_K);
      int s = 0; uint32_t ph = 0;
      int a = 0; uint32_t aph = 0;
      for (int base = tile0; base < num_tiles; base += gridDim.x) {
        mbar_wait(&acc_empty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * Cfg::BN);
        for (int kit = 0; kit < num_k; ++kit) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + (size_t)s * A_STAGE);
          const uint32_t b_addr = smem_u32(sB + (size_t)s * B_STAGE);
          constexpr int TERMS = Cfg::PLANES == 2 ? 3 : 1;     // (A plane, B plane): (hi,lo) (lo,hi) (hi,hi) — small terms first
          if (elect_one()) {
#pragma unroll
            for (int term = 0; term < TERMS; ++term) {
              if (Cfg::PLANES == 2 && term == 1 && p.a_lo_zero) continue;
              const uint32_t a_pl = a_addr + (Cfg::PLANES == 2 && term == 1 ? Cfg::A_BYTES : 0);
              const uint32_t b_pl = b_addr + (Cfg::PLANES == 2 && term == 0 ? Cfg::B_BYTES : 0);
#pragma unroll
              for (int k = 0; k < loader_ksteps<Loader>::value; ++k) {
                const uint64_t da = kmajor_desc(a_pl + k * 32);
#pragma unroll
                for (int ni = 0; ni < Cfg::NI; ++ni) {
                  const uint64_t db = kmajor_desc(b_pl + ni * Cfg::N_INSTR * 128 + k * 32);
                  umma_bf16(d_tmem + ni * Cfg::N_INSTR, da, db, idesc, (kit | term | k) != 0 ? 1u : 0u);
                }
              }
            }
            if constexpr (Cfg::CLUSTER > 1) umma_commit_mc(&empty[s], kClusterMask);   // ... in every CTA of the cluster
            else umma_commit(&empty[s]);                  // ring slot reusable once these MMAs retire
            if (kit == num_k - 1) umma_commit(&acc_full[a]);  // accumulator complete
          }
          __syncwarp();
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        if (++a == Cfg::ACC) { a = 0; aph ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;  // accumulator row owned by this thread
    typename Epi::State st;
    Epi::init(p.ep, st);
    int a = 0; uint32_t aph = 0;
    for (int base = tile0; base < num_tiles; base += gridDim.x) {
      const int tile = base + crank;
      const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
      mbar_wait(&acc_full[a], aph);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * Cfg::BN);
      if constexpr (Cfg::EPI_GROUPS > 1) Epi::tile(p.ep, st, mt * Cfg::TILE_M, nt * Cfg::BN, row_in_tile, t_acc, (warp - 2) >> 2, Cfg::EPI_GROUPS);
      else Epi::tile(p.ep, st, mt * Cfg::TILE_M, nt * Cfg::BN, row_in_tile, t_acc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
      if (++a == Cfg::ACC) { a = 0; aph ^= 1; }
    }
    Epi::finish(p.ep, st, threadIdx.x - 64);
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (Cfg::CLUSTER > 1) cluster_sync();   // no CTA leaves while a peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// host: how many clusters of `cluster` CTAs (one CTA per SM: `smem` bytes, `threads`) can be resident at once.  Clusters
// are placed inside one GPC, so this can be fewer than 148 / cluster; a persistent kernel must not launch more (a cluster
// that starts only after another has finished ALL its tiles would double the makespan).  Cached per kernel.
template <class Kern>
inline int max_resident_clusters(Kern kern, int cluster, int threads, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(kNumSMs / cluster * cluster));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = kNumSMs / cluster; }
  if (n > kNumSMs / cluster) n = kNumSMs / cluster;
  return n;
}

// host: CTAs to launch for `tiles` tiles (whole clusters; phantom tiles cover the remainder)
template <class Cfg, class Loader, class Epi>
inline int gemm_tn_grid(int tiles) {
  if constexpr (Cfg::CLUSTER > 1) {
    auto kern = gemm_tn_kernel<Cfg, Loader, Epi>;
    static const int resident = [&] {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      return max_resident_clusters(kern, Cfg::CLUSTER, Cfg::THREADS, Cfg::SMEM);
    }();
    int clusters = (tiles + Cfg::CLUSTER - 1) / Cfg::CLUSTER;
    if (clusters > resident) clusters = resident;
    return clusters * Cfg::CLUSTER;
  } else {
    return tiles < kNumSMs ? tiles : kNumSMs;
  }
}

// host: launch a gemm_tn kernel, as clusters of Cfg::CLUSTER CTAs when the configuration asks for them (grid from gemm_tn_grid)
template <class Cfg, class Loader, class Epi>
inline void launch_gemm_tn(const GemmParams<Loader, Epi>& p, int grid, cudaStream_t st) {
  auto kern = gemm_tn_kernel<Cfg, Loader, Epi>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  if constexpr (Cfg::CLUSTER > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = Cfg::CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, p);
  } else {
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
  }
}

// ------------------------------------------------------------------------------------------------
// Loader: plain K-major operands held as bf16 "plane" tensors [planes, rows, K] (3-D tensor maps,
// box {64, rows_per_box, 1}).  nterms = 1 (bf16) or 3 (bf16x3: planes (0,0), (0,1), (1,0)).
struct PlaneLoaderParams {
  CUtensorMap tmA, tmB;
  int k_blocks;   // ceil(K / 64)
  int nterms;     // 1 or 3
  int a_lo_zero = 0;   // PLANES == 2: do not fetch A's lo plane (see GemmParams::a_lo_zero)
};
template <class Cfg, int B_BOX_ROWS = (Cfg::BN > 256 ? Cfg::BN / 2 : Cfg::BN)>
struct PlaneLoader {
  using Params = PlaneLoaderParams;
  static constexpr uint32_t TX_BYTES = Cfg::STAGE_BYTES;
  static __device__ __forceinline__ int num_k_iters(const Params& p) { return Cfg::PLANES == 2 ? p.k_blocks : p.k_blocks * p.nterms; }
  static __device__ __forceinline__ void prefetch(const Params& p) {
    sm100::tma_prefetch_desc(&p.tmA);
    sm100::tma_prefetch_desc(&p.tmB);
  }
  static __device__ __forceinline__ void issue(const Params& p, int kit, int mt, int nt, uint8_t* sA, uint8_t* sB, uint64_t* bar) {
    if constexpr (Cfg::PLANES == 2) {       // both planes of both operands, once per K block
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        if (!(pl == 1 && p.a_lo_zero)) sm100::tma_load_3d(sA + (size_t)pl * Cfg::A_BYTES, &p.tmA, bar, kit * 64, mt * Cfg::BM, pl);
#pragma unroll
        for (int i = 0; i < Cfg::BN / B_BOX_ROWS; ++i)
          sm100::tma_load_3d(sB + (size_t)pl * Cfg::B_BYTES + (size_t)i * B_BOX_ROWS * 128, &p.tmB, bar, kit * 64,
                             nt * Cfg::BN + i * B_BOX_ROWS, pl);
      }
    } else {
      const int term = kit / p.k_blocks, kb = kit - term * p.k_blocks;
      int pa, pb;
      term_planes(term, p.nterms, pa, pb);
      sm100::tma_load_3d(sA, &p.tmA, bar, kb * 64, mt * Cfg::BM, pa);
#pragma unroll
      for (int i = 0; i < Cfg::BN / B_BOX_ROWS; ++i)
        sm100::tma_load_3d(sB + (size_t)i * B_BOX_ROWS * 128, &p.tmB, bar, kb * 64, nt * Cfg::BN + i * B_BOX_ROWS, pb);
    }
  }
};

// host helper: 3-D map over bf16 planes [planes][rows][cols] (cols contiguous), box {64, box_rows, 1}
inline int make_plane_tmap(CUtensorMap* out, const void* base, int64_t planes, int64_t rows, int64_t cols,
                           int64_t row_stride_elems, int64_t plane_stride_elems, int box_rows, const char* what) {
  const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)rows, (uint64_t)planes};
  const uint64_t strides[2] = {(uint64_t)row_stride_elems * 2, (uint64_t)plane_stride_elems * 2};
  const uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return make_tmap_bf16(out, base, 3, dims, strides, box, what);
}

}  // namespace dkd
