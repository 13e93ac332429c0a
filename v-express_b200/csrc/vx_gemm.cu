// tcgen05 GEMM / implicit-GEMM convolution for sm_100a (persistent, warp-specialised).
//
//   out[m, n] = (sum_k A[m, k] * W[n, k] + bias[n] + bias2[m / bias2_div, n]) * scale + residual[m, n]
//   GEGLU mode: W rows are packed per tile as (value | gate); out[m, j] = (v + bias_v) * gelu_erf(g + bias_g)
//
// A is bf16 row-major (K contiguous), W is bf16 [N, K] (K contiguous): both operands are K-major, so every
// 64-wide K block of a 128-row tile is one TMA box that lands in shared memory in the canonical 128B-swizzled
// K-major layout tcgen05.mma consumes.  Accumulation is fp32 in TMEM (two accumulator stages of 256 columns).
//
// Three producers share the same MMA + epilogue:
//   * plain GEMM, optionally split-K over two sources (A | A2) -- the `torch.cat([h, skip])` of the up blocks
//     (reference modules/unet_3d_blocks.py:694,831) folded into the K loop;
//   * 3x3 convolution, stride 1, pad 1, NHWC: K = 9 taps x Cin; for tap (dy,dx) the A box is the same pixel
//     rectangle shifted by (dy-1, dx-1) through a 4-D tensor map (C, W, H, N) whose out-of-bounds reads are
//     zero-filled by the TMA unit = the zero padding of nn.Conv2d (reference modules/resnet.py:9-17).
//
// One persistent CTA per SM walks output tiles (n fastest, so concurrently running CTAs share the A rows in L2).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..9 = epilogue: while the MMA warp fills accumulator stage s^1 they drain stage s:
//   tcgen05.ld -> bias/scale (+ residual read from a TMA-prefetched, 64B-swizzled smem tile) -> bf16 -> same smem
//   tile -> TMA store (coalesced, clipped at the M/N edges by the tensor map).
#include "vx_host.h"
#include "vx_ptx.cuh"

namespace vx {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kThreads = 352;   // + warp 10: TMA-store / residual-prefetch warp
constexpr int kStatThreads = 128;   // LNF instantiations: + warps 11..14, row statistics of the resident A tile (`ares`)
constexpr int kEpiThreads = 256;
constexpr int kPanelCols = 32;                         // staging panel: 32 bf16 = 64 B rows, 64B swizzle
constexpr int kPanelBytes = kBlockM * kPanelCols * 2;  // 8 KB

struct GemmArgs {
  int M, N;        // N = number of accumulator columns overall (2x the output width in GEGLU mode)
  int kblocks1;    // 64-wide K blocks taken from A (per tap for conv)
  int kblocks2;    // ... then from A2 (plain mode only)
  int taps;        // 1 = plain GEMM, 9 = 3x3 conv, 4 = nearest-2x upsample folded into the 3x3 conv (see `ups`)
  int rr;          // 3x3 conv "row reuse": one A box of hbox + 2 image rows per (dx, channel block) feeds the three dy taps
  int a_bytes;     // bytes of one A stage tile (128 rows x 128 B, or (hbox + 2) * W rows x 128 B with rr)
  int ups;         // 1: output parity classes (py, px) of conv3x3(upsample2x(x)) as four 2x2 convolutions on x (vx_upconv3x3_bf16)
  int block_n;     // UMMA N (multiple of 32, <= 256)
  int stages;
  int nbuf;        // staging tiles (2 when shared memory allows: TMA store/residual latency fully hidden)
  int rows_valid;  // output rows covered by one tile (128 for plain; wbox*hbox*nbox for conv)
  int W, H;        // conv OUTPUT image size (= input size at stride 1)
  int cstride;     // conv stride (1 | 2): tap (dy, dx) of output pixel (y, x) reads input pixel (cstride * y + dy, cstride * x + dx)
  int cpad;        // conv padding on the low side (1: nn.Conv2d(padding=1); 0: F.pad(x, (0, 1, 0, 1)) + padding 0)
  int tiles_m, tiles_n;
  int geglu;
  int has_residual;
  int out_f32;     // 1: `out` is float* written directly (attention scores feeding an fp32 softmax)
  const float* bias;
  const float* bias2;
  int bias2_div;
  float scale;
  float* out32;
  long long ldc;
  // LayerNorm folded into the epilogue (LNF instantiations only): out = rstd[m] * (acc - mean[m] * colsum[n]) + bias[n]
  // with W pre-multiplied by gamma, colsum[n] = sum_k W'[n,k], bias[n] = sum_k beta[k] W[n,k] + b[n]
  const float* ln_stats;    // [M][2] = (mean, rstd) of the un-normalised rows of A (null with `ares`)
  const float* ln_colsum;   // [N]
  // A-resident LayerNorm GEMM (LNF instantiations, K <= 512): a CTA (pair) keeps the K blocks of ONE 128-row tile of A in
  // shared memory while it walks ALL column tiles of that row tile (only W streams through the TMA ring), and four extra warps
  // compute the row statistics from the resident tile -- no statistics pass, no LayerNorm pass, 1/tiles_n of the A traffic.
  int ares;
  int ares_bytes;           // kblocks1 x 16 KB
  // LayerNorm statistics handed from the producer GEMM to the consumer GEMM (no statistics pass over the activations):
  //   producer (linear epilogue): rs_out[(tile_n * 2 + half) * rs_stride + m] = (sum, sum of squares) of the bf16-ROUNDED
  //     outputs this epilogue thread wrote for row m (its half of the column tile) -- 2 * tiles_n partials per row;
  //   consumer (LNF instantiations): mean / rstd of row m from ln_nparts such partials, summed in slot order
  //     (deterministic), variance = E[x^2] - mean^2 in fp32, eps = ln_eps, channel count 1 / ln_invK.
  float2* rs_out;
  long long rs_stride;
  const float2* ln_parts;
  long long ln_pstride;
  int ln_nparts;
  float ln_invK;
  // W multicast (CG = 1 only): the grid runs as clusters of two CTAs that own adjacent 128-row tiles and walk the same
  // column tiles in lock step per ring stage; each CTA fetches HALF of every W tile and multicasts it into both shared
  // memories (cp.async.bulk.tensor .multicast::cluster), the MMA commits release a stage in both CTAs.  L2 -> SM operand
  // traffic per 128 x bn x 64 MMA block drops from 16 KB + bn * 128 B to 16 KB + bn * 64 B like in the cta_group::2
  // kernel, but the two CTAs keep their own MMA stream, accumulators and epilogue.
  int mc;
  int strict_arrive;   // cluster-scope release on the pair's accumulator hand-back (default; VX_GEMM_STRICT_ARRIVE=0: .release.cta, A/B)
  float ln_eps;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// erf-GELU with erf from Abramowitz & Stegun 7.1.28 (|error| < 3e-7, i.e. exact at bf16 precision):
// erf(x) = 1 - (1 + a1 x + ... + a6 x^6)^-16 for x >= 0.  ~13 FMA-pipe instructions + one MUFU.RCP instead of
// libdevice erff (the GEGLU epilogue was erf-bound: 128 x 128 erf evaluations per tile).
// ---- cta_group::2 (CTA pair on one 256-row tile) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_ss_cg2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// completion of all prior MMAs of this thread -> arrive on the barrier at the same smem offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// The same arrival with the default (.release.cta) semantics: no cluster-scope fence in front of it.  Enough for the
// accumulator hand-back: what the leader's next MMAs must not overtake are this warp's tcgen05.ld reads, and those have
// completed (tcgen05.wait::ld) and are ordered by tcgen05.fence::before_thread_sync; no generic-proxy data is handed over.
// The cluster-scope release compiles to an ERRBAR in front of every arrival: 15 % of the warp samples of the K = 640 pair
// GEMM (ncu r02_gemm_k640).
__device__ __forceinline__ void mbar_arrive_cluster_cta(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA load of a 2-D box into the same shared-memory offset of every CTA in `mask`; each destination CTA's mbarrier at the
// offset of `bar` receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// completion of all prior MMAs of this thread (cta_group::1) -> arrive on the barrier at the same offset in both CTAs
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ float gelu_erf(float g) {
  const float x = fabsf(g) * 0.70710678118654752f;
  float pl = fmaf(x, 0.0000430638f, 0.0002765672f);
  pl = fmaf(x, pl, 0.0001520143f);
  pl = fmaf(x, pl, 0.0092705272f);
  pl = fmaf(x, pl, 0.0422820123f);
  pl = fmaf(x, pl, 0.0705230784f);
  pl = fmaf(x, pl, 1.0f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(pl));
  r *= r; r *= r; r *= r; r *= r;          // ^16
  const float erf_abs = 1.0f - r;
  return 0.5f * g * (1.0f + copysignf(erf_abs, g));
}

// CG = 1: one CTA per 128-row tile.  CG = 2: a CTA pair (cluster of 2) owns a 256-row tile: each CTA loads its own
// 128 rows of A and HALF of the W tile, the leader issues tcgen05.mma.cta_group::2 (M = 256) and every SM feeds only
// half of B from its shared memory -- the 1-CTA kernel is bound by the SS-MMA operand fetch (A 4 KB + B 8 KB per
// 128x256x16 MMA at ~64 B/clk = 192 clk vs 128 clk of math, profiles/r01e_final_ncu.md).
template <int CG, bool LNF>
__global__ void __launch_bounds__(LNF ? kThreads + kStatThreads : kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
                    const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapR,
                    const __grid_constant__ CUtensorMap mapC, const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atom
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = p.a_bytes;
  const int b_bytes = (p.block_n / CG) * kBlockK * 2;   // this CTA's share of one W tile
  const int nbt = p.rr ? 3 : 1;                         // W tiles per stage (rr: the three dy taps of one dx)
  const int stage_bytes = a_bytes + nbt * b_bytes;
  const bool mc = CG == 1 && p.mc;                 // W multicast between the two CTAs of a cluster (see GemmArgs::mc)
  const int G = (CG == 2 || mc) ? 2 : 1;           // row tiles per work item = CTAs per cluster
  const uint32_t cta_rank = G == 2 ? cluster_ctarank() : 0u;
  const bool pair_leader = CG == 2 ? cta_rank == 0 : true;   // CG = 1: every CTA issues its own MMAs
  uint8_t* ring = smem;                         // TMA ring (behind the resident A tile in `ares` mode)
  if constexpr (LNF) ring += p.ares ? p.ares_bytes : 0;
  uint8_t* sC = ring + p.stages * stage_bytes;  // staging: (block_n or block_n/2)/32 panels of 8 KB
  const int out_cols = p.geglu ? p.block_n / 2 : p.block_n;
  const int npanels = out_cols / kPanelCols;
  const int buf_bytes = npanels * kPanelBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sC + (p.out_f32 ? 0 : p.nbuf * buf_bytes));
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint64_t* c_ready = tmem_empty + 2;          // [2] staging tile b free (+ residual landed)
  uint64_t* staged = c_ready + 2;              // [2] staging tile b fully written by the epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(staged + 2);
  // `ares` only: per K block of the resident tile -- landed in THIS CTA / usable by the MMA (both CTAs landed, on the pair
  // leader) / released (MMAs of the last column tile done + statistics warps done); statistics double buffer
  uint64_t* a_land = staged + 3;          // [8]
  uint64_t* a_full = a_land + 8;          // [8]
  uint64_t* a_empty = a_full + 8;         // [8]
  uint64_t* stats_full = a_empty + 8;     // [2]
  uint64_t* stats_empty = stats_full + 2; // [2]
  float2* s_stats = reinterpret_cast<float2*>(stats_empty + 2);   // [2][128] (mean, rstd)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_kb = p.rr ? 3 * p.kblocks1 : p.taps * p.kblocks1 + p.kblocks2;   // ring stages per output tile
  // work items are (row-tile group of CG tiles, column tile); every CTA of a pair walks the same sequence
  const int tiles_per_par = ((p.tiles_m + G - 1) / G) * p.tiles_n;
  const int num_tiles = tiles_per_par * (p.ups ? 4 : 1);   // ups: (parity, row-tile group, column tile), parity slowest
  const int first_item = blockIdx.x / G, item_stride = gridDim.x / G;
  // the it-th output tile of this CTA (pair) as a flat (row-tile group, column tile) index, or -1 past the end.  Default:
  // tiles strided over the grid, n fastest.  `ares`: row-tile groups strided over the grid, each walked through all its
  // column tiles.
  const int groups = (p.tiles_m + G - 1) / G;
  auto item_at = [&](int it) -> int {
    if constexpr (LNF) {
      if (p.ares) {
        const int g = first_item + (it / p.tiles_n) * item_stride;
        return g < groups ? g * p.tiles_n + it % p.tiles_n : -1;
      }
    }
    const int t = first_item + it * item_stride;
    return t < num_tiles ? t : -1;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    if (p.kblocks2) tma_prefetch_desc(&mapA2);
    if (!p.out_f32) tma_prefetch_desc(&mapC);
    if (p.has_residual || p.ups) tma_prefetch_desc(&mapR);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], mc ? 2 : 1);   // mc: released by the MMA commits of BOTH CTAs (both read the shared W halves)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], CG * kEpiThreads / 32);
      mbar_init(&c_ready[s], 1);
      mbar_init(&staged[s], kEpiThreads);
    }
    if constexpr (LNF) {
      if (p.ares) {
        for (int s = 0; s < 8; ++s) {
          mbar_init(&a_land[s], 1);
          mbar_init(&a_full[s], CG);                    // one arrival per CTA of the pair (its statistics warp 0)
          mbar_init(&a_empty[s], 1 + kStatThreads / 32); // MMA commit + the four statistics warps
        }
        for (int s = 0; s < 2; ++s) {
          mbar_init(&stats_full[s], kStatThreads);
          mbar_init(&stats_empty[s], kEpiThreads / 32);
        }
      }
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_cg2(tmem_slot, 512);
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (G == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, tensor-map prefetch, TMEM allocation, the pair's cluster rendezvous) overlapped the
  // previous grid's tail; from here on every role touches global memory
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp runs the loop in lock-step so that tile / coordinate / barrier values stay warp-uniform (they live
    // in uniform registers); only the elected lane issues.  A divergent `if (lane == 0)` region makes the compiler wrap
    // every UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST loop (~100 cycles per instruction, measured).
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0, t = item_at(0); t >= 0; t = item_at(++it)) {
      const int par = t / tiles_per_par, tt = t - par * tiles_per_par;
      const int tile_n = tt % p.tiles_n, tile_m = (tt / p.tiles_n) * G + (int)cta_rank;
      int n0 = 0, y0 = 0, x0 = 0;
      const long long m0 = (long long)tile_m * p.rows_valid;
      if (p.taps != 1) {
        const long long hw = (long long)p.H * p.W;
        n0 = (int)(m0 / hw);
        const int rem = (int)(m0 % hw);
        y0 = rem / p.W;
        x0 = rem % p.W;
      }
      // the pair's loads all complete on the LEADER's full barrier (it alone waits for the operands)
      const uint32_t tx_bytes = (uint32_t)(CG * ((p.rr ? a_bytes : p.rows_valid * kBlockK * 2) + nbt * b_bytes));
      const int b_row = par * p.N + tile_n * p.block_n + (CG == 2 ? (int)cta_rank * (p.block_n / 2) : 0);
      // CG = 1: the whole W tile [b_row, b_row + block_n) lands at dst; with mc this CTA fetches its half of the rows for both
      auto load_w = [&](uint8_t* dst, int c0, int row) {
        if (mc)
          tma_load_2d_mc(dst + cta_rank * (uint32_t)(b_bytes / 2), &mapB, &full_bar[stage], c0, row + (int)cta_rank * (p.block_n / 2),
                         (uint16_t)3);
        else
          tma_load_2d(dst, &mapB, &full_bar[stage], c0, row);
      };
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = ring + stage * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        const int tap = p.taps != 1 ? kb / p.kblocks1 : 0;     // rr: tap = dx index (0..2)
        const int cb = kb - tap * p.kblocks1;
        if constexpr (LNF) {
          if (p.ares) {
            // resident A: K block kb of this row tile is loaded once, in front of the first column tile, into its own slot
            // (completing on THIS CTA's a_land: the statistics warps of each CTA read their own rows)
            if (tile_n == 0) {
              mbar_wait(&a_empty[kb], (uint32_t)(((it / p.tiles_n) & 1) ^ 1));
              if (leader) {
                mbar_expect_tx(&a_land[kb], (uint32_t)(kBlockM * kBlockK * 2));
                tma_load_2d(smem + kb * (kBlockM * kBlockK * 2), &mapA, &a_land[kb], kb * kBlockK, (int)m0);
              }
            }
            if (leader) {
              if (CG == 2) {
                const uint32_t fb = mapa_rank(smem_u32(&full_bar[stage]), 0);
                if (pair_leader) mbar_expect_tx(&full_bar[stage], (uint32_t)(CG * b_bytes));
                tma_load_2d_cg2(ring + stage * stage_bytes, &mapB, fb, kb * kBlockK, b_row);
              } else {
                mbar_expect_tx(&full_bar[stage], (uint32_t)b_bytes);
                tma_load_2d(ring + stage * stage_bytes, &mapB, &full_bar[stage], kb * kBlockK, b_row);
              }
            }
            __syncwarp();
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
        }
        if (p.rr) {
          // A: image rows y0 - 1 .. y0 + hbox of the column window shifted by dx (zero-filled outside the image = padding);
          // W: the three taps (dy, dx), dy = 0..2, of this channel block
          if (leader) {
            if (CG == 2) {
              const uint32_t fb = mapa_rank(smem_u32(&full_bar[stage]), 0);
              if (pair_leader) mbar_expect_tx(&full_bar[stage], tx_bytes);
              tma_load_4d_cg2(sa, &mapA, fb, cb * kBlockK, x0 + tap - 1, y0 - 1, n0);
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi)
                tma_load_2d_cg2(sb + dyi * b_bytes, &mapB, fb, ((dyi * 3 + tap) * p.kblocks1 + cb) * kBlockK, b_row);
            } else {
              mbar_expect_tx(&full_bar[stage], tx_bytes);
              tma_load_4d(sa, &mapA, &full_bar[stage], cb * kBlockK, x0 + tap - 1, y0 - 1, n0);
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi)
                load_w(sb + dyi * b_bytes, ((dyi * 3 + tap) * p.kblocks1 + cb) * kBlockK, b_row);
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        // 3x3: taps (dy, dx) in {-1, 0, 1}^2.  Folded upsample: output pixel (2i + py, 2j + px) reads the 2x2 input
        // neighbourhood rows i + py - 1 + {0, 1}, columns j + px - 1 + {0, 1} (weights pre-summed per parity on the host).
        const int dy = p.ups ? (tap >> 1) + (par >> 1) - 1 : tap / 3 - p.cpad;
        const int dx = p.ups ? (tap & 1) + (par & 1) - 1 : tap % 3 - p.cpad;
        if (leader) {
          if (CG == 2) {
            const uint32_t fb = mapa_rank(smem_u32(&full_bar[stage]), 0);
            if (pair_leader) mbar_expect_tx(&full_bar[stage], tx_bytes);
            if (p.taps != 1) {
              tma_load_4d_cg2(sa, &mapA, fb, cb * kBlockK, x0 * p.cstride + dx, y0 * p.cstride + dy, n0);
            } else if (kb < p.kblocks1) {
              tma_load_2d_cg2(sa, &mapA, fb, kb * kBlockK, (int)m0);
            } else {
              tma_load_2d_cg2(sa, &mapA2, fb, (kb - p.kblocks1) * kBlockK, (int)m0);
            }
            tma_load_2d_cg2(sb, &mapB, fb, kb * kBlockK, b_row);
          } else {
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            if (p.taps != 1) {
              tma_load_4d(sa, &mapA, &full_bar[stage], cb * kBlockK, x0 * p.cstride + dx, y0 * p.cstride + dy, n0);
            } else if (kb < p.kblocks1) {
              tma_load_2d(sa, &mapA, &full_bar[stage], kb * kBlockK, (int)m0);
            } else {
              tma_load_2d(sa, &mapA2, &full_bar[stage], (kb - p.kblocks1) * kBlockK, (int)m0);
            }
            load_w(sb, kb * kBlockK, b_row);
          }
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues)
    const bool leader = elect_one() && pair_leader;   // with CG = 2 only the pair leader issues MMAs
    const uint32_t idesc = make_idesc_bf16(kBlockM * CG, (uint32_t)p.block_n, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    if (pair_leader)
    for (int it = 0, t = item_at(0); t >= 0; t = item_at(++it)) {
      const int as = it & 1;
      mbar_wait(&tmem_empty[as], (uint32_t)(((it >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
      for (int kb = 0; kb < total_kb; ++kb) {
        bool a_last = false;   // ares: last column tile of the row tile -> its MMAs release the resident K block
        if constexpr (LNF) {
          if (p.ares) {
            const int tn = t % p.tiles_n;
            if (tn == 0) mbar_wait(&a_full[kb], (uint32_t)((it / p.tiles_n) & 1));
            a_last = tn == p.tiles_n - 1;
          }
        }
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        uint32_t sa = smem_u32(ring + stage * stage_bytes);
        const uint32_t sb = sa + a_bytes;
        if constexpr (LNF) {
          if (p.ares) sa = smem_u32(smem + kb * (kBlockM * kBlockK * 2));
        }
        const uint64_t da = make_smem_desc(sa, 16, 1024, SWZ_128B);
        const uint64_t db = make_smem_desc(sb, 16, 1024, SWZ_128B);
        if (leader) {
          if (p.rr) {
            // tap dy reads the 128 tile rows that start dy image rows (W x 128 B, a multiple of the 1024-B swizzle atom)
            // into the A box, against its own W tile
            const uint32_t a_step = (uint32_t)(p.W * 128) >> 4, b_step = (uint32_t)b_bytes >> 4;
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                const uint64_t a = da + (uint64_t)(dyi * a_step + k * 2), b = db + (uint64_t)(dyi * b_step + k * 2);
                if (CG == 2) umma_ss_cg2(d_tmem, a, b, idesc, (kb | dyi | k) != 0 ? 1u : 0u);
                else umma_ss(d_tmem, a, b, idesc, (kb | dyi | k) != 0 ? 1u : 0u);
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {  // +32 bytes per K step = +2 in the (addr >> 4) field
              if (CG == 2) umma_ss_cg2(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_ss(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if (CG == 2) umma_commit_pair(&empty_bar[stage]);
          else if (mc) umma_commit_mc(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (a_last) {
            if (CG == 2) umma_commit_pair(&a_empty[kb]); else umma_commit(&a_empty[kb]);
          }
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) {
        if (CG == 2) umma_commit_pair(&tmem_full[as]); else umma_commit(&tmem_full[as]);
      }
      __syncwarp();
    }
    pdl_trigger();   // last tile issued (or nothing to issue): the next grid may launch while the epilogue drains
  } else if (warp == 10) {
    // ------------------------------------------------------------ TMA store + residual prefetch (one lane)
    if (lane == 0 && !p.out_f32) {   // few TMA ops per tile: the single-lane form is good enough here
      const uint32_t res_bytes = (uint32_t)(p.rows_valid * out_cols * 2);
      auto arm = [&](int t, int b) {  // make staging tile b usable for output tile t
        if (p.has_residual) {   // (never with ups: the upsampler convs have no residual)
          const int tn_ = t % p.tiles_n, tm_ = (t / p.tiles_n) * G + (int)cta_rank;
          mbar_expect_tx(&c_ready[b], res_bytes);
          for (int pn = 0; pn < npanels; ++pn)
            tma_load_2d(sC + b * buf_bytes + pn * kPanelBytes, &mapR, &c_ready[b], tn_ * out_cols + pn * kPanelCols,
                        tm_ * p.rows_valid);
        } else {
          mbar_arrive(&c_ready[b]);
        }
      };
      for (int b = 0; b < p.nbuf; ++b)
        if (item_at(b) >= 0) arm(item_at(b), b);
      for (int it = 0, t = item_at(0); t >= 0; t = item_at(++it)) {
        const int b = it % p.nbuf;
        const int par = t / tiles_per_par, tt = t - par * tiles_per_par;
        const int tile_n = tt % p.tiles_n, tile_m = (tt / p.tiles_n) * G + (int)cta_rank;
        mbar_wait(&staged[b], (uint32_t)((it / p.nbuf) & 1));
        if (p.ups) {
          // rows of the tile = low-resolution pixels (n, i, j); they land on (n, 2i + py, 2j + px): 5-D map (c, j, i, n, py)
          // per px (mapC: px = 0, mapR: px = 1)
          const long long m0 = (long long)tile_m * p.rows_valid, hw = (long long)p.H * p.W;
          const int n0 = (int)(m0 / hw), rem = (int)(m0 % hw);
          const CUtensorMap* mo = (par & 1) ? &mapR : &mapC;
          if (m0 < p.M)
            for (int pn = 0; pn < npanels; ++pn)
              tma_store_5d(mo, sC + b * buf_bytes + pn * kPanelBytes, tile_n * out_cols + pn * kPanelCols, rem % p.W, rem / p.W,
                           n0, par >> 1);
        } else {
          for (int pn = 0; pn < npanels; ++pn)
            tma_store_2d(&mapC, sC + b * buf_bytes + pn * kPanelBytes, tile_n * out_cols + pn * kPanelCols,
                         tile_m * p.rows_valid);
        }
        tma_store_commit();
        const int tnext = item_at(it + p.nbuf);
        if (tnext >= 0) {
          tma_store_wait_read();  // the store has finished reading tile b
          arm(tnext, b);
        }
      }
      tma_store_wait_all();
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;    // TMEM lane quadrant this warp may access
    const int half = ew >> 2;  // which half of the output chunks this warp drains
    const int row = q * 32 + lane;
    const int nchunks = out_cols / 16;
    const int c_begin = half ? (nchunks + 1) / 2 : 0;
    const int c_end = half ? nchunks : (nchunks + 1) / 2;
    const uint32_t te_leader0 = CG == 2 ? mapa_rank(smem_u32(&tmem_empty[0]), 0) : 0u;
    const uint32_t te_leader1 = CG == 2 ? mapa_rank(smem_u32(&tmem_empty[1]), 0) : 0u;
    for (int it = 0, t = item_at(0); t >= 0; t = item_at(++it)) {
      const int tt = t % tiles_per_par;
      const int tile_n = tt % p.tiles_n, tile_m = (tt / p.tiles_n) * G + (int)cta_rank;
      const int as = it & 1;
      const long long m = (long long)tile_m * p.rows_valid + row;
      const bool row_ok = row < p.rows_valid && m < p.M;
      mbar_wait(&tmem_full[as], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const int sb = it % p.nbuf;
      if (!p.out_f32) mbar_wait(&c_ready[sb], (uint32_t)((it / p.nbuf) & 1));
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
      const int nbase = tile_n * p.block_n;  // accumulator column base (bias index)
      const float* b2 = p.bias2 ? p.bias2 + (row_ok ? (m / p.bias2_div) : 0) * (long long)p.N : nullptr;
      float ln_mean = 0.f, ln_rstd = 1.f;
      if constexpr (LNF) {
        if (p.ares) {
          // statistics of this row tile, computed by warps 11..14 from the resident tile (double-buffered by row tile)
          const int item = it / p.tiles_n, sbuf = item & 1;
          mbar_wait(&stats_full[sbuf], (uint32_t)((item >> 1) & 1));
          const float2 st = s_stats[sbuf * kBlockM + row];
          ln_mean = st.x;
          ln_rstd = st.y;
          if (tile_n == p.tiles_n - 1) {   // last use of this buffer: hand it back
            __syncwarp();
            if (lane == 0) mbar_arrive(&stats_empty[sbuf]);
          }
        } else if (p.ln_parts) {
          if (row_ok) {
            float s1 = 0.f, s2 = 0.f;
            for (int j = 0; j < p.ln_nparts; ++j) {
              const float2 v = __ldg(p.ln_parts + (long long)j * p.ln_pstride + m);
              s1 += v.x;
              s2 += v.y;
            }
            ln_mean = s1 * p.ln_invK;
            ln_rstd = rsqrtf(fmaxf(fmaf(-ln_mean, ln_mean, s2 * p.ln_invK), 0.f) + p.ln_eps);
          }
        } else if (row_ok) {
          const float2 st = *reinterpret_cast<const float2*>(p.ln_stats + 2 * m);
          ln_mean = st.x;
          ln_rstd = st.y;
        }
      }
      float rs_sum = 0.f, rs_sq = 0.f;   // row sums of the rounded outputs (producer side of the LayerNorm hand-over)
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t v[16];
        float f[16];
        tmem_ld16(tacc + (uint32_t)(c * 16), v);
        if (p.geglu) {
          uint32_t g[16];
          tmem_ld16(tacc + (uint32_t)(out_cols + c * 16), g);
          tmem_ld_wait();
          const int nv = nbase + c * 16, ng = nbase + out_cols + c * 16;
          float bv[16], bg[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 t0 = p.bias ? *reinterpret_cast<const float4*>(p.bias + nv + i) : make_float4(0, 0, 0, 0);
            const float4 t1 = p.bias ? *reinterpret_cast<const float4*>(p.bias + ng + i) : make_float4(0, 0, 0, 0);
            bv[i] = t0.x; bv[i + 1] = t0.y; bv[i + 2] = t0.z; bv[i + 3] = t0.w;
            bg[i] = t1.x; bg[i + 1] = t1.y; bg[i + 2] = t1.z; bg[i + 3] = t1.w;
          }
          if constexpr (LNF) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 s0 = *reinterpret_cast<const float4*>(p.ln_colsum + nv + i);
              const float4 s1 = *reinterpret_cast<const float4*>(p.ln_colsum + ng + i);
              const float sv[4] = {s0.x, s0.y, s0.z, s0.w}, sg[4] = {s1.x, s1.y, s1.z, s1.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[i + j] = __float_as_uint(ln_rstd * (__uint_as_float(v[i + j]) - ln_mean * sv[j]));
                g[i + j] = __float_as_uint(ln_rstd * (__uint_as_float(g[i + j]) - ln_mean * sg[j]));
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i)
            f[i] = (__uint_as_float(v[i]) + bv[i]) * gelu_erf(__uint_as_float(g[i]) + bg[i]);
        } else {
          tmem_ld_wait();
          const int n = nbase + c * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if constexpr (LNF) {
            if (n < p.N) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 s0 = *reinterpret_cast<const float4*>(p.ln_colsum + n + i);
                f[i] = ln_rstd * (f[i] - ln_mean * s0.x);
                f[i + 1] = ln_rstd * (f[i + 1] - ln_mean * s0.y);
                f[i + 2] = ln_rstd * (f[i + 2] - ln_mean * s0.z);
                f[i + 3] = ln_rstd * (f[i + 3] - ln_mean * s0.w);
              }
            }
          }
          if (n < p.N) {
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(p.bias + n + i);
                f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
              }
            }
            if (b2) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(b2 + n + i);
                f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
              }
            }
          }
          if (p.scale != 1.0f) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] *= p.scale;
          }
        }
        if (p.out_f32) {
          const int n = nbase + c * 16;
          if (row_ok && n < p.N) {
            float4* fp = reinterpret_cast<float4*>(p.out32 + m * p.ldc + n);
#pragma unroll
            for (int i = 0; i < 4; ++i) fp[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          }
          continue;
        }
        // staging tile: panel (c/2), 16-byte chunks (c&1)*2 + {0,1} of the 64-byte row, 64B swizzle:
        // byte offset o = row*64 + chunk*16 is stored at o ^ (((o >> 7) & 3) << 4)
        uint8_t* panel = sC + sb * buf_bytes + (c >> 1) * kPanelBytes;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t o = (uint32_t)(row * 64 + ((c & 1) * 2 + h) * 16);
          o ^= ((o >> 7) & 3u) << 4;
          uint4* sp = reinterpret_cast<uint4*>(panel + o);
          if (p.has_residual) {
            const uint4 r = *sp;
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 tt = unpack_bf16(rr[i]);
              f[h * 8 + 2 * i] += tt.x;
              f[h * 8 + 2 * i + 1] += tt.y;
            }
          }
          const uint4 pk = make_uint4(pack_bf16(f[h * 8], f[h * 8 + 1]), pack_bf16(f[h * 8 + 2], f[h * 8 + 3]),
                                      pack_bf16(f[h * 8 + 4], f[h * 8 + 5]), pack_bf16(f[h * 8 + 6], f[h * 8 + 7]));
          *sp = pk;
          if (p.rs_out) {
            const uint32_t pw[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 t = unpack_bf16(pw[i]);
              rs_sum += t.x + t.y;
              rs_sq = fmaf(t.x, t.x, rs_sq);
              rs_sq = fmaf(t.y, t.y, rs_sq);
            }
          }
        }
      }
      if (p.rs_out && row_ok) p.rs_out[(long long)(tile_n * 2 + half) * p.rs_stride + m] = make_float2(rs_sum, rs_sq);
      // accumulator stage drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) {   // the pair leader's MMA warp owns both halves
          if (p.strict_arrive) mbar_arrive_cluster(as ? te_leader1 : te_leader0);
          else mbar_arrive_cluster_cta(as ? te_leader1 : te_leader0);
        }
        else mbar_arrive(&tmem_empty[as]);
      }
      if (!p.out_f32) {
        fence_proxy_async_smem();     // staging writes -> visible to the TMA store (async proxy)
        mbar_arrive(&staged[sb]);
      }
    }
  }
  if constexpr (LNF) {
    if (warp >= 11 && p.ares) {
      // ---------------------------------------------------------- row statistics of the resident A tile (thread = row)
      // The tile sits in the 128B-swizzled K-major layout: row r of K block kb = 128 bytes at kb * 16 KB + r * 128, its eight
      // 16-byte chunks permuted by r & 7 -- sums do not care.  Chunk (j + r) & 7 at step j keeps the eight rows of a
      // quarter-warp on eight different bank groups.  Two passes (mean, then centred squares): the tile is in shared
      // memory anyway, and the result is the textbook variance rather than E[x^2] - mean^2.
      const int sw = warp - 11, r = sw * 32 + lane;
      const uint32_t a_full_leader = CG == 2 ? mapa_rank(smem_u32(&a_full[0]), 0) : 0u;
      const int K = p.kblocks1 * kBlockK;
      for (int item = 0; item_at(item * p.tiles_n) >= 0; ++item) {
        const int sbuf = item & 1;
        const uint32_t a_phase = (uint32_t)(item & 1);
        mbar_wait(&stats_empty[sbuf], (uint32_t)(((item >> 1) & 1) ^ 1));
        float sum = 0.f;
        for (int kb = 0; kb < p.kblocks1; ++kb) {
          mbar_wait(&a_land[kb], a_phase);
          if (sw == 0 && lane == 0) {   // this CTA's K block has landed: tell the MMA warp (of the pair leader)
            if (CG == 2) mbar_arrive_cluster(a_full_leader + (uint32_t)(kb * 8));
            else mbar_arrive(&a_full[kb]);
          }
          const uint8_t* rowp = smem + kb * (kBlockM * kBlockK * 2) + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(rowp + (((j + r) & 7) << 4));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f2 = unpack_bf16(w[i]);
              sum += f2.x + f2.y;
            }
          }
        }
        const float mean = sum / (float)K;
        float ssq = 0.f;
        for (int kb = 0; kb < p.kblocks1; ++kb) {
          const uint8_t* rowp = smem + kb * (kBlockM * kBlockK * 2) + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(rowp + (((j + r) & 7) << 4));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f2 = unpack_bf16(w[i]);
              const float d0 = f2.x - mean, d1 = f2.y - mean;
              ssq = fmaf(d0, d0, ssq);
              ssq = fmaf(d1, d1, ssq);
            }
          }
        }
        s_stats[sbuf * kBlockM + r] = make_float2(mean, rsqrtf(ssq / (float)K + p.ln_eps));
        __syncwarp();
        if (lane == 0)
          for (int kb = 0; kb < p.kblocks1; ++kb) mbar_arrive(&a_empty[kb]);   // this warp has finished reading the tile
        mbar_arrive(&stats_full[sbuf]);
      }
    }
  }
  tc_fence_before();
  if (G == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// A/B switches (bring-up only) are read ONCE per process: the launch path never touches the environment
// (the sweep tools re-read them through vx_gemm_reload_env).
struct GemmEnv {
  int cg, cg_minkb, pairs, stages, nbuf, bn, verbose, conv_rr, mc, strict_arrive;
  static int geti(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
  }
  GemmEnv() {
    cg = geti("VX_GEMM_CG", 0);
    cg_minkb = geti("VX_GEMM_CG_MINKB", 10);   // K >= 640: profiles/r02_gemm_notes.md (pairs +2..10 % at K = 640, -15 % at K = 320)
    pairs = geti("VX_GEMM_PAIRS", 0);
    stages = geti("VX_GEMM_STAGES", 0);
    nbuf = geti("VX_GEMM_NBUF", 0);
    bn = geti("VX_GEMM_BN", 0);
    verbose = geti("VX_GEMM_VERBOSE", 0);
    conv_rr = geti("VX_CONV_RR", 1);
    mc = geti("VX_GEMM_MC", 0);
    strict_arrive = geti("VX_GEMM_STRICT_ARRIVE", 1);   // 0 measured: +16 % on K = 320 pair tiles, -3..5 % on long-K convs, forward unchanged
  }
};
static GemmEnv& gemm_env() {
  static GemmEnv e;
  return e;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}


static bool pair_ok(int out_f32, int block_n, long long tiles_m, int total_kb) {
  // CTA pairs need an even split of the W tile into 8-row swizzle groups and at least two row tiles.  They pay off
  // once the K loop is long enough to be MMA/operand bound (K >= 640); the K = 320 loops are bound by the epilogue and
  // the output stores, where two independent CTAs overlap better (profiles/tools/gemm_sweep.py, profiles/r02_gemm_notes.md).
  const int mode = gemm_env().cg;   // 0 = auto, 1 = never, 2 = whenever legal
  if (out_f32 || block_n % 32 != 0 || tiles_m < 2 || mode == 1) return false;
  return mode == 2 || total_kb >= gemm_env().cg_minkb;   // threshold from profiles/tools/gemm_sweep.py
}

// W multicast between two 1-CTA tiles (GemmArgs::mc): whenever the pair kernel is not chosen, the W tile splits into two
// halves of whole 8-row swizzle groups and there are at least two row tiles
static bool mc_ok(int out_f32, int block_n, long long tiles_m, int total_kb) {
  return gemm_env().mc && !out_f32 && block_n % 32 == 0 && tiles_m >= 2 && !pair_ok(out_f32, block_n, tiles_m, total_kb);
}
// the W tensor-map box holds half a column tile (each CTA of a pair / cluster fetches one half)
static bool w_split(int out_f32, int block_n, long long tiles_m, int total_kb) {
  return pair_ok(out_f32, block_n, tiles_m, total_kb) || mc_ok(out_f32, block_n, tiles_m, total_kb);
}

// resident CTA pairs of the persistent cta_group::2 kernel (GPCs with an odd SM count strand one SM)
static int num_pairs();

// Pick the UMMA N: among the multiples of `gran` that divide N, minimise
//   waves(work items) x cycles per k-block, with cycles = max(MMA issue 2*bn, smem feed 128 + bn) for one CTA per
//   tile and max(2*bn, 128 + bn/2) for a CTA pair (each SM feeds only half of W).
static int pick_block_n(long long tiles_m, int N, int gran, int out_f32, int total_kb) {
  const int sms = num_sms();
  int best = 0;
  double best_cost = 1e30;
  for (int bn = 256; bn >= gran; bn -= gran) {
    if (N % bn) continue;
    const bool pair = pair_ok(out_f32, bn, tiles_m, total_kb);
    const long long items = (pair ? (tiles_m + 1) / 2 : tiles_m) * (N / bn);
    const int slots = pair ? num_pairs() : sms;
    const double waves = (double)((items + slots - 1) / slots);
    const double feed = pair ? 128.0 + bn / 2 : 128.0 + bn;
    const double cyc = (2.0 * bn > feed) ? 2.0 * bn : feed;
    const double cost = waves * (cyc + 40.0);  // + per-k-block issue overhead
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

static bool use_pair(const GemmArgs& a) {
  if (a.ares) return a.tiles_m >= 2;   // resident A: the pair halves the W stream, the only operand left in the ring
  return pair_ok(a.out_f32, a.block_n, a.tiles_m, a.taps * a.kblocks1 + a.kblocks2);
}

constexpr size_t kAresExtra = 2560;   // ares barriers + the statistics double buffer behind the ordinary barrier block

static int num_pairs() {
  static int n = 0;
  if (n <= 0) {
    cudaFuncSetAttribute(gemm_tcgen05_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * num_sms());
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 227 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int c = 0;
    if (cudaOccupancyMaxActiveClusters(&c, gemm_tcgen05_kernel<2, false>, &cfg) != cudaSuccess || c <= 0) {
      cudaGetLastError();
      c = num_sms() / 2;
    }
    n = gemm_env().pairs > 0 ? gemm_env().pairs : c;
    if (gemm_env().verbose) fprintf(stderr, "[vx_gemm] resident CTA pairs: %d (occupancy query %d)\n", n, c);
  }
  return n;
}

static int launch(const CUtensorMap& mA, const CUtensorMap& mA2, const CUtensorMap& mB, const CUtensorMap& mR,
                  const CUtensorMap& mC, GemmArgs& a, cudaStream_t st) {
  const int cg = use_pair(a) ? 2 : 1;
  a.strict_arrive = gemm_env().strict_arrive;
  a.mc = (cg == 1 && !a.ares && mc_ok(a.out_f32, a.block_n, a.tiles_m, a.taps * a.kblocks1 + a.kblocks2)) ? 1 : 0;
  if (a.ares) a.a_bytes = 0;   // A lives in its own resident slots, the ring stages hold W only
  else if (a.a_bytes <= 0) a.a_bytes = kBlockM * kBlockK * 2;
  const int stage_bytes = a.a_bytes + (a.rr ? 3 : 1) * (a.block_n / cg) * kBlockK * 2;
  const int out_cols = a.geglu ? a.block_n / 2 : a.block_n;
  const int buf_bytes = a.out_f32 ? 0 : out_cols / kPanelCols * kPanelBytes;
  const int total_kb = a.rr ? 3 * a.kblocks1 : a.taps * a.kblocks1 + a.kblocks2;
  const size_t cap = 227 * 1024 - 2048 - (a.ares ? kAresExtra + (size_t)a.ares_bytes : 0);
  // deep TMA rings only pay off for long K loops; short K loops need the second staging tile instead
  int want_stages = gemm_env().stages;
  if (want_stages <= 0) want_stages = total_kb < 6 ? (total_kb < 3 ? 3 : total_kb) : 6;
  int nbuf = (!a.out_f32 && (size_t)3 * stage_bytes + 2 * buf_bytes <= cap) ? 2 : 1;
  if (cg == 2 && nbuf == 2) {
    // a pair at full MMA rate pulls 64 B/clk/SM through the TMA ring: it needs >= 5 stages in flight.  Give up the
    // second staging tile for them unless a residual prefetch shares the staging tile and K is too short to hide it.
    const int st2 = (int)((cap - 2 * (size_t)buf_bytes) / stage_bytes);
    if (st2 < 5 && (!a.has_residual || total_kb >= 40)) nbuf = 1;
  }
  if (a.rr) {
    // a row-reuse stage is 3 taps deep (12 MMAs): three stages when they fit beside ONE staging tile, else two beside two
    nbuf = ((size_t)3 * stage_bytes + buf_bytes <= cap) ? 1 : 2;
    want_stages = nbuf == 1 ? 3 : 2;
    if ((size_t)want_stages * stage_bytes + (size_t)nbuf * buf_bytes > cap) nbuf = 1;
  }
  if (a.ares) {   // every column tile of a row tile streams the whole W panel: a deep ring, two staging tiles when they fit
    nbuf = ((size_t)4 * stage_bytes + 2 * (size_t)buf_bytes <= cap) ? 2 : 1;
    want_stages = 8;
  }
  const int force_nbuf = gemm_env().nbuf;
  if (force_nbuf == 1 || (force_nbuf == 2 && (size_t)2 * stage_bytes + 2 * buf_bytes <= cap)) nbuf = force_nbuf;
  int stages = want_stages;
  while (stages > 2 && (size_t)stages * stage_bytes + (size_t)nbuf * buf_bytes > cap) --stages;
  a.stages = stages;
  a.nbuf = nbuf;
  const size_t smem = (size_t)stages * stage_bytes + (size_t)nbuf * buf_bytes + 2048 + (a.ares ? kAresExtra + (size_t)a.ares_bytes : 0);
  VX_REQUIRE(smem <= (size_t)227 * 1024, "vx_gemm: %zu bytes of shared memory needed (bn=%d, K blocks=%d)", smem, a.block_n,
             a.kblocks1);
  if (gemm_env().verbose)
    fprintf(stderr, "[vx_gemm] M=%d N=%d kb=%d taps=%d cg=%d mc=%d bn=%d stages=%d nbuf=%d tiles=%dx%d\n", a.M, a.N, total_kb,
            a.taps, cg, a.mc, a.block_n, stages, nbuf, a.tiles_m, a.tiles_n);
  static bool configured = false;
  if (!configured) {
    VX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int npar = a.ups ? 4 : 1;
  const bool lnf = a.ln_stats != nullptr || a.ln_parts != nullptr || a.ares;
  const int threads = a.ares ? kThreads + kStatThreads : kThreads;
  if (cg == 1 && a.mc) {
    // clusters of two CTAs (adjacent row tiles), same residency as the pair kernel (one CTA per SM, two SMs of a TPC)
    const int items = ((a.tiles_m + 1) / 2) * a.tiles_n * npar;
    const int clusters = items < num_pairs() ? items : num_pairs();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (lnf) VX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<1, true>, mA, mA2, mB, mR, mC, a));
    else VX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<1, false>, mA, mA2, mB, mR, mC, a));
  } else if (cg == 1) {
    const int tiles = a.ares ? a.tiles_m : a.tiles_m * a.tiles_n * npar;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    if (lnf) VX_CHECK_CUDA(launch_k((gemm_tcgen05_kernel<1, true>), dim3(grid), dim3(threads), smem, st, mA, mA2, mB, mR, mC, a));
    else VX_CHECK_CUDA(launch_k((gemm_tcgen05_kernel<1, false>), dim3(grid), dim3(kThreads), smem, st, mA, mA2, mB, mR, mC, a));
  } else {
    const int items = a.ares ? (a.tiles_m + 1) / 2 : ((a.tiles_m + 1) / 2) * a.tiles_n * npar;
    const int pairs = items < num_pairs() ? items : num_pairs();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (lnf) VX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, true>, mA, mA2, mB, mR, mC, a));
    else VX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, false>, mA, mA2, mB, mR, mC, a));
  }
  VX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int make_out_maps(CUtensorMap* mR, CUtensorMap* mC, const void* residual, long long ldr, void* out,
                         long long ldc, long long M, int out_N, int rows_valid) {
  uint64_t dims[2] = {(uint64_t)out_N, (uint64_t)M};
  uint32_t box[2] = {kPanelCols, (uint32_t)rows_valid};
  {
    uint64_t str[1] = {(uint64_t)ldc * 2};
    if (make_tmap_bf16(mC, out, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  }
  if (residual) {
    uint64_t str[1] = {(uint64_t)ldr * 2};
    if (make_tmap_bf16(mR, residual, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  } else {
    *mR = *mC;
  }
  return 0;
}

}  // namespace vx

using namespace vx;

extern "C" void vx_gemm_reload_env() { gemm_env() = GemmEnv(); }   // sweep-tool hook (csrc/vx_bringup.h), not product ABI

// LayerNorm statistics hand-over between two GEMMs (GemmArgs::rs_out / ln_parts)
struct RowStatsIO {
  float* out;            // producer: [2 * tiles_n][stride] float2 partial (sum, sum of squares) per row, or null
  long long out_stride;  // >= M
  int out_cap;           // slots the caller allocated
  int* nparts_out;       // producer: receives 2 * tiles_n
  const float* parts;    // consumer: partials of the rows of A, or null
  long long parts_stride;
  int nparts;
};

// epilogue: 0 = linear (bias, bias2, scale, residual); 1 = GEGLU (W / bias packed per tile as value|gate halves,
// see vx_geglu_pack_rows; out has N/2 columns)
static int gemm_entry(const void* A, long long lda, int K1, const void* A2, long long lda2, int K2, const void* Wt,
                      long long ldw, int M, int N, const float* bias, const float* bias2, int bias2_div, float scale,
                      const void* residual, long long ldr, void* out, long long ldc, int out_f32, int block_n,
                      const float* ln_stats, const float* ln_colsum, void* stream, float ln_eps = 0.f,
                      const RowStatsIO* rs = nullptr) {
  const int geglu = out_f32 == 2 ? 1 : 0;  // out_f32: 0 bf16, 1 fp32, 2 bf16 + GEGLU epilogue
  const bool ares = ln_colsum && !ln_stats && !(rs && rs->parts);   // LayerNorm GEMM with in-kernel statistics (A tile resident)
  if (geglu) out_f32 = 0;
  VX_REQUIRE(M > 0 && N > 0 && K1 > 0, "vx_gemm_bf16: bad shape M=%d N=%d K1=%d", M, N, K1);
  const int gran = geglu ? 64 : (out_f32 ? 16 : 32);
  VX_REQUIRE(N % gran == 0, "vx_gemm_bf16: N=%d must be a multiple of %d", N, gran);
  VX_REQUIRE(K1 % 8 == 0 && K2 % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0,
             "vx_gemm_bf16: K/ld must be multiples of 8 elements (16-byte TMA strides)");
  VX_REQUIRE(K2 == 0 || (K1 % kBlockK == 0 && lda2 % 8 == 0), "vx_gemm_bf16: split-K needs K1 %% 64 == 0");
  VX_REQUIRE(!residual || (ldr % 8 == 0 && !out_f32 && !geglu), "vx_gemm_bf16: residual needs bf16 linear epilogue, ldr %%8");
  VX_REQUIRE(!geglu || (!bias2 && scale == 1.0f), "vx_gemm_bf16: GEGLU epilogue takes only the packed bias");
  const int tiles_m = (M + kBlockM - 1) / kBlockM;
  const int total_kb = (K1 + kBlockK - 1) / kBlockK + (K2 + kBlockK - 1) / kBlockK;
  if (block_n <= 0) block_n = gemm_env().bn;
  bool b_split = false;   // W tile split between the two CTAs of a pair
  if (ares) {
    VX_REQUIRE(K2 == 0 && K1 % kBlockK == 0 && K1 <= 8 * kBlockK && !out_f32,
               "vx_gemm_ln_bf16: K=%d must be a multiple of 64, <= 512 (the row tile stays in shared memory)", K1);
    b_split = tiles_m >= 2;
    const size_t cap = (size_t)227 * 1024 - 2048 - kAresExtra - (size_t)(K1 / kBlockK) * kBlockM * kBlockK * 2;
    auto fits = [&](int bn, int nbuf) {   // four ring stages + nbuf staging tiles
      const size_t b = (size_t)(b_split ? bn / 2 : bn) * kBlockK * 2, buf = (size_t)((geglu ? bn / 2 : bn) / kPanelCols) * kPanelBytes;
      return 4 * b + nbuf * buf <= cap;
    };
    if (block_n <= 0 || N % block_n || block_n % gran || !fits(block_n, 1)) {
      block_n = 0;
      for (int nbuf = 2; nbuf >= 1 && !block_n; --nbuf)   // widest column tile that keeps both staging tiles, else one
        for (int bn = 256; bn >= gran; bn -= gran)
          if (N % bn == 0 && fits(bn, nbuf)) {
            block_n = bn;
            break;
          }
    }
    VX_REQUIRE(block_n > 0, "vx_gemm_ln_bf16: no column tile of N=%d fits beside the resident K=%d tile", N, K1);
  } else {
    if (block_n <= 0) block_n = pick_block_n(tiles_m, N, gran, out_f32, total_kb);
    b_split = w_split(out_f32, block_n, tiles_m, total_kb);
  }
  VX_REQUIRE(block_n >= gran && block_n % gran == 0 && block_n <= 256 && N % block_n == 0,
             "vx_gemm_bf16: block_n=%d invalid for N=%d", block_n, N);
  CUtensorMap mA, mA2, mB, mR, mC;
  {
    uint64_t dims[2] = {(uint64_t)K1, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {kBlockK, kBlockM};
    if (make_tmap_bf16(&mA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  if (K2 > 0) {
    uint64_t dims[2] = {(uint64_t)K2, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda2 * 2};
    uint32_t box[2] = {kBlockK, kBlockM};
    if (make_tmap_bf16(&mA2, A2, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  } else {
    mA2 = mA;
  }
  {
    uint64_t dims[2] = {(uint64_t)(K1 + K2), (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {kBlockK, (uint32_t)(b_split ? block_n / 2 : block_n)};
    if (make_tmap_bf16(&mB, Wt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  if (!out_f32) {
    if (make_out_maps(&mR, &mC, residual, ldr, out, ldc, M, geglu ? N / 2 : N, kBlockM)) return 1;
  } else {
    mR = mA;
    mC = mA;
  }
  GemmArgs a{};
  a.M = M; a.N = N;
  a.kblocks1 = (K1 + kBlockK - 1) / kBlockK;
  a.kblocks2 = (K2 + kBlockK - 1) / kBlockK;
  a.taps = 1;
  a.block_n = block_n;
  a.rows_valid = kBlockM;
  a.W = a.H = 1;
  a.tiles_m = tiles_m;
  a.tiles_n = N / block_n;
  a.geglu = geglu;
  a.has_residual = residual ? 1 : 0;
  a.out_f32 = out_f32;
  a.bias = bias; a.bias2 = bias2; a.bias2_div = bias2_div > 0 ? bias2_div : 1; a.scale = scale;
  a.out32 = (float*)out; a.ldc = ldc;
  a.ln_stats = ln_stats; a.ln_colsum = ln_colsum;
  a.ares = ares ? 1 : 0;
  a.ares_bytes = ares ? a.kblocks1 * kBlockM * kBlockK * 2 : 0;
  a.ln_eps = ln_eps;
  if (rs && rs->out) {
    VX_REQUIRE(!out_f32 && !geglu && rs->out_stride >= M, "vx_gemm_rowsums_bf16: linear bf16 epilogue only, stride >= M");
    VX_REQUIRE(2 * a.tiles_n <= rs->out_cap, "vx_gemm_rowsums_bf16: %d partial slots needed, %d allocated", 2 * a.tiles_n, rs->out_cap);
    a.rs_out = reinterpret_cast<float2*>(rs->out);
    a.rs_stride = rs->out_stride;
    if (rs->nparts_out) *rs->nparts_out = 2 * a.tiles_n;
  }
  if (rs && rs->parts) {
    VX_REQUIRE(ln_colsum && K2 == 0 && rs->nparts > 0 && rs->parts_stride >= M, "vx_gemm_lnparts_bf16: bad statistics operands");
    a.ln_parts = reinterpret_cast<const float2*>(rs->parts);
    a.ln_pstride = rs->parts_stride;
    a.ln_nparts = rs->nparts;
    a.ln_invK = 1.0f / (float)K1;
  }
  return launch(mA, mA2, mB, mR, mC, a, (cudaStream_t)stream);
}

extern "C" int vx_gemm_bf16(const void* A, long long lda, int K1, const void* A2, long long lda2, int K2,
                            const void* Wt, long long ldw, int M, int N, const float* bias, const float* bias2,
                            int bias2_div, float scale, const void* residual, long long ldr, void* out,
                            long long ldc, int out_f32, int block_n, void* stream) {
  return gemm_entry(A, lda, K1, A2, lda2, K2, Wt, ldw, M, N, bias, bias2, bias2_div, scale, residual, ldr, out, ldc,
                    out_f32, block_n, nullptr, nullptr, stream);
}

// LayerNorm folded into the GEMM: A holds the UN-normalised rows, Wt = W * gamma (per input channel),
// out = rstd[m] * (A @ Wt^T - mean[m] * colsum) + bias (+ bias2, scale, residual, GEGLU as in vx_gemm_bf16) with
// stats[m] = (mean, rstd) from vx_row_stats, colsum[n] = sum_k Wt[n,k], bias[n] = sum_k beta[k] W[n,k] + b[n].
extern "C" int vx_gemm_lnfold_bf16(const void* A, long long lda, int K, const void* Wt, long long ldw, int M, int N,
                                   const float* stats, const float* colsum, const float* bias, const float* bias2,
                                   int bias2_div, float scale, const void* residual, long long ldr, void* out,
                                   long long ldc, int geglu, int block_n, void* stream) {
  VX_REQUIRE(stats && colsum, "vx_gemm_lnfold_bf16: stats / colsum missing (M=%d)", M);
  return gemm_entry(A, lda, K, nullptr, 0, 0, Wt, ldw, M, N, bias, bias2, bias2_div, scale, residual, ldr, out, ldc,
                    geglu ? 2 : 0, block_n, stats, colsum, stream);
}

// Producer side of the LayerNorm hand-over: vx_gemm_bf16's linear epilogue (bias, bias2, scale, residual; bf16 out) that also
// writes, per output row, 2 * ceil(N / block_n) partial (sum, sum of squares) pairs of the ROUNDED outputs:
// row_parts[slot * parts_stride + m] as float2, slot < *nparts.  The consumer GEMM (vx_gemm_lnparts_bf16) turns them into
// mean / rstd, so LayerNorm(out) needs neither a normalisation pass nor a statistics pass over `out`.
extern "C" int vx_gemm_rowsums_bf16(const void* A, long long lda, int K1, const void* A2, long long lda2, int K2,
                                    const void* Wt, long long ldw, int M, int N, const float* bias, const float* bias2,
                                    int bias2_div, float scale, const void* residual, long long ldr, void* out,
                                    long long ldc, int block_n, float* row_parts, long long parts_stride, int parts_cap,
                                    int* nparts, void* stream) {
  VX_REQUIRE(row_parts && nparts, "vx_gemm_rowsums_bf16: row_parts / nparts missing (M=%d)", M);
  RowStatsIO rs{row_parts, parts_stride, parts_cap, nparts, nullptr, 0, 0};
  return gemm_entry(A, lda, K1, A2, lda2, K2, Wt, ldw, M, N, bias, bias2, bias2_div, scale, residual, ldr, out, ldc, 0,
                    block_n, nullptr, nullptr, stream, 0.f, &rs);
}

// Consumer side: vx_gemm_lnfold_bf16 with the row statistics taken from a producer's partial sums instead of a
// vx_row_stats array: mean = sum / K, variance = sum of squares / K - mean^2 (fp32), rstd = rsqrt(variance + eps).
extern "C" int vx_gemm_lnparts_bf16(const void* A, long long lda, int K, const void* Wt, long long ldw, int M, int N,
                                    const float* row_parts, long long parts_stride, int nparts, float eps,
                                    const float* colsum, const float* bias, const float* bias2, int bias2_div, float scale,
                                    const void* residual, long long ldr, void* out, long long ldc, int geglu, int block_n,
                                    void* stream) {
  VX_REQUIRE(row_parts && colsum, "vx_gemm_lnparts_bf16: row_parts / colsum missing (M=%d)", M);
  RowStatsIO rs{nullptr, 0, 0, nullptr, row_parts, parts_stride, nparts};
  return gemm_entry(A, lda, K, nullptr, 0, 0, Wt, ldw, M, N, bias, bias2, bias2_div, scale, residual, ldr, out, ldc,
                    geglu ? 2 : 0, block_n, nullptr, colsum, stream, eps, &rs);
}

// LayerNorm -> Linear in ONE kernel: A holds the un-normalised rows (K <= 512), Wt / colsum / bias are the folded parameters
// of vx_gemm_lnfold_bf16; the row statistics are computed inside the kernel from the shared-memory resident row tile
// (two-pass mean / variance, eps as in nn.LayerNorm), so neither LayerNorm(A) nor a statistics array ever exists in HBM.
extern "C" int vx_gemm_ln_bf16(const void* A, long long lda, int K, const void* Wt, long long ldw, int M, int N,
                               const float* colsum, const float* bias, float eps, const float* bias2, int bias2_div,
                               float scale, const void* residual, long long ldr, void* out, long long ldc, int geglu,
                               int block_n, void* stream) {
  VX_REQUIRE(colsum, "vx_gemm_ln_bf16: colsum missing (M=%d)", M);
  return gemm_entry(A, lda, K, nullptr, 0, 0, Wt, ldw, M, N, bias, bias2, bias2_div, scale, residual, ldr, out, ldc,
                    geglu ? 2 : 0, block_n, nullptr, colsum, stream, eps);
}

// X: NHWC bf16 [NB, Hin, Win, C];  Wt: [Cout, 9*C] with K index = (ky*3+kx)*C + c;  out: [NB*H*W, ldc], H x W the output
// size.  stride 2: the A boxes are fetched through a tensor map with traversal stride 2 along x and y (the box covers
// 2 * wbox x 2 * hbox input pixels, every second one lands in shared memory), so the tile of output pixels (y, x) gets
// input pixels (2y + dy, 2x + dx) for tap (dy, dx) without any gathered copy of the input (no im2col tensor).
static int conv3x3_entry(const void* X, int NB, int Hin, int Win, int C, const void* Wt, int Cout, const float* bias,
                         const float* bias2, int bias2_div, float scale, const void* residual, long long ldr, void* out,
                         long long ldc, int block_n, int stride, int pad_lo, void* stream) {
  VX_REQUIRE(C % kBlockK == 0 && Cout % 32 == 0, "vx_conv3x3_bf16: C=%d must be %%64, Cout=%d %%32", C, Cout);
  VX_REQUIRE(ldc % 8 == 0 && (!residual || ldr % 8 == 0), "vx_conv3x3_bf16: ld must be %%8");
  VX_REQUIRE(stride == 1 || (stride == 2 && Hin % 2 == 0 && Win % 2 == 0), "vx_conv3x3: stride %d on %dx%d unsupported", stride,
             Hin, Win);
  VX_REQUIRE(pad_lo == 1 || (pad_lo == 0 && stride == 2), "vx_conv3x3: pad_lo=%d only with stride 2", pad_lo);
  const int H = Hin / stride, W = Win / stride;   // 3x3, pad 1 (or (0,1,0,1)), even sizes: H_out = H_in / stride
  // pixel rectangle of <= 128 output rows that is contiguous in NHWC row order
  int wbox, hbox = 1, nbox = 1;
  if (W >= kBlockM) {
    wbox = kBlockM;                 // widest divisor of W that fits a tile (192 -> 96 at the 768x768 VAE level)
    while (W % wbox) --wbox;
  } else {
    wbox = W;
    hbox = kBlockM / W;
    if (hbox > H) {
      hbox = H;
      nbox = kBlockM / (W * H);
      if (nbox > NB) nbox = NB;
      if (nbox < 1) nbox = 1;
      while (NB % nbox) --nbox;
    } else {
      while (H % hbox) --hbox;
    }
  }
  const int rows_valid = wbox * hbox * nbox;
  const long long M = (long long)NB * H * W;
  VX_REQUIRE(M % rows_valid == 0, "vx_conv3x3_bf16: NB*H*W=%lld not tileable by %d", M, rows_valid);
  const long long tiles_m = M / rows_valid;
  const int total_kb = 9 * (C / kBlockK);
  if (block_n <= 0) block_n = gemm_env().bn;
  if (block_n <= 0) block_n = pick_block_n(tiles_m, Cout, 32, 0, total_kb);
  VX_REQUIRE(block_n % 32 == 0 && block_n >= 32 && block_n <= 256 && Cout % block_n == 0,
             "vx_conv3x3_bf16: block_n=%d invalid for Cout=%d", block_n, Cout);
  // Row reuse: when a tile is hbox >= 2 whole image rows of one frame, ONE box of hbox + 2 rows per (dx, channel block)
  // serves the three dy taps (each tap's 128 rows start dy image rows further down, a multiple of the swizzle atom):
  // 3 (hbox + 2) / (9 hbox) of the A traffic.  The 3x3 convs run at the L2 -> SM cap (~13 TB/s,
  // profiles/r02_roofline.csv), so operand bytes are what they are bound by.
  bool rr = gemm_env().conv_rr && stride == 1 && nbox == 1 && wbox == W && hbox >= 2 && rows_valid == kBlockM && (W * 128) % 1024 == 0;
  if (rr) {   // two ring stages + one staging tile must fit
    const int cgx = pair_ok(0, block_n, tiles_m, total_kb) ? 2 : 1;
    const size_t st = (size_t)(hbox + 2) * W * kBlockK * 2 + (size_t)3 * (block_n / cgx) * kBlockK * 2;
    if (2 * st + (size_t)(block_n / kPanelCols) * kPanelBytes > (size_t)227 * 1024 - 2048) rr = false;
  }
  CUtensorMap mA, mB, mR, mC;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Win, (uint64_t)Hin, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)Win * C * 2, (uint64_t)Hin * Win * C * 2};
    uint32_t box[4] = {kBlockK, (uint32_t)(wbox * stride), (uint32_t)((rr ? hbox + 2 : hbox) * stride), (uint32_t)nbox};
    uint32_t est[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (make_tmap_bf16(&mA, X, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, est)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * C, (uint64_t)Cout};
    uint64_t str[1] = {(uint64_t)9 * C * 2};
    uint32_t box[2] = {kBlockK, (uint32_t)(w_split(0, block_n, tiles_m, total_kb) ? block_n / 2 : block_n)};
    if (make_tmap_bf16(&mB, Wt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  if (make_out_maps(&mR, &mC, residual, ldr, out, ldc, M, Cout, rows_valid)) return 1;
  GemmArgs a{};
  a.M = (int)M; a.N = Cout;
  a.kblocks1 = C / kBlockK;
  a.kblocks2 = 0;
  a.taps = 9;
  a.rr = rr ? 1 : 0;
  a.a_bytes = rr ? (hbox + 2) * W * kBlockK * 2 : kBlockM * kBlockK * 2;
  a.block_n = block_n;
  a.rows_valid = rows_valid;
  a.W = W; a.H = H;
  a.cstride = stride; a.cpad = pad_lo;
  a.tiles_m = (int)tiles_m;
  a.tiles_n = Cout / block_n;
  a.geglu = 0;
  a.has_residual = residual ? 1 : 0;
  a.out_f32 = 0;
  a.bias = bias; a.bias2 = bias2; a.bias2_div = bias2_div > 0 ? bias2_div : 1; a.scale = scale;
  a.out32 = (float*)out; a.ldc = ldc;
  return launch(mA, mA, mB, mR, mC, a, (cudaStream_t)stream);
}

extern "C" int vx_conv3x3_bf16(const void* X, int NB, int H, int W, int C, const void* Wt, int Cout,
                               const float* bias, const float* bias2, int bias2_div, float scale,
                               const void* residual, long long ldr, void* out, long long ldc, int block_n,
                               void* stream) {
  return conv3x3_entry(X, NB, H, W, C, Wt, Cout, bias, bias2, bias2_div, scale, residual, ldr, out, ldc, block_n, 1, 1, stream);
}

// 3x3 convolution with stride 2 on an even-sized NHWC image -> [NB * (H/2) * (W/2), ldc].  pad_lo = 1: nn.Conv2d(padding=1)
// (Downsample3D / Downsample2D of the UNets, reference modules/resnet.py:93-120); pad_lo = 0: F.pad(x, (0, 1, 0, 1)) + padding 0
// (diffusers Downsample2D(padding=0) of the VAE encoder).  Out-of-image taps are zero-filled by the TMA unit.
extern "C" int vx_conv3x3s2_bf16(const void* X, int NB, int H, int W, int C, const void* Wt, int Cout, const float* bias,
                                 int pad_lo, void* out, long long ldc, int block_n, void* stream) {
  return conv3x3_entry(X, NB, H, W, C, Wt, Cout, bias, nullptr, 1, 1.0f, nullptr, 0, out, ldc, block_n, 2, pad_lo, stream);
}

// conv3x3(nearest_upsample_2x(X)) without the upsampled tensor (reference modules/resnet.py:53-90 Upsample3D,
// diffusers Upsample2D in the VAE decoder): output pixel (2i + py, 2j + px) only ever sees the 2x2 input neighbourhood
// rows {i + py - 1, i + py}, columns {j + px - 1, j + px}, so each of the four output parity classes is a 2x2 convolution of
// X with the 3x3 weights summed over the taps that land on the same input pixel (host: vexpress_b200.ops.pack_upconv_weight)
// -- 4/9 of the FLOPs, and the 4x tensor is never written or read.
// X: NHWC bf16 [NB, H, W, C];  Wt: [4 * Cout, 4 * C] = parity-major (py, px) blocks, K index = (a * 2 + b) * C + c;
// out: [NB * 2H * 2W, ldc] (NHWC of the upsampled image).
extern "C" int vx_upconv3x3_bf16(const void* X, int NB, int H, int W, int C, const void* Wt, int Cout, const float* bias,
                                 void* out, long long ldc, int block_n, void* stream) {
  VX_REQUIRE(C % kBlockK == 0 && Cout % 32 == 0, "vx_upconv3x3_bf16: C=%d must be %%64, Cout=%d %%32", C, Cout);
  VX_REQUIRE(ldc % 8 == 0, "vx_upconv3x3_bf16: ldc must be %%8");
  int wbox, hbox = 1, nbox = 1;
  if (W >= kBlockM) {
    wbox = kBlockM;                 // widest divisor of W that fits a tile (192 -> 96 at the 768x768 VAE level)
    while (W % wbox) --wbox;
  } else {
    wbox = W;
    hbox = kBlockM / W;
    if (hbox > H) {
      hbox = H;
      nbox = kBlockM / (W * H);
      if (nbox > NB) nbox = NB;
      if (nbox < 1) nbox = 1;
      while (NB % nbox) --nbox;
    } else {
      while (H % hbox) --hbox;
    }
  }
  const int rows_valid = wbox * hbox * nbox;
  const long long M = (long long)NB * H * W;
  VX_REQUIRE(M % rows_valid == 0, "vx_upconv3x3_bf16: NB*H*W=%lld not tileable by %d", M, rows_valid);
  const long long tiles_m = M / rows_valid;
  const int total_kb = 4 * (C / kBlockK);
  if (block_n <= 0) block_n = gemm_env().bn;
  if (block_n <= 0) block_n = pick_block_n(tiles_m * 4, Cout, 32, 0, total_kb);
  VX_REQUIRE(block_n % 32 == 0 && block_n >= 32 && block_n <= 256 && Cout % block_n == 0,
             "vx_upconv3x3_bf16: block_n=%d invalid for Cout=%d", block_n, Cout);
  const bool pair = w_split(0, block_n, tiles_m, total_kb);   // half-tile W boxes (CTA pair or W multicast)
  CUtensorMap mA, mB, mC0, mC1;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {kBlockK, (uint32_t)wbox, (uint32_t)hbox, (uint32_t)nbox};
    if (make_tmap_bf16(&mA, X, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)4 * C, (uint64_t)4 * Cout};
    uint64_t str[1] = {(uint64_t)4 * C * 2};
    uint32_t box[2] = {kBlockK, (uint32_t)(pair ? block_n / 2 : block_n)};
    if (make_tmap_bf16(&mB, Wt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  for (int px = 0; px < 2; ++px) {
    // (c, j, i, n, py) view of the [NB, 2H, 2W, ldc] output for one column parity
    uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)NB, 2};
    uint64_t str[4] = {(uint64_t)2 * ldc * 2, (uint64_t)2 * (2 * W) * ldc * 2, (uint64_t)(2 * H) * (2 * W) * ldc * 2,
                       (uint64_t)(2 * W) * ldc * 2};
    uint32_t box[5] = {kPanelCols, (uint32_t)wbox, (uint32_t)hbox, (uint32_t)nbox, 1};
    if (make_tmap_bf16(px ? &mC1 : &mC0, (const __nv_bfloat16*)out + (long long)px * ldc, 5, dims, str, box,
                       CU_TENSOR_MAP_SWIZZLE_64B))
      return 1;
  }
  GemmArgs a{};
  a.M = (int)M; a.N = Cout;
  a.kblocks1 = C / kBlockK;
  a.kblocks2 = 0;
  a.taps = 4;
  a.ups = 1;
  a.cstride = 1; a.cpad = 1;
  a.block_n = block_n;
  a.rows_valid = rows_valid;
  a.W = W; a.H = H;
  a.tiles_m = (int)tiles_m;
  a.tiles_n = Cout / block_n;
  a.geglu = 0;
  a.has_residual = 0;
  a.out_f32 = 0;
  a.bias = bias; a.bias2 = nullptr; a.bias2_div = 1; a.scale = 1.0f;
  a.out32 = (float*)out; a.ldc = ldc;
  return launch(mA, mA, mB, mC1, mC0, a, (cudaStream_t)stream);
}
