// tcgen05 flash attention for the spatial attentions of the denoising UNet (sm_100a).
//
//   O[b, q, h, :] = softmax_k( Q[b, q, h, :] . K[b / kv_div, k, h, :] * hd^-0.5 ) @ V[b / kv_div, k, h, :]
//
// covers attn1 (self-attention, kv_div = 1; reference modules/mutual_self_attention.py:176-186) and attn1_5
// (reference attention: K/V projected from the ReferenceNet bank, one bank per CFG half shared by the f frames
// of the window, kv_div = f; :202-219) -- both diffusers AttnProcessor2_0 / F.scaled_dot_product_attention
// without mask.
//
// One CTA = 128 query rows of one (frame, head).  Per 128-key tile:
//   TMA   : K_j, V_j tiles -> smem ring (3-D view (8, rows, C/8) of the [rows, C] token matrix, so a head slice
//           lands as [hd/8][rows][8] = the no-swizzle core-matrix layout; hd = 40 is zero-padded to 48 in smem
//           only, never in HBM)
//   MMA   : S_j = Q K_j^T  (tcgen05.mma, M=128, N=kv_tile, K=hd)     -> TMEM (double buffered)
//   softmax warps (1 thread = 1 query row): tcgen05.ld S, online max / exp2 / row sum, P_j (bf16) -> smem,
//           lazy rescale of the O accumulator (only when the running max grew by > 2^8)
//   MMA   : O += P_j V_j  (A = P from smem, B = V tile MN-major)       -> TMEM
// Epilogue: O / l -> bf16 -> out[row, h*hd : (h+1)*hd].
#include "vx_host.h"
#include "vx_ptx.cuh"

namespace vx {

constexpr int kFaThreads = 192;
constexpr int kFaStages = 3;

struct FaArgs {
  int Nq, Nk, hd, hdp, kv_tile, kv_div, heads;
  float scale_log2;
  __nv_bfloat16* out;
  long long ldo;
  int q_bytes, kv_bytes, p_bytes;  // smem tile sizes (padded)
};

__global__ void __launch_bounds__(kFaThreads, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                  const __grid_constant__ CUtensorMap mapV, const FaArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + p.q_bytes;                     // kFaStages x kv_bytes
  uint8_t* sV = sK + kFaStages * p.kv_bytes;        // kFaStages x kv_bytes
  uint8_t* sP = sV + kFaStages * p.kv_bytes;        // 2 x p_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * p.p_bytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // stages
  uint64_t* kv_empty = kv_full + kFaStages;
  uint64_t* s_full = kv_empty + kFaStages;  // 2
  uint64_t* p_ready = s_full + 2;           // 2
  uint64_t* pv_done = p_ready + 2;          // 2
  uint64_t* o_full = pv_done + 2;           // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, bq = blockIdx.z;
  const int T = p.Nk / p.kv_tile;

  // zero the operand tiles once: the K-padding chunk (hd 40 -> 48) must read as exact zeros
  {
    const int total16 = (p.q_bytes + 2 * kFaStages * p.kv_bytes) / 16;
    uint4* z = reinterpret_cast<uint4*>(sQ);
    for (int i = threadIdx.x; i < total16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kFaStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 128);
      mbar_init(&pv_done[s], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();  // generic-proxy zero fill visible to TMA / tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // shared-memory fill, barrier init and the TMEM allocation overlapped the previous grid's tail
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&mapQ);
      tma_prefetch_desc(&mapK);
      tma_prefetch_desc(&mapV);
      const int col_chunk = head * p.hd / 8;
      mbar_expect_tx(q_full, 128 * p.hd * 2);
      tma_load_3d(sQ, &mapQ, q_full, 0, bq * p.Nq + q_tile * 128, col_chunk);
      const int kv_row0 = (bq / p.kv_div) * p.Nk;
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < T; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&kv_full[stage], 2 * p.kv_tile * p.hd * 2);
        tma_load_3d(sK + stage * p.kv_bytes, &mapK, &kv_full[stage], 0, kv_row0 + j * p.kv_tile, col_chunk);
        tma_load_3d(sV + stage * p.kv_bytes, &mapV, &kv_full[stage], 0, kv_row0 + j * p.kv_tile, col_chunk);
        if (++stage == kFaStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)p.kv_tile, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, (uint32_t)p.hdp, 0, 1);  // B = V is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t lbo_k = (uint32_t)p.kv_tile * 16;
      auto issue_s = [&](int j) {
        const int stage = j % kFaStages;
        mbar_wait(&kv_full[stage], (uint32_t)((j / kFaStages) & 1));
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + stage * p.kv_bytes);
        const uint32_t d = tmem_base + (uint32_t)((j & 1) * 128);
        for (int k = 0; k < p.hdp / 16; ++k) {
          const uint64_t da = make_smem_desc(q_addr + k * 2 * 2048, 2048, 128, SWZ_NONE);
          const uint64_t db = make_smem_desc(k_addr + k * 2 * lbo_k, lbo_k, 128, SWZ_NONE);
          umma_ss(d, da, db, idesc_s, k ? 1u : 0u);
        }
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        mbar_wait(&p_ready[j & 1], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const int stage = j % kFaStages;
        const uint32_t p_addr = smem_u32(sP + (j & 1) * p.p_bytes);
        const uint32_t v_addr = smem_u32(sV + stage * p.kv_bytes);
        for (int k = 0; k < p.kv_tile / 16; ++k) {
          const uint64_t da = make_smem_desc(p_addr + k * 2 * 2048, 2048, 128, SWZ_NONE);
          const uint64_t db = make_smem_desc(v_addr + k * 256, 128, lbo_k, SWZ_NONE);
          umma_ss(tmem_o, da, db, idesc_o, (j | k) ? 1u : 0u);
        }
        umma_commit(&kv_empty[stage]);
        umma_commit(&pv_done[j & 1]);
      }
      umma_commit(o_full);
    }
  } else {
    // ------------------------------------------------------------ softmax + correction + epilogue (warps 2..5)
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    float m_used = -INFINITY;  // max (in raw score units) the accumulators are currently scaled by
    float l = 0.f;
    const float c = p.scale_log2;
    for (int j = 0; j < T; ++j) {
      mbar_wait(&s_full[j & 1], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint32_t ts = tmem_base + lane_addr + (uint32_t)((j & 1) * 128);
      float mx = -INFINITY;
      for (int cb = 0; cb < p.kv_tile; cb += 16) {
        uint32_t v[16];
        tmem_ld16(ts + cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      if (j > 0) {
        mbar_wait(&pv_done[(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1));
        tc_fence_after();
      }
      const float m_new = fmaxf(m_used, mx);
      const bool need = (m_new - m_used) * c > 8.0f;  // also true for the first tile (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = (m_used == -INFINITY) ? 0.f : exp2f((m_used - m_new) * c);
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          for (int cb = 0; cb < p.hdp; cb += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_o + lane_addr + cb, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(tmem_o + lane_addr + cb, v);
          }
          tmem_st_wait();
        }
      }
      const float mc = m_used * c;
      uint8_t* pb = sP + (j & 1) * p.p_bytes + row * 16;
      for (int cb = 0; cb < p.kv_tile; cb += 16) {
        uint32_t v[16];
        tmem_ld16(ts + cb, v);
        tmem_ld_wait();
        float e[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          e[i] = exp2f(__uint_as_float(v[i]) * c - mc);
          l += e[i];
        }
        *reinterpret_cast<uint4*>(pb + (cb / 8) * 2048) =
            make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
        *reinterpret_cast<uint4*>(pb + (cb / 8 + 1) * 2048) = make_uint4(
            pack_bf16(e[8], e[9]), pack_bf16(e[10], e[11]), pack_bf16(e[12], e[13]), pack_bf16(e[14], e[15]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_ready[j & 1]);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int qrow = q_tile * 128 + row;
    const float inv = 1.f / l;
    __nv_bfloat16* op = p.out + ((long long)bq * p.Nq + qrow) * p.ldo + head * p.hd;
    for (int cb = 0; cb < p.hdp; cb += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_o + lane_addr + cb, v);
      tmem_ld_wait();
      if (qrow < p.Nq) {
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          if (cb + h8 * 8 < p.hd) {
            const int o = h8 * 8;
            *reinterpret_cast<uint4*>(op + cb + o) = make_uint4(
                pack_bf16(__uint_as_float(v[o]) * inv, __uint_as_float(v[o + 1]) * inv),
                pack_bf16(__uint_as_float(v[o + 2]) * inv, __uint_as_float(v[o + 3]) * inv),
                pack_bf16(__uint_as_float(v[o + 4]) * inv, __uint_as_float(v[o + 5]) * inv),
                pack_bf16(__uint_as_float(v[o + 6]) * inv, __uint_as_float(v[o + 7]) * inv));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// v2: 256 query rows per CTA as two 128-row tiles.  Measured on B200 (profiles/r01c_flash2_ncu_full.md and
// profiles/tools/fa_ablation.py) the hd = 40 kernel is bound neither by the tensor pipe (19 % active) nor by HBM
// but by the instruction stream of the softmax warps, so this version maximises softmax thread-level parallelism:
//   * 16 softmax warps (4 per SM sub-partition): every 128x128 S tile is split into two 64-column halves owned by
//     two warps of the same TMEM lane quadrant; they exchange the partial row maxima through shared memory;
//   * one MMA-issuing warp per query tile; S(q, j+1) is issued as soon as S(q, j) sits in registers; P is double
//     buffered so the softmax never waits for P.V (only the rare lazy O-rescale does);
//   * half of the exponentials run on the FMA pipe (ex2_poly) to balance the MUFU unit.
// Used when hd <= 128, Nq % 256 == 0 and Nk % 128 == 0 (the 64x64 and 32x32 levels, ~98 % of the attention FLOPs).
constexpr int kFa2Threads = 608;  // warps 0..15 softmax, 16 TMA, 17/18 MMA (query tile 0/1)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max rel. error 8e-5 -- far below the
// bf16 rounding of P).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;            // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float j = t - 12582912.0f;
  const float r = x - j;                      // in [-0.5, 0.5]
  float pl = fmaf(0.05519810691475868f, r, 0.24267712235450745f);
  pl = fmaf(pl, r, 0.6932618021965027f);
  pl = fmaf(pl, r, 0.9999227523803711f);
  return __uint_as_float(__float_as_uint(pl) + (__float_as_uint(t) << 23));   // scale by 2^round(x)
}
__device__ __forceinline__ void wg_bar_sync(int q) { asm volatile("bar.sync %0, 256;" ::"r"(q + 1) : "memory"); }
// named barrier of the two warps (column halves) that share the rows of one TMEM lane quadrant: ids 3..10
// baton between the two query tiles (256 softmax threads each): ids 11 / 12, 512 = 256 waiting + 256 arriving threads
__device__ __forceinline__ void baton_wait(int q) { asm volatile("bar.sync %0, 512;" ::"r"(11 + q) : "memory"); }
__device__ __forceinline__ void baton_pass(int q) { asm volatile("bar.arrive %0, 512;" ::"r"(12 - q) : "memory"); }
__device__ __forceinline__ void pair_bar_sync(int q, int qd) { asm volatile("bar.sync %0, 64;" ::"r"(3 + q * 4 + qd) : "memory"); }

struct Fa2Args {
  int Nq, Nk, hd, hdp, kv_div, stages;
  int pbufs;              // P tiles per query tile in shared memory (smem-P path)
  int p_tmem;             // 1: P lives in tensor memory and P.V is a TS-MMA (hd <= 64): no smem traffic for P
  float scale_log2;
  __nv_bfloat16* out;
  long long ldo;
  int q_bytes, kv_bytes;  // per 128-row tile
  long long* trace;       // bring-up only (VX_FA_TRACE): clock64 stamps of CTA (0,0,0)
  int stagger;            // clocks by which the softmax warps of query tile 1 start late (see the softmax loop)
  int pair_sync;          // exchange the half-row maxima under a 64-thread pair barrier instead of the 256-thread tile barrier
  int late_wait;          // p_tmem: wait for P.V(j-1) only right before P(j) is stored, not before the exponentials
};

#define FA_TR(slot) do { if (tr) tr[(j) * 16 + (slot)] = clock64(); } while (0)

// ONES: the head dim is padded (hd 40 -> 48) and column `hd` of every V tile is set to 1.0, so the P.V product also
// accumulates the softmax row sum in O[:, hd] (with exactly the bf16-rounded P it multiplies V with); the 64
// per-element FADDs of the row sum disappear from the issue-bound softmax loop.
// BATON (experiment, VX_FA_BATON=1, not yet run on hardware): the exponential phases of the two query tiles take turns
// through a pair of named barriers.  Reading of the clock64 timeline + A/B sweeps (profiles/r01f_final_ncu.md): the two
// tiles run in lock-step -- both groups of softmax warps sit in their MUFU-bound exponential phase together (~1550
// clk), then both tiles' MMAs queue on the tensor pipe together (~1850 clk), and the two costs add up to the 3400-clk
// iteration instead of overlapping; a one-off start offset does not survive.  With the baton one tile's exponentials
// (alone on the MUFU units: ~900 clk) overlap the other tile's S load / row max / MMAs by construction.
// PMASK: element i of every 8 takes the FMA-pipe exponential when (i & PMASK) == PMASK: 7 -> 1/8 (measured best without
// the baton), 3 -> 1/4, 1 -> 1/2 (to be swept with the baton, where one tile owns the MUFU units during its phase).
template <bool ONES, bool BATON, int PMASK>
__global__ void __launch_bounds__(kFa2Threads, 1)
flash_attn2_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV, const Fa2Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  constexpr int kP = 128 * 128 * 2;  // P tile bytes
  uint8_t* sQ = smem;                               // 2 x q_bytes
  uint8_t* sK = sQ + 2 * p.q_bytes;                 // stages x kv_bytes
  uint8_t* sV = sK + p.stages * p.kv_bytes;         // stages x kv_bytes
  uint8_t* sP = sV + p.stages * p.kv_bytes;         // 2 query tiles x pbufs x 32 KB
  float* xchg = reinterpret_cast<float*>(sP + (p.p_tmem ? 0 : 2 * p.pbufs * kP));  // 3 slots x [2 q][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 3 * 512);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;             // [stages <= 6]
  uint64_t* kv_empty = kv_full + 6;         // [6]
  uint64_t* s_full = kv_empty + 6;          // [2]
  uint64_t* p_ready = s_full + 2;           // [2]
  uint64_t* pv_done = p_ready + 2;          // [2 query tiles][2 P buffers]
  uint64_t* o_full = pv_done + 4;           // [2]
  uint64_t* s_free = o_full + 2;            // [2] S(q) has been pulled into registers -> may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x, head = blockIdx.y, bq = blockIdx.z;
  const int T = p.Nk / 128;
  {
    const int total16 = (2 * p.q_bytes + 2 * p.stages * p.kv_bytes) / 16;
    uint4* z = reinterpret_cast<uint4*>(sQ);
    for (int i = threadIdx.x; i < total16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (ONES) {
    __syncthreads();
    // V tiles are [hdp/8 chunks][128 keys][8]: element (key, col hd) = chunk hd/8, slot hd%8; TMA never touches it
    for (int i = threadIdx.x; i < p.stages * 128; i += blockDim.x) {
      const int st = i / 128, key = i % 128;
      reinterpret_cast<__nv_bfloat16*>(sV + st * p.kv_bytes + (p.hd / 8) * 2048 + key * 16)[p.hd % 8] = __float2bfloat16(1.0f);
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 6; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 256);
      mbar_init(&pv_done[2 * s], 1);
      mbar_init(&pv_done[2 * s + 1], 1);
      mbar_init(&s_free[s], 256);
      mbar_init(&o_full[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // shared-memory fill, barrier init and the TMEM allocation overlapped the previous grid's tail

  if (warp == 16) {
    if (lane == 0) {
      tma_prefetch_desc(&mapQ);
      tma_prefetch_desc(&mapK);
      tma_prefetch_desc(&mapV);
      const int col_chunk = head * p.hd / 8;
      mbar_expect_tx(q_full, 2 * 128 * p.hd * 2);
      tma_load_3d(sQ, &mapQ, q_full, 0, bq * p.Nq + q_pair * 256, col_chunk);
      tma_load_3d(sQ + p.q_bytes, &mapQ, q_full, 0, bq * p.Nq + q_pair * 256 + 128, col_chunk);
      const int kv_row0 = (bq / p.kv_div) * p.Nk;
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < T; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&kv_full[stage], 2 * 128 * p.hd * 2);
        tma_load_3d(sK + stage * p.kv_bytes, &mapK, &kv_full[stage], 0, kv_row0 + j * 128, col_chunk);
        tma_load_3d(sV + stage * p.kv_bytes, &mapV, &kv_full[stage], 0, kv_row0 + j * 128, col_chunk);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 17) {
    // one MMA-issuing warp per query tile: the two tiles advance independently.  The whole warp runs the loop so
    // that all descriptor / barrier values stay warp-uniform; only the elected lane issues (a divergent
    // `if (lane == 0)` region costs an ELECT + R2UR.BROADCAST loop per UTCHMMA: measured ~150 cycles per MMA,
    // 4700 instead of ~2000 cycles per KV iteration).
    const bool leader = elect_one();
    const int q = warp - 17;
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, (uint32_t)p.hdp, 0, 1);
    const uint64_t dq = make_smem_desc(smem_u32(sQ + q * p.q_bytes), 2048, 128, SWZ_NONE);
    const uint32_t p_base = smem_u32(sP + q * p.pbufs * kP);
    const uint32_t d_s = tmem_base + (uint32_t)(q * 128);
    const uint32_t d_o = tmem_base + 256u + (uint32_t)(q * (p.p_tmem ? 64 : 128));
    const uint32_t t_p = tmem_base + 384u + (uint32_t)(q * 64);   // P tile in TMEM (p_tmem mode): 64 packed columns
    const int ksteps = p.hdp / 16;
    long long* tr = (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && q == 0 && leader) ? p.trace : nullptr;
    auto issue_s = [&](int j) {
      const uint64_t dk = make_smem_desc(smem_u32(sK + (j % p.stages) * p.kv_bytes), 2048, 128, SWZ_NONE);
      if (leader) {
        for (int k = 0; k < ksteps; ++k)   // +4096 bytes per K step = +256 in the (addr >> 4) field
          umma_ss(d_s, dq + (uint64_t)(k * 256), dk + (uint64_t)(k * 256), idesc_s, k ? 1u : 0u);
        umma_commit(&s_full[q]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) {
        mbar_wait(&kv_full[(j + 1) % p.stages], (uint32_t)(((j + 1) / p.stages) & 1));
        FA_TR(8);
        mbar_wait(&s_free[q], (uint32_t)(j & 1));   // S(q, j) is in registers -> overwrite it with S(q, j+1)
        tc_fence_after();
        FA_TR(9);
        issue_s(j + 1);
        FA_TR(10);
      }
      const int stage = j % p.stages;
      mbar_wait(&p_ready[q], (uint32_t)(j & 1));
      tc_fence_after();
      FA_TR(11);
      const int pb_i = j % p.pbufs;
      const uint64_t dp = make_smem_desc(p_base + (uint32_t)(pb_i * kP), 2048, 128, SWZ_NONE);
      const uint64_t dv = make_smem_desc(smem_u32(sV + stage * p.kv_bytes), 128, 2048, SWZ_NONE);
      if (leader) {
        if (p.p_tmem) {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // A = P from tensor memory (8 packed columns per 16 keys), B = V from smem
            umma_ts(d_o, t_p + (uint32_t)(k * 8), dv + (uint64_t)(k * 16), idesc_o, (j | k) ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // P: +4096 B (= +256) per 16 keys; V: +256 B (= +16) per 16 keys
            umma_ss(d_o, dp + (uint64_t)(k * 256), dv + (uint64_t)(k * 16), idesc_o, (j | k) ? 1u : 0u);
        }
        umma_commit(&kv_empty[stage]);
        umma_commit(&pv_done[2 * q + pb_i]);
      }
      __syncwarp();
      FA_TR(12);
    }
    if (leader) umma_commit(&o_full[q]);
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax: warp = (q, column half, lane quadrant)
    const int q = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const uint32_t ts = tmem_base + lane_addr + (uint32_t)(q * 128 + half * 64);
    const uint32_t to = tmem_base + lane_addr + 256u + (uint32_t)(q * (p.p_tmem ? 64 : 128));
    const uint32_t tp = tmem_base + lane_addr + 384u + (uint32_t)(q * 64 + half * 32);
    float* my_x = xchg + (q * 2 + half) * 128 + row;
    const float* other_x = xchg + (q * 2 + (half ^ 1)) * 128 + row;
    // O columns (16-wide chunks) this warp owns for the rescale and the epilogue
    const int ochunks = p.hdp / 16;
    const int oc_begin = half ? (ochunks + 1) / 2 : 0, oc_end = half ? ochunks : (ochunks + 1) / 2;
    float m_used = -INFINITY, l = 0.f;
    const float c = p.scale_log2;
    uint8_t* pb0 = sP + q * p.pbufs * kP + row * 16 + half * 8 * 2048;
    long long* tr = (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 0 && lane == 0) ? p.trace : nullptr;
    // The four softmax warps of an SM sub-partition (two per query tile) share one MUFU unit.  Started together, the
    // two query tiles run in lock-step: all four warps sit in their exponential phase at the same time and the unit
    // idles during everybody's load / max / barrier phases (clock64 timeline, profiles/tools/fa_trace.py).  Starting
    // tile 1 half a period late makes the exponential phase of one tile overlap the other phases of the other tile.
    if (q == 1 && p.stagger > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < p.stagger) {
      }
    }
    const bool late_wait = p.p_tmem && p.late_wait;
    if constexpr (BATON) {
      if (q == 1) baton_pass(1);          // tile 0 exponentiates first
    }
    for (int j = 0; j < T; ++j) {
      mbar_wait(&s_full[q], (uint32_t)(j & 1));
      tc_fence_after();
      FA_TR(0);
      uint32_t v[2][32];
      tmem_ld32(ts, v[0]);
      tmem_ld32(ts + 32, v[1]);
      tmem_ld_wait();
      FA_TR(1);
      tc_fence_before();
      mbar_arrive(&s_free[q]);
      float mxs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) mxs[i] = __uint_as_float(v[0][i]);
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int i = 0; i < 32; ++i) mxs[i & 3] = fmaxf(mxs[i & 3], __uint_as_float(v[g][i]));
      float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
      // combine with the other column half of the same rows (double-buffered slot: one barrier per tile suffices)
      my_x[(j & 1) * 512] = mx;
      if (p.pair_sync) pair_bar_sync(q, qd); else wg_bar_sync(q);
      mx = fmaxf(mx, other_x[(j & 1) * 512]);
      FA_TR(2);
      // the P buffer of this tile was last read by P.V of tile j - pbufs
      uint8_t* pb = pb0 + (j % p.pbufs) * kP;
      if (j >= p.pbufs && !late_wait) {
        mbar_wait(&pv_done[2 * q + j % p.pbufs], (uint32_t)((j / p.pbufs - 1) & 1));
        tc_fence_after();
      }
      FA_TR(3);
      const float m_new = fmaxf(m_used, mx);
      const bool need = (m_new - m_used) * c > 8.0f;  // identical in both warps of a pair (same rows, same maxima)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = (m_used == -INFINITY) ? 0.f : ex2_approx((m_used - m_new) * c);
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          // rescale O: needs P.V of tile j-1 complete (rare: only when the running max grew by > 2^8)
          mbar_wait(&pv_done[2 * q + (j - 1) % p.pbufs], (uint32_t)(((j - 1) / p.pbufs) & 1));
          tc_fence_after();
          for (int oc = oc_begin; oc < oc_end; ++oc) {
            uint32_t o[16];
            tmem_ld16(to + oc * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(to + oc * 16, o);
          }
          tmem_st_wait();
        }
      }
      const float mc = m_used * c;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (BATON) baton_wait(q);
      if (p.p_tmem) {
        uint32_t pk[32];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float xx = fmaf(__uint_as_float(v[g][h * 8 + i]), c, -mc);
              e[i] = ((i & PMASK) == PMASK) ? ex2_poly(xx) : ex2_approx(xx);   // share of the FMA pipe: see PMASK
              if (!ONES) ls[i & 3] += e[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) pk[g * 16 + h * 4 + i] = pack_bf16(e[2 * i], e[2 * i + 1]);
          }
        }
        FA_TR(6);
        if constexpr (BATON) {
          if (q == 0 || j + 1 < T) baton_pass(q);      // tile 1's last pass would have no taker
        }
        if (j >= p.pbufs && late_wait) {   // P(j) overwrites the tile P.V(j - pbufs) reads: wait only now
          mbar_wait(&pv_done[2 * q + j % p.pbufs], (uint32_t)((j / p.pbufs - 1) & 1));
          tc_fence_after();
        }
        FA_TR(7);
        tmem_st32(tp, pk);       // this warp's 64 keys = 32 packed columns of P(q)
        tmem_st_wait();
      } else {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float xx = fmaf(__uint_as_float(v[g][h * 8 + i]), c, -mc);
              e[i] = ((i & 3) == 3) ? ex2_poly(xx) : ex2_approx(xx);
              if (!ONES) ls[i & 3] += e[i];
            }
            *reinterpret_cast<uint4*>(pb + (g * 4 + h) * 2048) =
                make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
          }
        }
        if constexpr (BATON) {
          if (q == 0 || j + 1 < T) baton_pass(q);
        }
        fence_proxy_async_smem();
      }
      l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      FA_TR(4);
      tc_fence_before();
      mbar_arrive(&p_ready[q]);
      FA_TR(5);
    }
    mbar_wait(&o_full[q], 0);
    tc_fence_after();
    if (ONES) {
      // the row sum sits in O[:, hd]; the warp owning that chunk publishes it to its partner
      const int lc = p.hd / 16;
      if (lc >= oc_begin && lc < oc_end) {
        uint32_t o[16];
        tmem_ld16(to + lc * 16, o);
        tmem_ld_wait();
        float lv = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i == (p.hd & 15)) lv = __uint_as_float(o[i]);
        l = lv;
        my_x[1024] = l;
      }
      if (p.pair_sync) pair_bar_sync(q, qd); else wg_bar_sync(q);
      if (!(lc >= oc_begin && lc < oc_end)) l = other_x[1024];
    } else {
      // total row sum = the two column halves
      my_x[1024] = l;
      if (p.pair_sync) pair_bar_sync(q, qd); else wg_bar_sync(q);
      l += other_x[1024];
    }
    const int qrow = q_pair * 256 + q * 128 + row;
    const float inv = 1.f / l;
    __nv_bfloat16* op = p.out + ((long long)bq * p.Nq + qrow) * p.ldo + head * p.hd;
    for (int oc = oc_begin; oc < oc_end; ++oc) {
      const int cb = oc * 16;
      uint32_t o[16];
      tmem_ld16(to + cb, o);
      tmem_ld_wait();
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        if (cb + h8 * 8 < p.hd) {
          const int b8 = h8 * 8;
          *reinterpret_cast<uint4*>(op + cb + b8) = make_uint4(
              pack_bf16(__uint_as_float(o[b8]) * inv, __uint_as_float(o[b8 + 1]) * inv),
              pack_bf16(__uint_as_float(o[b8 + 2]) * inv, __uint_as_float(o[b8 + 3]) * inv),
              pack_bf16(__uint_as_float(o[b8 + 4]) * inv, __uint_as_float(o[b8 + 5]) * inv),
              pack_bf16(__uint_as_float(o[b8 + 6]) * inv, __uint_as_float(o[b8 + 7]) * inv));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// v3: small independent CTAs whose softmax warps never wait.  Measured on B200 (profiles/tools/ubench.cu,
// profiles/r02_flash_notes.md): the hd = 40 attention is bound by the softmax instruction stream -- 16 ex2/clk/SM on the
// MUFU pipe, ~14 elements/clk/SM for a whole LDTM -> row max -> fma/ex2/pack -> STTM warp iteration when the warps run
// free of each other (2300 clk per 2 x 128 x 128 scores) -- while v2's 16 softmax warps advance in lock-step (pair
// barriers for the split-row maxima, 256-thread P hand-offs, two tiles sharing one MUFU phase): every phase that is not the
// exponential one is exposed and an iteration costs 3450 clk.  A first v3 with a single S buffer and three CTAs per SM
// showed the same convoy ACROSS CTAs (all resident CTAs sat in their softmax -> P.V -> S turnaround together: 2990 clk).
// So the rule is: a softmax warp must always find its next S tile already computed.
//   * one CTA = ONE 128-row query tile, 64 keys per step; 4 softmax warps (one per TMEM lane quadrant, 1 thread = 1 row =
//     64 scores: no split rows, no max exchange, no named barriers), 1 TMA warp, 1 MMA warp;
//   * S is DOUBLE-buffered in tensor memory: S(j+1) = Q K_{j+1}^T is issued before softmax(j) starts, P.V(j) and S(j+2)
//     are issued when the last of the four warps has stored P(j) -- by then every warp is already inside softmax(j+1),
//     i.e. the tensor-core turnaround has a whole iteration of slack and never stalls a softmax warp;
//   * P (bf16, 32 packed columns) overwrites the S columns the thread has just read and feeds P.V as a TS-MMA straight
//     from tensor memory; tensor memory per CTA: 2 x 64 (S) + hdp (O) <= 256 columns -> two CTAs per SM.
// Used when hd <= 128, Nq % 128 == 0 and Nk % 64 == 0.
constexpr int kFa3Threads = 192;

struct Fa3Args {
  int dbg;                   // bring-up switches (VX_FA3_DBG), see the kernel
  int noload;                // experiment (VX_FA3_NOLOAD): K/V tiles are loaded for the first ring pass only -> time without K/V traffic
  int Nq, Nk, hd, hdp, kv_div, stages;
  float scale_log2;
  __nv_bfloat16* out;
  long long ldo;
  int q_bytes, kv_bytes;     // Q tile (128 rows) / K or V tile (64 keys), padded head dim
  uint32_t tmem_cols;        // power of two >= 64 + hdp
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <bool ONES, int PMASK, int MINB>
__global__ void __maxnreg__(MINB == 3 ? 112 : 168)     // MINB resident CTAs x 192 threads within the 64 K registers of an SM
flash_attn3_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV, const Fa3Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + p.q_bytes;
  uint8_t* sV = sK + p.stages * p.kv_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + p.stages * p.kv_bytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;          // [<= 8]
  uint64_t* kv_empty = kv_full + 8;      // [8]
  uint64_t* s_full = kv_empty + 8;       // [2]
  uint64_t* p_ready = s_full + 2;        // [2]
  uint64_t* pv_done = p_ready + 2;       // one phase per step: P.V(j) complete
  uint64_t* o_full = pv_done + 1;        // single phase: the LAST P.V complete -> O final
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, bq = blockIdx.z;
  const int T = p.Nk / 64;
  if (p.hdp > p.hd) {
    // the padding chunk of Q / K / V (hd 40 -> 48) must read as exact zeros; TMA only ever writes chunks < hd / 8
    const int total16 = (p.q_bytes + 2 * p.stages * p.kv_bytes) / 16;
    uint4* z = reinterpret_cast<uint4*>(sQ);
    for (int i = threadIdx.x; i < total16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    if (ONES) {
      __syncthreads();
      // V tiles are [hdp/8 chunks][64 keys][8]: element (key, col hd) = chunk hd/8, slot hd%8
      for (int i = threadIdx.x; i < p.stages * 64; i += blockDim.x) {
        const int st = i / 64, key = i % 64;
        reinterpret_cast<__nv_bfloat16*>(sV + st * p.kv_bytes + (p.hd / 8) * 1024 + key * 16)[p.hd % 8] = __float2bfloat16(1.0f);
      }
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 8; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 128);
    }
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // shared-memory fill, barrier init and the TMEM allocation overlapped the previous grid's tail
  const uint32_t tmem_o = tmem_base + 128u;     // S buffers at columns [0, 64) and [64, 128)

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer (warp-uniform loop, elected lane issues)
    const bool leader = elect_one();
    const int col_chunk = head * p.hd / 8;
    const int kv_row0 = (bq / p.kv_div) * p.Nk;
    if (leader) {
      tma_prefetch_desc(&mapQ);
      tma_prefetch_desc(&mapK);
      tma_prefetch_desc(&mapV);
      mbar_expect_tx(q_full, 128 * p.hd * 2);
      tma_load_3d(sQ, &mapQ, q_full, 0, bq * p.Nq + q_tile * 128, col_chunk);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < T; ++j) {
      mbar_wait(&kv_empty[stage], phase ^ 1);
      if (leader) {
        if (p.noload && j >= p.stages) {
          mbar_arrive(&kv_full[stage]);          // timing experiment: reuse what the ring already holds
        } else {
          mbar_expect_tx(&kv_full[stage], 2 * 64 * p.hd * 2);
          tma_load_3d(sK + stage * p.kv_bytes, &mapK, &kv_full[stage], 0, kv_row0 + j * 64, col_chunk);
          tma_load_3d(sV + stage * p.kv_bytes, &mapV, &kv_full[stage], 0, kv_row0 + j * 64, col_chunk);
        }
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, (uint32_t)p.hdp, 0, 1);   // B = V is MN-major
    const uint64_t dq = make_smem_desc(smem_u32(sQ), 2048, 128, SWZ_NONE);
    const int ksteps = p.hdp / 16;
    mbar_wait(q_full, 0);
    // S(t) = Q K_t^T into S buffer t & 1, the buffer P(t-2) was stored into: S(t) is issued once P.V(t-2) has COMPLETED
    // (pv_done), not merely been issued -- S (D = S buffer, N = 64) and P.V (D = O, N = hdp, A = P in tensor memory) differ in
    // accumulator and shape, and only same-accumulator, same-shape MMAs are documented to execute in issue order.  The wait
    // costs nothing measurable (the softmax of step t - 1 runs in between) and also pins pv_done's phase for the softmax
    // warps' parity wait below.
    auto issue_s = [&](int t) {
      const int st = t % p.stages;
      mbar_wait(&kv_full[st], (uint32_t)((t / p.stages) & 1));
      tc_fence_after();
      const uint64_t dk = make_smem_desc(smem_u32(sK + st * p.kv_bytes), 1024, 128, SWZ_NONE);
      const uint32_t d_s = tmem_base + (uint32_t)((t & 1) * 64);
      if (leader) {
        for (int k = 0; k < ksteps; ++k)   // Q: +4096 B (= +256) per 16 dims; K: 2 chunks of 64 keys = +2048 B (= +128)
          umma_ss(d_s, dq + (uint64_t)(k * 256), dk + (uint64_t)(k * 128), idesc_s, k ? 1u : 0u);
        umma_commit(&s_full[t & 1]);
      }
      __syncwarp();
      if (p.dbg & 2) mbar_wait(&s_full[t & 1], (uint32_t)((t >> 1) & 1));   // S(t) complete before anything else is issued
    };
    issue_s(0);
    if (T > 1) issue_s(1);
    for (int j = 0; j < T; ++j) {
      const int stage = j % p.stages;
      mbar_wait(&p_ready[j & 1], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      if (p.dbg & 32) {   // bring-up: let the P stores settle
        const long long t0 = clock64();
        while (clock64() - t0 < 3000) {
        }
      }
      const uint64_t dv = make_smem_desc(smem_u32(sV + stage * p.kv_bytes), 128, 1024, SWZ_NONE);
      const uint32_t t_p = tmem_base + (uint32_t)((j & 1) * 64);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k)   // A = P from tensor memory (8 packed columns per 16 keys); V: +256 B (= +16) per 16 keys
          umma_ts(tmem_o, t_p + (uint32_t)(k * 8), dv + (uint64_t)(k * 16), idesc_o, (j | k) ? 1u : 0u);
        if (p.dbg & 16) {   // bring-up: release the K/V stage one step late (after P.V(j), free the stage of step j - 1)
          if (j > 0) umma_commit(&kv_empty[(j - 1) % p.stages]);
          if (j == T - 1) umma_commit(&kv_empty[stage]);
        } else {
          umma_commit(&kv_empty[stage]);
        }
        umma_commit(pv_done);
        if (j == T - 1) umma_commit(o_full);
      }
      __syncwarp();
      if (j + 2 < T) {
        mbar_wait(pv_done, (uint32_t)(j & 1));   // P(j) has been consumed out of S buffer j & 1
        issue_s(j + 2);
      }
    }
    pdl_trigger();   // the last P.V is issued: the next grid may launch behind this CTA's epilogue (matters in the last wave)
  } else {
    // ------------------------------------------------------------ softmax: warp = TMEM lane quadrant, thread = query row
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const uint32_t to = tmem_o + lane_addr;
    float m_used = -INFINITY, l = 0.f;
    const float c = p.scale_log2;
    uint32_t v[2][32];
    // S(0) into registers; inside the loop the load of S(j+1) is issued right behind the store of P(j), so that its latency
    // (and the barrier round trip in front of it) overlaps the wait for that store instead of opening the next iteration
    mbar_wait(&s_full[0], 0);
    tc_fence_after();
    tmem_ld32(tmem_base + lane_addr, v[0]);
    tmem_ld32(tmem_base + lane_addr + 32, v[1]);
    for (int j = 0; j < T; ++j) {
      const uint32_t ts = tmem_base + lane_addr + (uint32_t)((j & 1) * 64);
      if ((p.dbg & 1) && j > 0) {   // no prefetch: S(j) is loaded here
        mbar_wait(&s_full[j & 1], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        tmem_ld32(ts, v[0]);
        tmem_ld32(ts + 32, v[1]);
      }
      tmem_ld_wait();
      if ((p.dbg & 4) && j > 0) {
        mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
        tc_fence_after();
      }
      // row max of the 64 scores: 3-input max (sm_100), four independent chains
#define VX_SV(k) __uint_as_float(v[(k) >> 5][(k) & 31])
      float mxs[4] = {VX_SV(0), VX_SV(1), VX_SV(2), VX_SV(3)};
#pragma unroll
      for (int k = 4; k < 60; k += 8)
#pragma unroll
        for (int u = 0; u < 4; ++u) mxs[u] = fmax3(mxs[u], VX_SV(k + u), VX_SV(k + 4 + u));
#pragma unroll
      for (int u = 0; u < 4; ++u) mxs[u] = fmaxf(mxs[u], VX_SV(60 + u));
#undef VX_SV
      const float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
      const float m_new = fmaxf(m_used, mx);
      const bool need = (m_new - m_used) * c > 8.0f;   // lazy rescale: only when the running max grew by > 2^8
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = (m_used == -INFINITY) ? 0.f : ex2_approx((m_used - m_new) * c);
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          // P.V(j-1) has landed in O.  The parity is unambiguous here: S(j), which this thread has already seen complete,
          // was issued after P.V(j-2) had completed, so pv_done is in phase j - 1 or j.
          mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
          tc_fence_after();
          for (int cb = 0; cb < p.hdp; cb += 16) {
            uint32_t o[16];
            tmem_ld16(to + cb, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(to + cb, o);
          }
          tmem_st_wait();
        }
      }
      const float mc = m_used * c;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
      // (a software-pipelined order pinned with volatile asm -- scale FMA of element k + 4 between the MUFU ops of elements
      // k and k + 1 -- measured 1.51 vs 1.49 ms: the compiler's own schedule is kept)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xx = fmaf(__uint_as_float(v[g][h * 8 + i]), c, -mc);
            e[i] = ((i & PMASK) == PMASK) ? ex2_poly(xx) : ex2_approx(xx);
            if (!ONES) ls[i & 3] += e[i];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) pk[g * 16 + h * 4 + i] = pack_bf16(e[2 * i], e[2 * i + 1]);
        }
      }
      if (!ONES) l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      tmem_st32(ts, pk);        // P(j): 64 keys = 32 packed columns over the S columns this thread has consumed
      if (p.dbg & 8) {
        tmem_st_wait();
        tc_fence_before();
      }
      if (j + 1 < T && !(p.dbg & 1)) {          // S(j+1) was issued an iteration ago: normally complete, the wait is a formality
        mbar_wait(&s_full[(j + 1) & 1], (uint32_t)(((j + 1) >> 1) & 1));
        tc_fence_after();
        const uint32_t tn = tmem_base + lane_addr + (uint32_t)(((j + 1) & 1) * 64);
        tmem_ld32(tn, v[0]);
        tmem_ld32(tn + 32, v[1]);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[j & 1]);
    }
    // The last P.V has its own single-phase barrier.  (Waiting for phase T - 1 of pv_done by parity is only meaningful when
    // that barrier is at most ONE phase behind; a thread that has stored P(T-1) only knows that P.V(T-3) has completed, so
    // with a late MMA warp the parity test passed while P.V(T-2) was still pending and O was read half-accumulated: the
    // run-to-run differences at hd >= 80, where the K/V loads keep the MMA warp waiting -- profiles/r02_flash_notes.md 5.)
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int qrow = q_tile * 128 + row;
    __nv_bfloat16* op = p.out + ((long long)bq * p.Nq + qrow) * p.ldo + head * p.hd;
    if (ONES) {
      uint32_t o[16];
      tmem_ld16(to + (p.hd / 16) * 16, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i == (p.hd & 15)) l = __uint_as_float(o[i]);
    }
    const float inv = 1.f / l;
    for (int cb = 0; cb < p.hdp; cb += 16) {
      uint32_t o[16];
      tmem_ld16(to + cb, o);
      tmem_ld_wait();
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        if (cb + h8 * 8 < p.hd) {
          const int b8 = h8 * 8;
          *reinterpret_cast<uint4*>(op + cb + b8) = make_uint4(
              pack_bf16(__uint_as_float(o[b8]) * inv, __uint_as_float(o[b8 + 1]) * inv),
              pack_bf16(__uint_as_float(o[b8 + 2]) * inv, __uint_as_float(o[b8 + 3]) * inv),
              pack_bf16(__uint_as_float(o[b8 + 4]) * inv, __uint_as_float(o[b8 + 5]) * inv),
              pack_bf16(__uint_as_float(o[b8 + 6]) * inv, __uint_as_float(o[b8 + 7]) * inv));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// (A v4 with EIGHT softmax warps per CTA -- the 64 keys of a step split into two independent 32-key online-softmax
// streams with their own O accumulators, merged in the epilogue -- was built and measured this round: 1.56 ms vs 1.49 ms at
// level 0, and its 2-rank run was not bit-reproducible; it was removed again.  profiles/r02_flash_notes.md has the numbers.)

// Fallback for key counts the tensor-core tiling cannot express (Nk < 16 or Nk % 16 != 0, e.g. the 2x2 / 12x12
// maps of reduced test resolutions): one warp per (batch, head, query), exact softmax, CUDA cores.
struct GaArgs {
  const __nv_bfloat16* q; long long ldq;
  const __nv_bfloat16* k; long long ldk;
  const __nv_bfloat16* v; long long ldv;
  __nv_bfloat16* out; long long ldo;
  int Bq, Nq, Nk, heads, hd, kv_div;
  float scale;
};

__global__ void __launch_bounds__(128) generic_attn_kernel(const GaArgs p) {
  pdl_enter();
  extern __shared__ float sc_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = sc_all + warp * p.Nk;
  const long long item = (long long)blockIdx.x * 4 + warp;
  const long long total = (long long)p.Bq * p.heads * p.Nq;
  if (item >= total) return;
  const int qi = (int)(item % p.Nq);
  const int head = (int)((item / p.Nq) % p.heads);
  const int b = (int)(item / ((long long)p.Nq * p.heads));
  const __nv_bfloat16* qp = p.q + ((long long)b * p.Nq + qi) * p.ldq + head * p.hd;
  const long long kv0 = (long long)(b / p.kv_div) * p.Nk;
  float mx = -INFINITY;
  for (int j = lane; j < p.Nk; j += 32) {
    const __nv_bfloat16* kp = p.k + (kv0 + j) * p.ldk + head * p.hd;
    float acc = 0.f;
    for (int c = 0; c < p.hd; c += 2) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qp + c));
      const float2 bb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(kp + c));
      acc += a.x * bb.x + a.y * bb.y;
    }
    acc *= p.scale;
    sc[j] = acc;
    mx = fmaxf(mx, acc);
  }
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < p.Nk; j += 32) {
    const float e = __expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int c = lane * 2; c < p.hd; c += 64) {
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < p.Nk; ++j) {
      const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.v + (kv0 + j) * p.ldv + head * p.hd + c));
      o0 += sc[j] * vv.x;
      o1 += sc[j] * vv.y;
    }
    *reinterpret_cast<__nv_bfloat162*>(p.out + ((long long)b * p.Nq + qi) * p.ldo + head * p.hd + c) =
        __floats2bfloat162_rn(o0 * inv, o1 * inv);
  }
}

}  // namespace vx

using namespace vx;

// A/B switches (bring-up only), read ONCE per process: the launch path itself never touches the environment.
namespace {
struct FaEnv {
  int v1, v2, v3_noload, v3_dbg, psmem, noones, stages, stagger, pairsync, latewait, baton, poly, v3_stages;
  long long* trace;
  static int geti(const char* n, int d) {
    const char* e = getenv(n);
    return e ? atoi(e) : d;
  }
  FaEnv() {
    v1 = geti("VX_FA_V1", 0);
    v2 = geti("VX_FA_V2", 0);
    psmem = geti("VX_FA_PSMEM", 0);
    noones = geti("VX_FA_NOONES", 0);
    stages = geti("VX_FA_STAGES", 0);
    stagger = geti("VX_FA_STAGGER", 0);
    pairsync = geti("VX_FA_PAIRSYNC", 1);
    latewait = geti("VX_FA_LATEWAIT", 0);
    baton = geti("VX_FA_BATON", 0);
    poly = geti("VX_FA_POLY", 8);
    v3_stages = geti("VX_FA3_STAGES", 0);
    v3_noload = geti("VX_FA3_NOLOAD", 0);
    v3_dbg = geti("VX_FA3_DBG", 0);
    const char* t = getenv("VX_FA_TRACE");
    trace = t ? (long long*)strtoull(t, nullptr, 10) : nullptr;
  }
};
const FaEnv& fa_env() {
  static const FaEnv e;
  return e;
}
}  // namespace

extern "C" void vx_flash_reload_env() {   // test / sweep hook: re-read the switches (not used by the product)
  const_cast<FaEnv&>(fa_env()) = FaEnv();
}

// q: [Bq*Nq, ldq], k/v: [Bkv*Nk, ldk] / [.., ldv] (bf16, heads*hd columns used), out: [Bq*Nq, ldo].
// kv batch of query batch b is b / kv_div.
extern "C" int vx_flash_attention(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                  long long ldv, void* out, long long ldo, int Bq, int Nq, int Bkv, int Nk, int heads,
                                  int hd, int kv_div, void* stream) {
  VX_REQUIRE(hd % 8 == 0 && hd <= 256, "vx_flash_attention: hd=%d unsupported", hd);
  VX_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "vx_flash_attention: ld must be %%8");
  VX_REQUIRE(kv_div >= 1 && (Bq + kv_div - 1) / kv_div <= Bkv, "vx_flash_attention: kv_div=%d Bq=%d Bkv=%d", kv_div, Bq, Bkv);
  if (Nk < 16 || Nk % 16 != 0) {
    VX_REQUIRE(Nk * 4 * 4 <= 48 * 1024, "vx_flash_attention: Nk=%d not a multiple of 16 and too long for the fallback", Nk);
    GaArgs g{(const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv,
             (__nv_bfloat16*)out, ldo, Bq, Nq, Nk, heads, hd, kv_div, 1.0f / sqrtf((float)hd)};
    const long long items = (long long)Bq * heads * Nq;
    launch_k(generic_attn_kernel, dim3((unsigned)((items + 3) / 4)), dim3(128), (size_t)Nk * 4 * 4, (cudaStream_t)stream, g);
    VX_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int hdp = (hd + 15) / 16 * 16;
  const FaEnv& env = fa_env();
  if (hdp <= 128 && Nq % 128 == 0 && Nk % 64 == 0 && !env.v1 && !env.v2) {
    // ---- v3: one query tile per CTA, 64 keys per step, 2-3 CTAs per SM
    Fa3Args a{};
    a.noload = env.v3_noload;
    a.dbg = env.v3_dbg;
    a.Nq = Nq; a.Nk = Nk; a.hd = hd; a.hdp = hdp; a.kv_div = kv_div;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
    a.out = (__nv_bfloat16*)out; a.ldo = ldo;
    a.q_bytes = 128 * hdp * 2;
    a.kv_bytes = 64 * hdp * 2;
    a.tmem_cols = 256u;                       // 2 x 64 (S double buffer) + hdp (<= 128) columns
    const int ctas = 2;
    const size_t budget = (size_t)(227 * 1024) / ctas - 1024;          // 1 KB per CTA is reserved by the hardware
    auto need3 = [&](int st) { return (size_t)a.q_bytes + (size_t)2 * st * a.kv_bytes + 256 + 128; };
    a.stages = env.v3_stages > 0 ? env.v3_stages : 6;
    while (a.stages > 2 && need3(a.stages) > budget) --a.stages;
    if (a.stages > 8) a.stages = 8;
    const size_t smem3 = need3(a.stages);
    VX_REQUIRE(smem3 <= 227 * 1024, "vx_flash_attention: smem %zu too large (hd=%d)", smem3, hd);
    CUtensorMap mQ, mK, mV;
    const void* ptrs[3] = {q, k, v};
    const long long lds[3] = {ldq, ldk, ldv};
    const long long rows[3] = {(long long)Bq * Nq, (long long)Bkv * Nk, (long long)Bkv * Nk};
    CUtensorMap* maps[3] = {&mQ, &mK, &mV};
    for (int i = 0; i < 3; ++i) {
      uint64_t dims[3] = {8, (uint64_t)rows[i], (uint64_t)lds[i] / 8};
      uint64_t str[2] = {(uint64_t)lds[i] * 2, 16};
      uint32_t box[3] = {8, i == 0 ? 128u : 64u, (uint32_t)hd / 8};
      if (make_tmap_bf16(maps[i], ptrs[i], 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    }
    static bool cfg3 = false;
    if (!cfg3) {
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn3_kernel<true, 7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn3_kernel<true, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn3_kernel<true, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn3_kernel<false, 7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
      cfg3 = true;
    }
    dim3 grid3(Nq / 128, heads, Bq);
    const bool ones = hdp > hd && !env.noones;
    auto st3 = (cudaStream_t)stream;
    if (ones && env.poly == 4) launch_k((flash_attn3_kernel<true, 3, 2>), dim3(grid3), dim3(kFa3Threads), smem3, st3, mQ, mK, mV, a);
    else if (ones && env.poly == 2) launch_k((flash_attn3_kernel<true, 1, 2>), dim3(grid3), dim3(kFa3Threads), smem3, st3, mQ, mK, mV, a);
    else if (ones) launch_k((flash_attn3_kernel<true, 7, 2>), dim3(grid3), dim3(kFa3Threads), smem3, st3, mQ, mK, mV, a);
    else launch_k((flash_attn3_kernel<false, 7, 2>), dim3(grid3), dim3(kFa3Threads), smem3, st3, mQ, mK, mV, a);
    VX_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (hdp <= 128 && Nq % 256 == 0 && Nk % 128 == 0 && !env.v1) {
    Fa2Args a{};
    a.Nq = Nq; a.Nk = Nk; a.hd = hd; a.hdp = hdp; a.kv_div = kv_div;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
    a.out = (__nv_bfloat16*)out; a.ldo = ldo;
    a.q_bytes = 128 * hdp * 2;
    a.kv_bytes = 128 * hdp * 2;
    a.stages = 3;
    a.stagger = env.stagger;
    a.pair_sync = env.pairsync;
    a.late_wait = env.latewait;
    a.trace = env.trace;
    auto need = [&](int st, int pb) { return (size_t)2 * a.q_bytes + (size_t)2 * st * a.kv_bytes + (size_t)2 * pb * 32768 + 6144 + 512 + 128; };
    // P in tensor memory whenever O (2 x hdp) + S (2 x 128) + P (2 x 64) columns fit the 512-column TMEM
    a.p_tmem = (hdp <= 64 && !env.psmem) ? 1 : 0;
    a.pbufs = 1;
    a.stages = 6;
    const int psm = a.p_tmem ? 0 : 1;
    while (a.stages > 2 && need(a.stages, psm) > 227 * 1024) --a.stages;
    if (env.stages) a.stages = env.stages;
    const size_t smem2 = need(a.stages, psm);
    VX_REQUIRE(smem2 <= 227 * 1024, "vx_flash_attention: smem %zu too large (hd=%d)", smem2, hd);
    CUtensorMap mQ, mK, mV;
    const void* ptrs[3] = {q, k, v};
    const long long lds[3] = {ldq, ldk, ldv};
    const long long rows[3] = {(long long)Bq * Nq, (long long)Bkv * Nk, (long long)Bkv * Nk};
    CUtensorMap* maps[3] = {&mQ, &mK, &mV};
    for (int i = 0; i < 3; ++i) {
      uint64_t dims[3] = {8, (uint64_t)rows[i], (uint64_t)lds[i] / 8};
      uint64_t str[2] = {(uint64_t)lds[i] * 2, 16};
      uint32_t box[3] = {8, 128, (uint32_t)hd / 8};
      if (make_tmap_bf16(maps[i], ptrs[i], 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    }
    static bool cfg2 = false;
    if (!cfg2) {
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn2_kernel<true, false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn2_kernel<false, false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn2_kernel<true, true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn2_kernel<true, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn2_kernel<true, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn2_kernel<false, true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      cfg2 = true;
    }
    dim3 grid2(Nq / 256, heads, Bq);
    const bool ones = hdp > hd && !env.noones;
    const bool baton = env.baton != 0;
    const int poly = env.poly;      // 1/poly of the exponentials on the FMA pipe
    auto st2 = (cudaStream_t)stream;
    if (ones && baton && poly == 4) launch_k((flash_attn2_kernel<true, true, 3>), dim3(grid2), dim3(kFa2Threads), smem2, st2, mQ, mK, mV, a);
    else if (ones && baton && poly == 2) launch_k((flash_attn2_kernel<true, true, 1>), dim3(grid2), dim3(kFa2Threads), smem2, st2, mQ, mK, mV, a);
    else if (ones && baton) launch_k((flash_attn2_kernel<true, true, 7>), dim3(grid2), dim3(kFa2Threads), smem2, st2, mQ, mK, mV, a);
    else if (ones) launch_k((flash_attn2_kernel<true, false, 7>), dim3(grid2), dim3(kFa2Threads), smem2, st2, mQ, mK, mV, a);
    else if (baton) launch_k((flash_attn2_kernel<false, true, 7>), dim3(grid2), dim3(kFa2Threads), smem2, st2, mQ, mK, mV, a);
    else launch_k((flash_attn2_kernel<false, false, 7>), dim3(grid2), dim3(kFa2Threads), smem2, st2, mQ, mK, mV, a);
    VX_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  int kv_tile = 0;
  const int kv_max = hdp > 96 ? 64 : 128;
  for (int t = kv_max; t >= 16; t -= 16)
    if (Nk % t == 0) {
      kv_tile = t;
      break;
    }
  VX_REQUIRE(kv_tile > 0, "vx_flash_attention: Nk=%d must be a multiple of 16", Nk);
  FaArgs a{};
  a.Nq = Nq; a.Nk = Nk; a.hd = hd; a.hdp = hdp; a.kv_tile = kv_tile; a.kv_div = kv_div; a.heads = heads;
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  a.out = (__nv_bfloat16*)out; a.ldo = ldo;
  a.q_bytes = 128 * hdp * 2;
  a.kv_bytes = kv_tile * hdp * 2;
  a.p_bytes = 128 * kv_tile * 2;
  const size_t smem = (size_t)a.q_bytes + 2 * kFaStages * a.kv_bytes + 2 * a.p_bytes + 256 + 128;
  VX_REQUIRE(smem <= 227 * 1024, "vx_flash_attention: smem %zu too large (hd=%d)", smem, hd);
  CUtensorMap mQ, mK, mV;
  {
    uint64_t dims[3] = {8, (uint64_t)Bq * Nq, (uint64_t)ldq / 8};
    uint64_t str[2] = {(uint64_t)ldq * 2, 16};
    uint32_t box[3] = {8, 128, (uint32_t)hd / 8};
    if (make_tmap_bf16(&mQ, q, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  }
  {
    uint64_t dims[3] = {8, (uint64_t)Bkv * Nk, (uint64_t)ldk / 8};
    uint64_t str[2] = {(uint64_t)ldk * 2, 16};
    uint32_t box[3] = {8, (uint32_t)kv_tile, (uint32_t)hd / 8};
    if (make_tmap_bf16(&mK, k, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  }
  {
    uint64_t dims[3] = {8, (uint64_t)Bkv * Nk, (uint64_t)ldv / 8};
    uint64_t str[2] = {(uint64_t)ldv * 2, 16};
    uint32_t box[3] = {8, (uint32_t)kv_tile, (uint32_t)hd / 8};
    if (make_tmap_bf16(&mV, v, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  }
  static bool cfg = false;
  if (!cfg) {
    VX_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfg = true;
  }
  dim3 grid((Nq + 127) / 128, heads, Bq);
  launch_k(flash_attn_kernel, dim3(grid), dim3(kFaThreads), smem, (cudaStream_t)stream, mQ, mK, mV, a);
  VX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
