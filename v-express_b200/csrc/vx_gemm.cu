// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   out[m, n] = (sum_k A[m, k] * W[n, k] + bias[n] + bias2[m / bias2_div, n]) * scale + residual[m, n]
//
// A is bf16 row-major (K contiguous), W is bf16 [N, K] (K contiguous): both operands are K-major, so every
// 64-wide K block of a 128-row tile is one TMA box that lands in shared memory in the canonical 128B-swizzled
// K-major layout tcgen05.mma consumes.  Accumulation is fp32 in TMEM; the epilogue reads it back with
// tcgen05.ld, applies bias / per-sample bias (time embedding) / scale / residual and stores bf16.
//
// Three producers share the same MMA + epilogue:
//   * plain GEMM, optionally split-K over two sources (A | A2) -- the `torch.cat([h, skip])` of the up blocks
//     (reference modules/unet_3d_blocks.py:694,831) folded into the K loop;
//   * 3x3 convolution, stride 1, pad 1, NHWC: K = 9 taps x Cin; for tap (dy,dx) the A box is the same pixel
//     rectangle shifted by (dy-1, dx-1) through a 4-D tensor map (C, W, H, N) whose out-of-bounds reads are
//     zero-filled by the TMA unit = the zero padding of nn.Conv2d (reference modules/resnet.py:9-17).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (one TMEM lane quadrant each).
#include "vx_host.h"
#include "vx_ptx.cuh"

namespace vx {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kThreads = 192;

struct GemmArgs {
  int M, N;
  int kblocks1;   // 64-wide K blocks taken from A (per tap for conv)
  int kblocks2;   // ... then from A2 (plain mode only)
  int taps;       // 1 = plain GEMM, 9 = 3x3 conv
  int block_n;    // UMMA N (multiple of 16, <= 256)
  int stages;
  int rows_valid;  // output rows covered by one tile (128 for plain; wbox*hbox*nbox for conv)
  int W, H;        // conv image size
  int tmem_cols;
  const float* bias;
  const float* bias2;
  int bias2_div;
  float scale;
  const __nv_bfloat16* residual;
  long long ldr;
  __nv_bfloat16* out;
  long long ldc;
  int out_f32;  // 1: `out` is float* (used for attention scores that feed an fp32 softmax)
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
                    const __grid_constant__ CUtensorMap mapB, const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atom
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = kBlockM * kBlockK * 2;
  const int b_bytes = p.block_n * kBlockK * 2;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x;
  const int tile_m = blockIdx.y;
  const int total_kb = p.taps * p.kblocks1 + p.kblocks2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    if (p.kblocks2) tma_prefetch_desc(&mapA2);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int n0 = 0, y0 = 0, x0 = 0;
      const long long m0 = (long long)tile_m * p.rows_valid;
      if (p.taps == 9) {
        const long long hw = (long long)p.H * p.W;
        n0 = (int)(m0 / hw);
        const int rem = (int)(m0 % hw);
        y0 = rem / p.W;
        x0 = rem % p.W;
      }
      const uint32_t tx_bytes = (uint32_t)(p.rows_valid * kBlockK * 2 + b_bytes);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_expect_tx(&full_bar[stage], tx_bytes);
        if (p.taps == 9) {
          const int tap = kb / p.kblocks1;
          const int cb = kb - tap * p.kblocks1;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma_load_4d(sa, &mapA, &full_bar[stage], cb * kBlockK, x0 + dx, y0 + dy, n0);
        } else if (kb < p.kblocks1) {
          tma_load_2d(sa, &mapA, &full_bar[stage], kb * kBlockK, (int)m0);
        } else {
          tma_load_2d(sa, &mapA2, &full_bar[stage], (kb - p.kblocks1) * kBlockK, (int)m0);
        }
        tma_load_2d(sb, &mapB, &full_bar[stage], kb * kBlockK, tile_n * p.block_n);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kBlockM, (uint32_t)p.block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sb = sa + a_bytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024, SWZ_128B);
          const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024, SWZ_128B);
          umma_ss(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const long long m = (long long)tile_m * p.rows_valid + row;
    const bool row_ok = row < p.rows_valid && m < p.M;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int nbase = tile_n * p.block_n;
    const float* b2 = p.bias2 ? p.bias2 + (row_ok ? (m / p.bias2_div) : 0) * (long long)p.N : nullptr;
    for (int c = 0; c < p.block_n; c += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_ld_wait();
      const int n = nbase + c;
      if (row_ok && n < p.N) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
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
        if (p.scale != 1.0f) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] *= p.scale;
        }
        if (p.residual) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m * p.ldr + n);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 r = rp[h];
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 t = unpack_bf16(rr[i]);
              f[h * 8 + 2 * i] += t.x;
              f[h * 8 + 2 * i + 1] += t.y;
            }
          }
        }
        if (p.out_f32) {
          float4* fp = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + m * p.ldc + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) fp[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          continue;
        }
        uint4* op = reinterpret_cast<uint4*>(p.out + m * p.ldc + n);
        op[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        op[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]),
                           pack_bf16(f[14], f[15]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

static int pick_block_n(int M, int N, int rows_per_tile) {
  static const int cand[] = {256, 240, 224, 208, 192, 176, 160, 144, 128, 112, 96, 80, 64, 48, 32, 16};
  const long long tiles_m = (M + rows_per_tile - 1) / rows_per_tile;
  int best = 0;
  for (int bn : cand) {
    if (N % bn) continue;
    if (!best) best = bn;                      // largest divisor
    if (tiles_m * (N / bn) >= 148 && bn >= 128) return bn;  // largest divisor that still fills the chip
  }
  // small problem: prefer >=128-wide tiles when they exist, else the largest divisor
  for (int bn : cand)
    if (N % bn == 0 && bn <= 160 && bn >= 64) return bn;
  return best;
}

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

static int launch(const CUtensorMap& mA, const CUtensorMap& mA2, const CUtensorMap& mB, GemmArgs& a, cudaStream_t st) {
  const int stage_bytes = kBlockM * kBlockK * 2 + a.block_n * kBlockK * 2;
  int stages = env_int("VX_GEMM_STAGES", 0);
  if (stages <= 0) stages = (a.block_n > 128) ? 4 : 6;
  const int total_kb = a.taps * a.kblocks1 + a.kblocks2;
  if (stages > total_kb) stages = total_kb < 2 ? 2 : total_kb;
  while (stages * stage_bytes + 2048 > 227 * 1024) --stages;
  a.stages = stages;
  a.tmem_cols = pow2_cols(a.block_n);
  const size_t smem = (size_t)stages * stage_bytes + 2048;
  static size_t configured = 0;
  if (smem > configured) {
    VX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  dim3 grid((a.N + a.block_n - 1) / a.block_n, (a.M + a.rows_valid - 1) / a.rows_valid);
  gemm_tcgen05_kernel<<<grid, kThreads, smem, st>>>(mA, mA2, mB, a);
  VX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vx

using namespace vx;

extern "C" int vx_gemm_bf16(const void* A, long long lda, int K1, const void* A2, long long lda2, int K2,
                            const void* Wt, long long ldw, int M, int N, const float* bias, const float* bias2,
                            int bias2_div, float scale, const void* residual, long long ldr, void* out,
                            long long ldc, int out_f32, int block_n, void* stream) {
  VX_REQUIRE(M > 0 && N > 0 && K1 > 0, "vx_gemm_bf16: bad shape M=%d N=%d K1=%d", M, N, K1);
  VX_REQUIRE(N % 16 == 0, "vx_gemm_bf16: N=%d must be a multiple of 16", N);
  VX_REQUIRE(K1 % 8 == 0 && K2 % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0,
             "vx_gemm_bf16: K/ld must be multiples of 8 elements (16-byte TMA strides)");
  VX_REQUIRE(K2 == 0 || (K1 % kBlockK == 0 && lda2 % 8 == 0), "vx_gemm_bf16: split-K needs K1 %% 64 == 0");
  VX_REQUIRE(!residual || ldr % 8 == 0, "vx_gemm_bf16: ldr must be a multiple of 8");
  if (block_n <= 0) block_n = env_int("VX_GEMM_BN", 0);
  if (block_n <= 0) block_n = pick_block_n(M, N, kBlockM);
  VX_REQUIRE(block_n % 16 == 0 && block_n >= 16 && block_n <= 256, "vx_gemm_bf16: block_n=%d invalid", block_n);
  CUtensorMap mA, mA2, mB;
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
    uint32_t box[2] = {kBlockK, (uint32_t)block_n};
    if (make_tmap_bf16(&mB, Wt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  GemmArgs a{};
  a.M = M; a.N = N;
  a.kblocks1 = (K1 + kBlockK - 1) / kBlockK;
  a.kblocks2 = (K2 + kBlockK - 1) / kBlockK;
  a.taps = 1;
  a.block_n = block_n;
  a.rows_valid = kBlockM;
  a.W = a.H = 1;
  a.bias = bias; a.bias2 = bias2; a.bias2_div = bias2_div > 0 ? bias2_div : 1; a.scale = scale;
  a.residual = (const __nv_bfloat16*)residual; a.ldr = ldr;
  a.out = (__nv_bfloat16*)out; a.ldc = ldc; a.out_f32 = out_f32;
  return launch(mA, mA2, mB, a, (cudaStream_t)stream);
}

// X: NHWC bf16 [NB, H, W, C];  Wt: [Cout, 9*C] with K index = (ky*3+kx)*C + c;  out: [NB*H*W, ldc]
extern "C" int vx_conv3x3_bf16(const void* X, int NB, int H, int W, int C, const void* Wt, int Cout,
                               const float* bias, const float* bias2, int bias2_div, float scale,
                               const void* residual, long long ldr, void* out, long long ldc, int block_n,
                               void* stream) {
  VX_REQUIRE(C % 8 == 0 && Cout % 16 == 0, "vx_conv3x3_bf16: C=%d must be %%8, Cout=%d %%16", C, Cout);
  // pixel rectangle of <= 128 output rows that is contiguous in NHWC row order
  int wbox, hbox = 1, nbox = 1;
  if (W >= kBlockM) {
    VX_REQUIRE(W % kBlockM == 0, "vx_conv3x3_bf16: W=%d must be a multiple of 128 when >= 128", W);
    wbox = kBlockM;
  } else {
    wbox = W;
    hbox = kBlockM / W;
    if (hbox > H) {
      hbox = H;
      nbox = kBlockM / (W * H);
      if (nbox > NB) nbox = NB;
      if (nbox < 1) nbox = 1;
    } else {
      while (H % hbox) --hbox;
    }
  }
  const int rows_valid = wbox * hbox * nbox;
  const long long M = (long long)NB * H * W;
  VX_REQUIRE(M % rows_valid == 0, "vx_conv3x3_bf16: NB*H*W=%lld not tileable by %d", M, rows_valid);
  if (block_n <= 0) block_n = env_int("VX_GEMM_BN", 0);
  if (block_n <= 0) block_n = pick_block_n((int)M, Cout, rows_valid);
  VX_REQUIRE(block_n % 16 == 0 && block_n >= 16 && block_n <= 256, "vx_conv3x3_bf16: block_n=%d invalid", block_n);
  CUtensorMap mA, mB;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {kBlockK, (uint32_t)wbox, (uint32_t)hbox, (uint32_t)nbox};
    if (make_tmap_bf16(&mA, X, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * C, (uint64_t)Cout};
    uint64_t str[1] = {(uint64_t)9 * C * 2};
    uint32_t box[2] = {kBlockK, (uint32_t)block_n};
    if (make_tmap_bf16(&mB, Wt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  VX_REQUIRE(C % kBlockK == 0, "vx_conv3x3_bf16: C=%d must be a multiple of 64", C);
  GemmArgs a{};
  a.M = (int)M; a.N = Cout;
  a.kblocks1 = C / kBlockK;
  a.kblocks2 = 0;
  a.taps = 9;
  a.block_n = block_n;
  a.rows_valid = rows_valid;
  a.W = W; a.H = H;
  a.bias = bias; a.bias2 = bias2; a.bias2_div = bias2_div > 0 ? bias2_div : 1; a.scale = scale;
  a.residual = (const __nv_bfloat16*)residual; a.ldr = ldr;
  a.out = (__nv_bfloat16*)out; a.ldc = ldc;
  return launch(mA, mA, mB, a, (cudaStream_t)stream);
}
