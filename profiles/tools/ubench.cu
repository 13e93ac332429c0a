// Micro-benchmarks of the SM resources the flash-attention softmax leans on (B200, sm_100a):
// tcgen05.ld / tcgen05.st throughput per SM vs number of warps, MUFU ex2 throughput, FFMA / polynomial-exp throughput and
// MUFU + FMA co-issue.  One CTA per SM (148), clock64 around the measured loop, result = units per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I v-express_b200/csrc -o profiles/tools/build/ubench profiles/tools/ubench.cu
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include "vx_ptx.cuh"

using namespace vx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// mode 0: LDTM x32 (4 KB per warp instruction); 1: STTM x32; 2: LDTM x16
__global__ void __launch_bounds__(512, 1) tmem_kernel(int mode, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)(((it + warp) * 32) & 511);
    if (mode == 0) {
      tmem_ld32(base + (col & 480), v);
      tmem_ld_wait();
      acc ^= v[0] ^ v[17] ^ v[31];
    } else if (mode == 1) {
      v[5] += (uint32_t)it;
      tmem_st32(base + (col & 480), v);
      tmem_st_wait();
    } else {
      uint32_t w[16];
      tmem_ld16(base + (col & 496), w);
      tmem_ld_wait();
      acc ^= w[0] ^ w[9] ^ w[15];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc + v[3];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

// LDTM with 2 loads in flight before the wait (x32 each) -- what a softmax warp does for its 64 columns
__global__ void __launch_bounds__(512, 1) tmem_ld2_kernel(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[2][32];
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)(((it + warp) * 64) & 448);
    tmem_ld32(base + col, v[0]);
    tmem_ld32(base + col + 32, v[1]);
    tmem_ld_wait();
    acc ^= v[0][0] ^ v[0][31] ^ v[1][7] ^ v[1][31];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float j = t - 12582912.0f;
  const float r = x - j;
  float pl = fmaf(0.05519810691475868f, r, 0.24267712235450745f);
  pl = fmaf(pl, r, 0.6932618021965027f);
  pl = fmaf(pl, r, 0.9999227523803711f);
  return __uint_as_float(__float_as_uint(pl) + (__float_as_uint(t) << 23));
}

// mode 0: 16 independent ex2.approx per iteration; 1: 16 FFMA (3-reg); 2: 16 poly exps; 3: 12 MUFU + 4 poly;
// 4: 8 MUFU + 8 poly; 5: softmax-like: fma + ex2 + pack per element (16 elements); 6: same with 1/4 poly
__global__ void __launch_bounds__(1024, 1) alu_kernel(int mode, int iters, long long* out, float* sink) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i + 1);
  const float c = 0.999f, mc = 0.0001f;
  uint32_t pk = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex2a(x[i]) - 1.0001f;
    } else if (mode == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], c, x[(i + 1) & 15]);
    } else if (mode == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex2_poly(x[i]) - 1.0001f;
    } else if (mode == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ((i & 3) == 3 ? ex2_poly(x[i]) : ex2a(x[i])) - 1.0001f;
    } else if (mode == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ((i & 1) ? ex2_poly(x[i]) : ex2a(x[i])) - 1.0001f;
    } else if (mode == 5) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float a = ex2a(fmaf(x[i], c, -mc)), b = ex2a(fmaf(x[i + 1], c, -mc));
        pk ^= pack_bf16(a, b);
        x[i] = a - 1.0001f; x[i + 1] = b - 1.0001f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float a = ex2a(fmaf(x[i], c, -mc));
        const float b = (i & 2) ? ex2_poly(fmaf(x[i + 1], c, -mc)) : ex2a(fmaf(x[i + 1], c, -mc));
        pk ^= pack_bf16(a, b);
        x[i] = a - 1.0001f; x[i + 1] = b - 1.0001f;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 12345.f) sink[0] = s + __uint_as_float(pk);
}


// One softmax warp of the flash kernel in isolation, nw independent warps per CTA, no barriers between warps:
//   LDTM 64 S columns -> [row max] -> 64 x (fma, ex2, pack) -> STTM 32 packed P columns.
// VARIANT 0: max first (the exponentials depend on it).  1: stale max (exponentials use the previous tile's max; this
// tile's max is computed alongside).  2: no max at all.  3: like 0, all exponentials on MUFU (no polynomial share).
// 4: like 1 with the LDTM of the second 32 columns overlapped with the first 32 exponentials.
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) softmax_warp_kernel(int iters, float c, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64 & 255);
  const uint32_t pbase = base + 256;
  {  // fill the S region with something finite
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-0.01f * (float)((threadIdx.x * 7 + i * 13) & 255));
    tmem_st32(base, z);
    tmem_st32(base + 32, z);
    tmem_st_wait();
  }
  float m_used = 0.f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[2][32];
    tmem_ld32(base, v[0]);
    if (VARIANT != 4) tmem_ld32(base + 32, v[1]);
    tmem_ld_wait();
    if (VARIANT == 4) tmem_ld32(base + 32, v[1]);
    float mc;
    if (VARIANT == 0 || VARIANT == 3) {
      float mxs[4] = {__uint_as_float(v[0][0]), __uint_as_float(v[0][1]), __uint_as_float(v[0][2]), __uint_as_float(v[0][3])};
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int i = 0; i < 32; ++i) mxs[i & 3] = fmaxf(mxs[i & 3], __uint_as_float(v[g][i]));
      m_used = fmaxf(m_used, fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])));
      mc = m_used * c;
    } else {
      mc = m_used * c;
    }
    uint32_t pk[32];
    float mxs[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (VARIANT == 4 && g == 1) tmem_ld_wait();
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float sv = __uint_as_float(v[g][h * 8 + i]);
          if (VARIANT == 1 || VARIANT == 4) mxs[i & 3] = fmaxf(mxs[i & 3], sv);
          const float xx = fmaf(sv, c, -mc);
          e[i] = (VARIANT != 3 && i == 7) ? ex2_poly(xx) : ex2a(xx);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) pk[g * 16 + h * 4 + i] = pack_bf16(e[2 * i], e[2 * i + 1]);
      }
    }
    if (VARIANT == 1 || VARIANT == 4) m_used = fmaxf(m_used, fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])));
    tmem_st32(pbase, pk);
    tmem_st_wait();
    acc += pk[0];
  }
  const long long t1 = clock64();
  __syncthreads();
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&out[blockIdx.x], (unsigned long long)(t1 - t0));
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d_out;
  uint32_t* d_sink;
  CK(cudaMalloc(&d_out, sms * sizeof(long long)));
  CK(cudaMalloc(&d_sink, 64));
  std::vector<long long> h(sms);
  auto med = [&]() {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<long long> s(h);
    std::sort(s.begin(), s.end());
    return (double)s[s.size() / 2];
  };
  const int iters = 4096;
  printf("SMs: %d\n", sms);
  const char* tn[3] = {"LDTM 32x32b.x32 (wait each)", "STTM 32x32b.x32 (wait each)", "LDTM 32x32b.x16 (wait each)"};
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 4, 8, 16}) {
      tmem_kernel<<<sms, nw * 32, 0>>>(mode, 64, d_out, d_sink);
      tmem_kernel<<<sms, nw * 32, 0>>>(mode, iters, d_out, d_sink);
      const double clk = med();
      const double bytes = (double)nw * iters * (mode == 2 ? 2048.0 : 4096.0);
      printf("%-30s warps=%2d: %8.1f clk per instr per warp, %7.1f B/clk/SM\n", tn[mode], nw, clk / iters, bytes / clk);
    }
  for (int nw : {1, 4, 8, 16}) {
    tmem_ld2_kernel<<<sms, nw * 32, 0>>>(iters, d_out, d_sink);
    const double clk = med();
    printf("%-30s warps=%2d: %8.1f clk per 64-column row block, %7.1f B/clk/SM\n", "LDTM 2 x x32 then wait", nw, clk / iters,
           (double)nw * iters * 8192.0 / clk);
  }
  const char* an[7] = {"MUFU ex2.approx", "FFMA 3-reg", "poly exp2 (FMA/ALU pipes)", "3/4 MUFU + 1/4 poly", "1/2 MUFU + 1/2 poly",
                       "softmax elem: fma+ex2+pack", "softmax elem, 1/4 poly"};
  for (int mode = 0; mode < 7; ++mode)
    for (int nw : {4, 8, 16, 32}) {
      alu_kernel<<<sms, nw * 32, 0>>>(mode, 64, d_out, (float*)d_sink);
      alu_kernel<<<sms, nw * 32, 0>>>(mode, iters, d_out, (float*)d_sink);
      const double clk = med();
      printf("%-30s warps=%2d: %7.2f elements/clk/SM\n", an[mode], nw, (double)nw * 32 * iters * 16 / clk);
    }

  {
    const char* vn[5] = {"softmax warp: max first, 1/8 poly", "softmax warp: stale max, 1/8 poly", "softmax warp: no max, 1/8 poly",
                         "softmax warp: max first, all MUFU", "softmax warp: stale max, split LDTM"};
    for (int var = 0; var < 5; ++var)
      for (int nw : {4, 8, 12, 16}) {
        CK(cudaMemset(d_out, 0, sms * sizeof(long long)));
        auto run = [&](int n) {
          switch (var) {
            case 0: softmax_warp_kernel<0><<<sms, nw * 32>>>(n, 0.2f, d_out, d_sink); break;
            case 1: softmax_warp_kernel<1><<<sms, nw * 32>>>(n, 0.2f, d_out, d_sink); break;
            case 2: softmax_warp_kernel<2><<<sms, nw * 32>>>(n, 0.2f, d_out, d_sink); break;
            case 3: softmax_warp_kernel<3><<<sms, nw * 32>>>(n, 0.2f, d_out, d_sink); break;
            default: softmax_warp_kernel<4><<<sms, nw * 32>>>(n, 0.2f, d_out, d_sink); break;
          }
        };
        run(16);
        CK(cudaMemset(d_out, 0, sms * sizeof(long long)));
        run(2048);
        const double clk = med();
        printf("%-36s warps/SMSP=%d: %7.1f clk per (32 rows x 64 cols) per warp, %6.2f elements/clk/SM, 2 q-tiles x 128 keys = %6.0f clk\n",
               vn[var], nw / 4, clk / 2048, (double)nw * 2048 * 2048 / clk, clk / 2048 * 16.0 / nw);
      }
  }
  return 0;
}
