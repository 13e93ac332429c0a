// Bring-up probes (used only by tests/test_probe_gpu.py): run one tcgen05.mma chain on caller-provided
// shared-memory images + descriptor fields, and dump what a TMA box load leaves in shared memory.
// They pin the descriptor / layout conventions the production kernels rely on.
#include "vx_host.h"
#include "vx_bringup.h"
#include "vx_ptx.cuh"

namespace vx {

struct ProbeArgs {
  const uint8_t* a_img; int a_bytes;
  const uint8_t* b_img; int b_bytes;
  uint32_t lboA, sboA, layA, lboB, sboB, layB;
  uint32_t idesc;
  int ksteps, a_step, b_step;
  int N;
  float* out;  // [128, N]
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const ProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_bytes + 1023) & ~1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < p.a_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sa)[i] = reinterpret_cast<const uint32_t*>(p.a_img)[i];
  for (int i = threadIdx.x; i < p.b_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sb)[i] = reinterpret_cast<const uint32_t*>(p.b_img)[i];
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (threadIdx.x == 0) {
    for (int k = 0; k < p.ksteps; ++k) {
      const uint64_t da = make_smem_desc(smem_u32(sa) + k * p.a_step, p.lboA, p.sboA, p.layA);
      const uint64_t db = make_smem_desc(smem_u32(sb) + k * p.b_step, p.lboB, p.sboB, p.layB);
      umma_ss(tb, da, db, p.idesc, k ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < p.N; c += 16) {
    uint32_t v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) p.out[row * p.N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

// TS-MMA probe: A (128 x K bf16) comes from TENSOR MEMORY (written with tcgen05.st, row r in lane r, two K-adjacent
// bf16 per 32-bit column), B from shared memory.  Pins the layout the flash kernel uses to keep P in TMEM.
struct ProbeTsArgs {
  const uint32_t* a_packed;  // [128][K/2] 32-bit words: (a[r][2c] | a[r][2c+1] << 16)
  const uint8_t* b_img; int b_bytes;
  uint32_t lboB, sboB, layB;
  uint32_t idesc;
  int ksteps, b_step, K, N;
  float* out;  // [128, N]
};

__global__ void __launch_bounds__(128, 1) umma_ts_probe_kernel(const ProbeTsArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sb = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < p.b_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sb)[i] = reinterpret_cast<const uint32_t*>(p.b_img)[i];
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t a_tmem = tb + 256;  // A at columns [256, 256 + K/2)
  const int row = warp * 32 + lane;
  for (int c = 0; c < p.K / 2; c += 16) {
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = p.a_packed[row * (p.K / 2) + c + i];
    tmem_st16(a_tmem + ((uint32_t)(warp * 32) << 16) + c, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    for (int k = 0; k < p.ksteps; ++k) {
      const uint64_t db = make_smem_desc(smem_u32(sb) + k * p.b_step, p.lboB, p.sboB, p.layB);
      umma_ts(tb, a_tmem + (uint32_t)(k * 8), db, p.idesc, k ? 1u : 0u);   // 16 K elements = 8 columns per step
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < p.N; c += 16) {
    uint32_t v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) p.out[row * p.N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap map, int rank, int c0, int c1, int c2, int c3,
                                 int nbytes, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < nbytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xDEADBEEFu;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, nbytes);
    if (rank == 2) tma_load_2d(smem, &map, &bar, c0, c1);
    else if (rank == 3) tma_load_3d(smem, &map, &bar, c0, c1, c2);
    else tma_load_4d(smem, &map, &bar, c0, c1, c2, c3);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < nbytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(out)[i] = reinterpret_cast<uint32_t*>(smem)[i];
}

}  // namespace vx

using namespace vx;

extern "C" int vx_probe_umma(const void* a_img, int a_bytes, const void* b_img, int b_bytes, unsigned lboA,
                             unsigned sboA, unsigned layA, unsigned lboB, unsigned sboB, unsigned layB, int a_mn,
                             int b_mn, int N, int ksteps, int a_step, int b_step, float* out, void* stream) {
  VX_REQUIRE(N % 16 == 0 && N <= 256 && a_bytes % 4 == 0 && b_bytes % 4 == 0, "vx_probe_umma: bad args");
  ProbeArgs p{(const uint8_t*)a_img, a_bytes, (const uint8_t*)b_img, b_bytes, lboA, sboA, layA, lboB, sboB, layB,
              make_idesc_bf16(128, N, a_mn, b_mn), ksteps, a_step, b_step, N, out};
  const size_t smem = ((a_bytes + 1023) & ~1023) + ((b_bytes + 1023) & ~1023) + 2048;
  VX_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  VX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vx_probe_umma_ts(const void* a_packed, int K, const void* b_img, int b_bytes, unsigned lboB,
                                unsigned sboB, unsigned layB, int b_mn, int N, int b_step, float* out, void* stream) {
  VX_REQUIRE(N % 16 == 0 && N <= 256 && K % 32 == 0 && K <= 256 && b_bytes % 4 == 0, "vx_probe_umma_ts: bad args");
  ProbeTsArgs p{(const uint32_t*)a_packed, (const uint8_t*)b_img, b_bytes, lboB, sboB, layB,
                make_idesc_bf16(128, N, 0, b_mn), K / 16, b_step, K, N, out};
  VX_CHECK_CUDA(cudaFuncSetAttribute(umma_ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  umma_ts_probe_kernel<<<1, 128, ((b_bytes + 1023) & ~1023) + 2048, (cudaStream_t)stream>>>(p);
  VX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// swizzle: 0 none, 1 32B, 2 64B, 3 128B
extern "C" int vx_probe_tma(const void* base, int rank, const unsigned long long* dims,
                            const unsigned long long* strides_bytes, const unsigned* box, int swizzle,
                            const int* coords, int nbytes, void* out, void* stream) {
  CUtensorMap m;
  uint64_t d[5], s[4];
  for (int i = 0; i < rank; ++i) d[i] = dims[i];
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  static const CUtensorMapSwizzle sw[] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
  if (make_tmap_bf16(&m, base, rank, d, s, box, sw[swizzle & 3])) return 1;
  VX_CHECK_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  tma_probe_kernel<<<1, 128, nbytes + 2048, (cudaStream_t)stream>>>(m, rank, coords[0], coords[1], rank > 2 ? coords[2] : 0,
                                                                   rank > 3 ? coords[3] : 0, nbytes, (uint8_t*)out);
  VX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
