// Small-sequence attention kernels (CUDA cores; the contraction lengths are 5..32, far below a tensor-core tile):
//   * temporal self-attention of the motion modules: for every (b, pixel, head) an f x f attention over the
//     frames of the window, read straight from the (b f)(h w) c token layout with stride HW*ld between frames
//     (reference modules/motion_module.py:351-388: "(b f) d c -> (b d) f c", SDPA over f, and back) --
//     no transposing copies;
//   * audio cross-attention: every query row attends to the Lk (=5) audio tokens of its frame
//     (reference modules/mutual_self_attention.py:229-242 -> diffusers AttnProcessor2_0).
// Both are HBM-bound: each q/k/v element is read once, each output written once, fp32 math, exact softmax.
#include "vx_host.h"
#include "vx_ptx.cuh"

namespace vx {

// one warp per (b, pixel, head); 4 warps per CTA = 4 consecutive heads of one pixel (contiguous channels)
struct TemporalArgs {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v; long long ld;  // rows = (b f hw)
  __nv_bfloat16* out; long long ldo;
  int b, f, HW, heads, hd;
  float scale;
};

__global__ void __launch_bounds__(128) temporal_attn_kernel(const TemporalArgs p) {
  extern __shared__ uint8_t sm_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hdp = p.hd + 8;  // padded row (bank spread)
  const size_t per_warp = (size_t)3 * p.f * hdp * sizeof(__nv_bfloat16) + (size_t)p.f * p.f * sizeof(float);
  uint8_t* base = sm_raw + warp * ((per_warp + 15) & ~size_t(15));
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* sk = sq + p.f * hdp;
  __nv_bfloat16* sv = sk + p.f * hdp;
  float* ss = reinterpret_cast<float*>(sv + p.f * hdp);

  const long long item = (long long)blockIdx.x * 4 + warp;  // (b, pixel, head), head fastest
  const long long nitems = (long long)p.b * p.HW * p.heads;
  if (item >= nitems) return;
  const int head = (int)(item % p.heads);
  const long long bp = item / p.heads;
  const int px = (int)(bp % p.HW);
  const int bb = (int)(bp / p.HW);
  const int vec = p.hd / 8;
  const long long row0 = (long long)bb * p.f * p.HW + px;
  for (int idx = lane; idx < p.f * vec; idx += 32) {
    const int fr = idx / vec, c = (idx % vec) * 8;
    const long long off = (row0 + (long long)fr * p.HW) * p.ld + head * p.hd + c;
    *reinterpret_cast<uint4*>(sq + fr * hdp + c) = *reinterpret_cast<const uint4*>(p.q + off);
    *reinterpret_cast<uint4*>(sk + fr * hdp + c) = *reinterpret_cast<const uint4*>(p.k + off);
    *reinterpret_cast<uint4*>(sv + fr * hdp + c) = *reinterpret_cast<const uint4*>(p.v + off);
  }
  __syncwarp();
  for (int idx = lane; idx < p.f * p.f; idx += 32) {
    const int i = idx / p.f, j = idx % p.f;
    float acc = 0.f;
    for (int c = 0; c < p.hd; c += 2) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sq + i * hdp + c));
      const float2 bk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sk + j * hdp + c));
      acc += a.x * bk.x + a.y * bk.y;
    }
    ss[idx] = acc * p.scale;
  }
  __syncwarp();
  if (lane < p.f) {
    float mx = -INFINITY;
    for (int j = 0; j < p.f; ++j) mx = fmaxf(mx, ss[lane * p.f + j]);
    float sum = 0.f;
    for (int j = 0; j < p.f; ++j) {
      const float e = __expf(ss[lane * p.f + j] - mx);
      ss[lane * p.f + j] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    for (int j = 0; j < p.f; ++j) ss[lane * p.f + j] *= inv;
  }
  __syncwarp();
  for (int idx = lane; idx < p.f * vec; idx += 32) {
    const int i = idx / vec, c = (idx % vec) * 8;
    float o[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = 0.f;
    for (int j = 0; j < p.f; ++j) {
      const float w = ss[i * p.f + j];
      const uint4 u = *reinterpret_cast<const uint4*>(sv + j * hdp + c);
      const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 x = unpack_bf16(ww[t]);
        o[2 * t] += w * x.x;
        o[2 * t + 1] += w * x.y;
      }
    }
    const long long off = (row0 + (long long)i * p.HW) * p.ldo + head * p.hd + c;
    *reinterpret_cast<uint4*>(p.out + off) =
        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  }
}

// one thread per (query row, head); K/V of the frame ([Lk, C], Lk <= 8) stay cache resident
struct SmallKvArgs {
  const __nv_bfloat16* q; long long ldq;
  const __nv_bfloat16* k; const __nv_bfloat16* v; long long ldkv;  // rows = frame*Lk + token
  __nv_bfloat16* out; long long ldo;
  long long rows; int rows_per_frame, heads, hd, Lk;
  float scale;
};

__global__ void smallkv_attn_kernel(const SmallKvArgs p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.rows * p.heads) return;
  const int head = (int)(idx % p.heads);
  const long long row = idx / p.heads;
  const long long frame = row / p.rows_per_frame;
  const __nv_bfloat16* qp = p.q + row * p.ldq + head * p.hd;
  const __nv_bfloat16* kp = p.k + frame * p.Lk * p.ldkv + head * p.hd;
  const __nv_bfloat16* vp = p.v + frame * p.Lk * p.ldkv + head * p.hd;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  for (int c = 0; c < p.hd; c += 8) {
    const uint4 uq = *reinterpret_cast<const uint4*>(qp + c);
    const uint32_t wq[4] = {uq.x, uq.y, uq.z, uq.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < p.Lk) {
        const uint4 uk = *reinterpret_cast<const uint4*>(kp + j * p.ldkv + c);
        const uint32_t wk[4] = {uk.x, uk.y, uk.z, uk.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 a = unpack_bf16(wq[t]), b = unpack_bf16(wk[t]);
          s[j] += a.x * b.x + a.y * b.y;
        }
      }
    }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < p.Lk) mx = fmaxf(mx, s[j] * p.scale);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < p.Lk) {
      s[j] = __expf(s[j] * p.scale - mx);
      sum += s[j];
    }
  const float inv = 1.f / sum;
  __nv_bfloat16* op = p.out + row * p.ldo + head * p.hd;
  for (int c = 0; c < p.hd; c += 8) {
    float o[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < p.Lk) {
        const float w = s[j] * inv;
        const uint4 uv = *reinterpret_cast<const uint4*>(vp + j * p.ldkv + c);
        const uint32_t wv[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 x = unpack_bf16(wv[t]);
          o[2 * t] += w * x.x;
          o[2 * t + 1] += w * x.y;
        }
      }
    }
    *reinterpret_cast<uint4*>(op + c) =
        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  }
}

}  // namespace vx

using namespace vx;

// q/k/v: [(b f hw), ld] slices (same ld); out [(b f hw), ldo]; attention over f for each (b, pixel, head)
extern "C" int vx_temporal_attention(const void* q, const void* k, const void* v, long long ld, void* out,
                                     long long ldo, int b, int f, int HW, int heads, int hd, void* stream) {
  VX_REQUIRE(hd % 8 == 0 && f >= 1 && f <= 32 && ld % 8 == 0 && ldo % 8 == 0, "vx_temporal_attention: bad hd=%d f=%d", hd, f);
  TemporalArgs a{(const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ld, (__nv_bfloat16*)out,
                 ldo, b, f, HW, heads, hd, 1.0f / sqrtf((float)hd)};
  const size_t per_warp = (((size_t)3 * f * (hd + 8) * 2 + (size_t)f * f * 4) + 15) & ~size_t(15);
  const size_t smem = per_warp * 4;
  VX_REQUIRE(smem <= 200 * 1024, "vx_temporal_attention: smem %zu", smem);
  static bool cfg = false;
  if (!cfg) {
    VX_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cfg = true;
  }
  const long long items = (long long)b * HW * heads;
  temporal_attn_kernel<<<(unsigned)((items + 3) / 4), 128, smem, (cudaStream_t)stream>>>(a);
  VX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// q: [rows, ldq]; k/v: [frames*Lk, ldkv]; frame of a row = row / rows_per_frame
extern "C" int vx_smallkv_attention(const void* q, long long ldq, const void* k, const void* v, long long ldkv,
                                    void* out, long long ldo, long long rows, int rows_per_frame, int heads, int hd,
                                    int Lk, void* stream) {
  VX_REQUIRE(hd % 8 == 0 && Lk >= 1 && Lk <= 8, "vx_smallkv_attention: hd=%d Lk=%d", hd, Lk);
  SmallKvArgs a{(const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldkv,
                (__nv_bfloat16*)out, ldo, rows, rows_per_frame, heads, hd, Lk, 1.0f / sqrtf((float)hd)};
  const long long n = rows * heads;
  smallkv_attn_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  VX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
