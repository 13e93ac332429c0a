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

// One warp per (b, pixel); lane = (head slot, query frame): G = 32 / fpad heads are processed per pass
// (fpad = 16 for f <= 16, else 32).  K/V of the pass are staged in shared memory (bf16) and read as warp
// broadcasts; each thread keeps its q row packed in registers, does the f scores, an in-thread softmax and the
// P.V product -- no shuffles, fp32 math, FMA-bound.
struct TemporalArgs {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v; long long ld;  // rows = (b f hw)
  __nv_bfloat16* out; long long ldo;
  int b, f, HW, heads, hd;
  float scale;
};

template <int HD>
__global__ void __launch_bounds__(128) temporal_attn_kernel(const TemporalArgs p) {
  extern __shared__ uint8_t sm_raw[];
  constexpr int VEC = HD / 8;
  constexpr bool F32 = HD <= 80;                       // K/V staged as fp32 (no unpack in the inner loops)
  constexpr int ESZ = F32 ? 4 : 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fpad = p.f <= 16 ? 16 : 32;
  const int G = 32 / fpad;                             // heads per warp
  const int passes = p.heads / G;
  const size_t per_warp = (size_t)2 * G * p.f * HD * ESZ;
  uint8_t* base = sm_raw + warp * per_warp;
  const long long item = (long long)blockIdx.x * 4 + warp;   // (b, pixel, head group), head group fastest
  if (item >= (long long)p.b * p.HW * passes) return;
  const int h0 = (int)(item % passes) * G;
  const long long bp = item / passes;
  const int px = (int)(bp % p.HW);
  const int bb = (int)(bp / p.HW);
  const long long row0 = (long long)bb * p.f * p.HW + px;
  const int hs = lane / fpad, qi = lane % fpad;        // head slot, query frame of this thread
  const bool active = qi < p.f;
  const int head = h0 + hs;
  // stage K, V of heads [h0, h0+G): [G][f][HD]
  for (int idx = lane; idx < G * p.f * VEC; idx += 32) {
    const int c = (idx % VEC) * 8;
    const int fr = (idx / VEC) % p.f;
    const int g = idx / (VEC * p.f);
    const long long off = (row0 + (long long)fr * p.HW) * p.ld + (h0 + g) * HD + c;
    const uint4 uk = *reinterpret_cast<const uint4*>(p.k + off);
    const uint4 uv = *reinterpret_cast<const uint4*>(p.v + off);
    const int e = (g * p.f + fr) * HD + c;
    if (F32) {
      float* fk = reinterpret_cast<float*>(base) + e;
      float* fv = reinterpret_cast<float*>(base) + G * p.f * HD + e;
      const uint32_t wk[4] = {uk.x, uk.y, uk.z, uk.w}, wv[4] = {uv.x, uv.y, uv.z, uv.w};
      float a[8], bq[8];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 x = unpack_bf16(wk[t]), y = unpack_bf16(wv[t]);
        a[2 * t] = x.x; a[2 * t + 1] = x.y; bq[2 * t] = y.x; bq[2 * t + 1] = y.y;
      }
      *reinterpret_cast<float4*>(fk) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(fk + 4) = make_float4(a[4], a[5], a[6], a[7]);
      *reinterpret_cast<float4*>(fv) = make_float4(bq[0], bq[1], bq[2], bq[3]);
      *reinterpret_cast<float4*>(fv + 4) = make_float4(bq[4], bq[5], bq[6], bq[7]);
    } else {
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + e) = uk;
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + G * p.f * HD + e) = uv;
    }
  }
  const long long qoff = (row0 + (long long)(active ? qi : 0) * p.HW) * p.ld + head * HD;
  float s[32];
  float mx = -INFINITY;
  if (F32) {
    float qf[HD];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const uint4 u = *reinterpret_cast<const uint4*>(p.q + qoff + i * 8);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 x = unpack_bf16(w4[t]);
        qf[i * 8 + 2 * t] = x.x * p.scale;
        qf[i * 8 + 2 * t + 1] = x.y * p.scale;
      }
    }
    __syncwarp();
    const float* kh = reinterpret_cast<const float*>(base) + hs * p.f * HD;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {   // two keys per step: four independent FMA chains per thread
      if (j >= p.f) break;
      const bool two = j + 1 < p.f;
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
      for (int i = 0; i < HD / 4; ++i) {
        const float4 kk = *reinterpret_cast<const float4*>(kh + j * HD + i * 4);
        const float4 k2 = *reinterpret_cast<const float4*>(kh + (two ? j + 1 : j) * HD + i * 4);
        a0 = fmaf(qf[4 * i], kk.x, a0);
        b0 = fmaf(qf[4 * i], k2.x, b0);
        a1 = fmaf(qf[4 * i + 1], kk.y, a1);
        b1 = fmaf(qf[4 * i + 1], k2.y, b1);
        a0 = fmaf(qf[4 * i + 2], kk.z, a0);
        b0 = fmaf(qf[4 * i + 2], k2.z, b0);
        a1 = fmaf(qf[4 * i + 3], kk.w, a1);
        b1 = fmaf(qf[4 * i + 3], k2.w, b1);
      }
      s[j] = a0 + a1;
      mx = fmaxf(mx, s[j]);
      if (two) {
        s[j + 1] = b0 + b1;
        mx = fmaxf(mx, s[j + 1]);
      }
    }
  } else {
    uint4 qreg[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) qreg[i] = *reinterpret_cast<const uint4*>(p.q + qoff + i * 8);
    __syncwarp();
    const __nv_bfloat16* kh = reinterpret_cast<const __nv_bfloat16*>(base) + hs * p.f * HD;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j >= p.f) break;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const uint4 kk = *reinterpret_cast<const uint4*>(kh + j * HD + i * 8);
        const uint32_t a[4] = {qreg[i].x, qreg[i].y, qreg[i].z, qreg[i].w};
        const uint32_t bq[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 x = unpack_bf16(a[t]), y = unpack_bf16(bq[t]);
          acc = fmaf(x.x, y.x, acc);
          acc = fmaf(x.y, y.y, acc);
        }
      }
      s[j] = acc * p.scale;
      mx = fmaxf(mx, s[j]);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < p.f) {
      s[j] = __expf(s[j] - mx);
      sum += s[j];
    }
  }
  const float inv = 1.f / sum;
  __nv_bfloat16* op = p.out + (row0 + (long long)qi * p.HW) * p.ldo + head * HD;
#pragma unroll 1
  for (int i = 0; i < VEC; ++i) {
    float o[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = 0.f;
    if (F32) {
      const float* vh = reinterpret_cast<const float*>(base) + (G + hs) * p.f * HD + i * 8;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < p.f) {
          const float4 v0 = *reinterpret_cast<const float4*>(vh + j * HD);
          const float4 v1 = *reinterpret_cast<const float4*>(vh + j * HD + 4);
          o[0] = fmaf(s[j], v0.x, o[0]); o[1] = fmaf(s[j], v0.y, o[1]); o[2] = fmaf(s[j], v0.z, o[2]);
          o[3] = fmaf(s[j], v0.w, o[3]); o[4] = fmaf(s[j], v1.x, o[4]); o[5] = fmaf(s[j], v1.y, o[5]);
          o[6] = fmaf(s[j], v1.z, o[6]); o[7] = fmaf(s[j], v1.w, o[7]);
        }
      }
    } else {
      const __nv_bfloat16* vh = reinterpret_cast<const __nv_bfloat16*>(base) + (G + hs) * p.f * HD + i * 8;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < p.f) {
          const uint4 vv = *reinterpret_cast<const uint4*>(vh + j * HD);
          const uint32_t w4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 x = unpack_bf16(w4[t]);
            o[2 * t] = fmaf(s[j], x.x, o[2 * t]);
            o[2 * t + 1] = fmaf(s[j], x.y, o[2 * t + 1]);
          }
        }
      }
    }
    if (active)
      *reinterpret_cast<uint4*>(op + i * 8) = make_uint4(pack_bf16(o[0] * inv, o[1] * inv), pack_bf16(o[2] * inv, o[3] * inv),
                                                         pack_bf16(o[4] * inv, o[5] * inv), pack_bf16(o[6] * inv, o[7] * inv));
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
  VX_REQUIRE(f >= 1 && f <= 32 && ld % 8 == 0 && ldo % 8 == 0, "vx_temporal_attention: bad f=%d", f);
  TemporalArgs a{(const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ld, (__nv_bfloat16*)out,
                 ldo, b, f, HW, heads, hd, 1.0f / sqrtf((float)hd)};
  const int G = f <= 16 ? 2 : 1;
  VX_REQUIRE(heads % G == 0, "vx_temporal_attention: heads=%d must be even", heads);
  const size_t smem = (size_t)4 * 2 * G * f * hd * (hd <= 80 ? 4 : 2);
  VX_REQUIRE(smem <= 200 * 1024, "vx_temporal_attention: smem %zu", smem);
  const long long items = (long long)b * HW * (heads / G);
  const unsigned grid = (unsigned)((items + 3) / 4);
  auto st = (cudaStream_t)stream;
#define TA_LAUNCH(HD)                                                                                               \
  do {                                                                                                              \
    static bool cfg = false;                                                                                        \
    if (!cfg) {                                                                                                     \
      VX_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                         200 * 1024));                                                              \
      cfg = true;                                                                                                   \
    }                                                                                                               \
    temporal_attn_kernel<HD><<<grid, 128, smem, st>>>(a);                                                           \
  } while (0)
  switch (hd) {
    case 8: TA_LAUNCH(8); break;
    case 16: TA_LAUNCH(16); break;
    case 32: TA_LAUNCH(32); break;
    case 40: TA_LAUNCH(40); break;
    case 80: TA_LAUNCH(80); break;
    case 160: TA_LAUNCH(160); break;
    default: return fail("vx_temporal_attention: head dim %d not instantiated (8,16,32,40,80,160)", hd);
  }
#undef TA_LAUNCH
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
