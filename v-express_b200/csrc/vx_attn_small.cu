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
  pdl_enter();
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


// ---- tensor-core version for f <= 16: one warp per (b, pixel, head), everything in mma.sync fragments.
// S (16 queries x 16 keys) = Q K^T with m16n8k16 bf16 MMAs whose A / B fragments are read straight from global
// memory (row = frame, 4 lanes x 4 bytes = one 16-byte piece of the head's row; the second load of a pair takes the
// other half of the same 32-byte sector from L1), softmax across the 4 lanes of a quad, P packed to bf16 in place
// (the S accumulator layout is the A layout of the next MMA) and O = P V with V^T fragments produced by
// movmatrix.trans -- no shared memory, no transposing copies, ~11 MMAs instead of ~1300 FMAs per lane and head.
__device__ __forceinline__ uint32_t ldg_nc_b32(const __nv_bfloat16* p) {
  uint32_t r;
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// One 16 x 16 attention in fragments.  q0/q1, k0/k1, v0/v1, o0/o1 point at (row g | row g + 8, column 2t) of the
// head's slice; okq* / okk* say whether those query / key rows exist, nk = number of valid keys (<= 16).
template <int HD>
__device__ __forceinline__ void attn16_mma(const __nv_bfloat16* q0, const __nv_bfloat16* q1, const __nv_bfloat16* k0,
                                           const __nv_bfloat16* k1, const __nv_bfloat16* v0, const __nv_bfloat16* v1,
                                           __nv_bfloat16* o0, __nv_bfloat16* o1, bool okq0, bool okq1, bool okk0,
                                           bool okk1, int nk, float scale, int t) {
  constexpr int NB = HD / 8;          // 8-column blocks of the head
  constexpr int KS = (NB + 1) / 2;    // k-steps of 16 (the last one half empty when NB is odd)
  uint32_t qa[2 * KS][2], kb[2 * KS][2];
#pragma unroll
  for (int j = 0; j < 2 * KS; ++j) {
    const bool in = j < NB;
    qa[j][0] = (in && okq0) ? ldg_nc_b32(q0 + 8 * j) : 0u;
    qa[j][1] = (in && okq1) ? ldg_nc_b32(q1 + 8 * j) : 0u;
    kb[j][0] = (in && okk0) ? ldg_nc_b32(k0 + 8 * j) : 0u;
    kb[j][1] = (in && okk1) ? ldg_nc_b32(k1 + 8 * j) : 0u;
  }
  uint32_t vr[NB][2];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    vr[j][0] = okk0 ? ldg_nc_b32(v0 + 8 * j) : 0u;
    vr[j][1] = okk1 ? ldg_nc_b32(v1 + 8 * j) : 0u;
  }
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};   // keys 0..7 / 8..15
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    mma_bf16_16816(s0, qa[2 * ks][0], qa[2 * ks][1], qa[2 * ks + 1][0], qa[2 * ks + 1][1], kb[2 * ks][0], kb[2 * ks + 1][0]);
    mma_bf16_16816(s1, qa[2 * ks][0], qa[2 * ks][1], qa[2 * ks + 1][0], qa[2 * ks + 1][1], kb[2 * ks][1], kb[2 * ks + 1][1]);
  }
  // accumulator element i of s0/s1: row g (i < 2) or g + 8, key 2t + (i & 1) (+ 8 for s1)
  const float sl = scale * 1.4426950408889634f;
  const bool kv00 = 2 * t < nk, kv01 = 2 * t + 1 < nk, kv10 = 2 * t + 8 < nk, kv11 = 2 * t + 9 < nk;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s0[i] = ((i & 1) ? kv01 : kv00) ? s0[i] * sl : -INFINITY;
    s1[i] = ((i & 1) ? kv11 : kv10) ? s1[i] * sl : -INFINITY;
  }
  float m0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));
  float m1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s0[i] = exp2f(s0[i] - (i < 2 ? m0 : m1));    // key 0 is always valid, so the maxima are finite
    s1[i] = exp2f(s1[i] - (i < 2 ? m0 : m1));
  }
  float sum0 = s0[0] + s0[1] + s1[0] + s1[1], sum1 = s0[2] + s0[3] + s1[2] + s1[3];
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
  const uint32_t pa0 = pack_bf16(s0[0], s0[1]), pa1 = pack_bf16(s0[2], s0[3]);
  const uint32_t pa2 = pack_bf16(s1[0], s1[1]), pa3 = pack_bf16(s1[2], s1[3]);
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t b0 = movmatrix_trans(vr[j][0]);   // (keys 2t, 2t+1; d = 8j + g)
    const uint32_t b1 = movmatrix_trans(vr[j][1]);   // (keys 2t+8, 2t+9)
    mma_bf16_16816(o, pa0, pa1, pa2, pa3, b0, b1);
    if (okq0) *reinterpret_cast<uint32_t*>(o0 + 8 * j) = pack_bf16(o[0] * inv0, o[1] * inv0);
    if (okq1) *reinterpret_cast<uint32_t*>(o1 + 8 * j) = pack_bf16(o[2] * inv1, o[3] * inv1);
  }
}

template <int HD>
__global__ void __launch_bounds__(128) temporal_attn_mma_kernel(const TemporalArgs p) {
  pdl_enter();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long item = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);   // (b, pixel, head), head fastest
  if (item >= (long long)p.b * p.HW * p.heads) return;
  const int head = (int)(item % p.heads);
  const long long bp = item / p.heads;
  const int px = (int)(bp % p.HW);
  const int bb = (int)(bp / p.HW);
  const long long row0 = (long long)bb * p.f * p.HW + px;
  const bool ok0 = g < p.f, ok1 = g + 8 < p.f;           // frames g and g + 8 of this lane's fragment rows
  const long long r0 = row0 + (long long)(ok0 ? g : 0) * p.HW, r1 = row0 + (long long)(ok1 ? g + 8 : 0) * p.HW;
  const int col = head * HD + 2 * t;
  attn16_mma<HD>(p.q + r0 * p.ld + col, p.q + r1 * p.ld + col, p.k + r0 * p.ld + col, p.k + r1 * p.ld + col,
                 p.v + r0 * p.ld + col, p.v + r1 * p.ld + col, p.out + r0 * p.ldo + col, p.out + r1 * p.ldo + col,
                 ok0, ok1, ok0, ok1, p.f, p.scale, t);
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
  pdl_enter();
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

// tensor-core version: one warp per (16 consecutive query rows of one frame, head); the frame's Lk tokens are the
// key rows 0..Lk-1 of the 16-key tile (the rest masked), read once per 16 queries instead of once per query.
template <int HD>
__global__ void __launch_bounds__(128) smallkv_attn_mma_kernel(const SmallKvArgs p) {
  pdl_enter();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long item = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);   // (row tile, head), head fastest
  if (item >= (p.rows / 16) * p.heads) return;
  const int head = (int)(item % p.heads);
  const long long r0 = (item / p.heads) * 16 + g, r1 = r0 + 8;
  const long long frame = r0 / p.rows_per_frame;
  const int col = head * HD + 2 * t;
  const bool okk = g < p.Lk;
  const long long kr = frame * p.Lk + (okk ? g : 0);
  attn16_mma<HD>(p.q + r0 * p.ldq + col, p.q + r1 * p.ldq + col, p.k + kr * p.ldkv + col, p.k + kr * p.ldkv + col,
                 p.v + kr * p.ldkv + col, p.v + kr * p.ldkv + col, p.out + r0 * p.ldo + col, p.out + r1 * p.ldo + col,
                 true, true, okk, false, p.Lk, p.scale, t);
}

}  // namespace vx

using namespace vx;

// q/k/v: [(b f hw), ld] slices (same ld); out [(b f hw), ldo]; attention over f for each (b, pixel, head)
extern "C" int vx_temporal_attention(const void* q, const void* k, const void* v, long long ld, void* out,
                                     long long ldo, int b, int f, int HW, int heads, int hd, void* stream) {
  VX_REQUIRE(f >= 1 && f <= 32 && ld % 8 == 0 && ldo % 8 == 0, "vx_temporal_attention: bad f=%d", f);
  TemporalArgs a{(const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ld, (__nv_bfloat16*)out,
                 ldo, b, f, HW, heads, hd, 1.0f / sqrtf((float)hd)};
  auto st = (cudaStream_t)stream;
  static const bool temporal_v1 = getenv("VX_TEMPORAL_V1") != nullptr;   // A/B switch, read once
  if (f <= 16 && !temporal_v1) {
    const long long warps = (long long)b * HW * heads;
    const unsigned grid_m = (unsigned)((warps + 3) / 4);
    switch (hd) {
      case 8: launch_k(temporal_attn_mma_kernel<8>, dim3(grid_m), dim3(128), 0, st, a); break;
      case 16: launch_k(temporal_attn_mma_kernel<16>, dim3(grid_m), dim3(128), 0, st, a); break;
      case 32: launch_k(temporal_attn_mma_kernel<32>, dim3(grid_m), dim3(128), 0, st, a); break;
      case 40: launch_k(temporal_attn_mma_kernel<40>, dim3(grid_m), dim3(128), 0, st, a); break;
      case 80: launch_k(temporal_attn_mma_kernel<80>, dim3(grid_m), dim3(128), 0, st, a); break;
      case 160: launch_k(temporal_attn_mma_kernel<160>, dim3(grid_m), dim3(128), 0, st, a); break;
      default: return fail("vx_temporal_attention: head dim %d not instantiated (8,16,32,40,80,160)", hd);
    }
    VX_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int G = f <= 16 ? 2 : 1;
  VX_REQUIRE(heads % G == 0, "vx_temporal_attention: heads=%d must be even", heads);
  const size_t smem = (size_t)4 * 2 * G * f * hd * (hd <= 80 ? 4 : 2);
  VX_REQUIRE(smem <= 200 * 1024, "vx_temporal_attention: smem %zu", smem);
  const long long items = (long long)b * HW * (heads / G);
  const unsigned grid = (unsigned)((items + 3) / 4);
#define TA_LAUNCH(HD)                                                                                               \
  do {                                                                                                              \
    static bool cfg = false;                                                                                        \
    if (!cfg) {                                                                                                     \
      VX_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                         200 * 1024));                                                              \
      cfg = true;                                                                                                   \
    }                                                                                                               \
    launch_k(temporal_attn_kernel<HD>, dim3(grid), dim3(128), smem, st, a);                                                           \
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
  auto st = (cudaStream_t)stream;
  static const bool smallkv_v1 = getenv("VX_SMALLKV_V1") != nullptr;     // A/B switch, read once
  if (rows_per_frame % 16 == 0 && rows % 16 == 0 && !smallkv_v1 &&
      (hd == 8 || hd == 40 || hd == 80 || hd == 160)) {
    const long long warps = rows / 16 * heads;
    const unsigned grid_m = (unsigned)((warps + 3) / 4);
    switch (hd) {
      case 8: launch_k(smallkv_attn_mma_kernel<8>, dim3(grid_m), dim3(128), 0, st, a); break;
      case 40: launch_k(smallkv_attn_mma_kernel<40>, dim3(grid_m), dim3(128), 0, st, a); break;
      case 80: launch_k(smallkv_attn_mma_kernel<80>, dim3(grid_m), dim3(128), 0, st, a); break;
      default: launch_k(smallkv_attn_mma_kernel<160>, dim3(grid_m), dim3(128), 0, st, a); break;
    }
    VX_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const long long n = rows * heads;
  launch_k(smallkv_attn_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, a);
  VX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
