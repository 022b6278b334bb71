// kernels_attn.cu -- flash-style frame attention, SIMT fp32-accumulate version (exact-fp32 verify path and
// generic fallback).  One kernel serves self- and cross-frame attention: kv_frame = frame ^ cross
// (model/xunet.py:114-121).  No score matrix is materialised (the reference materialises (B,h,L,L),
// model/xunet.py:103).  Epilogue fuses the AttnBlock residual (o + h_in)/sqrt(2)  (model/xunet.py:127).
//
// qkv layout: (N, L, 3C) rows [q | k | v], head h occupies columns h*hd .. h*hd+hd of each third.
#include "common.cuh"
#include "kernels.h"
#include <math.h>

template <int LPQ>
__device__ __forceinline__ float part_sum(float v) {
#pragma unroll
  for (int o = LPQ / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------- forward
template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const T* __restrict__ qkv, const T* __restrict__ res,
                                                       T* __restrict__ out, float* __restrict__ lse, int L, int C,
                                                       int heads, int cross) {
  xu_grid_dep_sync();
  constexpr int LPQ = HD / 16;       // lanes cooperating on one query row (16 dims each)
  constexpr int QPB = 128 / LPQ;     // queries per block
  constexpr int KC = 32;             // keys per shared-memory chunk
  __shared__ __align__(16) float Ks[KC][HD];
  __shared__ __align__(16) float Vs[KC][HD];
  const int tid = threadIdx.x;
  const int part = tid % LPQ, ql = tid / LPQ;
  const int h = blockIdx.y, n = blockIdx.z;
  const int nkv = cross ? (n ^ 1) : n;
  const int qi = blockIdx.x * QPB + ql;
  const bool qvalid = qi < L;
  const int C3 = 3 * C;
  const float scale = rsqrtf((float)HD);
  float q[16], acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { q[i] = 0.f; acc[i] = 0.f; }
  if (qvalid) {
    const T* qp = qkv + ((long long)n * L + qi) * C3 + h * HD + part * 16;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float v[4];
      Vec4<T>::ld(qp + i, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) q[i + j] = v[j] * scale;
    }
  }
  float m = -INFINITY, l = 0.f;
  const T* kbase = qkv + (long long)nkv * L * C3 + C + h * HD;
  const T* vbase = qkv + (long long)nkv * L * C3 + 2 * C + h * HD;
  for (int k0 = 0; k0 < L; k0 += KC) {
    __syncthreads();
    for (int idx = tid; idx < KC * HD / 4; idx += 128) {
      const int kk = idx / (HD / 4), dv = (idx - kk * (HD / 4)) * 4;
      float kv[4] = {0.f, 0.f, 0.f, 0.f}, vv[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + kk < L) {
        Vec4<T>::ld(kbase + (long long)(k0 + kk) * C3 + dv, kv);
        Vec4<T>::ld(vbase + (long long)(k0 + kk) * C3 + dv, vv);
      }
      *reinterpret_cast<float4*>(&Ks[kk][dv]) = make_float4(kv[0], kv[1], kv[2], kv[3]);
      *reinterpret_cast<float4*>(&Vs[kk][dv]) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    __syncthreads();
#pragma unroll 1
    for (int j0 = 0; j0 < KC; j0 += 8) {
      float s[8];
      float mx = m;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 k4 = *reinterpret_cast<const float4*>(&Ks[j0 + jj][part * 16 + i]);
          d = fmaf(q[i], k4.x, d); d = fmaf(q[i + 1], k4.y, d); d = fmaf(q[i + 2], k4.z, d); d = fmaf(q[i + 3], k4.w, d);
        }
        d = part_sum<LPQ>(d);
        if (k0 + j0 + jj >= L) d = -INFINITY;
        s[jj] = d;
        mx = fmaxf(mx, d);
      }
      const float corr = __expf(m - mx);   // m=-inf first time -> 0
      l *= corr;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float p = __expf(s[jj] - mx);
        l += p;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 v4 = *reinterpret_cast<const float4*>(&Vs[j0 + jj][part * 16 + i]);
          acc[i] = fmaf(p, v4.x, acc[i]); acc[i + 1] = fmaf(p, v4.y, acc[i + 1]);
          acc[i + 2] = fmaf(p, v4.z, acc[i + 2]); acc[i + 3] = fmaf(p, v4.w, acc[i + 3]);
        }
      }
      m = mx;
    }
  }
  if (qvalid) {
    const float inv = 1.f / l;
    const long long o = ((long long)n * L + qi) * C + h * HD + part * 16;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float r[4], v[4];
      Vec4<T>::ld(res + o + i, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (acc[i + j] * inv + r[j]) * XU_RSQRT2;
      Vec4<T>::st(out + o + i, v);
    }
    if (part == 0) lse[((long long)n * heads + h) * L + qi] = m + __logf(l);
  }
}

// --------------------------------------------------------------------------------------------- backward
// dq kernel: one (query, part) per thread, loop over keys.  Also emits D = rowsum(dO * O) for the dk/dv kernel.
template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ res,
                                                          const T* __restrict__ out, const T* __restrict__ dout,
                                                          const float* __restrict__ lse, float* __restrict__ Dbuf,
                                                          T* __restrict__ dqkv, int L, int C, int heads, int cross) {
  xu_grid_dep_sync();
  constexpr int LPQ = HD / 16;
  constexpr int QPB = 128 / LPQ;
  constexpr int KC = 32;
  __shared__ __align__(16) float Ks[KC][HD];
  __shared__ __align__(16) float Vs[KC][HD];
  const int tid = threadIdx.x;
  const int part = tid % LPQ, ql = tid / LPQ;
  const int h = blockIdx.y, n = blockIdx.z;
  const int nkv = cross ? (n ^ 1) : n;
  const int qi = blockIdx.x * QPB + ql;
  const bool qvalid = qi < L;
  const int C3 = 3 * C;
  const float scale = rsqrtf((float)HD);
  float q[16], dO[16], dq[16];
  float Dp = 0.f, lse_q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { q[i] = 0.f; dO[i] = 0.f; dq[i] = 0.f; }
  if (qvalid) {
    const T* qp = qkv + ((long long)n * L + qi) * C3 + h * HD + part * 16;
    const long long o = ((long long)n * L + qi) * C + h * HD + part * 16;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float v[4], g[4], ov[4], rv[4];
      Vec4<T>::ld(qp + i, v);
      Vec4<T>::ld(dout + o + i, g);
      Vec4<T>::ld(out + o + i, ov);
      Vec4<T>::ld(res + o + i, rv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q[i + j] = v[j];
        dO[i + j] = g[j] * XU_RSQRT2;                       // d(attn out) = dout / sqrt2
        Dp = fmaf(dO[i + j], ov[j] * XU_SQRT2 - rv[j], Dp);  // attn out = out*sqrt2 - res
      }
    }
    lse_q = lse[((long long)n * heads + h) * L + qi];
  }
  const float D = part_sum<LPQ>(Dp);
  if (qvalid && part == 0) Dbuf[((long long)n * heads + h) * L + qi] = D;
  const T* kbase = qkv + (long long)nkv * L * C3 + C + h * HD;
  const T* vbase = qkv + (long long)nkv * L * C3 + 2 * C + h * HD;
  for (int k0 = 0; k0 < L; k0 += KC) {
    __syncthreads();
    for (int idx = tid; idx < KC * HD / 4; idx += 128) {
      const int kk = idx / (HD / 4), dv = (idx - kk * (HD / 4)) * 4;
      float kv[4] = {0.f, 0.f, 0.f, 0.f}, vv[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + kk < L) {
        Vec4<T>::ld(kbase + (long long)(k0 + kk) * C3 + dv, kv);
        Vec4<T>::ld(vbase + (long long)(k0 + kk) * C3 + dv, vv);
      }
      *reinterpret_cast<float4*>(&Ks[kk][dv]) = make_float4(kv[0], kv[1], kv[2], kv[3]);
      *reinterpret_cast<float4*>(&Vs[kk][dv]) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < KC; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 k4 = *reinterpret_cast<const float4*>(&Ks[j][part * 16 + i]);
        float4 v4 = *reinterpret_cast<const float4*>(&Vs[j][part * 16 + i]);
        s = fmaf(q[i], k4.x, s); s = fmaf(q[i + 1], k4.y, s); s = fmaf(q[i + 2], k4.z, s); s = fmaf(q[i + 3], k4.w, s);
        dp = fmaf(dO[i], v4.x, dp); dp = fmaf(dO[i + 1], v4.y, dp); dp = fmaf(dO[i + 2], v4.z, dp); dp = fmaf(dO[i + 3], v4.w, dp);
      }
      s = part_sum<LPQ>(s);
      dp = part_sum<LPQ>(dp);
      float p = (k0 + j < L) ? __expf(s * scale - lse_q) : 0.f;
      const float ds = p * (dp - D) * scale;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 k4 = *reinterpret_cast<const float4*>(&Ks[j][part * 16 + i]);
        dq[i] = fmaf(ds, k4.x, dq[i]); dq[i + 1] = fmaf(ds, k4.y, dq[i + 1]);
        dq[i + 2] = fmaf(ds, k4.z, dq[i + 2]); dq[i + 3] = fmaf(ds, k4.w, dq[i + 3]);
      }
    }
  }
  if (qvalid) {
    T* dqp = dqkv + ((long long)n * L + qi) * C3 + h * HD + part * 16;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float v[4] = {dq[i], dq[i + 1], dq[i + 2], dq[i + 3]};
      Vec4<T>::st(dqp + i, v);
    }
  }
}

// dk/dv kernel: one (key, part) per thread for kv-frame `m`; queries come from frame m ^ cross.
template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                           const float* __restrict__ lse, const float* __restrict__ Dbuf,
                                                           T* __restrict__ dqkv, int L, int C, int heads, int cross) {
  xu_grid_dep_sync();
  constexpr int LPQ = HD / 16;
  constexpr int QPB = 128 / LPQ;
  constexpr int KC = 32;
  __shared__ __align__(16) float Qs[KC][HD];
  __shared__ __align__(16) float Gs[KC][HD];
  __shared__ float sl[KC], sD[KC];
  const int tid = threadIdx.x;
  const int part = tid % LPQ, kl = tid / LPQ;
  const int h = blockIdx.y, mfr = blockIdx.z;
  const int nq = cross ? (mfr ^ 1) : mfr;
  const int ki = blockIdx.x * QPB + kl;
  const bool kvalid = ki < L;
  const int C3 = 3 * C;
  const float scale = rsqrtf((float)HD);
  float k[16], v[16], dk[16], dv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { k[i] = 0.f; v[i] = 0.f; dk[i] = 0.f; dv[i] = 0.f; }
  if (kvalid) {
    const T* kp = qkv + ((long long)mfr * L + ki) * C3 + C + h * HD + part * 16;
    const T* vp = kp + C;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float a[4], b[4];
      Vec4<T>::ld(kp + i, a);
      Vec4<T>::ld(vp + i, b);
#pragma unroll
      for (int j = 0; j < 4; ++j) { k[i + j] = a[j]; v[i + j] = b[j]; }
    }
  }
  const T* qbase = qkv + (long long)nq * L * C3 + h * HD;
  const T* gbase = dout + (long long)nq * L * C + h * HD;
  const float* lbase = lse + ((long long)nq * heads + h) * L;
  const float* dbase = Dbuf + ((long long)nq * heads + h) * L;
  for (int q0 = 0; q0 < L; q0 += KC) {
    __syncthreads();
    for (int idx = tid; idx < KC * HD / 4; idx += 128) {
      const int kk = idx / (HD / 4), dd = (idx - kk * (HD / 4)) * 4;
      float a[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
      if (q0 + kk < L) {
        Vec4<T>::ld(qbase + (long long)(q0 + kk) * C3 + dd, a);
        Vec4<T>::ld(gbase + (long long)(q0 + kk) * C + dd, g);
      }
      *reinterpret_cast<float4*>(&Qs[kk][dd]) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(&Gs[kk][dd]) =
          make_float4(g[0] * XU_RSQRT2, g[1] * XU_RSQRT2, g[2] * XU_RSQRT2, g[3] * XU_RSQRT2);
    }
    if (tid < KC) {
      const bool ok = q0 + tid < L;
      sl[tid] = ok ? lbase[q0 + tid] : INFINITY;   // exp(s - inf) = 0 for padded queries
      sD[tid] = ok ? dbase[q0 + tid] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < KC; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 q4 = *reinterpret_cast<const float4*>(&Qs[j][part * 16 + i]);
        float4 g4 = *reinterpret_cast<const float4*>(&Gs[j][part * 16 + i]);
        s = fmaf(k[i], q4.x, s); s = fmaf(k[i + 1], q4.y, s); s = fmaf(k[i + 2], q4.z, s); s = fmaf(k[i + 3], q4.w, s);
        dp = fmaf(v[i], g4.x, dp); dp = fmaf(v[i + 1], g4.y, dp); dp = fmaf(v[i + 2], g4.z, dp); dp = fmaf(v[i + 3], g4.w, dp);
      }
      s = part_sum<LPQ>(s);
      dp = part_sum<LPQ>(dp);
      const float p = __expf(s * scale - sl[j]);
      const float ds = p * (dp - sD[j]) * scale;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 q4 = *reinterpret_cast<const float4*>(&Qs[j][part * 16 + i]);
        float4 g4 = *reinterpret_cast<const float4*>(&Gs[j][part * 16 + i]);
        dv[i] = fmaf(p, g4.x, dv[i]); dv[i + 1] = fmaf(p, g4.y, dv[i + 1]);
        dv[i + 2] = fmaf(p, g4.z, dv[i + 2]); dv[i + 3] = fmaf(p, g4.w, dv[i + 3]);
        dk[i] = fmaf(ds, q4.x, dk[i]); dk[i + 1] = fmaf(ds, q4.y, dk[i + 1]);
        dk[i + 2] = fmaf(ds, q4.z, dk[i + 2]); dk[i + 3] = fmaf(ds, q4.w, dk[i + 3]);
      }
    }
  }
  if (kvalid) {
    T* dkp = dqkv + ((long long)mfr * L + ki) * C3 + C + h * HD + part * 16;
    T* dvp = dkp + C;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float a[4] = {dk[i], dk[i + 1], dk[i + 2], dk[i + 3]};
      float b[4] = {dv[i], dv[i + 1], dv[i + 2], dv[i + 3]};
      Vec4<T>::st(dkp + i, a);
      Vec4<T>::st(dvp + i, b);
    }
  }
}

template <typename T, int HD>
static void attn_fwd_launch(const AttnArgs& a, cudaStream_t s) {
  constexpr int QPB = 128 / (HD / 16);
  dim3 grid(cdiv(a.L, QPB), a.heads, a.N);
  xu_launch(attn_fwd_kernel<T, HD>, grid, 128, 0, s, (const T*)a.qkv, (const T*)a.res, (T*)a.out, a.lse, a.L, a.C, a.heads, a.cross);
}
template <typename T, int HD>
static void attn_bwd_launch(const AttnArgs& a, cudaStream_t s) {
  constexpr int QPB = 128 / (HD / 16);
  dim3 grid(cdiv(a.L, QPB), a.heads, a.N);
  xu_launch(attn_bwd_dq_kernel<T, HD>, grid, 128, 0, s, (const T*)a.qkv, (const T*)a.res, (const T*)a.out, (const T*)a.dout, a.lse,
                                                 a.dscratch, (T*)a.dqkv, a.L, a.C, a.heads, a.cross);
  xu_launch(attn_bwd_dkv_kernel<T, HD>, grid, 128, 0, s, (const T*)a.qkv, (const T*)a.dout, a.lse, a.dscratch, (T*)a.dqkv, a.L, a.C,
                                                  a.heads, a.cross);
}

template <typename T>
static bool attn_dispatch(const AttnArgs& a, cudaStream_t s, bool bwd) {
  const int hd = a.C / a.heads;
  switch (hd) {
    case 16: bwd ? attn_bwd_launch<T, 16>(a, s) : attn_fwd_launch<T, 16>(a, s); return true;
    case 32: bwd ? attn_bwd_launch<T, 32>(a, s) : attn_fwd_launch<T, 32>(a, s); return true;
    case 64: bwd ? attn_bwd_launch<T, 64>(a, s) : attn_fwd_launch<T, 64>(a, s); return true;
    case 128: bwd ? attn_bwd_launch<T, 128>(a, s) : attn_fwd_launch<T, 128>(a, s); return true;
    default: xu_set_kernel_error("attention: head_dim must be 16, 32, 64 or 128"); return false;
  }
}

void launch_attn_fwd_simt(int dtype, const AttnArgs& a, cudaStream_t s) {
  if (dtype == XU_F32) attn_dispatch<float>(a, s, false);
  else attn_dispatch<bf16>(a, s, false);
}
void launch_attn_bwd_simt(int dtype, const AttnArgs& a, cudaStream_t s) {
  if (dtype == XU_F32) attn_dispatch<float>(a, s, true);
  else attn_dispatch<bf16>(a, s, true);
}
