// kernels_conv.cu -- SIMT implicit-GEMM convolution (forward / data-gradient / weight-gradient).
//
// This is the exact-fp32 path ("verify mode", XUNET_DTYPE_F32) and the fallback for shapes the tcgen05
// kernel (conv_tc.cu) does not take (Cin=3 input conv, Cout=3 output conv, strided pose convs).
// Replaces flax nn.Conv / nn.Dense / nn.DenseGeneral as called from model/xunet.py:59,81,85-89,91,
// 100-102,199,229,276 and their XLA-autodiff gradients (train.py:70).
#include "common.cuh"
#include "kernels.h"

__device__ __forceinline__ long long conv_waddr(int tap, int ci, int co, int taps, int wCi, int segw) {
  int seg = co / segw;
  return (long long)seg * taps * wCi * segw + ((long long)tap * wCi + ci) * segw + (co - seg * segw);
}

template <typename T, int BM, int BN>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
  xu_grid_dep_sync();
  constexpr int BK = 16;
  constexpr int TX = BN / 4;
  __shared__ float As[BK][BM + 1];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ int pn[BM], py[BM], px[BM];
  const T* __restrict__ x = reinterpret_cast<const T*>(a.x);
  const int tid = threadIdx.x;
  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  for (int i = tid; i < BM; i += 256) {
    long long m = m0 + i;
    if (m < M) {
      int ox = (int)(m % a.Wo);
      long long t = m / a.Wo;
      px[i] = ox; py[i] = (int)(t % a.Ho); pn[i] = (int)(t / a.Ho);
    } else {
      pn[i] = -1; py[i] = 0; px[i] = 0;
    }
  }
  __syncthreads();
  const int taps = a.ks * a.ks;
  const int Ktot = taps * a.Ci;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tx = tid % TX, ty = tid / TX;

  for (int k0 = 0; k0 < Ktot; k0 += BK) {
    {  // A tile: BM pixels x 16 k
      const int kk = tid & 15;
      const int k = k0 + kk;
      const bool kvalid = k < Ktot;
      int tap = 0, ci = 0;
      if (kvalid) { tap = k / a.Ci; ci = k - tap * a.Ci; }
      const int dy = tap / a.ks, dx = tap - dy * a.ks;
#pragma unroll
      for (int r = 0; r < BM / 16; ++r) {
        const int i = (tid >> 4) + 16 * r;
        float v = 0.f;
        const int n = pn[i];
        if (kvalid && n >= 0) {
          int iy, ix;
          bool ok;
          if (a.mode == 0) {
            iy = py[i] * a.stride + dy - a.pad_h;
            ix = px[i] * a.stride + dx - a.pad_w;
            ok = iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi;
          } else {
            int ty_ = py[i] + a.pad_h - dy, tx_ = px[i] + a.pad_w - dx;
            ok = ty_ >= 0 && tx_ >= 0 && (ty_ % a.stride) == 0 && (tx_ % a.stride) == 0;
            iy = ty_ / a.stride; ix = tx_ / a.stride;
            ok = ok && iy < a.Hi && ix < a.Wi;
          }
          if (ok) v = ldf(x + (((long long)n * a.Hi + iy) * a.Wi + ix) * a.Ci + ci);
        }
        As[kk][i] = v;
      }
    }
    {  // B tile: 16 k x BN
#pragma unroll
      for (int r = 0; r < (BK * BN) / 256; ++r) {
        const int idx = tid + 256 * r;
        const int kk = idx / BN, nn = idx - kk * BN;
        const int k = k0 + kk, col = n0 + nn;
        float v = 0.f;
        if (k < Ktot && col < a.Co) {
          int tap = k / a.Ci, kc = k - tap * a.Ci;
          long long wa = (a.mode == 0) ? conv_waddr(tap, kc, col, taps, a.wCi, a.segw)
                                       : conv_waddr(tap, col, kc, taps, a.wCi, a.segw);
          v = a.w[wa];
        }
        Bs[kk][nn] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  T* __restrict__ y = reinterpret_cast<T*>(a.y);
  const T* __restrict__ res = reinterpret_cast<const T*>(a.res);
  const int c0 = n0 + tx * 4;
  const bool vec = (a.Co % 4 == 0) && (c0 + 3 < a.Co);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j];
      if (a.bias != nullptr && c0 + j < a.Co) v[j] += a.bias[c0 + j];
    }
    const long long base = m * a.Co + c0;
    if (vec) {
      if (res != nullptr) {
        float rv[4];
        Vec4<T>::ld(res + base, rv);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += rv[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= a.alpha;
      if (a.accumulate) {
        float ov[4];
        Vec4<T>::ld(y + base, ov);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += ov[j];
      }
      Vec4<T>::st(y + base, v);
    } else {
      for (int j = 0; j < 4; ++j) {
        if (c0 + j >= a.Co) break;
        float o = v[j];
        if (res != nullptr) o += ldf(res + base + j);
        o *= a.alpha;
        if (a.accumulate) o += ldf(y + base + j);
        stf(y + base + j, o);
      }
    }
  }
}

template <typename T>
static void conv_dispatch(const ConvArgs& a, cudaStream_t s) {
  const long long M = (long long)a.N * a.Ho * a.Wo;
  if (a.Co <= 32) {
    dim3 grid(cdiv(M, 128), cdiv(a.Co, 32));
    xu_launch(conv_simt_kernel<T, 128, 32>, grid, 256, 0, s, a);
  } else {
    dim3 grid(cdiv(M, 64), cdiv(a.Co, 64));
    xu_launch(conv_simt_kernel<T, 64, 64>, grid, 256, 0, s, a);
  }
}

void launch_conv_simt(int dtype, const ConvArgs& a, cudaStream_t s) {
  if (dtype == XU_F32) conv_dispatch<float>(a, s);
  else conv_dispatch<bf16>(a, s);
}

// ------------------------------------------------------------------------------------------------------
// weight gradient: split-K over output pixels, fp32 atomics into the flat gradient buffer
// ------------------------------------------------------------------------------------------------------
template <typename T, int TM, int TN, int RM, int RN>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradArgs a) {
  xu_grid_dep_sync();
  constexpr int PK = 16;
  constexpr int TXN = TN / RN;
  static_assert((TM / RM) * (TN / RN) == 256, "thread tiling");
  __shared__ float As[PK][TM + 1];
  __shared__ float Bs[PK][TN + 1];
  const T* __restrict__ x = reinterpret_cast<const T*>(a.x);
  const T* __restrict__ dy = reinterpret_cast<const T*>(a.dy);
  const int tid = threadIdx.x;
  const int tiles_co = (a.Co + TN - 1) / TN;
  const int tile_ci = blockIdx.x / tiles_co, tile_co = blockIdx.x - tile_ci * tiles_co;
  const int tap = blockIdx.y;
  const int tdy = tap / a.ks, tdx = tap - tdy * a.ks;
  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long chunks = (M + PK - 1) / PK;
  const long long per = (chunks + gridDim.z - 1) / gridDim.z;
  const long long cbeg = (long long)blockIdx.z * per;
  const long long cend = (cbeg + per < chunks) ? cbeg + per : chunks;
  const int txx = tid % TXN, tyy = tid / TXN;
  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;
  const bool do_bias = (a.dbias != nullptr) && blockIdx.y == 0 && tile_ci == 0;
  float bsum = 0.f;
  const int pp = tid >> 4, lane16 = tid & 15;

  for (long long ch = cbeg; ch < cend; ++ch) {
    const long long m = ch * PK + pp;
    bool ok = false;
    long long xbase = 0;
    if (m < M) {
      int ox = (int)(m % a.Wo);
      long long t = m / a.Wo;
      int oy = (int)(t % a.Ho);
      int n = (int)(t / a.Ho);
      int iy = oy * a.stride + tdy - a.pad_h, ix = ox * a.stride + tdx - a.pad_w;
      ok = iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi;
      xbase = (((long long)n * a.Hi + iy) * a.Wi + ix) * a.Ci;
    }
#pragma unroll
    for (int r = 0; r < TM / 16; ++r) {
      const int c = lane16 + 16 * r;
      const int ci = tile_ci * TM + c;
      float v = 0.f;
      if (ok && ci < a.Ci) v = ldf(x + xbase + ci);
      As[pp][c] = v;
    }
#pragma unroll
    for (int r = 0; r < TN / 16; ++r) {
      const int c = lane16 + 16 * r;
      const int co = tile_co * TN + c;
      float v = 0.f;
      if (m < M && co < a.Co) v = ldf(dy + m * a.Co + co);
      Bs[pp][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < PK; ++p) {
      float av[RM], bv[RN];
#pragma unroll
      for (int i = 0; i < RM; ++i) av[i] = As[p][tyy * RM + i];
#pragma unroll
      for (int j = 0; j < RN; ++j) bv[j] = Bs[p][txx * RN + j];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (do_bias && tid < TN) {
#pragma unroll
      for (int p = 0; p < PK; ++p) bsum += Bs[p][tid];
    }
    __syncthreads();
  }
  const int taps = a.ks * a.ks;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int ci = tile_ci * TM + tyy * RM + i;
    if (ci >= a.Ci) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int co = tile_co * TN + txx * RN + j;
      if (co >= a.Co) continue;
      atomicAdd(a.dw + conv_waddr(tap, ci, co, taps, a.Ci, a.segw), a.alpha * acc[i][j]);
    }
  }
  if (do_bias && tid < TN) {
    const int co = tile_co * TN + tid;
    if (co < a.Co) atomicAdd(a.dbias + co, a.alpha * bsum);
  }
}

template <typename T>
static void wgrad_dispatch(const WgradArgs& a, cudaStream_t s) {
  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long chunks = (M + 15) / 16;
  const int taps = a.ks * a.ks;
  if (a.Ci >= 48 && a.Co >= 48) {
    int tiles = cdiv(a.Ci, 64) * cdiv(a.Co, 64);
    long long ks = (4 * xu_num_sms() + (long long)tiles * taps - 1) / ((long long)tiles * taps);
    if (ks < 1) ks = 1;
    if (ks > chunks) ks = chunks;
    if (ks > 65535) ks = 65535;
    dim3 grid(tiles, taps, (unsigned)ks);
    xu_launch(wgrad_simt_kernel<T, 64, 64, 4, 4>, grid, 256, 0, s, a);
  } else {
    int tiles = cdiv(a.Ci, 32) * cdiv(a.Co, 32);
    long long ks = (4 * xu_num_sms() + (long long)tiles * taps - 1) / ((long long)tiles * taps);
    if (ks < 1) ks = 1;
    if (ks > chunks) ks = chunks;
    if (ks > 65535) ks = 65535;
    dim3 grid(tiles, taps, (unsigned)ks);
    xu_launch(wgrad_simt_kernel<T, 32, 32, 2, 2>, grid, 256, 0, s, a);
  }
}

void launch_wgrad_simt(int dtype, const WgradArgs& a, cudaStream_t s) {
  if (dtype == XU_F32) wgrad_dispatch<float>(a, s);
  else wgrad_dispatch<bf16>(a, s);
}

// ======================================================================================================
// Direct kernels for the two 3-channel convolutions: the input conv 3 -> ch (model/xunet.py:229) and the output
// conv ch -> 3 (model/xunet.py:276).  K (resp. N) = 3 is far too thin for a GEMM tile, so these stay on the SIMT
// pipes: one thread per output pixel (x 8-channel group), weights broadcast from shared memory.
// ======================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) conv_cin3_fwd_kernel(ConvArgs a) {
  xu_grid_dep_sync();
  extern __shared__ float sw[];                 // [27][Co] weights + [Co] bias
  const int Co = a.Co;
  for (int i = threadIdx.x; i < 27 * Co; i += 256) sw[i] = a.w[i];
  for (int i = threadIdx.x; i < Co; i += 256) sw[27 * Co + i] = a.bias ? a.bias[i] : 0.f;
  __syncthreads();
  const int G = Co >> 3;
  const long long total = (long long)a.N * a.Ho * a.Wo * G;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % G);
  const long long pix = idx / G;
  const int ox = (int)(pix % a.Wo);
  const int oy = (int)((pix / a.Wo) % a.Ho);
  const int n = (int)(pix / ((long long)a.Wo * a.Ho));
  const T* x = reinterpret_cast<const T*>(a.x);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = sw[27 * Co + g * 8 + j];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
    if (iy < 0 || iy >= a.Hi || ix < 0 || ix >= a.Wi) continue;
    const T* px = x + (((long long)n * a.Hi + iy) * a.Wi + ix) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = ldf(px + c);
      const float* wr = sw + (tap * 3 + c) * Co + g * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
    }
  }
  T* y = reinterpret_cast<T*>(a.y) + pix * Co + g * 8;
  float o0[4] = {acc[0] * a.alpha, acc[1] * a.alpha, acc[2] * a.alpha, acc[3] * a.alpha};
  float o1[4] = {acc[4] * a.alpha, acc[5] * a.alpha, acc[6] * a.alpha, acc[7] * a.alpha};
  Vec4<T>::st(y, o0);
  Vec4<T>::st(y + 4, o1);
}

// dW[27][Co] (+ dbias as row 27) = sum over pixels of patch (x) dY ; one block per chunk of 128 pixels
template <typename T>
__global__ void __launch_bounds__(256) conv_cin3_wgrad_kernel(WgradArgs a) {
  xu_grid_dep_sync();
  constexpr int PPB = 128;
  __shared__ float xs[PPB][28];
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* dy = reinterpret_cast<const T*>(a.dy);
  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long p0 = (long long)blockIdx.x * PPB;
  for (int i = threadIdx.x; i < PPB * 28; i += 256) {
    const int pp = i / 28, k = i - pp * 28;
    const long long m = p0 + pp;
    float v = 0.f;
    if (m < M) {
      if (k == 27) v = 1.f;
      else {
        const int tap = k / 3, c = k - tap * 3;
        const int ox = (int)(m % a.Wo);
        const int oy = (int)((m / a.Wo) % a.Ho);
        const int n = (int)(m / ((long long)a.Wo * a.Ho));
        const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
        if (iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi) v = ldf(x + (((long long)n * a.Hi + iy) * a.Wi + ix) * 3 + c);
      }
    }
    xs[pp][k] = v;
  }
  __syncthreads();
  const int G = a.Co >> 2;
  const int items = 28 * G;
  const int npix = (int)((M - p0) < PPB ? (M - p0) : PPB);
  for (int it = threadIdx.x; it < items; it += 256) {
    const int k = it / G, g = it - k * G;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int pp = 0; pp < npix; ++pp) {
      float d[4];
      Vec4<T>::ld(dy + (p0 + pp) * a.Co + g * 4, d);
      const float xv = xs[pp][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(xv, d[j], acc[j]);
    }
    float* dst = (k == 27) ? (a.dbias ? a.dbias + g * 4 : nullptr) : a.dw + (long long)k * a.Co + g * 4;
    if (dst)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(dst + j, a.alpha * acc[j]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) conv_cout3_fwd_kernel(ConvArgs a) {
  xu_grid_dep_sync();
  extern __shared__ float sw[];                 // [9][Ci][3]
  const int Ci = a.Ci;
  for (int i = threadIdx.x; i < 27 * Ci; i += 256) sw[i] = a.w[i];
  __syncthreads();
  const long long total = (long long)a.N * a.Ho * a.Wo;
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= total) return;
  const int ox = (int)(pix % a.Wo);
  const int oy = (int)((pix / a.Wo) % a.Ho);
  const int n = (int)(pix / ((long long)a.Wo * a.Ho));
  const T* x = reinterpret_cast<const T*>(a.x);
  float acc[3] = {a.bias ? a.bias[0] : 0.f, a.bias ? a.bias[1] : 0.f, a.bias ? a.bias[2] : 0.f};
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
    if (iy < 0 || iy >= a.Hi || ix < 0 || ix >= a.Wi) continue;
    const T* px = x + (((long long)n * a.Hi + iy) * a.Wi + ix) * Ci;
    const float* wr = sw + tap * Ci * 3;
    for (int c = 0; c < Ci; c += 4) {
      float v[4];
      Vec4<T>::ld(px + c, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[0] = fmaf(v[j], wr[(c + j) * 3 + 0], acc[0]);
        acc[1] = fmaf(v[j], wr[(c + j) * 3 + 1], acc[1]);
        acc[2] = fmaf(v[j], wr[(c + j) * 3 + 2], acc[2]);
      }
    }
  }
  T* y = reinterpret_cast<T*>(a.y) + pix * 3;
  stf(y, acc[0] * a.alpha); stf(y + 1, acc[1] * a.alpha); stf(y + 2, acc[2] * a.alpha);
}

// dX[pix][ci] (+)= alpha * sum_{tap,co} dO[pix + 1 - tap][co] W[tap][ci][co]      (a.x = dO (N,H,W,3), a.y = dX (N,H,W,Ci=a.Co))
template <typename T>
__global__ void __launch_bounds__(256) conv_cout3_dgrad_kernel(ConvArgs a) {
  xu_grid_dep_sync();
  extern __shared__ float sw[];                 // [9][Ci][3]
  const int Ci = a.Co;
  for (int i = threadIdx.x; i < 27 * Ci; i += 256) sw[i] = a.w[i];
  __syncthreads();
  const int G = Ci >> 3;
  const long long total = (long long)a.N * a.Ho * a.Wo * G;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % G);
  const long long pix = idx / G;
  const int ix0 = (int)(pix % a.Wo);
  const int iy0 = (int)((pix / a.Wo) % a.Ho);
  const int n = (int)(pix / ((long long)a.Wo * a.Ho));
  const T* d = reinterpret_cast<const T*>(a.x);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int oy = iy0 + 1 - tap / 3, ox = ix0 + 1 - tap % 3;
    if (oy < 0 || oy >= a.Hi || ox < 0 || ox >= a.Wi) continue;
    const T* pd = d + (((long long)n * a.Hi + oy) * a.Wi + ox) * 3;
    const float d0 = ldf(pd), d1 = ldf(pd + 1), d2 = ldf(pd + 2);
    const float* wr = sw + (tap * Ci + g * 8) * 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += d0 * wr[j * 3] + d1 * wr[j * 3 + 1] + d2 * wr[j * 3 + 2];
  }
  T* y = reinterpret_cast<T*>(a.y) + pix * Ci + g * 8;
  float o0[4], o1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { o0[j] = acc[j] * a.alpha; o1[j] = acc[4 + j] * a.alpha; }
  if (a.accumulate) {
    float p0[4], p1[4];
    Vec4<T>::ld(y, p0); Vec4<T>::ld(y + 4, p1);
#pragma unroll
    for (int j = 0; j < 4; ++j) { o0[j] += p0[j]; o1[j] += p1[j]; }
  }
  Vec4<T>::st(y, o0);
  Vec4<T>::st(y + 4, o1);
}

// dW[9][Ci][3] += alpha * sum_pix x[pix + tap - 1][ci] dO[pix][co];  dbias[3] += alpha * sum dO
template <typename T>
__global__ void __launch_bounds__(320) conv_cout3_wgrad_kernel(WgradArgs a) {
  xu_grid_dep_sync();
  constexpr int PPB = 128;
  __shared__ float ds[PPB][3];
  __shared__ int sy[PPB], sx[PPB], sn[PPB];
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* dy = reinterpret_cast<const T*>(a.dy);
  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long p0 = (long long)blockIdx.x * PPB;
  if (threadIdx.x < PPB) {
    const int pp = threadIdx.x;
    const long long m = p0 + pp;
    if (m < M) {
      sx[pp] = (int)(m % a.Wo); sy[pp] = (int)((m / a.Wo) % a.Ho); sn[pp] = (int)(m / ((long long)a.Wo * a.Ho));
      ds[pp][0] = ldf(dy + m * 3); ds[pp][1] = ldf(dy + m * 3 + 1); ds[pp][2] = ldf(dy + m * 3 + 2);
    } else { sn[pp] = -1; ds[pp][0] = ds[pp][1] = ds[pp][2] = 0.f; sx[pp] = sy[pp] = 0; }
  }
  __syncthreads();
  const int items = 9 * a.Ci;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int tap = it / a.Ci, ci = it - tap * a.Ci;
    const int dyy = tap / 3 - 1, dxx = tap % 3 - 1;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int pp = 0; pp < PPB; ++pp) {
      const int n = sn[pp];
      if (n < 0) break;
      const int iy = sy[pp] + dyy, ix = sx[pp] + dxx;
      if (iy < 0 || iy >= a.Hi || ix < 0 || ix >= a.Wi) continue;
      const float xv = ldf(x + (((long long)n * a.Hi + iy) * a.Wi + ix) * a.Ci + ci);
      acc[0] = fmaf(xv, ds[pp][0], acc[0]); acc[1] = fmaf(xv, ds[pp][1], acc[1]); acc[2] = fmaf(xv, ds[pp][2], acc[2]);
    }
    float* dst = a.dw + (long long)it * 3;
    atomicAdd(dst, a.alpha * acc[0]); atomicAdd(dst + 1, a.alpha * acc[1]); atomicAdd(dst + 2, a.alpha * acc[2]);
  }
  if (a.dbias != nullptr && threadIdx.x < 3) {
    float sacc = 0.f;
    for (int pp = 0; pp < PPB; ++pp) sacc += ds[pp][threadIdx.x];
    atomicAdd(a.dbias + threadIdx.x, a.alpha * sacc);
  }
}

bool conv_cin3_supported(const ConvArgs& a) { return a.mode == 0 && a.Ci == 3 && a.ks == 3 && a.stride == 1 && a.Co % 8 == 0 && a.segw == a.Co && a.res == nullptr && !a.accumulate && 28 * a.Co * 4 <= 96 * 1024; }
bool conv_cout3_supported(int Ci, int Co, int ks, int stride) { return Co == 3 && ks == 3 && stride == 1 && Ci % 8 == 0 && 27 * Ci * 4 <= 96 * 1024; }

template <typename T>
static void small_dispatch(int which, const ConvArgs* c, const WgradArgs* w, cudaStream_t s) {
  if (which == 0) {
    const long long total = (long long)c->N * c->Ho * c->Wo * (c->Co / 8);
    const size_t sm = sizeof(float) * 28 * c->Co;
    if (sm > 48 * 1024) cudaFuncSetAttribute(conv_cin3_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    xu_launch(conv_cin3_fwd_kernel<T>, cdiv(total, 256), 256, sm, s, *c);
  } else if (which == 1) {
    const long long M = (long long)w->N * w->Ho * w->Wo;
    xu_launch(conv_cin3_wgrad_kernel<T>, cdiv(M, 128), 256, 0, s, *w);
  } else if (which == 2) {
    const long long total = (long long)c->N * c->Ho * c->Wo;
    const size_t sm = sizeof(float) * 27 * c->Ci;
    if (sm > 48 * 1024) cudaFuncSetAttribute(conv_cout3_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    xu_launch(conv_cout3_fwd_kernel<T>, cdiv(total, 256), 256, sm, s, *c);
  } else if (which == 3) {
    const long long total = (long long)c->N * c->Ho * c->Wo * (c->Co / 8);
    const size_t sm = sizeof(float) * 27 * c->Co;
    if (sm > 48 * 1024) cudaFuncSetAttribute(conv_cout3_dgrad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    xu_launch(conv_cout3_dgrad_kernel<T>, cdiv(total, 256), 256, sm, s, *c);
  } else {
    const long long M = (long long)w->N * w->Ho * w->Wo;
    const int items = 9 * w->Ci;
    const int threads = items >= 320 ? 320 : ((items + 31) / 32 * 32 < 128 ? 128 : (items + 31) / 32 * 32);
    xu_launch(conv_cout3_wgrad_kernel<T>, cdiv(M, 128), threads, 0, s, *w);
  }
}
void launch_conv_small(int dtype, int which, const ConvArgs* c, const WgradArgs* w, cudaStream_t s) {
  if (dtype == XU_F32) small_dispatch<float>(which, c, w, s);
  else small_dispatch<bf16>(which, c, w, s);
}
