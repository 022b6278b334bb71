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
    conv_simt_kernel<T, 128, 32><<<grid, 256, 0, s>>>(a);
  } else {
    dim3 grid(cdiv(M, 64), cdiv(a.Co, 64));
    conv_simt_kernel<T, 64, 64><<<grid, 256, 0, s>>>(a);
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
    long long ks = (4 * 148 + (long long)tiles * taps - 1) / ((long long)tiles * taps);
    if (ks < 1) ks = 1;
    if (ks > chunks) ks = chunks;
    if (ks > 65535) ks = 65535;
    dim3 grid(tiles, taps, (unsigned)ks);
    wgrad_simt_kernel<T, 64, 64, 4, 4><<<grid, 256, 0, s>>>(a);
  } else {
    int tiles = cdiv(a.Ci, 32) * cdiv(a.Co, 32);
    long long ks = (4 * 148 + (long long)tiles * taps - 1) / ((long long)tiles * taps);
    if (ks < 1) ks = 1;
    if (ks > chunks) ks = chunks;
    if (ks > 65535) ks = 65535;
    dim3 grid(tiles, taps, (unsigned)ks);
    wgrad_simt_kernel<T, 32, 32, 2, 2><<<grid, 256, 0, s>>>(a);
  }
}

void launch_wgrad_simt(int dtype, const WgradArgs& a, cudaStream_t s) {
  if (dtype == XU_F32) wgrad_dispatch<float>(a, s);
  else wgrad_dispatch<bf16>(a, s);
}
