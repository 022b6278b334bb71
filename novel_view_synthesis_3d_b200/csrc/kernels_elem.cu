// kernels_elem.cu -- HBM-bound kernels of the X-UNet path: joint-frame GroupNorm (+SiLU/FiLM/dropout/
// resample) forward and backward, conditioning (log-SNR embedding, camera-ray NeRF posenc), resample,
// channel copies, loss, Adam, sampler update.  All 128-bit (fp32) / 64-bit (bf16) vectorised along C.
#include "common.cuh"
#include "kernels.h"
#include <string.h>

static thread_local char g_kernel_error[256] = "";
const char* xu_kernel_error() { return g_kernel_error; }
void xu_set_kernel_error(const char* msg) {
  strncpy(g_kernel_error, msg, sizeof(g_kernel_error) - 1);
  g_kernel_error[sizeof(g_kernel_error) - 1] = 0;
}

// lanes l, l' of a warp hold partial sums of the same channel vector iff (l - l') % TPB == 0 (when TPB divides 32):
// fold them with xor-shuffles so that only the first TPB lanes touch shared memory (cuts smem-atomic contention 4-32x)
__device__ __forceinline__ bool fold_same_channel_lanes(float (&v)[4], int TPB) {
  if (TPB >= 32 || (32 % TPB) != 0) return true;
  for (int off = TPB; off < 32; off <<= 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
  }
  return (threadIdx.x & 31) < TPB;
}

// ======================================================================================================
// GroupNorm  (model/xunet.py:46-52: nn.GroupNorm(32) on (B,2,H,W,C) -> statistics over F,H,W,C/32 jointly)
// ======================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, float* __restrict__ stats, int P, int C,
                                                       int ppb) {
  xu_grid_dep_sync();
  __shared__ float sg[XU_GROUPS], sq[XU_GROUPS];
  const int tid = threadIdx.x, b = blockIdx.y;
  if (tid < XU_GROUPS) { sg[tid] = 0.f; sq[tid] = 0.f; }
  __syncthreads();
  const int C4 = C >> 2, cpg = C / XU_GROUPS;
  const int TPB = C4 < 256 ? C4 : 256;
  const int PL = 256 / TPB;
  const int cv0 = tid % TPB, pl = tid / TPB;
  const int pbeg = blockIdx.x * ppb;
  const int pend = min(pbeg + ppb, P);
  {
    for (int cv = cv0; cv < C4; cv += TPB) {
      float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
      const T* base = x + ((long long)b * P) * C + cv * 4;
      if (pl < PL)
#pragma unroll 4
        for (int p = pbeg + pl; p < pend; p += PL) {
          float v[4];
          Vec4<T>::ldg(base + (long long)p * C, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
        }
      const bool owner = fold_same_channel_lanes(s, TPB);
      fold_same_channel_lanes(ss, TPB);
      if (owner && pl < PL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int g = (cv * 4 + j) / cpg;
          atomicAdd(&sg[g], s[j]);
          atomicAdd(&sq[g], ss[j]);
        }
      }
    }
  }
  __syncthreads();
  if (tid < XU_GROUPS) {
    atomicAdd(&stats[(b * XU_GROUPS + tid) * 2 + 0], sg[tid]);
    atomicAdd(&stats[(b * XU_GROUPS + tid) * 2 + 1], sq[tid]);
  }
}

static void gn_grid(int C, int P, int B, dim3& grid, int& ppb) {
  int C4 = C / 4;
  int TPB = C4 < 256 ? C4 : 256;
  int PL = 256 / TPB;
  static const int px = getenv("XUNET_GN_RED_PX") ? atoi(getenv("XUNET_GN_RED_PX")) : 8;
  ppb = PL * px;
  // enough blocks to cover the memory latency (reductions are latency-bound), but not absurdly many atomics
  while ((long long)cdiv(P, ppb) * B > xu_num_sms() * 8 && ppb < P) ppb *= 2;
  grid = dim3(cdiv(P, ppb), B);
}

void launch_gn_stats(int dtype, const GnArgs& a, cudaStream_t s) {
  const int B = a.N / 2, P = 2 * a.H * a.W;
  if (!a.skip_zero) cudaMemsetAsync(a.stats, 0, sizeof(float) * B * XU_GROUPS * 2, s);
  dim3 grid; int ppb;
  gn_grid(a.C, P, B, grid, ppb);
  if (dtype == XU_F32) xu_launch(gn_stats_kernel<float>, grid, 256, 0, s, (const float*)a.x, a.stats, P, a.C, ppb);
  else xu_launch(gn_stats_kernel<bf16>, grid, 256, 0, s, (const bf16*)a.x, a.stats, P, a.C, ppb);
}

// Per-(sample, channel) statistics for tensors whose producer cannot emit them in its epilogue: same traversal as
// gn_stats_kernel, channel sums kept in shared memory, one interleaved [sum, sumsq] red per channel and block.
template <typename T>
__global__ void __launch_bounds__(256) gn_cstats_kernel(const T* __restrict__ x, float* __restrict__ cstats, int P, int C,
                                                        int ppb) {
  xu_grid_dep_sync();
  extern __shared__ float scs[];   // [C][2]
  const int tid = threadIdx.x, b = blockIdx.y;
  for (int i = tid; i < 2 * C; i += 256) scs[i] = 0.f;
  __syncthreads();
  const int C4 = C >> 2;
  const int TPB = C4 < 256 ? C4 : 256;
  const int PL = 256 / TPB;
  const int cv0 = tid % TPB, pl = tid / TPB;
  const int pbeg = blockIdx.x * ppb;
  const int pend = min(pbeg + ppb, P);
  for (int cv = cv0; cv < C4; cv += TPB) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    const T* base = x + ((long long)b * P) * C + cv * 4;
    if (pl < PL)
#pragma unroll 4
      for (int p = pbeg + pl; p < pend; p += PL) {
        float v[4];
        Vec4<T>::ldg(base + (long long)p * C, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
      }
    const bool owner = fold_same_channel_lanes(s, TPB);
    fold_same_channel_lanes(ss, TPB);
    if (owner && pl < PL) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&scs[(cv * 4 + j) * 2 + 0], s[j]);
        atomicAdd(&scs[(cv * 4 + j) * 2 + 1], ss[j]);
      }
    }
  }
  __syncthreads();
  float* dst = cstats + (long long)b * C * 2;
  for (int i = tid; i < 2 * C; i += 256) atomicAdd(&dst[i], scs[i]);
}

void launch_gn_cstats(int dtype, const void* x, float* cstats, int N, int H, int W, int C, cudaStream_t s) {
  const int B = N / 2, P = 2 * H * W;
  dim3 grid; int ppb;
  gn_grid(C, P, B, grid, ppb);
  const size_t smem = sizeof(float) * 2 * C;
  if (dtype == XU_F32) xu_launch(gn_cstats_kernel<float>, grid, 256, smem, s, (const float*)x, cstats, P, C, ppb);
  else xu_launch(gn_cstats_kernel<bf16>, grid, 256, smem, s, (const bf16*)x, cstats, P, C, ppb);
}

struct GnDev {
  const void* x; void* y; const void* e; const void* dy; void* de;
  const float* gamma; const float* beta; float* dgamma; float* dbeta; float* stats; float* bstats;
  int N, H, W, C, Ho, Wo, mode, rs, train, op_index, accumulate, de_accumulate;
  float drop_rate, inv_cnt;
  int cpg, cpg_shift;   // channels per group; log2 if a power of two, else -1
  const void* extra; float extra_alpha;
  const unsigned long long* seed_dev;
  const float* cstatsA; const float* cstatsB; int csA;
  float4* params_out; const float* bcs;
};

static GnDev gn_dev(const GnArgs& a) {
  GnDev d;
  d.x = a.x; d.y = a.y; d.e = a.e; d.dy = a.dy; d.de = a.de; d.gamma = a.gamma; d.beta = a.beta;
  d.dgamma = a.dgamma; d.dbeta = a.dbeta; d.stats = a.stats; d.bstats = a.bstats;
  d.N = a.N; d.H = a.H; d.W = a.W; d.C = a.C;
  d.Ho = a.rs == RS_DOWN ? a.H / 2 : (a.rs == RS_UP ? a.H * 2 : a.H);
  d.Wo = a.rs == RS_DOWN ? a.W / 2 : (a.rs == RS_UP ? a.W * 2 : a.W);
  d.mode = a.mode; d.rs = a.rs; d.train = a.train; d.op_index = a.op_index; d.accumulate = a.accumulate;
  d.de_accumulate = a.de_accumulate;
  d.drop_rate = a.drop_rate;
  d.inv_cnt = 1.f / (2.f * a.H * a.W * (a.C / XU_GROUPS));
  d.cpg = a.C / XU_GROUPS;
  d.cpg_shift = -1;
  for (int sft = 0; sft < 12; ++sft) if ((1 << sft) == d.cpg) d.cpg_shift = sft;
  d.seed_dev = a.seed_dev;
  d.extra = a.extra; d.extra_alpha = a.extra_alpha;
  d.cstatsA = a.cstatsA; d.cstatsB = a.cstatsB; d.csA = a.csA;
  d.params_out = reinterpret_cast<float4*>(a.params_out); d.bcs = a.bcs;
  return d;
}

__device__ __forceinline__ int gn_group(const GnDev& d, int c) { return d.cpg_shift >= 0 ? (c >> d.cpg_shift) : c / d.cpg; }
__device__ __forceinline__ void gn_mean_rstd(const GnDev& d, int b, int c, float& mean, float& rstd) {
  const int g = gn_group(d, c);
  const float s = d.stats[(b * XU_GROUPS + g) * 2 + 0], q = d.stats[(b * XU_GROUPS + g) * 2 + 1];
  mean = s * d.inv_cnt;
  float var = fmaxf(q * d.inv_cnt - mean * mean, 0.f);
  rstd = rsqrtf(var + XU_GN_EPS);
}

// normalised+affine value of 4 channels of input pixel (n,y,x); optionally returns xhat
template <typename T>
__device__ __forceinline__ void gn_yhat4(const GnDev& d, int n, int y, int x, int c0, const float (&mean)[4],
                                         const float (&rstd)[4], const float (&gm)[4], const float (&bt)[4],
                                         float (&yh)[4], float (&xh)[4]) {
  float v[4];
  Vec4<T>::ldg(reinterpret_cast<const T*>(d.x) + (((long long)n * d.H + y) * d.W + x) * d.C + c0, v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    xh[j] = (v[j] - mean[j]) * rstd[j];
    yh[j] = fmaf(xh[j], gm[j], bt[j]);
  }
}

// grid (pixel chunks of the OUTPUT, B): every thread owns a fixed 4-channel vector, so the per-channel constants
// (mean, rstd, gamma, beta) are formed once and the inner loop is load -> fma -> activation -> store.
template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(GnDev d, int ppb) {
  xu_grid_dep_sync();
  const int tid = threadIdx.x, b = blockIdx.y;
  __shared__ float s_grp[XU_GROUPS][2];
  if (d.cstatsA != nullptr) {
    // the producers of x emitted per-channel sums: fold them into the 32 group sums here (8 lanes per group) -- there is no
    // statistics pass over x.  Block 0 of each sample also stores the group sums where the backward kernels read them.
    const int g = tid >> 3, sub = tid & 7;
    float s = 0.f, q = 0.f;
    for (int k = sub; k < d.cpg; k += 8) {
      const int c = g * d.cpg + k;
      const float* src = c < d.csA ? d.cstatsA + ((long long)b * d.csA + c) * 2
                                   : d.cstatsB + ((long long)b * (d.C - d.csA) + (c - d.csA)) * 2;
      const float2 v = *reinterpret_cast<const float2*>(src);
      s += v.x; q += v.y;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (sub == 0) {
      s_grp[g][0] = s; s_grp[g][1] = q;
      if (blockIdx.x == 0) { d.stats[(b * XU_GROUPS + g) * 2 + 0] = s; d.stats[(b * XU_GROUPS + g) * 2 + 1] = q; }
    }
    __syncthreads();
  }
  if (d.params_out != nullptr && blockIdx.x == 0) {
    // per-(sample, channel) {rstd, -mean*rstd, gamma, beta}: what the fused GroupNorm-backward epilogue of the consumer conv's
    // data gradient needs to rebuild xhat and the pre-activation from x (one 16-byte load per channel there)
    for (int c = tid; c < d.C; c += 256) {
      float mean, rstd;
      if (d.cstatsA != nullptr) {
        const int g = gn_group(d, c);
        mean = s_grp[g][0] * d.inv_cnt;
        rstd = rsqrtf(fmaxf(s_grp[g][1] * d.inv_cnt - mean * mean, 0.f) + XU_GN_EPS);
      } else gn_mean_rstd(d, b, c, mean, rstd);
      d.params_out[(long long)b * d.C + c] = make_float4(rstd, -mean * rstd, d.gamma[c], d.beta[c]);
    }
  }
  const int C4 = d.C >> 2;
  const int TPB = C4 < 256 ? C4 : 256;
  const int PL = 256 / TPB;
  const int cv0 = tid % TPB, pl = tid / TPB;
  if (pl >= PL) return;
  const int HWo = d.Ho * d.Wo, P = 2 * HWo;
  const int pbeg = blockIdx.x * ppb;
  const int pend = min(pbeg + ppb, P);
  const bool drop = d.mode == GN_FILM && d.train && d.drop_rate > 0.f;
  const unsigned long long seed = drop ? *d.seed_dev : 0ULL;
  const float keep_scale = 1.f / (1.f - d.drop_rate);
  for (int cv = cv0; cv < C4; cv += TPB) {
    const int c0 = cv * 4;
    float mean[4], rstd[4], gm[4], bt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (d.cstatsA != nullptr) {
        const int g = gn_group(d, c0 + j);
        mean[j] = s_grp[g][0] * d.inv_cnt;
        rstd[j] = rsqrtf(fmaxf(s_grp[g][1] * d.inv_cnt - mean[j] * mean[j], 0.f) + XU_GN_EPS);
      } else gn_mean_rstd(d, b, c0 + j, mean[j], rstd[j]);
      gm[j] = d.gamma[c0 + j];
      bt[j] = d.beta[c0 + j];
    }
#pragma unroll 2
    for (int p = pbeg + pl; p < pend; p += PL) {
      // without resampling the pixel index alone addresses everything ((2b)*HW + p): no div/mod in the hot loop
      int n = 2 * b, oy = 0, ox = p;
      if (d.rs != RS_NONE) {
        const int f = p / HWo, r = p - f * HWo;
        oy = r / d.Wo; ox = r - oy * d.Wo; n = b * 2 + f;
      }
      const long long pix = (long long)(2 * b) * HWo + p;
      float out[4], yh[4], xh[4];
      if (d.mode == GN_FILM) {
        gn_yhat4<T>(d, n, oy, ox, c0, mean, rstd, gm, bt, yh, xh);
        const T* e = reinterpret_cast<const T*>(d.e) + pix * (2LL * d.C);
        float sc[4], sh[4];
        Vec4<T>::ldg(e + c0, sc);
        Vec4<T>::ldg(e + d.C + c0, sh);
        const uint32_t km = drop ? xu_keep4(seed, d.op_index, (unsigned long long)(pix * d.C + c0) >> 2, d.drop_rate) : 0xFu;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float u = fmaf(yh[j], 1.f + sc[j], sh[j]);
          float sv = swishf_(u);
          if (drop) sv = ((km >> j) & 1u) ? sv * keep_scale : 0.f;
          out[j] = sv;
        }
      } else if (d.rs == RS_DOWN) {
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = 0.f;
        for (int i = 0; i < 2; ++i)
          for (int k = 0; k < 2; ++k) {
            gn_yhat4<T>(d, n, oy * 2 + i, ox * 2 + k, c0, mean, rstd, gm, bt, yh, xh);
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] += (d.mode == GN_SWISH) ? swishf_(yh[j]) : yh[j];
          }
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] *= 0.25f;
      } else {
        const int iy = d.rs == RS_UP ? oy >> 1 : oy, ix = d.rs == RS_UP ? ox >> 1 : ox;
        gn_yhat4<T>(d, n, iy, ix, c0, mean, rstd, gm, bt, yh, xh);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = (d.mode == GN_SWISH) ? swishf_(yh[j]) : yh[j];
      }
      Vec4<T>::st(reinterpret_cast<T*>(d.y) + pix * d.C + c0, out);
    }
  }
}

// pixels per block for the elementwise GroupNorm passes: ~16 pixels per thread, >= 2 waves of blocks when possible
static void gn_apply_grid(int C, int P, int B, dim3& grid, int& ppb) {
  int C4 = C / 4;
  int TPB = C4 < 256 ? C4 : 256;
  int PL = 256 / TPB;
  static const int px = getenv("XUNET_GN_APPLY_PX") ? atoi(getenv("XUNET_GN_APPLY_PX")) : 4;   // measured: 4 beats 8 and 2 (latency-bound: more blocks in flight)
  ppb = PL * 16;   // the per-thread prologue (statistics -> scale/shift) is ~200 instructions: amortise it over >= 8 pixels
  while (ppb > PL * px && (long long)cdiv(P, ppb) * B < xu_num_sms() * 2 * (8 / px)) ppb /= 2;
  grid = dim3(cdiv(P, ppb), B);
}


// ---- no-resample fast paths: 16-byte accesses for both dtypes (8 bf16 / 4 fp32 channels per thread), U pixels of loads in
// flight per thread before the first use, per-channel affine folded to one fma: yhat = x * a + b -----------------------------
template <typename T>
__device__ __forceinline__ void gn_group_prologue(const GnDev& d, int b, float (*s_grp)[2]) {
  const int tid = threadIdx.x;
  if (d.cstatsA != nullptr) {
    const int g = tid >> 3, sub = tid & 7;
    float s = 0.f, q = 0.f;
    for (int k = sub; k < d.cpg; k += 8) {
      const int c = g * d.cpg + k;
      const float* src = c < d.csA ? d.cstatsA + ((long long)b * d.csA + c) * 2
                                   : d.cstatsB + ((long long)b * (d.C - d.csA) + (c - d.csA)) * 2;
      const float2 v = *reinterpret_cast<const float2*>(src);
      s += v.x; q += v.y;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (sub == 0) {
      s_grp[g][0] = s; s_grp[g][1] = q;
      if (blockIdx.x == 0) { d.stats[(b * XU_GROUPS + g) * 2 + 0] = s; d.stats[(b * XU_GROUPS + g) * 2 + 1] = q; }
    }
  } else if (tid < XU_GROUPS) {
    s_grp[tid][0] = d.stats[(b * XU_GROUPS + tid) * 2 + 0];
    s_grp[tid][1] = d.stats[(b * XU_GROUPS + tid) * 2 + 1];
  }
  __syncthreads();
}
__device__ __forceinline__ void gn_group_mean_rstd(const GnDev& d, const float (*s_grp)[2], int c, float& mean, float& rstd) {
  const int g = gn_group(d, c);
  mean = s_grp[g][0] * d.inv_cnt;
  rstd = rsqrtf(fmaxf(s_grp[g][1] * d.inv_cnt - mean * mean, 0.f) + XU_GN_EPS);
}

template <typename T, bool FILM>
__global__ void __launch_bounds__(256, FILM ? 2 : 3) gn_apply_vec_kernel(GnDev d, int ppb) {
  xu_grid_dep_sync();
  constexpr int VW = VecW<T>::W;
  constexpr int U = FILM ? 4 : 8;
  const int tid = threadIdx.x, b = blockIdx.y;
  __shared__ float s_grp[XU_GROUPS][2];
  gn_group_prologue<T>(d, b, s_grp);
  if (d.params_out != nullptr && blockIdx.x == 0) {
    for (int c = tid; c < d.C; c += 256) {
      float mean, rstd;
      gn_group_mean_rstd(d, s_grp, c, mean, rstd);
      d.params_out[(long long)b * d.C + c] = make_float4(rstd, -mean * rstd, d.gamma[c], d.beta[c]);
    }
  }
  const int CV = d.C / VW;
  const int TPB = CV < 256 ? CV : 256;
  const int PL = 256 / TPB;
  const int cv0 = tid % TPB, pl = tid / TPB;
  if (pl >= PL) return;
  const int HW = d.H * d.W, P = 2 * HW;
  const int pbeg = blockIdx.x * ppb;
  const int pend = min(pbeg + ppb, P);
  constexpr bool film = FILM;
  const bool swish = d.mode != GN_PLAIN;
  const bool drop = film && d.train && d.drop_rate > 0.f;
  const unsigned long long seed = drop ? *d.seed_dev : 0ULL;
  const float keep_scale = 1.f / (1.f - d.drop_rate);
  const long long pix0 = (long long)(2 * b) * HW;
  for (int cv = cv0; cv < CV; cv += TPB) {
    const int c0 = cv * VW;
    float ka[VW], kb[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) {
      float mean, rstd;
      gn_group_mean_rstd(d, s_grp, c0 + j, mean, rstd);
      ka[j] = rstd * d.gamma[c0 + j];
      kb[j] = fmaf(-mean, ka[j], d.beta[c0 + j]);
    }
    const T* X = reinterpret_cast<const T*>(d.x) + pix0 * d.C + c0;
    const T* E = film ? reinterpret_cast<const T*>(d.e) + pix0 * (2LL * d.C) + c0 : nullptr;
    T* Y = reinterpret_cast<T*>(d.y) + pix0 * d.C + c0;
    for (int p0 = pbeg + pl; p0 < pend; p0 += U * PL) {
      typename VecW<T>::raw xr[U], scr[FILM ? U : 1], shr[FILM ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + u * PL;
        if (p < pend) {
          xr[u] = VecW<T>::ldg(X + (long long)p * d.C);
          if constexpr (film) {
            scr[u] = VecW<T>::ldg(E + (long long)p * (2 * d.C));
            shr[u] = VecW<T>::ldg(E + (long long)p * (2 * d.C) + d.C);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + u * PL;
        if (p >= pend) break;
        float v[VW], out[VW];
        VecW<T>::unpack(xr[u], v);
        if constexpr (film) {
          float sc[VW], sh[VW];
          VecW<T>::unpack(scr[u], sc);
          VecW<T>::unpack(shr[u], sh);
          uint32_t km = 0xFFu;
          if (drop) {
            const unsigned long long e4 = (unsigned long long)((pix0 + p) * d.C + c0) >> 2;
            km = xu_keep4(seed, d.op_index, e4, d.drop_rate);
            if (VW == 8) km |= xu_keep4(seed, d.op_index, e4 + 1, d.drop_rate) << 4;
          }
#pragma unroll
          for (int j = 0; j < VW; ++j) {
            const float yh = fmaf(v[j], ka[j], kb[j]);
            float sv = swishf_(fmaf(yh, 1.f + sc[j], sh[j]));
            if (drop) sv = ((km >> j) & 1u) ? sv * keep_scale : 0.f;
            out[j] = sv;
          }
        } else {
#pragma unroll
          for (int j = 0; j < VW; ++j) {
            const float yh = fmaf(v[j], ka[j], kb[j]);
            out[j] = swish ? swishf_(yh) : yh;
          }
        }
        VecW<T>::st(Y + (long long)p * d.C, out);
      }
    }
  }
}

// pixels per block for the vector kernels
static void gn_vec_grid(int C, int VW, int P, int B, dim3& grid, int& ppb) {
  const int CV = C / VW;
  const int TPB = CV < 256 ? CV : 256;
  const int PL = 256 / TPB;
  ppb = PL * 32;                    // several batches of U pixels per thread when the tensor is large
  while (ppb > PL * 4 && (long long)cdiv(P, ppb) * B < xu_num_sms() * 4) ppb /= 2;
  grid = dim3(cdiv(P, ppb), B);
}

void launch_gn_apply(int dtype, const GnArgs& a, cudaStream_t s) {
  GnDev d = gn_dev(a);
  dim3 grid; int ppb;
  static const bool old_path = getenv("XUNET_GN_APPLY_OLD") != nullptr;     // A/B switch: the 8-byte-access kernel
  if (a.rs == RS_NONE && !old_path && d.C % 8 == 0) {
    const bool film = a.mode == GN_FILM;
    if (dtype == XU_F32) {
      gn_vec_grid(d.C, 4, 2 * d.H * d.W, d.N / 2, grid, ppb);
      if (film) xu_launch(gn_apply_vec_kernel<float, true>, grid, 256, 0, s, d, ppb);
      else xu_launch(gn_apply_vec_kernel<float, false>, grid, 256, 0, s, d, ppb);
    } else {
      gn_vec_grid(d.C, 8, 2 * d.H * d.W, d.N / 2, grid, ppb);
      if (film) xu_launch(gn_apply_vec_kernel<bf16, true>, grid, 256, 0, s, d, ppb);
      else xu_launch(gn_apply_vec_kernel<bf16, false>, grid, 256, 0, s, d, ppb);
    }
    return;
  }
  gn_apply_grid(d.C, 2 * d.Ho * d.Wo, d.N / 2, grid, ppb);
  if (dtype == XU_F32) xu_launch(gn_apply_kernel<float>, grid, 256, 0, s, d, ppb);
  else xu_launch(gn_apply_kernel<bf16>, grid, 256, 0, s, d, ppb);
}

// gradient w.r.t. yhat (the GroupNorm output before activation/FiLM) for 4 channels of INPUT pixel (n,y,x).
// For GN_FILM also returns du (grad wrt pre-swish u) so the caller can emit de.
template <typename T>
__device__ __forceinline__ void gn_dyhat4(const GnDev& d, int n, int y, int x, int c0, const float (&yh)[4],
                                          unsigned long long seed, float (&dyh)[4], float (&du)[4]) {
  const T* dout = reinterpret_cast<const T*>(d.dy);
  float g[4];
  if (d.rs == RS_NONE) {
    Vec4<T>::ldg(dout + (((long long)n * d.H + y) * d.W + x) * d.C + c0, g);
  } else if (d.rs == RS_DOWN) {
    // forward averaged 2x2 -> each input pixel receives 0.25 * dout[y/2, x/2] (odd trailing row/col gets none)
    const int oy = y >> 1, ox = x >> 1;
    if (oy < d.Ho && ox < d.Wo) {
      Vec4<T>::ldg(dout + (((long long)n * d.Ho + oy) * d.Wo + ox) * d.C + c0, g);
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] *= 0.25f;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = 0.f;
    for (int i = 0; i < 2; ++i)
      for (int k = 0; k < 2; ++k) {
        float t[4];
        Vec4<T>::ldg(dout + (((long long)n * d.Ho + (2 * y + i)) * d.Wo + (2 * x + k)) * d.C + c0, t);
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] += t[j];
      }
  }
  if (d.mode == GN_PLAIN) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { dyh[j] = g[j]; du[j] = 0.f; }
  } else if (d.mode == GN_SWISH) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { dyh[j] = g[j] * swish_gradf_(yh[j]); du[j] = 0.f; }
  } else {
    const long long pix = ((long long)n * d.H + y) * d.W + x;
    const T* e = reinterpret_cast<const T*>(d.e) + pix * (2LL * d.C);
    float sc[4], sh[4];
    Vec4<T>::ldg(e + c0, sc);
    Vec4<T>::ldg(e + d.C + c0, sh);
    const bool drop = d.train && d.drop_rate > 0.f;
    const float keep_scale = 1.f / (1.f - d.drop_rate);
    const uint32_t km = drop ? xu_keep4(seed, d.op_index, (unsigned long long)(pix * d.C + c0) >> 2, d.drop_rate) : 0xFu;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float u = fmaf(yh[j], 1.f + sc[j], sh[j]);
      float gs = g[j];
      if (drop) gs = ((km >> j) & 1u) ? gs * keep_scale : 0.f;
      du[j] = gs * swish_gradf_(u);
      dyh[j] = du[j] * (1.f + sc[j]);
    }
  }
}

// compute-only part of gn_dyhat4 for the no-resample fast path (operands already in registers)
__device__ __forceinline__ void gn_dyhat_compute(int mode, const float (&g)[4], const float (&yh)[4], const float (&sc)[4],
                                                 const float (&sh)[4], uint32_t km, bool drop, float keep_scale,
                                                 float (&dyh)[4], float (&du)[4]) {
  if (mode == GN_PLAIN) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { dyh[j] = g[j]; du[j] = 0.f; }
  } else if (mode == GN_SWISH) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { dyh[j] = g[j] * swish_gradf_(yh[j]); du[j] = 0.f; }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float u = fmaf(yh[j], 1.f + sc[j], sh[j]);
      float gs = g[j];
      if (drop) gs = ((km >> j) & 1u) ? gs * keep_scale : 0.f;
      du[j] = gs * swish_gradf_(u);
      dyh[j] = du[j] * (1.f + sc[j]);
    }
  }
}

// pass A: per-channel sums  A_c = sum dyh*xhat (-> dgamma),  B_c = sum dyh (-> dbeta); group sums S1,S2; FiLM de.
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(GnDev d, int ppb) {
  xu_grid_dep_sync();
  extern __shared__ float sm[];  // sA[C], sB[C]
  float* sA = sm;
  float* sB = sm + d.C;
  const int tid = threadIdx.x, b = blockIdx.y;
  for (int i = tid; i < 2 * d.C; i += 256) sm[i] = 0.f;
  __syncthreads();
  const int C4 = d.C >> 2;
  const int TPB = C4 < 256 ? C4 : 256;
  const int PL = 256 / TPB;
  const int cv0 = tid % TPB, pl = tid / TPB;
  const int P = 2 * d.H * d.W, HW = d.H * d.W;
  const int pbeg = blockIdx.x * ppb;
  const int pend = min(pbeg + ppb, P);
  const unsigned long long seed = (d.mode == GN_FILM && d.train && d.drop_rate > 0.f) ? *d.seed_dev : 0ULL;
  {
    for (int cv = cv0; cv < C4; cv += TPB) {
      const int c0 = cv * 4;
      float mean[4], rstd[4], gm[4], bt[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        gn_mean_rstd(d, b, c0 + j, mean[j], rstd[j]);
        gm[j] = d.gamma[c0 + j];
        bt[j] = d.beta[c0 + j];
      }
      float A[4] = {0.f, 0.f, 0.f, 0.f}, Bc[4] = {0.f, 0.f, 0.f, 0.f};
      if (pl < PL && d.rs == RS_NONE) {
        // fast path (no resampling): 4 pixels per trip, all loads issued before the first use
        constexpr int U = 4;
        const bool film = d.mode == GN_FILM;
        const bool drop = film && d.train && d.drop_rate > 0.f;
        const float keep_scale = 1.f / (1.f - d.drop_rate);
        const long long pix0 = (long long)(2 * b) * HW;
        const T* X = reinterpret_cast<const T*>(d.x) + pix0 * d.C + c0;
        const T* G = reinterpret_cast<const T*>(d.dy) + pix0 * d.C + c0;
        const T* E = film ? reinterpret_cast<const T*>(d.e) + pix0 * (2LL * d.C) + c0 : nullptr;
        T* DE = film ? reinterpret_cast<T*>(d.de) + pix0 * (2LL * d.C) + c0 : nullptr;
        for (int p0 = pbeg + pl; p0 < pend; p0 += U * PL) {
          typename Vec4<T>::raw xr[U], gr[U], scr[U], shr[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = p0 + u * PL;
            if (p < pend) {
              xr[u] = Vec4<T>::ldg_raw(X + (long long)p * d.C);
              gr[u] = Vec4<T>::ldg_raw(G + (long long)p * d.C);
              if (film) {
                scr[u] = Vec4<T>::ldg_raw(E + (long long)p * (2 * d.C));
                shr[u] = Vec4<T>::ldg_raw(E + (long long)p * (2 * d.C) + d.C);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = p0 + u * PL;
            if (p >= pend) break;
            float v[4], g[4], sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f}, xh[4], yh[4], dyh[4], du[4];
            Vec4<T>::unpack(xr[u], v);
            Vec4<T>::unpack(gr[u], g);
            if (film) { Vec4<T>::unpack(scr[u], sc); Vec4<T>::unpack(shr[u], sh); }
#pragma unroll
            for (int j = 0; j < 4; ++j) { xh[j] = (v[j] - mean[j]) * rstd[j]; yh[j] = fmaf(xh[j], gm[j], bt[j]); }
            const uint32_t km = drop ? xu_keep4(seed, d.op_index, (unsigned long long)((pix0 + p) * d.C + c0) >> 2, d.drop_rate) : 0xFu;
            gn_dyhat_compute(d.mode, g, yh, sc, sh, km, drop, keep_scale, dyh, du);
#pragma unroll
            for (int j = 0; j < 4; ++j) { A[j] = fmaf(dyh[j], xh[j], A[j]); Bc[j] += dyh[j]; }
            if (film) {
              T* de = DE + (long long)p * (2 * d.C);
              float dsc[4], dsh[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) { dsc[j] = du[j] * yh[j]; dsh[j] = du[j]; }
              if (d.de_accumulate) {
                float o1[4], o2[4];
                Vec4<T>::ld(de, o1);
                Vec4<T>::ld(de + d.C, o2);
#pragma unroll
                for (int j = 0; j < 4; ++j) { dsc[j] += o1[j]; dsh[j] += o2[j]; }
              }
              Vec4<T>::st(de, dsc);
              Vec4<T>::st(de + d.C, dsh);
            }
          }
        }
      } else if (pl < PL)
#pragma unroll 2
      for (int p = pbeg + pl; p < pend; p += PL) {
        int n = 2 * b, y = 0, x = p;             // linear addressing unless the op resamples
        if (d.rs != RS_NONE) {
          const int f = p / HW, r = p - f * HW;
          y = r / d.W; x = r - y * d.W; n = b * 2 + f;
        }
        float yh[4], xh[4], dyh[4], du[4];
        gn_yhat4<T>(d, n, y, x, c0, mean, rstd, gm, bt, yh, xh);
        gn_dyhat4<T>(d, n, y, x, c0, yh, seed, dyh, du);
#pragma unroll
        for (int j = 0; j < 4; ++j) { A[j] = fmaf(dyh[j], xh[j], A[j]); Bc[j] += dyh[j]; }
        if (d.mode == GN_FILM) {
          T* de = reinterpret_cast<T*>(d.de) + (((long long)n * d.H + y) * d.W + x) * (2LL * d.C);
          float dsc[4], dsh[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) { dsc[j] = du[j] * yh[j]; dsh[j] = du[j]; }
          if (d.de_accumulate) {
            float o1[4], o2[4];
            Vec4<T>::ld(de + c0, o1);
            Vec4<T>::ld(de + d.C + c0, o2);
#pragma unroll
            for (int j = 0; j < 4; ++j) { dsc[j] += o1[j]; dsh[j] += o2[j]; }
          }
          Vec4<T>::st(de + c0, dsc);
          Vec4<T>::st(de + d.C + c0, dsh);
        }
      }
      const bool owner = fold_same_channel_lanes(A, TPB);
      fold_same_channel_lanes(Bc, TPB);
      if (owner && pl < PL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          atomicAdd(&sA[c0 + j], A[j]);
          atomicAdd(&sB[c0 + j], Bc[j]);
        }
      }
    }
  }
  __syncthreads();
  const int cpg = d.C / XU_GROUPS;
  // L2 atomics on a handful of lines shared by every block are the serial tail of this kernel: 128-bit vector reds
  // (4 channels per lane-op) when the leaves are 16-byte aligned
  if (((reinterpret_cast<uintptr_t>(d.dgamma) | reinterpret_cast<uintptr_t>(d.dbeta)) & 15) == 0) {
    for (int c = tid * 4; c < d.C; c += 1024) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d.dgamma + c), "f"(sA[c]), "f"(sA[c + 1]), "f"(sA[c + 2]), "f"(sA[c + 3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d.dbeta + c), "f"(sB[c]), "f"(sB[c + 1]), "f"(sB[c + 2]), "f"(sB[c + 3]) : "memory");
    }
  } else {
    for (int c = tid; c < d.C; c += 256) {
      atomicAdd(&d.dgamma[c], sA[c]);
      atomicAdd(&d.dbeta[c], sB[c]);
    }
  }
  if (tid < XU_GROUPS) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < cpg; ++k) {
      const int c = tid * cpg + k;
      const float gmc = d.gamma[c];
      s1 = fmaf(gmc, sB[c], s1);
      s2 = fmaf(gmc, sA[c], s2);
    }
    atomicAdd(&d.bstats[(b * XU_GROUPS + tid) * 2 + 0], s1);
    atomicAdd(&d.bstats[(b * XU_GROUPS + tid) * 2 + 1], s2);
  }
}

void launch_gn_bwd_reduce(int dtype, const GnArgs& a, cudaStream_t s) {
  GnDev d = gn_dev(a);
  const int B = a.N / 2, P = 2 * a.H * a.W;
  if (!a.skip_zero) cudaMemsetAsync(a.bstats, 0, sizeof(float) * B * XU_GROUPS * 2, s);
  dim3 grid; int ppb;
  gn_grid(a.C, P, B, grid, ppb);
  size_t smem = sizeof(float) * 2 * a.C;
  if (dtype == XU_F32) xu_launch(gn_bwd_reduce_kernel<float>, grid, 256, smem, s, d, ppb);
  else xu_launch(gn_bwd_reduce_kernel<bf16>, grid, 256, smem, s, d, ppb);
}

// pass B: dx = rstd * (gamma*dyh - S1/cnt - xhat*S2/cnt); same thread <-> channel-vector mapping as gn_apply_kernel
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(GnDev d, int ppb) {
  xu_grid_dep_sync();
  const int tid = threadIdx.x, b = blockIdx.y;
  const int C4 = d.C >> 2;
  const int TPB = C4 < 256 ? C4 : 256;
  const int PL = 256 / TPB;
  const int cv0 = tid % TPB, pl = tid / TPB;
  if (pl >= PL) return;
  const int HW = d.H * d.W, P = 2 * HW;
  const int pbeg = blockIdx.x * ppb;
  const int pend = min(pbeg + ppb, P);
    const unsigned long long seed = (d.mode == GN_FILM && d.train && d.drop_rate > 0.f) ? *d.seed_dev : 0ULL;
  for (int cv = cv0; cv < C4; cv += TPB) {
    const int c0 = cv * 4;
    float mean[4], rstd[4], gm[4], bt[4], s1[4], s2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gn_mean_rstd(d, b, c0 + j, mean[j], rstd[j]);
      gm[j] = d.gamma[c0 + j];
      bt[j] = d.beta[c0 + j];
      const int g = gn_group(d, c0 + j);
      s1[j] = d.bstats[(b * XU_GROUPS + g) * 2 + 0] * d.inv_cnt;
      s2[j] = d.bstats[(b * XU_GROUPS + g) * 2 + 1] * d.inv_cnt;
    }
    if (d.rs == RS_NONE) {
      constexpr int U = 4;
      const bool film = d.mode == GN_FILM;
      const bool drop = film && d.train && d.drop_rate > 0.f;
      const float keep_scale = 1.f / (1.f - d.drop_rate);
      const long long pix0 = (long long)(2 * b) * HW;
      const T* X = reinterpret_cast<const T*>(d.x) + pix0 * d.C + c0;
      const T* G = reinterpret_cast<const T*>(d.dy) + pix0 * d.C + c0;
      const T* E = film ? reinterpret_cast<const T*>(d.e) + pix0 * (2LL * d.C) + c0 : nullptr;
      T* DX = reinterpret_cast<T*>(d.y) + pix0 * d.C + c0;
      for (int p0 = pbeg + pl; p0 < pend; p0 += U * PL) {
        typename Vec4<T>::raw xr[U], gr[U], scr[U], shr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int p = p0 + u * PL;
          if (p < pend) {
            xr[u] = Vec4<T>::ldg_raw(X + (long long)p * d.C);
            gr[u] = Vec4<T>::ldg_raw(G + (long long)p * d.C);
            if (film) {
              scr[u] = Vec4<T>::ldg_raw(E + (long long)p * (2 * d.C));
              shr[u] = Vec4<T>::ldg_raw(E + (long long)p * (2 * d.C) + d.C);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int p = p0 + u * PL;
          if (p >= pend) break;
          float v[4], g[4], sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f}, xh[4], yh[4], dyh[4], du[4], out[4];
          Vec4<T>::unpack(xr[u], v);
          Vec4<T>::unpack(gr[u], g);
          if (film) { Vec4<T>::unpack(scr[u], sc); Vec4<T>::unpack(shr[u], sh); }
#pragma unroll
          for (int j = 0; j < 4; ++j) { xh[j] = (v[j] - mean[j]) * rstd[j]; yh[j] = fmaf(xh[j], gm[j], bt[j]); }
          const uint32_t km = drop ? xu_keep4(seed, d.op_index, (unsigned long long)((pix0 + p) * d.C + c0) >> 2, d.drop_rate) : 0xFu;
          gn_dyhat_compute(d.mode, g, yh, sc, sh, km, drop, keep_scale, dyh, du);
#pragma unroll
          for (int j = 0; j < 4; ++j) out[j] = rstd[j] * (gm[j] * dyh[j] - s1[j] - xh[j] * s2[j]);
          if (d.extra != nullptr) {   // fused residual-branch gradient (replaces a separate axpy pass over dx)
            float ex[4];
            Vec4<T>::ldg(reinterpret_cast<const T*>(d.extra) + pix0 * d.C + c0 + (long long)p * d.C, ex);
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] = fmaf(d.extra_alpha, ex[j], out[j]);
          }
          T* dx = DX + (long long)p * d.C;
          if (d.accumulate) {
            float o[4];
            Vec4<T>::ld(dx, o);
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] += o[j];
          }
          Vec4<T>::st(dx, out);
        }
      }
    } else
#pragma unroll 2
    for (int p = pbeg + pl; p < pend; p += PL) {
      int n = 2 * b, y = 0, x = p;
      if (d.rs != RS_NONE) {
        const int f = p / HW, r = p - f * HW;
        y = r / d.W; x = r - y * d.W; n = b * 2 + f;
      }
      float yh[4], xh[4], dyh[4], du[4], out[4];
      gn_yhat4<T>(d, n, y, x, c0, mean, rstd, gm, bt, yh, xh);
      gn_dyhat4<T>(d, n, y, x, c0, yh, seed, dyh, du);
#pragma unroll
      for (int j = 0; j < 4; ++j) out[j] = rstd[j] * (gm[j] * dyh[j] - s1[j] - xh[j] * s2[j]);
      if (d.extra != nullptr) {
        float ex[4];
        Vec4<T>::ldg(reinterpret_cast<const T*>(d.extra) + ((long long)(2 * b) * HW + p) * d.C + c0, ex);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = fmaf(d.extra_alpha, ex[j], out[j]);
      }
      T* dx = reinterpret_cast<T*>(d.y) + ((long long)(2 * b) * HW + p) * d.C + c0;
      if (d.accumulate) {
        float o[4];
        Vec4<T>::ld(dx, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] += o[j];
      }
      Vec4<T>::st(dx, out);
    }
  }
}

// Only pass of the GroupNorm backward when the consumer conv's data-gradient epilogue has already stored dyh (in a.dy) and
// accumulated the per-(sample, channel) sums [A = sum dyh*xhat, B = sum dyh] (a.bcs): fold them (dgamma, dbeta, group sums
// S1 = sum_c gamma B, S2 = sum_c gamma A), then dx = rstd * (gamma*dyh - S1/cnt - xhat*S2/cnt) [+ fused residual gradient].
template <typename T>
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_pre_kernel(GnDev d, int ppb) {
  xu_grid_dep_sync();
  constexpr int VW = VecW<T>::W;
  constexpr int U = 4;
  const int tid = threadIdx.x, b = blockIdx.y;
  __shared__ float s_s12[XU_GROUPS][2];
  __shared__ float s_grp[XU_GROUPS][2];
  {
    const int g = tid >> 3, sub = tid & 7;
    float s1 = 0.f, s2 = 0.f;
    for (int k = sub; k < d.cpg; k += 8) {
      const int c = g * d.cpg + k;
      const float2 ab = *reinterpret_cast<const float2*>(d.bcs + ((long long)b * d.C + c) * 2);
      const float gmc = d.gamma[c];
      s1 = fmaf(gmc, ab.y, s1);
      s2 = fmaf(gmc, ab.x, s2);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (sub == 0) { s_s12[g][0] = s1 * d.inv_cnt; s_s12[g][1] = s2 * d.inv_cnt; }
    if (tid < XU_GROUPS) {
      s_grp[tid][0] = d.stats[(b * XU_GROUPS + tid) * 2 + 0];
      s_grp[tid][1] = d.stats[(b * XU_GROUPS + tid) * 2 + 1];
    }
    if (blockIdx.x == 0) {     // one block per sample adds its channel sums to the parameter gradients
      for (int c = tid; c < d.C; c += 256) {
        const float2 ab = *reinterpret_cast<const float2*>(d.bcs + ((long long)b * d.C + c) * 2);
        atomicAdd(&d.dgamma[c], ab.x);
        atomicAdd(&d.dbeta[c], ab.y);
      }
    }
    __syncthreads();
  }
  const int CV = d.C / VW;
  const int TPB = CV < 256 ? CV : 256;
  const int PL = 256 / TPB;
  const int cv0 = tid % TPB, pl = tid / TPB;
  if (pl >= PL) return;
  const int HW = d.H * d.W, P = 2 * HW;
  const int pbeg = blockIdx.x * ppb;
  const int pend = min(pbeg + ppb, P);
  const long long pix0 = (long long)(2 * b) * HW;
  for (int cv = cv0; cv < CV; cv += TPB) {
    const int c0 = cv * VW;
    // dx = rstd*(gamma*dyh - S1 - xhat*S2)  with xhat = (x-mean)*rstd   ==   dyh*kg + x*kx + k0
    float kg[VW], kx[VW], k0[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) {
      float mean, rstd;
      gn_group_mean_rstd(d, s_grp, c0 + j, mean, rstd);
      const int g = gn_group(d, c0 + j);
      const float s1 = s_s12[g][0], s2 = s_s12[g][1];
      kg[j] = rstd * d.gamma[c0 + j];
      kx[j] = -rstd * rstd * s2;
      k0[j] = rstd * (mean * rstd * s2 - s1);
    }
    const T* X = reinterpret_cast<const T*>(d.x) + pix0 * d.C + c0;
    const T* G = reinterpret_cast<const T*>(d.dy) + pix0 * d.C + c0;
    const T* EX = d.extra ? reinterpret_cast<const T*>(d.extra) + pix0 * d.C + c0 : nullptr;
    T* DX = reinterpret_cast<T*>(d.y) + pix0 * d.C + c0;
    for (int p0 = pbeg + pl; p0 < pend; p0 += U * PL) {
      typename VecW<T>::raw xr[U], gr[U], er[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + u * PL;
        if (p < pend) {
          xr[u] = VecW<T>::ldg(X + (long long)p * d.C);
          gr[u] = VecW<T>::ldg(G + (long long)p * d.C);
          if (EX) er[u] = VecW<T>::ldg(EX + (long long)p * d.C);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + u * PL;
        if (p >= pend) break;
        float v[VW], g[VW], out[VW];
        VecW<T>::unpack(xr[u], v);
        VecW<T>::unpack(gr[u], g);
#pragma unroll
        for (int j = 0; j < VW; ++j) out[j] = fmaf(g[j], kg[j], fmaf(v[j], kx[j], k0[j]));
        if (EX) {
          float ex[VW];
          VecW<T>::unpack(er[u], ex);
#pragma unroll
          for (int j = 0; j < VW; ++j) out[j] = fmaf(d.extra_alpha, ex[j], out[j]);
        }
        T* dx = DX + (long long)p * d.C;
        if (d.accumulate) {
          float o[VW];
          VecW<T>::unpack(VecW<T>::ld(dx), o);
#pragma unroll
          for (int j = 0; j < VW; ++j) out[j] += o[j];
        }
        VecW<T>::st(dx, out);
      }
    }
  }
}

void launch_gn_bwd_apply_pre(int dtype, const GnArgs& a, cudaStream_t s) {
  GnDev d = gn_dev(a);
  dim3 grid; int ppb;
  gn_vec_grid(d.C, dtype == XU_F32 ? 4 : 8, 2 * d.H * d.W, d.N / 2, grid, ppb);
  if (dtype == XU_F32) xu_launch(gn_bwd_apply_pre_kernel<float>, grid, 256, 0, s, d, ppb);
  else xu_launch(gn_bwd_apply_pre_kernel<bf16>, grid, 256, 0, s, d, ppb);
}

void launch_gn_bwd_apply(int dtype, const GnArgs& a, cudaStream_t s) {
  GnDev d = gn_dev(a);
  dim3 grid; int ppb;
  gn_apply_grid(d.C, 2 * d.H * d.W, d.N / 2, grid, ppb);
  if (dtype == XU_F32) xu_launch(gn_bwd_apply_kernel<float>, grid, 256, 0, s, d, ppb);
  else xu_launch(gn_bwd_apply_kernel<bf16>, grid, 256, 0, s, d, ppb);
}

// ======================================================================================================
// log-SNR embedding  (model/xunet.py:153-157, posenc_ddpm :23-35)
// ======================================================================================================
// One Dense layer of the log-SNR MLP for one sample and 32 output columns per CTA: lane <-> column (128-byte coalesced weight rows),
// the 8 warps split K and are reduced through shared memory in a fixed order (deterministic).  FIRST: the input vector is the
// DDPM sinusoidal encoding of logsnr[b], built in shared memory (and exported to `pe` for the backward); otherwise swish(h1[b]).
// grid (E/32, B): E = 1024 -> 32*B CTAs instead of the B CTAs x 2E serial FMAs per thread of round 1 (581 us -> a few us).
template <bool FIRST>
__global__ void __launch_bounds__(256) logsnr_dense_kernel(const float* __restrict__ logsnr, const float* __restrict__ in,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ pe, float* __restrict__ out, int E) {
  xu_grid_dep_sync();
  extern __shared__ float sm[];  // sv[E] input vector, red[8][32]
  float* sv = sm;
  float* red = sm + E;
  const int b = blockIdx.y;
  const int j0 = blockIdx.x * 32;
  if (FIRST) {
    float l = fminf(fmaxf(logsnr[b], -20.f), 20.f);
    float t = 2.f * atanf(expf(-l * 0.5f)) / 3.14159265358979323846f;
    t *= 1000.f;  // posenc_ddpm(max_time=1.): timesteps *= 1000/max_time
    const int half = E / 2;
    const float c = (float)(-9.210340371976184 / (double)(half - 1));  // -log(10000)/(half-1)
    for (int k = threadIdx.x; k < E; k += blockDim.x) {
      int kk = k < half ? k : k - half;
      float f = expf((float)kk * c);
      float arg = t * f;
      float v = k < half ? sinf(arg) : cosf(arg);
      if (k >= 2 * half) v = 0.f;
      sv[k] = v;
      if (k >= j0 && k < j0 + 32) pe[b * E + k] = v;
    }
  } else {
    for (int k = threadIdx.x; k < E; k += blockDim.x) sv[k] = swishf_(in[b * E + k]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int j = j0 + lane;
  const int kper = (E + 7) / 8;
  const int k0 = wp * kper, k1 = min(E, k0 + kper);
  float acc = 0.f;
  if (j < E) {
#pragma unroll 8
    for (int k = k0; k < k1; ++k) acc = fmaf(sv[k], w[(size_t)k * E + j], acc);
  }
  red[wp * 32 + lane] = acc;
  __syncthreads();
  if (wp == 0 && j < E) {
    float r = bias[j];
#pragma unroll
    for (int q = 0; q < 8; ++q) r += red[q * 32 + lane];
    out[b * E + j] = r;
  }
}

void launch_logsnr_emb(const float* logsnr, const float* w0, const float* b0, const float* w1, const float* b1, float* pe,
                       float* h1, float* lemb, int B, int E, cudaStream_t s) {
  dim3 grid(cdiv(E, 32), B);
  const size_t smem = (size_t)(E + 256) * sizeof(float);
  xu_launch(logsnr_dense_kernel<true>, grid, 256, smem, s, logsnr, (const float*)nullptr, w0, b0, pe, h1, E);
  xu_launch(logsnr_dense_kernel<false>, grid, 256, smem, s, logsnr, (const float*)h1, w1, b1, pe, lemb, E);
}

// dh1[b,k] = swish'(h1[b,k]) * sum_j w1[k][j] dlemb[b,j]
__global__ void logsnr_bwd_dh1_kernel(const float* __restrict__ dlemb, const float* __restrict__ w1,
                                      const float* __restrict__ h1, float* __restrict__ dh1, int E) {
  xu_grid_dep_sync();
  const int b = blockIdx.y;
  const int k = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (k >= E) return;
  float acc = 0.f;
  for (int j = lane; j < E; j += 32) acc = fmaf(w1[k * E + j], dlemb[b * E + j], acc);
  acc = warp_sum(acc);
  if (lane == 0) dh1[b * E + k] = acc * swish_gradf_(h1[b * E + k]);
}
// dw1[k][j] = sum_b swish(h1[b,k]) dlemb[b,j];  dw0[k][j] = sum_b pe[b,k] dh1[b,j];  biases
__global__ void logsnr_bwd_w_kernel(const float* __restrict__ dlemb, const float* __restrict__ pe,
                                    const float* __restrict__ h1, const float* __restrict__ dh1, float* __restrict__ dw0,
                                    float* __restrict__ db0, float* __restrict__ dw1, float* __restrict__ db1, int B,
                                    int E) {
  xu_grid_dep_sync();
  const int k = blockIdx.x;
  for (int j = threadIdx.x; j < E; j += blockDim.x) {
    float a1 = 0.f, a0 = 0.f, s1 = 0.f, s0 = 0.f;
    for (int b = 0; b < B; ++b) {
      const float dl = dlemb[b * E + j], dh = dh1[b * E + j];
      a1 = fmaf(swishf_(h1[b * E + k]), dl, a1);
      a0 = fmaf(pe[b * E + k], dh, a0);
      s1 += dl;
      s0 += dh;
    }
    dw1[k * E + j] = a1;
    dw0[k * E + j] = a0;
    if (k == 0) { db1[j] = s1; db0[j] = s0; }
  }
}

void launch_logsnr_emb_bwd(const float* dlemb, const float* w1, const float* pe, const float* h1, float* dh1, float* dw0,
                           float* db0, float* dw1, float* db1, int B, int E, cudaStream_t s) {
  dim3 g1(cdiv(E, 8), B);
  xu_launch(logsnr_bwd_dh1_kernel, g1, 256, 0, s, dlemb, w1, h1, dh1, E);
  int threads = E < 256 ? ((E + 31) / 32) * 32 : 256;
  xu_launch(logsnr_bwd_w_kernel, E, threads, 0, s, dlemb, pe, h1, dh1, dw0, db0, dw1, db1, B, E);
}

// ======================================================================================================
// camera rays + NeRF positional encoding  (model/xunet.py:159-194, posenc_nerf :37-44)
// ======================================================================================================
__global__ void kinv_kernel(const float* __restrict__ K, float* __restrict__ kinv, int B) {
  xu_grid_dep_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double m[9];
  for (int i = 0; i < 9; ++i) m[i] = (double)K[b * 9 + i];
  double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  double id = 1.0 / det;
  double inv[9];
  inv[0] = c00 * id; inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  inv[3] = c01 * id; inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  inv[6] = c02 * id; inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  for (int i = 0; i < 9; ++i) kinv[b * 9 + i] = (float)inv[i];
}

// One block = 64 pixels of one frame.  The 93 position channels depend only on the frame (camera centre t), so they are
// evaluated once per block (their arguments reach 2^14 |t|: the slow range-reduction path of sinf) and broadcast; the ray
// direction of each pixel is formed once and shared by its 51 direction channels.
constexpr int kPosePix = 64;
__device__ __forceinline__ float pose_channel(int jj, int nd, const float (&v)[3]) {
  // posenc_nerf layout: [x (3) | sin(2^k x) k<nd (3 nd) | sin(2^k x + pi/2) (3 nd)]
  if (jj < 3) return v[jj];
  int k = jj - 3;
  const int phase = k >= 3 * nd;
  if (phase) k -= 3 * nd;
  const int sc = k / 3, d = k - sc * 3;
  float xb = v[d] * (float)(1 << sc);              // exact in fp32
  if (phase) xb = xb + 1.57079632679489661923f;    // fp32 add of pi/2, then accurate sin (not cos)
  return sinf(xb);
}

template <typename T>
__global__ void __launch_bounds__(256) pose_emb_kernel(const float* __restrict__ R1, const float* __restrict__ t1,
                                                       const float* __restrict__ R2, const float* __restrict__ t2,
                                                       const float* __restrict__ kinv, const float* __restrict__ cond_mask,
                                                       const float* __restrict__ pos_emb, const float* __restrict__ ref_first,
                                                       const float* __restrict__ ref_other, T* __restrict__ out, int B, int S,
                                                       int convention, const float* __restrict__ rays) {
  xu_grid_dep_sync();
  __shared__ float s_pos[93];
  __shared__ float s_dir[kPosePix][3];
  __shared__ float s_ppos[kPosePix][3];   // explicit-rays entry only: per-pixel ray origin
  const int tid = threadIdx.x;
  const int n = blockIdx.y, b = n >> 1, f = n & 1;
  const int HW = S * S;
  const int pix0 = blockIdx.x * kPosePix;
  const bool on = cond_mask[b] != 0.f;
  if (on && rays != nullptr) {
    // caller-supplied rays (B, 2, S, S, 6): v3d.Camera.rays() output fed straight in (model/xunet.py:159-161)
    if (tid < kPosePix && pix0 + tid < HW) {
      const float* r = rays + ((long long)n * HW + pix0 + tid) * 6;
      s_ppos[tid][0] = r[0]; s_ppos[tid][1] = r[1]; s_ppos[tid][2] = r[2];
      s_dir[tid][0] = r[3]; s_dir[tid][1] = r[4]; s_dir[tid][2] = r[5];
    }
  } else if (on) {
    const float* R = (f == 0 ? R1 : R2) + b * 9;
    const float* t = (f == 0 ? t1 : t2) + b * 3;
    if (tid < 93) {
      const float tv[3] = {t[0], t[1], t[2]};
      s_pos[tid] = pose_channel(tid, 15, tv);
    } else if (tid >= 128 && tid < 128 + kPosePix && pix0 + tid - 128 < HW) {
      const int pix = pix0 + tid - 128;
      const int row = pix / S, col = pix - row * S;
      const float p0 = (convention == 0 ? (float)row : (float)col) + 0.5f;
      const float p1 = (convention == 0 ? (float)col : (float)row) + 0.5f;
      const float* ki = kinv + b * 9;
      float cx = ki[0] * p0 + ki[1] * p1 + ki[2];
      float cy = ki[3] * p0 + ki[4] * p1 + ki[5];
      float cz = ki[6] * p0 + ki[7] * p1 + ki[8];
      float wx = R[0] * cx + R[1] * cy + R[2] * cz;
      float wy = R[3] * cx + R[4] * cy + R[5] * cz;
      float wz = R[6] * cx + R[7] * cy + R[8] * cz;
      float inv = 1.f / sqrtf(wx * wx + wy * wy + wz * wz);
      s_dir[tid - 128][0] = wx * inv; s_dir[tid - 128][1] = wy * inv; s_dir[tid - 128][2] = wz * inv;
    }
  }
  __syncthreads();
  const int npix = min(kPosePix, HW - pix0);
  T* o = out + ((long long)n * HW + pix0) * XU_POSE_DIM;
  const float* pe = pos_emb ? pos_emb + (long long)pix0 * XU_POSE_DIM : nullptr;
  const float* rf = ref_first ? (f == 0 ? ref_first : ref_other) : nullptr;
  for (int i = tid; i < npix * XU_POSE_DIM; i += 256) {
    const int pl = i / XU_POSE_DIM, j = i - pl * XU_POSE_DIM;
    float val = 0.f;
    if (on) {
      if (j < 93) {
        if (rays != nullptr) {
          const float pv[3] = {s_ppos[pl][0], s_ppos[pl][1], s_ppos[pl][2]};
          val = pose_channel(j, 15, pv);
        } else val = s_pos[j];
      } else {
        const float dv[3] = {s_dir[pl][0], s_dir[pl][1], s_dir[pl][2]};
        val = pose_channel(j - 93, 8, dv);
      }
    }
    if (pe) val += pe[i];
    if (rf) val += rf[j];
    stf(o + i, val);
  }
}

void launch_pose_emb(int dtype, const float* R1, const float* t1, const float* R2, const float* t2, const float* K,
                     const float* cond_mask, const float* pos_emb, const float* ref_first, const float* ref_other,
                     float* kinv_scratch, void* out, int B, int S, int convention, const float* rays, cudaStream_t s) {
  if (rays == nullptr) xu_launch(kinv_kernel, cdiv(B, 64), 64, 0, s, K, kinv_scratch, B);
  const dim3 grid(cdiv(S * S, kPosePix), 2 * B);
  if (dtype == XU_F32)
    xu_launch(pose_emb_kernel<float>, grid, 256, 0, s, R1, t1, R2, t2, kinv_scratch, cond_mask, pos_emb, ref_first,
                                                           ref_other, (float*)out, B, S, convention, rays);
  else
    xu_launch(pose_emb_kernel<bf16>, grid, 256, 0, s, R1, t1, R2, t2, kinv_scratch, cond_mask, pos_emb, ref_first,
                                                          ref_other, (bf16*)out, B, S, convention, rays);
}

template <typename T>
__global__ void pose_emb_bwd_kernel(const T* __restrict__ dpose, float* __restrict__ dpos_emb,
                                    float* __restrict__ dref_first, float* __restrict__ dref_other, int B, int S) {
  xu_grid_dep_sync();
  const long long per = (long long)S * S * XU_POSE_DIM;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= per) return;
  const int j = (int)(idx % XU_POSE_DIM);
  float s0 = 0.f, s1 = 0.f;
  for (int b = 0; b < B; ++b) {
    s0 += ldf(dpose + (long long)(2 * b) * per + idx);
    s1 += ldf(dpose + (long long)(2 * b + 1) * per + idx);
  }
  if (dpos_emb != nullptr) dpos_emb[idx] = s0 + s1;
  if (dref_first != nullptr) { atomicAdd(&dref_first[j], s0); atomicAdd(&dref_other[j], s1); }
}

void launch_pose_emb_bwd(int dtype, const void* dpose, float* dpos_emb, float* dref_first, float* dref_other, int B, int S,
                         cudaStream_t s) {
  const long long per = (long long)S * S * XU_POSE_DIM;
  if (dtype == XU_F32) xu_launch(pose_emb_bwd_kernel<float>, cdiv(per, 256), 256, 0, s, (const float*)dpose, dpos_emb, dref_first, dref_other, B, S);
  else xu_launch(pose_emb_bwd_kernel<bf16>, cdiv(per, 256), 256, 0, s, (const bf16*)dpose, dpos_emb, dref_first, dref_other, B, S);
}

// ======================================================================================================
// emb = swish(logsnr_emb[b] + pose_emb_level)   (FiLM's nonlinearity(emb), model/xunet.py:59,233)
// ======================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) emb_fwd_kernel(const float* __restrict__ lemb, const T* __restrict__ pe,
                                                      T* __restrict__ semb, int N, int HW, int E) {
  xu_grid_dep_sync();
  const int E4 = E >> 2;
  const long long total = (long long)N * HW * E4;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int cv = (int)(idx % E4);
  const long long pix = idx / E4;
  const int b = (int)(pix / HW) >> 1;
  float v[4];
  Vec4<T>::ld(pe + pix * E + cv * 4, v);
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = swishf_(v[j] + lemb[b * E + cv * 4 + j]);
  Vec4<T>::st(semb + pix * E + cv * 4, v);
}

void launch_emb_fwd(int dtype, const float* lemb, const void* pe, void* semb, int N, int HW, int E, cudaStream_t s) {
  const long long total = (long long)N * HW * (E / 4);
  if (dtype == XU_F32) xu_launch(emb_fwd_kernel<float>, cdiv(total, 256), 256, 0, s, lemb, (const float*)pe, (float*)semb, N, HW, E);
  else xu_launch(emb_fwd_kernel<bf16>, cdiv(total, 256), 256, 0, s, lemb, (const bf16*)pe, (bf16*)semb, N, HW, E);
}

// grid (chunks, B): dz = dsemb * swish'(lemb+pe); dpe = dz (optional); dlemb[b,c] += sum over this block's pixels
template <typename T>
__global__ void __launch_bounds__(256) emb_bwd_kernel(const float* __restrict__ lemb, const T* __restrict__ pe,
                                                      const T* __restrict__ dsemb, T* __restrict__ dpe,
                                                      float* __restrict__ dlemb, int P, int E, int ppb, int write_dpe) {
  xu_grid_dep_sync();
  extern __shared__ float sacc[];  // E
  const int tid = threadIdx.x, b = blockIdx.y;
  for (int i = tid; i < E; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int E4 = E >> 2;
  const int TPB = E4 < 256 ? E4 : 256;
  const int PL = 256 / TPB;
  const int cv0 = tid % TPB, pl = tid / TPB;
  const int pbeg = blockIdx.x * ppb, pend = min(pbeg + ppb, P);
  {
    for (int cv = cv0; cv < E4; cv += TPB) {
      float le[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) le[j] = lemb[b * E + cv * 4 + j];
      if (pl < PL)
      for (int p = pbeg + pl; p < pend; p += PL) {
        const long long off = ((long long)b * P + p) * E + cv * 4;
        float v[4], g[4];
        Vec4<T>::ld(pe + off, v);
        Vec4<T>::ld(dsemb + off, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) { g[j] *= swish_gradf_(v[j] + le[j]); acc[j] += g[j]; }
        if (write_dpe) Vec4<T>::st(dpe + off, g);
      }
      if (fold_same_channel_lanes(acc, TPB) && pl < PL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&sacc[cv * 4 + j], acc[j]);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < E; i += 256) atomicAdd(&dlemb[b * E + i], sacc[i]);
}

void launch_emb_bwd(int dtype, const float* lemb, const void* pe, const void* dsemb, void* dpe, float* dlemb, int N, int HW,
                    int E, int write_dpe, cudaStream_t s) {
  const int B = N / 2, P = 2 * HW;
  dim3 grid; int ppb;
  gn_grid(E, P, B, grid, ppb);
  size_t smem = sizeof(float) * E;
  if (dtype == XU_F32)
    xu_launch(emb_bwd_kernel<float>, grid, 256, smem, s, lemb, (const float*)pe, (const float*)dsemb, (float*)dpe, dlemb, P, E, ppb, write_dpe);
  else
    xu_launch(emb_bwd_kernel<bf16>, grid, 256, smem, s, lemb, (const bf16*)pe, (const bf16*)dsemb, (bf16*)dpe, dlemb, P, E, ppb, write_dpe);
}

// ======================================================================================================
// plumbing
// ======================================================================================================
template <typename T>
__global__ void pack_input_kernel(const float* __restrict__ x, const float* __restrict__ z, T* __restrict__ out, int B,
                                  long long per) {
  xu_grid_dep_sync();
  const long long total = (long long)B * 2 * per;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / per, r = idx - n * per;
  const long long b = n >> 1;
  stf(out + idx, (n & 1) ? z[b * per + r] : x[b * per + r]);
}
void launch_pack_input(int dtype, const float* x, const float* z, void* out, int B, int S, cudaStream_t s) {
  const long long per = (long long)S * S * 3, total = 2 * B * per;
  if (dtype == XU_F32) xu_launch(pack_input_kernel<float>, cdiv(total, 256), 256, 0, s, x, z, (float*)out, B, per);
  else xu_launch(pack_input_kernel<bf16>, cdiv(total, 256), 256, 0, s, x, z, (bf16*)out, B, per);
}

template <typename T>
__global__ void __launch_bounds__(256) resample_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int Hi, int Wi,
                                                       int C, int pool, float scale, int accumulate) {
  xu_grid_dep_sync();
  const int Ho = pool ? Hi / 2 : Hi * 2, Wo = pool ? Wi / 2 : Wi * 2;
  const int C4 = C >> 2;
  const long long total = (long long)N * Ho * Wo * C4;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int cv = (int)(idx % C4);
  const long long pix = idx / C4;
  const int ox = (int)(pix % Wo);
  const int oy = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((long long)Wo * Ho));
  float out[4] = {0.f, 0.f, 0.f, 0.f};
  if (pool) {
    for (int i = 0; i < 2; ++i)
      for (int k = 0; k < 2; ++k) {
        float v[4];
        Vec4<T>::ld(x + (((long long)n * Hi + 2 * oy + i) * Wi + 2 * ox + k) * C + cv * 4, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] += v[j];
      }
  } else {
    Vec4<T>::ld(x + (((long long)n * Hi + (oy >> 1)) * Wi + (ox >> 1)) * C + cv * 4, out);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) out[j] *= scale;
  T* dst = y + pix * C + cv * 4;
  if (accumulate) {
    float o[4];
    Vec4<T>::ld(dst, o);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] += o[j];
  }
  Vec4<T>::st(dst, out);
}
void launch_resample(int dtype, const void* x, void* y, int N, int Hi, int Wi, int C, int pool, float scale, int accumulate,
                     cudaStream_t s) {
  const int Ho = pool ? Hi / 2 : Hi * 2, Wo = pool ? Wi / 2 : Wi * 2;
  const long long total = (long long)N * Ho * Wo * (C / 4);
  if (dtype == XU_F32) xu_launch(resample_kernel<float>, cdiv(total, 256), 256, 0, s, (const float*)x, (float*)y, N, Hi, Wi, C, pool, scale, accumulate);
  else xu_launch(resample_kernel<bf16>, cdiv(total, 256), 256, 0, s, (const bf16*)x, (bf16*)y, N, Hi, Wi, C, pool, scale, accumulate);
}

template <typename T>
__global__ void __launch_bounds__(256) copy_channels_kernel(const T* __restrict__ src, T* __restrict__ dst, long long npix,
                                                            int Cs, int Cd, int so, int doff, int Cc, int accumulate) {
  xu_grid_dep_sync();
  const int C4 = Cc >> 2;
  const long long total = npix * C4;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int cv = (int)(idx % C4);
  const long long pix = idx / C4;
  float v[4];
  Vec4<T>::ld(src + pix * Cs + so + cv * 4, v);
  T* d = dst + pix * Cd + doff + cv * 4;
  if (accumulate) {
    float o[4];
    Vec4<T>::ld(d, o);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += o[j];
  }
  Vec4<T>::st(d, v);
}
void launch_copy_channels(int dtype, const void* src, void* dst, long long npix, int Cs, int Cd, int src_off, int dst_off,
                          int Cc, int accumulate, cudaStream_t s) {
  const long long total = npix * (Cc / 4);
  if (dtype == XU_F32)
    xu_launch(copy_channels_kernel<float>, cdiv(total, 256), 256, 0, s, (const float*)src, (float*)dst, npix, Cs, Cd, src_off, dst_off, Cc, accumulate);
  else
    xu_launch(copy_channels_kernel<bf16>, cdiv(total, 256), 256, 0, s, (const bf16*)src, (bf16*)dst, npix, Cs, Cd, src_off, dst_off, Cc, accumulate);
}

template <typename T>
__global__ void __launch_bounds__(256) scale_add_kernel(const T* __restrict__ src, T* __restrict__ dst, long long n4,
                                                        float alpha, int accumulate) {
  xu_grid_dep_sync();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n4) return;
  float v[4];
  Vec4<T>::ld(src + idx * 4, v);
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] *= alpha;
  if (accumulate) {
    float o[4];
    Vec4<T>::ld(dst + idx * 4, o);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += o[j];
  }
  Vec4<T>::st(dst + idx * 4, v);
}
void launch_scale_add(int dtype, const void* src, void* dst, long long n, float alpha, int accumulate, cudaStream_t s) {
  const long long n4 = n / 4;  // all activation tensors here have C % 4 == 0
  if (dtype == XU_F32) xu_launch(scale_add_kernel<float>, cdiv(n4, 256), 256, 0, s, (const float*)src, (float*)dst, n4, alpha, accumulate);
  else xu_launch(scale_add_kernel<bf16>, cdiv(n4, 256), 256, 0, s, (const bf16*)src, (bf16*)dst, n4, alpha, accumulate);
}

template <typename T>
__global__ void extract_frame1_kernel(const T* __restrict__ o, float* __restrict__ eps, int B, long long per) {
  xu_grid_dep_sync();
  const long long total = (long long)B * per;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const long long b = idx / per, r = idx - b * per;
  eps[idx] = ldf(o + (2 * b + 1) * per + r);
}
void launch_extract_frame1(int dtype, const void* o, float* eps, int B, int S, cudaStream_t s) {
  const long long per = (long long)S * S * 3, total = B * per;
  if (dtype == XU_F32) xu_launch(extract_frame1_kernel<float>, cdiv(total, 256), 256, 0, s, (const float*)o, eps, B, per);
  else xu_launch(extract_frame1_kernel<bf16>, cdiv(total, 256), 256, 0, s, (const bf16*)o, eps, B, per);
}

// loss = ||eps - noise||_F  (train.py:67)
__global__ void __launch_bounds__(256) loss_sumsq_kernel(const float* __restrict__ eps, const float* __restrict__ noise,
                                                         long long n, float* __restrict__ sumsq) {
  xu_grid_dep_sync();
  __shared__ float sw[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float r = eps[i] - noise[i];
    acc = fmaf(r, r, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = sw[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(sumsq, v);
  }
}
template <typename T>
__global__ void loss_grad_kernel(const float* __restrict__ eps, const float* __restrict__ noise,
                                 const float* __restrict__ sumsq, float* __restrict__ loss_out, T* __restrict__ dO, int B,
                                 long long per) {
  xu_grid_dep_sync();
  const long long total = (long long)B * 2 * per;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const float loss = sqrtf(*sumsq);
  if (idx == 0) *loss_out = loss;
  if (idx >= total) return;
  const long long n = idx / per, r = idx - n * per;
  float g = 0.f;
  if (n & 1) {
    const long long i = (n >> 1) * per + r;
    g = loss > 0.f ? (eps[i] - noise[i]) / loss : 0.f;
  }
  stf(dO + idx, g);
}
void launch_loss(int dtype, const float* eps, const float* noise, float* sumsq_scratch, float* loss_out, void* dO, int B,
                 int S, cudaStream_t s) {
  const long long per = (long long)S * S * 3, n = B * per;
  cudaMemsetAsync(sumsq_scratch, 0, sizeof(float), s);
  int blocks = cdiv(n, 256 * 4);
  if (blocks > 592) blocks = 592;
  xu_launch(loss_sumsq_kernel, blocks, 256, 0, s, eps, noise, n, sumsq_scratch);
  const long long total = 2 * n;
  if (dtype == XU_F32) xu_launch(loss_grad_kernel<float>, cdiv(total, 256), 256, 0, s, eps, noise, sumsq_scratch, loss_out, (float*)dO, B, per);
  else xu_launch(loss_grad_kernel<bf16>, cdiv(total, 256), 256, 0, s, eps, noise, sumsq_scratch, loss_out, (bf16*)dO, B, per);
}

// optax.adam (train.py:45, 74-76).  (1-b1), (1-b2) and the bias corrections are formed in double (optax forms them
// from python floats) and only then rounded to fp32.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, long long step,
                                                   const long long* __restrict__ step_dev, double lr, double b1d, double b2d,
                                                   float eps, float gs) {
  xu_grid_dep_sync();
  __shared__ float sc[2];
  if (threadIdx.x == 0) {
    const long long st = step_dev != nullptr ? *step_dev : step;
    sc[0] = (float)(1.0 / (1.0 - pow(b1d, (double)st)));
    sc[1] = (float)(1.0 / (1.0 - pow(b2d, (double)st)));
  }
  __syncthreads();
  const float c1 = sc[0], c2 = sc[1];
  const float b1 = (float)b1d, b2 = (float)b2d, omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d), lrf = (float)lr;
  // 16-byte accesses (7 streams of n floats: 4 read, 3 written -- the kernel is pure HBM traffic); scalar path for a ragged tail or
  // unaligned sub-ranges
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long n4 = vec ? (n >> 2) : 0;
  auto upd = [&](float gi, float& mi, float& vi, float& pi) {
    gi *= gs;
    mi = b1 * mi + omb1 * gi;
    vi = b2 * vi + omb2 * gi * gi;
    pi -= lrf * (mi * c1) / (sqrtf(vi * c2) + eps);
  };
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    upd(g4.x, m4.x, v4.x, p4.x); upd(g4.y, m4.y, v4.y, p4.y); upd(g4.z, m4.z, v4.z, p4.z); upd(g4.w, m4.w, v4.w, p4.w);
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    reinterpret_cast<float4*>(p)[i] = p4;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float mi = m[i], vi = v[i], pi = p[i];
    upd(g[i], mi, vi, pi);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
void launch_adam(float* p, const float* g, float* m, float* v, long long n, long long step, const long long* step_dev,
                 double lr, double b1, double b2, double eps, double grad_scale, cudaStream_t s) {
  int blocks = cdiv(n, 256 * 4);   // one float4 per thread and pass
  if (blocks > xu_num_sms() * 8) blocks = xu_num_sms() * 8;
  if (blocks < 1) blocks = 1;
  xu_launch(adam_kernel, blocks, 256, 0, s, p, g, m, v, n, step, step_dev, lr, b1, b2, (float)eps, (float)grad_scale);
}

// sampling.py:128-151 elementwise update
__global__ void sampler_update_kernel(const float* __restrict__ eps2, const float* __restrict__ z,
                                      const float* __restrict__ noise, float* __restrict__ z_out, long long n, float w,
                                      float c_recip, float c_recipm1, float c1, float c2, float sigma,
                                      unsigned long long seed) {
  xu_grid_dep_sync();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float e = (1.f + w) * eps2[i] - w * eps2[n + i];
  const float zi = z[i];
  float x0 = c_recip * zi - c_recipm1 * e;
  x0 = fminf(fmaxf(x0, -1.f), 1.f);
  float nz;
  if (noise != nullptr) nz = noise[i];
  else {  // Box-Muller on two hashed uniforms
    uint64_t h1 = xu_mix64(seed * 0x9E3779B97F4A7C15ULL + 2ULL * (uint64_t)i + 1ULL);
    uint64_t h2 = xu_mix64(h1 + 0xD1B54A32D192ED03ULL);
    float u1 = ((float)(h1 >> 40) + 1.f) * (1.0f / 16777216.0f);
    float u2 = (float)(h2 >> 40) * (1.0f / 16777216.0f);
    nz = sqrtf(-2.f * logf(u1)) * cosf(6.28318530717958647692f * u2);
  }
  z_out[i] = c1 * x0 + c2 * zi + sigma * nz;
}
void launch_sampler_update(const float* eps2, const float* z, const float* noise, float* z_out, long long n, float w,
                           float c_recip, float c_recipm1, float c1, float c2, float sigma, unsigned long long seed,
                           cudaStream_t s) {
  xu_launch(sampler_update_kernel, cdiv(n, 256), 256, 0, s, eps2, z, noise, z_out, n, w, c_recip, c_recipm1, c1, c2, sigma, seed);
}

// One ancestral step with the schedule ON THE DEVICE (sampling.py:128-151): the coefficients of loop position k = *pos_dev come
// from a table (k = 0 is the first executed step = highest t), and the kernel also writes the NEXT forward's inputs (z into
// both halves of the [cond ; uncond] batch, the lagging log-SNR), so one CUDA graph = forward + this kernel + counter++ is
// replayed per step with no host arithmetic or copies in between.  tab row: {c_recip, c_recipm1, c1, c2, sigma, logsnr_next, -, -}
__global__ void __launch_bounds__(256) sampler_step_table_kernel(const float* __restrict__ eps2, float* __restrict__ z, long long n,
                                                                 float w, const float* __restrict__ tab, const int* __restrict__ pos_dev,
                                                                 const unsigned long long* __restrict__ seed_dev, float* __restrict__ inp_z,
                                                                 float* __restrict__ inp_logsnr, int B2) {
  xu_grid_dep_sync();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const int k = *pos_dev;
  const float* t = tab + 8LL * k;
  if (i < B2) inp_logsnr[i] = t[5];
  if (i >= n) return;
  const float e = (1.f + w) * eps2[i] - w * eps2[n + i];
  const float zi = z[i];
  float x0 = t[0] * zi - t[1] * e;
  x0 = fminf(fmaxf(x0, -1.f), 1.f);
  const unsigned long long seed = *seed_dev - (unsigned long long)k;     // = seed0 * 1000003 + timestep index, as the host loop
  uint64_t h1 = xu_mix64(seed * 0x9E3779B97F4A7C15ULL + 2ULL * (uint64_t)i + 1ULL);
  uint64_t h2 = xu_mix64(h1 + 0xD1B54A32D192ED03ULL);
  float u1 = ((float)(h1 >> 40) + 1.f) * (1.0f / 16777216.0f);
  float u2 = (float)(h2 >> 40) * (1.0f / 16777216.0f);
  const float nz = sqrtf(-2.f * logf(u1)) * cosf(6.28318530717958647692f * u2);
  const float zn = t[2] * x0 + t[3] * zi + t[4] * nz;
  z[i] = zn;
  inp_z[i] = zn;
  inp_z[n + i] = zn;
}
void launch_sampler_step_table(const float* eps2, float* z, long long n, float w, const float* tab, const int* pos_dev,
                               const unsigned long long* seed_dev, float* inp_z, float* inp_logsnr, int B2, cudaStream_t s) {
  xu_launch(sampler_step_table_kernel, cdiv(n > B2 ? n : B2, 256), 256, 0, s, eps2, z, n, w, tab, pos_dev, seed_dev, inp_z, inp_logsnr, B2);
}

// dataset/data_loader.py:92-110 on the device
__global__ void __launch_bounds__(256) forward_diffusion_kernel(const float* __restrict__ x0, const float* __restrict__ noise_in,
                                                                const int* __restrict__ t_in, unsigned long long seed,
                                                                const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1mac,
                                                                float p_uncond, float* __restrict__ z, float* __restrict__ noise_out,
                                                                float* __restrict__ logsnr_out, int* __restrict__ t_out,
                                                                float* __restrict__ cond_mask_out, int B, long long per) {
  xu_grid_dep_sync();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * per) return;
  const int b = (int)(idx / per);
  const uint64_t hb = xu_mix64(seed * 0x9E3779B97F4A7C15ULL + 0xABCDEF0123ULL + (uint64_t)b);
  const int t = t_in != nullptr ? t_in[b] : (int)(hb % 1000ULL);
  float nz;
  if (noise_in != nullptr) nz = noise_in[idx];
  else {
    const uint64_t h1 = xu_mix64(seed * 0x9E3779B97F4A7C15ULL + 2ULL * (uint64_t)idx + 1ULL);
    const uint64_t h2 = xu_mix64(h1 + 0xD1B54A32D192ED03ULL);
    const float u1 = ((float)(h1 >> 40) + 1.f) * (1.0f / 16777216.0f);
    const float u2 = (float)(h2 >> 40) * (1.0f / 16777216.0f);
    nz = sqrtf(-2.f * logf(u1)) * cosf(6.28318530717958647692f * u2);
  }
  z[idx] = sqrt_ac[t] * x0[idx] + sqrt_1mac[t] * nz;
  if (noise_out != nullptr) noise_out[idx] = nz;
  if (idx == (long long)b * per) {
    // logsnr_schedule_cosine(t/1000), logsnr_min=-20, logsnr_max=20 (data_loader.py:94-97), in double like the reference
    const double bb = atan(exp(-10.0)), aa = atan(exp(10.0)) - bb;
    logsnr_out[b] = (float)(-2.0 * log(tan(aa * ((double)t / 1000.0) + bb)));
    if (t_out != nullptr) t_out[b] = t;
    if (cond_mask_out != nullptr) {
      const float u = (float)(xu_mix64(hb + 0x51ED27ULL) >> 40) * (1.0f / 16777216.0f);
      cond_mask_out[b] = u > p_uncond ? 1.f : 0.f;
    }
  }
}
void launch_forward_diffusion(const float* x0, const float* noise_in, const int* t_in, unsigned long long seed,
                              const float* sqrt_ac, const float* sqrt_1mac, float p_uncond, float* z, float* noise_out,
                              float* logsnr_out, int* t_out, float* cond_mask_out, int B, long long per, cudaStream_t s) {
  const long long n = (long long)B * per;
  xu_launch(forward_diffusion_kernel, cdiv(n, 256), 256, 0, s, x0, noise_in, t_in, seed, sqrt_ac, sqrt_1mac, p_uncond, z, noise_out,
                                                        logsnr_out, t_out, cond_mask_out, B, per);
}

__global__ void dropout_mask_kernel(float* __restrict__ out, long long n, int op_index, unsigned long long seed, float rate) {
  xu_grid_dep_sync();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = xu_keep(seed, op_index, (unsigned long long)i, rate) ? 1.f : 0.f;
}
void launch_dropout_mask(float* out, long long n, int op_index, unsigned long long seed, float rate, cudaStream_t s) {
  xu_launch(dropout_mask_kernel, cdiv(n, 256), 256, 0, s, out, n, op_index, seed, rate);
}
