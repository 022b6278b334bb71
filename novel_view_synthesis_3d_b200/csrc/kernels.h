// kernels.h -- host launchers of the hand-written kernels (all asynchronous on `s`).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

enum { XU_F32 = 0, XU_BF16 = 1 };

// ---- implicit-GEMM convolution / per-pixel dense (k=1), forward and data-gradient ------------------
// mode 0 (forward):  y[n,oy,ox,co] = alpha * ( sum_{tap,ci} x[n, oy*s+dy-ph, ox*s+dx-pw, ci] W[tap][ci][co] + bias[co] + res )
// mode 1 (dgrad):    y = dX (N,Ho,Wo,Co=wCi),  x = dY (N,Hi,Wi,Ci=wCo):
//                    y[n,iy,ix,ci] (+)= alpha * sum_{tap,co} dY[n,(iy+ph-dy)/s,(ix+pw-dx)/s,co] W[tap][ci][co]
// weights: Flax layout [tap][wCi][wCo] fp32, with wCo split into segments of width segw laid out one after
// the other ([seg][tap][wCi][segw]; nseg>1 only for the fused q|k|v projection with ks==1).
struct ConvArgs {
  const void* x; void* y; const void* res; const float* w; const float* bias;
  int N, Hi, Wi, Ci, Ho, Wo, Co;
  int ks, stride, pad_h, pad_w, mode;
  int wCi, wCo, segw;
  float alpha; int accumulate;
  // optional (tcgen05 forward only, see conv_tc_stats_supported): per-(sample, channel) [sum, sumsq] of the STORED output,
  // (N/2, Co, 2) fp32, accumulated with atomics into a zeroed buffer -- the GroupNorm statistics of the consumer norm
  float* cstats = nullptr;
  // optional (tcgen05 data gradient, mode 1): this conv consumes the output of a GroupNorm[+swish] without resampling.  The
  // epilogue then stores dyh = dy * swish'(.) instead of dy and accumulates the per-(sample, channel) sums [sum dyh*xhat,
  // sum dyh] into cstats (N/2, Co, 2).  gn_x: the norm's input (N,H,W,Co); gn_params: (N/2, Co) float4 {rstd, -mean*rstd,
  // gamma, beta} written by the forward launch_gn_apply (GnArgs::params_out)
  const void* gn_x = nullptr; const void* gn_params = nullptr; int gn_swish = 0;
  // ... and when that norm is the FiLM norm of a ResnetBlock (u = yhat*(1+scale)+shift -> swish -> dropout -> this conv):
  // gn_e (N,H,W,2*Co) [scale | shift], gn_de its gradient (written), dropout rate / residual-block index / device seed
  const void* gn_e = nullptr; void* gn_de = nullptr;
  float gn_drop_rate = 0.f; int gn_drop_op = 0; const unsigned long long* gn_seed_dev = nullptr; int gn_train = 0;
};
void launch_conv_simt(int dtype, const ConvArgs& a, cudaStream_t s);

// dW[tap][ci][co] += alpha * sum_{n,oy,ox} x[n,oy*s+dy-ph,ox*s+dx-pw,ci] dY[n,oy,ox,co];  dbias[co] += alpha*sum dY
struct WgradArgs {
  const void* x; const void* dy; float* dw; float* dbias;
  int N, Hi, Wi, Ci, Ho, Wo, Co;
  int ks, stride, pad_h, pad_w, segw;
  float alpha;
};
void launch_wgrad_simt(int dtype, const WgradArgs& a, cudaStream_t s);
// direct kernels for the 3-channel convs.  which: 0 = Cin3 forward, 1 = Cin3 wgrad, 2 = Cout3 forward, 3 = Cout3 dgrad
// (ConvArgs in mode-1 convention: x = dO, y = dX), 4 = Cout3 wgrad
bool conv_cin3_supported(const ConvArgs& a);
bool conv_cout3_supported(int Ci, int Co, int ks, int stride);
void launch_conv_small(int dtype, int which, const ConvArgs* c, const WgradArgs* w, cudaStream_t s);

// ---- GroupNorm(32) over both frames (+ SiLU / FiLM / dropout / resample) -----------------------------
enum { GN_PLAIN = 0, GN_SWISH = 1, GN_FILM = 2 };
enum { RS_NONE = 0, RS_DOWN = 1, RS_UP = 2 };
struct GnArgs {
  const void* x;      // (N,H,W,C) input
  void* y;            // forward output (N,Ho,Wo,C)  /  backward: dx (N,H,W,C)
  const void* e;      // FiLM (N,H,W,2C): [scale | shift]
  const void* dy;     // backward: grad of output (N,Ho,Wo,C)
  void* de;           // backward: grad of e
  const float* gamma; const float* beta;
  float* dgamma; float* dbeta;   // backward (atomic accumulate)
  float* stats;       // (B,32,2) sum, sumsq of x         (forward writes, backward reads)
  // producer-emitted statistics (forward apply only): channels [0, csA) of x come with per-channel sums cstatsA (B, csA, 2),
  // channels [csA, C) with cstatsB (B, C - csA, 2) (a channel concat of two producers); cstatsA == NULL: `stats` was filled
  // by launch_gn_stats.  With cstats the apply kernel reduces them to group sums itself and stores them to `stats`.
  const float* cstatsA; const float* cstatsB; int csA;
  float* params_out;   // forward apply, optional: (B, C) float4 {rstd, -mean*rstd, gamma, beta} for the fused backward epilogue
  const float* bcs;    // launch_gn_bwd_apply_pre: per-(sample, channel) [sum dyh*xhat, sum dyh] emitted by that epilogue
  float* bstats;      // (B,32,2) backward group sums S1,S2
  int N, H, W, C;
  int mode, rs;
  float drop_rate; int op_index; const unsigned long long* seed_dev; int train;
  int accumulate;     // backward: dx += instead of =
  int de_accumulate;
  const void* extra;  // backward: optional second gradient stream into dx: dx += extra_alpha * extra  (same shape as x)
  float extra_alpha;
  int skip_zero;      // stats / bstats were already zeroed by the caller (one memset for the whole plan)
};
void launch_gn_stats(int dtype, const GnArgs& a, cudaStream_t s);        // zeroes + fills a.stats
// per-(sample, channel) [sum, sumsq] of x (N,H,W,C) accumulated into cstats (N/2, C, 2): the stand-alone producer of the
// statistics a conv / attention epilogue emits for free (used after kernels that cannot: 3-channel input conv, SIMT paths)
void launch_gn_cstats(int dtype, const void* x, float* cstats, int N, int H, int W, int C, cudaStream_t s);
void launch_gn_apply(int dtype, const GnArgs& a, cudaStream_t s);
void launch_gn_bwd_reduce(int dtype, const GnArgs& a, cudaStream_t s);   // zeroes + fills a.bstats, dgamma/dbeta, de
void launch_gn_bwd_apply(int dtype, const GnArgs& a, cudaStream_t s);
// second (and only) pass of the GroupNorm backward when the consumer conv's data-gradient epilogue already produced dyh and the
// channel sums (a.dy = dyh, a.bcs): folds the sums into dgamma / dbeta / group sums, then dx = rstd*(gamma*dyh - S1 - xhat*S2)
void launch_gn_bwd_apply_pre(int dtype, const GnArgs& a, cudaStream_t s);

// ---- conditioning -------------------------------------------------------------------------------------
// logsnr (B) -> posenc_ddpm -> Dense -> swish -> Dense.  pe,h1: saved (B,E) fp32.  lemb (B,E) fp32
void launch_logsnr_emb(const float* logsnr, const float* w0, const float* b0, const float* w1, const float* b1,
                       float* pe, float* h1, float* lemb, int B, int E, cudaStream_t s);
void launch_logsnr_emb_bwd(const float* dlemb, const float* w1, const float* pe, const float* h1, float* dh1,
                           float* dw0, float* db0, float* dw1, float* db1, int B, int E, cudaStream_t s);
// rays + NeRF posenc -> (2B,S,S,144)
void launch_pose_emb(int dtype, const float* R1, const float* t1, const float* R2, const float* t2, const float* K,
                     const float* cond_mask, const float* pos_emb, const float* ref_first, const float* ref_other,
                     float* kinv_scratch, void* out, int B, int S, int convention, const float* rays, cudaStream_t s);
void launch_pose_emb_bwd(int dtype, const void* dpose, float* dpos_emb, float* dref_first, float* dref_other, int B, int S,
                         cudaStream_t s);
// semb = swish(lemb[b] + pe)   /   bwd: dpe (+)= dsemb*swish'(z), dlemb[b] += sum
void launch_emb_fwd(int dtype, const float* lemb, const void* pe, void* semb, int N, int HW, int E, cudaStream_t s);
void launch_emb_bwd(int dtype, const float* lemb, const void* pe, const void* dsemb, void* dpe, float* dlemb, int N, int HW,
                    int E, int write_dpe, cudaStream_t s);

// ---- plumbing -------------------------------------------------------------------------------------------
void launch_pack_input(int dtype, const float* x, const float* z, void* out, int B, int S, cudaStream_t s);
// pool (2x2 sum * scale) or replicate (x2 nearest * scale)
void launch_resample(int dtype, const void* x, void* y, int N, int Hi, int Wi, int C, int pool, float scale, int accumulate,
                     cudaStream_t s);
// dst[:, dst_off : dst_off+Cc] (+)= src[:, src_off : src_off+Cc]
void launch_copy_channels(int dtype, const void* src, void* dst, long long npix, int Cs, int Cd, int src_off, int dst_off,
                          int Cc, int accumulate, cudaStream_t s);
void launch_scale_add(int dtype, const void* src, void* dst, long long n, float alpha, int accumulate, cudaStream_t s);
void launch_extract_frame1(int dtype, const void* o, float* eps, int B, int S, cudaStream_t s);
// loss = ||eps - noise||_F ; dO (2B,S,S,3): frame0 = 0, frame1 = (eps-noise)/loss
void launch_loss(int dtype, const float* eps, const float* noise, float* sumsq_scratch, float* loss_out, void* dO, int B,
                 int S, cudaStream_t s);
void launch_adam(float* p, const float* g, float* m, float* v, long long n, long long step, const long long* step_dev,
                 double lr, double b1, double b2, double eps, double grad_scale, cudaStream_t s);
void launch_sampler_update(const float* eps2, const float* z, const float* noise, float* z_out, long long n, float w,
                           float c_recip, float c_recipm1, float c1, float c2, float sigma, unsigned long long seed,
                           cudaStream_t s);
void launch_sampler_step_table(const float* eps2, float* z, long long n, float w, const float* tab, const int* pos_dev,
                               const unsigned long long* seed_dev, float* inp_z, float* inp_logsnr, int B2, cudaStream_t s);
void launch_forward_diffusion(const float* x0, const float* noise_in, const int* t_in, unsigned long long seed,
                              const float* sqrt_ac, const float* sqrt_1mac, float p_uncond, float* z, float* noise_out,
                              float* logsnr_out, int* t_out, float* cond_mask_out, int B, long long per, cudaStream_t s);
void launch_dropout_mask(float* out, long long n, int op_index, unsigned long long seed, float rate, cudaStream_t s);

// ---- attention over frames ---------------------------------------------------------------------------------
struct AttnArgs {
  const void* qkv; const void* res; void* out; float* lse;
  const void* dout; float* dscratch; void* dqkv;   // backward
  int N, L, C, heads, cross;
  int scratch_zeroed;   // backward (head_dim <= 32): dscratch is all-zero on entry -> the fused kernel also forms D and rounds dQ
                        // itself (no prep / store kernels) and leaves dscratch all-zero again
  float* cstats;   // optional (tcgen05 forward): per-(sample, channel) [sum, sumsq] of the stored output, (N/2, C, 2) fp32, accumulated
};
void launch_attn_fwd_simt(int dtype, const AttnArgs& a, cudaStream_t s);
void launch_attn_bwd_simt(int dtype, const AttnArgs& a, cudaStream_t s);
// tcgen05/TMEM/TMA versions (bf16; L % 128 == 0; head_dim 16/32/64/128)
bool attn_tc_supported(int dtype, int L, int C, int heads);
void launch_attn_fwd_tc(const AttnArgs& a, cudaStream_t s);
void launch_attn_bwd_tc(const AttnArgs& a, cudaStream_t s);

const char* xu_kernel_error();  // last launch-configuration error recorded by a launcher ("" if none)
void xu_set_kernel_error(const char* msg);
