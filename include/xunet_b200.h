/* xunet_b200.h -- C-ABI of libxunet_b200.so: the B200 (sm_100a) X-UNet hot path.
 *
 * The reference (shiveshkhaitan/novel_view_synthesis_3d) has no FFI/plugin layer: its boundary is the
 * Flax functional API used by train.py / sampling.py.  Each entry point below names the reference
 * call it stands in for (file:line relative to the reference tree).  Conventions:
 *   - every pointer argument documented "device" is a CUDA device pointer owned by the caller
 *     (the PyTorch host allocates; the library allocates nothing after xunet_create);
 *   - `stream` is a cudaStream_t passed as void*; all launches are asynchronous on it and
 *     graph-capturable (no syncs, no allocations inside forward/backward/adam/sampler calls);
 *   - return 0 on success, non-zero on error; xunet_last_error() gives the message; nothing throws;
 *   - a handle is bound to the device current at creation and is NOT thread-safe.
 *   - inputs are float32 (JAX down-casts the float64 numpy batch with x64 off, train.py:132-140).
 */
#ifndef XUNET_B200_H
#define XUNET_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define XUNET_MAX_LEVELS 8
#define XUNET_DTYPE_F32 0   /* fp32 activations, fp32 SIMT math: the <=1e-3 parity ("verify") mode      */
#define XUNET_DTYPE_BF16 1  /* bf16 activations, fp32 accumulate / statistics / master weights           */

#define XUNET_RAYS_V3D130_IJ 0  /* visu3d<=1.3: K^-1 applied to (row+.5, col+.5)  (requirements.txt:17) */
#define XUNET_RAYS_OPENCV_UV 1  /* visu3d>=1.4: K^-1 applied to (col+.5, row+.5)                        */

/* XUNet module attributes, model/xunet.py:205-215 (ch_mult / attn_resolutions are class attributes there). */
typedef struct xunet_config {
  int ch;
  int n_levels;
  int ch_mult[XUNET_MAX_LEVELS];
  int emb_ch;
  int num_res_blocks;
  int n_attn_resolutions;
  int attn_resolutions[XUNET_MAX_LEVELS];
  int attn_heads;
  float dropout;
  int use_pos_emb;
  int use_ref_pose_emb;
  int ray_convention;
} xunet_config;

typedef struct xunet_handle xunet_handle;

/* Device pointers of one batch: the dict train.py:53-60 / sampling.py:156-165 builds.
 * x, z: (B,S,S,3); logsnr: (B); R1,R2,K: (B,3,3) row-major; t1,t2: (B,3); cond_mask: (B) 0/1 floats
 * (model/xunet.py:176-179).  R,t are cam->world (dataset/data_loader.py:105-108). */
typedef struct xunet_batch {
  const float* x;
  const float* z;
  const float* logsnr;
  const float* R1;
  const float* t1;
  const float* R2;
  const float* t2;
  const float* K;
  const float* cond_mask;
  /* OPTIONAL (NULL = compute rays from R,t,K with cfg.ray_convention): precomputed camera rays of both cameras,
   * (B, 2, S, S, 6) fp32 = [pos xyz | dir xyz] per pixel, camera 0 = (R1,t1), camera 1 = (R2,t2) -- exactly what
   * v3d.Camera(spec, world_from_cam).rays() returns at model/xunet.py:159-161,166-168 (.pos, .dir).  With this set the
   * network never depends on the library's restatement of visu3d's pixel convention; R,t,K are then unused. */
  const float* rays;
} xunet_batch;

const char* xunet_last_error(void);
int xunet_version(void);

/* XUNet() + shape inference of .init (train.py:39-43): builds the static execution plan for
 * (config, batch, side, dtype).  training!=0 reserves gradient storage in the workspace. */
int xunet_create(const xunet_config* cfg, int batch, int side, int dtype, int training, xunet_handle** out);
void xunet_destroy(xunet_handle* h);

/* Parameter tree: leaves in Flax naming ("XUNetBlock_3/ResnetBlock_0/Conv_0/kernel"),
 * Flax shapes and C-order element layout, concatenated into ONE flat fp32 buffer at `offset` (elements). */
long long xunet_param_count(const xunet_handle* h);
int xunet_param_leaves(const xunet_handle* h);
int xunet_param_leaf(const xunet_handle* h, int i, const char** name, int* ndim, long long shape[5],
                     long long* offset);

long long xunet_workspace_bytes(const xunet_handle* h);

/* Named activations kept in the workspace after a forward (for parity tests): dims = (N=2B, H, W, C),
 * offset in bytes; *is_f32 != 0 -> fp32 elements, else elements of the handle dtype.  After a backward,
 * *grad_byte_offset (>=0) locates the gradient of the same tensor (-1 if it has none). */
int xunet_tap_count(const xunet_handle* h);
int xunet_tap(const xunet_handle* h, int i, const char** name, int dims[4], long long* byte_offset, int* is_f32,
              long long* grad_byte_offset);

/* XUNet.apply({'params': p}, batch, cond_mask=, train=, rngs=)  (train.py:63-66, sampling.py:131-132).
 * params: device, flat fp32.  seed_dev: device pointer to ONE uint64 dropout seed (read at execution
 * time, so a captured graph sees updates); may be NULL when train==0.  eps_out: device (B,S,S,3) fp32. */
int xunet_forward(xunet_handle* h, const float* params, const xunet_batch* batch, int train,
                  const unsigned long long* seed_dev, void* workspace, float* eps_out, void* stream);

/* Sampler fast path: after one xunet_forward, declare poses / K / cond_mask / params unchanged; subsequent forwards skip the
 * ray + posenc kernel, the pose-embedding convs and the bf16 weight-shadow conversion (the reference recomputes all of it
 * 2000 times per view, sampling.py:131-132).  on=0 restores the full forward. */
int xunet_set_static_conditioning(xunet_handle* h, int on);

/* The value_and_grad half of apply_model (train.py:62-71): must follow xunet_forward on the same
 * workspace.  loss = ||eps_hat - noise||_F (train.py:67).  grads: device flat fp32 (param layout),
 * overwritten.  loss_out: device, 1 float. */
int xunet_backward(xunet_handle* h, const float* params, const xunet_batch* batch, const float* noise,
                   const unsigned long long* seed_dev, void* workspace, float* grads, float* loss_out,
                   void* stream);

/* Data-parallel hook (north_star: "one NCCL allreduce of the gradient bucket"; the reference's pmap step would place
 * lax.pmean between value_and_grad and apply_gradients, train.py:70-76).  With a callback installed, xunet_backward calls
 *     fn(user, elem_offset, n_elems)
 * on the HOST, in descending offset order, each time grads[elem_offset, elem_offset+n_elems) is final: every kernel that
 * writes it has been enqueued and ordered before the current point of `stream` (the library's internal side stream is
 * joined first).  The callback typically enqueues an all-reduce of that range on a communication stream ordered after
 * `stream`, so NVLink traffic overlaps the rest of the backward.  The ranges partition the whole flat buffer; buckets
 * are cut at residual-block boundaries once they reach min_bucket_bytes (the last one, offset 0, takes the remainder).
 * fn == NULL removes the hook.  Graph capture: the callback runs at capture time, so what it enqueues is captured too. */
typedef void (*xunet_bucket_fn)(void* user, long long elem_offset, long long n_elems);
int xunet_set_grad_bucket_callback(xunet_handle* h, xunet_bucket_fn fn, void* user, long long min_bucket_bytes);
int xunet_grad_bucket_count(const xunet_handle* h);

/* Number of kernel launches (graph kernel nodes) one xunet_forward / xunet_backward issues for this plan; counted by
 * capturing the call into a throw-away CUDA graph (nothing executes).  Arguments as for forward/backward. */
int xunet_count_kernels(xunet_handle* h, const float* params, const xunet_batch* batch, const float* noise,
                        const unsigned long long* seed_dev, void* workspace, float* grads, float* loss_out,
                        int* n_forward, int* n_backward);

/* update_model / TrainState.apply_gradients with optax.adam (train.py:45,74-76).  `step` is the 1-based
 * step count after increment; if step_dev != NULL the kernel reads *step_dev (device int64) instead.
 * grad_scale multiplies the gradient first (1/world_size after the all-reduce). */
int xunet_adam_step(float* params, const float* grads, float* m, float* v, long long n, long long step,
                    const long long* step_dev, double lr, double b1, double b2, double eps, double grad_scale,
                    void* stream);

/* One ancestral update of sampling.py:128-151 given the 2B-batched model output eps2 =
 * [eps_cond (B,S,S,3); eps_uncond (B,S,S,3)]:  eps=(1+w)eps_c-w eps_u; x0=clip(c_recip z - c_recipm1 eps);
 * z' = c1 x0 + c2 z + sigma * noise.  noise==NULL -> device philox-free hash normal from seed.  z_out may alias z. */
int xunet_sampler_update(const float* eps2, const float* z, const float* noise, float* z_out, long long n_per_half,
                         float w, float c_recip, float c_recipm1, float c1, float c2, float sigma,
                         unsigned long long seed, void* stream);

/* The same step with the schedule on the device, for a per-step CUDA graph (forward + this call + counter increment) that is
 * replayed with no host arithmetic in between (the reference does the schedule math in numpy between two un-jitted model
 * calls, sampling.py:133-151).  table: device (steps, 8) floats, row k (k = 0: first executed step = highest t) =
 * {sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2, sigma (0 at t=0),
 * log-SNR fed to the NEXT forward (sampling.py:151), 0, 0}; pos_dev: device int32 loop position (the caller increments it
 * after the call, in stream order).  z (n = B*S*S*3) is updated in place; the new z is also written to both halves of the
 * next forward's [cond ; uncond] input next_z2 (2n) and next_logsnr2 (batch2 = 2B entries).  Noise: hash normal of
 * (*seed_dev - k), i.e. the caller stores seed(first step) and the per-step seeds count down like the timestep index;
 * seed_dev is a device uint64 so a captured graph can be re-seeded. */
int xunet_sampler_step_table(const float* eps2, float* z, long long n, float w, const float* table, const int* pos_dev,
                             const unsigned long long* seed_dev, float* next_z2, float* next_logsnr2, int batch2, void* stream);

/* Device-side forward diffusion = the per-item work of SceneInstanceDataset.__getitem__ (dataset/data_loader.py:92-110):
 *   t ~ U{0..999};  noise ~ N(0,1);  z = sqrt_ac[t] * x0 + sqrt_1mac[t] * noise;  logsnr = logsnr_schedule_cosine(t/1000);
 *   cond_mask = (u > p_uncond)  (train.py:64).
 * sqrt_ac / sqrt_1mac: device tables of 1000 floats (cosine-beta schedule, data_loader.py:15-25,70-74).  t_in / noise_in may
 * be NULL (drawn on the device from `seed`) or given (parity tests).  All outputs are device buffers; x0 (B, per). */
int xunet_forward_diffusion(const float* x0, const float* noise_in, const int* t_in, unsigned long long seed,
                            const float* sqrt_ac, const float* sqrt_1mac, float p_uncond, float* z, float* noise_out,
                            float* logsnr_out, int* t_out, float* cond_mask_out, int B, long long per, void* stream);

/* The dropout keep-mask the kernels use for residual-block `op_index` (0/1 floats, device), so tests can
 * hand the identical mask to the oracle (nn.Dropout, model/xunet.py:84). */
int xunet_dropout_mask(float* mask_out, long long n, int op_index, unsigned long long seed, float rate,
                       void* stream);

/* ---- operator-level entry points (parity tests / micro-benchmarks of single kernels) ----
 * dtype as above; tensors channels-last (N,H,W,C). */
int xunet_op_conv(int dtype, int impl, const void* x, const float* w, const float* bias, const void* res, void* y,
                  int N, int Hi, int Wi, int Ci, int Co, int ksize, int stride, int nseg, float alpha, void* stream);
int xunet_op_conv_dgrad(int dtype, int impl, const void* dy, const float* w, void* dx, int N, int Hi, int Wi, int Ci,
                        int Co, int ksize, int stride, int nseg, float alpha, int accumulate, void* stream);
int xunet_op_conv_wgrad(int dtype, int impl, const void* x, const void* dy, float* dw, float* dbias, int N, int Hi, int Wi,
                        int Ci, int Co, int ksize, int stride, int nseg, float alpha, void* stream);
/* qkv: (N, L, 3C) [q | k | v], heads x hd; res/out: (N, L, C); lse: (N, heads, L) fp32. out=(attn+res)/sqrt2 */
int xunet_op_attention(int dtype, int impl, const void* qkv, const void* res, void* out, float* lse, int N, int L, int C,
                       int heads, int cross, void* stream);
/* dscratch: N*L*(heads + C) floats of scratch (row dots D, and the fp32 dQ accumulator of the fused head_dim<=32 path).
 * With XUNET_OP_ATTN_FOLD=1 in the environment the caller promises an all-zero dscratch and gets it back all-zero: the
 * head_dim<=32 path then runs as ONE kernel (D and the dQ rounding inside it), which is how the engine runs it. */
int xunet_op_attention_bwd(int dtype, int impl, const void* qkv, const void* res, const void* out, const void* dout,
                           const float* lse, float* dscratch, void* dqkv, int N, int L, int C, int heads, int cross,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif
