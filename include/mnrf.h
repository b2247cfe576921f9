/* mnrf.h -- C ABI of libmnrf.so, the B200 (sm_100a) renderer for Mirror-NeRF's render_rays hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain pointers and sizes (no torch types).
 * Unless a parameter is documented as HOST, every pointer is a DEVICE pointer to fp32 data on the
 * current CUDA device and `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * All functions return 0 on success, non-zero on error; mnrf_last_error() gives the message.
 * There is no CPU fallback anywhere: without a CUDA device every compute call fails.
 *
 * Reference interfaces replaced (R/ = zju3dv/Mirror-NeRF @ fc0d7911):
 *   mnrf_field_*            R/models/mirror_nerf.py:41-99   (MirrorNeRF.__init__ : parameter set/layout)
 *   mnrf_field_eval_*       R/models/mirror_nerf.py:101-212 (MirrorNeRF.forward + Embedding :6-38)
 *   mnrf_coarse_z           R/models/rendering.py:271-300   (stratified coarse depths)
 *   mnrf_composite          R/models/rendering.py:175-264   (inference(): quadrature / compositing)
 *   mnrf_sample_pdf         R/models/rendering.py:7-51,312-326 (inverse-CDF resampling + sort-merge)
 *   mnrf_searchsorted_right R/models/rendering.py:33        (torch.searchsorted(cdf,u,right=True))
 *   mnrf_render_level       R/models/rendering.py:54-369    (render_rays, one level)
 *   mnrf_render_level_host  same, HOST buffers in/out (what eval.py:1122-1138,735-736 does around it)
 *   mnrf_reflect_rays / mnrf_compact_rows / mnrf_blend_reflection
 *                           R/eval.py:295-320,515-548,676-697 and R/train.py:153-296 (Whitted bounce, one step at a time)
 *   mnrf_render_recursive   R/eval.py:114-740 batched_inference (the whole recursion incl. the roughness cone, device-side)
 *   mnrf_train_pass_fwd/bwd implicit torch.autograd of R/models/rendering.py:87-266 + mirror_nerf.py:101-212 (training)
 *   mnrf_adam_step / mnrf_peer_allreduce_adam
 *                           R/utils/__init__.py:47-58 (torch.optim.Adam) + PL DDP gradient all-reduce (R/train.py:582)
 *   mnrf_hash_field_create  R/models/mirror_nerf_tcnn.py:13-259 (MirrorNeRFTcnn: hash-grid field; its gradients come from
 *                           mnrf_train_pass_fwd/bwd with a hash-grid handle: tinycudann's grid backward + torch.autograd)
 */
#ifndef MNRF_H_
#define MNRF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNRF_ABI_VERSION 2

/* ---- field (MirrorNeRF MLP) ------------------------------------------------------------------ */

/* Parameter tensors in MirrorNeRF.state_dict() order, each fp32, [out,in] row-major as nn.Linear stores
 * them (R/models/mirror_nerf.py:60-99; key names in SURVEY.md section 5).  Index -> key:
 *   2*i, 2*i+1   xyz_encoding_{i+1}.0.weight / .bias        i = 0..7   (256x63, 256x256 x3, 256x319, 256x256 x3)
 *   16,17        xyz_encoding_final.weight / .bias          (256x256)
 *   18,19        dir_encoding.0.weight / .bias              (128x283)
 *   20,21        sigma.weight / .bias                       (1x256)
 *   22,23        rgb.0.weight / .bias                       (3x128)
 *   24..27       normal_net.0.weight/.bias, normal_net.1.weight/.bias   (128x256, 3x128)  or all NULL
 *   28..31       is_mirror_net.0.weight/.bias, is_mirror_net.2.weight/.bias (128x256, 1x128) or all NULL
 * Only the reference architecture D=8, W=256, skips=[4], N_emb_xyz=10, N_emb_dir=4 is supported. */
#define MNRF_NUM_PARAM_TENSORS 32

typedef struct mnrf_field mnrf_field; /* opaque: device-resident packed weights */

/* Pack the 32 tensors (HOST array of DEVICE pointers) into the kernel formats.  Asynchronous on stream. */
int mnrf_field_create(mnrf_field** out, const float* const* tensors, void* stream);
/* Re-pack after the parameters changed (optimizer step).  Same tensor set (heads may not appear/disappear).  For a hash-grid
 * field `tensors` are the 12 pointers of mnrf_hash_field_create (same table size). */
int mnrf_field_update(mnrf_field* f, const float* const* tensors, void* stream);
void mnrf_field_destroy(mnrf_field* f);
int mnrf_field_has_normal(const mnrf_field* f);
int mnrf_field_has_mirror(const mnrf_field* f);

/* Hash-grid field (BASELINE config 3; R/models/mirror_nerf_tcnn.py:13-259 with the arguments of R/train.py:73-100): 16-level
 * multiresolution hash encoding + small bias-free MLPs.  tensors = HOST array of 12 DEVICE pointers:
 *   0 encoder.params (table_floats) | 1,2 sigma_net.{0,1}.weight (64x32, 16x64) | 3,4,5 color_net.{0,1,2}.weight (64x31, 64x64,
 *   3x64) | 6,7 normal_net.{0,1}.weight (64x15, 3x64) or NULL | 8..11 is_mirror_net.0.{weight,bias} (32x15, 32),
 *   is_mirror_net.2.{weight,bias} (1x32, 1) or NULL.
 * The level table (HOST arrays of 16: grid scale, resolution, first table entry, entries) is computed by the caller so that it
 * is defined in one place (mirror_nerf_b200/mirror_nerf_tcnn.py::level_table, following tinycudann's grid.h).
 * The returned object is used wherever a mnrf_field is accepted (mnrf_field_eval_*, mnrf_render_level*), including analytic
 * normals (compute_normal) and the training pass (mnrf_train_pass_fwd/bwd).  Destroy with mnrf_field_destroy. */
int mnrf_hash_field_create(mnrf_field** out, const float* const* tensors, int64_t table_floats, float bound,
                           const float* level_scale, const int* level_res, const uint32_t* level_offset,
                           const uint32_t* level_size, void* stream);

/* which kernel evaluates the MLP */
#define MNRF_IMPL_TC3 3   /* tcgen05, fp16 hi/lo split operands, 3 MMAs per product (fp32-grade; parity mode) */
#define MNRF_IMPL_TC2 2   /* tcgen05, 1 fp16 pass + the two cross terms as e4m3 passes at twice the rate (2 pass-equivalents;
                             operand error ~2^-16: inside the 1e-3 parity bar on scene-like fields)      */
#define MNRF_IMPL_TC1 1   /* tcgen05, single fp16 pass (speed mode; does not meet the 1e-3 parity bar)      */
#define MNRF_IMPL_FP32 0  /* CUDA-core fp32 verification kernel (slow; the only one that exports geo_feat) */

/* raw per-point record written by the field kernels: 8 floats */
#define MNRF_RAW_STRIDE 8 /* [sigma, r, g, b, is_mirror, pred_nx, pred_ny, pred_nz] */

/* Evaluate the field at the samples of a ray batch: point (r,s) = o_r + d_r * z[r,s]  (rendering.py:302,326).
 *   rays (n_rays,8) = [o,d,near,far];  z (n_rays,S).
 *   sigma_only != 0: only `sigma_out` (n_rays*S) is written (coarse pass at test time, rendering.py:139-150).
 *   otherwise `raw` (n_rays*S, 8) is written; absent heads give 0.
 *   normal_out: optional (n_rays*S,3) analytic normal  normalize(-d sigma/d xyz) (mirror_nerf.py:136-146): the tensor-core
 *               kernels run the reverse chain through the trunk as 9 more GEMM steps with transposed weights. */
int mnrf_field_eval_rays(const mnrf_field* f, int impl, const float* rays, const float* z, int n_rays, int S,
                         int sigma_only, float* raw, float* sigma_out, float* normal_out, void* stream);

/* MirrorNeRF.forward on a flat batch (mirror_nerf.py:101-187): x (B, 3+27) = [xyz | embedded dir]
 * (or (B,3) when sigma_only).  Outputs (any may be NULL): sigma (B), rgb (B,3), is_mirror (B),
 * pred_normal (B,3, l2-normalised), normal (B,3, analytic; FP32 impl only), geo_feat (B,256). */
int mnrf_field_eval_points(const mnrf_field* f, int impl, const float* x, int B, int sigma_only,
                           float* sigma, float* rgb, float* is_mirror, float* pred_normal, float* normal,
                           float* geo_feat, void* stream);

/* Embedding.forward (mirror_nerf.py:21-38): out (n, 3+6*n_freqs) = [x, sin(2^k x), cos(2^k x) ...]. */
int mnrf_embed(const float* x, int n, int n_freqs, float* out, void* stream);

/* Rays of one pinhole view, generated on the device (R/datasets/ray_utils.py:6-53, R/datasets/blender.py:158-168):
 * c2w_host = 12 HOST floats (3x4 row-major camera-to-world); rays (H*W, 8) = [o, normalised d, near, far]. */
int mnrf_generate_rays(int H, int W, float focal, const float* c2w_host, float near, float far, float* rays,
                       void* stream);

/* ---- sampler ----------------------------------------------------------------------------------- */

/* z_steps: the S values of torch.linspace(0,1,S) (taken from the host library so they are bit-identical,
 * SURVEY.md appendix A).  perturb_u: (n,S) uniform draws or NULL when perturb == 0.  z_out (n,S).
 * Arithmetic is done with separately rounded fp32 mul/add so it is bit-exact against the CPU reference. */
int mnrf_coarse_z(const float* rays, int n, const float* z_steps, int S, int use_disp, float perturb,
                  const float* perturb_u, float* z_out, void* stream);

/* inds[i,j] = #{k : cdf[i,k] <= u[i,j]}  (int64, like torch).  u_stride = 0 broadcasts one row of u. */
int mnrf_searchsorted_right(const float* cdf, int n, int n_cdf, const float* u, int n_u, int u_stride,
                            int64_t* inds, void* stream);

/* sample_fine_points (rendering.py:312-326): bins = mid-points of z_coarse, pdf from weights[:,1:-1]+1e-5,
 * inverse CDF at u, then sort(cat(z_coarse, samples)).  u: (n_imp) if u_stride==0 else (n,n_imp).
 * z_fine (n, S+n_imp).  Optional: samples (n,n_imp), inds (n,n_imp) int64, cdf (n,S-1).  S <= 256, S+n_imp <= 512. */
int mnrf_sample_pdf(const float* z_coarse, const float* weights, int n, int S, int n_imp, const float* u,
                    int u_stride, float* z_fine, float* samples, int64_t* inds, float* cdf, void* stream);

/* sample_pdf(bins, weights, N_importance) exactly as rendering.py:7-51 takes it: bins (n, n_w+1), weights (n, n_w).
 * samples (n,n_imp); optional inds (n,n_imp) int64 and cdf (n,n_w+1). */
int mnrf_sample_pdf_bins(const float* bins, const float* weights, int n, int n_w, int n_imp, const float* u,
                         int u_stride, float* samples, int64_t* inds, float* cdf, void* stream);

/* ---- compositor -------------------------------------------------------------------------------- */

typedef struct mnrf_composite_out {
  float* weights;        /* (n,S)    */
  float* opacity;        /* (n)      */
  float* rgb;            /* (n,3)    or NULL */
  float* depth;          /* (n)      or NULL */
  float* mirror_mask;    /* (n)      or NULL */
  float* pred_normal;    /* (n,S,3)  or NULL : per-sample predicted normals (copied out of raw)      */
  float* surface_normal; /* (n,3)    or NULL : sum_s w * pred_normal (NOT re-normalised)              */
  float* surface_normal_grad; /* (n,3) or NULL : sum_s w * analytic normal                            */
  float* normal_dif;     /* (n)      or NULL : sum_s w * |n - n_pred|^2                               */
  float* x_surface;      /* (n,3)    or NULL : o + d * depth (rendering.py:363-367)                   */
} mnrf_composite_out;

/* sigma (n,S) with stride `sigma_stride` floats between consecutive samples (1 for a sigma array, 8 for raw);
 * raw (n,S,8) or NULL when only weights are wanted; normal (n,S,3) analytic normals or NULL;
 * noise (n,S) standard-normal draws or NULL (noise_std is then ignored).  */
int mnrf_composite(const float* rays, const float* z, const float* sigma, int sigma_stride, const float* raw,
                   const float* normal, const float* noise, float noise_std, int n, int S, int white_back,
                   const mnrf_composite_out* out, void* stream);

/* ---- one render level (render_rays) -------------------------------------------------------------- */

typedef struct mnrf_level_cfg {
  int n_samples;     /* N_samples                          */
  int n_importance;  /* N_importance (0 = coarse only)     */
  int use_disp;
  float perturb;
  float noise_std;
  int white_back;
  int test_time;     /* coarse pass sigma-only when a fine field exists (rendering.py:139,208)          */
  int compute_normal;/* analytic normals normalize(-d sigma/d xyz) for every pass that is not sigma-only         */
  int rerun_coarse_on_fine; /* only_one_field after only_one_field_fine_epoch (rendering.py:328-348)    */
  int impl;          /* MNRF_IMPL_*                                                                     */
  /* ---- ABI 2 (zero = the behaviour of ABI 1) ---- */
  float early_termination_eps; /* > 0: a ray stops once its transmittance prod(1 - alpha + 1e-10) falls below this; honoured by
                                  the fused compositor and only when no per-sample output (weights, pred_normal) is requested:
                                  every skipped sample has weight < eps (rendering.py:194-205)                                */
  int no_fused_composite;      /* 1: field kernel -> raw point records -> separate compositor (the ABI-1 launch sequence)      */
  const float* dir_source;     /* optional (n,8) rows whose columns 3..5 replace rays_d in the direction embedding only
                                  (rendering.py:276 `view_dir`); MLP field                                                     */
  unsigned long long* stats;   /* optional device counters of the fused fine pass: [0] += tiles executed, [1] += 32-sample
                                  chunks skipped by early termination                                                          */
} mnrf_level_cfg;

typedef struct mnrf_level_rng { /* explicit draws (device); NULL -> deterministic (perturb=0 / noise 0) */
  const float* perturb_u;    /* (n, n_samples)              */
  const float* noise_coarse; /* (n, n_samples)              */
  const float* u_pdf;        /* (n, n_importance)           */
  const float* noise_fine;   /* (n, n_samples+n_importance) */
} mnrf_level_rng;

typedef struct mnrf_level_out {
  float* z_coarse;             /* (n,Sc) required */
  mnrf_composite_out coarse;   /* weights+opacity required */
  float* normal_coarse;        /* (n,Sc,3) analytic normals or NULL */
  float* z_fine;               /* (n,Sc+Ni) required when a second pass runs; fine.weights may be NULL when that pass composites
                                  inside the field kernel (MLP field, tensor-core impl, no analytic normals, no sigma noise)  */
  mnrf_composite_out fine;
  float* normal_fine;          /* (n,Sc+Ni,3) or NULL */
} mnrf_level_out;

/* bytes of scratch mnrf_render_level needs for n rays: upper bound for any field pair ... */
int64_t mnrf_level_workspace_bytes(int n, const mnrf_level_cfg* cfg);
/* ... and the exact amount for these fields (a pass that composites inside the field kernel needs no per-point records;
 * with_sigma_noise != 0: the call will pass noise tensors, which selects the unfused compositor when noise_std != 0) */
int64_t mnrf_level_workspace_bytes_for(const mnrf_field* coarse, const mnrf_field* fine, int n, const mnrf_level_cfg* cfg,
                                       int with_sigma_noise);

/* fine may be NULL (coarse only, or only_one_field).  z_steps (n_samples), u_det (n_importance) are the
 * torch.linspace tables (device).  All launches go to `stream`; nothing synchronises. */
int mnrf_render_level(const mnrf_field* coarse, const mnrf_field* fine, const float* rays, int n,
                      const mnrf_level_cfg* cfg, const mnrf_level_rng* rng, const float* z_steps,
                      const float* u_det, void* workspace, int64_t workspace_bytes, const mnrf_level_out* out,
                      void* stream);

/* Same level with HOST buffers: rays_host (n,8) in; compact per-ray results out (each may be NULL):
 * rgb (n,3), depth (n), opacity (n), mirror_mask (n), surface_normal (n,3), x_surface (n,3) of the last pass.
 * Copies H2D/D2H on `stream` and synchronises it before returning.  z_steps/u_det are HOST here. */
int mnrf_render_level_host(const mnrf_field* coarse, const mnrf_field* fine, const float* rays_host, int n,
                           const mnrf_level_cfg* cfg, const float* z_steps_host, const float* u_det_host,
                           float* rgb, float* depth, float* opacity, float* mirror_mask, float* surface_normal,
                           float* x_surface, void* stream);

/* ---- training: one pass with saved activations and its backward -------------------
 * Replaces what torch.autograd records for R/models/rendering.py:87-266 (inference(): model call + quadrature) when
 * train.py:129-145 calls render_rays with gradients enabled: forward of one pass (field at the samples of a ray batch
 * -> compositor) that keeps every activation the backward needs in a caller-provided workspace, and the backward
 * that turns the gradients of the pass outputs into gradients of the 32 parameter tensors -- including the second-order
 * path through the analytic normal  n = normalize(-d sigma/d xyz)  (mirror_nerf.py:136-146, create_graph=True in
 * utils/func.py:10-25).  The GEMMs run on the tensor cores (tcgen05 kind::tf32, 3x split = fp32-grade; train_tc.cu).  Optionally also the gradient w.r.t. the rays'
 * origin and direction (train.py:194-243 builds secondary rays from x_surface / normals without detaching).  z carries no
 * gradient (z_fine is detached by the reference, rendering.py:335,353; near/far are constants). */
typedef struct mnrf_train_cfg {
  int S;              /* samples per ray of this pass */
  int compute_normal; /* analytic normals (and their double backward) */
  int white_back;
  float noise_std;    /* multiplies `noise` (ignored when noise == NULL) */
  int detach_density_for_mask_loss;   /* mirror_nerf.py:169-170, rendering.py:222-226 */
  int detach_density_for_normal_loss; /* mirror_nerf.py:158, rendering.py:245-247 */
} mnrf_train_cfg;

/* gradients of the pass outputs (device, any may be NULL = zero) */
typedef struct mnrf_train_grads {
  const float* rgb;                 /* (n,3) */
  const float* depth;               /* (n)   */
  const float* opacity;             /* (n)   */
  const float* mirror_mask;         /* (n)   */
  const float* surface_normal;      /* (n,3) */
  const float* surface_normal_grad; /* (n,3) */
  const float* normal_dif;          /* (n)   */
  const float* x_surface;           /* (n,3) */
  const float* weights;             /* (n,S) */
  const float* pred_normal;         /* (n,S,3) */
  const float* normal;              /* (n,S,3) */
} mnrf_train_grads;

int64_t mnrf_train_fwd_workspace_bytes(int n, int S, int compute_normal); /* MLP field */
int64_t mnrf_train_bwd_workspace_bytes(int n, int S, int compute_normal); /* MLP field */
/* the same for any field handle (the hash-grid field recomputes its forward in the backward and needs 44 / 80 bytes per point) */
int64_t mnrf_field_train_fwd_workspace_bytes(const mnrf_field* f, int n, int S, int compute_normal);
int64_t mnrf_field_train_bwd_workspace_bytes(const mnrf_field* f, int n, int S, int compute_normal);

/* Forward: rays (n,8), z (n,S), noise (n,S) or NULL.  Writes the compositor outputs (`out`, same meaning as
 * mnrf_composite) and, when cfg->compute_normal, the per-sample analytic normals normal_out (n,S,3).
 * ws: mnrf_train_fwd_workspace_bytes(n,S,compute_normal) bytes, must stay untouched until the backward has run. */
int mnrf_train_pass_fwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                        const mnrf_train_cfg* cfg, void* ws, int64_t ws_bytes, const mnrf_composite_out* out,
                        float* normal_out, void* stream);

/* Backward.  grad_tensors: HOST array of 32 DEVICE pointers in mnrf_field_create order, each the size of its
 * parameter ([out,in] layout), ACCUMULATED into (zero them first for plain gradients); entries of absent heads NULL.
 * For a hash-grid field: 12 pointers in mnrf_hash_field_create order (entry 0 = gradient of encoder.params, table_floats
 * elements, scatter-added with atomics as tinycudann's grid backward does; R/models/mirror_nerf_tcnn.py:151-259 under
 * torch.autograd, incl. the double backward through the analytic normal of :170-178).
 * ray_detach_mirror: optional (n) floats, != 0 marks rays whose density is detached from the mirror-mask loss
 * (detach_density_outside_mirror_for_mask_loss, mirror_nerf.py:171-183 / rendering.py:227-238: rays outside the
 * ground-truth mirror).  grad_rays: optional (n,8) OVERWRITTEN with dL/d[o, d, near, far] (near/far columns zero); it needs
 * `depth` (n), the forward's depth output, when grads->x_surface is set.
 * Accumulation uses atomics: results are not bit-reproducible run to run. */
int mnrf_train_pass_bwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                        const mnrf_train_cfg* cfg, const void* ws_fwd, int64_t ws_fwd_bytes, void* ws_bwd,
                        int64_t ws_bwd_bytes, const mnrf_train_grads* grads, const float* ray_detach_mirror,
                        float* const* grad_tensors, const float* depth, float* grad_rays, void* stream);

/* GEMM engine of the training path: 1 (default, parity mode) = tcgen05 tf32 3x-split kernels (train_tc.cu, fp32-grade products);
 * 0 = fp32 CUDA-core kernels (verification twin); 2 = tcgen05 single tf32 pass (speed mode: 10-bit mantissa operands, what
 * torch.backends.cuda.matmul.allow_tf32 = True would do -- NOT the reference's fp32 arithmetic).  The environment variable
 * MNRF_TRAIN_GEMM = simt | tf32 selects 0 | 2 at first use. */
int mnrf_train_set_gemm(int engine);

/* Bring-up aid: milliseconds per launch of one training GEMM on synthetic operands (kind 0 = layer GEMM `step` over P rows,
 * kind 1 = 256x256 weight-gradient GEMM over P rows; engine as above; dbg = train_tc.cu debug bits, 0 for a real run). */
int mnrf_debug_gemm_bench(const mnrf_field* f, int kind, int step, int P, int engine, int dbg, int iters, float* ms_out);

/* One torch.optim.Adam step (the reference's optimizer: R/utils/__init__.py:47-58, lr 5e-4, eps 1e-8, L2 weight decay) on flat
 * fp32 buffers of n elements: g = grads*grad_scale + weight_decay*p; m,v updated in place; p -= lr/(1-b1^t) * m/(sqrt(v)/
 * sqrt(1-b2^t) + eps).  `step` is t (1-based).  grad_scale = 1/world_size turns an all-reduce SUM into DDP's average. */
int mnrf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* The same data-parallel step as all-reduce + mnrf_adam_step, as ONE kernel over NVLink / NVSwitch peer memory: grad_ptrs /
 * param_ptrs = HOST arrays of `world` (<= 8) device pointers to every rank's flat gradient / parameter buffer, peer-mapped
 * into this process (e.g. torch.distributed._symmetric_memory buffer_ptrs).  Rank r owns the shard mnrf_peer_shard() returns:
 * it sums that shard of all ranks' gradients (P2P loads), divides by world, applies Adam with its OWN moment shards
 * (exp_avg_shard / exp_avg_sq_shard: hi-lo elements) and stores the new parameters into every rank's buffer (P2P stores).
 * n must be a multiple of 4.  The caller orders ranks with a device-side barrier before (all gradients written) and after
 * (all parameter shards visible).  Replaces PL DDP's all-reduce + torch.optim.Adam (R/train.py:582, R/utils/__init__.py:47-58). */
int mnrf_peer_shard(int64_t n, int world, int rank, int64_t* lo, int64_t* hi);
int mnrf_peer_allreduce_adam(const uint64_t* grad_ptrs, const uint64_t* param_ptrs, int world, int rank, float* exp_avg_shard,
                             float* exp_avg_sq_shard, int64_t n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int step, void* stream);

/* out[i] += alpha * in[i] (fp32, n elements): gradient-buffer plumbing for the data-parallel all-reduce */
int mnrf_axpy(float* out, const float* in, int64_t n, float alpha, void* stream);

/* ---- Whitted bounce helpers (callers of render_rays: eval.py:295-697, train.py:153-296) ---------- */

/* mask (n) is thresholded IN PLACE (>0.5 -> 1, <0.5 -> 0, exactly 0.5 kept; eval.py:305-306).
 * secondary (n,8) = [x_surface, 2(n.w)n - w, near_secondary, far_parent] with n = l2norm(normal),
 * w = l2norm(-d) (eval.py:515-540).  reflect_dir (n,3) optional.  any_mirror: device int, set to 1 if any mask==1. */
int mnrf_reflect_rays(const float* rays, const float* x_surface, const float* normal, float* mask, int n,
                      float near_secondary, float* secondary, float* reflect_dir, int* any_mirror, void* stream);

/* Stable compaction (same order as boolean-mask indexing): out (count,row_floats) = in[mask != 0].
 * index (n) int32 optional: destination row or -1.  count: device int.  */
int mnrf_compact_rows(const float* in, const float* mask, int n, int row_floats, float* out, int* index,
                      int* count, void* stream);

/* dense[i] = alpha*dense[i] + beta*compact[index[i]] for rows with index[i] >= 0 (index NULL: identity; compact NULL: scale
 * only).  Accumulates the extra jittered reflections of the roughness cone (eval.py:623-674) into the first one. */
int mnrf_axpy_rows(float* dense, const float* compact, const int* index, int n, int c, float alpha, float beta,
                   void* stream);

/* rgb = m*reflect + (1-m)*base (eval.py:680-697).  child_rgb/child_depth are the bounce results, either
 * dense (index == NULL, n rows) or compacted (index[i] = row in child or -1 -> reflect := base).
 * Outputs: rgb_out (n,3), rgb_reflect (n,3) and depth_reflect (n) (zeros where not traced). */
int mnrf_blend_reflection(const float* base_rgb, const float* mask, const float* child_rgb,
                          const float* child_depth, const int* index, int n, float* rgb_out, float* rgb_reflect,
                          float* depth_reflect, void* stream);

/* ---- Whitted recursion on the device (additive entry point; the drop-in render_rays stays single-level) ------------------
 * What the reference's callers do around render_rays at inference time -- R/eval.py::batched_inference:
 *   :132-160 level call, :295-320 mask threshold + trace condition, :336-360 normal, :506-548 jitter / reflect / secondary rays /
 *   compaction, :609-674 recursive call + roughness cone (--app_control_mirror_roughness), :676-723 blend
 * (R/train.py:129-348 has the same structure with `only_trace_rays_in_mirrors` from the hparams) -- as ONE call without any
 * host synchronisation: a level's mirror rays are counted, compacted and re-enqueued on the device, and the next level's
 * kernels are launched for a worst-case row count and read the live count from device memory.  The T+1 jittered reflections
 * of a level's mirror rays are rendered as ONE child batch and averaged per parent ray. */
typedef struct mnrf_trace_cfg {
  mnrf_level_cfg level;            /* per-level render_rays arguments; eval semantics: perturb = noise_std = 0, test_time = 1 */
  int max_recursive_level;         /* 0 = no bounce                                                                           */
  int only_trace_rays_in_mirrors;  /* -1: eval.py (level 0 re-traces ALL rays of a batch that has a mirror pixel, deeper levels
                                      only mirror rays, :159); 1: compact at every level (train.py hparam)                     */
  int trace_ray_times;             /* roughness cone: extra jittered reflections per mirror ray (0 = off)                      */
  float normal_noise_std;          /* normal += N(0, std^2) before reflecting (0 = off)                                        */
  uint64_t noise_seed;             /* Philox seed of the on-device normal noise                                                */
} mnrf_trace_cfg;

typedef struct mnrf_trace_out {    /* level-0 results, DEVICE; rgb, depth, opacity are required, the rest optional (NULL)      */
  float* rgb;               /* (n,3) blended colour  m*reflect + (1-m)*direct                                                  */
  float* rgb_direct;        /* (n,3) level-0 colour before the blend                                                           */
  float* rgb_reflect;       /* (n,3) colour seen along the reflected ray (zeros where nothing was traced)                      */
  float* depth;             /* (n)                                                                                             */
  float* depth_reflect;     /* (n)                                                                                             */
  float* opacity;           /* (n)                                                                                             */
  float* mirror_mask;       /* (n) hard-clipped mask (>0.5 -> 1, <0.5 -> 0)                                                    */
  float* surface_normal;    /* (n,3) composited normal used for the reflection (not normalised)                                */
  float* x_surface;         /* (n,3)                                                                                           */
  float* reflect_direction; /* (n,3) of the first (t = 0) reflection                                                           */
  int* level_rays;          /* (max_recursive_level + 1) ints: rays rendered per level, summed over that level's batches       */
} mnrf_trace_out;

/* Scratch for n primary rays.  `budget_bytes` > 0 caps it: deeper levels are then rendered in slabs of as many rows as fit
 * (more launches, less memory); 0 = everything in one batch per level.  Returns -1 on bad arguments. */
int64_t mnrf_recursive_workspace_bytes(const mnrf_field* coarse, const mnrf_field* fine, int n, const mnrf_trace_cfg* cfg,
                                       int64_t budget_bytes);
/* level0_normal_noise: optional (trace_ray_times+1, n, 3) standard-normal draws for the LEVEL-0 reflections (tests / replay of a
 * torch stream); deeper levels and a NULL pointer use the on-device generator.  z_steps / u_det as in mnrf_render_level. */
int mnrf_render_recursive(const mnrf_field* coarse, const mnrf_field* fine, const float* rays, int n,
                          const mnrf_trace_cfg* cfg, const float* z_steps, const float* u_det,
                          const float* level0_normal_noise, void* workspace, int64_t workspace_bytes,
                          const mnrf_trace_out* out, void* stream);

/* ---- misc ----------------------------------------------------------------------------------------- */
const char* mnrf_last_error(void);
int mnrf_abi_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t mnrf_launch_count(void);
/* Time every launch of the tcgen05 field kernel with CUDA events recorded on its own launch stream.
 * collect() waits for the recorded launches and returns their summed device time, summed ALGORITHMIC flops
 * (2 * points * mnrf_macs_*()) and count, then clears the log. */
int mnrf_profile_enable(int on);
int mnrf_profile_collect(double* total_ms, double* total_flops, int64_t* launches);
/* Bring-up aid: CTA 0 of the tcgen05 field kernel logs (clock64, tag) pairs into buf = uint64[1 + 2*capacity]
 * (buf[0] = event count; zero it first).  NULL disables. */
int mnrf_debug_set_trace(void* buf, int64_t capacity_events);
/* Schedule of the 256-wide layers in the MNRF_IMPL_TC2 kernels: 1 = N-split (two 128-column halves per layer, the first half's
 * epilogue overlaps the second half's MMAs), 0 = one N = 256 accumulation per layer, -1 = library default (environment
 * variable MNRF_TC_SPLIT, else the built-in default).  Same arithmetic per accumulator element either way; returns the
 * schedule now in force.  A measurement / regression knob, not part of the reference's interface. */
int mnrf_debug_set_tc_schedule(int split);
/* algorithmic MACs per point (unpadded layer dims): full forward / sigma-only+pred-normal */
int64_t mnrf_macs_full(void);
int64_t mnrf_macs_sigma_only(void);

#ifdef __cplusplus
}
#endif
#endif /* MNRF_H_ */
