// Shared declarations of libmnrf (internal).  See include/mnrf.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mnrf.h"

namespace mnrf {

// ------------------------------------------------------------------------------------------------
// architecture constants (R/models/mirror_nerf.py:41-99 defaults: D=8, W=256, skips=[4], 10/4 freqs)
// ------------------------------------------------------------------------------------------------
constexpr int W = 256;           // trunk width
constexpr int WH = 128;          // head width (W/2)
constexpr int NFREQ_XYZ = 10;
constexpr int NFREQ_DIR = 4;
constexpr int IN_XYZ = 3 + 6 * NFREQ_XYZ;  // 63
constexpr int IN_DIR = 3 + 6 * NFREQ_DIR;  // 27
constexpr int PE_PAD = 64;                 // 63 -> 64
constexpr float FP32_EPS = 1.1920928955078125e-07f;

// parameter tensor indices (mnrf.h)
enum {
  T_XYZ_W0 = 0,  // 2*i weight, 2*i+1 bias, i=0..7
  T_FINAL_W = 16, T_FINAL_B = 17, T_DIR_W = 18, T_DIR_B = 19, T_SIGMA_W = 20, T_SIGMA_B = 21,
  T_RGB_W = 22, T_RGB_B = 23, T_N0_W = 24, T_N0_B = 25, T_N1_W = 26, T_N1_B = 27,
  T_M0_W = 28, T_M0_B = 29, T_M2_W = 30, T_M2_B = 31
};

// ------------------------------------------------------------------------------------------------
// tensor-core GEMM steps of one 128-point tile (field_tc.cu) and their packed-weight blobs
// ------------------------------------------------------------------------------------------------
// Steps 11..19 are the TRANSPOSED trunk weights of the analytic-normal chain d sigma / d xyz (mirror_nerf.py:136-146):
//   11: W8^T  12: W7^T  13: W6^T  14: W5^T (h part, 256 inputs)  15: W5^T (PE part, 63->64 inputs)
//   16: W4^T  17: W3^T  18: W2^T  19: W1^T (63->64 inputs)          B operand [n = input feature][k = output feature]
constexpr int TC_FWD_STEPS = 11;
constexpr int TC_NUM_STEPS = 20;
// step:            0    1    2    3    4    5    6    7    8(final) 9(mirror0) 10(dir)
__host__ __device__ constexpr int tc_step_n(int s) { return s <= 8 ? 256 : (s <= 10 ? 128 : ((s == 15 || s == 19) ? 64 : 256)); }
__host__ __device__ constexpr int tc_step_k(int s) { return s == 0 ? 64 : (s == 4 ? 320 : 256); }
// forward trunk layer (0-based) whose weight matrix a transposed step uses
__host__ __device__ constexpr int tc_step_layer(int s) { return s <= 13 ? 18 - s : (s <= 15 ? 4 : 19 - s); }
__host__ __device__ constexpr int tc_step_chunks(int s) { return tc_step_k(s) / 32; }
// One blob = hi or lo part of one K32 chunk of a step's weight matrix: N rows x 32 halves (16 KB for N=256, 8 KB for
// N=128), stored as the shared-memory image of a tcgen05 K-major no-swizzle B operand.  Order inside a step: [chunk kc][hi, lo].
__host__ __device__ constexpr int tc_blob_bytes(int s) { return tc_step_n(s) * 64; }
__host__ __device__ constexpr int tc_step_bytes(int s) { return tc_step_chunks(s) * 2 * tc_blob_bytes(s); }
__host__ __device__ constexpr int tc_step_offset(int s) {
  int o = 0;
  for (int i = 0; i < s; ++i) o += tc_step_bytes(i);
  return o;
}
constexpr int TC_TOTAL_BYTES = tc_step_offset(TC_NUM_STEPS);
constexpr int TC_SPLIT_LAST = 8;   // steps 0..8 (the 256-wide forward layers) also exist in the N-split layout of the tc2 blobs (pack.cu)

// tf32 hi/lo blobs of the training GEMMs (train_tc.cu): steps 0..19 as above plus
//   20: normal_net.0 (N128,K256)   21: dir layer^T, feature part (N256,K128)   22: final^T (N256,K256)
//   23: normal_net.0^T (N256,K128) 24: is_mirror_net.0^T (N256,K128)
// One blob = hi or lo part of one K16 chunk: N rows x 16 floats (N*64 bytes) as the shared-memory image of a tcgen05 K-major
// no-swizzle B operand (core matrix = 8 rows x 4 tf32).  Order inside a step: [chunk kc][hi, lo].
constexpr int T32_NUM_STEPS = 25;
__host__ __device__ constexpr int t32_step_n(int s) { return s < TC_NUM_STEPS ? tc_step_n(s) : (s == 20 ? 128 : 256); }
__host__ __device__ constexpr int t32_step_k(int s) { return s < TC_NUM_STEPS ? tc_step_k(s) : ((s == 20 || s == 22) ? 256 : 128); }
__host__ __device__ constexpr long long t32_step_offset(int s) {
  long long o = 0;
  for (int i = 0; i < s; ++i) o += (long long)t32_step_n(i) * t32_step_k(i) * 8;
  return o;
}
constexpr long long T32_TOTAL_BYTES = t32_step_offset(T32_NUM_STEPS);

// ------------------------------------------------------------------------------------------------
// fp32 section layout (float offsets inside mnrf_field::f32)
// ------------------------------------------------------------------------------------------------
// Contiguous "epilogue table" inside the fp32 section (float offsets relative to F32Layout::epi_tab): everything the
// tensor-core kernel's epilogues read that is uniform across a warp.  It is copied to __constant__ memory before a launch,
// so those reads become constant-cache (LDC) accesses instead of global loads.
constexpr int ET_BIAS = 0;                      // [9][256]: trunk layers 1..8, xyz_encoding_final
constexpr int ET_HEADW = ET_BIAS + 9 * 256;     // float4[256]: {w_sigma[c], Wn_fold[0..2][c]}
constexpr int ET_B_M0 = ET_HEADW + 4 * 256;     // [128]
constexpr int ET_W_M2 = ET_B_M0 + 128;          // [128]
constexpr int ET_W_RGB = ET_W_M2 + 128;         // [3][128]
constexpr int ET_HEADB = ET_W_RGB + 3 * 128;    // {b_sigma, bn_fold[0..2]}
constexpr int ET_INV_SCALE = ET_HEADB + 4;      // [20] one per GEMM step
constexpr int ET_B_M2 = ET_INV_SCALE + 20;      // [4] (1 used)
constexpr int ET_B_RGB = ET_B_M2 + 4;           // [4] (3 used)
constexpr int ET_TOTAL = ET_B_RGB + 4;          // 3992 floats = 15,968 bytes

struct F32Layout {
  int epi_tab;
  // transposed, K padded to a multiple of 4 rows: Wt[k][n]
  int wt_trunk[8];   // layers 1..8
  int wt_final, wt_dir, wt_n0, wt_m0;
  // original [out,in] copies (for the analytic-normal backward chain and the tiny heads)
  int w_trunk[8];
  int w_sigma, w_rgb, w_n1, w_m2;
  // biases
  int b_trunk[8];
  int b_final, b_dir, b_sigma, b_rgb, b_n0, b_n1, b_m0, b_m2;
  // tensor-core epilogue tables
  int headw;      // float4[256]: {w_sigma[c], Wn_fold[0][c], Wn_fold[1][c], Wn_fold[2][c]}
  int headb;      // float4: {b_sigma, bn_fold[0..2]}
  int inv_scale;  // float[TC_NUM_STEPS]
  int absmax;     // uint[TC_NUM_STEPS] scratch for the scale computation
  // training path (train.cu): [out][in] copies with 16-byte aligned rows, the B operands of the dgrad / normal-chain GEMMs
  int tw_l1;      // [256][64]   W1, column 63 zero
  int tw_l5a;     // [256][64]   W5[:, :63] (PE part), column 63 zero
  int tw_l5b;     // [256][256]  W5[:, 63:] (h part)
  int tw_final;   // [256][256]
  int tw_dira;    // [128][256]  W_dir[:, :256] (feature part)
  int tw_n0;      // [128][256]
  int tw_m0;      // [128][256]
  int total;
};

__host__ __device__ inline int trunk_k(int l) { return l == 0 ? IN_XYZ : (l == 4 ? IN_XYZ + W : W); }
__host__ __device__ inline int pad4(int k) { return (k + 3) & ~3; }

inline F32Layout make_f32_layout() {
  F32Layout L;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };  // keep 16B alignment
  for (int l = 0; l < 8; ++l) L.wt_trunk[l] = take(pad4(trunk_k(l)) * W);
  L.wt_final = take(W * W);
  L.wt_dir = take(pad4(W + IN_DIR) * WH);
  L.wt_n0 = take(W * WH);
  L.wt_m0 = take(W * WH);
  for (int l = 0; l < 8; ++l) L.w_trunk[l] = take(W * trunk_k(l));
  L.w_sigma = take(W);
  L.w_n1 = take(3 * WH);
  L.b_dir = take(WH);
  L.b_sigma = take(4);
  L.b_n0 = take(WH);
  L.b_n1 = take(4);
  L.epi_tab = take(ET_TOTAL);
  for (int l = 0; l < 8; ++l) L.b_trunk[l] = L.epi_tab + ET_BIAS + 256 * l;
  L.b_final = L.epi_tab + ET_BIAS + 256 * 8;
  L.headw = L.epi_tab + ET_HEADW;
  L.b_m0 = L.epi_tab + ET_B_M0;
  L.w_m2 = L.epi_tab + ET_W_M2;
  L.w_rgb = L.epi_tab + ET_W_RGB;
  L.headb = L.epi_tab + ET_HEADB;
  L.inv_scale = L.epi_tab + ET_INV_SCALE;
  L.b_m2 = L.epi_tab + ET_B_M2;
  L.b_rgb = L.epi_tab + ET_B_RGB;
  L.absmax = take(32);
  L.tw_l1 = take(W * PE_PAD);
  L.tw_l5a = take(W * PE_PAD);
  L.tw_l5b = take(W * W);
  L.tw_final = take(W * W);
  L.tw_dira = take(WH * W);
  L.tw_n0 = take(WH * W);
  L.tw_m0 = take(WH * W);
  L.total = o;
  return L;
}

}  // namespace mnrf

namespace mnrf {
// hash-grid field (field_hash.cu; R/models/mirror_nerf_tcnn.py): level table + packed small MLPs
constexpr int HG_LEVELS = 16;
struct HashGridMeta {
  float bound;
  float scale[HG_LEVELS];
  int res[HG_LEVELS];
  unsigned int offset[HG_LEVELS];  // in table entries (2 floats each)
  unsigned int size[HG_LEVELS];
};
// float offsets inside the packed weight block (rows padded to multiples of 4 inputs; see field_hash.cu)
constexpr int HW_S0 = 0, HW_S1 = 2048, HW_C0 = 3072, HW_C1 = 5120, HW_C2 = 9216, HW_N0 = 9472, HW_N1 = 10496, HW_M0 = 10752,
              HW_M0B = 11264, HW_M2 = 11296, HW_M2B = 11328, HW_TOTAL = 11332;
constexpr int HASH_WREF_FLOATS = 11204;  // the same weights as [out][in] rows padded to 4 inputs (hash_train_math.cuh: HT_NW)
}  // namespace mnrf

struct mnrf_field {
  int kind;          // 0 = MirrorNeRF MLP field, 1 = hash-grid field
  float* hash_table; // kind 1: device copy of encoder.params
  float* hash_w;     // kind 1: HW_TOTAL packed floats
  float* hash_wref;  // kind 1: the small MLP weights as [out][in] rows padded to 4 inputs (hash_train_math.cuh offsets), for the backward
  mnrf::HashGridMeta hg;
  int has_normal;
  int has_mirror;
  float* f32;        // device, mnrf::F32Layout
  uint8_t* tc;       // device, TC_TOTAL_BYTES of fp16 hi/lo blobs
  uint8_t* tc8;      // device, 2 x TC_TOTAL_BYTES: per K32 chunk [fp16 hi | e4m3(2^-10 hi), e4m3(lo)] blobs of the fp8-corrected mode,
                     // then the same step offsets again with steps 0..TC_SPLIT_LAST in the N-split layout (pack.cu k_pack_tc8)
  uint8_t* t32;      // device, T32_TOTAL_BYTES of tf32 hi/lo blobs (training GEMMs)
  mnrf::F32Layout L;
  unsigned long long pack_stamp;  // unique per pack_field call (keys the constant-memory copies of the epilogue table)
};

namespace mnrf {

// error plumbing ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// One-time per-DEVICE launch setup (cudaFuncSetAttribute applies to the device that is current when it is called):
// returns true the first time it is called for `tag` on the current device.  `num_sms` (optional) = SM count of that device.
bool first_use_on_device(int tag, int* num_sms = nullptr);
enum { TAG_FIELD_TC = 0, TAG_FIELD_FP32, TAG_FIELD_HASH, TAG_TRAIN_TC_NN, TAG_TRAIN_TC_TN, TAG_TRAIN_HASH, TAG_COUNT };
#define MNRF_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      mnrf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)
#define MNRF_LAUNCH_OK()                                                                \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      mnrf::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                                         \
    }                                                                                   \
    mnrf::count_launch();                                                               \
  } while (0)
#define MNRF_REQUIRE(cond, ...)                                                         \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      mnrf::set_error(__VA_ARGS__);                                                     \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

// optional timing of the dominant kernel with CUDA events on its own launch stream (mnrf_profile_*)
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st, double flops);

// internal launchers (each returns 0 / error) ----------------------------------------------------
int pack_field(mnrf_field* f, const float* const* tensors, cudaStream_t st);

// per-ray (or per-point) additive term of the dir layer: b_dir + W_dir[:,256:283] . embed(dir)
// from_embedded == 0: src = rays (n,8), direction embedded in-kernel; else src = x (n,30), cols 3..29
int launch_dirbias(const mnrf_field* f, const float* src, int n, int src_stride, int from_embedded, float* out,
                   cudaStream_t st, const int* n_dev = nullptr);

struct FieldIO {
  // geometry: point p -> ray p / S, sample p % S;  xyz = o + d*z  (flat mode: S = 1, xyz read from x)
  const float* rays;    // (n_rays,8) or NULL in flat mode
  const float* z;       // (n_rays,S)
  const float* x;       // flat mode: (B, x_stride) rows, first 3 = xyz
  int x_stride;
  const float* dirbias; // (n_rays or B, 128) ; NULL when sigma_only
  int n_points;
  int S;
  int sigma_only;
  // optional device-side ray count (mnrf_render_recursive: a level is launched for its worst-case size and the kernels read
  // how many rays are really alive): n_points := min(n_points, *n_rays_dev * S)
  const int* n_rays_dev;
  // outputs (NULL = skip)
  float* raw;        // (n_points, 8)
  float* sigma_out;  // (n_points)
  float* normal_out; // (n_points, 3) analytic normal (fp32 kernel only)
  float* geo_out;    // (n_points, 256)
};

// `n_dev` (last argument of the launchers below, optional): device-side count; the kernel processes min(n, *n_dev) rows
int launch_coarse_z(const float* rays, int n, const float* z_steps, int S, int use_disp, float perturb,
                    const float* u, float* z_out, cudaStream_t st, const int* n_dev = nullptr);
int launch_generate_rays(int H, int W, float focal, const float* c2w_host, float near, float far, float* rays,
                         cudaStream_t st);
int launch_embed(const float* x, int n, int n_freqs, float* out, cudaStream_t st);
int launch_searchsorted(const float* cdf, int n, int m, const float* u, int n_u, int u_stride, int64_t* inds,
                        cudaStream_t st);
int launch_sample_pdf(const float* z_coarse, const float* bins, const float* weights, int w_stride, int w_off, int n,
                      int S, int n_imp, const float* u, int u_stride, float* z_fine, float* samples, int64_t* inds,
                      float* cdf, cudaStream_t st, const int* n_dev = nullptr);
int launch_composite(const float* rays, const float* z, const float* sigma, int sigma_stride, const float* raw,
                     const float* normal, const float* noise, float noise_std, int n, int S, int white_back,
                     const mnrf_composite_out& out, cudaStream_t st, const int* n_dev = nullptr);
// jitter (optional): normal += jitter_scale * jitter[i]  before normalising (R/eval.py:506-511 roughness cone)
int launch_reflect(const float* rays, const float* x_surface, const float* normal, float* mask, int n, float near2,
                   float* sec, float* refl, int* any_mirror, cudaStream_t st, const int* n_dev = nullptr,
                   const float* jitter = nullptr, float jitter_scale = 0.f, int threshold_mask = 1);
// scratch (optional): (n + 1023) / 1024 ints of caller-owned scratch instead of a stream-ordered allocation
int launch_compact(const float* in, const float* mask, int n, int row_floats, float* out, int* index, int* count,
                   cudaStream_t st, const int* n_dev = nullptr, int* scratch = nullptr);
int launch_axpy_rows(float* dense, const float* compact, const int* index, int n, int c, float alpha, float beta,
                     cudaStream_t st, const int* n_dev = nullptr);
// traced (optional, device): when *traced == 0 the level below was not rendered: rgb_out = base, reflect outputs = 0
int launch_blend(const float* base, const float* mask, const float* child_rgb, const float* child_depth,
                 const int* index, int n, float* rgb_out, float* rgb_reflect, float* depth_reflect, cudaStream_t st,
                 const int* n_dev = nullptr, const int* traced = nullptr);
// *out = (*flag != 0) ? min(n, n_dev ? *n_dev : n) : 0   -- size of an un-compacted child level (R/eval.py:159,311-320)
int launch_select_count(const int* flag, int n, const int* n_dev, int* out, cudaStream_t st);

int launch_field_fp32(const mnrf_field* f, const FieldIO& io, cudaStream_t st);
int launch_field_hash(const mnrf_field* f, const FieldIO& io, cudaStream_t st);
int pack_hash_field(mnrf_field* f, const float* const* tensors, long long table_floats, cudaStream_t st);
// training pass of the hash-grid field (train_hash.cu) and the compositor backward it shares with the MLP field (train.cu)
int64_t hash_train_fwd_workspace_bytes(int n, int S, int compute_normal);
int64_t hash_train_bwd_workspace_bytes(int n, int S, int compute_normal);
int hash_train_pass_fwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                        const mnrf_train_cfg& cfg, void* ws, const mnrf_composite_out& out, float* normal_out, cudaStream_t st);
int hash_train_pass_bwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                        const mnrf_train_cfg& cfg, const void* ws_fwd, void* ws_bwd, const mnrf_train_grads& g,
                        const float* ray_detach_mirror, float* const* gt, const float* depth, float* grad_rays, cudaStream_t st);
int launch_composite_bwd(const float* rays, const float* z, const float* raw, const float* normal, const float* noise,
                         const mnrf_train_cfg& cfg, int n, const float* ray_detach_mirror, const mnrf_train_grads& g, float* DR,
                         cudaStream_t st);

// epilogue of the training GEMMs:  v = acc [+ C] [+ bias[col]] [+ rowbias[row / rb_div][col]] [+ rvec[row] * cvec[col]];  act(v)
struct GemmEpi {
  const float* bias = nullptr;
  const float* rowbias = nullptr;
  int rb_div = 1, ld_rb = 0;
  const float* rvec = nullptr;
  int ld_rvec = 0;
  const float* cvec = nullptr;
  const float* mask = nullptr;  // act == 2: keep v where mask[row][col] > 0 (relu' of a saved activation), else 0
  int ld_mask = 0;
  // the same mask as one bit per element, 8 words per row of 256 columns (bit c&31 of word c>>5): the tensor-core kernels read
  // this (32 B per point, prefetched for a whole tile while the MMAs run) instead of the 1 KB activation row, and write it
  // (bits_out) from the bias+ReLU epilogue of the forward trunk
  const uint32_t* mask_bits = nullptr;
  uint32_t* bits_out = nullptr;
  int act = 0;                  // 0 none | 1 relu | 2 mask | 3 leaky relu (0.01)
  int accumulate = 0;
};
// C[M,N] = epi([A0 | A1] * B_step^T) on the tensor cores (3x tf32 split); A0 supplies the first K0 reduction columns
int gemm_nn_tc(const mnrf_field* f, int step, const float* A0, int lda0, int K0, const float* A1, int lda1, float* C,
               int ldc, int M, const GemmEpi& e, cudaStream_t st);
// Wg[NA rows][col0 + (0..valid)] += A[P,NA]^T B[P,NB] on the tensor cores (3x tf32 split), NA in {128,256}, NB in {64,128,256}
// colsum_out (optional): colsum_out[f] += sum_p A[p][f], the bias gradient, fused into the same pass over A
int gemm_tn_tc(const float* A, int lda, int NA, const float* B, int ldb, int NB, float* Wg, int ldw, int col0, int valid,
               int P, float* colsum_out, cudaStream_t st);

void set_train_tc_debug(int flags);
void set_train_tc_one_pass(int on);

// training path (train.cu): one field + compositor pass with saved activations, and its backward
int64_t train_fwd_workspace_bytes(int n, int S, int compute_normal);
int64_t train_bwd_workspace_bytes(int n, int S, int compute_normal);
int train_pass_fwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                   const mnrf_train_cfg& cfg, void* ws, const mnrf_composite_out& out, float* normal_out,
                   cudaStream_t st);
int train_pass_bwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                   const mnrf_train_cfg& cfg, const void* ws_fwd, void* ws_bwd, const mnrf_train_grads& g,
                   const float* ray_detach_mirror, float* const* grad_tensors, const float* depth, float* grad_rays,
                   cudaStream_t st);
bool can_fuse_composite(const mnrf_field* f, const mnrf_level_cfg* cfg, const float* noise, int S);
int render_level(const mnrf_field* coarse, const mnrf_field* fine, const float* rays, int n, const mnrf_level_cfg* cfg,
                 const mnrf_level_rng* rng, const float* z_steps, const float* u_det, void* workspace, int64_t workspace_bytes,
                 const mnrf_level_out* out, void* stream, const int* n_dev);
void set_tc_trace(unsigned long long* buf, unsigned int cap);
int set_tc_split(int split);   // -1 = default (MNRF_TC_SPLIT or built-in); returns the schedule in force
// compositing fused into the tensor-core field kernel (field_tc.cu FUSE): per-ray outputs straight from the epilogue registers
struct FusedComposite {
  mnrf_composite_out comp;     // opacity required; weights / pred_normal (per sample) optional
  int white_back;
  float term_eps;              // > 0: early ray termination (only when no per-sample output is requested)
  int* work_counter;           // device int (zeroed by the launcher)
  unsigned long long* stats;   // optional device counters: [0] tiles executed, [1] 32-sample chunks skipped
};
int launch_field_tc(const mnrf_field* f, const FieldIO& io, int precision /*1|2|3*/, cudaStream_t st,
                    const FusedComposite* fuse = nullptr);

}  // namespace mnrf
