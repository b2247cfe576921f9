// Training pass of the hash-grid field (BASELINE config 3, `--model_type nerf_tcnn`): what torch.autograd records for
// R/models/rendering.py:87-266 around MirrorNeRFTcnn.forward (R/models/mirror_nerf_tcnn.py:151-259) when R/train.py:129-145
// calls render_rays with gradients enabled.
//   forward  = k_field_hash (field_hash.cu; raw records + analytic normals kept for the backward) -> k_composite
//   backward = k_train_composite_bwd (train.cu; per-point gradient record) -> k_hash_bwd (this file) [-> k_hash_ray_grad]
// k_hash_bwd is a persistent kernel, one CTA per SM, one warp per 32-point tile.  It RECOMPUTES the forward of its tile (the
// field costs ~11 k MAC + 128 table reads per point; saved activations would cost 2 KB per point of HBM traffic) and runs the
// phases of hash_train_math.cuh: gradients of the four small MLPs are accumulated in shared memory (one 44 KB image per CTA,
// flushed once with atomics), table gradients are scattered with global atomics (what tinycudann's grid backward does, in
// fp32 here), and the double backward through the analytic normal (create_graph=True in the reference) is explicit.
#include "common.cuh"
#include "hash_train_math.cuh"
#include "hash_train_math2.cuh"
#include <stdlib.h>
#include <string.h>

namespace mnrf {
namespace {

using namespace ht;

static_assert(sizeof(HashGridMeta) == sizeof(float) * (1 + 4 * HT_LEVELS), "HashGridMeta layout");
static_assert(HG_LEVELS == HT_LEVELS, "level count");
static_assert(HT_NW % 4 == 0 && HASH_WREF_FLOATS == HT_NW, "padded weight image");

constexpr int HB_WARPS = 4;  // warps per CTA (each owns HT_WARP_FLOATS of shared memory)
constexpr int HB_SMEM_FLOATS = 2 * HT_NW + HB_WARPS * HT_WARP_FLOATS;

struct SmallPtrs { float* p[12]; };  // index = mnrf_hash_field_create tensor order (entry 0 = table)

__global__ void __launch_bounds__(HB_WARPS * 32, 1)
k_hash_bwd(const float* __restrict__ table, const float* __restrict__ wref, HashGridMeta M, const float* __restrict__ rays,
           const float* __restrict__ z, const float* __restrict__ DR, const float* __restrict__ ray_detach_mirror, int P, int S,
           Flags F, int second_order, float* __restrict__ gtable, SmallPtrs gsmall, float* __restrict__ dxd) {
  extern __shared__ __align__(16) float sm[];
  float* Wt = sm;
  float* G = sm + HT_NW;
  for (int i = threadIdx.x; i < HT_NW; i += blockDim.x) {
    Wt[i] = wref[i];
    G[i] = 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* B = sm + 2 * HT_NW + warp * HT_WARP_FLOATS;
  const int n_tiles = (P + 31) / 32;
  for (int tile = blockIdx.x * HB_WARPS + warp; tile < n_tiles; tile += gridDim.x * HB_WARPS) {
    Lane L;
    float J[HT_J];
    const int p_raw = tile * 32 + lane;
    L.valid = p_raw < P;
    const int p = L.valid ? p_raw : P - 1;  // tail lanes shadow the last point with a zero gradient record
    const int ray = p / S;
    const float* rr = rays + (size_t)ray * 8;
    const float zz = z[p];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __fadd_rn(rr[c], __fmul_rn(rr[3 + c], zz));
      L.u[c] = __fdiv_rn(__fadd_rn(x, M.bound), __fmul_rn(2.f, M.bound));
      L.d[c] = rr[3 + c];
    }
    {
      const float4* dr = reinterpret_cast<const float4*>(DR + (size_t)p * HT_DR_STRIDE);
      const float4 r0 = dr[0], r1 = dr[1], r2 = dr[2];
      const float k = L.valid ? 1.f : 0.f;
      L.dr[0] = k * r0.x; L.dr[1] = k * r0.y; L.dr[2] = k * r0.z; L.dr[3] = k * r0.w;
      L.dr[4] = k * r1.x; L.dr[5] = k * r1.y; L.dr[6] = k * r1.z; L.dr[7] = k * r1.w;
      L.dr[8] = k * r2.x; L.dr[9] = k * r2.y; L.dr[10] = k * r2.z; L.dr[11] = 0.f;
    }
    L.mirror_on = !F.detach_mask && !(ray_detach_mirror != nullptr && ray_detach_mirror[ray] != 0.f);
    L.dmp = 0.f;
    L.dnraw[0] = L.dnraw[1] = L.dnraw[2] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) L.dsh[i] = 0.f;

    phase_a(Wt, B, table, M, F, L, J, lane);
    __syncwarp();
    phase_b(G, B, lane);
    __syncwarp();
    phase_c(Wt, B, L, lane);
    __syncwarp();
    phase_d(G, B, lane);
    __syncwarp();
    phase_e(Wt, B, lane);
    __syncwarp();
    phase_f(G, B, lane);
    __syncwarp();
    phase_g(Wt, B, F, L, lane);
    __syncwarp();
    phase_h(G, B, F, lane);
    __syncwarp();
    phase_i(Wt, B, F, L, lane);
    __syncwarp();
    phase_j(G, B, F, lane);
    __syncwarp();
    phase_k(Wt, B, F, L, lane);
    __syncwarp();
    phase_l(G, B, lane);
    __syncwarp();
    phase_m(Wt, B, table, gtable, M, F, L, J, lane, second_order != 0);
    __syncwarp();
    if (second_order) phase_n(G, B, lane);
    if (dxd != nullptr && L.valid) {
      const float inv2b = 1.f / (2.f * M.bound);
      float gd[3];
      sh4_bwd(L.d, L.dsh, gd);
      float4* o = reinterpret_cast<float4*>(dxd + (size_t)p_raw * HT_DXD_STRIDE);
      o[0] = make_float4(L.du[0] * inv2b, L.du[1] * inv2b, L.du[2] * inv2b, gd[0]);
      o[1] = make_float4(gd[1], gd[2], 0.f, 0.f);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int t = 1; t < 12; ++t) {
    float* dst = gsmall.p[t];
    if (dst == nullptr) continue;
    const int off = small_offset(t), cols = small_cols(t), ld = small_ld(t), cnt = small_rows(t) * cols;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {  // padded rows -> the parameter's own [out][in] layout
      const int r = i / cols, c = i - r * cols;
      const float v = G[off + r * ld + small_col(t, c)];
      if (v != 0.f) atomicAdd(dst + i, v);
    }
  }
}

// Second layout (hash_train_math2.cuh): 16 points per warp tile, two lanes per point, EIGHT warps per CTA.
constexpr int HB2_WARPS = 8;
constexpr int HB2_SMEM_FLOATS = 2 * HT_NW + HB2_WARPS * ht2::HT2_WARP_FLOATS;
static_assert(HB2_SMEM_FLOATS * sizeof(float) <= 232448, "k_hash_bwd2: shared memory");

__global__ void __launch_bounds__(HB2_WARPS * 32, 1)
k_hash_bwd2(const float* __restrict__ table, const float* __restrict__ wref, HashGridMeta M, const float* __restrict__ rays,
            const float* __restrict__ z, const float* __restrict__ DR, const float* __restrict__ ray_detach_mirror, int P, int S,
            Flags F, int second_order, float* __restrict__ gtable, SmallPtrs gsmall, float* __restrict__ dxd) {
  using namespace ht2;
  extern __shared__ __align__(16) float sm[];
  float* Wt = sm;
  float* G = sm + HT_NW;
  for (int i = threadIdx.x; i < HT_NW; i += blockDim.x) {
    Wt[i] = wref[i];
    G[i] = 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = lane & 15, half = lane >> 4;
  float* B = sm + 2 * HT_NW + warp * HT2_WARP_FLOATS;
  const bool so = second_order != 0;
  const int n_tiles = (P + NP2 - 1) / NP2;
  for (int tile = blockIdx.x * HB2_WARPS + warp; tile < n_tiles; tile += gridDim.x * HB2_WARPS) {
    Lane L;
    float J[HT2_J];
    const int p_raw = tile * NP2 + col;
    L.valid = p_raw < P;
    const int p = L.valid ? p_raw : P - 1;  // tail lanes shadow the last point with a zero gradient record
    const int ray = p / S;
    const float* rr = rays + (size_t)ray * 8;
    const float zz = z[p];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __fadd_rn(rr[c], __fmul_rn(rr[3 + c], zz));
      L.u[c] = __fdiv_rn(__fadd_rn(x, M.bound), __fmul_rn(2.f, M.bound));
      L.d[c] = rr[3 + c];
    }
    {
      const float4* dr = reinterpret_cast<const float4*>(DR + (size_t)p * HT_DR_STRIDE);
      const float4 r0 = dr[0], r1 = dr[1], r2 = dr[2];
      const float k = L.valid ? 1.f : 0.f;
      L.dr[0] = k * r0.x; L.dr[1] = k * r0.y; L.dr[2] = k * r0.z; L.dr[3] = k * r0.w;
      L.dr[4] = k * r1.x; L.dr[5] = k * r1.y; L.dr[6] = k * r1.z; L.dr[7] = k * r1.w;
      L.dr[8] = k * r2.x; L.dr[9] = k * r2.y; L.dr[10] = k * r2.z; L.dr[11] = 0.f;
    }
    L.mirror_on = !F.detach_mask && !(ray_detach_mirror != nullptr && ray_detach_mirror[ray] != 0.f);
    L.dmp = 0.f;
    L.dnraw[0] = L.dnraw[1] = L.dnraw[2] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) L.dsh[i] = 0.f;

    step_a1(B, table, M, L, J, lane);
    __syncwarp();
    step_a2(Wt, B, lane);
    __syncwarp();
    step_a3(Wt, B, L, lane);
    __syncwarp();
    step_a4(Wt, B, lane);
    __syncwarp();
    step_a5(Wt, B, lane);
    __syncwarp();
    step_a6(Wt, B, L, lane);
    __syncwarp();
    step_b(G, B, lane);
    __syncwarp();
    step_c(Wt, B, L, lane);
    __syncwarp();
    step_d(G, B, lane);
    __syncwarp();
    step_e(Wt, B, lane);
    __syncwarp();
    step_f(G, B, lane);
    __syncwarp();
    step_g1(Wt, B, F, L, lane);
    __syncwarp();
    step_g2(Wt, B, F, lane);
    __syncwarp();
    step_g3(Wt, B, F, L, lane);
    __syncwarp();
    step_h(G, B, F, lane);
    __syncwarp();
    step_i(Wt, B, F, L, lane);
    __syncwarp();
    step_j(G, B, F, lane);
    __syncwarp();
    step_k1(Wt, B, F, L, lane);
    __syncwarp();
    step_k2(Wt, B, lane);
    __syncwarp();
    step_l(G, B, lane);
    __syncwarp();
    step_m1(Wt, B, lane, so);
    __syncwarp();
    if (so) {
      step_m2(Wt, B, J, lane);
      __syncwarp();
      step_m3(B, M, L, J, lane);
      __syncwarp();
    }
    step_m4(Wt, B, table, gtable, M, F, L, J, lane, so);
    __syncwarp();
    if (so) step_n(G, B, lane);
    step_o(B, L, lane);
    if (dxd != nullptr && L.valid && half == 0) {
      const float inv2b = 1.f / (2.f * M.bound);
      float gd[3];
      sh4_bwd(L.d, L.dsh, gd);
      float4* o = reinterpret_cast<float4*>(dxd + (size_t)p_raw * HT_DXD_STRIDE);
      o[0] = make_float4(L.du[0] * inv2b, L.du[1] * inv2b, L.du[2] * inv2b, gd[0]);
      o[1] = make_float4(gd[1], gd[2], 0.f, 0.f);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int t = 1; t < 12; ++t) {
    float* dst = gsmall.p[t];
    if (dst == nullptr) continue;
    const int off = small_offset(t), cols = small_cols(t), ld = small_ld(t), cnt = small_rows(t) * cols;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const int r = i / cols, c = i - r * cols;
      const float v = G[off + r * ld + small_col(t, c)];
      if (v != 0.f) atomicAdd(dst + i, v);
    }
  }
}

// which layout runs: 2 (default) = 16 points x 2 lanes, 8 warps per CTA; 1 = 32 points x 1 lane, 4 warps per CTA
// (MNRF_HASH_BWD_LAYOUT=32x1 selects the first one; both are pinned to the same oracle by tests/test_hash_train_emu.py)
int hash_bwd_layout() {
  static int layout = -1;
  if (layout < 0) {
    const char* e = getenv("MNRF_HASH_BWD_LAYOUT");
    layout = (e != nullptr && strcmp(e, "32x1") == 0) ? 1 : 2;
  }
  return layout;
}

// d L / d [o, d] of one ray from the per-point records: x = o + d z, SH(d), x_surface = o + d * depth
__global__ void k_hash_ray_grad(const float* __restrict__ z, const float* __restrict__ dxd, const float* __restrict__ g_xs,
                                const float* __restrict__ depth, int n, int S, float* __restrict__ grad_rays) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
  for (int s = 0; s < S; ++s) {
    const size_t p = (size_t)r * S + s;
    const float4 a = *reinterpret_cast<const float4*>(dxd + p * HT_DXD_STRIDE);
    const float4 b = *reinterpret_cast<const float4*>(dxd + p * HT_DXD_STRIDE + 4);
    const float zz = z[p];
    go[0] += a.x; go[1] += a.y; go[2] += a.z;
    gd[0] += a.x * zz + a.w; gd[1] += a.y * zz + b.x; gd[2] += a.z * zz + b.y;
  }
  if (g_xs != nullptr) {
    const float dep = depth[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) { go[c] += g_xs[(size_t)r * 3 + c]; gd[c] += g_xs[(size_t)r * 3 + c] * dep; }
  }
  float* o = grad_rays + (size_t)r * 8;
  o[0] = go[0]; o[1] = go[1]; o[2] = go[2]; o[3] = gd[0]; o[4] = gd[1]; o[5] = gd[2]; o[6] = 0.f; o[7] = 0.f;
}

inline size_t al(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

// forward workspace: raw (P,8) | analytic normals (P,3).  backward workspace: DR (P,12) | dxd (P,8)
int64_t hash_train_fwd_workspace_bytes(int n, int S, int compute_normal) {
  const size_t P = (size_t)n * S;
  return (int64_t)(al(P * 8 * sizeof(float)) + (compute_normal ? al(P * 3 * sizeof(float)) : 0));
}
int64_t hash_train_bwd_workspace_bytes(int n, int S, int compute_normal) {
  (void)compute_normal;
  const size_t P = (size_t)n * S;
  return (int64_t)(al(P * HT_DR_STRIDE * sizeof(float)) + al(P * HT_DXD_STRIDE * sizeof(float)));
}

int hash_train_pass_fwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                        const mnrf_train_cfg& cfg, void* ws, const mnrf_composite_out& out, float* normal_out,
                        cudaStream_t st) {
  const int S = cfg.S;
  const size_t P = (size_t)n * S;
  float* raw = reinterpret_cast<float*>(ws);
  float* nrm = cfg.compute_normal ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + al(P * 8 * sizeof(float))) : nullptr;
  FieldIO io{};
  io.rays = rays; io.z = z; io.x = nullptr; io.x_stride = 0; io.dirbias = nullptr;
  io.n_points = (int)P; io.S = S; io.sigma_only = 0;
  io.raw = raw; io.sigma_out = nullptr; io.normal_out = nrm; io.geo_out = nullptr;
  if (launch_field_hash(f, io, st)) return 1;
  if (nrm != nullptr && normal_out != nullptr)
    MNRF_CUDA_OK(cudaMemcpyAsync(normal_out, nrm, sizeof(float) * P * 3, cudaMemcpyDeviceToDevice, st));
  return launch_composite(rays, z, raw, 8, raw, nrm, noise, cfg.noise_std, n, S, cfg.white_back, out, st);
}

// gt: 12 device pointers in mnrf_hash_field_create order (0 = encoder.params gradient), ACCUMULATED into
int hash_train_pass_bwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                        const mnrf_train_cfg& cfg, const void* ws_fwd, void* ws_bwd, const mnrf_train_grads& g,
                        const float* ray_detach_mirror, float* const* gt, const float* depth, float* grad_rays,
                        cudaStream_t st) {
  const int S = cfg.S;
  const size_t P = (size_t)n * S;
  const float* raw = reinterpret_cast<const float*>(ws_fwd);
  const float* nrm = cfg.compute_normal
                         ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(ws_fwd) + al(P * 8 * sizeof(float)))
                         : nullptr;
  float* DR = reinterpret_cast<float*>(ws_bwd);
  float* dxd = grad_rays != nullptr
                   ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws_bwd) + al(P * HT_DR_STRIDE * sizeof(float)))
                   : nullptr;
  for (int i = 0; i < 6; ++i) MNRF_REQUIRE(gt[i] != nullptr, "hash train_pass_bwd: gradient tensor %d missing", i);
  if (f->has_normal) for (int i = 6; i < 8; ++i) MNRF_REQUIRE(gt[i] != nullptr, "hash train_pass_bwd: gradient tensor %d missing", i);
  if (f->has_mirror) for (int i = 8; i < 12; ++i) MNRF_REQUIRE(gt[i] != nullptr, "hash train_pass_bwd: gradient tensor %d missing", i);
  MNRF_REQUIRE(f->hash_wref != nullptr, "hash train_pass_bwd: field has no reference-layout weights");
  if (launch_composite_bwd(rays, z, raw, nrm, noise, cfg, n, ray_detach_mirror, g, DR, st)) return 1;

  Flags F;
  F.has_normal = f->has_normal; F.has_mirror = f->has_mirror; F.compute_normal = cfg.compute_normal;
  F.detach_normal = cfg.detach_density_for_normal_loss; F.detach_mask = cfg.detach_density_for_mask_loss;
  F.ray_grad = grad_rays != nullptr;
  const int second_order = cfg.compute_normal && (g.normal != nullptr || g.surface_normal_grad != nullptr || g.normal_dif != nullptr);
  SmallPtrs sp;
  for (int i = 0; i < 12; ++i) sp.p[i] = gt[i];
  if (first_use_on_device(TAG_TRAIN_HASH)) {
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_hash_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(HB_SMEM_FLOATS * sizeof(float))));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_hash_bwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(HB2_SMEM_FLOATS * sizeof(float))));
  }
  if (hash_bwd_layout() == 2) {
    const int n_tiles = (int)((P + ht2::NP2 - 1) / ht2::NP2);
    int blocks = (n_tiles + HB2_WARPS - 1) / HB2_WARPS;
    if (blocks > 148) blocks = 148;
    k_hash_bwd2<<<blocks, HB2_WARPS * 32, HB2_SMEM_FLOATS * sizeof(float), st>>>(f->hash_table, f->hash_wref, f->hg, rays, z, DR,
                                                                                ray_detach_mirror, (int)P, S, F, second_order,
                                                                                gt[0], sp, dxd);
  } else {
    const int n_tiles = (int)((P + 31) / 32);
    int blocks = (n_tiles + HB_WARPS - 1) / HB_WARPS;
    if (blocks > 148) blocks = 148;
    k_hash_bwd<<<blocks, HB_WARPS * 32, HB_SMEM_FLOATS * sizeof(float), st>>>(f->hash_table, f->hash_wref, f->hg, rays, z, DR,
                                                                             ray_detach_mirror, (int)P, S, F, second_order, gt[0],
                                                                             sp, dxd);
  }
  MNRF_LAUNCH_OK();
  if (grad_rays != nullptr) {
    MNRF_REQUIRE(g.x_surface == nullptr || depth != nullptr, "hash train_pass_bwd: ray gradients need the depth output");
    k_hash_ray_grad<<<(n + 127) / 128, 128, 0, st>>>(z, dxd, g.x_surface, depth, n, S, grad_rays);
    MNRF_LAUNCH_OK();
  }
  return 0;
}

}  // namespace mnrf
