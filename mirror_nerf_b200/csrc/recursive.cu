// mnrf_render_recursive: the Whitted recursion of the reference's callers as one device-side call (include/mnrf.h).
//   R/eval.py::batched_inference :132-160 (level call), :295-320 (mask threshold, trace condition), :336-360 (normal),
//   :506-548 (jitter, reflect, secondary rays, compaction), :609-674 (recursive call, roughness cone), :676-723 (blend);
//   R/train.py:129-348 is the same structure with only_trace_rays_in_mirrors taken from the hparams.
//
// No host synchronisation anywhere: the host walks a FIXED launch plan (it depends only on n, max_recursive_level,
// trace_ray_times and the slab size), every kernel of a level is launched for that level's row capacity and reads the live
// row count from device memory (common.cuh: n_dev / FieldIO::n_rays_dev), so a level without mirror rays costs a handful of
// empty launches.  The children of one level -- the first reflection of its rays (all rays at eval level 0, else the mirror
// rays) followed by the T extra jittered reflections of its mirror rays -- form ONE ray list:
//     block 0: rows [0, b0)                      b0 = traced ? (compact ? c : rays) : 0        c = number of mirror rays
//     block t: rows [b0 + (t-1) c, b0 + t c)     t = 1..T (always compacted, eval.py:647-649)
// which is rendered in slabs of at most `slab_rows` rows; the parent then averages its T+1 child colours in block order
// (the reference's summation order, eval.py:655-673) and blends.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace mnrf {
namespace {

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// ---- Philox4x32-10 + Box-Muller: three standard normals per (seed, stream, row) -------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ void normal3(unsigned long long seed, uint32_t stream, uint32_t row, float (&z)[3]) {
  uint32_t c[4] = {row, stream, 0x6d6e7266u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  const float u0 = ((float)(c[0] >> 8) + 0.5f) * (1.f / 16777216.f), u1 = ((float)(c[1] >> 8) + 0.5f) * (1.f / 16777216.f);
  const float u2 = ((float)(c[2] >> 8) + 0.5f) * (1.f / 16777216.f), u3 = ((float)(c[3] >> 8) + 0.5f) * (1.f / 16777216.f);
  const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
  float s, co;
  sincosf(6.283185307179586f * u1, &s, &co);
  z[0] = r0 * co; z[1] = r0 * s;
  z[2] = r1 * cosf(6.283185307179586f * u3);
}

// layout[0] = b0 (rows of block 0), [1] = total child rows, [2] = traced (0/1), [3] = c (mirror rays), [4 + s] = rows of slab s
__global__ void k_child_layout(const int* __restrict__ flag, const int* __restrict__ n_dev, int cap, const int* __restrict__ c_dev,
                               int compact0, int T, int slab, int n_slabs, int* __restrict__ layout, int* __restrict__ level_rays) {
  const int n = n_dev != nullptr ? min(cap, max(*n_dev, 0)) : cap;
  const int traced = (*flag != 0) ? 1 : 0;   // mirror_mask.any() (eval.py:311-318)
  const int c = traced ? *c_dev : 0;
  const int b0 = traced ? (compact0 ? c : n) : 0;
  const int total = b0 + T * c;
  layout[0] = b0; layout[1] = total; layout[2] = traced; layout[3] = c;
  for (int s = 0; s < n_slabs; ++s) layout[4 + s] = min(max(total - s * slab, 0), slab);
  if (level_rays != nullptr) atomicAdd(level_rays, total);
}

// reflection t of every parent ray that has a child row in block t (R/eval.py:506-540, 627-649)
__global__ void k_reflect_scatter(const float* __restrict__ rays, const float* __restrict__ x_surface,
                                  const float* __restrict__ normal, const int* __restrict__ index,
                                  const int* __restrict__ layout, const int* __restrict__ n_dev, int cap, int t, int compact0,
                                  const float* __restrict__ noise, float noise_std, unsigned long long seed, uint32_t stream,
                                  float near2, float* __restrict__ child_rays, float* __restrict__ refl_out) {
  const int n = n_dev != nullptr ? min(cap, max(__ldg(n_dev), 0)) : cap;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || layout[2] == 0) return;
  int dst;
  if (t == 0 && !compact0) dst = i;
  else {
    const int k = index[i];
    if (k < 0) return;
    dst = (t == 0 ? 0 : layout[0] + (t - 1) * layout[3]) + k;
  }
  const float* ry = rays + (size_t)i * 8;
  float nx = normal[i * 3], ny = normal[i * 3 + 1], nz = normal[i * 3 + 2];
  if (noise_std > 0.f) {
    float z[3];
    if (noise != nullptr) { z[0] = noise[(size_t)i * 3]; z[1] = noise[(size_t)i * 3 + 1]; z[2] = noise[(size_t)i * 3 + 2]; }
    else normal3(seed, stream, (uint32_t)i, z);
    nx = __fadd_rn(nx, __fmul_rn(z[0], noise_std));   // normal + randn * std, separately rounded as torch does
    ny = __fadd_rn(ny, __fmul_rn(z[1], noise_std));
    nz = __fadd_rn(nz, __fmul_rn(z[2], noise_std));
  }
  float nn = sqrtf(fmaxf(nx * nx + ny * ny + nz * nz, FP32_EPS));  // utils/func.py:5-7
  nx /= nn; ny /= nn; nz /= nn;
  float wx = -ry[3], wy = -ry[4], wz = -ry[5];
  float wn = sqrtf(fmaxf(wx * wx + wy * wy + wz * wz, FP32_EPS));
  wx /= wn; wy /= wn; wz /= wn;
  const float c2 = 2.f * (wx * nx + wy * ny + wz * nz);
  const float rx = c2 * nx - wx, ryy = c2 * ny - wy, rz = c2 * nz - wz;  // 2 (n.w) n - w
  float* o = child_rays + (size_t)dst * 8;
  o[0] = x_surface[i * 3]; o[1] = x_surface[i * 3 + 1]; o[2] = x_surface[i * 3 + 2];
  o[3] = rx; o[4] = ryy; o[5] = rz;
  o[6] = near2;   // ray_forward_offset = 0.1 (eval.py:529)
  o[7] = ry[7];   // parent far
  if (refl_out != nullptr) { refl_out[i * 3] = rx; refl_out[i * 3 + 1] = ryy; refl_out[i * 3 + 2] = rz; }
}

// average the T+1 child colours of a mirror ray in block order, then rgb = m * reflect + (1 - m) * base (eval.py:655-697)
__global__ void k_gather_blend(const float* __restrict__ base, const float* __restrict__ mask, const int* __restrict__ index,
                               const int* __restrict__ layout, const float* __restrict__ child_rgb,
                               const float* __restrict__ child_depth, const int* __restrict__ n_dev, int cap, int T, int compact0,
                               float* __restrict__ rgb_out, float* __restrict__ rgb_reflect, float* __restrict__ depth_reflect,
                               float* __restrict__ refl_dir) {
  const int n = n_dev != nullptr ? min(cap, max(__ldg(n_dev), 0)) : cap;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool traced = layout[2] != 0;
  const int k = traced ? index[i] : -1;                   // row inside a compacted block, -1 = not a mirror ray
  const int row0 = traced ? (compact0 ? k : i) : -1;      // row in block 0
  const float m = mask[i] != 0.f ? 1.f : 0.f;             // mirror_mask.bool().float() (eval.py:307,678)
  for (int ch = 0; ch < 3; ++ch) {
    const float b = base[i * 3 + ch];
    float rf = b;                                         // compacted child: rays outside the mirror keep the base colour (:684-688)
    if (row0 >= 0) {
      rf = child_rgb[(size_t)row0 * 3 + ch];
      if (T > 0 && k >= 0) {
        for (int t = 1; t <= T; ++t) rf = __fadd_rn(rf, child_rgb[((size_t)layout[0] + (size_t)(t - 1) * layout[3] + k) * 3 + ch]);
        rf = __fdiv_rn(rf, (float)(T + 1));               // reference: / (trace_ray_times + 1)
      }
    }
    if (rgb_reflect != nullptr) rgb_reflect[i * 3 + ch] = row0 >= 0 ? rf : 0.f;
    rgb_out[i * 3 + ch] = traced ? m * rf + (1.f - m) * b : b;
  }
  if (depth_reflect != nullptr) depth_reflect[i] = row0 >= 0 ? child_depth[row0] : 0.f;
  if (refl_dir != nullptr && !traced) { refl_dir[i * 3] = 0.f; refl_dir[i * 3 + 1] = 0.f; refl_dir[i * 3 + 2] = 0.f; }
}

__global__ void k_set_int(int* p, int v) { *p = v; }

struct Bump {
  uint8_t* base = nullptr;
  size_t off = 0, peak = 0, cap = 0;
  template <class T>
  T* take(size_t count) {
    const size_t o = off;
    off += align256(count * sizeof(T));
    peak = std::max(peak, off);
    return base == nullptr ? nullptr : reinterpret_cast<T*>(base + o);
  }
};

struct Ctx {
  const mnrf_field *coarse, *fine;
  const mnrf_trace_cfg* cfg;
  mnrf_level_cfg lc;
  const float *z_steps, *u_det, *noise0;
  const mnrf_trace_out* out;
  cudaStream_t st;
  int n;            // primary rays
  long long slab;   // rows per child batch
  bool dry;         // plan only: walk the allocation pattern, launch nothing
  bool second, has_mask, has_pred_normal;
  Bump bump;
  // per-sample scratch shared by every level (levels run one after the other on the stream), sized for `rows_max` rays
  long long rows_max;
  float *z_c, *w_c, *op_c, *z_f, *w_f, *nrm_c, *nrm_f;
  void* level_ws;
  int64_t level_ws_bytes;
  uint32_t node;    // running id of the (level batch) being processed: Philox stream = node * 64 + t
};

long long pow_ll(long long b, int e) { long long r = 1; while (e-- > 0) r *= b; return r; }

int rec(Ctx& C, int level, const float* rays, const int* cnt_dev, long long cap, float* rgb_io, float* depth_io, bool root) {
  const size_t mark = C.bump.off;
  const int T = C.cfg->normal_noise_std > 0.f ? C.cfg->trace_ray_times : 0;
  const mnrf_trace_out& O = *C.out;
  // ---- this level's per-ray results ----
  float* base_rgb = root ? (O.rgb_direct != nullptr ? O.rgb_direct : C.bump.take<float>(cap * 3)) : rgb_io;
  float* depth = root ? O.depth : depth_io;
  float* opacity = root ? O.opacity : C.bump.take<float>(cap);
  float* mask = !C.has_mask ? nullptr : (root && O.mirror_mask != nullptr ? O.mirror_mask : C.bump.take<float>(cap));
  float* normal = root && O.surface_normal != nullptr ? O.surface_normal : C.bump.take<float>(cap * 3);
  float* xs = root && O.x_surface != nullptr ? O.x_surface : C.bump.take<float>(cap * 3);
  if (!C.dry) {
    mnrf_level_out lo;
    memset(&lo, 0, sizeof(lo));
    lo.z_coarse = C.z_c; lo.coarse.weights = C.w_c; lo.coarse.opacity = C.op_c;
    mnrf_composite_out& last = C.second ? lo.fine : lo.coarse;
    if (C.second) { lo.z_fine = C.z_f; lo.fine.weights = C.w_f; }   // w_f == NULL: the pass composites inside the field kernel
    last.opacity = opacity; last.rgb = base_rgb; last.depth = depth; last.mirror_mask = mask; last.x_surface = xs;
    if (C.has_pred_normal) last.surface_normal = normal; else last.surface_normal_grad = normal;   // eval.py:337-360
    if (C.lc.compute_normal) { lo.normal_coarse = C.nrm_c; lo.normal_fine = C.nrm_f; }
    if (render_level(C.coarse, C.fine, rays, (int)cap, &C.lc, nullptr, C.z_steps, C.u_det, C.level_ws, C.level_ws_bytes, &lo,
                     C.st, cnt_dev))
      return 1;
  }
  const bool trace = C.has_mask && level < C.cfg->max_recursive_level;
  const unsigned grid = (unsigned)((cap + 255) / 256);
  if (!trace) {
    if (root && !C.dry) {
      // nothing below: the returned mask is still hard-clipped (eval.py:305-306), colour = direct colour, reflect outputs = 0
      if (C.has_mask && launch_reflect(rays, xs, normal, mask, (int)cap, 0.1f, nullptr, nullptr, nullptr, C.st, cnt_dev)) return 1;
      if (O.rgb != base_rgb) MNRF_CUDA_OK(cudaMemcpyAsync(O.rgb, base_rgb, sizeof(float) * cap * 3, cudaMemcpyDeviceToDevice, C.st));
      if (O.rgb_reflect) MNRF_CUDA_OK(cudaMemsetAsync(O.rgb_reflect, 0, sizeof(float) * cap * 3, C.st));
      if (O.depth_reflect) MNRF_CUDA_OK(cudaMemsetAsync(O.depth_reflect, 0, sizeof(float) * cap, C.st));
      if (O.reflect_direction) MNRF_CUDA_OK(cudaMemsetAsync(O.reflect_direction, 0, sizeof(float) * cap * 3, C.st));
    }
    C.bump.off = mark;
    return 0;
  }
  // ---- children ----
  const int compact0 = (C.cfg->only_trace_rays_in_mirrors == 1 || level >= 1) ? 1 : 0;   // eval.py:159
  const long long child_rows = (long long)(T + 1) * cap;
  const long long slab = std::min(child_rows, C.slab);
  const int n_slabs = (int)((child_rows + slab - 1) / slab);
  int* flag = C.bump.take<int>(1);
  int* ccount = C.bump.take<int>(1);
  int* index = C.bump.take<int>(cap);
  int* scan = C.bump.take<int>((cap + 1023) / 1024);
  int* layout = C.bump.take<int>(4 + n_slabs);
  float* child_rays = C.bump.take<float>(child_rows * 8);
  float* child_rgb = C.bump.take<float>(child_rows * 3);
  float* child_depth = C.bump.take<float>(child_rows);
  const uint32_t node = C.node++;
  if (!C.dry) {
    if (launch_reflect(rays, xs, normal, mask, (int)cap, 0.1f, nullptr, nullptr, flag, C.st, cnt_dev)) return 1;  // clip + any()
    if (launch_compact(nullptr, mask, (int)cap, 8, nullptr, index, ccount, C.st, cnt_dev, scan)) return 1;
    k_child_layout<<<1, 1, 0, C.st>>>(flag, cnt_dev, (int)cap, ccount, compact0, T, (int)slab, n_slabs, layout,
                                      O.level_rays != nullptr ? O.level_rays + level + 1 : nullptr);
    MNRF_LAUNCH_OK();
    for (int t = 0; t <= T; ++t) {
      const float* nz = (root && C.noise0 != nullptr) ? C.noise0 + (size_t)t * cap * 3 : nullptr;
      k_reflect_scatter<<<grid, 256, 0, C.st>>>(rays, xs, normal, index, layout, cnt_dev, (int)cap, t, compact0, nz,
                                                C.cfg->normal_noise_std, (unsigned long long)C.cfg->noise_seed, node * 64u + (uint32_t)t,
                                                0.1f, child_rays, (t == 0 && root) ? O.reflect_direction : nullptr);
      MNRF_LAUNCH_OK();
    }
  }
  for (int s = 0; s < n_slabs; ++s) {
    const long long lo_row = (long long)s * slab;
    const long long rows = std::min(slab, child_rows - lo_row);
    if (rec(C, level + 1, child_rays == nullptr ? nullptr : child_rays + lo_row * 8, layout == nullptr ? nullptr : layout + 4 + s, rows,
            child_rgb == nullptr ? nullptr : child_rgb + lo_row * 3, child_depth == nullptr ? nullptr : child_depth + lo_row, false))
      return 1;
  }
  if (!C.dry) {
    k_gather_blend<<<grid, 256, 0, C.st>>>(base_rgb, mask, index, layout, child_rgb, child_depth, cnt_dev, (int)cap, T, compact0,
                                           root ? O.rgb : rgb_io, root ? O.rgb_reflect : nullptr, root ? O.depth_reflect : nullptr,
                                           root ? O.reflect_direction : nullptr);
    MNRF_LAUNCH_OK();
  }
  C.bump.off = mark;
  return 0;
}

// fills the fixed part of the context and the shared per-sample scratch; returns the total bytes for slab size `slab`
int64_t plan(Ctx& C, long long slab) {
  C.slab = slab;
  const int T = C.cfg->normal_noise_std > 0.f ? C.cfg->trace_ray_times : 0;
  const long long worst = pow_ll(T + 1, std::max(C.cfg->max_recursive_level, 0)) * C.n;
  C.rows_max = C.cfg->max_recursive_level > 0 && C.has_mask ? std::max<long long>(C.n, std::min(worst, slab)) : C.n;
  if ((long long)C.rows_max * (C.lc.n_samples + C.lc.n_importance) >= (1ll << 31)) return -2;
  const size_t R = (size_t)C.rows_max, Sc = C.lc.n_samples, Sf = Sc + C.lc.n_importance;
  C.bump.off = 0; C.bump.peak = 0;
  C.z_c = C.bump.take<float>(R * Sc); C.w_c = C.bump.take<float>(R * Sc); C.op_c = C.bump.take<float>(R);
  C.z_f = C.second ? C.bump.take<float>(R * Sf) : nullptr;
  {
    const mnrf_field* last = C.lc.rerun_coarse_on_fine ? C.coarse : C.fine;
    const bool fused = C.second && last != nullptr && can_fuse_composite(last, &C.lc, nullptr, (int)Sf);
    C.w_f = C.second && !fused ? C.bump.take<float>(R * Sf) : nullptr;   // per-sample weights only for the unfused compositor
  }
  C.nrm_c = C.lc.compute_normal && !(C.lc.test_time && C.fine != nullptr) ? C.bump.take<float>(R * Sc * 3) : nullptr;
  C.nrm_f = C.lc.compute_normal && C.second ? C.bump.take<float>(R * Sf * 3) : nullptr;
  C.level_ws_bytes = mnrf_level_workspace_bytes_for(C.coarse, C.fine, (int)C.rows_max, &C.lc, 0);
  C.level_ws = C.bump.take<uint8_t>((size_t)C.level_ws_bytes);
  C.node = 0;
  const bool was_dry = C.dry;
  C.dry = true;
  rec(C, 0, nullptr, nullptr, C.n, nullptr, nullptr, true);
  C.dry = was_dry;
  return (int64_t)C.bump.peak;
}

int setup(Ctx& C, const mnrf_field* coarse, const mnrf_field* fine, int n, const mnrf_trace_cfg* cfg, const mnrf_trace_out* out) {
  MNRF_REQUIRE(coarse != nullptr && cfg != nullptr && n >= 0, "render_recursive: bad argument");
  MNRF_REQUIRE(cfg->max_recursive_level >= 0 && cfg->max_recursive_level <= 4, "render_recursive: 0 <= max_recursive_level <= 4");
  MNRF_REQUIRE(cfg->trace_ray_times >= 0 && cfg->trace_ray_times <= 63, "render_recursive: 0 <= trace_ray_times <= 63");
  MNRF_REQUIRE(cfg->only_trace_rays_in_mirrors == -1 || cfg->only_trace_rays_in_mirrors == 1,
               "render_recursive: only_trace_rays_in_mirrors must be -1 (eval.py) or 1");
  MNRF_REQUIRE(cfg->level.perturb == 0.f && cfg->level.noise_std == 0.f, "render_recursive: eval semantics (perturb = noise_std = 0)");
  C.coarse = coarse; C.fine = fine; C.cfg = cfg; C.lc = cfg->level; C.out = out; C.n = n;
  const mnrf_field* second = C.lc.rerun_coarse_on_fine ? coarse : fine;
  C.second = C.lc.n_importance > 0 && second != nullptr;
  if (!C.second) C.lc.n_importance = 0;
  const mnrf_field* last = C.second ? second : coarse;
  MNRF_REQUIRE(C.second || !(C.lc.test_time && fine != nullptr), "render_recursive: a sigma-only coarse pass needs a second pass");
  C.has_mask = last->has_mirror != 0;
  C.has_pred_normal = last->has_normal != 0;
  MNRF_REQUIRE(C.has_pred_normal || C.lc.compute_normal || !C.has_mask || cfg->max_recursive_level == 0,
               "render_recursive: a field without normal_net needs compute_normal = 1 to reflect (eval.py:351-360)");
  if (C.has_pred_normal) C.lc.compute_normal = 0;   // the analytic normal would not be used (eval.py:146-148)
  C.dry = true;
  return 0;
}

}  // namespace
}  // namespace mnrf

using namespace mnrf;

extern "C" {

int64_t mnrf_recursive_workspace_bytes(const mnrf_field* coarse, const mnrf_field* fine, int n, const mnrf_trace_cfg* cfg,
                                       int64_t budget_bytes) {
  Ctx C{};
  static const mnrf_trace_out no_out{};
  if (setup(C, coarse, fine, n, cfg, &no_out)) return -1;
  if (n == 0) return 256;
  const int T = cfg->normal_noise_std > 0.f ? cfg->trace_ray_times : 0;
  // slab candidates: (T+1)^k * n rows, largest first; the smallest (k = 0) is always accepted
  for (int k = std::max(cfg->max_recursive_level, 0); k >= 0; --k) {
    const int64_t b = plan(C, pow_ll(T + 1, k) * n);
    if (b > 0 && (budget_bytes <= 0 || b <= budget_bytes || k == 0)) return b;
    if (b == -2 && k == 0) { set_error("render_recursive: too many points per batch; split the ray batch"); return -1; }
  }
  return -1;
}

int mnrf_render_recursive(const mnrf_field* coarse, const mnrf_field* fine, const float* rays, int n,
                          const mnrf_trace_cfg* cfg, const float* z_steps, const float* u_det,
                          const float* level0_normal_noise, void* workspace, int64_t workspace_bytes,
                          const mnrf_trace_out* out, void* stream) {
  if (n == 0) return 0;
  MNRF_REQUIRE(rays && z_steps && out && workspace, "render_recursive: null argument");
  MNRF_REQUIRE(out->rgb && out->depth && out->opacity, "render_recursive: rgb, depth and opacity outputs are required");
  Ctx C{};
  if (setup(C, coarse, fine, n, cfg, out)) return 2;
  if (n == 0) return 0;
  MNRF_REQUIRE(C.lc.n_importance == 0 || u_det != nullptr, "render_recursive: u_det missing");
  const int T = cfg->normal_noise_std > 0.f ? cfg->trace_ray_times : 0;
  int64_t need = -1;
  for (int k = std::max(cfg->max_recursive_level, 0); k >= 0; --k) {   // the same choice as mnrf_recursive_workspace_bytes
    need = plan(C, pow_ll(T + 1, k) * n);
    if (need > 0 && need <= workspace_bytes) break;
    need = -1;
  }
  MNRF_REQUIRE(need > 0, "render_recursive: workspace too small (%lld bytes)", (long long)workspace_bytes);
  C.st = reinterpret_cast<cudaStream_t>(stream);
  C.z_steps = z_steps; C.u_det = u_det; C.noise0 = level0_normal_noise;
  C.bump.base = reinterpret_cast<uint8_t*>(workspace);
  plan(C, C.slab);   // again with real addresses for the shared per-sample scratch
  C.dry = false;
  C.node = 0;
  if (out->level_rays != nullptr) {
    MNRF_CUDA_OK(cudaMemsetAsync(out->level_rays, 0, sizeof(int) * (cfg->max_recursive_level + 1), C.st));
    k_set_int<<<1, 1, 0, C.st>>>(out->level_rays, n);
    MNRF_LAUNCH_OK();
  }
  return rec(C, 0, rays, nullptr, n, nullptr, nullptr, true);
}

}  // extern "C"
