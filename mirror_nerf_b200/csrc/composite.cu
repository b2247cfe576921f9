// Volume-rendering quadrature of one pass: R/models/rendering.py:175-264 (inference(), after the model call)
// plus x_surface (rendering.py:363-367).  One warp per ray; samples are lane-strided and the exclusive
// transmittance product T_i = prod_{j<i} (1 - alpha_j + 1e-10) is a warp inclusive scan per 32-sample block
// with a running carry (the CPU reference multiplies left to right; re-association is ~1e-7 relative).
#include "common.cuh"

namespace mnrf {
namespace {

constexpr int CW = 4;        // warps (rays) per block
constexpr int MAX_BLK = 16;  // up to 512 samples per ray

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(CW * 32)
k_composite(const float* __restrict__ rays, const float* __restrict__ z, const float* __restrict__ sigma,
            int sigma_stride, const float* __restrict__ raw, const float* __restrict__ normal,
            const float* __restrict__ noise, float noise_std, int n, int S, int white_back, mnrf_composite_out out,
            const int* __restrict__ n_dev) {
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * CW + (threadIdx.x >> 5);
  if (r >= n) return;
  const size_t base = (size_t)r * S;
  float carry = 1.f;  // product of all (1 - alpha + 1e-10) before the current 32-sample block
  float a_op = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f, a_d = 0.f, a_m = 0.f;
  float a_n0 = 0.f, a_n1 = 0.f, a_n2 = 0.f, a_g0 = 0.f, a_g1 = 0.f, a_g2 = 0.f, a_dif = 0.f;
  const int nblk = (S + 31) / 32;
  for (int b = 0; b < nblk; ++b) {
    const int s = b * 32 + lane;
    const bool ok = s < S;
    float zz = 0.f, alpha = 0.f;
    if (ok) {
      zz = z[base + s];
      const float delta = (s + 1 < S) ? __fsub_rn(z[base + s + 1], zz) : 1e10f;  // rendering.py:182-186
      float sg = sigma[(base + s) * sigma_stride];
      if (noise != nullptr) sg = __fadd_rn(sg, __fmul_rn(noise[base + s], noise_std));
      alpha = __fsub_rn(1.f, expf(-__fmul_rn(delta, fmaxf(sg, 0.f))));
    }
    const float f = ok ? __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f) : 1.f;
    float incl = f;
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= t;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    const float T = carry * excl;
    const float w = alpha * T;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    if (ok) {
      out.weights[base + s] = w;
      a_op += w;
      a_d = fmaf(w, zz, a_d);
      if (raw != nullptr) {
        const float4 q0 = *reinterpret_cast<const float4*>(raw + (base + s) * 8);
        const float4 q1 = *reinterpret_cast<const float4*>(raw + (base + s) * 8 + 4);
        a_r = fmaf(w, q0.y, a_r); a_g = fmaf(w, q0.z, a_g); a_b = fmaf(w, q0.w, a_b);
        a_m = fmaf(w, q1.x, a_m);
        a_n0 = fmaf(w, q1.y, a_n0); a_n1 = fmaf(w, q1.z, a_n1); a_n2 = fmaf(w, q1.w, a_n2);
        if (out.pred_normal != nullptr) {
          float* pn = out.pred_normal + (base + s) * 3;
          pn[0] = q1.y; pn[1] = q1.z; pn[2] = q1.w;
        }
        if (normal != nullptr) {
          const float* nn = normal + (base + s) * 3;
          const float e0 = nn[0] - q1.y, e1 = nn[1] - q1.z, e2 = nn[2] - q1.w;
          a_dif = fmaf(w, e0 * e0 + e1 * e1 + e2 * e2, a_dif);
        }
      }
      if (normal != nullptr) {
        const float* nn = normal + (base + s) * 3;
        a_g0 = fmaf(w, nn[0], a_g0); a_g1 = fmaf(w, nn[1], a_g1); a_g2 = fmaf(w, nn[2], a_g2);
      }
    }
  }
  a_op = warp_sum(a_op); a_d = warp_sum(a_d);
  a_r = warp_sum(a_r); a_g = warp_sum(a_g); a_b = warp_sum(a_b); a_m = warp_sum(a_m);
  a_n0 = warp_sum(a_n0); a_n1 = warp_sum(a_n1); a_n2 = warp_sum(a_n2);
  a_g0 = warp_sum(a_g0); a_g1 = warp_sum(a_g1); a_g2 = warp_sum(a_g2); a_dif = warp_sum(a_dif);
  if (lane == 0) {
    out.opacity[r] = a_op;
    if (white_back) { const float bg = 1.f - a_op; a_r += bg; a_g += bg; a_b += bg; }  // rendering.py:216-217
    if (out.rgb) { out.rgb[r * 3 + 0] = a_r; out.rgb[r * 3 + 1] = a_g; out.rgb[r * 3 + 2] = a_b; }
    if (out.depth) out.depth[r] = a_d;
    if (out.mirror_mask) out.mirror_mask[r] = a_m;
    if (out.surface_normal) { out.surface_normal[r * 3 + 0] = a_n0; out.surface_normal[r * 3 + 1] = a_n1; out.surface_normal[r * 3 + 2] = a_n2; }
    if (out.surface_normal_grad) { out.surface_normal_grad[r * 3 + 0] = a_g0; out.surface_normal_grad[r * 3 + 1] = a_g1; out.surface_normal_grad[r * 3 + 2] = a_g2; }
    if (out.normal_dif) out.normal_dif[r] = a_dif;
    if (out.x_surface) {
      const float* ry = rays + (size_t)r * 8;
      for (int c = 0; c < 3; ++c) out.x_surface[r * 3 + c] = __fadd_rn(ry[c], __fmul_rn(ry[3 + c], a_d));
    }
  }
}

}  // namespace

int launch_composite(const float* rays, const float* z, const float* sigma, int sigma_stride, const float* raw,
                     const float* normal, const float* noise, float noise_std, int n, int S, int white_back,
                     const mnrf_composite_out& out, cudaStream_t st, const int* n_dev) {
  if (n <= 0) return 0;
  MNRF_REQUIRE(S >= 1 && S <= 32 * MAX_BLK, "composite: 1 <= S <= %d", 32 * MAX_BLK);
  MNRF_REQUIRE(out.weights != nullptr && out.opacity != nullptr, "composite: weights/opacity outputs are required");
  MNRF_REQUIRE(raw != nullptr || (out.rgb == nullptr && out.mirror_mask == nullptr && out.pred_normal == nullptr &&
                                  out.surface_normal == nullptr && out.normal_dif == nullptr),
               "composite: colour/normal outputs need the raw field records");
  k_composite<<<(n + CW - 1) / CW, CW * 32, 0, st>>>(rays, z, sigma, sigma_stride, raw, normal, noise, noise_std, n, S,
                                                   white_back, out, n_dev);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace mnrf
