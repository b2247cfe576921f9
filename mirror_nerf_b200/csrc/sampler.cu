// Ray samplers: stratified coarse depths (R/models/rendering.py:271-300), positional embedding
// (mirror_nerf.py:21-38), searchsorted and inverse-CDF resampling + sort-merge (rendering.py:7-51,312-326).
//
// All index/depth arithmetic uses separately rounded fp32 operations (__fmul_rn/__fadd_rn/__fdiv_rn: no FMA
// contraction) in the reference's operation order, so coarse depths and -- given the same weights -- CDFs,
// bin indices and resampled depths are bit-identical to the CPU reference.
#include "common.cuh"

namespace mnrf {
namespace {

__device__ __forceinline__ float coarse_z_at(float near, float far, float t, int use_disp) {
  const float omt = __fsub_rn(1.f, t);
  if (!use_disp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
  const float a = __fmul_rn(__fdiv_rn(1.f, near), omt);
  const float b = __fmul_rn(__fdiv_rn(1.f, far), t);
  return __fdiv_rn(1.f, __fadd_rn(a, b));
}

__global__ void k_coarse_z(const float* __restrict__ rays, int n, const float* __restrict__ z_steps, int S,
                           int use_disp, float perturb, const float* __restrict__ u, float* __restrict__ z_out,
                           const int* __restrict__ n_dev) {
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * S) return;
  int r = (int)(i / S), s = (int)(i % S);
  const float near = rays[(size_t)r * 8 + 6], far = rays[(size_t)r * 8 + 7];
  float z = coarse_z_at(near, far, z_steps[s], use_disp);
  if (perturb > 0.f) {
    // lower = [z0, mids], upper = [mids, z_last]; z = lower + (upper - lower) * (perturb * U)
    float lower = z, upper = z;
    if (s > 0) lower = __fmul_rn(0.5f, __fadd_rn(coarse_z_at(near, far, z_steps[s - 1], use_disp), z));
    if (s < S - 1) upper = __fmul_rn(0.5f, __fadd_rn(z, coarse_z_at(near, far, z_steps[s + 1], use_disp)));
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __fmul_rn(perturb, u[i])));
  }
  z_out[i] = z;
}

__global__ void k_embed(const float* __restrict__ x, int n, int n_freqs, float* __restrict__ out) {
  const int C = 3 + 6 * n_freqs;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * C) return;
  int r = (int)(i / C), k = (int)(i % C);
  float v;
  if (k < 3) v = x[(size_t)r * 3 + k];
  else {
    int e = k - 3, f = e / 6, q = e % 6, c = q % 3;
    float a = ldexpf(x[(size_t)r * 3 + c], f);
    v = q < 3 ? sinf(a) : cosf(a);
  }
  out[i] = v;
}

// pinhole camera rays (R/datasets/ray_utils.py:6-53 + R/datasets/blender.py:158-168): pixel (i = column, j = row),
// direction [(i - W/2)/f, -(j - H/2)/f, -1] (no +0.5), rotated by c2w[:, :3], normalised; origin c2w[:, 3]
struct Pose { float m[12]; };
__global__ void k_generate_rays(int H, int W, float focal, Pose c2w, float near, float far, float* __restrict__ rays) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const int j = idx / W, i = idx % W;
  const float dx = __fdiv_rn(__fsub_rn((float)i, W * 0.5f), focal);
  const float dy = -__fdiv_rn(__fsub_rn((float)j, H * 0.5f), focal);
  const float dz = -1.f;
  float r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    r[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w.m[4 * k + 0]), __fmul_rn(dy, c2w.m[4 * k + 1])), __fmul_rn(dz, c2w.m[4 * k + 2]));
  const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r[0], r[0]), __fmul_rn(r[1], r[1])), __fmul_rn(r[2], r[2])));
  float* o = rays + (size_t)idx * 8;
  o[0] = c2w.m[3]; o[1] = c2w.m[7]; o[2] = c2w.m[11];
  o[3] = __fdiv_rn(r[0], n); o[4] = __fdiv_rn(r[1], n); o[5] = __fdiv_rn(r[2], n);
  o[6] = near; o[7] = far;
}

// count of cdf entries <= u  (torch.searchsorted(..., right=True)) on a sorted row of length m
__device__ __forceinline__ int upper_bound(const float* cdf, int m, float u) {
  int lo = 0, hi = m;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void k_searchsorted(const float* __restrict__ cdf, int n, int m, const float* __restrict__ u, int n_u,
                               int u_stride, int64_t* __restrict__ inds) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * n_u) return;
  int r = (int)(i / n_u), j = (int)(i % n_u);
  inds[i] = upper_bound(cdf + (size_t)r * m, m, u[(size_t)r * u_stride + j]);
}

// ---- sample_pdf: one warp per ray ------------------------------------------------------------------
constexpr int SP_WARPS = 4;
constexpr int SP_MAXS = 256;   // coarse samples
constexpr int SP_MAXT = 512;   // coarse + importance

// Row sum in the order ATen's CPU sum kernel uses for a contiguous fp32 row (8-lane vectors, 4 interleaved
// accumulators, scalar tail, then the 8 lanes sequentially) -- pinned against tests/golden/sample_pdf.npz.
__device__ float row_sum_aten(const float* w, int m, int lane) {
  const int V = 8;
  const int vec_size = m / V;
  const int size_ilp = vec_size / 4;
  float p0 = 0.f;
  if (lane < V) {
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < size_ilp; ++i)
      for (int k = 0; k < 4; ++k) part[k] = __fadd_rn(part[k], w[(i * 4 + k) * V + lane]);
    for (int i = size_ilp * 4; i < vec_size; ++i) part[0] = __fadd_rn(part[0], w[i * V + lane]);
    for (int k = 1; k < 4; ++k) part[0] = __fadd_rn(part[0], part[k]);
    p0 = part[0];
  }
  float acc = 0.f;
  for (int k = vec_size * V; k < m; ++k) acc = __fadd_rn(acc, w[k]);  // same on every lane
  for (int k = 0; k < V; ++k) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, p0, k));
  return acc;
}

__global__ void __launch_bounds__(SP_WARPS * 32)
k_sample_pdf(const float* __restrict__ z_coarse, const float* __restrict__ bins_in, const float* __restrict__ weights,
             int w_stride, int w_off, int n, int S, int n_imp,
             const float* __restrict__ u, int u_stride, float* __restrict__ z_fine, float* __restrict__ samples_out,
             int64_t* __restrict__ inds_out, float* __restrict__ cdf_out, const int* __restrict__ n_dev) {
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  __shared__ float s_bins[SP_WARPS][SP_MAXS];
  __shared__ float s_cdf[SP_WARPS][SP_MAXS];
  __shared__ float s_w[SP_WARPS][SP_MAXS];
  __shared__ float s_z[SP_WARPS][SP_MAXT];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * SP_WARPS + wib;
  if (r >= n) return;  // whole warp exits together
  float* bins = s_bins[wib];
  float* cdf = s_cdf[wib];
  float* w = s_w[wib];
  float* zs = s_z[wib];
  const int nb = S - 1;  // bins (mid-points)
  const int nw = S - 2;  // weights used: weights[:, 1:-1]
  const int T = S + n_imp;
  int Tpad = 1;
  while (Tpad < T) Tpad <<= 1;

  if (z_coarse != nullptr)
    for (int i = lane; i < S; i += 32) zs[i] = z_coarse[(size_t)r * S + i];
  for (int i = lane; i < nw; i += 32) w[i] = __fadd_rn(weights[(size_t)r * w_stride + w_off + i], 1e-5f);
  __syncwarp();
  if (bins_in != nullptr) {
    for (int i = lane; i < nb; i += 32) bins[i] = bins_in[(size_t)r * nb + i];
  } else {
    for (int i = lane; i < nb; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(zs[i], zs[i + 1]));
  }
  const float total = row_sum_aten(w, nw, lane);
  // cdf = [0, cumsum(pdf)]; ATen's CPU cumsum accumulates fp32 rows in double and rounds each prefix;
  // the double partial sums are exact here, so a warp scan gives the same bits as the sequential loop.
  {
    const int per = (nw + 31) / 32;
    const int i0 = lane * per;
    double loc = 0.0;
    for (int i = i0; i < min(i0 + per, nw); ++i) loc += (double)__fdiv_rn(w[i], total);
    double incl = loc;
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    double run = incl - loc;
    for (int i = i0; i < min(i0 + per, nw); ++i) {
      run += (double)__fdiv_rn(w[i], total);
      cdf[i + 1] = (float)run;
    }
    if (lane == 0) cdf[0] = 0.f;
  }
  __syncwarp();
  if (cdf_out != nullptr)
    for (int i = lane; i < nb; i += 32) cdf_out[(size_t)r * nb + i] = cdf[i];

  for (int j = lane; j < n_imp; j += 32) {
    const float uu = u[(size_t)r * u_stride + j];
    const int ind = upper_bound(cdf, nb, uu);
    const int below = max(ind - 1, 0), above = min(ind, nw);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = bins[below], b1 = bins[above];
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.f;
    const float smp = __fadd_rn(b0, __fmul_rn(__fdiv_rn(__fsub_rn(uu, c0), denom), __fsub_rn(b1, b0)));
    zs[S + j] = smp;
    if (samples_out != nullptr) samples_out[(size_t)r * n_imp + j] = smp;
    if (inds_out != nullptr) inds_out[(size_t)r * n_imp + j] = ind;
  }
  if (z_fine == nullptr) return;  // standalone sample_pdf(bins, weights): no merge with coarse depths
  for (int i = T + lane; i < Tpad; i += 32) zs[i] = __int_as_float(0x7f800000);  // +inf padding
  __syncwarp();
  // torch.sort(cat([z_coarse, samples]))[0]: bitonic network (values only, so stability is irrelevant)
  for (int k = 2; k <= Tpad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (Tpad >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const bool up = (i & k) == 0;
        const float a = zs[i], b = zs[p];
        if ((a > b) == up) { zs[i] = b; zs[p] = a; }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < T; i += 32) z_fine[(size_t)r * T + i] = zs[i];
}

}  // namespace

int launch_coarse_z(const float* rays, int n, const float* z_steps, int S, int use_disp, float perturb,
                    const float* u, float* z_out, cudaStream_t st, const int* n_dev) {
  if (n <= 0) return 0;
  MNRF_REQUIRE(S >= 1, "coarse_z: S must be >= 1");
  MNRF_REQUIRE(!(perturb > 0.f) || u != nullptr, "coarse_z: perturb > 0 needs perturb_u");
  long long total = (long long)n * S;
  k_coarse_z<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rays, n, z_steps, S, use_disp, perturb, u, z_out, n_dev);
  MNRF_LAUNCH_OK();
  return 0;
}

int launch_generate_rays(int H, int W, float focal, const float* c2w_host, float near, float far, float* rays,
                         cudaStream_t st) {
  if (H <= 0 || W <= 0) return 0;
  Pose p;
  for (int i = 0; i < 12; ++i) p.m[i] = c2w_host[i];
  k_generate_rays<<<(H * W + 255) / 256, 256, 0, st>>>(H, W, focal, p, near, far, rays);
  MNRF_LAUNCH_OK();
  return 0;
}

int launch_embed(const float* x, int n, int n_freqs, float* out, cudaStream_t st) {
  if (n <= 0) return 0;
  long long total = (long long)n * (3 + 6 * n_freqs);
  k_embed<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, n, n_freqs, out);
  MNRF_LAUNCH_OK();
  return 0;
}

int launch_searchsorted(const float* cdf, int n, int m, const float* u, int n_u, int u_stride, int64_t* inds,
                        cudaStream_t st) {
  if (n <= 0 || n_u <= 0) return 0;
  long long total = (long long)n * n_u;
  k_searchsorted<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cdf, n, m, u, n_u, u_stride, inds);
  MNRF_LAUNCH_OK();
  return 0;
}

int launch_sample_pdf(const float* z_coarse, const float* bins, const float* weights, int w_stride, int w_off, int n,
                      int S, int n_imp, const float* u, int u_stride, float* z_fine, float* samples, int64_t* inds,
                      float* cdf, cudaStream_t st, const int* n_dev) {
  if (n <= 0) return 0;
  MNRF_REQUIRE(S >= 3 && S <= SP_MAXS, "sample_pdf: need 3 <= N_samples <= %d (got %d)", SP_MAXS, S);
  MNRF_REQUIRE(n_imp >= 1 && S + n_imp <= SP_MAXT, "sample_pdf: need N_samples + N_importance <= %d", SP_MAXT);
  k_sample_pdf<<<(n + SP_WARPS - 1) / SP_WARPS, SP_WARPS * 32, 0, st>>>(z_coarse, bins, weights, w_stride, w_off, n, S, n_imp, u,
                                                                      u_stride, z_fine, samples, inds, cdf, n_dev);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace mnrf
