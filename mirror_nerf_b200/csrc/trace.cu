// Whitted-bounce helpers: what the callers of render_rays do between levels, moved on-device so that the
// recursion needs no boolean-mask indexing on the host path.
//   reflect   R/eval.py:295-307,515-540 == R/train.py:159-166,219-243
//   compact   R/eval.py:545-548 (secondary_rays[mirror_mask]) -- stable, same order as boolean indexing
//   blend     R/eval.py:676-697
#include "common.cuh"

namespace mnrf {
namespace {

__global__ void k_reflect(const float* __restrict__ rays, const float* __restrict__ x_surface,
                          const float* __restrict__ normal, float* __restrict__ mask, int n, float near2,
                          float* __restrict__ sec, float* __restrict__ refl, int* __restrict__ any_mirror,
                          const int* __restrict__ n_dev, const float* __restrict__ jitter, float jitter_scale,
                          int threshold_mask) {
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false;
  if (i < n) {
    float m = mask[i];
    if (threshold_mask) {
      if (m > 0.5f) m = 1.f; else if (m < 0.5f) m = 0.f;  // hard clip in place; exactly 0.5 is kept
      mask[i] = m;
    }
    hit = m != 0.f;
    const float* ry = rays + (size_t)i * 8;
    float nx = normal[i * 3], ny = normal[i * 3 + 1], nz = normal[i * 3 + 2];
    if (jitter != nullptr) {  // roughness cone: normal + N(0, std^2) noise, separately rounded (R/eval.py:506-511)
      nx = __fadd_rn(nx, __fmul_rn(jitter[i * 3], jitter_scale));
      ny = __fadd_rn(ny, __fmul_rn(jitter[i * 3 + 1], jitter_scale));
      nz = __fadd_rn(nz, __fmul_rn(jitter[i * 3 + 2], jitter_scale));
    }
    float nn = sqrtf(fmaxf(nx * nx + ny * ny + nz * nz, FP32_EPS));  // utils/func.py:5-7
    nx /= nn; ny /= nn; nz /= nn;
    float wx = -ry[3], wy = -ry[4], wz = -ry[5];
    float wn = sqrtf(fmaxf(wx * wx + wy * wy + wz * wz, FP32_EPS));
    wx /= wn; wy /= wn; wz /= wn;
    const float c2 = 2.f * (wx * nx + wy * ny + wz * nz);
    const float rx = c2 * nx - wx, ryy = c2 * ny - wy, rz = c2 * nz - wz;  // 2 (n.w) n - w
    if (sec != nullptr) {
      float* o = sec + (size_t)i * 8;
      o[0] = x_surface[i * 3]; o[1] = x_surface[i * 3 + 1]; o[2] = x_surface[i * 3 + 2];
      o[3] = rx; o[4] = ryy; o[5] = rz;
      o[6] = near2;   // ray_forward_offset = 0.1 in the reference
      o[7] = ry[7];   // parent far
    }
    if (refl != nullptr) { refl[i * 3] = rx; refl[i * 3 + 1] = ryy; refl[i * 3 + 2] = rz; }
  }
  if (any_mirror != nullptr && __any_sync(0xffffffffu, hit) && (threadIdx.x & 31) == 0) atomicOr(any_mirror, 1);
}

// ---- stable compaction: block counts -> exclusive scan -> scatter ----------------------------------------
constexpr int CB = 1024;

__global__ void k_count(const float* __restrict__ mask, int n, int* __restrict__ block_counts, const int* __restrict__ n_dev) {
  __shared__ int wsum[32];
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  int i = blockIdx.x * CB + threadIdx.x;
  bool f = i < n && mask[i] != 0.f;
  unsigned b = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wsum[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = v;
  }
}

// single block: exclusive scan of block_counts in place, total -> *count
__global__ void k_scan_blocks(int* __restrict__ block_counts, int nb, int* __restrict__ count) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += CB) {
    int i = base + threadIdx.x;
    int v = i < nb ? block_counts[i] : 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      int s = wsum[lane];
      int si = s;
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
      wsum[lane] = si - s;  // exclusive warp offsets
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + wsum[w] + incl - v;
    if (i < nb) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == CB - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry_s;
}

__global__ void k_scatter(const float* __restrict__ in, const float* __restrict__ mask, int n, int row_floats,
                          const int* __restrict__ block_offsets, float* __restrict__ out, int* __restrict__ index,
                          const int* __restrict__ n_dev) {
  __shared__ int wsum[32];
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  int i = blockIdx.x * CB + threadIdx.x;
  bool f = i < n && mask[i] != 0.f;
  unsigned b = __ballot_sync(0xffffffffu, f);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wsum[w] = __popc(b);
  __syncthreads();
  if (w == 0) {
    int s = wsum[lane], si = s;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
    wsum[lane] = si - s;
  }
  __syncthreads();
  int dst = block_offsets[blockIdx.x] + wsum[w] + __popc(b & ((1u << lane) - 1u));
  if (i < n) {
    if (index != nullptr) index[i] = f ? dst : -1;
    if (f && out != nullptr)
      for (int c = 0; c < row_floats; ++c) out[(size_t)dst * row_floats + c] = in[(size_t)i * row_floats + c];
  }
}

__global__ void k_blend(const float* __restrict__ base, const float* __restrict__ mask,
                        const float* __restrict__ child_rgb, const float* __restrict__ child_depth,
                        const int* __restrict__ index, int n, float* __restrict__ rgb_out,
                        float* __restrict__ rgb_reflect, float* __restrict__ depth_reflect,
                        const int* __restrict__ n_dev, const int* __restrict__ traced) {
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (traced != nullptr && __ldg(traced) == 0) {  // no mirror pixel in the batch: the level below was not rendered (eval.py:311-320)
    for (int c = 0; c < 3; ++c) {
      rgb_out[i * 3 + c] = base[i * 3 + c];
      if (rgb_reflect != nullptr) rgb_reflect[i * 3 + c] = 0.f;
    }
    if (depth_reflect != nullptr) depth_reflect[i] = 0.f;
    return;
  }
  const float m = mask[i] != 0.f ? 1.f : 0.f;  // mirror_mask.bool().float() (eval.py:307,678)
  const int src = index != nullptr ? index[i] : i;
  float d = 0.f;
  for (int c = 0; c < 3; ++c) {
    const float b = base[i * 3 + c];
    // compacted child: rays outside the mirror keep the base colour as "reflection" (eval.py:684-688)
    const float rf = src >= 0 ? child_rgb[(size_t)src * 3 + c] : b;
    if (rgb_reflect != nullptr) rgb_reflect[i * 3 + c] = src >= 0 ? rf : 0.f;
    rgb_out[i * 3 + c] = m * rf + (1.f - m) * b;
  }
  if (depth_reflect != nullptr) {
    if (src >= 0 && child_depth != nullptr) d = child_depth[src];
    depth_reflect[i] = d;
  }
}

// dense[i] = alpha * dense[i] + beta * compact[index[i]]  for rows with index[i] >= 0 (compact may be NULL: beta term dropped)
__global__ void k_axpy_rows(float* __restrict__ dense, const float* __restrict__ compact, const int* __restrict__ index, int n,
                            int c, float alpha, float beta, const int* __restrict__ n_dev) {
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int src = index != nullptr ? index[i] : i;
  if (src < 0) return;
  for (int k = 0; k < c; ++k) {
    float v = alpha * dense[(size_t)i * c + k];
    if (compact != nullptr) v += beta * compact[(size_t)src * c + k];
    dense[(size_t)i * c + k] = v;
  }
}

__global__ void k_select_count(const int* __restrict__ flag, int n, const int* __restrict__ n_dev, int* __restrict__ out) {
  if (n_dev != nullptr) n = min(n, *n_dev);
  *out = (*flag != 0) ? n : 0;
}

}  // namespace

int launch_select_count(const int* flag, int n, const int* n_dev, int* out, cudaStream_t st) {
  k_select_count<<<1, 1, 0, st>>>(flag, n, n_dev, out);
  MNRF_LAUNCH_OK();
  return 0;
}

int launch_axpy_rows(float* dense, const float* compact, const int* index, int n, int c, float alpha, float beta,
                     cudaStream_t st, const int* n_dev) {
  if (n <= 0) return 0;
  k_axpy_rows<<<(n + 255) / 256, 256, 0, st>>>(dense, compact, index, n, c, alpha, beta, n_dev);
  MNRF_LAUNCH_OK();
  return 0;
}

int launch_reflect(const float* rays, const float* x_surface, const float* normal, float* mask, int n, float near2,
                   float* sec, float* refl, int* any_mirror, cudaStream_t st, const int* n_dev, const float* jitter,
                   float jitter_scale, int threshold_mask) {
  if (any_mirror != nullptr) MNRF_CUDA_OK(cudaMemsetAsync(any_mirror, 0, sizeof(int), st));
  if (n <= 0) return 0;
  k_reflect<<<(n + 255) / 256, 256, 0, st>>>(rays, x_surface, normal, mask, n, near2, sec, refl, any_mirror, n_dev, jitter,
                                             jitter_scale, threshold_mask);
  MNRF_LAUNCH_OK();
  return 0;
}

int launch_compact(const float* in, const float* mask, int n, int row_floats, float* out, int* index, int* count,
                   cudaStream_t st, const int* n_dev, int* scratch) {
  if (n <= 0) {
    MNRF_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int), st));
    return 0;
  }
  int nb = (n + CB - 1) / CB;
  int* block_counts = scratch;
  if (block_counts == nullptr) MNRF_CUDA_OK(cudaMallocAsync(&block_counts, sizeof(int) * nb, st));
  k_count<<<nb, CB, 0, st>>>(mask, n, block_counts, n_dev);
  MNRF_LAUNCH_OK();
  k_scan_blocks<<<1, CB, 0, st>>>(block_counts, nb, count);
  MNRF_LAUNCH_OK();
  k_scatter<<<nb, CB, 0, st>>>(in, mask, n, row_floats, block_counts, out, index, n_dev);
  MNRF_LAUNCH_OK();
  if (scratch == nullptr) MNRF_CUDA_OK(cudaFreeAsync(block_counts, st));
  return 0;
}

int launch_blend(const float* base, const float* mask, const float* child_rgb, const float* child_depth,
                 const int* index, int n, float* rgb_out, float* rgb_reflect, float* depth_reflect, cudaStream_t st,
                 const int* n_dev, const int* traced) {
  if (n <= 0) return 0;
  k_blend<<<(n + 255) / 256, 256, 0, st>>>(base, mask, child_rgb, child_depth, index, n, rgb_out, rgb_reflect,
                                           depth_reflect, n_dev, traced);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace mnrf
