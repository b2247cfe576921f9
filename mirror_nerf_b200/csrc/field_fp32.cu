// fp32 CUDA-core field kernel: MirrorNeRF.forward (R/models/mirror_nerf.py:101-212) on 16-point tiles.
//
// Role: (1) the on-device verification twin of the tcgen05 kernel (same inputs/outputs, plain fp32 FMAs,
// un-folded normal head), (2) the only kernel that implements the analytic normal
// normalize(-d sigma / d xyz) (mirror_nerf.py:136-146, utils/func.py:10-25) -- done as an explicit
// reverse chain through the trunk (ReLU masks, skip split, PE Jacobian) instead of autograd.
// It is NOT the performance path.
#include "common.cuh"

namespace mnrf {
namespace {

constexpr int TP = 16;        // points per block
constexpr int NT = 256;       // threads per block

// acc[p] += sum_k in[p][k] * Wt[k][n]   (in: smem, row stride ld_in, K % 4 == 0; Wt: global [K][N])
template <int N>
__device__ __forceinline__ void dense_acc(float (&acc)[TP], const float* __restrict__ in, int ld_in, int K,
                                          const float* __restrict__ Wt, int n) {
  for (int k = 0; k < K; k += 4) {
    float w0 = __ldg(Wt + (k + 0) * N + n);
    float w1 = __ldg(Wt + (k + 1) * N + n);
    float w2 = __ldg(Wt + (k + 2) * N + n);
    float w3 = __ldg(Wt + (k + 3) * N + n);
#pragma unroll
    for (int p = 0; p < TP; ++p) {
      float4 a = *reinterpret_cast<const float4*>(in + p * ld_in + k);
      acc[p] = fmaf(a.x, w0, acc[p]);
      acc[p] = fmaf(a.y, w1, acc[p]);
      acc[p] = fmaf(a.z, w2, acc[p]);
      acc[p] = fmaf(a.w, w3, acc[p]);
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// smem carve (floats)
constexpr int SM_PE = 0;                       // [TP][64]
constexpr int SM_H = SM_PE + TP * 64;          // [8][TP][256]
constexpr int SM_T0 = SM_H + 8 * TP * W;       // [TP][256]   final feature / g ping
constexpr int SM_T1 = SM_T0 + TP * W;          // [TP][256]   head hidden  / g pong
constexpr int SM_GPE = SM_T1 + TP * W;         // [TP][64]
constexpr int SM_XYZ = SM_GPE + TP * 64;       // [TP][4]
constexpr int SM_TOTAL = SM_XYZ + TP * 4;

__global__ void __launch_bounds__(NT, 1)
k_field_fp32(const float* __restrict__ P, F32Layout L, FieldIO io, int has_normal, int has_mirror) {
  extern __shared__ float sm[];
  if (io.n_rays_dev != nullptr) {  // device-side ray count (mnrf_render_recursive)
    io.n_points = min(io.n_points, max(__ldg(io.n_rays_dev), 0) * io.S);
    if ((int)blockIdx.x * TP >= io.n_points) return;
  }
  float* pe = sm + SM_PE;
  float* H = sm + SM_H;
  float* t0 = sm + SM_T0;
  float* t1 = sm + SM_T1;
  float* gpe = sm + SM_GPE;
  float* xyz = sm + SM_XYZ;
  const int tid = threadIdx.x;
  const int p0 = blockIdx.x * TP;

  // ---- points + positional encoding (mirror_nerf.py:33-38) ----
  if (tid < TP) {
    int p = min(p0 + tid, io.n_points - 1);
    float x[3];
    if (io.rays != nullptr) {
      int r = p / io.S;
      float z = io.z[p];
      for (int c = 0; c < 3; ++c)
        x[c] = __fadd_rn(io.rays[r * 8 + c], __fmul_rn(io.rays[r * 8 + 3 + c], z));  // o + d*z, no FMA
    } else {
      for (int c = 0; c < 3; ++c) x[c] = io.x[(size_t)p * io.x_stride + c];
    }
    for (int c = 0; c < 3; ++c) xyz[tid * 4 + c] = x[c];
  }
  __syncthreads();
  for (int i = tid; i < TP * 64; i += NT) {
    int p = i >> 6, k = i & 63;
    float v = 0.f;
    if (k < 3) v = xyz[p * 4 + k];
    else if (k < IN_XYZ) {
      int e = k - 3, f = e / 6, r = e % 6, c = r % 3;
      float a = ldexpf(xyz[p * 4 + c], f);  // 2^f * x, exact
      v = (r < 3) ? sinf(a) : cosf(a);
    }
    pe[i] = v;
  }
  __syncthreads();

  // ---- trunk (mirror_nerf.py:189-197) ----
  {
    const int n = tid;
    for (int l = 0; l < 8; ++l) {
      float acc[TP];
      float b = P[L.b_trunk[l] + n];
#pragma unroll
      for (int p = 0; p < TP; ++p) acc[p] = b;
      const float* Wt = P + L.wt_trunk[l];
      if (l == 0) {
        dense_acc<W>(acc, pe, 64, 64, Wt, n);
      } else if (l == 4) {
        // input = [pe(63) | h4(256)]: the 64th pe entry is 0, so reading Wt row 63 there is harmless
        dense_acc<W>(acc, pe, 64, 64, Wt, n);
        dense_acc<W>(acc, H + 3 * TP * W, W, W, Wt + IN_XYZ * W, n);
      } else {
        dense_acc<W>(acc, H + (l - 1) * TP * W, W, W, Wt, n);
      }
      float* out = H + l * TP * W;
#pragma unroll
      for (int p = 0; p < TP; ++p) out[p * W + n] = fmaxf(acc[p], 0.f);
      __syncthreads();
    }
  }
  const float* h8 = H + 7 * TP * W;

  // ---- sigma (raw, mirror_nerf.py:195) : thread per point, 16 lanes per point ----
  float sigma_val = 0.f;
  {
    int p = tid >> 4, j = tid & 15;
    float s = 0.f;
    for (int k = j; k < W; k += 16) s = fmaf(h8[p * W + k], P[L.w_sigma + k], s);
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
    sigma_val = s + P[L.b_sigma];
    int gp = p0 + p;
    if (j == 0 && gp < io.n_points) {
      if (io.sigma_out) io.sigma_out[gp] = sigma_val;
      if (io.raw) io.raw[(size_t)gp * 8 + 0] = sigma_val;
    }
  }
  if (io.geo_out) {
    for (int i = tid; i < TP * W; i += NT) {
      int p = i / W;
      if (p0 + p < io.n_points) io.geo_out[(size_t)(p0 + p) * W + (i % W)] = h8[i];
    }
  }

  // ---- predicted normal: Linear(256,128) -> Linear(128,3), NO activation (mirror_nerf.py:85-88,206-208) ----
  // (computed by the reference even when sigma_only; we only need it when it is returned)
  if (has_normal && io.raw != nullptr) {
    if (tid < WH) {
      float acc[TP];
      float b = P[L.b_n0 + tid];
#pragma unroll
      for (int p = 0; p < TP; ++p) acc[p] = b;
      dense_acc<WH>(acc, h8, W, W, P + L.wt_n0, tid);
#pragma unroll
      for (int p = 0; p < TP; ++p) t1[p * WH + tid] = acc[p];
    }
    __syncthreads();
    if (tid < TP) {
      float v[3];
      for (int i = 0; i < 3; ++i) {
        float s = P[L.b_n1 + i];
        for (int k = 0; k < WH; ++k) s = fmaf(t1[tid * WH + k], P[L.w_n1 + i * WH + k], s);
        v[i] = s;
      }
      float nn = sqrtf(fmaxf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2], FP32_EPS));  // utils/func.py:5-7
      int gp = p0 + tid;
      if (gp < io.n_points)
        for (int i = 0; i < 3; ++i) io.raw[(size_t)gp * 8 + 5 + i] = v[i] / nn;
    }
    __syncthreads();
  } else if (io.raw != nullptr && tid < TP && p0 + tid < io.n_points) {
    for (int i = 0; i < 3; ++i) io.raw[(size_t)(p0 + tid) * 8 + 5 + i] = 0.f;
  }

  if (!io.sigma_only && io.raw != nullptr) {
    // ---- colour branch (mirror_nerf.py:199-204) ----
    {
      float acc[TP];
      float b = P[L.b_final + tid];
#pragma unroll
      for (int p = 0; p < TP; ++p) acc[p] = b;
      dense_acc<W>(acc, h8, W, W, P + L.wt_final, tid);
#pragma unroll
      for (int p = 0; p < TP; ++p) t0[p * W + tid] = acc[p];  // no activation
    }
    __syncthreads();
    if (tid < WH) {
      float acc[TP];
#pragma unroll
      for (int p = 0; p < TP; ++p) {
        int gp = min(p0 + p, io.n_points - 1);
        int r = (io.rays != nullptr) ? gp / io.S : gp;
        acc[p] = io.dirbias[(size_t)r * WH + tid];  // b_dir + W_dir[:,256:] . embed(dir)
      }
      dense_acc<WH>(acc, t0, W, W, P + L.wt_dir, tid);
#pragma unroll
      for (int p = 0; p < TP; ++p) t1[p * WH + tid] = fmaxf(acc[p], 0.f);
    }
    __syncthreads();
    if (tid < TP * 3) {
      int p = tid / 3, i = tid % 3;
      float s = P[L.b_rgb + i];
      for (int k = 0; k < WH; ++k) s = fmaf(t1[p * WH + k], P[L.w_rgb + i * WH + k], s);
      if (p0 + p < io.n_points) io.raw[(size_t)(p0 + p) * 8 + 1 + i] = sigmoidf_(s);
    }
    __syncthreads();
    // ---- mirror probability (mirror_nerf.py:94-99,210-212) ----
    if (has_mirror) {
      if (tid < WH) {
        float acc[TP];
        float b = P[L.b_m0 + tid];
#pragma unroll
        for (int p = 0; p < TP; ++p) acc[p] = b;
        dense_acc<WH>(acc, h8, W, W, P + L.wt_m0, tid);
#pragma unroll
        for (int p = 0; p < TP; ++p) t1[p * WH + tid] = acc[p] > 0.f ? acc[p] : 0.01f * acc[p];  // LeakyReLU
      }
      __syncthreads();
      if (tid < TP) {
        float s = P[L.b_m2];
        for (int k = 0; k < WH; ++k) s = fmaf(t1[tid * WH + k], P[L.w_m2 + k], s);
        if (p0 + tid < io.n_points) io.raw[(size_t)(p0 + tid) * 8 + 4] = sigmoidf_(s);
      }
      __syncthreads();
    } else if (tid < TP && p0 + tid < io.n_points) {
      io.raw[(size_t)(p0 + tid) * 8 + 4] = 0.f;
    }
  } else if (io.raw != nullptr && tid < TP && p0 + tid < io.n_points) {
    for (int i = 1; i < 5; ++i) io.raw[(size_t)(p0 + tid) * 8 + i] = 0.f;
  }

  // ---- analytic normal: reverse chain d sigma / d xyz (what autograd does for mirror_nerf.py:136-146) ----
  if (io.normal_out != nullptr) {
    __syncthreads();
    float* g = t0;
    float* gn = t1;
    for (int i = tid; i < TP * W; i += NT) g[i] = (h8[i] > 0.f) ? P[L.w_sigma + (i % W)] : 0.f;
    for (int i = tid; i < TP * 64; i += NT) gpe[i] = 0.f;
    __syncthreads();
    for (int l = 7; l >= 0; --l) {
      const int K = trunk_k(l);
      const float* Wl = P + L.w_trunk[l];  // [256][K]
      for (int k = tid; k < K; k += NT) {
        float acc[TP];
#pragma unroll
        for (int p = 0; p < TP; ++p) acc[p] = 0.f;
        for (int n = 0; n < W; ++n) {
          float w = __ldg(Wl + n * K + k);
#pragma unroll
          for (int p = 0; p < TP; ++p) acc[p] = fmaf(g[p * W + n], w, acc[p]);
        }
        if (l == 0) {
#pragma unroll
          for (int p = 0; p < TP; ++p) gpe[p * 64 + k] += acc[p];
        } else if (l == 4 && k < IN_XYZ) {
#pragma unroll
          for (int p = 0; p < TP; ++p) gpe[p * 64 + k] += acc[p];
        } else {
          int kk = (l == 4) ? k - IN_XYZ : k;
          const float* hprev = H + (l - 1) * TP * W;
#pragma unroll
          for (int p = 0; p < TP; ++p) gn[p * W + kk] = (hprev[p * W + kk] > 0.f) ? acc[p] : 0.f;
        }
      }
      __syncthreads();
      if (l > 0) { float* t = g; g = gn; gn = t; }
    }
    if (tid < TP * 3) {
      int p = tid / 3, c = tid % 3;
      float x = xyz[p * 4 + c];
      float gx = gpe[p * 64 + c];
      for (int f = 0; f < NFREQ_XYZ; ++f) {
        float fr = ldexpf(1.f, f);
        float a = fr * x;
        gx += fr * (gpe[p * 64 + 3 + 6 * f + c] * cosf(a) - gpe[p * 64 + 3 + 6 * f + 3 + c] * sinf(a));
      }
      gpe[p * 64 + c] = -gx;  // normal = normalize(-grad)
    }
    __syncthreads();
    if (tid < TP && p0 + tid < io.n_points) {
      float a = gpe[tid * 64 + 0], b = gpe[tid * 64 + 1], c = gpe[tid * 64 + 2];
      float nn = sqrtf(fmaxf(a * a + b * b + c * c, FP32_EPS));
      float* o = io.normal_out + (size_t)(p0 + tid) * 3;
      o[0] = a / nn; o[1] = b / nn; o[2] = c / nn;
    }
  }
}

// ---- dir-layer additive term ---------------------------------------------------------------------
// out[r][n] = b_dir[n] + sum_j W_dir[n][256+j] * embed(dir_r)[j]   (mirror_nerf.py:201-203, rendering.py:275-277)
__global__ void k_dirbias(const float* __restrict__ P, F32Layout L, const float* __restrict__ src, int n,
                          int src_stride, int from_embedded, float* __restrict__ out, const int* __restrict__ n_dev) {
  __shared__ float e[8][IN_DIR + 1];
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  int r0 = blockIdx.x * 8;
  if (r0 >= n) return;
  int tid = threadIdx.x;  // 128
  for (int i = tid; i < 8 * IN_DIR; i += 128) {
    int rr = i / IN_DIR, j = i % IN_DIR;
    int r = min(r0 + rr, n - 1);
    float v;
    if (from_embedded) v = src[(size_t)r * src_stride + 3 + j];
    else {
      const float* d = src + (size_t)r * src_stride + 3;
      if (j < 3) v = d[j];
      else {
        int q = j - 3, f = q / 6, s = q % 6, c = s % 3;
        float a = ldexpf(d[c], f);
        v = s < 3 ? sinf(a) : cosf(a);
      }
    }
    e[rr][j] = v;
  }
  __syncthreads();
  const float* Wt = P + L.wt_dir + W * WH;  // rows 256..282 of Wt_dir [284][128]
  float w[IN_DIR];
  for (int j = 0; j < IN_DIR; ++j) w[j] = Wt[j * WH + tid];
  float b = P[L.b_dir + tid];
  for (int rr = 0; rr < 8; ++rr) {
    if (r0 + rr >= n) break;
    float s = b;
    for (int j = 0; j < IN_DIR; ++j) s = fmaf(e[rr][j], w[j], s);
    out[(size_t)(r0 + rr) * WH + tid] = s;
  }
}

}  // namespace

int launch_dirbias(const mnrf_field* f, const float* src, int n, int src_stride, int from_embedded, float* out,
                   cudaStream_t st, const int* n_dev) {
  if (n <= 0) return 0;
  k_dirbias<<<(n + 7) / 8, 128, 0, st>>>(f->f32, f->L, src, n, src_stride, from_embedded, out, n_dev);
  MNRF_LAUNCH_OK();
  return 0;
}

int launch_field_fp32(const mnrf_field* f, const FieldIO& io, cudaStream_t st) {
  if (io.n_points <= 0) return 0;
  size_t smem = SM_TOTAL * sizeof(float);
  if (first_use_on_device(TAG_FIELD_FP32))
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = (io.n_points + TP - 1) / TP;
  k_field_fp32<<<blocks, NT, smem, st>>>(f->f32, f->L, io, f->has_normal, f->has_mirror);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace mnrf
