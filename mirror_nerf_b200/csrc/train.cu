// Training path: one render pass (field at the samples of a ray batch -> compositor) with every
// activation kept in HBM, and its hand-written backward -- what torch.autograd records for R/models/rendering.py:87-266 +
// R/models/mirror_nerf.py:101-212 when train.py:129-145 calls render_rays with gradients enabled.
//
// Layer-by-layer design (not the fused tcgen05 kernel of field_tc.cu): activations are [points, features] fp32 row-major
// matrices in a caller-provided workspace; every Linear is one launch of a CUDA-core fp32 GEMM (k_gemm_nn) whose epilogue
// fuses bias / ReLU / ReLU-mask / per-ray term / rank-1 term; every weight gradient is a split-K "A^T B" GEMM (k_gemm_tn)
// that accumulates with atomics straight into the reference's [out,in] gradient layout.
//
// The analytic normal n = normalize(-d sigma/d xyz) (mirror_nerf.py:136-146) is the explicit reverse chain
//   q8 = w_sigma * relu'(z8),  q_{k-1} = (q_k W_k) * relu'(z_{k-1}),  g_pe = q1 W1 + q5 W5[:, :63],  g_x = J_pe(x)^T g_pe.
// Its backward (the reference gets it from create_graph=True, utils/func.py:10-25) is linear in the weights because
// relu'' = 0: with t0 = J_pe dL/dg_x,  t_k = (t_{k-1} W_k^T) * relu'(z_k)  [a second "tangent" forward pass],
//   dW_k += q_k^T t_{k-1},   d w_sigma += sum_p t8.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mnrf {
namespace {

constexpr int GB_M = 128;  // rows of C per CTA
constexpr int GB_K = 16;   // reduction slab

__device__ __forceinline__ float apply_act(float v, int act, float m) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return m > 0.f ? v : 0.f;
  if (act == 3) return v > 0.f ? v : 0.01f * v;
  return v;
}

// C[M, gridDim.y*BN] = epi(A[M,K] * B[K,N]);  A, B, C row-major; K % 16 == 0; lda, ldb, ldc multiples of 4.
// 256 threads, 128 x BN tile, 8 x (BN/16) outputs per thread, register-prefetched double-buffered shared memory.
template <int BN>
__global__ void __launch_bounds__(256, 2) k_gemm_nn(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                 float* __restrict__ C, int ldc, int M, int K, GemmEpi e) {
  constexpr int TN = BN / 16;
  constexpr int NB4 = (GB_K * BN / 4) / 256;  // float4 loads of the B tile per thread (2 | 1)
  __shared__ __align__(16) float As[2][GB_K][GB_M];
  __shared__ __align__(16) float Bs[2][GB_K][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GB_M, n0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;

  // A loader: thread -> row (tid & 127), K quads (tid >> 7) and (tid >> 7) + 2
  const int a_row = tid & 127;
  const int a_q = tid >> 7;
  const float* a_ptr = A + (size_t)min(m0 + a_row, M - 1) * lda + a_q * 4;
  // B loader: rows k = b_k + i * (256 / (BN/4)), float4 column b_q
  constexpr int BQ = BN / 4;
  const int b_k = tid / BQ, b_q = tid % BQ;
  const float* b_ptr = B + (size_t)b_k * ldb + n0 + b_q * 4;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[NB4];
  auto gload = [&](int k0) {
    ra[0] = *reinterpret_cast<const float4*>(a_ptr + k0);
    ra[1] = *reinterpret_cast<const float4*>(a_ptr + k0 + 8);
#pragma unroll
    for (int i = 0; i < NB4; ++i)
      rb[i] = *reinterpret_cast<const float4*>(b_ptr + (size_t)(k0 + i * (256 / BQ)) * ldb);
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kq = (a_q + 2 * i) * 4;
      As[buf][kq + 0][a_row] = ra[i].x; As[buf][kq + 1][a_row] = ra[i].y;
      As[buf][kq + 2][a_row] = ra[i].z; As[buf][kq + 3][a_row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < NB4; ++i) *reinterpret_cast<float4*>(&Bs[buf][b_k + i * (256 / BQ)][b_q * 4]) = rb[i];
  };

  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / GB_K;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * GB_K);
#pragma unroll
    for (int kk = 0; kk < GB_K; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        if (TN == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][(BN / 2) + tx * 4]);
          b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
    const float rv = e.rvec != nullptr ? e.rvec[(size_t)row * e.ld_rvec] : 0.f;
#pragma unroll
    for (int g = 0; g < TN / 4; ++g) {
      const int col = n0 + g * (BN / 2) + tx * 4;
      float v[4] = {acc[i][4 * g + 0], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]};
      float* cp = C + (size_t)row * ldc + col;
      if (e.accumulate) {
        const float4 c = *reinterpret_cast<const float4*>(cp);
        v[0] += c.x; v[1] += c.y; v[2] += c.z; v[3] += c.w;
      }
      if (e.bias != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(e.bias + col);
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
      }
      if (e.rowbias != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(e.rowbias + (size_t)(row / e.rb_div) * e.ld_rb + col);
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
      }
      if (e.rvec != nullptr) {
        const float4 c = *reinterpret_cast<const float4*>(e.cvec + col);
        v[0] = fmaf(rv, c.x, v[0]); v[1] = fmaf(rv, c.y, v[1]); v[2] = fmaf(rv, c.z, v[2]); v[3] = fmaf(rv, c.w, v[3]);
      }
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e.act == 2) m = *reinterpret_cast<const float4*>(e.mask + (size_t)row * e.ld_mask + col);
      float4 o;
      o.x = apply_act(v[0], e.act, m.x); o.y = apply_act(v[1], e.act, m.y);
      o.z = apply_act(v[2], e.act, m.z); o.w = apply_act(v[3], e.act, m.w);
      *reinterpret_cast<float4*>(cp) = o;
    }
  }
}

// Wg[na0 + i][col0 + j] += sum_{p in this CTA's row range} A[p][na0 + i] * B[p][nb0 + j]   (j + nb0 < valid_cols)
// A [P, lda], B [P, ldb] row-major; output tile 128 x BN; blockIdx.z splits the P range; atomicAdd accumulation.
template <int BN>
__global__ void __launch_bounds__(256, 2) k_gemm_tn(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                 float* __restrict__ Wg, int ldw, int col0, int valid_cols, int P,
                                                 int rows_per_cta) {
  constexpr int TN = BN / 16;
  constexpr int BQ = BN / 4;
  constexpr int NB4 = (GB_K * BN / 4) / 256;
  __shared__ __align__(16) float As[2][GB_K][GB_M];
  __shared__ __align__(16) float Bs[2][GB_K][BN];
  const int tid = threadIdx.x;
  const int na0 = blockIdx.x * GB_M, nb0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;
  const int p_begin = blockIdx.z * rows_per_cta;
  const int p_end = min(P, p_begin + rows_per_cta);
  if (p_begin >= p_end) return;
  const int a_k = tid >> 5, a_q = tid & 31;  // rows a_k, a_k + 8
  const int b_k = tid / BQ, b_q = tid % BQ;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[NB4];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gload = [&](int p0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int p = p0 + a_k + 8 * i;
      ra[i] = p < p_end ? *reinterpret_cast<const float4*>(A + (size_t)p * lda + na0 + a_q * 4) : zero4;
    }
#pragma unroll
    for (int i = 0; i < NB4; ++i) {
      const int p = p0 + b_k + i * (256 / BQ);
      rb[i] = p < p_end ? *reinterpret_cast<const float4*>(B + (size_t)p * ldb + nb0 + b_q * 4) : zero4;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&As[buf][a_k + 8 * i][a_q * 4]) = ra[i];
#pragma unroll
    for (int i = 0; i < NB4; ++i) *reinterpret_cast<float4*>(&Bs[buf][b_k + i * (256 / BQ)][b_q * 4]) = rb[i];
  };

  gload(p_begin);
  sstore(0);
  __syncthreads();
  const int nk = (p_end - p_begin + GB_K - 1) / GB_K;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(p_begin + (kt + 1) * GB_K);
#pragma unroll
    for (int kk = 0; kk < GB_K; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        if (TN == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][(BN / 2) + tx * 4]);
          b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = na0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = nb0 + (j < 4 ? tx * 4 + j : (BN / 2) + tx * 4 + (j - 4));
      if (c < valid_cols) atomicAdd(Wg + (size_t)r * ldw + col0 + c, acc[i][j]);
    }
  }
}

// out[c] += sum_p X[p][c]   (bias gradients, d w_sigma of the tangent pass)
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ X, int ld, int P, int N, int rows_per_block,
                                                float* __restrict__ out) {
  const int c = threadIdx.x % N;
  const int sub = threadIdx.x / N, nsub = blockDim.x / N;
  const int p0 = blockIdx.x * rows_per_block, p1 = min(P, p0 + rows_per_block);
  float s = 0.f;
  for (int p = p0 + sub; p < p1; p += nsub) s += X[(size_t)p * ld + c];
  if (p0 < p1) atomicAdd(out + c, s);
}

// ---- pointwise kernels ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// x = o + d*z (separately rounded, rendering.py:302), PE row [x, sin(2^f x), cos(2^f x) ...] padded to 64 (mirror_nerf.py:33-38)
__global__ void k_train_pe(const float* __restrict__ rays, const float* __restrict__ z, int P, int S, float* __restrict__ pe) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float* r = rays + (size_t)(p / S) * 8;
  const float zz = z[p];
  float v[64];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = __fadd_rn(r[c], __fmul_rn(r[3 + c], zz));
#pragma unroll
  for (int f = 0; f < NFREQ_XYZ; ++f)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      sincosf(ldexpf(v[c], f), &s, &co);
      v[3 + 6 * f + c] = s;
      v[6 + 6 * f + c] = co;
    }
  v[63] = 0.f;
  float4* o = reinterpret_cast<float4*>(pe + (size_t)p * 64);
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// embedded ray direction (rendering.py:275: embedding_dir(rays_d)), zero-padded to 64 columns
__global__ void k_train_dir_pe(const float* __restrict__ rays, int n, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = rays[(size_t)r * 8 + 3 + c];
#pragma unroll
  for (int f = 0; f < NFREQ_DIR; ++f)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      sincosf(ldexpf(v[c], f), &s, &co);
      v[3 + 6 * f + c] = s;
      v[6 + 6 * f + c] = co;
    }
  float4* o = reinterpret_cast<float4*>(out + (size_t)r * 64);
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

struct HeadW {  // small head weights (fp32 section pointers)
  const float *w_sigma, *b_sigma, *w_rgb, *b_rgb, *w_n1, *b_n1, *w_m2, *b_m2;
};

// sigma / rgb / pred-normal / mirror outputs of one point from h8, d1, n1, m1 (mirror_nerf.py:196,199-212). Warp per point.
__global__ void __launch_bounds__(256) k_train_heads_fwd(const float* __restrict__ H8, const float* __restrict__ D1,
                                                         const float* __restrict__ N1, const float* __restrict__ M1, HeadW w,
                                                         int P, float* __restrict__ raw, float* __restrict__ n2out) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  float ws[8], wr[3][4], wn[3][4], wm[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) ws[i] = w.w_sigma[lane * 8 + i];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      wr[k][i] = w.w_rgb[k * WH + lane * 4 + i];
      wn[k][i] = N1 != nullptr ? w.w_n1[k * WH + lane * 4 + i] : 0.f;
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) wm[i] = M1 != nullptr ? w.w_m2[lane * 4 + i] : 0.f;
  for (int p = warp0; p < P; p += nwarps) {
    const float4 h0 = *reinterpret_cast<const float4*>(H8 + (size_t)p * W + lane * 8);
    const float4 h1 = *reinterpret_cast<const float4*>(H8 + (size_t)p * W + lane * 8 + 4);
    float sg = h0.x * ws[0] + h0.y * ws[1] + h0.z * ws[2] + h0.w * ws[3] + h1.x * ws[4] + h1.y * ws[5] + h1.z * ws[6] + h1.w * ws[7];
    const float4 d = *reinterpret_cast<const float4*>(D1 + (size_t)p * WH + lane * 4);
    float c[3], nn[3] = {0.f, 0.f, 0.f}, mm = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = d.x * wr[k][0] + d.y * wr[k][1] + d.z * wr[k][2] + d.w * wr[k][3];
    if (N1 != nullptr) {
      const float4 q = *reinterpret_cast<const float4*>(N1 + (size_t)p * WH + lane * 4);
#pragma unroll
      for (int k = 0; k < 3; ++k) nn[k] = q.x * wn[k][0] + q.y * wn[k][1] + q.z * wn[k][2] + q.w * wn[k][3];
    }
    if (M1 != nullptr) {
      const float4 q = *reinterpret_cast<const float4*>(M1 + (size_t)p * WH + lane * 4);
      mm = q.x * wm[0] + q.y * wm[1] + q.z * wm[2] + q.w * wm[3];
    }
    sg = warp_sum(sg); mm = warp_sum(mm);
#pragma unroll
    for (int k = 0; k < 3; ++k) { c[k] = warp_sum(c[k]); nn[k] = warp_sum(nn[k]); }
    if (lane == 0) {
      float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      o[0] = sg + w.b_sigma[0];
#pragma unroll
      for (int k = 0; k < 3; ++k) o[1 + k] = sigmoidf_(c[k] + w.b_rgb[k]);
      if (M1 != nullptr) o[4] = sigmoidf_(mm + w.b_m2[0]);
      float n2[4] = {0.f, 0.f, 0.f, 0.f};
      if (N1 != nullptr) {
#pragma unroll
        for (int k = 0; k < 3; ++k) n2[k] = nn[k] + w.b_n1[k];
        const float len = sqrtf(fmaxf(n2[0] * n2[0] + n2[1] * n2[1] + n2[2] * n2[2], FP32_EPS));  // utils/func.py:5-7
#pragma unroll
        for (int k = 0; k < 3; ++k) o[5 + k] = n2[k] / len;
      }
      float4* rp = reinterpret_cast<float4*>(raw + (size_t)p * 8);
      rp[0] = make_float4(o[0], o[1], o[2], o[3]);
      rp[1] = make_float4(o[4], o[5], o[6], o[7]);
      *reinterpret_cast<float4*>(n2out + (size_t)p * 4) = make_float4(n2[0], n2[1], n2[2], 0.f);
    }
  }
}

// q8 = w_sigma * relu'(h8)
__global__ void k_train_chain_start(const float* __restrict__ H8, const float* __restrict__ w_sigma, size_t n4,
                                    float* __restrict__ Q8) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 h = reinterpret_cast<const float4*>(H8)[i];
  const float4 w = *reinterpret_cast<const float4*>(w_sigma + (i % (W / 4)) * 4);
  reinterpret_cast<float4*>(Q8)[i] = make_float4(h.x > 0.f ? w.x : 0.f, h.y > 0.f ? w.y : 0.f, h.z > 0.f ? w.z : 0.f,
                                                 h.w > 0.f ? w.w : 0.f);
}

// g_x = J_pe(x)^T g_pe;  n = normalize(-g_x)   (mirror_nerf.py:143-145).  Thread per point.
__global__ void k_train_normal(const float* __restrict__ GPE, const float* __restrict__ PE, int P, float* __restrict__ gx_out,
                               float* __restrict__ normal) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float g[64], e[64];
  const float4* gp = reinterpret_cast<const float4*>(GPE + (size_t)p * 64);
  const float4* ep = reinterpret_cast<const float4*>(PE + (size_t)p * 64);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 a = gp[i], b = ep[i];
    g[4 * i] = a.x; g[4 * i + 1] = a.y; g[4 * i + 2] = a.z; g[4 * i + 3] = a.w;
    e[4 * i] = b.x; e[4 * i + 1] = b.y; e[4 * i + 2] = b.z; e[4 * i + 3] = b.w;
  }
  float gx[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = g[c];
#pragma unroll
    for (int f = 0; f < NFREQ_XYZ; ++f)
      s += ldexpf(1.f, f) * (g[3 + 6 * f + c] * e[6 + 6 * f + c] - g[6 + 6 * f + c] * e[3 + 6 * f + c]);
    gx[c] = s;
  }
  *reinterpret_cast<float4*>(gx_out + (size_t)p * 4) = make_float4(gx[0], gx[1], gx[2], 0.f);
  const float len = sqrtf(fmaxf(gx[0] * gx[0] + gx[1] * gx[1] + gx[2] * gx[2], FP32_EPS));
  normal[(size_t)p * 3 + 0] = -gx[0] / len;
  normal[(size_t)p * 3 + 1] = -gx[1] / len;
  normal[(size_t)p * 3 + 2] = -gx[2] / len;
}

// backward of y = v / sqrt(max(|v|^2, eps))
__device__ __forceinline__ void normalize_bwd(const float (&v)[3], const float (&dy)[3], float (&dv)[3]) {
  const float nn = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (nn > FP32_EPS) {
    const float inv = 1.f / sqrtf(nn);
    const float y0 = v[0] * inv, y1 = v[1] * inv, y2 = v[2] * inv;
    const float dot = y0 * dy[0] + y1 * dy[1] + y2 * dy[2];
    dv[0] = (dy[0] - y0 * dot) * inv; dv[1] = (dy[1] - y1 * dot) * inv; dv[2] = (dy[2] - y2 * dot) * inv;
  } else {
    const float inv = 1.f / sqrtf(FP32_EPS);
    dv[0] = dy[0] * inv; dv[1] = dy[1] * inv; dv[2] = dy[2] * inv;
  }
}

// ---- compositor backward (rendering.py:175-264).  Warp per ray.  Per-point record DR[p] (12 floats):
//      [0] d sigma  [1..3] d rgb  [4] d is_mirror  [5..7] d pred_normal (normalised)  [8..10] d analytic normal  [11] 0
constexpr int DR_STRIDE = 12;
constexpr int CB_WARPS = 4;
constexpr int CB_MAXS = 512;

__global__ void __launch_bounds__(CB_WARPS * 32)
k_train_composite_bwd(const float* __restrict__ rays, const float* __restrict__ z, const float* __restrict__ raw,
                      const float* __restrict__ normal, const float* __restrict__ noise, float noise_std, int n, int S,
                      int white_back, int detach_mask, int detach_normal, const float* __restrict__ ray_detach_mirror,
                      mnrf_train_grads g, float* __restrict__ DR) {
  __shared__ float s_gw[CB_WARPS][CB_MAXS], s_T[CB_WARPS][CB_MAXS], s_om[CB_WARPS][CB_MAXS], s_G[CB_WARPS][CB_MAXS];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int r = blockIdx.x * CB_WARPS + wi;
  if (r >= n) return;
  const size_t base = (size_t)r * S;
  // per-ray output gradients
  float g_rgb[3] = {0.f, 0.f, 0.f}, g_sn[3] = {0.f, 0.f, 0.f}, g_sng[3] = {0.f, 0.f, 0.f};
  float g_depth = g.depth ? g.depth[r] : 0.f;
  const float g_op = g.opacity ? g.opacity[r] : 0.f;
  const float g_m = g.mirror_mask ? g.mirror_mask[r] : 0.f;
  const float g_nd = (g.normal_dif && normal != nullptr) ? g.normal_dif[r] : 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (g.rgb) g_rgb[c] = g.rgb[(size_t)r * 3 + c];
    if (g.surface_normal) g_sn[c] = g.surface_normal[(size_t)r * 3 + c];
    if (g.surface_normal_grad && normal != nullptr) g_sng[c] = g.surface_normal_grad[(size_t)r * 3 + c];
    if (g.x_surface) g_depth += g.x_surface[(size_t)r * 3 + c] * rays[(size_t)r * 8 + 3 + c];  // x = o + d * depth
  }
  const bool mirror_w = !detach_mask && !(ray_detach_mirror != nullptr && ray_detach_mirror[r] != 0.f);
  const bool normal_w = !detach_normal;
  const float g_white = white_back ? (g_rgb[0] + g_rgb[1] + g_rgb[2]) : 0.f;

  // pass 1 (forward order): alpha, transmittance, weights, G_i = dL/dw_i
  float carry = 1.f;
  const int nblk = (S + 31) / 32;
  for (int b = 0; b < nblk; ++b) {
    const int s = b * 32 + lane;
    const bool ok = s < S;
    float zz = 0.f, alpha = 0.f;
    if (ok) {
      zz = z[base + s];
      const float delta = (s + 1 < S) ? __fsub_rn(z[base + s + 1], zz) : 1e10f;
      float sg = raw[(base + s) * 8];
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
      const float4 q0 = *reinterpret_cast<const float4*>(raw + (base + s) * 8);
      const float4 q1 = *reinterpret_cast<const float4*>(raw + (base + s) * 8 + 4);
      float G = (g.weights ? g.weights[base + s] : 0.f) + g_op - g_white + g_depth * zz +
                g_rgb[0] * q0.y + g_rgb[1] * q0.z + g_rgb[2] * q0.w;
      if (mirror_w) G += g_m * q1.x;
      if (normal_w) {
        G += g_sn[0] * q1.y + g_sn[1] * q1.z + g_sn[2] * q1.w;
        if (normal != nullptr) {
          const float* nn = normal + (base + s) * 3;
          const float e0 = nn[0] - q1.y, e1 = nn[1] - q1.z, e2 = nn[2] - q1.w;
          G += g_sng[0] * nn[0] + g_sng[1] * nn[1] + g_sng[2] * nn[2] + g_nd * (e0 * e0 + e1 * e1 + e2 * e2);
        }
      }
      s_G[wi][s] = G;
      s_gw[wi][s] = G * w;
      s_T[wi][s] = T;
      s_om[wi][s] = f;  // 1 - alpha + 1e-10
      // per-sample channel gradients
      float* dr = DR + (base + s) * DR_STRIDE;
      float dpn[3] = {w * g_sn[0], w * g_sn[1], w * g_sn[2]}, dan[3] = {w * g_sng[0], w * g_sng[1], w * g_sng[2]};
      if (g.pred_normal) { dpn[0] += g.pred_normal[(base + s) * 3]; dpn[1] += g.pred_normal[(base + s) * 3 + 1]; dpn[2] += g.pred_normal[(base + s) * 3 + 2]; }
      if (normal != nullptr) {
        const float* nn = normal + (base + s) * 3;
        if (g.normal) { dan[0] += g.normal[(base + s) * 3]; dan[1] += g.normal[(base + s) * 3 + 1]; dan[2] += g.normal[(base + s) * 3 + 2]; }
        const float k2 = 2.f * w * g_nd;
        const float e0 = nn[0] - q1.y, e1 = nn[1] - q1.z, e2 = nn[2] - q1.w;
        dan[0] += k2 * e0; dan[1] += k2 * e1; dan[2] += k2 * e2;
        dpn[0] -= k2 * e0; dpn[1] -= k2 * e1; dpn[2] -= k2 * e2;
      }
      dr[1] = w * g_rgb[0]; dr[2] = w * g_rgb[1]; dr[3] = w * g_rgb[2];
      dr[4] = w * g_m;
      dr[5] = dpn[0]; dr[6] = dpn[1]; dr[7] = dpn[2];
      dr[8] = dan[0]; dr[9] = dan[1]; dr[10] = dan[2];
      dr[11] = 0.f;
    }
  }
  __syncwarp();
  // pass 2 (reverse order): S_i = sum_{j>i} G_j w_j;  d alpha_i = G_i T_i - S_i / (1 - alpha_i + 1e-10)
  float tail = 0.f;
  for (int b = nblk - 1; b >= 0; --b) {
    const int s = b * 32 + lane;
    const bool ok = s < S;
    const float v = ok ? s_gw[wi][s] : 0.f;
    float incl = v;  // inclusive suffix sum inside the block
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    const float suffix = tail + incl - v;
    tail += __shfl_sync(0xffffffffu, incl, 0);
    if (ok) {
      const float zz = z[base + s];
      const float delta = (s + 1 < S) ? __fsub_rn(z[base + s + 1], zz) : 1e10f;
      float sg = raw[(base + s) * 8];
      if (noise != nullptr) sg = __fadd_rn(sg, __fmul_rn(noise[base + s], noise_std));
      const float d_alpha = s_G[wi][s] * s_T[wi][s] - suffix / s_om[wi][s];
      // alpha = 1 - exp(-delta * relu(sg)):  d alpha / d sg = delta * exp(-delta * sg) for sg > 0
      const float ds = sg > 0.f ? d_alpha * delta * expf(-__fmul_rn(delta, sg)) : 0.f;
      DR[(base + s) * DR_STRIDE] = ds;
    }
  }
}

// Heads backward: from the per-point record to the gradients of the 128-wide hidden rows, plus the small weights' gradients.
// Warp per point (grid-stride); lane-private accumulators for the small weight gradients, flushed with atomics at the end.
struct HeadG {  // gradient tensors of the small heads (reference layout) or NULL
  float *w_sigma, *b_sigma, *w_rgb, *b_rgb, *w_n1, *b_n1, *w_m2, *b_m2;
};
__global__ void __launch_bounds__(256)
k_train_heads_bwd(const float* __restrict__ H8, const float* __restrict__ D1, const float* __restrict__ N1,
                  const float* __restrict__ M1, const float* __restrict__ raw, const float* __restrict__ n2s,
                  const float* __restrict__ DR, HeadW w, int P, float* __restrict__ dD1, float* __restrict__ dN1,
                  float* __restrict__ dM1, HeadG hg) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  float wr[3][4], wn[3][4], wm[4];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      wr[k][i] = w.w_rgb[k * WH + lane * 4 + i];
      wn[k][i] = N1 != nullptr ? w.w_n1[k * WH + lane * 4 + i] : 0.f;
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) wm[i] = M1 != nullptr ? w.w_m2[lane * 4 + i] : 0.f;
  float a_ws[8], a_wr[3][4], a_wn[3][4], a_wm[4], a_bs = 0.f, a_br[3] = {0.f, 0.f, 0.f}, a_bn[3] = {0.f, 0.f, 0.f}, a_bm = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) a_ws[i] = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) { a_wr[k][i] = 0.f; a_wn[k][i] = 0.f; }
#pragma unroll
  for (int i = 0; i < 4; ++i) a_wm[i] = 0.f;

  for (int p = warp0; p < P; p += nwarps) {
    const float* dr = DR + (size_t)p * DR_STRIDE;
    const float4 r0 = *reinterpret_cast<const float4*>(dr), r1 = *reinterpret_cast<const float4*>(dr + 4),
                 r2 = *reinterpret_cast<const float4*>(dr + 8);
    const float4 q0 = *reinterpret_cast<const float4*>(raw + (size_t)p * 8);
    const float4 q1 = *reinterpret_cast<const float4*>(raw + (size_t)p * 8 + 4);
    (void)r2;
    // sigma head
    const float ds = r0.x;
    {
      const float4 h0 = *reinterpret_cast<const float4*>(H8 + (size_t)p * W + lane * 8);
      const float4 h1 = *reinterpret_cast<const float4*>(H8 + (size_t)p * W + lane * 8 + 4);
      a_ws[0] = fmaf(ds, h0.x, a_ws[0]); a_ws[1] = fmaf(ds, h0.y, a_ws[1]); a_ws[2] = fmaf(ds, h0.z, a_ws[2]); a_ws[3] = fmaf(ds, h0.w, a_ws[3]);
      a_ws[4] = fmaf(ds, h1.x, a_ws[4]); a_ws[5] = fmaf(ds, h1.y, a_ws[5]); a_ws[6] = fmaf(ds, h1.z, a_ws[6]); a_ws[7] = fmaf(ds, h1.w, a_ws[7]);
      a_bs += ds;
    }
    // rgb head: rgb = sigmoid(W_r d1 + b)
    {
      const float dp[3] = {r0.y * q0.y * (1.f - q0.y), r0.z * q0.z * (1.f - q0.z), r0.w * q0.w * (1.f - q0.w)};
      const float4 d = *reinterpret_cast<const float4*>(D1 + (size_t)p * WH + lane * 4);
      const float dv[4] = {d.x, d.y, d.z, d.w};
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i] = dv[i] > 0.f ? (wr[0][i] * dp[0] + wr[1][i] * dp[1] + wr[2][i] * dp[2]) : 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) a_wr[k][i] = fmaf(dp[k], dv[i], a_wr[k][i]);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) a_br[k] += dp[k];
      *reinterpret_cast<float4*>(dD1 + (size_t)p * WH + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
    // predicted-normal head: n_hat = normalize(W_n1 n1 + b)
    if (N1 != nullptr) {
      const float4 s4 = *reinterpret_cast<const float4*>(n2s + (size_t)p * 4);
      const float v[3] = {s4.x, s4.y, s4.z}, dy[3] = {r1.y, r1.z, r1.w};
      float dn2[3];
      normalize_bwd(v, dy, dn2);
      const float4 q = *reinterpret_cast<const float4*>(N1 + (size_t)p * WH + lane * 4);
      const float nv[4] = {q.x, q.y, q.z, q.w};
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i] = wn[0][i] * dn2[0] + wn[1][i] * dn2[1] + wn[2][i] * dn2[2];
#pragma unroll
        for (int k = 0; k < 3; ++k) a_wn[k][i] = fmaf(dn2[k], nv[i], a_wn[k][i]);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) a_bn[k] += dn2[k];
      *reinterpret_cast<float4*>(dN1 + (size_t)p * WH + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
    // mirror head: m = sigmoid(W_m2 leaky(m1pre) + b);  M1 holds leaky(m1pre) (same sign as m1pre)
    if (M1 != nullptr) {
      const float dmp = r1.x * q1.x * (1.f - q1.x);
      const float4 q = *reinterpret_cast<const float4*>(M1 + (size_t)p * WH + lane * 4);
      const float mv[4] = {q.x, q.y, q.z, q.w};
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i] = wm[i] * dmp * (mv[i] > 0.f ? 1.f : 0.01f);
        a_wm[i] = fmaf(dmp, mv[i], a_wm[i]);
      }
      a_bm += dmp;
      *reinterpret_cast<float4*>(dM1 + (size_t)p * WH + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(hg.w_sigma + lane * 8 + i, a_ws[i]);
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(hg.w_rgb + k * WH + lane * 4 + i, a_wr[k][i]);
      if (N1 != nullptr) atomicAdd(hg.w_n1 + k * WH + lane * 4 + i, a_wn[k][i]);
    }
  if (M1 != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(hg.w_m2 + lane * 4 + i, a_wm[i]);
  }
  if (lane == 0) {  // the per-point scalars are warp-uniform
    atomicAdd(hg.b_sigma, a_bs);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atomicAdd(hg.b_rgb + k, a_br[k]);
      if (N1 != nullptr) atomicAdd(hg.b_n1 + k, a_bn[k]);
    }
    if (M1 != nullptr) atomicAdd(hg.b_m2, a_bm);
  }
}

// t0 = J_pe(x) * dL/dg_x with dL/dg_x = -normalize_bwd(dL/dn; v = -g_x)   (start of the tangent pass). Thread per point.
__global__ void k_train_tangent_start(const float* __restrict__ DR, const float* __restrict__ GX, const float* __restrict__ PE,
                                      int P, float* __restrict__ T0) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float4 gx4 = *reinterpret_cast<const float4*>(GX + (size_t)p * 4);
  const float v[3] = {-gx4.x, -gx4.y, -gx4.z};
  const float dy[3] = {DR[(size_t)p * DR_STRIDE + 8], DR[(size_t)p * DR_STRIDE + 9], DR[(size_t)p * DR_STRIDE + 10]};
  float dv[3];
  normalize_bwd(v, dy, dv);
  const float dg[3] = {-dv[0], -dv[1], -dv[2]};
  float e[64], t[64];
  const float4* ep = reinterpret_cast<const float4*>(PE + (size_t)p * 64);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 b = ep[i];
    e[4 * i] = b.x; e[4 * i + 1] = b.y; e[4 * i + 2] = b.z; e[4 * i + 3] = b.w;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) t[c] = dg[c];
#pragma unroll
  for (int f = 0; f < NFREQ_XYZ; ++f)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float sc = ldexpf(1.f, f);
      t[3 + 6 * f + c] = sc * e[6 + 6 * f + c] * dg[c];   // d g_x / d g_pe[sin] = 2^f cos
      t[6 + 6 * f + c] = -sc * e[3 + 6 * f + c] * dg[c];  // d g_x / d g_pe[cos] = -2^f sin
    }
  t[63] = 0.f;
  float4* o = reinterpret_cast<float4*>(T0 + (size_t)p * 64);
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = make_float4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
}

// Gradient w.r.t. the rays (origin, direction): the reference gets it from autograd when train.py:194-243 builds secondary
// rays from x_surface / surface normals without detaching.  Per ray r (one block):
//   d xyz_p = J_pe(x_p)^T dPE_p  [+ second-order term through the analytic normal: d g_x / d x = d J_pe^T / dx * g_pe]
//   d o = sum_p d xyz_p + g_xs,   d d = sum_p z_p d xyz_p + depth * g_xs + J_dirpe(d)^T W_dir[:,256:]^T sum_s dD1pre
__global__ void __launch_bounds__(64)
k_train_ray_grad(const float* __restrict__ rays, const float* __restrict__ z, const float* __restrict__ PE,
                 const float* __restrict__ dPE, const float* __restrict__ GPE, const float* __restrict__ GX,
                 const float* __restrict__ DR, const float* __restrict__ rsum, const float* __restrict__ dirpe,
                 const float* __restrict__ wt_dir, const float* __restrict__ g_xs, const float* __restrict__ depth, int S,
                 int second_order, float* __restrict__ grad_rays) {
  __shared__ float red[6][64];
  __shared__ float ddir[IN_DIR];
  const int r = blockIdx.x, t = threadIdx.x;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = t; s < S; s += 64) {
    const size_t p = (size_t)r * S + s;
    const float* e = PE + p * 64;
    const float* dp = dPE + p * 64;
    float dx[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = dp[c];
      for (int f = 0; f < NFREQ_XYZ; ++f)
        v += ldexpf(1.f, f) * (dp[3 + 6 * f + c] * e[6 + 6 * f + c] - dp[6 + 6 * f + c] * e[3 + 6 * f + c]);
      dx[c] = v;
    }
    if (second_order) {
      const float4 gx4 = *reinterpret_cast<const float4*>(GX + p * 4);
      const float v[3] = {-gx4.x, -gx4.y, -gx4.z};
      const float dy[3] = {DR[p * DR_STRIDE + 8], DR[p * DR_STRIDE + 9], DR[p * DR_STRIDE + 10]};
      float dv[3];
      normalize_bwd(v, dy, dv);
      const float* gp = GPE + p * 64;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float h = 0.f;  // d g_x[c] / d x[c]
        for (int f = 0; f < NFREQ_XYZ; ++f)
          h -= ldexpf(1.f, 2 * f) * (gp[3 + 6 * f + c] * e[3 + 6 * f + c] + gp[6 + 6 * f + c] * e[6 + 6 * f + c]);
        dx[c] += -dv[c] * h;  // dL/dg_x = -dv
      }
    }
    const float zz = z[p];
#pragma unroll
    for (int c = 0; c < 3; ++c) { acc[c] += dx[c]; acc[3 + c] = fmaf(zz, dx[c], acc[3 + c]); }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) red[i][t] = acc[i];
  // d L / d embed(dir)[j] = sum_i W_dir[i][256 + j] * rsum[r][i]
  if (t < IN_DIR) {
    float v = 0.f;
    for (int i = 0; i < WH; ++i) v = fmaf(wt_dir[(size_t)(W + t) * WH + i], rsum[(size_t)r * WH + i], v);
    ddir[t] = v;
  }
  __syncthreads();
  if (t < 6) {
    float v = 0.f;
    for (int i = 0; i < 64; ++i) v += red[t][i];
    const int c = t % 3;
    if (g_xs != nullptr) v += (t < 3 ? 1.f : depth[r]) * g_xs[(size_t)r * 3 + c];
    if (t >= 3) {
      const float* de = dirpe + (size_t)r * 64;
      float w = ddir[c];
      for (int f = 0; f < NFREQ_DIR; ++f)
        w += ldexpf(1.f, f) * (ddir[3 + 6 * f + c] * de[6 + 6 * f + c] - ddir[6 + 6 * f + c] * de[3 + 6 * f + c]);
      v += w;
    }
    grad_rays[(size_t)r * 8 + t] = v;
  }
  if (t >= 6 && t < 8) grad_rays[(size_t)r * 8 + t] = 0.f;
}

// R[r][c] = sum_s X[r*S + s][c]   (c < 128): per-ray sum of the dir layer's pre-activation gradient
__global__ void __launch_bounds__(WH) k_train_sum_samples(const float* __restrict__ X, int S, float* __restrict__ R) {
  const int r = blockIdx.x, c = threadIdx.x;
  float s = 0.f;
  for (int i = 0; i < S; ++i) s += X[((size_t)r * S + i) * WH + c];
  R[(size_t)r * WH + c] = s;
}

// zero the rows of X (128 wide) that belong to flagged rays
__global__ void k_train_zero_rows(float* __restrict__ X, const float* __restrict__ flag, int S, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const size_t row = i / (WH / 4);
  if (flag[row / S] != 0.f) reinterpret_cast<float4*>(X)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void k_axpy(float* __restrict__ out, const float* __restrict__ in, long long n, float alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmaf(alpha, in[i], out[i]);
}

// torch.optim.Adam (the reference's optimizer, R/utils/__init__.py:47-58: lr, eps=1e-8, weight_decay as L2) on flat buffers;
// `grad_scale` folds the 1/world_size of the data-parallel gradient average into the same pass.
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       long long n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt,
                       float grad_scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i] * grad_scale;
  const float pi = p[i];
  if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}

// relu' bits of a [P,256] activation matrix: word w of row p holds columns 32w..32w+31 (CUDA-core engine; the tensor-core
// forward epilogue writes the same words itself)
__global__ void k_train_make_bits(const float* __restrict__ H, size_t n_words, uint32_t* __restrict__ bits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one warp per word: lane = column inside the word
  const size_t word = i >> 5;
  if (word >= n_words) return;
  const uint32_t b = __ballot_sync(0xffffffffu, H[word * 32 + (i & 31)] > 0.f);
  if ((i & 31) == 0) bits[word] = b;
}

// Data-parallel optimizer step over NVLink peer memory: reduce-scatter + Adam + all-gather in ONE kernel.  Every rank owns a
// contiguous shard of the flat parameter vector: it sums that shard of the gradient over all ranks' (peer-mapped) gradient
// buffers with 16-byte P2P loads, applies torch.optim.Adam to it (moment estimates exist only for the owned shard), and stores
// the new parameters into every rank's parameter buffer with P2P stores.  Cross-rank ordering (gradients complete before /
// parameters visible after) is the caller's device-side barrier on the symmetric-memory signal pads.
struct PeerPtrs { float* p[8]; };
__global__ void k_peer_allreduce_adam(PeerPtrs grads, PeerPtrs params, int world, long long lo, long long hi, float* __restrict__ m,
                                      float* __restrict__ v, float lr, float beta1, float beta2, float eps, float weight_decay,
                                      float bc1, float bc2_sqrt, float grad_scale, int rank) {
  const long long i = lo + 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= hi) return;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < world; ++r) {
    const float4 x = *reinterpret_cast<const float4*>(grads.p[r] + i);
    g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
  }
  const float4 p4 = *reinterpret_cast<const float4*>(params.p[rank] + i);
  float gv[4] = {g.x * grad_scale, g.y * grad_scale, g.z * grad_scale, g.w * grad_scale};
  float pv[4] = {p4.x, p4.y, p4.z, p4.w};
  const long long j = i - lo;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (i + k < hi) {
      float gi = gv[k];
      if (weight_decay != 0.f) gi = fmaf(weight_decay, pv[k], gi);
      const float mi = beta1 * m[j + k] + (1.f - beta1) * gi;
      const float vi = beta2 * v[j + k] + (1.f - beta2) * gi * gi;
      m[j + k] = mi;
      v[j + k] = vi;
      pv[k] = pv[k] - (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
  }
  const float4 out = make_float4(pv[0], pv[1], pv[2], pv[3]);
  for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(params.p[r] + i) = out;
}

// ---- host-side helpers -----------------------------------------------------------------------------------------------
inline size_t al(size_t b) { return (b + 255) & ~(size_t)255; }

struct FwdWs {  // float offsets into the forward workspace
  size_t pe, h[8], hb[8], f, d1, n1, m1, raw, n2, dirbias, q[8], gpe, gx, nrm, total;
};
FwdWs fwd_layout(int n, int S, int compute_normal) {
  const size_t P = (size_t)n * S;
  FwdWs L;
  size_t o = 0;
  auto take = [&](size_t floats) { size_t r = o; o += al(floats * sizeof(float)) / sizeof(float); return r; };
  L.pe = take(P * 64);
  for (int l = 0; l < 8; ++l) L.h[l] = take(P * W);
  for (int l = 0; l < 8; ++l) L.hb[l] = take(P * 8);  // relu' bit masks (uint32 words)
  L.f = take(P * W);
  L.d1 = take(P * WH);
  L.n1 = take(P * WH);
  L.m1 = take(P * WH);
  L.raw = take(P * 8);
  L.n2 = take(P * 4);
  L.dirbias = take((size_t)n * WH);
  for (int l = 0; l < 8; ++l) L.q[l] = compute_normal ? take(P * W) : 0;
  L.gpe = compute_normal ? take(P * 64) : 0;
  L.gx = compute_normal ? take(P * 4) : 0;
  L.nrm = compute_normal ? take(P * 3) : 0;
  L.total = o;
  return L;
}
struct BwdWs {
  size_t dr, dd1, dn1, dm1, df, dz[2], t0, t[2], rsum, dirpe, dpe, total;
};
BwdWs bwd_layout(int n, int S, int compute_normal) {
  const size_t P = (size_t)n * S;
  BwdWs L;
  size_t o = 0;
  auto take = [&](size_t floats) { size_t r = o; o += al(floats * sizeof(float)) / sizeof(float); return r; };
  L.dr = take(P * DR_STRIDE);
  L.dd1 = take(P * WH);
  L.dn1 = take(P * WH);
  L.dm1 = take(P * WH);
  L.df = take(P * W);
  L.dz[0] = take(P * W);
  L.dz[1] = take(P * W);
  L.t0 = compute_normal ? take(P * 64) : 0;
  L.t[0] = compute_normal ? take(P * W) : 0;
  L.t[1] = compute_normal ? take(P * W) : 0;
  L.rsum = take((size_t)n * WH);
  L.dirpe = take((size_t)n * 64);
  L.dpe = take(P * 64);
  L.total = o;
  return L;
}

int gemm_nn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K, const GemmEpi& e,
            cudaStream_t st) {
  MNRF_REQUIRE(K % GB_K == 0 && (N % 64) == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "train gemm_nn: bad shape %d %d %d", M, N, K);
  if (N % 128 == 0) {
    k_gemm_nn<128><<<dim3((M + GB_M - 1) / GB_M, N / 128), 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, K, e);
  } else {
    k_gemm_nn<64><<<dim3((M + GB_M - 1) / GB_M, N / 64), 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, K, e);
  }
  MNRF_LAUNCH_OK();
  return 0;
}

// Wg[NA rows][col0 + (0..valid)] += A[P,NA]^T B[P,NB]
int gemm_tn(const float* A, int lda, int NA, const float* B, int ldb, int NB, float* Wg, int ldw, int col0, int valid, int P,
            cudaStream_t st) {
  if (Wg == nullptr || P <= 0) return 0;
  MNRF_REQUIRE(NA % GB_M == 0 && NB % 64 == 0 && lda % 4 == 0 && ldb % 4 == 0, "train gemm_tn: bad shape %d %d", NA, NB);
  const int bn = (NB % 128 == 0) ? 128 : 64;
  const int tiles = (NA / GB_M) * (NB / bn);
  int splits = (3 * 148 + tiles - 1) / tiles;
  int rows = (P + splits - 1) / splits;
  rows = ((rows + GB_K - 1) / GB_K) * GB_K;
  if (rows < 64) rows = 64;
  splits = (P + rows - 1) / rows;
  if (bn == 128) k_gemm_tn<128><<<dim3(NA / GB_M, NB / 128, splits), 256, 0, st>>>(A, lda, B, ldb, Wg, ldw, col0, valid, P, rows);
  else           k_gemm_tn<64><<<dim3(NA / GB_M, NB / 64, splits), 256, 0, st>>>(A, lda, B, ldb, Wg, ldw, col0, valid, P, rows);
  MNRF_LAUNCH_OK();
  return 0;
}

int colsum(const float* X, int ld, int P, int N, float* out, cudaStream_t st) {
  if (out == nullptr || P <= 0) return 0;
  MNRF_REQUIRE(N <= 256 && 256 % N == 0, "train colsum: bad width %d", N);
  int rows = (P + 4 * 148 - 1) / (4 * 148);
  if (rows < 32) rows = 32;
  k_colsum<<<(P + rows - 1) / rows, 256, 0, st>>>(X, ld, P, N, rows, out);
  MNRF_LAUNCH_OK();
  return 0;
}

// ---- GEMM engine: tcgen05 (train_tc.cu) by default; MNRF_TRAIN_GEMM=simt or mnrf_train_set_gemm(0) selects the fp32 CUDA-core
// kernels above (verification twin).
int g_engine = -1;
bool use_tc() {
  if (g_engine < 0) {
    const char* e = getenv("MNRF_TRAIN_GEMM");
    g_engine = (e != nullptr && strcmp(e, "simt") == 0) ? 0 : ((e != nullptr && strcmp(e, "tf32") == 0) ? 2 : 1);
    set_train_tc_one_pass(g_engine == 2);
  }
  return g_engine != 0;
}

inline int chain_step(int l) { return l >= 5 ? 18 - l : (l == 4 ? 14 : 19 - l); }  // W_l^T (h part for l == 4); PE part of l == 4: 15

// C[P, N(step)] = epi([A0 | A1] * B_step^T): B_step is the weight matrix of GEMM step `step` (common.cuh T32 numbering)
int gemm_w(const mnrf_field* f, int step, const float* A0, int lda0, int K0, const float* A1, int lda1, float* C, int ldc,
           int P, const GemmEpi& e, cudaStream_t st) {
  if (use_tc()) return gemm_nn_tc(f, step, A0, lda0, K0, A1, lda1, C, ldc, P, e, st);
  const float* F = f->f32;
  const F32Layout& L = f->L;
  const int N = t32_step_n(step), K = t32_step_k(step);
  const float* B0 = nullptr;
  const float* B1 = nullptr;  // second K segment (only the skip layer has one)
  int ldb = W;
  if (step < 8) {
    B0 = F + L.wt_trunk[step];
    if (step == 4) B1 = F + L.wt_trunk[4] + (size_t)IN_XYZ * W;
  } else if (step == 8) B0 = F + L.wt_final;
  else if (step == 9) { B0 = F + L.wt_m0; ldb = WH; }
  else if (step == 10) { B0 = F + L.wt_dir; ldb = WH; }
  else if (step == 20) { B0 = F + L.wt_n0; ldb = WH; }
  else if (step == 14) B0 = F + L.tw_l5b;
  else if (step == 15) { B0 = F + L.tw_l5a; ldb = 64; }
  else if (step == 19) { B0 = F + L.tw_l1; ldb = 64; }
  else if (step < TC_NUM_STEPS) B0 = F + L.w_trunk[tc_step_layer(step)];
  else if (step == 21) B0 = F + L.tw_dira;
  else if (step == 22) B0 = F + L.tw_final;
  else if (step == 23) B0 = F + L.tw_n0;
  else B0 = F + L.tw_m0;
  if (K0 == K) return gemm_nn(A0, lda0, B0, ldb, C, ldc, P, N, K, e, st);
  MNRF_REQUIRE(B1 != nullptr && A1 != nullptr, "train gemm_w: step %d has no second K segment", step);
  GemmEpi e0;
  e0.accumulate = e.accumulate;
  if (gemm_nn(A0, lda0, B0, ldb, C, ldc, P, N, K0, e0, st)) return 1;
  GemmEpi e1 = e;
  e1.accumulate = 1;
  return gemm_nn(A1, lda1, B1, ldb, C, ldc, P, N, K - K0, e1, st);
}

// weight gradient Wg += A^T B; bias_g (optional) += column sums of A (fused into the tensor-core kernel's pass over A)
int gemm_g(const float* A, int lda, int NA, const float* B, int ldb, int NB, float* Wg, int ldw, int col0, int valid, int P,
           cudaStream_t st, float* bias_g = nullptr) {
  if (use_tc()) return gemm_tn_tc(A, lda, NA, B, ldb, NB, Wg, ldw, col0, valid, P, bias_g, st);
  if (bias_g != nullptr && colsum(A, lda, P, NA, bias_g, st)) return 1;
  return gemm_tn(A, lda, NA, B, ldb, NB, Wg, ldw, col0, valid, P, st);
}

}  // namespace

// compositor backward alone (also the first step of the hash-grid field's backward, train_hash.cu): DR (n*S, 12)
int launch_composite_bwd(const float* rays, const float* z, const float* raw, const float* normal, const float* noise,
                         const mnrf_train_cfg& cfg, int n, const float* ray_detach_mirror, const mnrf_train_grads& g, float* DR,
                         cudaStream_t st) {
  static_assert(DR_STRIDE == 12, "train_hash.cu reads 12-float gradient records");
  MNRF_REQUIRE(cfg.S <= CB_MAXS, "train_pass_bwd: S <= %d", CB_MAXS);
  k_train_composite_bwd<<<(n + CB_WARPS - 1) / CB_WARPS, CB_WARPS * 32, 0, st>>>(
      rays, z, raw, normal, noise, cfg.noise_std, n, cfg.S, cfg.white_back, cfg.detach_density_for_mask_loss,
      cfg.detach_density_for_normal_loss, ray_detach_mirror, g, DR);
  MNRF_LAUNCH_OK();
  return 0;
}

int64_t train_fwd_workspace_bytes(int n, int S, int compute_normal) {
  return (int64_t)(fwd_layout(n, S, compute_normal).total * sizeof(float));
}
int64_t train_bwd_workspace_bytes(int n, int S, int compute_normal) {
  return (int64_t)(bwd_layout(n, S, compute_normal).total * sizeof(float));
}

int train_pass_fwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                   const mnrf_train_cfg& cfg, void* ws, const mnrf_composite_out& out, float* normal_out, cudaStream_t st) {
  const int S = cfg.S;
  const int P = n * S;
  const FwdWs L = fwd_layout(n, S, cfg.compute_normal);
  float* w = reinterpret_cast<float*>(ws);
  const float* F = f->f32;
  const F32Layout& FL = f->L;
  float* PE = w + L.pe;
  float* H[8];
  for (int l = 0; l < 8; ++l) H[l] = w + L.h[l];

  k_train_pe<<<(P + 127) / 128, 128, 0, st>>>(rays, z, P, S, PE);
  MNRF_LAUNCH_OK();
  // trunk (mirror_nerf.py:189-197); the skip layer reads [pe | h4] as two K segments (PE column 63 is zero)
  uint32_t* HB[8];
  for (int l = 0; l < 8; ++l) HB[l] = reinterpret_cast<uint32_t*>(w + L.hb[l]);
  for (int l = 0; l < 8; ++l) {
    GemmEpi e;
    e.bias = F + FL.b_trunk[l];
    e.act = 1;
    e.bits_out = HB[l];
    if (l == 0) {
      if (gemm_w(f, 0, PE, 64, 64, nullptr, 0, H[0], W, P, e, st)) return 1;
    } else if (l == 4) {
      if (gemm_w(f, 4, PE, 64, 64, H[3], W, H[4], W, P, e, st)) return 1;
    } else {
      if (gemm_w(f, l, H[l - 1], W, W, nullptr, 0, H[l], W, P, e, st)) return 1;
    }
    if (!use_tc()) {  // the CUDA-core GEMM has no bit-mask epilogue
      const size_t nw = (size_t)P * 8;
      k_train_make_bits<<<(unsigned)((nw * 32 + 255) / 256), 256, 0, st>>>(H[l], nw, HB[l]);
      MNRF_LAUNCH_OK();
    }
  }
  // colour branch (mirror_nerf.py:199-204)
  {
    GemmEpi e;
    e.bias = F + FL.b_final;
    if (gemm_w(f, 8, H[7], W, W, nullptr, 0, w + L.f, W, P, e, st)) return 1;
    if (launch_dirbias(f, rays, n, 8, 0, w + L.dirbias, st)) return 1;  // b_dir + W_dir[:,256:] embed(d) per ray
    GemmEpi e2;
    e2.rowbias = w + L.dirbias; e2.rb_div = S; e2.ld_rb = WH; e2.act = 1;
    if (gemm_w(f, 10, w + L.f, W, W, nullptr, 0, w + L.d1, WH, P, e2, st)) return 1;
  }
  if (f->has_normal) {  // normal_net.0 (no activation, mirror_nerf.py:85-88)
    GemmEpi e;
    e.bias = F + FL.b_n0;
    if (gemm_w(f, 20, H[7], W, W, nullptr, 0, w + L.n1, WH, P, e, st)) return 1;
  }
  if (f->has_mirror) {  // is_mirror_net.0 + LeakyReLU (mirror_nerf.py:94-96)
    GemmEpi e;
    e.bias = F + FL.b_m0; e.act = 3;
    if (gemm_w(f, 9, H[7], W, W, nullptr, 0, w + L.m1, WH, P, e, st)) return 1;
  }
  HeadW hw{F + FL.w_sigma, F + FL.b_sigma, F + FL.w_rgb, F + FL.b_rgb, F + FL.w_n1, F + FL.b_n1, F + FL.w_m2, F + FL.b_m2};
  k_train_heads_fwd<<<148 * 4, 256, 0, st>>>(H[7], w + L.d1, f->has_normal ? w + L.n1 : nullptr,
                                             f->has_mirror ? w + L.m1 : nullptr, hw, P, w + L.raw, w + L.n2);
  MNRF_LAUNCH_OK();
  const float* normal = nullptr;
  if (cfg.compute_normal) {
    // reverse chain d sigma / d xyz (mirror_nerf.py:136-146); q_k kept for the double backward
    float* Q[8];
    for (int l = 0; l < 8; ++l) Q[l] = w + L.q[l];
    const size_t n4 = (size_t)P * (W / 4);
    k_train_chain_start<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(H[7], F + FL.w_sigma, n4, Q[7]);
    MNRF_LAUNCH_OK();
    for (int l = 7; l >= 1; --l) {  // q_{l-1} = (q_l W_l) * relu'(h_{l-1});  W_l = layer l+1 in 1-based naming
      GemmEpi e;
      e.act = 2; e.mask = H[l - 1]; e.ld_mask = W; e.mask_bits = HB[l - 1];
      if (gemm_w(f, chain_step(l), Q[l], W, W, nullptr, 0, Q[l - 1], W, P, e, st)) return 1;
    }
    GemmEpi e0;
    if (gemm_w(f, 19, Q[0], W, W, nullptr, 0, w + L.gpe, 64, P, e0, st)) return 1;
    GemmEpi e1;
    e1.accumulate = 1;
    if (gemm_w(f, 15, Q[4], W, W, nullptr, 0, w + L.gpe, 64, P, e1, st)) return 1;
    k_train_normal<<<(P + 127) / 128, 128, 0, st>>>(w + L.gpe, PE, P, w + L.gx, w + L.nrm);
    MNRF_LAUNCH_OK();
    normal = w + L.nrm;
    if (normal_out != nullptr)
      MNRF_CUDA_OK(cudaMemcpyAsync(normal_out, normal, sizeof(float) * (size_t)P * 3, cudaMemcpyDeviceToDevice, st));
  }
  return launch_composite(rays, z, w + L.raw, 8, w + L.raw, normal, noise, cfg.noise_std, n, S, cfg.white_back, out, st);
}

int train_pass_bwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                   const mnrf_train_cfg& cfg, const void* ws_fwd, void* ws_bwd, const mnrf_train_grads& g,
                   const float* ray_detach_mirror, float* const* gt, const float* depth, float* grad_rays,
                   cudaStream_t st) {
  const int S = cfg.S;
  const int P = n * S;
  const FwdWs L = fwd_layout(n, S, cfg.compute_normal);
  const BwdWs B = bwd_layout(n, S, cfg.compute_normal);
  const float* w = reinterpret_cast<const float*>(ws_fwd);
  float* b = reinterpret_cast<float*>(ws_bwd);
  const float* F = f->f32;
  const F32Layout& FL = f->L;
  const float* PE = w + L.pe;
  const float* H[8];
  for (int l = 0; l < 8; ++l) H[l] = w + L.h[l];
  const uint32_t* HB[8];
  for (int l = 0; l < 8; ++l) HB[l] = reinterpret_cast<const uint32_t*>(w + L.hb[l]);
  const float* normal = cfg.compute_normal ? w + L.nrm : nullptr;
  const bool hn = f->has_normal != 0, hm = f->has_mirror != 0;
  for (int i = 0; i < 24; ++i) MNRF_REQUIRE(gt[i] != nullptr, "train_pass_bwd: gradient tensor %d missing", i);
  if (hn) for (int i = 24; i < 28; ++i) MNRF_REQUIRE(gt[i] != nullptr, "train_pass_bwd: gradient tensor %d missing", i);
  if (hm) for (int i = 28; i < 32; ++i) MNRF_REQUIRE(gt[i] != nullptr, "train_pass_bwd: gradient tensor %d missing", i);

  // 1. compositor backward -> per-point record
  if (launch_composite_bwd(rays, z, w + L.raw, normal, noise, cfg, n, ray_detach_mirror, g, b + B.dr, st)) return 1;

  // 2. heads backward (small weights' gradients; gradients of the 128-wide hidden rows)
  HeadW hw{F + FL.w_sigma, F + FL.b_sigma, F + FL.w_rgb, F + FL.b_rgb, F + FL.w_n1, F + FL.b_n1, F + FL.w_m2, F + FL.b_m2};
  HeadG hg{gt[T_SIGMA_W], gt[T_SIGMA_B], gt[T_RGB_W], gt[T_RGB_B], gt[T_N1_W], gt[T_N1_B], gt[T_M2_W], gt[T_M2_B]};
  k_train_heads_bwd<<<148 * 2, 256, 0, st>>>(H[7], w + L.d1, hn ? w + L.n1 : nullptr, hm ? w + L.m1 : nullptr, w + L.raw,
                                             w + L.n2, b + B.dr, hw, P, b + B.dd1, b + B.dn1, b + B.dm1, hg);
  MNRF_LAUNCH_OK();

  // 3. colour branch: dir layer and final linear
  if (gemm_g(b + B.dd1, WH, WH, w + L.f, W, W, gt[T_DIR_W], W + IN_DIR, 0, W, P, st, gt[T_DIR_B])) return 1;
  k_train_sum_samples<<<n, WH, 0, st>>>(b + B.dd1, S, b + B.rsum);
  MNRF_LAUNCH_OK();
  k_train_dir_pe<<<(n + 127) / 128, 128, 0, st>>>(rays, n, b + B.dirpe);
  MNRF_LAUNCH_OK();
  if (gemm_g(b + B.rsum, WH, WH, b + B.dirpe, 64, 64, gt[T_DIR_W], W + IN_DIR, W, IN_DIR, n, st)) return 1;
  {
    GemmEpi e;  // dF = dD1pre * W_dir[:, :256]
    if (gemm_w(f, 21, b + B.dd1, WH, WH, nullptr, 0, b + B.df, W, P, e, st)) return 1;
  }
  if (gemm_g(b + B.df, W, W, H[7], W, W, gt[T_FINAL_W], W, 0, W, P, st, gt[T_FINAL_B])) return 1;
  // 4. normal / mirror head first layers
  if (hn) {
    if (gemm_g(b + B.dn1, WH, WH, H[7], W, W, gt[T_N0_W], W, 0, W, P, st, gt[T_N0_B])) return 1;
  }
  if (hm) {
    if (gemm_g(b + B.dm1, WH, WH, H[7], W, W, gt[T_M0_W], W, 0, W, P, st, gt[T_M0_B])) return 1;
  }
  // 5. dH8 = dF W_final + [dN1 W_n0] + [dM1 W_m0] + d sigma (x) w_sigma, then * relu'(h8) -> dZ8
  const bool use_n = hn && !cfg.detach_density_for_normal_loss;
  const bool use_m = hm && !cfg.detach_density_for_mask_loss;
  if (use_m && ray_detach_mirror != nullptr) {
    const size_t n4 = (size_t)P * (WH / 4);
    k_train_zero_rows<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(b + B.dm1, ray_detach_mirror, S, n4);
    MNRF_LAUNCH_OK();
  }
  float* dZ = b + B.dz[0];
  float* dZn = b + B.dz[1];
  {
    GemmEpi last;  // the last GEMM of the sum carries the rank-1 sigma term and the relu' mask
    last.rvec = b + B.dr; last.ld_rvec = DR_STRIDE; last.cvec = F + FL.w_sigma;
    last.act = 2; last.mask = H[7]; last.ld_mask = W; last.mask_bits = HB[7];
    GemmEpi plain;
    const int n_terms = 1 + (use_n ? 1 : 0) + (use_m ? 1 : 0);
    int term = 0;
    auto epi_for = [&](int t) {
      GemmEpi e = (t == n_terms - 1) ? last : plain;
      e.accumulate = t > 0;
      return e;
    };
    if (gemm_w(f, 22, b + B.df, W, W, nullptr, 0, dZ, W, P, epi_for(term++), st)) return 1;
    if (use_n) { if (gemm_w(f, 23, b + B.dn1, WH, WH, nullptr, 0, dZ, W, P, epi_for(term++), st)) return 1; }
    if (use_m) { if (gemm_w(f, 24, b + B.dm1, WH, WH, nullptr, 0, dZ, W, P, epi_for(term++), st)) return 1; }
  }
  // 6. trunk backward
  for (int l = 7; l >= 0; --l) {
    // bias and weight gradients of layer l (0-based) from dZ = dL/dz_l
    if (l == 0) {
      if (gemm_g(dZ, W, W, PE, 64, 64, gt[0], IN_XYZ, 0, IN_XYZ, P, st, gt[1])) return 1;
    } else if (l == 4) {
      if (gemm_g(dZ, W, W, PE, 64, 64, gt[8], IN_XYZ + W, 0, IN_XYZ, P, st, gt[9])) return 1;
      if (gemm_g(dZ, W, W, H[3], W, W, gt[8], IN_XYZ + W, IN_XYZ, W, P, st)) return 1;
    } else {
      if (gemm_g(dZ, W, W, H[l - 1], W, W, gt[2 * l], W, 0, W, P, st, gt[2 * l + 1])) return 1;
    }
    if (grad_rays != nullptr && (l == 4 || l == 0)) {  // dL/dPE = dZ_5 W_5[:, :63] + dZ_1 W_1
      GemmEpi e;
      e.accumulate = l == 0;
      if (gemm_w(f, l == 4 ? 15 : 19, dZ, W, W, nullptr, 0, b + B.dpe, 64, P, e, st)) return 1;
    }
    if (l > 0) {  // dZ_{l-1} = (dZ_l W_l) * relu'(h_{l-1})
      GemmEpi e;
      e.act = 2; e.mask = H[l - 1]; e.ld_mask = W; e.mask_bits = HB[l - 1];
      if (gemm_w(f, chain_step(l), dZ, W, W, nullptr, 0, dZn, W, P, e, st)) return 1;
      float* t = dZ; dZ = dZn; dZn = t;
    }
  }
  // 7. double backward through the analytic normal (tangent pass)
  const bool normal_grads = cfg.compute_normal && (g.normal != nullptr || g.surface_normal_grad != nullptr ||
                                                   (g.normal_dif != nullptr));
  if (normal_grads) {
    const float* Q[8];
    for (int l = 0; l < 8; ++l) Q[l] = w + L.q[l];
    float* T0 = b + B.t0;
    k_train_tangent_start<<<(P + 127) / 128, 128, 0, st>>>(b + B.dr, w + L.gx, PE, P, T0);
    MNRF_LAUNCH_OK();
    float* Tc = b + B.t[0];
    float* Tn = b + B.t[1];
    const float* Tprev = T0;  // t_{l-1} (t0 is 64 wide)
    for (int l = 0; l < 8; ++l) {
      // dW_l += q_l^T t_{l-1}
      if (l == 0) {
        if (gemm_g(Q[0], W, W, T0, 64, 64, gt[0], IN_XYZ, 0, IN_XYZ, P, st)) return 1;
      } else if (l == 4) {
        if (gemm_g(Q[4], W, W, T0, 64, 64, gt[8], IN_XYZ + W, 0, IN_XYZ, P, st)) return 1;
        if (gemm_g(Q[4], W, W, Tprev, W, W, gt[8], IN_XYZ + W, IN_XYZ, W, P, st)) return 1;
      } else {
        if (gemm_g(Q[l], W, W, Tprev, W, W, gt[2 * l], W, 0, W, P, st)) return 1;
      }
      // t_l = (t_{l-1} W_l^T) * relu'(h_l)
      GemmEpi e;
      e.act = 2; e.mask = H[l]; e.ld_mask = W; e.mask_bits = HB[l];
      if (l == 0) {
        if (gemm_w(f, 0, T0, 64, 64, nullptr, 0, Tc, W, P, e, st)) return 1;
      } else if (l == 4) {
        if (gemm_w(f, 4, T0, 64, 64, Tprev, W, Tc, W, P, e, st)) return 1;
      } else {
        if (gemm_w(f, l, Tprev, W, W, nullptr, 0, Tc, W, P, e, st)) return 1;
      }
      Tprev = Tc;
      float* t = Tc; Tc = Tn; Tn = t;
    }
    // q8 = w_sigma * relu'(h8):  d w_sigma += sum_p t8
    if (colsum(Tprev, W, P, W, gt[T_SIGMA_W], st)) return 1;
  }
  // 8. gradient w.r.t. the rays
  if (grad_rays != nullptr) {
    MNRF_REQUIRE(g.x_surface == nullptr || depth != nullptr, "train_pass_bwd: ray gradients need the depth output");
    k_train_ray_grad<<<n, 64, 0, st>>>(rays, z, PE, b + B.dpe, cfg.compute_normal ? w + L.gpe : nullptr,
                                       cfg.compute_normal ? w + L.gx : nullptr, b + B.dr, b + B.rsum, b + B.dirpe,
                                       F + FL.wt_dir, g.x_surface, depth, S, normal_grads ? 1 : 0, grad_rays);
    MNRF_LAUNCH_OK();
  }
  return 0;
}

}  // namespace mnrf

using namespace mnrf;

extern "C" {

int64_t mnrf_train_fwd_workspace_bytes(int n, int S, int compute_normal) {
  if (n < 0 || S < 1) return -1;
  return train_fwd_workspace_bytes(n, S, compute_normal);
}
int64_t mnrf_train_bwd_workspace_bytes(int n, int S, int compute_normal) {
  if (n < 0 || S < 1) return -1;
  return train_bwd_workspace_bytes(n, S, compute_normal);
}

int64_t mnrf_field_train_fwd_workspace_bytes(const mnrf_field* f, int n, int S, int compute_normal) {
  if (f == nullptr || n < 0 || S < 1) return -1;
  return f->kind == 1 ? hash_train_fwd_workspace_bytes(n, S, compute_normal) : train_fwd_workspace_bytes(n, S, compute_normal);
}
int64_t mnrf_field_train_bwd_workspace_bytes(const mnrf_field* f, int n, int S, int compute_normal) {
  if (f == nullptr || n < 0 || S < 1) return -1;
  return f->kind == 1 ? hash_train_bwd_workspace_bytes(n, S, compute_normal) : train_bwd_workspace_bytes(n, S, compute_normal);
}

int mnrf_train_pass_fwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                        const mnrf_train_cfg* cfg, void* ws, int64_t ws_bytes, const mnrf_composite_out* out,
                        float* normal_out, void* stream) {
  MNRF_REQUIRE(f && rays && z && cfg && ws && out, "train_pass_fwd: null argument");
  MNRF_REQUIRE(n >= 0 && cfg->S >= 1 && (long long)n * cfg->S < (1ll << 31) / 256, "train_pass_fwd: bad sizes");
  MNRF_REQUIRE(ws_bytes >= mnrf_field_train_fwd_workspace_bytes(f, n, cfg->S, cfg->compute_normal), "train_pass_fwd: workspace too small");
  if (n == 0) return 0;
  if (f->kind == 1)
    return hash_train_pass_fwd(f, rays, z, noise, n, *cfg, ws, *out, normal_out, reinterpret_cast<cudaStream_t>(stream));
  return train_pass_fwd(f, rays, z, noise, n, *cfg, ws, *out, normal_out, reinterpret_cast<cudaStream_t>(stream));
}

int mnrf_train_pass_bwd(const mnrf_field* f, const float* rays, const float* z, const float* noise, int n,
                        const mnrf_train_cfg* cfg, const void* ws_fwd, int64_t ws_fwd_bytes, void* ws_bwd,
                        int64_t ws_bwd_bytes, const mnrf_train_grads* grads, const float* ray_detach_mirror,
                        float* const* grad_tensors, const float* depth, float* grad_rays, void* stream) {
  MNRF_REQUIRE(f && rays && z && cfg && ws_fwd && ws_bwd && grads && grad_tensors, "train_pass_bwd: null argument");
  MNRF_REQUIRE(n >= 0 && cfg->S >= 1 && (long long)n * cfg->S < (1ll << 31) / 256, "train_pass_bwd: bad sizes");
  MNRF_REQUIRE(ws_fwd_bytes >= mnrf_field_train_fwd_workspace_bytes(f, n, cfg->S, cfg->compute_normal) &&
                   ws_bwd_bytes >= mnrf_field_train_bwd_workspace_bytes(f, n, cfg->S, cfg->compute_normal),
               "train_pass_bwd: workspace too small");
  if (n == 0) return 0;
  if (f->kind == 1)
    return hash_train_pass_bwd(f, rays, z, noise, n, *cfg, ws_fwd, ws_bwd, *grads, ray_detach_mirror, grad_tensors, depth,
                               grad_rays, reinterpret_cast<cudaStream_t>(stream));
  return train_pass_bwd(f, rays, z, noise, n, *cfg, ws_fwd, ws_bwd, *grads, ray_detach_mirror, grad_tensors, depth,
                        grad_rays, reinterpret_cast<cudaStream_t>(stream));
}

// Bring-up aid: time `iters` launches of one training GEMM on synthetic operands.  kind 0: NN step `step` on P rows; kind 1:
// TN 256x256 over P rows.  engine 1 = tcgen05, 0 = CUDA cores.  dbg: train_tc.cu debug bits.  Returns milliseconds per launch.
int mnrf_debug_gemm_bench(const mnrf_field* f, int kind, int step, int P, int engine, int dbg, int iters, float* ms_out) {
  MNRF_REQUIRE(f && ms_out && P > 0 && iters > 0 && f->kind == 0, "gemm_bench: bad argument");
  float *A = nullptr, *C = nullptr, *Wg = nullptr;
  // zero bias of the widest layer: with a bias the NN launch takes the bias + ReLU epilogue flavour of the forward trunk
  // (k_gemm_tc_nn<1>), which is what a training step runs; without one it would fall back to the generic run-time flavour
  float* bias_z = nullptr;
  MNRF_CUDA_OK(cudaMalloc(&bias_z, sizeof(float) * W));
  MNRF_CUDA_OK(cudaMemset(bias_z, 0, sizeof(float) * W));
  MNRF_CUDA_OK(cudaMalloc(&A, sizeof(float) * (size_t)P * W));
  MNRF_CUDA_OK(cudaMalloc(&C, sizeof(float) * (size_t)P * W));
  MNRF_CUDA_OK(cudaMalloc(&Wg, sizeof(float) * W * W));
  MNRF_CUDA_OK(cudaMemset(A, 0, sizeof(float) * (size_t)P * W));
  MNRF_CUDA_OK(cudaMemset(C, 0, sizeof(float) * (size_t)P * W));
  MNRF_CUDA_OK(cudaMemset(Wg, 0, sizeof(float) * W * W));
  const int saved = g_engine;
  g_engine = engine;
  set_train_tc_one_pass(engine == 2);
  set_train_tc_debug(dbg);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = 0;
  for (int it = -1; it < iters && rc == 0; ++it) {
    if (it == 0) cudaEventRecord(e0, 0);
    if (kind == 0) {
      GemmEpi e;
      e.act = 1;
      e.bias = bias_z;
      const int K = t32_step_k(step);
      rc = gemm_w(f, step, A, W, K > W ? 64 : K, K > W ? A : nullptr, W, C, W, P, e, 0);
    } else {
      rc = gemm_g(A, W, W, C, W, W, Wg, W, 0, W, P, 0);
    }
  }
  cudaEventRecord(e1, 0);
  cudaError_t err = cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_out = ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  set_train_tc_debug(0);
  g_engine = saved;
  set_train_tc_one_pass(saved == 2);
  cudaFree(A); cudaFree(C); cudaFree(Wg); cudaFree(bias_z);
  if (err != cudaSuccess) { set_error("gemm_bench: %s", cudaGetErrorString(err)); return 1; }
  return rc;
}

int mnrf_train_set_gemm(int engine) {
  MNRF_REQUIRE(engine >= 0 && engine <= 2, "train_set_gemm: engine must be 0 (CUDA cores), 1 (tf32 x3) or 2 (tf32 x1)");
  g_engine = engine;
  set_train_tc_one_pass(engine == 2);
  return 0;
}

int mnrf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  MNRF_REQUIRE(params && grads && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "adam_step: bad argument");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  k_adam<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, (long long)n, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2),
      grad_scale);
  MNRF_LAUNCH_OK();
  return 0;
}

int mnrf_peer_shard(int64_t n, int world, int rank, int64_t* lo, int64_t* hi) {
  MNRF_REQUIRE(n >= 0 && world >= 1 && rank >= 0 && rank < world && lo && hi, "peer_shard: bad argument");
  int64_t chunk = (n + world - 1) / world;
  chunk = (chunk + 3) / 4 * 4;  // 16-byte granules
  *lo = std::min<int64_t>(n, chunk * rank);
  *hi = std::min<int64_t>(n, chunk * (rank + 1));
  return 0;
}

int mnrf_peer_allreduce_adam(const uint64_t* grad_ptrs, const uint64_t* param_ptrs, int world, int rank, float* exp_avg_shard,
                             float* exp_avg_sq_shard, int64_t n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int step, void* stream) {
  MNRF_REQUIRE(grad_ptrs && param_ptrs && exp_avg_shard && exp_avg_sq_shard && step >= 1, "peer_allreduce_adam: bad argument");
  MNRF_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "peer_allreduce_adam: 1 <= world <= 8");
  MNRF_REQUIRE(n % 4 == 0, "peer_allreduce_adam: the flat buffers must be padded to a multiple of 4 elements");
  int64_t lo, hi;
  if (mnrf_peer_shard(n, world, rank, &lo, &hi)) return 2;
  if (hi <= lo) return 0;
  PeerPtrs g{}, p{};
  for (int r = 0; r < world; ++r) {
    MNRF_REQUIRE(grad_ptrs[r] != 0 && param_ptrs[r] != 0 && grad_ptrs[r] % 16 == 0 && param_ptrs[r] % 16 == 0,
                 "peer_allreduce_adam: peer buffer %d missing or not 16-byte aligned", r);
    g.p[r] = reinterpret_cast<float*>(grad_ptrs[r]);
    p.p[r] = reinterpret_cast<float*>(param_ptrs[r]);
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const long long quads = (hi - lo + 3) / 4;
  k_peer_allreduce_adam<<<(unsigned)((quads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      g, p, world, lo, hi, exp_avg_shard, exp_avg_sq_shard, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2),
      1.f / (float)world, rank);
  MNRF_LAUNCH_OK();
  return 0;
}

int mnrf_axpy(float* out, const float* in, int64_t n, float alpha, void* stream) {
  MNRF_REQUIRE(out && in && n >= 0, "axpy: bad argument");
  if (n == 0) return 0;
  k_axpy<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, in, (long long)n, alpha);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // extern "C"
