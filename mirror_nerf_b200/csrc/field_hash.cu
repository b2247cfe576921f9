// Hash-grid field (BASELINE config 3): MirrorNeRFTcnn.forward with compute_normal=False (R/models/mirror_nerf_tcnn.py:151-259)
// for the samples of a ray batch, one thread per point:
//   x = o + d*z -> [0,1]^3 -> 16-level multiresolution hash encoding (2 features per level, 8 trilinear corners per level; the
//   algorithm of tinycudann's HashGrid: dense indexing while a level fits its table, coherent prime hash otherwise) ->
//   sigma_net 32->64->16 (sigma raw = out[0], geo_feat = out[1:16]) -> colour net [SH4(d) | geo] 31->64->64->3 sigmoid,
//   normal net 15->64->3 l2-normalised, mirror net 15->32->1 (LeakyReLU, biases, sigmoid)    -> 8 floats / point.
// This path is gather-bound (128 table reads of 8 bytes per point against a 46.5 MB fp32 table that lives mostly in the 126 MB
// L2) with ~11 k MAC per point, so it runs on the CUDA cores.  Gather phase: one thread per point.  MLP phase: each warp treats
// its 32 points as a small GEMM batch -- activations live in a per-warp shared-memory buffer (feature-major [k][32 points]),
// weights (45 KB, transposed to [k][out]) in shared memory, and every lane accumulates a register tile of 4 points x 16 (8, 4)
// outputs, i.e. 64 FMAs per 5 shared-memory loads instead of 4 per load with one point per thread (the first version of this
// kernel was shared-memory-bandwidth bound: ncu L1/TEX 77 %).  The tiny last layers (64->3, 64->3, 32->1) are fused into
// the tile epilogue of the layer before them (partial dots + two shuffles).  fp32 throughout (tinycudann evaluates in fp16).
#include "common.cuh"
#include "hash_train_math.cuh"

namespace mnrf {
namespace {

static_assert(HASH_WREF_FLOATS == ht::HT_NW, "row-padded weight block size");
constexpr int HB = 128;        // threads per block = 4 warps x 32 points
constexpr int HX = 64 * 32;    // floats of one per-warp activation buffer: [64 features][32 points]
constexpr int H_SMEM_FLOATS = HW_TOTAL + 4 * 2 * HX;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// acc[p][j] += sum_k X[k][4*pg + p] * Wt[k][o0 + j]      X: [K][32] floats (feature-major), Wt: [K][N] floats
template <int NO>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ X, const float* __restrict__ Wt, int N, int K, int pg, int o0,
                                          float (&acc)[4][NO]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 x = *reinterpret_cast<const float4*>(X + k * 32 + 4 * pg);
#pragma unroll
    for (int j = 0; j < NO; j += 4) {
      const float4 w = *reinterpret_cast<const float4*>(Wt + k * N + o0 + j);
      acc[0][j] = fmaf(x.x, w.x, acc[0][j]); acc[0][j + 1] = fmaf(x.x, w.y, acc[0][j + 1]); acc[0][j + 2] = fmaf(x.x, w.z, acc[0][j + 2]); acc[0][j + 3] = fmaf(x.x, w.w, acc[0][j + 3]);
      acc[1][j] = fmaf(x.y, w.x, acc[1][j]); acc[1][j + 1] = fmaf(x.y, w.y, acc[1][j + 1]); acc[1][j + 2] = fmaf(x.y, w.z, acc[1][j + 2]); acc[1][j + 3] = fmaf(x.y, w.w, acc[1][j + 3]);
      acc[2][j] = fmaf(x.z, w.x, acc[2][j]); acc[2][j + 1] = fmaf(x.z, w.y, acc[2][j + 1]); acc[2][j + 2] = fmaf(x.z, w.z, acc[2][j + 2]); acc[2][j + 3] = fmaf(x.z, w.w, acc[2][j + 3]);
      acc[3][j] = fmaf(x.w, w.x, acc[3][j]); acc[3][j + 1] = fmaf(x.w, w.y, acc[3][j + 1]); acc[3][j + 2] = fmaf(x.w, w.z, acc[3][j + 2]); acc[3][j + 3] = fmaf(x.w, w.w, acc[3][j + 3]);
    }
  }
}
template <int NO>
__device__ __forceinline__ void tile_zero(float (&acc)[4][NO]) {
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < NO; ++j) acc[p][j] = 0.f;
}
// Y[o0 + j][4*pg .. 4*pg+3] = acc[.][j]
template <int NO>
__device__ __forceinline__ void tile_store(float* __restrict__ Y, int pg, int o0, const float (&acc)[4][NO]) {
#pragma unroll
  for (int j = 0; j < NO; ++j)
    *reinterpret_cast<float4*>(Y + (o0 + j) * 32 + 4 * pg) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
}
// fused last layer: out[c][point] = sum_o Wf[c][o] * hidden[o][point] for c < NC; partial dots over this lane's NO outputs, summed
// over the 4 lanes that share the point group (lane ^ 8, lane ^ 16), delivered to the owning threads through Y[c][point]
template <int NO, int NC>
__device__ __forceinline__ void tile_final(const float (&hid)[4][NO], const float* __restrict__ Wf, int ldw, int pg, int og, int o0,
                                           float* __restrict__ Y) {
  float part[4][NC];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < NC; ++c) part[p][c] = 0.f;
#pragma unroll
  for (int j = 0; j < NO; ++j)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float w = Wf[c * ldw + o0 + j];
#pragma unroll
      for (int p = 0; p < 4; ++p) part[p][c] = fmaf(hid[p][j], w, part[p][c]);
    }
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      part[p][c] += __shfl_xor_sync(0xffffffffu, part[p][c], 8);
      part[p][c] += __shfl_xor_sync(0xffffffffu, part[p][c], 16);
    }
  if (og == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
      *reinterpret_cast<float4*>(Y + c * 32 + 4 * pg) = make_float4(part[0][c], part[1][c], part[2][c], part[3][c]);
  }
}

// NORMALS: also the analytic normal normalize(-d sigma / d xyz) (mirror_nerf_tcnn.py:170-178: autograd through the encoder in the
// reference): d sigma/d enc = W0^T (W1[0,:] * relu'(h)), then a second pass over the 8 corners of every level with the derivative
// of the trilinear weights (d frac / d u = scale_l, d u / d x = 1 / (2 bound); floor() has no gradient).
template <bool NORMALS>
__global__ void __launch_bounds__(HB) k_field_hash(const float* __restrict__ table, const float* __restrict__ wpack, HashGridMeta M,
                                                   FieldIO io, int has_normal, int has_mirror) {
  extern __shared__ __align__(16) float sw[];
  if (io.n_rays_dev != nullptr) {  // device-side ray count (mnrf_render_recursive)
    io.n_points = min(io.n_points, max(__ldg(io.n_rays_dev), 0) * io.S);
    if ((long long)blockIdx.x * HB >= io.n_points) return;
  }
  for (int i = threadIdx.x; i < HW_TOTAL / 4; i += HB)
    reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(wpack)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pg = lane & 7, og = lane >> 3;
  float* X0 = sw + HW_TOTAL + warp * 2 * HX;
  float* X1 = X0 + HX;
  const long long p_raw = (long long)blockIdx.x * HB + threadIdx.x;
  const bool valid = p_raw < io.n_points;
  const long long p = valid ? p_raw : (long long)io.n_points - 1;  // tail threads shadow the last point (warp-collective MLP)
  float x[3], d[3] = {0.f, 0.f, 0.f};
  if (io.rays != nullptr) {
    const float* rr = io.rays + (p / io.S) * 8;
    const float z = io.z[p];
#pragma unroll
    for (int c = 0; c < 3; ++c) { x[c] = __fadd_rn(rr[c], __fmul_rn(rr[3 + c], z)); d[c] = rr[3 + c]; }
  } else {
    const float* xr = io.x + p * io.x_stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) { x[c] = xr[c]; if (!io.sigma_only) d[c] = xr[3 + c]; }
  }
  // to [0,1] (mirror_nerf_tcnn.py:224)
  float u[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) u[c] = __fdiv_rn(__fadd_rn(x[c], M.bound), __fmul_rn(2.f, M.bound));

  // ---- multiresolution hash encoding (thread = point) -> X0[2l + f][lane] ----
  const float2* tab = reinterpret_cast<const float2*>(table);
#pragma unroll 1
  for (int l = 0; l < HG_LEVELS; ++l) {
    const float scale = M.scale[l];
    const unsigned int res = (unsigned int)M.res[l], size = M.size[l];
    unsigned int g[3];
    float fr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float pos = __fadd_rn(__fmul_rn(u[c], scale), 0.5f);
      const float fl = floorf(pos);
      g[c] = (unsigned int)(int)fl;
      fr[c] = __fsub_rn(pos, fl);
    }
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
      float w = 1.f;
      unsigned int c3[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int bit = (corner >> c) & 1;
        w = __fmul_rn(w, bit ? fr[c] : __fsub_rn(1.f, fr[c]));
        c3[c] = g[c] + bit;
      }
      unsigned int stride = 1, index = 0;
      int dim = 0;
      for (; dim < 3 && stride <= size; ++dim) { index += c3[dim] * stride; stride *= res; }
      if (size < stride) index = (c3[0] * 1u) ^ (c3[1] * 2654435761u) ^ (c3[2] * 805459861u);
      index %= size;
      const float2 f = __ldg(tab + M.offset[l] + index);  // through L1: __ldcg (L2 only) measured 15 % slower, coarse levels hit in L1
      a0 = __fadd_rn(a0, __fmul_rn(w, f.x));
      a1 = __fadd_rn(a1, __fmul_rn(w, f.y));
    }
    X0[(2 * l) * 32 + lane] = a0;
    X0[(2 * l + 1) * 32 + lane] = a1;
  }
  __syncwarp();

  // ---- sigma_net: 32 -> 64 (ReLU) -> 16; [sigma, geo_feat(15)] ends up in X0 rows 0..15 ----
  {
    float acc[4][16];
    tile_zero<16>(acc);
    tile_gemm<16>(X0, sw + HW_S0, 64, 32, pg, 16 * og, acc);
#pragma unroll
    for (int pp = 0; pp < 4; ++pp)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[pp][j] = fmaxf(acc[pp][j], 0.f);
    tile_store<16>(X1, pg, 16 * og, acc);
  }
  __syncwarp();
  {
    float acc[4][4];
    tile_zero<4>(acc);
    tile_gemm<4>(X1, sw + HW_S1, 16, 64, pg, 4 * og, acc);
    __syncwarp();  // every lane is done reading X0 (the encoding) before it is overwritten
    tile_store<4>(X0, pg, 4 * og, acc);
  }
  __syncwarp();
  float o_an[3] = {0.f, 0.f, 0.f};
  if (NORMALS && !io.sigma_only) {
    // g_h = W1[0,:] * relu'(h) (thread = point), g_enc[k] = sum_o W0[o][k] g_h[o] -> X1 rows 0..31 (the hidden layer is consumed)
    float gh[64];
#pragma unroll
    for (int o = 0; o < 64; ++o) gh[o] = X1[o * 32 + lane] > 0.f ? sw[HW_S1 + o * 16] : 0.f;
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int o = 0; o < 64; o += 4) {
        const float4 w = *reinterpret_cast<const float4*>(sw + HW_S0 + k * 64 + o);
        acc = fmaf(w.x, gh[o], acc); acc = fmaf(w.y, gh[o + 1], acc); acc = fmaf(w.z, gh[o + 2], acc); acc = fmaf(w.w, gh[o + 3], acc);
      }
      X1[k * 32 + lane] = acc;
    }
    __syncwarp();
    float gu[3] = {0.f, 0.f, 0.f};  // d sigma / d u
#pragma unroll 1
    for (int l = 0; l < HG_LEVELS; ++l) {
      const float scale = M.scale[l];
      const unsigned int res = (unsigned int)M.res[l], size = M.size[l];
      unsigned int g[3];
      float fr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float pos = __fadd_rn(__fmul_rn(u[c], scale), 0.5f);
        const float fl = floorf(pos);
        g[c] = (unsigned int)(int)fl;
        fr[c] = __fsub_rn(pos, fl);
      }
      const float ge0 = X1[(2 * l) * 32 + lane], ge1 = X1[(2 * l + 1) * 32 + lane];
      float dl[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int corner = 0; corner < 8; ++corner) {
        unsigned int c3[3];
        float wd[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int bit = (corner >> c) & 1;
          wd[c] = bit ? fr[c] : 1.f - fr[c];
          c3[c] = g[c] + bit;
        }
        unsigned int stride = 1, index = 0;
        int dim = 0;
        for (; dim < 3 && stride <= size; ++dim) { index += c3[dim] * stride; stride *= res; }
        if (size < stride) index = (c3[0] * 1u) ^ (c3[1] * 2654435761u) ^ (c3[2] * 805459861u);
        index %= size;
        const float2 f = __ldg(tab + M.offset[l] + index);
        const float v = ge0 * f.x + ge1 * f.y;
        dl[0] += ((corner & 1) ? v : -v) * wd[1] * wd[2];
        dl[1] += ((corner & 2) ? v : -v) * wd[0] * wd[2];
        dl[2] += ((corner & 4) ? v : -v) * wd[0] * wd[1];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) gu[c] = fmaf(scale, dl[c], gu[c]);
    }
    const float inv2b = 1.f / (2.f * M.bound);
    const float a = -gu[0] * inv2b, b = -gu[1] * inv2b, cc = -gu[2] * inv2b;
    const float len = sqrtf(fmaxf(a * a + b * b + cc * cc, FP32_EPS));  // utils/func.py:5-7
    o_an[0] = a / len; o_an[1] = b / len; o_an[2] = cc / len;
    __syncwarp();
  }
  const float sigma = X0[lane];
  if (io.sigma_only) {
    if (io.sigma_out != nullptr && valid) io.sigma_out[p_raw] = sigma;
    if (io.raw == nullptr) return;  // warp-uniform
  }
  const float* GEO = X0 + 32;  // rows 1..15 = geo_feat
  float o_n[3] = {0.f, 0.f, 0.f}, o_rgb[3] = {0.f, 0.f, 0.f}, o_m = 0.f;

  if (has_normal) {  // 15 -> 64 (ReLU) -> 3, l2-normalised
    float acc[4][16];
    tile_zero<16>(acc);
    tile_gemm<16>(GEO, sw + HW_N0, 64, 15, pg, 16 * og, acc);
#pragma unroll
    for (int pp = 0; pp < 4; ++pp)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[pp][j] = fmaxf(acc[pp][j], 0.f);
    tile_final<16, 3>(acc, sw + HW_N1, 64, pg, og, 16 * og, X1);
    __syncwarp();
    const float n0 = X1[lane], n1 = X1[32 + lane], n2 = X1[64 + lane];
    const float len = sqrtf(fmaxf(n0 * n0 + n1 * n1 + n2 * n2, FP32_EPS));
    o_n[0] = n0 / len; o_n[1] = n1 / len; o_n[2] = n2 / len;
    __syncwarp();
  }
  if (!io.sigma_only) {
    if (has_mirror) {  // 15 -> 32 (+bias, LeakyReLU 0.01) -> 1 (+bias) sigmoid
      float acc[4][8];
      tile_zero<8>(acc);
      tile_gemm<8>(GEO, sw + HW_M0, 32, 15, pg, 8 * og, acc);
#pragma unroll
      for (int pp = 0; pp < 4; ++pp)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = acc[pp][j] + sw[HW_M0B + 8 * og + j];
          acc[pp][j] = v > 0.f ? v : 0.01f * v;
        }
      tile_final<8, 1>(acc, sw + HW_M2, 32, pg, og, 8 * og, X1);
      __syncwarp();
      o_m = sigmoidf_(X1[lane] + sw[HW_M2B]);
      __syncwarp();
    }
    // colour net input [SH4(d) (16) | geo_feat (15) | 0] -> X1 rows 0..31
    {
      const float X = d[0], Y = d[1], Z = d[2];
      const float xy = X * Y, xz = X * Z, yz = Y * Z, x2 = X * X, y2 = Y * Y, z2 = Z * Z;
      float sh[16];
      sh[0] = 0.28209479177387814f;
      sh[1] = -0.48860251190291987f * Y; sh[2] = 0.48860251190291987f * Z; sh[3] = -0.48860251190291987f * X;
      sh[4] = 1.0925484305920792f * xy; sh[5] = -1.0925484305920792f * yz;
      sh[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
      sh[7] = -1.0925484305920792f * xz; sh[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
      sh[9] = 0.59004358992664352f * Y * (-3.0f * x2 + y2); sh[10] = 2.8906114426405538f * xy * Z;
      sh[11] = 0.45704579946446572f * Y * (1.0f - 5.0f * z2); sh[12] = 0.3731763325901154f * Z * (5.0f * z2 - 3.0f);
      sh[13] = 0.45704579946446572f * X * (1.0f - 5.0f * z2); sh[14] = 1.4453057213202769f * Z * (x2 - y2);
      sh[15] = 0.59004358992664352f * X * (-x2 + 3.0f * y2);
#pragma unroll
      for (int i = 0; i < 16; ++i) X1[i * 32 + lane] = sh[i];
#pragma unroll
      for (int i = 0; i < 15; ++i) X1[(16 + i) * 32 + lane] = GEO[i * 32 + lane];
      X1[31 * 32 + lane] = 0.f;
    }
    __syncwarp();
    {
      float acc[4][16];
      tile_zero<16>(acc);
      tile_gemm<16>(X1, sw + HW_C0, 64, 32, pg, 16 * og, acc);
#pragma unroll
      for (int pp = 0; pp < 4; ++pp)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[pp][j] = fmaxf(acc[pp][j], 0.f);
      tile_store<16>(X0, pg, 16 * og, acc);  // geo_feat is no longer needed
    }
    __syncwarp();
    {
      float acc[4][16];
      tile_zero<16>(acc);
      tile_gemm<16>(X0, sw + HW_C1, 64, 64, pg, 16 * og, acc);
#pragma unroll
      for (int pp = 0; pp < 4; ++pp)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[pp][j] = fmaxf(acc[pp][j], 0.f);
      tile_final<16, 3>(acc, sw + HW_C2, 64, pg, og, 16 * og, X1);
    }
    __syncwarp();
    o_rgb[0] = sigmoidf_(X1[lane]); o_rgb[1] = sigmoidf_(X1[32 + lane]); o_rgb[2] = sigmoidf_(X1[64 + lane]);
  }
  if (io.raw != nullptr && valid) {
    float4* o = reinterpret_cast<float4*>(io.raw + p_raw * 8);
    o[0] = make_float4(sigma, o_rgb[0], o_rgb[1], o_rgb[2]);
    o[1] = make_float4(o_m, o_n[0], o_n[1], o_n[2]);
  }
  if (io.sigma_out != nullptr && !io.sigma_only && valid) io.sigma_out[p_raw] = sigma;
  if (NORMALS && io.normal_out != nullptr && valid) {
    io.normal_out[p_raw * 3 + 0] = o_an[0]; io.normal_out[p_raw * 3 + 1] = o_an[1]; io.normal_out[p_raw * 3 + 2] = o_an[2];
  }
}

// hidden-layer weights are stored transposed: dst[k][o] (N wide) = k < K && o < N ? src[o][k] : 0
__global__ void k_hash_pack_t(const float* __restrict__ src, float* __restrict__ dst, int N, int K, int Kpad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kpad * N) return;
  const int k = i / N, o = i % N;
  dst[i] = (src != nullptr && k < K) ? src[(size_t)o * K + k] : 0.f;
}

// dst[r][c] (cols_pad wide) = c < cols ? src[r][c0 + c] : 0
__global__ void k_hash_pack_rows(const float* __restrict__ src, float* __restrict__ dst, int rows, int ld, int cols, int cols_pad,
                                 int rows_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_pad * cols_pad) return;
  const int r = i / cols_pad, c = i % cols_pad;
  dst[i] = (src != nullptr && r < rows && c < cols) ? src[(size_t)r * ld + c] : 0.f;
}

// the row-padded [out][in] image of small tensor t (hash_train_math.cuh): dst[r][small_col(t, c)] = src[r][c]
__global__ void k_hash_pack_wref(const float* __restrict__ src, float* __restrict__ dst, int t, int rows, int cols, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i - r * cols;
  dst[r * ld + ht::small_col(t, c)] = src[i];
}

int pack_rows(const float* src, float* dst, int rows, int cols, int cols_pad, int rows_pad, cudaStream_t st) {
  const int n = rows_pad * cols_pad;
  k_hash_pack_rows<<<(n + 255) / 256, 256, 0, st>>>(src, dst, rows, cols, cols, cols_pad, rows_pad);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace

// tensors (device, fp32): 0 encoder.params | 1,2 sigma_net.{0,1}.weight | 3,4,5 color_net.{0,1,2}.weight |
//   6,7 normal_net.{0,1}.weight or NULL | 8,9 is_mirror_net.0.{weight,bias}, 10,11 is_mirror_net.2.{weight,bias} or NULL
int pack_hash_field(mnrf_field* f, const float* const* t, long long table_floats, cudaStream_t st) {
  MNRF_CUDA_OK(cudaMemcpyAsync(f->hash_table, t[0], sizeof(float) * (size_t)table_floats, cudaMemcpyDeviceToDevice, st));
  float* w = f->hash_w;
  auto pack_t = [&](const float* src, float* dst, int N, int K, int Kpad) {  // dst[k][o] = src[o][k]
    k_hash_pack_t<<<(Kpad * N + 255) / 256, 256, 0, st>>>(src, dst, N, K, Kpad);
    MNRF_LAUNCH_OK();
    return 0;
  };
  if (pack_t(t[1], w + HW_S0, 64, 32, 32)) return 1;
  if (pack_t(t[2], w + HW_S1, 16, 64, 64)) return 1;
  if (pack_t(t[3], w + HW_C0, 64, 31, 32)) return 1;
  if (pack_t(t[4], w + HW_C1, 64, 64, 64)) return 1;
  if (pack_rows(t[5], w + HW_C2, 3, 64, 64, 4, st)) return 1;
  if (pack_t(t[6], w + HW_N0, 64, 15, 16)) return 1;
  if (pack_rows(t[7], w + HW_N1, 3, 64, 64, 4, st)) return 1;
  if (pack_t(t[8], w + HW_M0, 32, 15, 16)) return 1;
  if (pack_rows(t[9], w + HW_M0B, 1, 32, 32, 1, st)) return 1;
  if (pack_rows(t[10], w + HW_M2, 1, 32, 32, 1, st)) return 1;
  if (pack_rows(t[11], w + HW_M2B, 1, 1, 4, 1, st)) return 1;
  // [out][in] image with 16-byte aligned rows for the backward (train_hash.cu); pad columns and absent heads stay zero
  MNRF_CUDA_OK(cudaMemsetAsync(f->hash_wref, 0, sizeof(float) * HASH_WREF_FLOATS, st));
  for (int i = 1; i < 12; ++i) {
    if (t[i] == nullptr) continue;
    const int n = ht::small_rows(i) * ht::small_cols(i);
    k_hash_pack_wref<<<(n + 255) / 256, 256, 0, st>>>(t[i], f->hash_wref + ht::small_offset(i), i, ht::small_rows(i),
                                                      ht::small_cols(i), ht::small_ld(i));
    MNRF_LAUNCH_OK();
  }
  return 0;
}

int launch_field_hash(const mnrf_field* f, const FieldIO& io, cudaStream_t st) {
  if (io.n_points <= 0) return 0;
  MNRF_REQUIRE(io.normal_out == nullptr || !io.sigma_only, "hash-grid field: analytic normals need the full (non sigma-only) pass");
  MNRF_REQUIRE(io.geo_out == nullptr, "hash-grid field: geo_feat export is not built");
  if (first_use_on_device(TAG_FIELD_HASH)) {
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_hash<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(H_SMEM_FLOATS * sizeof(float))));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_hash<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(H_SMEM_FLOATS * sizeof(float))));
  }
  const long long blocks = ((long long)io.n_points + HB - 1) / HB;
  if (io.normal_out != nullptr)
    k_field_hash<true><<<(unsigned)blocks, HB, H_SMEM_FLOATS * sizeof(float), st>>>(f->hash_table, f->hash_w, f->hg, io, f->has_normal, f->has_mirror);
  else
    k_field_hash<false><<<(unsigned)blocks, HB, H_SMEM_FLOATS * sizeof(float), st>>>(f->hash_table, f->hash_w, f->hg, io, f->has_normal, f->has_mirror);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace mnrf
