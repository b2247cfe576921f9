// Hash-grid field (BASELINE config 3): MirrorNeRFTcnn.forward with compute_normal=False (R/models/mirror_nerf_tcnn.py:151-259)
// for the samples of a ray batch, one thread per point:
//   x = o + d*z -> [0,1]^3 -> 16-level multiresolution hash encoding (2 features per level, 8 trilinear corners per level; the
//   algorithm of tinycudann's HashGrid: dense indexing while a level fits its table, coherent prime hash otherwise) ->
//   sigma_net 32->64->16 (sigma raw = out[0], geo_feat = out[1:16]) -> colour net [SH4(d) | geo] 31->64->64->3 sigmoid,
//   normal net 15->64->3 l2-normalised, mirror net 15->32->1 (LeakyReLU, biases, sigmoid)    -> 8 floats / point.
// This path is gather-bound (128 table reads of 8 bytes per point against a 46.5 MB fp32 table that lives mostly in the 126 MB
// L2) with ~11 k MAC per point, so it runs on the CUDA cores: weights (45 KB) are staged in shared memory and read as
// warp-uniform float4 broadcasts; every layer after the first of each net is consumed output-by-output, so only one 64-wide
// activation vector is live in registers.  fp32 throughout (tinycudann evaluates in fp16).
#include "common.cuh"

namespace mnrf {
namespace {

constexpr int HB = 128;  // threads per block

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// dot of a shared-memory weight row (K floats, K % 4 == 0) with a register vector
template <int K>
__device__ __forceinline__ float dotw(const float* __restrict__ w, const float (&v)[K]) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < K; k += 4) {
    const float4 q = *reinterpret_cast<const float4*>(w + k);
    acc = fmaf(q.x, v[k], acc); acc = fmaf(q.y, v[k + 1], acc); acc = fmaf(q.z, v[k + 2], acc); acc = fmaf(q.w, v[k + 3], acc);
  }
  return acc;
}

__global__ void __launch_bounds__(HB) k_field_hash(const float* __restrict__ table, const float* __restrict__ wpack, HashGridMeta M,
                                                   FieldIO io, int has_normal, int has_mirror) {
  extern __shared__ __align__(16) float sw[];
  for (int i = threadIdx.x; i < HW_TOTAL / 4; i += HB)
    reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(wpack)[i];
  __syncthreads();
  const long long p = (long long)blockIdx.x * HB + threadIdx.x;
  if (p >= io.n_points) return;
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

  // ---- multiresolution hash encoding ----
  float enc[32];
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
      const float2 f = __ldg(tab + M.offset[l] + index);
      a0 = __fadd_rn(a0, __fmul_rn(w, f.x));
      a1 = __fadd_rn(a1, __fmul_rn(w, f.y));
    }
    enc[2 * l] = a0;
    enc[2 * l + 1] = a1;
  }

  // ---- sigma_net: 32 -> 64 (ReLU) -> 16 ----
  float h[64];
#pragma unroll
  for (int o = 0; o < 64; ++o) h[o] = fmaxf(dotw<32>(sw + HW_S0 + o * 32, enc), 0.f);
  float geo[16];  // [sigma, geo_feat(15)]
#pragma unroll
  for (int o = 0; o < 16; ++o) geo[o] = dotw<64>(sw + HW_S1 + o * 64, h);
  const float sigma = geo[0];
  if (io.sigma_only) {
    if (io.sigma_out != nullptr) io.sigma_out[p] = sigma;
    if (io.raw == nullptr) return;
  }
  float g16[16];  // geo_feat padded to 16
#pragma unroll
  for (int i = 0; i < 15; ++i) g16[i] = geo[1 + i];
  g16[15] = 0.f;

  float o_n[3] = {0.f, 0.f, 0.f}, o_rgb[3] = {0.f, 0.f, 0.f}, o_m = 0.f;
  if (has_normal) {  // 15 -> 64 (ReLU) -> 3, l2-normalised
    float n2[3] = {0.f, 0.f, 0.f};
#pragma unroll 8
    for (int o = 0; o < 64; ++o) {
      const float v = fmaxf(dotw<16>(sw + HW_N0 + o * 16, g16), 0.f);
      n2[0] = fmaf(sw[HW_N1 + o], v, n2[0]); n2[1] = fmaf(sw[HW_N1 + 64 + o], v, n2[1]); n2[2] = fmaf(sw[HW_N1 + 128 + o], v, n2[2]);
    }
    const float len = sqrtf(fmaxf(n2[0] * n2[0] + n2[1] * n2[1] + n2[2] * n2[2], FP32_EPS));
    o_n[0] = n2[0] / len; o_n[1] = n2[1] / len; o_n[2] = n2[2] / len;
  }
  if (!io.sigma_only) {
    // colour net: [SH4(d) (16) | geo_feat (15) | 0] -> 64 (ReLU) -> 64 (ReLU) -> 3 sigmoid
    float in[32];
    {
      const float X = d[0], Y = d[1], Z = d[2];
      const float xy = X * Y, xz = X * Z, yz = Y * Z, x2 = X * X, y2 = Y * Y, z2 = Z * Z;
      in[0] = 0.28209479177387814f;
      in[1] = -0.48860251190291987f * Y; in[2] = 0.48860251190291987f * Z; in[3] = -0.48860251190291987f * X;
      in[4] = 1.0925484305920792f * xy; in[5] = -1.0925484305920792f * yz;
      in[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
      in[7] = -1.0925484305920792f * xz; in[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
      in[9] = 0.59004358992664352f * Y * (-3.0f * x2 + y2); in[10] = 2.8906114426405538f * xy * Z;
      in[11] = 0.45704579946446572f * Y * (1.0f - 5.0f * z2); in[12] = 0.3731763325901154f * Z * (5.0f * z2 - 3.0f);
      in[13] = 0.45704579946446572f * X * (1.0f - 5.0f * z2); in[14] = 1.4453057213202769f * Z * (x2 - y2);
      in[15] = 0.59004358992664352f * X * (-x2 + 3.0f * y2);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) in[16 + i] = g16[i];
#pragma unroll
    for (int o = 0; o < 64; ++o) h[o] = fmaxf(dotw<32>(sw + HW_C0 + o * 32, in), 0.f);
    float c[3] = {0.f, 0.f, 0.f};
#pragma unroll 4
    for (int o = 0; o < 64; ++o) {
      const float v = fmaxf(dotw<64>(sw + HW_C1 + o * 64, h), 0.f);
      c[0] = fmaf(sw[HW_C2 + o], v, c[0]); c[1] = fmaf(sw[HW_C2 + 64 + o], v, c[1]); c[2] = fmaf(sw[HW_C2 + 128 + o], v, c[2]);
    }
    o_rgb[0] = sigmoidf_(c[0]); o_rgb[1] = sigmoidf_(c[1]); o_rgb[2] = sigmoidf_(c[2]);
    if (has_mirror) {  // 15 -> 32 (+bias, LeakyReLU 0.01) -> 1 (+bias) sigmoid
      float acc = sw[HW_M2B];
#pragma unroll 8
      for (int o = 0; o < 32; ++o) {
        float v = dotw<16>(sw + HW_M0 + o * 16, g16) + sw[HW_M0B + o];
        v = v > 0.f ? v : 0.01f * v;
        acc = fmaf(sw[HW_M2 + o], v, acc);
      }
      o_m = sigmoidf_(acc);
    }
  }
  if (io.raw != nullptr) {
    float4* o = reinterpret_cast<float4*>(io.raw + p * 8);
    o[0] = make_float4(sigma, o_rgb[0], o_rgb[1], o_rgb[2]);
    o[1] = make_float4(o_m, o_n[0], o_n[1], o_n[2]);
  }
  if (io.sigma_out != nullptr && !io.sigma_only) io.sigma_out[p] = sigma;
}

// dst[r][c] (cols_pad wide) = c < cols ? src[r][c0 + c] : 0
__global__ void k_hash_pack_rows(const float* __restrict__ src, float* __restrict__ dst, int rows, int ld, int cols, int cols_pad,
                                 int rows_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_pad * cols_pad) return;
  const int r = i / cols_pad, c = i % cols_pad;
  dst[i] = (src != nullptr && r < rows && c < cols) ? src[(size_t)r * ld + c] : 0.f;
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
  if (pack_rows(t[1], w + HW_S0, 64, 32, 32, 64, st)) return 1;
  if (pack_rows(t[2], w + HW_S1, 16, 64, 64, 16, st)) return 1;
  if (pack_rows(t[3], w + HW_C0, 64, 31, 32, 64, st)) return 1;
  if (pack_rows(t[4], w + HW_C1, 64, 64, 64, 64, st)) return 1;
  if (pack_rows(t[5], w + HW_C2, 3, 64, 64, 4, st)) return 1;
  if (pack_rows(t[6], w + HW_N0, 64, 15, 16, 64, st)) return 1;
  if (pack_rows(t[7], w + HW_N1, 3, 64, 64, 4, st)) return 1;
  if (pack_rows(t[8], w + HW_M0, 32, 15, 16, 32, st)) return 1;
  if (pack_rows(t[9], w + HW_M0B, 1, 32, 32, 1, st)) return 1;
  if (pack_rows(t[10], w + HW_M2, 1, 32, 32, 1, st)) return 1;
  if (pack_rows(t[11], w + HW_M2B, 1, 1, 4, 1, st)) return 1;
  return 0;
}

int launch_field_hash(const mnrf_field* f, const FieldIO& io, cudaStream_t st) {
  if (io.n_points <= 0) return 0;
  MNRF_REQUIRE(io.normal_out == nullptr, "hash-grid field: analytic normals (compute_normal=True) are not built; use the predicted normals");
  MNRF_REQUIRE(io.geo_out == nullptr, "hash-grid field: geo_feat export is not built");
  static bool attr = false;
  if (!attr) {
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_hash, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(HW_TOTAL * sizeof(float))));
    attr = true;
  }
  const long long blocks = ((long long)io.n_points + HB - 1) / HB;
  k_field_hash<<<(unsigned)blocks, HB, HW_TOTAL * sizeof(float), st>>>(f->hash_table, f->hash_w, f->hg, io, f->has_normal, f->has_mirror);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace mnrf
