// Weight packing: MirrorNeRF.state_dict() tensors ([out,in] fp32, R/models/mirror_nerf.py:60-99)
//   -> (a) fp32 section: transposed copies, biases, folded normal head, epilogue tables
//   -> (b) tensor-core section: per GEMM step, per K32 chunk, an fp16 "hi" blob and an fp16 "lo" blob
//          (W*2^s = hi + lo) in the tcgen05 K-major no-swizzle core-matrix layout, so that one
//          cp.async.bulk drops a ready-to-use B operand stage into shared memory.
// Everything runs on the device, asynchronously on the caller's stream (no host sync: the optimizer
// can step and the next render can repack without draining the GPU).
#include <atomic>

#include <cuda_fp8.h>

#include "common.cuh"

namespace mnrf {

// ---- one launch for all the plain copies of a pack ---------------------------------------------------------------
// pack_field used to issue ~80 tiny launches per field (k_copy / k_transpose_pad / k_copy_block_pad / k_fill), i.e. ~160 per
// training step (both fields are re-packed after every optimizer step).  They are described by a table in the kernel
// parameters and executed by ONE grid-stride kernel: element i of the concatenated destination ranges -> (op, local index).
enum { PK_COPY = 0, PK_TRANSPOSE = 1, PK_BLOCK = 2 };
struct PackOp {
  const float* src;   // nullptr: the destination range is zero-filled
  int dst;            // offset into the fp32 section
  int kind;
  int p0, p1, p2, p3; // COPY: -; TRANSPOSE: N, K (dst is [Kpad][N]); BLOCK: ld, c0, cols, cols_pad
  int end;            // exclusive prefix sum of destination elements up to and including this op
};
constexpr int PK_MAX_OPS = 72;
struct PackOps {
  PackOp op[PK_MAX_OPS];
  int n;
};
__global__ void k_pack_ops(const __grid_constant__ PackOps T, float* __restrict__ f32) {
  const int total = T.op[T.n - 1].end;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int lo = 0, hi = T.n - 1;   // first op whose range ends behind i
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (T.op[mid].end > i) hi = mid; else lo = mid + 1;
    }
    const PackOp& o = T.op[lo];
    const int j = i - (lo > 0 ? T.op[lo - 1].end : 0);
    float v = 0.f;
    if (o.src != nullptr) {
      if (o.kind == PK_COPY) {
        v = o.src[j];
      } else if (o.kind == PK_TRANSPOSE) {      // dst[k][n] = src[n][k], rows k >= K are padding
        const int k = j / o.p0, n = j - k * o.p0;
        v = k < o.p1 ? o.src[n * o.p1 + k] : 0.f;
      } else {                                  // dst[r][c] = c < cols ? src[r*ld + c0 + c] : 0
        const int r = j / o.p3, c = j - r * o.p3;
        v = c < o.p2 ? o.src[(size_t)r * o.p0 + o.p1 + c] : 0.f;
      }
    }
    f32[o.dst + j] = v;
  }
}

// normal_net has no activation between its two Linears (mirror_nerf.py:85-88), so
//   n = W1 (W0 g + b0) + b1 = (W1 W0) g + (W1 b0 + b1):  fold to one 3x256 map (double accumulation).
// headw[c] = {w_sigma[c], Wf[0][c], Wf[1][c], Wf[2][c]},  headb = {b_sigma, bf[0..2]}
__global__ void k_fold_heads(const float* __restrict__ w_sigma, const float* __restrict__ b_sigma,
                             const float* __restrict__ n0w, const float* __restrict__ n0b,
                             const float* __restrict__ n1w, const float* __restrict__ n1b, float4* __restrict__ headw,
                             float* __restrict__ headb) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < W) {
    float4 o;
    o.x = w_sigma[c];
    float v[3] = {0.f, 0.f, 0.f};
    if (n0w != nullptr) {
      for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
        for (int j = 0; j < WH; ++j) acc += (double)n1w[i * WH + j] * (double)n0w[j * W + c];
        v[i] = (float)acc;
      }
    }
    o.y = v[0]; o.z = v[1]; o.w = v[2];
    headw[c] = o;
  }
  if (c < 4) {
    float r;
    if (c == 0) r = b_sigma[0];
    else if (n0w == nullptr) r = 0.f;
    else {
      double acc = (double)n1b[c - 1];
      for (int j = 0; j < WH; ++j) acc += (double)n1w[(c - 1) * WH + j] * (double)n0b[j];
      r = (float)acc;
    }
    headb[c] = r;
  }
}

// ---- tensor-core blobs -------------------------------------------------------------------------
struct TcSrc {
  const float* w[TC_NUM_STEPS];  // source matrices [out][ld] as nn.Linear stores them
  int ld[TC_NUM_STEPS];          // row stride (in-features of the nn.Linear)
};

// value of element (n, k) of step s's B operand, or 0 for padding.  Forward steps: B[n][k] = W[n][col(k)].
// Transposed steps (analytic-normal chain): B[n][k] = W[k][col(n)], n = input feature, k = output feature.
__device__ __forceinline__ float tc_src_value(const TcSrc& src, int s, int n, int k);

// column of the source matrix feeding padded K index k of step s, or -1 for zero padding
__device__ __forceinline__ int tc_src_col(int s, int k) {
  if (s == 0) return k < IN_XYZ ? k : -1;                       // 63 -> 64
  if (s == 4) return k < IN_XYZ ? k : (k == IN_XYZ ? -1 : k - 1);  // [pe(63) pad | h(256)]
  return k;                                                      // 256 (dir layer: feature part only)
}

__device__ __forceinline__ float tc_src_value(const TcSrc& src, int s, int n, int k) {
  const float* w = src.w[s];
  if (w == nullptr) return 0.f;
  if (s < TC_FWD_STEPS) {
    const int c = tc_src_col(s, k);
    return c >= 0 ? w[n * src.ld[s] + c] : 0.f;
  }
  int c;  // input-feature column of the layer's weight matrix
  if (s == 15) c = n < IN_XYZ ? n : -1;          // PE part of layer 5 (63 -> 64)
  else if (s == 14) c = IN_XYZ + n;              // h part of layer 5
  else if (s == 19) c = n < IN_XYZ ? n : -1;     // layer 1 (63 -> 64)
  else c = n;
  return c >= 0 ? w[k * src.ld[s] + c] : 0.f;
}

__global__ void k_absmax(TcSrc src, unsigned int* __restrict__ absmax) {
  int s = blockIdx.y;
  if (src.w[s] == nullptr) return;
  int N = tc_step_n(s), K = tc_step_k(s);
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K; i += gridDim.x * blockDim.x) {
    int n = i / K, k = i % K;
    m = fmaxf(m, fabsf(tc_src_value(src, s, n, k)));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(absmax + s, __float_as_uint(m));
}

// scale = 2^(9 - floor(log2(max)))  ->  max|W| * scale in [2^9, 2^10): fp16 "lo" parts of all but the
// tiniest weights stay in the normal range, and hi never overflows.
__global__ void k_scales(const unsigned int* __restrict__ absmax, float* __restrict__ inv_scale) {
  int s = threadIdx.x;
  if (s >= TC_NUM_STEPS) return;
  float m = __uint_as_float(absmax[s]);
  float inv = 1.f;
  if (m > 0.f && isfinite(m)) {
    int e = ilogbf(m);
    inv = ldexpf(1.f, e - 9);
  }
  inv_scale[s] = inv;
}

__global__ void k_pack_tc(TcSrc src, const float* __restrict__ inv_scale, uint8_t* __restrict__ tc) {
  int s = blockIdx.y;
  int N = tc_step_n(s), K = tc_step_k(s);
  uint8_t* base = tc + tc_step_offset(s);
  const int blob = tc_blob_bytes(s);
  float scale = 1.f / inv_scale[s];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K; i += gridDim.x * blockDim.x) {
    int n = i / K, k = i % K;
    const float v = tc_src_value(src, s, n, k) * scale;
    __half hi = __float2half_rn(v);
    __half lo = __float2half_rn(v - __half2float(hi));
    int kc = k >> 5, kk = k & 31;
    // K-major no-swizzle core matrices: 8 rows x 16 B contiguous; 8-row groups 128 B apart; K-groups N*16 B apart
    int off = (kk >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
    uint8_t* chunk = base + (size_t)kc * 2 * blob;
    *reinterpret_cast<__half*>(chunk + off) = hi;
    *reinterpret_cast<__half*>(chunk + blob + off) = lo;
  }
}

// tc2 blobs (field_tc.cu PREC == 2): same chunk geometry as k_pack_tc, 2^5 more scale (max |W| * scale in [2^14, 2^15): the
// fp16 lo part of typical weights lands in e4m3's normal range), second half of a chunk = [e4m3(2^-10 W_hi) | e4m3(W_lo)] with
// 16 K values per 16-byte core-matrix row.
__global__ void k_pack_tc8(TcSrc src, const float* __restrict__ inv_scale, uint8_t* __restrict__ tc8) {
  int s = blockIdx.y;
  if (s >= TC_FWD_STEPS) return;  // the analytic-normal chain has no fp8 variant
  int N = tc_step_n(s), K = tc_step_k(s);
  uint8_t* base = tc8 + tc_step_offset(s);
  const int blob = tc_blob_bytes(s);
  float scale = 32.f / inv_scale[s];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K; i += gridDim.x * blockDim.x) {
    int n = i / K, k = i % K;
    const float v = tc_src_value(src, s, n, k) * scale;
    const __half hi = __float2half_rn(v);
    const float lo = v - __half2float(hi);
    int kc = k >> 5, kk = k & 31;
    uint8_t* chunk = base + (size_t)kc * 2 * blob;
    *reinterpret_cast<__half*>(chunk + (kk >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2) = hi;
    const int off8 = (kk >> 4) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 15);
    const uint8_t hi8 = (uint8_t)__nv_cvt_float_to_fp8(__half2float(hi) * 0.0009765625f, __NV_SATFINITE, __NV_E4M3);
    const uint8_t lo8 = (uint8_t)__nv_cvt_float_to_fp8(lo, __NV_SATFINITE, __NV_E4M3);
    chunk[blob + off8] = hi8;
    chunk[blob + blob / 2 + off8] = lo8;
    if (s <= TC_SPLIT_LAST) {
      // second copy for the N-split schedule of field_tc.cu: the 256 output rows as two 128-row halves, half-major, each
      // (half, K32 chunk) one 16 KB stage in the N = 128 blob format [fp16 hi 8 KB | e4m3(2^-10 hi) 4 KB | e4m3(lo) 4 KB]
      const int h = n >> 7, n2 = n & 127, nch = K >> 5;
      uint8_t* c2 = tc8 + TC_TOTAL_BYTES + tc_step_offset(s) + (size_t)(h * nch + kc) * 16384;
      *reinterpret_cast<__half*>(c2 + (kk >> 3) * 2048 + (n2 >> 3) * 128 + (n2 & 7) * 16 + (kk & 7) * 2) = hi;
      const int o8 = (kk >> 4) * 2048 + (n2 >> 3) * 128 + (n2 & 7) * 16 + (kk & 15);
      c2[8192 + o8] = hi8;
      c2[8192 + 4096 + o8] = lo8;
    }
  }
}

// ---- tf32 blobs of the training GEMMs (common.cuh T32_*) ------------------------------------------------------------
struct T32Src {
  TcSrc base;
  const float *w_n0, *w_dir, *w_final, *w_m0;
};
__device__ __forceinline__ float t32_value(const T32Src& src, int s, int n, int k) {
  if (s < TC_NUM_STEPS) return tc_src_value(src.base, s, n, k);
  if (s == 20) return src.w_n0 != nullptr ? src.w_n0[n * W + k] : 0.f;             // normal_net.0:      B[n][k] = W[n][k]
  if (s == 21) return src.w_dir[k * (W + IN_DIR) + n];                              // dir layer^T:       B[n=f][k=j] = W[j][f]
  if (s == 22) return src.w_final[k * W + n];                                       // final^T
  if (s == 23) return src.w_n0 != nullptr ? src.w_n0[k * W + n] : 0.f;              // normal_net.0^T
  return src.w_m0 != nullptr ? src.w_m0[k * W + n] : 0.f;                           // is_mirror_net.0^T
}
__global__ void k_pack_t32(T32Src src, uint8_t* __restrict__ t32) {
  const int s = blockIdx.y;
  const int N = t32_step_n(s), K = t32_step_k(s);
  uint8_t* base = t32 + t32_step_offset(s);
  const int blob = N * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i % K;
    const float v = t32_value(src, s, n, k);
    uint32_t hi, lo;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    const float r = v - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
    const int kc = k >> 4, kk = k & 15;
    // K-major no-swizzle core matrices of 8 rows x 4 tf32: 8-row groups 128 B apart, K-groups N*16 B apart
    const int off = (kk >> 2) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 3) * 4;
    uint8_t* chunk = base + (size_t)kc * 2 * blob;
    *reinterpret_cast<uint32_t*>(chunk + off) = hi;
    *reinterpret_cast<uint32_t*>(chunk + blob + off) = lo;
  }
}

int pack_field(mnrf_field* f, const float* const* t, cudaStream_t st) {
  static std::atomic<unsigned long long> g_pack_stamp{0};
  f->pack_stamp = ++g_pack_stamp;  // invalidates the constant-memory copies of the epilogue table (field_tc.cu)
  const F32Layout& L = f->L;
  float* d = f->f32;
  PackOps ops;
  ops.n = 0;
  int total = 0;
  auto add = [&](const float* src, int dst, int kind, int n, int p0, int p1, int p2, int p3) {
    PackOp& o = ops.op[ops.n++];
    o.src = src; o.dst = dst; o.kind = kind; o.p0 = p0; o.p1 = p1; o.p2 = p2; o.p3 = p3;
    total += n;
    o.end = total;
  };
  auto copy_to = [&](const float* src, int dst, int n) { add(src, dst, PK_COPY, n, 0, 0, 0, 0); };
  auto transpose_to = [&](const float* src, int dst, int N, int K) { add(src, dst, PK_TRANSPOSE, pad4(K) * N, N, K, 0, 0); };
  auto block_to = [&](const float* src, int dst, int rows, int ld, int c0, int cols, int cols_pad) {
    add(src, dst, PK_BLOCK, rows * cols_pad, ld, c0, cols, cols_pad);
  };
  for (int l = 0; l < 8; ++l) {
    transpose_to(t[2 * l], L.wt_trunk[l], W, trunk_k(l));
    copy_to(t[2 * l], L.w_trunk[l], W * trunk_k(l));
    copy_to(t[2 * l + 1], L.b_trunk[l], W);
  }
  transpose_to(t[T_FINAL_W], L.wt_final, W, W);
  copy_to(t[T_FINAL_B], L.b_final, W);
  transpose_to(t[T_DIR_W], L.wt_dir, WH, W + IN_DIR);
  copy_to(t[T_DIR_B], L.b_dir, WH);
  copy_to(t[T_SIGMA_W], L.w_sigma, W);
  copy_to(t[T_SIGMA_B], L.b_sigma, 1);
  copy_to(t[T_RGB_W], L.w_rgb, 3 * WH);
  copy_to(t[T_RGB_B], L.b_rgb, 3);
  transpose_to(t[T_N0_W], L.wt_n0, WH, W);
  copy_to(t[T_N0_B], L.b_n0, WH);
  copy_to(t[T_N1_W], L.w_n1, 3 * WH);
  copy_to(t[T_N1_B], L.b_n1, 3);
  transpose_to(t[T_M0_W], L.wt_m0, WH, W);
  copy_to(t[T_M0_B], L.b_m0, WH);
  copy_to(t[T_M2_W], L.w_m2, WH);
  copy_to(t[T_M2_B], L.b_m2, 1);
  // aligned [out][in] copies for the training path (train.cu)
  block_to(t[0], L.tw_l1, W, IN_XYZ, 0, IN_XYZ, PE_PAD);
  block_to(t[8], L.tw_l5a, W, IN_XYZ + W, 0, IN_XYZ, PE_PAD);
  block_to(t[8], L.tw_l5b, W, IN_XYZ + W, IN_XYZ, W, W);
  block_to(t[T_FINAL_W], L.tw_final, W, W, 0, W, W);
  block_to(t[T_DIR_W], L.tw_dira, WH, W + IN_DIR, 0, W, W);
  block_to(t[T_N0_W], L.tw_n0, WH, W, 0, W, W);
  block_to(t[T_M0_W], L.tw_m0, WH, W, 0, W, W);
  MNRF_REQUIRE(ops.n <= PK_MAX_OPS, "pack_field: op table overflow (%d)", ops.n);
  k_pack_ops<<<(total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184, 256, 0, st>>>(ops, d);
  MNRF_LAUNCH_OK();

  k_fold_heads<<<1, 256, 0, st>>>(t[T_SIGMA_W], t[T_SIGMA_B], t[T_N0_W], t[T_N0_B], t[T_N1_W], t[T_N1_B],
                                  reinterpret_cast<float4*>(d + L.headw), d + L.headb);
  MNRF_LAUNCH_OK();

  TcSrc src;
  for (int s = 0; s < 8; ++s) { src.w[s] = t[2 * s]; src.ld[s] = trunk_k(s); }
  src.w[8] = t[T_FINAL_W]; src.ld[8] = W;
  src.w[9] = t[T_M0_W];    src.ld[9] = W;
  src.w[10] = t[T_DIR_W];  src.ld[10] = W + IN_DIR;
  for (int s = TC_FWD_STEPS; s < TC_NUM_STEPS; ++s) { const int l = tc_step_layer(s); src.w[s] = t[2 * l]; src.ld[s] = trunk_k(l); }
  unsigned int* absmax = reinterpret_cast<unsigned int*>(d + L.absmax);
  MNRF_CUDA_OK(cudaMemsetAsync(absmax, 0, 32 * sizeof(unsigned int), st));
  k_absmax<<<dim3(32, TC_NUM_STEPS), 256, 0, st>>>(src, absmax);
  MNRF_LAUNCH_OK();
  k_scales<<<1, 32, 0, st>>>(absmax, d + L.inv_scale);
  MNRF_LAUNCH_OK();
  k_pack_tc<<<dim3(64, TC_NUM_STEPS), 256, 0, st>>>(src, d + L.inv_scale, f->tc);
  MNRF_LAUNCH_OK();
  k_pack_tc8<<<dim3(64, TC_FWD_STEPS), 256, 0, st>>>(src, d + L.inv_scale, f->tc8);
  MNRF_LAUNCH_OK();
  T32Src s32;
  s32.base = src;
  s32.w_n0 = t[T_N0_W]; s32.w_dir = t[T_DIR_W]; s32.w_final = t[T_FINAL_W]; s32.w_m0 = t[T_M0_W];
  k_pack_t32<<<dim3(64, T32_NUM_STEPS), 256, 0, st>>>(s32, f->t32);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace mnrf
