// Backward of the hash-grid field (MirrorNeRFTcnn, R/models/mirror_nerf_tcnn.py:151-259) for one warp = 32 points: everything
// torch.autograd derives for that module when R/train.py:129-145 trains the nerf_tcnn model family -- gradients of the hash table
// (scatter-add through the trilinear weights), of the four small MLPs, of the rays, and the DOUBLE backward through the analytic
// normal n = normalize(-d sigma/d xyz) (mirror_nerf_tcnn.py:170-178 with create_graph=True in R/utils/func.py:10-25).
//
// The forward is recomputed here (it costs ~11 k MAC and 128 table reads per point; saving its activations would cost 2 KB per
// point of HBM traffic instead).  The code is written as a sequence of PHASES.  A "lane phase" touches only the lane's own column
// of the per-warp buffers (thread = point); a "warp phase" forms weight gradients, where lanes own weight elements and read all
// 32 columns.  Phases are separated by a warp barrier.  The same source compiles as plain C++ (tests/emu/hash_train_emu.cpp runs
// the phases lane by lane on the CPU), which is how this math is checked against the oracle's autograd without a GPU.
//
// Per-warp buffers are feature-major [row][HT_LD] floats with HT_LD = 33 so that both access patterns (lane = point column,
// lane = feature row) are bank-conflict free.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define HT_DEV __device__ __forceinline__
#define HT_HD __host__ __device__ __forceinline__
#define HT_MUL(a, b) __fmul_rn((a), (b))
#define HT_ADD(a, b) __fadd_rn((a), (b))
#define HT_SUB(a, b) __fsub_rn((a), (b))
#define HT_DIV(a, b) __fdiv_rn((a), (b))
#define HT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define HT_LDG2(p) __ldg(p)
#else
#define HT_DEV inline
#define HT_HD inline
#define HT_MUL(a, b) ((a) * (b))
#define HT_ADD(a, b) ((a) + (b))
#define HT_SUB(a, b) ((a) - (b))
#define HT_DIV(a, b) ((a) / (b))
#define HT_ATOMIC_ADD(p, v) (*(p) += (v))
#define HT_LDG2(p) (*(p))
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
#endif

namespace mnrf {
namespace ht {

constexpr int HT_LEVELS = 16;
constexpr int HT_LD = 33;
constexpr float HT_EPS = 1.1920928955078125e-07f;
constexpr int HT_DR_STRIDE = 12;  // per-point gradient record of k_train_composite_bwd (train.cu)
constexpr int HT_DXD_STRIDE = 8;  // per-point ray-gradient record: [dx(3), dSH/dd^T g (3), 0, 0]

// rows of the per-warp buffer (260 rows = 34 KB: four warps per CTA next to the two 45 KB weight / gradient images)
constexpr int R_E = 0;     // 32: encoding; later g_e = d sigma / d enc
constexpr int R_H = 32;    // 64: hidden layer of sigma_net (post ReLU); later the masked tangent
constexpr int R_P = 96;    // 64: scratch
constexpr int R_Q = 160;   // 64: scratch
constexpr int R_R = 224;   // 32: colour-net input [SH4 (16) | sigma | geo_feat (15)]; later d enc
constexpr int R_S = 240;   // 16: [sigma, geo_feat] = rows 16..31 of R (the colour net's weight column 16 is zero); later their gradients
constexpr int R_D = 256;   // 4:  small head gradients
constexpr int HT_ROWS = 260;
constexpr int HT_WARP_FLOATS = HT_ROWS * HT_LD;

// The small weights ([out][in] row-major, R/models/mirror_nerf_tcnn.py:52-149) concatenated, every row padded to a multiple of
// 4 inputs (31 -> 32, 15 -> 16; pad columns are zero) so that rows are 16-byte aligned for vector loads.  The field keeps this
// image on the device (mnrf_field::hash_wref); the kernel's gradient accumulator has the same layout.
constexpr int O_S0 = 0;               // sigma_net.0.weight   64 x 32
constexpr int O_S1 = O_S0 + 64 * 32;  // sigma_net.1.weight   16 x 64
constexpr int O_C0 = O_S1 + 16 * 64;  // color_net.0.weight   64 x 31 (ld 32; the zero pad column sits at index 16, see R_S)
constexpr int O_C1 = O_C0 + 64 * 32;  // color_net.1.weight   64 x 64
constexpr int O_C2 = O_C1 + 64 * 64;  // color_net.2.weight    3 x 64
constexpr int O_N0 = O_C2 + 3 * 64;   // normal_net.0.weight  64 x 15 (ld 16)
constexpr int O_N1 = O_N0 + 64 * 16;  // normal_net.1.weight   3 x 64
constexpr int O_M0 = O_N1 + 3 * 64;   // is_mirror_net.0.weight 32 x 15 (ld 16)
constexpr int O_M0B = O_M0 + 32 * 16; // is_mirror_net.0.bias  32
constexpr int O_M2 = O_M0B + 32;      // is_mirror_net.2.weight 1 x 32
constexpr int O_M2B = O_M2 + 32;      // is_mirror_net.2.bias  1
constexpr int HT_NW = O_M2B + 4;      // 11204
constexpr int LD_C0 = 32, LD_N0 = 16, LD_M0 = 16;
// tensor index (mnrf_hash_field_create order, 1..11) -> offset / rows / columns / padded row stride
HT_HD int small_offset(int i) {
  switch (i) {
    case 1: return O_S0; case 2: return O_S1; case 3: return O_C0; case 4: return O_C1; case 5: return O_C2; case 6: return O_N0;
    case 7: return O_N1; case 8: return O_M0; case 9: return O_M0B; case 10: return O_M2; default: return O_M2B;
  }
}
HT_HD int small_rows(int i) {
  switch (i) {
    case 1: return 64; case 2: return 16; case 3: return 64; case 4: return 64; case 5: return 3; case 6: return 64;
    case 7: return 3; case 8: return 32; default: return 1;
  }
}
HT_HD int small_cols(int i) {
  switch (i) {
    case 1: return 32; case 2: return 64; case 3: return 31; case 4: return 64; case 5: return 64; case 6: return 15;
    case 7: return 64; case 8: return 15; case 9: return 32; case 10: return 32; default: return 1;
  }
}
HT_HD int small_ld(int i) { return (small_cols(i) + 3) / 4 * 4; }
// column of the padded image that holds reference column c of tensor i
HT_HD int small_col(int i, int c) { return (i == 3 && c >= 16) ? c + 1 : c; }

struct Flags {
  int has_normal, has_mirror;
  int compute_normal;  // analytic normals were produced by the forward (their gradient may be non-zero)
  int detach_normal;   // detach_density_for_normal_loss (mirror_nerf_tcnn.py:188-190)
  int detach_mask;     // detach_density_for_mask_loss   (mirror_nerf_tcnn.py:201-202)
  int ray_grad;        // also d L / d [o, d]
};

// J[3*k + c] = d enc_k / d u_c is a separate per-lane array of HT_J floats (indexed dynamically, so it lives in local memory; only
// its own lane reads it, which is why it is not in the shared buffer).
constexpr int HT_J = 96;
// per-lane state that lives across phases (registers on the GPU)
struct Lane {
  float u[3];      // position in the unit cube
  float d[3];      // ray direction (input of the SH encoding)
  float dr[HT_DR_STRIDE];  // [0] d sigma [1..3] d rgb [4] d is_mirror [5..7] d pred_normal [8..10] d analytic normal
  float dgeo[15];
  float dco[3];    // d (colour pre-sigmoid)
  float dnraw[3];  // d (normal_net output before l2-normalise)
  float dmp;       // d (mirror pre-sigmoid)
  float t[3];      // d L / d g_u  (g_u = d sigma / d u)
  float du[3];     // d L / d u
  float dsh[16];   // d L / d SH(d)
  int mirror_on;   // the mirror-mask loss reaches the density for this ray
  int valid;
};

HT_DEV float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

HT_DEV void normalize_bwd(const float (&v)[3], const float (&dy)[3], float (&dv)[3]) {  // y = v / sqrt(max(|v|^2, eps))
  const float nn = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (nn > HT_EPS) {
    const float inv = 1.f / sqrtf(nn);
    const float y0 = v[0] * inv, y1 = v[1] * inv, y2 = v[2] * inv;
    const float dot = y0 * dy[0] + y1 * dy[1] + y2 * dy[2];
    dv[0] = (dy[0] - y0 * dot) * inv; dv[1] = (dy[1] - y1 * dot) * inv; dv[2] = (dy[2] - y2 * dot) * inv;
  } else {
    const float inv = 1.f / sqrtf(HT_EPS);
    dv[0] = dy[0] * inv; dv[1] = dy[1] * inv; dv[2] = dy[2] * inv;
  }
}

// out[j] = sum_{k<K} W[(o0+j)*ldw + k] * X[k][lane]           (thread = point; NB independent accumulators)
// ldw % 4 == 0 and W 16-byte aligned: the weights (warp-uniform addresses) are read four inputs at a time.
template <int NB, int LD = HT_LD>
HT_DEV void rows_dot(const float* W, int ldw, int o0, const float* X, int K, int lane, float (&out)[NB]) {
#pragma unroll
  for (int j = 0; j < NB; ++j) out[j] = 0.f;
  const int K4 = K & ~3;
  for (int k = 0; k < K4; k += 4) {
    const float x0 = X[k * LD + lane], x1 = X[(k + 1) * LD + lane], x2 = X[(k + 2) * LD + lane], x3 = X[(k + 3) * LD + lane];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const float4 w = *reinterpret_cast<const float4*>(W + (o0 + j) * ldw + k);
      out[j] = fmaf(w.x, x0, out[j]); out[j] = fmaf(w.y, x1, out[j]); out[j] = fmaf(w.z, x2, out[j]); out[j] = fmaf(w.w, x3, out[j]);
    }
  }
  for (int k = K4; k < K; ++k) {
    const float x = X[k * LD + lane];
#pragma unroll
    for (int j = 0; j < NB; ++j) out[j] = fmaf(W[(o0 + j) * ldw + k], x, out[j]);
  }
}
// out[j] = sum_{o<O} W[o*ldw + k0 + j] * A[o][lane], j < 8    (transposed weights; k0 % 4 == 0, ldw % 4 == 0)
template <int LD = HT_LD>
HT_DEV void cols_dot8(const float* W, int ldw, int k0, const float* A, int O, int lane, float (&out)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) out[j] = 0.f;
  for (int o = 0; o < O; ++o) {
    const float a = A[o * LD + lane];
    const float4 w0 = *reinterpret_cast<const float4*>(W + o * ldw + k0);
    const float4 w1 = *reinterpret_cast<const float4*>(W + o * ldw + k0 + 4);
    out[0] = fmaf(w0.x, a, out[0]); out[1] = fmaf(w0.y, a, out[1]); out[2] = fmaf(w0.z, a, out[2]); out[3] = fmaf(w0.w, a, out[3]);
    out[4] = fmaf(w1.x, a, out[4]); out[5] = fmaf(w1.y, a, out[5]); out[6] = fmaf(w1.z, a, out[6]); out[7] = fmaf(w1.w, a, out[7]);
  }
}
// warp phase: G[o*ldg + k] += sum_p A[o][p] * B[k][p]   for o < O, k < K.  Every lane owns register tiles of 4 x 4 elements
// (o = ob + nb_o*i, k = kb + nb_k*j: strided, so that the lanes of a warp read consecutive rows = distinct banks) and walks the 32
// points with 8 shared loads per 16 FMAs.
template <int LD = HT_LD, int NP = 32>
HT_DEV void wgrad(float* G, int ldg, const float* A, int O, const float* B, int K, int lane) {
  const int nb_o = (O + 3) >> 2, nb_k = (K + 3) >> 2;
  const int nblk = nb_o * nb_k;
  for (int blk = lane; blk < nblk; blk += 32) {
    const int ob = blk / nb_k, kb = blk - ob * nb_k;
    const float* a[4];
    const float* b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = ob + nb_o * i, k = kb + nb_k * i;
      a[i] = A + (o < O ? o : O - 1) * LD;  // clamped rows: their products are discarded below
      b[i] = B + (k < K ? k : K - 1) * LD;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int p = 0; p < NP; ++p) {
      const float a0 = a[0][p], a1 = a[1][p], a2 = a[2][p], a3 = a[3][p];
      const float b0 = b[0][p], b1 = b[1][p], b2 = b[2][p], b3 = b[3][p];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]); acc[0][2] = fmaf(a0, b2, acc[0][2]); acc[0][3] = fmaf(a0, b3, acc[0][3]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]); acc[1][2] = fmaf(a1, b2, acc[1][2]); acc[1][3] = fmaf(a1, b3, acc[1][3]);
      acc[2][0] = fmaf(a2, b0, acc[2][0]); acc[2][1] = fmaf(a2, b1, acc[2][1]); acc[2][2] = fmaf(a2, b2, acc[2][2]); acc[2][3] = fmaf(a2, b3, acc[2][3]);
      acc[3][0] = fmaf(a3, b0, acc[3][0]); acc[3][1] = fmaf(a3, b1, acc[3][1]); acc[3][2] = fmaf(a3, b2, acc[3][2]); acc[3][3] = fmaf(a3, b3, acc[3][3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = ob + nb_o * i;
      if (o >= O) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kb + nb_k * j;
        if (k < K) HT_ATOMIC_ADD(G + o * ldg + k, acc[i][j]);
      }
    }
  }
}
// warp phase: G[o] += sum_p A[o][p]
template <int LD = HT_LD, int NP = 32>
HT_DEV void wcolsum(float* G, const float* A, int O, int lane) {
  for (int o = lane; o < O; o += 32) {
    float acc = 0.f;
    for (int p = 0; p < NP; ++p) acc += A[o * LD + p];
    HT_ATOMIC_ADD(G + o, acc);
  }
}

template <class MetaT>
HT_DEV void level_cell(const MetaT& M, int l, const float (&u)[3], unsigned int (&g)[3], float (&fr)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float pos = HT_ADD(HT_MUL(u[c], M.scale[l]), 0.5f);
    const float fl = floorf(pos);
    g[c] = (unsigned int)(int)fl;
    fr[c] = HT_SUB(pos, fl);
  }
}
HT_DEV unsigned int grid_index(const unsigned int (&c3)[3], unsigned int res, unsigned int size) {
  unsigned int stride = 1, index = 0;
  int dim = 0;
  for (; dim < 3 && stride <= size; ++dim) { index += c3[dim] * stride; stride *= res; }
  if (size < stride) index = (c3[0] * 1u) ^ (c3[1] * 2654435761u) ^ (c3[2] * 805459861u);
  return (size & (size - 1u)) == 0u ? (index & (size - 1u)) : (index % size);  // the hashed levels hold 2^19 entries
}

// ---- lane phase A: encoding + its Jacobian, sigma_net, colour net forward ---------------------------------------------------
template <class MetaT>
HT_DEV void phase_a(const float* Wt, float* B, const float* table, const MetaT& M, const Flags& F, Lane& L, float* J, int lane) {
  const float2* tab = reinterpret_cast<const float2*>(table);
  float* E = B + R_E * HT_LD;
#pragma unroll 2
  for (int l = 0; l < HT_LEVELS; ++l) {  // two levels = 16 independent table reads in flight per lane
    const float scale = M.scale[l];
    const unsigned int res = (unsigned int)M.res[l], size = M.size[l];
    unsigned int g[3];
    float fr[3];
    level_cell(M, l, L.u, g, fr);
    float a0 = 0.f, a1 = 0.f, j0[3] = {0.f, 0.f, 0.f}, j1[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
      float wd[3], w = 1.f;
      unsigned int c3[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int bit = (corner >> c) & 1;
        wd[c] = bit ? fr[c] : HT_SUB(1.f, fr[c]);
        w = HT_MUL(w, wd[c]);
        c3[c] = g[c] + bit;
      }
      const float2 f = HT_LDG2(tab + M.offset[l] + grid_index(c3, res, size));
      a0 = HT_ADD(a0, HT_MUL(w, f.x));
      a1 = HT_ADD(a1, HT_MUL(w, f.y));
      const float d0 = ((corner & 1) ? 1.f : -1.f) * wd[1] * wd[2];
      const float d1 = ((corner & 2) ? 1.f : -1.f) * wd[0] * wd[2];
      const float d2 = ((corner & 4) ? 1.f : -1.f) * wd[0] * wd[1];
      j0[0] = fmaf(d0, f.x, j0[0]); j0[1] = fmaf(d1, f.x, j0[1]); j0[2] = fmaf(d2, f.x, j0[2]);
      j1[0] = fmaf(d0, f.y, j1[0]); j1[1] = fmaf(d1, f.y, j1[1]); j1[2] = fmaf(d2, f.y, j1[2]);
    }
    E[(2 * l) * HT_LD + lane] = a0;
    E[(2 * l + 1) * HT_LD + lane] = a1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      J[3 * (2 * l) + c] = scale * j0[c];
      J[3 * (2 * l + 1) + c] = scale * j1[c];
    }
  }
  // sigma_net: 32 -> 64 (ReLU) -> 16
  float* H = B + R_H * HT_LD;
  float* S = B + R_S * HT_LD;
  for (int o0 = 0; o0 < 64; o0 += 8) {
    float acc[8];
    rows_dot<8>(Wt + O_S0, 32, o0, E, 32, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) H[(o0 + j) * HT_LD + lane] = fmaxf(acc[j], 0.f);
  }
  for (int o0 = 0; o0 < 16; o0 += 8) {
    float acc[8];
    rows_dot<8>(Wt + O_S1, 64, o0, H, 64, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) S[(o0 + j) * HT_LD + lane] = acc[j];
  }
#pragma unroll
  for (int j = 0; j < 15; ++j) L.dgeo[j] = 0.f;
  // colour net: [SH4(d) | geo_feat] (31) -> 64 -> 64 -> 3, sigmoid
  float* Rb = B + R_R * HT_LD;
  float* P = B + R_P * HT_LD;
  float* Q = B + R_Q * HT_LD;
  {
    const float X = L.d[0], Y = L.d[1], Z = L.d[2];
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
    for (int i = 0; i < 16; ++i) Rb[i * HT_LD + lane] = sh[i];  // rows 16..31 = [sigma | geo_feat] are already in place (R_S)
  }
  for (int o0 = 0; o0 < 64; o0 += 8) {
    float acc[8];
    rows_dot<8>(Wt + O_C0, LD_C0, o0, Rb, 32, lane, acc);  // input row 16 is sigma: its weight column is zero
#pragma unroll
    for (int j = 0; j < 8; ++j) P[(o0 + j) * HT_LD + lane] = fmaxf(acc[j], 0.f);
  }
  for (int o0 = 0; o0 < 64; o0 += 8) {
    float acc[8];
    rows_dot<8>(Wt + O_C1, 64, o0, P, 64, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) Q[(o0 + j) * HT_LD + lane] = fmaxf(acc[j], 0.f);
  }
  {
    float acc[3];
    rows_dot<3>(Wt + O_C2, 64, 0, Q, 64, lane, acc);
    float* D = B + R_D * HT_LD;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float rgb = sigmoidf_(acc[c]);
      L.dco[c] = L.dr[1 + c] * rgb * (1.f - rgb);
      D[c * HT_LD + lane] = L.dco[c];
    }
  }
  (void)F;
}
// warp phase B: d color_net.2
HT_DEV void phase_b(float* G, const float* B, int lane) { wgrad(G + O_C2, 64, B + R_D * HT_LD, 3, B + R_Q * HT_LD, 64, lane); }
// lane phase C: d c2 (in place over c2)
HT_DEV void phase_c(const float* Wt, float* B, const Lane& L, int lane) {
  float* Q = B + R_Q * HT_LD;
  for (int o = 0; o < 64; ++o) {
    const float g = Wt[O_C2 + o] * L.dco[0] + Wt[O_C2 + 64 + o] * L.dco[1] + Wt[O_C2 + 128 + o] * L.dco[2];
    Q[o * HT_LD + lane] = Q[o * HT_LD + lane] > 0.f ? g : 0.f;
  }
}
// warp phase D: d color_net.1
HT_DEV void phase_d(float* G, const float* B, int lane) { wgrad(G + O_C1, 64, B + R_Q * HT_LD, 64, B + R_P * HT_LD, 64, lane); }
// lane phase E: d c1 (in place over c1)
HT_DEV void phase_e(const float* Wt, float* B, int lane) {
  float* P = B + R_P * HT_LD;
  const float* Q = B + R_Q * HT_LD;
  for (int k0 = 0; k0 < 64; k0 += 8) {
    float acc[8];
    cols_dot8(Wt + O_C1, 64, k0, Q, 64, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) P[(k0 + j) * HT_LD + lane] = P[(k0 + j) * HT_LD + lane] > 0.f ? acc[j] : 0.f;
  }
}
// warp phase F: d color_net.0
HT_DEV void phase_f(float* G, const float* B, int lane) { wgrad(G + O_C0, LD_C0, B + R_P * HT_LD, 64, B + R_R * HT_LD, 32, lane); }  // column 16 is never read back
// lane phase G: d [SH | geo] of the colour net; forward of the normal and mirror heads up to their output gradients
HT_DEV void phase_g(const float* Wt, float* B, const Flags& F, Lane& L, int lane) {
  float* P = B + R_P * HT_LD;
  float* Q = B + R_Q * HT_LD;
  float* Rb = P;  // the mirror head's hidden layer goes to rows 0..31 of P (d c1 has been consumed by then)
  const float* S = B + R_S * HT_LD;
  float* D = B + R_D * HT_LD;
  {
    float acc[8];
    cols_dot8(Wt + O_C0, LD_C0, 16, P, 64, lane, acc);  // column 16 is padding
#pragma unroll
    for (int j = 1; j < 8; ++j) L.dgeo[j - 1] += acc[j];
    cols_dot8(Wt + O_C0, LD_C0, 24, P, 64, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dgeo[7 + j] += acc[j];
  }
  if (F.ray_grad) {
    float acc[8];
    cols_dot8(Wt + O_C0, LD_C0, 0, P, 64, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dsh[j] = acc[j];
    cols_dot8(Wt + O_C0, LD_C0, 8, P, 64, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dsh[8 + j] = acc[j];
  }
  const float* GEO = S + HT_LD;  // rows 1..15
  if (F.has_normal) {  // 15 -> 64 (ReLU) -> 3, l2-normalised
    for (int o0 = 0; o0 < 64; o0 += 8) {
      float acc[8];
      rows_dot<8>(Wt + O_N0, LD_N0, o0, GEO, 15, lane, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) Q[(o0 + j) * HT_LD + lane] = fmaxf(acc[j], 0.f);
    }
    float nraw[3];
    rows_dot<3>(Wt + O_N1, 64, 0, Q, 64, lane, nraw);
    const float dy[3] = {L.dr[5], L.dr[6], L.dr[7]};
    normalize_bwd(nraw, dy, L.dnraw);
#pragma unroll
    for (int c = 0; c < 3; ++c) D[c * HT_LD + lane] = L.dnraw[c];
  }
  if (F.has_mirror) {  // 15 -> 32 (+bias, LeakyReLU 0.01) -> 1 (+bias), sigmoid
    for (int o0 = 0; o0 < 32; o0 += 8) {
      float acc[8];
      rows_dot<8>(Wt + O_M0, LD_M0, o0, GEO, 15, lane, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = acc[j] + Wt[O_M0B + o0 + j];
        Rb[(o0 + j) * HT_LD + lane] = v > 0.f ? v : 0.01f * v;
      }
    }
    float mp[1];
    rows_dot<1>(Wt + O_M2, 32, 0, Rb, 32, lane, mp);
    const float m = sigmoidf_(mp[0] + Wt[O_M2B]);
    L.dmp = L.dr[4] * m * (1.f - m);
    D[3 * HT_LD + lane] = L.dmp;
  }
}
// warp phase H: d normal_net.1, d is_mirror_net.2 (+ bias)
HT_DEV void phase_h(float* G, const float* B, const Flags& F, int lane) {
  if (F.has_normal) wgrad(G + O_N1, 64, B + R_D * HT_LD, 3, B + R_Q * HT_LD, 64, lane);
  if (F.has_mirror) {
    wgrad(G + O_M2, 32, B + (R_D + 3) * HT_LD, 1, B + R_P * HT_LD, 32, lane);
    wcolsum(G + O_M2B, B + (R_D + 3) * HT_LD, 1, lane);
  }
}
// lane phase I: gradients of the heads' hidden layers (in place)
HT_DEV void phase_i(const float* Wt, float* B, const Flags& F, const Lane& L, int lane) {
  if (F.has_normal) {
    float* Q = B + R_Q * HT_LD;
    for (int o = 0; o < 64; ++o) {
      const float g = Wt[O_N1 + o] * L.dnraw[0] + Wt[O_N1 + 64 + o] * L.dnraw[1] + Wt[O_N1 + 128 + o] * L.dnraw[2];
      Q[o * HT_LD + lane] = Q[o * HT_LD + lane] > 0.f ? g : 0.f;
    }
  }
  if (F.has_mirror) {
    float* Rb = B + R_P * HT_LD;
    for (int o = 0; o < 32; ++o) {
      const float a = Rb[o * HT_LD + lane];  // leaky(pre): same sign as pre
      Rb[o * HT_LD + lane] = Wt[O_M2 + o] * L.dmp * (a > 0.f ? 1.f : 0.01f);
    }
  }
}
// warp phase J: d normal_net.0, d is_mirror_net.0 (+ bias)
HT_DEV void phase_j(float* G, const float* B, const Flags& F, int lane) {
  const float* GEO = B + (R_S + 1) * HT_LD;
  if (F.has_normal) wgrad(G + O_N0, LD_N0, B + R_Q * HT_LD, 64, GEO, 15, lane);
  if (F.has_mirror) {
    wgrad(G + O_M0, LD_M0, B + R_P * HT_LD, 32, GEO, 15, lane);
    wcolsum(G + O_M0B, B + R_P * HT_LD, 32, lane);
  }
}
// lane phase K: d geo_feat from the heads, d [sigma, geo] -> S rows, d hidden -> P
HT_DEV void phase_k(const float* Wt, float* B, const Flags& F, Lane& L, int lane) {
  if (F.has_normal && !F.detach_normal) {
    float acc[8];
    cols_dot8(Wt + O_N0, LD_N0, 0, B + R_Q * HT_LD, 64, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dgeo[j] += acc[j];
    cols_dot8(Wt + O_N0, LD_N0, 8, B + R_Q * HT_LD, 64, lane, acc);  // column 15 is padding
#pragma unroll
    for (int j = 0; j < 7; ++j) L.dgeo[8 + j] += acc[j];
  }
  if (F.has_mirror && L.mirror_on) {
    float acc[8];
    cols_dot8(Wt + O_M0, LD_M0, 0, B + R_P * HT_LD, 32, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dgeo[j] += acc[j];
    cols_dot8(Wt + O_M0, LD_M0, 8, B + R_P * HT_LD, 32, lane, acc);
#pragma unroll
    for (int j = 0; j < 7; ++j) L.dgeo[8 + j] += acc[j];
  }
  float* S = B + R_S * HT_LD;
  S[lane] = L.dr[0];
#pragma unroll
  for (int j = 0; j < 15; ++j) S[(1 + j) * HT_LD + lane] = L.dgeo[j];
  float* P = B + R_P * HT_LD;
  const float* H = B + R_H * HT_LD;
  for (int k0 = 0; k0 < 64; k0 += 8) {
    float acc[8];
    cols_dot8(Wt + O_S1, 64, k0, S, 16, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) P[(k0 + j) * HT_LD + lane] = H[(k0 + j) * HT_LD + lane] > 0.f ? acc[j] : 0.f;
  }
}
// warp phase L: d sigma_net.1, d sigma_net.0 (first-order parts)
HT_DEV void phase_l(float* G, const float* B, int lane) {
  wgrad(G + O_S1, 64, B + R_S * HT_LD, 16, B + R_H * HT_LD, 64, lane);
  wgrad(G + O_S0, 32, B + R_P * HT_LD, 64, B + R_E * HT_LD, 32, lane);
}
// lane phase M: d enc; the double backward through the analytic normal; table scatter; ray-gradient record
template <class MetaT>
HT_DEV void phase_m(const float* Wt, float* B, const float* table, float* gtable, const MetaT& M, const Flags& F, Lane& L,
                    const float* J, int lane, bool second_order) {
  float* E = B + R_E * HT_LD;
  float* H = B + R_H * HT_LD;
  float* P = B + R_P * HT_LD;
  float* Q = B + R_Q * HT_LD;
  float* Rb = B + R_R * HT_LD;
  // d enc = W0^T d hidden -> R rows
  for (int k0 = 0; k0 < 32; k0 += 8) {
    float acc[8];
    cols_dot8(Wt + O_S0, 32, k0, P, 64, lane, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) Rb[(k0 + j) * HT_LD + lane] = acc[j];
  }
  L.t[0] = L.t[1] = L.t[2] = 0.f;
  const float inv2b = 1.f / (2.f * M.bound);
  if (second_order) {
    // q = W1[0,:] * relu'(h) -> Q;  g_e = W0^T q -> E (the encoding itself is no longer needed)
    for (int o = 0; o < 64; ++o) Q[o * HT_LD + lane] = H[o * HT_LD + lane] > 0.f ? Wt[O_S1 + o] : 0.f;
    for (int k0 = 0; k0 < 32; k0 += 8) {
      float acc[8];
      cols_dot8(Wt + O_S0, 32, k0, Q, 64, lane, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) E[(k0 + j) * HT_LD + lane] = acc[j];
    }
    // g_u = J^T g_e;  n = normalize(-g_u / (2 bound));  t = d L / d g_u
    float gu[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < 32; ++k) {
      const float ge = E[k * HT_LD + lane];
#pragma unroll
      for (int c = 0; c < 3; ++c) gu[c] = fmaf(ge, J[3 * k + c], gu[c]);
    }
    const float v[3] = {-gu[0] * inv2b, -gu[1] * inv2b, -gu[2] * inv2b};
    const float dy[3] = {L.dr[8], L.dr[9], L.dr[10]};
    float dv[3];
    normalize_bwd(v, dy, dv);
#pragma unroll
    for (int c = 0; c < 3; ++c) L.t[c] = -dv[c] * inv2b;
    // r = J t -> P rows 0..31 (d L / d g_e);  masked tangent (W0 r) * relu'(h) -> H (in place: d L / d W1[0,:] summand)
    for (int k = 0; k < 32; ++k)
      P[k * HT_LD + lane] = L.t[0] * J[3 * k] + L.t[1] * J[3 * k + 1] + L.t[2] * J[3 * k + 2];
    for (int o0 = 0; o0 < 64; o0 += 8) {
      float acc[8];
      rows_dot<8>(Wt + O_S0, 32, o0, P, 32, lane, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) H[(o0 + j) * HT_LD + lane] = H[(o0 + j) * HT_LD + lane] > 0.f ? acc[j] : 0.f;
    }
  }
  // d L / d u, first-order part
  L.du[0] = L.du[1] = L.du[2] = 0.f;
  if (F.ray_grad) {
    for (int k = 0; k < 32; ++k) {
      const float de = Rb[k * HT_LD + lane];
#pragma unroll
      for (int c = 0; c < 3; ++c) L.du[c] = fmaf(de, J[3 * k + c], L.du[c]);
    }
  }
  // table scatter: d T[idx] += w * d enc + scale * g_e * (t . d w / d frac)   [+ second-order ray term]
  if (!L.valid) return;
  const float2* tab = reinterpret_cast<const float2*>(table);
  for (int l = 0; l < HT_LEVELS; ++l) {
    const float scale = M.scale[l];
    const unsigned int res = (unsigned int)M.res[l], size = M.size[l];
    unsigned int g[3];
    float fr[3];
    level_cell(M, l, L.u, g, fr);
    const float de0 = Rb[(2 * l) * HT_LD + lane], de1 = Rb[(2 * l + 1) * HT_LD + lane];
    const float ge0 = second_order ? E[(2 * l) * HT_LD + lane] : 0.f, ge1 = second_order ? E[(2 * l + 1) * HT_LD + lane] : 0.f;
    float hx[3] = {0.f, 0.f, 0.f};  // sum_corner v * d2 w / (d frac_c d frac_c') t_c'
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
      float wd[3], sg[3], w = 1.f;
      unsigned int c3[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int bit = (corner >> c) & 1;
        wd[c] = bit ? fr[c] : HT_SUB(1.f, fr[c]);
        sg[c] = bit ? 1.f : -1.f;
        w = HT_MUL(w, wd[c]);
        c3[c] = g[c] + bit;
      }
      const unsigned int idx = M.offset[l] + grid_index(c3, res, size);
      float g0 = w * de0, g1 = w * de1;
      if (second_order) {
        const float dwt = L.t[0] * sg[0] * wd[1] * wd[2] + L.t[1] * sg[1] * wd[0] * wd[2] + L.t[2] * sg[2] * wd[0] * wd[1];
        g0 = fmaf(scale * ge0, dwt, g0);
        g1 = fmaf(scale * ge1, dwt, g1);
        if (F.ray_grad) {
          const float2 f = HT_LDG2(tab + idx);
          const float v = ge0 * f.x + ge1 * f.y;
          hx[0] += v * sg[0] * (sg[1] * wd[2] * L.t[1] + sg[2] * wd[1] * L.t[2]);
          hx[1] += v * sg[1] * (sg[0] * wd[2] * L.t[0] + sg[2] * wd[0] * L.t[2]);
          hx[2] += v * sg[2] * (sg[0] * wd[1] * L.t[0] + sg[1] * wd[0] * L.t[1]);
        }
      }
      HT_ATOMIC_ADD(gtable + 2 * (size_t)idx, g0);
      HT_ATOMIC_ADD(gtable + 2 * (size_t)idx + 1, g1);
    }
    if (second_order && F.ray_grad) {
#pragma unroll
      for (int c = 0; c < 3; ++c) L.du[c] = fmaf(scale * scale, hx[c], L.du[c]);
    }
  }
}
// warp phase N: second-order parts of d sigma_net.0 (q (x) r) and d sigma_net.1[0,:] (column sums of the masked tangent)
HT_DEV void phase_n(float* G, const float* B, int lane) {
  wgrad(G + O_S0, 32, B + R_Q * HT_LD, 64, B + R_P * HT_LD, 32, lane);
  wcolsum(G + O_S1, B + R_H * HT_LD, 64, lane);
}

// d L / d (direction) through the SH encoding: (d SH / d d)^T dsh
HT_DEV void sh4_bwd(const float (&d)[3], const float (&g)[16], float (&out)[3]) {
  const float X = d[0], Y = d[1], Z = d[2];
  const float a1 = 0.48860251190291987f, a2 = 1.0925484305920792f, a3 = 0.94617469575755997f, a5 = 0.54627421529603959f,
              a6 = 0.59004358992664352f, a7 = 2.8906114426405538f, a8 = 0.45704579946446572f, a9 = 0.3731763325901154f,
              a10 = 1.4453057213202769f;
  out[0] = -a1 * g[3] + a2 * Y * g[4] - a2 * Z * g[7] + 2.f * a5 * X * g[8] - 6.f * a6 * X * Y * g[9] + a7 * Y * Z * g[10] +
           a8 * (1.f - 5.f * Z * Z) * g[13] + 2.f * a10 * X * Z * g[14] + a6 * (-3.f * X * X + 3.f * Y * Y) * g[15];
  out[1] = -a1 * g[1] + a2 * X * g[4] - a2 * Z * g[5] - 2.f * a5 * Y * g[8] + a6 * (-3.f * X * X + 3.f * Y * Y) * g[9] +
           a7 * X * Z * g[10] + a8 * (1.f - 5.f * Z * Z) * g[11] - 2.f * a10 * Y * Z * g[14] + 6.f * a6 * X * Y * g[15];
  out[2] = a1 * g[2] - a2 * Y * g[5] + 2.f * a3 * Z * g[6] - a2 * X * g[7] + a7 * X * Y * g[10] - 10.f * a8 * Y * Z * g[11] +
           a9 * (15.f * Z * Z - 3.f) * g[12] - 10.f * a8 * X * Z * g[13] + a10 * (X * X - Y * Y) * g[14];
}

}  // namespace ht
}  // namespace mnrf
