// Second layout of the hash-grid field's backward (see hash_train_math.cuh for the math and the buffer map): a warp works on a
// tile of 16 points with TWO lanes per point (lane = 16 * half + point).  The two lanes of a point split every layer's outputs,
// the 16 grid levels and the 32 encoding features between them, so the per-warp shared buffer is [260 rows][17] = 17.7 KB and
// EIGHT warps fit next to the two 44.8 KB weight / gradient images of a CTA (the first layout fits four and is bound by issue
// latency with one warp per scheduler: profiles/r01_v4_hash_bwd_ncu_metrics.txt).  Instruction count per point is unchanged: a
// warp instruction still covers 32 (point, output-half) pairs.  Because two lanes now share a buffer column, every step that
// reads what the other half wrote is separated by a warp barrier (the "steps" below); partial sums over the encoding features
// (g_u, d L / d u) are exchanged through rows 32.. of the P scratch.  Compiles as plain C++ like the first layout
// (tests/emu/hash_train_emu.cpp runs both).
#pragma once
#include "hash_train_math.cuh"

namespace mnrf {
namespace ht2 {

using namespace ht;

constexpr int LD2 = 17;   // row stride (odd: conflict-free for column and row access)
constexpr int NP2 = 16;   // points per warp tile
constexpr int HT2_WARP_FLOATS = HT_ROWS * LD2;
constexpr int HT2_J = 48; // Jacobian of the lane's own 8 levels
constexpr int R_X = R_P + 32;  // 12 scratch rows inside P: partial g_u of half 0 / 1 (6), partial d L / d u (6)

// ---- step a1: encoding + Jacobian of the lane's own 8 levels -------------------------------------------------------------------
template <class MetaT>
HT_DEV void step_a1(float* B, const float* table, const MetaT& M, Lane& L, float* J, int lane) {
  const int col = lane & 15, h = lane >> 4;
  const float2* tab = reinterpret_cast<const float2*>(table);
  float* E = B + R_E * LD2;
#pragma unroll 2
  for (int li = 0; li < 8; ++li) {
    const int l = 8 * h + li;
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
    E[(2 * l) * LD2 + col] = a0;
    E[(2 * l + 1) * LD2 + col] = a1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      J[3 * (2 * li) + c] = scale * j0[c];
      J[3 * (2 * li + 1) + c] = scale * j1[c];
    }
  }
}
// ---- step a2: hidden layer of sigma_net, the lane's 32 outputs -------------------------------------------------------------------
HT_DEV void step_a2(const float* Wt, float* B, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float* H = B + R_H * LD2;
  for (int o0 = 32 * h; o0 < 32 * h + 32; o0 += 8) {
    float acc[8];
    rows_dot<8, LD2>(Wt + O_S0, 32, o0, B + R_E * LD2, 32, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) H[(o0 + j) * LD2 + col] = fmaxf(acc[j], 0.f);
  }
}
// ---- step a3: [sigma | geo_feat] (8 rows per lane) and the SH rows of the colour-net input ---------------------------------------
HT_DEV void step_a3(const float* Wt, float* B, Lane& L, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float* S = B + R_S * LD2;
  float* Rb = B + R_R * LD2;
  {
    float acc[8];
    rows_dot<8, LD2>(Wt + O_S1, 64, 8 * h, B + R_H * LD2, 64, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) S[(8 * h + j) * LD2 + col] = acc[j];
  }
#pragma unroll
  for (int j = 0; j < 15; ++j) L.dgeo[j] = 0.f;
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
  for (int i = 0; i < 16; ++i)
    if ((i >> 3) == h) Rb[i * LD2 + col] = sh[i];
}
// ---- steps a4 / a5: the two hidden layers of the colour net, the lane's 32 outputs each --------------------------------------------
HT_DEV void step_a4(const float* Wt, float* B, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float* P = B + R_P * LD2;
  for (int o0 = 32 * h; o0 < 32 * h + 32; o0 += 8) {
    float acc[8];
    rows_dot<8, LD2>(Wt + O_C0, LD_C0, o0, B + R_R * LD2, 32, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) P[(o0 + j) * LD2 + col] = fmaxf(acc[j], 0.f);
  }
}
HT_DEV void step_a5(const float* Wt, float* B, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float* Q = B + R_Q * LD2;
  for (int o0 = 32 * h; o0 < 32 * h + 32; o0 += 8) {
    float acc[8];
    rows_dot<8, LD2>(Wt + O_C1, 64, o0, B + R_P * LD2, 64, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) Q[(o0 + j) * LD2 + col] = fmaxf(acc[j], 0.f);
  }
}
// ---- step a6: colour output (both lanes of the point), d colour pre-activation -----------------------------------------------------
HT_DEV void step_a6(const float* Wt, float* B, Lane& L, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float acc[3];
  rows_dot<3, LD2>(Wt + O_C2, 64, 0, B + R_Q * LD2, 64, col, acc);
  float* D = B + R_D * LD2;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float rgb = sigmoidf_(acc[c]);
    L.dco[c] = L.dr[1 + c] * rgb * (1.f - rgb);
    if (h == 0) D[c * LD2 + col] = L.dco[c];
  }
}
HT_DEV void step_b(float* G, const float* B, int lane) { wgrad<LD2, NP2>(G + O_C2, 64, B + R_D * LD2, 3, B + R_Q * LD2, 64, lane); }
HT_DEV void step_c(const float* Wt, float* B, const Lane& L, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float* Q = B + R_Q * LD2;
  for (int o = 32 * h; o < 32 * h + 32; ++o) {
    const float g = Wt[O_C2 + o] * L.dco[0] + Wt[O_C2 + 64 + o] * L.dco[1] + Wt[O_C2 + 128 + o] * L.dco[2];
    Q[o * LD2 + col] = Q[o * LD2 + col] > 0.f ? g : 0.f;
  }
}
HT_DEV void step_d(float* G, const float* B, int lane) { wgrad<LD2, NP2>(G + O_C1, 64, B + R_Q * LD2, 64, B + R_P * LD2, 64, lane); }
HT_DEV void step_e(const float* Wt, float* B, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float* P = B + R_P * LD2;
  for (int k0 = 32 * h; k0 < 32 * h + 32; k0 += 8) {
    float acc[8];
    cols_dot8<LD2>(Wt + O_C1, 64, k0, B + R_Q * LD2, 64, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) P[(k0 + j) * LD2 + col] = P[(k0 + j) * LD2 + col] > 0.f ? acc[j] : 0.f;
  }
}
HT_DEV void step_f(float* G, const float* B, int lane) { wgrad<LD2, NP2>(G + O_C0, LD_C0, B + R_P * LD2, 64, B + R_R * LD2, 32, lane); }
// ---- step g1: d [SH | geo] of the colour net (both lanes); hidden layer of the normal head, the lane's 32 outputs -------------------
HT_DEV void step_g1(const float* Wt, float* B, const Flags& F, Lane& L, int lane) {
  const int col = lane & 15, h = lane >> 4;
  const float* P = B + R_P * LD2;
  {
    float acc[8];
    cols_dot8<LD2>(Wt + O_C0, LD_C0, 16, P, 64, col, acc);  // column 16 is padding
#pragma unroll
    for (int j = 1; j < 8; ++j) L.dgeo[j - 1] += acc[j];
    cols_dot8<LD2>(Wt + O_C0, LD_C0, 24, P, 64, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dgeo[7 + j] += acc[j];
  }
  if (F.ray_grad) {
    float acc[8];
    cols_dot8<LD2>(Wt + O_C0, LD_C0, 0, P, 64, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dsh[j] = acc[j];
    cols_dot8<LD2>(Wt + O_C0, LD_C0, 8, P, 64, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dsh[8 + j] = acc[j];
  }
  if (F.has_normal) {
    float* Q = B + R_Q * LD2;
    const float* GEO = B + (R_S + 1) * LD2;
    for (int o0 = 32 * h; o0 < 32 * h + 32; o0 += 8) {
      float acc[8];
      rows_dot<8, LD2>(Wt + O_N0, LD_N0, o0, GEO, 15, col, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) Q[(o0 + j) * LD2 + col] = fmaxf(acc[j], 0.f);
    }
  }
}
// ---- step g2: hidden layer of the mirror head (rows 0..31 of P: d c1 has been consumed by both lanes), the lane's 16 outputs -------
HT_DEV void step_g2(const float* Wt, float* B, const Flags& F, int lane) {
  const int col = lane & 15, h = lane >> 4;
  if (!F.has_mirror) return;
  float* P = B + R_P * LD2;
  const float* GEO = B + (R_S + 1) * LD2;
  for (int o0 = 16 * h; o0 < 16 * h + 16; o0 += 8) {
    float acc[8];
    rows_dot<8, LD2>(Wt + O_M0, LD_M0, o0, GEO, 15, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = acc[j] + Wt[O_M0B + o0 + j];
      P[(o0 + j) * LD2 + col] = v > 0.f ? v : 0.01f * v;
    }
  }
}
// ---- step g3: outputs of the normal and mirror heads (both lanes) and their gradients ----------------------------------------------
HT_DEV void step_g3(const float* Wt, float* B, const Flags& F, Lane& L, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float* D = B + R_D * LD2;
  if (F.has_normal) {
    float nraw[3];
    rows_dot<3, LD2>(Wt + O_N1, 64, 0, B + R_Q * LD2, 64, col, nraw);
    const float dy[3] = {L.dr[5], L.dr[6], L.dr[7]};
    normalize_bwd(nraw, dy, L.dnraw);
    if (h == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) D[c * LD2 + col] = L.dnraw[c];
    }
  }
  if (F.has_mirror) {
    float mp[1];
    rows_dot<1, LD2>(Wt + O_M2, 32, 0, B + R_P * LD2, 32, col, mp);
    const float m = sigmoidf_(mp[0] + Wt[O_M2B]);
    L.dmp = L.dr[4] * m * (1.f - m);
    if (h == 0) D[3 * LD2 + col] = L.dmp;
  }
}
HT_DEV void step_h(float* G, const float* B, const Flags& F, int lane) {
  if (F.has_normal) wgrad<LD2, NP2>(G + O_N1, 64, B + R_D * LD2, 3, B + R_Q * LD2, 64, lane);
  if (F.has_mirror) {
    wgrad<LD2, NP2>(G + O_M2, 32, B + (R_D + 3) * LD2, 1, B + R_P * LD2, 32, lane);
    wcolsum<LD2, NP2>(G + O_M2B, B + (R_D + 3) * LD2, 1, lane);
  }
}
HT_DEV void step_i(const float* Wt, float* B, const Flags& F, const Lane& L, int lane) {
  const int col = lane & 15, h = lane >> 4;
  if (F.has_normal) {
    float* Q = B + R_Q * LD2;
    for (int o = 32 * h; o < 32 * h + 32; ++o) {
      const float g = Wt[O_N1 + o] * L.dnraw[0] + Wt[O_N1 + 64 + o] * L.dnraw[1] + Wt[O_N1 + 128 + o] * L.dnraw[2];
      Q[o * LD2 + col] = Q[o * LD2 + col] > 0.f ? g : 0.f;
    }
  }
  if (F.has_mirror) {
    float* P = B + R_P * LD2;
    for (int o = 16 * h; o < 16 * h + 16; ++o) {
      const float a = P[o * LD2 + col];
      P[o * LD2 + col] = Wt[O_M2 + o] * L.dmp * (a > 0.f ? 1.f : 0.01f);
    }
  }
}
HT_DEV void step_j(float* G, const float* B, const Flags& F, int lane) {
  const float* GEO = B + (R_S + 1) * LD2;
  if (F.has_normal) wgrad<LD2, NP2>(G + O_N0, LD_N0, B + R_Q * LD2, 64, GEO, 15, lane);
  if (F.has_mirror) {
    wgrad<LD2, NP2>(G + O_M0, LD_M0, B + R_P * LD2, 32, GEO, 15, lane);
    wcolsum<LD2, NP2>(G + O_M0B, B + R_P * LD2, 32, lane);
  }
}
// ---- step k1: d geo_feat from the heads (both lanes), d [sigma, geo] -> S rows (8 per lane) ----------------------------------------
HT_DEV void step_k1(const float* Wt, float* B, const Flags& F, Lane& L, int lane) {
  const int col = lane & 15, h = lane >> 4;
  if (F.has_normal && !F.detach_normal) {
    float acc[8];
    cols_dot8<LD2>(Wt + O_N0, LD_N0, 0, B + R_Q * LD2, 64, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dgeo[j] += acc[j];
    cols_dot8<LD2>(Wt + O_N0, LD_N0, 8, B + R_Q * LD2, 64, col, acc);  // column 15 is padding
#pragma unroll
    for (int j = 0; j < 7; ++j) L.dgeo[8 + j] += acc[j];
  }
  if (F.has_mirror && L.mirror_on) {
    float acc[8];
    cols_dot8<LD2>(Wt + O_M0, LD_M0, 0, B + R_P * LD2, 32, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) L.dgeo[j] += acc[j];
    cols_dot8<LD2>(Wt + O_M0, LD_M0, 8, B + R_P * LD2, 32, col, acc);
#pragma unroll
    for (int j = 0; j < 7; ++j) L.dgeo[8 + j] += acc[j];
  }
  float* S = B + R_S * LD2;
  if (h == 0) {
    S[col] = L.dr[0];
#pragma unroll
    for (int j = 0; j < 7; ++j) S[(1 + j) * LD2 + col] = L.dgeo[j];
  } else {
#pragma unroll
    for (int j = 7; j < 15; ++j) S[(1 + j) * LD2 + col] = L.dgeo[j];
  }
}
// ---- step k2: d hidden of sigma_net, the lane's 32 rows -> P -----------------------------------------------------------------------
HT_DEV void step_k2(const float* Wt, float* B, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float* P = B + R_P * LD2;
  const float* H = B + R_H * LD2;
  for (int k0 = 32 * h; k0 < 32 * h + 32; k0 += 8) {
    float acc[8];
    cols_dot8<LD2>(Wt + O_S1, 64, k0, B + R_S * LD2, 16, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) P[(k0 + j) * LD2 + col] = H[(k0 + j) * LD2 + col] > 0.f ? acc[j] : 0.f;
  }
}
HT_DEV void step_l(float* G, const float* B, int lane) {
  wgrad<LD2, NP2>(G + O_S1, 64, B + R_S * LD2, 16, B + R_H * LD2, 64, lane);
  wgrad<LD2, NP2>(G + O_S0, 32, B + R_P * LD2, 64, B + R_E * LD2, 32, lane);
}
// ---- step m1: d enc of the lane's 16 features -> R; q = W1[0,:] relu'(h), the lane's 32 rows -> Q ----------------------------------
HT_DEV void step_m1(const float* Wt, float* B, int lane, bool second_order) {
  const int col = lane & 15, h = lane >> 4;
  float* Rb = B + R_R * LD2;
  for (int k0 = 16 * h; k0 < 16 * h + 16; k0 += 8) {
    float acc[8];
    cols_dot8<LD2>(Wt + O_S0, 32, k0, B + R_P * LD2, 64, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) Rb[(k0 + j) * LD2 + col] = acc[j];
  }
  if (second_order) {
    float* Q = B + R_Q * LD2;
    const float* H = B + R_H * LD2;
    for (int o = 32 * h; o < 32 * h + 32; ++o) Q[o * LD2 + col] = H[o * LD2 + col] > 0.f ? Wt[O_S1 + o] : 0.f;
  }
}
// ---- step m2: g_e of the lane's 16 features -> E; partial g_u over them -> scratch ----------------------------------------------------
HT_DEV void step_m2(const float* Wt, float* B, const float* J, int lane) {
  const int col = lane & 15, h = lane >> 4;
  float* E = B + R_E * LD2;
  for (int k0 = 16 * h; k0 < 16 * h + 16; k0 += 8) {
    float acc[8];
    cols_dot8<LD2>(Wt + O_S0, 32, k0, B + R_Q * LD2, 64, col, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) E[(k0 + j) * LD2 + col] = acc[j];
  }
  float gu[3] = {0.f, 0.f, 0.f};
  for (int kk = 0; kk < 16; ++kk) {
    const float ge = E[(16 * h + kk) * LD2 + col];
#pragma unroll
    for (int c = 0; c < 3; ++c) gu[c] = fmaf(ge, J[3 * kk + c], gu[c]);
  }
  float* X = B + R_X * LD2;
#pragma unroll
  for (int c = 0; c < 3; ++c) X[(3 * h + c) * LD2 + col] = gu[c];
}
// ---- step m3: t = d L / d g_u (both lanes); r = J t for the lane's 16 features -> P rows 0..31 -------------------------------------
template <class MetaT>
HT_DEV void step_m3(float* B, const MetaT& M, Lane& L, const float* J, int lane) {
  const int col = lane & 15, h = lane >> 4;
  const float* X = B + R_X * LD2;
  const float inv2b = 1.f / (2.f * M.bound);
  float gu[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) gu[c] = X[c * LD2 + col] + X[(3 + c) * LD2 + col];
  const float v[3] = {-gu[0] * inv2b, -gu[1] * inv2b, -gu[2] * inv2b};
  const float dy[3] = {L.dr[8], L.dr[9], L.dr[10]};
  float dv[3];
  normalize_bwd(v, dy, dv);
#pragma unroll
  for (int c = 0; c < 3; ++c) L.t[c] = -dv[c] * inv2b;
  float* P = B + R_P * LD2;
  for (int kk = 0; kk < 16; ++kk)
    P[(16 * h + kk) * LD2 + col] = L.t[0] * J[3 * kk] + L.t[1] * J[3 * kk + 1] + L.t[2] * J[3 * kk + 2];
}
// ---- step m4: masked tangent (the lane's 32 rows, in place over H); table scatter of the lane's 8 levels; partial d L / d u ---------
template <class MetaT>
HT_DEV void step_m4(const float* Wt, float* B, const float* table, float* gtable, const MetaT& M, const Flags& F, Lane& L,
                    const float* J, int lane, bool second_order) {
  const int col = lane & 15, h = lane >> 4;
  float* E = B + R_E * LD2;
  float* H = B + R_H * LD2;
  const float* Rb = B + R_R * LD2;
  if (second_order) {
    for (int o0 = 32 * h; o0 < 32 * h + 32; o0 += 8) {
      float acc[8];
      rows_dot<8, LD2>(Wt + O_S0, 32, o0, B + R_P * LD2, 32, col, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) H[(o0 + j) * LD2 + col] = H[(o0 + j) * LD2 + col] > 0.f ? acc[j] : 0.f;
    }
  } else {
    L.t[0] = L.t[1] = L.t[2] = 0.f;
  }
  float du[3] = {0.f, 0.f, 0.f};
  if (F.ray_grad) {
    for (int kk = 0; kk < 16; ++kk) {
      const float de = Rb[(16 * h + kk) * LD2 + col];
#pragma unroll
      for (int c = 0; c < 3; ++c) du[c] = fmaf(de, J[3 * kk + c], du[c]);
    }
  }
  if (L.valid) {
    const float2* tab = reinterpret_cast<const float2*>(table);
    for (int li = 0; li < 8; ++li) {
      const int l = 8 * h + li;
      const float scale = M.scale[l];
      const unsigned int res = (unsigned int)M.res[l], size = M.size[l];
      unsigned int g[3];
      float fr[3];
      level_cell(M, l, L.u, g, fr);
      const float de0 = Rb[(2 * l) * LD2 + col], de1 = Rb[(2 * l + 1) * LD2 + col];
      const float ge0 = second_order ? E[(2 * l) * LD2 + col] : 0.f, ge1 = second_order ? E[(2 * l + 1) * LD2 + col] : 0.f;
      float hx[3] = {0.f, 0.f, 0.f};
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
        for (int c = 0; c < 3; ++c) du[c] = fmaf(scale * scale, hx[c], du[c]);
      }
    }
  }
  float* X = B + R_X * LD2;
#pragma unroll
  for (int c = 0; c < 3; ++c) X[(6 + 3 * h + c) * LD2 + col] = du[c];
}
// ---- step n: second-order weight gradients; step o: d L / d u of the point (lane half 0 uses it) ------------------------------------
HT_DEV void step_n(float* G, const float* B, int lane) {
  wgrad<LD2, NP2>(G + O_S0, 32, B + R_Q * LD2, 64, B + R_P * LD2, 32, lane);
  wcolsum<LD2, NP2>(G + O_S1, B + R_H * LD2, 64, lane);
}
HT_DEV void step_o(const float* B, Lane& L, int lane) {
  const int col = lane & 15;
  const float* X = B + R_X * LD2;
#pragma unroll
  for (int c = 0; c < 3; ++c) L.du[c] = X[(6 + c) * LD2 + col] + X[(9 + c) * LD2 + col];
}

}  // namespace ht2
}  // namespace mnrf
