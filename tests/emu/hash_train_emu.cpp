// TEST INFRASTRUCTURE ONLY.  Compiles mirror_nerf_b200/csrc/hash_train_math.cuh (the per-warp math of the hash-grid field's
// backward kernel, csrc/train_hash.cu::k_hash_bwd) as plain C++ and runs its phases lane by lane, so that the hand-derived
// backward (table scatter, small MLPs, double backward through the analytic normal, ray gradients) can be compared with the
// oracle's torch autograd in the CPU-only build container.  Nothing in the product path links or calls this.
#include <string.h>
#include <vector>

#include "../../mirror_nerf_b200/csrc/hash_train_math.cuh"

using namespace mnrf::ht;

struct Meta {
  float bound;
  float scale[16];
  int res[16];
  unsigned int offset[16];
  unsigned int size[16];
};

extern "C" int hash_bwd_emu(const float* table, const float* wref, const Meta* Mp, const float* x, const float* d, const float* DR,
                            const int* mirror_on, int P, int has_normal, int has_mirror, int compute_normal, int detach_normal,
                            int detach_mask, int ray_grad, int second_order, float* gtable, float* gsmall, float* dxd) {
  const Meta& M = *Mp;
  Flags F;
  F.has_normal = has_normal; F.has_mirror = has_mirror; F.compute_normal = compute_normal;
  F.detach_normal = detach_normal; F.detach_mask = detach_mask; F.ray_grad = ray_grad;
  // wref / gsmall: the 11 small tensors concatenated in their own [out][in] layouts; the kernel works on row-padded images
  std::vector<float> Wt(HT_NW, 0.f), G(HT_NW, 0.f), B(HT_WARP_FLOATS, 0.f);
  {
    int pos = 0;
    for (int t = 1; t < 12; ++t)
      for (int r = 0; r < small_rows(t); ++r)
        for (int c = 0; c < small_cols(t); ++c) Wt[small_offset(t) + r * small_ld(t) + small_col(t, c)] = wref[pos++];
  }
  const int n_tiles = (P + 31) / 32;
  for (int tile = 0; tile < n_tiles; ++tile) {
    Lane L[32];
    static float J[32][HT_J];
    for (int lane = 0; lane < 32; ++lane) {
      Lane& l = L[lane];
      const int p_raw = tile * 32 + lane;
      l.valid = p_raw < P;
      const int p = l.valid ? p_raw : P - 1;
      for (int c = 0; c < 3; ++c) {
        l.u[c] = (x[p * 3 + c] + M.bound) / (2.f * M.bound);
        l.d[c] = d[p * 3 + c];
      }
      for (int i = 0; i < HT_DR_STRIDE; ++i) l.dr[i] = l.valid ? DR[(size_t)p * HT_DR_STRIDE + i] : 0.f;
      l.dr[11] = 0.f;
      l.mirror_on = !detach_mask && mirror_on[p];
      l.dmp = 0.f;
      l.dnraw[0] = l.dnraw[1] = l.dnraw[2] = 0.f;
      for (int i = 0; i < 16; ++i) l.dsh[i] = 0.f;
    }
#define LANES(call) for (int lane = 0; lane < 32; ++lane) { call; }
    LANES(phase_a(Wt.data(), B.data(), table, M, F, L[lane], J[lane], lane))
    LANES(phase_b(G.data(), B.data(), lane))
    LANES(phase_c(Wt.data(), B.data(), L[lane], lane))
    LANES(phase_d(G.data(), B.data(), lane))
    LANES(phase_e(Wt.data(), B.data(), lane))
    LANES(phase_f(G.data(), B.data(), lane))
    LANES(phase_g(Wt.data(), B.data(), F, L[lane], lane))
    LANES(phase_h(G.data(), B.data(), F, lane))
    LANES(phase_i(Wt.data(), B.data(), F, L[lane], lane))
    LANES(phase_j(G.data(), B.data(), F, lane))
    LANES(phase_k(Wt.data(), B.data(), F, L[lane], lane))
    LANES(phase_l(G.data(), B.data(), lane))
    LANES(phase_m(Wt.data(), B.data(), table, gtable, M, F, L[lane], J[lane], lane, second_order != 0))
    if (second_order) LANES(phase_n(G.data(), B.data(), lane))
    if (dxd != nullptr) {
      const float inv2b = 1.f / (2.f * M.bound);
      for (int lane = 0; lane < 32; ++lane) {
        if (!L[lane].valid) continue;
        float gd[3];
        sh4_bwd(L[lane].d, L[lane].dsh, gd);
        float* o = dxd + (size_t)(tile * 32 + lane) * HT_DXD_STRIDE;
        o[0] = L[lane].du[0] * inv2b; o[1] = L[lane].du[1] * inv2b; o[2] = L[lane].du[2] * inv2b;
        o[3] = gd[0]; o[4] = gd[1]; o[5] = gd[2]; o[6] = 0.f; o[7] = 0.f;
      }
    }
  }
  {
    int pos = 0;
    for (int t = 1; t < 12; ++t)
      for (int r = 0; r < small_rows(t); ++r)
        for (int c = 0; c < small_cols(t); ++c) gsmall[pos++] += G[small_offset(t) + r * small_ld(t) + small_col(t, c)];
  }
  return 0;
}

extern "C" int hash_bwd_emu_nw(void) {
  int n = 0;
  for (int t = 1; t < 12; ++t) n += small_rows(t) * small_cols(t);
  return n;
}
