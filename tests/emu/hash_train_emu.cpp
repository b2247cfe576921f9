// TEST INFRASTRUCTURE ONLY.  Compiles mirror_nerf_b200/csrc/hash_train_math.cuh (the per-warp math of the hash-grid field's
// backward kernel, csrc/train_hash.cu::k_hash_bwd) as plain C++ and runs its phases lane by lane, so that the hand-derived
// backward (table scatter, small MLPs, double backward through the analytic normal, ray gradients) can be compared with the
// oracle's torch autograd in the CPU-only build container.  Nothing in the product path links or calls this.
#include <string.h>
#include <vector>

#include "../../mirror_nerf_b200/csrc/hash_train_math.cuh"
#include "../../mirror_nerf_b200/csrc/hash_train_math2.cuh"

using namespace mnrf::ht;

// Lanes of a step run one after the other here; on the GPU they run concurrently.  A step that reads, in one lane, what another
// lane writes in the same step would be a race there: running every step in both lane orders makes such a step give different
// results here (tests/test_hash_train_emu.py compares both orders bit for bit).
static int g_reverse = 0;
extern "C" void hash_bwd_emu_set_reverse(int r) { g_reverse = r; }

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
#define LANES(call) for (int li_ = 0; li_ < 32; ++li_) { const int lane = g_reverse ? 31 - li_ : li_; call; }
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


// second layout (hash_train_math2.cuh): 16 points per warp tile, two lanes per point, a barrier after every step
extern "C" int hash_bwd_emu2(const float* table, const float* wref, const Meta* Mp, const float* x, const float* d, const float* DR,
                             const int* mirror_on, int P, int has_normal, int has_mirror, int compute_normal, int detach_normal,
                             int detach_mask, int ray_grad, int second_order, float* gtable, float* gsmall, float* dxd) {
  using namespace mnrf::ht2;
  const Meta& M = *Mp;
  Flags F;
  F.has_normal = has_normal; F.has_mirror = has_mirror; F.compute_normal = compute_normal;
  F.detach_normal = detach_normal; F.detach_mask = detach_mask; F.ray_grad = ray_grad;
  std::vector<float> Wt(HT_NW, 0.f), G(HT_NW, 0.f), B(HT2_WARP_FLOATS, 0.f);
  {
    int pos = 0;
    for (int t = 1; t < 12; ++t)
      for (int r = 0; r < small_rows(t); ++r)
        for (int c = 0; c < small_cols(t); ++c) Wt[small_offset(t) + r * small_ld(t) + small_col(t, c)] = wref[pos++];
  }
  const bool so = second_order != 0;
  const int n_tiles = (P + NP2 - 1) / NP2;
  for (int tile = 0; tile < n_tiles; ++tile) {
    Lane L[32];
    static float J[32][HT2_J];
    for (int lane = 0; lane < 32; ++lane) {
      Lane& l = L[lane];
      const int p_raw = tile * NP2 + (lane & 15);
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
    float* b = B.data();
    const float* w = Wt.data();
    float* g = G.data();
    LANES(step_a1(b, table, M, L[lane], J[lane], lane))
    LANES(step_a2(w, b, lane))
    LANES(step_a3(w, b, L[lane], lane))
    LANES(step_a4(w, b, lane))
    LANES(step_a5(w, b, lane))
    LANES(step_a6(w, b, L[lane], lane))
    LANES(step_b(g, b, lane))
    LANES(step_c(w, b, L[lane], lane))
    LANES(step_d(g, b, lane))
    LANES(step_e(w, b, lane))
    LANES(step_f(g, b, lane))
    LANES(step_g1(w, b, F, L[lane], lane))
    LANES(step_g2(w, b, F, lane))
    LANES(step_g3(w, b, F, L[lane], lane))
    LANES(step_h(g, b, F, lane))
    LANES(step_i(w, b, F, L[lane], lane))
    LANES(step_j(g, b, F, lane))
    LANES(step_k1(w, b, F, L[lane], lane))
    LANES(step_k2(w, b, lane))
    LANES(step_l(g, b, lane))
    LANES(step_m1(w, b, lane, so))
    if (so) {
      LANES(step_m2(w, b, J[lane], lane))
      LANES(step_m3(b, M, L[lane], J[lane], lane))
    }
    LANES(step_m4(w, b, table, gtable, M, F, L[lane], J[lane], lane, so))
    if (so) LANES(step_n(g, b, lane))
    LANES(step_o(b, L[lane], lane))
    if (dxd != nullptr) {
      const float inv2b = 1.f / (2.f * M.bound);
      for (int lane = 0; lane < 16; ++lane) {
        if (!L[lane].valid) continue;
        float gd[3];
        sh4_bwd(L[lane].d, L[lane].dsh, gd);
        float* o = dxd + (size_t)(tile * NP2 + lane) * HT_DXD_STRIDE;
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
