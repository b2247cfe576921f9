// C ABI of libmnrf.so (include/mnrf.h): argument checking, scratch management and the per-level launch
// sequence of render_rays (R/models/rendering.py:54-369).  No torch types, no CPU compute path.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace mnrf {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool first_use_on_device(int tag, int* num_sms) {
  constexpr int MAX_DEV = 64;
  static std::atomic<unsigned char> done[MAX_DEV][TAG_COUNT];
  static std::atomic<int> sms[MAX_DEV];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) dev = 0;
  if (num_sms != nullptr) {
    int n = sms[dev].load(std::memory_order_relaxed);
    if (n == 0) {
      if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
      sms[dev].store(n, std::memory_order_relaxed);
    }
    *num_sms = n;
  }
  return done[dev][tag].exchange(1, std::memory_order_relaxed) == 0;
}

// ---- per-launch timing of the tcgen05 field kernel (roofline numbers of bench.py) ------------------
struct ProfRec { cudaEvent_t e0, e1; double flops; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t g_prof_pending = nullptr;

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(cudaStream_t st) {
  if (!g_prof_on) return;
  g_prof_pending = prof_event();
  cudaEventRecord(g_prof_pending, st);
}
void prof_end(cudaStream_t st, double flops) {
  if (!g_prof_on || g_prof_pending == nullptr) return;
  ProfRec r{g_prof_pending, prof_event(), flops};
  cudaEventRecord(r.e1, st);
  g_prof.push_back(r);
  g_prof_pending = nullptr;
}

bool can_fuse_composite(const mnrf_field* f, const mnrf_level_cfg* cfg, const float* noise, int S) {
  return f->kind == 0 && cfg->impl != MNRF_IMPL_FP32 && !cfg->compute_normal && !cfg->no_fused_composite &&
         (noise == nullptr || cfg->noise_std == 0.f) && S <= 512;
}

namespace {

inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

__global__ void k_unpack_raw(const float* __restrict__ raw, int n, float* __restrict__ sigma, float* __restrict__ rgb,
                             float* __restrict__ mirror, float* __restrict__ pn) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = *reinterpret_cast<const float4*>(raw + (size_t)i * 8);
  const float4 b = *reinterpret_cast<const float4*>(raw + (size_t)i * 8 + 4);
  if (sigma) sigma[i] = a.x;
  if (rgb) { rgb[i * 3] = a.y; rgb[i * 3 + 1] = a.z; rgb[i * 3 + 2] = a.w; }
  if (mirror) mirror[i] = b.x;
  if (pn) { pn[i * 3] = b.y; pn[i * 3 + 1] = b.z; pn[i * 3 + 2] = b.w; }
}

int check_impl(int impl) {
  MNRF_REQUIRE(impl == MNRF_IMPL_TC3 || impl == MNRF_IMPL_TC2 || impl == MNRF_IMPL_TC1 || impl == MNRF_IMPL_FP32,
               "unknown impl %d", impl);
  return 0;
}

int run_field(const mnrf_field* f, int impl, const FieldIO& io, cudaStream_t st) {
  if (f->kind == 1) return launch_field_hash(f, io, st);
  if (impl == MNRF_IMPL_FP32) return launch_field_fp32(f, io, st);
  return launch_field_tc(f, io, impl, st);  // MNRF_IMPL_TC1/2/3 == number of fp16-pass equivalents
}

// A full (non sigma-only) pass of the MLP field can composite inside the tensor-core kernel (field_tc.cu FUSE): no raw point
// records, no k_composite launch.  Not for analytic normals (their chain has its own kernel variant), sigma noise, or more
// than 512 samples per ray.
bool can_fuse(const mnrf_field* f, const mnrf_level_cfg* cfg, const float* noise, int S) {
  return can_fuse_composite(f, cfg, noise, S);
}

// field + compositor of one full pass: fused when possible, else field kernel -> raw records -> k_composite
int run_full_pass(const mnrf_field* f, const mnrf_level_cfg* cfg, FieldIO io, const float* noise, float* raw_buf, int n, int S,
                  const mnrf_composite_out& out, int* counter, unsigned long long* stats, cudaStream_t st, const int* n_dev) {
  if (can_fuse(f, cfg, noise, S)) {
    FusedComposite fc;
    fc.comp = out;
    fc.white_back = cfg->white_back;
    // early termination only when no per-sample output is wanted (the skipped samples would have weights < eps)
    fc.term_eps = (out.weights == nullptr && out.pred_normal == nullptr) ? cfg->early_termination_eps : 0.f;
    fc.work_counter = counter;
    fc.stats = stats;
    io.raw = nullptr;
    return launch_field_tc(f, io, cfg->impl, st, &fc);
  }
  MNRF_REQUIRE(out.weights != nullptr, "render_level: the unfused compositor needs the per-sample weights buffer");
  io.raw = raw_buf;
  if (run_field(f, cfg->impl, io, st)) return 1;
  return launch_composite(io.rays, io.z, raw_buf, 8, raw_buf, io.normal_out, noise, cfg->noise_std, n, S, cfg->white_back, out,
                          st, n_dev);
}

}  // namespace
}  // namespace mnrf

using namespace mnrf;

extern "C" {

const char* mnrf_last_error(void) { return g_err; }
int mnrf_abi_version(void) { return MNRF_ABI_VERSION; }
int64_t mnrf_launch_count(void) { return (int64_t)g_launches.load(); }
int mnrf_profile_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}
int mnrf_profile_collect(double* total_ms, double* total_flops, int64_t* launches) {
  double ms = 0.0, fl = 0.0;
  for (ProfRec& r : g_prof) {
    MNRF_CUDA_OK(cudaEventSynchronize(r.e1));
    float t = 0.f;
    MNRF_CUDA_OK(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t;
    fl += r.flops;
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = (int64_t)g_prof.size();
  g_prof.clear();
  return 0;
}
int mnrf_debug_set_trace(void* buf, int64_t capacity_events) {
  set_tc_trace(reinterpret_cast<unsigned long long*>(buf), (unsigned int)capacity_events);
  return 0;
}
int mnrf_debug_set_tc_schedule(int split) { return set_tc_split(split); }
int64_t mnrf_macs_full(void) {
  // trunk 491,264 + colour 102,144 + normal 33,152 + mirror 32,896
  return 659456;
}
int64_t mnrf_macs_sigma_only(void) { return 524416; }

int mnrf_field_create(mnrf_field** out, const float* const* tensors, void* stream) {
  MNRF_REQUIRE(out != nullptr && tensors != nullptr, "field_create: null argument");
  for (int i = 0; i < 24; ++i) MNRF_REQUIRE(tensors[i] != nullptr, "field_create: tensor %d is required", i);
  const bool hn = tensors[T_N0_W] != nullptr, hm = tensors[T_M0_W] != nullptr;
  for (int i = 24; i < 28; ++i) MNRF_REQUIRE((tensors[i] != nullptr) == hn, "field_create: normal_net tensors must be all set or all NULL");
  for (int i = 28; i < 32; ++i) MNRF_REQUIRE((tensors[i] != nullptr) == hm, "field_create: is_mirror_net tensors must be all set or all NULL");
  mnrf_field* f = new mnrf_field();
  f->kind = 0;
  f->hash_table = nullptr;
  f->hash_w = nullptr;
  f->hash_wref = nullptr;
  f->has_normal = hn;
  f->has_mirror = hm;
  f->L = make_f32_layout();
  f->f32 = nullptr;
  f->tc = nullptr;
  f->tc8 = nullptr;
  f->t32 = nullptr;
  if (cudaMalloc(&f->f32, sizeof(float) * f->L.total) != cudaSuccess || cudaMalloc(&f->tc, TC_TOTAL_BYTES) != cudaSuccess ||
      cudaMalloc(&f->tc8, 2 * (size_t)TC_TOTAL_BYTES) != cudaSuccess || cudaMalloc(&f->t32, T32_TOTAL_BYTES) != cudaSuccess) {
    set_error("field_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    mnrf_field_destroy(f);
    return 1;
  }
  // padding floats of the fp32 section (e.g. behind b_rgb / b_m2 in the epilogue table) are copied around but never used
  if (cudaMemsetAsync(f->f32, 0, sizeof(float) * f->L.total, S_(stream)) != cudaSuccess) { mnrf_field_destroy(f); return 1; }
  if (pack_field(f, tensors, S_(stream))) { mnrf_field_destroy(f); return 1; }
  *out = f;
  return 0;
}

int mnrf_hash_field_create(mnrf_field** out, const float* const* tensors, int64_t table_floats, float bound,
                           const float* level_scale, const int* level_res, const uint32_t* level_offset,
                           const uint32_t* level_size, void* stream) {
  MNRF_REQUIRE(out && tensors && level_scale && level_res && level_offset && level_size, "hash_field_create: null argument");
  for (int i = 0; i < 6; ++i) MNRF_REQUIRE(tensors[i] != nullptr, "hash_field_create: tensor %d is required", i);
  const bool hn = tensors[6] != nullptr, hm = tensors[8] != nullptr;
  MNRF_REQUIRE((tensors[7] != nullptr) == hn, "hash_field_create: normal_net tensors must be all set or all NULL");
  for (int i = 9; i < 12; ++i) MNRF_REQUIRE((tensors[i] != nullptr) == hm, "hash_field_create: is_mirror_net tensors must be all set or all NULL");
  MNRF_REQUIRE(bound > 0.f && table_floats > 0, "hash_field_create: bad bound / table size");
  long long need = 0;
  for (int l = 0; l < HG_LEVELS; ++l) {
    MNRF_REQUIRE(level_size[l] > 0 && level_res[l] > 0, "hash_field_create: bad level table");
    need = std::max(need, ((long long)level_offset[l] + level_size[l]) * 2);
  }
  MNRF_REQUIRE(need == table_floats, "hash_field_create: encoder.params has %lld floats, the level table needs %lld",
               (long long)table_floats, need);
  mnrf_field* f = new mnrf_field();
  f->kind = 1;
  f->has_normal = hn;
  f->has_mirror = hm;
  f->f32 = nullptr; f->tc = nullptr; f->tc8 = nullptr; f->t32 = nullptr;
  f->hash_table = nullptr; f->hash_w = nullptr; f->hash_wref = nullptr;
  f->hg.bound = bound;
  for (int l = 0; l < HG_LEVELS; ++l) {
    f->hg.scale[l] = level_scale[l]; f->hg.res[l] = level_res[l]; f->hg.offset[l] = level_offset[l]; f->hg.size[l] = level_size[l];
  }
  if (cudaMalloc(&f->hash_table, sizeof(float) * (size_t)table_floats) != cudaSuccess ||
      cudaMalloc(&f->hash_w, sizeof(float) * HW_TOTAL) != cudaSuccess ||
      cudaMalloc(&f->hash_wref, sizeof(float) * HASH_WREF_FLOATS) != cudaSuccess) {
    set_error("hash_field_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    mnrf_field_destroy(f);
    return 1;
  }
  if (pack_hash_field(f, tensors, table_floats, S_(stream))) { mnrf_field_destroy(f); return 1; }
  *out = f;
  return 0;
}

int mnrf_field_update(mnrf_field* f, const float* const* tensors, void* stream) {
  MNRF_REQUIRE(f != nullptr && tensors != nullptr, "field_update: null argument");
  if (f->kind == 1) {  // hash-grid field: tensors = the 12 pointers of mnrf_hash_field_create; same table size and head set
    for (int i = 0; i < 6; ++i) MNRF_REQUIRE(tensors[i] != nullptr, "field_update: hash-grid tensor %d is required", i);
    MNRF_REQUIRE((tensors[6] != nullptr) == (f->has_normal != 0) && (tensors[7] != nullptr) == (f->has_normal != 0) &&
                     (tensors[8] != nullptr) == (f->has_mirror != 0),
                 "field_update: head set changed");
    for (int i = 9; i < 12; ++i) MNRF_REQUIRE((tensors[i] != nullptr) == (f->has_mirror != 0), "field_update: head set changed");
    long long table_floats = 0;
    for (int l = 0; l < HG_LEVELS; ++l) table_floats = std::max(table_floats, ((long long)f->hg.offset[l] + f->hg.size[l]) * 2);
    return pack_hash_field(f, tensors, table_floats, S_(stream));
  }
  MNRF_REQUIRE((tensors[T_N0_W] != nullptr) == (f->has_normal != 0) && (tensors[T_M0_W] != nullptr) == (f->has_mirror != 0),
               "field_update: head set changed");
  return pack_field(f, tensors, S_(stream));
}

void mnrf_field_destroy(mnrf_field* f) {
  if (f == nullptr) return;
  if (f->f32) cudaFree(f->f32);
  if (f->tc) cudaFree(f->tc);
  if (f->tc8) cudaFree(f->tc8);
  if (f->t32) cudaFree(f->t32);
  if (f->hash_table) cudaFree(f->hash_table);
  if (f->hash_w) cudaFree(f->hash_w);
  if (f->hash_wref) cudaFree(f->hash_wref);
  delete f;
}
int mnrf_field_has_normal(const mnrf_field* f) { return f ? f->has_normal : 0; }
int mnrf_field_has_mirror(const mnrf_field* f) { return f ? f->has_mirror : 0; }

int mnrf_field_eval_rays(const mnrf_field* f, int impl, const float* rays, const float* z, int n_rays, int S,
                         int sigma_only, float* raw, float* sigma_out, float* normal_out, void* stream) {
  MNRF_REQUIRE(f && rays && z, "field_eval_rays: null argument");
  MNRF_REQUIRE(n_rays >= 0 && S >= 1, "field_eval_rays: bad sizes");
  if (check_impl(impl)) return 2;
  MNRF_REQUIRE(sigma_only ? sigma_out != nullptr : raw != nullptr, "field_eval_rays: output buffer missing");
  MNRF_REQUIRE((long long)n_rays * S < (1ll << 31), "field_eval_rays: too many points for one call");
  if (n_rays == 0) return 0;
  cudaStream_t st = S_(stream);
  float* dirbias = nullptr;
  if (!sigma_only && f->kind == 0) {
    MNRF_CUDA_OK(cudaMallocAsync(&dirbias, sizeof(float) * (size_t)n_rays * WH, st));
    if (launch_dirbias(f, rays, n_rays, 8, 0, dirbias, st)) return 1;
  }
  FieldIO io{};
  io.rays = rays; io.z = z; io.x = nullptr; io.x_stride = 0; io.dirbias = dirbias;
  io.n_points = n_rays * S; io.S = S; io.sigma_only = sigma_only;
  io.raw = sigma_only ? nullptr : raw; io.sigma_out = sigma_out; io.normal_out = normal_out; io.geo_out = nullptr;
  int rc = run_field(f, impl, io, st);
  if (dirbias) MNRF_CUDA_OK(cudaFreeAsync(dirbias, st));
  return rc;
}

int mnrf_field_eval_points(const mnrf_field* f, int impl, const float* x, int B, int sigma_only, float* sigma,
                           float* rgb, float* is_mirror, float* pred_normal, float* normal, float* geo_feat,
                           void* stream) {
  MNRF_REQUIRE(f && x, "field_eval_points: null argument");
  if (check_impl(impl)) return 2;
  if (B <= 0) return 0;
  if (geo_feat != nullptr) impl = MNRF_IMPL_FP32;  // only the fp32 kernel exports the 256-wide feature
  cudaStream_t st = S_(stream);
  if (f->kind == 1) {
    // hash-grid field: x (B, 3+3) = [xyz | d] (identity embeddings, R/train.py:69-70) or (B,3) when sigma_only
    MNRF_REQUIRE(geo_feat == nullptr, "field_eval_points: the hash-grid field has no geo_feat export");
    MNRF_REQUIRE(normal == nullptr || !sigma_only, "field_eval_points: hash-grid analytic normals need the full pass");
    float* rawh = nullptr;
    MNRF_CUDA_OK(cudaMallocAsync(&rawh, sizeof(float) * (size_t)B * 8, st));
    FieldIO ioh{};
    ioh.x = x; ioh.x_stride = sigma_only ? 3 : 6; ioh.n_points = B; ioh.S = 1; ioh.sigma_only = sigma_only; ioh.raw = rawh;
    ioh.normal_out = normal;
    int rch = launch_field_hash(f, ioh, st);
    if (rch == 0) {
      k_unpack_raw<<<(B + 255) / 256, 256, 0, st>>>(rawh, B, sigma, sigma_only ? nullptr : rgb, sigma_only ? nullptr : is_mirror, pred_normal);
      MNRF_LAUNCH_OK();
    }
    MNRF_CUDA_OK(cudaFreeAsync(rawh, st));
    return rch;
  }
  const int stride = sigma_only ? 3 : 3 + IN_DIR;
  float* dirbias = nullptr;
  float* raw = nullptr;
  MNRF_CUDA_OK(cudaMallocAsync(&raw, sizeof(float) * (size_t)B * 8, st));
  if (!sigma_only) {
    MNRF_CUDA_OK(cudaMallocAsync(&dirbias, sizeof(float) * (size_t)B * WH, st));
    if (launch_dirbias(f, x, B, stride, 1, dirbias, st)) return 1;
  }
  FieldIO io{};
  io.rays = nullptr; io.z = nullptr; io.x = x; io.x_stride = stride; io.dirbias = dirbias;
  io.n_points = B; io.S = 1; io.sigma_only = sigma_only;
  // the predicted normal is returned even for sigma_only (mirror_nerf.py:154-161): evaluate the heads we need
  const bool need_heads = !sigma_only || pred_normal != nullptr;
  int rc;
  if (sigma_only && need_heads) {
    // normal head without the colour branch: only the fp32 kernel has that combination
    io.raw = raw; io.sigma_out = nullptr; io.normal_out = normal; io.geo_out = geo_feat;
    rc = launch_field_fp32(f, io, st);
  } else {
    io.raw = sigma_only ? nullptr : raw; io.sigma_out = sigma_only ? sigma : nullptr;
    io.normal_out = normal; io.geo_out = geo_feat;
    rc = run_field(f, impl, io, st);
  }
  if (rc == 0 && io.raw != nullptr) {
    k_unpack_raw<<<(B + 255) / 256, 256, 0, st>>>(raw, B, sigma, sigma_only ? nullptr : rgb,
                                                 sigma_only ? nullptr : is_mirror, pred_normal);
    MNRF_LAUNCH_OK();
  }
  if (dirbias) MNRF_CUDA_OK(cudaFreeAsync(dirbias, st));
  MNRF_CUDA_OK(cudaFreeAsync(raw, st));
  return rc;
}

int mnrf_generate_rays(int H, int W, float focal, const float* c2w_host, float near, float far, float* rays,
                       void* stream) {
  MNRF_REQUIRE(c2w_host && rays && focal > 0.f, "generate_rays: bad argument");
  return launch_generate_rays(H, W, focal, c2w_host, near, far, rays, S_(stream));
}

int mnrf_embed(const float* x, int n, int n_freqs, float* out, void* stream) {
  MNRF_REQUIRE(x && out && n_freqs >= 0 && n_freqs <= 32, "embed: bad argument");
  return launch_embed(x, n, n_freqs, out, S_(stream));
}

int mnrf_coarse_z(const float* rays, int n, const float* z_steps, int S, int use_disp, float perturb,
                  const float* perturb_u, float* z_out, void* stream) {
  MNRF_REQUIRE(rays && z_steps && z_out, "coarse_z: null argument");
  return launch_coarse_z(rays, n, z_steps, S, use_disp, perturb, perturb_u, z_out, S_(stream));
}

int mnrf_searchsorted_right(const float* cdf, int n, int n_cdf, const float* u, int n_u, int u_stride, int64_t* inds,
                            void* stream) {
  MNRF_REQUIRE(cdf && u && inds && n_cdf >= 1, "searchsorted: bad argument");
  return launch_searchsorted(cdf, n, n_cdf, u, n_u, u_stride, inds, S_(stream));
}

int mnrf_sample_pdf(const float* z_coarse, const float* weights, int n, int S, int n_imp, const float* u, int u_stride,
                    float* z_fine, float* samples, int64_t* inds, float* cdf, void* stream) {
  MNRF_REQUIRE(z_coarse && weights && u && z_fine, "sample_pdf: null argument");
  MNRF_REQUIRE(u_stride == 0 || u_stride == n_imp, "sample_pdf: u_stride must be 0 or n_imp");
  return launch_sample_pdf(z_coarse, nullptr, weights, S, 1, n, S, n_imp, u, u_stride, z_fine, samples, inds, cdf,
                           S_(stream));
}

int mnrf_sample_pdf_bins(const float* bins, const float* weights, int n, int n_w, int n_imp, const float* u,
                         int u_stride, float* samples, int64_t* inds, float* cdf, void* stream) {
  MNRF_REQUIRE(bins && weights && u && samples, "sample_pdf_bins: null argument");
  MNRF_REQUIRE(u_stride == 0 || u_stride == n_imp, "sample_pdf_bins: u_stride must be 0 or n_imp");
  return launch_sample_pdf(nullptr, bins, weights, n_w, 0, n, n_w + 2, n_imp, u, u_stride, nullptr, samples, inds, cdf,
                           S_(stream));
}

int mnrf_composite(const float* rays, const float* z, const float* sigma, int sigma_stride, const float* raw,
                   const float* normal, const float* noise, float noise_std, int n, int S, int white_back,
                   const mnrf_composite_out* out, void* stream) {
  MNRF_REQUIRE(rays && z && sigma && out, "composite: null argument");
  return launch_composite(rays, z, sigma, sigma_stride, raw, normal, noise, noise_std, n, S, white_back, *out,
                          S_(stream));
}

// scratch layout of one level: [dirbias n x 128][coarse records][second-pass records][work counter 256 B].  A sigma-only coarse
// pass keeps one float per point, a full pass 8 floats per point unless it composites inside the field kernel (no records).
struct LevelScratch { size_t dirbias, buf_c, buf_f, total; };
static LevelScratch level_scratch(const mnrf_field* coarse, const mnrf_field* fine, int n, const mnrf_level_cfg* cfg, bool has_noise) {
  const size_t Sc = cfg->n_samples, Sf = cfg->n_samples + cfg->n_importance;
  const bool sig_only = cfg->test_time && fine != nullptr;
  const mnrf_field* second = cfg->rerun_coarse_on_fine ? coarse : fine;
  const bool two = cfg->n_importance > 0 && second != nullptr;
  const float* noise = has_noise ? reinterpret_cast<const float*>(1) : nullptr;
  // without the field handles (mnrf_level_workspace_bytes) assume the unfused sequence: an upper bound
  const bool fuse_c = coarse != nullptr && !sig_only && can_fuse_composite(coarse, cfg, noise, (int)Sc);
  const bool fuse_f = coarse != nullptr && two && can_fuse_composite(second, cfg, noise, (int)Sf);
  LevelScratch L;
  L.dirbias = align256(sizeof(float) * (size_t)n * WH);
  L.buf_c = (coarse != nullptr && sig_only) ? align256(sizeof(float) * (size_t)n * Sc) : (fuse_c ? 0 : align256(sizeof(float) * (size_t)n * Sc * 8));
  L.buf_f = (cfg->n_importance > 0 && (coarse == nullptr || two)) ? (fuse_f ? 0 : align256(sizeof(float) * (size_t)n * Sf * 8)) : 0;
  L.total = L.dirbias + L.buf_c + L.buf_f + 256;
  return L;
}

int64_t mnrf_level_workspace_bytes(int n, const mnrf_level_cfg* cfg) {
  if (cfg == nullptr || n < 0) return -1;
  return (int64_t)level_scratch(nullptr, nullptr, n, cfg, true).total;
}

int64_t mnrf_level_workspace_bytes_for(const mnrf_field* coarse, const mnrf_field* fine, int n, const mnrf_level_cfg* cfg,
                                       int with_sigma_noise) {
  if (cfg == nullptr || coarse == nullptr || n < 0) return -1;
  return (int64_t)level_scratch(coarse, fine, n, cfg, with_sigma_noise != 0 && cfg->noise_std != 0.f).total;
}

int mnrf_render_level(const mnrf_field* coarse, const mnrf_field* fine, const float* rays, int n,
                      const mnrf_level_cfg* cfg, const mnrf_level_rng* rng, const float* z_steps, const float* u_det,
                      void* workspace, int64_t workspace_bytes, const mnrf_level_out* out, void* stream) {
  return mnrf::render_level(coarse, fine, rays, n, cfg, rng, z_steps, u_det, workspace, workspace_bytes, out, stream, nullptr);
}

}  // extern "C"

// One render level; `n_dev` (optional, device) = number of rays that are really alive: every kernel is launched for n rays and
// processes min(n, *n_dev) of them (mnrf_render_recursive sizes deeper levels on the device, no host read-back).
int mnrf::render_level(const mnrf_field* coarse, const mnrf_field* fine, const float* rays, int n,
                       const mnrf_level_cfg* cfg, const mnrf_level_rng* rng, const float* z_steps, const float* u_det,
                       void* workspace, int64_t workspace_bytes, const mnrf_level_out* out, void* stream, const int* n_dev) {
  MNRF_REQUIRE(coarse && rays && cfg && z_steps && out, "render_level: null argument");
  if (check_impl(cfg->impl)) return 2;
  if (n <= 0) return 0;
  const int Sc = cfg->n_samples, Ni = cfg->n_importance, Sf = Sc + Ni;
  MNRF_REQUIRE(Sc >= 1 && Ni >= 0, "render_level: bad sample counts");
  MNRF_REQUIRE((long long)n * Sf < (1ll << 31), "render_level: too many points; split the ray batch");
  static const mnrf_level_rng no_rng0 = {nullptr, nullptr, nullptr, nullptr};
  const bool has_noise = (rng != nullptr ? rng : &no_rng0)->noise_coarse != nullptr || (rng != nullptr ? rng : &no_rng0)->noise_fine != nullptr;
  const LevelScratch LS = level_scratch(coarse, fine, n, cfg, has_noise && cfg->noise_std != 0.f);
  MNRF_REQUIRE(workspace != nullptr && workspace_bytes >= (int64_t)LS.total,
               "render_level: workspace too small (%lld bytes needed: mnrf_level_workspace_bytes_for)", (long long)LS.total);
  MNRF_REQUIRE(out->z_coarse && out->coarse.weights && out->coarse.opacity, "render_level: coarse outputs missing");
  MNRF_REQUIRE(cfg->dir_source == nullptr || coarse->kind == 0, "render_level: view_dir needs the MLP field");
  const float* dir_src = cfg->dir_source != nullptr ? cfg->dir_source : rays;   // rendering.py:276 view_dir
  static const mnrf_level_rng no_rng = {nullptr, nullptr, nullptr, nullptr};
  if (rng == nullptr) rng = &no_rng;
  cudaStream_t st = S_(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* dirbias = reinterpret_cast<float*>(ws); ws += LS.dirbias;
  float* buf_c = reinterpret_cast<float*>(ws);   ws += LS.buf_c;
  float* buf_f = reinterpret_cast<float*>(ws);   ws += LS.buf_f;
  int* counter = reinterpret_cast<int*>(ws);
  const int impl = cfg->impl;

  // ---- coarse pass (rendering.py:271-305) ----
  if (launch_coarse_z(rays, n, z_steps, Sc, cfg->use_disp, cfg->perturb, rng->perturb_u, out->z_coarse, st, n_dev)) return 1;
  const bool sig_only = cfg->test_time && fine != nullptr;  // rendering.py:139
  FieldIO io{};
  io.rays = rays; io.z = out->z_coarse; io.n_points = n * Sc; io.S = Sc; io.sigma_only = sig_only; io.n_rays_dev = n_dev;
  if (sig_only) {
    io.sigma_out = buf_c;
    if (run_field(coarse, impl, io, st)) return 1;
    if (launch_composite(rays, out->z_coarse, buf_c, 1, nullptr, nullptr, rng->noise_coarse, cfg->noise_std, n, Sc,
                         cfg->white_back, out->coarse, st, n_dev))
      return 1;
  } else {
    if (coarse->kind == 0 && launch_dirbias(coarse, dir_src, n, 8, 0, dirbias, st, n_dev)) return 1;
    io.dirbias = dirbias;
    if (cfg->compute_normal) {
      MNRF_REQUIRE(out->normal_coarse != nullptr, "render_level: compute_normal needs normal_coarse");
      io.normal_out = out->normal_coarse;
    }
    if (run_full_pass(coarse, cfg, io, rng->noise_coarse, buf_c, n, Sc, out->coarse, counter, nullptr, st, n_dev)) return 1;
  }

  // ---- importance resampling + second pass (rendering.py:307-361) ----
  const mnrf_field* second = cfg->rerun_coarse_on_fine ? coarse : fine;
  if (Ni > 0 && second != nullptr) {
    MNRF_REQUIRE(out->z_fine && out->fine.opacity, "render_level: fine outputs missing");
    const float* u = rng->u_pdf != nullptr ? rng->u_pdf : u_det;
    MNRF_REQUIRE(u != nullptr, "render_level: need u_pdf or u_det");
    if (launch_sample_pdf(out->z_coarse, nullptr, out->coarse.weights, Sc, 1, n, Sc, Ni, u,
                          rng->u_pdf != nullptr ? Ni : 0, out->z_fine, nullptr, nullptr, nullptr, st, n_dev))
      return 1;
    if (second->kind == 0 && launch_dirbias(second, dir_src, n, 8, 0, dirbias, st, n_dev)) return 1;
    FieldIO io2{};
    io2.rays = rays; io2.z = out->z_fine; io2.n_points = n * Sf; io2.S = Sf; io2.sigma_only = 0; io2.n_rays_dev = n_dev;
    io2.dirbias = dirbias;
    if (cfg->compute_normal) {
      MNRF_REQUIRE(out->normal_fine != nullptr, "render_level: compute_normal needs normal_fine");
      io2.normal_out = out->normal_fine;
    }
    if (run_full_pass(second, cfg, io2, rng->noise_fine, buf_f, n, Sf, out->fine, counter, cfg->stats, st, n_dev)) return 1;
  }
  return 0;
}

extern "C" {

int mnrf_render_level_host(const mnrf_field* coarse, const mnrf_field* fine, const float* rays_host, int n,
                           const mnrf_level_cfg* cfg, const float* z_steps_host, const float* u_det_host, float* rgb,
                           float* depth, float* opacity, float* mirror_mask, float* surface_normal, float* x_surface,
                           void* stream) {
  MNRF_REQUIRE(coarse && rays_host && cfg && z_steps_host, "render_level_host: null argument");
  if (n <= 0) return 0;
  cudaStream_t st = S_(stream);
  const int Sc = cfg->n_samples, Ni = cfg->n_importance, Sf = Sc + Ni;
  const bool second = Ni > 0 && (cfg->rerun_coarse_on_fine || fine != nullptr);
  const int64_t wsb = mnrf_level_workspace_bytes(n, cfg);
  // one stream-ordered allocation for everything that lives on the device during the call
  const size_t fN = (size_t)n;
  size_t off = 0;
  auto take = [&](size_t floats) { size_t o = off; off += align256(floats * sizeof(float)); return o; };
  const size_t o_rays = take(fN * 8), o_zs = take(Sc), o_u = take(Ni > 0 ? Ni : 1), o_zc = take(fN * Sc),
               o_wc = take(fN * Sc), o_opc = take(fN), o_zf = take(second ? fN * Sf : 1),
               o_wf = take(second ? fN * Sf : 1), o_op = take(fN), o_rgb = take(fN * 3), o_dep = take(fN),
               o_mm = take(fN), o_sn = take(fN * 3), o_xs = take(fN * 3), o_ws = take((size_t)wsb / 4 + 64);
  uint8_t* d = nullptr;
  MNRF_CUDA_OK(cudaMallocAsync(&d, off, st));
  auto P = [&](size_t o) { return reinterpret_cast<float*>(d + o); };
  MNRF_CUDA_OK(cudaMemcpyAsync(P(o_rays), rays_host, fN * 8 * sizeof(float), cudaMemcpyHostToDevice, st));
  MNRF_CUDA_OK(cudaMemcpyAsync(P(o_zs), z_steps_host, Sc * sizeof(float), cudaMemcpyHostToDevice, st));
  if (Ni > 0) {
    MNRF_REQUIRE(u_det_host != nullptr, "render_level_host: u_det missing");
    MNRF_CUDA_OK(cudaMemcpyAsync(P(o_u), u_det_host, Ni * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  mnrf_level_out lo;
  memset(&lo, 0, sizeof(lo));
  lo.z_coarse = P(o_zc);
  lo.coarse.weights = P(o_wc);
  lo.coarse.opacity = P(o_opc);
  mnrf_composite_out& last = second ? lo.fine : lo.coarse;
  if (second) { lo.z_fine = P(o_zf); lo.fine.weights = P(o_wf); }
  const bool last_full = second || !(cfg->test_time && fine != nullptr);
  last.opacity = P(o_op);
  if (last_full) {
    last.rgb = P(o_rgb); last.depth = P(o_dep); last.mirror_mask = P(o_mm);
    last.surface_normal = P(o_sn); last.x_surface = P(o_xs);
  }
  mnrf_level_cfg c = *cfg;
  c.compute_normal = 0;
  int rc = mnrf_render_level(coarse, fine, P(o_rays), n, &c, nullptr, P(o_zs), P(o_u), P(o_ws), wsb, &lo, stream);
  if (rc == 0) {
    auto back = [&](float* h, size_t o, size_t floats) {
      return h == nullptr ? cudaSuccess : cudaMemcpyAsync(h, P(o), floats * sizeof(float), cudaMemcpyDeviceToHost, st);
    };
    MNRF_CUDA_OK(back(opacity, o_op, fN));
    if (last_full) {
      MNRF_CUDA_OK(back(rgb, o_rgb, fN * 3));
      MNRF_CUDA_OK(back(depth, o_dep, fN));
      MNRF_CUDA_OK(back(mirror_mask, o_mm, fN));
      MNRF_CUDA_OK(back(surface_normal, o_sn, fN * 3));
      MNRF_CUDA_OK(back(x_surface, o_xs, fN * 3));
    }
  }
  MNRF_CUDA_OK(cudaFreeAsync(d, st));
  MNRF_CUDA_OK(cudaStreamSynchronize(st));
  return rc;
}

int mnrf_reflect_rays(const float* rays, const float* x_surface, const float* normal, float* mask, int n,
                      float near_secondary, float* secondary, float* reflect_dir, int* any_mirror, void* stream) {
  MNRF_REQUIRE(rays && x_surface && normal && mask && secondary, "reflect_rays: null argument");
  return launch_reflect(rays, x_surface, normal, mask, n, near_secondary, secondary, reflect_dir, any_mirror,
                        S_(stream));
}

int mnrf_compact_rows(const float* in, const float* mask, int n, int row_floats, float* out, int* index, int* count,
                      void* stream) {
  MNRF_REQUIRE(mask && count && (out == nullptr || in != nullptr), "compact_rows: null argument");
  return launch_compact(in, mask, n, row_floats, out, index, count, S_(stream));
}

int mnrf_axpy_rows(float* dense, const float* compact, const int* index, int n, int c, float alpha, float beta,
                   void* stream) {
  MNRF_REQUIRE(dense != nullptr && c >= 1, "axpy_rows: bad argument");
  return launch_axpy_rows(dense, compact, index, n, c, alpha, beta, S_(stream));
}

int mnrf_blend_reflection(const float* base_rgb, const float* mask, const float* child_rgb, const float* child_depth,
                          const int* index, int n, float* rgb_out, float* rgb_reflect, float* depth_reflect,
                          void* stream) {
  MNRF_REQUIRE(base_rgb && mask && child_rgb && rgb_out, "blend_reflection: null argument");
  return launch_blend(base_rgb, mask, child_rgb, child_depth, index, n, rgb_out, rgb_reflect, depth_reflect,
                      S_(stream));
}

}  // extern "C"
