// tcgen05 field kernel (v2): MirrorNeRF.forward (R/models/mirror_nerf.py:101-212) fused per 128-point tile:
//   o + d*z  ->  positional encoding  ->  8x256 trunk (skip at layer 5)  ->  sigma / folded normal head
//   ->  mirror head  ->  final 256x256  ->  dir layer (+ per-ray dir term)  ->  rgb      -> 8 floats/point.
//
// One persistent CTA per SM, 12 warps:
//   warp 0        weight producer: cp.async.bulk (TMA 1-D) of pre-packed 8 KB B-operand blobs, 8-stage mbarrier ring
//   warp 1        MMA issuer: one thread issues tcgen05.mma (M=128, N=128, K=16, fp16 -> fp32 in TMEM)
//   warps 4..11   epilogue/PE: TMEM -> registers (tcgen05.ld) -> bias/ReLU -> fp16 hi/lo split -> next layer's
//                 A operand written in place into shared memory in the UMMA K-major core-matrix layout.
//
// Pipelining: every 256-wide layer is accumulated as two 128-column halves, half-major, each half with its own
// "accumulator complete" mbarrier; the epilogue of half h produces two 64-column K chunks of the next layer, each with its
// own "chunk ready" mbarrier, and the two 256-column TMEM accumulators alternate by layer.  So the epilogue of layer l
// overlaps the second half of layer l's MMAs and the first K chunks of layer l+1's (simulated period max(T_mma, T_epi)
// for T_epi <= 0.75 T_mma; v1, with one barrier per layer, measured T_epi + T_mma/2).
//
// Precision: operands are split x = hi + lo in fp16 (weights pre-scaled by 2^s per layer) and
// each product is formed as hi*hi + lo*hi + hi*lo with fp32 accumulation ("3x" mode, fp32-grade);
// PREC3 == false drops the lo terms (speed mode).
//
// Shared memory (1 CTA/SM): A hi/lo 2x64 KB | PE hi/lo 2x16 KB | 8 x 8 KB weight stages | 2 KB partial sums.
#include "common.cuh"

namespace mnrf {
namespace {

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 384;
// warp 0: weight producer; warp 1: MMA issuer; warps 4..11: epilogue/PE (TMEM lane quarter = warp % 4).
// (Measured: putting the issuer at the highest warp id of its scheduler partition is 7-16 % slower.)
constexpr int WARP_PRODUCER = 0;
constexpr int WARP_MMA = 1;
constexpr int NUM_WSTAGES = 4;      // 3x mode; the 1x mode also uses the idle A_lo region: 8 stages
constexpr uint32_t WSTAGE_BYTES = 2 * TC_BLOB_BYTES;  // 16 KB: tc3 = [hi,lo] of one K32 chunk; tc1 = hi of two K32 chunks

constexpr uint32_t SM_A_HI = 0;
constexpr uint32_t SM_A_LO = 65536;
constexpr uint32_t SM_PE_HI = 131072;
constexpr uint32_t SM_PE_LO = 147456;
constexpr uint32_t SM_WST = 163840;
constexpr uint32_t SM_PART = SM_WST + NUM_WSTAGES * WSTAGE_BYTES;  // 229376: float4[128]
constexpr uint32_t SM_BAR = SM_PART + 2048;                        // 231424
constexpr uint32_t SM_TOTAL = SM_BAR + 256;                        // 231680 <= 232448

// barrier slots (8 bytes each)
constexpr int BAR_W_FULL = 0;    // [4] (8 slots reserved)
constexpr int BAR_W_EMPTY = 8;   // [4]
constexpr int BAR_PE = 16;       // PE chunk written (8 warp arrivals)
constexpr int BAR_A = 17;        // [4] A 64-column chunk written (4 warp arrivals)
constexpr int BAR_ACC = 21;      // [4] accumulator half complete (tcgen05.commit): [buffer][half]
constexpr int BAR_AFREE = 25;    // [4] A 64-column chunk no longer read by any issued MMA (tcgen05.commit)
constexpr int BAR_TMEM_SLOT = 30;

constexpr uint32_t IDESC_N128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, K-major

struct TcParams {
  const float* f32;        // fp32 section
  const uint8_t* tc;       // packed blobs
  int b_trunk[8];
  int b_final, b_m0, w_m2, b_m2, w_rgb, b_rgb, headw, headb, inv_scale;
  int has_normal, has_mirror;
  FieldIO io;
  int n_tiles;
  unsigned long long* trace;  // optional device-side event trace of CTA 0: [count, (clock, tag)...]
  unsigned int trace_cap;
  int debug;  // timing experiments (MNRF_TC_DEBUG): 1 = 16-byte weight copies, 2 = no MMA issue, 4 = no epilogue math
};

// device-side tracing (mnrf_debug_set_trace): lane 0 of a warp of CTA 0 logs (clock64, tag) with plain stores into its own
// region of the buffer (no atomics, so the perturbation is one clock read + one store): region r = who, 8192 events each
struct TraceCtx { unsigned int n; };
__device__ __forceinline__ void trace_ev(const TcParams& P, TraceCtx& tc, int lane, int who, int ev, int a, int b) {
#ifdef MNRF_TC_TRACE
  if (P.trace != nullptr && blockIdx.x == 0 && lane == 0 && tc.n < 8192u) {
    unsigned long long* r = P.trace + 1 + (size_t)who * 2 * 8192 + 2 * tc.n;
    r[0] = (unsigned long long)clock64();
    r[1] = ((unsigned long long)who << 24) | ((unsigned long long)ev << 16) | ((unsigned long long)a << 8) | (unsigned long long)b;
    ++tc.n;
  }
#endif
}
// timing-experiment knobs (MNRF_TC_DEBUG) are compiled in only for bring-up builds (make EXTRA=-DMNRF_TC_TRACE)
#ifdef MNRF_TC_TRACE
#define MNRF_DBG(P, bit) ((P).debug & (bit))
#else
#define MNRF_DBG(P, bit) 0
#endif

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("mnrf field_tc: mbarrier timeout (block %d thread %d bar@%u parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// wait for this thread's outstanding tcgen05.ld, then pin the destination registers behind the wait
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void pin32(uint32_t (&r)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(r[i]));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
// one lane of a converged warp (the pattern ptxas turns into ELECT + predicated uniform-datapath instructions)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar_sync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }

// K-major, no-swizzle operand descriptor.  Core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes;
// LBO = byte distance between K-adjacent core matrices, SBO = byte distance between 8-row groups (M/N direction).
// Both operand kinds here are 128 rows tall: LBO = 128*16 = 2048, SBO = 128; one K16 step = 4096 bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}

// All operands share the descriptor's high word (SBO = 128 B, descriptor version 1); the low word is
// (address >> 4) | (LBO >> 4) << 16, so stepping through an operand is an integer add in 16-byte units
// (K16 step = 256, K32 chunk = 512, K64 chunk = 1024).
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | ((2048u >> 4) << 16); }
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %4};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(accumulate), "r"(DESC_HI), "r"(IDESC_N128)
      : "memory");
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- x = hi + lo in fp16, two values per 32-bit word (element 0 in the low half) -----------------------------------
template <bool RELU, bool PREC3>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (PREC3) {
    // hi rounded toward zero so that the residual of a non-negative value is non-negative: both ReLUs ride on the cvt
    if (RELU) asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    else      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float2 hf = __half22float2(*reinterpret_cast<__half2*>(&hi));
    if (RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hf.y), "f"(a - hf.x));
    else      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hf.y), "f"(a - hf.x));
  } else {
    if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    else      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    lo = 0;
  }
}

// store 8 consecutive K values (one 16-byte core-matrix row) of an A-type operand, hi and lo parts
template <bool RELU, bool PREC3>
__device__ __forceinline__ void store_a8(uint32_t hi_addr, uint32_t lo_addr, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2<RELU, PREC3>(v[2 * i], v[2 * i + 1], h[i], l[i]);
  st_shared_v4(hi_addr, h[0], h[1], h[2], h[3]);
  if (PREC3) st_shared_v4(lo_addr, l[0], l[1], l[2], l[3]);
}

// accurate sin/cos (arguments reach 2^9 * |x|): shared, not inlined 30 times
__device__ __noinline__ float2 sincos_pe(float a) {
  float s, c;
  sincosf(a, &s, &c);
  return make_float2(s, c);
}

// ---- positional encoding of one row, K range [32*HALF, 32*HALF+32) (mirror_nerf.py:33-38) ----------------
template <int HALF, bool PREC3>
__device__ __forceinline__ void pe_fill(const float (&x)[3], uint32_t pe_hi, uint32_t pe_lo, uint32_t rowoff) {
  constexpr int K0 = 32 * HALF;
  float vals[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) vals[i] = 0.f;  // k = 63 stays 0 (padding)
  if (HALF == 0) {
    vals[0] = x[0]; vals[1] = x[1]; vals[2] = x[2];
  }
#pragma unroll
  for (int f = 0; f < NFREQ_XYZ; ++f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int ks = 3 + 6 * f + c, kc = ks + 3;
      const bool need_s = (ks >= K0 && ks < K0 + 32), need_c = (kc >= K0 && kc < K0 + 32);
      if (need_s || need_c) {
        const float2 sc = sincos_pe(ldexpf(x[c], f));  // 2^f * x is exact
        if (need_s) vals[ks - K0] = sc.x;
        if (need_c) vals[kc - K0] = sc.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = vals[8 * j + i];
    const uint32_t off = (uint32_t)(4 * HALF + j) * 2048u + rowoff;
    store_a8<false, PREC3>(pe_hi + off, pe_lo + off, v);
  }
}

// ---- epilogue of 32 accumulator columns of one row ------------------------------------------------------------------
template <bool RELU, bool DOTS, bool WRITE_A, bool PREC3>
__device__ __forceinline__ void epi32(const uint32_t (&r)[32], const float4 (&b)[8], float inv, uint32_t s_hi,
                                      uint32_t s_lo, const float4* __restrict__ hw, float (&d)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 b0 = b[2 * j], b1 = b[2 * j + 1];
    float v[8];
    v[0] = fmaf(__uint_as_float(r[8 * j + 0]), inv, b0.x);
    v[1] = fmaf(__uint_as_float(r[8 * j + 1]), inv, b0.y);
    v[2] = fmaf(__uint_as_float(r[8 * j + 2]), inv, b0.z);
    v[3] = fmaf(__uint_as_float(r[8 * j + 3]), inv, b0.w);
    v[4] = fmaf(__uint_as_float(r[8 * j + 4]), inv, b1.x);
    v[5] = fmaf(__uint_as_float(r[8 * j + 5]), inv, b1.y);
    v[6] = fmaf(__uint_as_float(r[8 * j + 6]), inv, b1.z);
    v[7] = fmaf(__uint_as_float(r[8 * j + 7]), inv, b1.w);
    if (DOTS) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = fmaxf(v[i], 0.f);
        const float4 w = __ldg(hw + 8 * j + i);
        d[0] = fmaf(v[i], w.x, d[0]); d[1] = fmaf(v[i], w.y, d[1]);
        d[2] = fmaf(v[i], w.z, d[2]); d[3] = fmaf(v[i], w.w, d[3]);
      }
    }
    if (WRITE_A) store_a8<RELU, PREC3>(s_hi + (uint32_t)j * 2048u, s_lo + (uint32_t)j * 2048u, v);
  }
}

// 64 accumulator columns (this warp's share of a 128-column half): both TMEM loads in flight, bias prefetched
template <bool RELU, bool DOTS, bool WRITE_A, bool PREC3>
__device__ __forceinline__ void epi64(uint32_t tacc, const float* __restrict__ bias, float inv, uint32_t s_hi,
                                      uint32_t s_lo, const float4* __restrict__ hw, float (&d)[4], uint32_t free_bar,
                                      uint32_t free_parity) {
  uint32_t ra[32], rb[32];
  tmem_ld32(tacc, ra);
  tmem_ld32(tacc + 32u, rb);
  const float4* b4 = reinterpret_cast<const float4*>(bias);
  float4 b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = __ldg(b4 + i);
  tmem_wait_ld();
  pin32(ra);
  pin32(rb);
  if (WRITE_A && free_bar != 0u) mbar_wait(free_bar, free_parity);  // in-place overwrite: wait for the chunk's last reader
  epi32<RELU, DOTS, WRITE_A, PREC3>(ra, b, inv, s_hi, s_lo, hw, d);
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = __ldg(b4 + 8 + i);
  epi32<RELU, DOTS, WRITE_A, PREC3>(rb, b, inv, s_hi + 4u * 2048u, s_lo + 4u * 2048u, hw + 32, d);
}

// ---- step geometry --------------------------------------------------------------------------------
// TMEM columns: trunk layers alternate [0,256) / [256,512); mirror head (step 9) -> [0,128); final (step 8) -> [128,384);
// dir layer (step 10) -> [384,512).  Issue order: 0..7, 9, 8, 10.
__device__ __forceinline__ uint32_t acc_col(int s, int h) {
  return (s <= 7 ? (uint32_t)(s & 1) * 256u : (s == 9 ? 0u : (s == 8 ? 128u : 384u))) + (uint32_t)h * 128u;
}
__device__ __forceinline__ int acc_bar(int s, int h) { return s <= 7 ? 2 * (s & 1) + h : (s == 9 ? 0 : (s == 8 ? 1 + h : 3)); }
__device__ __forceinline__ int step_at(int i) { return i < 8 ? i : (i == 8 ? 9 : (i == 9 ? 8 : 10)); }

// ================================================================================================
template <bool PREC3>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_field_tc(const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bars = sbase + SM_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * BAR_TMEM_SLOT);
  const int n_issue = P.io.sigma_only ? 8 : 11;
  constexpr uint32_t NST = PREC3 ? NUM_WSTAGES : 8;  // weight stages (the 1x mode leaves the A_lo region free)
  auto stage_addr = [&](uint32_t st) { return sbase + (st < 4u ? SM_WST + st * WSTAGE_BYTES : SM_A_LO + (st - 4u) * WSTAGE_BYTES); };

  if (threadIdx.x == 0) {
    if (sbase & 127u) { printf("mnrf field_tc: unaligned dynamic smem base %u\n", sbase); __trap(); }
    for (int i = 0; i < 8; ++i) { mbar_init(bar(BAR_W_FULL + i), 1); mbar_init(bar(BAR_W_EMPTY + i), 1); }
    mbar_init(bar(BAR_PE), 8);
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_A + i), 4);
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_ACC + i), 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_AFREE + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == WARP_PRODUCER) {
    // =========================== weight producer (whole warp in lock-step, one elected lane issues) ===========
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      for (int i = 0; i < n_issue; ++i) {
        const int s = step_at(i);
        if (s == 9 && !P.has_mirror) continue;
        const uint8_t* src = P.tc + tc_step_offset(s);
        const int nst = tc_step_halves(s) * tc_step_chunks(s) / (PREC3 ? 1 : 2);  // stages of this step
        for (int si = 0; si < nst; ++si) {
          mbar_wait(bar(BAR_W_EMPTY + stage), phase ^ 1u);
          if (elect_one()) {
            const uint32_t dst = stage_addr(stage);
            const uint32_t fb = bar(BAR_W_FULL + stage);
            if (MNRF_DBG(P, 1)) {
              mbar_expect_tx(fb, 16);
              bulk_g2s(dst, src, 16, fb);
            } else if (PREC3) {  // [hi, lo] blobs of one K32 chunk are contiguous
              mbar_expect_tx(fb, WSTAGE_BYTES);
              const uint8_t* g = src + (size_t)si * 2 * TC_BLOB_BYTES;
#pragma unroll
              for (int piece = 0; piece < 4; ++piece)  // four 4 KB copies in flight per stage
                bulk_g2s(dst + piece * 4096u, g + piece * 4096, 4096u, fb);
            } else {      // hi blobs of two consecutive K32 chunks
              mbar_expect_tx(fb, WSTAGE_BYTES);
#pragma unroll
              for (int piece = 0; piece < 2; ++piece) {
                bulk_g2s(dst + piece * 4096u, src + (size_t)(2 * si) * 2 * TC_BLOB_BYTES + piece * 4096, 4096u, fb);
                bulk_g2s(dst + TC_BLOB_BYTES + piece * 4096u, src + (size_t)(2 * si + 1) * 2 * TC_BLOB_BYTES + piece * 4096, 4096u, fb);
              }
            }
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // =========================== MMA issuer: ONE elected thread runs the whole role ===========
    // (waits included: the tensor pipe queue hides them; per-block elect/reconverge cost ~25 cycles per MMA otherwise)
    if (elect_one()) {
    uint32_t stage = 0, phase = 0;
    uint32_t pe_phase = 0, a_phase = 0;  // a_phase: one parity bit per 64-column A chunk
    TraceCtx trc{0};
    const uint32_t dl_a_hi = desc_lo(sbase + SM_A_HI), dl_a_lo = desc_lo(sbase + SM_A_LO);
    const uint32_t dl_pe_hi = desc_lo(sbase + SM_PE_HI), dl_pe_lo = desc_lo(sbase + SM_PE_LO);
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      for (int i = 0; i < n_issue; ++i) {
        const int s = step_at(i);
        if (s == 9 && !P.has_mirror) continue;
        const int npairs = tc_step_chunks(s) >> 1;           // K64 chunks (= 64-column A chunks) of this step
        const int pe_pairs = (s == 0 || s == 4) ? 1 : 0;     // the leading K64 chunk comes from the PE buffer
        const bool a_reused = (s == 8 && P.has_mirror);      // h8 was already awaited by the mirror GEMM
        // the epilogue of this step overwrites the A buffer in place while this step's second half is still
        // reading it: release each 64-column chunk as soon as its last MMA has been issued
        const bool a_release = (s >= 1 && s <= 8) && !(s == 7 && P.io.sigma_only);
        const int nhalves = tc_step_halves(s);
        for (int h = 0; h < nhalves; ++h) {
          const uint32_t d_tmem = tmem + acc_col(s, h);
          const uint32_t acc_done = bar(BAR_ACC + acc_bar(s, h));
          const bool release = a_release && h == nhalves - 1;
          uint32_t accumulate = 0;
          trace_ev(P, trc, 0, 1, 1, s, h);
          for (int kp = 0; kp < npairs; ++kp) {
            // descriptor low words (16-byte units) of the A operand's hi / lo parts for this K64 chunk
            uint32_t ah, al;
            const int c = kp - pe_pairs;
            if (c < 0) {
              if (s == 0 && h == 0) { mbar_wait(bar(BAR_PE), pe_phase); pe_phase ^= 1u; }
              ah = dl_pe_hi; al = dl_pe_lo;
            } else {
              if (h == 0 && !a_reused) {  // first touch of a 64-column chunk of a new activation version
                mbar_wait(bar(BAR_A + c), (a_phase >> c) & 1u);
                a_phase ^= 1u << c;
                trace_ev(P, trc, 0, 1, 2, s, c);
              }
              ah = dl_a_hi + (uint32_t)c * 1024u; al = dl_a_lo + (uint32_t)c * 1024u;
            }
            const bool last = kp == npairs - 1;
            if (last) trace_ev(P, trc, 0, 1, 3, s, h);
            if (PREC3) {
              // two 16 KB stages ([W_hi | W_lo] of one K32 chunk each) per K64 chunk, issued from one elected block
              const uint32_t st0 = stage, ph0 = phase;
              if (++stage == NST) { stage = 0; phase ^= 1u; }
              const uint32_t st1 = stage, ph1 = phase;
              if (++stage == NST) { stage = 0; phase ^= 1u; }
              mbar_wait(bar(BAR_W_FULL + st0), ph0);
              mbar_wait(bar(BAR_W_FULL + st1), ph1);
              tc_fence_after();
              const uint32_t w0 = desc_lo(stage_addr(st0)), w1 = desc_lo(stage_addr(st1));
              {
                if (!(MNRF_DBG(P, 2))) {
                  tc_mma2(d_tmem, ah, w0, accumulate);                 // A_hi * W_hi  (k 0..15)
                  tc_mma2(d_tmem, al, w0, 1u);                         // A_lo * W_hi
                  tc_mma2(d_tmem, ah + 256u, w0 + 256u, 1u);           // (k 16..31)
                  tc_mma2(d_tmem, al + 256u, w0 + 256u, 1u);
                  tc_mma2(d_tmem, ah, w0 + 512u, 1u);                  // A_hi * W_lo
                  tc_mma2(d_tmem, ah + 256u, w0 + 768u, 1u);
                }
                tc_commit(bar(BAR_W_EMPTY + st0));
                if (!(MNRF_DBG(P, 2))) {
                  tc_mma2(d_tmem, ah + 512u, w1, 1u);                  // (k 32..47)
                  tc_mma2(d_tmem, al + 512u, w1, 1u);
                  tc_mma2(d_tmem, ah + 768u, w1 + 256u, 1u);           // (k 48..63)
                  tc_mma2(d_tmem, al + 768u, w1 + 256u, 1u);
                  tc_mma2(d_tmem, ah + 512u, w1 + 512u, 1u);
                  tc_mma2(d_tmem, ah + 768u, w1 + 768u, 1u);
                }
                tc_commit(bar(BAR_W_EMPTY + st1));
                if (release && c >= 0) tc_commit(bar(BAR_AFREE + c));
                if (last) tc_commit(acc_done);
              }
              accumulate = 1u;
            } else {
              mbar_wait(bar(BAR_W_FULL + stage), phase);  // one 16 KB stage per K64 chunk: W_hi of two K32 chunks
              tc_fence_after();
              const uint32_t wb = desc_lo(stage_addr(stage));
              {
                if (!(MNRF_DBG(P, 2))) {
                tc_mma2(d_tmem, ah, wb, accumulate);
                tc_mma2(d_tmem, ah + 256u, wb + 256u, 1u);
                tc_mma2(d_tmem, ah + 512u, wb + 512u, 1u);
                tc_mma2(d_tmem, ah + 768u, wb + 768u, 1u);
                }
                tc_commit(bar(BAR_W_EMPTY + stage));
                if (release && c >= 0) tc_commit(bar(BAR_AFREE + c));
                if (last) tc_commit(acc_done);
              }
              accumulate = 1u;
              if (++stage == NST) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
    }
  } else if (warp >= 4) {
    // =========================== epilogue / PE warps ===========================
    const int ew = warp - 4;
    const int q = ew & 3;          // TMEM lane quarter == warp_id % 4
    const int g = ew >> 2;         // 64-column group inside a 128-column half
    const int row = q * 32 + lane;
    const uint32_t rowoff = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const float* F = P.f32;
    const float4* headw = reinterpret_cast<const float4*>(F + P.headw);
    float4* part = reinterpret_cast<float4*>(smem + SM_PART);
    uint32_t acc_phase = 0;  // one parity bit per accumulator barrier
    TraceCtx trc{0};
    uint32_t free_phase = 0; // one parity bit per A-chunk release barrier
    auto wait_acc = [&](int s, int h) {
      const int b = acc_bar(s, h);
      mbar_wait(bar(BAR_ACC + b), (acc_phase >> b) & 1u);
      acc_phase ^= 1u << b;
      tc_fence_after();
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 10, s, h);
    };
    auto a_ready = [&](int c) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_A + c));
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 11, c, 0);
    };
    // xyz + positional encoding of this thread's row of `tile` -> PE operand buffer
    auto pe_tile = [&](int tile) {
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 12, 0, 0);
      const long long pr = (long long)tile * TILE_M + row;
      const long long p = pr < P.io.n_points ? pr : (long long)P.io.n_points - 1;
      float x[3];
      if (P.io.rays != nullptr) {
        const float* rr = P.io.rays + (p / P.io.S) * 8;
        const float z = __ldg(P.io.z + p);
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(__ldg(rr + c), __fmul_rn(__ldg(rr + 3 + c), z));
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = __ldg(P.io.x + p * P.io.x_stride + c);
      }
      if (g == 0) pe_fill<0, PREC3>(x, sbase + SM_PE_HI, sbase + SM_PE_LO, rowoff);
      else        pe_fill<1, PREC3>(x, sbase + SM_PE_HI, sbase + SM_PE_LO, rowoff);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_PE));
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 13, 0, 0);
    };

    if ((int)blockIdx.x < P.n_tiles) pe_tile(blockIdx.x);
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      const long long p_raw = (long long)tile * TILE_M + row;
      const bool valid = p_raw < P.io.n_points;
      const long long p = valid ? p_raw : (long long)P.io.n_points - 1;
      const long long ray = (P.io.rays != nullptr) ? p / P.io.S : p;
      float o_sigma = 0.f, o_n[3] = {0.f, 0.f, 0.f}, o_mirror = 0.f, o_rgb[3] = {0.f, 0.f, 0.f};
      float d[4] = {0.f, 0.f, 0.f, 0.f};

      // ---- trunk layers 1..8 (steps 0..7) ----
      for (int s = 0; s < 8; ++s) {
        const float* bias = F + P.b_trunk[s] + g * 64;
        const float inv = __ldg(F + P.inv_scale + s);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          wait_acc(s, h);
          const int c = 2 * h + g;  // 64-column chunk of the next layer's K produced by this warp
          const uint32_t tacc = tlane + acc_col(s, h) + (uint32_t)g * 64u;
          const uint32_t s_hi = sbase + SM_A_HI + (uint32_t)c * 16384u + rowoff;
          const uint32_t s_lo = sbase + SM_A_LO + (uint32_t)c * 16384u + rowoff;
          // steps >= 1 overwrite the activations their own second-half MMAs may still be reading
          const uint32_t fb = s >= 1 ? bar(BAR_AFREE + c) : 0u;
          const uint32_t fp = (free_phase >> c) & 1u;
          if (MNRF_DBG(P, 4)) {
            if (s >= 1 && !(s == 7 && P.io.sigma_only)) free_phase ^= 1u << c;
            if (!(s == 7 && P.io.sigma_only)) a_ready(c);
          } else if (s < 7) {
            epi64<true, false, true, PREC3>(tacc, bias + h * 128, inv, s_hi, s_lo, nullptr, d, fb, fp);
            if (s >= 1) free_phase ^= 1u << c;
            a_ready(c);
          } else if (!P.io.sigma_only) {
            epi64<true, true, true, PREC3>(tacc, bias + h * 128, inv, s_hi, s_lo, headw + c * 64, d, fb, fp);
            free_phase ^= 1u << c;
            a_ready(c);
          } else {
            epi64<true, true, false, PREC3>(tacc, bias + h * 128, inv, s_hi, s_lo, headw + c * 64, d, 0u, 0u);
          }
        }
        // the PE buffer is free once layer 5's MMAs are done: encode the next tile while the tensor pipe is busy
        if (s == 5 && tile + (int)gridDim.x < P.n_tiles) pe_tile(tile + gridDim.x);
      }
      // combine the two column groups' partial dot products (sigma + folded normal head)
      {
        if (g == 1) part[row] = make_float4(d[0], d[1], d[2], d[3]);
        epi_bar_sync(1);
        if (g == 0) {
          const float4 o = part[row];
          const float4 hb = __ldg(reinterpret_cast<const float4*>(F + P.headb));
          o_sigma = d[0] + o.x + hb.x;
          if (P.has_normal) {
            const float a = d[1] + o.y + hb.y, b = d[2] + o.z + hb.z, cc = d[3] + o.w + hb.w;
            const float nn = sqrtf(fmaxf(a * a + b * b + cc * cc, FP32_EPS));  // utils/func.py:5-7
            o_n[0] = a / nn; o_n[1] = b / nn; o_n[2] = cc / nn;
          }
        }
        epi_bar_sync(2);
      }

      if (!P.io.sigma_only) {
        // ---- mirror head (step 9): LeakyReLU(0.01) -> Linear(128,1) -> sigmoid (mirror_nerf.py:94-99) ----
        if (P.has_mirror) {
          wait_acc(9, 0);
          const float inv = __ldg(F + P.inv_scale + 9);
          float dm = 0.f;
          uint32_t ra[32], rb[32];
          tmem_ld32(tlane + acc_col(9, 0) + (uint32_t)g * 64u, ra);
          tmem_ld32(tlane + acc_col(9, 0) + (uint32_t)g * 64u + 32u, rb);
          tmem_wait_ld();
          pin32(ra);
          pin32(rb);
          const float* bm = F + P.b_m0 + g * 64;
          const float* wm = F + P.w_m2 + g * 64;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = fmaf(__uint_as_float(ra[i]), inv, __ldg(bm + i));
            v = v > 0.f ? v : 0.01f * v;
            dm = fmaf(v, __ldg(wm + i), dm);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = fmaf(__uint_as_float(rb[i]), inv, __ldg(bm + 32 + i));
            v = v > 0.f ? v : 0.01f * v;
            dm = fmaf(v, __ldg(wm + 32 + i), dm);
          }
          if (g == 1) part[row].x = dm;
          epi_bar_sync(1);
          if (g == 0) o_mirror = sigmoidf_(dm + part[row].x + __ldg(F + P.b_m2));
          epi_bar_sync(2);
        }
        // ---- final linear (step 8): f = W h8 + b, written over h8 chunk by chunk as its readers (steps 9, 8) finish ----
        {
          const float inv = __ldg(F + P.inv_scale + 8);
          const float* bias = F + P.b_final + g * 64;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            wait_acc(8, h);
            const int c = 2 * h + g;
            epi64<false, false, true, PREC3>(tlane + acc_col(8, h) + (uint32_t)g * 64u, bias + h * 128, inv,
                                             sbase + SM_A_HI + (uint32_t)c * 16384u + rowoff,
                                             sbase + SM_A_LO + (uint32_t)c * 16384u + rowoff, nullptr, d,
                                             bar(BAR_AFREE + c), (free_phase >> c) & 1u);
            free_phase ^= 1u << c;
            a_ready(c);
          }
        }
        // ---- dir layer (step 10): relu(W_f f + [b + W_d embed(dir)]) -> rgb (mirror_nerf.py:199-204) ----
        wait_acc(10, 0);
        {
          const float inv = __ldg(F + P.inv_scale + 10);
          const float* db = P.io.dirbias + ray * WH + g * 64;
          const float* wr = F + P.w_rgb + g * 64;
          float d0 = 0.f, d1 = 0.f, d2 = 0.f;
          uint32_t ra[32], rb[32];
          tmem_ld32(tlane + acc_col(10, 0) + (uint32_t)g * 64u, ra);
          tmem_ld32(tlane + acc_col(10, 0) + (uint32_t)g * 64u + 32u, rb);
          tmem_wait_ld();
          pin32(ra);
          pin32(rb);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = fmaxf(fmaf(__uint_as_float(ra[i]), inv, __ldg(db + i)), 0.f);
            d0 = fmaf(v, __ldg(wr + i), d0);
            d1 = fmaf(v, __ldg(wr + WH + i), d1);
            d2 = fmaf(v, __ldg(wr + 2 * WH + i), d2);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = fmaxf(fmaf(__uint_as_float(rb[i]), inv, __ldg(db + 32 + i)), 0.f);
            d0 = fmaf(v, __ldg(wr + 32 + i), d0);
            d1 = fmaf(v, __ldg(wr + WH + 32 + i), d1);
            d2 = fmaf(v, __ldg(wr + 2 * WH + 32 + i), d2);
          }
          if (g == 1) part[row] = make_float4(d0, d1, d2, 0.f);
          epi_bar_sync(1);
          if (g == 0) {
            const float4 o = part[row];
            o_rgb[0] = sigmoidf_(d0 + o.x + __ldg(F + P.b_rgb + 0));
            o_rgb[1] = sigmoidf_(d1 + o.y + __ldg(F + P.b_rgb + 1));
            o_rgb[2] = sigmoidf_(d2 + o.z + __ldg(F + P.b_rgb + 2));
          }
          epi_bar_sync(2);
        }
      }

      // ---- write the point record ----
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 14, 0, 0);
      tc_fence_before();
      if (g == 0 && valid) {
        if (P.io.sigma_out != nullptr) P.io.sigma_out[p_raw] = o_sigma;
        if (P.io.raw != nullptr) {
          float4* o = reinterpret_cast<float4*>(P.io.raw + p_raw * 8);
          o[0] = make_float4(o_sigma, o_rgb[0], o_rgb[1], o_rgb[2]);
          o[1] = make_float4(o_mirror, o_n[0], o_n[1], o_n[2]);
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == WARP_MMA) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

static unsigned long long* g_trace_buf = nullptr;
static unsigned int g_trace_cap = 0;
void set_tc_trace(unsigned long long* buf, unsigned int cap) { g_trace_buf = buf; g_trace_cap = cap; }

int launch_field_tc(const mnrf_field* f, const FieldIO& io, int precision, cudaStream_t st) {
  if (io.n_points <= 0) return 0;
  MNRF_REQUIRE(precision == 1 || precision == 3, "field_tc: precision must be 1 or 3");
  MNRF_REQUIRE(io.normal_out == nullptr, "field_tc: analytic normals need MNRF_IMPL_FP32");
  MNRF_REQUIRE(io.geo_out == nullptr, "field_tc: geo_feat output needs MNRF_IMPL_FP32");
  MNRF_REQUIRE(io.sigma_only || io.dirbias != nullptr, "field_tc: dirbias missing");
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    MNRF_CUDA_OK(cudaGetDevice(&dev));
    MNRF_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
  }
  TcParams P;
  const F32Layout& L = f->L;
  P.f32 = f->f32;
  P.tc = f->tc;
  for (int l = 0; l < 8; ++l) P.b_trunk[l] = L.b_trunk[l];
  P.b_final = L.b_final; P.b_m0 = L.b_m0; P.w_m2 = L.w_m2; P.b_m2 = L.b_m2;
  P.w_rgb = L.w_rgb; P.b_rgb = L.b_rgb; P.headw = L.headw; P.headb = L.headb; P.inv_scale = L.inv_scale;
  P.has_normal = f->has_normal; P.has_mirror = f->has_mirror;
  P.io = io;
  P.trace = g_trace_buf;
  P.trace_cap = g_trace_cap;
  const char* dbg = getenv("MNRF_TC_DEBUG");
  P.debug = dbg != nullptr ? atoi(dbg) : 0;
  P.n_tiles = (io.n_points + TILE_M - 1) / TILE_M;
  const int grid = P.n_tiles < num_sms ? P.n_tiles : num_sms;
  // algorithmic MACs of this launch (unpadded reference layer sizes)
  const double macs = (double)io.n_points * (io.sigma_only ? (double)mnrf_macs_sigma_only() : (double)mnrf_macs_full());
  prof_begin(st);
  if (precision == 3) k_field_tc<true><<<grid, NUM_THREADS, SM_TOTAL, st>>>(P);
  else                k_field_tc<false><<<grid, NUM_THREADS, SM_TOTAL, st>>>(P);
  prof_end(st, 2.0 * macs);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace mnrf
