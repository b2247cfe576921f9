// tcgen05 field kernel: MirrorNeRF.forward (R/models/mirror_nerf.py:101-212) fused per 128-point tile:
//   o + d*z  ->  positional encoding  ->  8x256 trunk (skip at layer 5)  ->  sigma / folded normal head
//   ->  final 256x256  ->  mirror head  ->  dir layer (+ per-ray dir term)  ->  rgb      -> 8 floats/point.
//
// One persistent CTA per SM, 12 warps:
//   warp 0        weight producer: cp.async.bulk (TMA 1-D) of pre-packed B-operand blobs, 4-stage mbarrier ring
//   warp 1        MMA issuer: one thread issues tcgen05.mma (M=128, N=256|128, K=16, fp16 -> fp32 in TMEM)
//   warps 4..11   epilogue/PE: TMEM -> registers (tcgen05.ld) -> bias/ReLU -> fp16 hi/lo split -> next layer's
//                 A operand written in place into shared memory in the UMMA K-major core-matrix layout;
//                 per-64-column "chunk ready" mbarriers let the next layer's MMAs start while the rest of the
//                 epilogue is still running (two 256-column TMEM accumulators alternate by layer).
//
// Precision: operands are split x = hi + lo in fp16 (weights pre-scaled by 2^s per layer) and
// each product is formed as hi*hi + lo*hi + hi*lo with fp32 accumulation ("3x" mode, fp32-grade);
// precision == 1 drops the lo terms (speed mode).
//
// Shared memory (1 CTA/SM): A hi/lo 2x64 KB | PE hi/lo 2x16 KB | 4 x 16 KB weight stages | 2 KB partial sums.
#include "common.cuh"

namespace mnrf {
namespace {

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 384;
constexpr int NUM_WSTAGES = 4;
constexpr int WSTAGE_BYTES = 16384;

constexpr uint32_t SM_A_HI = 0;
constexpr uint32_t SM_A_LO = 65536;
constexpr uint32_t SM_PE_HI = 131072;
constexpr uint32_t SM_PE_LO = 147456;
constexpr uint32_t SM_WST = 163840;
constexpr uint32_t SM_PART = SM_WST + NUM_WSTAGES * WSTAGE_BYTES;  // 229376: float4[128]
constexpr uint32_t SM_BAR = SM_PART + 2048;                        // 231424
constexpr uint32_t SM_TOTAL = SM_BAR + 256;                        // 231680 <= 232448

// barrier slots (8 bytes each)
constexpr int BAR_W_FULL = 0;    // [4]
constexpr int BAR_W_EMPTY = 4;   // [4]
constexpr int BAR_PE = 8;        // PE chunk written (8 warp arrivals)
constexpr int BAR_A = 9;         // [4] A 64-column chunk written (4 warp arrivals)
constexpr int BAR_ACC = 13;      // [2] GEMM step complete (tcgen05.commit)
constexpr int BAR_TMEM_SLOT = 15;

constexpr uint32_t IDESC_N256 = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);  // fp16 x fp16 -> fp32, K-major
constexpr uint32_t IDESC_N128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

struct TcParams {
  const float* f32;        // fp32 section
  const uint8_t* tc;       // packed blobs
  int b_trunk[8];
  int b_final, b_m0, w_m2, b_m2, w_rgb, b_rgb, headw, headb, inv_scale;
  int has_normal, has_mirror;
  int precision;           // 1 | 3
  int desc_swap;           // debug: swap LBO/SBO roles
  FieldIO io;
  int n_tiles;
};

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
  long long t0 = clock64();
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void epi_bar_sync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }

// K-major, no-swizzle operand descriptor.  Core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes;
// 8-row groups (M/N direction) are `mn_stride` bytes apart, K-adjacent core matrices `k_stride` bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t k_stride, uint32_t mn_stride, int swap) {
  uint32_t lbo = swap ? mn_stride : k_stride;
  uint32_t sbo = swap ? k_stride : mn_stride;
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// x = hi + lo in fp16; two values packed per 32-bit word
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(a, b);
  float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ float clampf16(float v) { return fminf(fmaxf(v, -60000.f), 60000.f); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// store 8 consecutive K values (one 16-byte core-matrix row) of an A-type operand, hi and lo parts
__device__ __forceinline__ void store_a8(uint32_t hi_addr, uint32_t lo_addr, const float (&v)[8], bool with_lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  st_shared_v4(hi_addr, h[0], h[1], h[2], h[3]);
  if (with_lo) st_shared_v4(lo_addr, l[0], l[1], l[2], l[3]);
}

// ---- positional encoding of one row, K range [32*HALF, 32*HALF+32) (mirror_nerf.py:33-38) ----------------
template <int HALF>
__device__ __forceinline__ void pe_fill(const float (&x)[3], uint32_t pe_hi, uint32_t pe_lo, uint32_t rowoff,
                                        bool with_lo) {
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
        float s, co;
        sincosf(ldexpf(x[c], f), &s, &co);  // accurate path; 2^f * x is exact
        if (need_s) vals[ks - K0] = s;
        if (need_c) vals[kc - K0] = co;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = vals[8 * j + i];
    uint32_t off = (uint32_t)(4 * HALF + j) * 2048u + rowoff;
    store_a8(pe_hi + off, pe_lo + off, v, with_lo);
  }
}

// ---- step geometry --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t acc_col(int s) { return s <= 7 ? (uint32_t)(s & 1) * 256u : (s == 8 ? 0u : (s == 9 ? 256u : 384u)); }

// ================================================================================================
__global__ void __launch_bounds__(NUM_THREADS, 1) k_field_tc(const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bars = sbase + SM_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * BAR_TMEM_SLOT);
  const bool prec3 = P.precision == 3;
  const int last_step = P.io.sigma_only ? 7 : 10;

  if (threadIdx.x == 0) {
    if (sbase & 127u) { printf("mnrf field_tc: unaligned dynamic smem base %u\n", sbase); __trap(); }
    for (int i = 0; i < NUM_WSTAGES; ++i) { mbar_init(bar(BAR_W_FULL + i), 1); mbar_init(bar(BAR_W_EMPTY + i), 1); }
    mbar_init(bar(BAR_PE), 8);
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_A + i), 4);
    mbar_init(bar(BAR_ACC + 0), 1);
    mbar_init(bar(BAR_ACC + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
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

  if (warp == 0) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        for (int s = 0; s <= last_step; ++s) {
          if (s == 9 && !P.has_mirror) continue;
          const uint8_t* src = P.tc + tc_step_offset(s);
          const uint32_t blob = (uint32_t)tc_blob_bytes(s);
          const int nch = tc_step_chunks(s);
          for (int kc = 0; kc < nch; ++kc) {
            for (int part = 0; part < (prec3 ? 2 : 1); ++part) {
              mbar_wait(bar(BAR_W_EMPTY + stage), phase ^ 1u);
              mbar_expect_tx(bar(BAR_W_FULL + stage), blob);
              bulk_g2s(sbase + SM_WST + stage * WSTAGE_BYTES, src + (size_t)(2 * kc + part) * blob, blob,
                       bar(BAR_W_FULL + stage));
              if (++stage == NUM_WSTAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t pe_phase = 0, a_phase[4] = {0, 0, 0, 0};
      const int sw = P.desc_swap;
      for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        for (int s = 0; s <= last_step; ++s) {
          if (s == 9 && !P.has_mirror) continue;
          const int N = tc_step_n(s);
          const uint32_t idesc = (N == 256) ? IDESC_N256 : IDESC_N128;
          const uint32_t d_tmem = tmem + acc_col(s);
          const int nch = tc_step_chunks(s);
          const int n_pe = (s == 0 || s == 4) ? 2 : 0;  // leading K32 chunks that come from the PE buffer
          uint32_t accumulate = 0;
          for (int kc = 0; kc < nch; ++kc) {
            uint32_t a_hi, a_lo;
            if (kc < n_pe) {
              if (s == 0 && kc == 0) { mbar_wait(bar(BAR_PE), pe_phase); pe_phase ^= 1u; tc_fence_after(); }
              a_hi = sbase + SM_PE_HI + (uint32_t)kc * 8192u;
              a_lo = sbase + SM_PE_LO + (uint32_t)kc * 8192u;
            } else {
              const int ka = kc - n_pe;  // K32 chunk inside the A buffer
              if ((ka & 1) == 0 && s != 9) {  // first touch of a 64-column chunk of a new activation version
                const int c = ka >> 1;
                mbar_wait(bar(BAR_A + c), a_phase[c]); a_phase[c] ^= 1u; tc_fence_after();
              }
              a_hi = sbase + SM_A_HI + (uint32_t)ka * 8192u;
              a_lo = sbase + SM_A_LO + (uint32_t)ka * 8192u;
            }
            // ---- hi weights: A_hi*W_hi (+ A_lo*W_hi) ----
            mbar_wait(bar(BAR_W_FULL + stage), phase);
            tc_fence_after();
            {
              const uint32_t b0 = sbase + SM_WST + stage * WSTAGE_BYTES;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint64_t bd = make_desc(b0 + (uint32_t)j * (uint32_t)N * 32u, (uint32_t)N * 16u, 128u, sw);
                tc_mma(d_tmem, make_desc(a_hi + (uint32_t)j * 4096u, 2048u, 128u, sw), bd, idesc, accumulate);
                accumulate = 1;
                if (prec3) tc_mma(d_tmem, make_desc(a_lo + (uint32_t)j * 4096u, 2048u, 128u, sw), bd, idesc, 1);
              }
            }
            tc_commit(bar(BAR_W_EMPTY + stage));
            if (++stage == NUM_WSTAGES) { stage = 0; phase ^= 1u; }
            // ---- lo weights: A_hi*W_lo ----
            if (prec3) {
              mbar_wait(bar(BAR_W_FULL + stage), phase);
              tc_fence_after();
              const uint32_t b0 = sbase + SM_WST + stage * WSTAGE_BYTES;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint64_t bd = make_desc(b0 + (uint32_t)j * (uint32_t)N * 32u, (uint32_t)N * 16u, 128u, sw);
                tc_mma(d_tmem, make_desc(a_hi + (uint32_t)j * 4096u, 2048u, 128u, sw), bd, idesc, 1);
              }
              tc_commit(bar(BAR_W_EMPTY + stage));
              if (++stage == NUM_WSTAGES) { stage = 0; phase ^= 1u; }
            }
          }
          tc_commit(bar(BAR_ACC + (s & 1)));
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue / PE warps ===========================
    const int ew = warp - 4;
    const int q = ew & 3;          // TMEM lane quarter == warp_id % 4
    const int g = ew >> 2;         // column group
    const int row = q * 32 + lane;
    const uint32_t rowoff = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const float* F = P.f32;
    const float4* headw = reinterpret_cast<const float4*>(F + P.headw);
    float4* part = reinterpret_cast<float4*>(smem + SM_PART);
    uint32_t acc_phase[2] = {0, 0};
    auto wait_acc = [&](int s) { mbar_wait(bar(BAR_ACC + (s & 1)), acc_phase[s & 1]); acc_phase[s & 1] ^= 1u; tc_fence_after(); };

    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      const long long p_raw = (long long)tile * TILE_M + row;
      const bool valid = p_raw < P.io.n_points;
      const long long p = valid ? p_raw : (long long)P.io.n_points - 1;
      const long long ray = (P.io.rays != nullptr) ? p / P.io.S : p;

      // ---- xyz + positional encoding -> PE operand buffer ----
      {
        float x[3];
        if (P.io.rays != nullptr) {
          const float* rr = P.io.rays + ray * 8;
          const float z = __ldg(P.io.z + p);
#pragma unroll
          for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(__ldg(rr + c), __fmul_rn(__ldg(rr + 3 + c), z));
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) x[c] = __ldg(P.io.x + p * P.io.x_stride + c);
        }
        if (g == 0) pe_fill<0>(x, sbase + SM_PE_HI, sbase + SM_PE_LO, rowoff, prec3);
        else        pe_fill<1>(x, sbase + SM_PE_HI, sbase + SM_PE_LO, rowoff, prec3);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_PE));
      }

      float o_sigma = 0.f, o_n[3] = {0.f, 0.f, 0.f}, o_mirror = 0.f, o_rgb[3] = {0.f, 0.f, 0.f};

      // ---- trunk layers 1..8 (steps 0..7) and the final linear (step 8) ----
      for (int s = 0; s <= (P.io.sigma_only ? 7 : 8); ++s) {
        wait_acc(s);
        if (s == 8 && P.has_mirror) wait_acc(9);  // h8 (A buffer) is still being read by the mirror GEMM
        const bool relu = s < 8;
        const bool write_a = !(P.io.sigma_only && s == 7);
        const bool dots = s == 7;
        const float* bias = F + (s < 8 ? P.b_trunk[s] : P.b_final);
        const float inv = __ldg(F + P.inv_scale + s);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        for (int ci = 0; ci < 2; ++ci) {
          const int c = g + 2 * ci;  // 64-column chunk owned by this warp group
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            const int col0 = c * 64 + sub * 32;
            uint32_t r[32];
            tmem_ld32(tlane + acc_col(s) + (uint32_t)col0, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + 8 * j));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + 8 * j + 4));
              float v[8];
              v[0] = fmaf(__uint_as_float(r[8 * j + 0]), inv, b0.x);
              v[1] = fmaf(__uint_as_float(r[8 * j + 1]), inv, b0.y);
              v[2] = fmaf(__uint_as_float(r[8 * j + 2]), inv, b0.z);
              v[3] = fmaf(__uint_as_float(r[8 * j + 3]), inv, b0.w);
              v[4] = fmaf(__uint_as_float(r[8 * j + 4]), inv, b1.x);
              v[5] = fmaf(__uint_as_float(r[8 * j + 5]), inv, b1.y);
              v[6] = fmaf(__uint_as_float(r[8 * j + 6]), inv, b1.z);
              v[7] = fmaf(__uint_as_float(r[8 * j + 7]), inv, b1.w);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = relu ? fminf(fmaxf(v[i], 0.f), 60000.f) : clampf16(v[i]);
              if (dots) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 hw = __ldg(headw + col0 + 8 * j + i);
                  d0 = fmaf(v[i], hw.x, d0); d1 = fmaf(v[i], hw.y, d1);
                  d2 = fmaf(v[i], hw.z, d2); d3 = fmaf(v[i], hw.w, d3);
                }
              }
              if (write_a) {
                const uint32_t off = (uint32_t)((col0 >> 3) + j) * 2048u + rowoff;
                store_a8(sbase + SM_A_HI + off, sbase + SM_A_LO + off, v, prec3);
              }
            }
          }
          if (write_a) {
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_A + c));
          }
        }
        if (dots) {
          // combine the two column groups' partial dot products (sigma + folded normal head)
          if (g == 1) part[row] = make_float4(d0, d1, d2, d3);
          epi_bar_sync(1);
          if (g == 0) {
            const float4 o = part[row];
            const float4 hb = __ldg(reinterpret_cast<const float4*>(F + P.headb));
            o_sigma = d0 + o.x + hb.x;
            if (P.has_normal) {
              float a = d1 + o.y + hb.y, b = d2 + o.z + hb.z, cc = d3 + o.w + hb.w;
              float nn = sqrtf(fmaxf(a * a + b * b + cc * cc, FP32_EPS));  // utils/func.py:5-7
              o_n[0] = a / nn; o_n[1] = b / nn; o_n[2] = cc / nn;
            }
          }
          epi_bar_sync(2);
        }
      }

      if (!P.io.sigma_only) {
        // ---- mirror head (step 9): LeakyReLU(0.01) -> Linear(128,1) -> sigmoid (mirror_nerf.py:94-99) ----
        if (P.has_mirror) {
          const float inv = __ldg(F + P.inv_scale + 9);
          float d = 0.f;
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            const int col0 = g * 64 + sub * 32;
            uint32_t r[32];
            tmem_ld32(tlane + acc_col(9) + (uint32_t)col0, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float v = fmaf(__uint_as_float(r[i]), inv, __ldg(F + P.b_m0 + col0 + i));
              v = v > 0.f ? v : 0.01f * v;
              d = fmaf(v, __ldg(F + P.w_m2 + col0 + i), d);
            }
          }
          if (g == 1) part[row].x = d;
          epi_bar_sync(1);
          if (g == 0) o_mirror = sigmoidf_(d + part[row].x + __ldg(F + P.b_m2));
          epi_bar_sync(2);
        }
        // ---- dir layer (step 10): relu(W_f f + [b + W_d embed(dir)]) -> rgb (mirror_nerf.py:199-204) ----
        wait_acc(10);
        {
          const float inv = __ldg(F + P.inv_scale + 10);
          const float* db = P.io.dirbias + ray * WH;
          float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            const int col0 = g * 64 + sub * 32;
            uint32_t r[32];
            tmem_ld32(tlane + acc_col(10) + (uint32_t)col0, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float v = fmaxf(fmaf(__uint_as_float(r[i]), inv, __ldg(db + col0 + i)), 0.f);
              d0 = fmaf(v, __ldg(F + P.w_rgb + col0 + i), d0);
              d1 = fmaf(v, __ldg(F + P.w_rgb + WH + col0 + i), d1);
              d2 = fmaf(v, __ldg(F + P.w_rgb + 2 * WH + col0 + i), d2);
            }
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
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

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
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
  }
  TcParams P;
  const F32Layout& L = f->L;
  P.f32 = f->f32;
  P.tc = f->tc;
  for (int l = 0; l < 8; ++l) P.b_trunk[l] = L.b_trunk[l];
  P.b_final = L.b_final; P.b_m0 = L.b_m0; P.w_m2 = L.w_m2; P.b_m2 = L.b_m2;
  P.w_rgb = L.w_rgb; P.b_rgb = L.b_rgb; P.headw = L.headw; P.headb = L.headb; P.inv_scale = L.inv_scale;
  P.has_normal = f->has_normal; P.has_mirror = f->has_mirror;
  P.precision = precision;
  const char* sw = getenv("MNRF_TC_DESC_SWAP");
  P.desc_swap = (sw != nullptr && sw[0] == '1') ? 1 : 0;
  P.io = io;
  P.n_tiles = (io.n_points + TILE_M - 1) / TILE_M;
  int grid = P.n_tiles < num_sms ? P.n_tiles : num_sms;
  // algorithmic MACs of this launch (unpadded reference layer sizes)
  const double macs = (double)io.n_points * (io.sigma_only ? (double)mnrf_macs_sigma_only() : (double)mnrf_macs_full());
  prof_begin(st);
  k_field_tc<<<grid, NUM_THREADS, SM_TOTAL, st>>>(P);
  prof_end(st, 2.0 * macs);
  MNRF_LAUNCH_OK();
  return 0;
}

}  // namespace mnrf
