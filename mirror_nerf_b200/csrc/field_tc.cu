// tcgen05 field kernel (v3): MirrorNeRF.forward (R/models/mirror_nerf.py:101-212) fused per 128-point tile:
//   o + d*z  ->  positional encoding  ->  8x256 trunk (skip at layer 5)  ->  sigma / folded normal head
//   ->  mirror head  ->  final 256x256  ->  dir layer (+ per-ray dir term)  ->  rgb      -> 8 floats/point.
//
// One persistent CTA per SM, 20 warps:
//   warps 0..15   epilogue/PE (TMEM lane quarter = warp % 4, column group = warp / 4): TMEM -> registers (tcgen05.ld) ->
//                 bias/ReLU -> operand split -> next layer's A operand written in place into shared memory in the UMMA K-major
//                 core-matrix layout; heads, positional encoding of the next tile, and (fused mode) compositing
//   warp 16       weight producer: cp.async.bulk (TMA 1-D) of pre-packed B-operand blobs, 16 KB mbarrier-ring stages
//   warp 17       MMA issuer: ONE elected thread runs the whole role and issues tcgen05.mma
//                 (M=128, N=256 | 128, K=16 | 32, fp16 / e4m3 operands in shared memory, fp32 accumulators in TMEM)
//
// Pipelining.  Measured on B200 (tools/mma_bench*.cu, profiles/): the tensor pipe needs 128 cycles per M128xN256xK16 and 64
// per N128; the 256-wide layers are issued as N=256 instructions, one accumulator barrier per layer, two 256-column TMEM
// accumulators alternating by layer.  The epilogue of layer l starts when its accumulator is complete (every reader of the
// in-place activations is done by then) and signals per-chunk mbarriers, so layer l+1's MMAs start after the first chunk.
// An optional N-split schedule (issue_split_step) issues every 256-wide layer as two 128-column halves so that the first
// half's epilogue overlaps the second half's MMAs.  Measured neutral (profiles/r02_v5_*): while the sixteen epilogue warps
// convert they saturate all four schedulers, and the two single-thread roles that share those schedulers then get an issue
// slot every ~45 cycles (device timeline) -- tensor pipe and epilogue time add up whichever way they are ordered.
//
// Precision: operands are split x = hi + lo in fp16 (weights pre-scaled by 2^s per layer) and
// each product is formed as hi*hi + lo*hi + hi*lo with fp32 accumulation ("3x" mode, fp32-grade);
// PREC == 1 drops the lo terms (speed mode).
// PREC == 2 ("tc2"): the two cross terms need only a few bits of their own, so they run on the fp8 datapath at twice the fp16
// rate: D = A_hi W_hi (fp16, K16) + e4m3(2^10 A_lo) e4m3(2^-10 W_hi) + e4m3(A_hi) e4m3(W_lo) (kind::f8f6f4, K32), all three
// accumulated into the same fp32 TMEM accumulator (measured: the accumulator keeps full fp32 precision across kinds,
// profiles/r02_v1_fp8_mma_microbench.txt) -- 2.0 fp16-pass equivalents per algorithmic MAC instead of 3, operand error ~2^-16.
//
// Shared memory (1 CTA/SM): A hi/lo 2x64 KB | PE hi/lo 2x16 KB | 4 x 16 KB weight stages | 2 KB partial sums.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mnrf {
namespace {
using namespace tcx;

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 640;
// warps 0..15: epilogue/PE (TMEM lane quarter = warp % 4, column group = warp / 4); warp 16: weight producer; warp 17: MMA issuer.
// The two single-thread roles sit on the HIGHEST warp ids of their schedulers: they share the issue slots with four epilogue
// warps each, and under that contention the arbiter serves higher warp ids first (tools/mma_bench4.cu: an ALU-saturated
// scheduler slows the issuer 2.1x with the roles on warps 0/1 and 1.5x on warps 16/17; on the bench +0.6..1.1 %).
// MNRF_TC_ROLES_LOW restores the old placement for A/B measurements.
#ifdef MNRF_TC_ROLES_LOW
constexpr int EPI_WARP0 = 4;
constexpr int WARP_PRODUCER = 0;
constexpr int WARP_MMA = 1;
#else
constexpr int EPI_WARP0 = 0;
constexpr int WARP_PRODUCER = 16;
constexpr int WARP_MMA = 17;
#endif
constexpr uint32_t WSTAGE_BYTES = 16384;

constexpr uint32_t SM_A_HI = 0;
constexpr uint32_t SM_A_LO = 65536;
constexpr uint32_t SM_PE_HI = 131072;
constexpr uint32_t SM_PE_LO = 147456;
constexpr uint32_t SM_WST = 163840;
constexpr uint32_t SM_PART = SM_WST + 4 * WSTAGE_BYTES;  // 229376: float4[128]
constexpr uint32_t SM_BAR = SM_PART + 2048;              // 231424
constexpr uint32_t SM_FUSE = SM_BAR + 256;               // 231680: work descriptors of the fused (ray-tile) mode, 256 B
constexpr uint32_t SM_TOTAL = SM_FUSE + 256;             // 231936 <= 232448

// barrier slots (8 bytes each)
constexpr int BAR_W_FULL = 0;    // [8]
constexpr int BAR_W_EMPTY = 8;   // [8]
constexpr int BAR_PE = 16;       // PE chunk written (16 warp arrivals)
constexpr int BAR_A = 17;        // [5] A columns written (4 warp arrivals: one column group): [0] cols 0-31, [4] cols 32-63, [1..3] 64-col chunks 1..3
constexpr int BAR_ACC = 22;      // [4] accumulator of a GEMM step complete (tcgen05.commit)
constexpr int BAR_GO = 26;       // fused mode: "next tile is decided" for the weight producer (8 warp arrivals, like BAR_PE)
constexpr int BAR_ACCH = 27;     // N-split schedule: the first 128 accumulator columns of a layer are complete (tcgen05.commit)
constexpr int BAR_AFREE = 28;    // [2] N-split schedule: the layer's MMAs have read A chunk 0 / 1 for the last time (tcgen05.commit)
constexpr int BAR_TMEM_SLOT = 30;

constexpr uint32_t IDESC_N256 = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, K-major
constexpr uint32_t IDESC_N128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_N64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

struct TcParams {
  const float* f32;        // fp32 section
  const uint8_t* tc;       // packed blobs
  const uint8_t* tc_split; // tc2 only: the same step offsets with steps 0..TC_SPLIT_LAST in the N-split layout (pack.cu)
  int split;               // tc2, no analytic normals: run the 256-wide layers as two 128-column halves (see issue_split_step)
  int b_trunk[8];
  int b_final, b_m0, w_m2, b_m2, w_rgb, b_rgb, headw, headb, inv_scale;
  int has_normal, has_mirror;
  FieldIO io;
  int n_tiles;
  int slot;                   // which copy of the epilogue table (c_epi) belongs to this launch's field
  // fused (ray-tile) mode: tile = 4 rays x 32 consecutive samples, compositing in the epilogue registers
  mnrf_composite_out comp;    // per-ray outputs (+ optional per-sample weights / pred_normal)
  int n_rays;                 // rays of the launch (the device count io.n_rays_dev clamps it)
  int white_back;
  float term_eps;             // > 0: stop a ray once its transmittance falls below this (eval, compact outputs only)
  int* work_counter;          // device int, zero at launch: next ray to hand out
  unsigned long long* stats;  // optional device counters: [0] tiles executed, [1] chunks skipped by early termination
  unsigned long long* trace;  // optional device-side event trace of CTA 0 (bring-up builds)
  unsigned int trace_cap;
  int debug;
};

// Epilogue table (biases, head weights, scales: common.cuh ET_*) in constant memory, so that the warp-uniform epilogue reads are
// constant-cache accesses instead of global loads.  The host side (epi_slot_acquire) keys the copy by the field's pack stamp
// and orders a rewrite after the last kernel (on any stream) that read it, so two fields rendered on different streams never
// race on the table.  ONE slot: with a compile-time slot index the biases are immediate constant operands of the epilogue's
// FFMAs; indexing several slots at run time turned every bias read into a register-indexed LDC (measured: 976 LDC in the
// tc2 kernel, 0.5 per element), which costs more than re-copying 16 KB when the coarse and the fine field alternate.
constexpr int EPI_SLOTS = 1;
__constant__ float c_epi_slots[EPI_SLOTS][ET_TOTAL];

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

// trace builds only: cycles the MMA issuer spends blocked on one barrier class, accumulated over a tile (payload of events 5..7)
#ifdef MNRF_TC_TRACE
#define TR_T0() const long long _tr_t0 = clock64()
#define TR_ADD(x) (x) += (unsigned int)(clock64() - _tr_t0)
__device__ __forceinline__ void trace_val(const TcParams& P, TraceCtx& tc, int who, int ev, unsigned int val) {
  if (P.trace != nullptr && blockIdx.x == 0 && tc.n < 8192u) {
    unsigned long long* r = P.trace + 1 + (size_t)who * 2 * 8192 + 2 * tc.n;
    r[0] = (unsigned long long)clock64();
    r[1] = ((unsigned long long)val << 32) | ((unsigned long long)who << 24) | ((unsigned long long)ev << 16);
    ++tc.n;
  }
}
#else
#define TR_T0() do {} while (0)
#define TR_ADD(x) do {} while (0)
__device__ __forceinline__ void trace_val(const TcParams&, TraceCtx&, int, int, unsigned int) {}
#endif

// PTX wrappers (mbarrier, bulk copy, tcgen05 fences / commit / ld / st, descriptors): tc_ptx.cuh, shared with train_tc.cu

// fp8 (e4m3 x e4m3 -> f32, K = 32 per instruction): same descriptor and instruction-descriptor bits as the fp16 form (format 0 is
// F16 for kind::f16 and E4M3 for kind::f8f6f4); K-major core matrices hold 16 K values per 16-byte row
template <int N>
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %4};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(accumulate), "r"(DESC_HI), "r"(N == 256 ? IDESC_N256 : (N == 128 ? IDESC_N128 : IDESC_N64))
      : "memory");
}
template <int N>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %4};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(accumulate), "r"(DESC_HI), "r"(N == 256 ? IDESC_N256 : (N == 128 ? IDESC_N128 : IDESC_N64))
      : "memory");
}

// ---- N-split schedule of a 256-wide layer (tc2) ------------------------------------------------------------------------------
// The layer is issued as two N = 128 GEMMs over the full K: columns [0,128) first, then [128,256), each into its own half of the
// layer's accumulator and with its own completion barrier.  The epilogue of the first half (the next layer's A chunks 0 and 1)
// then runs WHILE the tensor pipe works on the second half, so the next layer's MMAs find half of their operand in place the
// moment this layer's last MMA is issued; the hand-over bubble of the unsplit schedule (accumulator drain -> TMEM load ->
// convert -> proxy fence -> barrier, ~2,000 cycles per layer with the pipe idle) shrinks to the tail of the second half.
// The in-place A operand needs one more ordering: the first half's epilogue overwrites A chunks 0/1 while the second half's
// MMAs still read the OLD operand, so the issuer commits BAR_AFREE[c] right after the second half's last MMAs on chunk c.
// One 16 KB weight stage per (half, K32 chunk) in the N = 128 blob format (pack.cu); every step consumes a multiple of 4 stages,
// so a step starts at ring position 0 and the ring position of every stage is a compile-time constant of the unrolled loop
// (measured, tools/mma_bench4.cu: 282 vs 363 cycles per N = 128 stage against the run-time stage index).  The PE operand of
// step 0 is awaited once in front of the step (WAIT_PE: static tiles; fused tiles wait at the tile start).
struct SplitCtx {
  uint32_t bars;                        // shared address of barrier slot 0
  uint32_t wdesc0;                      // descriptor low word of weight stage 0 (LBO 2048); stage st is + st * 1024
  uint32_t dl_a_hi, dl_a8, dl_a8r;      // A operand: fp16 hi, e4m3 copy, e4m3 residual
  uint32_t dl_pe_hi, dl_pe8, dl_pe8r;   // PE operand
};
template <int NCH, int NPE, bool WAIT_PE, class Tracer>
__device__ __forceinline__ void issue_split_step(const SplitCtx& c, uint32_t d_base, int acc_slot, bool a_reused, uint32_t& phase,
                                                 uint32_t& a_phase, uint32_t& pe_phase, Tracer&& tr) {
  static_assert((2 * NCH) % 4 == 0, "a split step must consume whole trips of the 4-stage ring");
  constexpr int K0 = (NPE + 1 < NCH - 1) ? NPE + 1 : NCH - 1;   // last K32 chunk that reads A chunk 0 / chunk 1 (PE-only
  constexpr int K1 = (NPE + 3 < NCH - 1) ? NPE + 3 : NCH - 1;   // layer: nothing reads the A buffer, signal at the end)
  auto bar = [&](int i) { return c.bars + 8u * (uint32_t)i; };
  if (WAIT_PE) { mbar_spin(bar(BAR_PE), pe_phase); pe_phase ^= 1u; }
  // Fully unrolled: half, K chunk and ring position of every stage are compile-time constants, so each descriptor is a uniform
  // base register plus an immediate -- the issuing thread shares its scheduler with four epilogue warps and its dependent
  // address arithmetic (14 R2UR per stage with run-time indices), not the tensor pipe, set the stage time.
#pragma unroll
  for (int t = 0; t < (2 * NCH) / 4; ++t) {   // one trip around the ring
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const int i = 4 * t + st;
      const int h = i >= NCH ? 1 : 0, kc = i - h * NCH;
      uint32_t ah, a8, a8r;
      if (kc < NPE) {
        ah = c.dl_pe_hi + (uint32_t)kc * 512u; a8 = c.dl_pe8 + (uint32_t)kc * 256u; a8r = c.dl_pe8r + (uint32_t)kc * 256u;
      } else {
        const int ka = kc - NPE;
        if (h == 0 && ((ka & 1) == 0 || ka == 1) && !a_reused) {   // first touch of freshly written activation columns
          const int cb = ka == 1 ? 4 : (ka >> 1);
          mbar_spin(bar(BAR_A + cb), (a_phase >> cb) & 1u);
          a_phase ^= 1u << cb;
        }
        ah = c.dl_a_hi + (uint32_t)ka * 512u; a8 = c.dl_a8 + (uint32_t)ka * 256u; a8r = c.dl_a8r + (uint32_t)ka * 256u;
      }
      mbar_spin(bar(BAR_W_FULL + st), phase);
      tc_fence_after();
      tr(4, i);   // trace builds: this stage's operands are in place, its MMAs are issued now
      const uint32_t wb = c.wdesc0 + (uint32_t)st * 1024u;
      const uint32_t d = d_base + 128u * (uint32_t)h;
      tc_mma<128>(d, ah, wb, kc > 0 ? 1u : 0u);
      tc_mma<128>(d, ah + 256u, wb + 256u, 1u);
      tc_mma_f8<128>(d, a8r, wb + 512u, 1u);     // (2^10 A_lo) * (2^-10 W_hi)
      tc_mma_f8<128>(d, a8, wb + 768u, 1u);      // A_hi * W_lo
      tc_commit(bar(BAR_W_EMPTY + st));
      if (h == 1 && kc == K0) tc_commit(bar(BAR_AFREE + 0));
      if (h == 1 && kc == K1) tc_commit(bar(BAR_AFREE + 1));
      if (kc == NCH - 1) tc_commit(h == 0 ? bar(BAR_ACCH) : bar(BAR_ACC + acc_slot));
    }
    phase ^= 1u;
  }
}

// ---- one 256-wide layer of the unsplit tc2 schedule, fully unrolled -------------------------------------------------------
// Same MMAs in the same order as the generic issue loop of the kernel (per K32 chunk: stage [W_hi fp16] -> two f16 K16 MMAs,
// stage [e4m3(2^-10 W_hi) | e4m3(W_lo)] -> two f8 K32 MMAs), but chunk index and ring position of every stage are compile-time
// constants, so every descriptor is a base register plus an immediate: about half the instructions per stage.  That matters
// where the issuer overlaps the epilogue of the previous layer: there it gets an issue slot only every few tens of cycles
// (device timeline: 350-410 cycles per 256-cycle stage under the epilogue, 260 once the epilogue is done).
template <int NCH, int NPE, bool WAIT_PE, class Tracer>
__device__ __forceinline__ void issue_wide_step(const SplitCtx& c, uint32_t d, int acc_slot, bool a_reused, uint32_t& phase,
                                                uint32_t& a_phase, uint32_t& pe_phase, Tracer&& tr) {
  static_assert((2 * NCH) % 4 == 0, "a 256-wide tc2 step must consume whole trips of the 4-stage ring");
  auto bar = [&](int i) { return c.bars + 8u * (uint32_t)i; };
  if (WAIT_PE) { mbar_spin(bar(BAR_PE), pe_phase); pe_phase ^= 1u; }
#pragma unroll
  for (int t = 0; t < (2 * NCH) / 4; ++t) {   // one trip around the ring = two K32 chunks
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const int i = 4 * t + st, kc = i >> 1, half = i & 1;
      uint32_t ah, a8, a8r;
      if (kc < NPE) {
        ah = c.dl_pe_hi + (uint32_t)kc * 512u; a8 = c.dl_pe8 + (uint32_t)kc * 256u; a8r = c.dl_pe8r + (uint32_t)kc * 256u;
      } else {
        const int ka = kc - NPE;
        if (half == 0 && ((ka & 1) == 0 || ka == 1) && !a_reused) {   // first touch of freshly written activation columns
          const int cb = ka == 1 ? 4 : (ka >> 1);
          mbar_spin(bar(BAR_A + cb), (a_phase >> cb) & 1u);
          a_phase ^= 1u << cb;
          tr(2, cb);
        }
        ah = c.dl_a_hi + (uint32_t)ka * 512u; a8 = c.dl_a8 + (uint32_t)ka * 256u; a8r = c.dl_a8r + (uint32_t)ka * 256u;
      }
      mbar_spin(bar(BAR_W_FULL + st), phase);
      tc_fence_after();
      // descriptor low word of stage st with the N = 256 leading-byte offset (4096): wdesc0 carries LBO 2048 in bits 16+
      const uint32_t wb = c.wdesc0 + (uint32_t)st * 1024u + (128u << 16);
      if (half == 0) {
        tc_mma<256>(d, ah, wb, kc > 0 ? 1u : 0u);        // A_hi * W_hi  (k 0..15)
        tc_mma<256>(d, ah + 256u, wb + 512u, 1u);        // (k 16..31)
      } else {
        tc_mma_f8<256>(d, a8r, wb, 1u);                  // (2^10 A_lo) * (2^-10 W_hi)
        tc_mma_f8<256>(d, a8, wb + 512u, 1u);            // A_hi * W_lo
      }
      tc_commit(bar(BAR_W_EMPTY + st));
    }
    phase ^= 1u;
  }
  tc_commit(bar(BAR_ACC + acc_slot));
}

// ---- a 128-wide step (mirror head, dir layer) of the tc2 kernels, fully unrolled: K = 256 = 8 stages of
// [W_hi fp16 8 KB | e4m3(2^-10 W_hi) 4 KB | e4m3(W_lo) 4 KB], four N = 128 MMAs each (64 pipe cycles apiece, so the issue
// work per stage counts twice as much as in the 256-wide steps: the generic loop ran them at ~440 cycles per 256-cycle stage)
template <class Tracer>
__device__ __forceinline__ void issue_narrow_step(const SplitCtx& c, uint32_t d, int acc_slot, bool a_reused, uint32_t& phase,
                                                  uint32_t& a_phase, Tracer&& tr) {
  auto bar = [&](int i) { return c.bars + 8u * (uint32_t)i; };
#pragma unroll
  for (int t = 0; t < 2; ++t) {
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const int ka = 4 * t + st;
      if (((ka & 1) == 0 || ka == 1) && !a_reused) {   // first touch of freshly written activation columns
        const int cb = ka == 1 ? 4 : (ka >> 1);
        mbar_spin(bar(BAR_A + cb), (a_phase >> cb) & 1u);
        a_phase ^= 1u << cb;
        tr(2, cb);
      }
      const uint32_t ah = c.dl_a_hi + (uint32_t)ka * 512u, a8 = c.dl_a8 + (uint32_t)ka * 256u, a8r = c.dl_a8r + (uint32_t)ka * 256u;
      mbar_spin(bar(BAR_W_FULL + st), phase);
      tc_fence_after();
      const uint32_t wb = c.wdesc0 + (uint32_t)st * 1024u;
      tc_mma<128>(d, ah, wb, ka > 0 ? 1u : 0u);
      tc_mma<128>(d, ah + 256u, wb + 256u, 1u);
      tc_mma_f8<128>(d, a8r, wb + 512u, 1u);
      tc_mma_f8<128>(d, a8, wb + 768u, 1u);
      tc_commit(bar(BAR_W_EMPTY + st));
    }
    phase ^= 1u;
  }
  tc_commit(bar(BAR_ACC + acc_slot));
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- x = hi + lo in fp16, two values per 32-bit word (element 0 in the low half) -----------------------------------
template <bool RELU, int PREC>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (PREC == 3) {
    // hi rounded toward zero so that the residual of a non-negative value is non-negative: both ReLUs ride on the cvt
    if (RELU) asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    else      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float2 hf = __half22float2(*reinterpret_cast<__half2*>(&hi));
    float ra, rb;   // both residuals with one packed subtraction (FADD2)
    f2_unpack(f2_sub(f2_pack(a, b), f2_pack(hf.x, hf.y)), ra, rb);
    if (RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
    else      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
  } else {
    if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    else      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    lo = 0;
  }
}

// store 8 consecutive K values (one 16-byte core-matrix row) of an A-type operand, hi and lo parts
template <bool RELU, int PREC>
__device__ __forceinline__ void store_a8(uint32_t hi_addr, uint32_t lo_addr, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2<RELU, PREC>(v[2 * i], v[2 * i + 1], h[i], l[i]);
  st_shared_v4(hi_addr, h[0], h[1], h[2], h[3]);
  if (PREC == 3) st_shared_v4(lo_addr, l[0], l[1], l[2], l[3]);
}

// ---- tc2: 16 consecutive K values of an A operand -> fp16 hi (two core-matrix rows) + e4m3(x) + e4m3(2^10 (x - hi)) -----------
// hi16_addr: the core-matrix row of the first 8 values (the next 8 are one K-group = 2048 B further); a8_addr: the 16-byte row
// of the e4m3 copy of x; the residual copy lives A8_LO_OFF bytes behind it.
template <bool RELU>
__device__ __forceinline__ uint32_t cvt_e4m3x2(float a, float b) {  // {low byte = e4m3(a), high byte = e4m3(b)}
  uint16_t r;
  if (RELU) asm("cvt.rn.satfinite.relu.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(b), "f"(a));
  else      asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(b), "f"(a));
  return (uint32_t)r;
}
template <bool RELU>
__device__ __forceinline__ void store_a16_tc2(uint32_t hi16_addr, uint32_t a8_addr, uint32_t lo8_off, const f32x2 (&v)[8]) {
  uint32_t h[8], x8[4], l8[4];
  const f32x2 k_pos = f2_pack(1024.f, 1024.f), k_neg = f2_pack(-1024.f, -1024.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a, b;
    f2_unpack(v[i], a, b);
    // The ReLU rides on the conversions (no separate max): hi = rz(relu(x)) so that the residual of a positive value is >= 0;
    // for x < 0: hi = 0 and the residual x - 0 < 0 is clamped by the relu of its own conversion.  Signed values: rn, no relu.
    if (RELU) asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(b), "f"(a));
    else      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(b), "f"(a));
    const float2 hf = __half22float2(*reinterpret_cast<__half2*>(&h[i]));
    const uint32_t xa = cvt_e4m3x2<RELU>(a, b);
    // 2^10 (x - hi), both lanes at once: x - hi is exact (hi is x with low mantissa bits removed) and so is the scaling, hence
    // fma(x, 1024, -1024 hi) returns the same bits as the scalar (x - hi) * 1024
    float la_, lb_;
    f2_unpack(f2_fma(v[i], k_pos, f2_mul(f2_pack(hf.x, hf.y), k_neg)), la_, lb_);
    const uint32_t la = cvt_e4m3x2<RELU>(la_, lb_);
    if (i & 1) { x8[i >> 1] |= xa << 16; l8[i >> 1] |= la << 16; }
    else       { x8[i >> 1] = xa;        l8[i >> 1] = la; }
  }
  st_shared_v4(hi16_addr, h[0], h[1], h[2], h[3]);
  st_shared_v4(hi16_addr + 2048u, h[4], h[5], h[6], h[7]);
  st_shared_v4(a8_addr, x8[0], x8[1], x8[2], x8[3]);
  st_shared_v4(a8_addr + lo8_off, l8[0], l8[1], l8[2], l8[3]);
}

// accurate sin/cos (arguments reach 2^9 * |x|): shared, not inlined 30 times
__device__ __noinline__ float2 sincos_pe(float a) {
  float s, c;
  sincosf(a, &s, &c);
  return make_float2(s, c);
}

// ---- positional encoding of one row, K range [16*QUARTER, 16*QUARTER+16) (mirror_nerf.py:33-38) --------
template <int QUARTER, int PREC>
__device__ __forceinline__ void pe_fill(const float (&x)[3], uint32_t pe_hi, uint32_t pe_lo, uint32_t rowoff) {
  constexpr int K0 = 16 * QUARTER;
  float vals[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) vals[i] = 0.f;  // k = 63 stays 0 (padding)
  if (QUARTER == 0) {
    vals[0] = x[0]; vals[1] = x[1]; vals[2] = x[2];
  }
#pragma unroll
  for (int f = 0; f < NFREQ_XYZ; ++f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int ks = 3 + 6 * f + c, kc = ks + 3;
      const bool need_s = (ks >= K0 && ks < K0 + 16), need_c = (kc >= K0 && kc < K0 + 16);
      if (need_s || need_c) {
        const float2 sc = sincos_pe(ldexpf(x[c], f));  // 2^f * x is exact
        if (need_s) vals[ks - K0] = sc.x;
        if (need_c) vals[kc - K0] = sc.y;
      }
    }
  }
  if (PREC == 2) {
    // pe_lo = base of the e4m3 copies: [x (8 KB) | 2^10 residual (8 KB)], 16 K values per core-matrix row
    f32x2 v2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v2[i] = f2_pack(vals[2 * i], vals[2 * i + 1]);
    store_a16_tc2<false>(pe_hi + (uint32_t)(2 * QUARTER) * 2048u + rowoff, pe_lo + (uint32_t)QUARTER * 2048u + rowoff, 8192u, v2);
    return;
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = vals[8 * j + i];
    const uint32_t off = (uint32_t)(2 * QUARTER + j) * 2048u + rowoff;
    store_a8<false, PREC>(pe_hi + off, pe_lo + off, v);
  }
}

// ---- epilogue of NC accumulator columns of one row ------------------------------------------------------------------
// MASK: 0 = none; 1 = record (bit i of `mbits` from `bit0` up := value > 0, the ReLU derivative needed by the analytic-normal
// chain); 2 = apply (value := bit ? value : 0, no bias: a step of the chain g_{l-1} = (g_l W_l) * relu'(h_{l-1})).
template <int NC, bool RELU, bool DOTS, bool WRITE_A, int PREC, int MASK>
__device__ __forceinline__ void epi_cols(const uint32_t (&r)[NC], const float4 (&b)[NC / 4], float inv, uint32_t s_hi,
                                         uint32_t s_lo, const float4* __restrict__ hw, float (&d)[4], uint32_t& mbits,
                                         int bit0) {
  if (PREC == 2) {
    // s_lo = address of this thread's e4m3 row for the first 16 columns (see layer_epilogue); residual copy 32 KB behind
    const f32x2 inv2 = f2_pack(inv, inv);
#pragma unroll
    for (int j = 0; j < NC / 16; ++j) {
      f32x2 v[8];   // 16 values as packed fp32 pairs (FFMA2: half the issue slots of the scalar bias / scale FMAs)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 bb = b[4 * j + i];
        v[2 * i + 0] = f2_fma(f2_pack(__uint_as_float(r[16 * j + 4 * i + 0]), __uint_as_float(r[16 * j + 4 * i + 1])), inv2, f2_pack(bb.x, bb.y));
        v[2 * i + 1] = f2_fma(f2_pack(__uint_as_float(r[16 * j + 4 * i + 2]), __uint_as_float(r[16 * j + 4 * i + 3])), inv2, f2_pack(bb.z, bb.w));
      }
      if (DOTS) {
        f32x2 d01 = f2_pack(d[0], d[1]), d23 = f2_pack(d[2], d[3]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a0, a1;
          f2_unpack(v[i], a0, a1);
          a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f);
          v[i] = f2_pack(a0, a1);
          const float4 w0 = hw[16 * j + 2 * i], w1 = hw[16 * j + 2 * i + 1];
          d01 = f2_fma(f2_pack(a0, a0), f2_pack(w0.x, w0.y), d01); d23 = f2_fma(f2_pack(a0, a0), f2_pack(w0.z, w0.w), d23);
          d01 = f2_fma(f2_pack(a1, a1), f2_pack(w1.x, w1.y), d01); d23 = f2_fma(f2_pack(a1, a1), f2_pack(w1.z, w1.w), d23);
        }
        f2_unpack(d01, d[0], d[1]);
        f2_unpack(d23, d[2], d[3]);
      }
      if (WRITE_A) store_a16_tc2<RELU>(s_hi + (uint32_t)j * 4096u, s_lo + (uint32_t)j * 2048u, 32768u, v);
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < NC / 8; ++j) {
    float v[8];
    if (MASK == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ((mbits >> (bit0 + 8 * j + i)) & 1u) ? __uint_as_float(r[8 * j + i]) * inv : 0.f;
    } else {
      const float4 b0 = b[2 * j], b1 = b[2 * j + 1];
      const f32x2 inv2 = f2_pack(inv, inv);   // scale + bias as packed pairs (FFMA2)
      f2_unpack(f2_fma(f2_pack(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1])), inv2, f2_pack(b0.x, b0.y)), v[0], v[1]);
      f2_unpack(f2_fma(f2_pack(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])), inv2, f2_pack(b0.z, b0.w)), v[2], v[3]);
      f2_unpack(f2_fma(f2_pack(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])), inv2, f2_pack(b1.x, b1.y)), v[4], v[5]);
      f2_unpack(f2_fma(f2_pack(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])), inv2, f2_pack(b1.z, b1.w)), v[6], v[7]);
    }
    if (MASK == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mbits |= (v[i] > 0.f ? 1u : 0u) << (bit0 + 8 * j + i);
    }
    if (DOTS) {
      f32x2 d01 = f2_pack(d[0], d[1]), d23 = f2_pack(d[2], d[3]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = fmaxf(v[i], 0.f);
        const float4 w = hw[8 * j + i];
        const f32x2 vv = f2_pack(v[i], v[i]);
        d01 = f2_fma(vv, f2_pack(w.x, w.y), d01);
        d23 = f2_fma(vv, f2_pack(w.z, w.w), d23);
      }
      f2_unpack(d01, d[0], d[1]);
      f2_unpack(d23, d[2], d[3]);
    }
    if (WRITE_A) store_a8<RELU, PREC>(s_hi + (uint32_t)j * 2048u, s_lo + (uint32_t)j * 2048u, v);
  }
}

// epilogue flavours of a 256-wide layer
struct TagTrunk { static constexpr bool relu = true, dots = false, write_a = true; static constexpr int mask = 0; };   // layers 1..7
struct TagLast  { static constexpr bool relu = true, dots = true, write_a = true; static constexpr int mask = 0; };    // layer 8 (+ sigma / normal dots)
struct TagSigma { static constexpr bool relu = true, dots = true, write_a = false; static constexpr int mask = 0; };   // layer 8 of a sigma-only launch
struct TagFinal { static constexpr bool relu = false, dots = false, write_a = true; static constexpr int mask = 0; };  // xyz_encoding_final (no activation)
struct TagTrunkM { static constexpr bool relu = true, dots = false, write_a = true; static constexpr int mask = 1; };  // ... recording relu' bits
struct TagLastM  { static constexpr bool relu = true, dots = true, write_a = true; static constexpr int mask = 1; };
struct TagSigmaM { static constexpr bool relu = true, dots = true, write_a = false; static constexpr int mask = 1; };
struct TagChain  { static constexpr bool relu = false, dots = false, write_a = true; static constexpr int mask = 2; }; // analytic-normal chain step

// ---- step geometry --------------------------------------------------------------------------------
// TMEM columns: trunk layers alternate [0,256) / [256,512); mirror head (step 9) -> [0,128); final (step 8) -> [128,384);
// dir layer (step 10) -> [384,512).  Issue order: 0..7, 9, 8, 10.
// Analytic-normal chain (steps 11..19, issued in numeric order after step 10): the seven 256-wide steps alternate
// [0,256) / [256,512) again; the two 64-wide PE-gradient steps use 64 columns of the buffer that is idle at that point.
__device__ __forceinline__ int chain_w(int s) { return s <= 14 ? s - 11 : s - 12; }  // index among the wide chain steps (s != 15, 19)
__device__ __forceinline__ uint32_t acc_col(int s) {
  if (s <= 7) return (uint32_t)(s & 1) * 256u;
  if (s <= 10) return s == 9 ? 0u : (s == 8 ? 128u : 384u);
  if (s == 15) return 0u;     // W5^T PE part: after step 14 (which owns [256,512))
  if (s == 19) return 256u;   // W1^T: after step 18 (which owns [0,256))
  return (uint32_t)(chain_w(s) & 1) * 256u;
}
__device__ __forceinline__ int acc_bar(int s) {  // steps 9 and 10 share slot 2; chain: wide -> 0/1, 15 -> 2, 19 -> 3
  if (s <= 7) return s & 1;
  if (s <= 10) return s == 8 ? 3 : 2;
  if (s == 15) return 2;
  if (s == 19) return 3;
  return chain_w(s) & 1;
}
__device__ __forceinline__ int step_at(int i) { return i < 8 ? i : (i == 8 ? 9 : (i == 9 ? 8 : i)); }

// ================================================================================================
// FUSE (ray-tile mode, full non-NORMALS pass only): a tile is 4 rays x 32 consecutive samples -- TMEM lane quarter q = one ray's
// chunk -- handed out dynamically (two rays in flight per quarter, alternating tiles, so that the (ray, chunk) of a tile is
// known one tile ahead for the positional encoding); the epilogue composites the chunk in registers with the arithmetic of
// composite.cu (same warp scan, same running carry: bit-identical), optionally stops a ray whose transmittance fell below
// term_eps, and writes per-ray outputs -- no per-point record goes through HBM.  R/models/rendering.py:175-264,363-367.
template <int PREC, bool NORMALS, bool FUSE = false, bool SPLIT = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_field_tc(const TcParams P) {
  static_assert(!SPLIT || (PREC == 2 && !NORMALS), "the N-split schedule exists for the tc2 kernels without analytic normals");
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bars = sbase + SM_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * BAR_TMEM_SLOT);
  const int n_issue = P.io.sigma_only ? 8 : (NORMALS ? 20 : 11);
  // device-side ray count (mnrf_render_recursive): every role derives the same tile count from it
  int n_points = P.io.n_points;
  if (P.io.n_rays_dev != nullptr) {
    const long long alive = (long long)__ldg(P.io.n_rays_dev) * P.io.S;
    if (alive < n_points) n_points = (int)(alive < 0 ? 0 : alive);
  }
  const int n_tiles = (n_points + TILE_M - 1) / TILE_M;
  const float* const c_epi = c_epi_slots[0];
  // fused mode: work descriptors of the two alternating tile slots, [slot][quarter]; stop flag for the producer / MMA roles
  volatile int* f_ray = reinterpret_cast<volatile int*>(smem + SM_FUSE);        // ray index or -1
  volatile int* f_chunk = f_ray + 8;                                              // 32-sample chunk of that ray
  volatile int* f_stop = f_ray + 16;
  constexpr bool PREC3 = PREC == 3;                 // three fp16 passes
  constexpr bool TWO_BLOBS = PREC != 1;              // weight chunk = two 16 KB halves (3x: hi|lo; tc2: hi16 | hi8,lo8)
  constexpr uint32_t NST = TWO_BLOBS ? 4u : 8u;  // weight stages (the 1x mode also uses the idle A_lo region)  // weight stages (the 1x mode also uses the idle A_lo region)
  auto stage_addr = [&](uint32_t st) { return sbase + (st < 4u ? SM_WST + st * WSTAGE_BYTES : SM_A_LO + (st - 4u) * WSTAGE_BYTES); };

  if (threadIdx.x == 0) {
    if (sbase & 127u) { printf("mnrf field_tc: unaligned dynamic smem base %u\n", sbase); __trap(); }
    for (int i = 0; i < 8; ++i) { mbar_init(bar(BAR_W_FULL + i), 1); mbar_init(bar(BAR_W_EMPTY + i), 1); }
    mbar_init(bar(BAR_PE), 16);   // one arrival per epilogue warp
    mbar_init(bar(BAR_GO), 16);
    // a chunk (or half of chunk 0) is written by one column group: 4 warps; N-split schedule: every warp writes 16 columns of
    // every chunk (the halves of chunk 0 come from groups 0,1 / 2,3)
    for (int i = 0; i < 5; ++i) mbar_init(bar(BAR_A + i), SPLIT ? ((i == 0 || i == 4) ? 8 : 16) : 4);
    mbar_init(bar(BAR_ACCH), 1);
    mbar_init(bar(BAR_AFREE), 1);
    mbar_init(bar(BAR_AFREE + 1), 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_ACC + i), 1);
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
    // =========================== weight producer (one elected thread) ===========================
    // Stage contents.  3x mode: N=256 steps -> one blob (hi or lo of a K32 chunk, 16 KB); N=128 steps -> [hi|lo] of a K32
    // chunk (2 x 8 KB, contiguous).  1x mode: N=256 -> hi blob of a K32 chunk; N=128 -> hi blobs of two K32 chunks.
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, go_phase = 0;
      for (int tile = blockIdx.x; FUSE || tile < n_tiles; tile += gridDim.x) {
        if (FUSE) {  // the epilogue warps decide tile by tile whether there is another one
          mbar_spin(bar(BAR_GO), go_phase);
          go_phase ^= 1u;
          if (*f_stop) break;
        }
        for (int i = 0; i < n_issue; ++i) {
          const int s = step_at(i);
          if (s == 9 && !P.has_mirror) continue;
          // N-split steps: 2 * nch stages of 16 KB, contiguous in (half, K32 chunk) order -- the same walk as the unsplit layout
          const uint8_t* src = ((SPLIT && s <= TC_SPLIT_LAST) ? P.tc_split : P.tc) + tc_step_offset(s);
          const int sn = tc_step_n(s);
          const bool wide = sn == 256;
          const int nch = tc_step_chunks(s);
          // N = 64 steps (4 KB blobs): 3x -> [hi|lo] of two K32 chunks per stage (contiguous); 1x -> hi of four chunks
          const int nst = sn == 64 ? (TWO_BLOBS ? nch / 2 : nch / 4) : (TWO_BLOBS ? (wide ? 2 * nch : nch) : (wide ? nch : nch / 2));
          if constexpr (TWO_BLOBS) {
            // tc3 / tc2: every stage is 16 contiguous KB and every step takes a multiple of 4 stages (4, 8, 16 or 20), so the
            // ring position is unrolled (constant stage / barrier addresses): like the issuer, this thread shares its scheduler
            // with four epilogue warps and its reaction time is part of every stage's round trip.  The generic issue loops
            // walk the same ring in the same order.
            if (nst & 3) { printf("mnrf field_tc: step %d takes %d weight stages\n", s, nst); __trap(); }
            for (int t = 0; t < nst / 4; ++t) {
#pragma unroll
              for (int st = 0; st < 4; ++st) {
                mbar_spin(bar(BAR_W_EMPTY + st), phase ^ 1u);
                const uint32_t fb = bar(BAR_W_FULL + st);
                mbar_expect_tx(fb, WSTAGE_BYTES);
                bulk_g2s(sbase + SM_WST + (uint32_t)st * WSTAGE_BYTES, src, WSTAGE_BYTES, fb);
                src += WSTAGE_BYTES;
              }
              phase ^= 1u;
            }
            continue;
          }
          for (int si = 0; si < nst; ++si) {
            mbar_spin(bar(BAR_W_EMPTY + stage), phase ^ 1u);
            const uint32_t dst = stage_addr(stage);
            const uint32_t fb = bar(BAR_W_FULL + stage);
            if (P.debug & 1) {   // timing experiment only (wrong results): a quarter of the weight bytes per stage
              mbar_expect_tx(fb, 4096u);
              bulk_g2s(dst, src + (size_t)si * 4096u, 4096u, fb);
              if (++stage == NST) { stage = 0; phase ^= 1u; }
              continue;
            }
            mbar_expect_tx(fb, WSTAGE_BYTES);
            if (sn == 64 && !TWO_BLOBS) {
#pragma unroll
              for (int piece = 0; piece < 4; ++piece) bulk_g2s(dst + piece * 4096u, src + (size_t)(4 * si + piece) * 8192, 4096u, fb);
            } else if (TWO_BLOBS || wide) {
              // 16 contiguous KB: blob si (3x wide), blobs 2si,2si+1 (3x narrow), blob 2si = hi of chunk si (1x wide)
              const uint8_t* g = src + (size_t)(TWO_BLOBS ? si : 2 * si) * WSTAGE_BYTES;
              bulk_g2s(dst, g, WSTAGE_BYTES, fb);   // one 16 KB copy (four 4 KB pieces measured the same)
            } else {
              // 1x narrow: hi blobs (8 KB) of chunks 2si and 2si+1; a chunk's [hi|lo] pair is 16 KB
              bulk_g2s(dst, src + (size_t)(2 * si) * 16384, 8192u, fb);
              bulk_g2s(dst + 8192u, src + (size_t)(2 * si + 1) * 16384, 8192u, fb);
            }
            if (++stage == NST) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // =========================== MMA issuer: ONE elected thread runs the whole role ===========================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      uint32_t pe_phase = 0, a_phase = 0;  // a_phase: one parity bit per 64-column A chunk
      TraceCtx trc{0};
      const uint32_t dl_a_hi = desc_lo(sbase + SM_A_HI, 2048), dl_a_lo = desc_lo(sbase + SM_A_LO, 2048);
      const uint32_t dl_pe_hi = desc_lo(sbase + SM_PE_HI, 2048), dl_pe_lo = desc_lo(sbase + SM_PE_LO, 2048);
      // tc2: e4m3 copies of the A operand (x at +0, 2^10 residual behind it); a K32 chunk = two 2048-byte K-groups = 256 units
      const uint32_t dl_a8 = dl_a_lo, dl_a8r = desc_lo(sbase + SM_A_LO + 32768u, 2048);
      const uint32_t dl_pe8 = dl_pe_lo, dl_pe8r = desc_lo(sbase + SM_PE_LO + 8192u, 2048);
      auto next_stage = [&]() { if (++stage == NST) { stage = 0; phase ^= 1u; } };
      SplitCtx sc;
      sc.bars = bars; sc.wdesc0 = desc_lo(sbase + SM_WST, 2048);
      sc.dl_a_hi = dl_a_hi; sc.dl_a8 = dl_a8; sc.dl_a8r = dl_a8r; sc.dl_pe_hi = dl_pe_hi; sc.dl_pe8 = dl_pe8; sc.dl_pe8r = dl_pe8r;
      for (int tile = blockIdx.x; FUSE || tile < n_tiles; tile += gridDim.x) {
        unsigned int w_pe = 0, w_a = 0, w_w = 0;   // trace builds: cycles blocked on the PE / activation / weight barriers
        (void)w_pe; (void)w_a; (void)w_w;
        if (FUSE) {  // PE barrier = "the next tile's encoding is in place" or "stop"
          mbar_spin(bar(BAR_PE), pe_phase);
          pe_phase ^= 1u;
          if (*f_stop) break;
        }
        for (int i = 0; i < n_issue; ++i) {
          const int s = step_at(i);
          if (s == 9 && !P.has_mirror) continue;
          const bool wide = tc_step_n(s) == 256;
          const int nch = tc_step_chunks(s);
          const int n_pe = (s == 0 || s == 4) ? 2 : 0;       // leading K32 chunks that come from the PE buffer
          bool a_reused = (s == 8 && P.has_mirror) || s == 15;  // operand already awaited by the previous GEMM
          // Two steps write accumulator columns that the epilogue producing their A operand still READS while it runs (the
          // four column groups convert their chunks concurrently): the final layer without a mirror head ([128,384) overlaps the
          // last trunk layer's [256,512)) and chain step 16 ([0,256) overlaps step 15's PE-gradient columns, read by every group
          // in front of step 14's epilogue).  They wait for the WHOLE operand first; every other step starts on chunk 0.
          if ((s == 8 && !P.has_mirror) || s == 16) {
#pragma unroll 1
            for (int c = 0; c < 5; ++c) { mbar_spin(bar(BAR_A + c), (a_phase >> c) & 1u); a_phase ^= 1u << c; }
            a_reused = true;
          }
          const uint32_t d_tmem = tmem + acc_col(s);
          uint32_t accumulate = 0;
          trace_ev(P, trc, 0, 1, 1, s, 0);
          if constexpr (SPLIT) if (s <= TC_SPLIT_LAST) {
            // the ring is at position 0 here: every step of this mode consumes a multiple of 4 stages
            auto tr = [&](int ev, int v) { trace_ev(P, trc, 0, 1, ev, s, v); };
            if (s == 0) issue_split_step<2, 2, !FUSE>(sc, d_tmem, acc_bar(s), a_reused, phase, a_phase, pe_phase, tr);
            else if (s == 4) issue_split_step<10, 2, false>(sc, d_tmem, acc_bar(s), a_reused, phase, a_phase, pe_phase, tr);
            else issue_split_step<8, 0, false>(sc, d_tmem, acc_bar(s), a_reused, phase, a_phase, pe_phase, tr);
            trace_ev(P, trc, 0, 1, 3, s, 0);
            continue;
          }
          if constexpr (PREC == 2 && !NORMALS) if ((s == 9 || s == 10) && !(P.debug & 3)) {
            auto tr = [&](int ev, int v) { trace_ev(P, trc, 0, 1, ev, s, v); };
            issue_narrow_step(sc, d_tmem, acc_bar(s), a_reused, phase, a_phase, tr);
            trace_ev(P, trc, 0, 1, 3, s, 0);
            continue;
          }
          if constexpr (PREC == 2 && !NORMALS && !SPLIT) if (s <= 8 && !(P.debug & 3)) {
            // unsplit tc2 schedule of the 256-wide steps: the unrolled issue loop (the ring is at position 0 here)
            auto tr = [&](int ev, int v) { trace_ev(P, trc, 0, 1, ev, s, v); };
            if (s == 0) issue_wide_step<2, 2, !FUSE>(sc, d_tmem, acc_bar(s), a_reused, phase, a_phase, pe_phase, tr);
            else if (s == 4) issue_wide_step<10, 2, false>(sc, d_tmem, acc_bar(s), a_reused, phase, a_phase, pe_phase, tr);
            else issue_wide_step<8, 0, false>(sc, d_tmem, acc_bar(s), a_reused, phase, a_phase, pe_phase, tr);
            trace_ev(P, trc, 0, 1, 3, s, 0);
            continue;
          }
          for (int kc = 0; kc < nch; ++kc) {
            uint32_t ah, al;  // descriptor low words of this K32 chunk of the A operand (hi / lo parts)
            uint32_t a8 = 0, a8r = 0;  // tc2: e4m3 copy of the chunk and of its residual
            if (kc < n_pe) {
              if (!FUSE && s == 0 && kc == 0) { TR_T0(); mbar_spin(bar(BAR_PE), pe_phase); pe_phase ^= 1u; TR_ADD(w_pe); }
              ah = dl_pe_hi + (uint32_t)kc * 512u; al = dl_pe_lo + (uint32_t)kc * 512u;
              a8 = dl_pe8 + (uint32_t)kc * 256u; a8r = dl_pe8r + (uint32_t)kc * 256u;
            } else {
              const int ka = kc - n_pe;
              // first touch of freshly written activation columns: chunk 0 is signalled in two 32-column halves
              if (((ka & 1) == 0 || ka == 1) && !a_reused) {
                const int c = ka == 1 ? 4 : (ka >> 1);
                TR_T0();
                mbar_spin(bar(BAR_A + c), (a_phase >> c) & 1u);
                TR_ADD(w_a);
                a_phase ^= 1u << c;
                trace_ev(P, trc, 0, 1, 2, s, c);
              }
              ah = dl_a_hi + (uint32_t)ka * 512u; al = dl_a_lo + (uint32_t)ka * 512u;
              a8 = dl_a8 + (uint32_t)ka * 256u; a8r = dl_a8r + (uint32_t)ka * 256u;
            }
            if (wide) {
              // ---- N = 256: K16 step of the B operand = 8192 B = 512 units ----
              { TR_T0(); mbar_spin(bar(BAR_W_FULL + stage), phase); TR_ADD(w_w); }
              tc_fence_after();
              uint32_t wb = desc_lo(stage_addr(stage), 4096);
              tc_mma<256>(d_tmem, ah, wb, accumulate);                        // A_hi * W_hi  (k 0..15)
              if (PREC3) tc_mma<256>(d_tmem, al, wb, 1u);                     // A_lo * W_hi
              tc_mma<256>(d_tmem, ah + 256u, wb + 512u, 1u);                  // (k 16..31)
              if (PREC3) tc_mma<256>(d_tmem, al + 256u, wb + 512u, 1u);
              tc_commit(bar(BAR_W_EMPTY + stage));
              next_stage();
              if (PREC3) {
                mbar_spin(bar(BAR_W_FULL + stage), phase);
                tc_fence_after();
                wb = desc_lo(stage_addr(stage), 4096);
                tc_mma<256>(d_tmem, ah, wb, 1u);                              // A_hi * W_lo
                tc_mma<256>(d_tmem, ah + 256u, wb + 512u, 1u);
                tc_commit(bar(BAR_W_EMPTY + stage));
                next_stage();
              }
              if (PREC == 2) {
                // second stage of the chunk = [e4m3(2^-10 W_hi) 8 KB | e4m3(W_lo) 8 KB], 32 K values per instruction
                { TR_T0(); mbar_spin(bar(BAR_W_FULL + stage), phase); TR_ADD(w_w); }
                tc_fence_after();
                wb = desc_lo(stage_addr(stage), 4096);
                tc_mma_f8<256>(d_tmem, a8r, wb, 1u);                          // (2^10 A_lo) * (2^-10 W_hi)
                tc_mma_f8<256>(d_tmem, a8, wb + 512u, 1u);                    // A_hi * W_lo
                tc_commit(bar(BAR_W_EMPTY + stage));
                next_stage();
              }
            } else if (tc_step_n(s) == 64) {
              // ---- N = 64 (PE-gradient steps of the normal chain): K16 step of B = 2048 B = 128 units, LBO = 1024 ----
              if (PREC3) {  // stage = [hi|lo] of chunks 2j, 2j+1 (4 x 4 KB)
                if ((kc & 1) == 0) { mbar_spin(bar(BAR_W_FULL + stage), phase); tc_fence_after(); }
                const uint32_t wb = desc_lo(stage_addr(stage), 1024) + (uint32_t)(kc & 1) * 512u;
                tc_mma<64>(d_tmem, ah, wb, accumulate);
                tc_mma<64>(d_tmem, al, wb, 1u);
                tc_mma<64>(d_tmem, ah + 256u, wb + 128u, 1u);
                tc_mma<64>(d_tmem, al + 256u, wb + 128u, 1u);
                tc_mma<64>(d_tmem, ah, wb + 256u, 1u);
                tc_mma<64>(d_tmem, ah + 256u, wb + 384u, 1u);
                if (kc & 1) { tc_commit(bar(BAR_W_EMPTY + stage)); next_stage(); }
              } else {      // stage = hi of chunks 4j..4j+3
                if ((kc & 3) == 0) { mbar_spin(bar(BAR_W_FULL + stage), phase); tc_fence_after(); }
                const uint32_t wb = desc_lo(stage_addr(stage), 1024) + (uint32_t)(kc & 3) * 256u;
                tc_mma<64>(d_tmem, ah, wb, accumulate);
                tc_mma<64>(d_tmem, ah + 256u, wb + 128u, 1u);
                if ((kc & 3) == 3) { tc_commit(bar(BAR_W_EMPTY + stage)); next_stage(); }
              }
            } else if (PREC == 2) {
              // ---- N = 128, tc2: stage = [W_hi fp16 8 KB | e4m3(2^-10 W_hi) 4 KB | e4m3(W_lo) 4 KB] of this K32 chunk ----
              mbar_spin(bar(BAR_W_FULL + stage), phase);
              tc_fence_after();
              const uint32_t wb = desc_lo(stage_addr(stage), 2048);
              tc_mma<128>(d_tmem, ah, wb, accumulate);
              tc_mma<128>(d_tmem, ah + 256u, wb + 256u, 1u);
              tc_mma_f8<128>(d_tmem, a8r, wb + 512u, 1u);
              tc_mma_f8<128>(d_tmem, a8, wb + 768u, 1u);
              tc_commit(bar(BAR_W_EMPTY + stage));
              next_stage();
            } else if (PREC3) {
              // ---- N = 128, 3x: stage = [W_hi | W_lo] of this K32 chunk; K16 step = 4096 B = 256 units ----
              mbar_spin(bar(BAR_W_FULL + stage), phase);
              tc_fence_after();
              const uint32_t wb = desc_lo(stage_addr(stage), 2048);
              tc_mma<128>(d_tmem, ah, wb, accumulate);
              tc_mma<128>(d_tmem, al, wb, 1u);
              tc_mma<128>(d_tmem, ah + 256u, wb + 256u, 1u);
              tc_mma<128>(d_tmem, al + 256u, wb + 256u, 1u);
              tc_mma<128>(d_tmem, ah, wb + 512u, 1u);
              tc_mma<128>(d_tmem, ah + 256u, wb + 768u, 1u);
              tc_commit(bar(BAR_W_EMPTY + stage));
              next_stage();
            } else {
              // ---- N = 128, 1x: stage = W_hi of two K32 chunks ----
              if ((kc & 1) == 0) { mbar_spin(bar(BAR_W_FULL + stage), phase); tc_fence_after(); }
              const uint32_t wb = desc_lo(stage_addr(stage), 2048) + (uint32_t)(kc & 1) * 512u;
              tc_mma<128>(d_tmem, ah, wb, accumulate);
              tc_mma<128>(d_tmem, ah + 256u, wb + 256u, 1u);
              if (kc & 1) { tc_commit(bar(BAR_W_EMPTY + stage)); next_stage(); }
            }
            accumulate = 1u;
          }
          tc_commit(bar(BAR_ACC + acc_bar(s)));
          trace_ev(P, trc, 0, 1, 3, s, 0);
        }
        trace_val(P, trc, 1, 5, w_pe); trace_val(P, trc, 1, 6, w_a); trace_val(P, trc, 1, 7, w_w);
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 16) {
    // =========================== epilogue / PE warps ===========================
    const int ew = warp - EPI_WARP0;       // 0..15
    const int q = ew & 3;          // TMEM lane quarter == warp_id % 4
    const int g = ew >> 2;         // 16-column quarter of every 64-column chunk (32-column quarter of the 128-wide heads)
    const int row = q * 32 + lane;
    const uint32_t rowoff = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const float4* headw = reinterpret_cast<const float4*>(c_epi + ET_HEADW);
    uint32_t acc_phase = 0;  // one parity bit per accumulator barrier
    TraceCtx trc{0};
    auto wait_acc = [&](int s) {
      const int b = acc_bar(s);
      mbar_wait(bar(BAR_ACC + b), (acc_phase >> b) & 1u);
      acc_phase ^= 1u << b;
      tc_fence_after();
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 10, s, 0);
    };
    auto a_ready = [&](int c) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_A + c));
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 11, c, 0);
    };
    // Sum of up to 4 floats per row over the four column groups, in the fixed order g = 0, 1, 2, 3.  The four warps of a lane
    // quarter own the same TMEM lanes, so the exchange goes through 16 accumulator columns that are idle at that moment
    // (`xcol`): groups 1..3 store, group 0 loads and adds.  Returns the sum in v for g == 0.  Only the four warps of the quarter
    // synchronise (named barrier 3 + q, 128 threads).
    auto quarter_sync = [&]() {   // the four warps (column groups) that own this TMEM lane quarter
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(3 + q) : "memory");
      tc_fence_after();
    };
    auto xreduce = [&](float (&v)[4], uint32_t xcol) {
      quarter_sync();   // every group is done reading whatever accumulator these columns belonged to
      if (g != 0) {
        tmem_st4(tlane + xcol + 4u * (uint32_t)g, v[0], v[1], v[2], v[3]);
        tmem_wait_st();
      }
      quarter_sync();
      if (g == 0) {
        uint32_t r[16];
        tmem_ld16(tlane + xcol, r);
        tmem_wait_ld();
        pin<16>(r);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] += __uint_as_float(r[4 * k + i]);
        }
      }
      quarter_sync();   // the columns may be stored to again (next reduction) or overwritten by a later MMA
    };
    auto pe_point = [&](const float (&x)[3]) {
      if (g == 0) pe_fill<0, PREC>(x, sbase + SM_PE_HI, sbase + SM_PE_LO, rowoff);
      else if (g == 1) pe_fill<1, PREC>(x, sbase + SM_PE_HI, sbase + SM_PE_LO, rowoff);
      else if (g == 2) pe_fill<2, PREC>(x, sbase + SM_PE_HI, sbase + SM_PE_LO, rowoff);
      else pe_fill<3, PREC>(x, sbase + SM_PE_HI, sbase + SM_PE_LO, rowoff);
    };
    // xyz + positional encoding of this thread's row of `tile` -> PE operand buffer
    auto pe_tile = [&](int tile) {
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 12, 0, 0);
      const long long pr = (long long)tile * TILE_M + row;
      const long long p = pr < n_points ? pr : (long long)n_points - 1;
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
      pe_point(x);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_PE));
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 13, 0, 0);
    };
    // address of this thread's row in the second operand buffer for the columns that start at byte offset `off` of the fp16
    // buffer: 3x -> the fp16 lo part (same layout); tc2 -> the e4m3 copy (16 K values per 16-byte row: half the K-group count)
    auto lo_addr = [&](uint32_t off) {
      return PREC == 2 ? sbase + SM_A_LO + ((off - rowoff) >> 1) + rowoff : sbase + SM_A_LO + off;
    };
    // One 256-wide layer: column group g owns the whole 64-column K chunk g of the next layer's A operand (its 32 rows of it),
    // converted as four 16-column pieces with the next piece's TMEM load in flight.  The four chunks are therefore produced
    // CONCURRENTLY by the four groups instead of one after the other by everybody: the next layer's MMAs find chunk 0 after two
    // pieces of one warp and all other chunks shortly after (measured before: the serial chain of five load -> convert ->
    // proxy fence -> barrier rounds per layer, not instruction issue, set the layer period).  Group 0 signals its chunk in two
    // 32-column halves (BAR_A[0], BAR_A[4]); groups 1..3 signal BAR_A[g] once.
    // mk: this thread's two 32-bit relu' words of the layer (piece j -> bits [16 (j & 1), +16) of mk[j >> 1])
    auto layer_epilogue = [&](auto tag, int s, const float* bias256, float (&d)[4], uint32_t (&mk)[2]) {
      constexpr bool RELU = decltype(tag)::relu, DOTS = decltype(tag)::dots, WRITE_A = decltype(tag)::write_a;
      constexpr int MASK = decltype(tag)::mask;
      if (MASK == 1) { mk[0] = 0u; mk[1] = 0u; }
      const float4* b4 = reinterpret_cast<const float4*>(bias256) + 16 * g;
      // everything that does not depend on the accumulator is fetched before waiting for it
      const float inv = c_epi[ET_INV_SCALE + s] * (PREC == 2 ? 0.03125f : 1.f);  // tc2 blobs carry 2^5 more scale (pack.cu)
      float4 b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = b4[i];
      wait_acc(s);
      // Stagger the groups (named barriers 7..9, arrive = non-blocking): the next layer's MMAs wait for chunk 0, so group 0
      // converts it with the schedulers to itself (measured: with all 16 warps converting at once a 16-column piece took 800
      // cycles and chunk 0 appeared 2,300 cycles after the accumulator); group 1 starts when group 0 is half done, group 2 when
      // group 0 is done, group 3 when group 1 is done -- each chunk is still ready before the MMAs reach it.
      if (g == 1) asm volatile("bar.sync 7, 256;" ::: "memory");
      else if (g == 2) asm volatile("bar.sync 8, 256;" ::: "memory");
      else if (g == 3) asm volatile("bar.sync 9, 256;" ::: "memory");
      const uint32_t tcol = tlane + acc_col(s) + 64u * (uint32_t)g;
      uint32_t ra[16], rb[16];
      tmem_ld16(tcol, ra);
      tmem_wait_ld();
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 20, s, 0);   // first TMEM load landed
      tmem_ld16(tcol + 16u, rb);  // piece 1
      pin<16>(ra);
      {
        const uint32_t off = (uint32_t)(8 * g) * 2048u + rowoff;
        epi_cols<16, RELU, DOTS, WRITE_A, PREC, MASK>(ra, b, inv, sbase + SM_A_HI + off, lo_addr(off), headw + 64 * g, d, mk[0], 0);
      }
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 21, s, 0);   // piece 0 converted and stored
#pragma unroll
      for (int j = 1; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = b4[4 * j + i];
        tmem_wait_ld();  // the load of piece j (issued one iteration ago) has landed
        if (j < 3) {     // next piece's load flies while this piece is converted
          if (j & 1) tmem_ld16(tcol + (uint32_t)(j + 1) * 16u, ra);
          else       tmem_ld16(tcol + (uint32_t)(j + 1) * 16u, rb);
        }
        const uint32_t off = (uint32_t)(8 * g + 2 * j) * 2048u + rowoff;
        if (j & 1) { pin<16>(rb); epi_cols<16, RELU, DOTS, WRITE_A, PREC, MASK>(rb, b, inv, sbase + SM_A_HI + off, lo_addr(off), headw + 64 * g + 16 * j, d, mk[j >> 1], 16 * (j & 1)); }
        else       { pin<16>(ra); epi_cols<16, RELU, DOTS, WRITE_A, PREC, MASK>(ra, b, inv, sbase + SM_A_HI + off, lo_addr(off), headw + 64 * g + 16 * j, d, mk[j >> 1], 16 * (j & 1)); }
        if (q == 0) trace_ev(P, trc, lane, 4 + g, 22, s, j);  // piece j converted and stored
        if (WRITE_A && g == 0 && j == 1) a_ready(0);   // columns 0..31: the next layer's first K32 chunk
        if (WRITE_A && j == 3) a_ready(g == 0 ? 4 : g);
        if (g == 0 && j == 1) asm volatile("bar.arrive 7, 256;" ::: "memory");   // releases group 1
        if (g == 0 && j == 3) asm volatile("bar.arrive 8, 256;" ::: "memory");   // releases group 2
        if (g == 1 && j == 3) asm volatile("bar.arrive 9, 256;" ::: "memory");   // releases group 3
      }
    };

    // N-split schedule (issue_split_step): this warp owns columns [64c + 16g, +16) of every 64-column chunk c, chunks in K
    // order.  Chunks 0 and 1 come from the first accumulator half (ready while the second half's MMAs still run; their stores
    // wait until those MMAs have read the old A chunk for the last time), chunks 2 and 3 from the second half.  Chunk 0 is
    // signalled in two 32-column halves (groups 0,1 -> BAR_A[0], groups 2,3 -> BAR_A[4]).
    uint32_t acch_phase = 0, afree_phase = 0;
    auto layer_epilogue_split = [&](auto tag, int s, const float* bias256, float (&d)[4]) {
      constexpr bool RELU = decltype(tag)::relu, DOTS = decltype(tag)::dots, WRITE_A = decltype(tag)::write_a;
      const float4* b4 = reinterpret_cast<const float4*>(bias256) + 4 * g;
      const float inv = c_epi[ET_INV_SCALE + s] * 0.03125f;   // tc2 blobs carry 2^5 more scale (pack.cu)
      uint32_t mk = 0u;
      float4 b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = b4[i];
      const uint32_t tcol = tlane + acc_col(s) + 16u * (uint32_t)g;
      uint32_t ra[16], rb[16];
      // ---- first half: chunks 0, 1 ----
      mbar_wait(bar(BAR_ACCH), acch_phase);
      acch_phase ^= 1u;
      tc_fence_after();
      if (q == 0) trace_ev(P, trc, lane, 4 + g, 10, s, 0);
      tmem_ld16(tcol, ra);
      tmem_wait_ld();
      tmem_ld16(tcol + 64u, rb);
      pin<16>(ra);
      mbar_wait(bar(BAR_AFREE + 0), afree_phase & 1u);   // the second half's MMAs are past A chunk 0
      tc_fence_after();
      {
        const uint32_t off = (uint32_t)(2 * g) * 2048u + rowoff;
        epi_cols<16, RELU, DOTS, WRITE_A, PREC, 0>(ra, b, inv, sbase + SM_A_HI + off, lo_addr(off), headw + 16 * g, d, mk, 0);
        if (WRITE_A) a_ready(g < 2 ? 0 : 4);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = b4[16 + i];
      tmem_wait_ld();
      pin<16>(rb);
      mbar_wait(bar(BAR_AFREE + 1), (afree_phase >> 1) & 1u);
      afree_phase ^= 3u;
      tc_fence_after();
      {
        const uint32_t off = (uint32_t)(8 + 2 * g) * 2048u + rowoff;
        epi_cols<16, RELU, DOTS, WRITE_A, PREC, 0>(rb, b, inv, sbase + SM_A_HI + off, lo_addr(off), headw + 64 + 16 * g, d, mk, 0);
        if (WRITE_A) a_ready(1);
      }
      // ---- second half: chunks 2, 3 ----
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = b4[32 + i];
      {
        const int ab = acc_bar(s);
        mbar_wait(bar(BAR_ACC + ab), (acc_phase >> ab) & 1u);
        acc_phase ^= 1u << ab;
        tc_fence_after();
        if (q == 0) trace_ev(P, trc, lane, 4 + g, 20, s, 0);
      }
      tmem_ld16(tcol + 128u, ra);
      tmem_wait_ld();
      tmem_ld16(tcol + 192u, rb);
      pin<16>(ra);
      {
        const uint32_t off = (uint32_t)(16 + 2 * g) * 2048u + rowoff;
        epi_cols<16, RELU, DOTS, WRITE_A, PREC, 0>(ra, b, inv, sbase + SM_A_HI + off, lo_addr(off), headw + 128 + 16 * g, d, mk, 0);
        if (WRITE_A) a_ready(2);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = b4[48 + i];
      tmem_wait_ld();
      pin<16>(rb);
      {
        const uint32_t off = (uint32_t)(24 + 2 * g) * 2048u + rowoff;
        epi_cols<16, RELU, DOTS, WRITE_A, PREC, 0>(rb, b, inv, sbase + SM_A_HI + off, lo_addr(off), headw + 192 + 16 * g, d, mk, 0);
        if (WRITE_A) a_ready(3);
      }
    };

    // One tile: p = this thread's point (clamped), ray = its ray; pe_next() is called while layer 6 runs (the PE buffer is free
    // then) to encode the following tile; emit(...) receives this thread's per-point results (g == 0 threads hold them).
    auto tile_body = [&](const long long p_raw, const bool valid, const long long p, const long long ray, auto&& pe_next,
                         auto&& emit) {
      float o_sigma = 0.f, o_n[3] = {0.f, 0.f, 0.f}, o_mirror = 0.f, o_rgb[3] = {0.f, 0.f, 0.f};
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t masks[NORMALS ? 8 : 1][2];  // relu' bits of this thread's (row, columns) for every trunk layer
      float o_an[3] = {0.f, 0.f, 0.f};     // analytic normal

      // ---- trunk layers 1..8 (steps 0..7) ----
#pragma unroll 1
      for (int s = 0; s < 8; ++s) {
        const float* bias = c_epi + ET_BIAS + 256 * s;
        if constexpr (NORMALS) {
          if (s < 7) layer_epilogue(TagTrunkM{}, s, bias, d, masks[s]);
          else if (!P.io.sigma_only) layer_epilogue(TagLastM{}, s, bias, d, masks[s]);
          else layer_epilogue(TagSigmaM{}, s, bias, d, masks[s]);
        } else if constexpr (SPLIT) {
          if (s < 7) layer_epilogue_split(TagTrunk{}, s, bias, d);
          else if (!P.io.sigma_only) layer_epilogue_split(TagLast{}, s, bias, d);
          else layer_epilogue_split(TagSigma{}, s, bias, d);
        } else {
          if (s < 7) layer_epilogue(TagTrunk{}, s, bias, d, masks[0]);
          else if (!P.io.sigma_only) layer_epilogue(TagLast{}, s, bias, d, masks[0]);
          else layer_epilogue(TagSigma{}, s, bias, d, masks[0]);
        }
        // the PE buffer is free once layer 5's MMAs are done: encode the next tile while the tensor pipe is busy
        if (s == 5) pe_next();
      }
      // combine the four column groups' partial dot products (sigma + folded normal head); columns [384,400) are idle here
      // (step 7's accumulator is consumed, the dir layer's MMAs come after the final layer's epilogue)
      {
        xreduce(d, 384u);
        if (g == 0) {
          const float4 hb = *reinterpret_cast<const float4*>(c_epi + ET_HEADB);
          o_sigma = d[0] + hb.x;
          if (P.has_normal) {
            const float a = d[1] + hb.y, b = d[2] + hb.z, cc = d[3] + hb.w;
            const float nn = sqrtf(fmaxf(a * a + b * b + cc * cc, FP32_EPS));  // utils/func.py:5-7
            o_n[0] = a / nn; o_n[1] = b / nn; o_n[2] = cc / nn;
          }
        }
      }

      if (!P.io.sigma_only) {
        // ---- mirror head (step 9): LeakyReLU(0.01) -> Linear(128,1) -> sigmoid (mirror_nerf.py:94-99) ----
        if (P.has_mirror) {
          wait_acc(9);
          const float inv = c_epi[ET_INV_SCALE + 9] * (PREC == 2 ? 0.03125f : 1.f);
          float dm[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t ra[32];
          tmem_ld32(tlane + acc_col(9) + (uint32_t)g * 32u, ra);
          const float4* bm = reinterpret_cast<const float4*>(c_epi + ET_B_M0 + g * 32);
          const float4* wm = reinterpret_cast<const float4*>(c_epi + ET_W_M2 + g * 32);
          tmem_wait_ld();
          pin32(ra);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = bm[i], ww = wm[i];
            float v0 = fmaf(__uint_as_float(ra[4 * i + 0]), inv, bb.x), v1 = fmaf(__uint_as_float(ra[4 * i + 1]), inv, bb.y);
            float v2 = fmaf(__uint_as_float(ra[4 * i + 2]), inv, bb.z), v3 = fmaf(__uint_as_float(ra[4 * i + 3]), inv, bb.w);
            v0 = v0 > 0.f ? v0 : 0.01f * v0; v1 = v1 > 0.f ? v1 : 0.01f * v1;
            v2 = v2 > 0.f ? v2 : 0.01f * v2; v3 = v3 > 0.f ? v3 : 0.01f * v3;
            dm[0] = fmaf(v0, ww.x, dm[0]); dm[0] = fmaf(v1, ww.y, dm[0]); dm[0] = fmaf(v2, ww.z, dm[0]); dm[0] = fmaf(v3, ww.w, dm[0]);
          }
          xreduce(dm, 384u);
          if (g == 0) o_mirror = sigmoidf_(dm[0] + c_epi[ET_B_M2]);
        }
        // ---- final linear (step 8): f = W h8 + b, written over h8 (its readers, steps 9 and 8, are complete) ----
        if constexpr (SPLIT) layer_epilogue_split(TagFinal{}, 8, c_epi + ET_BIAS + 256 * 8, d);
        else layer_epilogue(TagFinal{}, 8, c_epi + ET_BIAS + 256 * 8, d, masks[0]);
        // ---- dir layer (step 10): relu(W_f f + [b + W_d embed(dir)]) -> rgb (mirror_nerf.py:199-204) ----
        wait_acc(10);
        {
          const float inv = c_epi[ET_INV_SCALE + 10] * (PREC == 2 ? 0.03125f : 1.f);
          const float4* db = reinterpret_cast<const float4*>(P.io.dirbias + ray * WH + g * 32);
          const float4* wr = reinterpret_cast<const float4*>(c_epi + ET_W_RGB + g * 32);
          float dc[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t ra[32];
          tmem_ld32(tlane + acc_col(10) + (uint32_t)g * 32u, ra);
          tmem_wait_ld();
          pin32(ra);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(db + i), w0 = wr[i], w1 = wr[WH / 4 + i], w2 = wr[2 * (WH / 4) + i];
            const float v0 = fmaxf(fmaf(__uint_as_float(ra[4 * i + 0]), inv, bb.x), 0.f);
            const float v1 = fmaxf(fmaf(__uint_as_float(ra[4 * i + 1]), inv, bb.y), 0.f);
            const float v2 = fmaxf(fmaf(__uint_as_float(ra[4 * i + 2]), inv, bb.z), 0.f);
            const float v3 = fmaxf(fmaf(__uint_as_float(ra[4 * i + 3]), inv, bb.w), 0.f);
            dc[0] = fmaf(v0, w0.x, dc[0]); dc[0] = fmaf(v1, w0.y, dc[0]); dc[0] = fmaf(v2, w0.z, dc[0]); dc[0] = fmaf(v3, w0.w, dc[0]);
            dc[1] = fmaf(v0, w1.x, dc[1]); dc[1] = fmaf(v1, w1.y, dc[1]); dc[1] = fmaf(v2, w1.z, dc[1]); dc[1] = fmaf(v3, w1.w, dc[1]);
            dc[2] = fmaf(v0, w2.x, dc[2]); dc[2] = fmaf(v1, w2.y, dc[2]); dc[2] = fmaf(v2, w2.z, dc[2]); dc[2] = fmaf(v3, w2.w, dc[2]);
          }
          // columns [256,272): inside the final layer's accumulator (consumed), not touched by the next tile's first layer
          xreduce(dc, 256u);
          if (g == 0) {
            o_rgb[0] = sigmoidf_(dc[0] + c_epi[ET_B_RGB + 0]);
            o_rgb[1] = sigmoidf_(dc[1] + c_epi[ET_B_RGB + 1]);
            o_rgb[2] = sigmoidf_(dc[2] + c_epi[ET_B_RGB + 2]);
          }
        }
      }

      // ---- analytic normal: n = normalize(-d sigma / d xyz) as a reverse chain on the tensor cores (mirror_nerf.py:136-146) ----
      if (NORMALS && !P.io.sigma_only) {
        // g7 = w_sigma * relu'(h8), written straight into the A buffer (its last reader, the dir GEMM, is complete)
        {
          const float4* hw = headw;
#pragma unroll
          for (int c = 0; c < 4; ++c) {  // same column ownership as layer_epilogue: piece c of chunk g
            const int col0 = 64 * g + 16 * c;
            const uint32_t bits = masks[7][c >> 1] >> (16 * (c & 1));
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = ((bits >> (8 * j + i)) & 1u) ? hw[col0 + 8 * j + i].x : 0.f;
              const uint32_t off = (uint32_t)(col0 / 8 + j) * 2048u + rowoff;
              store_a8<false, PREC>(sbase + SM_A_HI + off, sbase + SM_A_LO + off, v);
            }
            if (g == 0 && c == 1) a_ready(0);
            if (c == 3) a_ready(g == 0 ? 4 : g);
          }
        }
        float gpe[16];  // this thread's 16 of the 64 PE-gradient entries: columns [16g, 16g+16)
        // wide chain steps: 11 (W8^T) .. 18 (W2^T), masks of the layer whose output the gradient now refers to
#pragma unroll 1
        for (int s = 11; s <= 18; ++s) {
          if (s == 15) continue;
          if (s == 14) {
            // the PE part of layer 5 (step 15) lands in 64 columns that the NEXT wide step overwrites: read it first
            wait_acc(15);
            uint32_t r[16];
            tmem_ld16(tlane + acc_col(15) + (uint32_t)g * 16u, r);
            tmem_wait_ld();
            pin<16>(r);
            const float inv15 = c_epi[ET_INV_SCALE + 15];
#pragma unroll
            for (int i = 0; i < 16; ++i) gpe[i] = __uint_as_float(r[i]) * inv15;
          }
          const int layer = tc_step_layer(s);  // 0-based trunk layer of this transposed weight; gradient is w.r.t. its input
          layer_epilogue(TagChain{}, s, c_epi, d, masks[layer - 1]);
        }
        // last step 19 (W1^T): PE gradient of layer 1; total PE gradient -> xyz gradient through the PE Jacobian
        wait_acc(19);
        {
          uint32_t r[16];
          tmem_ld16(tlane + acc_col(19) + (uint32_t)g * 16u, r);
          tmem_wait_ld();
          pin<16>(r);
          const float inv19 = c_epi[ET_INV_SCALE + 19];
#pragma unroll
          for (int i = 0; i < 16; ++i) gpe[i] = fmaf(__uint_as_float(r[i]), inv19, gpe[i]);
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
          float dx[4] = {0.f, 0.f, 0.f, 0.f};
          const int K0 = 16 * g;
          if (g == 0) { dx[0] = gpe[0]; dx[1] = gpe[1]; dx[2] = gpe[2]; }
#pragma unroll
          for (int f = 0; f < NFREQ_XYZ; ++f) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int ks = 3 + 6 * f + c, kc = ks + 3;  // PE columns of sin(2^f x_c) and cos(2^f x_c)
              const bool need_s = (ks >= K0 && ks < K0 + 16), need_c = (kc >= K0 && kc < K0 + 16);
              if (need_s || need_c) {
                const float fr = (float)(1 << f);
                const float2 sc = sincos_pe(fr * x[c]);
                if (need_s) dx[c] = fmaf(fr * gpe[(ks - K0) & 15], sc.y, dx[c]);    // d sin = +f cos
                if (need_c) dx[c] = fmaf(-fr * gpe[(kc - K0) & 15], sc.x, dx[c]);   // d cos = -f sin
              }
            }
          }
          xreduce(dx, 384u);   // [384,512) is idle after the chain (step 19 sits in [256,320))
          if (g == 0) {
            const float a = -dx[0], b = -dx[1], cc = -dx[2];
            const float nn = sqrtf(fmaxf(a * a + b * b + cc * cc, FP32_EPS));  // utils/func.py:5-7
            o_an[0] = a / nn; o_an[1] = b / nn; o_an[2] = cc / nn;
          }
        }
      }

      if (q == 0) trace_ev(P, trc, lane, 4 + g, 14, 0, 0);
      tc_fence_before();
      emit(o_sigma, o_rgb, o_mirror, o_n, o_an);
      (void)p_raw; (void)valid;
    };

    if (!FUSE) {
      // ---- static tiles: 128 consecutive points, one record per point ----
      if ((int)blockIdx.x < n_tiles) pe_tile(blockIdx.x);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p_raw = (long long)tile * TILE_M + row;
        const bool valid = p_raw < n_points;
        const long long p = valid ? p_raw : (long long)n_points - 1;
        const long long ray = (P.io.rays != nullptr) ? p / P.io.S : p;
        tile_body(p_raw, valid, p, ray, [&]() { if (tile + (int)gridDim.x < n_tiles) pe_tile(tile + gridDim.x); },
                  [&](float o_sigma, const float (&o_rgb)[3], float o_mirror, const float (&o_n)[3], const float (&o_an)[3]) {
          if (g == 0 && valid) {
            if (P.io.sigma_out != nullptr) P.io.sigma_out[p_raw] = o_sigma;
            if (P.io.raw != nullptr) {
              float4* o = reinterpret_cast<float4*>(P.io.raw + p_raw * 8);
              o[0] = make_float4(o_sigma, o_rgb[0], o_rgb[1], o_rgb[2]);
              o[1] = make_float4(o_mirror, o_n[0], o_n[1], o_n[2]);
            }
            if (NORMALS && P.io.normal_out != nullptr) {
              float* no = P.io.normal_out + p_raw * 3;
              no[0] = o_an[0]; no[1] = o_an[1]; no[2] = o_an[2];
            }
          }
        });
      }
    } else {
      // ---- fused ray tiles ----
      const int S = P.io.S;
      const int nch = (S + 31) >> 5;                  // 32-sample chunks per ray
      int n_rays = P.n_rays;
      if (P.io.n_rays_dev != nullptr) n_rays = min(n_rays, max(__ldg(P.io.n_rays_dev), 0));
      // per-slot compositing state of this quarter's ray (meaningful in the g == 0 warp: lane = sample within the chunk)
      struct RayAcc { float carry, op, r, gg, b, d, m, n0, n1, n2; };
      RayAcc st[2];
      auto reset = [](RayAcc& a) { a.carry = 1.f; a.op = a.r = a.gg = a.b = a.d = a.m = a.n0 = a.n1 = a.n2 = 0.f; };
      reset(st[0]); reset(st[1]);
      auto grab = [&]() {   // warp-uniform: next ray index or -1
        int r = 0;
        if (lane == 0) { r = atomicAdd(P.work_counter, 1); if (r >= n_rays) r = -1; }
        return __shfl_sync(0xffffffffu, r, 0);
      };
      if (g == 0) {
        const int r0 = grab(), r1 = grab();
        if (lane == 0) { f_ray[q] = r0; f_chunk[q] = 0; f_ray[4 + q] = r1; f_chunk[4 + q] = 0; }
        if (warp == EPI_WARP0 && lane == 0) *f_stop = 0;
      }
      epi_bar_sync<512>(1);
      auto slot_alive = [&](int sl) { return (f_ray[4 * sl] >= 0) || (f_ray[4 * sl + 1] >= 0) || (f_ray[4 * sl + 2] >= 0) || (f_ray[4 * sl + 3] >= 0); };
      // encode the tile of slot `sl` (or announce the stop) and release the MMA issuer / weight producer for it
      auto pe_slot = [&](int sl, bool stop) {
        if (!stop) {
          const int rr_ = f_ray[4 * sl + q], ch = f_chunk[4 * sl + q];
          const long long ray = rr_ >= 0 ? rr_ : 0;
          const int smp = min(ch * 32 + lane, S - 1);
          const float* rr = P.io.rays + ray * 8;
          const float z = __ldg(P.io.z + ray * S + smp);
          float x[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(__ldg(rr + c), __fmul_rn(__ldg(rr + 3 + c), z));
          pe_point(x);
        } else if (warp == EPI_WARP0 && lane == 0) {
          *f_stop = 1;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar(BAR_PE)); mbar_arrive(bar(BAR_GO)); }
      };
      int cur = 0;
      bool alive_cur = slot_alive(0), alive_other = slot_alive(1);
      pe_slot(0, !alive_cur);
      while (alive_cur) {
        const int rr_ = f_ray[4 * cur + q], ch = f_chunk[4 * cur + q];
        const bool has_ray = rr_ >= 0;
        const long long ray = has_ray ? rr_ : 0;
        const int smp_raw = ch * 32 + lane;
        const bool valid = has_ray && smp_raw < S;
        const long long p = ray * S + min(smp_raw, S - 1);
        bool pe_done = false;
        tile_body(p, valid, p, ray, [&]() { if (alive_other) { pe_slot(cur ^ 1, false); pe_done = true; } },
                  [&](float o_sigma, const float (&o_rgb)[3], float o_mirror, const float (&o_n)[3], const float (&o_an)[3]) {
          (void)o_an;
          if (g == 0) {
            // ---- composite this chunk (composite.cu::k_composite, one 32-sample block) ----
            RayAcc& A = st[cur];
            float zz = 0.f, alpha = 0.f;
            if (valid) {
              zz = __ldg(P.io.z + p);
              const float delta = (smp_raw + 1 < S) ? __fsub_rn(__ldg(P.io.z + p + 1), zz) : 1e10f;  // rendering.py:182-186
              alpha = __fsub_rn(1.f, expf(-__fmul_rn(delta, fmaxf(o_sigma, 0.f))));
            }
            const float f = valid ? __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f) : 1.f;
            float incl = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const float t = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl *= t;
            }
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.f;
            const float Tr = A.carry * excl;
            const float w = alpha * Tr;
            A.carry *= __shfl_sync(0xffffffffu, incl, 31);
            if (valid) {
              if (P.comp.weights != nullptr) P.comp.weights[p] = w;
              A.op += w;
              A.d = fmaf(w, zz, A.d);
              A.r = fmaf(w, o_rgb[0], A.r); A.gg = fmaf(w, o_rgb[1], A.gg); A.b = fmaf(w, o_rgb[2], A.b);
              A.m = fmaf(w, o_mirror, A.m);
              A.n0 = fmaf(w, o_n[0], A.n0); A.n1 = fmaf(w, o_n[1], A.n1); A.n2 = fmaf(w, o_n[2], A.n2);
              if (P.comp.pred_normal != nullptr) {
                float* pn = P.comp.pred_normal + p * 3;
                pn[0] = o_n[0]; pn[1] = o_n[1]; pn[2] = o_n[2];
              }
            }
            // ---- next work item of this (quarter, slot) ----
            if (has_ray) {
              const bool terminated = P.term_eps > 0.f && A.carry < P.term_eps;   // warp-uniform
              if (ch + 1 >= nch || terminated) {
                auto wsum = [](float v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; };
                float a_op = wsum(A.op), a_d = wsum(A.d), a_r = wsum(A.r), a_g = wsum(A.gg), a_b = wsum(A.b), a_m = wsum(A.m);
                const float a_n0 = wsum(A.n0), a_n1 = wsum(A.n1), a_n2 = wsum(A.n2);
                if (lane == 0) {
                  const mnrf_composite_out& O = P.comp;
                  O.opacity[ray] = a_op;
                  if (P.white_back) { const float bg = 1.f - a_op; a_r += bg; a_g += bg; a_b += bg; }  // rendering.py:216-217
                  if (O.rgb) { O.rgb[ray * 3 + 0] = a_r; O.rgb[ray * 3 + 1] = a_g; O.rgb[ray * 3 + 2] = a_b; }
                  if (O.depth) O.depth[ray] = a_d;
                  if (O.mirror_mask) O.mirror_mask[ray] = a_m;
                  if (O.surface_normal) { O.surface_normal[ray * 3 + 0] = a_n0; O.surface_normal[ray * 3 + 1] = a_n1; O.surface_normal[ray * 3 + 2] = a_n2; }
                  if (O.x_surface) {
                    const float* ry = P.io.rays + ray * 8;
                    for (int c = 0; c < 3; ++c) O.x_surface[ray * 3 + c] = __fadd_rn(ry[c], __fmul_rn(ry[3 + c], a_d));
                  }
                  if (P.stats != nullptr && ch + 1 < nch) atomicAdd(P.stats + 1, (unsigned long long)(nch - 1 - ch));
                }
                reset(A);
                const int nr = grab();
                if (lane == 0) { f_ray[4 * cur + q] = nr; f_chunk[4 * cur + q] = 0; }
              } else if (lane == 0) {
                f_chunk[4 * cur + q] = ch + 1;
              }
            }
          }
        });
        if (warp == EPI_WARP0 && lane == 0 && P.stats != nullptr) atomicAdd(P.stats, 1ull);
        epi_bar_sync<512>(1);   // every quarter's next work item of slot `cur` is published
        const bool alive_this = slot_alive(cur);
        if (pe_done) {           // the other slot's tile is already encoded and released: it runs next
          alive_other = alive_this;
          cur ^= 1;
          alive_cur = true;
        } else {                 // the other slot had run dry: stay on this one (its encoding could not be prepared ahead)
          alive_cur = alive_this;
          pe_slot(cur, !alive_this);
        }
        epi_bar_sync<512>(2);   // nobody re-reads the descriptors of the finished tile after this point
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
  if (FUSE && threadIdx.x == 0) {
    // the last CTA to finish rewinds the work queue, so the launch can be replayed as is (ncu kernel replay, CUDA graphs)
    if (atomicAdd(P.work_counter + 1, 1) == (int)gridDim.x - 1) {
      P.work_counter[0] = 0;
      P.work_counter[1] = 0;
    }
  }
}

}  // namespace

static int g_tc_split = -1;                 // -1: not decided yet (environment / built-in default at the first launch)
constexpr int TC_SPLIT_DEFAULT = 0;
static unsigned long long* g_trace_buf = nullptr;
static unsigned int g_trace_cap = 0;
void set_tc_trace(unsigned long long* buf, unsigned int cap) { g_trace_buf = buf; g_trace_cap = cap; }
int set_tc_split(int split) {
  if (split < 0) {
    const char* e = getenv("MNRF_TC_SPLIT");
    split = e != nullptr ? (atoi(e) != 0) : TC_SPLIT_DEFAULT;
  }
  g_tc_split = split != 0 ? 1 : 0;
  return g_tc_split;
}

// ---- constant-memory slots of the epilogue table (per device) -------------------------------------------------------------
// A slot is keyed by the field's pack stamp (unique per pack_field call, so a re-packed or re-created field never hits a stale
// copy).  `filled` orders the table copy before readers on other streams, `used` orders a rewrite after the last reader.
namespace {
struct EpiSlot { unsigned long long stamp = 0; unsigned long long tick = 0; cudaEvent_t filled = nullptr, used = nullptr; };
struct EpiDevice { EpiSlot s[EPI_SLOTS]; unsigned long long clock = 0; };
std::mutex g_epi_mu;
EpiDevice g_epi[64];

int epi_slot_acquire(const mnrf_field* f, cudaStream_t st, int* slot_out) {
  int dev = 0;
  MNRF_CUDA_OK(cudaGetDevice(&dev));
  MNRF_REQUIRE(dev >= 0 && dev < 64, "field_tc: device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_epi_mu);
  EpiDevice& D = g_epi[dev];
  int hit = -1, lru = 0;
  for (int i = 0; i < EPI_SLOTS; ++i) {
    if (D.s[i].stamp == f->pack_stamp && f->pack_stamp != 0) hit = i;
    if (D.s[i].tick < D.s[lru].tick) lru = i;
  }
  const int k = hit >= 0 ? hit : lru;
  EpiSlot& S = D.s[k];
  if (S.filled == nullptr) {
    MNRF_CUDA_OK(cudaEventCreateWithFlags(&S.filled, cudaEventDisableTiming));
    MNRF_CUDA_OK(cudaEventCreateWithFlags(&S.used, cudaEventDisableTiming));
    MNRF_CUDA_OK(cudaEventRecord(S.used, st));
    MNRF_CUDA_OK(cudaEventRecord(S.filled, st));
  }
  if (hit < 0) {
    MNRF_CUDA_OK(cudaStreamWaitEvent(st, S.used, 0));   // the last kernel that read this slot (any stream) is done first
    MNRF_CUDA_OK(cudaMemcpyToSymbolAsync(c_epi_slots, f->f32 + f->L.epi_tab, sizeof(float) * ET_TOTAL,
                                         sizeof(float) * ET_TOTAL * (size_t)k, cudaMemcpyDeviceToDevice, st));
    MNRF_CUDA_OK(cudaEventRecord(S.filled, st));
    S.stamp = f->pack_stamp;
  } else {
    MNRF_CUDA_OK(cudaStreamWaitEvent(st, S.filled, 0)); // the copy may have been issued on another stream
  }
  S.tick = ++D.clock;
  *slot_out = k;
  return 0;
}

int epi_slot_release(int slot, cudaStream_t st) {
  int dev = 0;
  MNRF_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_epi_mu);
  MNRF_CUDA_OK(cudaEventRecord(g_epi[dev].s[slot].used, st));
  return 0;
}
}  // namespace

int launch_field_tc(const mnrf_field* f, const FieldIO& io, int precision, cudaStream_t st, const FusedComposite* fuse) {
  if (io.n_points <= 0) return 0;
  if (fuse != nullptr) {
    MNRF_REQUIRE(!io.sigma_only && io.normal_out == nullptr && io.rays != nullptr && io.z != nullptr,
                 "field_tc: the fused compositor needs a full ray pass without analytic normals");
    MNRF_REQUIRE(fuse->comp.opacity != nullptr && fuse->work_counter != nullptr, "field_tc: fused compositor outputs missing");
  }
  MNRF_REQUIRE(precision >= 1 && precision <= 3, "field_tc: precision must be 1, 2 or 3");
  if (precision == 2 && io.normal_out != nullptr) precision = 3;  // the analytic-normal chain has no fp8 variant
  MNRF_REQUIRE(io.normal_out == nullptr || !io.sigma_only, "field_tc: analytic normals need the full (non sigma-only) pass");
  MNRF_REQUIRE(io.geo_out == nullptr, "field_tc: geo_feat output needs MNRF_IMPL_FP32");
  MNRF_REQUIRE(io.sigma_only || io.dirbias != nullptr, "field_tc: dirbias missing");
  int num_sms = 0;
  if (first_use_on_device(TAG_FIELD_TC, &num_sms)) {
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_field_tc<2, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
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
  memset(&P.comp, 0, sizeof(P.comp));
  P.n_rays = 0; P.white_back = 0; P.term_eps = 0.f; P.work_counter = nullptr; P.stats = nullptr;
  if (fuse != nullptr) {
    P.comp = fuse->comp; P.n_rays = io.n_points / io.S; P.white_back = fuse->white_back; P.term_eps = fuse->term_eps;
    P.work_counter = fuse->work_counter; P.stats = fuse->stats;
    MNRF_CUDA_OK(cudaMemsetAsync(fuse->work_counter, 0, 2 * sizeof(int), st));   // [0] next ray, [1] CTAs finished
  }
  {
    static const int dbg = getenv("MNRF_TC_DEBUG") != nullptr ? atoi(getenv("MNRF_TC_DEBUG")) : 0;
    P.debug = dbg;   // bit 0: timing experiment, a quarter of the weight bytes per stage (results are garbage);
                     // bit 1: the generic (run-time stage index) issue loop instead of the unrolled one of the tc2 kernels
    P.split = set_tc_split(g_tc_split);   // N-split schedule of the tc2 kernels (issue_split_step)
    P.tc_split = f->tc8 + TC_TOTAL_BYTES;
  }
  P.n_tiles = (io.n_points + TILE_M - 1) / TILE_M;
  const int grid = P.n_tiles < num_sms ? P.n_tiles : num_sms;
  // algorithmic MACs of this launch (unpadded reference layer sizes)
  const double macs = (double)io.n_points * (io.sigma_only ? (double)mnrf_macs_sigma_only() : (double)mnrf_macs_full());
  MNRF_REQUIRE(num_sms > 0, "field_tc: no CUDA device");
  int slot = 0;
  if (epi_slot_acquire(f, st, &slot)) return 1;
  P.slot = slot;
  prof_begin(st);
  const bool normals = io.normal_out != nullptr;
  if (fuse != nullptr) {
    // ray tiles: 4 rays x 32 samples, work handed out dynamically; at most one CTA per SM, no more CTAs than ray quartets
    const int quartets = (P.n_rays + 3) / 4;
    const int fgrid = quartets < num_sms ? quartets : num_sms;
    if (precision == 2) {
      P.tc = f->tc8;
      if (P.split) k_field_tc<2, false, true, true><<<fgrid, NUM_THREADS, SM_TOTAL, st>>>(P);
      else k_field_tc<2, false, true><<<fgrid, NUM_THREADS, SM_TOTAL, st>>>(P);
    }
    else if (precision == 3) k_field_tc<3, false, true><<<fgrid, NUM_THREADS, SM_TOTAL, st>>>(P);
    else k_field_tc<1, false, true><<<fgrid, NUM_THREADS, SM_TOTAL, st>>>(P);
  } else if (precision == 2) {
    P.tc = f->tc8;
    if (P.split) k_field_tc<2, false, false, true><<<grid, NUM_THREADS, SM_TOTAL, st>>>(P);
    else k_field_tc<2, false><<<grid, NUM_THREADS, SM_TOTAL, st>>>(P);
  } else if (precision == 3) {
    if (normals) k_field_tc<3, true><<<grid, NUM_THREADS, SM_TOTAL, st>>>(P);
    else         k_field_tc<3, false><<<grid, NUM_THREADS, SM_TOTAL, st>>>(P);
  } else {
    if (normals) k_field_tc<1, true><<<grid, NUM_THREADS, SM_TOTAL, st>>>(P);
    else         k_field_tc<1, false><<<grid, NUM_THREADS, SM_TOTAL, st>>>(P);
  }
  prof_end(st, 2.0 * macs);
  MNRF_LAUNCH_OK();
  return epi_slot_release(slot, st);
}

}  // namespace mnrf
