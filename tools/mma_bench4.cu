// Micro-benchmark 4: the weight ring of k_field_tc in isolation (no epilogue): one producer thread streams B-operand stages from
// L2 with cp.async.bulk into an mbarrier ring, one issuer thread waits / issues the tc2 MMA mix / commits the stage back.
// All SMs run at once (same L2 pressure as the real kernel).  Questions:
//   (1) what does the ring cost per K32 chunk as a function of stage size and count (4 x 16 KB today vs 8 x 8 KB)?
//   (2) how fast can ONE thread issue N = 128 MMAs (64 cycles of pipe each) with a wait + commit per stage -- the kernel's loop
//       (run-time stage index) against a loop unrolled over the ring (compile-time stage addresses)?
// tc2 mix per K32 chunk: N = 256: [W_hi fp16 16 KB] 2 x f16 K16 + [e4m3 | e4m3 16 KB] 2 x f8 K32 = 512 pipe cycles / 32 KB;
//                        N = 128: [fp16 8 KB | e4m3 4 KB | e4m3 4 KB] 2 x f16 + 2 x f8 = 256 pipe cycles / 16 KB.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_bench4 mma_bench4.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16); }
template <int N>
__device__ __forceinline__ void mma_f16(uint32_t d, uint32_t a, uint32_t b, uint32_t acc) {
  constexpr uint32_t id = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
               "setp.ne.b32 p, %3, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a),
               "r"(b), "r"(acc), "r"(DESC_HI), "r"(id) : "memory");
}
template <int N>
__device__ __forceinline__ void mma_f8(uint32_t d, uint32_t a, uint32_t b, uint32_t acc) {
  constexpr uint32_t id = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
               "setp.ne.b32 p, %3, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a),
               "r"(b), "r"(acc), "r"(DESC_HI), "r"(id) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) { while (!mbar_test(bar, parity)) {} }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// wm = 0: poll with test_wait (never suspends); 1: try_wait (the hardware may park the warp until the phase flips)
__device__ __forceinline__ void mbar_w(uint32_t bar, uint32_t parity, int wm) {
  if (wm) { while (!mbar_try(bar, parity)) {} } else { while (!mbar_test(bar, parity)) {} }
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

constexpr uint32_t SM_A_HI = 0, SM_A_LO = 65536, SM_WST = 131072;   // ring: up to 64 KB behind the A operand
constexpr uint32_t SM_TOTAL = SM_WST + 65536;

// the MMAs of ring stage `role` (position of the stage inside its K32 chunk) -- addresses as in field_tc.cu
template <int N, int SB>
__device__ __forceinline__ void issue_stage(uint32_t d, uint32_t ah, uint32_t a8, uint32_t a8r, uint32_t st_addr, int role, uint32_t acc) {
  if (N == 256 && SB == 16384) {
    const uint32_t wb = desc_lo(st_addr, 4096);
    if (role == 0) { mma_f16<256>(d, ah, wb, acc); mma_f16<256>(d, ah + 256u, wb + 512u, 1u); }
    else           { mma_f8<256>(d, a8r, wb, 1u);  mma_f8<256>(d, a8, wb + 512u, 1u); }
  } else if (N == 256 && SB == 8192) {
    const uint32_t wb = desc_lo(st_addr, 4096);
    if (role == 0) mma_f16<256>(d, ah, wb, acc);
    else if (role == 1) mma_f16<256>(d, ah + 256u, wb, 1u);
    else if (role == 2) mma_f8<256>(d, a8r, wb, 1u);
    else mma_f8<256>(d, a8, wb, 1u);
  } else if (N == 128 && SB == 16384) {
    const uint32_t wb = desc_lo(st_addr, 2048);
    mma_f16<128>(d, ah, wb, acc); mma_f16<128>(d, ah + 256u, wb + 256u, 1u);
    mma_f8<128>(d, a8r, wb + 512u, 1u); mma_f8<128>(d, a8, wb + 768u, 1u);
  } else {  // N == 128, 8 KB stages
    const uint32_t wb = desc_lo(st_addr, 2048);
    if (role == 0) { mma_f16<128>(d, ah, wb, acc); mma_f16<128>(d, ah + 256u, wb + 256u, 1u); }
    else           { mma_f8<128>(d, a8r, wb, 1u);  mma_f8<128>(d, a8, wb + 256u, 1u); }
  }
}

// STYLE 0: the kernel's loop (one elected thread runs everything, run-time stage index)
// STYLE 1: the same thread, ring unrolled (compile-time stage addresses and barrier offsets)
// STYLE 2: like 1, but the whole warp runs the loop and waits; only MMAs + commit are elected (uniform-datapath friendly)
// STYLE 3: the whole warp runs the loop (uniform control flow and descriptors), the leader lane is elected ONCE and only the
//          tcgen05.mma / commit instructions are predicated on it
// busy: warps 4..19 (four per scheduler, like the kernel's epilogue warps) run an ALU loop with 4 independent chains
template <int N, int SB, int NST, int STYLE, int PROD>
__global__ void __launch_bounds__(640, 1) k(int chunks, const uint8_t* __restrict__ wsrc, uint32_t wbytes, long long* out, int busy,
                                            float* sink, int wp, int wi, int wm, int ym) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ int stop;
  __shared__ uint64_t bars[2 * 8 + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  constexpr int SPC = (N == 256 ? 32768 : 16384) / SB;   // stages per K32 chunk
  static_assert(NST % SPC == 0, "ring must hold whole chunks for the unrolled style");
  for (int i = threadIdx.x; i < (int)(SM_TOTAL / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    stop = 0;
    for (int i = 0; i < 17; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == wi) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), done = smem_u32(&bars[16]);
  const int total_stages = chunks * SPC;

  if (warp == wp) {
    if (PROD && elect_one()) {
      uint32_t stage = 0, phase = 0, off = (blockIdx.x * 65536u) % wbytes;
      for (int s = 0; s < total_stages; ++s) {
        mbar_w(empty0 + 8u * stage, phase ^ 1u, wm);
        const uint32_t fb = full0 + 8u * stage, dst = sbase + SM_WST + stage * SB;
        expect_tx(fb, SB);
#pragma unroll
        for (int piece = 0; piece < SB / 4096; ++piece) bulk_g2s(dst + piece * 4096u, wsrc + off + piece * 4096u, 4096u, fb);
        off += SB;
        if (off + SB > wbytes) off = 0;
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == wi) {
    const uint32_t dl_a_hi = desc_lo(sbase + SM_A_HI, 2048), dl_a8 = desc_lo(sbase + SM_A_LO, 2048),
                   dl_a8r = desc_lo(sbase + SM_A_LO + 32768u, 2048);
    long long t0 = 0, t1 = 0;
    if (STYLE == 0) {
      if (elect_one()) {
        t0 = clock64();
        uint32_t stage = 0, phase = 0;
        for (int c = 0; c < chunks; ++c) {
          const uint32_t ka = (uint32_t)(c & 7);
          const uint32_t ah = dl_a_hi + ka * 512u, a8 = dl_a8 + ka * 256u, a8r = dl_a8r + ka * 256u;
          const uint32_t d = tmem + (N == 128 ? (uint32_t)((c >> 3) & 1) * 128u : 0u);
          for (int r = 0; r < SPC; ++r) {
            if (PROD) mbar_w(full0 + 8u * stage, phase, wm);
            fence_after();
            issue_stage<N, SB>(d, ah, a8, a8r, sbase + SM_WST + stage * SB, r, (c & 7) != 0 || r != 0);
            commit(empty0 + 8u * stage);
            if (++stage == NST) { stage = 0; phase ^= 1u; }
          }
        }
        commit(done);
        mbar_spin(done, 0);
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
      }
    } else {
      const bool leader = (STYLE == 1) ? elect_one() : true;   // styles 2, 3: every lane runs the loop
      const bool lead3 = elect_one();
      if (leader) {
        t0 = clock64();
        uint32_t phase = 0;
        constexpr int CPR = NST / SPC;   // chunks per trip around the ring
        for (int c0 = 0; c0 < chunks; c0 += CPR) {
#pragma unroll
          for (int cc = 0; cc < CPR; ++cc) {
            const int c = c0 + cc;
            const uint32_t ka = (uint32_t)(c & 7);
            const uint32_t ah = dl_a_hi + ka * 512u, a8 = dl_a8 + ka * 256u, a8r = dl_a8r + ka * 256u;
            const uint32_t d = tmem + (N == 128 ? (uint32_t)((c >> 3) & 1) * 128u : 0u);
#pragma unroll
            for (int r = 0; r < SPC; ++r) {
              constexpr int dummy = 0; (void)dummy;
              const int stage = cc * SPC + r;   // compile-time after unrolling
              if (PROD) mbar_w(full0 + 8u * stage, phase, wm);
              fence_after();
              if (STYLE == 3) {
                if (lead3) {
                  issue_stage<N, SB>(d, ah, a8, a8r, sbase + SM_WST + stage * SB, r, (c & 7) != 0 || r != 0);
                  commit(empty0 + 8u * stage);
                }
              } else if (STYLE == 2) {
                if (elect_one()) {
                  issue_stage<N, SB>(d, ah, a8, a8r, sbase + SM_WST + stage * SB, r, (c & 7) != 0 || r != 0);
                  commit(empty0 + 8u * stage);
                }
                __syncwarp();
              } else {
                issue_stage<N, SB>(d, ah, a8, a8r, sbase + SM_WST + stage * SB, r, (c & 7) != 0 || r != 0);
                commit(empty0 + 8u * stage);
              }
            }
          }
          phase ^= 1u;
        }
        if (STYLE == 2) { if (elect_one()) commit(done); __syncwarp(); }
        else if (STYLE == 3) { if (lead3) commit(done); }
        else commit(done);
        mbar_spin(done, 0);
        t1 = clock64();
        if (STYLE == 1 || lane == 0) out[blockIdx.x] = t1 - t0;
      }
    }
    __syncwarp();
    if (lane == 0) { volatile int* vs = &stop; *vs = 1; }
  } else if (busy && (wi < 4 ? warp >= 4 : warp < 16)) {
    volatile int* vs = &stop;
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
    while (!*vs) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 0.9999f, 0.25f); a2 = fmaf(a2, 1.0002f, 0.125f); a3 = fmaf(a3, 0.9998f, 1.f); }
      if (ym == 1) __nanosleep(0);
      else if (ym == 2) asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
    }
    if (a0 + a1 + a2 + a3 == 12345.f) sink[threadIdx.x] = a0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == wi) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static uint8_t* g_w = nullptr;
static float* g_sink = nullptr;
constexpr uint32_t WBYTES = 4u << 20;

template <int N, int SB, int NST, int STYLE, int PROD>
void run(int busy = 0, int wp = 0, int wi = 1, int wm = 0, int ym = 0) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * sms);
  cudaMemset(d, 0, sizeof(long long) * sms);
  auto kern = k<N, SB, NST, STYLE, PROD>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
  const int chunks = 8 * 400;   // multiple of every ring trip
  kern<<<sms, 640, SM_TOTAL>>>(64, g_w, WBYTES, d, busy, g_sink, wp, wi, wm, ym);
  kern<<<sms, 640, SM_TOTAL>>>(chunks, g_w, WBYTES, d, busy, g_sink, wp, wi, wm, ym);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0, mx = 0;
  for (int i = 0; i < sms; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
  avg /= sms;
  const double ideal = N == 256 ? 512.0 : 256.0;
  printf("N=%3d stage=%5d B x %d  style=%d producer=%d busy=%d roles=warp %d,%d wait=%s yield=%d : %7.1f cycles per K32 chunk (max SM %7.1f), ideal %3.0f -> %.2fx   [%s]\n", N, SB, NST,
         STYLE, PROD, busy, wp, wi, wm ? "try" : "test", ym, avg / chunks, mx / chunks, ideal, avg / chunks / ideal, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  cudaMalloc(&g_w, WBYTES);
  cudaMemset(g_w, 0, WBYTES);
  cudaMalloc(&g_sink, 4096);
  // contention study: ALU-saturating warps (four per scheduler) against the producer / issuer pair
  for (int wm = 0; wm < 2; ++wm)
    for (int ym = 0; ym < 3; ++ym) {
      run<256, 16384, 4, 0, 1>(1, 0, 1, wm, ym);
      run<128, 16384, 4, 1, 1>(1, 0, 1, wm, ym);
      run<256, 16384, 4, 0, 1>(1, 16, 17, wm, ym);
      run<128, 16384, 4, 1, 1>(1, 16, 17, wm, ym);
    }
  run<256, 16384, 4, 0, 1>(0, 0, 1, 1, 0);
  run<128, 16384, 4, 1, 1>(0, 0, 1, 1, 0);
  return 0;
}
