// Micro-benchmark 2: tcgen05.mma (M128 N128 K16, SS) issue/execute rate under the interference the field kernel creates:
//   bit 0: a producer warp streams 16 KB cp.async.bulk copies from L2 into a 4-stage ring (42 B/clk needed by the kernel)
//   bit 1: 8 "epilogue" warps loop on tcgen05.ld x32 + 16-byte shared stores
//   bit 2: the issuing warp does an mbarrier try_wait (on a completed barrier) and a tcgen05.commit every 6 MMAs
//   bit 3: the 8 "epilogue" warps loop on FFMA/cvt ALU work only (issue-slot pressure, no memory)
// prints cycles per MMA for every combination of interest.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | ((2048u >> 4) << 16); }
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint32_t b, uint32_t acc) {
  const uint32_t idesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
               "setp.ne.b32 p, %3, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a),
               "r"(b), "r"(acc), "r"(DESC_HI), "r"(idesc) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(384, 1) k(int iters, int mode, const uint8_t* __restrict__ wsrc, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bars[16];
  __shared__ int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  for (int i = threadIdx.x; i < 224 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  volatile int* vs = &stop;
  if (warp == 1) {
    // operands like the field kernel: A hi at 0, A lo at 64 KB (128 rows x 256), weight stages at 160 KB
    const uint32_t ah = desc_lo(sbase), al = desc_lo(sbase + 65536), w = desc_lo(sbase + 163840);
    const uint32_t done_bar = smem_u32(&bars[8]);
    // pre-complete a barrier for the try_wait test
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[9])) : "memory");
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t st = (uint32_t)(it & 3) * 1024u, ka = (uint32_t)(it & 7) * 512u;
      if (mode & 4) { while (!mbar_try(smem_u32(&bars[9]), 0)) {} asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      if (elect_one()) {
        mma(tmem, ah + ka, w + st, 1);
        mma(tmem, al + ka, w + st, 1);
        mma(tmem, ah + ka + 256, w + st + 256, 1);
        mma(tmem, al + ka + 256, w + st + 256, 1);
        mma(tmem, ah + ka, w + st + 512, 1);
        mma(tmem, ah + ka + 256, w + st + 768, 1);
        if (mode & 4) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[10])) : "memory");
      }
      __syncwarp();
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done_bar) : "memory");
    __syncwarp();
    while (!mbar_try(done_bar, 0)) {}
    long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; stop = 1; }
    __threadfence_block();
  } else if (warp == 0 && (mode & 1)) {
    // weight streaming: 4-stage ring of 16 KB bulk copies, each waits for its own previous copy
    if (lane == 0) {
      uint32_t ph[4] = {0, 0, 0, 0};
      int issued = 0;
      size_t off = 0;
      while (!*vs) {
        const int s = issued & 3;
        const uint32_t bar = smem_u32(&bars[s]);
        if (issued >= 4) { while (!mbar_try(bar, ph[s]) && !*vs) {} ph[s] ^= 1u; }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16384) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         sbase + 163840 + s * 16384), "l"(wsrc + off), "r"(16384), "r"(bar) : "memory");
        off = (off + 16384) % (2490368);
        ++issued;
      }
      // drain outstanding copies before exit
      for (int s = 0; s < 4; ++s) { if (issued > s) { int tries = 0; while (!mbar_try(smem_u32(&bars[s]), ph[s]) && ++tries < 100000) {} } }
    }
  } else if (warp >= 4 && (mode & 2)) {
    const int q = warp & 3;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 256;
    const uint32_t saddr = sbase + 131072 + (threadIdx.x - 128) * 16;
    uint32_t r[32];
    float acc = 0.f;
    while (!*vs) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr + (j & 3) * 4096), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                     "r"(r[4 * j + 2]), "r"(r[4 * j + 3]));
      acc += __uint_as_float(r[0]);
    }
    if (acc == 123.456f) sink[0] = acc;
  } else if (warp >= 4 && (mode & 8)) {
    float a = threadIdx.x, b = 1.0001f, c = 0.5f;
    while (!*vs) {
#pragma unroll
      for (int j = 0; j < 64; ++j) { a = fmaf(a, b, c); c = fmaf(c, b, a); }
    }
    if (a == 123.456f) sink[0] = a + c;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  float* sink;
  uint8_t* w;
  cudaMalloc(&d, sizeof(long long) * sms);
  cudaMalloc(&sink, 16);
  cudaMalloc(&w, 4 << 20);
  cudaMemset(w, 0, 4 << 20);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  const int iters = 20000;
  const int modes[] = {0, 4, 1, 2, 8, 5, 6, 3, 7, 15};
  for (int m : modes) {
    k<<<sms, 384, 224 * 1024>>>(100, m, w, d, sink);
    k<<<sms, 384, 224 * 1024>>>(iters, m, w, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += h[i];
    avg /= sms;
    printf("mode %2d [%s%s%s%s]: %.1f cycles per MMA  [%s]\n", m, (m & 1) ? " tma" : "", (m & 2) ? " ldtm+sts" : "",
           (m & 4) ? " wait+commit" : "", (m & 8) ? " alu" : "", avg / (iters * 6.0), cudaGetErrorString(e));
  }
  return 0;
}
