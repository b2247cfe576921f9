// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (M=128, K=16) for N = 64/128/256 with both operands in shared
// memory (SS) and with A in tensor memory (TS), on every SM at once.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o mma_bench tools/mma_bench.cu ; prints cycles per MMA.
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
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16); }

template <int TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc) {
  if (TS) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %4};\n\tsetp.ne.b32 p, %3, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %5, p;\n\t}" ::"r"(d), "r"(a), "r"(b), "r"(acc),
                 "r"(DESC_HI), "r"(idesc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
                 "setp.ne.b32 p, %3, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a),
                 "r"(b), "r"(acc), "r"(DESC_HI), "r"(idesc) : "memory");
  }
}

// mode: N, TS flag, whether the other 8 warps hammer shared memory with 16-byte stores (epilogue-like traffic)
template <int N, int TS>
__global__ void __launch_bounds__(384, 1) k(int iters, int traffic, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
  __shared__ int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 1) {
    // A: 128 rows x K=256 at smem 0 (LBO 2048); B: N rows x 256 at smem 64 KB (LBO N*16)
    const uint32_t a0 = desc_lo(sbase, 2048), b0 = desc_lo(sbase + 65536, N * 16);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          mma<TS>(tmem, TS ? (tmem + 256 + j * 8) : (a0 + j * 256), b0 + j * (N * 32 / 16), idesc, j > 0 || it > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      // wait for this batch (keeps at most one batch of 16 in flight + the next being issued would need 2 barriers;
      // to measure pure throughput we only wait every 4th batch)
      if ((it & 3) == 3) {
        uint32_t ok = 0, parity = ((it >> 2) * 4 + 3) & 1;  // phases complete once per commit
        (void)parity;
      }
    }
    // drain: wait until the last commit's phase: iters commits -> parity of (iters-1)
    {
      uint32_t ok = 0;
      const uint32_t parity = (uint32_t)((iters - 1) & 1);
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
    }
    long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; stop = 1; }
    __threadfence_block();
  } else if (warp >= 4 && traffic) {
    // epilogue-like smem traffic: 16-byte stores into the upper part of smem until the MMA warp is done
    volatile int* vs = &stop;
    uint32_t addr = sbase + 140 * 1024 + (threadIdx.x - 128) * 16;
    int n = 0;
    while (!*vs) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(addr + (n & 7) * 4096), "r"(n));
      ++n;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N, int TS>
void run(const char* name, int traffic) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * sms);
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  k<N, TS><<<sms, 384, 200 * 1024>>>(10, traffic, d);
  k<N, TS><<<sms, 384, 200 * 1024>>>(iters, traffic, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += h[i];
  avg /= sms;
  printf("%-28s traffic=%d  %.1f cycles per MMA (M128 N%d K16)  -> %.0f MAC/cycle/SM   [%s]\n", name, traffic,
         avg / (iters * 16.0), N, 128.0 * N * 16 / (avg / (iters * 16.0)), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<256, 0>("SS N=256", 0);
  run<128, 0>("SS N=128", 0);
  run<64, 0>("SS N=64", 0);
  run<256, 1>("TS N=256 (A in TMEM)", 0);
  run<128, 1>("TS N=128 (A in TMEM)", 0);
  run<256, 0>("SS N=256", 1);
  run<128, 0>("SS N=128", 1);
  run<256, 1>("TS N=256 (A in TMEM)", 1);
  return 0;
}
