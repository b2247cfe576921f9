// Micro-benchmark 3: can fp8 (kind::f8f6f4, E4M3) tcgen05.mma carry the two small cross terms of the split-precision product
// into the SAME fp32 TMEM accumulator that a kind::f16 MMA wrote?  Two questions:
//   (1) numerics: D = A16*B16^T (fp16, K16) then D += A8*B8^T (e4m3, K32): is the fp32 accumulator kept at full precision
//       (no Hopper-style ~14-bit accumulation)?  Compared against a double reference, error in ulps of the result.
//   (2) rate: cycles per M128 N256 K32 e4m3 MMA alone and interleaved 1:2 with fp16 K16 MMAs (the "2-pass-equivalent" mix).
// Operands are K-major no-swizzle core matrices (8 rows x 16 bytes): fp16 -> 8 k per core row, fp8 -> 16 k per core row.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint32_t idesc(int n, int afmt, int bfmt) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint32_t a, uint32_t b, uint32_t acc, uint32_t id) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
               "setp.ne.b32 p, %3, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a),
               "r"(b), "r"(acc), "r"(DESC_HI), "r"(id) : "memory");
}
__device__ __forceinline__ void mma_f8(uint32_t d, uint32_t a, uint32_t b, uint32_t acc, uint32_t id) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
               "setp.ne.b32 p, %3, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a),
               "r"(b), "r"(acc), "r"(DESC_HI), "r"(id) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// smem: A16 [128 x 16 fp16] 4 KB @0 | B16 [256 x 16 fp16] 8 KB @4096 | A8 [128 x 32 e4m3] 4 KB @12288 | B8 [256 x 32] 8 KB @16384
// mode 0: numerics (do16, then n8 fp8 accumulations); mode 1: rate of fp8 only; 2: rate of fp16 only; 3: 1 fp16 + 2 fp8 interleaved
__global__ void __launch_bounds__(128, 1) k(int mode, int iters, int n8, const uint8_t* __restrict__ ops, float* dout, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bars[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < 24576 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(ops)[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t a16 = desc_lo(sbase, 2048), b16 = desc_lo(sbase + 4096, 4096);
  const uint32_t a8 = desc_lo(sbase + 12288, 2048), b8 = desc_lo(sbase + 16384, 4096);
  const uint32_t id16 = idesc(256, 0, 0), id8 = idesc(256, 0, 0);
  const uint32_t bar = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    if (mode == 0) {
      mma_f16(tmem, a16, b16, 0, id16);
      for (int i = 0; i < n8; ++i) mma_f8(tmem, a8, b8, 1, id8);
    } else {
      for (int it = 0; it < iters; ++it) {
        if (mode == 1) { mma_f8(tmem, a8, b8, 1, id8); mma_f8(tmem, a8, b8, 1, id8); mma_f8(tmem, a8, b8, 1, id8); }
        if (mode == 2) { mma_f16(tmem, a16, b16, 1, id16); mma_f16(tmem, a16, b16, 1, id16); mma_f16(tmem, a16, b16, 1, id16); }
        if (mode == 3) { mma_f16(tmem, a16, b16, 1, id16); mma_f8(tmem, a8, b8, 1, id8); mma_f16(tmem, a16, b16, 1, id16); }
        if (mode == 4) { mma_f16(tmem, a16, b16, 1, id16); mma_f16(tmem, a16, b16, 1, id16); mma_f8(tmem, a8, b8, 1, id8); }
      }
    }
    commit(bar);
    while (!mbar_try(bar, 0)) {}
    cyc[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (mode == 0) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 256; c0 += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) dout[row * 256 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static float e4m3_to_float(uint8_t v) {
  __half_raw h = __nv_cvt_fp8_to_halfraw(v, __NV_E4M3);
  __half hh; memcpy(&hh, &h, 2);
  return __half2float(hh);
}

int main() {
  std::vector<uint8_t> ops(24576, 0);
  std::vector<double> A16(128 * 16), B16(256 * 16), A8(128 * 32), B8(256 * 32);
  srand(1);
  auto rnd = []() { return (rand() / (double)RAND_MAX) * 2.0 - 1.0; };
  // case table: scale of the fp16 part vs the fp8 part
  const double s16[] = {1.0, 1024.0, 1.0, 32768.0};
  const double s8[] = {1.0, 1.0, 1.0 / 1024.0, 1.0 / 16.0};
  float* dout; long long* cyc; uint8_t* dops;
  cudaMalloc(&dout, 128 * 256 * 4); cudaMalloc(&cyc, 8); cudaMalloc(&dops, 24576);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int cs = 0; cs < 4; ++cs) {
    for (int n8 = 0; n8 <= 8; n8 += 4) {
      for (int r = 0; r < 128; ++r) for (int kk = 0; kk < 16; ++kk) {
        __half h = __float2half((float)(rnd() * 4.0)); A16[r * 16 + kk] = __half2float(h);
        memcpy(&ops[0 + (kk / 8) * 2048 + (r / 8) * 128 + (r % 8) * 16 + (kk % 8) * 2], &h, 2);
      }
      for (int r = 0; r < 256; ++r) for (int kk = 0; kk < 16; ++kk) {
        __half h = __float2half((float)(rnd() * s16[cs])); B16[r * 16 + kk] = __half2float(h);
        memcpy(&ops[4096 + (kk / 8) * 4096 + (r / 8) * 128 + (r % 8) * 16 + (kk % 8) * 2], &h, 2);
      }
      for (int r = 0; r < 128; ++r) for (int kk = 0; kk < 32; ++kk) {
        uint8_t v = __nv_cvt_float_to_fp8((float)(rnd() * 4.0), __NV_SATFINITE, __NV_E4M3); A8[r * 32 + kk] = e4m3_to_float(v);
        ops[12288 + (kk / 16) * 2048 + (r / 8) * 128 + (r % 8) * 16 + (kk % 16)] = v;
      }
      for (int r = 0; r < 256; ++r) for (int kk = 0; kk < 32; ++kk) {
        uint8_t v = __nv_cvt_float_to_fp8((float)(rnd() * s8[cs]), __NV_SATFINITE, __NV_E4M3); B8[r * 32 + kk] = e4m3_to_float(v);
        ops[16384 + (kk / 16) * 4096 + (r / 8) * 128 + (r % 8) * 16 + (kk % 16)] = v;
      }
      cudaMemcpy(dops, ops.data(), 24576, cudaMemcpyHostToDevice);
      k<<<1, 128, 32768>>>(0, 0, n8, dops, dout, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> D(128 * 256);
      cudaMemcpy(D.data(), dout, D.size() * 4, cudaMemcpyDeviceToHost);
      double max_ulp = 0, max_rel8 = 0, sum_ulp = 0;
      for (int i = 0; i < 128; ++i) for (int j = 0; j < 256; ++j) {
        double m = 0, c = 0;
        for (int kk = 0; kk < 16; ++kk) m += A16[i * 16 + kk] * B16[j * 16 + kk];
        for (int kk = 0; kk < 32; ++kk) c += A8[i * 32 + kk] * B8[j * 32 + kk];
        const double ref = m + n8 * c;
        const double ulp = ldexp(1.0, (int)floor(log2(fabs(ref) + 1e-300)) - 23);
        const double err = fabs((double)D[i * 256 + j] - ref);
        max_ulp = fmax(max_ulp, err / ulp); sum_ulp += err / ulp;
        if (n8) max_rel8 = fmax(max_rel8, err / (fabs(n8 * c) + 1e-30));
      }
      printf("numerics case %d (|B16|~%g, |B8|~%g) fp8 accumulations=%d: max err %.2f ulp of result, mean %.3f ulp  [%s]\n", cs,
             s16[cs], s8[cs], n8, max_ulp, sum_ulp / (128 * 256), cudaGetErrorString(e));
    }
  }
  const char* names[] = {"", "fp8 K32 only", "fp16 K16 only", "f16,f8,f16", "f16,f16,f8"};
  for (int mode = 1; mode <= 4; ++mode) {
    const int iters = 20000;
    k<<<1, 128, 32768>>>(mode, 100, 0, dops, dout, cyc);
    k<<<1, 128, 32768>>>(mode, iters, 0, dops, dout, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("rate mode %d [%s]: %.1f cycles per MMA (M128 N256)  [%s]\n", mode, names[mode], h / (iters * 3.0), cudaGetErrorString(e));
  }
  return 0;
}
