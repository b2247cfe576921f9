// Tensor-core GEMMs of the training path (train.cu): tcgen05.mma kind::tf32 with 3x split operands (x = hi + lo, both tf32;
// hi*hi + lo*hi + hi*lo accumulated in fp32 TMEM) -- fp32-grade products without any scaling, because tf32 keeps the fp32
// exponent range (gradients span many decades; the fp16 split of field_tc.cu would need per-row scales).
//
// k_gemm_tc_nn   C[M,N] = epi([A0 | A1][M,K] * W^T)      activations x packed weight blobs (pack.cu k_pack_t32)
//   persistent CTA per SM, 128-row tiles, 18 warps:
//     warp 0       B producer: cp.async.bulk of the pre-packed tf32 hi|lo blob of a K16 chunk into the stage
//     warp 1       MMA issuer (one elected thread): 6 MMAs (2 K8 steps x 3 passes) per stage, accumulators double-buffered in TMEM
//     warps 2..9   A loaders: global fp32 -> registers (4 chunks in flight per thread) -> tf32 hi/lo -> UMMA K-major core matrices
//     warps 10..17 epilogue: tcgen05.ld -> bias / per-ray term / rank-1 term / ReLU / ReLU-mask / accumulate -> global fp32
// k_gemm_tc_tn   dW[NA,NB] += A[P,NA]^T B[P,NB]          weight gradients: the reduction runs over the points
//     both operands are transposed on the fly by the loaders (a core-matrix row = one feature x 4 consecutive points), the
//     whole dW tile (<= 256 x 256) lives in TMEM for the CTA's slab of points and is flushed once with red.global.add.
//
// Per-unit cost: a 128 x 256 x 256 layer tile = 32 K8 steps x 3 passes x 128 cycles = 12.3k tensor cycles against
// 128 KB read + 128 KB written (+128 KB mask) of HBM traffic: the unfused layer GEMMs sit at the HBM/tensor balance point.
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mnrf {
namespace {

using namespace tcx;

// ------------------------------------------------------------------------------------------------ NN
constexpr int NN_STAGES = 3;   // UMMA operand stages (A hi/lo + B hi/lo of one K16 chunk)
constexpr int NN_RAW = 6;      // raw fp32 A chunks (128 rows x 16 columns, 8 KB) in flight behind them: TMA tensor-map loads
// A operand: K-adjacent core matrices are LBO = 2048 + 32 bytes apart (not 2048): with the loaders' mapping (4 lanes = the four
// 16-byte K quads of a row, 8 rows per instruction -> 64-byte runs of global memory) the padding makes the st.shared.v4
// phases bank-conflict free.
constexpr uint32_t NN_A_LBO = 2080;
constexpr uint32_t NN_A_PART = 4 * NN_A_LBO;   // 8320: 128 rows x 16 tf32 (padded)
constexpr uint32_t NN_B_PART = 16384;          // up to 256 rows x 16 tf32
constexpr uint32_t NN_STAGE = 2 * NN_A_PART + 2 * NN_B_PART;  // 49408
constexpr uint32_t NN_SM_RAW = NN_STAGES * NN_STAGE;          // 148224: raw ring
constexpr uint32_t NN_RAW_SLOT = 8192;
constexpr uint32_t NN_SM_EPI = NN_SM_RAW + NN_RAW * NN_RAW_SLOT;  // 197376: 8 epilogue warps x 4 KB transposition tiles
constexpr uint32_t NN_SM_BAR = NN_SM_EPI + 8 * 4096;          // 230144
constexpr uint32_t NN_SM_TOTAL = NN_SM_BAR + 256;
constexpr int NN_THREADS = 640;  // warp 0 B producer | 1 MMA | 2 A producer (TMA) | 3 idle | 4..11 converters | 12..19 epilogue
constexpr int NN_LOADER_WARP0 = 4, NN_EPI_WARP0 = 12;
// barrier slots
constexpr int NB_A_FULL = 0, NB_B_FULL = 4, NB_EMPTY = 8, NB_ACC_FULL = 12, NB_ACC_EMPTY = 14, NB_TMEM_SLOT = 16, NB_RAW_FULL = 18,
              NB_RAW_EMPTY = 24;

struct NNParams {
  const float* A0; int lda0; int K0;
  const float* A1; int lda1;
  const uint8_t* blob;   // this step's blobs: [kc][hi | lo], N*64 bytes each
  float* C; int ldc;
  int M, N, K, n_tiles;
  GemmEpi e;
  int one_pass;  // speed mode: single tf32 pass (hi parts only; 10-bit mantissa operands) instead of the 3-pass split
  int dbg;  // bring-up: 1 = no global loads of A, 2 = no conversion/stores of A, 4 = no MMAs, 8 = no epilogue global traffic, 16 = no B copies
};

__device__ __forceinline__ float act_apply(float v, int act, float m) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return m > 0.f ? v : 0.f;
  if (act == 3) return v > 0.f ? v : 0.01f * v;
  return v;
}

// Epilogue of a 32 x 32 block held in the warp's transposition tile: lane -> (row 4*i + rsub, columns col..col+3), i = 0..7.
// All global operands (C for accumulate, ReLU mask, per-ray term, rank-1 row factor) are fetched for four rows at a time BEFORE
// the dependent arithmetic and stores, so that their latencies overlap instead of serialising the 8 iterations.
// MODE: 0 = generic (every feature decided at run time), 1 = bias + ReLU (forward trunk), 2 = ReLU-mask only (normal chain,
// dgrad, tangent pass), 3 = plain linear with optional bias (final, normal_net.0, dF, PE gradients) -- the hot flavours are
// compiled without the unused operand streams.
template <int MODE>
__device__ __forceinline__ void nn_epi_block(const NNParams& P, uint32_t tb, long long row_base, int rsub, int c4, int col,
                                             const uint32_t (&mbits)[8], bool have_bits) {
  const GemmEpi& e = P.e;
  const bool use_acc = MODE == 0 && e.accumulate, use_bias = MODE == 1 || ((MODE == 0 || MODE == 3) && e.bias != nullptr);
  const bool use_rb = MODE == 0 && e.rowbias != nullptr, use_rv = MODE == 0 && e.rvec != nullptr;
  const int act = MODE == 1 ? 1 : (MODE == 2 ? 2 : (MODE == 3 ? 0 : e.act));
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), cv = bias;
  if (use_bias) bias = __ldg(reinterpret_cast<const float4*>(e.bias + col));
  if (use_rv) cv = __ldg(reinterpret_cast<const float4*>(e.cvec + col));
  // rows handled by this lane: row_base + rsub + 4*i; element offsets advance by 4 rows per i
  const long long r0 = row_base + rsub;
  float* cp = P.C + (size_t)r0 * P.ldc + col;
  const bool bitmask = act == 2 && have_bits;
  const float* mp = (act == 2 && !bitmask) ? e.mask + (size_t)r0 * e.ld_mask + col : nullptr;
  const int bsh = col & 31;
  const size_t cstep = (size_t)4 * P.ldc, mstep = (size_t)4 * e.ld_mask;
  const int rows_left = (int)min((long long)32, P.M - row_base) - rsub;  // rows r0 + 4*i with 4*i < rows_left exist
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float4 cin[4], mk[4], rbv[4];
    float rv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * half + k;
      const bool ok = 4 * i < rows_left;
      cin[k] = make_float4(0.f, 0.f, 0.f, 0.f); mk[k] = cin[k]; rbv[k] = cin[k]; rv[k] = 0.f;
      if (ok) {
        if (use_acc) cin[k] = *reinterpret_cast<const float4*>(cp + i * cstep);
        if (bitmask) {
          const uint32_t wb = mbits[i] >> bsh;
          mk[k] = make_float4((wb & 1u) ? 1.f : 0.f, (wb & 2u) ? 1.f : 0.f, (wb & 4u) ? 1.f : 0.f, (wb & 8u) ? 1.f : 0.f);
        } else if (act == 2) {
          mk[k] = *reinterpret_cast<const float4*>(mp + i * mstep);
        }
        if (use_rb) rbv[k] = *reinterpret_cast<const float4*>(e.rowbias + (size_t)((r0 + 4 * i) / e.rb_div) * e.ld_rb + col);
        if (use_rv) rv[k] = e.rvec[(size_t)(r0 + 4 * i) * e.ld_rvec];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * half + k;
      const int r = 4 * i + rsub;
      const float4 a = ld_shared_v4(tb + (uint32_t)r * 128u + (uint32_t)((c4 ^ (r & 7)) * 16));
      float v[4] = {a.x, a.y, a.z, a.w};
      if (use_acc) { v[0] += cin[k].x; v[1] += cin[k].y; v[2] += cin[k].z; v[3] += cin[k].w; }
      if (use_bias) { v[0] += bias.x; v[1] += bias.y; v[2] += bias.z; v[3] += bias.w; }
      if (use_rb) { v[0] += rbv[k].x; v[1] += rbv[k].y; v[2] += rbv[k].z; v[3] += rbv[k].w; }
      if (use_rv) { v[0] = fmaf(rv[k], cv.x, v[0]); v[1] = fmaf(rv[k], cv.y, v[1]); v[2] = fmaf(rv[k], cv.z, v[2]); v[3] = fmaf(rv[k], cv.w, v[3]); }
      float4 o;
      o.x = act_apply(v[0], act, mk[k].x); o.y = act_apply(v[1], act, mk[k].y);
      o.z = act_apply(v[2], act, mk[k].z); o.w = act_apply(v[3], act, mk[k].w);
      if (4 * i < rows_left) *reinterpret_cast<float4*>(cp + i * cstep) = o;
      if (MODE == 1 && e.bits_out != nullptr) {
        // relu' bits of this row's 32-column chunk: 4 bits per lane, OR-combined over the 8 lanes that share the row
        uint32_t nib = ((o.x > 0.f ? 1u : 0u) | (o.y > 0.f ? 2u : 0u) | (o.z > 0.f ? 4u : 0u) | (o.w > 0.f ? 8u : 0u)) << (4 * c4);
        nib |= __shfl_xor_sync(0xffffffffu, nib, 1);
        nib |= __shfl_xor_sync(0xffffffffu, nib, 2);
        nib |= __shfl_xor_sync(0xffffffffu, nib, 4);
        if (c4 == 0 && 4 * i < rows_left) e.bits_out[(size_t)(r0 + 4 * i) * 8 + (col >> 5)] = nib;
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(NN_THREADS, 1)
k_gemm_tc_nn(const NNParams P, const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bars = sbase + NN_SM_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + NN_SM_BAR + 8 * NB_TMEM_SLOT);
  const int nkc = P.K >> 4;
  const uint32_t N = (uint32_t)P.N;

  if (threadIdx.x == 0) {
    if (sbase & 127u) { printf("mnrf train_tc: unaligned dynamic smem base %u\n", sbase); __trap(); }
    for (int i = 0; i < NN_STAGES; ++i) { mbar_init(bar(NB_A_FULL + i), 8); mbar_init(bar(NB_B_FULL + i), 1); mbar_init(bar(NB_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(NB_ACC_FULL + i), 1); mbar_init(bar(NB_ACC_EMPTY + i), 8); }
    for (int i = 0; i < NN_RAW; ++i) { mbar_init(bar(NB_RAW_FULL + i), 1); mbar_init(bar(NB_RAW_EMPTY + i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int my_tiles = ((int)blockIdx.x < P.n_tiles) ? (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ------------------------------ B producer ------------------------------
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t chunk_bytes = 2u * N * 64u;
      const uint32_t bytes = P.one_pass ? N * 64u : chunk_bytes;
      for (int t = 0; t < my_tiles; ++t) {
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_wait(bar(NB_EMPTY + stage), phase ^ 1u);
          const uint32_t dst = sbase + stage * NN_STAGE + 2 * NN_A_PART;
          const uint32_t fb = bar(NB_B_FULL + stage);
          const uint8_t* src = P.blob + (size_t)kc * chunk_bytes;
          if (P.dbg & 16) {
            mbar_expect_tx(fb, 0);
          } else {
            mbar_expect_tx(fb, bytes);
            for (uint32_t o = 0; o < bytes; o += 8192u) bulk_g2s(dst + o, src + o, min(8192u, bytes - o), fb);
          }
          if (++stage == NN_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t idesc = idesc_tf32(N);
      const uint32_t bstep = 2u * N;  // one K8 step of B = 2 core columns = 2 * N*16 bytes, in 16-byte units
      for (int t = 0; t < my_tiles; ++t) {
        const int buf = t & 1;
        if (t >= 2) mbar_wait(bar(NB_ACC_EMPTY + buf), (uint32_t)(((t >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)buf * 256u;
        uint32_t accumulate = 0;
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_wait(bar(NB_A_FULL + stage), phase);
          mbar_wait(bar(NB_B_FULL + stage), phase);
          tc_fence_after();
          const uint32_t sa = sbase + stage * NN_STAGE;
          const uint32_t a_hi = desc_lo(sa, NN_A_LBO), a_lo = desc_lo(sa + NN_A_PART, NN_A_LBO);
          const uint32_t b_hi = desc_lo(sa + 2 * NN_A_PART, N * 16u), b_lo = desc_lo(sa + 2 * NN_A_PART + N * 64u, N * 16u);
          constexpr uint32_t astep = 2u * NN_A_LBO / 16u;  // one K8 step of A = 2 core columns
          if (!(P.dbg & 4)) {
#pragma unroll
            for (uint32_t j = 0; j < 2; ++j) {
              mma_tf32(d_tmem, a_hi + j * astep, b_hi + j * bstep, idesc, accumulate);
              if (!P.one_pass) {
                mma_tf32(d_tmem, a_lo + j * astep, b_hi + j * bstep, idesc, 1u);
                mma_tf32(d_tmem, a_hi + j * astep, b_lo + j * bstep, idesc, 1u);
              }
              accumulate = 1u;
            }
          }
          tc_commit(bar(NB_EMPTY + stage));
          if (++stage == NN_STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(bar(NB_ACC_FULL + buf));
      }
    }
  } else if (warp == 2) {
    // ------------------------------ A producer: one tensor-map load (128 rows x 16 fp32 columns) per K chunk ------------------------------
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int row = (int)(((long long)blockIdx.x + (long long)t * gridDim.x) * 128);
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_wait(bar(NB_RAW_EMPTY + slot), phase ^ 1u);
          const uint32_t fb = bar(NB_RAW_FULL + slot);
          const int k = kc * 16;
          mbar_expect_tx(fb, NN_RAW_SLOT);
          if (P.dbg & 1) {
            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(fb), "r"(NN_RAW_SLOT) : "memory");
          } else if (k < P.K0) {
            tma_load_2d(sbase + NN_SM_RAW + slot * NN_RAW_SLOT, &tmA0, k, row, fb);   // rows past M are zero-filled by the TMA unit
          } else {
            tma_load_2d(sbase + NN_SM_RAW + slot * NN_RAW_SLOT, &tmA1, k - P.K0, row, fb);
          }
          if (++slot == NN_RAW) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= NN_LOADER_WARP0 && warp < NN_EPI_WARP0) {
    // ------------------------------ A converters: raw fp32 chunk -> tf32 hi/lo UMMA core matrices ------------------------------
    // thread -> K quad (lt & 3) of rows (lt >> 2) and (lt >> 2) + 64 (raw chunk: [128 rows][16 floats], 64 bytes per row)
    const int lt = threadIdx.x - NN_LOADER_WARP0 * 32;
    const int quad = lt & 3, row0 = lt >> 2;
    const int total = my_tiles * nkc;
    uint32_t stage = 0, phase = 0, slot = 0, rphase = 0;
    uint32_t so[2], ro[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = row0 + 64 * h;
      so[h] = (uint32_t)quad * NN_A_LBO + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
      ro[h] = (uint32_t)row * 64u + (uint32_t)quad * 16u;
    }
    for (int i = 0; i < total; ++i) {
      mbar_wait(bar(NB_RAW_FULL + slot), rphase);
      const uint32_t rs = sbase + NN_SM_RAW + slot * NN_RAW_SLOT;
      const float4 v0 = ld_shared_v4(rs + ro[0]), v1 = ld_shared_v4(rs + ro[1]);
      mbar_wait(bar(NB_EMPTY + stage), phase ^ 1u);
      const uint32_t sa = sbase + stage * NN_STAGE;
      if (!(P.dbg & 2)) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        split_tf32(v0.x, h0, l0); split_tf32(v0.y, h1, l1); split_tf32(v0.z, h2, l2); split_tf32(v0.w, h3, l3);
        st_shared_v4(sa + so[0], h0, h1, h2, h3);
        if (!P.one_pass) st_shared_v4(sa + so[0] + NN_A_PART, l0, l1, l2, l3);
        split_tf32(v1.x, h0, l0); split_tf32(v1.y, h1, l1); split_tf32(v1.z, h2, l2); split_tf32(v1.w, h3, l3);
        st_shared_v4(sa + so[1], h0, h1, h2, h3);
        if (!P.one_pass) st_shared_v4(sa + so[1] + NN_A_PART, l0, l1, l2, l3);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(NB_A_FULL + stage)); mbar_arrive(bar(NB_RAW_EMPTY + slot)); }
      if (++stage == NN_STAGES) { stage = 0; phase ^= 1u; }
      if (++slot == NN_RAW) { slot = 0; rphase ^= 1u; }
    }
  } else if (warp >= NN_EPI_WARP0) {
    // ------------------------------ epilogue ------------------------------
    // TMEM -> registers (thread = row) -> per-warp 32x32 transposition tile in shared memory (16-byte XOR swizzle) -> each
    // global access of the warp then covers 4 rows x 128 contiguous bytes (C, mask, accumulate and per-ray operands alike)
    const int q = warp & 3;                       // TMEM lane quarter = warp_id % 4
    const int ew = warp - NN_EPI_WARP0;
    const int g = ew >> 2;                        // column half
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t tb = sbase + NN_SM_EPI + (uint32_t)ew * 4096u;
    const int ncol = P.N >> 1;
    const int nch = ncol >> 5;                    // 32-column chunks of this warp: 4 | 2 | 1
    const int rsub = lane >> 3, c4 = lane & 7;
    for (int t = 0; t < my_tiles; ++t) {
      const int buf = t & 1;
      const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
      const long long row_base = tile * 128 + q * 32;
      // relu' bit masks of this lane's 8 rows x this warp's 32-column chunks: fetched while the MMAs of the tile still run
      uint32_t mw[4][8];
      const bool have_bits = (MODE == 2 || MODE == 0) && P.e.act == 2 && P.e.mask_bits != nullptr;
      if (have_bits) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long grow = row_base + rsub + 4 * i;
            mw[c][i] = (c < nch && grow < P.M) ? __ldg(P.e.mask_bits + (size_t)grow * 8 + ((g * ncol + c * 32) >> 5)) : 0u;
          }
      }
      mbar_wait(bar(NB_ACC_FULL + buf), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      const uint32_t tcol = tlane + (uint32_t)buf * 256u + (uint32_t)(g * ncol);
      for (int c = 0; c < nch; ++c) {
        uint32_t ra[32];
        tmem_ld32(tcol + (uint32_t)c * 32u, ra);
        tmem_wait_ld();
        if (c + 1 == nch) {
          // every column of this warp has left TMEM: hand the accumulator buffer back before the global traffic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(NB_ACC_EMPTY + buf));
        }
        pin32(ra);
        const uint32_t wrow = tb + (uint32_t)lane * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) st_shared_v4(wrow + (uint32_t)((j ^ (lane & 7)) * 16), ra[4 * j], ra[4 * j + 1], ra[4 * j + 2], ra[4 * j + 3]);
        __syncwarp();
        const int col = g * ncol + c * 32 + c4 * 4;
        if (!(P.dbg & 8)) {
          // the chunk index selects the prefetched words at compile time (register array)
          if (c == 0) nn_epi_block<MODE>(P, tb, row_base, rsub, c4, col, mw[0], have_bits);
          else if (c == 1) nn_epi_block<MODE>(P, tb, row_base, rsub, c4, col, mw[1], have_bits);
          else if (c == 2) nn_epi_block<MODE>(P, tb, row_base, rsub, c4, col, mw[2], have_bits);
          else nn_epi_block<MODE>(P, tb, row_base, rsub, c4, col, mw[3], have_bits);
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ TN (weight gradients)
constexpr int TN_STAGES = 2;   // UMMA operand stages (hi/lo of both operands, 72 KB each)
constexpr int TN_RAW = 2;      // raw fp32 chunks in flight behind them (TMA bulk copies, 32 KB each)
// Operand rows are features, K = points.  8-row groups are SBO = 144 bytes apart (not 128): the loaders read float4 along the
// features (coalesced) and each thread then owns 4 core-matrix rows 64 bytes apart; the 16-byte pad per group makes those
// st.shared.v4 phases bank-conflict free.  K-adjacent core matrices: LBO = (features / 8) * 144.
constexpr uint32_t TN_SBO = 144;
constexpr uint32_t TN_PART = 4 * 32 * TN_SBO;       // 18432: up to 256 features x 16 points, one tf32 part
constexpr uint32_t TN_STAGE = 4 * TN_PART;          // A_hi | A_lo | B_hi | B_lo = 73728
constexpr uint32_t TN_SM_RAW = TN_STAGES * TN_STAGE;  // 147456: raw ring, per slot [A 16 KB | B 16 KB]
constexpr uint32_t TN_RAW_SLOT = 32768;
constexpr uint32_t TN_SM_BAR = TN_SM_RAW + TN_RAW * TN_RAW_SLOT;  // 212992
constexpr uint32_t TN_SM_TOTAL = TN_SM_BAR + 256;
constexpr int TN_THREADS = 576;                     // warp 1: MMA; warps 2..17: loaders (2..9 operand A, 10..17 operand B), then epilogue
constexpr int TB_FULL = 0, TB_EMPTY = 4, TB_ACC = 8, TB_TMEM_SLOT = 10, TB_RAW_FULL = 12, TB_RAW_EMPTY = 16;

struct TNParams {
  const float* A; int lda; int NA;
  const float* B; int ldb; int NB;
  float* Wg; int ldw; int col0; int valid;
  int P; int rows_per_cta;
  int one_pass;
  float* colsum;  // optional: colsum[f] += sum_p A[p][f] (bias gradient, fused into the operand-A converters)
  int dbg;  // bring-up: 1 = no global loads, 2 = no conversion/stores, 4 = no MMAs, 8 = no flush
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(TN_THREADS, 1) k_gemm_tc_tn(const TNParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bars = sbase + TN_SM_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + TN_SM_BAR + 8 * TB_TMEM_SLOT);
  const int NA = P.NA, NB = P.NB;
  const int p_begin = blockIdx.x * P.rows_per_cta;
  const int p_end = min(P.P, p_begin + P.rows_per_cta);
  const int nchunks = p_end > p_begin ? (p_end - p_begin + 15) >> 4 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TN_STAGES; ++i) { mbar_init(bar(TB_FULL + i), 16); mbar_init(bar(TB_EMPTY + i), 1); }
    for (int i = 0; i < TN_RAW; ++i) { mbar_init(bar(TB_RAW_FULL + i), 1); mbar_init(bar(TB_RAW_EMPTY + i), 16); }
    mbar_init(bar(TB_ACC), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // raw producer: the 16 points of a chunk are contiguous in both operands (lda == NA, ldb == NB): two bulk copies per chunk
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bar(TB_RAW_EMPTY + slot), phase ^ 1u);
        const int p0 = p_begin + c * 16;
        const uint32_t rows = (uint32_t)min(16, p_end - p0);
        const uint32_t ba = rows * (uint32_t)NA * 4u, bb = rows * (uint32_t)NB * 4u;
        const uint32_t fb = bar(TB_RAW_FULL + slot);
        const uint32_t dst = sbase + TN_SM_RAW + slot * TN_RAW_SLOT;
        mbar_expect_tx(fb, ba + bb);
        if (!(P.dbg & 1)) {
          const uint8_t* ga = reinterpret_cast<const uint8_t*>(P.A + (size_t)p0 * NA);
          const uint8_t* gb = reinterpret_cast<const uint8_t*>(P.B + (size_t)p0 * NB);
          for (uint32_t o = 0; o < ba; o += 8192u) bulk_g2s(dst + o, ga + o, min(8192u, ba - o), fb);
          for (uint32_t o = 0; o < bb; o += 8192u) bulk_g2s(dst + 16384u + o, gb + o, min(8192u, bb - o), fb);
        } else {
          asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(fb), "r"(ba + bb) : "memory");
        }
        if (++slot == TN_RAW) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one() && nchunks > 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t idesc = idesc_tf32((uint32_t)NB);
      const uint32_t lbo_a = (uint32_t)(NA >> 3) * TN_SBO, lbo_b = (uint32_t)(NB >> 3) * TN_SBO;
      const uint32_t astep = 2u * lbo_a / 16u, bstep = 2u * lbo_b / 16u;  // one K8 step = 2 core columns (16-byte units)
      const uint32_t dhi = desc_hi(TN_SBO);
      const int MT = NA >> 7;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bar(TB_FULL + stage), phase);
        tc_fence_after();
        const uint32_t sa = sbase + stage * TN_STAGE;
        const uint32_t a_hi = desc_lo(sa, lbo_a), a_lo = desc_lo(sa + TN_PART, lbo_a);
        const uint32_t b_hi = desc_lo(sa + 2 * TN_PART, lbo_b), b_lo = desc_lo(sa + 3 * TN_PART, lbo_b);
        for (int mt = 0; mt < MT && !(P.dbg & 4); ++mt) {
          const uint32_t d_tmem = tmem + (uint32_t)(mt * NB);
          const uint32_t mo = (uint32_t)mt * (16u * TN_SBO / 16u);  // 128 rows = 16 row groups
#pragma unroll
          for (uint32_t j = 0; j < 2; ++j) {
            mma_tf32(d_tmem, a_hi + mo + j * astep, b_hi + j * bstep, idesc, (c > 0 || j > 0) ? 1u : 0u, dhi, dhi);
            if (!P.one_pass) {
              mma_tf32(d_tmem, a_lo + mo + j * astep, b_hi + j * bstep, idesc, 1u, dhi, dhi);
              mma_tf32(d_tmem, a_hi + mo + j * astep, b_lo + j * bstep, idesc, 1u, dhi, dhi);
            }
          }
        }
        tc_commit(bar(TB_EMPTY + stage));
        if (++stage == TN_STAGES) { stage = 0; phase ^= 1u; }
      }
      tc_commit(bar(TB_ACC));
    }
  } else if (warp >= 2) {
    // loaders: warps 2..9 stage operand A, warps 10..17 operand B.  thread -> (feature quad fq, point group j): 4 float4 loads along
    // the features (coalesced), a 4x4 register transpose, 4 core-matrix rows (one feature x 4 points each).  Four chunks in flight.
    const int lt = (threadIdx.x - 64) & 255;
    const bool is_b = warp >= 10;
    const int NX = is_b ? NB : NA;
    const float* X = is_b ? P.B : P.A;
    const size_t ldx = is_b ? (size_t)P.ldb : (size_t)P.lda;
    const int sh = 31 - __clz(NX >> 2);  // log2(feature quads)
    const bool on = lt < NX;
    const int fq = lt & ((NX >> 2) - 1), j = lt >> sh;
    const uint32_t lbo = (uint32_t)(NX >> 3) * TN_SBO;
    const uint32_t part0 = is_b ? 2 * TN_PART : 0u;
    const uint32_t raw_off = (is_b ? 16384u : 0u) + (uint32_t)(4 * j * NX + 4 * fq) * 4u;  // (point 4j, feature 4fq) of the raw slot
    auto put = [&](uint32_t base, const float4 (&x)[4]) {
      const float r[4][4] = {{x[0].x, x[1].x, x[2].x, x[3].x}, {x[0].y, x[1].y, x[2].y, x[3].y},
                             {x[0].z, x[1].z, x[2].z, x[3].z}, {x[0].w, x[1].w, x[2].w, x[3].w}};
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int f = 4 * fq + cc;
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        split_tf32(r[cc][0], h0, l0); split_tf32(r[cc][1], h1, l1); split_tf32(r[cc][2], h2, l2); split_tf32(r[cc][3], h3, l3);
        const uint32_t a = base + (uint32_t)j * lbo + (uint32_t)(f >> 3) * TN_SBO + (uint32_t)(f & 7) * 16u;
        st_shared_v4(a, h0, h1, h2, h3);
        if (!P.one_pass) st_shared_v4(a + TN_PART, l0, l1, l2, l3);
      }
    };
    uint32_t stage = 0, phase = 0, slot = 0, rphase = 0;
    (void)X; (void)ldx;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};  // column sums of operand A over this thread's points
    for (int c = 0; c < nchunks; ++c) {
      const int p0 = p_begin + c * 16 + 4 * j;
      mbar_wait(bar(TB_RAW_FULL + slot), rphase);
      float4 x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        x[i] = (on && p0 + i < p_end && !(P.dbg & 1)) ? ld_shared_v4(sbase + TN_SM_RAW + slot * TN_RAW_SLOT + raw_off + (uint32_t)(i * NX) * 4u)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      if (!is_b && P.colsum != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { cs[0] += x[i].x; cs[1] += x[i].y; cs[2] += x[i].z; cs[3] += x[i].w; }
      }
      mbar_wait(bar(TB_EMPTY + stage), phase ^ 1u);
      if (on && !(P.dbg & 2)) put(sbase + stage * TN_STAGE + part0, x);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(TB_FULL + stage)); mbar_arrive(bar(TB_RAW_EMPTY + slot)); }
      if (++stage == TN_STAGES) { stage = 0; phase ^= 1u; }
      if (++slot == TN_RAW) { slot = 0; rphase ^= 1u; }
    }
    if (!is_b && on && P.colsum != nullptr) {
#pragma unroll
      for (int c = 0; c < 4; ++c) atomicAdd(P.colsum + 4 * fq + c, cs[c]);
    }
    // epilogue (same 16 warps): flush the dW tile with reductions into global memory.  The MT x NB accumulator columns are split
    // into four column groups (one per set of 4 warps with distinct TMEM lane quarters).
    if (nchunks > 0 && !(P.dbg & 8)) {
      const int q = warp & 3, g4 = (warp - 2) >> 2;
      const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
      mbar_wait(bar(TB_ACC), 0u);
      tc_fence_after();
      const int tot = (NA >> 7) * NB;                  // TMEM columns in use (64 .. 512)
      const int ng = tot >= 128 ? 4 : tot >> 5;        // column groups that get work
      const int per = tot / ng;                        // columns per group (multiple of 32)
      const bool vec = ((P.ldw & 3) == 0) && ((P.col0 & 3) == 0) && P.valid == NB;
      for (int tc0 = g4 * per; g4 < ng && tc0 < (g4 + 1) * per; tc0 += 32) {
        const int mt = tc0 / NB, cb = tc0 - mt * NB;
        const int r = mt * 128 + q * 32 + lane;
        float* wrow = P.Wg + (size_t)r * P.ldw + P.col0;
        uint32_t v[32];
        tmem_ld32(tlane + (uint32_t)tc0, v);
        tmem_wait_ld();
        pin32(v);
        if (vec) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            red_add_v4(wrow + cb + 4 * jj, __uint_as_float(v[4 * jj]), __uint_as_float(v[4 * jj + 1]), __uint_as_float(v[4 * jj + 2]),
                       __uint_as_float(v[4 * jj + 3]));
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)
            if (cb + jj < P.valid) atomicAdd(wrow + cb + jj, __uint_as_float(v[jj]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int g_dbg = 0;
int g_one_pass = 0;
// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// fp32 matrix [rows, cols] with row stride ld (floats) -> boxes of 128 rows x 16 columns, no swizzle, zero fill outside
int make_a_map(CUtensorMap* m, const float* A, int rows, int cols, int ld) {
  EncodeTiledFn enc = encode_tiled();
  MNRF_REQUIRE(enc != nullptr, "gemm_nn_tc: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {16, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MNRF_REQUIRE(r == CUDA_SUCCESS, "gemm_nn_tc: cuTensorMapEncodeTiled failed (%d) for a %d x %d matrix, ld %d", (int)r, rows, cols, ld);
  return 0;
}



}  // namespace

int gemm_nn_tc(const mnrf_field* f, int step, const float* A0, int lda0, int K0, const float* A1, int lda1, float* C,
               int ldc, int M, const GemmEpi& e, cudaStream_t st) {
  if (M <= 0) return 0;
  MNRF_REQUIRE(step >= 0 && step < T32_NUM_STEPS, "gemm_nn_tc: bad step %d", step);
  const int N = t32_step_n(step), K = t32_step_k(step);
  MNRF_REQUIRE(K0 % 16 == 0 && K0 <= K && (K0 == K || A1 != nullptr), "gemm_nn_tc: bad K split %d of %d", K0, K);
  MNRF_REQUIRE(lda0 % 4 == 0 && (A1 == nullptr || lda1 % 4 == 0) && ldc % 4 == 0, "gemm_nn_tc: leading dimensions must be multiples of 4");
  int sms = 0;
  if (first_use_on_device(TAG_TRAIN_TC_NN, &sms)) {
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_nn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NN_SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_nn<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NN_SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_nn<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NN_SM_TOTAL));
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_nn<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NN_SM_TOTAL));
  }
  MNRF_REQUIRE(sms > 0, "gemm_nn_tc: no CUDA device");
  NNParams P;
  P.A0 = A0; P.lda0 = lda0; P.K0 = K0; P.A1 = A1; P.lda1 = lda1;
  P.blob = f->t32 + t32_step_offset(step);
  P.C = C; P.ldc = ldc; P.M = M; P.N = N; P.K = K;
  P.n_tiles = (M + 127) / 128;
  P.e = e;
  if (getenv("MNRF_NO_BITS") != nullptr) P.e.mask_bits = nullptr;  // A/B switch: read the fp32 activation as the mask
  P.one_pass = g_one_pass;
  P.dbg = g_dbg;
  const int grid = P.n_tiles < sms ? P.n_tiles : sms;
  CUtensorMap m0, m1;
  MNRF_REQUIRE(((uintptr_t)A0 & 15) == 0 && (A1 == nullptr || ((uintptr_t)A1 & 15) == 0), "gemm_nn_tc: A must be 16-byte aligned");
  if (make_a_map(&m0, A0, M, K0, lda0)) return 2;
  if (K0 < K) { if (make_a_map(&m1, A1, M, K - K0, lda1)) return 2; } else m1 = m0;
  const bool plain = !e.accumulate && e.rowbias == nullptr && e.rvec == nullptr;
  if (plain && e.act == 1 && e.bias != nullptr) k_gemm_tc_nn<1><<<grid, NN_THREADS, NN_SM_TOTAL, st>>>(P, m0, m1);
  else if (plain && e.act == 2 && e.bias == nullptr) k_gemm_tc_nn<2><<<grid, NN_THREADS, NN_SM_TOTAL, st>>>(P, m0, m1);
  else if (plain && e.act == 0) k_gemm_tc_nn<3><<<grid, NN_THREADS, NN_SM_TOTAL, st>>>(P, m0, m1);
  else k_gemm_tc_nn<0><<<grid, NN_THREADS, NN_SM_TOTAL, st>>>(P, m0, m1);
  MNRF_LAUNCH_OK();
  return 0;
}

int gemm_tn_tc(const float* A, int lda, int NA, const float* B, int ldb, int NB, float* Wg, int ldw, int col0, int valid,
               int Pn, float* colsum_out, cudaStream_t st) {
  if (Wg == nullptr || Pn <= 0) return 0;
  MNRF_REQUIRE((NA == 128 || NA == 256) && (NB == 64 || NB == 128 || NB == 256), "gemm_tn_tc: bad shape %d x %d", NA, NB);
  MNRF_REQUIRE(lda == NA && ldb == NB, "gemm_tn_tc: operands must be contiguous (lda == NA, ldb == NB): their chunks are bulk-copied");
  int sms = 0;
  if (first_use_on_device(TAG_TRAIN_TC_TN, &sms))
    MNRF_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_tn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TN_SM_TOTAL));
  MNRF_REQUIRE(sms > 0, "gemm_tn_tc: no CUDA device");
  TNParams P;
  P.A = A; P.lda = lda; P.NA = NA; P.B = B; P.ldb = ldb; P.NB = NB;
  P.Wg = Wg; P.ldw = ldw; P.col0 = col0; P.valid = valid; P.P = Pn;
  int rows = (Pn + sms - 1) / sms;
  rows = ((rows + 15) / 16) * 16;
  if (rows < 256) rows = 256;  // small problems: fewer CTAs, fewer atomics
  P.rows_per_cta = rows;
  P.colsum = colsum_out;
  P.one_pass = g_one_pass;
  P.dbg = g_dbg;
  const int grid = (Pn + rows - 1) / rows;
  k_gemm_tc_tn<<<grid, TN_THREADS, TN_SM_TOTAL, st>>>(P);
  MNRF_LAUNCH_OK();
  return 0;
}

void set_train_tc_debug(int flags) { g_dbg = flags; }
void set_train_tc_one_pass(int on) { g_one_pass = on ? 1 : 0; }

}  // namespace mnrf
