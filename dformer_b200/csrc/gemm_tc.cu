// tcgen05 / TMEM / TMA GEMM for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
//   C[M,N] = epi( opA(A) * opB(B) + bias ),   same operand conventions as gemm_simt.cu.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  : cp.async.bulk.tensor tiles (128B swizzle) into a 4-stage smem ring
//   warp 1      MMA issuer    : one thread issues tcgen05.mma (UMMA 128 x BN x 16), accumulators in TMEM,
//                               double-buffered (2 x 256 columns) so the epilogue of tile i overlaps tile i+1
//   warps 2..9  epilogue      : tcgen05.ld (32 lanes x 32 columns per warp) -> bias / activation -> global
//                               (two warps per TMEM lane quadrant, each owning half of the tile's columns)
//
// Both operands may be K-major (reduction dim contiguous) or MN-major (reduction dim strided), which is
// what lets one kernel serve forward (A K-major, W K-major), dgrad (dY K-major, W MN-major) and wgrad
// (dY MN-major, X MN-major) without any transposed copies in HBM.  The wgrad case reduces over the pixel
// dimension, so the reduction can be split across CTAs (split-K) with fp32 red.global.add epilogues.
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"
#include "dfb200_internal.h"

namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int MAX_STAGES = 8;                     // ring depth is chosen per launch: as deep as shared memory allows for this BN
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KB
constexpr int MAX_SMEM = 227 * 1024;
constexpr int SMALL_SMEM = 112 * 1024;            // EPI_WARPS = 4 variant: two CTAs per SM
constexpr int EPI_WARP_BYTES = 4096;              // per-epilogue-warp [32 rows x 128 B] staging tile of the TMA store
constexpr int EPI_BIAS_FLOATS = 256;              // per-epilogue-warp bias slice (<= 4 boxes of 64 columns per warp and tile)

struct TcParams {
  int M, N, K;               // problem (K = reduction length)
  int BN;                    // tile N (multiple of 16, <= 256)
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  int a_mn_major, b_mn_major;
  int batch; long strideC;
  int stages, stage_bytes;
  void* C; long ldc;
  const float* bias;
  int out_bf16, act, act_col_start, accumulate;
  int acc_stages, acc_stride, tmem_cols;     // accumulator ring in tensor memory: 1 or 2 stages, `acc_stride` columns apart
  int tma_store;                             // bf16 C written by cp.async.bulk.tensor (tmC valid)
  // fused epilogue (TMA-store path only), see dfb200.h:
  //   1 "gate": C = (acc + bias) * aux (bf16 [M, N], leading dimension ld_aux); second output (optional) = acc + bias
  int epi_mode, has_out2;
  int bias_smem;                             // the epilogue warps stage their bias slices in shared memory (when that costs no pipeline stage)
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("dfb200 gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), 128-byte swizzle:
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell)   [61,64) layout type = 2 (SWIZZLE_128B)
// K-major : rows of 128 B, 8-row groups 1024 B apart           -> LBO (unused) = 1, SBO = 1024
// MN-major: 64-element MN chunks are [64 k-rows x 128 B] boxes -> LBO = 8192 (next MN chunk), SBO = 1024 (next 8 k-rows)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, bool mn_major) {
  const uint64_t lbo = mn_major ? (8192u >> 4) : 1u;
  const uint64_t sbo = 1024u >> 4;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint32_t* w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
// Generic (rarely taken) epilogue: ragged chunks / row tails, fp32 outputs, accumulate, scalar split-K.  Bias and activation
// have already been applied.  Kept out of line so the hot loop stays small in the instruction cache.
__device__ __noinline__ void epilogue_generic(const TcParams& p, float* v, int row, int col0, int ncols, long c_off) {
  const bool full = ncols == 32;
  if (p.splits > 1) {                       // split-K partial sums: fp32 reductions into a zeroed / accumulating C
    float* dst = reinterpret_cast<float*>(p.C) + c_off + (long)row * p.ldc + col0;
    if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) atomicAdd(dst + j, v[j]);
    }
    return;
  }
  if (p.out_bf16) {
    bf16* dst = reinterpret_cast<bf16*>(p.C) + c_off + (long)row * p.ldc + col0;
    if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) Vec8<bf16>::store(dst + j, v + j);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) dst[j] = __float2bfloat16_rn(v[j]);
    }
  } else {
    float* dst = reinterpret_cast<float*>(p.C) + c_off + (long)row * p.ldc + col0;
    const bool vec = full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    if (p.accumulate) {
      if (vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = *reinterpret_cast<float4*>(dst + j);
          o.x += v[j]; o.y += v[j + 1]; o.z += v[j + 2]; o.w += v[j + 3];
          *reinterpret_cast<float4*>(dst + j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) dst[j] += v[j];
      }
    } else if (vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) dst[j] = v[j];
    }
  }
}

// bias (first split only) on one 32-column accumulator chunk held in registers
__device__ __forceinline__ void add_bias_chunk(const TcParams& p, float* v, int col0, int ncols) {
  const float* bp = p.bias + col0;
  if (ncols == 32 && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + j));
      v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) v[j] += __ldg(bp + j);
  }
}
// activation on the columns >= act_col_start
__device__ __forceinline__ void act_chunk(const TcParams& p, float* v, int col0) {
  if (p.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (col0 + j >= p.act_col_start) ? gelu_f(v[j]) : v[j];
  } else if (p.act == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (col0 + j >= p.act_col_start) ? fmaxf(v[j], 0.f) : v[j];
  }
}
__device__ __forceinline__ void bias_act(const TcParams& p, float* v, int col0, int ncols, bool add_bias) {
  if (add_bias) add_bias_chunk(p, v, col0, ncols);
  act_chunk(p, v, col0);
}
// 32 fp32 values of row `lane` -> bf16 -> half `h` of the warp's [32 rows x 128 B] staging tile in the 128-byte swizzle pattern
// (the row's eight 16-byte chunks land at chunk ^ (row & 7): conflict-free, and what the TMA store expects)
__device__ __forceinline__ void stage_half(uint32_t sbase, int lane, int h, const float* v) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[u * 8 + 2 * q], v[u * 8 + 2 * q + 1]);
      w[q] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    sts128(sbase + (uint32_t)(lane * 8 + ((h * 4 + u) ^ (lane & 7))) * 16u, w);
  }
}

// ------------------------------------------------------------------ kernel
// EPI_WARPS = 8: one CTA per SM with the whole shared memory (deep ring, 2 x 256 TMEM columns), two epilogue warps per TMEM
//                lane quadrant (column halves).
// EPI_WARPS = 4: 192 threads, <= 112 KB shared memory and <= 256 TMEM columns per CTA so that TWO CTAs are resident per SM:
//                150 tiles of an M = 9600 layer fit in one wave of 296 slots, and GEMMs of different streams (RGB / depth /
//                weight-gradient) overlap one CTA's load phase with the other's epilogue.
// GATE / BIAS_SMEM are compile-time: carried as run-time branches inside the epilogue loop they cost the plain store-bound GEMMs
// 16 - 20 % (fc1-type (153600, 768, 96): 61.6 -> 74.7 us) -- the hot loop of the common case must contain nothing but its own work.
template <int EPI_WARPS, bool GATE, bool BIAS_SMEM>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, EPI_WARPS == 4 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmE, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + p.stages * p.stage_bytes;                 // EPI_WARPS x [32 rows x 128 B], 1024-byte aligned (128-byte swizzle atom)
  uint8_t* epi_stage2 = epi_stage + EPI_WARPS * EPI_WARP_BYTES;         // second output's staging tiles (present when has_out2)
  const bool has_out2 = GATE && p.has_out2 != 0;
  uint8_t* epi_aux = epi_stage2 + (has_out2 ? EPI_WARPS * EPI_WARP_BYTES : 0);       // gate tiles (present in gate mode), TMA-loaded per box
  float* epi_bias = reinterpret_cast<float*>(epi_aux + (GATE ? EPI_WARPS * EPI_WARP_BYTES : 0));   // per warp: the bias of its boxes
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_bias + (BIAS_SMEM ? EPI_WARPS * EPI_BIAS_FLOATS : 0));
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint64_t* aux_bar = tmem_empty + 2;           // [EPI_WARPS]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(aux_bar + EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_mn * p.splits * p.batch;     // tile = ((batch * splits + split) * m_tiles + m) * n_tiles + n

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    if (has_out2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
    if (GATE) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmE) : "memory");
    for (int s = 0; s < EPI_WARPS; ++s) mbar_init(&aux_bar[s], 1);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // barrier init / tensor-memory allocation above touch no global memory: under programmatic dependent launch they overlap the
  // tail of the previous kernel of the stream; everything below (TMA loads, bias reads, stores) waits for it
  pdl_sync();

  const int a_boxes = p.a_mn_major ? 2 : 1;                       // MN-major: one [64 k x 64 mn] box per 64 MN elements
  const int b_boxes = p.b_mn_major ? (p.BN + 63) / 64 : 1;
  const uint32_t stage_tx = (uint32_t)A_STAGE_BYTES + (uint32_t)(p.b_mn_major ? b_boxes * 8192 : p.BN * 128);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles, m_blk = (tile / p.n_tiles) % p.m_tiles;
        const int sp = (tile / tiles_mn) % p.splits, bz = tile / (tiles_mn * p.splits);
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int ncol0 = n_blk * p.BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * p.stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], stage_tx);
          if (p.a_mn_major) {
            for (int i = 0; i < a_boxes; ++i) tma_load_3d(&tmA, &full_bar[stage], sa + i * 8192, m_blk * BM + i * 64, kb * BK, bz);
          } else {
            tma_load_3d(&tmA, &full_bar[stage], sa, kb * BK, m_blk * BM, bz);
          }
          if (p.b_mn_major) {
            for (int i = 0; i < b_boxes; ++i) tma_load_3d(&tmB, &full_bar[stage], sb + i * 8192, ncol0 + i * 64, kb * BK, bz);
          } else {
            tma_load_3d(&tmB, &full_bar[stage], sb, kb * BK, ncol0, bz);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      // Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
      // a_major bit 15, b_major bit 16, N>>3 at [17,23), M>>4 at [24,29).
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn_major << 15) |
                             ((uint32_t)p.b_mn_major << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_kstep = p.a_mn_major ? 2048u : 32u;   // bytes per UMMA_K = 16 reduction elements
      const uint32_t b_kstep = p.b_mn_major ? 2048u : 32u;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int sp = (tile / tiles_mn) % p.splits;
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = make_smem_desc(sa + k * a_kstep, p.a_mn_major);
            const uint64_t bd = make_smem_desc(sb + k * b_kstep, p.b_mn_major);
            umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);          // frees the smem slot once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);              // accumulator complete -> epilogue
        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue (warps 2 .. 2 + EPI_WARPS) ===============================
    // warp w reads TMEM lane quadrant (w & 3) (hardware rule: a warp may only touch lanes 32*(w%4)..+31); with 8 warps the two
    // warps of a quadrant split the tile's columns.
    constexpr int PARTS = EPI_WARPS / 4;
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;
    const uint32_t sbase = smem_u32(epi_stage + (warp - 2) * EPI_WARP_BYTES);
    const uint32_t sbase2 = smem_u32(epi_stage2 + (warp - 2) * EPI_WARP_BYTES);
    const uint32_t sbase_aux = smem_u32(epi_aux + (warp - 2) * EPI_WARP_BYTES);
    uint64_t* my_aux_bar = &aux_bar[warp - 2];
    float* bias_w = epi_bias + (warp - 2) * EPI_BIAS_FLOATS;
    uint32_t aux_phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    bool store_pending = false;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.n_tiles, m_blk = (tile / p.n_tiles) % p.m_tiles;
      const int bz = tile / (tiles_mn * p.splits);
      const int row = m_blk * BM + quad * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.acc_stride);
      const bool row_ok = row < p.M;
      const bool add_bias = p.bias != nullptr && (p.splits == 1 || ((tile / tiles_mn) % p.splits) == 0);
      const int ncol0 = n_blk * p.BN;
      const int tn = min(p.BN, p.N - ncol0);                        // valid columns of this tile
      if (p.tma_store) {
        // ---- bf16 output through TMA: 64-column boxes, [32 rows x 128 B] staged in the 128-byte swizzle pattern (lane = row writes
        // its eight 16-byte chunks at chunk ^ (row & 7): conflict-free), one cp.async.bulk.tensor store per box; rows >= M and
        // columns >= N are clipped by the hardware
        const int nboxes = (tn + 63) >> 6;
        const int b_begin = (PARTS == 1 || part == 0) ? 0 : (nboxes + 1) >> 1;
        const int b_end = PARTS == 1 ? nboxes : (part == 0 ? (nboxes + 1) >> 1 : nboxes);
        // gate: the [32 rows x 64 columns] gate tile of a box is TMA-loaded into this warp's aux buffer (same swizzled layout as the
        // staging tile).  It does not depend on the accumulator: the first box's load is issued before waiting for the MMAs of the
        // tile, the next box's as soon as the current one has been consumed.
        auto load_gate = [&](int box) {
          if (GATE && box < b_end && lane == 0) {
            mbar_expect_tx(my_aux_bar, EPI_WARP_BYTES);
            tma_load_3d(&tmE, my_aux_bar, epi_aux + (warp - 2) * EPI_WARP_BYTES, ncol0 + box * 64, m_blk * BM + quad * 32, bz);
          }
        };
        load_gate(b_begin);
        // bias of this warp's boxes -> shared memory while the MMAs of the tile are still running: an L2 round trip between
        // "accumulator ready" and the first store is the critical path of every one-tile-per-CTA launch (the small-M layers)
        if (BIAS_SMEM && add_bias) {
          const int c_lo = ncol0 + b_begin * 64;
          for (int i = lane; i < (b_end - b_begin) * 64; i += 32) bias_w[i] = (c_lo + i < p.N) ? __ldg(p.bias + c_lo + i) : 0.f;
          __syncwarp();
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        for (int box = b_begin; box < b_end; ++box) {
          const int cb = ncol0 + box * 64;                            // first global column of this box
          const bool sec = has_out2;                                  // gate mode keeps acc + bias as a second output
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c0 = box * 64 + h * 32;
            uint32_t r[32];
            __syncwarp();
            tmem_ld32(t_row + (uint32_t)c0, r);
            tmem_ld_wait();
            if (h == 0 && store_pending) {          // the previous box of this warp must have left the staging tile(s)
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
            const int col0 = ncol0 + c0;
            const int ncols = min(32, p.N - col0);
            if (ncols <= 0) continue;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (BIAS_SMEM && add_bias) {
              const float4* bq = reinterpret_cast<const float4*>(bias_w + (box - b_begin) * 64 + h * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = bq[j];
                v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
              }
            } else if (add_bias) {
              add_bias_chunk(p, v, col0, ncols);
            }
            act_chunk(p, v, col0);
            if (GATE) {
              // ---- gate: second output keeps acc + bias (needed by the backward pass), C = (acc + bias) * aux
              if (sec) stage_half(sbase2, lane, h, v);
              if (h == 0) { mbar_wait(my_aux_bar, aux_phase); aux_phase ^= 1; }
#pragma unroll
              for (int u = 0; u < 4; ++u) {          // out-of-range rows / columns were zero-filled by the load and are clipped by the store
                float g8[8];
                Vec8<bf16>::unpack(lds128(sbase_aux + (uint32_t)(lane * 8 + ((h * 4 + u) ^ (lane & 7))) * 16u), g8);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[u * 8 + j] *= g8[j];
              }
            }
            stage_half(sbase, lane, h, v);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          load_gate(box + 1);
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(&tmC), "r"(sbase), "r"(cb), "r"(m_blk * BM + quad * 32), "r"(bz) : "memory");
            if (sec)
              asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                           ::"l"(&tmD), "r"(sbase2), "r"(cb), "r"(m_blk * BM + quad * 32), "r"(bz) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          store_pending = true;
        }
      } else {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const long c_off = (long)bz * p.strideC;
        const int chunks = (tn + 31) >> 5;
        const int c_begin = (PARTS == 1 || part == 0) ? 0 : (chunks + 1) >> 1;
        const int c_end = PARTS == 1 ? chunks : (part == 0 ? (chunks + 1) >> 1 : chunks);
        for (int ch = c_begin; ch < c_end; ++ch) {
          const int c0 = ch << 5;
          uint32_t r[32];
          __syncwarp();                             // tcgen05.ld is warp-collective: reconverge after the per-row predicated stores
          tmem_ld32(t_row + (uint32_t)c0, r);
          tmem_ld_wait();
          const int col0 = ncol0 + c0;
          const int ncols = min(32, tn - c0);
          if (!row_ok) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          bias_act(p, v, col0, ncols, add_bias);
          // ---- split-K partial sums -> vector reductions into the fp32 gradient
          if (p.splits > 1 && (ncols & 3) == 0) {
            float* dst = reinterpret_cast<float*>(p.C) + c_off + (long)row * p.ldc + col0;
            if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (j < ncols)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
              continue;
            }
          }
          // ---- fp32 output without split-K (small wgrads accumulate into the arena; downsample convs write fp32)
          if (p.splits == 1 && !p.out_bf16 && (ncols & 3) == 0) {
            float* dst = reinterpret_cast<float*>(p.C) + c_off + (long)row * p.ldc + col0;
            if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (j < ncols) {
                  float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                  if (p.accumulate) {
                    const float4 old = *reinterpret_cast<const float4*>(dst + j);
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                  }
                  *reinterpret_cast<float4*>(dst + j) = o;
                }
              }
              continue;
            }
          }
          // ---- bf16 output that cannot go through TMA (unaligned leading dimension / base): whole 16-byte vectors per lane
          if (p.splits == 1 && p.out_bf16 && (ncols & 7) == 0) {
            bf16* dst = reinterpret_cast<bf16*>(p.C) + c_off + (long)row * p.ldc + col0;
            if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                if (j < ncols) Vec8<bf16>::store(dst + j, v + j);
              continue;
            }
          }
          {
            float vv[32];                            // address-taken copy: keeps v[] itself in registers for the hot paths
#pragma unroll
            for (int j = 0; j < 32; ++j) vv[j] = v[j];
            epilogue_generic(p, vv, row, col0, ncols, c_off);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
    // the staging tiles must outlive the bulk stores' reads of them, nothing more: completion of the grid (what the next kernel's
    // griddepcontrol.wait / stream order waits for) covers the writes themselves
    if (store_pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; long inner, outer, ld, batch, bstride; int box_inner, box_outer;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = 1469598103934665603ull;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
    for (size_t i = 0; i < sizeof(MapKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
  }
};

// 3-D bf16 tensor map over `batch` row-major [outer, inner] matrices (leading dimension ld, batch stride bstride, elements).
int make_map(CUtensorMap* out, const void* ptr, long inner, long outer, long ld, int box_inner, int box_outer, long batch, long bstride) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.inner = inner; key.outer = outer; key.ld = ld; key.box_inner = box_inner; key.box_outer = box_outer;
  key.batch = batch; key.bstride = bstride;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return DFB_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { dfb_set_error("cuTensorMapEncodeTiled entry point not available"); return DFB_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bstride * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    dfb_set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%ld outer=%ld ld=%ld box=%dx%d", (int)r, ptr, inner, outer, ld, box_inner, box_outer);
    return DFB_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, *out);
  }
  return DFB_OK;
}

// Tile width: minimise the estimated makespan of the static round-robin tile schedule,
//   waves(BN) * bytes-moved-per-tile(BN)   with  waves = ceil(m_tiles * n_tiles * batch / slots),  slots = #SM * CTAs per SM,
// instead of only minimising padding.  `bn_cap` is the widest tile the variant's shared-memory budget leaves a pipeline for;
// `step64`: the TMA-store epilogue works on 64-column boxes, so interior tile boundaries must be multiples of 64 (a single
// tile may have any width that is a multiple of 16: its ragged last box is clipped by the tensor map).
int pick_bn_padding(int N, int bn_cap) {      // widest tile with the least padding (split-K problems: the reduction split fills the machine)
  if (N <= bn_cap) return ((N + 15) / 16) * 16;
  int best = bn_cap; long best_cost = (long)dfb_cdiv(N, bn_cap) * bn_cap;
  for (int bn = bn_cap - 16; bn >= 64; bn -= 16) {
    const long cost = (long)dfb_cdiv(N, bn) * bn;
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

int pick_bn(int N, int m_tiles, int kb, int batch, int slots, bool b_mn_major, int bn_cap, bool step64) {
  const int n16 = ((N + 15) / 16) * 16;
  int best = 0;
  double best_cost = 1e300;
  auto consider = [&](int bn) {
    if (bn > bn_cap || bn < 16) return;
    if (b_mn_major && bn < 64 && N >= 64) return;
    const long n_tiles = (N + bn - 1) / bn;
    const long tiles = (long)m_tiles * n_tiles * batch;
    const long waves = (tiles + slots - 1) / slots;
    const double per_tile = (double)kb * (16384.0 + 128.0 * bn) + 256.0 * bn + 24000.0;
    const double cost = (double)waves * per_tile;
    if (cost < best_cost * 0.97) { best_cost = cost; best = bn; }     // candidates come widest first: prefer wider tiles unless clearly (3 %) worse
  };
  if (n16 <= bn_cap) consider(n16);                                    // one tile spanning N
  for (int bn = bn_cap / (step64 ? 64 : 16) * (step64 ? 64 : 16); bn >= (step64 ? 64 : 32); bn -= (step64 ? 64 : 16)) consider(bn);
  if (best == 0) best = n16 <= bn_cap ? n16 : (step64 ? 64 : 32);
  return best;
}

int next_pow2_cols(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

}  // namespace

// SM budgets of the persistent grids: [0] forward / dgrad, [1] split-K (weight gradients).  g_sms_budget = 0 -> the measured defaults.
static int g_sms_default[2] = {0, 0};
static int g_sms_budget[2] = {0, 0};
extern "C" int dfb200_gemm_sm_budget(int main_sms, int split_sms) {
  g_sms_budget[0] = main_sms > 0 ? main_sms : 0;
  g_sms_budget[1] = split_sms > 0 ? split_sms : 0;
  return DFB_OK;
}

bool dfb_gemm_tc_supported(const dfb200_gemm_args& g) {
  if (g.a_dtype != 1 || g.b_dtype != 1 || g.batch < 1 || g.batch_inner > 1) return false;
  if (g.batch > 1 && ((g.strideA % 8) || (g.strideB % 8) || g.bias != nullptr)) return false;
  if (g.alpha != 0.f && g.alpha != 1.f) return false;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
  if ((g.lda % 8) || (g.ldb % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) return false;
  if (g.accumulate && g.out_dtype != 0) return false;
  if (g.epi_mode != 0) {                       // fused "gate" epilogue: bf16 output through the TMA-store path only
    if (g.epi_mode != 1) return false;
    if (g.batch != 1 || g.out_dtype != 1 || g.accumulate || g.splitk > 1) return false;
    if ((g.ldc % 8) || (reinterpret_cast<uintptr_t>(g.C) & 15)) return false;
    if (g.out2 && ((g.ld_out2 % 8) || (reinterpret_cast<uintptr_t>(g.out2) & 15))) return false;
    if (!g.aux || (g.ld_aux % 8) || (reinterpret_cast<uintptr_t>(g.aux) & 15) || (g.N % 8)) return false;
  }
  return true;
}

int dfb_gemm_tc(const dfb200_gemm_args& g, cudaStream_t st) {
  DFB_REQUIRE(dfb_gemm_tc_supported(g), "gemm_tc: unsupported arguments (dtype=%d/%d batch=%d lda=%ld ldb=%ld epi_mode=%d)", g.a_dtype, g.b_dtype, g.batch,
              g.lda, g.ldb, g.epi_mode);
  static int num_sms_all = 0;
  static bool attr_set = false;
  static int forced_bn = 0, forced_stages = 0, forced_epi = 0;
  if (!attr_set) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms_all, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaSuccess;
#define TC_ATTR(G, S)                                                                                                                       \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<8, G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM);           \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<4, G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM);         \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<4, G, S>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    TC_ATTR(false, false) TC_ATTR(false, true) TC_ATTR(true, false) TC_ATTR(true, true)
#undef TC_ATTR
    if (e != cudaSuccess) { dfb_set_error("gemm_tc smem attr: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    // tuning aids (tools/gemm_bn_sweep.py, tools/gemm_replay.py), read once
    if (const char* v = getenv("DFB200_TC_BN")) forced_bn = atoi(v);      // whole-process overrides for tools/gemm_replay.py runs
    if (const char* v = getenv("DFB200_TC_STAGES")) forced_stages = atoi(v);
    if (const char* v = getenv("DFB200_TC_EPI")) forced_epi = atoi(v);
    // Persistent grids deliberately do NOT cover the whole chip.  The training step runs four streams (RGB chain, depth chain,
    // attention branch, weight gradients): a GEMM whose CTAs sit on every SM keeps the other streams' kernels waiting for its
    // whole duration (and a one-wave kernel that cannot place all its CTAs runs a second wave).  Measured on the DFormer-L step
    // (profiles/r02_sm_budget_sweep.txt): forward / dgrad GEMMs on 5/6 of the SMs and the split-K weight-gradient GEMMs (off the
    // critical path) on half of them: 22.9 -> 22.1 ms/step, although the same GEMMs timed alone get 9 % slower.
    g_sms_default[0] = num_sms_all - num_sms_all / 6;
    g_sms_default[1] = (num_sms_all * 49) / 100;
    if (const char* v = getenv("DFB200_TC_SM_RESERVE")) { const int r = atoi(v); if (r > 0 && r < num_sms_all) g_sms_default[0] = g_sms_default[1] = num_sms_all - r; }
    if (const char* v = getenv("DFB200_TC_SM_SPLIT")) { const int r = atoi(v); if (r > 0 && r <= num_sms_all) g_sms_default[1] = r; }
    attr_set = true;
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.m_tiles = dfb_cdiv(g.M, BM);
  p.kb_total = dfb_cdiv(g.K, BK);
  p.a_mn_major = g.transA ? 1 : 0;   // A stored [K, M]  -> M contiguous
  p.b_mn_major = g.transB ? 0 : 1;   // B stored [K, N]  -> N contiguous
  p.C = g.C; p.ldc = g.ldc; p.bias = g.bias;
  p.batch = g.batch; p.strideC = g.strideC;
  p.out_bf16 = g.out_dtype == 1; p.act = g.act; p.act_col_start = g.act_col_start; p.accumulate = g.accumulate;
  p.epi_mode = g.epi_mode; p.has_out2 = (g.epi_mode != 0 && g.out2) ? 1 : 0;
  // wgrad-like problems are split along K; their tile count is multiplied by the split factor later, so only the un-split
  // (forward / dgrad) shapes are tuned for wave quantisation
  const bool will_split = (g.splitk == 0 && g.out_dtype == 0 && g.act == 0 && p.kb_total >= 16) || g.splitk > 1;
  const int budget = g_sms_budget[will_split ? 1 : 0];
  const int num_sms = (budget > 0 && budget <= num_sms_all) ? budget : g_sms_default[will_split ? 1 : 0];
  // bf16 C through TMA: 16-byte aligned base / leading dimension / batch stride
  const bool can_tma_store = p.out_bf16 && !will_split && (g.ldc % 8) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 &&
                             (g.batch == 1 || (g.strideC % 8) == 0);
  // variant: two 192-thread CTAs per SM (EPI_WARPS = 4) when that puts every tile of an un-split problem into ONE wave that the
  // one-CTA-per-SM variant would need two for (M = 9600 layers: 150 tiles on 148 SMs); everything else runs the deep-ring variant
  // (measured per shape with tools/gemm_replay.py, profiles/r02_gemm_replay_variants.txt)
  int epi_warps = 8;
  if (!will_split) {
    const int bn8 = pick_bn(g.N, p.m_tiles, p.kb_total, g.batch, num_sms, p.b_mn_major, 256, can_tma_store);
    const int bn4 = pick_bn(g.N, p.m_tiles, p.kb_total, g.batch, 2 * num_sms, p.b_mn_major, 192, can_tma_store);
    const long t8 = (long)p.m_tiles * dfb_cdiv(g.N, bn8) * g.batch, t4 = (long)p.m_tiles * dfb_cdiv(g.N, bn4) * g.batch;
    if (t8 > num_sms && t4 <= 2L * num_sms) epi_warps = 4;
  }
  if (forced_epi == 4 || forced_epi == 8) epi_warps = forced_epi;
  const int ctas_per_sm = epi_warps == 4 ? 2 : 1;
  const int bn_cap = epi_warps == 4 ? 192 : 256;
  p.BN = will_split ? pick_bn_padding(g.N, bn_cap) : pick_bn(g.N, p.m_tiles, p.kb_total, g.batch, num_sms * ctas_per_sm, p.b_mn_major, bn_cap, can_tma_store);
  if (forced_bn >= 16 && forced_bn <= bn_cap && forced_bn % 16 == 0 && !(p.b_mn_major && forced_bn < 64 && g.N >= 64)) p.BN = forced_bn;
  p.n_tiles = dfb_cdiv(g.N, p.BN);
  p.tma_store = can_tma_store && ((p.BN % 64) == 0 || p.n_tiles == 1);
  // split-K when the output has too few tiles to fill the machine and the reduction is long (wgrad)
  int splits = 1;
  const long tiles = (long)p.m_tiles * p.n_tiles * g.batch;
  if (g.splitk > 1) splits = g.splitk;
  else if (g.splitk == 0 && g.out_dtype == 0 && g.act == 0 && tiles * 2 <= num_sms && p.kb_total >= 16)
    splits = (int)min((long)p.kb_total / 4, (long)(num_sms / tiles));
  if (splits < 1) splits = 1;
  if (splits > 1) DFB_REQUIRE(g.out_dtype == 0 && g.act == 0, "gemm_tc split-K needs fp32 output without activation");
  p.kb_per_split = dfb_cdiv(p.kb_total, splits);
  p.splits = dfb_cdiv(p.kb_total, p.kb_per_split);
  if (p.splits > 1 && !g.accumulate && g.ldc == g.N && (g.batch == 1 || g.strideC == (long)g.M * g.N)) {      // contiguous C: one memset node
    cudaError_t e = cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)g.batch * g.M * g.N, st);
    if (e != cudaSuccess) { dfb_set_error("memset: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
  } else if (p.splits > 1 && !g.accumulate) {
    for (int b = 0; b < g.batch; ++b) {
      cudaError_t e = cudaMemset2DAsync((float*)g.C + (long)b * g.strideC, g.ldc * sizeof(float), 0, (size_t)g.N * sizeof(float), g.M, st);
      if (e != cudaSuccess) { dfb_set_error("memset2d: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    }
  }
  DFB_REQUIRE(p.epi_mode == 0 || p.tma_store, "gemm_tc: fused epilogue %d needs the TMA-store path (BN %d, N %d)", p.epi_mode, p.BN, g.N);
  CUtensorMap tmA, tmB, tmC, tmD, tmE;
  int rc;
  const long sA = g.batch > 1 ? g.strideA : (long)g.lda * (g.transA ? g.K : g.M);
  const long sB = g.batch > 1 ? g.strideB : (long)g.ldb * (g.transB ? g.N : g.K);
  if (g.transA) rc = make_map(&tmA, g.A, g.M, g.K, g.lda, 64, 64, g.batch, sA);          // stored [K, M]
  else rc = make_map(&tmA, g.A, g.K, g.M, g.lda, 64, BM, g.batch, sA);                  // stored [M, K]
  if (rc) return rc;
  if (g.transB) rc = make_map(&tmB, g.B, g.K, g.N, g.ldb, 64, p.BN, g.batch, sB);        // stored [N, K]
  else rc = make_map(&tmB, g.B, g.N, g.K, g.ldb, 64, 64, g.batch, sB);                  // stored [K, N]
  if (rc) return rc;
  if (p.tma_store) {
    rc = make_map(&tmC, g.C, g.N, g.M, g.ldc, 64, 32, g.batch, g.batch > 1 ? g.strideC : (long)g.ldc * g.M);     // [32 rows x 64 columns] boxes
    if (rc) return rc;
  } else {
    tmC = tmA;                                  // unused by the kernel
  }
  if (p.has_out2) {
    rc = make_map(&tmD, g.out2, g.N, g.M, g.ld_out2, 64, 32, 1, (long)g.ld_out2 * g.M);
    if (rc) return rc;
  } else {
    tmD = tmA;                                  // unused by the kernel
  }
  if (p.epi_mode == 1) {
    rc = make_map(&tmE, g.aux, g.N, g.M, g.ld_aux, 64, 32, 1, (long)g.ld_aux * g.M);
    if (rc) return rc;
  } else {
    tmE = tmA;                                  // unused by the kernel
  }
  const long total = tiles * p.splits;
  const int grid = (int)min((long)num_sms * ctas_per_sm, total);
  // accumulator ring in tensor memory: stages `acc_stride` columns apart (64-column granules: the TMA-store epilogue reads whole boxes)
  p.acc_stride = ((p.BN + 63) / 64) * 64;
  const int tmem_budget = epi_warps == 4 ? 256 : 512;
  p.acc_stages = (2 * p.acc_stride <= tmem_budget && total > grid) ? 2 : 1;     // one tile per CTA needs no second stage
  p.tmem_cols = next_pow2_cols(p.acc_stages * p.acc_stride);
  // pipeline depth: as deep as the variant's shared-memory budget allows, but no deeper than the k-blocks this CTA will ever load
  const int b_bytes = p.b_mn_major ? ((p.BN + 63) / 64) * 8192 : ((p.BN * 128 + 1023) / 1024) * 1024;   // 1024-B aligned (swizzle atom)
  p.stage_bytes = A_STAGE_BYTES + b_bytes;
  const int fixed0 = epi_warps * EPI_WARP_BYTES * (1 + (p.has_out2 ? 1 : 0) + (p.epi_mode == 1 ? 1 : 0)) + 1024 + 512;
  const int budget_smem = epi_warps == 4 ? SMALL_SMEM : MAX_SMEM;
  const long kb_per_cta = (long)dfb_cdiv(total, grid) * p.kb_per_split;
  // bias slices in shared memory (fetched before the accumulator is ready): only where the 4 / 8 KB do not cost this launch a pipeline
  // stage it would use (the deep-K, many-tile GEMMs hide the bias fetch behind the next tile's main loop anyway)
  const int bias_bytes = epi_warps * EPI_BIAS_FLOATS * 4;
  const long st_without = min((long)min((budget_smem - fixed0) / p.stage_bytes, MAX_STAGES), kb_per_cta);
  const long st_with = min((long)min((budget_smem - fixed0 - bias_bytes) / p.stage_bytes, MAX_STAGES), kb_per_cta);
  p.bias_smem = (p.tma_store && p.bias != nullptr && st_with >= 1 && st_with == st_without) ? 1 : 0;
  const int fixed = fixed0 + (p.bias_smem ? bias_bytes : 0);
  p.stages = (budget_smem - fixed) / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.stages > kb_per_cta) p.stages = (int)kb_per_cta;
  if (forced_stages > 0 && forced_stages < p.stages) p.stages = forced_stages;
  DFB_REQUIRE(p.stages >= 1, "gemm_tc: tile %d x %d does not fit the shared-memory budget", BM, p.BN);
  const int smem_bytes = p.stages * p.stage_bytes + fixed;
  const bool gate = p.epi_mode == 1, bsm = p.bias_smem != 0;
#define TC_LAUNCH(G, S)                                                                                                   \
  do {                                                                                                                    \
    if (epi_warps == 4) dfb_launch(gemm_tc_kernel<4, G, S>, grid, 64 + 32 * 4, smem_bytes, st, tmA, tmB, tmC, tmD, tmE, p);  \
    else dfb_launch(gemm_tc_kernel<8, G, S>, grid, 64 + 32 * 8, smem_bytes, st, tmA, tmB, tmC, tmD, tmE, p);                 \
  } while (0)
  if (gate && bsm) TC_LAUNCH(true, true);
  else if (gate) TC_LAUNCH(true, false);
  else if (bsm) TC_LAUNCH(false, true);
  else TC_LAUNCH(false, false);
#undef TC_LAUNCH
  return dfb_check_launch("gemm_tc");
}
