// Generic SIMT GEMM (fp32 accumulate) -- the exact-fp32 path and the fallback shape coverage
// for the tcgen05 kernel in gemm_tc.cu.  C[M,N] = epi(opA(A) * opB(B) + bias).
//   transA = 0: A stored [M,K] (lda)   transA = 1: A stored [K,M] (lda)
//   transB = 0: B stored [K,N] (ldb)   transB = 1: B stored [N,K] (ldb)   (nn.Linear weight)
#include "common.cuh"
#include "dfb200_internal.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename TAe, typename TBe, typename TO, bool TA, bool TB>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(dfb200_gemm_args g, int splits, int kchunk) {
  pdl_sync();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int zb = blockIdx.z / splits, zs = blockIdx.z % splits;
  const int bi = g.batch_inner > 1 ? g.batch_inner : 1;
  const int zo = zb / bi, zi = zb % bi;
  const TAe* __restrict__ A = reinterpret_cast<const TAe*>(g.A) + (long)zo * g.strideA + (long)zi * g.strideA_in;
  const TBe* __restrict__ B = reinterpret_cast<const TBe*>(g.B) + (long)zo * g.strideB + (long)zi * g.strideB_in;
  TO* __restrict__ C = reinterpret_cast<TO*>(g.C) + (long)zo * g.strideC + (long)zi * g.strideC_in;
  const float alpha = g.alpha == 0.f ? 1.f : g.alpha;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = zs * kchunk, k_end = min(g.K, k_begin + kchunk);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // A tile: BM x BK
#pragma unroll
    for (int i = tid; i < BM * BK; i += NT) {
      int m, k;
      if (TA) { m = i % BM; k = i / BM; } else { k = i % BK; m = i / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < k_end) v = to_f(TA ? A[(long)gk * g.lda + gm] : A[(long)gm * g.lda + gk]);
      As[k][m] = v;
    }
#pragma unroll
    for (int i = tid; i < BN * BK; i += NT) {
      int n, k;
      if (TB) { k = i % BK; n = i / BK; } else { n = i % BN; k = i / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < k_end) v = to_f(TB ? B[(long)gn * g.ldb + gk] : B[(long)gk * g.ldb + gn]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j] * alpha;
      if (splits > 1) {   // split-K: raw partial sums, bias from split 0; no activation allowed
        if (zs == 0 && g.bias) v += g.bias[n];
        atomicAdd(reinterpret_cast<float*>(C) + (long)m * g.ldc + n, v);
        continue;
      }
      if (g.bias) v += g.bias[n];
      if (n >= g.act_col_start) {
        if (g.act == 1) v = gelu_f(v);
        else if (g.act == 2) v = fmaxf(v, 0.f);
      }
      TO* dst = C + (long)m * g.ldc + n;
      if (g.accumulate) v += to_f(*dst);
      *dst = from_f<TO>(v);
    }
  }
}

template <typename TAe, typename TBe, typename TO>
int launch(const dfb200_gemm_args& g, cudaStream_t st) {
  int splits = 1;
  const int nbatch = g.batch * (g.batch_inner > 1 ? g.batch_inner : 1);
  const long tiles = (long)dfb_cdiv(g.M, BM) * dfb_cdiv(g.N, BN) * nbatch;
  if (g.splitk > 1) splits = g.splitk;
  else if (g.splitk == 0 && g.out_dtype == 0 && g.act == 0 && g.alpha == 0.f && tiles < 148 && g.K >= 2048) {
    splits = (int)min((long)dfb_cdiv(g.K, 512), (2 * 148 + tiles - 1) / tiles);
  }
  if (splits > 1) {
    DFB_REQUIRE(g.out_dtype == 0 && g.act == 0, "split-K gemm needs fp32 output without activation");
    if (!g.accumulate) {
      const int bi = g.batch_inner > 1 ? g.batch_inner : 1;
      for (int b = 0; b < nbatch; ++b) {
        float* c = (float*)g.C + (long)(b / bi) * g.strideC + (long)(b % bi) * g.strideC_in;
        cudaError_t e = cudaMemset2DAsync(c, g.ldc * sizeof(float), 0, (size_t)g.N * sizeof(float), g.M, st);
        if (e != cudaSuccess) { dfb_set_error("memset2d: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
      }
    }
  }
  int kchunk = dfb_cdiv(dfb_cdiv(g.K, splits), BK) * BK;
  if (kchunk == 0) kchunk = BK;
  dim3 grid(dfb_cdiv(g.N, BN), dfb_cdiv(g.M, BM), nbatch * splits);
  DFB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm_simt grid too large (M=%d batch=%d)", g.M, g.batch);
#define L(TA, TB) dfb_launch(gemm_simt_kernel<TAe, TBe, TO, TA, TB>, grid, NT, 0, st, g, splits, kchunk)
  if (g.transA) { if (g.transB) L(true, true); else L(true, false); }
  else { if (g.transB) L(false, true); else L(false, false); }
#undef L
  return dfb_check_launch("gemm_simt");
}

}  // namespace

int dfb_gemm_simt(const dfb200_gemm_args& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return DFB_OK;
  DFB_REQUIRE((unsigned)g.a_dtype < 2 && (unsigned)g.b_dtype < 2 && (unsigned)g.out_dtype < 2, "gemm: bad dtype");
  const int key = g.a_dtype * 4 + g.b_dtype * 2 + g.out_dtype;
  switch (key) {
    case 0: return launch<float, float, float>(g, st);
    case 1: return launch<float, float, bf16>(g, st);
    case 2: return launch<float, bf16, float>(g, st);
    case 3: return launch<float, bf16, bf16>(g, st);
    case 4: return launch<bf16, float, float>(g, st);
    case 5: return launch<bf16, float, bf16>(g, st);
    case 6: return launch<bf16, bf16, float>(g, st);
    case 7: return launch<bf16, bf16, bf16>(g, st);
  }
  dfb_set_error("gemm_simt: unsupported dtype combination %d/%d/%d", g.a_dtype, g.b_dtype, g.out_dtype);
  return DFB_ERR_ARG;
}
