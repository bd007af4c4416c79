#pragma once
#include <cuda_runtime.h>
#include "../../include/dfb200.h"

int dfb_gemm_simt(const dfb200_gemm_args& g, cudaStream_t st);
int dfb_gemm_tc(const dfb200_gemm_args& g, cudaStream_t st);
bool dfb_gemm_tc_supported(const dfb200_gemm_args& g);

// TMA-fed bf16 depthwise 7x7 (dw7.cu): y = dw7x7(x) + bias (flip = 0) or the data gradient dw7x7^T (flip = 1); weight/bias gradient
int dfb_dw7_conv(const void* x, const float* weight, const float* bias, int B, int H, int W, int C, int flip, void* y, cudaStream_t st);
int dfb_dw7_wgrad(const void* dz, const void* x, int B, int H, int W, int C, float* dweight, float* dbias, cudaStream_t st);

// tensor-core (mma.sync bf16) Global Awareness Attention core (gaa_mma.cu); same buffers as dfb200_gaa_fused_fwd / _bwd
int dfb_gaa_mma_fwd(const void* m, const void* kv, int B, int HW, int heads, int d, float* out, float* lse, float* scratch, int* counters, cudaStream_t st);
int dfb_gaa_mma_bwd(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int B, int HW, int heads, int d, float* dm,
                    void* dkv, float* dkv_colsum, float* dm_colsum, void* dm_lo, cudaStream_t st);
