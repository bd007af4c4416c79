#pragma once
#include <cuda_runtime.h>
#include "../../include/dfb200.h"

int dfb_gemm_simt(const dfb200_gemm_args& g, cudaStream_t st);
int dfb_gemm_tc(const dfb200_gemm_args& g, cudaStream_t st);
bool dfb_gemm_tc_supported(const dfb200_gemm_args& g);
