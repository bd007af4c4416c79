// C-ABI glue: error reporting, version, GEMM backend dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "dfb200_internal.h"

static thread_local char g_err[512] = "";

void dfb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool dfb_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DFB200_PDL"); v = e ? atoi(e) : 1; }
  return v != 0;
}

static long g_launches = 0;
extern "C" long dfb200_launch_count(void) { return g_launches; }

int dfb_check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    dfb_set_error("%s: %s", what, cudaGetErrorString(e));
    return DFB_ERR_CUDA;
  }
  return DFB_OK;
}

extern "C" const char* dfb200_last_error(void) { return g_err; }
extern "C" int dfb200_version(void) { return 100; }

extern "C" int dfb200_gemm(const dfb200_gemm_args* a, void* stream) {
  if (!a) { dfb_set_error("gemm: null args"); return DFB_ERR_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->M <= 0 || a->N <= 0) return DFB_OK;
  if (a->backend == DFB200_BACKEND_TCGEN05) return dfb_gemm_tc(*a, st);
  if (a->backend == DFB200_BACKEND_AUTO && dfb_gemm_tc_supported(*a)) return dfb_gemm_tc(*a, st);
  if (a->epi_mode != 0) { dfb_set_error("gemm: fused epilogue %d needs the tcgen05 backend (bf16 operands and output, aligned C / aux / out2)", a->epi_mode); return DFB_ERR_UNSUPPORTED; }
  return dfb_gemm_simt(*a, st);
}
