// C-ABI plumbing: error reporting, one-time initialisation.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <string.h>

#include "../../include/fvqa_debug.h"
#include "common.cuh"

namespace fvqa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return FVQA_ERR_CUDA;
  }
  return FVQA_OK;
}

int gemm_init();
int attn_init();

static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load() != 0; }

}  // namespace fvqa

extern "C" int fvqa_abi_version(void) { return FVQA_ABI_VERSION; }
extern "C" int fvqa_operand_dtype(void) { return FVQA_OPERAND_DTYPE; }
extern "C" const char* fvqa_last_error(void) { return fvqa::g_err; }

/* Tuning hook (fvqa_debug.h): 1 (default; FVQA_PDL=0 in the environment disables it at fvqa_init) = hot kernels are launched with
 * programmatic dependent launch. Returns the previous setting. */
extern "C" int fvqa_debug_pdl(int on) {
  const int prev = fvqa::g_pdl.load();
  fvqa::g_pdl = on ? 1 : 0;
  return prev;
}

extern "C" int fvqa_init(void) {
  if (const char* e = getenv("FVQA_PDL")) fvqa::g_pdl = (e[0] == '0') ? 0 : 1;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    fvqa::set_error("fvqa_init: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
    return FVQA_ERR_CUDA;
  }
  int rc = fvqa::gemm_init();
  if (rc) return rc;
  return fvqa::attn_init();
}
