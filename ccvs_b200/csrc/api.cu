// api.cu — version / error-string plumbing of the C ABI (include/ccvsq.h).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"

namespace ccvsq {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// (kernel, device) -> largest dynamic shared-memory size already granted
struct SmemGrant { const void* kern; int dev; size_t bytes; };
static SmemGrant g_grants[128];
static int g_num_grants = 0;
static std::mutex g_grant_mu;

int enable_smem_impl(const void* kern, size_t bytes) {
  CCVSQ_REQUIRE(bytes <= 227 * 1024, CCVSQ_UNSUPPORTED,
                "kernel needs %zu bytes of shared memory (> 227 KiB)", bytes);
  if (bytes <= 48 * 1024) return CCVSQ_OK;
  int dev = 0;
  CCVSQ_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_grant_mu);
  SmemGrant* slot = nullptr;
  for (int i = 0; i < g_num_grants; ++i)
    if (g_grants[i].kern == kern && g_grants[i].dev == dev) { slot = &g_grants[i]; break; }
  if (slot && slot->bytes >= bytes) return CCVSQ_OK;
  // (the hardware limit is 227 KiB minus the kernel's static shared memory: ask for what is needed)
  CCVSQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (!slot && g_num_grants < 128) slot = &g_grants[g_num_grants++];
  if (slot) { slot->kern = kern; slot->dev = dev; slot->bytes = bytes; }
  return CCVSQ_OK;
}
int pdl_off_mask() {
  static const int m = [] { const char* e = getenv("CCVSQ_NO_PDL_KERNELS"); return e ? atoi(e) : 0; }();
  return m;
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CCVSQ_NO_PDL"); return !(e && atoi(e) != 0); }();
  return on;
}
}  // namespace ccvsq

extern "C" int ccvsq_version(void) { return CCVSQ_VERSION; }
extern "C" const char* ccvsq_last_error(void) { return ccvsq::g_err; }
