// api.cu — version / error-string plumbing of the C ABI (include/ccvsq.h).
#include <stdarg.h>
#include <stdio.h>
#include "common.cuh"

namespace ccvsq {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace ccvsq

extern "C" int ccvsq_version(void) { return CCVSQ_VERSION; }
extern "C" const char* ccvsq_last_error(void) { return ccvsq::g_err; }
