// ABI version + thread-local error text.
#include <stdarg.h>

#include "common.cuh"

namespace dqrm {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dqrm

extern "C" int dqrm_abi_version(void) { return DQRM_ABI_VERSION; }
extern "C" const char* dqrm_last_error(void) { return dqrm::g_err; }
