// Error reporting for the C ABI (thread-local last-error string) and version query.
#include "common.cuh"
#include "../../include/margipose_b200.h"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void mp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {
int mp_abi_version(void) { return 1; }
const char* mp_last_error(void) { return g_err; }
}
