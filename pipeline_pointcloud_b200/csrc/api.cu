// Library-level entry points: version and the thread-local error string.
#include <stdarg.h>

#include "common.cuh"

namespace b2s {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace b2s

extern "C" {
int b2s_version(void) { return B2S_VERSION; }
const char* b2s_last_error(void) { return b2s::g_err; }
}
