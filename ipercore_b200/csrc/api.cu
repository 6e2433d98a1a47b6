// C-ABI plumbing: thread-local last-error string and ABI version.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "iper_b200.h"

namespace iper {
static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}
}  // namespace iper

extern "C" const char* iper_last_error(void) { return iper::g_last_error; }
extern "C" int iper_abi_version(void) { return 1; }
