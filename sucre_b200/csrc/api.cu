// ABI bookkeeping: version and the thread-local error string.
#include <string>

#include "common.cuh"

namespace sucre {
static thread_local std::string g_error;

int set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return 1;
}
void clear_error() { g_error.clear(); }
}  // namespace sucre

extern "C" int sucre_abi_version(void) { return SUCRE_ABI_VERSION; }
extern "C" int sucre_segment_views(void) { return SUCRE_SEGMENT_VIEWS; }
extern "C" const char* sucre_last_error(void) { return sucre::g_error.c_str(); }
