// Thread-local error string of the C ABI (the library's only mutable global state).  Its own translation unit so that the
// test-only library (csrc/testlib) links the same plumbing without the product entry points.
#include <stdlib.h>

#include "mg_common.cuh"

namespace mg {

static thread_local char g_error[512] = "";

char *error_buffer() { return g_error; }

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("MG_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

}  // namespace mg
