// Handle, error plumbing and buffer binding of liblls.
#include "lls_internal.h"
#include <cstring>

namespace lls {
static thread_local char g_err[512] = "";
void set_last_error(const char* file, int line, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s:%d: %s", file, line, msg);
}
}  // namespace lls

extern "C" const char* lls_last_error(void) { return lls::g_err; }
extern "C" int lls_abi_version(void) { return 1; }
