// Error plumbing and library-level entry points of libipdm_b200.so (include/ipdm_b200.h).
#include "common.cuh"

#include <cstring>

namespace ipdm {

static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

}  // namespace ipdm

extern "C" const char* ipdm_last_error(void) { return ipdm::last_error(); }
extern "C" int ipdm_abi_version(void) { return 1; }
extern "C" unsigned long long ipdm_launch_count(void) { return ipdm::g_launch_count; }
extern "C" void ipdm_launch_count_reset(void) { ipdm::g_launch_count = 0; }
