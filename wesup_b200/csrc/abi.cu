// ABI bookkeeping: version and the thread-local error message.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace wesup {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace wesup

extern "C" int wesup_abi_version(void) { return WESUP_ABI_VERSION; }
extern "C" const char *wesup_last_error(void) { return wesup::g_err; }
extern "C" unsigned long long wesup_kernel_launches(void) { return wesup::g_launches.load(); }
