// Library-wide C-ABI plumbing: version, thread-local error string, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace drg {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace drg

extern "C" int drg_version(void) { return 100; }
extern "C" const char* drg_last_error(void) { return drg::g_err; }
extern "C" unsigned long long drg_launch_count(void) { return drg::g_launches.load(); }
