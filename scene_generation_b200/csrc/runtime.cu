// Runtime entry points of libsg_b200: error reporting, version, launch accounting.
#include <atomic>
#include <stdarg.h>
#include "common.cuh"
#include "../../include/sg_b200.h"

thread_local char sg_err_buf[512] = {0};
static std::atomic<unsigned long long> g_launches{0};

int sg_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(sg_err_buf, sizeof(sg_err_buf), fmt, ap);
  va_end(ap);
  return code;
}

void sg_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" const char* sg_last_error(void) { return sg_err_buf; }
extern "C" const char* sg_version(void) { return "sg_b200 0.1.0 (sm_100a)"; }
extern "C" int sg_arch(void) { return 100; }
extern "C" unsigned long long sg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void sg_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }
// launches replayed from a captured CUDA graph (the host enqueued them once, at capture): the caller accounts them
extern "C" void sg_add_launch_count(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
