// Library-wide state: error string, launch counter, device properties.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};

void egb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void egb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int egb_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

extern "C" {
const char* egb_last_error(void) { return g_err; }
int egb_version(void) { return 1; }
int64_t egb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
}
