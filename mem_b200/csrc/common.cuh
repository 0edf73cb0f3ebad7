// Shared host/device helpers for libmemb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "memb.h"

namespace memb {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define MEMB_CUDA_OK(expr)                                                                    \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return ::memb::fail(MEMB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                                \
  } while (0)

#define MEMB_LAUNCH_OK(what)                                                                 \
  do {                                                                                        \
    ::memb::count_launch();                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess)                                                                   \
      return ::memb::fail(MEMB_ECUDA, "launch of %s failed: %s (%s:%d)", what,                \
                          cudaGetErrorString(e__), __FILE__, __LINE__);                       \
  } while (0)

#define MEMB_REQUIRE(cond, ...)                                 \
  do {                                                          \
    if (!(cond)) return ::memb::fail(MEMB_EINVAL, __VA_ARGS__); \
  } while (0)

inline int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

template <typename T>
inline T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
inline T round_up(T a, T b) { return ceil_div(a, b) * b; }

}  // namespace memb
