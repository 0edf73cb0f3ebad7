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

// Programmatic dependent launch.  A kernel launched through launch_dependent() may become resident while its predecessor in
// the stream is still running (as SMs free up); it must call grid_dependency_wait() before it touches global memory the
// predecessor writes or reads -- the call returns once the predecessor has completed and its writes are visible.  A kernel
// that calls grid_dependency_trigger() early lets such a successor start its prologue (barrier init, TMEM allocation,
// tensor-map prefetch) under this kernel's tail instead of after the launch gap; for a normally launched successor the
// trigger is a no-op, and so is the wait in a normally launched kernel.
#ifdef __CUDACC__
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

template <typename T>
inline T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
inline T round_up(T a, T b) { return ceil_div(a, b) * b; }

}  // namespace memb
