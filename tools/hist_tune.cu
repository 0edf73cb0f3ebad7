// Tuning microbench for the RED scatter kernel (not product code): sweeps block size, unroll,
// occupancy and grid size on 10M uniform events at 640x480.  nvcc -arch=sm_100a -O3.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

__device__ __forceinline__ void ld256(const double* p, double& x, double& y, double& t, double& q) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(t), "=d"(q) : "l"(p));
}
__device__ __forceinline__ void ld256_plain(const double* p, double& x, double& y, double& t, double& q) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(t), "=d"(q) : "l"(p));
}

template <int THREADS, int UNROLL, int MINB, int MODE>
__global__ void __launch_bounds__(THREADS, MINB) scatter(const double* __restrict__ ev, long long n, int W, long long npix,
                                                         unsigned int* __restrict__ acc) {
  const long long step = (long long)gridDim.x * THREADS * UNROLL;
  for (long long base = blockIdx.x * (long long)(THREADS * UNROLL); base < n; base += step) {
    double x[UNROLL], y[UNROLL], t[UNROLL], p[UNROLL];
    bool live[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      long long r = base + u * THREADS + threadIdx.x;
      live[u] = r < n;
      if (live[u]) { if (MODE == 1) ld256_plain(ev + 4 * r, x[u], y[u], t[u], p[u]); else ld256(ev + 4 * r, x[u], y[u], t[u], p[u]); }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (!live[u]) continue;
      const bool pos = p[u] == 1.0, neg = p[u] == -1.0;
      if (!(pos || neg)) continue;
      if (!(fabs(x[u]) < 1.0995116e12) || !(fabs(y[u]) < 1.0995116e12)) continue;
      long long i = __double2ll_rz(x[u]) + (long long)W * __double2ll_rz(y[u]);
      if (i < -npix || i >= npix) continue;
      if (i < 0) i += npix;
      if (MODE == 2) {  // packed 16-bit halves in one word per pixel
        atomicAdd(acc + i, pos ? 1u : 0x10000u);
      } else {
        atomicAdd(acc + (neg ? npix : 0) + i, 1u);
      }
    }
  }
}

// read-only ceiling: same loads, no atomics
template <int THREADS, int UNROLL, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) readonly(const double* __restrict__ ev, long long n, double* sink) {
  const long long step = (long long)gridDim.x * THREADS * UNROLL;
  double s = 0;
  for (long long base = blockIdx.x * (long long)(THREADS * UNROLL); base < n; base += step) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      long long r = base + u * THREADS + threadIdx.x;
      if (r < n) { double x, y, t, p; ld256(ev + 4 * r, x, y, t, p); s += x + p; }
    }
  }
  if (s == 12345.678) *sink = s;
}

__global__ void gen(double* ev, long long n, int W, int H) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  ev[4 * i] = (double)(h % W);
  ev[4 * i + 1] = (double)((h >> 20) % H);
  ev[4 * i + 2] = (double)i;
  ev[4 * i + 3] = (h >> 50 & 1) ? 1.0 : -1.0;
}

template <class F>
float bench(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  std::vector<float> ts;
  for (int i = 0; i < 15; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); ts.push_back(ms); }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

int main() {
  const long long n = 10000000; const int W = 640, H = 480; const long long npix = (long long)W * H;
  double* ev; unsigned int* acc; double* sink;
  cudaMalloc(&ev, n * 32); cudaMalloc(&acc, npix * 8); cudaMalloc(&sink, 8);
  gen<<<(n + 255) / 256, 256>>>(ev, n, W, H);
  cudaMemset(acc, 0, npix * 8);
  cudaDeviceSynchronize();
  int sms = 148;
#define RUN(T, U, MB, MODE, GM)                                                                 \
  {                                                                                             \
    int grid = sms * GM;                                                                        \
    float ms = bench([&] { scatter<T, U, MB, MODE><<<grid, T>>>(ev, n, W, npix, acc); });       \
    printf("scatter T=%d U=%d minB=%d mode=%d grid=%dxSM : %.1f us  %.0f GB/s\n", T, U, MB, MODE, GM, ms * 1e3, n * 32 / ms / 1e6); \
  }
#define RO(T, U, MB, GM)                                                                        \
  {                                                                                             \
    int grid = sms * GM;                                                                        \
    float ms = bench([&] { readonly<T, U, MB><<<grid, T>>>(ev, n, sink); });                    \
    printf("readonly T=%d U=%d minB=%d grid=%dxSM : %.1f us  %.0f GB/s\n", T, U, MB, GM, ms * 1e3, n * 32 / ms / 1e6); \
  }
  RO(256, 4, 8, 8) RO(256, 8, 4, 8) RO(512, 4, 4, 4) RO(256, 4, 8, 16) RO(256, 2, 8, 32)
  RUN(256, 4, 1, 0, 8) RUN(256, 4, 8, 0, 8) RUN(256, 4, 8, 0, 16) RUN(256, 4, 8, 0, 32)
  RUN(256, 8, 4, 0, 8) RUN(256, 8, 4, 0, 4) RUN(256, 8, 6, 0, 6) RUN(256, 2, 8, 0, 16) RUN(256, 2, 8, 0, 64)
  RUN(512, 4, 4, 0, 4) RUN(512, 4, 4, 0, 8) RUN(1024, 2, 2, 0, 2) RUN(1024, 4, 2, 0, 2) RUN(128, 4, 16, 0, 16) RUN(128, 8, 8, 0, 16)
  RUN(256, 4, 8, 1, 8) RUN(256, 4, 8, 2, 8) RUN(256, 8, 4, 2, 8) RUN(256, 1, 8, 0, 64) RUN(256, 1, 8, 0, 256)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
