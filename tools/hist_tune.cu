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


// ---- bulk-copy staged variant: a producer warp streams CHUNK-event blocks into a smem ring with cp.async.bulk
// (mbarrier complete_tx), consumer warps read rows from smem and fire REDs; per-warp release, no CTA-wide sync.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int CWARPS, int STAGES, int CHUNK>
__global__ void __launch_bounds__((CWARPS + 1) * 32, 1) scatter_bulk(const double* __restrict__ ev, long long n, int W, long long npix,
                                                                   unsigned int* __restrict__ acc) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* ring = reinterpret_cast<double*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK * 32);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nchunks = (n + CHUNK - 1) / CHUNK;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == CWARPS) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        const long long r0 = c * CHUNK;
        const uint32_t bytes = (uint32_t)(min((long long)CHUNK, n - r0) * 32);
        mbar_expect(&full[s], bytes);
        bulk_load(ring + (size_t)s * CHUNK * 4, ev + 4 * r0, bytes, &full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    int s = 0; uint32_t ph = 0;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
      mbar_wait(&full[s], ph);
      const long long r0 = c * CHUNK;
      const int cnt = (int)min((long long)CHUNK, n - r0);
      const double* st = ring + (size_t)s * CHUNK * 4;
#pragma unroll 4
      for (int i = warp * 32 + lane; i < cnt; i += CWARPS * 32) {
        const double2 a = *reinterpret_cast<const double2*>(st + 4 * i);
        const double2 b = *reinterpret_cast<const double2*>(st + 4 * i + 2);
        const bool pos = b.y == 1.0, neg = b.y == -1.0;
        if (!(pos || neg)) continue;
        if (!(fabs(a.x) < 1.0995116e12) || !(fabs(a.y) < 1.0995116e12)) continue;
        long long idx = __double2ll_rz(a.x) + (long long)W * __double2ll_rz(a.y);
        if (idx < -npix || idx >= npix) continue;
        if (idx < 0) idx += npix;
        atomicAdd(acc + (neg ? npix : 0) + idx, 1u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
}


// ---- ceilings: REDs alone (no event loads), to L2 and to distributed shared memory of an 8-CTA cluster
__device__ __forceinline__ unsigned long long mix(unsigned long long i) {
  unsigned long long h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32; return h;
}
__global__ void __launch_bounds__(256) red_only(long long n, long long nslots, unsigned int* __restrict__ acc) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    atomicAdd(acc + mix(i) % nslots, 1u);
}
template <int WORDS>  // words of smem per CTA; cluster of 8 holds 8*WORDS counters
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(1024, 1) dsmem_only(long long n, unsigned int* __restrict__ out) {
  extern __shared__ unsigned int cnt[];
  for (int i = threadIdx.x; i < WORDS; i += 1024) cnt[i] = 0;
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(cnt);
  for (long long i = blockIdx.x * 1024ll + threadIdx.x; i < n; i += (long long)gridDim.x * 1024) {
    const unsigned long long h = mix(i) % (8ull * WORDS);
    const uint32_t rank = (uint32_t)(h / WORDS), off = (uint32_t)(h % WORDS);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + off * 4), "r"(rank));
    asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  unsigned int s = 0;
  for (int i = threadIdx.x; i < WORDS; i += 1024) s += cnt[i];
  if (s == 0xdeadbeef) out[0] = s;
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
  RUN(256, 4, 1, 0, 8)
  RUN(256, 8, 4, 0, 8) RUN(256, 8, 4, 0, 4) RUN(256, 8, 6, 0, 6) RUN(256, 2, 8, 0, 16) RUN(256, 2, 8, 0, 64)
  RUN(512, 4, 4, 0, 4) RUN(512, 4, 4, 0, 8) RUN(1024, 2, 2, 0, 2) RUN(1024, 4, 2, 0, 2) RUN(128, 4, 16, 0, 16) RUN(128, 8, 8, 0, 16)
  RUN(256, 4, 8, 1, 8) RUN(256, 4, 8, 2, 8) RUN(256, 8, 4, 2, 8) RUN(256, 1, 8, 0, 64) RUN(256, 1, 8, 0, 256)

#define RUNB(CW, ST, CH, GM)                                                                    \
  {                                                                                             \
    int grid = sms * GM; size_t sm = (size_t)ST * CH * 32 + 256;                                \
    cudaFuncSetAttribute(scatter_bulk<CW, ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    float ms = bench([&] { scatter_bulk<CW, ST, CH><<<grid, (CW + 1) * 32, sm>>>(ev, n, W, npix, acc); }); \
    printf("bulk cwarps=%d stages=%d chunk=%d grid=%dxSM smem=%zu : %.1f us  %.0f GB/s\n", CW, ST, CH, GM, sm, ms * 1e3, n * 32 / ms / 1e6); \
  }
  RUNB(8, 4, 512, 2) RUNB(8, 4, 512, 3) RUNB(16, 4, 1024, 1) RUNB(16, 6, 1024, 1) RUNB(8, 3, 1024, 2) RUNB(8, 6, 512, 2) RUNB(4, 4, 256, 4) RUNB(4, 4, 256, 6) RUNB(8, 8, 256, 3) RUNB(31, 6, 1024, 1) RUNB(16, 3, 512, 4)

  {
    float ms = bench([&] { red_only<<<sms * 64, 256>>>(n, npix * 2, acc); });
    printf("red_only L2 random, 10M: %.1f us\n", ms * 1e3);
    ms = bench([&] { red_only<<<sms * 64, 256>>>(n, 4096, acc); });
    printf("red_only L2 4096 hot slots, 10M: %.1f us\n", ms * 1e3);
    constexpr int WORDS = 38400;  // 8 * 38400 = 307200 packed pixels
    cudaFuncSetAttribute(dsmem_only<WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4);
    ms = bench([&] { dsmem_only<WORDS><<<144, 1024, WORDS * 4>>>(n, acc); });
    printf("dsmem_only cluster8 random, 10M: %.1f us\n", ms * 1e3);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
