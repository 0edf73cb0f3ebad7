// Event stream -> polarity-count image on sm_100a.
//
// Replaces EventArrToImg.__call__ (reference mem/datasets.py:566-595).  HBM-bound
// integer work: one 32-byte float64[4] row per event is read exactly once with a
// single 256-bit load (LDG.E.256), the scatter goes either to L2-resident u32
// accumulators with fire-and-forget RED (one long stream) or to a shared-memory
// privatised sensor tile (ragged training batches), and the uint8 image is
// written once, coalesced.  Counters are wider than a byte and reduced modulo 256
// at the end, which is exactly numpy's uint8 wrap.
#include <algorithm>
#include <climits>
#include <cstdlib>


#include "common.cuh"

namespace memb {
namespace hist {

constexpr int kHeaderBytes = 256;
constexpr int kThreads = 256;
constexpr int kUnroll = 2;   // measured on B200: short bursts + a large grid beat deep unrolling (tools/hist_tune.cu)
constexpr int kTileThreads = 1024;
constexpr int kTileMaxWords = 50 * 1024;        // 200 KB of packed u16x2 counters per CTA
constexpr int kTileUnroll = 4;
constexpr int kTileChunk = kTileThreads * 60;   // 61440 rows between folds (< 65535 - 255)

struct Header {
  int oob;            // set when an event fell outside [-H*W, H*W)
  int pad;
  long long max_x;    // memb_hist_extent
  long long max_y;
};

struct Event { double x, y, t, p; };

// Programmatic dependent launch (init -> scatter -> finalize of the GLOBAL strategy): a kernel launched with the
// programmatic-serialisation attribute may start while its predecessor drains; pdl_wait() blocks until the predecessor
// has completed and its writes are visible (a no-op for an ordinary launch), pdl_launch_dependents() lets the successor
// start as soon as every CTA of this grid has passed the call.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static_assert(sizeof(memb_event_aug) == 64, "memb_event_aug is part of the ABI: 64 bytes");

template <bool kAligned>
__device__ __forceinline__ Event load_event(const double* __restrict__ ev, long long row) {
  Event e;
  const double* p = ev + 4 * row;
  if constexpr (kAligned) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(e.x), "=d"(e.y), "=d"(e.t), "=d"(e.p) : "l"(p));
  } else {
    e.x = __ldg(p); e.y = __ldg(p + 1); e.t = __ldg(p + 2); e.p = __ldg(p + 3);
  }
  return e;
}

// Flat pixel index following numpy: trunc toward zero, x + W*y, one negative wrap.
// Returns false (and leaves idx untouched) when the reference would raise IndexError.
__device__ __forceinline__ bool pixel_index(double x, double y, int W, long long npix, long long& idx) {
  // |coord| >= 2^40 (or NaN) can never index a sensor; numpy's cast gives INT64_MIN there.
  if (!(fabs(x) < 1.0995116e12) || !(fabs(y) < 1.0995116e12)) return false;
  long long i = __double2ll_rz(x) + (long long)W * __double2ll_rz(y);
  if (i < -npix || i >= npix) return false;
  idx = i < 0 ? i + npix : i;
  return true;
}

// Same result, common case first: a row inside the sensor needs four compares and two 32-bit conversions.
__device__ __forceinline__ bool pixel_index_fast(double x, double y, int W, int H, long long npix, int& idx) {
  if (x >= 0.0 && x < (double)W && y >= 0.0 && y < (double)H) {
    idx = __double2int_rz(x) + W * __double2int_rz(y);
    return true;
  }
  long long i = 0;
  const bool ok = pixel_index(x, y, W, npix, i);
  idx = (int)i;
  return ok;
}

// ... also returning the integer coordinates (for the 2-D granules of the HYBRID strategy)
__device__ __forceinline__ bool pixel_xy_fast(double x, double y, int W, int H, long long npix, int& idx, int& xi, int& yi) {
  if (x >= 0.0 && x < (double)W && y >= 0.0 && y < (double)H) {
    xi = __double2int_rz(x);
    yi = __double2int_rz(y);
    idx = xi + W * yi;
    return true;
  }
  long long i = 0;
  const bool ok = pixel_index(x, y, W, npix, i);
  idx = (int)i;
  yi = ok ? idx / W : 0;
  xi = ok ? idx - yi * W : 0;
  return ok;
}

__device__ __forceinline__ unsigned long long order_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k ^ 0x8000000000000000ull) : ~k;
  return __longlong_as_double((long long)b);
}

// ---------------------------------------------------------------- fused event-space augmentation
// One correctly rounded float64 operation per reference statement, in the reference's order
// (datasets.py:482-484 scale, :603-606 time flip, :518-519 x flip, :541-546 shift + cull).
// Returns false for a row the shift stage drops.
__device__ __forceinline__ bool apply_aug(Event& e, const memb_event_aug& a) {
  e.x = __dmul_rn(e.x, a.scale_x);
  e.y = __dmul_rn(e.y, a.scale_y);
  if (a.time_flip) e.p = -e.p;
  if (a.flip_x) e.x = __dsub_rn((double)(a.flip_w - 1), e.x);
  if (a.cull) {
    e.x = __dadd_rn(e.x, (double)a.shift_x);
    e.y = __dadd_rn(e.y, (double)a.shift_y);
    return e.x >= 0.0 && e.x < (double)a.cull_w && e.y >= 0.0 && e.y < (double)a.cull_h;
  }
  return true;
}

// Row range of stream b after the SliceRandomMaxEvs window (datasets.py:494-497).
__device__ __forceinline__ void aug_window(const memb_event_aug& a, long long& begin, long long& end) {
  begin = min(end, begin + max((long long)a.start, 0LL));
  if (a.count >= 0) end = min(end, begin + (long long)a.count);
}

// ---------------------------------------------------------------- init
__global__ void __launch_bounds__(256) hist_init(uint4* __restrict__ ws, long long n_vec,
                                                 long long tkeys_vec, int B) {
  // One 16-byte vector per stream holds {min key, max key}; everything else starts at zero.
  pdl_launch_dependents();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n_vec; i += stride) {
    const bool key = tkeys_vec >= 0 && i >= tkeys_vec && i < tkeys_vec + B;
    ws[i] = key ? make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// ---------------------------------------------------------------- per-stream min/max of t
template <bool kAligned>
__global__ void __launch_bounds__(kThreads) hist_time_range(const double* __restrict__ ev,
                                                            const long long* __restrict__ offsets,
                                                            long long n_total,
                                                            unsigned long long* __restrict__ tkeys,
                                                            const memb_event_aug* __restrict__ aug) {
  // min / max of t over the rows the rasteriser will see: the stream's window, minus the rows the shift stage culls
  // (EventArrToImg normalises the time surface over the array it is handed, datasets.py:587-589)
  const int b = blockIdx.y;
  long long begin = offsets ? offsets[b] : 0, end = offsets ? offsets[b + 1] : n_total;
  memb_event_aug a;
  const bool has_aug = aug != nullptr;
  if (has_aug) {
    a = aug[b];
    aug_window(a, begin, end);
  }
  unsigned long long lo = ~0ull, hi = 0ull;
  for (long long r = begin + blockIdx.x * (long long)kThreads + threadIdx.x; r < end;
       r += (long long)gridDim.x * kThreads) {
    Event e = load_event<kAligned>(ev, r);
    const double t = e.t;
    if (has_aug && !apply_aug(e, a)) continue;
    unsigned long long k = order_key(t);
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
  }
  for (int o = 16; o; o >>= 1) {
    unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(&tkeys[2 * b], lo);
    atomicMax(&tkeys[2 * b + 1], hi);
  }
}

// ---------------------------------------------------------------- strategy GLOBAL
// acc: u32 [B][2][npix] (pos plane, neg plane); last: u64 [B][npix] (row index + 1 of the last
// writer, time surface only).
template <bool kAligned, bool kAggregate, bool kTss>
__global__ void __launch_bounds__(kThreads) hist_scatter_global(
    const double* __restrict__ ev, const long long* __restrict__ offsets, long long n_total, int W,
    long long npix, unsigned int* __restrict__ acc, unsigned long long* __restrict__ last,
    Header* __restrict__ hdr, const memb_event_aug* __restrict__ aug, int replicas, const int* __restrict__ skip_if) {
  const int b = blockIdx.y;
  if (skip_if) {                // HYBRID chain: the prepare kernel decides on the device which rasteriser runs
    pdl_launch_dependents();
    pdl_wait();
    if (*skip_if) return;
  }
  long long begin = offsets ? offsets[b] : 0, end = offsets ? offsets[b + 1] : n_total;
  memb_event_aug a;
  const bool has_aug = aug != nullptr;
  if (has_aug) {
    a = aug[b];
    aug_window(a, begin, end);
  }
  // GLOBAL_REPL: every warp adds into one of `replicas` copies of the planes (neighbouring warps use different
  // copies), so that REDs to one hot sector are spread over `replicas` sectors; hist_finalize adds the copies.
  const int rep = replicas > 1 ? (int)((blockIdx.x * (unsigned)(kThreads / 32) + (threadIdx.x >> 5)) % (unsigned)replicas) : 0;
  unsigned int* acc_b = acc + ((long long)b * replicas + rep) * 2 * npix;
  unsigned long long* last_b = kTss ? last + (long long)b * npix : nullptr;
  bool bad = false;
  bool first = skip_if == nullptr;
  pdl_launch_dependents();      // the finalize grid may take the SM slots this grid's tail leaves free

  const long long step = (long long)gridDim.x * kThreads * kUnroll;
  // Whole-warp iterations so that match.any sees a converged warp.
  for (long long base = begin + blockIdx.x * (long long)(kThreads * kUnroll); base < end; base += step) {
    Event e[kUnroll];
    bool live[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      long long r = base + u * kThreads + threadIdx.x;
      live[u] = r < end;
      if (live[u]) e[u] = load_event<kAligned>(ev, r);
    }
    if (first) {                // the first rows were requested while the zero-fill kernel was still draining
      pdl_wait();
      first = false;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (has_aug && live[u]) live[u] = apply_aug(e[u], a);
      long long idx = 0;
      const bool pos = live[u] && e[u].p == 1.0, neg = live[u] && e[u].p == -1.0;
      const bool need = kTss ? live[u] : (pos || neg);
      const bool ok = need && pixel_index(e[u].x, e[u].y, W, npix, idx);
      bad |= need && !ok;
      const bool count = ok && (pos || neg);
      const long long slot = (neg ? npix : 0) + idx;
      if constexpr (kAggregate) {
        const unsigned int active = __ballot_sync(0xffffffffu, count);
        if (count) {
          const unsigned int peers = __match_any_sync(active, slot);
          if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(acc_b + slot, (unsigned int)__popc(peers));
        }
      } else {
        if (count) atomicAdd(acc_b + slot, 1u);
      }
      if constexpr (kTss) {
        // the LAST row (in array order) that hits a pixel writes its time surface value (numpy fancy assignment); after
        // RandomTimeFlip the array runs backwards, so the winner is the row with the smallest index: key = end - row
        const long long r = base + u * kThreads + threadIdx.x;
        if (ok) atomicMax(last_b + idx, (unsigned long long)((has_aug && a.time_flip) ? end - r : r - begin + 1));
      }
    }
  }
  if (first) pdl_wait();         // no rows for this CTA: still order the flag write after the zero-fill
  if (bad) hdr->oob = 1;
}

// 4 pixels per thread; out is uint8 [B][npix][C].
template <bool kTss>
__global__ void __launch_bounds__(256) hist_finalize(const unsigned int* __restrict__ acc,
                                                     const unsigned long long* __restrict__ last,
                                                     const unsigned long long* __restrict__ tkeys,
                                                     const double* __restrict__ ev,
                                                     const long long* __restrict__ offsets,
                                                     long long npix, int C, uint8_t* __restrict__ out, int replicas,
                                                     const memb_event_aug* __restrict__ aug, long long n_total) {
  const int b = blockIdx.y;
  const unsigned int* pos = acc + (long long)b * replicas * 2 * npix;
  const unsigned int* neg = pos + npix;
  uint8_t* o = out + (long long)b * npix * C;
  pdl_wait();
  double tmin = 0.0, span = 0.0, t_last = 0.0;
  long long begin = 0, end = 0;
  bool flip = false;
  if constexpr (kTss) {
    begin = offsets ? offsets[b] : 0;
    end = offsets ? offsets[b + 1] : n_total;
    if (aug) {
      const memb_event_aug a = aug[b];
      aug_window(a, begin, end);
      flip = a.time_flip != 0;
    }
    const double t_lo = key_value(tkeys[2 * b]), t_hi = key_value(tkeys[2 * b + 1]);
    if (flip && end > begin) {
      // RandomTimeFlip (datasets.py:603-606): t' = t[last row of the window] - t, one float64 subtraction per row
      t_last = ev[4 * (end - 1) + 2];
      tmin = t_last - t_hi;
      span = (t_last - t_lo) - tmin;
    } else {
      tmin = t_lo;
      span = t_hi - tmin;
    }
  }
  for (long long px = blockIdx.x * (long long)blockDim.x + threadIdx.x; px < npix;
       px += (long long)gridDim.x * blockDim.x) {
    unsigned int sp = pos[px], sn = neg[px];
    for (int k = 1; k < replicas; ++k) {
      sp += pos[(long long)k * 2 * npix + px];
      sn += neg[(long long)k * 2 * npix + px];
    }
    const uint8_t cp = (uint8_t)(sp & 0xffu), cn = (uint8_t)(sn & 0xffu);
    if (C == 2) {
      o[2 * px] = cp;
      o[2 * px + 1] = cn;
    } else {
      uint8_t ts = 0;
      if constexpr (kTss) {
        const unsigned long long w = last[(long long)b * npix + px];
        if (w) {
          // (t - tmin) / (tmax - tmin) * 255, float64, same operation order as datasets.py:588-589.
          const long long r = flip ? end - (long long)w : begin + (long long)w - 1;
          const double t = flip ? t_last - ev[4 * r + 2] : ev[4 * r + 2];
          const double v = (t - tmin) / span * 255.0;
          ts = (v == v) ? (uint8_t)(long long)v : (uint8_t)0;
        }
      }
      o[3 * px] = cp;
      o[3 * px + 1] = ts;
      o[3 * px + 2] = cn;
    }
  }
}

// ---------------------------------------------------------------- strategy TILE
// CTA (tile, b) owns pixels [tile*tile_pix, ...) of stream b, keeps one packed word per pixel in
// shared memory (low half: +1 events, high half: -1 events), reads the whole stream (L2 hits after
// the first tile) and writes its part of the uint8 image directly: no global atomics, no zero-fill,
// no finalize pass.  Halves are folded mod 256 between chunks so they can never carry.
template <bool kAligned>
__global__ void __launch_bounds__(kTileThreads, 1) hist_tile_smem(
    const double* __restrict__ ev, const long long* __restrict__ offsets, long long n_total, int W,
    long long npix, int tile_pix, int C, uint8_t* __restrict__ out, Header* __restrict__ hdr,
    const memb_event_aug* __restrict__ aug) {
  extern __shared__ unsigned int tile[];
  const int b = blockIdx.y;
  long long begin = offsets ? offsets[b] : 0, end = offsets ? offsets[b + 1] : n_total;
  memb_event_aug a;
  const bool has_aug = aug != nullptr;
  if (has_aug) {
    a = aug[b];
    aug_window(a, begin, end);
  }
  const long long lo = (long long)blockIdx.x * tile_pix;
  const int mine = (int)min((long long)tile_pix, npix - lo);
  for (int i = threadIdx.x; i < tile_pix; i += kTileThreads) tile[i] = 0u;
  __syncthreads();

  bool bad = false;
  for (long long chunk = begin; chunk < end; chunk += kTileChunk) {
    const long long stop = min(end, chunk + (long long)kTileChunk);
    for (long long base = chunk; base < stop; base += kTileThreads * kTileUnroll) {
      Event e[kTileUnroll];
      bool live[kTileUnroll];
#pragma unroll
      for (int u = 0; u < kTileUnroll; ++u) {
        long long r = base + u * kTileThreads + threadIdx.x;
        live[u] = r < stop;
        if (live[u]) e[u] = load_event<kAligned>(ev, r);
      }
#pragma unroll
      for (int u = 0; u < kTileUnroll; ++u) {
        if (has_aug && live[u]) live[u] = apply_aug(e[u], a);
        const bool pos = live[u] && e[u].p == 1.0, neg = live[u] && e[u].p == -1.0;
        if (pos || neg) {
          long long idx;
          if (!pixel_index(e[u].x, e[u].y, W, npix, idx)) {
            bad = true;
          } else {
            idx -= lo;
            if (idx >= 0 && idx < mine) atomicAdd(&tile[idx], pos ? 1u : 0x10000u);
          }
        }
      }
    }
    if (stop < end) {
      __syncthreads();
      for (int i = threadIdx.x; i < tile_pix; i += kTileThreads) tile[i] &= 0x00ff00ffu;
      __syncthreads();
    }
  }
  if (bad) hdr->oob = 1;
  __syncthreads();

  uint8_t* o = out + ((long long)b * npix + lo) * C;
  if (C == 2 && (((uintptr_t)o) & 7u) == 0) {
    // 4 pixels -> 8 bytes per thread per step
    for (int i = threadIdx.x * 4; i < mine; i += kTileThreads * 4) {
      if (i + 3 < mine) {
        unsigned int w0 = tile[i], w1 = tile[i + 1], w2 = tile[i + 2], w3 = tile[i + 3];
        uint2 v;
        v.x = (w0 & 0xffu) | ((w0 >> 16 & 0xffu) << 8) | ((w1 & 0xffu) << 16) | ((w1 >> 16 & 0xffu) << 24);
        v.y = (w2 & 0xffu) | ((w2 >> 16 & 0xffu) << 8) | ((w3 & 0xffu) << 16) | ((w3 >> 16 & 0xffu) << 24);
        *reinterpret_cast<uint2*>(o + 2 * i) = v;
      } else {
        for (int j = i; j < mine; ++j) {
          o[2 * j] = (uint8_t)(tile[j] & 0xffu);
          o[2 * j + 1] = (uint8_t)(tile[j] >> 16 & 0xffu);
        }
      }
    }
  } else if (C == 3 && (((uintptr_t)o) & 3u) == 0) {
    // 4 pixels -> 12 bytes per thread per step
    for (int i = threadIdx.x * 4; i < mine; i += kTileThreads * 4) {
      if (i + 3 < mine) {
        unsigned int w0 = tile[i], w1 = tile[i + 1], w2 = tile[i + 2], w3 = tile[i + 3];
        unsigned int* o32 = reinterpret_cast<unsigned int*>(o + 3 * i);
        o32[0] = (w0 & 0xffu) | ((w0 >> 16 & 0xffu) << 16) | ((w1 & 0xffu) << 24);
        o32[1] = (w1 >> 16 & 0xffu) << 8 | ((w2 & 0xffu) << 16);
        o32[2] = (w2 >> 16 & 0xffu) | ((w3 & 0xffu) << 8) | ((w3 >> 16 & 0xffu) << 24);
      } else {
        for (int j = i; j < mine; ++j) {
          o[3 * j] = (uint8_t)(tile[j] & 0xffu);
          o[3 * j + 1] = 0;
          o[3 * j + 2] = (uint8_t)(tile[j] >> 16 & 0xffu);
        }
      }
    }
  } else {
    for (int j = threadIdx.x; j < mine; j += kTileThreads) {
      if (C == 2) {
        o[2 * j] = (uint8_t)(tile[j] & 0xffu);
        o[2 * j + 1] = (uint8_t)(tile[j] >> 16 & 0xffu);
      } else {
        o[3 * j] = (uint8_t)(tile[j] & 0xffu);
        o[3 * j + 1] = 0;
        o[3 * j + 2] = (uint8_t)(tile[j] >> 16 & 0xffu);
      }
    }
  }
}

// ---------------------------------------------------------------- fused event pipeline
// One CTA per stream does the whole per-sample chain of build_transformNPY(train) (datasets.py:611-660):
// augment -> rasterise -> crop -> ToTensor -> RemoveTimesurface -> RemoveHotPixels -> NormalizeEvent.
// The cropped outH x outW raster (<= kTileMaxWords pixels, e.g. 224 x 224) lives in shared memory as one packed
// word per pixel, so the uint8 image never exists in HBM: the stream window is read once (32 B / event) and the
// float32 planes are written once.  Crop is applied to the integer pixel index (after numpy's truncation and
// negative wrap), which is what cropping the rasterised image does.
constexpr int kFuseUnroll = 2;                                   // rows per thread per iteration (double-buffered)
constexpr int kFuseAhead = 4;                                    // iterations (64 KB each) the L2 prefetch runs ahead

// Ask the copy engine to pull [p, p + bytes) into L2: no registers, no shared memory, nothing to wait on.
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned int bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
constexpr int kFuseFold = kTileChunk / (kTileThreads * kFuseUnroll);   // iterations between mod-256 folds

template <bool kAligned>
__global__ void __launch_bounds__(kTileThreads, 1) event_pipeline_fused(
    const double* __restrict__ ev, const long long* __restrict__ offsets, long long n_total,
    const memb_event_aug* __restrict__ aug, const int* __restrict__ crop_tl, int H, int W, int pad_t, int pad_l,
    int outH, int outW, int C, float hot_num_stds, int normalize, const float* __restrict__ value_lut,
    float* __restrict__ out, Header* __restrict__ hdr) {
  extern __shared__ unsigned int tile[];
  __shared__ unsigned long long red[2][kTileThreads / 32];   // per-warp partial sums (no 64-bit smem atomics)
  __shared__ unsigned int present[8];
  __shared__ float lut[256];      // ToTensor value of count c (what the hot-pixel statistics see)
  __shared__ float vlut[256];     // value written for count c: lut, or the caller's table (LogTransform / GammaTransform of lut)
  __shared__ int params[3];
  const int b = blockIdx.x;
  long long begin = offsets ? offsets[b] : 0, end = offsets ? offsets[b + 1] : n_total;
  const memb_event_aug a = aug[b];
  aug_window(a, begin, end);
  const long long npix = (long long)H * W;
  const int npx = outH * outW, npx4 = (npx + 3) & ~3;
  const int y0 = (crop_tl ? crop_tl[2 * b] : 0) - pad_t, x0 = (crop_tl ? crop_tl[2 * b + 1] : 0) - pad_l;
  // After the shift stage's cull every surviving row lies inside [0,cull_w) x [0,cull_h); when that window is
  // inside the raster the truncated coordinates ARE the pixel (no wrap, no range test, 32-bit conversions).
  const bool fast = a.cull && a.cull_w <= W && a.cull_h <= H;

  // The register double buffer below keeps only ~64 KB per SM in flight, too little to cover HBM latency at
  // full bandwidth; a bulk L2 prefetch running kFuseAhead iterations ahead turns the loads into L2 hits.
  constexpr int kStep = kTileThreads * kFuseUnroll;
  if (kAligned && threadIdx.x == 0 && begin < end)
    prefetch_l2(ev + 4 * begin, (unsigned int)(min((long long)kFuseAhead * kStep, end - begin) * 32));
  // first rows are requested before the tile is cleared: the clear overlaps their latency
  Event nxt[kFuseUnroll];
#pragma unroll
  for (int u = 0; u < kFuseUnroll; ++u) {
    const long long r = begin + u * kTileThreads + threadIdx.x;
    if (r < end) nxt[u] = load_event<kAligned>(ev, r);
  }
  if (threadIdx.x < 8) present[threadIdx.x] = 0u;
  if (threadIdx.x < 256) {
    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);   // ToTensor value of count c
    vlut[threadIdx.x] = value_lut ? __ldg(value_lut + threadIdx.x) : lut[threadIdx.x];
  }
  for (int i = threadIdx.x * 4; i < npx4; i += kTileThreads * 4) *reinterpret_cast<uint4*>(tile + i) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  bool bad = false;
  int it = 0;
  for (long long base = begin; base < end; base += kStep, ++it) {
    Event cur[kFuseUnroll];
    bool live[kFuseUnroll];
#pragma unroll
    for (int u = 0; u < kFuseUnroll; ++u) {
      cur[u] = nxt[u];
      live[u] = base + u * kTileThreads + threadIdx.x < end;
    }
    if (kAligned && threadIdx.x == 0) {
      const long long far = base + (long long)kFuseAhead * kStep;
      if (far < end) prefetch_l2(ev + 4 * far, (unsigned int)(min((long long)kStep, end - far) * 32));
    }
#pragma unroll
    for (int u = 0; u < kFuseUnroll; ++u) {            // next iteration's rows into registers
      const long long r = base + kStep + u * kTileThreads + threadIdx.x;
      if (r < end) nxt[u] = load_event<kAligned>(ev, r);
    }
#pragma unroll
    for (int u = 0; u < kFuseUnroll; ++u) {
      if (live[u]) live[u] = apply_aug(cur[u], a);
      const bool pos = live[u] && cur[u].p == 1.0, neg = live[u] && cur[u].p == -1.0;
      if (pos || neg) {
        int x, y;
        bool ok = true;
        if (fast) {
          x = __double2int_rz(cur[u].x);
          y = __double2int_rz(cur[u].y);
        } else {
          int idx = 0;
          ok = pixel_index_fast(cur[u].x, cur[u].y, W, H, npix, idx);
          y = idx / W;
          x = idx - y * W;
        }
        if (!ok) {
          bad = true;
        } else {
          y -= y0;
          x -= x0;
          if ((unsigned)y < (unsigned)outH && (unsigned)x < (unsigned)outW)
            atomicAdd(&tile[y * outW + x], pos ? 1u : 0x10000u);
        }
      }
    }
    if ((it + 1) % kFuseFold == 0 && base + kStep < end) {     // halves can never carry: fold them mod 256
      __syncthreads();
      for (int i = threadIdx.x; i < npx4; i += kTileThreads) tile[i] &= 0x00ff00ffu;
      __syncthreads();
    }
  }
  if (bad) hdr->oob = 1;
  __syncthreads();

  // ---- pass 1: counts mod 256 (uint8 wrap) in place, exact integer sums for RemoveHotPixels, and a 256-bit
  //      presence map of max(pos, neg) per pixel (what NormalizeEvent's maximum is read from once the hot
  //      threshold is known -- no second scan).  Event images are sparse: all-zero quads cost one 16-byte read.
  const bool filter = hot_num_stds >= 0.0f;
  {
    unsigned long long s1 = 0, s2 = 0;
    unsigned int seen0 = 0u;                            // counts < 32 (nearly all of them); larger ones go to smem
    for (int i = threadIdx.x * 4; i < npx4; i += kTileThreads * 4) {
      uint4 q = *reinterpret_cast<uint4*>(tile + i);
      if ((q.x | q.y | q.z | q.w) == 0u) continue;
      q.x &= 0x00ff00ffu; q.y &= 0x00ff00ffu; q.z &= 0x00ff00ffu; q.w &= 0x00ff00ffu;
      *reinterpret_cast<uint4*>(tile + i) = q;
      const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned int cp = w[k] & 0xffu, cn = w[k] >> 16, m = max(cp, cn);
        s1 += cp + cn;
        s2 += cp * cp + cn * cn;
        if (m < 32u) seen0 |= 1u << m;
        else atomicOr(&present[m >> 5], 1u << (m & 31u));
      }
    }
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    seen0 = __reduce_or_sync(0xffffffffu, seen0);
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = s1;
      red[1][threadIdx.x >> 5] = s2;
      if (seen0) atomicOr(&present[0], seen0);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // warp 0: threshold -> largest count that is NOT hot -> maximum surviving count -> scale factor
    int c_keep = 255;
    if (filter) {
      // exact sums, one rounding to float32 (see raster_post.cu for the rounding argument)
      unsigned long long t1 = red[0][threadIdx.x], t2 = red[1][threadIdx.x];
      for (int o = 16; o; o >>= 1) {
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
      }
      const double s1 = (double)t1, s2 = (double)t2, n = 2.0 * (double)npx;
      const double mean = s1 / (255.0 * n);
      double var = (s2 - s1 * s1 / n) / ((n - 1.0) * 255.0 * 255.0);
      var = var > 0.0 ? var : 0.0;
      const float thr = (float)(mean + (double)hot_num_stds * sqrt(var));
      int cold = 0;                                     // lut is monotone: count the entries not above thr
#pragma unroll
      for (int k = 0; k < 8; ++k) cold += !(lut[threadIdx.x * 8 + k] > thr);
      cold = __reduce_add_sync(0xffffffffu, cold);
      c_keep = cold - 1;                                // counts 0 .. c_keep survive (-1: everything is hot)
    }
    if (threadIdx.x == 0) {
      int m = 0;
      for (int c = min(c_keep, 255); c > 0 && m == 0; --c)
        if (present[c >> 5] >> (c & 31) & 1u) m = c;
      params[0] = c_keep;
      params[1] = (normalize && m != 0) ? 1 : 0;
      // factor = 1.0 / x.max(), x.max() = fl32(cmax / 255) (transforms.py:234-236), or its transformed value (the
      // transforms are monotone, and a zero maximum leaves the image alone like `if x.max() != 0`)
      const bool on = normalize && m != 0 && vlut[m] != 0.0f;
      params[1] = on ? 1 : 0;
      params[2] = __float_as_int(on ? __fdiv_rn(1.0f, vlut[m]) : 1.0f);
    }
  }
  __syncthreads();
  const int c_keep = params[0];
  const bool scale = params[1] != 0;
  const float factor = __int_as_float(params[2]);
  // ---- pass 3: float32 planes, 4 pixels (16 B) per thread per plane, streaming stores
  float* o = out + (long long)b * C * npx;
  float* o_neg = o + (long long)(C - 1) * npx;
  const bool vec = (npx & 3) == 0;
  for (int i = threadIdx.x * 4; i < npx4; i += kTileThreads * 4) {
    const uint4 q = *reinterpret_cast<const uint4*>(tile + i);
    float vp[4] = {0.f, 0.f, 0.f, 0.f}, vn[4] = {0.f, 0.f, 0.f, 0.f};
    if ((q.x | q.y | q.z | q.w) != 0u) {
      const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cp = (int)(w[k] & 0xffu), cn = (int)(w[k] >> 16);
        if (w[k] != 0u && cp <= c_keep && cn <= c_keep) {
          vp[k] = vlut[cp];
          vn[k] = vlut[cn];
          if (scale) {
            vp[k] = __fmul_rn(vp[k], factor);
            vn[k] = __fmul_rn(vn[k], factor);
          }
        }
      }
    }
    if (vec) {
      __stcs(reinterpret_cast<float4*>(o + i), make_float4(vp[0], vp[1], vp[2], vp[3]));
      __stcs(reinterpret_cast<float4*>(o_neg + i), make_float4(vn[0], vn[1], vn[2], vn[3]));
      if (C == 3) __stcs(reinterpret_cast<float4*>(o + npx + i), make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
      for (int k = 0; k < 4 && i + k < npx; ++k) {
        o[i + k] = vp[k];
        o_neg[i + k] = vn[k];
        if (C == 3) o[npx + i + k] = 0.0f;
      }
    }
  }
}

// ---------------------------------------------------------------- fused event pipeline, VARIABLE sensor size
// The N-Caltech101 / N-Cars branch of build_transformNPY (datasets.py:611-660 with H = W = None): recordings differ in
// extent, every stage infers its size from the rows it sees, and the raster is resized to the model input:
//   SliceRandomMaxEvs (:488-498) -> RandomTimeFlip (:598-608) -> Aug_FlipEvsAlongX (W = int(max x) + 1, :513-519) ->
//   Aug_RandomShiftEvs (H, W = int(max) + 1 BEFORE the shift; shift; drop rows outside, :538-547) ->
//   EventArrToImg(None, None) (H, W = int(max) + 1 over the surviving rows, :571-575) -> ToTensor ->
//   Resize((outH, outW), BILINEAR, antialias=True) -> RemoveTimesurface -> RemoveHotPixels -> NormalizeEvent.
// One CTA per stream: pass 1 over the window finds min x / max x / max y, pass 2 rasterises the augmented rows into a
// shared-memory tile of the cull window (H2 x W2 <= canvas) and tracks the survivors' extent (H3 x W3, a top-left
// sub-rectangle of the tile), the resize then evaluates ATen's separable anti-aliased triangle filter
// (UpSampleKernel.cpp: support = max(scale, 1), taps normalised per output index, horizontal then vertical) straight
// from the tile, and the float32 statistics / filter / normalisation run over this stream's own output.
// status codes left in the header: 1 = index error (never here: rows are culled), 2 = a stage saw an empty stream
// (the reference raises ValueError from max() of an empty array), 3 = the recording does not fit the canvas.
constexpr int kVarMaxTaps = 16;      // ceil(max(scale, 1)) * 2 + 1 taps per axis: input extent <= 7 x output extent

struct AaTaps { int lo, n; float w[kVarMaxTaps]; };
__device__ __forceinline__ AaTaps aa_taps(int o, int in_size, int out_size) {
  // at::native upsample_bilinear2d_aa index / weight computation, align_corners = false
  AaTaps t;
  const float scale = (float)in_size / (float)out_size;
  const float support = scale >= 1.0f ? scale : 1.0f;
  const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  const float center = scale * ((float)o + 0.5f);
  t.lo = max((int)(center - support + 0.5f), 0);
  t.n = min(min((int)(center + support + 0.5f), in_size) - t.lo, kVarMaxTaps);
  float total = 0.f;
  for (int j = 0; j < t.n; ++j) {
    const float x = fabsf(((float)(j + t.lo) - center + 0.5f) * invscale);
    const float w = x < 1.0f ? 1.0f - x : 0.0f;
    t.w[j] = w;
    total += w;
  }
  if (total != 0.f)
    for (int j = 0; j < t.n; ++j) t.w[j] /= total;
  return t;
}

template <bool kAligned>
__global__ void __launch_bounds__(kTileThreads, 1) event_pipeline_var_fused(
    const double* __restrict__ ev, const long long* __restrict__ offsets, long long n_total,
    const memb_event_aug* __restrict__ aug, int Hc, int Wc, int outH, int outW, int C, float hot_num_stds, int normalize,
    int logtrafo, int gammatrafo, float gamma, int timesurface, float* __restrict__ out, Header* __restrict__ hdr) {
  extern __shared__ unsigned int tile[];
  __shared__ double redd[3][kTileThreads / 32];
  __shared__ unsigned long long redk[2][kTileThreads / 32];
  __shared__ float redf[kTileThreads / 32];
  __shared__ float lut[256];
  __shared__ int ext[2];
  __shared__ float s_thr, s_factor;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long begin = offsets ? offsets[b] : 0, end = offsets ? offsets[b + 1] : n_total;
  const memb_event_aug a = aug[b];
  aug_window(a, begin, end);
  if (begin >= end) { if (threadIdx.x == 0) hdr->oob = 2; return; }
  if (threadIdx.x < 256) lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
  if (threadIdx.x < 2) ext[threadIdx.x] = -1;

  // ---- pass 1: extent of the window (numpy max / min over every row, whatever its polarity)
  double mnx = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (long long r = begin + threadIdx.x; r < end; r += kTileThreads) {
    const Event e = load_event<kAligned>(ev, r);
    mnx = fmin(mnx, e.x); mxx = fmax(mxx, e.x); mxy = fmax(mxy, e.y);
  }
  for (int o = 16; o; o >>= 1) {
    mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if (lane == 0) { redd[0][warp] = mnx; redd[1][warp] = mxx; redd[2][warp] = mxy; }
  __syncthreads();
  for (int w = 0; w < kTileThreads / 32; ++w) { mnx = fmin(mnx, redd[0][w]); mxx = fmax(mxx, redd[1][w]); mxy = fmax(mxy, redd[2][w]); }
  // sizes as the reference derives them: W1 for the flip, (H2, W2) for the cull window (both from the data BEFORE the shift)
  const long long W1 = __double2ll_rz(mxx) + 1;
  const double mxx_flipped = a.flip_x ? __dsub_rn((double)(W1 - 1), mnx) : mxx;
  const long long W2 = __double2ll_rz(mxx_flipped) + 1, H2 = __double2ll_rz(mxy) + 1;
  if (!(W2 >= 1 && H2 >= 1 && W2 <= Wc && H2 <= Hc) || !(mnx >= 0.0)) {      // (negative coordinates: outside the format)
    if (threadIdx.x == 0) hdr->oob = 3;
    return;
  }
  const int tw = (int)W2, th = (int)H2, words = (tw * th + 3) & ~3;
  for (int i = threadIdx.x * 4; i < words; i += kTileThreads * 4) *reinterpret_cast<uint4*>(tile + i) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  // ---- pass 2: augment, cull, rasterise into the (H2 x W2) tile; extent of the surviving rows
  int sx = -1, sy = -1, it = 0;
  for (long long base = begin; base < end; base += kTileThreads, ++it) {
    const long long r = base + threadIdx.x;
    if (r < end) {
      Event e = load_event<kAligned>(ev, r);
      if (a.time_flip) e.p = -e.p;
      if (a.flip_x) e.x = __dsub_rn((double)(W1 - 1), e.x);
      bool keep = true;
      if (a.cull) {
        e.x = __dadd_rn(e.x, (double)a.shift_x);
        e.y = __dadd_rn(e.y, (double)a.shift_y);
        keep = e.x >= 0.0 && e.x < (double)W2 && e.y >= 0.0 && e.y < (double)H2;
      }
      if (keep) {
        const int x = __double2int_rz(e.x), y = __double2int_rz(e.y);
        sx = max(sx, x); sy = max(sy, y);
        if (e.p == 1.0) atomicAdd(&tile[y * tw + x], 1u);
        else if (e.p == -1.0) atomicAdd(&tile[y * tw + x], 0x10000u);
      }
    }
    if ((it + 1) % (kTileChunk / kTileThreads) == 0 && base + kTileThreads < end) {     // halves can never carry
      __syncthreads();
      for (int i = threadIdx.x; i < words; i += kTileThreads) tile[i] &= 0x00ff00ffu;
      __syncthreads();
    }
  }
  sx = __reduce_max_sync(0xffffffffu, sx);
  sy = __reduce_max_sync(0xffffffffu, sy);
  if (lane == 0) { atomicMax(&ext[0], sx); atomicMax(&ext[1], sy); }
  __syncthreads();
  const int W3 = ext[0] + 1, H3 = ext[1] + 1;
  if (W3 < 1 || H3 < 1) { if (threadIdx.x == 0) hdr->oob = 2; return; }     // every row was culled

  // ---- resize (H3 x W3) -> (outH x outW), anti-aliased bilinear, horizontal then vertical; raw planes + statistics
  const int npx = outH * outW;
  float* o_pos = out + (long long)b * C * npx;
  float* o_neg = o_pos + (long long)(C - 1) * npx;
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < npx; i += kTileThreads) {
    const int oy = i / outW, ox = i - oy * outW;
    const AaTaps tx = aa_taps(ox, W3, outW), ty = aa_taps(oy, H3, outH);
    float vp = 0.f, vn = 0.f;
    for (int jy = 0; jy < ty.n; ++jy) {
      const unsigned int* row = tile + (ty.lo + jy) * tw + tx.lo;
      float rp = 0.f, rn = 0.f;
      for (int jx = 0; jx < tx.n; ++jx) {
        const unsigned int w = row[jx];
        rp = fmaf(lut[w & 0xffu], tx.w[jx], rp);
        rn = fmaf(lut[(w >> 16) & 0xffu], tx.w[jx], rn);
      }
      vp = fmaf(rp, ty.w[jy], vp);
      vn = fmaf(rn, ty.w[jy], vn);
    }
    o_pos[i] = vp;
    o_neg[i] = vn;
    if (C == 3 && !timesurface) o_pos[npx + i] = 0.0f;       // RemoveTimesurface
    s1 += (double)vp + (double)vn;
    s2 += (double)vp * vp + (double)vn * vn;
  }
  if (C == 3 && timesurface) {
    // ---- time surface (EventArrToImg(timeSurface=True), datasets.py:585-589): the tile is reused for the row that writes
    // each pixel last -- every surviving row whatever its polarity; after RandomTimeFlip the array runs backwards, so the
    // winner is the smallest row index -- then holds that row's (t - t.min()) / (t - t.min()).max() * 255 as uint8 and is
    // resized like the polarity planes.  The filter / log / gamma / normalisation below leave this plane alone.
    __syncthreads();
    for (int i = threadIdx.x * 4; i < words; i += kTileThreads * 4) *reinterpret_cast<uint4*>(tile + i) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    unsigned long long lo = ~0ull, hi = 0ull;
    for (long long r = begin + threadIdx.x; r < end; r += kTileThreads) {
      Event e = load_event<kAligned>(ev, r);
      if (a.flip_x) e.x = __dsub_rn((double)(W1 - 1), e.x);
      if (a.cull) {
        e.x = __dadd_rn(e.x, (double)a.shift_x);
        e.y = __dadd_rn(e.y, (double)a.shift_y);
        if (!(e.x >= 0.0 && e.x < (double)W2 && e.y >= 0.0 && e.y < (double)H2)) continue;
      }
      const unsigned long long k = order_key(e.t);
      lo = k < lo ? k : lo;
      hi = k > hi ? k : hi;
      atomicMax(&tile[__double2int_rz(e.y) * tw + __double2int_rz(e.x)], (unsigned int)(a.time_flip ? end - r : r - begin + 1));
    }
    for (int o = 16; o; o >>= 1) {
      const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
      lo = l2 < lo ? l2 : lo;
      hi = h2 > hi ? h2 : hi;
    }
    if (lane == 0) { redk[0][warp] = lo; redk[1][warp] = hi; }
    __syncthreads();
    for (int w = 0; w < kTileThreads / 32; ++w) { lo = redk[0][w] < lo ? redk[0][w] : lo; hi = redk[1][w] > hi ? redk[1][w] : hi; }
    const double t_lo = key_value(lo), t_hi = key_value(hi);
    const bool flip = a.time_flip != 0;
    const double t_last = ev[4 * (end - 1) + 2];          // RandomTimeFlip: t' = t[last row of the window] - t (datasets.py:603-606)
    const double tmin = flip ? t_last - t_hi : t_lo;
    const double span = flip ? (t_last - t_lo) - tmin : t_hi - tmin;
    for (int i = threadIdx.x; i < tw * th; i += kTileThreads) {
      const unsigned int w = tile[i];
      if (w) {
        const long long r = flip ? end - (long long)w : begin + (long long)w - 1;
        const double t = flip ? t_last - ev[4 * r + 2] : ev[4 * r + 2];
        const double v = (t - tmin) / span * 255.0;     // float64, the reference's operation order; numpy's cast truncates
        tile[i] = (v == v) ? (unsigned int)(uint8_t)(long long)v : 0u;
      }
    }
    __syncthreads();
    float* o_tss = o_pos + npx;
    for (int i = threadIdx.x; i < npx; i += kTileThreads) {
      const int oy = i / outW, ox = i - oy * outW;
      const AaTaps tx = aa_taps(ox, W3, outW), ty = aa_taps(oy, H3, outH);
      float v = 0.f;
      for (int jy = 0; jy < ty.n; ++jy) {
        const unsigned int* row = tile + (ty.lo + jy) * tw + tx.lo;
        float rv = 0.f;
        for (int jx = 0; jx < tx.n; ++jx) rv = fmaf(lut[row[jx] & 0xffu], tx.w[jx], rv);
        v = fmaf(rv, ty.w[jy], v);
      }
      o_tss[i] = v;
    }
  }
  // ---- RemoveHotPixels threshold: mean + k * std (unbiased) over both polarity planes
  const bool filter = hot_num_stds >= 0.0f;
  for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if (lane == 0) { redd[0][warp] = s1; redd[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int w = 0; w < kTileThreads / 32; ++w) { t1 += redd[0][w]; t2 += redd[1][w]; }
    const double n = 2.0 * (double)npx, mean = t1 / n;
    double var = (t2 - t1 * t1 / n) / (n - 1.0);
    var = var > 0.0 ? var : 0.0;
    s_thr = filter ? (float)(mean + (double)hot_num_stds * sqrt(var)) : INFINITY;
  }
  __syncthreads();
  const float thr = s_thr;
  const bool transform = logtrafo || gammatrafo;
  if (!filter && !normalize && !transform) return;
  // LogTransform / GammaTransform act on the resized float32 planes here (transforms.py:200-222): log(x + 1) and x ** gamma,
  // evaluated in double and rounded once (torch's float32 kernels are within 1 ulp of that); gamma = 0.5 is a square root in
  // torch (pow_tensor_scalar's special case), correctly rounded on both sides.
  auto value_transform = [&](float v) {
    if (logtrafo) v = (float)log((double)__fadd_rn(v, 1.0f));
    if (gammatrafo) v = gamma == 0.5f ? __fsqrt_rn(v) : (float)pow((double)v, (double)gamma);
    return v;
  };
  // ---- maximum of what survives the filter (this CTA re-reads its own stores: plain loads after the barrier)
  float mx = 0.f;
  for (int i = threadIdx.x; i < npx; i += kTileThreads) {
    const float vp = __ldcg(o_pos + i), vn = __ldcg(o_neg + i);
    if (!(vp > thr || vn > thr)) mx = fmaxf(mx, fmaxf(vp, vn));
  }
  if (transform) mx = value_transform(mx);        // both maps are non-decreasing: the maximum commutes with them
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) redf[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int w = 0; w < kTileThreads / 32; ++w) m = fmaxf(m, redf[w]);
    s_factor = (normalize && m != 0.f) ? __fdiv_rn(1.0f, m) : 1.0f;      // factor = 1.0 / x.max()  (transforms.py:234-236)
  }
  __syncthreads();
  const float factor = s_factor;
  for (int i = threadIdx.x; i < npx; i += kTileThreads) {
    float vp = __ldcg(o_pos + i), vn = __ldcg(o_neg + i);
    if (vp > thr || vn > thr) { vp = 0.f; vn = 0.f; }
    if (transform) { vp = value_transform(vp); vn = value_transform(vn); }
    o_pos[i] = __fmul_rn(vp, factor);
    o_neg[i] = __fmul_rn(vn, factor);
  }
}

// ---------------------------------------------------------------- row sources for the PRIVATE strategy
// kSrc: 0 = float64 rows, 32-byte aligned; -1 = float64 rows, 8-byte aligned; MEMB_RAW_NCALTECH101 / MEMB_RAW_NCARS =
// raw records decoded on the fly (bit fields as in decode_record below, process_dataset.py:52-60, :88-99).
template <int kSrc>
struct RowSource {
  static constexpr int kBytes = kSrc == 1 ? 5 : (kSrc == 2 ? 8 : 32);
  static constexpr bool kPrefetch = kSrc != -1;
  __device__ __forceinline__ static Event fetch(const void* __restrict__ base, long long r, long long n) {
    if constexpr (kSrc == 0) {
      return load_event<true>(static_cast<const double*>(base), r);
    } else if constexpr (kSrc == -1) {
      return load_event<false>(static_cast<const double*>(base), r);
    } else if constexpr (kSrc == 2) {
      const uint2 w = __ldg(static_cast<const uint2*>(base) + r);       // records are 8-byte aligned
      Event e;
      e.x = (double)(w.y & 0x3fffu);
      e.y = (double)((w.y & 0x0fffc000u) >> 14);
      e.t = (double)w.x;
      e.p = (w.y & 0x10000000u) ? 1.0 : 0.0;
      return e;
    } else {
      // 5-byte record at byte 5r: the two aligned words around it, or single bytes at the very end of the buffer
      const long long at = 5 * r, wi = at >> 2, bytes = 5 * n;
      unsigned long long v;
      if ((wi + 2) * 4 <= bytes) {
        const unsigned int* w = static_cast<const unsigned int*>(base) + wi;
        v = (((unsigned long long)__ldg(w + 1) << 32) | __ldg(w)) >> ((at & 3) * 8);
      } else {
        const uint8_t* b = static_cast<const uint8_t*>(base) + at;
        v = (unsigned long long)b[0] | ((unsigned long long)b[1] << 8) | ((unsigned long long)b[2] << 16) |
            ((unsigned long long)b[3] << 24) | ((unsigned long long)b[4] << 32);
      }
      const unsigned int b2 = (unsigned int)(v >> 16) & 0xffu;
      Event e;
      e.x = (double)((unsigned int)v & 0xffu);
      e.y = (double)((unsigned int)(v >> 8) & 0xffu);
      e.p = (b2 & 0x80u) ? 1.0 : -1.0;
      e.t = (double)(((b2 & 0x7fu) << 16) | (((unsigned int)(v >> 24) & 0xffu) << 8) | ((unsigned int)(v >> 32) & 0xffu));
      return e;
    }
  }
  // copy-engine prefetch of rows [r, r + rows) into L2 (16-byte aligned start for every chunk this kernel uses)
  __device__ __forceinline__ static void prefetch(const void* __restrict__ base, long long r, long long rows) {
    const unsigned int bytes = (unsigned int)(rows * kBytes) & ~15u;
    if (bytes) prefetch_l2(static_cast<const char*>(base) + r * kBytes, bytes);
  }
};

// ---------------------------------------------------------------- strategy PRIVATE (one long stream, small sensor)
// Every CTA (one per SM) keeps a private copy of the WHOLE sensor in shared memory (one packed word per pixel,
// <= kTileMaxWords pixels: N-Caltech101 240x180, N-Cars 120x100) and rasterises a strided share of the stream into
// it: hot pixels and edges cost shared-memory atomics instead of same-address L2 REDs, and nothing needs a zero-fill.
// Each CTA then stores its copy, already reduced mod 256, as one u16 per pixel (pos | neg << 8) into its own slice
// of the workspace with plain coalesced stores; hist_private_finalize adds the slices up.  (A first version summed
// the copies of a thread-block cluster through distributed shared memory: the DSMEM reads (~20 B/clk/SM) and the
// gpu-scope fences of cluster.sync cost more than writing 86 KB per CTA, and clusters left 28 SMs idle.)
constexpr int kPrivMaxCtas = 160;      // upper bound of the grid the workspace is sized for (>= SM count)

template <int kSrc>
__global__ void __launch_bounds__(kTileThreads, 1) hist_private(
    const void* __restrict__ ev, long long n, int W, int H, long long npix, unsigned short* __restrict__ slices,
    int* __restrict__ flags) {
  extern __shared__ unsigned int tile[];
  const int words = ((int)npix + 7) & ~7;
  constexpr int kStep = kTileThreads * kFuseUnroll;
  const long long stride = (long long)gridDim.x * kStep;
  const long long first = (long long)blockIdx.x * kStep;

  if (RowSource<kSrc>::kPrefetch && threadIdx.x == 0) {          // the copy engine pulls this CTA's next chunks into L2
    for (int k = 0; k < kFuseAhead; ++k) {
      const long long far = first + k * stride;
      if (far < n) RowSource<kSrc>::prefetch(ev, far, min((long long)kStep, n - far));
    }
  }
  Event nxt[kFuseUnroll];
#pragma unroll
  for (int u = 0; u < kFuseUnroll; ++u) {
    const long long r = first + u * kTileThreads + threadIdx.x;
    if (r < n) nxt[u] = RowSource<kSrc>::fetch(ev, r, n);
  }
  for (int i = threadIdx.x * 4; i < words; i += kTileThreads * 4) *reinterpret_cast<uint4*>(tile + i) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  bool bad = false;
  int it = 0;
  for (long long base = first; base < n; base += stride, ++it) {
    Event cur[kFuseUnroll];
    bool live[kFuseUnroll];
#pragma unroll
    for (int u = 0; u < kFuseUnroll; ++u) {
      cur[u] = nxt[u];
      live[u] = base + u * kTileThreads + threadIdx.x < n;
    }
    if (RowSource<kSrc>::kPrefetch && threadIdx.x == 0) {
      const long long far = base + (long long)kFuseAhead * stride;
      if (far < n) RowSource<kSrc>::prefetch(ev, far, min((long long)kStep, n - far));
    }
#pragma unroll
    for (int u = 0; u < kFuseUnroll; ++u) {
      const long long r = base + stride + u * kTileThreads + threadIdx.x;
      if (r < n) nxt[u] = RowSource<kSrc>::fetch(ev, r, n);
    }
#pragma unroll
    for (int u = 0; u < kFuseUnroll; ++u) {
      const bool pos = live[u] && cur[u].p == 1.0, neg = live[u] && cur[u].p == -1.0;
      if (pos || neg) {
        int idx;
        if (!pixel_index_fast(cur[u].x, cur[u].y, W, H, npix, idx)) bad = true;
        else atomicAdd(&tile[idx], pos ? 1u : 0x10000u);
      }
    }
    if ((it + 1) % kFuseFold == 0 && base + stride < n) {     // halves can never carry: fold them mod 256
      __syncthreads();
      for (int i = threadIdx.x; i < words; i += kTileThreads) tile[i] &= 0x00ff00ffu;
      __syncthreads();
    }
  }
  const int any_bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) flags[blockIdx.x] = any_bad;           // plain store: nothing to initialise beforehand
  // this CTA's slice: 8 pixels -> 8 x u16 (pos | neg << 8) -> one 16-byte store
  uint4* slice = reinterpret_cast<uint4*>(slices + (long long)blockIdx.x * words);
  for (int i = threadIdx.x * 8; i < words; i += kTileThreads * 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(tile + i), b = *reinterpret_cast<const uint4*>(tile + i + 4);
    auto pack = [](unsigned int lo, unsigned int hi) {
      return (lo & 0xffu) | ((lo >> 8) & 0xff00u) | ((hi & 0xffu) << 16) | ((hi << 8) & 0xff000000u);
    };
    slice[i >> 3] = make_uint4(pack(a.x, a.y), pack(a.z, a.w), pack(b.x, b.y), pack(b.z, b.w));
  }
}

// out[px] = (sum over the CTA slices) mod 256; also publishes the out-of-range flag.  blockDim = (32, kFinRows): lane x
// owns 8 consecutive pixels, row y sums slices y, y + kFinRows, ... (all its loads in flight at once); the partial sums
// meet in shared memory.
constexpr int kFinRows = 32;
__global__ void __launch_bounds__(32 * kFinRows) hist_private_finalize(const unsigned short* __restrict__ slices, int n_slices, int words,
                                                             const int* __restrict__ flags, long long npix, int C,
                                                             uint8_t* __restrict__ out, Header* __restrict__ hdr) {
  __shared__ unsigned int part[kFinRows][32][8];      // [slice group][lane][pixel]: pos | neg << 16 (sums < 65536)
  if (blockIdx.x == 0 && threadIdx.y == 0) {
    int f = 0;
    for (int i = threadIdx.x; i < n_slices; i += 32) f |= flags[i];
    f = (int)__reduce_or_sync(0xffffffffu, (unsigned)f);
    if (threadIdx.x == 0) hdr->oob = f ? 1 : 0;
  }
  const long long oct = (long long)blockIdx.x * 32 + threadIdx.x;       // pixel octet
  unsigned int acc[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  if (oct * 8 < words) {
#pragma unroll 5
    for (int k = threadIdx.y; k < n_slices; k += kFinRows) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(slices + (long long)k * words) + oct);
      const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[2 * q] += (w[q] & 0xffu) | ((w[q] & 0xff00u) << 8);
        acc[2 * q + 1] += ((w[q] >> 16) & 0xffu) | ((w[q] >> 8) & 0xff0000u);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) part[threadIdx.y][threadIdx.x][q] = acc[q];
  __syncthreads();
  // the first 256 threads -> the block's 256 pixels
  const int lane = threadIdx.y * 4 + (threadIdx.x >> 3), q = threadIdx.x & 7;   // (octet in block, pixel in octet)
  const long long px = ((long long)blockIdx.x * 32 + lane) * 8 + q;
  if (threadIdx.y < 8 && px < npix) {
    unsigned int w = 0;
#pragma unroll
    for (int g = 0; g < kFinRows; ++g) w += part[g][lane][q];
    const uint8_t cp = (uint8_t)(w & 0xffu), cn = (uint8_t)((w >> 16) & 0xffu);
    if (C == 2) {
      out[2 * px] = cp;
      out[2 * px + 1] = cn;
    } else {
      out[3 * px] = cp;
      out[3 * px + 1] = 0;
      out[3 * px + 2] = cn;
    }
  }
}

// ---------------------------------------------------------------- strategy HYBRID (one long stream, sensor too large for PRIVATE)
// A 640x480 sensor (1.2 MB of packed counters) does not fit one SM's shared memory, and the GLOBAL strategy collapses on
// concentrated streams (edges, hot regions): REDs to the same few L2 sectors serialise (10 M edge-like events: 270-430 us
// against 85 us uniform).  HYBRID privatises only the HOT part of the sensor, chosen per call ON THE DEVICE:
//   1. hist_hybrid_prepare zero-fills the L2 accumulators; its CTA 0 counts a sample of kHybSamples rows spread over the
//      stream per granule (an 8 x 8 pixel block: a line at any angle crosses few of them, a 64-pixel row segment would
//      catch ~5 hot pixels of a steep edge), picks the most frequent granules (as many as fit a shared-memory tile), writes the
//      granule -> slot map and decides the mode: privatise (1) when the sample is concentrated -- its effective number
//      of granules 1 / sum p_g^2 is below half of the sensor's -- or the caller forced HYBRID, else plain L2 REDs
//      (0: a uniform stream gains nothing from privatising a sixth of the sensor);
//   2. hist_scatter_global runs when mode == 0 and hist_hybrid when mode == 1 (both are launched; the other one's CTAs
//      return at once): every CTA of hist_hybrid (one per SM) rasterises a strided share of the stream, sends events of
//      mapped granules to shared-memory atomics (its private copy of the hot pixels) and the rest to L2 REDs, and stores
//      its copy (mod 256, u16 per pixel) into its own slice of the workspace;
//   3. hist_hybrid_finalize adds accumulators + slices.
// The four launches are chained as programmatic dependent launches.
constexpr int kHybGranule = 64;                 // pixels per granule: a block of 8 rows x 8 columns
constexpr int kHybSamples = 4096;               // rows the selector looks at (4 per thread, one batch of loads)
constexpr int kHybMaxGranules = 16384;          // sensors up to 1 Mpixel (1280x720 = 14400 granules)
constexpr int kHybHistBins = 1024;
constexpr int kHybSmemBytes = 222 * 1024;       // dynamic shared memory of hist_hybrid: slot map (u16 per granule) + tile

struct HybState { int nsel; int mode; int pad[2]; };

__host__ __device__ inline int hyb_gw(int W) { return (W + 7) >> 3; }                       // granules per row of blocks
__host__ __device__ inline long long hyb_granules(int W, int H) { return (long long)hyb_gw(W) * ((H + 7) >> 3); }
__device__ __forceinline__ int hyb_granule_of(int xi, int yi, int gw) { return (yi >> 3) * gw + (xi >> 3); }
__device__ __forceinline__ int hyb_offset_of(int xi, int yi) { return ((yi & 7) << 3) | (xi & 7); }
// tile capacity in granules once the slot map has taken its share of the dynamic shared memory
__host__ __device__ inline int hyb_tile_granules(int granules) {
  const int map_bytes = ((granules * 2 + 15) / 16) * 16;
  const int words = (kHybSmemBytes - map_bytes) / 4;
  return words / kHybGranule < 0xffff ? words / kHybGranule : 0xfffe;
}

// Block-wide exclusive scan of one int per thread (kTileThreads threads); returns the exclusive prefix, *total = sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_sums[lane] = w;           // inclusive over warps
  }
  __syncthreads();
  *total = warp_sums[kTileThreads / 32 - 1];
  return inc - v + (warp ? warp_sums[warp - 1] : 0);
}

template <bool kAligned>
__global__ void __launch_bounds__(kTileThreads) hist_hybrid_prepare(const double* __restrict__ ev, long long n, int W, int H,
                                                                    long long npix, uint4* __restrict__ zero_from,
                                                                    long long zero_vecs, int granules, int tile_granules,
                                                                    int force_mode, unsigned short* __restrict__ slot_map,
                                                                    unsigned short* __restrict__ sel_list,
                                                                    HybState* __restrict__ state) {
  extern __shared__ unsigned int cnt[];           // sample count per granule
  __shared__ int hist[kHybHistBins];
  __shared__ int warp_sums[32];
  __shared__ int s_thr, s_cover[32], s_seen[32];
  pdl_launch_dependents();
  if (blockIdx.x != 0) {                          // zero-fill of header + accumulators
    for (long long i = (blockIdx.x - 1) * (long long)kTileThreads + threadIdx.x; i < zero_vecs; i += (long long)(gridDim.x - 1) * kTileThreads)
      zero_from[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  // ---- CTA 0: sample, select, decide
  constexpr int kPer = kHybSamples / kTileThreads, kBatch = 4;
  const long long step = max(1LL, n / kHybSamples);
  const int gw = hyb_gw(W);
  for (int j = 0; j < kPer; j += kBatch) {
    Event e[kBatch];
    bool live[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const long long r = ((long long)(j + u) * kTileThreads + threadIdx.x) * step;
      live[u] = r < n;
      if (live[u]) e[u] = load_event<kAligned>(ev, r);
    }
    if (j == 0) {                                 // the counters are cleared while the (cold) sample rows are on their way
      for (int g = threadIdx.x; g < granules; g += kTileThreads) cnt[g] = 0u;
      for (int i = threadIdx.x; i < kHybHistBins; i += kTileThreads) hist[i] = 0;
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      int idx, xi, yi;
      if (live[u] && (e[u].p == 1.0 || e[u].p == -1.0) && pixel_xy_fast(e[u].x, e[u].y, W, H, npix, idx, xi, yi)) {
        atomicAdd(&cnt[hyb_granule_of(xi, yi, gw)], 1u);
      }
    }
  }
  __syncthreads();
  if (force_mode < 0) {
    // Is the stream concentrated?  The effective number of granules N_eff = 1 / sum_g p_g^2, estimated from the sample
    // counts c_g (S = sum c_g, Poisson sampling: E[sum c_g^2] = S + S^2 sum p_g^2): a uniform stream gives N_eff ~ the
    // granule count, 8 line segments ~ 500, 50 segments ~ 1500.  The L2-RED rasteriser starts to serialise on hot sectors
    // once the events sit on fewer than about half of the sensor's granules; above that it is the faster one: mode 0.
    long long sum_c = 0, sum_c2 = 0;
    for (int g = threadIdx.x; g < granules; g += kTileThreads) {
      const long long c = cnt[g];
      sum_c += c;
      sum_c2 += c * c;
    }
    const int s1 = __reduce_add_sync(0xffffffffu, (int)sum_c), s2 = __reduce_add_sync(0xffffffffu, (int)min(sum_c2, 0x3ffffffLL));
    if ((threadIdx.x & 31) == 0) { s_seen[threadIdx.x >> 5] = s1; s_cover[threadIdx.x >> 5] = s2; }
    __syncthreads();
    long long S = 0, Q = 0;
    for (int w = 0; w < kTileThreads / 32; ++w) { S += s_seen[w]; Q += s_cover[w]; }
    __syncthreads();
    // N_eff >= granules / 2   <=>   S^2 >= (Q - S) * granules / 2
    if (S == 0 || 2 * S * S >= (Q - S) * (long long)granules) {
      if (threadIdx.x == 0) { state->nsel = 0; state->mode = 0; }
      return;
    }
  }
  {   // histogram of the counts; nearly all of them are 0..3: those are counted in registers, one atomic per warp and bin
    int small[4] = {0, 0, 0, 0};
    for (int g = threadIdx.x; g < granules; g += kTileThreads) {
      const unsigned int c = min(cnt[g], (unsigned int)(kHybHistBins - 1));
      if (c < 4u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) small[k] += c == (unsigned int)k;
      } else {
        atomicAdd(&hist[c], 1);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int v = __reduce_add_sync(0xffffffffu, small[k]);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&hist[k], v);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // smallest threshold thr >= 1 with #{granules: count >= thr} <= tile_granules
    const int lane = threadIdx.x;
    int part = 0;
    for (int k = 0; k < 32; ++k) part += hist[lane * 32 + k];
    int suf = part;                               // inclusive suffix sum over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += t;
    }
    int run = suf - part, best = kHybHistBins;
    for (int k = 31; k >= 0; --k) {
      run += hist[lane * 32 + k];
      if (run <= tile_granules) best = lane * 32 + k; else break;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) s_thr = max(best, 1);
  }
  __syncthreads();
  const int thr = s_thr;
  const int per_thread = (granules + kTileThreads - 1) / kTileThreads;
  const int g0 = threadIdx.x * per_thread, g1 = min(granules, g0 + per_thread);
  int mine = 0;
  for (int g = g0; g < g1; ++g) mine += (int)min(cnt[g], (unsigned int)(kHybHistBins - 1)) >= thr;
  int nsel = 0;
  int slot = block_exclusive_scan(mine, warp_sums, &nsel);
  for (int g = g0; g < g1; ++g) {
    const bool sel = (int)min(cnt[g], (unsigned int)(kHybHistBins - 1)) >= thr;
    if (sel) sel_list[slot] = (unsigned short)g;          // slot -> granule, for the finalize pass
    slot_map[g] = sel ? (unsigned short)slot++ : (unsigned short)0xffffu;
  }
  if (threadIdx.x == 0) { state->nsel = nsel; state->mode = 1; }      // forced, or the sample is concentrated (see above)
}

template <bool kAligned>
__global__ void __launch_bounds__(kTileThreads, 1) hist_hybrid(
    const double* __restrict__ ev, long long n, int W, int H, long long npix, unsigned int* __restrict__ acc,
    const unsigned short* __restrict__ slot_map_g, int granules, int tile_granules, const HybState* __restrict__ state,
    unsigned short* __restrict__ slices, Header* __restrict__ hdr) {
  extern __shared__ __align__(16) unsigned char hyb_smem[];
  const int map_bytes = ((granules * 2 + 15) / 16) * 16;
  unsigned short* slot_map = reinterpret_cast<unsigned short*>(hyb_smem);
  unsigned int* tile = reinterpret_cast<unsigned int*>(hyb_smem + map_bytes);
  constexpr int kStep = kTileThreads * kFuseUnroll;
  const long long stride = (long long)gridDim.x * kStep;
  const long long first = (long long)blockIdx.x * kStep;
  const int gw = hyb_gw(W);
  pdl_launch_dependents();
  pdl_wait();                                   // map, mode and zeroed accumulators are visible from here on
  if (state->mode == 0) return;                 // the stream is not concentrated: hist_scatter_global did the work
  const int nsel = state->nsel;

  if (kAligned && threadIdx.x == 0) {           // the copy engine pulls this CTA's next chunks into L2
    for (int k = 0; k < kFuseAhead; ++k) {
      const long long far = first + k * stride;
      if (far < n) prefetch_l2(ev + 4 * far, (unsigned int)(min((long long)kStep, n - far) * 32));
    }
  }
  Event nxt[kFuseUnroll];
#pragma unroll
  for (int u = 0; u < kFuseUnroll; ++u) {
    const long long r = first + u * kTileThreads + threadIdx.x;
    if (r < n) nxt[u] = load_event<kAligned>(ev, r);
  }
  for (int i = threadIdx.x * 8; i < granules; i += kTileThreads * 8)      // map_bytes is a multiple of 16
    *reinterpret_cast<uint4*>(slot_map + i) = __ldg(reinterpret_cast<const uint4*>(slot_map_g + i));
  const int words = nsel * kHybGranule;
  for (int i = threadIdx.x * 4; i < words; i += kTileThreads * 4) *reinterpret_cast<uint4*>(tile + i) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  bool bad = false;
  int it = 0;
  for (long long base = first; base < n; base += stride, ++it) {
    Event cur[kFuseUnroll];
    bool live[kFuseUnroll];
#pragma unroll
    for (int u = 0; u < kFuseUnroll; ++u) {
      cur[u] = nxt[u];
      live[u] = base + u * kTileThreads + threadIdx.x < n;
    }
    if (kAligned && threadIdx.x == 0) {
      const long long far = base + (long long)kFuseAhead * stride;
      if (far < n) prefetch_l2(ev + 4 * far, (unsigned int)(min((long long)kStep, n - far) * 32));
    }
#pragma unroll
    for (int u = 0; u < kFuseUnroll; ++u) {
      const long long r = base + stride + u * kTileThreads + threadIdx.x;
      if (r < n) nxt[u] = load_event<kAligned>(ev, r);
    }
#pragma unroll
    for (int u = 0; u < kFuseUnroll; ++u) {
      const bool pos = live[u] && cur[u].p == 1.0, neg = live[u] && cur[u].p == -1.0;
      if (pos || neg) {
        int idx, xi, yi;
        if (!pixel_xy_fast(cur[u].x, cur[u].y, W, H, npix, idx, xi, yi)) {
          bad = true;
        } else {
          const unsigned int s = slot_map[hyb_granule_of(xi, yi, gw)];
          if (s != 0xffffu) atomicAdd(&tile[s * kHybGranule + hyb_offset_of(xi, yi)], pos ? 1u : 0x10000u);
          else atomicAdd(acc + (neg ? npix : 0) + idx, 1u);
        }
      }
    }
    if ((it + 1) % kFuseFold == 0 && base + stride < n) {     // halves can never carry: fold them mod 256
      __syncthreads();
      for (int i = threadIdx.x; i < words; i += kTileThreads) tile[i] &= 0x00ff00ffu;
      __syncthreads();
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) hdr->oob = 1;
  // this CTA's slice: 8 pixels -> 8 x u16 (pos | neg << 8) -> one 16-byte store
  uint4* slice = reinterpret_cast<uint4*>(slices + (long long)blockIdx.x * tile_granules * kHybGranule);
  for (int i = threadIdx.x * 8; i < words; i += kTileThreads * 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(tile + i), b = *reinterpret_cast<const uint4*>(tile + i + 4);
    auto pack = [](unsigned int lo, unsigned int hi) {
      return (lo & 0xffu) | ((lo >> 8) & 0xff00u) | ((hi & 0xffu) << 16) | ((hi << 8) & 0xff000000u);
    };
    slice[i >> 3] = make_uint4(pack(a.x, a.y), pack(a.z, a.w), pack(b.x, b.y), pack(b.z, b.w));
  }
}

// out[px] = (L2 accumulators + sum over the CTA slices of mapped granules) mod 256.  blockDim = (32, kHybFinRows).
// Blocks [0, hot_blocks): four mapped granules each (slots 4j .. 4j+3 of the slot -> granule list): lane x owns 8 consecutive
// pixels, row y sums slices y, y + kHybFinRows, ...; the partial sums meet in shared memory (these blocks do the long
// work, so they are scheduled first).  Blocks beyond: 256 consecutive pixels each, accumulators -> image for the pixels
// of UNMAPPED granules (every pixel in mode 0).
constexpr int kHybFinRows = 16;
__device__ __forceinline__ void hyb_store_px(uint8_t* __restrict__ out, long long px, int C, unsigned int cp, unsigned int cn) {
  if (C == 2) {
    out[2 * px] = (uint8_t)cp;
    out[2 * px + 1] = (uint8_t)cn;
  } else {
    out[3 * px] = (uint8_t)cp;
    out[3 * px + 1] = 0;
    out[3 * px + 2] = (uint8_t)cn;
  }
}
__global__ void __launch_bounds__(32 * kHybFinRows) hist_hybrid_finalize(
    const unsigned int* __restrict__ acc, const unsigned short* __restrict__ slot_map, const unsigned short* __restrict__ sel_list,
    const HybState* __restrict__ state, const unsigned short* __restrict__ slices, int n_slices, int tile_granules, int hot_blocks,
    long long npix, int W, int H, int C, uint8_t* __restrict__ out) {
  __shared__ unsigned int part[kHybFinRows][32][8];      // pos | neg << 16 (sums < 65536: at most 160 slices x 255)
  pdl_wait();
  const int t = threadIdx.y * 32 + threadIdx.x;
  if ((int)blockIdx.x >= hot_blocks) {
    if (t < 256) {
      const long long px = (long long)((int)blockIdx.x - hot_blocks) * 256 + t;
      if (px < npix) {
        const int yi = (int)(px / W), xi = (int)(px - (long long)yi * W);
        if (state->mode == 0 || slot_map[hyb_granule_of(xi, yi, hyb_gw(W))] == 0xffffu)
          hyb_store_px(out, px, C, acc[px] & 0xffu, acc[npix + px] & 0xffu);
      }
    }
    return;
  }
  const int slot = (int)blockIdx.x * 4 + (threadIdx.x >> 3);      // lane x: octet (x & 7) of this slot's granule
  const int nsel = state->nsel;
  if ((int)blockIdx.x * 4 >= nsel) return;
  unsigned int sum[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  if (slot < nsel) {
    const long long off = ((long long)slot * kHybGranule) / 8 + (threadIdx.x & 7);   // uint4 index inside a slice
    const long long slice_vecs = (long long)tile_granules * kHybGranule / 8;
#pragma unroll 5
    for (int k = threadIdx.y; k < n_slices; k += kHybFinRows) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(slices) + k * slice_vecs + off);
      const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        sum[2 * q] += (w[q] & 0xffu) | ((w[q] & 0xff00u) << 8);
        sum[2 * q + 1] += ((w[q] >> 16) & 0xffu) | ((w[q] >> 8) & 0xff0000u);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) part[threadIdx.y][threadIdx.x][q] = sum[q];
  __syncthreads();
  if (t < 256) {                                            // 256 threads -> the 4 x 64 pixels of this block's granules
    const int lane = t >> 3, q = t & 7;                     // (slot in block * 8 + octet = row of the 8 x 8 block, column)
    const int sl = (int)blockIdx.x * 4 + (lane >> 3);
    if (sl < nsel) {
      const int g = sel_list[sl], gw = hyb_gw(W), gy = g / gw, gx = g - gy * gw;
      const int yi = gy * 8 + (lane & 7), xi = gx * 8 + q;
      if (xi < W && yi < H) {
        const long long px = (long long)yi * W + xi;
        unsigned int w = 0;
#pragma unroll
        for (int g = 0; g < kHybFinRows; ++g) w += part[g][lane][q];
        hyb_store_px(out, px, C, (w + acc[px]) & 0xffu, ((w >> 16) + acc[npix + px]) & 0xffu);
      }
    }
  }
}

// ---------------------------------------------------------------- strategy SORT (one long stream, large sensor)
// GLOBAL pays one L2 RED per event (reads alone 53 us, REDs alone 52 us, together 84 us per 10 M events at 640x480,
// 2-4x more on edge-like streams where REDs to one sector serialise); HYBRID moves the hot granules into shared memory
// but keeps the cold REDs.  SORT has no global atomics at all:
//   1. hist_sort_partition (two CTAs per SM, equal chunks of <= 4096 events): every event becomes a 16-bit key -- its
//      place inside one of kSortBins interleaved pixel classes (class idx % kSortBins, key = (idx / kSortBins) << 1 |
//      polarity) -- and the chunk's keys are counting-sorted by class in shared memory (one shared-memory atomic per
//      event returns its rank) and stored as one contiguous block, with the class boundaries of the chunk in a
//      transposed table off[class][chunk].  HBM traffic: 32 B read + 2 B written per event.
//   2. hist_sort_accumulate (one CTA per class, two resident per SM): gathers the class's segment of every chunk with
//      16-byte loads (the vectors of all segments are dealt round-robin to the threads, so long and short segments cost
//      the same), counts into its private shared-memory image of the class's pixels and writes their uint8 values
//      straight into the image: no zero-fill, no finalize pass.
// Classes interleave the sensor pixel by pixel, so edges and blobs spread evenly; a class far above its share (a hot
// pixel) switches its CTA to warp-aggregated atomics.  Out-of-range rows are reported through one flag per chunk, so
// nothing needs zeroing before the chain starts.
constexpr int kSortThreads = 512;
constexpr int kSortCap = 4096;                                       // key slots per chunk
constexpr int kSortBatch = 4;                                        // rows per thread in flight (reloaded as soon as consumed)
constexpr int kSortStep = kSortThreads * kSortBatch;                 // rows per batch (64 KB)
constexpr int kSortBatches = kSortCap / kSortStep;                   // batches per chunk
constexpr int kSortBins = 296;                                       // pixel classes
constexpr unsigned int kSortMagic = (unsigned int)((0x100000000ULL + kSortBins - 1) / kSortBins);   // idx / kSortBins, exact for idx < 2^23
constexpr long long kSortMaxPixels = 1LL << 22;                      // key = (idx / kSortBins) << 1 | polarity < 2^16
constexpr int kSortPerLane = (kSortBins + 31) / 32;                  // classes per lane in the one-warp scan
constexpr int kSortSlices = 4;                                       // accumulate CTAs per class (each takes a quarter of the chunks)
constexpr int kSortAccThreads = 512;
constexpr int kSortSmemBytes = kSortCap * 4 + kSortCap * 2 + kSortCap * 2;   // unsorted words, ranks, sorted keys

__host__ __device__ inline int sort_local_words(long long npix) {    // counters per class: pixels per class x 2, in whole uint4
  return (int)(((((npix - 1) / kSortBins + 1) * 2 + 3) / 4) * 4);
}

// float64 row through the read-only path, no L1 allocation, first in line for L2 eviction: the 320 MB stream must not
// push the sorted keys (re-read by the accumulate pass) out of the L2
template <bool kAligned>
__device__ __forceinline__ Event load_event_stream(const double* __restrict__ ev, long long row, unsigned long long policy) {
  Event e;
  const double* p = ev + 4 * row;
  if constexpr (kAligned) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(e.x), "=d"(e.y), "=d"(e.t), "=d"(e.p) : "l"(p), "l"(policy));
  } else {
    e.x = __ldg(p); e.y = __ldg(p + 1); e.t = __ldg(p + 2); e.p = __ldg(p + 3);
  }
  return e;
}

template <bool kAligned>
__global__ void __launch_bounds__(kSortThreads, 2) hist_sort_partition(
    const double* __restrict__ ev, long long n, int W, int H, long long npix, int nchunks, int chunk_len,
    unsigned short* __restrict__ keys, unsigned short* __restrict__ offs, unsigned char* __restrict__ bad_chunk,
    unsigned int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  unsigned int* unsorted = reinterpret_cast<unsigned int*>(sort_smem);                               // key | class << 16
  unsigned short* rank = reinterpret_cast<unsigned short*>(sort_smem + kSortCap * 4);
  unsigned short* sorted = reinterpret_cast<unsigned short*>(sort_smem + kSortCap * 6);
  __shared__ unsigned int cnt[kSortPerLane * 32];
  __shared__ unsigned int start[kSortPerLane * 32 + 1];
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  unsigned long long stream_policy, keep_policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream_policy));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
  for (int i = tid; i < kSortPerLane * 32; i += kSortThreads) cnt[i] = 0u;
  if (blockIdx.x == 0) {                          // hand-off flags of the accumulate pass (it starts after this grid has completed)
    for (int i = tid; i < kSortBins * (kSortSlices - 1); i += kSortThreads) flags[i] = 0u;
  }
  // this CTA's it-th batch: chunk blockIdx.x + (it / kSortBatches) * gridDim.x, rows [chunk * chunk_len + batch * kSortStep, ...)
  auto batch_rows = [&](int it, long long& first, long long& end) {
    const long long chunk = (long long)blockIdx.x + (long long)(it / kSortBatches) * gridDim.x;
    end = min(n, (chunk + 1) * (long long)chunk_len);
    first = chunk * (long long)chunk_len + (long long)(it % kSortBatches) * kSortStep;
  };
  Event nxt[kSortBatch];
  {
    long long f, e;
    batch_rows(0, f, e);
#pragma unroll
    for (int u = 0; u < kSortBatch; ++u) {
      const long long r = f + u * kSortThreads + tid;
      if (r < e) nxt[u] = load_event_stream<kAligned>(ev, r, stream_policy);
    }
  }
  __syncthreads();
  int it = 0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    bool bad = false;
#pragma unroll 1
    for (int j = 0; j < kSortBatches; ++j, ++it) {
      long long first, end, nfirst, nend;
      batch_rows(it, first, end);
      batch_rows(it + 1, nfirst, nend);
#pragma unroll
      for (int u = 0; u < kSortBatch; ++u) {
        const int slot = (j * kSortBatch + u) * kSortThreads + tid;
        const Event cur = nxt[u];
        {                                         // the register is free again: request the row of the next batch
          const long long r = nfirst + u * kSortThreads + tid;
          if (r < nend) nxt[u] = load_event_stream<kAligned>(ev, r, stream_policy);
        }
        unsigned int w = 0xffffffffu;
        if (first + u * kSortThreads + tid < end) {
          const bool pos = cur.p == 1.0, neg = cur.p == -1.0;
          if (pos || neg) {
            int idx;
            if (!pixel_index_fast(cur.x, cur.y, W, H, npix, idx)) {
              bad = true;
            } else {
              const unsigned int q = __umulhi((unsigned int)idx, kSortMagic);
              const unsigned int cls = (unsigned int)idx - q * kSortBins;
              rank[slot] = (unsigned short)atomicAdd(&cnt[cls], 1u);
              w = (q << 1) | (neg ? 1u : 0u) | (cls << 16);
            }
          }
        }
        unsorted[slot] = w;
      }
    }
    __syncthreads();
    if (tid < 32) {                               // exclusive scan of the class counts by one warp; counts reset for the next chunk
      unsigned int c[kSortPerLane], sum = 0;
#pragma unroll
      for (int k = 0; k < kSortPerLane; ++k) {
        c[k] = cnt[tid * kSortPerLane + k];
        cnt[tid * kSortPerLane + k] = 0u;
        sum += c[k];
      }
      unsigned int inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      unsigned int run = inc - sum;
#pragma unroll
      for (int k = 0; k < kSortPerLane; ++k) {
        const int cls = tid * kSortPerLane + k;
        if (cls <= kSortBins) {                   // cls == kSortBins: the chunk's key count (classes beyond hold zero)
          start[cls] = run;
          offs[(size_t)cls * nchunks + chunk] = (unsigned short)run;
        }
        run += c[k];
      }
    }
    const int any_bad = __syncthreads_or(bad);
    if (tid == 0) bad_chunk[chunk] = any_bad ? 1 : 0;
#pragma unroll
    for (int k = 0; k < kSortCap / kSortThreads; ++k) {
      const int slot = k * kSortThreads + tid;
      const unsigned int w = unsorted[slot];
      if (w != 0xffffffffu) sorted[start[w >> 16] + rank[slot]] = (unsigned short)w;
    }
    __syncthreads();
    static_assert(kSortCap * 2 / 16 == kSortThreads, "one 16-byte vector of sorted keys per thread");
    const uint4 v = reinterpret_cast<const uint4*>(sorted)[tid];
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(reinterpret_cast<uint4*>(keys + (size_t)chunk * kSortCap) + tid), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
                   "l"(keep_policy) : "memory");
  }
}

// CTA (class, slice): blockIdx.x = (kSortSlices - 1 - slice) * kSortBins + class, so the CTAs that only produce a partial
// image (slices 1 ..) are dispatched before the CTAs that wait for them (slice 0).
__global__ void __launch_bounds__(kSortAccThreads, 2) hist_sort_accumulate(
    const unsigned short* __restrict__ keys, const unsigned short* __restrict__ offs, const unsigned char* __restrict__ bad_chunk,
    int nchunks, int local_words, long long npix, int C, long long n, unsigned int* __restrict__ partial,
    unsigned int* __restrict__ flags, uint8_t* __restrict__ out, Header* __restrict__ hdr) {
  extern __shared__ __align__(16) unsigned int sort_acc[];
  __shared__ unsigned int s_total;
  const int cls = blockIdx.x % kSortBins, slice = kSortSlices - 1 - blockIdx.x / kSortBins, tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < local_words; i += kSortAccThreads) sort_acc[i] = 0u;
  if (tid == 0) s_total = 0u;
  pdl_wait();                                    // keys, boundaries and flags of the partition pass are visible from here on
  __syncthreads();
  const unsigned short* o0 = offs + (size_t)cls * nchunks;
  const unsigned short* o1 = o0 + nchunks;
  const int c_begin = (int)((long long)nchunks * slice / kSortSlices), c_end = (int)((long long)nchunks * (slice + 1) / kSortSlices);
  const int rounds = (c_end - c_begin + kSortAccThreads - 1) / kSortAccThreads;
  // a class far above its share (a hot pixel): merge equal keys inside the warp before the shared-memory atomic
  // (same-address shared-memory atomics serialise at ~8 clocks each); decided from this slice's key count
  unsigned int mine = 0;
  for (int c = c_begin + tid; c < c_end; c += kSortAccThreads) mine += (unsigned int)o1[c] - (unsigned int)o0[c];
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0 && mine) atomicAdd(&s_total, mine);
  __syncthreads();
  const bool aggregate = (unsigned long long)s_total * (kSortBins * kSortSlices) > 2ull * (unsigned long long)n;
  for (int r = 0; r < rounds; ++r) {             // every lane of a warp runs the same trip counts (full-mask match below)
    const int c = c_begin + r * kSortAccThreads + tid;
    unsigned int s = 0, e = 0;
    if (c < c_end) { s = o0[c]; e = o1[c]; }
    const unsigned short* kb = keys + (size_t)c * kSortCap;
    const unsigned int k0 = s & ~7u;
    const unsigned int nv = e > s ? ((e + 7u) >> 3) - (s >> 3) : 0u;
    const unsigned int maxv = __reduce_max_sync(0xffffffffu, nv);
    constexpr int kU = 2;
    for (unsigned int j0 = 0; j0 < maxv; j0 += kU) {
      uint4 vec[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        vec[u] = make_uint4(0u, 0u, 0u, 0u);
        if (j0 + u < nv) vec[u] = __ldcg(reinterpret_cast<const uint4*>(kb + k0 + (j0 + u) * 8u));
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const unsigned int k = k0 + (j0 + u) * 8u;
        const unsigned int w[4] = {vec[u].x, vec[u].y, vec[u].z, vec[u].w};
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const bool live = j0 + u < nv && k + m >= s && k + m < e;
          const unsigned int key = (m & 1) ? (w[m >> 1] >> 16) : (w[m >> 1] & 0xffffu);
          if (aggregate) {                         // CTA-uniform branch, the warp is converged here
            const unsigned int peers = __match_any_sync(0xffffffffu, live ? key : 0x10000u + (unsigned int)lane);
            if (live && (int)(__ffs(peers) - 1) == lane) atomicAdd(&sort_acc[key], (unsigned int)__popc(peers));
          } else if (live) {
            atomicAdd(&sort_acc[key], 1u);
          }
        }
      }
    }
  }
  __syncthreads();
  const int vecs = local_words >> 2;
  if (slice > 0) {
    // partial image of this slice, one byte per counter (the image is a count mod 256), then the hand-off flag
    unsigned int* dst = partial + ((size_t)(slice - 1) * kSortBins + cls) * vecs;
    for (int i = tid; i < vecs; i += kSortAccThreads) {
      const uint4 a = reinterpret_cast<const uint4*>(sort_acc)[i];
      __stcg(dst + i, (a.x & 0xffu) | ((a.y & 0xffu) << 8) | ((a.z & 0xffu) << 16) | (a.w << 24));
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(&flags[(slice - 1) * kSortBins + cls], 1u);
    return;
  }
  if (tid < kSortSlices - 1) {
    // the producers of this class have lower block indices, so they were dispatched before this CTA; the wait is bounded
    // all the same (a stuck hand-off reports an error through the status header instead of hanging the device)
    volatile unsigned int* f = flags + tid * kSortBins + cls;
    const long long t0 = clock64();
    while (*f == 0u) {
      __nanosleep(64);
      if (clock64() - t0 > 4000000000LL) { hdr->oob = 1; break; }
    }
    __threadfence();
  }
  __syncthreads();
  for (int i = tid; i < vecs; i += kSortAccThreads) {
    unsigned int add = 0;                         // byte-wise sums mod 256 of the other slices' partial images
#pragma unroll
    for (int sl = 0; sl < kSortSlices - 1; ++sl) {
      const unsigned int b = __ldcg(partial + ((size_t)sl * kSortBins + cls) * vecs + i);
      add = ((add & 0x00ff00ffu) + (b & 0x00ff00ffu)) & 0x00ff00ffu | ((add & 0xff00ff00u) + (b & 0xff00ff00u)) & 0xff00ff00u;
    }
    const uint4 a = reinterpret_cast<const uint4*>(sort_acc)[i];
    const unsigned int v[4] = {a.x + (add & 0xffu), a.y + ((add >> 8) & 0xffu), a.z + ((add >> 16) & 0xffu), a.w + (add >> 24)};
#pragma unroll
    for (int h = 0; h < 2; ++h) {                 // counters 4i + 2h, 4i + 2h + 1: pixel 2i + h of this class (pos, neg)
      const long long px = (long long)(2 * i + h) * kSortBins + cls;
      if (px < npix) hyb_store_px(out, px, C, v[2 * h] & 0xffu, v[2 * h + 1] & 0xffu);
    }
  }
  if (cls == 0) {
    int any = 0;
    for (int c = tid; c < nchunks; c += kSortAccThreads) any |= bad_chunk[c];
    any = __syncthreads_or(any);
    if (tid == 0) hdr->oob = any ? 1 : 0;
  }
}

// ---------------------------------------------------------------- raw record formats
// The reference turns raw recordings into float64 [N,4] .npy files with byte-by-byte Python loops
// (process_data/process_dataset.py): N-Caltech101 40-bit big-endian records (:47-60) and N-Cars / Prophesee
// .dat 64-bit little-endian records (:85-99).  decode_record restates the bit fields; the decoder kernel
// writes the reference's rows, the raw rasteriser feeds the same fields straight into the scatter so that a
// recording costs 5 or 8 bytes per event of HBM (and PCIe) traffic instead of 32.
template <int kFmt>
__device__ __forceinline__ Event decode_record(const uint8_t* __restrict__ rec) {
  Event e;
  if constexpr (kFmt == MEMB_RAW_NCALTECH101) {
    // column 0 = byte 0, column 1 = byte 1, p = bit 7 of byte 2 -> {-1,+1}, t = low 23 bits, big endian
    e.x = (double)rec[0];
    e.y = (double)rec[1];
    e.p = (rec[2] & 0x80u) ? 1.0 : -1.0;
    e.t = (double)(((unsigned int)(rec[2] & 0x7fu) << 16) | ((unsigned int)rec[3] << 8) | (unsigned int)rec[4]);
  } else {
    // uint32 timestamp, then uint32 data: column 0 = bits 0-13, column 1 = bits 14-27, p = bit 28 -> {0,1}
    const unsigned int t = *reinterpret_cast<const unsigned int*>(rec);
    const unsigned int d = *reinterpret_cast<const unsigned int*>(rec + 4);
    e.x = (double)(d & 0x3fffu);
    e.y = (double)((d & 0x0fffc000u) >> 14);
    e.t = (double)t;
    e.p = (d & 0x10000000u) ? 1.0 : 0.0;
  }
  return e;
}

constexpr int kRawRecords = 4096;   // records staged per CTA iteration (20 KB of 5-byte records)

// Stage `count` records starting at record `first` into shared memory with 16-byte loads (raw is 16-byte
// aligned; the first / last partial vectors are read bytewise so nothing outside the buffer is touched).
template <int kBytes>
__device__ __forceinline__ const uint8_t* stage_records(const uint8_t* __restrict__ raw, long long first, int count,
                                                        long long n_total, uint8_t* smem) {
  const long long b0 = first * kBytes, b1 = (first + count) * kBytes;
  const long long v0 = b0 & ~15LL;                       // aligned start (>= 0, inside the buffer)
  const long long v1 = min((b1 + 15) & ~15LL, (n_total * kBytes) & ~15LL);   // last full vector boundary inside the buffer
  for (long long v = v0 + 16LL * threadIdx.x; v < v1; v += 16LL * blockDim.x)
    *reinterpret_cast<uint4*>(smem + (v - v0)) = __ldg(reinterpret_cast<const uint4*>(raw + v));
  for (long long q = max(v1, v0) + threadIdx.x; q < b1; q += blockDim.x) smem[q - v0] = raw[q];   // tail bytes
  __syncthreads();
  return smem + (b0 - v0);
}

template <int kFmt>
__global__ void __launch_bounds__(kThreads) decode_events_kernel(const uint8_t* __restrict__ raw, long long n,
                                                                 double* __restrict__ out) {
  constexpr int kBytes = kFmt == MEMB_RAW_NCALTECH101 ? 5 : 8;
  __shared__ __align__(16) uint8_t stage[kRawRecords * kBytes + 32];
  for (long long first = (long long)blockIdx.x * kRawRecords; first < n; first += (long long)gridDim.x * kRawRecords) {
    const int count = (int)min((long long)kRawRecords, n - first);
    const uint8_t* rec = stage_records<kBytes>(raw, first, count, n, stage);
    for (int r = threadIdx.x; r < count; r += kThreads) {
      const Event e = decode_record<kFmt>(rec + r * kBytes);
      double* o = out + 4 * (first + r);
      asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(o), "d"(e.x), "d"(e.y), "d"(e.t), "d"(e.p) : "memory");
    }
    __syncthreads();
  }
}

// GLOBAL strategy straight from raw records: acc u32 [2][npix] as in hist_scatter_global.
template <int kFmt>
__global__ void __launch_bounds__(kThreads) hist_scatter_raw(const uint8_t* __restrict__ raw, long long n, int W,
                                                             long long npix, unsigned int* __restrict__ acc,
                                                             Header* __restrict__ hdr) {
  constexpr int kBytes = kFmt == MEMB_RAW_NCALTECH101 ? 5 : 8;
  __shared__ __align__(16) uint8_t stage[kRawRecords * kBytes + 32];
  bool bad = false;
  for (long long first = (long long)blockIdx.x * kRawRecords; first < n; first += (long long)gridDim.x * kRawRecords) {
    const int count = (int)min((long long)kRawRecords, n - first);
    const uint8_t* rec = stage_records<kBytes>(raw, first, count, n, stage);
    for (int r = threadIdx.x; r < count; r += kThreads) {
      const Event e = decode_record<kFmt>(rec + r * kBytes);
      const bool pos = e.p == 1.0, neg = e.p == -1.0;
      if (pos || neg) {
        long long idx;
        if (!pixel_index(e.x, e.y, W, npix, idx)) bad = true;
        else atomicAdd(acc + (neg ? npix : 0) + idx, 1u);
      }
    }
    __syncthreads();
  }
  if (bad) hdr->oob = 1;
}

// ---------------------------------------------------------------- extent (H/W = None)
template <bool kAligned>
__global__ void __launch_bounds__(kThreads) hist_extent(const double* __restrict__ ev, long long n,
                                                        Header* __restrict__ hdr) {
  long long mx = LLONG_MIN, my = LLONG_MIN;
  for (long long r = blockIdx.x * (long long)kThreads + threadIdx.x; r < n; r += (long long)gridDim.x * kThreads) {
    Event e = load_event<kAligned>(ev, r);
    // numpy: astype(int) of NaN / huge is INT64_MIN; keep that ordering.
    long long xi = (fabs(e.x) < 9.2e18) ? __double2ll_rz(e.x) : LLONG_MIN;
    long long yi = (fabs(e.y) < 9.2e18) ? __double2ll_rz(e.y) : LLONG_MIN;
    mx = max(mx, xi);
    my = max(my, yi);
  }
  for (int o = 16; o; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    my = max(my, __shfl_xor_sync(0xffffffffu, my, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&hdr->max_x, mx);
    atomicMax(&hdr->max_y, my);
  }
}
__global__ void hist_extent_init(Header* hdr) {
  hdr->max_x = LLONG_MIN;
  hdr->max_y = LLONG_MIN;
}

// ---------------------------------------------------------------- host side
constexpr int kReplDefault = 8;
// copies of the accumulator planes the GLOBAL_REPL strategy keeps (MEMB_HIST_REPLICAS overrides: tuning only)
static int repl_count() {
  static const int k = [] {
    const char* e = getenv("MEMB_HIST_REPLICAS");
    const int v = e ? atoi(e) : kReplDefault;
    return std::max(1, std::min(v, 64));
  }();
  return k;
}

static bool auto_large_is_sort() {
  static const bool v = [] {
    const char* e = getenv("MEMB_HIST_AUTO_LARGE");
    return e != nullptr && e[0] == 's';
  }();
  return v;
}

constexpr int kSortGridSms = 148;   // the SORT chunking is laid out for a B200 (the workspace size must not depend on a device query)

struct Plan {
  int strategy;
  int replicas;
  int tiles, tile_pix;
  size_t ws_bytes;
  size_t off_tkeys, off_acc, off_last;
  // HYBRID: slot map [granules] u16, slot -> granule list [tile_granules] u16, HybState, CTA slices
  int granules, tile_granules;
  size_t off_map, off_sel, off_state, off_slices;
  // SORT: class boundaries [kSortBins + 1][nchunks] u16, one out-of-range flag per chunk, sorted keys [nchunks][kSortChunk] u16
  int nchunks, chunk_len, local_words;
  size_t off_offs, off_bad, off_flags, off_partial, off_keys;
};

constexpr long long kHybMinEvents = 1LL << 20;   // below this the three-launch GLOBAL chain is as fast and needs no sampling

static Plan make_plan(int B, int64_t n, int H, int W, int timesurface, int strategy) {
  Plan p{};
  const long long npix = (long long)H * W;
  if (timesurface) strategy = (strategy == MEMB_HIST_GLOBAL_AGG) ? MEMB_HIST_GLOBAL_AGG : MEMB_HIST_GLOBAL;
  if (strategy == MEMB_HIST_AUTO) {
    // Ragged training batches (many short streams): one privatised tile set per stream.
    // One long stream: every SM streams its share and REDs into L2.
    const long long per_stream = B > 0 ? n / B : n;
    strategy = (B >= 16 && per_stream <= (1 << 20)) ? MEMB_HIST_TILE : MEMB_HIST_GLOBAL;
    // one long stream on a sensor that fits a shared-memory tile: a privatised copy per SM (immune to hot pixels)
    if (B == 1 && npix <= kTileMaxWords && n >= (1 << 18)) strategy = MEMB_HIST_PRIVATE;
    // one long stream on a larger sensor: privatise the hot granules, RED the rest
    if (B == 1 && npix > kTileMaxWords && n >= kHybMinEvents && hyb_granules(W, H) <= kHybMaxGranules) strategy = MEMB_HIST_HYBRID;
    // (SORT -- events binned by pixel class first, no global atomics -- is correct but measured slower than HYBRID on every
    // distribution, profiles/r02_hist_sort_attempt.txt; MEMB_HIST_AUTO_LARGE=sort selects it here: tuning only)
    if (B == 1 && npix > kTileMaxWords && n >= kHybMinEvents && npix <= kSortMaxPixels && auto_large_is_sort()) strategy = MEMB_HIST_SORT;
  }
  if (strategy == MEMB_HIST_SORT && (B != 1 || timesurface || npix > kSortMaxPixels || n < 1)) strategy = MEMB_HIST_GLOBAL;
  if (strategy == MEMB_HIST_HYBRID && (B != 1 || timesurface || hyb_granules(W, H) > kHybMaxGranules || n < kHybSamples))
    strategy = MEMB_HIST_GLOBAL;
  if (strategy == MEMB_HIST_PRIVATE && (B != 1 || npix > kTileMaxWords)) strategy = MEMB_HIST_GLOBAL;
  p.replicas = 1;
  if (strategy == MEMB_HIST_GLOBAL_REPL) {   // the copies only pay for one long stream
    strategy = MEMB_HIST_GLOBAL;
    if (B == 1) p.replicas = repl_count();
  }
  p.strategy = strategy;
  p.tiles = (int)ceil_div<long long>(npix, kTileMaxWords);
  p.tile_pix = (int)round_up<long long>(ceil_div<long long>(npix, p.tiles), 4);
  p.off_tkeys = kHeaderBytes;
  p.off_acc = p.off_tkeys + (timesurface ? round_up<size_t>((size_t)B * 16, 256) : 0);
  p.off_last = p.off_acc + (strategy == MEMB_HIST_TILE ? 0 : (size_t)B * p.replicas * 2 * npix * 4);
  if (strategy == MEMB_HIST_PRIVATE)   // CTA slices [kPrivMaxCtas][words] u16 + one flag per CTA
    p.off_last = p.off_acc + (size_t)kPrivMaxCtas * round_up<size_t>((size_t)npix, 8) * 2 + kPrivMaxCtas * 4;
  p.off_last = round_up<size_t>(p.off_last, 16);
  p.ws_bytes = round_up<size_t>(p.off_last + (timesurface ? (size_t)B * npix * 8 : 0), 16);
  if (strategy == MEMB_HIST_HYBRID) {
    p.granules = (int)hyb_granules(W, H);
    p.tile_granules = hyb_tile_granules(p.granules);
    if (const char* e = getenv("MEMB_HYB_TILE_GRANULES")) p.tile_granules = std::max(1, std::min(p.tile_granules, atoi(e)));   // tuning only
    p.off_map = round_up<size_t>(p.off_acc + (size_t)2 * npix * 4, 16);
    p.off_sel = round_up<size_t>(p.off_map + (size_t)p.granules * 2, 16);
    p.off_state = round_up<size_t>(p.off_sel + (size_t)p.tile_granules * 2, 16);
    p.off_slices = p.off_state + sizeof(HybState);
    p.ws_bytes = round_up<size_t>(p.off_slices + (size_t)kPrivMaxCtas * p.tile_granules * kHybGranule * 2, 16);
  }
  if (strategy == MEMB_HIST_SORT) {
    // equal chunks: every CTA of the partition grid (two per SM) gets the same number of chunks of the same length
    const long long grid = 2LL * kSortGridSms;
    const long long rounds = std::max<long long>(1, ceil_div<long long>(n, grid * kSortCap));
    p.chunk_len = (int)std::min<long long>(kSortCap, round_up<long long>(ceil_div<long long>(std::max<long long>(n, 1), grid * rounds), 16));
    p.nchunks = (int)ceil_div<long long>(std::max<long long>(n, 1), p.chunk_len);
    p.local_words = sort_local_words(npix);
    p.off_offs = kHeaderBytes;
    p.off_bad = round_up<size_t>(p.off_offs + (size_t)(kSortBins + 1) * p.nchunks * 2, 16);
    p.off_flags = round_up<size_t>(p.off_bad + (size_t)p.nchunks, 16);
    p.off_partial = round_up<size_t>(p.off_flags + (size_t)kSortBins * (kSortSlices - 1) * 4, 16);
    p.off_keys = round_up<size_t>(p.off_partial + (size_t)(kSortSlices - 1) * kSortBins * p.local_words, 16);
    p.ws_bytes = round_up<size_t>(p.off_keys + (size_t)p.nchunks * kSortCap * 2, 16);
  }
  return p;
}

}  // namespace hist
}  // namespace memb

using namespace memb;
using namespace memb::hist;

extern "C" size_t memb_hist_workspace_bytes(int B, int64_t n, int H, int W, int timesurface, int strategy) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return make_plan(B, n, H, W, timesurface, strategy).ws_bytes;
}

// Launch with the programmatic-stream-serialisation attribute (see pdl_wait above).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, memb_stream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// PRIVATE strategy: rasterise + finalize (two launches, no initialisation of the workspace needed).
template <int kSrc>
static int run_private(const void* rows, long long n, int W, int H, int C, unsigned int* acc, Header* hdr, uint8_t* out,
                       memb_stream_t stream) {
  const long long npix = (long long)H * W;
  auto kern = hist_private<kSrc>;
  static bool attr_set = false;
  if (!attr_set) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileMaxWords * 4));
    attr_set = true;
  }
  const int words = (int)round_up<long long>(npix, 8);
  // one CTA per SM; fewer when the stream is short (every CTA costs one sensor's worth of slice traffic)
  const long long want = ceil_div<long long>(n, (long long)kTileThreads * kFuseUnroll * 4);
  const int ctas = (int)std::max<long long>(1, std::min<long long>(std::min(num_sms(), kPrivMaxCtas), want));
  unsigned short* slices = reinterpret_cast<unsigned short*>(acc);
  int* flags = reinterpret_cast<int*>(slices + (size_t)kPrivMaxCtas * words);
  kern<<<ctas, kTileThreads, (size_t)words * 4, stream>>>(rows, n, W, H, npix, slices, flags);
  MEMB_LAUNCH_OK("hist_private");
  const int fblocks = (int)ceil_div<long long>(words / 8, 32);
  hist_private_finalize<<<fblocks, dim3(32, kFinRows), 0, stream>>>(slices, ctas, words, flags, npix, C, out, hdr);
  MEMB_LAUNCH_OK("hist_private_finalize");
  return MEMB_OK;
}

// HYBRID strategy: prepare (zero-fill, sample, select, decide) -> GLOBAL scatter | privatised rasteriser -> finalize, chained as
// programmatic dependent launches.  force_mode: -1 = decided on the device from the sample (AUTO), 1 = always privatise.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_smem(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, memb_stream_t stream, Args... args) {
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

template <bool kAligned>
static int run_hybrid(const double* ev, long long n, int W, int H, int C, const Plan& p, int force_mode, char* wsb, uint8_t* out,
                      memb_stream_t stream) {
  const long long npix = (long long)H * W;
  Header* hdr = reinterpret_cast<Header*>(wsb);
  unsigned int* acc = reinterpret_cast<unsigned int*>(wsb + p.off_acc);
  unsigned short* slot_map = reinterpret_cast<unsigned short*>(wsb + p.off_map);
  unsigned short* sel_list = reinterpret_cast<unsigned short*>(wsb + p.off_sel);
  HybState* state = reinterpret_cast<HybState*>(wsb + p.off_state);
  unsigned short* slices = reinterpret_cast<unsigned short*>(wsb + p.off_slices);
  static bool attr_set = false;
  if (!attr_set) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(hist_hybrid<kAligned>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHybSmemBytes));
    MEMB_CUDA_OK(cudaFuncSetAttribute(hist_hybrid_prepare<kAligned>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHybMaxGranules * 4));
    attr_set = true;
  }
  const int sms = num_sms();
  const long long zero_vecs = (long long)((p.off_acc + (size_t)2 * npix * 4 + 15) / 16);
  const int prep_ctas = 1 + std::max(1, std::min(sms - 1, (int)ceil_div<long long>(zero_vecs, kTileThreads)));
  hist_hybrid_prepare<kAligned><<<prep_ctas, kTileThreads, (size_t)p.granules * 4, stream>>>(
      ev, n, W, H, npix, reinterpret_cast<uint4*>(wsb), zero_vecs, p.granules, p.tile_granules, force_mode, slot_map, sel_list, state);
  MEMB_LAUNCH_OK("hist_hybrid_prepare");
  if (force_mode < 0) {        // the plain L2-RED rasteriser, for streams the sample finds spread out (skipped when mode == 1)
    const long long per_cta = (long long)kThreads * kUnroll;
    const long long gx = std::max<long long>(1, std::min<long long>(ceil_div<long long>(n, per_cta), (long long)sms * 32));   // every CTA costs a skip in mode 1
    MEMB_CUDA_OK(launch_pdl(hist_scatter_global<kAligned, false, false>, dim3((unsigned)gx), dim3(kThreads), stream, ev,
                            (const long long*)nullptr, n, W, npix, acc, (unsigned long long*)nullptr, hdr,
                            (const memb_event_aug*)nullptr, 1, (const int*)&state->mode));
    MEMB_LAUNCH_OK("hist_scatter_global");
  }
  const int ctas = std::min(sms, kPrivMaxCtas);
  MEMB_CUDA_OK(launch_pdl_smem(hist_hybrid<kAligned>, dim3(ctas), dim3(kTileThreads), (size_t)kHybSmemBytes, stream, ev, n, W, H, npix,
                               acc, (const unsigned short*)slot_map, p.granules, p.tile_granules, (const HybState*)state, slices, hdr));
  MEMB_LAUNCH_OK("hist_hybrid");
  const int cold_blocks = (int)ceil_div<long long>(npix, 256), hot_blocks = ceil_div(p.tile_granules, 4);
  MEMB_CUDA_OK(launch_pdl_smem(hist_hybrid_finalize, dim3(hot_blocks + cold_blocks), dim3(32, kHybFinRows), (size_t)0, stream,
                               (const unsigned int*)acc, (const unsigned short*)slot_map, (const unsigned short*)sel_list,
                               (const HybState*)state, (const unsigned short*)slices, ctas, p.tile_granules, hot_blocks, npix, W, H, C,
                               out));
  MEMB_LAUNCH_OK("hist_hybrid_finalize");
  return MEMB_OK;
}

// SORT strategy: partition (counting sort of 16-bit keys per 8192-event chunk) -> accumulate (one CTA per pixel class), chained
// as a programmatic dependent launch; nothing in the workspace needs initialising.
template <bool kAligned>
static int run_sort(const double* ev, long long n, int W, int H, int C, const Plan& p, char* wsb, uint8_t* out, memb_stream_t stream) {
  const long long npix = (long long)H * W;
  Header* hdr = reinterpret_cast<Header*>(wsb);
  unsigned short* offs = reinterpret_cast<unsigned short*>(wsb + p.off_offs);
  unsigned char* bad = reinterpret_cast<unsigned char*>(wsb + p.off_bad);
  unsigned short* keys = reinterpret_cast<unsigned short*>(wsb + p.off_keys);
  unsigned int* flags = reinterpret_cast<unsigned int*>(wsb + p.off_flags);
  unsigned int* partial = reinterpret_cast<unsigned int*>(wsb + p.off_partial);
  static bool attr_set = false;
  if (!attr_set) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(hist_sort_partition<kAligned>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemBytes));
    MEMB_CUDA_OK(cudaFuncSetAttribute(hist_sort_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    attr_set = true;
  }
  const int ctas = std::max(1, std::min(2 * kSortGridSms, p.nchunks));
  hist_sort_partition<kAligned><<<ctas, kSortThreads, kSortSmemBytes, stream>>>(ev, n, W, H, npix, p.nchunks, p.chunk_len, keys, offs, bad, flags);
  MEMB_LAUNCH_OK("hist_sort_partition");
  MEMB_CUDA_OK(launch_pdl_smem(hist_sort_accumulate, dim3(kSortBins * kSortSlices), dim3(kSortAccThreads), (size_t)p.local_words * 4, stream,
                               (const unsigned short*)keys, (const unsigned short*)offs, (const unsigned char*)bad, p.nchunks,
                               p.local_words, npix, C, (long long)n, partial, flags, out, hdr));
  MEMB_LAUNCH_OK("hist_sort_accumulate");
  return MEMB_OK;
}

static int run_hist(const double* ev, int64_t n, const int64_t* offsets, int B, int64_t max_stream_len,
                    const memb_event_aug* aug, int H, int W, int C, int timesurface, int strategy,
                    uint8_t* out, void* ws, size_t ws_bytes, memb_stream_t stream) {
  MEMB_REQUIRE(B >= 1 && H >= 1 && W >= 1, "hist: B, H, W must be positive (B=%d H=%d W=%d)", B, H, W);
  MEMB_REQUIRE(C == 2 || C == 3, "hist: C must be 2 or 3, got %d", C);
  MEMB_REQUIRE(!(timesurface && C != 3), "hist: the time surface needs C == 3");
  MEMB_REQUIRE(n >= 0 && (n == 0 || ev != nullptr), "hist: null event pointer");
  MEMB_REQUIRE(offsets != nullptr || B == 1, "hist: a batch needs row offsets");
  MEMB_REQUIRE(out != nullptr && ws != nullptr, "hist: null output / workspace");
  MEMB_REQUIRE((((uintptr_t)ev) & 7u) == 0 && (((uintptr_t)ws) & 15u) == 0, "hist: misaligned pointer");
  MEMB_REQUIRE(strategy >= MEMB_HIST_AUTO && strategy <= MEMB_HIST_SORT, "hist: unknown strategy %d", strategy);
  const long long npix = (long long)H * W;
  const Plan p = make_plan(B, n, H, W, timesurface, strategy);
  if (ws_bytes < p.ws_bytes)
    return fail(MEMB_EWORKSPACE, "hist: workspace %zu B < required %zu B", ws_bytes, p.ws_bytes);
  if (max_stream_len <= 0 || max_stream_len > n) max_stream_len = n;

  char* wsb = static_cast<char*>(ws);
  Header* hdr = reinterpret_cast<Header*>(wsb);
  unsigned long long* tkeys = timesurface ? reinterpret_cast<unsigned long long*>(wsb + p.off_tkeys) : nullptr;
  unsigned int* acc = reinterpret_cast<unsigned int*>(wsb + p.off_acc);
  unsigned long long* last = reinterpret_cast<unsigned long long*>(wsb + p.off_last);
  const bool aligned = (((uintptr_t)ev) & 31u) == 0;
  const long long* offs = reinterpret_cast<const long long*>(offsets);
  const int sms = num_sms();

  if (p.strategy == MEMB_HIST_HYBRID && aug == nullptr && n > 0) {
    const int force = strategy == MEMB_HIST_HYBRID ? 1 : -1;      // AUTO lets the sample decide
    return aligned ? run_hybrid<true>(ev, n, W, H, C, p, force, wsb, out, stream) : run_hybrid<false>(ev, n, W, H, C, p, force, wsb, out, stream);
  }
  if (p.strategy == MEMB_HIST_SORT && aug == nullptr && n > 0)
    return aligned ? run_sort<true>(ev, n, W, H, C, p, wsb, out, stream) : run_sort<false>(ev, n, W, H, C, p, wsb, out, stream);
  const bool use_private = n > 0 && p.strategy == MEMB_HIST_PRIVATE && aug == nullptr && !timesurface;
  if (!use_private) {  // zero the header (+ accumulators) and seed the min/max keys
    const long long n_vec = (long long)((p.strategy == MEMB_HIST_TILE ? (size_t)kHeaderBytes : p.ws_bytes) / 16);
    const int blocks = (int)std::min<long long>(ceil_div<long long>(n_vec, 256), (long long)sms * 8);
    hist_init<<<blocks, 256, 0, stream>>>(reinterpret_cast<uint4*>(wsb), n_vec,
                                          timesurface ? (long long)(p.off_tkeys / 16) : -1LL, B);
    MEMB_LAUNCH_OK("hist_init");
  }

  if (p.strategy == MEMB_HIST_TILE) {
    const size_t smem = (size_t)p.tile_pix * 4;
    auto kern = aligned ? hist_tile_smem<true> : hist_tile_smem<false>;
    static bool attr_set[2] = {false, false};
    if (!attr_set[aligned]) {
      MEMB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileMaxWords * 4));
      attr_set[aligned] = true;
    }
    kern<<<dim3(p.tiles, B), kTileThreads, smem, stream>>>(ev, offs, n, W, npix, p.tile_pix, C, out, hdr, aug);
    MEMB_LAUNCH_OK("hist_tile_smem");
    return MEMB_OK;
  }

  // grid.x: enough CTAs to cover the longest stream once, capped at 64 CTAs per SM overall
  const long long per_cta = (long long)kThreads * kUnroll;
  long long gx = std::max<long long>(1, ceil_div<long long>(max_stream_len, per_cta));
  const long long cap = std::max<long long>(1, ((long long)sms * 64 + B - 1) / B);
  gx = std::min(gx, cap);
  const dim3 grid((unsigned)gx, (unsigned)B);

  if (timesurface && n > 0) {
    if (aligned) hist_time_range<true><<<grid, kThreads, 0, stream>>>(ev, offs, n, tkeys, aug);
    else hist_time_range<false><<<grid, kThreads, 0, stream>>>(ev, offs, n, tkeys, aug);
    MEMB_LAUNCH_OK("hist_time_range");
  }
  if (use_private) {
    return aligned ? run_private<0>(ev, n, W, H, C, acc, hdr, out, stream) : run_private<-1>(ev, n, W, H, C, acc, hdr, out, stream);
  } else if (n > 0) {
    const bool agg = p.strategy == MEMB_HIST_GLOBAL_AGG;
#define MEMB_SCATTER(A, G, T)                                                                        \
  MEMB_CUDA_OK(launch_pdl(hist_scatter_global<A, G, T>, grid, dim3(kThreads), stream, ev, offs, (long long)n, W, npix, acc, \
                          last, hdr, aug, p.replicas, (const int*)nullptr))
    if (timesurface) {
      if (aligned) { if (agg) MEMB_SCATTER(true, true, true); else MEMB_SCATTER(true, false, true); }
      else { if (agg) MEMB_SCATTER(false, true, true); else MEMB_SCATTER(false, false, true); }
    } else {
      if (aligned) { if (agg) MEMB_SCATTER(true, true, false); else MEMB_SCATTER(true, false, false); }
      else { if (agg) MEMB_SCATTER(false, true, false); else MEMB_SCATTER(false, false, false); }
    }
#undef MEMB_SCATTER
    MEMB_LAUNCH_OK("hist_scatter_global");
  }
  {
    long long fx = std::min<long long>(ceil_div<long long>(npix, 256), std::max<long long>(1, (long long)sms * 8 / B));
    const dim3 fgrid((unsigned)std::max<long long>(1, fx), (unsigned)B);
    if (timesurface)
      MEMB_CUDA_OK(launch_pdl(hist_finalize<true>, fgrid, dim3(256), stream, acc, last, tkeys, ev, offs, npix, C, out, p.replicas,
                              aug, (long long)n));
    else
      MEMB_CUDA_OK(launch_pdl(hist_finalize<false>, fgrid, dim3(256), stream, acc, nullptr, nullptr, ev, offs, npix, C, out, p.replicas,
                              (const memb_event_aug*)nullptr, (long long)n));
    MEMB_LAUNCH_OK("hist_finalize");
  }
  return MEMB_OK;
}

extern "C" int memb_hist_u8(const double* ev, int64_t n, const int64_t* offsets, int B,
                            int64_t max_stream_len, int H, int W, int C, int timesurface, int strategy,
                            uint8_t* out, void* ws, size_t ws_bytes, memb_stream_t stream) {
  return run_hist(ev, n, offsets, B, max_stream_len, nullptr, H, W, C, timesurface, strategy, out, ws, ws_bytes,
                  stream);
}

extern "C" int memb_hist_aug_u8(const double* ev, int64_t n, const int64_t* offsets, int B,
                                int64_t max_stream_len, const memb_event_aug* aug, int H, int W, int C,
                                int strategy, uint8_t* out, void* ws, size_t ws_bytes, memb_stream_t stream) {
  MEMB_REQUIRE(aug != nullptr, "hist_aug: null augmentation array");
  MEMB_REQUIRE((((uintptr_t)aug) & 7u) == 0, "hist_aug: misaligned augmentation array");
  return run_hist(ev, n, offsets, B, max_stream_len, aug, H, W, C, 0, strategy, out, ws, ws_bytes, stream);
}

extern "C" int memb_hist_aug_tss_u8(const double* ev, int64_t n, const int64_t* offsets, int B,
                                    int64_t max_stream_len, const memb_event_aug* aug, int H, int W, int timesurface,
                                    uint8_t* out, void* ws, size_t ws_bytes, memb_stream_t stream) {
  MEMB_REQUIRE(aug != nullptr, "hist_aug_tss: null augmentation array");
  MEMB_REQUIRE((((uintptr_t)aug) & 7u) == 0, "hist_aug_tss: misaligned augmentation array");
  return run_hist(ev, n, offsets, B, max_stream_len, aug, H, W, 3, timesurface, MEMB_HIST_GLOBAL, out, ws, ws_bytes, stream);
}

extern "C" int memb_event_pipeline_f32(const double* ev, int64_t n, const int64_t* offsets, int B,
                                       const memb_event_aug* aug, const int32_t* crop_tl, int H, int W, int pad_t,
                                       int pad_l, int outH, int outW, int C, float hot_num_stds, int normalize,
                                       float* out, void* ws, size_t ws_bytes, memb_stream_t stream) {
  return memb_event_pipeline_lut_f32(ev, n, offsets, B, aug, crop_tl, H, W, pad_t, pad_l, outH, outW, C, hot_num_stds, normalize,
                                     nullptr, out, ws, ws_bytes, stream);
}

extern "C" int memb_event_pipeline_lut_f32(const double* ev, int64_t n, const int64_t* offsets, int B,
                                           const memb_event_aug* aug, const int32_t* crop_tl, int H, int W, int pad_t,
                                           int pad_l, int outH, int outW, int C, float hot_num_stds, int normalize,
                                           const float* value_lut, float* out, void* ws, size_t ws_bytes,
                                           memb_stream_t stream) {
  MEMB_REQUIRE(B >= 1 && H >= 1 && W >= 1 && outH >= 1 && outW >= 1, "event_pipeline: bad shape");
  MEMB_REQUIRE(C == 2 || C == 3, "event_pipeline: C must be 2 or 3, got %d", C);
  MEMB_REQUIRE((long long)outH * outW <= kTileMaxWords,
               "event_pipeline: the %dx%d output raster does not fit one shared-memory tile (%d pixels); use "
               "memb_hist_aug_u8 + memb_raster_post_f32", outH, outW, kTileMaxWords);
  MEMB_REQUIRE(n >= 0 && (n == 0 || ev != nullptr), "event_pipeline: null event pointer");
  MEMB_REQUIRE(offsets != nullptr || B == 1, "event_pipeline: a batch needs row offsets");
  MEMB_REQUIRE(aug != nullptr && (((uintptr_t)aug) & 7u) == 0, "event_pipeline: null / misaligned augmentation array");
  MEMB_REQUIRE(out != nullptr && (((uintptr_t)out) & 15u) == 0, "event_pipeline: null / misaligned output");
  MEMB_REQUIRE(ws != nullptr && (((uintptr_t)ws) & 15u) == 0 && ws_bytes >= (size_t)kHeaderBytes,
               "event_pipeline: workspace must hold the %d-byte status header", kHeaderBytes);
  MEMB_REQUIRE((((uintptr_t)ev) & 7u) == 0 && pad_t >= 0 && pad_l >= 0, "event_pipeline: misaligned pointer / bad padding");
  MEMB_CUDA_OK(cudaMemsetAsync(ws, 0, kHeaderBytes, stream));
  const bool aligned = (((uintptr_t)ev) & 31u) == 0;
  auto kern = aligned ? event_pipeline_fused<true> : event_pipeline_fused<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[aligned]) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileMaxWords * 4));
    attr_set[aligned] = true;
  }
  const size_t smem = (size_t)round_up<long long>((long long)outH * outW, 4) * 4;
  kern<<<B, kTileThreads, smem, stream>>>(ev, reinterpret_cast<const long long*>(offsets), n, aug, crop_tl, H, W, pad_t,
                                          pad_l, outH, outW, C, hot_num_stds, normalize, value_lut, out,
                                          reinterpret_cast<Header*>(ws));
  MEMB_LAUNCH_OK("event_pipeline_fused");
  return MEMB_OK;
}

extern "C" int memb_event_pipeline_var_f32(const double* ev, int64_t n, const int64_t* offsets, int B, const memb_event_aug* aug,
                                           int canvas_H, int canvas_W, int outH, int outW, int C, float hot_num_stds,
                                           int normalize, float* out, void* ws, size_t ws_bytes, memb_stream_t stream) {
  return memb_event_pipeline_var_tf_f32(ev, n, offsets, B, aug, canvas_H, canvas_W, outH, outW, C, hot_num_stds, normalize, 0, 0,
                                        0.5f, 0, out, ws, ws_bytes, stream);
}

extern "C" int memb_event_pipeline_var_tf_f32(const double* ev, int64_t n, const int64_t* offsets, int B,
                                              const memb_event_aug* aug, int canvas_H, int canvas_W, int outH, int outW, int C,
                                              float hot_num_stds, int normalize, int logtrafo, int gammatrafo, float gamma,
                                              int timesurface, float* out, void* ws, size_t ws_bytes, memb_stream_t stream) {
  MEMB_REQUIRE(!gammatrafo || gamma > 0.0f, "event_pipeline_var: gamma must be positive, got %g", (double)gamma);
  MEMB_REQUIRE(!timesurface || C == 3, "event_pipeline_var: the time surface needs C = 3, got %d", C);
  MEMB_REQUIRE(!timesurface || n < 0xffffffffLL, "event_pipeline_var: too many rows for the time surface's row keys");
  MEMB_REQUIRE(B >= 1 && canvas_H >= 1 && canvas_W >= 1 && outH >= 1 && outW >= 1, "event_pipeline_var: bad shape");
  MEMB_REQUIRE(C == 2 || C == 3, "event_pipeline_var: C must be 2 or 3, got %d", C);
  MEMB_REQUIRE((long long)canvas_H * canvas_W <= kTileMaxWords,
               "event_pipeline_var: the %dx%d canvas does not fit one shared-memory tile (%d pixels)", canvas_H, canvas_W, kTileMaxWords);
  MEMB_REQUIRE(canvas_H <= 7 * outH && canvas_W <= 7 * outW, "event_pipeline_var: down-scaling by more than 7x is not supported");
  MEMB_REQUIRE(n >= 0 && (n == 0 || ev != nullptr), "event_pipeline_var: null event pointer");
  MEMB_REQUIRE(offsets != nullptr || B == 1, "event_pipeline_var: a batch needs row offsets");
  MEMB_REQUIRE(aug != nullptr && (((uintptr_t)aug) & 7u) == 0, "event_pipeline_var: null / misaligned augmentation array");
  MEMB_REQUIRE(out != nullptr && (((uintptr_t)out) & 15u) == 0, "event_pipeline_var: null / misaligned output");
  MEMB_REQUIRE(ws != nullptr && (((uintptr_t)ws) & 15u) == 0 && ws_bytes >= (size_t)kHeaderBytes,
               "event_pipeline_var: workspace must hold the %d-byte status header", kHeaderBytes);
  MEMB_REQUIRE((((uintptr_t)ev) & 7u) == 0, "event_pipeline_var: misaligned event pointer");
  MEMB_CUDA_OK(cudaMemsetAsync(ws, 0, kHeaderBytes, stream));
  const bool aligned = (((uintptr_t)ev) & 31u) == 0;
  auto kern = aligned ? event_pipeline_var_fused<true> : event_pipeline_var_fused<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[aligned]) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileMaxWords * 4));
    attr_set[aligned] = true;
  }
  const size_t smem = (size_t)round_up<long long>((long long)canvas_H * canvas_W, 4) * 4;
  kern<<<B, kTileThreads, smem, stream>>>(ev, reinterpret_cast<const long long*>(offsets), n, aug, canvas_H, canvas_W, outH, outW,
                                          C, hot_num_stds, normalize, logtrafo, gammatrafo, gamma, timesurface, out,
                                          reinterpret_cast<Header*>(ws));
  MEMB_LAUNCH_OK("event_pipeline_var_fused");
  return MEMB_OK;
}

extern "C" int memb_decode_events_f64(const uint8_t* raw, int64_t n_records, int format, double* out,
                                      memb_stream_t stream) {
  MEMB_REQUIRE(format == MEMB_RAW_NCALTECH101 || format == MEMB_RAW_NCARS, "decode: unknown record format %d", format);
  MEMB_REQUIRE(n_records >= 0, "decode: negative record count");
  if (n_records == 0) return MEMB_OK;
  MEMB_REQUIRE(raw != nullptr && out != nullptr, "decode: null pointer");
  MEMB_REQUIRE((((uintptr_t)raw) & 15u) == 0 && (((uintptr_t)out) & 31u) == 0,
               "decode: raw records must be 16-byte aligned and the output 32-byte aligned");
  const int blocks = (int)std::min<long long>(ceil_div<long long>(n_records, kRawRecords), (long long)num_sms() * 8);
  if (format == MEMB_RAW_NCALTECH101) decode_events_kernel<MEMB_RAW_NCALTECH101><<<blocks, kThreads, 0, stream>>>(raw, n_records, out);
  else decode_events_kernel<MEMB_RAW_NCARS><<<blocks, kThreads, 0, stream>>>(raw, n_records, out);
  MEMB_LAUNCH_OK("decode_events_kernel");
  return MEMB_OK;
}

extern "C" int memb_hist_raw_u8(const uint8_t* raw, int64_t n_records, int format, int H, int W, int C, uint8_t* out,
                                void* ws, size_t ws_bytes, memb_stream_t stream) {
  MEMB_REQUIRE(format == MEMB_RAW_NCALTECH101 || format == MEMB_RAW_NCARS, "hist_raw: unknown record format %d", format);
  MEMB_REQUIRE(H >= 1 && W >= 1 && (C == 2 || C == 3), "hist_raw: bad shape (H=%d W=%d C=%d)", H, W, C);
  MEMB_REQUIRE(n_records >= 0 && (n_records == 0 || raw != nullptr), "hist_raw: null record pointer");
  MEMB_REQUIRE(out != nullptr && ws != nullptr, "hist_raw: null output / workspace");
  MEMB_REQUIRE((((uintptr_t)raw) & 15u) == 0 && (((uintptr_t)ws) & 15u) == 0, "hist_raw: misaligned pointer");
  const long long npix = (long long)H * W;
  const Plan p = make_plan(1, n_records, H, W, 0, MEMB_HIST_GLOBAL);
  if (ws_bytes < p.ws_bytes)
    return fail(MEMB_EWORKSPACE, "hist_raw: workspace %zu B < required %zu B", ws_bytes, p.ws_bytes);
  char* wsb = static_cast<char*>(ws);
  Header* hdr = reinterpret_cast<Header*>(wsb);
  unsigned int* acc = reinterpret_cast<unsigned int*>(wsb + p.off_acc);
  // small sensor + long recording + a workspace sized with MEMB_HIST_AUTO: the privatised strategy, decoding on the fly
  const Plan pa = make_plan(1, n_records, H, W, 0, MEMB_HIST_AUTO);
  if (pa.strategy == MEMB_HIST_PRIVATE && ws_bytes >= pa.ws_bytes && n_records > 0) {
    unsigned int* acc_p = reinterpret_cast<unsigned int*>(wsb + pa.off_acc);
    return format == MEMB_RAW_NCALTECH101 ? run_private<MEMB_RAW_NCALTECH101>(raw, n_records, W, H, C, acc_p, hdr, out, stream)
                                          : run_private<MEMB_RAW_NCARS>(raw, n_records, W, H, C, acc_p, hdr, out, stream);
  }
  const int sms = num_sms();
  {
    const long long n_vec = (long long)(p.ws_bytes / 16);
    const int blocks = (int)std::min<long long>(ceil_div<long long>(n_vec, 256), (long long)sms * 8);
    hist_init<<<blocks, 256, 0, stream>>>(reinterpret_cast<uint4*>(wsb), n_vec, -1LL, 1);
    MEMB_LAUNCH_OK("hist_init");
  }
  if (n_records > 0) {
    const int blocks = (int)std::min<long long>(ceil_div<long long>(n_records, kRawRecords), (long long)sms * 8);
    if (format == MEMB_RAW_NCALTECH101)
      hist_scatter_raw<MEMB_RAW_NCALTECH101><<<blocks, kThreads, 0, stream>>>(raw, n_records, W, npix, acc, hdr);
    else
      hist_scatter_raw<MEMB_RAW_NCARS><<<blocks, kThreads, 0, stream>>>(raw, n_records, W, npix, acc, hdr);
    MEMB_LAUNCH_OK("hist_scatter_raw");
  }
  {
    const long long fx = std::min<long long>(ceil_div<long long>(npix, 256), (long long)sms * 8);
    hist_finalize<false><<<(unsigned)std::max<long long>(1, fx), 256, 0, stream>>>(acc, nullptr, nullptr, nullptr, nullptr,
                                                                                  npix, C, out, 1, nullptr, 0LL);
    MEMB_LAUNCH_OK("hist_finalize");
  }
  return MEMB_OK;
}

extern "C" int memb_hist_status(const void* ws, memb_stream_t stream) {
  MEMB_REQUIRE(ws != nullptr, "hist_status: null workspace");
  int flag = 0;
  MEMB_CUDA_OK(cudaMemcpyAsync(&flag, ws, sizeof(int), cudaMemcpyDeviceToHost, stream));
  MEMB_CUDA_OK(cudaStreamSynchronize(stream));
  if (flag == 2)
    return fail(MEMB_EINVAL, "zero-size array to reduction operation maximum which has no identity (a stream was empty, or "
                             "every row was dropped by the shift augmentation: the reference raises here)");
  if (flag == 3) return fail(MEMB_EINVAL, "event pipeline: a recording's extent does not fit the canvas (or has negative coordinates)");
  if (flag) return fail(MEMB_EOOB, "hist: event index out of bounds for the sensor (reference: IndexError)");
  return MEMB_OK;
}

extern "C" int memb_hist_extent(const double* ev, int64_t n, int64_t* max_xy_host, void* ws, size_t ws_bytes,
                                memb_stream_t stream) {
  MEMB_REQUIRE(ev != nullptr && n > 0, "hist_extent: empty stream (reference: ValueError on max of empty)");
  MEMB_REQUIRE(ws != nullptr && ws_bytes >= (size_t)kHeaderBytes && max_xy_host != nullptr, "hist_extent: bad workspace");
  Header* hdr = reinterpret_cast<Header*>(ws);
  hist_extent_init<<<1, 1, 0, stream>>>(hdr);
  MEMB_LAUNCH_OK("hist_extent_init");
  const int blocks = (int)std::min<long long>(ceil_div<long long>(n, kThreads), (long long)num_sms() * 8);
  if ((((uintptr_t)ev) & 31u) == 0) hist_extent<true><<<blocks, kThreads, 0, stream>>>(ev, n, hdr);
  else hist_extent<false><<<blocks, kThreads, 0, stream>>>(ev, n, hdr);
  MEMB_LAUNCH_OK("hist_extent");
  long long host[3];
  MEMB_CUDA_OK(cudaMemcpyAsync(host, reinterpret_cast<char*>(ws) + 8, 16, cudaMemcpyDeviceToHost, stream));
  MEMB_CUDA_OK(cudaStreamSynchronize(stream));
  max_xy_host[0] = host[0];
  max_xy_host[1] = host[1];
  return MEMB_OK;
}
