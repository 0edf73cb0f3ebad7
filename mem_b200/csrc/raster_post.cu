// Post-raster tensor transforms on sm_100a: uint8 polarity counts -> the float32 model input.
//
// Replaces, for a whole batch, the per-sample torch chain that follows EventArrToImg in
// build_transformNPY (reference mem/datasets.py:639-660): torchvision ToTensor (uint8 HWC -> float32
// CHW / 255), RandomCrop(pad_if_needed=True) with host-drawn offsets, RemoveTimesurface
// (mem/transforms.py:239-247), RemoveHotPixels(num_stds) (:249-275) and NormalizeEvent (:225-237).
//
// HBM-bound byte work: the uint8 batch (L2-sized) is read by up to three small passes, the float32
// output is written once, coalesced.  Bit-exactness: every output value is fl32(c / 255) or
// fl32(c / 255) * fl32(1 / fl32(cmax / 255)) with single correctly rounded float32 operations, the
// same ones the reference executes.  The hot-pixel threshold mean + k*std is computed from exact
// integer sums (sum c, sum c^2) in float64 and rounded to float32 once; the reference accumulates in
// float32, so the two thresholds can differ in the last bits -- the hot set differs only if a count
// c/255 lies within that rounding distance of the threshold (never observed; tests compare exactly).
#include <algorithm>

#include "common.cuh"

namespace memb {
namespace rpost {

constexpr int kThreads = 256;

struct ImgStats {            // one per image, 32 bytes
  unsigned long long s1;     // sum of counts over the cropped polarity channels (padding counts as 0)
  unsigned long long s2;     // sum of squared counts
  unsigned int cmax;         // maximum count that survived the hot-pixel filter
  unsigned int pad[3];
};

struct Geom {
  int H, W, C, outH, outW, pad_t, pad_l;
};

__device__ __forceinline__ void crop_origin(const int32_t* __restrict__ crop_tl, int b, const Geom& g, int& y0, int& x0) {
  y0 = (crop_tl ? crop_tl[2 * b] : 0) - g.pad_t;   // source row of output row 0
  x0 = (crop_tl ? crop_tl[2 * b + 1] : 0) - g.pad_l;
}

// counts of the two polarity channels at output pixel (oy, ox); zero in the padding
__device__ __forceinline__ void load_pol(const uint8_t* __restrict__ img, const Geom& g, int y0, int x0, int oy, int ox,
                                         unsigned int& cp, unsigned int& cn, unsigned int& cm) {
  const int y = y0 + oy, x = x0 + ox;
  cp = cn = cm = 0u;
  if (y >= 0 && y < g.H && x >= 0 && x < g.W) {
    const uint8_t* px = img + ((long long)y * g.W + x) * g.C;
    cp = px[0];
    if (g.C == 3) {
      cm = px[1];
      cn = px[2];
    } else {
      cn = px[1];
    }
  }
}

__device__ __forceinline__ float hot_threshold(const ImgStats& st, long long n, float num_stds) {
  // torch.mean / torch.std (unbiased) of x = c/255 over n values, then mean + num_stds * std
  const double mean = (double)st.s1 / (255.0 * (double)n);
  double var = ((double)st.s2 - (double)st.s1 * (double)st.s1 / (double)n) / ((double)(n - 1) * 255.0 * 255.0);
  var = var > 0.0 ? var : 0.0;
  return (float)(mean + (double)num_stds * sqrt(var));
}

__global__ void __launch_bounds__(kThreads) stats_kernel(const uint8_t* __restrict__ hist, const int32_t* __restrict__ crop_tl,
                                                         Geom g, ImgStats* __restrict__ stats) {
  const int b = blockIdx.y;
  int y0, x0;
  crop_origin(crop_tl, b, g, y0, x0);
  const uint8_t* img = hist + (long long)b * g.H * g.W * g.C;
  unsigned long long s1 = 0, s2 = 0;
  const int npx = g.outH * g.outW;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < npx; i += gridDim.x * kThreads) {
    unsigned int cp, cn, cm;
    load_pol(img, g, y0, x0, i / g.outW, i % g.outW, cp, cn, cm);
    s1 += cp + cn;
    s2 += cp * cp + cn * cn;
  }
  for (int o = 16; o; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0 && (s1 | s2)) {
    atomicAdd(&stats[b].s1, s1);
    atomicAdd(&stats[b].s2, s2);
  }
}

__global__ void __launch_bounds__(kThreads) max_kernel(const uint8_t* __restrict__ hist, const int32_t* __restrict__ crop_tl,
                                                       Geom g, float hot_num_stds, ImgStats* __restrict__ stats) {
  const int b = blockIdx.y;
  int y0, x0;
  crop_origin(crop_tl, b, g, y0, x0);
  const uint8_t* img = hist + (long long)b * g.H * g.W * g.C;
  const int npx = g.outH * g.outW;
  const bool filter = hot_num_stds >= 0.0f;
  const float thr = filter ? hot_threshold(stats[b], 2LL * npx, hot_num_stds) : 0.0f;
  unsigned int m = 0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < npx; i += gridDim.x * kThreads) {
    unsigned int cp, cn, cm;
    load_pol(img, g, y0, x0, i / g.outW, i % g.outW, cp, cn, cm);
    const bool hot = filter && (__fdiv_rn((float)cp, 255.0f) > thr || __fdiv_rn((float)cn, 255.0f) > thr);
    if (!hot) m = max(m, max(cp, cn));
  }
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(&stats[b].cmax, m);
}

// One thread per output pixel: all channels of the pixel, planar float32 stores (coalesced per plane).
__global__ void __launch_bounds__(kThreads) write_kernel(const uint8_t* __restrict__ hist, const int32_t* __restrict__ crop_tl,
                                                         Geom g, int remove_ts, float hot_num_stds, int normalize,
                                                         const ImgStats* __restrict__ stats,
                                                         const float* __restrict__ value_lut, float* __restrict__ out) {
  // value_lut: LogTransform / GammaTransform of the 256 possible c / 255 (evaluated on the host, see memb.h); the filter
  // compares c / 255, the maps act on what survives, NormalizeEvent divides by the mapped maximum (both maps are monotone)
  __shared__ float vlut[256];
  if (threadIdx.x < 256) vlut[threadIdx.x] = value_lut ? value_lut[threadIdx.x] : __fdiv_rn((float)threadIdx.x, 255.0f);
  __syncthreads();
  const int b = blockIdx.y;
  int y0, x0;
  crop_origin(crop_tl, b, g, y0, x0);
  const uint8_t* img = hist + (long long)b * g.H * g.W * g.C;
  const int npx = g.outH * g.outW;
  const bool filter = hot_num_stds >= 0.0f;
  float thr = 0.0f, factor = 1.0f;
  bool scale = false;
  if (filter || normalize) {
    const ImgStats st = stats[b];
    if (filter) thr = hot_threshold(st, 2LL * npx, hot_num_stds);
    if (normalize && st.cmax != 0u && vlut[st.cmax & 0xffu] != 0.0f) {
      // factor = 1.0 / x.max() with x.max() = fl32(cmax / 255) or its mapped value   (transforms.py:234-236)
      factor = __fdiv_rn(1.0f, vlut[st.cmax & 0xffu]);
      scale = true;
    }
  }
  float* o = out + (long long)b * g.C * npx;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < npx; i += gridDim.x * kThreads) {
    unsigned int cp, cn, cm;
    load_pol(img, g, y0, x0, i / g.outW, i % g.outW, cp, cn, cm);
    float vp = __fdiv_rn((float)cp, 255.0f), vn = __fdiv_rn((float)cn, 255.0f);
    if (filter && (vp > thr || vn > thr)) { cp = 0u; cn = 0u; }
    vp = vlut[cp];
    vn = vlut[cn];
    if (scale) {
      vp = __fmul_rn(vp, factor);
      vn = __fmul_rn(vn, factor);
    }
    o[i] = vp;
    if (g.C == 3) {
      o[npx + i] = remove_ts ? 0.0f : __fdiv_rn((float)cm, 255.0f);
      o[2 * npx + i] = vn;
    } else {
      o[npx + i] = vn;
    }
  }
}

}  // namespace rpost
}  // namespace memb

using namespace memb;
using namespace memb::rpost;

extern "C" size_t memb_raster_post_workspace_bytes(int B) {
  return B > 0 ? (size_t)B * sizeof(ImgStats) : 0;
}

extern "C" int memb_raster_post_f32(const uint8_t* hist, int B, int H, int W, int C, const int32_t* crop_tl, int pad_t,
                                    int pad_l, int outH, int outW, int remove_ts, float hot_num_stds, int normalize,
                                    float* out, void* ws, size_t ws_bytes, memb_stream_t stream) {
  return memb_raster_post_lut_f32(hist, B, H, W, C, crop_tl, pad_t, pad_l, outH, outW, remove_ts, hot_num_stds, normalize, nullptr,
                                  out, ws, ws_bytes, stream);
}

extern "C" int memb_raster_post_lut_f32(const uint8_t* hist, int B, int H, int W, int C, const int32_t* crop_tl, int pad_t,
                                        int pad_l, int outH, int outW, int remove_ts, float hot_num_stds, int normalize,
                                        const float* value_lut, float* out, void* ws, size_t ws_bytes, memb_stream_t stream) {
  MEMB_REQUIRE(B >= 1 && H >= 1 && W >= 1 && outH >= 1 && outW >= 1, "raster_post: bad shape");
  MEMB_REQUIRE(C == 2 || C == 3, "raster_post: C must be 2 or 3, got %d", C);
  MEMB_REQUIRE(hist != nullptr && out != nullptr, "raster_post: null pointer");
  MEMB_REQUIRE(pad_t >= 0 && pad_l >= 0, "raster_post: negative padding");
  MEMB_REQUIRE((long long)outH * outW < (1LL << 30), "raster_post: output image too large");
  const bool filter = hot_num_stds >= 0.0f;
  const bool need_stats = filter || normalize;
  if (need_stats) {
    MEMB_REQUIRE(ws != nullptr && (((uintptr_t)ws) & 15u) == 0, "raster_post: null / misaligned workspace");
    if (ws_bytes < memb_raster_post_workspace_bytes(B))
      return fail(MEMB_EWORKSPACE, "raster_post: workspace %zu B < required %zu B", ws_bytes,
                  memb_raster_post_workspace_bytes(B));
  }
  const Geom g{H, W, C, outH, outW, pad_t, pad_l};
  ImgStats* stats = reinterpret_cast<ImgStats*>(ws);
  const int npx = outH * outW;
  // enough CTAs per image to fill the GPU about twice over the batch
  const int per_img = std::max(1, std::min(ceil_div(npx, kThreads), ceil_div(2 * num_sms() * 8, B)));
  const dim3 grid((unsigned)per_img, (unsigned)B);
  if (need_stats) {
    MEMB_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)B * sizeof(ImgStats), stream));
    if (filter) {
      stats_kernel<<<grid, kThreads, 0, stream>>>(hist, crop_tl, g, stats);
      MEMB_LAUNCH_OK("raster_post stats");
    }
    if (normalize) {
      max_kernel<<<grid, kThreads, 0, stream>>>(hist, crop_tl, g, hot_num_stds, stats);
      MEMB_LAUNCH_OK("raster_post max");
    }
  }
  write_kernel<<<grid, kThreads, 0, stream>>>(hist, crop_tl, g, remove_ts, hot_num_stds, normalize, stats, value_lut, out);
  MEMB_LAUNCH_OK("raster_post write");
  return MEMB_OK;
}
