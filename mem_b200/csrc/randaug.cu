// EventRandAugment on the device (SURVEY 8f N1 remainder).
//
// Replaces the reference's per-sample CPU module (mem/transforms.py:351-471: `num_ops` operations per sample, each one
// torchvision functional call on a uint8 [3, H, W] tensor, wrapped in ToUnit8 / ToFloat32, mem/datasets.py:655-658).  One CTA
// per sample keeps the whole image in shared memory (3 x 224 x 224 = 147 KB), applies the sample's operations one after the
// other -- the host has drawn them in the reference's generator order -- and writes the result once.  Point-wise operations
// run in place; the two gathers (affine / rotate resampling, the 3x3 blur of Sharpness) write to the sample's own output slab
// and the image is read back from there.
//
// Arithmetic follows torchvision 0.26's tensor code path (torchvision/transforms/_functional_tensor.py) float32 operation by
// float32 operation (explicit round-to-nearest intrinsics: no FMA contraction), so every photometric operation is bit-exact
// with the reference; the resampling fixes one evaluation order where the reference's is its BLAS's (see oracle/randaug_ref.py).
#include <algorithm>
#include <cstdint>

#include "common.cuh"

namespace memb {
namespace randaug {

constexpr int kThreads = 1024;
constexpr int kMaxImageBytes = 200 * 1024;

struct Scratch {
  float unit[256];            // ToFloat32 value of every count: c / 255 (one correctly rounded division each)
  unsigned int hist[3][256];
  unsigned int lut[3][256];
  unsigned int sum;
  unsigned int lo[3], hi[3];
  int lut_on[3];
};

__device__ __forceinline__ uint8_t to_u8_trunc(float v) { return (uint8_t)(int)v; }            // .to(torch.uint8) of an in-range value
__device__ __forceinline__ uint8_t clamp_u8(float v) { return to_u8_trunc(fminf(fmaxf(v, 0.f), 255.f)); }
__device__ __forceinline__ uint8_t gray_u8(uint8_t r, uint8_t g, uint8_t b) {
  // (0.2989 * r + 0.587 * g + 0.114 * b).to(uint8): three rounded products, two rounded sums, truncation
  const float v = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, (float)r), __fmul_rn(0.587f, (float)g)), __fmul_rn(0.114f, (float)b));
  return to_u8_trunc(v);
}
__device__ __forceinline__ uint8_t blend_u8(float r0, float r1, uint8_t a, float other) {
  // (ratio * img1 + (1 - ratio) * img2).clamp(0, 255).to(uint8)
  return clamp_u8(__fadd_rn(__fmul_rn(r0, (float)a), __fmul_rn(r1, other)));
}

// In-place point-wise operation over the image bytes, four per thread when possible: f(value, byte index) -> value
template <typename F>
__device__ __forceinline__ void map_bytes(uint8_t* img, int n, bool vec, int tid, F f) {
  if (vec) {
    for (int i = tid; i < n / 4; i += 1024) {
      uchar4 v = reinterpret_cast<uchar4*>(img)[i];
      v.x = f(v.x, 4 * i); v.y = f(v.y, 4 * i + 1); v.z = f(v.z, 4 * i + 2); v.w = f(v.w, 4 * i + 3);
      reinterpret_cast<uchar4*>(img)[i] = v;
    }
  } else {
    for (int i = tid; i < n; i += 1024) img[i] = f(img[i], i);
  }
}

// grid_sample(bilinear, zeros, align_corners=False) of one output pixel for the three channels
__device__ __forceinline__ void sample3(const uint8_t* __restrict__ img, int H, int W, float ix, float iy, float (&out)[3]) {
  const float x0 = floorf(ix), y0 = floorf(iy), x1 = __fadd_rn(x0, 1.f), y1 = __fadd_rn(y0, 1.f);
  const float wx1 = __fsub_rn(x1, ix), wx0 = __fsub_rn(ix, x0), wy1 = __fsub_rn(y1, iy), wy0 = __fsub_rn(iy, y0);
  const float w[4] = {__fmul_rn(wx1, wy1), __fmul_rn(wx0, wy1), __fmul_rn(wx1, wy0), __fmul_rn(wx0, wy0)};
  const float cx[4] = {x0, x1, x0, x1}, cy[4] = {y0, y0, y1, y1};
  out[0] = out[1] = out[2] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (cx[k] >= 0.f && cx[k] < (float)W && cy[k] >= 0.f && cy[k] < (float)H) {
      const int o = (int)cy[k] * W + (int)cx[k];
#pragma unroll
      for (int c = 0; c < 3; ++c) out[c] = __fadd_rn(out[c], __fmul_rn((float)img[c * H * W + o], w[k]));
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) event_randaug(const void* __restrict__ in, int in_f32, int H, int W,
                                                             const memb_randaug_op* __restrict__ ops, int num_ops,
                                                             void* __restrict__ out, int out_f32) {
  extern __shared__ __align__(16) uint8_t ra_smem[];
  const int HW = H * W, CHW = 3 * HW, tid = threadIdx.x, b = blockIdx.x;
  uint8_t* img = ra_smem;
  Scratch* sc = reinterpret_cast<Scratch*>(ra_smem + ((CHW + 15) / 16) * 16);
  uint8_t* stash = static_cast<uint8_t*>(out) + (size_t)b * CHW * (out_f32 ? 4 : 1);   // this sample's slab of the output
  if (tid < 256) sc->unit[tid] = __fdiv_rn((float)tid, 255.f);

  // ---- load (ToUnit8: (255 * x).to(uint8)); four pixels per thread when the slab allows 16-byte / 4-byte accesses
  const bool vec = (CHW & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  // (eight independent loads per thread in flight: a load -> convert -> st.shared loop issues one load per round trip)
  constexpr int kLd = 8;
  if (in_f32) {
    const float* src = static_cast<const float*>(in) + (size_t)b * CHW;
    if (vec) {
      const int n4 = CHW / 4;
      for (int i0 = tid; i0 < n4; i0 += kThreads * kLd) {
        float4 v[kLd];
#pragma unroll
        for (int u = 0; u < kLd; ++u)
          if (i0 + u * kThreads < n4) v[u] = __ldcs(reinterpret_cast<const float4*>(src) + i0 + u * kThreads);
#pragma unroll
        for (int u = 0; u < kLd; ++u)
          if (i0 + u * kThreads < n4)
            reinterpret_cast<uchar4*>(img)[i0 + u * kThreads] =
                make_uchar4(to_u8_trunc(__fmul_rn(255.f, v[u].x)), to_u8_trunc(__fmul_rn(255.f, v[u].y)),
                            to_u8_trunc(__fmul_rn(255.f, v[u].z)), to_u8_trunc(__fmul_rn(255.f, v[u].w)));
      }
    } else {
      for (int i = tid; i < CHW; i += kThreads) img[i] = to_u8_trunc(__fmul_rn(255.f, src[i]));
    }
  } else {
    const uint8_t* src = static_cast<const uint8_t*>(in) + (size_t)b * CHW;
    if (vec) {
      const int n4 = CHW / 4;
      for (int i0 = tid; i0 < n4; i0 += kThreads * kLd) {
        uchar4 v[kLd];
#pragma unroll
        for (int u = 0; u < kLd; ++u)
          if (i0 + u * kThreads < n4) v[u] = reinterpret_cast<const uchar4*>(src)[i0 + u * kThreads];
#pragma unroll
        for (int u = 0; u < kLd; ++u)
          if (i0 + u * kThreads < n4) reinterpret_cast<uchar4*>(img)[i0 + u * kThreads] = v[u];
      }
    } else {
      for (int i = tid; i < CHW; i += kThreads) img[i] = src[i];
    }
  }
  __syncthreads();

  for (int s = 0; s < num_ops; ++s) {
    const memb_randaug_op op = ops[(size_t)b * num_ops + s];
    bool gathered = false;
    switch (op.op) {
      case MEMB_RA_IDENTITY:
        break;
      case MEMB_RA_AFFINE: {
        // _gen_affine_grid: base grid of pixel centres relative to the image centre, theta^T / (w/2, h/2); grid_sample
        const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
        const float r00 = __fdiv_rn(op.theta[0], hw), r10 = __fdiv_rn(op.theta[1], hw), r20 = __fdiv_rn(op.theta[2], hw);
        const float r01 = __fdiv_rn(op.theta[3], hh), r11 = __fdiv_rn(op.theta[4], hh), r21 = __fdiv_rn(op.theta[5], hh);
        const float xo = -hw + 0.5f, yo = -hh + 0.5f;
        auto pixel = [&](int i, float (&v)[3]) {
          const int y = i / W, x = i - y * W;
          const float xb = (float)x + xo, yb = (float)y + yo;
          const float gx = __fadd_rn(__fadd_rn(__fmul_rn(xb, r00), __fmul_rn(yb, r10)), r20);
          const float gy = __fadd_rn(__fadd_rn(__fmul_rn(xb, r01), __fmul_rn(yb, r11)), r21);
          const float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 2.f);
          const float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 2.f);
          sample3(img, H, W, ix, iy, v);
        };
        if (vec && (HW & 3) == 0) {                  // four pixels per thread: one 4-byte store per channel
          for (int i = tid; i < HW / 4; i += kThreads) {
            uint8_t r[3][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float v[3];
              pixel(4 * i + k, v);
#pragma unroll
              for (int c = 0; c < 3; ++c) r[c][k] = (uint8_t)(int)rintf(v[c]);       // torch.round, then .to(uint8)
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
              reinterpret_cast<uchar4*>(stash + c * HW)[i] = make_uchar4(r[c][0], r[c][1], r[c][2], r[c][3]);
          }
        } else {
          for (int i = tid; i < HW; i += kThreads) {
            float v[3];
            pixel(i, v);
#pragma unroll
            for (int c = 0; c < 3; ++c) stash[c * HW + i] = (uint8_t)(int)rintf(v[c]);
          }
        }
        gathered = true;
        break;
      }
      case MEMB_RA_BRIGHTNESS:
        map_bytes(img, CHW, vec, tid, [&](uint8_t v, int) { return blend_u8(op.f0, op.f1, v, 0.f); });
        break;
      case MEMB_RA_COLOR:
        for (int i = tid; i < HW; i += kThreads) {
          const uint8_t r = img[i], g = img[HW + i], bl = img[2 * HW + i];
          const float gr = (float)gray_u8(r, g, bl);
          img[i] = blend_u8(op.f0, op.f1, r, gr);
          img[HW + i] = blend_u8(op.f0, op.f1, g, gr);
          img[2 * HW + i] = blend_u8(op.f0, op.f1, bl, gr);
        }
        break;
      case MEMB_RA_CONTRAST: {
        if (tid == 0) sc->sum = 0u;
        __syncthreads();
        unsigned int part = 0;
        for (int i = tid; i < HW; i += kThreads) part += gray_u8(img[i], img[HW + i], img[2 * HW + i]);
        part = __reduce_add_sync(0xffffffffu, part);
        if ((tid & 31) == 0) atomicAdd(&sc->sum, part);
        __syncthreads();
        const float mean = __fdiv_rn((float)sc->sum, (float)HW);       // integer sum < 2^24: exact in float32 in any order
        map_bytes(img, CHW, vec, tid, [&](uint8_t v, int) { return blend_u8(op.f0, op.f1, v, mean); });
        break;
      }
      case MEMB_RA_SHARPNESS: {
        if (H <= 2 || W <= 2) break;
        const float k1 = __fdiv_rn(1.f, 13.f), k5 = __fdiv_rn(5.f, 13.f);
        auto sharpen = [&](int i) -> uint8_t {
          const int c = i / HW, r = i - c * HW, y = r / W, x = r - y * W;
          float other = (float)img[i];
          if (y > 0 && y < H - 1 && x > 0 && x < W - 1) {
            float acc = 0.f;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
              for (int dx = -1; dx <= 1; ++dx)
                acc = __fadd_rn(acc, __fmul_rn((float)img[i + dy * W + dx], (dy == 0 && dx == 0) ? k5 : k1));
            other = (float)(uint8_t)(int)rintf(acc);
          }
          return blend_u8(op.f0, op.f1, img[i], other);
        };
        if (vec) {
          for (int i = tid; i < CHW / 4; i += kThreads)
            reinterpret_cast<uchar4*>(stash)[i] = make_uchar4(sharpen(4 * i), sharpen(4 * i + 1), sharpen(4 * i + 2), sharpen(4 * i + 3));
        } else {
          for (int i = tid; i < CHW; i += kThreads) stash[i] = sharpen(i);
        }
        gathered = true;
        break;
      }
      case MEMB_RA_POSTERIZE: {
        const uint8_t mask = (uint8_t)((-(1 << (8 - op.ival))) & 0xff);
        map_bytes(img, CHW, vec, tid, [&](uint8_t v, int) { return (uint8_t)(v & mask); });
        break;
      }
      case MEMB_RA_SOLARIZE:
        map_bytes(img, CHW, vec, tid, [&](uint8_t v, int) { return ((float)v >= op.f0) ? (uint8_t)(255 - v) : v; });
        break;
      case MEMB_RA_AUTOCONTRAST: {
        if (tid < 3) { sc->lo[tid] = 255u; sc->hi[tid] = 0u; }
        __syncthreads();
        for (int c = 0; c < 3; ++c) {
          unsigned int lo = 255u, hi = 0u;
          for (int i = tid; i < HW; i += kThreads) {
            const unsigned int v = img[c * HW + i];
            lo = min(lo, v);
            hi = max(hi, v);
          }
          lo = __reduce_min_sync(0xffffffffu, lo);
          hi = __reduce_max_sync(0xffffffffu, hi);
          if ((tid & 31) == 0) { atomicMin(&sc->lo[c], lo); atomicMax(&sc->hi[c], hi); }
        }
        __syncthreads();
        float lo3[3], sc3[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          lo3[c] = (float)sc->lo[c];
          sc3[c] = __fdiv_rn(255.f, __fsub_rn((float)sc->hi[c], lo3[c]));
          if (!isfinite(sc3[c])) { lo3[c] = 0.f; sc3[c] = 1.f; }
        }
        map_bytes(img, CHW, vec, tid, [&](uint8_t v, int i) {
          const int c = i >= 2 * HW ? 2 : (i >= HW ? 1 : 0);
          return clamp_u8(__fmul_rn(__fsub_rn((float)v, c == 0 ? lo3[0] : (c == 1 ? lo3[1] : lo3[2])),
                                    c == 0 ? sc3[0] : (c == 1 ? sc3[1] : sc3[2])));
        });
        break;
      }
      case MEMB_RA_EQUALIZE: {
        for (int i = tid; i < 3 * 256; i += kThreads) (&sc->hist[0][0])[i] = 0u;
        __syncthreads();
        for (int i = tid; i < CHW; i += kThreads) atomicAdd(&sc->hist[i / HW][img[i]], 1u);
        __syncthreads();
        if (tid < 3) {       // _scale_channel: step = floor(sum of the non-zero bins but the last / 255); lut = shifted cumulative sum
          const unsigned int* h = sc->hist[tid];
          int last = -1;
          for (int v = 255; v >= 0; --v) if (h[v]) { last = v; break; }
          unsigned int total = 0;
          for (int v = 0; v < 256; ++v) if (h[v] && v != last) total += h[v];
          const unsigned int step = total / 255u;
          sc->lut_on[tid] = step != 0u;
          if (step) {
            unsigned int run = 0;
            for (int v = 0; v < 256; ++v) {
              sc->lut[tid][v] = min(255u, (run + step / 2u) / step);      // lut[v] uses the cumulative sum BEFORE bin v
              run += h[v];
            }
          }
        }
        __syncthreads();
        map_bytes(img, CHW, vec, tid, [&](uint8_t v, int i) {
          const int c = i >= 2 * HW ? 2 : (i >= HW ? 1 : 0);
          return sc->lut_on[c] ? (uint8_t)sc->lut[c][v] : v;
        });
        break;
      }
      default:
        break;
    }
    __syncthreads();
    if (gathered) {                                   // read the gathered image back from this sample's slab
      if (vec) {
        const int n4 = CHW / 4;
        for (int i0 = tid; i0 < n4; i0 += kThreads * kLd) {
          uchar4 v[kLd];
#pragma unroll
          for (int u = 0; u < kLd; ++u)
            if (i0 + u * kThreads < n4) v[u] = __ldcg(reinterpret_cast<const uchar4*>(stash) + i0 + u * kThreads);
#pragma unroll
          for (int u = 0; u < kLd; ++u)
            if (i0 + u * kThreads < n4) reinterpret_cast<uchar4*>(img)[i0 + u * kThreads] = v[u];
        }
      } else {
        for (int i = tid; i < CHW; i += kThreads) img[i] = stash[i];
      }
      __syncthreads();
    }
  }
  // ---- store (ToFloat32: x.to(float32) / 255)
  if (out_f32) {
    float* dst = static_cast<float*>(out) + (size_t)b * CHW;
    if (vec) {
      for (int i = tid; i < CHW / 4; i += kThreads) {
        const uchar4 v = reinterpret_cast<const uchar4*>(img)[i];
        __stcs(reinterpret_cast<float4*>(dst) + i, make_float4(sc->unit[v.x], sc->unit[v.y], sc->unit[v.z], sc->unit[v.w]));
      }
    } else {
      for (int i = tid; i < CHW; i += kThreads) dst[i] = sc->unit[img[i]];
    }
  } else {
    uint8_t* dst = static_cast<uint8_t*>(out) + (size_t)b * CHW;
    if (vec) {
      for (int i = tid; i < CHW / 4; i += kThreads) reinterpret_cast<uchar4*>(dst)[i] = reinterpret_cast<const uchar4*>(img)[i];
    } else {
      for (int i = tid; i < CHW; i += kThreads) dst[i] = img[i];
    }
  }
}

}  // namespace randaug
}  // namespace memb

using namespace memb;

extern "C" int memb_event_randaug(const void* in, int in_f32, int B, int C, int H, int W, const memb_randaug_op* ops,
                                  int num_ops, void* out, int out_f32, memb_stream_t stream) {
  MEMB_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && num_ops >= 0 && (num_ops == 0 || ops), "event_randaug: bad arguments");
  MEMB_REQUIRE(C == 3, "event_randaug: images must have 3 channels ([pos, 0, neg]), got %d", C);
  const long long chw = 3LL * H * W;
  MEMB_REQUIRE(chw <= randaug::kMaxImageBytes, "event_randaug: a %dx%d image does not fit one SM's shared memory", H, W);
  const size_t smem = (size_t)((chw + 15) / 16) * 16 + sizeof(randaug::Scratch);
  static bool configured = false;
  if (!configured) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(randaug::event_randaug, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      randaug::kMaxImageBytes + (int)sizeof(randaug::Scratch) + 16));
    configured = true;
  }
  randaug::event_randaug<<<B, randaug::kThreads, smem, stream>>>(in, in_f32, H, W, ops, num_ops, out, out_f32);
  MEMB_LAUNCH_OK("event_randaug");
  return MEMB_OK;
}
