// fp32-faithful convolution for the dVAE tokenizer as an implicit GEMM on tcgen05 (kind::tf32, 3xTF32).
//
// The dVAE that produces the visual-token targets runs in fp32 in the reference (the call sits
// outside autocast, mem/engine_for_pretraining.py:139-145) and its output is an ARGMAX over 8192
// logits whose top-1/top-2 margins go down to ~1e-6 -- plain TF32 or bf16 tensor-core
// arithmetic flips tokens.  Every operand is therefore carried as a TF32 pair  v = hi + lo  and each
// K block issues three MMAs into the same TMEM accumulator:  lo*hi + hi*lo + hi*hi  (the 2^-22 lo*lo
// term is dropped), fp32 accumulation.
//
// Convolutions (eventvae/vae/vae_model.py:29-41 ResBlock, :91 Conv2d(4, stride 2, pad 1) + ReLU,
// :101 Conv2d 1x1) are expressed over "pixel-slot" activations  [R slots rows, X slot columns, inner]
// fp32 (separate hi and lo buffers):
//   * 4x4 / stride 2 / pad 1  ==  2x2 / stride 1 over the space-to-depth(2) image of the zero-padded
//     input: slot (Y, X) holds the 2x2 pixel block, inner = (dy, dx, c);
//   * 3x3 / pad 1             ==  3x3 taps over the zero-padded input, inner = c;
//   * 1x1                     ==  one tap (at slot offset (1,1) when reading a padded layout).
// So the A operand of K block (tap, kc) is ONE 3-D TMA box  [BR slot rows][BW slot columns][32 floats]
// at slot offset (tap_y, tap_x): no im2col buffer, zero padding is physically present in the layout
// and TMA zero-fills the tile tails.  An M tile is BR x BW = 128 output pixels; image boundaries are
// handled by giving every image `rows_per_img` virtual rows (the extra ones are computed and dropped).
// The epilogue adds bias (+ fp32 residual), applies ReLU, splits the result into TF32 hi / lo and
// writes it directly in the slot layout the NEXT layer reads (space-to-depth, padded or plain).
//
// Accumulation is two-level.  Tensor-core accumulators round toward zero at every MMA, so a long K loop
// (K = 6144 -> 2304 dependent accumulations) drifts by ~1e-4 relative -- far above the ~1e-6 argmax
// margins.  The K loop is therefore cut into segments of `seg_kblocks` K blocks; each segment accumulates
// into a fresh TMEM buffer and the epilogue warps add the segments in registers with round-to-nearest
// FADDs while the tensor core works on the next segment (the two TMEM buffers alternate).
//
// Pipeline: warp 0 TMA producer (A_hi, A_lo, W_hi, W_lo per stage), warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-11 epilogue (two warps per TMEM lane quarter, 96 columns each); persistent over tiles.
#include <algorithm>

#include "common.cuh"
#include "sm100.cuh"

namespace memb {
namespace conv {

using namespace memb::ptx;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 192;  // 384 output channels = 2 tiles
constexpr int BLOCK_K = 32;   // floats: one 128-byte swizzle row
constexpr int UMMA_K = 8;
constexpr int kThreads = 384;
constexpr int kEpiCols = BLOCK_N / 2;  // columns per epilogue warp
constexpr int A_BYTES = BLOCK_M * 128;
constexpr int B_BYTES = BLOCK_N * 128;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 80 KB
constexpr int STAGES = 2;
constexpr int TMEM_COLS = 512;  // 2 x 192 accumulator columns, power-of-two allocation
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

struct Params {
  int B, OH, OW, Cout;
  int rows_per_img, BW, BR, x_tiles, m_tiles, n_tiles, num_tiles;
  int taps_x, ntaps, tap_y0, tap_x0, kc_per_tap, K, seg;
  const float* bias;
  const float* aux;
  float* d_full;
  float* d_hi;
  float* d_lo;
  unsigned long long* keys;
  long long sB, sy_major, sy_minor, sx_major, sx_minor;
  int pad, shift, relu;
  int* err_flag;
};

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
  // larger value wins; on equal values the smaller index wins (torch.argmax: first maximum)
  uint32_t b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}
__device__ __forceinline__ void st_row32(float* dst, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

__global__ void __launch_bounds__(kThreads, 1)
conv_tf32x3(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
            const __grid_constant__ CUtensorMap tmap_w, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a_hi);
    prefetch_tmap(&tmap_a_lo);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int iters = p.ntaps * p.kc_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
        const int x0 = (m_tile % p.x_tiles) * p.BW, r0 = (m_tile / p.x_tiles) * p.BR;
        for (int it = 0; it < iters; ++it) {
          const int tap = it / p.kc_per_tap, kc = it - tap * p.kc_per_tap;
          const int ty = tap / p.taps_x, tx = tap - ty * p.taps_x;
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 11);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* s0 = smem + stage * STAGE_BYTES;
          tma_load_3d(s0, &tmap_a_hi, &full_bar[stage], kc * BLOCK_K, x0 + tx + p.tap_x0, r0 + ty + p.tap_y0);
          tma_load_3d(s0 + A_BYTES, &tmap_a_lo, &full_bar[stage], kc * BLOCK_K, x0 + tx + p.tap_x0, r0 + ty + p.tap_y0);
          tma_load_2d(s0 + 2 * A_BYTES, &tmap_w, &full_bar[stage], it * BLOCK_K, n_tile * BLOCK_N);
          tma_load_2d(s0 + 2 * A_BYTES + B_BYTES, &tmap_w, &full_bar[stage], p.K + it * BLOCK_K, n_tile * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(2, false, false, BLOCK_M, BLOCK_N);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int it0 = 0; it0 < iters; it0 += p.seg) {
          const int it1 = min(iters, it0 + p.seg);
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, p.err_flag, 12);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int it = it0; it < it1; ++it) {
            mbar_wait(&full_bar[stage], phase, p.err_flag, 13);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + stage * STAGE_BYTES), a_lo = a_hi + A_BYTES;
            const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
            // small terms first: lo*hi, hi*lo, then hi*hi
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_tf32(d_tmem, make_smem_desc_sw128(a_lo + k * 32, 0, 1024), make_smem_desc_sw128(b_hi + k * 32, 0, 1024), idesc, (it != it0) || k != 0);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_tf32(d_tmem, make_smem_desc_sw128(a_hi + k * 32, 0, 1024), make_smem_desc_sw128(b_lo + k * 32, 0, 1024), idesc, 1);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_tf32(d_tmem, make_smem_desc_sw128(a_hi + k * 32, 0, 1024), make_smem_desc_sw128(b_hi + k * 32, 0, 1024), idesc, 1);
            umma_commit(&empty_bar[stage]);
            if (it == it1 - 1) umma_commit(&tmem_full_bar[acc]);  // this segment's partial sum is complete
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3, half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
      const int x0 = (m_tile % p.x_tiles) * p.BW, r0 = (m_tile / p.x_tiles) * p.BR;
      const int i = quad * 32 + lane;
      const int ox = x0 + i % p.BW, r = r0 + i / p.BW;
      const int b = r / p.rows_per_img, oy = r - b * p.rows_per_img;
      const bool live = ox < p.OW && oy < p.OH && b < p.B;
      const long long plain = ((long long)b * p.OH + oy) * p.OW + ox;
      const int py = oy + p.pad, px = ox + p.pad, msk = (1 << p.shift) - 1;
      const long long off = (long long)b * p.sB + (long long)(py >> p.shift) * p.sy_major + (long long)(py & msk) * p.sy_minor +
                            (long long)(px >> p.shift) * p.sx_major + (long long)(px & msk) * p.sx_minor;
      float sum[kEpiCols];
#pragma unroll
      for (int j = 0; j < kEpiCols; ++j) sum[j] = 0.f;
      // ---- add the K segments with round-to-nearest fp32 adds
      for (int it0 = 0; it0 < iters; it0 += p.seg) {
        mbar_wait(&tmem_full_bar[acc], acc_phase, p.err_flag, 14);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BLOCK_N + half * kEpiCols;
#pragma unroll
        for (int c = 0; c < kEpiCols / 32; ++c) {
          uint32_t raw[32];
          tmem_ld32(taddr + c * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c * 32 + j] += __uint_as_float(raw[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      // ---- fused epilogue from registers
      unsigned long long best = 0ull;
#pragma unroll
      for (int c = 0; c < kEpiCols / 32; ++c) {
        const int col0 = n_tile * BLOCK_N + half * kEpiCols + c * 32;
        if (live && col0 < p.Cout) {  // Cout % 32 == 0
          float v[32], lo[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = sum[c * 32 + j] + __ldg(p.bias + col0 + j);
          if (p.aux) {
            const float4* a4 = reinterpret_cast<const float4*>(p.aux + plain * p.Cout + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 q = a4[j];
              v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.d_full) st_row32(p.d_full + plain * p.Cout + col0, v);
          if (p.keys) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const unsigned long long key = argmax_key(v[j], col0 + j);
              best = key > best ? key : best;
            }
          }
          if (p.d_hi) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float h = tf32_rna(v[j]);
              lo[j] = tf32_rna(v[j] - h);
              v[j] = h;
            }
            st_row32(p.d_hi + off + col0, v);
            st_row32(p.d_lo + off + col0, lo);
          }
        }
      }
      if (p.keys && live) atomicMax(p.keys + plain, best);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------- small helpers
// First dVAE layer (C in {2,3}): explicit im2col of the 4x4 / stride 2 / pad 1 window, k = c*16 + ky*4 + kx
// (Conv2d weight order), zero for padding and for k >= C*16; optional per-channel normalisation
// (DiscreteVAE.norm, vae_model.py:133-141) applied to in-bounds pixels.  One thread per (pixel, 4 k's).
__global__ void __launch_bounds__(256) im2col_l1(const float* __restrict__ img, int B, int C, int H, int W, int Kpad,
                                                 const float* __restrict__ mean, const float* __restrict__ stdv,
                                                 float* __restrict__ a_hi, float* __restrict__ a_lo) {
  const int OH = H / 2, OW = W / 2, kq = Kpad / 4;
  const long long total = (long long)B * OH * OW * kq;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t % kq);
    const long long pix = t / kq;
    const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), b = (int)(pix / ((long long)OW * OH));
    const int k0 = q * 4, c = k0 / 16, ky = (k0 % 16) / 4;  // the 4 k's share (c, ky); kx = 0..3
    float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
    const int iy = 2 * oy + ky - 1;
    if (c < C && iy >= 0 && iy < H) {
      const float* row = img + (((long long)b * C + c) * H + iy) * W;
      const float mu = mean ? mean[c] : 0.f, sd = stdv ? stdv[c] : 1.f;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int ix = 2 * ox + kx - 1;
        if (ix >= 0 && ix < W) {
          float v = row[ix];
          if (mean) v = (v - mu) / sd;
          hi[kx] = tf32_rna(v);
          lo[kx] = tf32_rna(v - hi[kx]);
        }
      }
    }
    *reinterpret_cast<float4*>(a_hi + pix * Kpad + k0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(a_lo + pix * Kpad + k0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(256) split_tf32(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                                  long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = src[i], h = tf32_rna(v);
    hi[i] = h;
    lo[i] = tf32_rna(v - h);
  }
}

// keys from the MEMB_EPI_ARGMAX epilogue -> int64 indices
__global__ void __launch_bounds__(256) argmax_decode(const unsigned long long* __restrict__ keys, long long* __restrict__ idx,
                                                     long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    idx[i] = (long long)(0xffffffffu - (uint32_t)(keys[i] & 0xffffffffull));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(sym);
    return (EncodeTiledFn) nullptr;
  }();
  return fn;
}

// fp32 tensor [d2][d1][d0] (d0 innermost, dense), box [b2][b1][32 floats], 128B swizzle.
static int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const long long* dims, const int* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled is unavailable");
  MEMB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && dims[0] % 4 == 0, "conv: operand must be 16-byte aligned");
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bdim[3], estr[3] = {1, 1, 1};
  unsigned long long stride = (unsigned long long)dims[0] * 4;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bdim[i] = (cuuint32_t)box[i];
    if (i > 0) { gstr[i - 1] = stride; stride *= (unsigned long long)dims[i]; }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled (conv) failed with CUresult %d", (int)r);
  return MEMB_OK;
}

}  // namespace conv
}  // namespace memb

using namespace memb;
using namespace memb::conv;

extern "C" int memb_conv_tf32x3(const memb_conv_desc* dp, memb_stream_t stream) {
  MEMB_REQUIRE(dp != nullptr, "conv: null descriptor");
  const memb_conv_desc& d = *dp;
  MEMB_REQUIRE(d.a_hi && d.a_lo && d.w && d.bias, "conv: null operand");
  MEMB_REQUIRE((d.d_hi != nullptr) == (d.d_lo != nullptr), "conv: d_hi and d_lo go together");
  MEMB_REQUIRE(d.d_hi || d.d_full || d.keys, "conv: no output requested");
  MEMB_REQUIRE(d.inner > 0 && d.inner % BLOCK_K == 0, "conv: inner slot length must be a multiple of 32 floats, got %d", d.inner);
  MEMB_REQUIRE(d.Cout > 0 && d.Cout % 32 == 0, "conv: Cout must be a multiple of 32, got %d", d.Cout);
  MEMB_REQUIRE(d.B > 0 && d.OH > 0 && d.OW > 0 && d.taps_x > 0 && d.taps_y > 0 && d.rows_per_img >= d.OH, "conv: bad geometry");
  MEMB_REQUIRE(d.shift == 0 || d.shift == 1, "conv: output shift must be 0 or 1");
  Params p{};
  p.B = d.B; p.OH = d.OH; p.OW = d.OW; p.Cout = d.Cout;
  p.rows_per_img = d.rows_per_img;
  int bw = 1;
  while (bw < 128 && d.OW % (bw * 2) == 0) bw *= 2;
  p.BW = bw; p.BR = BLOCK_M / bw;
  p.x_tiles = ceil_div(d.OW, p.BW);
  const long long vrows = (long long)d.B * d.rows_per_img;
  p.m_tiles = p.x_tiles * (int)ceil_div<long long>(vrows, p.BR);
  p.n_tiles = ceil_div(d.Cout, BLOCK_N);
  p.num_tiles = p.m_tiles * p.n_tiles;
  p.taps_x = d.taps_x; p.ntaps = d.taps_x * d.taps_y; p.tap_y0 = d.tap_y0; p.tap_x0 = d.tap_x0;
  p.kc_per_tap = d.inner / BLOCK_K;
  p.K = p.ntaps * d.inner;
  p.seg = d.seg_kblocks > 0 ? d.seg_kblocks : 4;
  p.keys = reinterpret_cast<unsigned long long*>(d.keys);
  p.bias = d.bias; p.aux = d.aux; p.d_full = d.d_full; p.d_hi = d.d_hi; p.d_lo = d.d_lo;
  p.sB = d.sB; p.sy_major = d.sy_major; p.sy_minor = d.sy_minor; p.sx_major = d.sx_major; p.sx_minor = d.sx_minor;
  p.pad = d.pad; p.shift = d.shift; p.relu = d.relu; p.err_flag = d.err_flag;

  CUtensorMap ta_hi, ta_lo, tw;
  const long long adims[3] = {d.inner, d.x_slots, d.r_slots};
  const int abox[3] = {BLOCK_K, p.BW, p.BR};
  if (int rc = make_tmap_f32(&ta_hi, d.a_hi, 3, adims, abox)) return rc;
  if (int rc = make_tmap_f32(&ta_lo, d.a_lo, 3, adims, abox)) return rc;
  const long long wdims[2] = {2LL * p.K, d.Cout};
  const int wbox[2] = {BLOCK_K, BLOCK_N};
  if (int rc = make_tmap_f32(&tw, d.w, 2, wdims, wbox)) return rc;

  static bool configured = false;
  if (!configured) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(conv_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int grid = std::min(p.num_tiles, num_sms());
  conv_tf32x3<<<grid, kThreads, SMEM_BYTES, stream>>>(ta_hi, ta_lo, tw, p);
  MEMB_LAUNCH_OK("conv_tf32x3");
  return MEMB_OK;
}

extern "C" int memb_dvae_im2col_l1(const float* img, int B, int C, int H, int W, int Kpad, const float* mean,
                                   const float* stdv, float* a_hi, float* a_lo, memb_stream_t s) {
  MEMB_REQUIRE(img && a_hi && a_lo && B > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "im2col_l1: bad arguments");
  MEMB_REQUIRE(Kpad % 32 == 0 && Kpad >= 16 * C, "im2col_l1: Kpad must be a multiple of 32 and >= 16*C");
  MEMB_REQUIRE((mean == nullptr) == (stdv == nullptr), "im2col_l1: mean and std go together");
  const long long total = (long long)B * (H / 2) * (W / 2) * (Kpad / 4);
  const int grid = (int)std::min<long long>(ceil_div<long long>(total, 256), (long long)num_sms() * 32);
  im2col_l1<<<grid, 256, 0, s>>>(img, B, C, H, W, Kpad, mean, stdv, a_hi, a_lo);
  MEMB_LAUNCH_OK("im2col_l1");
  return MEMB_OK;
}

extern "C" int memb_split_tf32(const float* src, float* hi, float* lo, int64_t n, memb_stream_t s) {
  MEMB_REQUIRE(src && hi && lo && n > 0, "split_tf32: bad arguments");
  const int grid = (int)std::min<long long>(ceil_div<long long>(n, 256), (long long)num_sms() * 16);
  split_tf32<<<grid, 256, 0, s>>>(src, hi, lo, n);
  MEMB_LAUNCH_OK("split_tf32");
  return MEMB_OK;
}

extern "C" int memb_argmax_decode(const uint64_t* keys, int64_t* idx, int64_t n, memb_stream_t s) {
  MEMB_REQUIRE(keys && idx && n > 0, "argmax_decode: bad arguments");
  const int grid = (int)std::min<long long>(ceil_div<long long>(n, 256), (long long)num_sms() * 8);
  argmax_decode<<<grid, 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(keys), reinterpret_cast<long long*>(idx), n);
  MEMB_LAUNCH_OK("argmax_decode");
  return MEMB_OK;
}
