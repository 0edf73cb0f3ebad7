// fp32-faithful convolution for the dVAE tokenizer on the fp16 tensor path: 2 x fp16 split operands on
// tcgen05 (kind::f16, fp32 accumulation), CTA pairs (cta_group::2).
//
// Same algorithm as conv.cu (implicit GEMM over "pixel-slot" activations, two-level accumulation, fused
// bias / residual / ReLU / split / argmax epilogue -- see the header of that file for the layouts; reference:
// eventvae/vae/vae_model.py:29-41 ResBlock, :91 Conv2d(4, s2, p1)+ReLU, :101 Conv2d 1x1), with a cheaper
// operand representation.  conv.cu carries every operand as a TF32 pair (4 B + 4 B per element, TF32 MMAs run at
// half the 16-bit rate) and is bound by L2 -> SMEM operand traffic (~5.2 KB/clk chip-wide, the LTS ceiling).
// Here a value v is carried as
//        v * 2^e  =  hi + lo,      hi = fp16(v * 2^e),  lo = fp16(v * 2^e - hi)
// with a per-tensor power-of-two exponent e chosen by the host (activations: calibrated on the first batch with a
// 16x overflow margin and monitored through `absmax`; weights: from their own maximum).  fp16 carries 11
// significand bits like TF32, so hi + lo holds 22 bits exactly as the TF32 pair does -- when |v 2^e| >= 2^-3; below
// that lo enters the fp16 subnormals and the ABSOLUTE error floors at 2^-25 (in scaled units), i.e. tiny operands
// lose relative but not absolute accuracy, which is what a dot product needs.  fp16 x fp16 products are exact in
// fp32 (22 bits), the three MMAs  lo*hi + hi*lo + hi*hi  accumulate in fp32 in TMEM exactly like the TF32 path.
// Per K element this halves the operand bytes and doubles the MMA rate.
//
// CTA pair: UMMA M = 256 (two 128-pixel M tiles, one per CTA), N = 192; each CTA stages its own A tiles and half
// (96 rows) of the weight tile.  Stage = A_hi + A_lo + W_hi/2 + W_lo/2 = 56 KB, 4 stages.
#include <algorithm>

#include <cuda_fp16.h>

#include "common.cuh"
#include "sm100.cuh"

namespace memb {
namespace conv16 {

using namespace memb::ptx;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 192;  // 384 output channels = 2 tiles
constexpr int BLOCK_K = 64;   // halves: one 128-byte swizzle row
constexpr int UMMA_K = 16;
// Epilogue warps: 4 lane quarters x (kEpiWarps / 4) column parts.  General tiles keep a 96-column partial sum per thread in
// registers (two-level accumulation) and run 8 epilogue warps; tiles whose whole K fits one accumulation segment (the
// first layer: K = 16*C padded to one 64-wide block) need no partial sums, so 12 warps with 64 columns each fit the
// register file: that layer is bound by the epilogue (1.2 GB of hi/lo output per 64 images), not by the tensor pipe.
template <int kEpiWarps> struct Cfg {
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  static constexpr int kEpiCols = BLOCK_N / (kEpiWarps / 4);  // columns per epilogue warp
};
constexpr int A_BYTES = BLOCK_M * 128;
constexpr int B_BYTES = (BLOCK_N / 2) * 128;  // this CTA's half of the weight tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 56 KB
constexpr int STAGES = 4;
constexpr int TMEM_COLS = 512;  // 2 x 192 accumulator columns, power-of-two allocation
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;

struct Params {
  int B, OH, OW, Cout;
  int rows_per_img, BW, BR, x_tiles, m_tiles, n_tiles, num_tiles;
  int taps_x, ntaps, tap_y0, tap_x0, kc_per_tap, K, seg;
  const float* bias;
  const float* aux;
  float* d_full;
  __half* d_hi;
  __half* d_lo;
  unsigned long long* keys;
  unsigned int* absmax;
  long long sB, sy_major, sy_minor, sx_major, sx_minor;
  int pad, shift, relu;
  float acc_scale, out_scale;
  int* err_flag;
};

__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
  // larger value wins; on equal values the smaller index wins (torch.argmax: first maximum)
  uint32_t b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// In every group of four lanes, lane r holds quarters q[0..3] of its own row; afterwards lane r holds quarter r of the
// rows of lanes 4g, 4g+1, 4g+2, 4g+3 (q[j] = old q[r] of lane 4g+j): two butterfly steps of 2 x 16-byte exchanges.
__device__ __forceinline__ uint4 shfl_xor_u4(uint4 v, int m) {
  return make_uint4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                    __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
__device__ __forceinline__ void transpose4x4_quarters(uint4 (&q)[4], int lane) {
  const bool odd = lane & 1, up = lane & 2;
#pragma unroll
  for (int k = 0; k < 4; k += 2) {          // 2x2 blocks: (r, k..k+1) x (r^1)
    const uint4 recv = shfl_xor_u4(odd ? q[k] : q[k + 1], 1);
    if (odd) q[k] = recv; else q[k + 1] = recv;
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {             // blocks of 2: (r, k / k+2) x (r^2)
    const uint4 recv = shfl_xor_u4(up ? q[k] : q[k + 2], 2);
    if (up) q[k] = recv; else q[k + 2] = recv;
  }
}

template <int kEpiWarps, bool kSingle>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg<kEpiWarps>::kThreads, 1)
conv_f16x2(const __grid_constant__ CUtensorMap tmap_a_hi, const __grid_constant__ CUtensorMap tmap_a_lo,
           const __grid_constant__ CUtensorMap tmap_w, const Params p) {
  constexpr int kEpiCols = Cfg<kEpiWarps>::kEpiCols;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0 && p.err_flag) atomicExch(p.err_flag, 91);
    return;
  }
  grid_dependency_trigger();      // the next layer's launch sets itself up under this one's tail
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a_hi);
    prefetch_tmap(&tmap_a_lo);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dependency_wait();         // launched as a programmatic dependent: the previous layer's output is complete from here on
  const int iters = p.ntaps * p.kc_per_tap;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
      const int n_tile = tile % p.n_tiles, m_tile = (tile / p.n_tiles) * 2 + (int)cta_rank;
      const int x0 = (m_tile % p.x_tiles) * p.BW, r0 = (m_tile / p.x_tiles) * p.BR;
      const int wrow = n_tile * BLOCK_N + (int)cta_rank * (BLOCK_N / 2);
      for (int it = 0; it < iters; ++it) {
        const int tap = it / p.kc_per_tap, kc = it - tap * p.kc_per_tap;
        const int ty = tap / p.taps_x, tx = tap - ty * p.taps_x;
        mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 11);
        if (elect_one()) {
          const uint32_t fb = mapa(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t s0 = smem_u32(smem + stage * STAGE_BYTES);
          tma_load_3d_pair(s0, &tmap_a_hi, fb, kc * BLOCK_K, x0 + tx + p.tap_x0, r0 + ty + p.tap_y0);
          tma_load_3d_pair(s0 + A_BYTES, &tmap_a_lo, fb, kc * BLOCK_K, x0 + tx + p.tap_x0, r0 + ty + p.tap_y0);
          tma_load_2d_pair(s0 + 2 * A_BYTES, &tmap_w, fb, it * BLOCK_K, wrow);
          tma_load_2d_pair(s0 + 2 * A_BYTES + B_BYTES, &tmap_w, fb, p.K + it * BLOCK_K, wrow);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA)
      constexpr uint32_t idesc = make_idesc(0, false, false, 2 * BLOCK_M, BLOCK_N);  // f16 x f16 -> f32
      const uint64_t desc0 = make_smem_desc_sw128(smem_u32(smem), 0, 1024);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
        for (int it0 = 0; it0 < iters; it0 += p.seg) {
          const int it1 = min(iters, it0 + p.seg);
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, p.err_flag, 12);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int it = it0; it < it1; ++it) {
            mbar_wait(&full_bar[stage], phase, p.err_flag, 13);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t a_hi = desc0 + (uint64_t)((stage * STAGE_BYTES) >> 4), a_lo = a_hi + (A_BYTES >> 4);
              const uint64_t b_hi = a_hi + (2 * A_BYTES >> 4), b_lo = b_hi + (B_BYTES >> 4);
              // small terms first: lo*hi, hi*lo, then hi*hi
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16_pair(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, (it != it0) || k != 0);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16_pair(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16_pair(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1);
              umma_commit_pair(&empty_bar[stage], 3);
              if (it == it1 - 1) umma_commit_pair(&tmem_full_bar[acc], 3);  // this segment's partial sum is complete
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------------- epilogue (lane quarter x column half)
    const int quad = warp & 3, half = (warp - 4) >> 2;      // lane quarter, column part (kEpiCols columns each)
    const uint32_t tmem_empty_leader0 = mapa(smem_u32(&tmem_empty_bar[0]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    float amax = 0.f;
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
      const int n_tile = tile % p.n_tiles, m_tile = (tile / p.n_tiles) * 2 + (int)cta_rank;
      const int x0 = (m_tile % p.x_tiles) * p.BW, r0 = (m_tile / p.x_tiles) * p.BR;
      const int i = quad * 32 + lane;
      const int ox = x0 + i % p.BW, r = r0 + i / p.BW;
      const int b = r / p.rows_per_img, oy = r - b * p.rows_per_img;
      const bool live = ox < p.OW && oy < p.OH && b < p.B;
      const long long plain = ((long long)b * p.OH + oy) * p.OW + ox;
      const int py = oy + p.pad, px = ox + p.pad, msk = (1 << p.shift) - 1;
      const long long off = (long long)b * p.sB + (long long)(py >> p.shift) * p.sy_major + (long long)(py & msk) * p.sy_minor +
                            (long long)(px >> p.shift) * p.sx_major + (long long)(px & msk) * p.sx_minor;
      float sum[kSingle ? 32 : kEpiCols];
      uint32_t taddr_single = 0;
      if constexpr (!kSingle) {
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
          if (lane == 0) mbar_arrive_cluster(tmem_empty_leader0 + acc * 8);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      } else {
        // one segment per tile: the accumulator is read column group by column group inside the epilogue loop below
        mbar_wait(&tmem_full_bar[acc], acc_phase, p.err_flag, 14);
        tc_fence_after();
        taddr_single = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BLOCK_N + half * kEpiCols;
      }
      // ---- fused epilogue from registers
      float best_v = -INFINITY;
      int best_i = -1;
      // hi / lo rows leave through a 4x4 transpose of 16-byte quarters inside each group of four lanes: one store
      // instruction then writes 64 contiguous bytes of each of 8 rows (8 lines, full sectors) instead of 16 bytes of
      // each of 32 rows (32 lines, half sectors).  Row bases / liveness of the group's four rows:
      long long off_g[4];
      bool live_g[4];
      if (p.d_hi) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          off_g[j] = __shfl_sync(0xffffffffu, off, (lane & ~3) + j);
          live_g[j] = __shfl_sync(0xffffffffu, (int)live, (lane & ~3) + j) != 0;
        }
      }
#pragma unroll
      for (int c = 0; c < kEpiCols / 32; ++c) {
        const int col0 = n_tile * BLOCK_N + half * kEpiCols + c * 32;
        if constexpr (kSingle) {
          uint32_t raw[32];
          tmem_ld32(taddr_single + c * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(raw[j]);
          if (c == kEpiCols / 32 - 1) {          // last read of this accumulator stage: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tmem_empty_leader0 + acc * 8);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
        constexpr int kSumBase = kSingle ? 0 : 32;   // sum[] holds one column group (single) or all of them
        if (col0 < p.Cout) {  // warp-uniform (Cout % 32 == 0); rows that are not live compute on zeros and store nothing
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bq = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
            v[4 * j] = fmaf(sum[c * kSumBase + 4 * j], p.acc_scale, bq.x);
            v[4 * j + 1] = fmaf(sum[c * kSumBase + 4 * j + 1], p.acc_scale, bq.y);
            v[4 * j + 2] = fmaf(sum[c * kSumBase + 4 * j + 2], p.acc_scale, bq.z);
            v[4 * j + 3] = fmaf(sum[c * kSumBase + 4 * j + 3], p.acc_scale, bq.w);
          }
          if (p.aux && live) {
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
          if (live && p.d_hi) {          // range monitor of the stored fp16 operands (logits are never stored as fp16)
#pragma unroll
            for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(v[j]));
          }
          if (p.d_full && live) {
            float4* dst = reinterpret_cast<float4*>(p.d_full + plain * p.Cout + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (p.keys && live) {          // running maximum in float (columns ascend: the first maximum wins), key built once
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (v[j] > best_v) { best_v = v[j]; best_i = col0 + j; }
            }
          }
          if (p.d_hi) {
            uint4 qh[4], ql[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float s0 = v[8 * j + 2 * k] * p.out_scale, s1 = v[8 * j + 2 * k + 1] * p.out_scale;
                const __half2 h = __floats2half2_rn(s0, s1);            // one cvt.rn.f16x2.f32
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
                hw[k] = *reinterpret_cast<const uint32_t*>(&h);
                lw[k] = *reinterpret_cast<const uint32_t*>(&l);
              }
              qh[j] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              ql[j] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            transpose4x4_quarters(qh, lane);
            transpose4x4_quarters(ql, lane);
            const int q = lane & 3;     // this lane now holds quarter q of rows 4g .. 4g+3
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (live_g[j]) {
                *reinterpret_cast<uint4*>(p.d_hi + off_g[j] + col0 + 8 * q) = qh[j];
                *reinterpret_cast<uint4*>(p.d_lo + off_g[j] + col0 + 8 * q) = ql[j];
              }
            }
          }
        }
      }
      if (p.keys && live && best_i >= 0) atomicMax(p.keys + plain, argmax_key(best_v, best_i));
    }
    if (p.absmax) {  // |result| maximum seen by this warp -> one atomic (float bits order like unsigned for x >= 0)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0) atomicMax(p.absmax, __float_as_uint(amax));
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------- small helpers
// First dVAE layer (C in {2,3}): explicit im2col of the 4x4 / stride 2 / pad 1 window, k = c*16 + ky*4 + kx
// (Conv2d weight order), zero for padding and for k >= C*16; optional per-channel normalisation
// (DiscreteVAE.norm, vae_model.py:133-141) applied to in-bounds pixels; fp16 hi / lo of value * scale.
__global__ void __launch_bounds__(256) im2col_l1_f16(const float* __restrict__ img, int B, int C, int H, int W, int Kpad,
                                                     const float* __restrict__ mean, const float* __restrict__ stdv, float scale,
                                                     __half* __restrict__ a_hi, __half* __restrict__ a_lo,
                                                     unsigned int* __restrict__ absmax) {
  const int OH = H / 2, OW = W / 2, kq = Kpad / 4;
  const unsigned int total = (unsigned int)B * OH * OW * kq;   // < 2^31 (checked by the launcher): 32-bit index math
  float amax = 0.f;
  for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int q = (int)(t % (unsigned int)kq);
    const unsigned int pix = t / (unsigned int)kq;
    const unsigned int rowi = pix / (unsigned int)OW;            // b * OH + oy
    const int ox = (int)(pix - rowi * OW), oy = (int)(rowi % (unsigned int)OH), b = (int)(rowi / (unsigned int)OH);
    const int k0 = q * 4, c = k0 / 16, ky = (k0 % 16) / 4;  // the 4 k's share (c, ky); kx = 0..3
    __half hi[4], lo[4];
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) hi[kx] = lo[kx] = __float2half_rn(0.f);
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
          amax = fmaxf(amax, fabsf(v));
          const float s = v * scale;
          hi[kx] = __float2half_rn(s);
          lo[kx] = __float2half_rn(s - __half2float(hi[kx]));
        }
      }
    }
    *reinterpret_cast<uint2*>(a_hi + (long long)pix * Kpad + k0) = make_uint2(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]));
    *reinterpret_cast<uint2*>(a_lo + (long long)pix * Kpad + k0) = make_uint2(pack_h2(lo[0], lo[1]), pack_h2(lo[2], lo[3]));
  }
  if (absmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(absmax, __float_as_uint(amax));
  }
}

__global__ void __launch_bounds__(256) split_f16(const float* __restrict__ src, float scale, __half* __restrict__ hi,
                                                 __half* __restrict__ lo, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = src[i] * scale;
    const __half h = __float2half_rn(s);
    hi[i] = h;
    lo[i] = __float2half_rn(s - __half2float(h));
  }
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

// fp16 tensor [d2][d1][d0] (d0 innermost, dense), box [b2][b1][64 halves], 128B swizzle.
static int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const long long* dims, const int* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled is unavailable");
  MEMB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && dims[0] % 8 == 0, "conv16: operand must be 16-byte aligned");
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bdim[3], estr[3] = {1, 1, 1};
  unsigned long long stride = (unsigned long long)dims[0] * 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bdim[i] = (cuuint32_t)box[i];
    if (i > 0) { gstr[i - 1] = stride; stride *= (unsigned long long)dims[i]; }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled (conv16) failed with CUresult %d", (int)r);
  return MEMB_OK;
}

}  // namespace conv16
}  // namespace memb

using namespace memb;
using namespace memb::conv16;

extern "C" int memb_conv_f16x2(const memb_conv16_desc* dp, memb_stream_t stream) {
  MEMB_REQUIRE(dp != nullptr, "conv16: null descriptor");
  const memb_conv16_desc& d = *dp;
  MEMB_REQUIRE(d.a_hi && d.a_lo && d.w && d.bias, "conv16: null operand");
  MEMB_REQUIRE((d.d_hi != nullptr) == (d.d_lo != nullptr), "conv16: d_hi and d_lo go together");
  MEMB_REQUIRE(d.d_hi || d.d_full || d.keys, "conv16: no output requested");
  MEMB_REQUIRE(d.inner > 0 && d.inner % BLOCK_K == 0, "conv16: inner slot length must be a multiple of 64 halves, got %d", d.inner);
  MEMB_REQUIRE(d.Cout > 0 && d.Cout % 32 == 0, "conv16: Cout must be a multiple of 32, got %d", d.Cout);
  MEMB_REQUIRE(d.B > 0 && d.OH > 0 && d.OW > 0 && d.taps_x > 0 && d.taps_y > 0 && d.rows_per_img >= d.OH, "conv16: bad geometry");
  MEMB_REQUIRE(d.shift == 0 || d.shift == 1, "conv16: output shift must be 0 or 1");
  MEMB_REQUIRE(abs(d.a_exp + d.w_exp) < 120 && abs(d.out_exp) < 120, "conv16: exponents out of range");
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
  p.num_tiles = ceil_div(p.m_tiles, 2) * p.n_tiles;
  p.taps_x = d.taps_x; p.ntaps = d.taps_x * d.taps_y; p.tap_y0 = d.tap_y0; p.tap_x0 = d.tap_x0;
  p.kc_per_tap = d.inner / BLOCK_K;
  p.K = p.ntaps * d.inner;
  p.seg = d.seg_kblocks > 0 ? d.seg_kblocks : 2;
  p.keys = reinterpret_cast<unsigned long long*>(d.keys);
  p.absmax = d.absmax;
  p.bias = d.bias; p.aux = d.aux; p.d_full = d.d_full;
  p.d_hi = reinterpret_cast<__half*>(d.d_hi); p.d_lo = reinterpret_cast<__half*>(d.d_lo);
  p.sB = d.sB; p.sy_major = d.sy_major; p.sy_minor = d.sy_minor; p.sx_major = d.sx_major; p.sx_minor = d.sx_minor;
  p.pad = d.pad; p.shift = d.shift; p.relu = d.relu; p.err_flag = d.err_flag;
  p.acc_scale = ldexpf(1.0f, -(d.a_exp + d.w_exp));
  p.out_scale = ldexpf(1.0f, d.out_exp);

  CUtensorMap ta_hi, ta_lo, tw;
  const long long adims[3] = {d.inner, d.x_slots, d.r_slots};
  const int abox[3] = {BLOCK_K, p.BW, p.BR};
  if (int rc = make_tmap_f16(&ta_hi, d.a_hi, 3, adims, abox)) return rc;
  if (int rc = make_tmap_f16(&ta_lo, d.a_lo, 3, adims, abox)) return rc;
  const long long wdims[2] = {2LL * p.K, d.Cout};
  const int wbox[2] = {BLOCK_K, BLOCK_N / 2};
  if (int rc = make_tmap_f16(&tw, d.w, 2, wdims, wbox)) return rc;

  static bool configured = false;
  if (!configured) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(conv_f16x2<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    MEMB_CUDA_OK(cudaFuncSetAttribute(conv_f16x2<12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int clusters = std::max(1, std::min(p.num_tiles, num_sms() / 2));
  const int iters = p.ntaps * p.kc_per_tap;
  if (iters <= p.seg && !d.keys)
    MEMB_CUDA_OK(launch_dependent(conv_f16x2<12, true>, dim3(2 * clusters), dim3(Cfg<12>::kThreads), (size_t)SMEM_BYTES, stream, ta_hi, ta_lo, tw, p));
  else
    MEMB_CUDA_OK(launch_dependent(conv_f16x2<8, false>, dim3(2 * clusters), dim3(Cfg<8>::kThreads), (size_t)SMEM_BYTES, stream, ta_hi, ta_lo, tw, p));
  MEMB_LAUNCH_OK("conv_f16x2");
  return MEMB_OK;
}

extern "C" int memb_dvae_im2col_l1_f16(const float* img, int B, int C, int H, int W, int Kpad, const float* mean,
                                       const float* stdv, int exp2, void* a_hi, void* a_lo, uint32_t* absmax,
                                       memb_stream_t s) {
  MEMB_REQUIRE(img && a_hi && a_lo && B > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "im2col_l1_f16: bad arguments");
  MEMB_REQUIRE(Kpad % 64 == 0 && Kpad >= 16 * C, "im2col_l1_f16: Kpad must be a multiple of 64 and >= 16*C");
  MEMB_REQUIRE((mean == nullptr) == (stdv == nullptr), "im2col_l1_f16: mean and std go together");
  const long long total = (long long)B * (H / 2) * (W / 2) * (Kpad / 4);
  MEMB_REQUIRE(total < (1LL << 31), "im2col_l1_f16: %lld work items do not fit 32-bit indexing (tokenise in smaller chunks)", total);
  const int grid = (int)std::min<long long>(ceil_div<long long>(total, 256), (long long)num_sms() * 32);
  im2col_l1_f16<<<grid, 256, 0, s>>>(img, B, C, H, W, Kpad, mean, stdv, ldexpf(1.0f, exp2), reinterpret_cast<__half*>(a_hi),
                                     reinterpret_cast<__half*>(a_lo), absmax);
  MEMB_LAUNCH_OK("im2col_l1_f16");
  return MEMB_OK;
}

extern "C" int memb_split_f16(const float* src, int exp2, void* hi, void* lo, int64_t n, memb_stream_t s) {
  MEMB_REQUIRE(src && hi && lo && n > 0, "split_f16: bad arguments");
  const int grid = (int)std::min<long long>(ceil_div<long long>(n, 256), (long long)num_sms() * 16);
  split_f16<<<grid, 256, 0, s>>>(src, ldexpf(1.0f, exp2), reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), n);
  MEMB_LAUNCH_OK("split_f16");
  return MEMB_OK;
}
