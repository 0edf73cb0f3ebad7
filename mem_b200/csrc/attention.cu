// Fused multi-head self-attention with an additive (relative-position) bias on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM, operands staged by TMA), forward and backward, for the short
// sequences of a ViT (N <= 208 tokens, head dim 64): a whole head -- Q, K, V (and dO) -- is resident in
// 128B-swizzled shared memory, nothing of size N x N goes to HBM in the forward, and the backward writes
// dS^T once (bf16, by TMA store) so that the bias gradient can be reduced over the batch.
//
// Reference: Attention.forward, mem/modeling_finetune.py:128-157 (q*scale, q@k^T, + rel_pos_bias,
// softmax, @v) and its autograd backward.
//
// Layouts: qkv / dqkv bf16 [B, N, 3, H, 64] (the QKV GEMM output viewed as in modeling_finetune.py:134);
// out / dout bf16 [B, N, H*64]; bias (forward) and its transpose (backward) in the packed layout written by
// memb_attention_pack_bias: fp32 [H][52][256][4] = (head, column group, row, 4 columns), times log2(e), -inf past N;
// lse fp32 [B, H, N] (natural log); dsT bf16 [B, H, N(key), ldb(query)].
//
// Forward (persistent, one CTA per SM, 288 threads):
//   warp 8   control: TMA loads (Q, K, V of the NEXT head are fetched while the current one is in softmax / PV),
//            tcgen05.mma issue: S_p = Q_p K^T (M=128 query tile p, N=208 keys, fp32 in TMEM), O_p = P_p V
//   warps 0-7 softmax: two warps per TMEM lane quarter, each thread owns half a row (104 keys) in registers:
//            x = s*scale*log2e + bias*log2e, row max exchanged through smem, p = 2^(x-m) -> bf16 P tile in
//            swizzled smem (the A operand of the PV MMA), then the O epilogue (1/l, bf16, global) and lse.
// Backward (one CTA per (b, h), 288 threads), keys on the M axis so that P^T / dS^T come out of TMEM in the
// orientation dV = P^T dO, dK = dS^T Q and dQ = dS K need; query chunks of 128 (then 80) columns:
//   S^T = K_t Q_c^T, dP^T = V_t dO_c^T  ->  p = 2^(s*c1 + b*log2e - lse2), ds = p (dp - delta)  ->  bf16 P^T, dS^T
//   tiles in smem  ->  dV_t += P^T dO_c, dK_t += dS^T Q_c, dQ_c += dS K_t (MN-major A straight from the dS^T tile);
//   TMEM: S^T 128 + dP^T 128 + dV 64 + dK 64 + dQ 2x64 = 512 columns.
#include <algorithm>

#include <cuda_bf16.h>

#include "common.cuh"
#include "sm100.cuh"

namespace memb {
namespace attn {

using namespace memb::ptx;
using bf16 = __nv_bfloat16;

constexpr int kHeadDim = 64;
constexpr int kMaxN = 208;                    // 13 x 16
constexpr int kLoadBytes = kMaxN * 128;       // one TMA box: [208 rows][64 bf16]
constexpr int kThreads = 288;                 // 8 compute warps + 1 control warp
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kBiasGroups = MEMB_ATTN_BIAS_GROUPS;  // packed bias: [H][52 column groups][256 rows][4]

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// [rows][128 B] tile, 16-byte chunks XOR-swizzled by (row & 7)  (== TMA SWIZZLE_128B, UMMA layout type 2)
__device__ __forceinline__ uint32_t sw128(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void tmem_ld32f(uint32_t taddr, float* dst) {  // dst[0..31], constant-indexed by the caller
  uint32_t r[32];
  tmem_ld32(taddr, r);
#pragma unroll
  for (int i = 0; i < 32; ++i) dst[i] = __uint_as_float(r[i]);
}

// =========================================================================================== forward
namespace fwd {
constexpr int OFF_Q = 0;                       // [256][128 B]  (rows >= 208 stale: they only feed unused S rows)
constexpr int OFF_K = 32768;                   // [208][128 B]
constexpr int OFF_V = OFF_K + kLoadBytes;      // [208][128 B]
constexpr int OFF_P = OFF_V + kLoadBytes;      // 2 query tiles x 4 key blocks x [128][128 B]
constexpr int P_TILE = 4 * 16384;
constexpr int OFF_MAX = OFF_P + 2 * P_TILE;    // float [2 tiles][2 halves][128]
constexpr int OFF_SUM = OFF_MAX + 2048;        // float [2 tiles][2 halves][128]
constexpr int OFF_BAR = OFF_SUM + 2048;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;  // + alignment slack
enum { B_QK = 0, B_V, B_S0, B_S1, B_P0, B_P1, B_O0, B_O1, B_OFREE, B_COUNT };
__host__ __device__ constexpr int s_col(int t) { return t * 256; }  // TMEM columns of S_p; O_p reuses the first 64
}  // namespace fwd

struct FwdParams {
  const float* bias;
  int ldb, B, N, H;
  float c1;  // scale * log2(e)
  bf16* out;
  float* lse;
};

template <int TILE>
__device__ __forceinline__ void fwd_softmax_tile(const FwdParams& p, uint8_t* smem, uint32_t tmem_base, int h, int quarter,
                                                 int half, int lane, float& m_out) {
  using namespace fwd;
  const int rl = quarter * 32 + lane;          // row within the tile == TMEM lane
  const int row = TILE * 128 + rl;             // query index
  const bool rv = row < p.N;
  float* smax = reinterpret_cast<float*>(smem + OFF_MAX) + TILE * 256;
  float* ssum = reinterpret_cast<float*>(smem + OFF_SUM) + TILE * 256;
  float x[104];
  {
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + s_col(TILE) + half * 104;
    tmem_ld32f(taddr, x);
    tmem_ld32f(taddr + 32, x + 32);
    tmem_ld32f(taddr + 64, x + 64);
    uint32_t r8[8];
    tmem_ld8(taddr + 96, r8);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[96 + i] = __uint_as_float(r8[i]);
    tmem_ld_wait();
  }
  float mx = -INFINITY;
  if (rv) {
    // packed bias: float4 (h, column group g, row) -> lanes of a warp read 512 contiguous bytes
    const float4* bp = p.bias ? reinterpret_cast<const float4*>(p.bias) + ((long long)h * kBiasGroups + half * 26) * 256 + row
                              : nullptr;
#pragma unroll
    for (int i = 0; i < 26; ++i) {
      if (bp) {  // pre-multiplied by log2(e); columns >= N hold -inf
        const float4 b = __ldg(bp + i * 256);
        x[4 * i + 0] = fmaf(x[4 * i + 0], p.c1, b.x);
        x[4 * i + 1] = fmaf(x[4 * i + 1], p.c1, b.y);
        x[4 * i + 2] = fmaf(x[4 * i + 2], p.c1, b.z);
        x[4 * i + 3] = fmaf(x[4 * i + 3], p.c1, b.w);
      } else {
        const int c = half * 104 + 4 * i;  // warp-uniform
#pragma unroll
        for (int j = 0; j < 4; ++j) x[4 * i + j] = (c + j < p.N) ? x[4 * i + j] * p.c1 : -INFINITY;
      }
      mx = fmaxf(mx, fmaxf(fmaxf(x[4 * i], x[4 * i + 1]), fmaxf(x[4 * i + 2], x[4 * i + 3])));
    }
  }
  smax[half * 128 + rl] = mx;
  named_bar_sync(1, 256);
  const float m = fmaxf(mx, smax[(half ^ 1) * 128 + rl]);
  m_out = m;
  if (rv) {
    float sum = 0.f;
    const uint32_t sp = smem_u32(smem + OFF_P) + TILE * P_TILE;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
      float e[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        e[j] = ex2(x[8 * i + j] - m);
        sum += e[j];
      }
      const int c0 = half * 104 + 8 * i;
      st_shared_v4(sw128(sp + (c0 >> 6) * 16384, rl, (c0 & 63) >> 3), pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]),
                   pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
    }
    ssum[half * 128 + rl] = sum;
  }
}

template <int TILE>
__device__ __forceinline__ void fwd_out_tile(const FwdParams& p, uint8_t* smem, uint32_t tmem_base, int b, int h, int quarter,
                                             int half, int lane, float m) {
  using namespace fwd;
  const int rl = quarter * 32 + lane, row = TILE * 128 + rl;
  float o[32];
  tmem_ld32f(tmem_base + ((uint32_t)(quarter * 32) << 16) + s_col(TILE) + half * 32, o);
  tmem_ld_wait();
  if (row < p.N) {
    const float* ssum = reinterpret_cast<const float*>(smem + OFF_SUM) + TILE * 256;
    const float l = ssum[rl] + ssum[128 + rl];
    const float inv = 1.f / l;
    bf16* dst = p.out + ((long long)b * p.N + row) * p.H * kHeadDim + h * kHeadDim + half * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<uint4*>(dst)[i] =
          make_uint4(pack_bf16(o[8 * i] * inv, o[8 * i + 1] * inv), pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                     pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
    if (half == 0) p.lse[((long long)b * p.H + h) * p.N + row] = (m + lg2(l)) * kLn2;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
attention_fwd_tc(const __grid_constant__ CUtensorMap tmap_qkv, const FwdParams p) {
  using namespace fwd;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nheads = p.B * p.H;
  const bool two = p.N > 128;  // second query tile in use

  if (warp == 8) {
    if (lane == 0) {
      prefetch_tmap(&tmap_qkv);
      mbar_init(&bar[B_QK], 1); mbar_init(&bar[B_V], 1);
      mbar_init(&bar[B_S0], 1); mbar_init(&bar[B_S1], 1);
      mbar_init(&bar[B_P0], 256); mbar_init(&bar[B_P1], 256);
      mbar_init(&bar[B_O0], 1); mbar_init(&bar[B_O1], 1);
      mbar_init(&bar[B_OFREE], 256);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      // -------------------------------------------------------------------------- control thread
      const uint32_t sq = smem_u32(smem + OFF_Q), sk = smem_u32(smem + OFF_K), sv = smem_u32(smem + OFF_V);
      const uint32_t spp = smem_u32(smem + OFF_P);
      constexpr uint32_t idesc_s = make_idesc(1, false, false, 128, kMaxN);
      constexpr uint32_t idesc_o = make_idesc(1, false, true, 128, kHeadDim);
      const int nks = (p.N + 15) / 16;  // key steps of the PV product
      auto load_qk = [&](int head) {
        const int b = head / p.H, h = head % p.H;
        mbar_arrive_expect_tx(&bar[B_QK], 2 * kLoadBytes);
        tma_load_2d(smem + OFF_Q, &tmap_qkv, &bar[B_QK], h * kHeadDim, b * p.N);
        tma_load_2d(smem + OFF_K, &tmap_qkv, &bar[B_QK], (p.H + h) * kHeadDim, b * p.N);
      };
      auto load_v = [&](int head) {
        const int b = head / p.H, h = head % p.H;
        mbar_arrive_expect_tx(&bar[B_V], kLoadBytes);
        tma_load_2d(smem + OFF_V, &tmap_qkv, &bar[B_V], (2 * p.H + h) * kHeadDim, b * p.N);
      };
      if ((int)blockIdx.x < nheads) { load_qk(blockIdx.x); load_v(blockIdx.x); }
      uint32_t ph = 0;
      for (int head = blockIdx.x; head < nheads; head += gridDim.x, ph ^= 1) {
        mbar_wait(&bar[B_QK], ph, nullptr, 1);
        if (head != (int)blockIdx.x) mbar_wait(&bar[B_OFREE], ph ^ 1, nullptr, 2);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t == 0 || two) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tmem_base + s_col(t), make_smem_desc_sw128(sq + t * 16384 + ks * 32, 0, 1024),
                        make_smem_desc_sw128(sk + ks * 32, 0, 1024), idesc_s, ks != 0);
          }
          umma_commit(&bar[B_S0 + t]);
        }
        mbar_wait(&bar[B_S1], ph, nullptr, 3);  // both S tiles done: Q and K are free
        const int next = head + gridDim.x;
        if (next < nheads) load_qk(next);
        mbar_wait(&bar[B_V], ph, nullptr, 4);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&bar[B_P0 + t], ph, nullptr, 5);
          tc_fence_after();
          if (t == 0 || two) {
            for (int ks = 0; ks < nks; ++ks)
              umma_bf16(tmem_base + s_col(t),
                        make_smem_desc_sw128(spp + t * P_TILE + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024),
                        make_smem_desc_sw128(sv + ks * 2048, 8192, 1024), idesc_o, ks != 0);
          }
          umma_commit(&bar[B_O0 + t]);
        }
        mbar_wait(&bar[B_O1], ph, nullptr, 6);  // V is free
        if (next < nheads) load_v(next);
      }
    }
  } else {
    // ------------------------------------------------------------------------------ softmax / epilogue warps
    const int quarter = warp & 3, half = warp >> 2;
    uint32_t ph = 0;
    for (int head = blockIdx.x; head < nheads; head += gridDim.x, ph ^= 1) {
      const int b = head / p.H, h = head % p.H;
      float m0 = 0.f, m1 = 0.f;
      mbar_wait(&bar[B_S0], ph, nullptr, 7);
      tc_fence_after();
      fwd_softmax_tile<0>(p, smem, tmem_base, h, quarter, half, lane, m0);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar[B_P0]);
      mbar_wait(&bar[B_S1], ph, nullptr, 8);
      tc_fence_after();
      if (two) fwd_softmax_tile<1>(p, smem, tmem_base, h, quarter, half, lane, m1);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar[B_P1]);
      mbar_wait(&bar[B_O0], ph, nullptr, 9);
      tc_fence_after();
      fwd_out_tile<0>(p, smem, tmem_base, b, h, quarter, half, lane, m0);
      mbar_wait(&bar[B_O1], ph, nullptr, 10);
      tc_fence_after();
      if (two) fwd_out_tile<1>(p, smem, tmem_base, b, h, quarter, half, lane, m1);
      tc_fence_before();
      mbar_arrive(&bar[B_OFREE]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =========================================================================================== backward
namespace bwd {
constexpr int OFF_Q = 0;                        // [208][128 B]
constexpr int OFF_DO = kLoadBytes;              // [208][128 B]
constexpr int OFF_K = 2 * kLoadBytes;           // [256][128 B] (rows >= 208 stale: unused S^T rows only)
constexpr int OFF_V = OFF_K + 32768;            // [256][128 B]
constexpr int OFF_PT = OFF_V + 32768;           // 2 query blocks x [128 keys][64 queries]
constexpr int OFF_DST = OFF_PT + 32768;         // same shape: dS^T
constexpr int OFF_NL = OFF_DST + 32768;         // float[208]: -lse*log2e  (-inf for q >= N)
constexpr int OFF_DELTA = OFF_NL + 1024;        // float[208]: rowsum(dO * O)
constexpr int OFF_BAR = OFF_DELTA + 1024;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
enum { B_LOAD = 0, B_S, B_PD, B_MMA2, B_ACCFREE, B_COUNT };
constexpr int COL_ST = 0, COL_DPT = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;
}  // namespace bwd

struct BwdParams {
  const bf16* out;
  const bf16* dout;
  const float* lse;
  const float* biasT;
  int ldb, B, N, H;
  float scale, c1;
  bf16* dqkv;
  int write_ds;
};

// One group of G (32 or 8) query columns of this thread's key row: S^T, dP^T -> P^T, dS^T (bf16, swizzled smem).
template <int G>
__device__ __forceinline__ void bwd_group(const BwdParams& p, uint8_t* smem, uint32_t taddr, const float4* btrow, bool kv,
                                          int q0 /* global query index */, int cl0 /* column within the chunk */, int rl) {
  using namespace bwd;
  float s[G], dp[G];
  if constexpr (G == 32) {
    tmem_ld32f(taddr + COL_ST, s);
    tmem_ld32f(taddr + COL_DPT, dp);
  } else {
    uint32_t a[8], c[8];
    tmem_ld8(taddr + COL_ST, a);
    tmem_ld8(taddr + COL_DPT, c);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = __uint_as_float(a[i]); dp[i] = __uint_as_float(c[i]); }
  }
  tmem_ld_wait();
  const float* nl = reinterpret_cast<const float*>(smem + OFF_NL);
  const float* dl = reinterpret_cast<const float*>(smem + OFF_DELTA);
  const uint32_t spt = smem_u32(smem + OFF_PT), sdst = smem_u32(smem + OFF_DST);
#pragma unroll
  for (int c = 0; c < G / 8; ++c) {
    const int q = q0 + 8 * c, cl = cl0 + 8 * c;
    uint32_t pw[4] = {0u, 0u, 0u, 0u}, dw[4] = {0u, 0u, 0u, 0u};
    if (kv && q < p.N) {
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
      if (btrow) {  // packed: (column group, row = key) float4s, pre-multiplied by log2(e)
        b0 = __ldg(btrow + (q >> 2) * 256);
        b1 = __ldg(btrow + ((q >> 2) + 1) * 256);
      }
      const float4 n0 = *reinterpret_cast<const float4*>(nl + q), n1 = *reinterpret_cast<const float4*>(nl + q + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(dl + q), d1 = *reinterpret_cast<const float4*>(dl + q + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float nn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
      const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      float pe[8], de[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        pe[j] = ex2(fmaf(s[8 * c + j], p.c1, bb[j] + nn[j]));  // nn = -inf for q >= N -> p = 0
        de[j] = pe[j] * (dp[8 * c + j] - dd[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pw[j] = pack_bf16(pe[2 * j], pe[2 * j + 1]);
        dw[j] = pack_bf16(de[2 * j], de[2 * j + 1]);
      }
    }
    const int blk = cl >> 6, chunk = (cl & 63) >> 3;
    st_shared_v4(sw128(spt + blk * 16384, rl, chunk), pw[0], pw[1], pw[2], pw[3]);
    st_shared_v4(sw128(sdst + blk * 16384, rl, chunk), dw[0], dw[1], dw[2], dw[3]);
  }
}

// 64 fp32 accumulator columns of this thread's TMEM lane -> bf16 row (128 B) in global memory
__device__ __forceinline__ void store_acc_row(uint32_t taddr, float mul, bf16* dst, bool valid) {
  float v[64];
  tmem_ld32f(taddr, v);
  tmem_ld32f(taddr + 32, v + 32);
  tmem_ld_wait();
  if (valid) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      reinterpret_cast<uint4*>(dst)[i] =
          make_uint4(pack_bf16(v[8 * i] * mul, v[8 * i + 1] * mul), pack_bf16(v[8 * i + 2] * mul, v[8 * i + 3] * mul),
                     pack_bf16(v[8 * i + 4] * mul, v[8 * i + 5] * mul), pack_bf16(v[8 * i + 6] * mul, v[8 * i + 7] * mul));
  }
}

__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_tc(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                 const __grid_constant__ CUtensorMap tmap_ds, const BwdParams p) {
  using namespace bwd;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int nt = p.N > 128 ? 2 : 1;  // key tiles == query chunks in use

  if (warp == 8) {
    if (lane == 0) {
      prefetch_tmap(&tmap_qkv); prefetch_tmap(&tmap_do); prefetch_tmap(&tmap_ds);
      mbar_init(&bar[B_LOAD], 1);
      mbar_init(&bar[B_S], 1);
      mbar_init(&bar[B_PD], 256);
      mbar_init(&bar[B_MMA2], 1);
      mbar_init(&bar[B_ACCFREE], 256);
      fence_barrier_init();
      mbar_arrive_expect_tx(&bar[B_LOAD], 4 * kLoadBytes);
      tma_load_2d(smem + OFF_Q, &tmap_qkv, &bar[B_LOAD], h * kHeadDim, b * p.N);
      tma_load_2d(smem + OFF_K, &tmap_qkv, &bar[B_LOAD], (p.H + h) * kHeadDim, b * p.N);
      tma_load_2d(smem + OFF_V, &tmap_qkv, &bar[B_LOAD], (2 * p.H + h) * kHeadDim, b * p.N);
      tma_load_2d(smem + OFF_DO, &tmap_do, &bar[B_LOAD], h * kHeadDim, b * p.N);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  } else {
    // delta[q] = sum_d dO[q,d] * O[q,d] (straight from global), -lse*log2e -> smem
    float* nl = reinterpret_cast<float*>(smem + OFF_NL);
    float* dl = reinterpret_cast<float*>(smem + OFF_DELTA);
    const long long os = (long long)p.H * kHeadDim;
    for (int q = threadIdx.x; q < kMaxN; q += 256) {
      float d = 0.f, l = -INFINITY;
      if (q < p.N) {
        const uint4* o4 = reinterpret_cast<const uint4*>(p.out + ((long long)b * p.N + q) * os + h * kHeadDim);
        const uint4* g4 = reinterpret_cast<const uint4*>(p.dout + ((long long)b * p.N + q) * os + h * kHeadDim);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 ov = __ldg(o4 + i), gv = __ldg(g4 + i);
          const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov);
          const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = __bfloat1622float2(oh[j]), g = __bfloat1622float2(gh[j]);
            d = fmaf(a.x, g.x, fmaf(a.y, g.y, d));
          }
        }
        l = -p.lse[((long long)b * p.H + h) * p.N + q] * kLog2e;
      }
      dl[q] = d;
      nl[q] = l;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      // -------------------------------------------------------------------------- control thread
      const uint32_t sq = smem_u32(smem + OFF_Q), sdo = smem_u32(smem + OFF_DO), sk = smem_u32(smem + OFF_K);
      const uint32_t sv = smem_u32(smem + OFF_V), spt = smem_u32(smem + OFF_PT), sdst = smem_u32(smem + OFF_DST);
      constexpr uint32_t idesc_s128 = make_idesc(1, false, false, 128, 128);
      constexpr uint32_t idesc_s80 = make_idesc(1, false, false, 128, 80);
      constexpr uint32_t idesc_kv = make_idesc(1, false, true, 128, kHeadDim);  // A K-major (P^T / dS^T), B MN-major
      constexpr uint32_t idesc_dq = make_idesc(1, true, true, 128, kHeadDim);   // A MN-major (dS), B MN-major (K)
      mbar_wait(&bar[B_LOAD], 0, nullptr, 1);
      tc_fence_after();
      uint32_t it = 0;
      for (int t = 0; t < nt; ++t) {
        for (int c = 0; c < nt; ++c, ++it) {
          const uint32_t ph = it & 1;
          // ---- S^T = K_t Q_c^T, dP^T = V_t dO_c^T
          const uint32_t idesc_s = c ? idesc_s80 : idesc_s128;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_base + COL_ST, make_smem_desc_sw128(sk + t * 16384 + ks * 32, 0, 1024),
                      make_smem_desc_sw128(sq + c * 16384 + ks * 32, 0, 1024), idesc_s, ks != 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_base + COL_DPT, make_smem_desc_sw128(sv + t * 16384 + ks * 32, 0, 1024),
                      make_smem_desc_sw128(sdo + c * 16384 + ks * 32, 0, 1024), idesc_s, ks != 0);
          umma_commit(&bar[B_S]);
          mbar_wait(&bar[B_PD], ph, nullptr, 2);  // P^T, dS^T tiles written; S^T, dP^T consumed
          tc_fence_after();
          if (t == 1 && c == 0) mbar_wait(&bar[B_ACCFREE], 0, nullptr, 3);  // dV_0 / dK_0 read out
          // ---- dV_t += P^T dO_c, dK_t += dS^T Q_c   (reduction over the chunk's queries)
          const int nq = c ? 5 : 8;
          for (int ks = 0; ks < nq; ++ks) {
            const uint32_t aoff = (ks >> 2) * 16384 + (ks & 3) * 32, boff = (c * 128 + ks * 16) * 128;
            umma_bf16(tmem_base + COL_DV, make_smem_desc_sw128(spt + aoff, 0, 1024),
                      make_smem_desc_sw128(sdo + boff, 8192, 1024), idesc_kv, (c | ks) != 0);
            umma_bf16(tmem_base + COL_DK, make_smem_desc_sw128(sdst + aoff, 0, 1024),
                      make_smem_desc_sw128(sq + boff, 8192, 1024), idesc_kv, (c | ks) != 0);
          }
          // ---- dQ_c += dS K_t   (reduction over the tile's keys; A = dS read MN-major out of the dS^T tile)
          const int nk = t ? 5 : 8;
          for (int ks = 0; ks < nk; ++ks)
            umma_bf16(tmem_base + COL_DQ + 64 * c, make_smem_desc_sw128(sdst + ks * 2048, 16384, 1024),
                      make_smem_desc_sw128(sk + (t * 128 + ks * 16) * 128, 8192, 1024), idesc_dq, (t | ks) != 0);
          umma_commit(&bar[B_MMA2]);
          if (p.write_ds) {  // dS^T[b, h, keys of tile t, queries of chunk c]  (rows >= N, columns >= ldb clipped)
            for (int j = 0; j < 2; ++j)
              if (c * 128 + j * 64 < p.ldb) tma_store_3d(&tmap_ds, sdst + j * 16384, c * 128 + j * 64, t * 128, blockIdx.x);
            bulk_commit();
            bulk_wait_read<0>();  // the next S commit (below) then also covers "dS^T tile is reusable"
          }
        }
      }
      if (p.write_ds) bulk_wait<0>();
    }
  } else {
    // ------------------------------------------------------------------------------ elementwise / epilogue warps
    const int quarter = warp & 3, half = warp >> 2;
    const int rl = quarter * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const long long rs = 3LL * p.H * kHeadDim;
    bf16* dq_base = p.dqkv + (long long)b * p.N * rs + h * kHeadDim;
    uint32_t it = 0;
    for (int t = 0; t < nt; ++t) {
      const int key = t * 128 + rl;
      const bool kv = key < p.N;
      const float4* btrow = (p.biasT && kv) ? reinterpret_cast<const float4*>(p.biasT) + (long long)h * kBiasGroups * 256 + key
                                            : nullptr;
      for (int c = 0; c < nt; ++c, ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(&bar[B_S], ph, nullptr, 4);
        tc_fence_after();
        if (c == 0) {  // 128 queries: 64 per thread
          bwd_group<32>(p, smem, tlane + half * 64, btrow, kv, half * 64, half * 64, rl);
          bwd_group<32>(p, smem, tlane + half * 64 + 32, btrow, kv, half * 64 + 32, half * 64 + 32, rl);
        } else {       // 80 queries: 40 per thread
          bwd_group<32>(p, smem, tlane + half * 40, btrow, kv, 128 + half * 40, half * 40, rl);
          bwd_group<8>(p, smem, tlane + half * 40 + 32, btrow, kv, 128 + half * 40 + 32, half * 40 + 32, rl);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&bar[B_PD]);
        mbar_wait(&bar[B_MMA2], ph, nullptr, 5);
        tc_fence_after();
        if (c == nt - 1) {  // dV_t (half 0) / dK_t (half 1) complete
          bf16* dst = dq_base + (long long)key * rs + (half ? 1 : 2) * p.H * kHeadDim;
          store_acc_row(tlane + (half ? COL_DK : COL_DV), half ? p.scale : 1.f, dst, kv);
          tc_fence_before();
          mbar_arrive(&bar[B_ACCFREE]);
        }
      }
    }
    // dQ: query tile `half`
    if (half < nt) {
      const int q = half * 128 + rl;
      store_acc_row(tlane + COL_DQ + 64 * half, p.scale, dq_base + (long long)q * rs, q < p.N);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ host side
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
// bf16 tensor, dims[0] innermost (dense), 128B swizzle, box[0] = 64 elements
static int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const long long* dims, const long long* strides_elems,
                          const int* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled is unavailable");
  MEMB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, "attention: tensors must be 16-byte aligned");
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bdim[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bdim[i] = (cuuint32_t)box[i];
    if (i > 0) {
      MEMB_REQUIRE((strides_elems[i - 1] * 2) % 16 == 0, "attention: row strides must be multiples of 16 bytes");
      gstr[i - 1] = (cuuint64_t)strides_elems[i - 1] * 2;
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled (attention) failed with CUresult %d", (int)r);
  return MEMB_OK;
}

}  // namespace attn
}  // namespace memb

using namespace memb;
using namespace memb::attn;

static int check_shape(int B, int N, int H, int head_dim, int ldb, bool has_bias) {
  MEMB_REQUIRE(B > 0 && H > 0 && N > 0 && N <= kMaxN, "attention: N must be in [1, %d], got %d", kMaxN, N);
  MEMB_REQUIRE(head_dim == kHeadDim, "attention: head dim must be %d, got %d", kHeadDim, head_dim);
  MEMB_REQUIRE(!has_bias || (ldb % 8 == 0 && ldb >= N), "attention: dS^T row stride must be a multiple of 8 and >= N");
  return MEMB_OK;
}

extern "C" int memb_attention_fwd(const void* qkv, const float* bias, int ldb, int B, int N, int H, int head_dim,
                                  float scale, void* out, float* lse, memb_stream_t s) {
  if (int rc = check_shape(B, N, H, head_dim, ldb, false)) return rc;
  MEMB_REQUIRE(qkv && out && lse, "attention_fwd: null pointer");
  MEMB_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15u) == 0, "attention_fwd: bias must be 16-byte aligned");
  CUtensorMap tq;
  const long long dims[2] = {3LL * H * kHeadDim, (long long)B * N}, str[1] = {3LL * H * kHeadDim};
  const int box[2] = {64, kMaxN};
  if (int rc = make_tmap_bf16(&tq, qkv, 2, dims, str, box)) return rc;
  static bool configured = false;
  if (!configured) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(attention_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::SMEM_BYTES));
    configured = true;
  }
  FwdParams p{bias, ldb, B, N, H, scale * kLog2e, (bf16*)out, lse};
  const int grid = std::min(B * H, num_sms());
  attention_fwd_tc<<<grid, kThreads, fwd::SMEM_BYTES, s>>>(tq, p);
  MEMB_LAUNCH_OK("attention_fwd_tc");
  return MEMB_OK;
}

extern "C" int memb_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const float* bias,
                                  const float* biasT, int ldb, int B, int N, int H, int head_dim, float scale, void* dqkv,
                                  void* dsT, memb_stream_t s) {
  (void)bias;  // the backward reads the transposed copy only
  if (int rc = check_shape(B, N, H, head_dim, ldb, dsT != nullptr)) return rc;
  MEMB_REQUIRE(qkv && out && dout && lse && dqkv, "attention_bwd: null pointer");
  MEMB_REQUIRE((bias == nullptr) == (biasT == nullptr), "attention_bwd: bias and its transpose go together");
  MEMB_REQUIRE(!biasT || (reinterpret_cast<uintptr_t>(biasT) & 15u) == 0, "attention_bwd: biasT must be 16-byte aligned");
  CUtensorMap tq, tdo, tds;
  {
    const long long dims[2] = {3LL * H * kHeadDim, (long long)B * N}, str[1] = {3LL * H * kHeadDim};
    const int box[2] = {64, kMaxN};
    if (int rc = make_tmap_bf16(&tq, qkv, 2, dims, str, box)) return rc;
  }
  {
    const long long dims[2] = {(long long)H * kHeadDim, (long long)B * N}, str[1] = {(long long)H * kHeadDim};
    const int box[2] = {64, kMaxN};
    if (int rc = make_tmap_bf16(&tdo, dout, 2, dims, str, box)) return rc;
  }
  if (dsT) {
    const long long dims[3] = {ldb, N, (long long)B * H}, str[2] = {ldb, (long long)N * ldb};
    const int box[3] = {64, 128, 1};
    if (int rc = make_tmap_bf16(&tds, dsT, 3, dims, str, box)) return rc;
  } else {
    tds = tdo;  // never dereferenced
  }
  static bool configured = false;
  if (!configured) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(attention_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::SMEM_BYTES));
    configured = true;
  }
  BwdParams p{(const bf16*)out, (const bf16*)dout, lse, biasT, ldb, B, N, H, scale, scale * kLog2e, (bf16*)dqkv, dsT ? 1 : 0};
  attention_bwd_tc<<<B * H, kThreads, bwd::SMEM_BYTES, s>>>(tq, tdo, tds, p);
  MEMB_LAUNCH_OK("attention_bwd_tc");
  return MEMB_OK;
}
