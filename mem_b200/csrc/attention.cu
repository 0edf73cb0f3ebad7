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
// Forward (persistent, one CTA per SM, 416 threads):
//   warp 12  control: TMA loads (Q, K, V of the NEXT head are fetched while the current one is in softmax / PV),
//            tcgen05.mma issue: S_p = Q_p K^T (M=128 query tile p, N=208 keys, fp32 in TMEM), O_p = P_p V
//   warps 0-11 softmax: three warps per TMEM lane quarter, each thread owns a third of a row (72/64 keys) in
//            registers, preloaded with the bias: x = s*scale*log2e + bias*log2e, row max exchanged through smem,
//            p = 2^(x-m) -> bf16 P tile in swizzled smem (the A operand of the PV MMA), then the O epilogue
//            (1/l, bf16, global) and lse.
// Backward (one CTA per (b, h), 576 threads: 16 elementwise warps, one warp issuing batch 1, one issuing batch 2), keys on the M axis so that P^T / dS^T come out of TMEM in the
// orientation dV = P^T dO, dK = dS^T Q and dQ = dS K need; per key tile t the queries go by in chunks of 64:
//   batch 1 (chunk g):  S^T = K_t Q_g^T, dP^T = V_t dO_g^T                    (TMEM buffer g & 1)
//   elementwise:        p = 2^(s*c1 + b - lse2), ds = p (dp - delta) -> bf16 P^T (slot g & 1), dS^T (4 slots) in smem
//   batch 2 (chunk g):  dV_t += P^T dO_g, dK_t += dS^T Q_g; after a pair of chunks dQ_pair += dS K_t with dS read
//                       MN-major straight out of the two adjacent dS^T slots; dS^T also leaves by TMA store.
//   Batch 1 of chunk g+2 is issued before batch 2 of chunk g, so the tensor pipe works in the shadow of the
//   elementwise warps.  TMEM: 2 x (S^T 64 + dP^T 64) + dV 64 + dK 64 + dQ 2x64 = 512 columns.
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
constexpr int kFwdThreads = 512;              // 12 softmax warps + a control warpgroup (1 working warp)
constexpr int kBwdThreads = 576;              // 16 elementwise warps + 2 MMA-issuing warps
constexpr int kBwdEw = 512;                   // elementwise threads
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
__device__ __forceinline__ float4 ldg_f4(const float4* p) {  // volatile: stays where it is written (issued early)
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
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
constexpr int OFF_MAX = OFF_P + 2 * P_TILE;    // float [2 tiles][<=4 column parts][128]
constexpr int OFF_SUM = OFF_MAX + 4096;        // float [2 tiles][<=4 column parts][128]
constexpr int OFF_BAR = OFF_SUM + 4096;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;  // + alignment slack
constexpr int kSoftmaxThreads = 384;
enum { B_QK = 0, B_V, B_S0, B_S1, B_P0, B_P1, B_O0, B_O1, B_OFREE0, B_OFREE1, B_COUNT };
__host__ __device__ constexpr int s_col(int t) { return t * 256; }  // TMEM columns of S_p; O_p reuses the first 64
}  // namespace fwd

struct FwdParams {
  const float* bias;
  int ldb, B, N, H;
  float c1;  // scale * log2(e)
  bf16* out;
  float* lse;
};

// Column third ct of a 208-key row: [0,72) [72,136) [136,208)  -> 9 / 8 / 9 chunks of 8
__device__ __forceinline__ int ct_start(int ct) { return ct == 0 ? 0 : (ct == 1 ? 72 : 136); }

// x <- this thread's bias values (times log2 e; -inf past N).  Issued BEFORE the wait for S so that the L2
// latency hides behind the MMA / the previous epilogue.
template <int TILE>
__device__ __forceinline__ void fwd_load_bias(const FwdParams& p, int h, int quarter, int ct, int lane, float (&x)[72]) {
  const int row = TILE * 128 + quarter * 32 + lane;
  const int c0 = ct_start(ct), ng = ct == 1 ? 16 : 18;
  if (p.bias && row < p.N) {
    // packed bias: float4 (h, column group g, row) -> the lanes of a warp read 512 contiguous bytes
    const float4* bp = reinterpret_cast<const float4*>(p.bias) + ((long long)h * kBiasGroups + (c0 >> 2)) * 256 + row;
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      if (i < ng) {
        const float4 b = ldg_f4(bp + i * 256);
        x[4 * i] = b.x; x[4 * i + 1] = b.y; x[4 * i + 2] = b.z; x[4 * i + 3] = b.w;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 72; ++i) x[i] = (c0 + i < p.N) ? 0.f : -INFINITY;
  }
}

template <int TILE>
__device__ __forceinline__ void fwd_softmax_tile(const FwdParams& p, uint8_t* smem, uint32_t tmem_base, int quarter, int ct,
                                                 int lane, float (&x)[72], float& m_out) {
  using namespace fwd;
  const int rl = quarter * 32 + lane;          // row within the tile == TMEM lane
  const int row = TILE * 128 + rl;             // query index
  const bool rv = row < p.N;
  const int c0 = ct_start(ct), nc8 = ct == 1 ? 8 : 9;
  float* smax = reinterpret_cast<float*>(smem + OFF_MAX) + TILE * 512;
  float* ssum = reinterpret_cast<float*>(smem + OFF_SUM) + TILE * 512;
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + s_col(TILE) + c0;
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) {  // 8 columns at a time keeps the live set small
    if (k < nc8) {
      uint32_t r[8];
      tmem_ld8(taddr + 8 * k, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[8 * k + i] = fmaf(__uint_as_float(r[i]), p.c1, x[8 * k + i]);
        mx = fmaxf(mx, x[8 * k + i]);
      }
    }
  }
  if (!rv) mx = -INFINITY;  // stale rows may hold NaN
  smax[ct * 128 + rl] = mx;
  named_bar_sync(1, kSoftmaxThreads);
  const float m = fmaxf(fmaxf(smax[rl], smax[128 + rl]), smax[256 + rl]);
  m_out = m;
  if (rv) {
    float sum = 0.f;
    const uint32_t sp = smem_u32(smem + OFF_P) + TILE * P_TILE;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      if (i < nc8) {
        float e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          e[j] = ex2(x[8 * i + j] - m);
          sum += e[j];
        }
        const int c = c0 + 8 * i;
        st_shared_v4(sw128(sp + (c >> 6) * 16384, rl, (c & 63) >> 3), pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]),
                     pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
      }
    }
    ssum[ct * 128 + rl] = sum;
  }
}

// O epilogue: column third ct takes head-dim columns [0,24) [24,48) [48,64)
template <int TILE>
__device__ __forceinline__ void fwd_out_tile(const FwdParams& p, uint8_t* smem, uint32_t tmem_base, int b, int h, int quarter,
                                             int ct, int lane, float m) {
  using namespace fwd;
  const int rl = quarter * 32 + lane, row = TILE * 128 + rl;
  uint32_t o[24];
  uint32_t (&o0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&o[0]);
  uint32_t (&o1)[8] = *reinterpret_cast<uint32_t(*)[8]>(&o[16]);
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + s_col(TILE) + ct * 24;
  tmem_ld16(taddr, o0);
  if (ct < 2) tmem_ld8(taddr + 16, o1);
  tmem_ld_wait();
  if (row < p.N) {
    const float* ssum = reinterpret_cast<const float*>(smem + OFF_SUM) + TILE * 512;
    const float l = ssum[rl] + ssum[128 + rl] + ssum[256 + rl];
    const float inv = 1.f / l;
    bf16* dst = p.out + ((long long)b * p.N + row) * p.H * kHeadDim + h * kHeadDim + ct * 24;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i < 2 || ct < 2)
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(
            pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
            pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
            pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
            pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
    if (ct == 0) p.lse[((long long)b * p.H + h) * p.N + row] = (m + lg2(l)) * kLn2;
  }
}

__global__ void __launch_bounds__(kFwdThreads, 1)
attention_fwd_tc(const __grid_constant__ CUtensorMap tmap_qkv, const FwdParams p) {
  using namespace fwd;
  grid_dependency_trigger();   // the proj GEMM (a programmatic dependent) sets itself up under this kernel's tail
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nheads = p.B * p.H;
  const bool two = p.N > 128;  // second query tile in use

  if (warp == 12) {
    if (lane == 0) {
      prefetch_tmap(&tmap_qkv);
      mbar_init(&bar[B_QK], 1); mbar_init(&bar[B_V], 1);
      mbar_init(&bar[B_S0], 1); mbar_init(&bar[B_S1], 1);
      mbar_init(&bar[B_P0], kSoftmaxThreads); mbar_init(&bar[B_P1], kSoftmaxThreads);
      mbar_init(&bar[B_O0], 1); mbar_init(&bar[B_O1], 1);
      mbar_init(&bar[B_OFREE0], kSoftmaxThreads); mbar_init(&bar[B_OFREE1], kSoftmaxThreads);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");  // the control warpgroup hands its registers to the softmax warps
    if (warp == 12 && lane == 0) {
      // -------------------------------------------------------------------------- control thread
      // (a single-thread branch: the `if (elect_one())` form makes ptxas 12.9 predicate these UTCHMMAs, the shape it
      //  mis-schedules; the backward issuers, whose MMAs it leaves unpredicated, do use it)
      const uint32_t sq = smem_u32(smem + OFF_Q), sk = smem_u32(smem + OFF_K), sv = smem_u32(smem + OFF_V);
      const uint32_t spp = smem_u32(smem + OFF_P);
      constexpr uint32_t idesc_s = make_idesc(1, false, false, 128, kMaxN);
      constexpr uint32_t idesc_o = make_idesc(1, false, true, 128, kHeadDim);
      const uint64_t dq0 = make_smem_desc_sw128(sq, 0, 1024), dk0 = make_smem_desc_sw128(sk, 0, 1024);
      const uint64_t dp0 = make_smem_desc_sw128(spp, 0, 1024), dv0 = make_smem_desc_sw128(sv, 8192, 1024);
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
        const bool first = head == (int)blockIdx.x;
        const int next = head + gridDim.x;
        mbar_wait(&bar[B_QK], ph, nullptr, 1);
        // Both query tiles are always issued (tile 1 of a short sequence computes rows nobody reads): ptxas 12.9
        // mis-schedules the descriptor moves of a predicated UTCHMMA, so the MMAs below carry no predicate.
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (!first) mbar_wait(&bar[B_OFREE0 + t], ph ^ 1, nullptr, 2);  // O_t of the previous head has been read
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)  // descriptor address field counts 16-byte units
            umma_bf16(tmem_base + s_col(t), dq0 + (uint64_t)(t * 1024 + ks * 2), dk0 + (uint64_t)(ks * 2), idesc_s, ks != 0);
          umma_commit(&bar[B_S0 + t]);
        }
        mbar_wait(&bar[B_S1], ph, nullptr, 3);  // both S tiles done: Q and K are free
        if (next < nheads) load_qk(next);
        mbar_wait(&bar[B_V], ph, nullptr, 4);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&bar[B_P0 + t], ph, nullptr, 5);
          tc_fence_after();
          uint64_t da = dp0 + (uint64_t)(t * (P_TILE >> 4)), db = dv0;
#pragma unroll 1
          for (int ks = 0; ks < nks; ++ks) {  // rolled on purpose: the control thread lives on few registers
            umma_bf16(tmem_base + s_col(t), da, db, idesc_o, ks != 0);
            da += ((ks & 3) == 3) ? (uint64_t)(1024 - 6) : (uint64_t)2;  // next 32 B, or next 64-key block
            db += 128;
          }
          umma_commit(&bar[B_O0 + t]);
        }
        mbar_wait(&bar[B_O1], ph, nullptr, 6);  // V is free
        if (next < nheads) load_v(next);
      }
    }
  } else {
    // ------------------------------------------------------------------------------ softmax / epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int quarter = warp & 3, cq = warp >> 2;  // cq: column third of the row
    uint32_t ph = 0;
    for (int head = blockIdx.x; head < nheads; head += gridDim.x, ph ^= 1) {
      const int b = head / p.H, h = head % p.H;
      float m0 = 0.f, m1 = 0.f;
      float x[72];
      fwd_load_bias<0>(p, h, quarter, cq, lane, x);
      mbar_wait(&bar[B_S0], ph, nullptr, 7);
      tc_fence_after();
      fwd_softmax_tile<0>(p, smem, tmem_base, quarter, cq, lane, x, m0);
      if (two) fwd_load_bias<1>(p, h, quarter, cq, lane, x);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar[B_P0]);
      mbar_wait(&bar[B_S1], ph, nullptr, 8);
      tc_fence_after();
      if (two) fwd_softmax_tile<1>(p, smem, tmem_base, quarter, cq, lane, x, m1);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar[B_P1]);
      mbar_wait(&bar[B_O0], ph, nullptr, 9);
      tc_fence_after();
      fwd_out_tile<0>(p, smem, tmem_base, b, h, quarter, cq, lane, m0);
      tc_fence_before();
      mbar_arrive(&bar[B_OFREE0]);
      mbar_wait(&bar[B_O1], ph, nullptr, 10);
      tc_fence_after();
      if (two) fwd_out_tile<1>(p, smem, tmem_base, b, h, quarter, cq, lane, m1);
      tc_fence_before();
      mbar_arrive(&bar[B_OFREE1]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =========================================================================================== backward
#ifdef MEMB_ATTN_TRACE
// debug timeline of one CTA: three writers (elementwise thread 0, issuer 1, issuer 2), 512 slots each, no atomics
__device__ long long g_trace[3 * 512 * 2];
#define TRACE_DECL(role) int trace_i_ = (role) * 512
#define TRACE(id) do { if (blockIdx.x == MEMB_ATTN_TRACE) { g_trace[2 * trace_i_] = (id); g_trace[2 * trace_i_ + 1] = clock64(); ++trace_i_; } } while (0)
#else
#define TRACE_DECL(role) do {} while (0)
#define TRACE(id) do {} while (0)
#endif
namespace bwd {
constexpr int OFF_Q = 0;                        // [208][128 B]
constexpr int OFF_DO = kLoadBytes;              // [208][128 B]
constexpr int OFF_K = 2 * kLoadBytes;           // [208][128 B]; tile 1 as an A operand runs 48 rows past the end into
constexpr int OFF_V = 3 * kLoadBytes;           // the next buffer: those rows only feed S^T / dP^T rows that nobody reads
constexpr int OFF_PT = 4 * kLoadBytes;          // 2 slots x [128 keys][64 queries] bf16
constexpr int OFF_DST = OFF_PT + 2 * 16384;     // 4 slots (2 pairs) of the same shape: dS^T
constexpr int OFF_NL = OFF_DST + 4 * 16384;     // float[256]: -lse*log2e  (-inf for q >= N)
constexpr int OFF_DELTA = OFF_NL + 1024;        // float[256]: rowsum(dO * O)
constexpr int OFF_BAR = OFF_DELTA + 1024;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
enum { B_LOAD = 0, B_S0, B_S1, B_PD0, B_PD1, B_B20, B_B21, B_ACC, B_ACCFREE, B_COUNT };
constexpr int COL_DV = 256, COL_DK = 320, COL_DQ = 384;  // S^T at 128*buf, dP^T at 128*buf + 64
}  // namespace bwd

struct BwdParams {
  const float* lse;       // [B*H*N] natural-log row logsumexp of the forward
  const float* delta;     // [B*H*N] rowsum(dO*O): attn_rowdot_heads, or the producer of dO (memb_gemm MEMB_EPI_STORE_ROWDOT)
  const float* biasT;
  int ldb, B, N, H;
  float scale, c1;
  bf16* dqkv;
  int write_ds;
};

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]: eight lanes per (b, q, h) row of 64, one 16-byte load of O and of dO
// each (a warp reads 512 contiguous bytes per instruction), 3 shuffles, lane 0 writes.
__global__ void __launch_bounds__(256) attn_rowdot_heads(const bf16* __restrict__ out, const bf16* __restrict__ dout,
                                                         int B, int N, int H, float* __restrict__ delta) {
  static_assert(kHeadDim == 64, "eight 16-byte loads cover one head row");
  const long long total = (long long)B * N * H;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long i = t >> 3;                                            // (b, q, h), h fastest
  const int part = (int)(t & 7);
  float d = 0.f;
  if (i < total) {
    const uint4 ov = __ldcs(reinterpret_cast<const uint4*>(out + i * kHeadDim) + part);
    const uint4 gv = __ldcs(reinterpret_cast<const uint4*>(dout + i * kHeadDim) + part);
    const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov);
    const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __bfloat1622float2(oh[j]), g = __bfloat1622float2(gh[j]);
      d = fmaf(a.x, g.x, fmaf(a.y, g.y, d));
    }
  }
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 4);
  if (i < total && part == 0) {
    const unsigned int iu = (unsigned int)i;                           // B*N*H < 2^31 (checked by the launcher)
    const unsigned int bq = iu / (unsigned int)H, h = iu - bq * H;
    const unsigned int b = bq / (unsigned int)N, q = bq - b * N;
    const long long o = ((long long)b * H + h) * N + q;
    delta[o] = d;
  }
}

// G (16 or 8) query columns of this thread's key row: S^T, dP^T -> P^T, dS^T (bf16, swizzled smem slots).
template <int G>
__device__ __forceinline__ void bwd_group(const BwdParams& p, uint8_t* smem, uint32_t taddr, const float4 (&breg)[4], bool kv,
                                          int q0 /* global query index */, int cl0 /* column within the chunk */, int rl,
                                          uint32_t spt, uint32_t sdst) {
  using namespace bwd;
  uint32_t sr[G], dr[G];
  if constexpr (G == 16) {
    tmem_ld16(taddr, sr);
    tmem_ld16(taddr + 64, dr);
  } else {
    tmem_ld8(taddr, sr);
    tmem_ld8(taddr + 64, dr);
  }
  tmem_ld_wait();
  const float* nl = reinterpret_cast<const float*>(smem + OFF_NL);
  const float* dl = reinterpret_cast<const float*>(smem + OFF_DELTA);
#pragma unroll
  for (int c = 0; c < G / 8; ++c) {
    const int q = q0 + 8 * c, cl = cl0 + 8 * c;
    uint32_t pw[4] = {0u, 0u, 0u, 0u}, dw[4] = {0u, 0u, 0u, 0u};
    if (kv) {
      const float4 b0 = breg[2 * c], b1 = breg[2 * c + 1];
      const float4 n0 = *reinterpret_cast<const float4*>(nl + q), n1 = *reinterpret_cast<const float4*>(nl + q + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(dl + q), d1 = *reinterpret_cast<const float4*>(dl + q + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float nn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
      const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      float pe[8], de[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        pe[j] = ex2(fmaf(__uint_as_float(sr[8 * c + j]), p.c1, bb[j] + nn[j]));  // nn = -inf for q >= N -> p = 0
        de[j] = pe[j] * (__uint_as_float(dr[8 * c + j]) - dd[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pw[j] = pack_bf16(pe[2 * j], pe[2 * j + 1]);
        dw[j] = pack_bf16(de[2 * j], de[2 * j + 1]);
      }
    }
    st_shared_v4(sw128(spt, rl, cl >> 3), pw[0], pw[1], pw[2], pw[3]);
    st_shared_v4(sw128(sdst, rl, cl >> 3), dw[0], dw[1], dw[2], dw[3]);
  }
}

// 32 fp32 accumulator columns of this thread's TMEM lane -> bf16, four 16-byte chunks of row `rl` of a swizzled
// [128][64] staging tile (the layout a SWIZZLE_128B TMA store expects)
__device__ __forceinline__ void stage_acc32(uint32_t taddr, float mul, uint32_t stage, int rl, int chunk0) {
  float v[32];
  tmem_ld32f(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i)
    st_shared_v4(sw128(stage, rl, chunk0 + i), pack_bf16(v[8 * i] * mul, v[8 * i + 1] * mul),
                 pack_bf16(v[8 * i + 2] * mul, v[8 * i + 3] * mul), pack_bf16(v[8 * i + 4] * mul, v[8 * i + 5] * mul),
                 pack_bf16(v[8 * i + 6] * mul, v[8 * i + 7] * mul));
}

// NQ k-steps of 16 queries: dV_t += P^T dO_j, dK_t += dS^T Q_j.  Descriptor address fields count 16-byte units.
template <int NQ>
__device__ __forceinline__ void issue_dv_dk(uint32_t tmem_base, uint64_t dpt, uint64_t ddst, uint64_t ddo, uint64_t dq,
                                            uint32_t idesc, bool first_chunk) {
#pragma unroll
  for (int ks = 0; ks < NQ; ++ks) {
    umma_bf16(tmem_base + bwd::COL_DV, dpt + (uint64_t)(ks * 2), ddo + (uint64_t)(ks * 128), idesc, !(first_chunk && ks == 0));
    umma_bf16(tmem_base + bwd::COL_DK, ddst + (uint64_t)(ks * 2), dq + (uint64_t)(ks * 128), idesc, !(first_chunk && ks == 0));
  }
}
template <int NK>
__device__ __forceinline__ void issue_dq(uint32_t tmem_d, uint64_t dds, uint64_t dk, uint32_t idesc, bool first_tile) {
#pragma unroll
  for (int ks = 0; ks < NK; ++ks)
    umma_bf16(tmem_d, dds + (uint64_t)(ks * 128), dk + (uint64_t)(ks * 128), idesc, !(first_tile && ks == 0));
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attention_bwd_tc(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                 const __grid_constant__ CUtensorMap tmap_ds, const __grid_constant__ CUtensorMap tmap_dqkv, const BwdParams p) {
  using namespace bwd;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int nt = p.N > 128 ? 2 : 1;               // key tiles
  const int nch = min(4, (p.N + 63) / 64);        // query chunks (64 wide; the 4th is 16 wide)
  const int npair = (nch + 1) >> 1;

  if (warp == 16) {
    if (lane == 0) {
      prefetch_tmap(&tmap_qkv); prefetch_tmap(&tmap_do); prefetch_tmap(&tmap_ds); prefetch_tmap(&tmap_dqkv);
      mbar_init(&bar[B_LOAD], 1);
      mbar_init(&bar[B_S0], 1); mbar_init(&bar[B_S1], 1);
      mbar_init(&bar[B_PD0], kBwdEw / 2); mbar_init(&bar[B_PD1], kBwdEw / 2);
      mbar_init(&bar[B_B20], 1); mbar_init(&bar[B_B21], 1);
      mbar_init(&bar[B_ACC], 1);
      mbar_init(&bar[B_ACCFREE], kBwdEw);
      fence_barrier_init();
      mbar_arrive_expect_tx(&bar[B_LOAD], 4 * kLoadBytes);
      tma_load_2d(smem + OFF_K, &tmap_qkv, &bar[B_LOAD], (p.H + h) * kHeadDim, b * p.N);
      tma_load_2d(smem + OFF_Q, &tmap_qkv, &bar[B_LOAD], h * kHeadDim, b * p.N);
      tma_load_2d(smem + OFF_V, &tmap_qkv, &bar[B_LOAD], (2 * p.H + h) * kHeadDim, b * p.N);
      tma_load_2d(smem + OFF_DO, &tmap_do, &bar[B_LOAD], h * kHeadDim, b * p.N);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  } else if (warp < 8) {  // 256 threads fill the two 256-entry vectors
    float* nl = reinterpret_cast<float*>(smem + OFF_NL);
    float* dl = reinterpret_cast<float*>(smem + OFF_DELTA);
    const int q = threadIdx.x;
    const long long o = (long long)blockIdx.x * p.N + q;
    nl[q] = q < p.N ? -__ldg(p.lse + o) * kLog2e : -INFINITY;
    dl[q] = q < p.N ? __ldg(p.delta + o) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sq = smem_u32(smem + OFF_Q), sdo = smem_u32(smem + OFF_DO), sk = smem_u32(smem + OFF_K);
  const uint32_t sv = smem_u32(smem + OFF_V), spt = smem_u32(smem + OFF_PT), sdst = smem_u32(smem + OFF_DST);

  if (warp == 16) {
    // ---------------------------------------------------- issuer 1: S^T = K_t Q_j^T, dP^T = V_t dO_j^T -> TMEM buffer g & 1
    // (the whole warp walks the loop; one elected lane issues, so ptxas emits straight-line UTCHMMA)
    constexpr uint32_t idesc_s64 = make_idesc(1, false, false, 128, 64);
    constexpr uint32_t idesc_s16 = make_idesc(1, false, false, 128, 16);
    const uint64_t dk0 = make_smem_desc_sw128(sk, 0, 1024), dq0 = make_smem_desc_sw128(sq, 0, 1024);
    const uint64_t dv0 = make_smem_desc_sw128(sv, 0, 1024), ddo0 = make_smem_desc_sw128(sdo, 0, 1024);
    mbar_wait(&bar[B_LOAD], 0, nullptr, 1);
    tc_fence_after();
    TRACE_DECL(1);
    TRACE(2);
    int g = 0;
    for (int t = 0; t < nt; ++t)
      for (int j = 0; j < nch; ++j, ++g) {
        const int buf = g & 1;
        if (g >= 2) {  // chunk g-2 has left this TMEM buffer
          mbar_wait(&bar[B_PD0 + buf], ((g - 2) >> 1) & 1, nullptr, 2);
          tc_fence_after();
        }
        TRACE(100 + g);
        if (elect_one()) {
          const uint32_t idesc = j < 3 ? idesc_s64 : idesc_s16;
          const uint64_t a_off = (uint64_t)(t * 1024), b_off = (uint64_t)(j * 512);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_base + 128 * buf, dk0 + a_off + (uint64_t)(ks * 2), dq0 + b_off + (uint64_t)(ks * 2), idesc, ks != 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_base + 128 * buf + 64, dv0 + a_off + (uint64_t)(ks * 2), ddo0 + b_off + (uint64_t)(ks * 2), idesc, ks != 0);
          umma_commit(&bar[B_S0 + buf]);
        }
        __syncwarp();
        TRACE(200 + g);
      }
  } else if (warp == 17) {
    // ---------------------------------------------------- issuer 2: dV, dK, dQ accumulation + dS^T store
    constexpr uint32_t idesc_kv = make_idesc(1, false, true, 128, kHeadDim);  // A K-major (P^T / dS^T), B MN-major
    constexpr uint32_t idesc_dq = make_idesc(1, true, true, 128, kHeadDim);   // A MN-major (dS), B MN-major (K)
    const uint64_t dpt0 = make_smem_desc_sw128(spt, 0, 1024), ddst0 = make_smem_desc_sw128(sdst, 0, 1024);
    const uint64_t ddo0 = make_smem_desc_sw128(sdo, 8192, 1024), dq0 = make_smem_desc_sw128(sq, 8192, 1024);
    const uint64_t dds0 = make_smem_desc_sw128(sdst, 16384, 1024), dk0 = make_smem_desc_sw128(sk, 8192, 1024);
    mbar_wait(&bar[B_LOAD], 0, nullptr, 3);
    TRACE_DECL(2);
    int g = 0;
    for (int t = 0; t < nt; ++t)
      for (int j = 0; j < nch; ++j, ++g) {
        const int buf = g & 1;
        const int slot = 2 * ((t * npair + (j >> 1)) & 1) + (j & 1);
        mbar_wait(&bar[B_PD0 + buf], (g >> 1) & 1, nullptr, 4);  // P^T, dS^T of chunk g are in smem
        TRACE(300 + g);
        if (t == 1 && j == 0) mbar_wait(&bar[B_ACCFREE], 0, nullptr, 5);  // dV_0 / dK_0 have been read out
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_pt = dpt0 + (uint64_t)(buf * 1024), a_dst = ddst0 + (uint64_t)(slot * 1024);
          const uint64_t b_off = (uint64_t)(j * 512);
          if (j < 3) issue_dv_dk<4>(tmem_base, a_pt, a_dst, ddo0 + b_off, dq0 + b_off, idesc_kv, j == 0);
          else issue_dv_dk<1>(tmem_base, a_pt, a_dst, ddo0 + b_off, dq0 + b_off, idesc_kv, false);
          if ((j & 1) || j == nch - 1) {  // a pair of chunks is complete: dQ_pair += dS K_t
            const uint64_t a_ds = dds0 + (uint64_t)((slot & 2) * 1024), b_k = dk0 + (uint64_t)(t * 1024);
            const uint32_t d = tmem_base + COL_DQ + 64 * (j >> 1);
            if (t == 0) issue_dq<8>(d, a_ds, b_k, idesc_dq, true);
            else issue_dq<5>(d, a_ds, b_k, idesc_dq, false);
          }
          if (p.write_ds) bulk_wait_read<0>();  // earlier dS^T stores have left smem before the commit frees slots
          umma_commit(&bar[B_B20 + buf]);
          if (j == nch - 1) umma_commit(&bar[B_ACC]);
          if (p.write_ds && j * 64 < p.ldb) {  // dS^T[b, h, keys of tile t, queries of chunk j]  (clipped at N / ldb)
            tma_store_3d(&tmap_ds, sdst + slot * 16384, j * 64, t * 128, blockIdx.x);
            bulk_commit();
          }
        }
        __syncwarp();
        TRACE(400 + g);
      }
    if (p.write_ds && elect_one()) bulk_wait_read<0>();
    __syncwarp();
  } else if (warp < 16) {
    // ------------------------------------------------------------------------------ elementwise / epilogue warps
    // two groups of 8 warps take the even / odd chunks (each on its own TMEM buffer and P^T slot), so that one
    // group's barrier round trips hide behind the other group's math
    const int grp = warp >> 3, w8 = warp & 7;
    const int quarter = warp & 3, half = w8 >> 2;
    const int part = warp >> 2;  // epilogue role: 4 x (tile half / 32-column half)
    const int rl = quarter * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    TRACE_DECL(0);
    int g = 0;
    for (int t = 0; t < nt; ++t) {
      const int key = t * 128 + rl;
      const bool kv = key < p.N;
      for (int j = 0; j < nch; ++j, ++g) {
        if ((g & 1) != grp) continue;
        const int buf = g & 1;
        const int slot = 2 * ((t * npair + (j >> 1)) & 1) + (j & 1);
        const bool wide = j < 3;                       // 64-query chunk: 32 columns per thread; last chunk: 8
        const int cl0 = wide ? half * 32 : half * 8;   // first column (within the chunk) of this thread
        const int q0 = j * 64 + cl0;
        // bias^T values of this thread's columns, issued before the waits
        float4 br0[4], br1[4];
        {
          const float4* bt = reinterpret_cast<const float4*>(p.biasT) + ((long long)h * kBiasGroups + (q0 >> 2)) * 256 + key;
          const bool on = p.biasT && kv;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            br0[i] = (on && (wide || i < 2) && q0 + 4 * i < p.N) ? ldg_f4(bt + i * 256) : make_float4(0.f, 0.f, 0.f, 0.f);
            br1[i] = (on && wide && q0 + 16 + 4 * i < p.N) ? ldg_f4(bt + (4 + i) * 256) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (warp == 0 && lane == 0) TRACE(500 + g);
        mbar_wait(&bar[B_S0 + buf], (g >> 1) & 1, nullptr, 6);
        if (g >= 2) mbar_wait(&bar[B_B20 + buf], ((g - 2) >> 1) & 1, nullptr, 7);  // P^T slot (and dS^T slot) reusable
        tc_fence_after();
        if (warp == 0 && lane == 0) TRACE(700 + g);
        const uint32_t ta = tlane + 128 * buf + cl0, s_pt = spt + buf * 16384, s_ds = sdst + slot * 16384;
        if (wide) {
          bwd_group<16>(p, smem, ta, br0, kv, q0, cl0, rl, s_pt, s_ds);
          bwd_group<16>(p, smem, ta + 16, br1, kv, q0 + 16, cl0 + 16, rl, s_pt, s_ds);
        } else {
          bwd_group<8>(p, smem, ta, br0, kv, q0, cl0, rl, s_pt, s_ds);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&bar[B_PD0 + buf]);
        if (warp == 0 && lane == 0) TRACE(800 + g);
      }
      // tile finished: dV_t (parts 0, 1) / dK_t (parts 2, 3), 32 columns each -> staging tiles in the two (now idle)
      // P^T slots -> coalesced TMA stores, clipped at the sequence end
      mbar_wait(&bar[B_ACC], t & 1, nullptr, 8);
      tc_fence_after();
      if (threadIdx.x == 0) TRACE(900 + t);
      {
        const bool is_k = part >= 2;
        stage_acc32(tlane + (is_k ? COL_DK : COL_DV) + (part & 1) * 32, is_k ? p.scale : 1.f, spt + (is_k ? 16384 : 0), rl,
                    (part & 1) * 4);
        tc_fence_before();
        mbar_arrive(&bar[B_ACCFREE]);
        fence_proxy_async();
        named_bar_sync(2, kBwdEw);
        if (threadIdx.x == 0) {
          tma_store_3d(&tmap_dqkv, spt, (2 * p.H + h) * kHeadDim, t * 128, b);
          tma_store_3d(&tmap_dqkv, spt + 16384, (p.H + h) * kHeadDim, t * 128, b);
          bulk_commit();
          bulk_wait_read<0>();
        }
        named_bar_sync(2, kBwdEw);  // the P^T slots are writable again
      }
    }
    // dQ: query tile (part >> 1), column half (part & 1)   (the last B_ACC phase covers every MMA of issuer 2)
    if ((part >> 1) < nt)
      stage_acc32(tlane + COL_DQ + 64 * (part >> 1) + (part & 1) * 32, p.scale, spt + (part >> 1) * 16384, rl, (part & 1) * 4);
    fence_proxy_async();
    named_bar_sync(2, kBwdEw);
    if (threadIdx.x == 0) {
      for (int t = 0; t < nt; ++t) tma_store_3d(&tmap_dqkv, spt + t * 16384, h * kHeadDim, t * 128, b);
      bulk_commit();
      bulk_wait_read<0>();  // smem may be released once the TMA unit has read it; the writes land before the grid ends
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
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
  attention_fwd_tc<<<grid, kFwdThreads, fwd::SMEM_BYTES, s>>>(tq, p);
  MEMB_LAUNCH_OK("attention_fwd_tc");
  return MEMB_OK;
}

#ifdef MEMB_ATTN_TRACE
extern "C" int memb_attention_trace_dump(long long* host, int max_pairs) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * 2 * 1536);
  static long long zeros[2 * 1536];
  cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros));
  return 1536;
}
#endif

extern "C" int memb_rowdot_heads(const void* a, const void* b, int B, int N, int H, float* delta, memb_stream_t s) {
  MEMB_REQUIRE(a && b && delta && B > 0 && N > 0 && H > 0, "rowdot_heads: bad arguments");
  MEMB_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0, "rowdot_heads: operands must be 16-byte aligned");
  const long long rows = (long long)B * N * H;
  MEMB_REQUIRE(rows < (1LL << 28), "rowdot_heads: B*N*H = %lld is beyond the kernel's 32-bit row indexing", rows);
  attn_rowdot_heads<<<(unsigned)ceil_div<long long>(rows * 8, 256), 256, 0, s>>>((const bf16*)a, (const bf16*)b, B, N, H, delta);
  MEMB_LAUNCH_OK("attn_rowdot_heads");
  return MEMB_OK;
}

extern "C" size_t memb_attention_bwd_workspace_bytes(int B, int N, int H) {
  return (size_t)2 * (size_t)B * (size_t)N * (size_t)H * sizeof(float);
}

extern "C" int memb_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const float* bias,
                                  const float* biasT, int ldb, int B, int N, int H, int head_dim, float scale, void* dqkv,
                                  void* dsT, void* workspace, size_t ws_bytes, memb_stream_t s) {
  (void)bias;  // the backward reads the transposed copy only
  if (int rc = check_shape(B, N, H, head_dim, ldb, dsT != nullptr)) return rc;
  MEMB_REQUIRE(qkv && dout && lse && dqkv && workspace, "attention_bwd: null pointer");
  MEMB_REQUIRE(ws_bytes >= memb_attention_bwd_workspace_bytes(B, N, H), "attention_bwd: workspace too small");
  MEMB_REQUIRE((bias == nullptr) == (biasT == nullptr), "attention_bwd: bias and its transpose go together");
  MEMB_REQUIRE(!biasT || (reinterpret_cast<uintptr_t>(biasT) & 15u) == 0, "attention_bwd: biasT must be 16-byte aligned");
  CUtensorMap tq, tdo, tds, tdq;
  {
    const long long dims[3] = {3LL * H * kHeadDim, N, B}, str[2] = {3LL * H * kHeadDim, (long long)N * 3 * H * kHeadDim};
    const int box[3] = {64, 128, 1};
    if (int rc = make_tmap_bf16(&tdq, dqkv, 3, dims, str, box)) return rc;
  }
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
  const long long rows = (long long)B * N * H;
  MEMB_REQUIRE(rows < (1LL << 28), "attention_bwd: B*N*H = %lld is beyond the 32-bit row indexing of the prep kernel", rows);
  if (out != nullptr) {       // out == NULL: the caller's producer of dO already left delta in the workspace
    attn_rowdot_heads<<<(unsigned)ceil_div<long long>(rows * 8, 256), 256, 0, s>>>((const bf16*)out, (const bf16*)dout, B, N, H,
                                                                               (float*)workspace);
    MEMB_LAUNCH_OK("attn_rowdot_heads");
  }
  BwdParams p{lse, (const float*)workspace, biasT, ldb, B, N, H, scale, scale * kLog2e, (bf16*)dqkv, dsT ? 1 : 0};
  attention_bwd_tc<<<B * H, kBwdThreads, bwd::SMEM_BYTES, s>>>(tq, tdo, tds, tdq, p);
  MEMB_LAUNCH_OK("attention_bwd_tc");
  return MEMB_OK;
}
