// Fused multi-head self-attention with an additive (relative-position) bias, forward and backward,
// for short sequences (N <= 208 tokens, head dim 64): one CTA per (batch, head), the whole K/V (and
// Q/dO) of the head resident in swizzled shared memory, one warp per 16-row tile, bf16 tensor-core
// MMAs (m16n8k16) with fp32 softmax.  Nothing of size N x N is written to HBM in the forward; the
// backward writes dS once (bf16) so that the bias gradient can be reduced over the batch.
//
// Reference: Attention.forward, mem/modeling_finetune.py:128-157  (q*scale, q@k^T, + rel_pos_bias,
// softmax, @v) and its autograd backward.
//
// Layouts: qkv / dqkv bf16 [B, N, 3, H, 64] (the QKV GEMM output viewed as in modeling_finetune.py:134);
// out / dout bf16 [B, N, H*64]; bias fp32 [H, N, ldb] (+ transposed copy for the backward);
// lse fp32 [B, H, N]; ds bf16 [B, H, N, ldb].
#include <cuda_bf16.h>

#include "common.cuh"

namespace memb {
namespace attn_mma {

using bf16 = __nv_bfloat16;
constexpr int kHeadDim = 64;
constexpr int kMaxTiles = 13;               // 13 x 16 = 208 >= 197 tokens
constexpr int kMaxRows = kMaxTiles * 16;
constexpr int kThreads = kMaxTiles * 32;    // one warp per 16-row tile
constexpr int kTileBytes = kMaxRows * 128;  // [208][64] bf16, 128 B per row, 16B chunks XOR-swizzled by row&7
constexpr int kHalfChunks = 4;              // forward keeps 4 key chunks (64 keys) of S in registers per pass
// Register budget: 13 warps are allocated as 16 (granularity 4), so 16 * 32 * regs <= 65536 -> 128 per thread.

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Copy rows [0,N) x 64 bf16 (row stride `stride` elements) into a swizzled tile; rows [N, 208) are zeroed.
__device__ __forceinline__ void load_tile(uint32_t tile, const bf16* src, long long stride, int N) {
  for (int t = threadIdx.x; t < kMaxRows * 8; t += kThreads) {
    const int row = t >> 3, chunk = t & 7;
    const uint32_t dst = tile_addr(tile, row, chunk);
    if (row < N) cp_async16(dst, src + (long long)row * stride + chunk * 8);
    else asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
  }
}

// A fragments (16 rows x 64 d) of rows [row0, row0+16) from a tile: 4 k-steps x 4 regs.
__device__ __forceinline__ void load_a_frags(uint32_t tile, int row0, uint32_t (&f)[4][4]) {
  const int lane = threadIdx.x & 31, mat = lane >> 3, r = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(tile_addr(tile, row0 + (mat & 1) * 8 + r, 2 * ks + (mat >> 1)), f[ks]);
}
__device__ __forceinline__ void load_a_frag(uint32_t tile, int row0, int ks, uint32_t (&f)[4]) {
  const int lane = threadIdx.x & 31, mat = lane >> 3, r = lane & 7;
  ldsm_x4(tile_addr(tile, row0 + (mat & 1) * 8 + r, 2 * ks + (mat >> 1)), f);
}
// B fragments "rows are n, columns are k" (K/Q/V/dO used as the [n][k] operand): rows [n0, n0+16), k-step ks.
// r[0],r[1] -> n8 tile n0..n0+7; r[2],r[3] -> n8 tile n0+8..n0+15.
__device__ __forceinline__ void load_b_nk(uint32_t tile, int n0, int ks, uint32_t (&r)[4]) {
  const int lane = threadIdx.x & 31, mat = lane >> 3, rr = lane & 7;
  ldsm_x4(tile_addr(tile, n0 + (mat >> 1) * 8 + rr, 2 * ks + (mat & 1)), r);
}
// B fragments "rows are k, columns are n" (tile row = reduction index): k rows [k0, k0+16), n8 tiles 2*np, 2*np+1.
__device__ __forceinline__ void load_b_kn(uint32_t tile, int k0, int np, uint32_t (&r)[4]) {
  const int lane = threadIdx.x & 31, mat = lane >> 3, rr = lane & 7;
  ldsm_x4_t(tile_addr(tile, k0 + (mat & 1) * 8 + rr, 2 * np + (mat >> 1)), r);
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __maxnreg__(128)
attention_fwd(const bf16* __restrict__ qkv, const float* __restrict__ bias, int ldb, int B, int N, int H, float scale,
              bf16* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sq = smem_u32(smem), sk = sq + kTileBytes, sv = sk + kTileBytes;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const long long rs = 3LL * H * kHeadDim;
  const bf16* qp = qkv + (long long)b * N * rs + h * kHeadDim;
  load_tile(sq, qp, rs, N);
  load_tile(sk, qp + H * kHeadDim, rs, N);
  load_tile(sv, qp + 2 * H * kHeadDim, rs, N);
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int ntiles = (N + 15) / 16;
  if (warp >= ntiles) return;
  const int row0 = warp * 16;
  uint32_t qf[4][4];
  load_a_frags(sq, row0, qf);

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows row0+g and row0+g+8
  const int qa = row0 + g, qb = row0 + g + 8;
  const float* bias_a = bias ? bias + ((long long)h * N + min(qa, N - 1)) * ldb : nullptr;
  const float* bias_b = bias ? bias + ((long long)h * N + min(qb, N - 1)) * ldb : nullptr;

#pragma unroll 1
  for (int c0 = 0; c0 < ntiles; c0 += kHalfChunks) {
    float s[2 * kHalfChunks][4];
#pragma unroll
    for (int kc = 0; kc < kHalfChunks; ++kc) {
      s[2 * kc][0] = s[2 * kc][1] = s[2 * kc][2] = s[2 * kc][3] = 0.f;
      s[2 * kc + 1][0] = s[2 * kc + 1][1] = s[2 * kc + 1][2] = s[2 * kc + 1][3] = 0.f;
      if (c0 + kc < ntiles) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t bb[4];
          load_b_nk(sk, (c0 + kc) * 16, ks, bb);
          mma16816(s[2 * kc], qf[ks], bb[0], bb[1]);
          mma16816(s[2 * kc + 1], qf[ks], bb[2], bb[3]);
        }
      }
    }
    // scale + bias + key mask, running max
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int t = 0; t < 2 * kHalfChunks; ++t) {
      const int key = (c0 + t / 2) * 16 + (t & 1) * 8 + 2 * c;
      float2 ba = make_float2(0.f, 0.f), bb2 = make_float2(0.f, 0.f);
      if (bias && key < N) {  // ldb is even and >= N+1 rounded, so the float2 is in bounds and aligned
        ba = *reinterpret_cast<const float2*>(bias_a + key);
        bb2 = *reinterpret_cast<const float2*>(bias_b + key);
      }
      s[t][0] = key < N ? s[t][0] * scale + ba.x : -INFINITY;
      s[t][1] = key + 1 < N ? s[t][1] * scale + ba.y : -INFINITY;
      s[t][2] = key < N ? s[t][2] * scale + bb2.x : -INFINITY;
      s[t][3] = key + 1 < N ? s[t][3] * scale + bb2.y : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[t][0], s[t][1]));
      mx1 = fmaxf(mx1, fmaxf(s[t][2], s[t][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float r0 = __expf(m0 - mx0), r1 = __expf(m1 - mx1);  // first half: exp(-inf) = 0
    m0 = mx0; m1 = mx1;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int t = 0; t < 2 * kHalfChunks; ++t) {
      s[t][0] = __expf(s[t][0] - m0); s[t][1] = __expf(s[t][1] - m0);
      s[t][2] = __expf(s[t][2] - m1); s[t][3] = __expf(s[t][3] - m1);
      sum0 += s[t][0] + s[t][1];
      sum1 += s[t][2] + s[t][3];
    }
    l0 = l0 * r0 + sum0;
    l1 = l1 * r1 + sum1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] *= r0; o[i][1] *= r0; o[i][2] *= r1; o[i][3] *= r1; }
    // O += P V
#pragma unroll
    for (int kc = 0; kc < kHalfChunks; ++kc) {
      if (c0 + kc < ntiles) {
        uint32_t a[4];
        a[0] = pack_bf16(s[2 * kc][0], s[2 * kc][1]);
        a[1] = pack_bf16(s[2 * kc][2], s[2 * kc][3]);
        a[2] = pack_bf16(s[2 * kc + 1][0], s[2 * kc + 1][1]);
        a[3] = pack_bf16(s[2 * kc + 1][2], s[2 * kc + 1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bb[4];
          load_b_kn(sv, (c0 + kc) * 16, np, bb);
          mma16816(o[2 * np], a, bb[0], bb[1]);
          mma16816(o[2 * np + 1], a, bb[2], bb[3]);
        }
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  bf16* oa = out + ((long long)b * N + qa) * H * kHeadDim + h * kHeadDim;
  bf16* ob = out + ((long long)b * N + qb) * H * kHeadDim + h * kHeadDim;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (qa < N) *reinterpret_cast<uint32_t*>(oa + 8 * i + 2 * c) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
    if (qb < N) *reinterpret_cast<uint32_t*>(ob + 8 * i + 2 * c) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
  }
  if (c == 0) {
    if (qa < N) lse[((long long)b * H + h) * N + qa] = m0 + __logf(l0);
    if (qb < N) lse[((long long)b * H + h) * N + qb] = m1 + __logf(l1);
  }
}

// ------------------------------------------------------------------------------------------ backward
__global__ void __maxnreg__(128)
attention_bwd(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
              const float* __restrict__ lse, const float* __restrict__ bias, const float* __restrict__ biasT, int ldb,
              int B, int N, int H, float scale, bf16* __restrict__ dqkv, bf16* __restrict__ ds) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sq = smem_u32(smem), sk = sq + kTileBytes, sv = sk + kTileBytes, sdo = sv + kTileBytes;
  float* s_lse = reinterpret_cast<float*>(smem + 4 * kTileBytes);
  float* s_delta = s_lse + kMaxRows;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const long long rs = 3LL * H * kHeadDim, os = (long long)H * kHeadDim;
  const bf16* qp = qkv + (long long)b * N * rs + h * kHeadDim;
  const bf16* dop = dout + (long long)b * N * os + h * kHeadDim;
  const bf16* op = out + (long long)b * N * os + h * kHeadDim;
  load_tile(sq, qp, rs, N);
  load_tile(sk, qp + H * kHeadDim, rs, N);
  load_tile(sv, qp + 2 * H * kHeadDim, rs, N);
  load_tile(sdo, dop, os, N);
  // delta[q] = sum_d dO[q,d] * O[q,d]  (read straight from global), lse -> smem
  for (int q = threadIdx.x; q < kMaxRows; q += kThreads) {
    float d = 0.f, l = 0.f;
    if (q < N) {
      const uint4* o4 = reinterpret_cast<const uint4*>(op + (long long)q * os);
      const uint4* g4 = reinterpret_cast<const uint4*>(dop + (long long)q * os);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 ov = o4[i], gv = g4[i];
        const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov);
        const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = __bfloat1622float2(oh[j]), bb = __bfloat1622float2(gh[j]);
          d += a.x * bb.x + a.y * bb.y;
        }
      }
      l = lse[((long long)b * H + h) * N + q];
    }
    s_delta[q] = d;
    s_lse[q] = l;
  }
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int ntiles = (N + 15) / 16;
  if (warp >= ntiles) return;
  const int row0 = warp * 16;
  const int ra = row0 + g, rb = row0 + g + 8;
  bf16* dq_base = dqkv + (long long)b * N * rs + h * kHeadDim;

  // ---- phase A: this warp owns keys [row0, row0+16): dK, dV
  {
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
    const float* bt_a = biasT ? biasT + ((long long)h * N + min(ra, N - 1)) * ldb : nullptr;
    const float* bt_b = biasT ? biasT + ((long long)h * N + min(rb, N - 1)) * ldb : nullptr;
#pragma unroll 1
    for (int qc = 0; qc < ntiles; ++qc) {
      float st[2][4], dpt[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t) { st[t][0] = st[t][1] = st[t][2] = st[t][3] = 0.f; dpt[t][0] = dpt[t][1] = dpt[t][2] = dpt[t][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bb[4], af[4];
        load_a_frag(sk, row0, ks, af);   // K/V fragments are re-read from smem (register budget)
        load_b_nk(sq, qc * 16, ks, bb);  // S^T = K_j Q^T
        mma16816(st[0], af, bb[0], bb[1]);
        mma16816(st[1], af, bb[2], bb[3]);
        load_a_frag(sv, row0, ks, af);
        load_b_nk(sdo, qc * 16, ks, bb);  // dP^T = V_j dO^T
        mma16816(dpt[0], af, bb[0], bb[1]);
        mma16816(dpt[1], af, bb[2], bb[3]);
      }
      uint32_t pa[4], dsa[4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int q = qc * 16 + t * 8 + 2 * c;  // columns q, q+1
        float2 ba = make_float2(0.f, 0.f), bb2 = make_float2(0.f, 0.f);
        if (biasT && q < N) {
          ba = *reinterpret_cast<const float2*>(bt_a + q);
          bb2 = *reinterpret_cast<const float2*>(bt_b + q);
        }
        const float l0 = s_lse[q], l1 = s_lse[q + 1], d0 = s_delta[q], d1 = s_delta[q + 1];
        const bool q0 = q < N, q1 = q + 1 < N, ka = ra < N, kb = rb < N;
        const float p00 = (q0 && ka) ? __expf(st[t][0] * scale + ba.x - l0) : 0.f;
        const float p01 = (q1 && ka) ? __expf(st[t][1] * scale + ba.y - l1) : 0.f;
        const float p10 = (q0 && kb) ? __expf(st[t][2] * scale + bb2.x - l0) : 0.f;
        const float p11 = (q1 && kb) ? __expf(st[t][3] * scale + bb2.y - l1) : 0.f;
        pa[2 * t] = pack_bf16(p00, p01);
        pa[2 * t + 1] = pack_bf16(p10, p11);
        dsa[2 * t] = pack_bf16(p00 * (dpt[t][0] - d0), p01 * (dpt[t][1] - d1));
        dsa[2 * t + 1] = pack_bf16(p10 * (dpt[t][2] - d0), p11 * (dpt[t][3] - d1));
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        load_b_kn(sdo, qc * 16, np, bb);  // dV += P^T dO
        mma16816(dv[2 * np], pa, bb[0], bb[1]);
        mma16816(dv[2 * np + 1], pa, bb[2], bb[3]);
        load_b_kn(sq, qc * 16, np, bb);   // dK += dS^T Q
        mma16816(dk[2 * np], dsa, bb[0], bb[1]);
        mma16816(dk[2 * np + 1], dsa, bb[2], bb[3]);
      }
    }
    bf16* dk_a = dq_base + (long long)ra * rs + H * kHeadDim;
    bf16* dk_b = dq_base + (long long)rb * rs + H * kHeadDim;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (ra < N) {
        *reinterpret_cast<uint32_t*>(dk_a + 8 * i + 2 * c) = pack_bf16(dk[i][0] * scale, dk[i][1] * scale);
        *reinterpret_cast<uint32_t*>(dk_a + H * kHeadDim + 8 * i + 2 * c) = pack_bf16(dv[i][0], dv[i][1]);
      }
      if (rb < N) {
        *reinterpret_cast<uint32_t*>(dk_b + 8 * i + 2 * c) = pack_bf16(dk[i][2] * scale, dk[i][3] * scale);
        *reinterpret_cast<uint32_t*>(dk_b + H * kHeadDim + 8 * i + 2 * c) = pack_bf16(dv[i][2], dv[i][3]);
      }
    }
  }

  // ---- phase B: this warp owns queries [row0, row0+16): dQ, dS
  {
    uint32_t qf[4][4], dof[4][4];
    load_a_frags(sq, row0, qf);
    load_a_frags(sdo, row0, dof);
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    const float* b_a = bias ? bias + ((long long)h * N + min(ra, N - 1)) * ldb : nullptr;
    const float* b_b = bias ? bias + ((long long)h * N + min(rb, N - 1)) * ldb : nullptr;
    const float la = s_lse[ra], lb = s_lse[rb], da = s_delta[ra], db = s_delta[rb];
    bf16* ds_a = ds ? ds + (((long long)b * H + h) * N + min(ra, N - 1)) * ldb : nullptr;
    bf16* ds_b = ds ? ds + (((long long)b * H + h) * N + min(rb, N - 1)) * ldb : nullptr;
#pragma unroll 1
    for (int kc = 0; kc < ntiles; ++kc) {
      float s[2][4], dp[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t) { s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f; dp[t][0] = dp[t][1] = dp[t][2] = dp[t][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bb[4];
        load_b_nk(sk, kc * 16, ks, bb);  // S = Q_i K^T
        mma16816(s[0], qf[ks], bb[0], bb[1]);
        mma16816(s[1], qf[ks], bb[2], bb[3]);
        load_b_nk(sv, kc * 16, ks, bb);  // dP = dO_i V^T
        mma16816(dp[0], dof[ks], bb[0], bb[1]);
        mma16816(dp[1], dof[ks], bb[2], bb[3]);
      }
      uint32_t dsa[4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int k = kc * 16 + t * 8 + 2 * c;
        float2 ba = make_float2(0.f, 0.f), bb2 = make_float2(0.f, 0.f);
        if (bias && k < N) {
          ba = *reinterpret_cast<const float2*>(b_a + k);
          bb2 = *reinterpret_cast<const float2*>(b_b + k);
        }
        const bool k0 = k < N, k1 = k + 1 < N, qa = ra < N, qb = rb < N;
        const float p00 = (k0 && qa) ? __expf(s[t][0] * scale + ba.x - la) : 0.f;
        const float p01 = (k1 && qa) ? __expf(s[t][1] * scale + ba.y - la) : 0.f;
        const float p10 = (k0 && qb) ? __expf(s[t][2] * scale + bb2.x - lb) : 0.f;
        const float p11 = (k1 && qb) ? __expf(s[t][3] * scale + bb2.y - lb) : 0.f;
        dsa[2 * t] = pack_bf16(p00 * (dp[t][0] - da), p01 * (dp[t][1] - da));
        dsa[2 * t + 1] = pack_bf16(p10 * (dp[t][2] - db), p11 * (dp[t][3] - db));
        if (ds && k < ldb) {  // padded columns [N, ldb) receive zeros
          if (qa) *reinterpret_cast<uint32_t*>(ds_a + k) = dsa[2 * t];
          if (qb) *reinterpret_cast<uint32_t*>(ds_b + k) = dsa[2 * t + 1];
        }
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        load_b_kn(sk, kc * 16, np, bb);  // dQ += dS K
        mma16816(dq[2 * np], dsa, bb[0], bb[1]);
        mma16816(dq[2 * np + 1], dsa, bb[2], bb[3]);
      }
    }
    bf16* dq_a = dq_base + (long long)ra * rs;
    bf16* dq_b = dq_base + (long long)rb * rs;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (ra < N) *reinterpret_cast<uint32_t*>(dq_a + 8 * i + 2 * c) = pack_bf16(dq[i][0] * scale, dq[i][1] * scale);
      if (rb < N) *reinterpret_cast<uint32_t*>(dq_b + 8 * i + 2 * c) = pack_bf16(dq[i][2] * scale, dq[i][3] * scale);
    }
  }
}

}  // namespace attn_mma
}  // namespace memb

using namespace memb;
using namespace memb::attn_mma;

static int check_shape(int B, int N, int H, int head_dim, int ldb, bool has_bias) {
  MEMB_REQUIRE(B > 0 && H > 0 && N > 0 && N <= kMaxRows, "attention: N must be in [1, %d], got %d", kMaxRows, N);
  MEMB_REQUIRE(head_dim == kHeadDim, "attention: head dim must be %d, got %d", kHeadDim, head_dim);
  MEMB_REQUIRE(!has_bias || (ldb % 2 == 0 && ldb >= N), "attention: bias row stride must be even and >= N");
  return MEMB_OK;
}

extern "C" int memb_attention_fwd_mma(const void* qkv, const float* bias, int ldb, int B, int N, int H, int head_dim,
                                  float scale, void* out, float* lse, memb_stream_t s) {
  if (int rc = check_shape(B, N, H, head_dim, ldb, bias != nullptr)) return rc;
  MEMB_REQUIRE(qkv && out && lse, "attention_fwd: null pointer");
  const int smem = 3 * kTileBytes;
  static bool configured = false;
  if (!configured) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(attention_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  attention_fwd<<<B * H, kThreads, smem, s>>>((const bf16*)qkv, bias, ldb, B, N, H, scale, (bf16*)out, lse);
  MEMB_LAUNCH_OK("attention_fwd");
  return MEMB_OK;
}

extern "C" int memb_attention_bwd_mma(const void* qkv, const void* out, const void* dout, const float* lse, const float* bias,
                                  const float* biasT, int ldb, int B, int N, int H, int head_dim, float scale, void* dqkv,
                                  void* ds, memb_stream_t s) {
  if (int rc = check_shape(B, N, H, head_dim, ldb, bias != nullptr)) return rc;
  MEMB_REQUIRE(qkv && out && dout && lse && dqkv, "attention_bwd: null pointer");
  MEMB_REQUIRE((bias == nullptr) == (biasT == nullptr), "attention_bwd: bias and its transpose go together");
  MEMB_REQUIRE(!ds || (ldb % 2 == 0 && ldb >= N), "attention_bwd: dS row stride must be even and >= N");
  const int smem = 4 * kTileBytes + 2 * kMaxRows * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(attention_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  attention_bwd<<<B * H, kThreads, smem, s>>>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, bias, biasT, ldb, B,
                                              N, H, scale, (bf16*)dqkv, (bf16*)ds);
  MEMB_LAUNCH_OK("attention_bwd");
  return MEMB_OK;
}
