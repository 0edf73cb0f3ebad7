// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )        bf16 x bf16 -> fp32 (kind::f16) or
//                                                 fp32-as-tf32 (kind::tf32, optional 3xTF32 split)
//
// One CTA per SM (persistent over output tiles), 256 threads:
//   warp 0   TMA producer   : cp.async.bulk.tensor -> 128B-swizzled smem ring (mbarrier complete_tx)
//   warp 1   MMA issuer     : one elected lane issues tcgen05.mma (M=128, N=BLOCK_N, K=32 bytes per
//                             instruction), accumulators live in TMEM, tcgen05.commit frees smem slots
//   warp 2   TMEM allocator : 2 x BLOCK_N columns = two accumulator stages (epilogue of tile i
//                             overlaps the main loop of tile i+1)
//   warps 4-11 epilogue     : two warps per TMEM lane quarter (each takes half of the tile's columns);
//                             tcgen05.ld of 32-column chunks is software-pipelined one chunk ahead of the
//                             fused epilogue math -> global stores
//
// Operands may be K-major (row-major [rows,K]) or MN-major (row-major [K,rows]); the latter is what
// dgrad (B = W as stored) and wgrad (A = dY, B = X as stored) need, so backward needs no transposes.
// Used by: ViT linears (reference mem/modeling_finetune.py:61-71,130-155, modeling_pretrain.py:126),
// patch embedding (:203-209) and the dVAE encoder convolutions as implicit/explicit GEMMs
// (eventvae/vae/vae_model.py:29-41,91-101).
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

#include <cuda_bf16.h>

#include "common.cuh"
#include "epilogue_math.cuh"
#include "sm100.cuh"

namespace memb {
namespace gemm_pair {
int try_launch(const memb_gemm_desc& g, cudaStream_t stream, bool* handled);  // gemm_pair.cu
}
namespace gemm {

using namespace memb::ptx;

constexpr int BLOCK_M = 128;
constexpr int kThreads = 384;  // warps 0-2 producer / MMA / TMEM alloc, warp 3 idle, warps 4-11 epilogue
constexpr int kSwizzleBytes = 128;  // one swizzle atom row = BLOCK_K elements

struct Params {
  int M, N, K;
  int num_n_tiles, num_tiles, splits, kb_per_split, kb_total;
  int split_precision;  // 3xTF32: A = [hi | lo] along K, B = [hi | lo] along K (K = logical K)
  // epilogue operands
  void* d;
  long long ldd;
  void* d2;
  long long ldd2;
  const float* bias;
  const void* aux;
  long long ldaux;
  const float* colscale;
  const float* rowscale;
  int rows_per_group;
  int out_group_rows, out_group_stride, out_row_offset;
  const unsigned char* rowmask;
  const float* maskvec;
  float alpha;
  const float* alpha_dev;
  int act;        // STORE: 0 none, 1 relu
  int out_split;  // STORE fp32: also write tf32 hi at [n] and lo at [N + n]
  int* err_flag;
};

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
  // larger value wins; on equal values the smaller index wins (torch.argmax first-max rule)
  uint32_t b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}

template <typename T>
__device__ __forceinline__ void store_row32(T* dst, const float (&v)[32], int valid);
template <>
__device__ __forceinline__ void store_row32<float>(float* dst, const float (&v)[32], int valid) {
  if (valid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < valid) dst[i] = v[i];
  }
}
template <>
__device__ __forceinline__ void store_row32<__nv_bfloat16>(__nv_bfloat16* dst, const float (&v)[32], int valid) {
  if (valid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 q;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
      __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
      q.x = *reinterpret_cast<uint32_t*>(&h0);
      q.y = *reinterpret_cast<uint32_t*>(&h1);
      q.z = *reinterpret_cast<uint32_t*>(&h2);
      q.w = *reinterpret_cast<uint32_t*>(&h3);
      reinterpret_cast<uint4*>(dst)[i] = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < valid) dst[i] = __float2bfloat16_rn(v[i]);
  }
}
__device__ __forceinline__ void load_row32(const float* src, float (&v)[32], int valid) {
  if (valid == 32 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 q = reinterpret_cast<const float4*>(src)[i];
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i < valid ? src[i] : 0.f;
  }
}
__device__ __forceinline__ void load_row32(const __nv_bfloat16* src, float (&v)[32], int valid) {
  if (valid == 32 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 q = reinterpret_cast<const uint4*>(src)[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(h[j]);
        v[8 * i + 2 * j] = f.x; v[8 * i + 2 * j + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i < valid ? __bfloat162float(src[i]) : 0.f;
  }
}

// 32 consecutive floats of a per-column vector (bias, LayerScale, mask token): the address is warp-uniform, so
// the eight 16-byte loads are broadcasts served by L1.
__device__ __forceinline__ void load_vec32(const float* src, float (&v)[32], int valid) {
  if (valid == 32 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(src) + i);
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i < valid ? __ldg(src + i) : 0.f;
  }
}

// Fused epilogue for one (row, 32-column chunk).  `v` holds the fp32 accumulators.
template <int EPI, typename OutT>
__device__ __forceinline__ void epilogue_chunk(const Params& p, int row, int col0, float (&v)[32], int valid) {
  if constexpr (EPI == MEMB_EPI_STORE) {
    if (p.bias) {
      float b[32];
      load_vec32(p.bias + col0, b, valid);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += b[i];
    }
    const float alpha = p.alpha_dev ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
    if (alpha != 1.0f) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= alpha;
    }
    if (p.aux) {  // residual add (fp32), e.g. dVAE ResBlock skip connection
      float r[32];
      load_row32(reinterpret_cast<const float*>(p.aux) + (long long)row * p.ldaux + col0, r, valid);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += r[i];
    }
    if (p.act == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    long long orow = row;
    if (p.out_group_rows > 0)
      orow = (long long)(row / p.out_group_rows) * p.out_group_stride + p.out_row_offset + row % p.out_group_rows;
    if (p.rowmask && p.rowmask[row]) load_vec32(p.maskvec + col0, v, valid);  // masked patch: the row becomes the mask token
    if constexpr (sizeof(OutT) == 4) {
      if (p.out_split) {
        float hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { hi[i] = tf32_round(v[i]); lo[i] = tf32_round(v[i] - hi[i]); }
        store_row32(reinterpret_cast<float*>(p.d) + orow * p.ldd + col0, hi, valid);
        store_row32(reinterpret_cast<float*>(p.d) + orow * p.ldd + p.N + col0, lo, valid);
        if (p.d2) store_row32(reinterpret_cast<float*>(p.d2) + orow * p.ldd2 + col0, v, valid);
        return;
      }
    }
    store_row32(reinterpret_cast<OutT*>(p.d) + orow * p.ldd + col0, v, valid);
  } else if constexpr (EPI == MEMB_EPI_BIAS_GELU) {
    float pre[32];
    if (p.bias) load_vec32(p.bias + col0, pre, valid);
    else {
#pragma unroll
      for (int i = 0; i < 32; ++i) pre[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      pre[i] += v[i];
      v[i] = epi::gelu_fwd(pre[i]);
    }
    store_row32(reinterpret_cast<OutT*>(p.d) + (long long)row * p.ldd + col0, v, valid);
    if (p.d2) store_row32(reinterpret_cast<__nv_bfloat16*>(p.d2) + (long long)row * p.ldd2 + col0, pre, valid);
  } else if constexpr (EPI == MEMB_EPI_RESIDUAL) {
    // x_out = x_in + rowscale[row / rows_per_group] * colscale[n] * (acc + bias[n]); fp32 stream.
    float r[32];
    load_row32(reinterpret_cast<const float*>(p.aux) + (long long)row * p.ldaux + col0, r, valid);
    const float rs = p.rowscale ? __ldg(p.rowscale + row / p.rows_per_group) : 1.0f;
    if (p.bias) {
      float b[32];
      load_vec32(p.bias + col0, b, valid);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += b[i];
    }
    if (p.colscale) {
      float g[32];
      load_vec32(p.colscale + col0, g, valid);
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = fmaf(rs * g[i], v[i], r[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = fmaf(rs, v[i], r[i]);
    }
    store_row32(reinterpret_cast<float*>(p.d) + (long long)row * p.ldd + col0, r, valid);
    if (p.d2) store_row32(reinterpret_cast<__nv_bfloat16*>(p.d2) + (long long)row * p.ldd2 + col0, v, valid);
  } else if constexpr (EPI == MEMB_EPI_ATOMIC_ADD) {
    float* dst = reinterpret_cast<float*>(p.d) + (long long)row * p.ldd + col0;
    const float alpha = p.alpha_dev ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
    if (valid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * i), "f"(v[4 * i] * alpha),
                     "f"(v[4 * i + 1] * alpha), "f"(v[4 * i + 2] * alpha), "f"(v[4 * i + 3] * alpha)
                     : "memory");
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < valid) atomicAdd(dst + i, v[i] * alpha);
    }
  } else if constexpr (EPI == MEMB_EPI_DGELU) {
    float pre[32];
    load_row32(reinterpret_cast<const __nv_bfloat16*>(p.aux) + (long long)row * p.ldaux + col0, pre, valid);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= epi::gelu_grad(pre[i]);
    store_row32(reinterpret_cast<OutT*>(p.d) + (long long)row * p.ldd + col0, v, valid);
  }
}

template <int BLOCK_N, int kElemBytes>
struct Cfg {
  static constexpr int BLOCK_K = kSwizzleBytes / kElemBytes;  // 64 bf16 / 32 tf32
  static constexpr int UMMA_K = 32 / kElemBytes;              // 16 bf16 / 8 tf32
  static constexpr int A_BYTES = BLOCK_M * kSwizzleBytes;     // 16 KB
  static constexpr int B_BYTES = BLOCK_N * kSwizzleBytes;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, int kElemBytes, bool A_MN, bool B_MN, int EPI, typename OutT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  using C = Cfg<BLOCK_N, kElemBytes>;
  static_assert(!(A_MN || B_MN) || kElemBytes == 2, "MN-major operands are implemented for bf16 only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  grid_dependency_trigger();      // the next kernel of the stream may set itself up under this kernel's tail
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dependency_wait();         // everything above ran while the previous kernel drained; operands / outputs from here on

  const int kb_mult = p.split_precision ? 3 : 1;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.num_n_tiles;
        const int rest = tile / p.num_n_tiles;
        const int split = rest % p.splits;
        const int m_tile = rest / p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int it = 0; it < (kb1 - kb0) * kb_mult; ++it) {
          int kb = kb0 + it, a_koff = 0, b_koff = 0;
          if (p.split_precision) {  // order: lo*hi, hi*lo, hi*hi (small terms first)
            const int seg = it / (kb1 - kb0);
            kb = kb0 + it % (kb1 - kb0);
            a_koff = (seg == 0) ? p.K : 0;
            b_koff = (seg == 1) ? p.K : 0;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * C::A_BYTES;
          uint8_t* sb = smem_b + stage * C::B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], a_koff + kb * C::BLOCK_K, m_tile * BLOCK_M);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)  // [64 MN-elements x BLOCK_K rows] boxes
              tma_load_2d(sa + j * (64 * C::BLOCK_K * 2), &tmap_a, &full_bar[stage], m_tile * BLOCK_M + j * 64, kb * C::BLOCK_K);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], b_koff + kb * C::BLOCK_K, n_tile * BLOCK_N);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * (64 * C::BLOCK_K * 2), &tmap_b, &full_bar[stage], n_tile * BLOCK_N + j * 64, kb * C::BLOCK_K);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = make_idesc(kElemBytes == 2 ? 1 : 2, A_MN, B_MN, BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int rest = tile / p.num_n_tiles;
        const int split = rest % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, p.err_flag, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        const int iters = (kb1 - kb0) * kb_mult;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < C::BLOCK_K / C::UMMA_K; ++k) {
            // K-major: 8-row groups are 1024 B apart (SBO), a K step of 32 B moves the start address.
            // MN-major: 64-element column blocks are one TMA box apart (LBO), 8-row K groups 1024 B (SBO),
            //           a K step of 16 rows moves the start address by 2048 B.
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 64 * C::BLOCK_K * 2, 1024)
                                        : make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 64 * C::BLOCK_K * 2, 1024)
                                        : make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            if constexpr (kElemBytes == 2) umma_bf16(d_tmem, adesc, bdesc, idesc, (it | k) != 0);
            else umma_tf32(d_tmem, adesc, bdesc, idesc, (it | k) != 0);
          }
          umma_commit(&empty_bar[stage]);                      // smem slot reusable once these MMAs retire
          if (it == iters - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------- epilogue (8 warps: lane quarter x column half)
    const int quad = warp & 3, half = (warp - 4) >> 2;
    constexpr int kCols = BLOCK_N / 2, kChunks = kCols / 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = (tile / p.num_n_tiles) / p.splits;
      mbar_wait(&tmem_full_bar[acc], acc_phase, p.err_flag, 4);
      tc_fence_after();
      const int row = m_tile * BLOCK_M + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BLOCK_N + half * kCols;
      const int colbase = n_tile * BLOCK_N + half * kCols;
      uint32_t r[2][32];
      tmem_ld32(taddr, r[0]);  // warp-collective: every lane takes part even for rows >= M
      unsigned long long best = 0ull;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        tmem_ld_wait();
        if (c + 1 < kChunks) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);  // next chunk in flight during the math
        const int col0 = colbase + c * 32;
        const int valid = min(32, p.N - col0);
        if constexpr (EPI == MEMB_EPI_ARGMAX) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (col0 + i < p.N) {
              const float val = __uint_as_float(r[c & 1][i]) + (p.bias ? __ldg(p.bias + col0 + i) : 0.f);
              const unsigned long long key = argmax_key(val, col0 + i);
              best = key > best ? key : best;
            }
          }
        } else {
          if (row < p.M && valid > 0) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[c & 1][i]);
            epilogue_chunk<EPI, OutT>(p, row, col0, v, valid);
          }
        }
      }
      if constexpr (EPI == MEMB_EPI_ARGMAX) {
        if (row < p.M) atomicMax(reinterpret_cast<unsigned long long*>(p.d) + row, best);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D row-major array [rows, cols] of `elem_bytes` elements with leading dimension ld (elements);
// box = [box_rows, box_cols]; 128B swizzle (box_cols * elem_bytes must be 128).
static int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, long long rows, long long cols,
                        long long ld, int box_rows, int box_cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  MEMB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, "gemm: operand base must be 16-byte aligned");
  MEMB_REQUIRE((ld * elem_bytes) % 16 == 0, "gemm: operand leading dimension must be a multiple of 16 bytes");
  MEMB_REQUIRE(box_cols * elem_bytes == kSwizzleBytes && box_rows <= 256, "gemm: bad TMA box");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * elem_bytes)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MEMB_OK;
}

template <int BLOCK_N, int kElemBytes, bool A_MN, bool B_MN, int EPI, typename OutT>
static int launch(const memb_gemm_desc& g, const Params& p, const CUtensorMap& ta, const CUtensorMap& tb, int grid,
                  cudaStream_t stream) {
  using C = Cfg<BLOCK_N, kElemBytes>;
  auto kern = gemm_tcgen05<BLOCK_N, kElemBytes, A_MN, B_MN, EPI, OutT>;
  static bool configured = false;
  if (!configured) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  MEMB_CUDA_OK(launch_dependent(kern, dim3(grid), dim3(kThreads), (size_t)C::SMEM_BYTES, stream, ta, tb, p));
  MEMB_LAUNCH_OK("gemm_tcgen05");
  return MEMB_OK;
}

template <int BLOCK_N, int kElemBytes, bool A_MN, bool B_MN>
static int dispatch_epi(const memb_gemm_desc& g, const Params& p, const CUtensorMap& ta, const CUtensorMap& tb,
                        int grid, cudaStream_t s) {
  const bool f32 = g.out_dtype == MEMB_DT_F32;
  switch (g.epilogue) {
    case MEMB_EPI_STORE:
      return f32 ? launch<BLOCK_N, kElemBytes, A_MN, B_MN, MEMB_EPI_STORE, float>(g, p, ta, tb, grid, s)
                 : launch<BLOCK_N, kElemBytes, A_MN, B_MN, MEMB_EPI_STORE, __nv_bfloat16>(g, p, ta, tb, grid, s);
    case MEMB_EPI_ATOMIC_ADD:
      return launch<BLOCK_N, kElemBytes, A_MN, B_MN, MEMB_EPI_ATOMIC_ADD, float>(g, p, ta, tb, grid, s);
    case MEMB_EPI_ARGMAX:
      return launch<BLOCK_N, kElemBytes, A_MN, B_MN, MEMB_EPI_ARGMAX, float>(g, p, ta, tb, grid, s);
    default:
      break;
  }
  if constexpr (kElemBytes == 2) {
    switch (g.epilogue) {
      case MEMB_EPI_BIAS_GELU:
        return launch<BLOCK_N, 2, A_MN, B_MN, MEMB_EPI_BIAS_GELU, __nv_bfloat16>(g, p, ta, tb, grid, s);
      case MEMB_EPI_RESIDUAL:
        return launch<BLOCK_N, 2, A_MN, B_MN, MEMB_EPI_RESIDUAL, float>(g, p, ta, tb, grid, s);
      case MEMB_EPI_DGELU:
        return launch<BLOCK_N, 2, A_MN, B_MN, MEMB_EPI_DGELU, __nv_bfloat16>(g, p, ta, tb, grid, s);
      default:
        break;
    }
  }
  return fail(MEMB_EINVAL, "gemm: unsupported epilogue %d for this operand type", g.epilogue);
}

template <int BLOCK_N>
static int dispatch_layout(const memb_gemm_desc& g, const Params& p, const CUtensorMap& ta, const CUtensorMap& tb,
                           int grid, cudaStream_t s) {
  if (g.in_dtype == MEMB_DT_F32) {
    MEMB_REQUIRE(!g.a_layout && !g.b_layout, "gemm: tf32 operands must be K-major");
    return dispatch_epi<BLOCK_N, 4, false, false>(g, p, ta, tb, grid, s);
  }
  if (!g.a_layout && !g.b_layout) return dispatch_epi<BLOCK_N, 2, false, false>(g, p, ta, tb, grid, s);
  if (!g.a_layout && g.b_layout) return dispatch_epi<BLOCK_N, 2, false, true>(g, p, ta, tb, grid, s);
  if (g.a_layout && g.b_layout) return dispatch_epi<BLOCK_N, 2, true, true>(g, p, ta, tb, grid, s);
  return dispatch_epi<BLOCK_N, 2, true, false>(g, p, ta, tb, grid, s);
}

}  // namespace gemm
}  // namespace memb

using namespace memb;
using namespace memb::gemm;

extern "C" int memb_gemm(const memb_gemm_desc* gp, memb_stream_t stream) {
  MEMB_REQUIRE(gp != nullptr, "gemm: null descriptor");
  const memb_gemm_desc& g = *gp;
  MEMB_REQUIRE(g.m > 0 && g.n > 0 && g.k > 0, "gemm: m, n, k must be positive (%d, %d, %d)", g.m, g.n, g.k);
  MEMB_REQUIRE(g.a && g.b && g.d, "gemm: null operand");
  MEMB_REQUIRE(g.in_dtype == MEMB_DT_BF16 || g.in_dtype == MEMB_DT_F32, "gemm: in_dtype must be bf16 or fp32(tf32)");
  MEMB_REQUIRE(g.epilogue >= MEMB_EPI_STORE && g.epilogue <= MEMB_EPI_STORE_ROWDOT, "gemm: unknown epilogue %d", g.epilogue);
  if (g.epilogue == MEMB_EPI_STORE_ROWDOT)
    MEMB_REQUIRE(g.aux && g.rowdot && g.rows_per_group > 0 && g.n % 64 == 0 && g.m % g.rows_per_group == 0 &&
                     g.out_dtype == MEMB_DT_BF16 && g.ldd == g.n && g.ldaux == g.n && !g.bias,
                 "gemm: STORE_ROWDOT needs bf16 d and aux (dense [M,N], N %% 64 == 0), rowdot, rows_per_group dividing M, no bias");
  {  // large bf16 K-major-A problems with a fused epilogue run on the CTA-pair kernel
    bool handled = false;
    if (int rc = gemm_pair::try_launch(g, stream, &handled)) return rc;
    if (handled) return MEMB_OK;
  }
  if (g.epilogue == MEMB_EPI_STORE_ROWDOT) {   // this kernel has no fused row dots: the plain GEMM, then one pass over d and aux
    memb_gemm_desc plain = g;
    plain.epilogue = MEMB_EPI_STORE;
    plain.aux = nullptr; plain.ldaux = 0; plain.rowdot = nullptr; plain.rows_per_group = 0;
    if (int rc = memb_gemm(&plain, stream)) return rc;
    return memb_rowdot_heads(g.d, g.aux, g.m / g.rows_per_group, g.rows_per_group, g.n / 64, g.rowdot, stream);
  }
  if (g.colsum != nullptr) {
    MEMB_REQUIRE(g.epilogue == MEMB_EPI_DGELU && g.out_dtype == MEMB_DT_BF16, "gemm: colsum rides on the DGELU epilogue with a bf16 output");
    memb_gemm_desc plain = g;        // this kernel has no fused column sums: the same GEMM, then one pass over its output
    plain.colsum = nullptr;
    if (int rc = memb_gemm(&plain, stream)) return rc;
    return memb_colsum_bf16(g.d, g.ldd, g.m, g.n, g.colsum, stream);
  }
  const int eb = g.in_dtype == MEMB_DT_BF16 ? 2 : 4;
  const int block_k = kSwizzleBytes / eb;
  const int block_n = (g.block_n == 128 || g.block_n == 256) ? g.block_n : ((g.n % 256 == 0 || g.n > 1024) ? 256 : 128);
  if (g.split_precision) MEMB_REQUIRE(g.in_dtype == MEMB_DT_F32 && g.k % block_k == 0, "gemm: 3xTF32 needs fp32 operands and K % 32 == 0");
  if (g.epilogue == MEMB_EPI_RESIDUAL || g.epilogue == MEMB_EPI_DGELU) MEMB_REQUIRE(g.aux != nullptr, "gemm: epilogue needs aux");
  if (g.epilogue == MEMB_EPI_RESIDUAL && g.rowscale) MEMB_REQUIRE(g.rows_per_group > 0, "gemm: rows_per_group must be positive");

  Params p{};
  p.M = g.m; p.N = g.n; p.K = g.k;
  p.num_n_tiles = ceil_div(g.n, block_n);
  const int num_m_tiles = ceil_div(g.m, BLOCK_M);
  p.kb_total = ceil_div(g.k, block_k);
  const int sms = num_sms();
  int splits = g.splits;
  if (g.epilogue != MEMB_EPI_ATOMIC_ADD) splits = 1;
  else if (splits <= 0) {
    const int out_tiles = num_m_tiles * p.num_n_tiles;
    splits = std::max(1, std::min(p.kb_total / 4, (2 * sms) / std::max(1, out_tiles)));
  }
  p.kb_per_split = ceil_div(p.kb_total, std::max(1, splits));
  p.splits = ceil_div(p.kb_total, p.kb_per_split);
  p.num_tiles = num_m_tiles * p.num_n_tiles * p.splits;
  p.split_precision = g.split_precision;
  p.d = g.d; p.ldd = g.ldd; p.d2 = g.d2; p.ldd2 = g.ldd2;
  p.bias = g.bias; p.aux = g.aux; p.ldaux = g.ldaux;
  p.colscale = g.colscale; p.rowscale = g.rowscale; p.rows_per_group = g.rows_per_group;
  p.out_group_rows = g.out_group_rows; p.out_group_stride = g.out_group_stride; p.out_row_offset = g.out_row_offset;
  p.rowmask = g.rowmask; p.maskvec = g.maskvec;
  p.alpha = g.alpha == 0.0f ? 1.0f : g.alpha;
  p.alpha_dev = g.alpha_dev;
  p.act = g.act; p.out_split = g.out_split;
  p.err_flag = g.err_flag;

  CUtensorMap ta, tb;
  const long long kcols = (long long)g.k * (g.split_precision ? 2 : 1);
  int rc;
  if (!g.a_layout) rc = make_tmap_2d(&ta, g.a, eb, g.m, kcols, g.lda, BLOCK_M, block_k);
  else rc = make_tmap_2d(&ta, g.a, eb, g.k, g.m, g.lda, block_k, 64);
  if (rc) return rc;
  if (!g.b_layout) rc = make_tmap_2d(&tb, g.b, eb, g.n, kcols, g.ldb, block_n, block_k);
  else rc = make_tmap_2d(&tb, g.b, eb, g.k, g.n, g.ldb, block_k, 64);
  if (rc) return rc;

  const int grid = std::min(p.num_tiles, sms);
  if (block_n == 256) return dispatch_layout<256>(g, p, ta, tb, grid, stream);
  return dispatch_layout<128>(g, p, ta, tb, grid, stream);
}
