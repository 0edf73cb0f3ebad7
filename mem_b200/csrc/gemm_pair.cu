// CTA-pair (cta_group::2) tcgen05 GEMM with a TMA-staged fused epilogue -- the workhorse of the ViT step.
//
//   D[M,N] = epilogue( A[M,K] * B^T ),  A bf16 K-major, B bf16 K-major [N,K] (forward) or MN-major [K,N] (dgrad)
//
// A cluster of two CTAs (one SM each) computes a 256 x BLOCK_N tile: UMMA M = 256, each CTA stages its own
// 128 rows of A and HALF of the B tile (the tensor core reads the other half from the peer's shared memory),
// so the L2 -> SMEM traffic per flop is 2/3 of the single-CTA 128 x 256 tile and a stage is 32 KB instead of
// 48 KB.  The freed shared memory holds per-warp epilogue staging:
//
//   warp 0      TMA producer (both CTAs; completion bytes land on the LEADER's full barrier)
//   warp 1      MMA issuer (leader CTA only; tcgen05.commit multicasts the arrivals to both CTAs)
//   warp 2      TMEM allocator (2 accumulator stages x BLOCK_N columns, in both CTAs)
//   warps 4-11  epilogue: tcgen05.ld 32-column chunks -> fused math in registers -> swizzled staging tile
//               -> cp.async.bulk.tensor store (coalesced, clipped at the M tail by the TMA unit).  Epilogue
//               inputs (fp32 residual stream, bf16 pre-activation) are prefetched into the same staging
//               ring by TMA loads one or two chunks ahead; per-column vectors (bias, LayerScale) are staged
//               once per tile.
//
// Epilogues (reference: mem/modeling_finetune.py:66-71 Mlp, :128-157 Attention, :182-189 Block):
//   STORE      d = bf16(acc + bias)
//   BIAS_GELU  d = bf16(gelu(acc + bias)),  d2 = bf16(acc + bias)
//   RESIDUAL   d = aux + rowscale[row / rows_per_group] * colscale[n] * (acc + bias)  (fp32), d2 = bf16(acc + bias)
//   DGELU      d = bf16(acc * gelu'(aux))            (+ optional column sums of d: the bias gradient)
//   ROWDOT     d = bf16(acc), rowdot = per-head row sums of d * aux   (the attention backward's rowsum(dO * O))
#include <algorithm>
#include <cstdlib>

#include <cuda_bf16.h>

#include "common.cuh"
#include "epilogue_math.cuh"
#include "sm100.cuh"

namespace memb {
namespace gemm_pair {

using namespace memb::ptx;
using namespace memb::epi;

constexpr int BLOCK_M = 128;  // rows per CTA; the pair covers 256
constexpr int BLOCK_K = 64;   // bf16: one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int A_BYTES = BLOCK_M * 128;
constexpr int kSmemMax = 232448;  // 227 KB opt-in limit per CTA

struct Params {
  int M, N, K;
  int num_n_tiles, num_tiles, kb_total;
  const float* bias;
  const float* colscale;
  const float* rowscale;
  int rows_per_group;
  int has_d2;
  int* err_flag;
  float* colsum;   // DGELU: += column sums of the output (may be NULL)
  float* rowdot;   // STORE_ROWDOT: [M / rows_per_group][N / 64][rows_per_group]
};
// epilogues that stream a bf16 aux tile of the output's shape through the per-warp staging ring
template <int EPI> constexpr bool kAuxBf16 = (EPI == MEMB_EPI_DGELU || EPI == MEMB_EPI_STORE_ROWDOT);

template <int BLOCK_N, int EPI>
struct Cfg {
  static constexpr int B_BYTES = (BLOCK_N / 2) * 128;  // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARP_BYTES = (EPI == MEMB_EPI_RESIDUAL) ? 10240 : 8192;
  static constexpr int VEC_FLOATS = BLOCK_N * ((EPI == MEMB_EPI_RESIDUAL) ? 2 : 1);  // bias (+ LayerScale)
  static constexpr int VEC_BYTES = 2 * VEC_FLOATS * 4;                                // per accumulator stage
  static constexpr int BAR_BYTES = 512;
  static constexpr int FIXED = kEpiWarps * EPI_WARP_BYTES + VEC_BYTES + BAR_BYTES;
  static constexpr int STAGES_RAW = (kSmemMax - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 256) ? 256 : 512;
  static_assert(STAGES >= 3, "not enough shared memory for a useful pipeline");
};

// staging tiles: rows of 64 B (bf16 x 32, TMA SWIZZLE_64B) or 128 B (fp32 x 32, SWIZZLE_128B); lane = row
__device__ __forceinline__ uint32_t sw64(uint32_t base, int row, int chunk) {
  return base + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ uint32_t sw128(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}

template <int BLOCK_N, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_d2,
                 const __grid_constant__ CUtensorMap tmap_aux, const Params p) {
  using C = Cfg<BLOCK_N, EPI>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * A_BYTES;
  uint8_t* epi_stage = smem + C::STAGES * C::STAGE_BYTES;
  float* vec = reinterpret_cast<float*>(epi_stage + kEpiWarps * C::EPI_WARP_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(vec) + C::VEC_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* aux_bar = tmem_empty_bar + 2;  // [kEpiWarps][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + kEpiWarps * 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if ((smem_u32(smem) & 1023u) != 0) {  // the swizzled layouts assume a 1024-byte aligned base
    if (threadIdx.x == 0 && p.err_flag) atomicExch(p.err_flag, 90);
    return;
  }
  grid_dependency_trigger();      // the next kernel of the stream may set itself up under this kernel's tail
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_d);
    if (EPI == MEMB_EPI_BIAS_GELU || EPI == MEMB_EPI_RESIDUAL) prefetch_tmap(&tmap_d2);
    if (EPI == MEMB_EPI_RESIDUAL || kAuxBf16<EPI>) prefetch_tmap(&tmap_aux);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(&empty_bar[s], 1);  // one multicast commit from the leader's MMA warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * kEpiWarps);  // leader's: every epilogue warp of both CTAs
    }
    for (int s = 0; s < kEpiWarps * 4; ++s) mbar_init(&aux_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dependency_wait();         // everything above ran while the previous kernel drained; operands / outputs from here on

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
      const int n_tile = tile % p.num_n_tiles, m_pair = tile / p.num_n_tiles;
      const int m0 = (m_pair * 2 + (int)cta_rank) * BLOCK_M;
      const int n0 = n_tile * BLOCK_N + (int)cta_rank * (BLOCK_N / 2);
      for (int kb = 0; kb < p.kb_total; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
        if (elect_one()) {
          const uint32_t fb = mapa(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          const uint32_t sa = smem_u32(smem_a + stage * A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * C::B_BYTES);
          tma_load_2d_pair(sa, &tmap_a, fb, kb * BLOCK_K, m0);
          if constexpr (!B_MN) {
            tma_load_2d_pair(sb, &tmap_b, fb, kb * BLOCK_K, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 128; ++j)  // [64 N-elements x BLOCK_K rows] boxes
              tma_load_2d_pair(sb + j * (64 * BLOCK_K * 2), &tmap_b, fb, n0 + j * 64, kb * BLOCK_K);
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // -------------------------------------------------------------- MMA issuer (leader CTA)
      constexpr uint32_t idesc = make_idesc(1, false, B_MN, 2 * BLOCK_M, BLOCK_N);
      const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
      // descriptor templates: K-major {LBO 0, SBO 1024}; MN-major {LBO = one 64-column box, SBO 1024}
      const uint64_t adesc0 = make_smem_desc_sw128(a_base, 0, 1024);
      const uint64_t bdesc0 = B_MN ? make_smem_desc_sw128(b_base, 64 * BLOCK_K * 2, 1024) : make_smem_desc_sw128(b_base, 0, 1024);
      constexpr uint32_t kStepA = 32 >> 4;                    // 16 K elements = 32 B along a K-major row
      constexpr uint32_t kStepB = (B_MN ? 2048 : 32) >> 4;    // MN-major: 16 K rows = 2048 B
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, p.err_flag, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = adesc0 + (uint64_t)((stage * A_BYTES) >> 4);
            const uint64_t bd = bdesc0 + (uint64_t)((stage * C::B_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16_pair(d_tmem, ad + k * kStepA, bd + k * kStepB, idesc, (kb | k) != 0);
            umma_commit_pair(&empty_bar[stage], 3);                            // frees this slot in both CTAs
            if (kb == p.kb_total - 1) umma_commit_pair(&tmem_full_bar[acc], 3);  // accumulator complete (both CTAs)
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (8 warps: lane quarter x column half)
    const int ew = warp - 4, quad = warp & 3, half = ew >> 2;
    constexpr int kCols = BLOCK_N / 2, kChunks = kCols / 32;
    const uint32_t stg = smem_u32(epi_stage + ew * C::EPI_WARP_BYTES);
    uint64_t* abar = aux_bar + ew * 4;
    const uint32_t tmem_empty_leader0 = mapa(smem_u32(&tmem_empty_bar[0]), 0);
    const int et = threadIdx.x - 128;  // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t ring = 0;        // running chunk counter: selects staging slots across tiles
    uint32_t aux_phases = 0;  // one parity bit per aux barrier slot
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
      const int n_tile = tile % p.num_n_tiles, m_pair = tile / p.num_n_tiles;
      const int row0 = (m_pair * 2 + (int)cta_rank) * BLOCK_M + quad * 32;
      const int colw = n_tile * BLOCK_N + half * kCols;
      const bool live = row0 < p.M;  // warp-uniform; a dead warp still drains its TMEM quarter
      // per-tile column vectors -> smem (double-buffered with the accumulator stage)
      float* vb = vec + acc * C::VEC_FLOATS;
      if (et < BLOCK_N) {
        const int col = n_tile * BLOCK_N + et;
        vb[et] = (p.bias && col < p.N) ? __ldg(p.bias + col) : 0.f;
        if constexpr (EPI == MEMB_EPI_RESIDUAL) vb[BLOCK_N + et] = (col < p.N) ? (p.colscale ? __ldg(p.colscale + col) : 1.f) : 0.f;
      }
      named_bar_sync(1, kEpiWarps * 32);
      const uint32_t vbs = smem_u32(vb) + (half * kCols) * 4;

      // aux prefetch for the first chunk(s) of this tile (overlaps the tail of the main loop)
      if constexpr (kAuxBf16<EPI>) {
        if (elect_one()) {
          bulk_wait_read<0>();
#pragma unroll
          for (int c = 0; c < 2 && c < kChunks; ++c) {
            const uint32_t slot = (ring + c) & 3;
            fence_proxy_async();
            mbar_arrive_expect_tx(&abar[slot], 32 * 64);
            tma_load_2d_addr(stg + slot * 2048, &tmap_aux, smem_u32(&abar[slot]), colw + c * 32, row0);
          }
        }
        __syncwarp();
      } else if constexpr (EPI == MEMB_EPI_RESIDUAL) {
        if (elect_one()) {
          bulk_wait_read<0>();
          const uint32_t slot = ring & 1;
          fence_proxy_async();
          mbar_arrive_expect_tx(&abar[slot], 32 * 128);
          tma_load_2d_addr(stg + slot * 4096, &tmap_aux, smem_u32(&abar[slot]), colw, row0);
        }
        __syncwarp();
      }
      float dot = 0.f;               // ROWDOT: this lane's row, running over the 64 columns of one head
      long long dot_base = 0;        // rowdot index of (image, head 0, row in image)
      if constexpr (EPI == MEMB_EPI_STORE_ROWDOT) {
        const int row = min(row0 + lane, p.M - 1), img = row / p.rows_per_group;
        dot_base = (long long)img * (p.N >> 6) * p.rows_per_group + (row - img * p.rows_per_group);
      }
      float rs = 1.0f;
      if constexpr (EPI == MEMB_EPI_RESIDUAL) {
        const int row = min(row0 + lane, p.M - 1);
        rs = p.rowscale ? __ldg(p.rowscale + row / p.rows_per_group) : 1.0f;
      }

      mbar_wait(&tmem_full_bar[acc], acc_phase, p.err_flag, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BLOCK_N + half * kCols;
      uint32_t r[2][32];
      tmem_ld32(taddr, r[0]);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        tmem_ld_wait();
        if (c + 1 < kChunks) {
          tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
        } else {  // the accumulator stage is drained: hand it back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tmem_empty_leader0 + acc * 8);
        }
        const uint32_t (&a)[32] = r[c & 1];
        const int col = colw + c * 32;
        const bool store_ok = live && col < p.N;
        const uint32_t bsm = vbs + c * 128;

        if constexpr (EPI == MEMB_EPI_STORE) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 b = ld_shared_v4(bsm + j * 16);
            pk[2 * j] = pack_bf16x2(__uint_as_float(a[4 * j]) + __uint_as_float(b.x), __uint_as_float(a[4 * j + 1]) + __uint_as_float(b.y));
            pk[2 * j + 1] = pack_bf16x2(__uint_as_float(a[4 * j + 2]) + __uint_as_float(b.z), __uint_as_float(a[4 * j + 3]) + __uint_as_float(b.w));
          }
          const uint32_t buf = stg + (ring & 3) * 2048;
          if (elect_one()) bulk_wait_read<3>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) st_shared_v4(sw64(buf, lane, j), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) {
            if (store_ok) tma_store_2d(&tmap_d, buf, col, row0);
            bulk_commit();
          }
        } else if constexpr (EPI == MEMB_EPI_BIAS_GELU) {
          uint32_t pa[16], pp[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 b = ld_shared_v4(bsm + j * 16);
            const float x0 = __uint_as_float(a[4 * j]) + __uint_as_float(b.x), x1 = __uint_as_float(a[4 * j + 1]) + __uint_as_float(b.y);
            const float x2 = __uint_as_float(a[4 * j + 2]) + __uint_as_float(b.z), x3 = __uint_as_float(a[4 * j + 3]) + __uint_as_float(b.w);
            pp[2 * j] = pack_bf16x2(x0, x1);
            pp[2 * j + 1] = pack_bf16x2(x2, x3);
            pa[2 * j] = pack_bf16x2(gelu_fwd(x0), gelu_fwd(x1));
            pa[2 * j + 1] = pack_bf16x2(gelu_fwd(x2), gelu_fwd(x3));
          }
          const uint32_t buf_a = stg + (ring & 1) * 2048, buf_p = stg + 4096 + (ring & 1) * 2048;
          if (elect_one()) bulk_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            st_shared_v4(sw64(buf_a, lane, j), pa[4 * j], pa[4 * j + 1], pa[4 * j + 2], pa[4 * j + 3]);
            st_shared_v4(sw64(buf_p, lane, j), pp[4 * j], pp[4 * j + 1], pp[4 * j + 2], pp[4 * j + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) {
            if (store_ok) {
              tma_store_2d(&tmap_d, buf_a, col, row0);
              if (p.has_d2) tma_store_2d(&tmap_d2, buf_p, col, row0);
            }
            bulk_commit();
          }
        } else if constexpr (EPI == MEMB_EPI_DGELU) {
          const uint32_t slot = ring & 3;
          const uint32_t buf = stg + slot * 2048;
          // prefetch the pre-activation two chunks ahead (its slot held the store of two chunks ago)
          if (c + 2 < kChunks) {
            if (elect_one()) {
              const uint32_t s2 = (ring + 2) & 3;
              bulk_wait_read<1>();
              fence_proxy_async();
              mbar_arrive_expect_tx(&abar[s2], 32 * 64);
              tma_load_2d_addr(stg + s2 * 2048, &tmap_aux, smem_u32(&abar[s2]), col + 64, row0);
            }
            __syncwarp();
          }
          mbar_wait(&abar[slot], (aux_phases >> slot) & 1u, p.err_flag, 5);
          aux_phases ^= 1u << slot;
          uint32_t pk[16];
          float cs[32];     // this lane's row of the chunk, kept for the column sums
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = ld_shared_v4(sw64(buf, lane, j));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float g0 = __uint_as_float(a[8 * j + 2 * t]) * gelu_grad(bf16_lo(w[t]));
              const float g1 = __uint_as_float(a[8 * j + 2 * t + 1]) * gelu_grad(bf16_hi(w[t]));
              pk[4 * j + t] = pack_bf16x2(g0, g1);
              cs[8 * j + 2 * t] = g0;
              cs[8 * j + 2 * t + 1] = g1;
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) st_shared_v4(sw64(buf, lane, j), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          if (p.colsum != nullptr) {
            // column sums over the warp's 32 rows: a butterfly that halves the values a lane holds at every step (31 shuffles);
            // lane l ends with column l, one 128-byte red.add per warp and chunk
            const bool row_ok = live && row0 + lane < p.M;
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
              const bool up = (lane & s) != 0;
#pragma unroll
              for (int k = 0; k < s; ++k) {
                float lo = cs[k], hi = cs[k + s];
                if (s == 16 && !row_ok) { lo = 0.f; hi = 0.f; }
                cs[k] = (up ? hi : lo) + __shfl_xor_sync(0xffffffffu, up ? lo : hi, s);
              }
            }
            if (col + lane < p.N) atomicAdd(p.colsum + col + lane, cs[0]);
          }
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) {
            if (store_ok) tma_store_2d(&tmap_d, buf, col, row0);
            bulk_commit();
          }
        } else if constexpr (EPI == MEMB_EPI_STORE_ROWDOT) {
          const uint32_t slot = ring & 3;
          const uint32_t buf = stg + slot * 2048;
          if (c + 2 < kChunks) {   // the aux tile two chunks ahead (its slot held the store of two chunks ago)
            if (elect_one()) {
              const uint32_t s2 = (ring + 2) & 3;
              bulk_wait_read<1>();
              fence_proxy_async();
              mbar_arrive_expect_tx(&abar[s2], 32 * 64);
              tma_load_2d_addr(stg + s2 * 2048, &tmap_aux, smem_u32(&abar[s2]), col + 64, row0);
            }
            __syncwarp();
          }
          mbar_wait(&abar[slot], (aux_phases >> slot) & 1u, p.err_flag, 5);
          aux_phases ^= 1u << slot;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = ld_shared_v4(sw64(buf, lane, j));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t v = pack_bf16x2(__uint_as_float(a[8 * j + 2 * t]), __uint_as_float(a[8 * j + 2 * t + 1]));
              pk[4 * j + t] = v;
              dot = fmaf(bf16_lo(v), bf16_lo(w[t]), dot);      // the stored (rounded) value times aux, fp32 accumulation
              dot = fmaf(bf16_hi(v), bf16_hi(w[t]), dot);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) st_shared_v4(sw64(buf, lane, j), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) {
            if (store_ok) tma_store_2d(&tmap_d, buf, col, row0);
            bulk_commit();
          }
          if (c & 1) {             // two 32-column chunks = one head
            if (store_ok && row0 + lane < p.M) p.rowdot[dot_base + (long long)(col >> 6) * p.rows_per_group] = dot;
            dot = 0.f;
          }
        } else if constexpr (EPI == MEMB_EPI_RESIDUAL) {
          const uint32_t slot = ring & 1;
          const uint32_t buf = stg + slot * 4096, buf2 = stg + 8192;
          mbar_wait(&abar[slot], (aux_phases >> slot) & 1u, p.err_flag, 5);
          aux_phases ^= 1u << slot;
          float o[32];
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 x = ld_shared_v4(sw128(buf, lane, j));
            const uint4 b = ld_shared_v4(bsm + j * 16);
            const uint4 g = ld_shared_v4(bsm + BLOCK_N * 4 + j * 16);
            const float v0 = __uint_as_float(a[4 * j]) + __uint_as_float(b.x), v1 = __uint_as_float(a[4 * j + 1]) + __uint_as_float(b.y);
            const float v2 = __uint_as_float(a[4 * j + 2]) + __uint_as_float(b.z), v3 = __uint_as_float(a[4 * j + 3]) + __uint_as_float(b.w);
            pk[2 * j] = pack_bf16x2(v0, v1);
            pk[2 * j + 1] = pack_bf16x2(v2, v3);
            o[4 * j] = fmaf(rs * __uint_as_float(g.x), v0, __uint_as_float(x.x));
            o[4 * j + 1] = fmaf(rs * __uint_as_float(g.y), v1, __uint_as_float(x.y));
            o[4 * j + 2] = fmaf(rs * __uint_as_float(g.z), v2, __uint_as_float(x.z));
            o[4 * j + 3] = fmaf(rs * __uint_as_float(g.w), v3, __uint_as_float(x.w));
          }
          // the previous chunk's stores must have left their buffers: then prefetch the next residual chunk
          if (elect_one()) {
            bulk_wait_read<0>();
            if (c + 1 < kChunks) {
              const uint32_t s1 = slot ^ 1;
              fence_proxy_async();
              mbar_arrive_expect_tx(&abar[s1], 32 * 128);
              tma_load_2d_addr(stg + s1 * 4096, &tmap_aux, smem_u32(&abar[s1]), col + 32, row0);
            }
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(sw128(buf, lane, j), __float_as_uint(o[4 * j]), __float_as_uint(o[4 * j + 1]), __float_as_uint(o[4 * j + 2]),
                         __float_as_uint(o[4 * j + 3]));
#pragma unroll
          for (int j = 0; j < 4; ++j) st_shared_v4(sw64(buf2, lane, j), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) {
            if (store_ok) {
              tma_store_2d(&tmap_d, buf, col, row0);
              if (p.has_d2) tma_store_2d(&tmap_d2, buf2, col, row0);
            }
            bulk_commit();
          }
        }
        ++ring;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (elect_one()) bulk_wait<0>();  // all staged tiles have been written out before the CTA may exit
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
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

// 2-D row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, box_cols].
static int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, long long rows, long long cols, long long ld,
                     int box_rows, int box_cols, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * elem_bytes)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MEMB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MEMB_OK;
}

static bool aligned16(const void* p, long long ld, int elem_bytes) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (ld * elem_bytes) % 16 == 0;
}

template <int BLOCK_N, bool B_MN, int EPI>
static int launch(const Params& p, const CUtensorMap* t, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, EPI>;
  auto kern = gemm_pair_kernel<BLOCK_N, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    MEMB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  const int clusters = std::max(1, std::min(p.num_tiles, num_sms() / 2));
  MEMB_CUDA_OK(launch_dependent(kern, dim3(2 * clusters), dim3(kThreads), (size_t)C::SMEM_BYTES, stream, t[0], t[1], t[2], t[3],
                                t[4], p));
  MEMB_LAUNCH_OK("gemm_pair_kernel");
  return MEMB_OK;
}

template <int BLOCK_N, bool B_MN>
static int dispatch_epi(int epi, const Params& p, const CUtensorMap* t, cudaStream_t s) {
  switch (epi) {
    case MEMB_EPI_STORE: return launch<BLOCK_N, B_MN, MEMB_EPI_STORE>(p, t, s);
    case MEMB_EPI_BIAS_GELU: return launch<BLOCK_N, B_MN, MEMB_EPI_BIAS_GELU>(p, t, s);
    case MEMB_EPI_RESIDUAL: return launch<BLOCK_N, B_MN, MEMB_EPI_RESIDUAL>(p, t, s);
    case MEMB_EPI_DGELU: return launch<BLOCK_N, B_MN, MEMB_EPI_DGELU>(p, t, s);
    case MEMB_EPI_STORE_ROWDOT:
      if constexpr (BLOCK_N == 192) return fail(MEMB_EINVAL, "gemm_pair: row dots need tiles of whole heads");
      else return launch<BLOCK_N, B_MN, MEMB_EPI_STORE_ROWDOT>(p, t, s);
    default: return fail(MEMB_EINVAL, "gemm_pair: unsupported epilogue %d", epi);
  }
}

// Tile width: fewest (waves x tile cost), slight preference for wider tiles (less L2 traffic per flop).
static int pick_block_n(int m_pairs, int n, bool b_mn, int forced) {
  const int clusters = std::max(1, num_sms() / 2);
  const int cand[3] = {256, 192, 128};
  const double penalty[3] = {1.0, 1.04, 1.08};
  int best = 256;
  double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (forced && forced != bn) continue;
    if (b_mn && bn == 192) continue;  // MN-major halves are loaded as 64-column boxes
    const long long tiles = (long long)m_pairs * ceil_div(n, bn);
    const double cost = (double)ceil_div<long long>(tiles, clusters) * bn * penalty[i];
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// Returns MEMB_OK and sets *handled when the descriptor was launched on the CTA-pair kernel.
int try_launch(const memb_gemm_desc& g, cudaStream_t stream, bool* handled) {
  *handled = false;
  static const bool disabled = [] { const char* e = std::getenv("MEMB_GEMM_SINGLE_CTA"); return e && e[0] == '1'; }();
  if (disabled) return MEMB_OK;
  if (g.in_dtype != MEMB_DT_BF16 || g.a_layout != 0 || g.split_precision) return MEMB_OK;
  const int epi = g.epilogue;
  if (!(epi == MEMB_EPI_STORE || epi == MEMB_EPI_BIAS_GELU || epi == MEMB_EPI_RESIDUAL || epi == MEMB_EPI_DGELU ||
        epi == MEMB_EPI_STORE_ROWDOT)) return MEMB_OK;
  if (g.m < 512 || g.n < 128 || g.n % 32 != 0 || g.k % 8 != 0) return MEMB_OK;
  if (g.act || g.out_split || g.rowmask || g.out_group_rows > 0 || g.alpha_dev || (g.alpha != 0.0f && g.alpha != 1.0f)) return MEMB_OK;
  const bool d_f32 = epi == MEMB_EPI_RESIDUAL;
  if ((g.out_dtype == MEMB_DT_F32) != d_f32) return MEMB_OK;
  if (epi == MEMB_EPI_STORE && g.aux) return MEMB_OK;
  if (!aligned16(g.a, g.lda, 2) || !aligned16(g.b, g.ldb, 2) || !aligned16(g.d, g.ldd, d_f32 ? 4 : 2)) return MEMB_OK;
  if (g.d2 && !aligned16(g.d2, g.ldd2, 2)) return MEMB_OK;
  if (epi == MEMB_EPI_RESIDUAL && (!g.aux || !aligned16(g.aux, g.ldaux, 4))) return MEMB_OK;
  if ((epi == MEMB_EPI_DGELU || epi == MEMB_EPI_STORE_ROWDOT) && (!g.aux || !aligned16(g.aux, g.ldaux, 2))) return MEMB_OK;
  if (epi == MEMB_EPI_STORE_ROWDOT && (!g.rowdot || g.n % 64 != 0 || g.rows_per_group <= 0)) return MEMB_OK;
  if (epi == MEMB_EPI_RESIDUAL && g.rowscale && g.rows_per_group <= 0) return MEMB_OK;

  const bool b_mn = g.b_layout != 0;
  const int m_pairs = ceil_div(g.m, 2 * BLOCK_M);
  // (192-wide tiles split a warp's columns at 96: not whole heads for the row dots; MN-major B halves come as 64-column boxes)
  const int bn = pick_block_n(m_pairs, g.n, b_mn || epi == MEMB_EPI_STORE_ROWDOT,
                              (g.block_n == 128 || g.block_n == 192 || g.block_n == 256) ? g.block_n : 0);
  Params p{};
  p.M = g.m; p.N = g.n; p.K = g.k;
  p.num_n_tiles = ceil_div(g.n, bn);
  p.num_tiles = m_pairs * p.num_n_tiles;
  p.kb_total = ceil_div(g.k, BLOCK_K);
  p.bias = g.bias; p.colscale = g.colscale; p.rowscale = g.rowscale; p.rows_per_group = std::max(1, g.rows_per_group);
  p.has_d2 = g.d2 != nullptr;
  p.err_flag = g.err_flag;
  p.colsum = (epi == MEMB_EPI_DGELU) ? g.colsum : nullptr;
  p.rowdot = (epi == MEMB_EPI_STORE_ROWDOT) ? g.rowdot : nullptr;

  CUtensorMap t[5];
  if (int rc = make_tmap(&t[0], g.a, 2, g.m, g.k, g.lda, BLOCK_M, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (!b_mn) {
    if (int rc = make_tmap(&t[1], g.b, 2, g.n, g.k, g.ldb, bn / 2, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  } else {
    if (int rc = make_tmap(&t[1], g.b, 2, g.k, g.n, g.ldb, BLOCK_K, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  if (d_f32) {
    if (int rc = make_tmap(&t[2], g.d, 4, g.m, g.n, g.ldd, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  } else {
    if (int rc = make_tmap(&t[2], g.d, 2, g.m, g.n, g.ldd, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  t[3] = t[2];
  t[4] = t[2];
  if (g.d2) {
    if (int rc = make_tmap(&t[3], g.d2, 2, g.m, g.n, g.ldd2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  if (epi == MEMB_EPI_RESIDUAL) {
    if (int rc = make_tmap(&t[4], g.aux, 4, g.m, g.n, g.ldaux, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  } else if (epi == MEMB_EPI_DGELU || epi == MEMB_EPI_STORE_ROWDOT) {
    if (int rc = make_tmap(&t[4], g.aux, 2, g.m, g.n, g.ldaux, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }

  int rc;
  if (b_mn) rc = (bn == 256) ? dispatch_epi<256, true>(epi, p, t, stream) : dispatch_epi<128, true>(epi, p, t, stream);
  else rc = (bn == 256) ? dispatch_epi<256, false>(epi, p, t, stream)
          : (bn == 192) ? dispatch_epi<192, false>(epi, p, t, stream) : dispatch_epi<128, false>(epi, p, t, stream);
  if (rc == MEMB_OK) *handled = true;
  return rc;
}

}  // namespace gemm_pair
}  // namespace memb
