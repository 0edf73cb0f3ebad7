// HBM-bound kernels of the masked-ViT step (everything that is not a GEMM or attention):
// LayerNorm fwd/bwd (+ residual-gradient accumulate, + masked-row gather/scatter), LayerScale /
// DropPath branch backward, bias column sums, patchify, cls/pos assembly, mask compaction,
// 8192-way cross-entropy (loss, accuracy, dlogits in one pass), relative-position-bias gather /
// scatter, and the flat-buffer grad-norm + clip + AdamW pass.
//
// Reference semantics: mem/modeling_finetune.py:166-189 (Block), :203-247 (PatchEmbed,
// RelativePositionBias); mem/modeling_pretrain.py:97-126; mem/engine_for_pretraining.py:152,233
// (CrossEntropyLoss, mlm_acc); mem/utils.py:357-371 (clip + step); torch.optim.AdamW.
#include <algorithm>

#include <cuda_bf16.h>

#include "common.cuh"

namespace memb {
namespace vit {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4_bf16(const bf16* p) {
  uint2 q = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4_bf16(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 q;
  q.x = *reinterpret_cast<uint32_t*>(&a);
  q.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = q;
}

constexpr int kMaxChunks = 16;  // D <= 2048, D % 128 == 0: lane owns float4 chunks (c*32 + lane)*4, CH = D/128
constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;

// Reduce per-lane column partials across the warps of a block, then one atomicAdd per column.
template <int NV, int CH>
__device__ __forceinline__ void block_column_atomic(float (&acc)[NV][CH][4], int D,
                                                    float* const (&dst)[NV], float* smem /*[kRowWarps][D]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (dst[v] == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = (c * 32 + lane) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) smem[warp * D + col + j] = acc[v][c][j];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < D; col += kRowThreads) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) s += smem[w * D + col];
      atomicAdd(dst[v] + col, s);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm forward
// One warp per row.  Optional gather: source row = row_index[r]; rows >= *count produce zeros.
template <int CH>
__global__ void __launch_bounds__(kRowThreads) layernorm_fwd(const float* __restrict__ x, long long ldx,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int rows, int D,
                                                             bf16* __restrict__ y, long long ldy, float* __restrict__ mean,
                                                             float* __restrict__ rstd, const int* __restrict__ row_index,
                                                             const int* __restrict__ count) {
  grid_dependency_trigger();   // a GEMM launched as a programmatic dependent sets itself up under this kernel's tail
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int live = count ? min(rows, *count) : rows;
  for (int r = blockIdx.x * kRowWarps + warp; r < rows; r += gridDim.x * kRowWarps) {
    bf16* yr = y + (long long)r * ldy;
    if (r >= live) {
      for (int c = 0; c < CH; ++c) st4_bf16(yr + (c * 32 + lane) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      if (lane == 0) { mean[r] = 0.f; rstd[r] = 0.f; }
      continue;
    }
    const float* xr = x + (long long)(row_index ? row_index[r] : r) * ldx;
    float4 v[CH];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
      { v[c] = ld4(xr + (c * 32 + lane) * 4); s += v[c].x + v[c].y + v[c].z + v[c].w; }
    const float mu = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
      {
        const float a = v[c].x - mu, b = v[c].y - mu, cc = v[c].z - mu, d = v[c].w - mu;
        q += a * a + b * b + cc * cc + d * d;
      }
    const float rs = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
    for (int c = 0; c < CH; ++c)
      {
        const int col = (c * 32 + lane) * 4;
        const float4 g = ld4(gamma + col), b = ld4(beta + col);
        st4_bf16(yr + col, make_float4((v[c].x - mu) * rs * g.x + b.x, (v[c].y - mu) * rs * g.y + b.y,
                                       (v[c].z - mu) * rs * g.z + b.z, (v[c].w - mu) * rs * g.w + b.w));
      }
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// dx[target row] += LNbwd(dy);  dgamma += sum dy*xhat;  dbeta += sum dy.   dy is bf16 or fp32.
// The column sums live in per-warp shared-memory rows (plain read-modify-write, a lane owns its columns), not in
// registers: 48 fewer registers per thread lets two blocks share an SM, which is what this HBM-bound kernel needs
// (16 warps x ~4.5 KB of loads in flight instead of 8).
// kBranch: the branch backward of the NEXT sub-block (see branch_bwd below) runs on the freshly updated residual gradient
// row while it is still in registers -- dz = rowscale*colscale*dx (bf16), dcolscale += sum rowscale*dx*branch, dbias += sum dz
// -- which saves that kernel's read of the fp32 gradient (77 MB per sub-block at ViT-B/16, batch 128) and its launch.
struct BranchArgs {
  const bf16* branch; long long ldb;
  const float* colscale; const float* rowscale; int rows_per_group;
  bf16* dz; long long lddz;
  float* dcolscale; float* dbias;
};

template <typename DyT, int CH, bool kBranch>
__global__ void __launch_bounds__(kRowThreads, 2) layernorm_bwd(const DyT* __restrict__ dy, long long lddy,
                                                                const float* __restrict__ x, long long ldx,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, int rows, int D,
                                                                float* __restrict__ dx, long long lddx,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                const int* __restrict__ row_index,
                                                                const int* __restrict__ count, const BranchArgs br) {
  grid_dependency_trigger();
  extern __shared__ float smem[];                 // [2 (+2)][kRowWarps][D]: dgamma, dbeta (, dcolscale, dbias) partials
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int live = count ? min(rows, *count) : rows;
  float* sg = smem + (size_t)warp * D;
  float* sb = smem + (size_t)(kRowWarps + warp) * D;
  float* sc = smem + (size_t)(2 * kRowWarps + warp) * D;
  float* sz = smem + (size_t)(3 * kRowWarps + warp) * D;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int col = (c * 32 + lane) * 4;
    st4(sg + col, make_float4(0.f, 0.f, 0.f, 0.f));
    st4(sb + col, make_float4(0.f, 0.f, 0.f, 0.f));
    if constexpr (kBranch) {
      st4(sc + col, make_float4(0.f, 0.f, 0.f, 0.f));
      st4(sz + col, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }

  for (int r = blockIdx.x * kRowWarps + warp; r < live; r += gridDim.x * kRowWarps) {
    const long long src = row_index ? row_index[r] : r;
    const float* xr = x + src * ldx;
    const DyT* dyr = dy + (long long)r * lddy;
    float* dxr = dx + src * lddx;
    const float mu = mean[r], rs = rstd[r];
    float4 xh[CH], dg[CH], o[CH];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = (c * 32 + lane) * 4;
      float4 d;
      if constexpr (sizeof(DyT) == 2) d = ld4_bf16(reinterpret_cast<const bf16*>(dyr) + col);
      else d = ld4(reinterpret_cast<const float*>(dyr) + col);
      const float4 xv = ld4(xr + col), g = ld4(gamma + col);
      o[c] = ld4(dxr + col);                      // requested early: only needed after the two warp sums
      xh[c] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      float4 a = ld4(sg + col), b = ld4(sb + col);
      a.x += d.x * xh[c].x; a.y += d.y * xh[c].y; a.z += d.z * xh[c].z; a.w += d.w * xh[c].w;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
      st4(sg + col, a);
      st4(sb + col, b);
      dg[c] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      s1 += dg[c].x + dg[c].y + dg[c].z + dg[c].w;
      s2 += dg[c].x * xh[c].x + dg[c].y * xh[c].y + dg[c].z * xh[c].z + dg[c].w * xh[c].w;
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = (c * 32 + lane) * 4;
      o[c].x += rs * (dg[c].x - s1 - xh[c].x * s2);
      o[c].y += rs * (dg[c].y - s1 - xh[c].y * s2);
      o[c].z += rs * (dg[c].z - s1 - xh[c].z * s2);
      o[c].w += rs * (dg[c].w - s1 - xh[c].w * s2);
      st4(dxr + col, o[c]);
    }
    if constexpr (kBranch) {
      const float brs = br.rowscale ? br.rowscale[r / br.rows_per_group] : 1.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int col = (c * 32 + lane) * 4;
        const float4 cs = br.colscale ? ld4(br.colscale + col) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 z = make_float4(brs * cs.x * o[c].x, brs * cs.y * o[c].y, brs * cs.z * o[c].z, brs * cs.w * o[c].w);
        st4_bf16(br.dz + (long long)r * br.lddz + col, z);
        if (br.dcolscale) {
          const float4 b = ld4_bf16(br.branch + (long long)r * br.ldb + col);
          float4 a = ld4(sc + col);
          a.x += brs * o[c].x * b.x; a.y += brs * o[c].y * b.y; a.z += brs * o[c].z * b.z; a.w += brs * o[c].w * b.w;
          st4(sc + col, a);
        }
        float4 zb = ld4(sz + col);
        zb.x += z.x; zb.y += z.y; zb.z += z.z; zb.w += z.w;
        st4(sz + col, zb);
      }
    }
  }
  __syncthreads();
  for (int col = threadIdx.x; col < D; col += kRowThreads) {
    float g = 0.f, b = 0.f, cs = 0.f, z = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) {
      g += smem[(size_t)w * D + col];
      b += smem[(size_t)(kRowWarps + w) * D + col];
      if constexpr (kBranch) {
        cs += smem[(size_t)(2 * kRowWarps + w) * D + col];
        z += smem[(size_t)(3 * kRowWarps + w) * D + col];
      }
    }
    if (dgamma) atomicAdd(dgamma + col, g);
    if (dbeta) atomicAdd(dbeta + col, b);
    if constexpr (kBranch) {
      if (br.dcolscale) atomicAdd(br.dcolscale + col, cs);
      if (br.dbias) atomicAdd(br.dbias + col, z);
    }
  }
}

// ------------------------------------------------------------------ branch backward
// Forward was  x_out = x_in + rowscale[row/g] * colscale[n] * branch[row,n]  (LayerScale x DropPath).
// dz = rowscale*colscale*gout (bf16);  dcolscale += sum rowscale*gout*branch;  dbias += sum dz.
template <int CH>
__global__ void __launch_bounds__(kRowThreads) branch_bwd(const float* __restrict__ gout, long long ldg,
                                                          const bf16* __restrict__ branch, long long ldb,
                                                          const float* __restrict__ colscale,
                                                          const float* __restrict__ rowscale, int rows_per_group,
                                                          int rows, int D, bf16* __restrict__ dz, long long lddz,
                                                          float* __restrict__ dcolscale, float* __restrict__ dbias) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[2][CH][4];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][c][j] = acc[1][c][j] = 0.f;
  for (int r = blockIdx.x * kRowWarps + warp; r < rows; r += gridDim.x * kRowWarps) {
    const float rs = rowscale ? rowscale[r / rows_per_group] : 1.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
      {
        const int col = (c * 32 + lane) * 4;
        const float4 g = ld4(gout + (long long)r * ldg + col);
        const float4 cs = colscale ? ld4(colscale + col) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 z = make_float4(rs * cs.x * g.x, rs * cs.y * g.y, rs * cs.z * g.z, rs * cs.w * g.w);
        st4_bf16(dz + (long long)r * lddz + col, z);
        if (dcolscale) {
          const float4 b = ld4_bf16(branch + (long long)r * ldb + col);
          acc[0][c][0] += rs * g.x * b.x; acc[0][c][1] += rs * g.y * b.y; acc[0][c][2] += rs * g.z * b.z; acc[0][c][3] += rs * g.w * b.w;
        }
        acc[1][c][0] += z.x; acc[1][c][1] += z.y; acc[1][c][2] += z.z; acc[1][c][3] += z.w;
      }
  }
  float* const dst[2] = {dcolscale, dbias};
  block_column_atomic<2, CH>(acc, D, dst, smem);
}

// v_bias gradient without a pass over dV.  The rows of a softmax sum to one, so  sum_keys dV[key, :] = sum_queries dAO[query, :]
// (dV = P^T dAO), and dAO = dZ W_proj, so  colsum(dV) = colsum(dZ) W_proj  where colsum(dZ) is the proj bias gradient of
// THIS backward pass (t, accumulated by the branch backward into a scratch vector): dv_bias[j] += sum_i t[i] W[i, j]  (W the
// fp32 master weight [D, D], row i = output feature i), dproj_bias[i] += t[i].  Block (x, y): 256 columns j, rows i of slice y.
constexpr int kChainSlices = 16;
__global__ void __launch_bounds__(256) vbias_chain(const float* __restrict__ t, const float* __restrict__ W, int D,
                                                   float* __restrict__ dproj_bias, float* __restrict__ dv_bias) {
  grid_dependency_trigger();
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int per = (D + kChainSlices - 1) / kChainSlices;
  const int i0 = blockIdx.y * per, i1 = min(D, i0 + per);
  if (j < D) {
    float acc = 0.f;
#pragma unroll 4
    for (int i = i0; i < i1; ++i) acc = fmaf(__ldg(t + i), __ldg(W + (long long)i * D + j), acc);
    atomicAdd(dv_bias + j, acc);
    if (blockIdx.y == 0) atomicAdd(dproj_bias + j, __ldg(t + j));
  }
}

// out[n] += sum_rows x[row, n]   (bias gradients of fc1 / qkv); N % 8 == 0, rows 16-byte aligned.
// Block = 32 column lanes (8 columns = one 16-byte load each) x 8 row phases; every thread keeps 8 independent loads
// in flight, the 8 phases meet in shared memory and one atomic per column leaves the block.
constexpr int kColsumRows = 8;
__global__ void __launch_bounds__(32 * kColsumRows) colsum_bf16(const bf16* __restrict__ x, long long ld, int rows, int N,
                                                                int rows_per_block, float* __restrict__ out) {
  grid_dependency_trigger();
  __shared__ float part[kColsumRows][32][8];
  const int col = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    const bf16* p = x + col;
    int r = r0 + threadIdx.y;
    for (; r + 7 * kColsumRows < r1; r += 8 * kColsumRows) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const uint4*>(p + (long long)(r + u * kColsumRows) * ld));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(h[k]);
          s[2 * k] += f.x;
          s[2 * k + 1] += f.y;
        }
      }
    }
    for (; r < r1; r += kColsumRows) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p + (long long)r * ld));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        s[2 * k] += f.x;
        s[2 * k + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) part[threadIdx.y][threadIdx.x][k] = s[k];
  __syncthreads();
  // 256 threads -> the block's 256 columns
  const int t = threadIdx.y * 32 + threadIdx.x, c = blockIdx.x * 256 + t;
  if (c < N) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < kColsumRows; ++g) a += part[g][t >> 3][t & 7];
    atomicAdd(out + c, a);
  }
}

// ------------------------------------------------------------------ patch embedding helpers
// img fp32 [B,C,H,W] -> rows bf16 [B*(H/P)*(W/P), C*P*P], column = c*P*P + i*P + j (Conv2d weight order).
__global__ void __launch_bounds__(256) patchify(const float* __restrict__ img, int B, int C, int H, int W, int P,
                                                bf16* __restrict__ out) {
  const int per_row = P / 8;  // 8 output elements per thread
  const long long total = (long long)B * C * H * W / 8;
  const int gw = W / P, gh = H / P;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int j8 = (int)(r % per_row); r /= per_row;
    const int i = (int)(r % P); r /= P;
    const int c = (int)(r % C); r /= C;
    const int px = (int)(r % gw); r /= gw;
    const int py = (int)(r % gh); r /= gh;
    const int b = (int)r;
    const float* src = img + (((long long)b * C + c) * H + (py * P + i)) * W + px * P + j8 * 8;
    const float4 v0 = ld4(src), v1 = ld4(src + 4);
    bf16* dst = out + t * 8;
    st4_bf16(dst, v0);
    st4_bf16(dst + 4, v1);
  }
}

// x[b,0,:] = cls (+ pos[0]);  x[b,1+p,:] += pos[1+p] when pos != null.
__global__ void __launch_bounds__(256) cls_pos(float* __restrict__ x, const float* __restrict__ cls,
                                               const float* __restrict__ pos, int B, int N, int D) {
  const long long total = pos ? (long long)B * N * D : (long long)B * D;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    if (pos) {
      const int d = (int)(t % D);
      const int n = (int)((t / D) % N);
      if (n == 0) x[t] = cls[d] + pos[d];
      else x[t] += pos[(long long)n * D + d];
    } else {
      const int d = (int)(t % D);
      const long long b = t / D;
      x[b * N * D + d] = cls[d];
    }
  }
}

// Backward of the embedding assembly.  g0 fp32 [B,N,D] (N = P+1), mask u8 [B,P]:
//   dpatch[b*P+p,:] = mask ? 0 : g0[b,1+p,:] (bf16);  dmask_token += masked rows;  dcls += g0[b,0,:];
//   dbias += unmasked rows;  dpos[n,:] += g0[b,n,:] when dpos != null.
template <int CH>
__global__ void __launch_bounds__(kRowThreads) embed_bwd(const float* __restrict__ g0, const unsigned char* __restrict__ mask,
                                                         int B, int P, int D, bf16* __restrict__ dpatch,
                                                         float* __restrict__ dmask_token, float* __restrict__ dcls,
                                                         float* __restrict__ dbias, float* __restrict__ dpos) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = P + 1;
  float acc[3][CH][4];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][c][j] = acc[1][c][j] = acc[2][c][j] = 0.f;
  const int rows = B * N;
  for (int r = blockIdx.x * kRowWarps + warp; r < rows; r += gridDim.x * kRowWarps) {
    const int b = r / N, n = r % N;
    const bool is_cls = n == 0;
    const bool masked = !is_cls && mask[b * P + n - 1] != 0;
#pragma unroll
    for (int c = 0; c < CH; ++c)
      {
        const int col = (c * 32 + lane) * 4;
        const float4 g = ld4(g0 + (long long)r * D + col);
        const int which = is_cls ? 1 : (masked ? 0 : 2);
        acc[which][c][0] += g.x; acc[which][c][1] += g.y; acc[which][c][2] += g.z; acc[which][c][3] += g.w;
        if (!is_cls) st4_bf16(dpatch + ((long long)b * P + n - 1) * D + col, masked ? make_float4(0.f, 0.f, 0.f, 0.f) : g);
        if (dpos) {
          atomicAdd(dpos + (long long)n * D + col, g.x); atomicAdd(dpos + (long long)n * D + col + 1, g.y);
          atomicAdd(dpos + (long long)n * D + col + 2, g.z); atomicAdd(dpos + (long long)n * D + col + 3, g.w);
        }
      }
  }
  float* const dst[3] = {dmask_token, dcls, dbias};
  block_column_atomic<3, CH>(acc, D, dst, smem);
}

// ------------------------------------------------------------------ mask compaction
// mask u8 [B,P] -> row_index[i] = b*(P+1)+1+p and patch_index[i] = b*P+p of the i-th masked patch in
// row-major (b,p) order (the order boolean indexing produces, modeling_pretrain.py:126 /
// engine_for_pretraining.py:145); *count = number of masked patches (clamped to cap).
__global__ void __launch_bounds__(1024) mask_compact(const unsigned char* __restrict__ mask, int total, int P,
                                                     int* __restrict__ row_index, int* __restrict__ patch_index,
                                                     int* __restrict__ count, int cap) {
  __shared__ int part[1024];
  const int per = (total + 1023) / 1024;
  const int lo = min(total, (int)threadIdx.x * per), hi = min(total, lo + per);
  int mine = 0;
  for (int i = lo; i < hi; ++i) mine += mask[i] != 0;
  part[threadIdx.x] = mine;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int pos = part[threadIdx.x] - mine;
  for (int i = lo; i < hi; ++i)
    if (mask[i] != 0) {
      if (pos < cap) {
        const int b = i / P, p = i % P;
        row_index[pos] = b * (P + 1) + 1 + p;
        patch_index[pos] = i;
      }
      ++pos;
    }
  if (threadIdx.x == 1023) *count = min(part[1023], cap);
  // deterministic tail for rows >= count
  __syncthreads();
  const int n = min(part[1023], cap);
  for (int i = n + threadIdx.x; i < cap; i += 1024) { row_index[i] = 0; patch_index[i] = 0; }
}

// ------------------------------------------------------------------ cross entropy (8192-way)
// One block per row.  logits fp32 [cap,V] (row stride ld).  Writes dlogits = (softmax - onehot) *
// grad_scale / count (bf16, zero rows beyond count) and accumulates stats[0] += loss_i,
// stats[1] += (argmax == label), stats[2] = count.
template <int kPerThread>
__global__ void __launch_bounds__(256) cross_entropy(const float* __restrict__ logits, long long ld,
                                                     const long long* __restrict__ tokens,
                                                     const int* __restrict__ patch_index,
                                                     const int* __restrict__ count, int V, bf16* __restrict__ dlogits,
                                                     long long ldd, float* __restrict__ stats, float grad_scale) {
  __shared__ float red[8];
  __shared__ int redi[8];
  const int r = blockIdx.x;
  const int live = *count;
  bf16* dr = dlogits ? dlogits + (long long)r * ldd : nullptr;
  if (r >= live) {
    if (dr)
      for (int c = threadIdx.x * 4; c < V; c += 256 * 4) st4_bf16(dr + c, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  const float* lr = logits + (long long)r * ld;
  float4 v[kPerThread];
  float mx = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int i = 0; i < kPerThread; ++i) {
    const int c = (i * 256 + threadIdx.x) * 4;
    if (c < V) {
      v[i] = ld4(lr + c);
      if (v[i].x > mx) { mx = v[i].x; arg = c; }
      if (v[i].y > mx) { mx = v[i].y; arg = c + 1; }
      if (v[i].z > mx) { mx = v[i].z; arg = c + 2; }
      if (v[i].w > mx) { mx = v[i].w; arg = c + 3; }
    }
  }
  // block arg-max (first index wins ties)
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mx; redi[threadIdx.x >> 5] = arg; }
  __syncthreads();
  mx = red[0]; arg = redi[0];
#pragma unroll
  for (int w = 1; w < 8; ++w)
    if (red[w] > mx || (red[w] == mx && redi[w] < arg)) { mx = red[w]; arg = redi[w]; }
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kPerThread; ++i) {
    const int c = (i * 256 + threadIdx.x) * 4;
    if (c < V) {
      v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx); v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  const int label = (int)tokens[patch_index[r]];
  if (threadIdx.x == 0) {
    const float loss = logf(s) + mx - lr[label];
    atomicAdd(stats, loss);
    atomicAdd(stats + 1, arg == label ? 1.f : 0.f);
    stats[2] = (float)live;
  }
  if (dr) {
    const float k = grad_scale / (float)live, inv = 1.f / s;
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int c = (i * 256 + threadIdx.x) * 4;
      if (c < V) {
        float4 g = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
        if (label == c) g.x -= 1.f; else if (label == c + 1) g.y -= 1.f; else if (label == c + 2) g.z -= 1.f; else if (label == c + 3) g.w -= 1.f;
        st4_bf16(dr + c, make_float4(g.x * k, g.y * k, g.z * k, g.w * k));
      }
    }
  }
}

// ------------------------------------------------------------------ relative position bias
// bias[h][q][k] = table[index[q*N+k]][h]; also the transposed copy biasT[h][k][q] for attention backward.
__global__ void __launch_bounds__(256) relpos_gather(const float* __restrict__ table, const long long* __restrict__ index,
                                                     int N, int heads, int ldk, float* __restrict__ bias,
                                                     float* __restrict__ biasT) {
  const int total = N * ldk * heads;  // padded columns [N, ldk) are written as zeros
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int h = t / (N * ldk), qk = t % (N * ldk), q = qk / ldk, k = qk % ldk;
    const float v = k < N ? table[index[q * N + k] * heads + h] : 0.f;
    bias[t] = v;
    if (biasT) biasT[t] = k < N ? table[index[k * N + q] * heads + h] : 0.f;  // biasT[h][q'][k'] = bias[h][k'][q']
  }
}
// Dense bias [H][N][ld] -> the attention kernels' packed layout [H][52][256][4] (head, column group, row, 4 columns),
// multiplied by log2(e); columns >= N hold -inf (they mask the padded keys), rows >= N hold 0.
__global__ void __launch_bounds__(256) attn_bias_pack(const float* __restrict__ dense, int ld, int N, int heads,
                                                      float4* __restrict__ packed) {
  const int total = heads * MEMB_ATTN_BIAS_GROUPS * 256;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int r = t & 255, g = (t >> 8) % MEMB_ATTN_BIAS_GROUPS, h = t / (256 * MEMB_ATTN_BIAS_GROUPS);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * g + j;
      v[j] = c >= N ? -INFINITY : (r < N ? dense[((long long)h * N + r) * ld + c] * 1.4426950408889634f : 0.f);
    }
    packed[t] = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// dtable[index[q*N+k]][h] += dbias[h][q][k (row stride ldk)]
__global__ void __launch_bounds__(256) relpos_scatter(const float* __restrict__ dbias, int ldk,
                                                      const long long* __restrict__ index, int N, int heads,
                                                      float* __restrict__ dtable) {
  const int total = N * N * heads;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int h = t / (N * N), qk = t % (N * N), q = qk / N, k = qk % N;
    atomicAdd(dtable + index[qk] * heads + h, dbias[((long long)h * N + q) * ldk + k]);
  }
}
// acc[i] += sum_b x[b*inner + i]   (dS summed over the batch)
__global__ void __launch_bounds__(256) batch_reduce_bf16(const bf16* __restrict__ x, int B, long long inner,
                                                         float* __restrict__ acc) {
  const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  if (i >= inner) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float4 v = ld4_bf16(x + b * inner + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float4 a = ld4(acc + i);
  st4(acc + i, make_float4(a.x + s.x, a.y + s.y, a.z + s.z, a.w + s.w));
}


// ------------------------------------------------------------------ classification head (ft_vit, config 5)
// VisionTransformer.forward_features with mean pooling (modeling_finetune.py:343-352): pooled[b] = mean of the patch
// tokens x[b, 1:, :]; the backward spreads dpooled / P over the patch rows (the cls row gets zero).
__global__ void __launch_bounds__(256) meanpool_fwd(const float* __restrict__ x, int B, int N, int D, float* __restrict__ out) {
  const int b = blockIdx.y, col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (col >= D) return;
  const float* xb = x + ((long long)b * N + 1) * D + col;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < N - 1; ++t) {
    const float4 v = ld4(xb + (long long)t * D);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const float inv = 1.f / (float)(N - 1);
  st4(out + (long long)b * D + col, make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv));
}
__global__ void __launch_bounds__(256) meanpool_bwd(const float* __restrict__ dpool, int B, int N, int D, float* __restrict__ gres) {
  const long long total = (long long)B * N * D / 4;
  const float inv = 1.f / (float)(N - 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4, row = e / D;
    const int col = (int)(e % D), t = (int)(row % N), b = (int)(row / N);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t > 0) { v = ld4(dpool + (long long)b * D + col); v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv; }
    st4(gres + e, v);
  }
}
// Small dense head (nn.Linear D -> C, C from 2 to ~1000): fp32 CUDA-core kernels, one warp per output.
__global__ void __launch_bounds__(256) linear_small_fwd(const bf16* __restrict__ z, const float* __restrict__ W,
                                                        const float* __restrict__ bias, int B, int D, int C,
                                                        float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * C) return;
  const int b = warp / C, c = warp % C;
  float s = 0.f;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 a = ld4_bf16(z + (long long)b * D + d), w = ld4(W + (long long)c * D + d);
    s += a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
  }
  s = warp_sum(s);
  if (lane == 0) out[(long long)b * C + c] = s + (bias ? bias[c] : 0.f);
}
// dW[c,d] += sum_b dl[b,c] z[b,d];  db[c] += sum_b dl[b,c];  dz[b,d] = sum_c dl[b,c] W[c,d]
__global__ void __launch_bounds__(256) linear_small_bwd(const float* __restrict__ dl, const bf16* __restrict__ z,
                                                        const float* __restrict__ W, int B, int D, int C,
                                                        float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dz) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n_w = (long long)C * D, n_z = (long long)B * D;
  if (i < n_w) {
    const int c = (int)(i / D), d = (int)(i % D);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dl[(long long)b * C + c] * __bfloat162float(z[(long long)b * D + d]);
    dW[i] += s;
    if (d == 0 && db) {
      float t = 0.f;
      for (int b = 0; b < B; ++b) t += dl[(long long)b * C + c];
      db[c] += t;
    }
  } else if (i < n_w + n_z) {
    const long long j = i - n_w;
    const int b = (int)(j / D), d = (int)(j % D);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += dl[(long long)b * C + c] * W[(long long)c * D + d];
    dz[j] = s;
  }
}

// ------------------------------------------------------------------ optimizer
__global__ void __launch_bounds__(256) fill_f32(float* __restrict__ p, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void __launch_bounds__(256) cast_bf16(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}
// out[0] += sum g^2 * scale^2
__global__ void __launch_bounds__(256) sqnorm(const float* __restrict__ g, long long n, float scale, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i] * scale;
    s += v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}
// ------------------------------------------------------------------ classification criteria (finetuning)
// Soft-target cross entropy on [B, C] logits, one warp per row: loss_i = sum_c t_c * (lse_i - x_ic) with
//   t = soft[i, :]                                            (timm SoftTargetCrossEntropy: mixup / cutmix targets)
//   t = (1 - eps) * onehot(label_i) + eps / C                 (timm LabelSmoothingCrossEntropy; eps = 0: CrossEntropyLoss)
// (the criteria mem/run_class_finetuning.py:551-559 picks).  Accumulates loss_out[0] += mean_i loss_i and writes
// dlogits[i, c] = (softmax_ic * sum_c t_c - t_c) / B, the gradient of that mean.
__global__ void __launch_bounds__(128) soft_ce(const float* __restrict__ logits, int B, int C, const long long* __restrict__ labels,
                                               const float* __restrict__ soft, float eps, float* __restrict__ loss_out,
                                               float* __restrict__ dlogits) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* x = logits + (long long)row * C;
  const float* t = soft ? soft + (long long)row * C : nullptr;
  const int label = labels ? (int)labels[row] : -1;
  const float base = eps / (float)C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f, st = 0.f, stx = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float xv = x[c];
    const float tv = t ? t[c] : (base + (c == label ? 1.f - eps : 0.f));
    se += expf(xv - mx);
    st += tv;
    stx += tv * (xv - mx);
  }
  se = warp_sum(se); st = warp_sum(st); stx = warp_sum(stx);
  const float lse = logf(se);
  const float inv_b = 1.f / (float)B;
  if (lane == 0) atomicAdd(loss_out, (st * lse - stx) * inv_b);
  if (dlogits) {
    float* d = dlogits + (long long)row * C;
    for (int c = lane; c < C; c += 32) {
      const float tv = t ? t[c] : (base + (c == label ? 1.f - eps : 0.f));
      d[c] = (expf(x[c] - mx - lse) * st - tv) * inv_b;
    }
  }
}

// Same over a flat parameter-gradient buffer laid out in 1024-element chunks, skipping chunks whose group is 255
// (padding / requires_grad = False parameters: clip_grad_norm_ only sees trainable tensors, mem/utils.py:380-392).
__global__ void __launch_bounds__(256) sqnorm_groups(const float* __restrict__ g, long long n, float scale,
                                                     const unsigned char* __restrict__ chunk_group, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  const long long nchunks = (n + 1023) / 1024;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    if (chunk_group[ch] == 255) continue;
    const long long i = ch * 1024 + threadIdx.x * 4;
    if (i >= n) continue;
    const float4 v = ld4(g + i);
    s += (v.x * scale) * (v.x * scale) + (v.y * scale) * (v.y * scale) + (v.z * scale) * (v.z * scale) + (v.w * scale) * (v.w * scale);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}
// AdamW over the whole flat parameter buffer (torch.optim.AdamW semantics).  Tensors are laid out on
// 1024-element boundaries; chunk_group[i / 1024] selects the parameter group (lr, weight decay), >= 64 =
// never updated (255: padding / frozen, 254: the status chunk, counted by sqnorm_groups).  The global-norm clip
// coefficient is applied on the fly and the bf16 shadow copy of the weights is refreshed in the same pass.
// A step whose gradient norm is not finite is skipped entirely (parameters and moments untouched): what the
// reference's GradScaler.step does with inf / NaN gradients (mem/utils.py:357-371).
struct AdamGroups { float lr[64]; float wd[64]; };
__global__ void __launch_bounds__(256) adamw(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                             float* __restrict__ v, bf16* __restrict__ shadow, long long n,
                                             const unsigned char* __restrict__ chunk_group, const AdamGroups groups,
                                             float beta1, float beta2, float eps, float bc1, float bc2_sqrt,
                                             float grad_scale, float max_norm, const float* __restrict__ sq) {
  float coef = grad_scale;
  if (sq) {
    const float norm = sqrtf(*sq);
    if (!isfinite(norm)) return;
    if (max_norm > 0.f) coef *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const long long nchunks = (n + 1023) / 1024;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int gid = chunk_group[ch];
    if (gid >= 64) continue;
    const float lr = groups.lr[gid], wd = groups.wd[gid];
    const long long i = ch * 1024 + threadIdx.x * 4;
    if (i >= n) continue;
    float4 pv = ld4(p + i), mv = ld4(m + i), vv = ld4(v + i);
    const float4 gv = ld4(g + i);
    float* pp = &pv.x; float* mp = &mv.x; float* vp = &vv.x; const float* gp = &gv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gi = gp[j] * coef;
      float pi = pp[j] * (1.f - lr * wd);
      const float mi = beta1 * mp[j] + (1.f - beta1) * gi;
      const float vi = beta2 * vp[j] + (1.f - beta2) * gi * gi;
      pi -= (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
      pp[j] = pi; mp[j] = mi; vp[j] = vi;
    }
    st4(p + i, pv); st4(m + i, mv); st4(v + i, vv);
    if (shadow) st4_bf16(shadow + i, pv);
  }
}

static inline int row_grid(int rows) { return std::max(1, std::min(ceil_div(rows, kRowWarps), num_sms() * 8)); }
static inline int flat_grid(long long n, int per_thread = 1) {
  return (int)std::max<long long>(1, std::min<long long>(ceil_div<long long>(n, 256LL * per_thread), (long long)num_sms() * 16));
}

}  // namespace vit
}  // namespace memb

using namespace memb;
using namespace memb::vit;

#define MEMB_CH_DISPATCH(D, CALL)                        \
  switch ((D) / 128) {                                   \
    case 1: { constexpr int CH = 1; CALL; } break;       \
    case 2: { constexpr int CH = 2; CALL; } break;       \
    case 3: { constexpr int CH = 3; CALL; } break;       \
    case 4: { constexpr int CH = 4; CALL; } break;       \
    case 6: { constexpr int CH = 6; CALL; } break;       \
    case 8: { constexpr int CH = 8; CALL; } break;       \
    case 10: { constexpr int CH = 10; CALL; } break;     \
    case 16: { constexpr int CH = 16; CALL; } break;     \
    default: return ::memb::fail(MEMB_EINVAL, "unsupported embedding dim %d", (D)); \
  }
#define MEMB_REQ_D(D) MEMB_REQUIRE((D) % 128 == 0 && (D) <= 128 * kMaxChunks, "embedding dim must be a multiple of 128 and <= 2048, got %d", (D))

extern "C" int memb_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int rows,
                                  int D, void* y, int64_t ldy, float* mean, float* rstd, const int32_t* row_index,
                                  const int32_t* count, memb_stream_t s) {
  MEMB_REQ_D(D);
  MEMB_REQUIRE(x && gamma && beta && y && mean && rstd && rows > 0, "layernorm_fwd: bad arguments");
  MEMB_CH_DISPATCH(D, (layernorm_fwd<CH><<<row_grid(rows), kRowThreads, 0, s>>>(x, ldx, gamma, beta, eps, rows, D, (bf16*)y, ldy, mean, rstd, row_index, count)));
  MEMB_LAUNCH_OK("layernorm_fwd");
  return MEMB_OK;
}

static int layernorm_bwd_launch(const void* dy, int dy_dtype, int64_t lddy, const float* x, int64_t ldx,
                                const float* gamma, const float* mean, const float* rstd, int rows, int D, float* dx,
                                int64_t lddx, float* dgamma, float* dbeta, const int32_t* row_index,
                                const int32_t* count, const BranchArgs* br, memb_stream_t s) {
  MEMB_REQ_D(D);
  MEMB_REQUIRE(dy && x && gamma && mean && rstd && dx && rows > 0, "layernorm_bwd: bad arguments");
  const size_t smem = (size_t)(br ? 4 : 2) * kRowWarps * D * sizeof(float);
  MEMB_REQUIRE(smem <= 113 * 1024, "layernorm_bwd: D = %d needs %zu B of shared memory per block", D, smem);
  const int grid = std::max(1, std::min(ceil_div(rows, kRowWarps), num_sms() * 2));   // persistent: 2 blocks per SM
  const BranchArgs none{};
#define MEMB_LN_BWD(T, BR)                                                                                              \
  MEMB_CH_DISPATCH(D, (cudaFuncSetAttribute(layernorm_bwd<T, CH, BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                       layernorm_bwd<T, CH, BR><<<grid, kRowThreads, smem, s>>>((const T*)dy, lddy, x, ldx, gamma, mean, rstd, rows, D, dx, \
                                                                                lddx, dgamma, dbeta, row_index, count, br ? *br : none)))
  if (dy_dtype == MEMB_DT_BF16) {
    if (br) { MEMB_LN_BWD(bf16, true); } else { MEMB_LN_BWD(bf16, false); }
  } else {
    if (br) { MEMB_LN_BWD(float, true); } else { MEMB_LN_BWD(float, false); }
  }
#undef MEMB_LN_BWD
  MEMB_LAUNCH_OK("layernorm_bwd");
  return MEMB_OK;
}

extern "C" int memb_layernorm_bwd(const void* dy, int dy_dtype, int64_t lddy, const float* x, int64_t ldx,
                                  const float* gamma, const float* mean, const float* rstd, int rows, int D, float* dx,
                                  int64_t lddx, float* dgamma, float* dbeta, const int32_t* row_index,
                                  const int32_t* count, memb_stream_t s) {
  return layernorm_bwd_launch(dy, dy_dtype, lddy, x, ldx, gamma, mean, rstd, rows, D, dx, lddx, dgamma, dbeta, row_index, count,
                              nullptr, s);
}

extern "C" int memb_layernorm_bwd_branch(const void* dy, int dy_dtype, int64_t lddy, const float* x, int64_t ldx,
                                         const float* gamma, const float* mean, const float* rstd, int rows, int D, float* dx,
                                         int64_t lddx, float* dgamma, float* dbeta, const void* branch, int64_t ldb,
                                         const float* colscale, const float* rowscale, int rows_per_group, void* dz,
                                         int64_t lddz, float* dcolscale, float* dbias, memb_stream_t s) {
  MEMB_REQUIRE(dz && (!dcolscale || branch), "layernorm_bwd_branch: bad arguments");
  MEMB_REQUIRE(!rowscale || rows_per_group > 0, "layernorm_bwd_branch: rows_per_group must be positive");
  const BranchArgs br{(const bf16*)branch, ldb, colscale, rowscale, rows_per_group, (bf16*)dz, lddz, dcolscale, dbias};
  return layernorm_bwd_launch(dy, dy_dtype, lddy, x, ldx, gamma, mean, rstd, rows, D, dx, lddx, dgamma, dbeta, nullptr, nullptr,
                              &br, s);
}

extern "C" int memb_branch_bwd(const float* gout, int64_t ldg, const void* branch, int64_t ldb, const float* colscale,
                               const float* rowscale, int rows_per_group, int rows, int D, void* dz, int64_t lddz,
                               float* dcolscale, float* dbias, memb_stream_t s) {
  MEMB_REQ_D(D);
  MEMB_REQUIRE(gout && dz && rows > 0 && (!dcolscale || branch), "branch_bwd: bad arguments");
  MEMB_REQUIRE(!rowscale || rows_per_group > 0, "branch_bwd: rows_per_group must be positive");
  MEMB_CH_DISPATCH(D, (branch_bwd<CH><<<row_grid(rows), kRowThreads, (size_t)kRowWarps * D * sizeof(float), s>>>(
      gout, ldg, (const bf16*)branch, ldb, colscale, rowscale, rows_per_group, rows, D, (bf16*)dz, lddz, dcolscale, dbias)));
  MEMB_LAUNCH_OK("branch_bwd");
  return MEMB_OK;
}

extern "C" int memb_colsum_bf16(const void* x, int64_t ld, int rows, int N, float* out, memb_stream_t s) {
  MEMB_REQUIRE(x && out && rows > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0 && (((uintptr_t)x) & 15u) == 0,
               "colsum: N and ld must be multiples of 8 and x 16-byte aligned");
  const int gx = ceil_div(N, 256);
  // about 8 resident blocks per SM over the whole grid, at least 64 rows (8 loads per thread) per block
  const int gy = std::max(1, std::min(ceil_div(rows, 64), ceil_div(num_sms() * 8, gx)));
  const int rpb = ceil_div(rows, gy);
  colsum_bf16<<<dim3(gx, ceil_div(rows, rpb)), dim3(32, kColsumRows), 0, s>>>((const bf16*)x, ld, rows, N, rpb, out);
  MEMB_LAUNCH_OK("colsum_bf16");
  return MEMB_OK;
}

extern "C" int memb_vbias_chain(const float* t, const float* w_proj, int D, float* dproj_bias, float* dv_bias, memb_stream_t s) {
  MEMB_REQUIRE(t && w_proj && dproj_bias && dv_bias && D > 0, "vbias_chain: bad arguments");
  vbias_chain<<<dim3(ceil_div(D, 256), kChainSlices), 256, 0, s>>>(t, w_proj, D, dproj_bias, dv_bias);
  MEMB_LAUNCH_OK("vbias_chain");
  return MEMB_OK;
}

extern "C" int memb_patchify(const float* img, int B, int C, int H, int W, int P, void* out, memb_stream_t s) {
  MEMB_REQUIRE(img && out && B > 0 && C > 0 && P % 8 == 0 && H % P == 0 && W % P == 0, "patchify: bad arguments");
  patchify<<<flat_grid((long long)B * C * H * W / 8), 256, 0, s>>>(img, B, C, H, W, P, (bf16*)out);
  MEMB_LAUNCH_OK("patchify");
  return MEMB_OK;
}

extern "C" int memb_cls_pos(float* x, const float* cls, const float* pos, int B, int N, int D, memb_stream_t s) {
  MEMB_REQUIRE(x && cls && B > 0, "cls_pos: bad arguments");
  cls_pos<<<flat_grid(pos ? (long long)B * N * D : (long long)B * D), 256, 0, s>>>(x, cls, pos, B, N, D);
  MEMB_LAUNCH_OK("cls_pos");
  return MEMB_OK;
}

extern "C" int memb_embed_bwd(const float* g0, const uint8_t* mask, int B, int P, int D, void* dpatch, float* dmask_token,
                              float* dcls, float* dbias, float* dpos, memb_stream_t s) {
  MEMB_REQ_D(D);
  MEMB_REQUIRE(g0 && mask && dpatch, "embed_bwd: bad arguments");
  MEMB_CH_DISPATCH(D, (embed_bwd<CH><<<row_grid(B * (P + 1)), kRowThreads, (size_t)kRowWarps * D * sizeof(float), s>>>(g0, mask, B, P, D, (bf16*)dpatch, dmask_token, dcls, dbias, dpos)));
  MEMB_LAUNCH_OK("embed_bwd");
  return MEMB_OK;
}

extern "C" int memb_mask_compact(const uint8_t* mask, int B, int P, int32_t* row_index, int32_t* patch_index,
                                 int32_t* count, int cap, memb_stream_t s) {
  MEMB_REQUIRE(mask && row_index && patch_index && count && B > 0 && P > 0 && cap > 0, "mask_compact: bad arguments");
  mask_compact<<<1, 1024, 0, s>>>(mask, B * P, P, row_index, patch_index, count, cap);
  MEMB_LAUNCH_OK("mask_compact");
  return MEMB_OK;
}

extern "C" int memb_cross_entropy(const float* logits, int64_t ld, const int64_t* tokens, const int32_t* patch_index,
                                  const int32_t* count, int cap, int V, void* dlogits, int64_t ldd, float* stats,
                                  float grad_scale, memb_stream_t s) {
  MEMB_REQUIRE(logits && tokens && patch_index && count && stats && cap > 0, "cross_entropy: bad arguments");
  MEMB_REQUIRE(V % 4 == 0 && V <= 256 * 4 * 16, "cross_entropy: V must be a multiple of 4 and <= 16384, got %d", V);
  const long long* tok = reinterpret_cast<const long long*>(tokens);
  if (V <= 256 * 4 * 2) cross_entropy<2><<<cap, 256, 0, s>>>(logits, ld, tok, patch_index, count, V, (bf16*)dlogits, ldd, stats, grad_scale);
  else if (V <= 256 * 4 * 8) cross_entropy<8><<<cap, 256, 0, s>>>(logits, ld, tok, patch_index, count, V, (bf16*)dlogits, ldd, stats, grad_scale);
  else cross_entropy<16><<<cap, 256, 0, s>>>(logits, ld, tok, patch_index, count, V, (bf16*)dlogits, ldd, stats, grad_scale);
  MEMB_LAUNCH_OK("cross_entropy");
  return MEMB_OK;
}

extern "C" int memb_relpos_gather(const float* table, const int64_t* index, int N, int heads, int ldk, float* bias,
                                  float* biasT, memb_stream_t s) {
  MEMB_REQUIRE(table && index && bias && N > 0 && heads > 0 && ldk >= N, "relpos_gather: bad arguments");
  relpos_gather<<<flat_grid((long long)N * ldk * heads), 256, 0, s>>>(table, (const long long*)index, N, heads, ldk, bias, biasT);
  MEMB_LAUNCH_OK("relpos_gather");
  return MEMB_OK;
}
extern "C" int memb_attention_pack_bias(const float* dense, int ld, int N, int heads, float* packed, memb_stream_t s) {
  MEMB_REQUIRE(dense && packed && N > 0 && N <= 4 * MEMB_ATTN_BIAS_GROUPS && heads > 0 && ld >= N, "attention_pack_bias: bad arguments");
  MEMB_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15u) == 0, "attention_pack_bias: packed must be 16-byte aligned");
  attn_bias_pack<<<flat_grid((long long)heads * MEMB_ATTN_BIAS_GROUPS * 256), 256, 0, s>>>(dense, ld, N, heads, (float4*)packed);
  MEMB_LAUNCH_OK("attn_bias_pack");
  return MEMB_OK;
}
extern "C" int memb_relpos_scatter(const float* dbias, int ldk, const int64_t* index, int N, int heads, float* dtable,
                                   memb_stream_t s) {
  MEMB_REQUIRE(dbias && index && dtable && N > 0 && heads > 0 && ldk >= N, "relpos_scatter: bad arguments");
  relpos_scatter<<<flat_grid((long long)N * N * heads), 256, 0, s>>>(dbias, ldk, (const long long*)index, N, heads, dtable);
  MEMB_LAUNCH_OK("relpos_scatter");
  return MEMB_OK;
}
extern "C" int memb_batch_reduce_bf16(const void* x, int B, int64_t inner, float* acc, memb_stream_t s) {
  MEMB_REQUIRE(x && acc && B > 0 && inner > 0 && inner % 4 == 0, "batch_reduce: bad arguments");
  batch_reduce_bf16<<<(unsigned)ceil_div<long long>(inner / 4, 256), 256, 0, s>>>((const bf16*)x, B, inner, acc);
  MEMB_LAUNCH_OK("batch_reduce_bf16");
  return MEMB_OK;
}

extern "C" int memb_meanpool_fwd(const float* x, int B, int N, int D, float* out, memb_stream_t s) {
  MEMB_REQUIRE(x && out && B > 0 && N > 1 && D > 0 && D % 4 == 0, "meanpool_fwd: bad arguments");
  meanpool_fwd<<<dim3((unsigned)ceil_div(D / 4, 256), (unsigned)B), 256, 0, s>>>(x, B, N, D, out);
  MEMB_LAUNCH_OK("meanpool_fwd");
  return MEMB_OK;
}
extern "C" int memb_meanpool_bwd(const float* dpool, int B, int N, int D, float* gres, memb_stream_t s) {
  MEMB_REQUIRE(dpool && gres && B > 0 && N > 1 && D > 0 && D % 4 == 0, "meanpool_bwd: bad arguments");
  meanpool_bwd<<<flat_grid((long long)B * N * D / 4), 256, 0, s>>>(dpool, B, N, D, gres);
  MEMB_LAUNCH_OK("meanpool_bwd");
  return MEMB_OK;
}
extern "C" int memb_linear_small_fwd(const void* z, const float* W, const float* bias, int B, int D, int C, float* out,
                                     memb_stream_t s) {
  MEMB_REQUIRE(z && W && out && B > 0 && C > 0 && D > 0 && D % 4 == 0, "linear_small_fwd: bad arguments");
  linear_small_fwd<<<(unsigned)ceil_div<long long>((long long)B * C * 32, 256), 256, 0, s>>>((const bf16*)z, W, bias, B, D, C, out);
  MEMB_LAUNCH_OK("linear_small_fwd");
  return MEMB_OK;
}
extern "C" int memb_linear_small_bwd(const float* dl, const void* z, const float* W, int B, int D, int C, float* dW, float* db,
                                     float* dz, memb_stream_t s) {
  MEMB_REQUIRE(dl && z && W && dW && dz && B > 0 && C > 0 && D > 0, "linear_small_bwd: bad arguments");
  const long long n = (long long)C * D + (long long)B * D;
  linear_small_bwd<<<(unsigned)ceil_div<long long>(n, 256), 256, 0, s>>>(dl, (const bf16*)z, W, B, D, C, dW, db, dz);
  MEMB_LAUNCH_OK("linear_small_bwd");
  return MEMB_OK;
}

extern "C" int memb_fill_f32(float* p, int64_t n, float v, memb_stream_t s) {
  MEMB_REQUIRE(p && n > 0, "fill: bad arguments");
  fill_f32<<<flat_grid(n, 4), 256, 0, s>>>(p, n, v);
  MEMB_LAUNCH_OK("fill_f32");
  return MEMB_OK;
}
extern "C" int memb_cast_bf16(const float* src, void* dst, int64_t n, memb_stream_t s) {
  MEMB_REQUIRE(src && dst && n > 0, "cast: bad arguments");
  cast_bf16<<<flat_grid(n, 4), 256, 0, s>>>(src, (bf16*)dst, n);
  MEMB_LAUNCH_OK("cast_bf16");
  return MEMB_OK;
}
extern "C" int memb_sqnorm(const float* g, int64_t n, float scale, float* out, memb_stream_t s) {
  MEMB_REQUIRE(g && out && n > 0, "sqnorm: bad arguments");
  sqnorm<<<flat_grid(n, 8), 256, 0, s>>>(g, n, scale, out);
  MEMB_LAUNCH_OK("sqnorm");
  return MEMB_OK;
}
extern "C" int memb_soft_ce(const float* logits, int B, int C, const int64_t* labels, const float* soft_targets, float smoothing,
                            float* loss_out, float* dlogits, memb_stream_t s) {
  MEMB_REQUIRE(logits && loss_out && B > 0 && C > 0, "soft_ce: bad arguments");
  MEMB_REQUIRE((labels != nullptr) != (soft_targets != nullptr), "soft_ce: give integer labels or soft targets, not both");
  MEMB_REQUIRE(smoothing >= 0.f && smoothing < 1.f, "soft_ce: smoothing must be in [0, 1)");
  soft_ce<<<ceil_div(B, 4), 128, 0, s>>>(logits, B, C, reinterpret_cast<const long long*>(labels), soft_targets, smoothing, loss_out, dlogits);
  MEMB_LAUNCH_OK("soft_ce");
  return MEMB_OK;
}
extern "C" int memb_sqnorm_groups(const float* g, int64_t n, float scale, const uint8_t* chunk_group, float* out, memb_stream_t s) {
  MEMB_REQUIRE(g && out && chunk_group && n > 0 && n % 4 == 0, "sqnorm_groups: bad arguments");
  const long long nchunks = (n + 1023) / 1024;
  sqnorm_groups<<<(int)std::min<long long>(nchunks, (long long)num_sms() * 16), 256, 0, s>>>(g, n, scale, chunk_group, out);
  MEMB_LAUNCH_OK("sqnorm_groups");
  return MEMB_OK;
}
extern "C" int memb_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n,
                          const uint8_t* chunk_group, const float* group_lr, const float* group_wd, int ngroups,
                          float beta1, float beta2, float eps, int step, float grad_scale, float max_norm,
                          const float* sqnorm_dev, memb_stream_t s) {
  MEMB_REQUIRE(p && g && m && v && chunk_group && group_lr && group_wd && n > 0 && step >= 1, "adamw: bad arguments");
  MEMB_REQUIRE(ngroups >= 1 && ngroups <= 64 && n % 4 == 0, "adamw: 1..64 parameter groups, n %% 4 == 0");
  AdamGroups groups{};
  for (int i = 0; i < ngroups; ++i) { groups.lr[i] = group_lr[i]; groups.wd[i] = group_wd[i]; }
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const long long nchunks = (n + 1023) / 1024;
  const int grid = (int)std::min<long long>(nchunks, (long long)num_sms() * 16);
  adamw<<<grid, 256, 0, s>>>(p, g, m, v, (bf16*)shadow_bf16, n, chunk_group, groups, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale, max_norm, sqnorm_dev);
  MEMB_LAUNCH_OK("adamw");
  return MEMB_OK;
}
