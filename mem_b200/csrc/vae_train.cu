// dVAE training step and decoder: the HBM-bound kernels around the bf16 tcgen05 GEMMs.
//
// Reference: eventvae/vae/vae_model.py -- DiscreteVAE.forward (:173-213: norm, encoder, gumbel_softmax over the token axis,
// einsum with the codebook, decoder, MSE / smooth-L1 reconstruction loss + KL to the uniform prior), decode (:160-171),
// ResBlock (:29-41), Conv2d(4, stride 2, padding 1) + ReLU (:91), ConvTranspose2d(4, stride 2, padding 1) + ReLU (:92);
// the training loop eventvae/train_vae.py:304-392 calls ``loss, recons = vae(images, return_loss=True, return_recons=True,
// temp=temp); loss.backward()``.
//
// Training is not index-exact work (unlike the tokenizer path in conv_f16.cu): activations are NHWC bf16 matrices
// [B*H*W, C], every convolution is an explicit im2col / col2im around gemm_tcgen05 (bf16 operands, fp32 accumulation):
//   Conv2d           fwd  y  = im2col(x) W^T                 bwd  dW += dy^T im2col(x),  dx = col2im(dy W)
//   ConvTranspose2d  fwd  y  = col2im(x Wt)                  bwd  dWt += x^T im2col(dy), dx = im2col(dy) Wt^T
// so the kernels here are: layout changes (NCHW fp32 <-> NHWC bf16, with DiscreteVAE.norm folded in), im2col, col2im
// (gather form: no atomics), ReLU / residual elementwise passes, the gumbel-softmax row kernel pair with the KL term fused
// in, the reconstruction loss, and the codebook row gather of decode().
#include <algorithm>

#include <cuda_bf16.h>

#include "common.cuh"

namespace memb {
namespace vaet {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block reductions for 256-thread blocks
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]);
  return t;
}

static inline int grid_for(long long n, int per_block = 256) {
  return (int)std::max<long long>(1, std::min<long long>(ceil_div<long long>(n, per_block), (long long)num_sms() * 32));
}

// ---------------------------------------------------------------- layout changes
// img fp32 [B,C,H,W] -> NHWC bf16 [B,H,W,Cpad] (channels >= C are zero), optional (x - mean[c]) / std[c] (DiscreteVAE.norm);
// also keeps the normalised image as fp32 NHWC [B,H,W,C] for the reconstruction loss.
__global__ void __launch_bounds__(256) nchw_to_nhwc(const float* __restrict__ img, int B, int C, int H, int W, int Cpad,
                                                    const float* __restrict__ mean, const float* __restrict__ stdv,
                                                    bf16* __restrict__ out, float* __restrict__ out_f32) {
  const long long total = (long long)B * H * W * Cpad;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % Cpad);
    const long long pix = i / Cpad;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const long long b = pix / ((long long)W * H);
    float v = 0.f;
    if (c < C) {
      v = img[((b * C + c) * H + y) * W + x];
      if (mean) v = (v - mean[c]) / stdv[c];
      if (out_f32) out_f32[pix * C + c] = v;
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// x fp32 NHWC [B,H,W,ld] (first C channels) -> fp32 NCHW [B,C,H,W]
__global__ void __launch_bounds__(256) nhwc_to_nchw(const float* __restrict__ x, int B, int C, int H, int W, int ld,
                                                    float* __restrict__ out) {
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int xx = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % C);
    const long long b = i / ((long long)W * H * C);
    out[i] = x[((b * H + y) * W + xx) * ld + c];
  }
}

// ---------------------------------------------------------------- im2col / col2im (NHWC, K ordered (ky, kx, c))
// col[(b, oy, ox), (ky, kx, c)] = x[b, oy*s + ky - p, ox*s + kx - p, c] (zero outside); 8 channels (16 bytes) per thread.
__global__ void __launch_bounds__(256) im2col_nhwc(const bf16* __restrict__ x, int B, int H, int W, int C, int kh, int kw, int s,
                                                   int p, int OH, int OW, bf16* __restrict__ col) {
  const int c8 = C / 8;
  const long long total = (long long)B * OH * OW * kh * kw * c8;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cv = (int)(i % c8);
    long long r = i / c8;
    const int kx = (int)(r % kw); r /= kw;
    const int ky = (int)(r % kh); r /= kh;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const long long b = r / OH;
    const int iy = oy * s + ky - p, ix = ox * s + kx - p;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = *reinterpret_cast<const uint4*>(x + ((b * H + iy) * W + ix) * C + cv * 8);
    *reinterpret_cast<uint4*>(col + i * 8) = v;
  }
}

// Inverse of im2col as a GATHER: out[b, y, x, c] = sum over taps (ky, kx) with (y + p - ky) % s == 0 and (x + p - kx) % s == 0
// of col[(b, (y + p - ky)/s, (x + p - kx)/s), (ky, kx, c)]   (+ bias[c]) (ReLU) (* (gate[b,y,x,c] > 0): ReLU backward of the
// layer below).  col is fp32 [B*OH*OW, ldc]; out bf16 and / or fp32 NHWC [B,H,W,C].  Used for Conv2d dgrad and for the
// ConvTranspose2d forward.
__global__ void __launch_bounds__(256) col2im_nhwc(const float* __restrict__ col, long long ldc, int B, int H, int W, int C, int kh,
                                                   int kw, int s, int p, int OH, int OW, const float* __restrict__ bias,
                                                   int relu, const bf16* __restrict__ gate, bf16* __restrict__ out,
                                                   float* __restrict__ out_f32, int ld_out) {
  const long long total = (long long)B * H * W * C;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const long long b = r / H;
    float acc = bias ? bias[c] : 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int ty = y + p - ky;
      if (ty < 0 || ty % s) continue;
      const int oy = ty / s;
      if (oy >= OH) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int tx = x + p - kx;
        if (tx < 0 || tx % s) continue;
        const int ox = tx / s;
        if (ox >= OW) continue;
        acc += col[((b * OH + oy) * OW + ox) * ldc + (ky * kw + kx) * C + c];
      }
    }
    if (relu) acc = fmaxf(acc, 0.f);
    const long long o = ((b * H + y) * W + x) * ld_out + c;
    if (gate && !(__bfloat162float(gate[o]) > 0.f)) acc = 0.f;
    if (out) out[o] = __float2bfloat16_rn(acc);
    if (out_f32) out_f32[o] = acc;
  }
}

// ---------------------------------------------------------------- elementwise passes on bf16 matrices
// op 0: out = relu(a)            op 1: out = a * (b > 0)  (ReLU backward, b = forward output)
// op 2: out = a + b              op 3: out = relu(a) and the input is fp32 (GEMM output with bias) -- unused placeholder
__global__ void __launch_bounds__(256) ew_bf16(int op, const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out,
                                               long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float av = __bfloat162float(a[i]);
    float r;
    if (op == 0) r = fmaxf(av, 0.f);
    else if (op == 1) r = __bfloat162float(b[i]) > 0.f ? av : 0.f;
    else r = av + __bfloat162float(b[i]);
    out[i] = __float2bfloat16_rn(r);
  }
}

// rows of a [V, D] fp32 table -> bf16 [n, ld] (decode(): codebook lookup; columns >= D are zero)
__global__ void __launch_bounds__(256) gather_rows(const float* __restrict__ table, int D, const long long* __restrict__ idx,
                                                   long long n, int V, int ld, bf16* __restrict__ out, int* __restrict__ err) {
  const long long total = n * ld;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % ld);
    const long long r = idx[i / ld];
    float v = 0.f;
    if (r < 0 || r >= V) { if (err) *err = 1; }
    else if (c < D) v = table[r * D + c];
    out[i] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------- gumbel softmax (+ KL to the uniform prior)
// One block per token position (row of logits [M, N] fp32).  y = softmax((logits + g) / tau) (F.gumbel_softmax, soft
// sample); hard: the forward value is onehot(argmax y) (straight-through, the gradient is the soft one).  Also
//   kl_out[0] += sum_n q_n (log q_n + log N),  q = softmax(logits)        (vae_model.py:204-208: F.kl_div(log_uniform, log_qy,
//   'batchmean', log_target=True) with a one-element input: the "batch" it divides by is 1).
// Saves the two row log-sum-exps for the backward pass.
__global__ void __launch_bounds__(256) gumbel_fwd(const float* __restrict__ logits, const float* __restrict__ noise, int N,
                                                  float inv_tau, int hard, bf16* __restrict__ y, bf16* __restrict__ y_fwd,
                                                  float* __restrict__ lse, float* __restrict__ kl_out) {
  __shared__ float red[8];
  __shared__ int redi[8];
  const long long row = blockIdx.x;
  const float* l = logits + row * N;
  const float* g = noise + row * N;
  float m1 = -INFINITY, m2 = -INFINITY;
  for (int c = threadIdx.x; c < N; c += 256) {
    const float a = l[c];
    m1 = fmaxf(m1, a);
    m2 = fmaxf(m2, (a + g[c]) * inv_tau);
  }
  m1 = block_max(m1, red);
  m2 = block_max(m2, red);
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < N; c += 256) {
    const float a = l[c];
    s1 += expf(a - m1);
    s2 += expf((a + g[c]) * inv_tau - m2);
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  const float lse1 = m1 + logf(s1), lse2 = m2 + logf(s2), logN = logf((float)N);
  float kl = 0.f, best = -INFINITY;
  int arg = 0;
  for (int c = threadIdx.x; c < N; c += 256) {
    const float a = l[c];
    const float lq = a - lse1;
    kl += expf(lq) * (lq + logN);
    const float z = (a + g[c]) * inv_tau;
    const float yv = expf(z - lse2);
    y[row * N + c] = __float2bfloat16_rn(yv);
    if (z > best) { best = z; arg = c; }
  }
  kl = block_sum(kl, red);
  if (threadIdx.x == 0) {
    lse[2 * row] = lse1;
    lse[2 * row + 1] = lse2;
    atomicAdd(kl_out, kl);
  }
  if (hard) {     // block arg-max, first index wins ties (torch.max)
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = best; redi[threadIdx.x >> 5] = arg; }
    __syncthreads();
    best = red[0]; arg = redi[0];
    for (int w = 1; w < 8; ++w)
      if (red[w] > best || (red[w] == best && redi[w] < arg)) { best = red[w]; arg = redi[w]; }
    for (int c = threadIdx.x; c < N; c += 256) y_fwd[row * N + c] = __float2bfloat16_rn(c == arg ? 1.f : 0.f);
  }
}

// dlogits = (1/tau) * y * (dy - sum_n y_n dy_n)  +  kl_scale * q * ((log q + log N) - sum_n q_n (log q_n + log N))
// with y = softmax((logits + g)/tau), q = softmax(logits) recomputed from the saved log-sum-exps; dy fp32 [M, N].
__global__ void __launch_bounds__(256) gumbel_bwd(const float* __restrict__ logits, const float* __restrict__ noise,
                                                  const float* __restrict__ lse, const float* __restrict__ dy, int N, float inv_tau,
                                                  const float* __restrict__ gscale, float kl_weight, bf16* __restrict__ dlogits) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const float* l = logits + row * N;
  const float* g = noise + row * N;
  const float* d = dy + row * N;
  const float lse1 = lse[2 * row], lse2 = lse[2 * row + 1], logN = logf((float)N);
  float ydy = 0.f, qk = 0.f;
  for (int c = threadIdx.x; c < N; c += 256) {
    const float a = l[c];
    ydy += expf((a + g[c]) * inv_tau - lse2) * d[c];
    const float lq = a - lse1;
    qk += expf(lq) * (lq + logN);
  }
  ydy = block_sum(ydy, red);
  qk = block_sum(qk, red);
  const float gs = gscale ? *gscale : 1.f;
  for (int c = threadIdx.x; c < N; c += 256) {
    const float a = l[c];
    const float yv = expf((a + g[c]) * inv_tau - lse2);
    const float lq = a - lse1;
    const float v = inv_tau * yv * (d[c] - ydy) + gs * kl_weight * expf(lq) * (lq + logN - qk);
    dlogits[row * N + c] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------- reconstruction loss
// kind 0: mse_loss(img, out) = mean (img - out)^2 ; kind 1: smooth_l1_loss (beta 1).  target / recon fp32 NHWC [n] with row
// strides (C valid channels of ld_r).  loss_out[0] += loss; dout = d loss / d recon (bf16, padded channels zero).
__global__ void __launch_bounds__(256) recon_loss(const float* __restrict__ target, const float* __restrict__ recon, long long pixels,
                                                  int C, int ld_r, int kind, float* __restrict__ loss_out, bf16* __restrict__ dout) {
  __shared__ float red[8];
  const long long total = pixels * ld_r;
  const float inv_n = 1.f / (float)(pixels * C);
  float acc = 0.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % ld_r);
    float gval = 0.f;
    if (c < C) {
      const float d = recon[i] - target[(i / ld_r) * C + c];
      if (kind == 0) {
        acc += d * d;
        gval = 2.f * d * inv_n;
      } else {
        const float ad = fabsf(d);
        acc += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
        gval = (ad < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * inv_n;
      }
    }
    if (dout) dout[i] = __float2bfloat16_rn(gval);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss_out, acc * inv_n);
}

// dst (+)= scale[0] * src   (fp32; parameter-gradient hand-off with the upstream gradient of the loss)
__global__ void __launch_bounds__(256) axpy_f32(const float* __restrict__ src, const float* __restrict__ scale, float* __restrict__ dst,
                                                long long n, int accumulate) {
  const float s = scale ? *scale : 1.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    dst[i] = (accumulate ? dst[i] : 0.f) + s * src[i];
}

}  // namespace vaet
}  // namespace memb

using namespace memb;
using namespace memb::vaet;

extern "C" int memb_vae_nchw_to_nhwc(const float* img, int B, int C, int H, int W, int Cpad, const float* mean, const float* stdv,
                                     void* out_bf16, float* out_f32, memb_stream_t s) {
  MEMB_REQUIRE(img && out_bf16 && B > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "vae_nchw_to_nhwc: bad arguments");
  MEMB_REQUIRE((mean == nullptr) == (stdv == nullptr), "vae_nchw_to_nhwc: mean and std go together");
  nchw_to_nhwc<<<grid_for((long long)B * H * W * Cpad), 256, 0, s>>>(img, B, C, H, W, Cpad, mean, stdv, (bf16*)out_bf16, out_f32);
  MEMB_LAUNCH_OK("vae_nchw_to_nhwc");
  return MEMB_OK;
}

extern "C" int memb_vae_nhwc_to_nchw(const float* x, int B, int C, int H, int W, int ld, float* out, memb_stream_t s) {
  MEMB_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0 && ld >= C, "vae_nhwc_to_nchw: bad arguments");
  nhwc_to_nchw<<<grid_for((long long)B * C * H * W), 256, 0, s>>>(x, B, C, H, W, ld, out);
  MEMB_LAUNCH_OK("vae_nhwc_to_nchw");
  return MEMB_OK;
}

extern "C" int memb_vae_im2col(const void* x_bf16, int B, int H, int W, int C, int kh, int kw, int stride, int pad, void* col_bf16,
                               memb_stream_t s) {
  MEMB_REQUIRE(x_bf16 && col_bf16 && B > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "vae_im2col: bad arguments");
  MEMB_REQUIRE(C > 0 && C % 8 == 0, "vae_im2col: channels must be a multiple of 8 (pad the tensor), got %d", C);
  const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
  MEMB_REQUIRE(OH > 0 && OW > 0, "vae_im2col: empty output");
  im2col_nhwc<<<grid_for((long long)B * OH * OW * kh * kw * (C / 8)), 256, 0, s>>>((const bf16*)x_bf16, B, H, W, C, kh, kw, stride, pad, OH,
                                                                                 OW, (bf16*)col_bf16);
  MEMB_LAUNCH_OK("vae_im2col");
  return MEMB_OK;
}

extern "C" int memb_vae_col2im(const float* col, int64_t ldc, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                               const float* bias, int relu, const void* gate_bf16, void* out_bf16, float* out_f32, int ld_out,
                               memb_stream_t s) {
  MEMB_REQUIRE(col && (out_bf16 || out_f32) && B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0,
               "vae_col2im: bad arguments");
  const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
  MEMB_REQUIRE(OH > 0 && OW > 0 && ldc >= (int64_t)kh * kw * C && ld_out >= C, "vae_col2im: bad geometry");
  col2im_nhwc<<<grid_for((long long)B * H * W * C), 256, 0, s>>>(col, ldc, B, H, W, C, kh, kw, stride, pad, OH, OW, bias, relu,
                                                                (const bf16*)gate_bf16, (bf16*)out_bf16, out_f32, ld_out);
  MEMB_LAUNCH_OK("vae_col2im");
  return MEMB_OK;
}

extern "C" int memb_vae_ew_bf16(int op, const void* a, const void* b, void* out, int64_t n, memb_stream_t s) {
  MEMB_REQUIRE(a && out && n > 0 && op >= 0 && op <= 2 && (op == 0 || b), "vae_ew_bf16: bad arguments");
  ew_bf16<<<grid_for(n), 256, 0, s>>>(op, (const bf16*)a, (const bf16*)b, (bf16*)out, n);
  MEMB_LAUNCH_OK("vae_ew_bf16");
  return MEMB_OK;
}

extern "C" int memb_vae_gather_rows(const float* table, int V, int D, const int64_t* idx, int64_t n, int ld, void* out_bf16,
                                    int* err_flag, memb_stream_t s) {
  MEMB_REQUIRE(table && idx && out_bf16 && V > 0 && D > 0 && n > 0 && ld >= D, "vae_gather_rows: bad arguments");
  gather_rows<<<grid_for(n * ld), 256, 0, s>>>(table, D, reinterpret_cast<const long long*>(idx), n, V, ld, (bf16*)out_bf16, err_flag);
  MEMB_LAUNCH_OK("vae_gather_rows");
  return MEMB_OK;
}

extern "C" int memb_vae_gumbel_fwd(const float* logits, const float* noise, int64_t rows, int N, float tau, int hard, void* y_bf16,
                                   void* y_fwd_bf16, float* lse, float* kl_out, memb_stream_t s) {
  MEMB_REQUIRE(logits && noise && y_bf16 && lse && kl_out && rows > 0 && N > 0 && tau > 0.f, "vae_gumbel_fwd: bad arguments");
  MEMB_REQUIRE(!hard || y_fwd_bf16, "vae_gumbel_fwd: the straight-through sample needs y_fwd");
  gumbel_fwd<<<(unsigned)rows, 256, 0, s>>>(logits, noise, N, 1.0f / tau, hard, (bf16*)y_bf16, (bf16*)y_fwd_bf16, lse, kl_out);
  MEMB_LAUNCH_OK("vae_gumbel_fwd");
  return MEMB_OK;
}

extern "C" int memb_vae_gumbel_bwd(const float* logits, const float* noise, const float* lse, const float* dy, int64_t rows, int N,
                                   float tau, const float* grad_scale_dev, float kl_weight, void* dlogits_bf16, memb_stream_t s) {
  MEMB_REQUIRE(logits && noise && lse && dy && dlogits_bf16 && rows > 0 && N > 0 && tau > 0.f, "vae_gumbel_bwd: bad arguments");
  gumbel_bwd<<<(unsigned)rows, 256, 0, s>>>(logits, noise, lse, dy, N, 1.0f / tau, grad_scale_dev, kl_weight, (bf16*)dlogits_bf16);
  MEMB_LAUNCH_OK("vae_gumbel_bwd");
  return MEMB_OK;
}

extern "C" int memb_vae_recon_loss(const float* target, const float* recon, int64_t pixels, int C, int ld_recon, int kind,
                                   float* loss_out, void* dout_bf16, memb_stream_t s) {
  MEMB_REQUIRE(target && recon && loss_out && pixels > 0 && C > 0 && ld_recon >= C && (kind == 0 || kind == 1),
               "vae_recon_loss: bad arguments");
  recon_loss<<<grid_for(pixels * ld_recon), 256, 0, s>>>(target, recon, pixels, C, ld_recon, kind, loss_out, (bf16*)dout_bf16);
  MEMB_LAUNCH_OK("vae_recon_loss");
  return MEMB_OK;
}

extern "C" int memb_axpy_f32(const float* src, const float* scale_dev, float* dst, int64_t n, int accumulate, memb_stream_t s) {
  MEMB_REQUIRE(src && dst && n > 0, "axpy_f32: bad arguments");
  axpy_f32<<<grid_for(n), 256, 0, s>>>(src, scale_dev, dst, n, accumulate);
  MEMB_LAUNCH_OK("axpy_f32");
  return MEMB_OK;
}
