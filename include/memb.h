/*
 * memb.h -- C ABI of libmemb.so, the sm_100a kernels behind the MEM (masked
 * event modelling) pretraining hot path.
 *
 * The reference (tum-vision/mem) is pure Python/PyTorch and has no FFI layer;
 * the drop-in boundary is its Python call surface (mem_b200/ mirrors it) and
 * this C ABI sits underneath.  Every entry point names the reference code it
 * replaces.  Conventions:
 *
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers
 *     unless the name ends in _host;
 *   - the caller owns every buffer including workspaces; the library never
 *     allocates, frees or retains device memory across calls (the only
 *     process-wide state is a per-device cache of TMA descriptors keyed by
 *     pointer/shape, which holds no memory);
 *   - every call takes the CUDA stream to launch on and returns without
 *     synchronising unless documented otherwise;
 *   - return value: MEMB_OK (0) or a negative memb_status; the text of the last
 *     failure on the calling thread is available from memb_last_error().
 */
#ifndef MEMB_H_
#define MEMB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* memb_stream_t; /* == cudaStream_t */

typedef enum memb_status {
  MEMB_OK = 0,
  MEMB_EINVAL = -1,     /* bad shape / alignment / null pointer            */
  MEMB_EOOB = -2,       /* an event indexed outside [-H*W, H*W): the       */
                        /* reference raises IndexError (datasets.py:581)   */
  MEMB_ECUDA = -3,      /* CUDA runtime / driver error                     */
  MEMB_EWORKSPACE = -4  /* workspace too small                             */
} memb_status;

const char* memb_last_error(void);
int memb_version(void);
/* Number of kernels this library has launched on the calling process so far
 * (bench.py reports the delta over its timed region as gpu_launches). */
int64_t memb_launch_count(void);

/* ------------------------------------------------------------------------
 * Event rasteriser.  Replaces EventArrToImg.__call__, mem/datasets.py:566-595
 * (x,y truncation :568-569, np.add.at scatter :581-582, time surface :587-589,
 * (H,W,3) layout :591-595).
 * ---------------------------------------------------------------------- */

/* strategy values for memb_hist_u8 */
#define MEMB_HIST_AUTO 0
#define MEMB_HIST_GLOBAL 1      /* L2-resident u32 accumulators + RED            */
#define MEMB_HIST_GLOBAL_AGG 2  /* same, duplicates merged per warp (match.any)  */
#define MEMB_HIST_TILE 3        /* shared-memory privatised sensor tiles         */
#define MEMB_HIST_PRIVATE 4     /* one stream, sensor <= 51200 px: a whole-sensor copy per SM in shared      */
                                /* memory, copies summed from per-CTA slices (else falls back to GLOBAL)    */
#define MEMB_HIST_GLOBAL_REPL 5 /* GLOBAL with 8 copies of the accumulator planes (one stream only): warps RED */
                                /* into different copies so that a concentrated stream (edges, hot pixels)    */
                                /* does not serialise on a few L2 sectors; the finalize pass adds the copies  */
#define MEMB_HIST_HYBRID 6      /* one stream, sensor too large for PRIVATE (up to 1 Mpixel): a sampled pass    */
                                /* picks the most frequent 64-pixel granules, every SM keeps a private copy of  */
                                /* those in shared memory and sends the remaining events to L2 REDs; AUTO picks */
                                /* it for >= 2^20 events (else falls back to GLOBAL)                            */
#define MEMB_HIST_SORT 7        /* one stream, sensor up to 4 Mpixel: no global atomics.  Pass 1 turns every    */
                                /* event into a 16-bit key inside one of 296 interleaved pixel classes and      */
                                /* counting-sorts each 8192-event chunk by class in shared memory (2 B written  */
                                /* per event); pass 2 runs one CTA per class over its segments with shared-     */
                                /* memory counters and writes the image directly.  Bit-exact, but measured      */
                                /* slower than HYBRID (10 M events at 640x480: 118-126 us vs 94), so AUTO does  */
                                /* not pick it                                                                  */

/* Bytes memb_hist_u8 needs for this problem (n = total rows; same strategy value as the call). */
size_t memb_hist_workspace_bytes(int B, int64_t n, int H, int W, int timesurface, int strategy);

/* ev      : float64 [n,4] rows [x,y,t,p], row-major.
 * offsets : NULL (one stream, B must be 1) or int64 [B+1] row offsets of a ragged batch.
 * out     : uint8 [B,H,W,C], C in {2,3}: C==3 -> [pos, time-surface|0, neg], C==2 -> [pos, neg].
 * max_stream_len: upper bound of the longest stream (sizing of the launch grid; n if unknown).
 * The out-of-range flag is left in the workspace; read it with memb_hist_status. */
int memb_hist_u8(const double* ev, int64_t n, const int64_t* offsets, int B, int64_t max_stream_len,
                 int H, int W, int C, int timesurface, int strategy, uint8_t* out, void* ws,
                 size_t ws_bytes, memb_stream_t stream);

/* Synchronises `stream`, then returns MEMB_EOOB if the last memb_hist_u8 on this
 * workspace saw an event outside the sensor (reference: IndexError), else MEMB_OK. */
int memb_hist_status(const void* ws, memb_stream_t stream);

/* max(trunc(x)), max(trunc(y)) over all rows (datasets.py:571-575, H/W = None).
 * Synchronises; result in host int64[2].  n must be > 0. */
int memb_hist_extent(const double* ev, int64_t n, int64_t* max_xy_host, void* ws, size_t ws_bytes,
                     memb_stream_t stream);

/* ------------------------------------------------------------------------
 * Event-space augmentations fused into the rasteriser, and the post-raster
 * tensor transforms.  Replaces, for one batch in one pass,
 * the per-sample numpy chain of build_transformNPY (mem/datasets.py:611-660):
 *   ReshapeScaleXandY (:464-485)  x *= scale_x, y *= scale_y
 *   SliceRandomMaxEvs (:488-498)  rows [start, start+count) of the stream
 *   RandomTimeFlip    (:598-608)  p = -p (row order and t only matter to the time surface)
 *   Aug_FlipEvsAlongX (:501-521)  x = flip_w - 1 - x
 *   Aug_RandomShiftEvs(:524-549)  x += shift_x, y += shift_y, rows outside [0,cull_w)x[0,cull_h) dropped
 *   EventArrToImg     (:552-595)  -> uint8 counts
 * and then ToTensor (/255), RandomCrop(pad_if_needed), RemoveTimesurface, RemoveHotPixels(num_stds),
 * NormalizeEvent (mem/transforms.py:225-275) -> float32 [B,C,outH,outW].
 * The random draws stay on the host (they consume Python's / numpy's / torch's global generators in the
 * reference's order, mem_b200/event_pipeline.py); the device applies them.  All float64 event arithmetic
 * is done with single correctly rounded operations in the reference's order, so the counts are bit-exact.
 * ---------------------------------------------------------------------- */
typedef struct memb_event_aug {   /* one per stream, 64 bytes, device memory */
  double scale_x, scale_y;        /* 1.0 = off */
  int64_t start, count;           /* window inside the stream; count < 0: to the end of the stream */
  int32_t time_flip;              /* != 0: polarity inverted */
  int32_t flip_x, flip_w;         /* != 0: x = (flip_w - 1) - x */
  int32_t cull;                   /* != 0: shift, then drop rows outside the cull window */
  int32_t shift_x, shift_y, cull_w, cull_h;
} memb_event_aug;

/* As memb_hist_u8 (same workspace size, strategies AUTO / GLOBAL / GLOBAL_AGG / TILE, no time surface),
 * with stream b transformed by aug[b] on the fly.  aug: device array [B]. */
int memb_hist_aug_u8(const double* ev, int64_t n, const int64_t* offsets, int B, int64_t max_stream_len,
                     const memb_event_aug* aug, int H, int W, int C, int strategy, uint8_t* out,
                     void* ws, size_t ws_bytes, memb_stream_t stream);

/* The same with the time surface in the middle channel (C == 3, strategy GLOBAL; workspace:
 * memb_hist_workspace_bytes(B, n, H, W, 1, MEMB_HIST_GLOBAL)): EventArrToImg(timeSurface=True) after the event-space
 * augmentations (mem/datasets.py:583-590) -- the timestamps are normalised over the rows that survive the window and the
 * shift stage's cull; after RandomTimeFlip (t' = t[last row of the window] - t, array reversed, :603-606) the row that
 * writes a pixel last is the one with the smallest index. */
int memb_hist_aug_tss_u8(const double* ev, int64_t n, const int64_t* offsets, int B, int64_t max_stream_len,
                         const memb_event_aug* aug, int H, int W, int timesurface, uint8_t* out, void* ws,
                         size_t ws_bytes, memb_stream_t stream);

/* hist    : uint8 [B,H,W,C] counts, C == 3 [pos, ts, neg] or C == 2 [pos, neg].
 * crop_tl : device int32 [B,2] (top, left) of the outH x outW window in the (virtually) padded image, or NULL
 *           for (0,0); pad_t / pad_l: rows / columns of zero padding torchvision's RandomCrop(pad_if_needed)
 *           put before the image (the window may also run past the bottom / right edge into padding).
 * remove_ts     : zero the middle channel (RemoveTimesurface; C == 3 only).
 * hot_num_stds  : < 0 off; else zero, in both polarity channels, every pixel where either exceeds
 *                 mean + num_stds * std (unbiased) of the cropped polarity channels (RemoveHotPixels).
 * normalize     : divide the polarity channels by their maximum when it is not 0 (NormalizeEvent).
 * out     : float32 [B,C,outH,outW].  Workspace: memb_raster_post_workspace_bytes(B). */
/* The whole chain in ONE kernel, one CTA per stream, for output rasters that fit a shared-memory tile
 * (outH * outW <= 51200 pixels, e.g. the 224 x 224 training crop): augment -> rasterise at H x W -> crop ->
 * /255 -> RemoveTimesurface -> RemoveHotPixels -> NormalizeEvent -> float32 [B,C,outH,outW].  The uint8 image
 * never reaches HBM.  Arguments as memb_hist_aug_u8 / memb_raster_post_f32 (the middle channel of C == 3 is
 * always zero: no time surface on this path).  ws: >= 256 bytes (status header for memb_hist_status).
 * Returns MEMB_EINVAL when the output raster is too large: use the two calls above instead. */
int memb_event_pipeline_f32(const double* ev, int64_t n, const int64_t* offsets, int B, const memb_event_aug* aug,
                            const int32_t* crop_tl, int H, int W, int pad_t, int pad_l, int outH, int outW, int C,
                            float hot_num_stds, int normalize, float* out, void* ws, size_t ws_bytes,
                            memb_stream_t stream);

/* memb_event_pipeline_f32 with a value table: count c of a surviving pixel is written as value_lut[c] instead of c / 255
 * (device float32 [256], value_lut[0] must be 0).  This is how LogTransform / GammaTransform (mem/transforms.py:200-222,
 * applied between RemoveHotPixels and NormalizeEvent, mem/datasets.py:647-652) ride on the fused kernel bit-exactly: the image
 * holds 256 distinct values, so the host evaluates torch.log(x + 1) / x ** gamma on them once with the reference's own CPU
 * routines; the hot-pixel statistics keep using c / 255, NormalizeEvent divides by the table value of the largest count. */
int memb_event_pipeline_lut_f32(const double* ev, int64_t n, const int64_t* offsets, int B, const memb_event_aug* aug,
                                const int32_t* crop_tl, int H, int W, int pad_t, int pad_l, int outH, int outW, int C,
                                float hot_num_stds, int normalize, const float* value_lut, float* out, void* ws,
                                size_t ws_bytes, memb_stream_t stream);

/* The same chain for recordings WITHOUT a fixed sensor size (the N-Caltech101 / N-Cars branch of build_transformNPY,
 * H = W = None): flip width, cull window and raster size are inferred per stream from the rows (int(max) + 1 at the
 * stage where the reference infers them, datasets.py:513-515, :538-541, :571-575), the (H3 x W3) raster is resized to
 * (outH, outW) with torchvision's Resize(BILINEAR, antialias=True) (ATen upsample_bilinear2d_aa), then RemoveTimesurface,
 * RemoveHotPixels (float32 statistics of the resized planes) and NormalizeEvent.  aug[b]: start / count / time_flip /
 * flip_x / cull / shift_x / shift_y are used (scale_*, flip_w, cull_w, cull_h ignored).  canvas_H x canvas_W bounds the
 * recordings' extent (<= 51200 pixels: one shared-memory tile).  One CTA per stream; memb_hist_status afterwards returns
 * MEMB_EINVAL for an empty stream (the reference raises ValueError) or a recording larger than the canvas. */
int memb_event_pipeline_var_f32(const double* ev, int64_t n, const int64_t* offsets, int B, const memb_event_aug* aug,
                                int canvas_H, int canvas_W, int outH, int outW, int C, float hot_num_stds, int normalize,
                                float* out, void* ws, size_t ws_bytes, memb_stream_t stream);

/* ... with LogTransform (logtrafo != 0: log(x + 1)) and / or GammaTransform (gammatrafo != 0: x ** gamma, gamma > 0) between
 * RemoveHotPixels and NormalizeEvent (mem/datasets.py:648-651, mem/transforms.py:200-222).  On this branch they act on the
 * resized float32 planes: evaluated per pixel in double and rounded once, within 1 ulp of torch's float32 log / pow per
 * map; gamma == 0.5 is torch's square-root special case (bit-exact).  timesurface != 0 (C = 3 only): the middle plane is
 * EventArrToImg's time surface of the augmented rows (mem/datasets.py:585-589: last writer per pixel, every polarity,
 * normalised over the surviving rows; RandomTimeFlip reverses order and time), resized like the polarity planes and kept
 * (no RemoveTimesurface, mem/datasets.py:644); the filter, the maps and the normalisation do not touch it. */
int memb_event_pipeline_var_tf_f32(const double* ev, int64_t n, const int64_t* offsets, int B, const memb_event_aug* aug,
                                   int canvas_H, int canvas_W, int outH, int outW, int C, float hot_num_stds, int normalize,
                                   int logtrafo, int gammatrafo, float gamma, int timesurface, float* out, void* ws,
                                   size_t ws_bytes, memb_stream_t stream);

/* EventRandAugment on uint8 [pos, 0, neg] images (mem/transforms.py:351-471, applied between ToUnit8 and ToFloat32 at the
 * end of build_transformNPY when args.rand_aug is set, mem/datasets.py:655-658).  The HOST draws each sample's operations
 * in the reference's generator order (three torch.randint calls per operation) and describes them with one record per
 * (sample, operation); the kernel applies them in order, one CTA per sample with the image in shared memory
 * (3 * H * W <= 200 KB).  Arithmetic follows torchvision's tensor code path operation by operation: the photometric
 * operations are bit-exact with the reference, the bilinear resampling of the geometric ones up to the summation order
 * of its affine-grid GEMM (a few pixels per image can differ by one count). */
#define MEMB_RA_IDENTITY 0
#define MEMB_RA_AFFINE 1        /* ShearX, ShearY, TranslateX, TranslateY, Rotate: theta = torchvision's inverse matrix      */
#define MEMB_RA_BRIGHTNESS 2    /* f0 = float32(ratio), f1 = float32(1 - ratio), ratio = 1 + magnitude (also 3, 4, 5)      */
#define MEMB_RA_COLOR 3
#define MEMB_RA_CONTRAST 4
#define MEMB_RA_SHARPNESS 5
#define MEMB_RA_POSTERIZE 6     /* ival = bits                                                                             */
#define MEMB_RA_SOLARIZE 7      /* f0 = threshold                                                                          */
#define MEMB_RA_AUTOCONTRAST 8
#define MEMB_RA_EQUALIZE 9
typedef struct memb_randaug_op {  /* 40 bytes, device memory, [B][num_ops] */
  int32_t op;
  int32_t ival;
  float f0, f1;
  float theta[6];
} memb_randaug_op;
/* in  : uint8 [B,3,H,W], or float32 [B,3,H,W] converted like ToUnit8 ((255 * x).to(uint8)) when in_f32 != 0.
 * out : uint8 [B,3,H,W], or float32 converted like ToFloat32 (x / 255) when out_f32 != 0; may alias `in`. */
int memb_event_randaug(const void* in, int in_f32, int B, int C, int H, int W, const memb_randaug_op* ops, int num_ops,
                       void* out, int out_f32, memb_stream_t stream);

size_t memb_raster_post_workspace_bytes(int B);
int memb_raster_post_f32(const uint8_t* hist, int B, int H, int W, int C, const int32_t* crop_tl, int pad_t,
                         int pad_l, int outH, int outW, int remove_ts, float hot_num_stds, int normalize,
                         float* out, void* ws, size_t ws_bytes, memb_stream_t stream);
/* ... with LogTransform / GammaTransform (mem/transforms.py:200-222) as the 256-entry value table of
 * memb_event_pipeline_lut_f32 (NULL: c / 255): applied to the polarity channels after the filter, the middle channel (time
 * surface, when kept) is left at c / 255 as in the reference. */
int memb_raster_post_lut_f32(const uint8_t* hist, int B, int H, int W, int C, const int32_t* crop_tl, int pad_t,
                             int pad_l, int outH, int outW, int remove_ts, float hot_num_stds, int normalize,
                             const float* value_lut, float* out, void* ws, size_t ws_bytes, memb_stream_t stream);


/* ------------------------------------------------------------------------
 * Raw recording formats.  Replaces the byte-by-byte Python decoders of
 * process_data/process_dataset.py: ncaltech101 (:24-63, 40-bit big-endian records) and ncars (:66-103,
 * Prophesee .dat: uint32 timestamp + uint32 data, little endian; the '%' header lines and the two
 * type/size bytes are skipped by the caller, mem_b200/process_data.py).
 * ---------------------------------------------------------------------- */
#define MEMB_RAW_NCALTECH101 1 /* 5 bytes: col0 = b0, col1 = b1, p = 2*(b2>>7)-1, t = ((b2&0x7f)<<16)|(b3<<8)|b4 */
#define MEMB_RAW_NCARS 2       /* 8 bytes: t = u32, d = u32: col0 = d&0x3fff, col1 = (d>>14)&0x3fff, p = (d>>28)&1 */

/* raw: device bytes, 16-byte aligned, n_records records.  out: float64 [n_records,4] rows exactly as the
 * reference's .npy files hold them ([col0, col1, t, p]), 32-byte aligned. */
int memb_decode_events_f64(const uint8_t* raw, int64_t n_records, int format, double* out, memb_stream_t stream);

/* Rasterise one recording straight from its raw records (no float64 rows in HBM): identical to
 * memb_hist_u8(decode(raw)), no time surface.  Workspace: at least
 * memb_hist_workspace_bytes(1, n_records, H, W, 0, MEMB_HIST_GLOBAL); with the MEMB_HIST_AUTO size a long recording
 * on a sensor of <= 51200 pixels takes the PRIVATE strategy (records decoded inside it).  Status via memb_hist_status. */
int memb_hist_raw_u8(const uint8_t* raw, int64_t n_records, int format, int H, int W, int C, uint8_t* out,
                     void* ws, size_t ws_bytes, memb_stream_t stream);

/* ------------------------------------------------------------------------
 * tcgen05 GEMM with fused epilogues:  D[M,N] = epi( A[M,K] * B[N,K]^T ).
 * Replaces the cuBLASLt / cuDNN calls behind nn.Linear / nn.Conv2d on the hot
 * path: mem/modeling_finetune.py:61-71 (Mlp), :130-155 (qkv, proj), :203-209
 * (patch embedding), mem/modeling_pretrain.py:126 (lm_head) and the dVAE
 * encoder convolutions eventvae/vae/vae_model.py:29-41,91-101.
 * ---------------------------------------------------------------------- */
#define MEMB_DT_BF16 0
#define MEMB_DT_F32 1 /* as an input type: consumed by the tensor cores as TF32 */

#define MEMB_EPI_STORE 0      /* d = act(alpha*(acc + bias) [+ aux fp32]); optional row remap / mask rows */
#define MEMB_EPI_BIAS_GELU 1  /* d = gelu(acc + bias) (bf16); d2 = acc + bias (bf16, optional)            */
#define MEMB_EPI_RESIDUAL 2   /* d = aux + rowscale[row/g]*colscale[n]*(acc + bias) (fp32); d2 = acc+bias  */
#define MEMB_EPI_ATOMIC_ADD 3 /* d += alpha*acc (fp32, split-K, red.global.add)                           */
#define MEMB_EPI_DGELU 4      /* d = acc * gelu'(aux) (aux = bf16 pre-activation)                          */
#define MEMB_EPI_ARGMAX 5     /* d[row] = max over n of key(acc + bias, n) (uint64, atomicMax)             */
#define MEMB_EPI_STORE_ROWDOT 6 /* d = bf16(acc); rowdot[r / g][n / 64][r % g] = sum over the 64 columns of one head of
                                   d * aux (aux: bf16 [M,N]; g = rows_per_group; N % 64 == 0): the attention backward's
                                   rowsum(dO * O) leaves the GEMM that produces dO (proj dgrad)                */

typedef struct memb_gemm_desc {
  const void* a;  /* a_layout 0: [M,K] row-major (K-major); 1: [K,M] row-major (MN-major, bf16 only) */
  const void* b;  /* b_layout 0: [N,K] row-major (K-major); 1: [K,N] row-major (MN-major, bf16 only) */
  int64_t lda, ldb; /* leading dimensions in elements */
  int32_t m, n, k;
  int32_t a_layout, b_layout;
  int32_t in_dtype;        /* MEMB_DT_BF16 | MEMB_DT_F32 (tf32) */
  int32_t out_dtype;       /* MEMB_DT_BF16 | MEMB_DT_F32 */
  int32_t epilogue;        /* MEMB_EPI_* */
  int32_t splits;          /* split-K factor for MEMB_EPI_ATOMIC_ADD (0 = auto) */
  int32_t block_n;         /* 0 = auto, 128 or 256 */
  int32_t split_precision; /* 1: 3xTF32 -- a = [hi|lo] (M x 2K), b = [hi|lo] (N x 2K), k = logical K */
  int32_t act;             /* MEMB_EPI_STORE: 0 none, 1 relu */
  int32_t out_split;       /* MEMB_EPI_STORE fp32: write tf32 hi at d[row, n] and lo at d[row, N + n]; d2 = full */
  void* d;
  int64_t ldd;
  void* d2;
  int64_t ldd2;
  const float* bias;       /* [N] or NULL */
  const void* aux;
  int64_t ldaux;
  const float* colscale;   /* [N] or NULL */
  const float* rowscale;   /* [ceil(M / rows_per_group)] or NULL */
  int32_t rows_per_group;
  int32_t out_group_rows, out_group_stride, out_row_offset; /* STORE: out row = (r/g)*stride + off + r%g when g > 0 */
  const uint8_t* rowmask;  /* STORE: rows with mask != 0 are replaced by maskvec */
  const float* maskvec;    /* [N] */
  float alpha;             /* 0 is read as 1 */
  const float* alpha_dev;  /* optional device scalar multiplied into alpha (STORE / ATOMIC_ADD) */
  int32_t* err_flag;       /* device int, set before a trap if an internal wait times out (may be NULL) */
  float* colsum;           /* MEMB_EPI_DGELU, bf16 output: optional [N] accumulator, += column sums of d over the M rows (the
                              bias gradient of the layer whose GELU this is, mem/modeling_finetune.py:62-71 backward); fused
                              into the CTA-pair kernel's epilogue (one red.add per column per 32 rows), else a separate pass */
  float* rowdot;           /* MEMB_EPI_STORE_ROWDOT: fp32 [ceil(M / g)][N / 64][g], written (not accumulated) */
} memb_gemm_desc;

int memb_gemm(const memb_gemm_desc* desc, memb_stream_t stream);

/* ------------------------------------------------------------------------
 * dVAE tokenizer (fp32-faithful: every operand is a TF32 hi/lo pair, 3 MMAs per K block).
 * Replaces DiscreteVAE.get_codebook_indices / forward(return_logits) / ResBlock,
 * eventvae/vae/vae_model.py:29-41, 91-101, 153-158, 182-189.
 * ---------------------------------------------------------------------- */
typedef struct memb_conv_desc {
  const float* a_hi;  /* input activations, "pixel-slot" layout [r_slots][x_slots][inner], TF32 hi parts */
  const float* a_lo;  /* same layout, lo parts */
  int32_t inner, x_slots, r_slots;
  int32_t rows_per_img;            /* virtual output rows per image (>= OH); rows >= OH are dropped */
  int32_t taps_y, taps_x;          /* tap grid: K = taps_y*taps_x*inner, tap-major */
  int32_t tap_y0, tap_x0;          /* slot offset of tap (0,0) relative to the output pixel */
  const float* w;                  /* weights [Cout][2K]: hi parts then lo parts, K ordered (tap, inner) */
  const float* bias;               /* [Cout] */
  int32_t B, OH, OW, Cout;
  int32_t relu;
  const float* aux;                /* optional fp32 residual, plain [B*OH*OW][Cout] */
  float* d_full;                   /* optional fp32 result, plain [B*OH*OW][Cout] (may alias aux) */
  uint64_t* keys;                  /* optional: per-pixel argmax keys over Cout (atomicMax; zero before the call) */
  int32_t seg_kblocks;             /* K blocks (32 floats) accumulated in the tensor core before an fp32 RN add; 0 = 4 */
  float* d_hi;                     /* optional result hi / lo parts at  b*sB + f(oy) + f(ox) + n  with         */
  float* d_lo;                     /* f(o) = ((o+pad)>>shift)*s_major + ((o+pad)&(2^shift-1))*s_minor          */
  int64_t sB, sy_major, sy_minor, sx_major, sx_minor;
  int32_t pad, shift;
  int32_t* err_flag;
} memb_conv_desc;
int memb_conv_tf32x3(const memb_conv_desc* desc, memb_stream_t stream);
/* First encoder layer: im2col of the 4x4/s2/p1 window of img fp32 [B,C,H,W] -> hi/lo [B*(H/2)*(W/2)][Kpad],
 * k = c*16+ky*4+kx; optional per-channel (x-mean)/std (DiscreteVAE.norm, vae_model.py:133-141). */
int memb_dvae_im2col_l1(const float* img, int B, int C, int H, int W, int Kpad, const float* mean, const float* stdv,
                        float* a_hi, float* a_lo, memb_stream_t stream);
/* hi = tf32(v), lo = tf32(v - hi)  (weight preparation). */
int memb_split_tf32(const float* src, float* hi, float* lo, int64_t n, memb_stream_t stream);
/* MEMB_EPI_ARGMAX keys -> int64 indices (logits.argmax(dim=1), first maximum wins). */
int memb_argmax_decode(const uint64_t* keys, int64_t* idx, int64_t n, memb_stream_t stream);

/* fp16-pair variant of the same convolution (csrc/conv_f16.cu): operands are fp16 hi / lo parts of value * 2^exp
 * (11 + 11 significand bits like the TF32 pair, half the bytes, twice the MMA rate), CTA-pair tcgen05 MMAs.
 * Exponents are the caller's: a_exp / w_exp scale the inputs, out_exp the hi / lo result; d_full, keys and the
 * residual `aux` stay unscaled fp32.  `absmax` (optional) receives atomicMax(float bits of |result|) so the caller
 * can calibrate / monitor the exponents (fp16 overflows at 65504). */
typedef struct memb_conv16_desc {
  const void* a_hi;   /* fp16 [r_slots][x_slots][inner] */
  const void* a_lo;
  int32_t inner, x_slots, r_slots;   /* inner % 64 == 0 */
  int32_t rows_per_img;
  int32_t taps_y, taps_x;
  int32_t tap_y0, tap_x0;
  const void* w;      /* fp16 [Cout][2K]: hi parts then lo parts of weight * 2^w_exp */
  const float* bias;  /* [Cout], unscaled */
  int32_t B, OH, OW, Cout;
  int32_t relu;
  int32_t a_exp, w_exp, out_exp;
  const float* aux;
  float* d_full;
  uint64_t* keys;
  int32_t seg_kblocks;   /* K blocks (64 halves) per tensor-core accumulation segment; 0 = 2 */
  void* d_hi;            /* fp16 result parts, addressing as in memb_conv_desc (strides in elements) */
  void* d_lo;
  int64_t sB, sy_major, sy_minor, sx_major, sx_minor;
  int32_t pad, shift;
  uint32_t* absmax;
  int32_t* err_flag;
} memb_conv16_desc;
int memb_conv_f16x2(const memb_conv16_desc* desc, memb_stream_t stream);
/* First encoder layer im2col with fp16 hi / lo of value * 2^exp2 (Kpad % 64 == 0); absmax as above (of the
 * normalised pixels). */
int memb_dvae_im2col_l1_f16(const float* img, int B, int C, int H, int W, int Kpad, const float* mean, const float* stdv,
                            int exp2, void* a_hi, void* a_lo, uint32_t* absmax, memb_stream_t stream);
/* hi = fp16(v * 2^exp2), lo = fp16(v * 2^exp2 - hi)  (weight preparation). */
int memb_split_f16(const float* src, int exp2, void* hi, void* lo, int64_t n, memb_stream_t stream);

/* ------------------------------------------------------------------------
 * Masked-ViT step kernels (HBM-bound pieces).  bf16 pointers are `void*`.
 * Reference: mem/modeling_finetune.py:166-189 (Block: pre-LN, LayerScale, DropPath),
 * :203-247 (PatchEmbed, RelativePositionBias); mem/modeling_pretrain.py:97-126.
 * ---------------------------------------------------------------------- */

/* y(bf16)[r] = LN(x[row_index ? row_index[r] : r]) * gamma + beta; saves mean / rstd.  With `count`
 * (device int) rows >= *count are written as zeros.  D % 128 == 0.  nn.LayerNorm, modeling_finetune.py:166,172. */
int memb_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int rows, int D,
                       void* y, int64_t ldy, float* mean, float* rstd, const int32_t* row_index, const int32_t* count,
                       memb_stream_t stream);
/* dx[target row] += LN backward of dy (bf16 or fp32); dgamma += ..., dbeta += ... (may be NULL). */
int memb_layernorm_bwd(const void* dy, int dy_dtype, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                       const float* mean, const float* rstd, int rows, int D, float* dx, int64_t lddx, float* dgamma,
                       float* dbeta, const int32_t* row_index, const int32_t* count, memb_stream_t stream);
/* memb_layernorm_bwd (without row_index / count) followed, on the updated dx row while it is still in registers, by
 * memb_branch_bwd of the NEXT sub-block of the backward pass with gout = dx (see below): one launch and one read of the
 * fp32 residual gradient less per sub-block.  4 * 8 * D floats of shared memory per block: D <= 896. */
int memb_layernorm_bwd_branch(const void* dy, int dy_dtype, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                              const float* mean, const float* rstd, int rows, int D, float* dx, int64_t lddx,
                              float* dgamma, float* dbeta, const void* branch, int64_t ldb, const float* colscale,
                              const float* rowscale, int rows_per_group, void* dz, int64_t lddz, float* dcolscale,
                              float* dbias, memb_stream_t stream);
/* Backward of x_out = x_in + rowscale[row/g]*colscale[n]*branch (modeling_finetune.py:187-188):
 * dz(bf16) = rowscale*colscale*gout; dcolscale += sum rowscale*gout*branch; dbias += sum dz. */
int memb_branch_bwd(const float* gout, int64_t ldg, const void* branch, int64_t ldb, const float* colscale,
                    const float* rowscale, int rows_per_group, int rows, int D, void* dz, int64_t lddz, float* dcolscale,
                    float* dbias, memb_stream_t stream);
/* out[n] += sum over rows of x(bf16)[row, n]  (bias gradients). */
int memb_colsum_bf16(const void* x, int64_t ld, int rows, int N, float* out, memb_stream_t stream);
/* v_bias gradient from the proj bias gradient of the same backward pass (softmax rows sum to one: colsum(dV) = colsum(dAO) =
 * colsum(dZ) W_proj; Attention.forward, modeling_finetune.py:128-157): dv_bias[j] += sum_i t[i] * w_proj[i, j],
 * dproj_bias[i] += t[i].  t: fp32 [D] (this pass's column sum of the proj output gradient), w_proj: fp32 [D, D] as stored. */
int memb_vbias_chain(const float* t, const float* w_proj, int D, float* dproj_bias, float* dv_bias, memb_stream_t stream);
/* img fp32 [B,C,H,W] -> bf16 [B*(H/P)*(W/P), C*P*P] in Conv2d weight order (PatchEmbed.proj, modeling_finetune.py:203,209). */
int memb_patchify(const float* img, int B, int C, int H, int W, int P, void* out, memb_stream_t stream);
/* x[b,0,:] = cls (+pos[0]); x[b,n,:] += pos[n] if pos (modeling_pretrain.py:101-112). */
int memb_cls_pos(float* x, const float* cls, const float* pos, int B, int N, int D, memb_stream_t stream);
/* Backward of the embedding assembly (mask-token blend, cls concat, pos add). */
int memb_embed_bwd(const float* g0, const uint8_t* mask, int B, int P, int D, void* dpatch, float* dmask_token, float* dcls,
                   float* dbias, float* dpos, memb_stream_t stream);
/* Boolean-mask compaction in row-major (b,p) order (x[:,1:][bool_masked_pos], modeling_pretrain.py:126). */
int memb_mask_compact(const uint8_t* mask, int B, int P, int32_t* row_index, int32_t* patch_index, int32_t* count, int cap,
                      memb_stream_t stream);
/* Mean cross entropy over the *count live rows + top-1 hits + dlogits (engine_for_pretraining.py:152,233).
 * stats[0] += sum loss_i, stats[1] += hits, stats[2] = count. */
int memb_cross_entropy(const float* logits, int64_t ld, const int64_t* tokens, const int32_t* patch_index,
                       const int32_t* count, int cap, int V, void* dlogits, int64_t ldd, float* stats, float grad_scale,
                       memb_stream_t stream);
/* RelativePositionBias.forward (modeling_finetune.py:242-247): bias[h][q][k] (row stride ldk, padded) and its transpose. */
int memb_relpos_gather(const float* table, const int64_t* index, int N, int heads, int ldk, float* bias, float* biasT,
                       memb_stream_t stream);
int memb_relpos_scatter(const float* dbias, int ldk, const int64_t* index, int N, int heads, float* dtable,
                        memb_stream_t stream);
/* acc[i] += sum_b x(bf16)[b*inner + i]. */
int memb_batch_reduce_bf16(const void* x, int B, int64_t inner, float* acc, memb_stream_t stream);

/* Fused attention with additive bias on tcgen05, N <= 208, head_dim == 64 (Attention.forward,
 * modeling_finetune.py:128-157).  `bias` / `biasT` (nullable, together) are the bias and its transpose in the packed
 * layout of memb_attention_pack_bias: fp32 [H][MEMB_ATTN_BIAS_GROUPS][256][4] = (head, group of 4 columns, row, column
 * in group), pre-multiplied by log2(e), -inf in columns >= N.  lse is the natural-log row logsumexp [B, H, N].
 * The backward writes dsT = dS^T as bf16 [B, H, N (key), ld_ds (query)] (nullable) for the bias gradient. */
#define MEMB_ATTN_BIAS_GROUPS 52
#define MEMB_ATTN_BIAS_FLOATS_PER_HEAD (MEMB_ATTN_BIAS_GROUPS * 256 * 4)
int memb_attention_pack_bias(const float* dense /* [H, N, ld] */, int ld, int N, int heads, float* packed,
                             memb_stream_t stream);
int memb_attention_fwd(const void* qkv, const float* bias, int ld_ds, int B, int N, int H, int head_dim, float scale,
                       void* out, float* lse, memb_stream_t stream);
/* delta[b][h][q] = sum over d of a[b, q, h, d] * b[b, q, h, d] (bf16 [B*N, H*64] both): rowsum(dO * O) of the attention
 * backward as a separate pass. */
int memb_rowdot_heads(const void* a, const void* b, int B, int N, int H, float* delta, memb_stream_t stream);
/* out == NULL: the first B*H*N floats of the workspace already hold rowsum(dO * O) (memb_rowdot_heads layout), e.g. left
 * there by the GEMM that produced dout (MEMB_EPI_STORE_ROWDOT); otherwise the call computes it from out and dout. */
size_t memb_attention_bwd_workspace_bytes(int B, int N, int H);
int memb_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const float* bias,
                       const float* biasT, int ld_ds, int B, int N, int H, int head_dim, float scale, void* dqkv,
                       void* dsT, void* workspace, size_t ws_bytes, memb_stream_t stream);

/* Classification head of ft_vit (VisionTransformer.forward_features / forward with mean pooling,
 * modeling_finetune.py:343-352, 354-357): pooled = mean over patch tokens; small fp32 nn.Linear (D -> num_classes). */
int memb_meanpool_fwd(const float* x /* [B*N, D] */, int B, int N, int D, float* out /* [B, D] */, memb_stream_t stream);
int memb_meanpool_bwd(const float* dpool /* [B, D] */, int B, int N, int D, float* gres /* [B*N, D], overwritten */,
                      memb_stream_t stream);
int memb_linear_small_fwd(const void* z_bf16 /* [B, D] */, const float* W /* [C, D] */, const float* bias, int B, int D, int C,
                          float* out /* [B, C] */, memb_stream_t stream);
int memb_linear_small_bwd(const float* dl /* [B, C] */, const void* z_bf16, const float* W, int B, int D, int C,
                          float* dW /* += */, float* db /* +=, nullable */, float* dz /* [B, D], overwritten */,
                          memb_stream_t stream);

/* Classification criteria of the finetuning loop on [B, C] fp32 logits (mem/run_class_finetuning.py:551-559 picks
 * timm's SoftTargetCrossEntropy under mixup, LabelSmoothingCrossEntropy for smoothing > 0, else CrossEntropyLoss):
 * loss_out[0] += mean_i sum_c t_ic * (logsumexp_i - x_ic), dlogits (nullable) = d(mean loss)/d logits.
 * Exactly one of labels (int64 [B], t = (1 - smoothing) * onehot + smoothing / C) and soft_targets (fp32 [B, C]). */
int memb_soft_ce(const float* logits, int B, int C, const int64_t* labels, const float* soft_targets, float smoothing,
                 float* loss_out, float* dlogits, memb_stream_t stream);

/* Flat-buffer optimizer pass (mem/utils.py:357-371 + torch.optim.AdamW with optim_factory.py:121 betas). */
int memb_fill_f32(float* p, int64_t n, float v, memb_stream_t stream);
int memb_cast_bf16(const float* src, void* dst, int64_t n, memb_stream_t stream);
int memb_sqnorm(const float* g, int64_t n, float scale, float* out, memb_stream_t stream);
/* out[0] += sum (g * scale)^2 over the 1024-element chunks whose chunk_group is not 255 (255 = padding or a
 * requires_grad = False tensor: get_grad_norm_ / clip_grad_norm_ only see trainable tensors, mem/utils.py:380-392). */
int memb_sqnorm_groups(const float* g, int64_t n, float scale, const uint8_t* chunk_group, float* out,
                       memb_stream_t stream);
/* One pass over the flat parameter buffer; tensors start on 1024-element boundaries and
 * chunk_group[i/1024] (device uint8, >= 64 = never updated) selects (group_lr_host[g], group_wd_host[g]).
 * When sqnorm_dev is given and *sqnorm_dev is not finite the step is skipped (nothing is written): the
 * reference's GradScaler.step skips inf / NaN gradients (mem/utils.py:357-371). */
int memb_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, const uint8_t* chunk_group,
               const float* group_lr_host, const float* group_wd_host, int ngroups, float beta1, float beta2, float eps,
               int step, float grad_scale, float max_norm, const float* sqnorm_dev, memb_stream_t stream);

/* ------------------------------------------------------------------------
 * dVAE training step and decoder.  Replaces DiscreteVAE.forward(return_loss / return_recons) and
 * DiscreteVAE.decode, eventvae/vae/vae_model.py:160-213, as driven by eventvae/train_vae.py:304-392.  Activations are
 * NHWC bf16 matrices [B*H*W, C]; the convolutions are im2col / col2im around memb_gemm (bf16 operands, fp32 accumulate):
 * Conv2d fwd = im2col + GEMM, dgrad = GEMM + col2im, wgrad = GEMM on im2col; ConvTranspose2d swaps the roles.
 * ---------------------------------------------------------------------- */
/* fp32 [B,C,H,W] -> bf16 [B,H,W,Cpad] (zero channels beyond C), (x - mean[c]) / std[c] when given (DiscreteVAE.norm, :133-141);
 * out_f32 (nullable): the normalised image as fp32 [B,H,W,C], the reconstruction target. */
int memb_vae_nchw_to_nhwc(const float* img, int B, int C, int H, int W, int Cpad, const float* mean, const float* stdv,
                          void* out_bf16, float* out_f32, memb_stream_t stream);
/* fp32 [B,H,W,ld] (first C channels) -> fp32 [B,C,H,W] */
int memb_vae_nhwc_to_nchw(const float* x, int B, int C, int H, int W, int ld, float* out, memb_stream_t stream);
/* col[(b,oy,ox), (ky,kx,c)] = x[b, oy*stride+ky-pad, ox*stride+kx-pad, c], zero outside; C % 8 == 0. */
int memb_vae_im2col(const void* x_bf16, int B, int H, int W, int C, int kh, int kw, int stride, int pad, void* col_bf16,
                    memb_stream_t stream);
/* The adjoint of im2col as a gather (no atomics): out[b,y,x,c] = sum of the col entries that im2col would have read from that
 * element, + bias[c], ReLU, then zeroed where gate <= 0 (ReLU backward of the producing layer).  col fp32 [B*OH*OW, ldc]
 * (OH, OW from H, W, k, stride, pad); out bf16 and / or fp32 [B,H,W,ld_out]. */
int memb_vae_col2im(const float* col, int64_t ldc, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                    const float* bias, int relu, const void* gate_bf16, void* out_bf16, float* out_f32, int ld_out,
                    memb_stream_t stream);
/* op 0: out = relu(a); op 1: out = a where b > 0 else 0; op 2: out = a + b (bf16, n elements) */
int memb_vae_ew_bf16(int op, const void* a, const void* b, void* out, int64_t n, memb_stream_t stream);
/* out[i, :] = table[idx[i], :] as bf16 [n, ld] (nn.Embedding lookup of decode(), :164); err_flag set on an index outside [0, V). */
int memb_vae_gather_rows(const float* table, int V, int D, const int64_t* idx, int64_t n, int ld, void* out_bf16, int* err_flag,
                         memb_stream_t stream);
/* F.gumbel_softmax(logits, tau, dim=token axis, hard) on rows of logits fp32 [rows, N] with the Gumbel noise given (fp32, same
 * shape): y = softmax((logits + noise) / tau) (bf16); hard: y_fwd = onehot(argmax) (straight-through forward value).
 * lse: fp32 [rows, 2] saved for the backward pass.  kl_out[0] += sum_rows sum_n q_n (log q_n + log N), q = softmax(logits)
 * (the reference's F.kl_div(log_uniform, log_qy, 'batchmean', log_target=True), :204-208). */
int memb_vae_gumbel_fwd(const float* logits, const float* noise, int64_t rows, int N, float tau, int hard, void* y_bf16,
                        void* y_fwd_bf16, float* lse, float* kl_out, memb_stream_t stream);
/* dlogits (bf16) = d/dlogits [ <dy, y> + kl_weight * KL ]; dy fp32 [rows, N]. */
int memb_vae_gumbel_bwd(const float* logits, const float* noise, const float* lse, const float* dy, int64_t rows, int N, float tau,
                        const float* grad_scale_dev, float kl_weight, void* dlogits_bf16, memb_stream_t stream);
/* loss_out[0] += mse_loss (kind 0) or smooth_l1_loss (kind 1) of recon fp32 [pixels, ld_recon] (first C channels) against
 * target fp32 [pixels, C]; dout (nullable, bf16 [pixels, ld_recon]) = d loss / d recon. */
int memb_vae_recon_loss(const float* target, const float* recon, int64_t pixels, int C, int ld_recon, int kind, float* loss_out,
                        void* dout_bf16, memb_stream_t stream);
/* dst = (accumulate ? dst : 0) + scale_dev[0] * src (fp32; scale_dev NULL = 1) */
int memb_axpy_f32(const float* src, const float* scale_dev, float* dst, int64_t n, int accumulate, memb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MEMB_H_ */
