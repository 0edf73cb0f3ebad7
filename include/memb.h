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

#ifdef __cplusplus
}
#endif
#endif /* MEMB_H_ */
