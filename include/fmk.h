/*
 * fmk.h -- C ABI of libfmk.so: the B200-native tick -> bars -> features -> labels hot path.
 *
 * The reference (quantscious/finmlkit v0.1.11) is pure Python + Numba and has NO FFI; this header is the boundary a
 * maintainer would bind with ctypes.  Every entry point cites the reference function it replaces
 * (paths relative to the reference tree).  Conventions:
 *   - plain C types only; host arrays are caller-allocated (NumPy), device memory lives behind opaque handles;
 *   - every function returns an int status: 0 = ok, negative = fmk_status (fmk_last_error(ctx) gives the text, which
 *     for argument errors is the reference's own ValueError message);
 *   - a ctx is bound to one device and one stream and is single-threaded; distinct contexts are independent;
 *   - there is no CPU fallback anywhere: without a CUDA device every compute entry point fails with FMK_ERR_CUDA.
 */
#ifndef FMK_H
#define FMK_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    FMK_OK = 0,
    FMK_ERR_CUDA = -1,        /* CUDA runtime error / no device */
    FMK_ERR_ARG = -2,         /* invalid argument (message = the reference's ValueError text) */
    FMK_ERR_ALLOC = -3,
    FMK_ERR_CAPACITY = -4,    /* caller buffer too small */
    FMK_ERR_LEVEL = -5,       /* "Something went wrong! Invalid price level index!" (bar/base.py:719) */
    FMK_ERR_INTERNAL = -6
} fmk_status;

typedef struct fmk_ctx fmk_ctx;         /* device + stream + scratch */
typedef struct fmk_trades fmk_trades;   /* device SoA: ts i64[n], price f64[n], amount f64[n], side i8[n] */
typedef struct fmk_index fmk_index;     /* device bar-close arrays: close_ts i64[m], close_idx i64[m] (m = n_bars + 1) */
typedef struct fmk_buf fmk_buf;         /* generic device array */
typedef struct fmk_footprint fmk_footprint; /* device CSR footprint (bar/data_model.py:775 FootprintData) */

/* ---- library / context ------------------------------------------------------------------------------------------ */
const char *fmk_version(void);
int fmk_device_count(void);
int fmk_ctx_create(int device, fmk_ctx **out);
/* Same, but every kernel is launched on the caller's CUDA stream (cudaStream_t passed as void*), so a host that
 * already owns a stream (e.g. the one NCCL collectives are enqueued on) gets one ordered timeline. */
int fmk_ctx_create_on_stream(int device, void *stream, fmk_ctx **out);
void fmk_ctx_destroy(fmk_ctx *ctx);
const char *fmk_last_error(fmk_ctx *ctx);
int fmk_ctx_sync(fmk_ctx *ctx);
/* Large device scratch blocks are cached per ctx between calls (a repeated build never allocates); this releases them. */
int fmk_ctx_trim(fmk_ctx *ctx);
/* CUDA-event timer on the ctx stream (the stream every kernel of this ctx is launched on). */
int fmk_timer_start(fmk_ctx *ctx);
int fmk_timer_stop(fmk_ctx *ctx, float *ms_out);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
int64_t fmk_launch_count(fmk_ctx *ctx);
/* per-kernel CUDA-event timing: enable, run, then drain (names_out: cap*64 bytes; returns number of distinct kernels) */
int fmk_prof_enable(fmk_ctx *ctx, int on);
int fmk_prof_report(fmk_ctx *ctx, char *names_out, int64_t *counts_out, float *ms_out, int cap);
/* device columns of the last fmk_bar_ohlcv_device call (for a host that gathers bar frames over NCCL) */
int fmk_result_cols(fmk_ctx *ctx, void **ptr, int64_t *n_bars, int64_t *bytes);
/* Writes > L2-size bytes so the next timed step starts with a cold L2. */
int fmk_flush_l2(fmk_ctx *ctx);
/* pinned host memory for the end-to-end path */
int fmk_host_alloc(void **out, int64_t bytes);
void fmk_host_free(void *p);

/* ---- trades: TradesData columns (bar/data_model.py:121-244) as device SoA --------------------------------------- */
/* side may be NULL (directional/footprint calls then fail with FMK_ERR_ARG). amount is float64.
 * ts may be NULL for tick/volume/dollar bars + reductions: close timestamps are then gathered by the host from its own
 * array (ts[close_idx]), which saves a third of the H2D traffic; time/CUSUM bars, lagged returns, ewmst and TBM need ts. */
int fmk_trades_upload(fmk_ctx *ctx, const int64_t *ts, const double *price, const double *amount, const int8_t *side,
                      int64_t n, fmk_trades **out);
/* Same with float32 amounts (TradesData after the reference's split-trade merge holds float32 amounts,
 * bar/data_model.py:326-344): 4 B/tick over PCIe, widened exactly to float64 on the device. */
int fmk_trades_upload_f32amt(fmk_ctx *ctx, const int64_t *ts, const double *price, const float *amount, const int8_t *side,
                             int64_t n, fmk_trades **out);
/* Month-store loader (the reference's /trades/YYYY-MM tables, bar/data_model.py:420-574): allocate one handle for a whole
 * time range, then write each partition's columns at its offset (any pointer may be NULL to skip that column; amount may be
 * float32, widened exactly on the device).  No concatenated host frame is needed. */
int fmk_trades_alloc(fmk_ctx *ctx, int64_t n, int with_ts, int with_side, fmk_trades **out);
int fmk_trades_write(fmk_ctx *ctx, fmk_trades *t, int64_t offset, int64_t count, const int64_t *ts, const double *price,
                     const void *amount, int amount_is_f32, const int8_t *side);
/* Adds a column to a handle uploaded without it (which: 0 = timestamps int64[n], 1 = side int8[n]). */
int fmk_trades_add_column(fmk_ctx *ctx, fmk_trades *t, int which, const void *host);
/* Device-side synthetic BTCUSDT-like stream (SURVEY 8d shape) for bench-size runs. */
int fmk_trades_synth(fmk_ctx *ctx, int64_t n, uint64_t seed, fmk_trades **out);
/* Re-fill an existing handle from host arrays (async H2D on the ctx stream; arrays should be pinned). */
int fmk_trades_refill(fmk_ctx *ctx, fmk_trades *t, const int64_t *ts, const double *price, const double *amount,
                      const int8_t *side, int64_t n);
int fmk_trades_download(fmk_ctx *ctx, const fmk_trades *t, int64_t *ts, double *price, double *amount, int8_t *side);
int64_t fmk_trades_size(const fmk_trades *t);
void fmk_trades_free(fmk_ctx *ctx, fmk_trades *t);

/* ---- generic device arrays -------------------------------------------------------------------------------------- */
int fmk_buf_upload(fmk_ctx *ctx, const void *host, int64_t bytes, fmk_buf **out);
int fmk_buf_alloc(fmk_ctx *ctx, int64_t bytes, fmk_buf **out);
int fmk_buf_download(fmk_ctx *ctx, const fmk_buf *b, void *host, int64_t bytes);
int64_t fmk_buf_bytes(const fmk_buf *b);
void *fmk_buf_devptr(const fmk_buf *b);
void fmk_buf_free(fmk_ctx *ctx, fmk_buf *b);
/* out[k] = src[idx[k]] for m host indices into a device array of 8-byte elements (negative indices wrap like NumPy);
 * e.g. CUSUMBarKit.get_sigma, bar/kit.py:176-181 (sigma[bar_close_indices]) without downloading sigma. */
int fmk_buf_gather8(fmk_ctx *ctx, const fmk_buf *src, const int64_t *idx, int64_t m, void *out);

/* ---- bar indexers (bar/logic.py) -> device index handle ----------------------------------------------------------
 * close_idx follows the reference exactly: element 0 is the "open" marker (-1 possible for time bars). */
int fmk_time_bar_index(fmk_ctx *ctx, const fmk_trades *t, double interval_seconds, fmk_index **out); /* logic.py:12-51 */
int fmk_tick_bar_index(fmk_ctx *ctx, const fmk_trades *t, int64_t threshold, fmk_index **out);       /* logic.py:54-84 */
int fmk_volume_bar_index(fmk_ctx *ctx, const fmk_trades *t, double threshold, fmk_index **out);      /* logic.py:87-115 */
int fmk_dollar_bar_index(fmk_ctx *ctx, const fmk_trades *t, double threshold, fmk_index **out);      /* logic.py:118-149 */
/* sigma: device f64[n] buffer; it is forward-filled in place like the reference (logic.py:181-189). */
int fmk_cusum_bar_index(fmk_ctx *ctx, const fmk_trades *t, fmk_buf *sigma, double sigma_floor, double sigma_mult,
                        fmk_index **out);                                                             /* logic.py:152-221 */
/* number of NaNs of sigma the last fmk_cusum_bar_index call forward-filled (0: the buffer was left untouched) */
int64_t fmk_cusum_filled_count(fmk_ctx *ctx);
/* Tick-imbalance (kind 0) / tick-run (kind 1) bars.  The reference only has stubs (logic.py:224-261 raise
 * NotImplementedError), so the semantics are this library's own and are pinned by its own CPU oracle only ("parity
 * unpinned"): b_t = side column (use_side != 0 and present) or the tick rule on the prices (bar/utils.py:12-46); the index
 * list starts with 0; kind 0 closes at the first tick with |sum b_t since the last close| >= threshold, kind 1 at the first
 * tick with max(#buys, #sells since the last close) >= threshold; the accumulators restart from 0 after a close. */
int fmk_imbalance_bar_index(fmk_ctx *ctx, const fmk_trades *t, double threshold, int use_side, int kind, fmk_index **out);
/* wrap caller-provided close indices (comp_bar_* called directly, MockBarBuilder-style tests) */
int fmk_index_from_host(fmk_ctx *ctx, const fmk_trades *t, const int64_t *close_idx, int64_t m, fmk_index **out);
int64_t fmk_index_size(const fmk_index *ix);      /* m = n_bars + 1 */
int fmk_index_download(fmk_ctx *ctx, const fmk_index *ix, int64_t *close_ts, int64_t *close_idx);
void fmk_index_free(fmk_ctx *ctx, fmk_index *ix);
/* diagnostics of the last dollar/volume index build: [0]=speculative tasks, [1]=serial repairs, [2]=passes */
int fmk_index_stats(fmk_ctx *ctx, int64_t *stats3);

/* ---- per-bar reductions (bar/base.py:303-850); outputs are host arrays of n_bars elements ----------------------- */
/* comp_bar_ohlcv, base.py:306-407.  Any output pointer may be NULL to skip its D2H copy. */
int fmk_bar_ohlcv(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, double *open, double *high, double *low,
                  double *close, float *volume, double *vwap, int64_t *trades, double *median_trade_size);
/* Device-resident variant: computes into ctx-owned device columns and returns nothing to the host (bench `value`). */
int fmk_bar_ohlcv_device(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, int with_median);
/* comp_bar_directional_features, base.py:409-546 (tuple order of the reference) */
int fmk_bar_directional(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, int64_t *ticks_buy, int64_t *ticks_sell,
                        float *volume_buy, float *volume_sell, float *dollars_buy, float *dollars_sell,
                        float *mean_spread, float *max_spread, int64_t *cum_ticks_min, int64_t *cum_ticks_max,
                        float *cum_volume_min, float *cum_volume_max, float *cum_dollars_min, float *cum_dollars_max);
/* comp_bar_trade_size_features, base.py:549-612 */
int fmk_bar_trade_size(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, const double *theta, int64_t n_theta,
                       double theta_mult, float *mean_size_rel, float *size_95_rel, float *pct_block, float *size_gini);
/* comp_bar_footprints + comp_footprint_features, base.py:615-850.  Two-phase: build on device, then download CSR. */
int fmk_bar_footprints(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, double price_tick_size,
                       const double *bar_lows, const double *bar_highs, double imbalance_factor, fmk_footprint **out);
int64_t fmk_footprint_levels(const fmk_footprint *fp);   /* total number of (bar, level) rows */
int fmk_footprint_download(fmk_ctx *ctx, const fmk_footprint *fp, int64_t *level_offsets, int32_t *price_levels,
                           float *buy_volumes, float *sell_volumes, int32_t *buy_ticks, int32_t *sell_ticks,
                           uint8_t *buy_imbalances, uint8_t *sell_imbalances, uint16_t *buy_imb_sum,
                           uint16_t *sell_imb_sum, int32_t *cot_price_level, int16_t *imb_max_run_signed,
                           double *vp_skew, double *vp_gini);
void fmk_footprint_free(fmk_ctx *ctx, fmk_footprint *fp);

/* ---- device-resident bar frame: every per-bar output of BarBuilderBase (bar/base.py:132-300) in one pass ------------
 * build_ohlcv + build_directional_features + build_trade_size_features + build_footprints against one index, results
 * kept on the device in two blocks (per-bar columns, per-level CSR columns) so that a host can download them with two
 * copies or hand them to fmk_comm_gather_submit.  flags select the feature groups; FMK_F_FOOTPRINT needs FMK_F_OHLCV
 * (bar lows / highs are taken from the OHLCV columns on the device, base.py:257-259) and FMK_F_TRADE_SIZE with
 * theta == NULL uses each bar's own median trade size as theta (needs FMK_F_MEDIAN). */
typedef struct fmk_frame fmk_frame;
enum {
    FMK_F_OHLCV = 1, FMK_F_MEDIAN = 2, FMK_F_DIRECTIONAL = 4, FMK_F_TRADE_SIZE = 8, FMK_F_FOOTPRINT = 16
};
typedef enum {
    /* per-bar block */
    FMK_COL_CLOSE_TS = 0, FMK_COL_CLOSE_IDX,                                                  /* i64 */
    FMK_COL_OPEN, FMK_COL_HIGH, FMK_COL_LOW, FMK_COL_CLOSE, FMK_COL_VWAP, FMK_COL_MEDIAN,     /* f64 */
    FMK_COL_TRADES,                                                                           /* i64 */
    FMK_COL_VOLUME,                                                                           /* f32 */
    FMK_COL_TICKS_BUY, FMK_COL_TICKS_SELL, FMK_COL_CUM_TICKS_MIN, FMK_COL_CUM_TICKS_MAX,      /* i64 */
    FMK_COL_VOLUME_BUY, FMK_COL_VOLUME_SELL, FMK_COL_DOLLARS_BUY, FMK_COL_DOLLARS_SELL, FMK_COL_MEAN_SPREAD,
    FMK_COL_MAX_SPREAD, FMK_COL_CUM_VOLUME_MIN, FMK_COL_CUM_VOLUME_MAX, FMK_COL_CUM_DOLLARS_MIN,
    FMK_COL_CUM_DOLLARS_MAX,                                                                  /* f32 */
    FMK_COL_MEAN_SIZE_REL, FMK_COL_SIZE_95_REL, FMK_COL_PCT_BLOCK, FMK_COL_SIZE_GINI,         /* f32 */
    FMK_COL_FP_LEVEL_OFFSETS,                                                                 /* i64[n_bars + 1] */
    FMK_COL_FP_VP_SKEW, FMK_COL_FP_VP_GINI,                                                   /* f64 */
    FMK_COL_FP_COT,                                                                           /* i32 */
    FMK_COL_FP_BUY_IMB_SUM, FMK_COL_FP_SELL_IMB_SUM,                                          /* u16 */
    FMK_COL_FP_RUN_SIGNED,                                                                    /* i16 */
    /* per-level block (CSR rows) */
    FMK_COL_FP_PRICE_LEVELS,                                                                  /* i32 */
    FMK_COL_FP_BUY_VOL, FMK_COL_FP_SELL_VOL,                                                  /* f32 */
    FMK_COL_FP_BUY_TICKS, FMK_COL_FP_SELL_TICKS,                                              /* i32 */
    FMK_COL_FP_BUY_IMB, FMK_COL_FP_SELL_IMB,                                                  /* u8 */
    FMK_COL_COUNT
} fmk_col;
/* The per-bar block begins with a self-describing header of FMK_FRAME_HEADER_BYTES: int64[0] = FMK_FRAME_MAGIC, [1] = n_bars,
 * [2] = n_levels, [3] = bytes of the per-bar block (header included), [4] = bytes of the per-level block, [5] = flags,
 * [6] = FMK_COL_COUNT, [8 + k] = byte offset of column k inside its block (-1 = absent).  A gathered frame is the per-bar
 * block followed (at the next 16-byte boundary) by the per-level block. */
#define FMK_FRAME_HEADER_BYTES 512
#define FMK_FRAME_MAGIC 0x464D4B4652414D45ll
int fmk_bar_features_device(fmk_ctx *ctx, const fmk_trades *t, const fmk_index *ix, int flags, const double *theta,
                            int64_t n_theta, double theta_mult, double price_tick_size, double imbalance_factor,
                            fmk_frame **out);
/* n_bars, n_levels, byte sizes of the two blocks, and col_offsets[FMK_COL_COUNT] (byte offset of each column inside its
 * block, -1 when the column is absent) */
int fmk_frame_info(const fmk_frame *f, int64_t *n_bars, int64_t *n_levels, int64_t *bar_block_bytes,
                   int64_t *level_block_bytes, int64_t *col_offsets);
int fmk_frame_devptrs(const fmk_frame *f, void **bar_block, void **level_block);
/* either host pointer may be NULL */
int fmk_frame_download(fmk_ctx *ctx, const fmk_frame *f, void *bar_block_host, void *level_block_host);
void fmk_frame_free(fmk_ctx *ctx, fmk_frame *f);

/* ---- multi-GPU: one gather-v of the finished frames to one rank over NCCL ------------------
 * One process per GPU.  Rank 0 calls fmk_comm_unique_id and hands the 128 bytes to the other ranks (any side channel);
 * every rank then calls fmk_comm_init with its own ctx.  libnccl.so.2 is loaded with dlopen at that point.
 * max_ctas > 0 caps the SMs NCCL may occupy.  A gather step packs up to 8 device segments into this rank's frame and
 * sends exactly that many bytes: ncclAllGather of the byte counts, then every rank pushes its frame into the destination's
 * receive buffer (mapped with CUDA IPC) with a peer-to-peer copy over NVLink -- copy engines, no SM -- fenced by a tiny
 * all-gather; grouped ncclSend / ncclRecv is the fallback (FMK_COMM_P2P=0 or a rank that cannot map the buffer).  The transfer
 * of step k overlaps the kernels of step k+1.  fmk_comm_gather_finish makes the ctx stream wait for all transfers. */
typedef struct fmk_comm fmk_comm;
int fmk_comm_unique_id(void *out128);
int fmk_comm_init(fmk_ctx *ctx, const void *id128, int rank, int world, int max_ctas, fmk_comm **out);
void fmk_comm_destroy(fmk_comm *c);
int fmk_comm_rank(const fmk_comm *c);
int fmk_comm_world(const fmk_comm *c);
int fmk_comm_nccl_version(void);
/* 1: payloads travel as peer-to-peer pushes (CUDA IPC + copy engines); 0: grouped ncclSend / ncclRecv */
int fmk_comm_p2p_active(const fmk_comm *c);
int fmk_comm_barrier(fmk_comm *c);
/* op: 0 = max, 1 = min, 2 = sum over ranks of n <= 64 host doubles (in place) */
int fmk_comm_allreduce_f64(fmk_comm *c, double *inout, int n, int op);
int fmk_comm_gather_submit(fmk_comm *c, const void *const *seg_ptrs, const int64_t *seg_bytes, int nseg, int dst);
int fmk_comm_gather_finish(fmk_comm *c);
/* finish + release the staging / receive buffers; the next submit sizes the pipeline from its own frame.  COLLECTIVE. */
int fmk_comm_gather_reset(fmk_comm *c);
int fmk_comm_gather_result(fmk_comm *c, int rank, void **dev_ptr, int64_t *bytes);
int fmk_comm_gather_download(fmk_comm *c, int rank, void *host, int64_t cap);

/* ---- tick-level series (feature/core) --------------------------------------------------------------------------- */
/* comp_lagged_returns, feature/core/utils.py:12-64: host in / host out */
int fmk_lagged_returns(fmk_ctx *ctx, const int64_t *ts, const double *close, int64_t n, double window_sec, int is_log,
                       double *out);
/* ewmst, feature/core/volatility.py:139-219 */
int fmk_ewmst(fmk_ctx *ctx, const int64_t *ts, const double *y, int64_t n, double half_life, double sigma_floor,
              double *out);
/* device-resident variants used by the fused sigma pipeline (ts/price taken from the trades handle) */
int fmk_lagged_returns_dev(fmk_ctx *ctx, const fmk_trades *t, double window_sec, int is_log, fmk_buf **out);
int fmk_ewmst_dev(fmk_ctx *ctx, const fmk_trades *t, const fmk_buf *y, double half_life, double sigma_floor,
                  fmk_buf **out);

/* ---- bar-level features on n_bars-length host arrays ------------------------------------------------- */
int fmk_realized_vol(fmk_ctx *ctx, const double *r, int64_t n, int64_t window, int is_sample, double *out);   /* volatility.py:256-286 */
int fmk_ewms(fmk_ctx *ctx, const double *y, int64_t n, int64_t span, double *out);                               /* volatility.py:9-69 */
int fmk_vpin(fmk_ctx *ctx, const double *volume_buy, const double *volume_sell, int64_t n, int64_t window,
             float *out);                                                                                         /* volume.py:610-641 */
int fmk_flow_acceleration(fmk_ctx *ctx, const double *volumes, int64_t n, int64_t window, int64_t recent_periods,
                          double *out);                                                                           /* volume.py:572-607 */

/* ---- labels: triple_barrier, label/tbm.py:11-158 ------------------------------------------------------------------
 * side may be NULL (side prediction). Skipped events get touch_idx = event_idx (the reference leaves it uninitialised). */
int fmk_triple_barrier(fmk_ctx *ctx, const fmk_trades *t, const int64_t *event_idx, const double *targets,
                       int64_t n_events, int64_t n_targets, double bottom_mult, double top_mult,
                       double vertical_barrier_s, double min_close_time_s, const int8_t *side, int64_t n_side,
                       double min_ret, int8_t *labels, int64_t *touch_idx, double *rets, double *max_rb_ratios);

/* ---- sample weights on ticks (label/weights.py) ----------------------------------------------------------
 * Indices must lie in [0, n).  concurrency is int16 with the reference's silent wrap-around. */
/* average_uniqueness, label/weights.py:7-49: weights f64[n_events], concurrency i16[n] (either may be NULL) */
int fmk_average_uniqueness(fmk_ctx *ctx, int64_t n, const int64_t *event_idx, const int64_t *touch_idx, int64_t n_events,
                           int64_t n_touch, double *weights, int16_t *concurrency);
/* return_attribution, label/weights.py:52-103 (host close / concurrency arrays, as the reference passes them) */
int fmk_return_attribution(fmk_ctx *ctx, const int64_t *event_idx, const int64_t *touch_idx, int64_t n_events,
                           const double *close, const int16_t *concurrency, int64_t n, int normalize, double *weights);
/* both in one pass over the device-resident price column (SampleWeights.compute_info_weights, label/kit.py:329-366) */
int fmk_sample_weights(fmk_ctx *ctx, const fmk_trades *t, const int64_t *event_idx, const int64_t *touch_idx,
                       int64_t n_events, int normalize, double *avg_uniqueness, double *return_attribution,
                       int16_t *concurrency);

/* ---- event sampler and ingest scans ------------------------------------------------------------------- */
/* cusum_filter, sampling/filters.py:6-70: threshold has 1 or n elements; *events_out is a device buffer of n_events int64
 * indices (fmk_buf_download / fmk_buf_free). */
int fmk_cusum_filter(fmk_ctx *ctx, const double *series, int64_t n, const double *threshold, int64_t n_thr,
                     fmk_buf **events_out, int64_t *n_events);
/* comp_trade_side_vector (tick rule), bar/utils.py:12-46 */
int fmk_trade_side_vector(fmk_ctx *ctx, const double *prices, int64_t n, int8_t *sides_out);
/* merge_split_trades, bar/utils.py:263-329: inputs ordered by (timestamp, price, side); outputs are caller-allocated with
 * n elements (the reference allocates n and trims); is_buyer_maker / sides_out may be NULL. */
int fmk_merge_split_trades(fmk_ctx *ctx, const int64_t *ts, const double *prices, const float *amounts,
                           const uint8_t *is_buyer_maker, int64_t n, int64_t *ts_out, double *prices_out,
                           float *amounts_out, int8_t *sides_out, int64_t *n_out);

/* ---- rolling volume profile (feature/core/volume.py:133-456) --------------------------------------------
 * n_bins <= 0 means None (no bucketing).  Outputs have n_bars elements (zeros before the first full window, like the
 * reference): POC / HVA / LVA in integer price-tick units, fraction of the volume above the POC as float32. */
int fmk_volume_profile_rolling(fmk_ctx *ctx, const int64_t *level_offsets, const int32_t *price_levels,
                               const float *buy_volumes, const float *sell_volumes, int64_t n_bars, const int64_t *bar_ts,
                               const double *highs, const double *lows, double window_sec, int64_t n_bins,
                               double price_tick, double va_pct, int32_t *poc, int32_t *hva, int32_t *lva,
                               float *pct_above_poc);
/* same on the device-resident CSR footprint of fmk_bar_footprints */
int fmk_volume_profile_rolling_fp(fmk_ctx *ctx, const fmk_footprint *fp, const int64_t *bar_ts, const double *highs,
                                  const double *lows, double window_sec, int64_t n_bins, double price_tick, double va_pct,
                                  int32_t *poc, int32_t *hva, int32_t *lva, float *pct_above_poc);

#ifdef __cplusplus
}
#endif
#endif
