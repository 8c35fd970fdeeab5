/*
 * rvdd_bridge.h -- C ABI of libBridge.so, the B200-native (sm_100a) replacement for the frame-alignment hot path
 * of centreborelli/RVDD-release: dual TV-L1 optical flow + flow-based backward warp.
 *
 * Plain C, plain pointers and sizes; no torch or C++ types cross this boundary.  `dev` pointers are CUDA device
 * pointers (e.g. torch.Tensor.data_ptr()), `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * File:line references are to the reference repository (/root/reference).
 *
 * Every function except `tvl1flow` returns 0 on success and a non-zero code on failure; the message is
 * available from rvdd_last_error().  Nothing in this library aborts or exits the process (the reference C does:
 * xmalloc.c:15-17, mask.c:229-232).  There is no CPU fallback: without a usable CUDA device every call fails.
 */
#ifndef RVDD_BRIDGE_H
#define RVDD_BRIDGE_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RVDD_API __attribute__((visibility("default")))
#else
#define RVDD_API
#endif

#define RVDD_ABI_VERSION 1
#define RVDD_TRACE_SCALES 16 /* second dimension of the iteration trace */

/* ---------------------------------------------------------------------------------------------------------
 * Drop-in symbol.  Replaces `void tvl1flow(float*, float*, float*, int, int)` of libBridge.cpp:44, which
 * library.py:145-148 binds with argtypes [c_void_p, c_void_p, c_void_p, c_int, c_int], restype None and calls
 * with HOST float buffers (library.py:172-173).
 *   I0: target/reference image, nx*ny floats, row-major;  I1: image to be warped onto I0;
 *   u : caller-allocated 2*nx*ny floats; on return plane 0 = x-displacement, plane 1 = y-displacement, with
 *       I1(x + u) ~ I0(x).  Parameters are the hard-wired defaults of libBridge.cpp:27-36.
 * Buffers are neither retained nor modified (except u).  On failure u is left untouched and the error is
 * recorded (rvdd_last_error); the process is not killed.  The call runs on the CUDA device current in the calling
 * thread (one lazily created context per device) and is serialised by a mutex. */
RVDD_API void tvl1flow(float *I0, float *I1, float *u, int nx, int ny);

/* TV-L1 parameters (libBridge.cpp:27-36, tvl1flow_lib.c:343-359).  MAX_ITERATIONS=300, the presmoothing sigma
 * 0.8 and GRAD_IS_ZERO are compile-time constants in the reference (tvl1flow_lib.c:22-24) and here. */
typedef struct rvdd_tvl1_params {
    float tau;     /* 0.25 */
    float lambda;  /* 0.15 */
    float theta;   /* 0.3  */
    int nscales;   /* 100: clamped to 1 + log(hypot(nx,ny)/16)/log(1/zfactor) as libBridge.cpp:134-136 */
    int fscale;    /* 0 */
    float zfactor; /* 0.5 */
    int nwarps;    /* 5 */
    float epsilon; /* 0.01 */
} rvdd_tvl1_params;

RVDD_API void rvdd_default_params(rvdd_tvl1_params *p);

/* Number of pyramid scales the bridge would use for an nx*ny image (libBridge.cpp:131-138) and their sizes
 * (zoom.c:22-34).  nxs/nys must have room for RVDD_TRACE_SCALES ints.  Returns the number of scales. */
RVDD_API int rvdd_pyramid(int nx, int ny, const rvdd_tvl1_params *p, int *nxs, int *nys);

/* ---------------------------------------------------------------------------------------------------------
 * Context: owns the device workspace (pyramids, solver scratch, barrier words) so that no call allocates in
 * steady state -- the arena that replaces the reference's per-call xmalloc/free (tvl1flow_lib.c:110-129,
 * :364-401).  A context is bound to the CUDA device current at creation and must not be used from two threads
 * at once.  Calls on different streams are safe: the workspace is shared, so a call queued on another stream than
 * the previous one first waits (on the device, cudaStreamWaitEvent) for that one's kernels.  n_groups: how many frame pairs the persistent solver works on concurrently (0 = choose from the
 * batch size). */
typedef struct rvdd_ctx rvdd_ctx;
RVDD_API int rvdd_create(rvdd_ctx **out);
RVDD_API int rvdd_destroy(rvdd_ctx *ctx);
RVDD_API int rvdd_set_groups(rvdd_ctx *ctx, int n_groups);
/* Which instantiation of the persistent solver a launch uses.  One iterates the primal-dual loop one pass per iteration; the
 * other runs TWO iterations per pass on levels of at least `min_px` pixels (half the HBM traffic, a speculative exact stop
 * with a one-iteration replay) and is faster when the inner loops run many iterations (noisy frames) and slower when they stop
 * after one or two (clean frames).  Both return the same bits.  mode 0 = auto (decided per launch from the previous launch's
 * iteration counts on this context; the default), 1 = never fuse, 2 = always.  min_px < 0 keeps the current threshold
 * (default 600000).  Environment RVDD_FUSE=auto|0|1 sets the initial mode. */
RVDD_API int rvdd_set_fuse(rvdd_ctx *ctx, int mode, int min_px);
/* 1 if the last rvdd_tvl1_flow_dev launch on this context used the two-iterations-per-pass instantiation, 0 if not. */
RVDD_API int rvdd_last_solver_fused(rvdd_ctx *ctx);
/* Watchdog of the persistent solver: a group barrier that waits longer than `ticks` SM clock cycles (default 4e9, about
 * 2 s; also settable with the environment variable RVDD_WATCHDOG_TICKS at context creation) makes the launch unwind.
 * Raise it under time-slicing / MPS / a debugger.  When it fires, every flow of that call is overwritten with NaN on
 * the device (visible without a synchronisation) and rvdd_solver_status / the host entry points report the error. */
RVDD_API int rvdd_set_watchdog(rvdd_ctx *ctx, long long ticks);
RVDD_API const char *rvdd_last_error(void);
RVDD_API int rvdd_abi_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Gray conversion of `nimg` packed HWC images on the device, as CPPbridge.TVL1_flow does on the host
 * (library.py:162-170): c==4 mean of the channels, c==3 rgb2gray weights, c==1 copy.
 *   img_dev : [nimg][h][w][c] float32,  gray_dev : [nimg][h][w] float32 */
RVDD_API int rvdd_gray_dev(const float *img_dev, float *gray_dev, int nimg, int h, int w, int c, void *stream);

/* Batched TV-L1 on the device (Dual_TVL1_optic_flow_multiscale, tvl1flow_lib.c:343-472, for every pair).
 *   gray_dev : [nframes][ny][nx] gray frames;  pair k uses I0 = gray[tgt[k]] (frame t), I1 = gray[src[k]]
 *              (frame t-1 or t+1), i.e. compute_flow_and_warp(img1=src, img2=tgt) of flow_utils.py:138-156;
 *   src, tgt : HOST int arrays of length npairs;
 *   flow_dev : [npairs][2][ny][nx], plane 0 = u (x), plane 1 = v (y) -- the layout libBridge.cpp:150 writes;
 *   iters_dev: optional [npairs][RVDD_TRACE_SCALES][nwarps] int32 inner-iteration counts per (scale, warp)
 *              (what the reference prints with verbose=1, tvl1flow_lib.c:246-249); may be NULL;
 *   params   : NULL = defaults.
 * Asynchronous with respect to the host: work is queued on `stream`. */
RVDD_API int rvdd_tvl1_flow_dev(rvdd_ctx *ctx, const float *gray_dev, int nframes, int nx, int ny, const int *src,
                       const int *tgt, int npairs, const rvdd_tvl1_params *params, float *flow_dev, int *iters_dev,
                       void *stream);

/* Watchdog status of the last solver launch on this context: 0 ok, non-zero = a group barrier timed out and
 * the results are invalid.  Synchronises `stream`. */
RVDD_API int rvdd_solver_status(rvdd_ctx *ctx, void *stream);

/* Timing of the dominant kernel for the roofline report: with profiling enabled every rvdd_tvl1_flow_dev call
 * brackets its persistent-solver launch with a pair of CUDA events on the call's stream.  rvdd_profile_read waits
 * for them, writes the elapsed milliseconds of up to `cap` launches (oldest first), clears the list and returns
 * how many it wrote (negative on error). */
RVDD_API int rvdd_profile(rvdd_ctx *ctx, int enable);
RVDD_API int rvdd_profile_read(rvdd_ctx *ctx, float *solver_ms, int cap);
/* With profiling enabled the solver also stamps %globaltimer when a pair enters each pyramid level; this returns the
 * mean time (ms) a pair of the last launch spent at level s in ms[s] (0 = finest).  Returns the number of levels. */
RVDD_API int rvdd_profile_scales(rvdd_ctx *ctx, float *ms, int cap);
/* Same launch split by phase: ms[2*s] = warp-constants phases of level s (tvl1flow_lib.c:143-159), ms[2*s+1] = its
 * iteration loops (:161-244).  Returns the number of levels, negative on error. */
RVDD_API int rvdd_profile_phases(rvdd_ctx *ctx, float *ms, int cap);

/* Test hook: run `blocks` x 256 threads x `iters` pseudo-random trials of the kernels' straight-line exact
 * division / hypot fast paths against IEEE division and the double-precision square root on the device.
 * counters_host[6] = {hypot trials, hypot fallbacks, hypot MISMATCHES, div trials, div rejected, div MISMATCHES};
 * the two mismatch counts must be zero. */
RVDD_API int rvdd_selftest_fastmath(unsigned long long seed, int blocks, int iters, unsigned long long *counters_host);

/* Test hook: copy level `level` of the normalised + presmoothed pyramid (which: 0 = I0, 1 = I1) that the last
 * rvdd_tvl1_flow_dev call built for pair `pair` (tvl1flow_lib.c:380-401) into dst_dev (nx[level]*ny[level]
 * floats), so each pyramid stage can be checked against image_normalization / gaussian / zoom_out. */
RVDD_API int rvdd_debug_level_dev(rvdd_ctx *ctx, int pair, int which, int level, float *dst_dev, void *stream);

/* Backward warp, the CUDA side of util/flow_utils.py:70-102 (`warp(x, flow, interp)`):
 *   out[b,c,y,x] = grid_sample(x, base_grid + flow, padding_mode="border", mode=interp, align_corners=True)
 *   mask[b,0,y,x] = 1 if the normalised sampling position lies in [-1,1]^2 else 0 (may be NULL).
 *   x_dev   : B*C*H*W floats addressed with element strides (xs_b, xs_c, xs_h, xs_w)  (NCHW or HWC views);
 *   flow_dev: [B][2][fh][fw], ch 0 = x-displacement, ch 1 = y-displacement.  (fh, fw) == (H, W), or
 *             (H/2, W/2) to fuse upsample_factor_2 (flow_utils.py:159-174, bilinear x2, align_corners=True);
 *   flow_mul: multiplies the (upsampled) flow (multiply_by; 1 for a plain warp, 2 at recurrent_model.py:129);
 *   interp  : 0 = bilinear, 1 = bicubic.  out must not alias x. */
RVDD_API int rvdd_warp_dev(const float *x_dev, const float *flow_dev, float *out_dev, float *mask_dev, int B, int C, int H,
                  int W, long long xs_b, long long xs_c, long long xs_h, long long xs_w, long long os_b,
                  long long os_c, long long os_h, long long os_w, int fh, int fw, float flow_mul, int interp,
                  void *stream);

/* upsample_factor_2 (flow_utils.py:159-174): [planes][h][w] -> [planes][2h][2w], bilinear, align_corners=True,
 * times `mul`. */
RVDD_API int rvdd_upsample2_dev(const float *in_dev, float *out_dev, long long planes, int h, int w, float mul, void *stream);

/* Hamilton-Adams demosaicking of packed Bayer raw, the step right before the warp in the inference loop
 * (models/recurrent_model.py:126 -> util/Hamilton_Adam_demo.py:249-289, `HamiltonAdam(pattern).forward`), one kernel:
 *   x_dev : [B][4][H][W] packed raw, plane k = CFA sample at cell position (k / 2, k % 2) (pack_in_one, :226-234);
 *   y_dev : [B][3][2H][2W] RGB;
 *   pattern: "grbg", "rggb", "gbrg" or "bggr" (colour of cell positions (0,0), (0,1), (1,0), (1,1); the model uses
 *            "gbrg", recurrent_model.py:98).  A tensor [B, 4k, H, W] of k stacked frames is B*k images here. */
RVDD_API int rvdd_demosaic_ha_dev(const float *x_dev, float *y_dev, int B, int H, int W, const char *pattern, void *stream);

/* remosaick (Hamilton_Adam_demo.py:237-246) fused with the value mapping (v + add) * mul (singleiT, library.py:67:
 * add = 1, mul = 0.5) and the mean over the 4 packed channels (library.py:165-167): the gray image the online-flow
 * path (validate.py:29-33, --val_flow_from_denoised) hands to TV-L1, without leaving the GPU.
 *   rgb_dev : [B][3][2H][2W];  gray_dev : [B][H][W]. */
RVDD_API int rvdd_remosaick_gray_dev(const float *rgb_dev, float *gray_dev, int B, int H, int W, const char *pattern,
                                     float add, float mul, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * End-to-end with HOST buffers: exactly what data/base_dataset.py:159-180 does per (source, target) pair
 * (compute_flow_and_warp + the (h, w, 2) flow the .tif files hold), for a whole batch:
 *   frames_host : [nframes][h][w][c] float32 packed frames (c = 4 raw, 3 RGB or 1);
 *   flow_host   : [npairs][h][w][2] float32, [...,0] = u, [...,1] = v  (library.py:175 / base_dataset.py:180);
 *   warped_host : optional [npairs][h][w][c], bicubic warp of the SOURCE frame by that flow (single_warp,
 *                 flow_utils.py:105-122); NULL to skip;
 *   iters_host  : optional [npairs][RVDD_TRACE_SCALES][nwarps] int32.
 * Copies host->device, computes and copies back inside the call; returns when the results are in host memory. */
RVDD_API int rvdd_flow_and_warp_host(rvdd_ctx *ctx, const float *frames_host, int nframes, int h, int w, int c,
                            const int *src, const int *tgt, int npairs, const rvdd_tvl1_params *params,
                            float *flow_host, float *warped_host, int *iters_host);

/* The same work as rvdd_flow_and_warp_host, split into submit + wait over two staging slots (0, 1) so that a stream
 * of batches (e.g. one video after another in the offline precompute, base_dataset.py:143-189) keeps three CUDA
 * streams busy: batch i+1 uploads while batch i computes and batch i-1 downloads.  Host buffers must be pinned for
 * the copies to overlap and must stay untouched until the slot has been waited for.  A slot has to be waited for
 * before it is submitted again. */
RVDD_API int rvdd_flow_and_warp_host_submit(rvdd_ctx *ctx, int slot, const float *frames_host, int nframes, int h, int w,
                                            int c, const int *src, const int *tgt, int npairs,
                                            const rvdd_tvl1_params *params, float *flow_host, float *warped_host,
                                            int *iters_host);
RVDD_API int rvdd_flow_and_warp_host_wait(rvdd_ctx *ctx, int slot);

/* The same submission with an explicit policy for the warped frames, following the reference's dataset constructor
 * (data/base_dataset.py:178-189): `compute_flow_and_warp` always warps the source frame, but the result is only kept
 * (written to <wFolder>) when gen_warp is set.
 *   RVDD_WARP_SKIP     no warp at all (warped_host ignored);
 *   RVDD_WARP_DOWNLOAD warp on the device and copy it to warped_host (what rvdd_flow_and_warp_host_submit does when
 *                      warped_host is not NULL);
 *   RVDD_WARP_DISCARD  warp on the device, leave it there (gen_warp = False: the work of the reference's call, without
 *                      moving frames nobody reads). */
#define RVDD_WARP_SKIP 0
#define RVDD_WARP_DOWNLOAD 1
#define RVDD_WARP_DISCARD 2
RVDD_API int rvdd_flow_and_warp_host_submit_ex(rvdd_ctx *ctx, int slot, const float *frames_host, int nframes, int h, int w,
                                               int c, const int *src, const int *tgt, int npairs,
                                               const rvdd_tvl1_params *params, float *flow_host, float *warped_host,
                                               int *iters_host, int warp_mode);

#ifdef __cplusplus
}
#endif
#endif /* RVDD_BRIDGE_H */
