// internal.h -- host-side declarations shared by the .cu files of libBridge.so (not part of the C ABI).
#pragma once
#include <cuda.h>            // CUtensorMap (type only; the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

#include "exact_math.h"

namespace rvdd {

// Geometry of one pyramid, computed on the host exactly as zoom_size / libBridge.cpp do.
struct Pyramid {
    int S;                              // number of scales actually used
    int nx[RVDD_MAX_SCALES], ny[RVDD_MAX_SCALES];
    long long off[RVDD_MAX_SCALES];     // element offset of level s inside a per-image pyramid buffer
    long long total;                    // elements per image pyramid (padded so every level is 16 B aligned)
};

struct GaussTaps {
    int size;                           // taps B[0..size-1] (mask.c:225)
    double B[RVDD_MAX_TAPS];
};

// Arguments of the persistent solver kernel (solver.cu).
struct SolverArgs {
    // TMA descriptors of the whole solver scratch viewed as a 2-D tensor [ngroups * RVDD_NPLANES planes][plane
    // floats]; the boxes are one staged row segment (136 floats) of 4 adjacent planes (dual variable), 3 (per-warp
    // constants) or 2 (flow).
    alignas(64) CUtensorMap tm4;
    alignas(64) CUtensorMap tm3;
    alignas(64) CUtensorMap tm2;
    int npairs, S, fscale, nwarps;
    int nx[RVDD_MAX_SCALES], ny[RVDD_MAX_SCALES];
    long long off[RVDD_MAX_SCALES];
    float zfx[RVDD_MAX_SCALES], zfy[RVDD_MAX_SCALES];   // zoom_in factors towards level s (from s+1), zoom.c:95-96
    float l_t, theta, taut, eps2, zoom_mul, g0f;   // g0f: GRAD_IS_ZERO as a float threshold (exact_math.h)
    const float *pyr0, *pyr1;           // [npairs][pyr_stride]
    long long pyr_stride;
    float *flow_out;                    // [npairs][2][nx0*ny0]
    float *scratch;                     // [ngroups][scratch_stride]
    long long scratch_stride, plane;    // plane = padded nx0*ny0
    int *iters_out;                     // [npairs][RVDD_MAX_SCALES][nwarps] or null
    float *err_out;                     // same shape, error at loop exit, or null
    unsigned *bar;                      // [ngroups * 32] (one counter per 128 B)
    double *partials;                   // [ngroups][2 slots][2 sums][ctas_per_group]
    unsigned long long *scale_ns;       // optional [npairs][RVDD_MAX_SCALES + 1] globaltimer stamps (profiling)
    int *status;                        // [0]: watchdog flag, [1]: finest-level inner iterations of the launch (summed over pairs)
    int ngroups, ctas_per_group;
    long long spin_limit;               // watchdog, in clock64 ticks
    int fuse_min_px;                    // levels with at least this many pixels (and nx % 4 == 0) run two iterations per pass
    int fuse_first;                     // ... after this many single iterations of every inner loop
    int fuse_min_rows;                  // ... and only if every warp of the group gets at least this many segment-rows
    int fuse_hint;                      // inner iterations per warp on the finest level in the previous launch (0 = unknown)
};

// bridge.cu: 2-D float32 TMA descriptor (dims d0 innermost / d1, row pitch in bytes, box b0 x b1, zero fill out of range)
cudaError_t encode_map_2d(CUtensorMap *tm, const float *base, unsigned long long d0, unsigned long long d1,
                          unsigned long long pitch_bytes, unsigned b0, unsigned b1);

// prep.cu
cudaError_t launch_setup(int *minmax_slots, int npairs, unsigned *bar, int nbar, int *status, cudaStream_t st);
cudaError_t launch_minmax(const float *const *I0, const float *const *I1, int n, int npairs, int *slots, cudaStream_t st);
// dst[z] = gaussian(normalise?(src[z])); z < nimg.  Sources come from the device pointer table `srcs` or, when
// it is null, from src_base + z * src_stride; dst images are dst_base + z * dst_stride.  When slots != null image z is normalised with the min/max of pair z % npairs.
cudaError_t launch_gauss(const float *const *srcs, const float *src_base, long long src_stride, float *dst_base,
                         long long dst_stride, int nx, int ny, int nimg, const GaussTaps &taps, const int *slots,
                         int npairs, cudaStream_t st);
cudaError_t launch_gauss_decimate(const float *src_base, long long src_stride, float *dst_base, long long dst_stride, int nx,
                                  int ny, int nxx, int nyy, int nimg, const GaussTaps &taps, cudaStream_t st);
cudaError_t launch_resample(const float *src_base, long long src_stride, int nx, int ny, float *dst_base,
                            long long dst_stride, int nxx, int nyy, float fx, float fy, int nimg, cudaStream_t st);
cudaError_t launch_gray(const float *img, float *gray, long long npix_total, int c, cudaStream_t st);

// solver.cu
cudaError_t solver_max_ctas(int *ctas_per_sm, int *sms);
cudaError_t launch_solver(const SolverArgs &args, bool fused_kernel, cudaStream_t st);
int solver_threads();

// warp.cu
struct WarpArgs {
    const float *x;
    const float *flow;
    float *out;
    float *mask;                        // nullable, [B][1][H][W]
    int B, C, H, W;
    long long xs_b, xs_c, xs_h, xs_w;   // element strides of x
    long long os_b, os_c, os_h, os_w;   // element strides of out
    int fh, fw;                         // flow grid; (H, W) or (H/2, W/2) with fused upsample_factor_2
    float flow_mul;
    int interp;                         // 0 bilinear, 1 bicubic
};
cudaError_t launch_warp(const WarpArgs &a, cudaStream_t st);
cudaError_t launch_upsample2(const float *in, float *out, long long planes, int h, int w, float mul, cudaStream_t st);

// demosaic.cu: (ry, rx) / (by, bx) = position of the red / blue sample inside the 2x2 Bayer cell
cudaError_t launch_demosaic_ha(const float *x, float *y, int B, int H, int W, int ry, int rx, int by, int bx, cudaStream_t st);
cudaError_t launch_remosaick_gray(const float *rgb, float *gray, int B, int H, int W, int ry, int rx, int by, int bx,
                                  float add, float mul, cudaStream_t st);

}  // namespace rvdd
