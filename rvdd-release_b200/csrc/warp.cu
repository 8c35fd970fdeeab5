// warp.cu -- flow-based backward warp of multi-channel tensors, the CUDA side of util/flow_utils.py:
//   warp(x, flow, interp)            (flow_utils.py:70-102)  -> grid_sample(bicubic|bilinear, border, align_corners)
//   upsample_factor_2(t, multiply)   (flow_utils.py:159-174) -> bilinear x2, align_corners, optionally fused
// One thread owns one output pixel: the sampling position, the 16 tap offsets and the 8 cubic weights are
// computed once from the flow and reused for every channel, so a C-channel warp reads the flow once instead
// of once per channel, rebuilds no meshgrid, and writes the validity mask in the same pass (no D2H sync).
//
// Semantics follow ATen's grid_sampler_2d (float32): normalise 2*v/(W-1)-1 as flow_utils.py:93-94 does,
// un-normalise ((g+1)/2)*(W-1), bicubic = Keys A=-0.75 with the centre left unclipped and every tap clamped to
// the image, bilinear = centre clipped to the border first.  Results agree with torch to ~1e-6 relative (the
// float expression order of the two ATen back ends differs by that much already); the gate is 1e-4.
#include <stdlib.h>

#include "internal.h"

namespace rvdd {

__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4])
{
    const float A = -0.75f;
    float x = t + 1.0f;
    c[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
    x = t;
    c[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
    x = 1.0f - t;
    c[2] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
    x = 2.0f - t;
    c[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}

// bilinear x2 upsampling with align_corners=True (ATen upsample_bilinear2d): src = dst * (in-1)/(out-1)
__device__ __forceinline__ float up2_sample(const float *__restrict__ p, int h, int w, int y, int x, float sy, float sx)
{
    const float fy = sy * (float)y, fx = sx * (float)x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    return hy * (hx * p[(long long)y0 * w + x0] + lx * p[(long long)y0 * w + x1]) +
           ly * (hx * p[(long long)y1 * w + x0] + lx * p[(long long)y1 * w + x1]);
}

template <int INTERP>
__global__ void __launch_bounds__(256) warp_kernel(const WarpArgs a)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= a.W || y >= a.H) return;

    // flow at this pixel (optionally bilinearly upsampled from the half-resolution grid and scaled)
    float fu, fv;
    const float *fl = a.flow + (long long)b * 2 * a.fh * a.fw;
    if (a.fh == a.H && a.fw == a.W) {
        fu = fl[(long long)y * a.W + x];
        fv = fl[(long long)a.H * a.W + (long long)y * a.W + x];
    } else {
        const float sy = a.H > 1 ? (float)(a.fh - 1) / (float)(a.H - 1) : 0.f;
        const float sx = a.W > 1 ? (float)(a.fw - 1) / (float)(a.W - 1) : 0.f;
        fu = up2_sample(fl, a.fh, a.fw, y, x, sy, sx);
        fv = up2_sample(fl + (long long)a.fh * a.fw, a.fh, a.fw, y, x, sy, sx);
    }
    fu *= a.flow_mul;
    fv *= a.flow_mul;

    // flow_utils.py:90-96: vgrid, normalisation, validity mask
    const float gxn = 2.0f * ((float)x + fu) / (float)(a.W - 1) - 1.0f;
    const float gyn = 2.0f * ((float)y + fv) / (float)(a.H - 1) - 1.0f;
    if (a.mask)
        a.mask[((long long)b * a.H + y) * a.W + x] = (gxn >= -1.f && gxn <= 1.f && gyn >= -1.f && gyn <= 1.f) ? 1.f : 0.f;
    // grid_sampler_unnormalize, align_corners=True
    float ix = ((gxn + 1.f) / 2.f) * (float)(a.W - 1);
    float iy = ((gyn + 1.f) / 2.f) * (float)(a.H - 1);

    const float *xb = a.x + (long long)b * a.xs_b;
    float *ob = a.out + (long long)b * a.os_b + (long long)y * a.os_h + (long long)x * a.os_w;

    if (INTERP == 1) {
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        float cx[4], cy[4];
        cubic_coeffs(ix - fx0, cx);
        cubic_coeffs(iy - fy0, cy);
        long long ox[4], oy[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // get_value_bounded: clip the float tap coordinate to [0, size-1], then truncate
            const float tx = fminf((float)(a.W - 1), fmaxf(fx0 - 1.f + (float)k, 0.f));
            const float ty = fminf((float)(a.H - 1), fmaxf(fy0 - 1.f + (float)k, 0.f));
            ox[k] = (long long)(int)tx * a.xs_w;
            oy[k] = (long long)(int)ty * a.xs_h;
        }
        for (int c = 0; c < a.C; c++) {
            const float *p = xb + (long long)c * a.xs_c;
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const float *q = p + oy[r];
                const float row = __ldg(q + ox[0]) * cx[0] + __ldg(q + ox[1]) * cx[1] + __ldg(q + ox[2]) * cx[2] +
                                  __ldg(q + ox[3]) * cx[3];
                acc += row * cy[r];
            }
            ob[(long long)c * a.os_c] = acc;
        }
    } else {
        // bilinear, padding border: clip the centre, then the usual 4 taps (out-of-range taps have weight 0)
        ix = fminf((float)(a.W - 1), fmaxf(ix, 0.f));
        iy = fminf((float)(a.H - 1), fmaxf(iy, 0.f));
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        const int x1 = min(x0 + 1, a.W - 1), y1 = min(y0 + 1, a.H - 1);
        const float lx = ix - fx0, ly = iy - fy0;
        const float wnw = (1.f - lx) * (1.f - ly), wne = lx * (1.f - ly), wsw = (1.f - lx) * ly, wse = lx * ly;
        for (int c = 0; c < a.C; c++) {
            const float *p = xb + (long long)c * a.xs_c;
            ob[(long long)c * a.os_c] = __ldg(p + (long long)y0 * a.xs_h + (long long)x0 * a.xs_w) * wnw +
                                        __ldg(p + (long long)y0 * a.xs_h + (long long)x1 * a.xs_w) * wne +
                                        __ldg(p + (long long)y1 * a.xs_h + (long long)x0 * a.xs_w) * wsw +
                                        __ldg(p + (long long)y1 * a.xs_h + (long long)x1 * a.xs_w) * wse;
        }
    }
}

// ------------------------------------------------------------------------------------------------ tiled variant
//
// warp_kernel issues 16 scattered global loads per pixel and channel; a warp's 32 pixels touch the same ~5 cache lines
// 16 times over, and the kernel ends up bound by L1 wavefronts at ~0.8 TB/s of useful traffic.  For plane-contiguous
// inputs (NCHW, xs_w == 1) the tiled kernel below stages, per 32x8 output tile and per group of 4 channels, the
// bounding box of all taps of the tile in shared memory and gathers from there.  The staged box is CHANNEL-INTERLEAVED
// (one float4 = the 4 channels of a pixel): a tap is ONE 128-bit shared-memory load for 4 channels instead of four
// 32-bit ones, always 16-byte aligned whatever the flow, and conflict-free (the lanes of a quarter-warp read
// consecutive 16-byte words; the row pitch is a multiple of 128 B, so lanes whose taps fall on different rows still
// hit different banks).  The sampling position, the clamped tap offsets and the cubic weights are computed once per
// pixel and reused for every channel; the next channel group streams in with cp.async while the current one is
// consumed.  If the flow varies so much inside a tile that the box does not fit (more than 13 px horizontally or
// vertically), the tile falls back to direct global gathers.  Same arithmetic as warp_kernel.

#define WT_X 32
#define WT_Y 8
#define WT_BW 48            // staged box: at most 48 x 24 pixels
#define WT_BH 24
#define WT_CC 4             // channels staged per pass = one float4 (two passes in flight: cp.async double buffering)

__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int INTERP>
__global__ void __launch_bounds__(256) warp_tile_kernel(const WarpArgs a)
{
    __shared__ float4 s_box[2][WT_BH * WT_BW];
    __shared__ int s_lim[4];
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int x = blockIdx.x * WT_X + lane, y = blockIdx.y * WT_Y + wy, b = blockIdx.z;
    const bool inside = (x < a.W && y < a.H);
    if (threadIdx.x == 0) { s_lim[0] = 0x7fffffff; s_lim[1] = -0x7fffffff; s_lim[2] = 0x7fffffff; s_lim[3] = -0x7fffffff; }

    // sampling position of this pixel (identical to warp_kernel)
    float ix = 0.f, iy = 0.f;
    if (inside) {
        float fu, fv;
        const float *fl = a.flow + (long long)b * 2 * a.fh * a.fw;
        if (a.fh == a.H && a.fw == a.W) {
            fu = fl[(long long)y * a.W + x];
            fv = fl[(long long)a.H * a.W + (long long)y * a.W + x];
        } else {
            const float sy = a.H > 1 ? (float)(a.fh - 1) / (float)(a.H - 1) : 0.f;
            const float sx = a.W > 1 ? (float)(a.fw - 1) / (float)(a.W - 1) : 0.f;
            fu = up2_sample(fl, a.fh, a.fw, y, x, sy, sx);
            fv = up2_sample(fl + (long long)a.fh * a.fw, a.fh, a.fw, y, x, sy, sx);
        }
        fu *= a.flow_mul;
        fv *= a.flow_mul;
        const float gxn = 2.0f * ((float)x + fu) / (float)(a.W - 1) - 1.0f;
        const float gyn = 2.0f * ((float)y + fv) / (float)(a.H - 1) - 1.0f;
        if (a.mask)
            a.mask[((long long)b * a.H + y) * a.W + x] = (gxn >= -1.f && gxn <= 1.f && gyn >= -1.f && gyn <= 1.f) ? 1.f : 0.f;
        ix = ((gxn + 1.f) / 2.f) * (float)(a.W - 1);
        iy = ((gyn + 1.f) / 2.f) * (float)(a.H - 1);
    }
    constexpr int NT = (INTERP == 1) ? 4 : 2;          // taps per axis
    float cx[NT], cy[NT];
    int tx[NT], ty[NT];                                // clamped absolute tap coordinates
    if (INTERP == 1) {
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        float wx[4], wyy[4];
        cubic_coeffs(ix - fx0, wx);
        cubic_coeffs(iy - fy0, wyy);
#pragma unroll
        for (int k = 0; k < NT; k++) {
            cx[k] = wx[k];
            cy[k] = wyy[k];
            tx[k] = (int)fminf((float)(a.W - 1), fmaxf(fx0 - 1.f + (float)k, 0.f));
            ty[k] = (int)fminf((float)(a.H - 1), fmaxf(fy0 - 1.f + (float)k, 0.f));
        }
    } else {
        ix = fminf((float)(a.W - 1), fmaxf(ix, 0.f));
        iy = fminf((float)(a.H - 1), fmaxf(iy, 0.f));
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        tx[0] = (int)fx0; ty[0] = (int)fy0;
        tx[NT - 1] = min(tx[0] + 1, a.W - 1); ty[NT - 1] = min(ty[0] + 1, a.H - 1);
        cx[NT - 1] = ix - fx0; cy[NT - 1] = iy - fy0;
        cx[0] = 1.f - cx[NT - 1]; cy[0] = 1.f - cy[NT - 1];
    }
    __syncthreads();
    // bounding box of the tile's taps (taps are monotone in k, so first / last suffice)
    {
        int lox = inside ? tx[0] : 0x7fffffff, hix = inside ? tx[NT - 1] : -0x7fffffff;
        int loy = inside ? ty[0] : 0x7fffffff, hiy = inside ? ty[NT - 1] : -0x7fffffff;
        lox = __reduce_min_sync(0xffffffffu, lox); hix = __reduce_max_sync(0xffffffffu, hix);
        loy = __reduce_min_sync(0xffffffffu, loy); hiy = __reduce_max_sync(0xffffffffu, hiy);
        if (lane == 0) { atomicMin(&s_lim[0], lox); atomicMax(&s_lim[1], hix); atomicMin(&s_lim[2], loy); atomicMax(&s_lim[3], hiy); }
    }
    __syncthreads();
    const int bx0 = s_lim[0], by0 = s_lim[2], bw = s_lim[1] - s_lim[0] + 1, bh = s_lim[3] - s_lim[2] + 1;
    const bool staged = (bw <= WT_BW && bh <= WT_BH);      // block-uniform
    const float *xb = a.x + (long long)b * a.xs_b;
    float *ob = a.out + (long long)b * a.os_b + (long long)y * a.os_h + (long long)x * a.os_w;

    if (!staged) {
        if (!inside) return;
        for (int c = 0; c < a.C; c++) {
            const float *p = xb + (long long)c * a.xs_c;
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < NT; r++) {
                const float *q = p + (long long)ty[r] * a.xs_h;
                float row = 0.f;
#pragma unroll
                for (int k = 0; k < NT; k++) row = row + __ldg(q + tx[k]) * cx[k];
                acc += row * cy[r];
            }
            ob[(long long)c * a.os_c] = acc;
        }
        return;
    }
    int off[NT][NT];
#pragma unroll
    for (int r = 0; r < NT; r++)
#pragma unroll
        for (int k = 0; k < NT; k++) off[r][k] = (ty[r] - by0) * WT_BW + (tx[k] - bx0);

    // stage the box of the 4 channels starting at c0 into buffer `buf`.  A lane owns (column lane / 4, channel lane % 4):
    // a warp reads 8 consecutive pixels (one full 32-byte sector) of each of the 4 channel planes and writes 32
    // consecutive floats of the interleaved box -- conflict-free stores, fully used sectors.
    const int s_cc = lane & 3, s_col = lane >> 2;
    auto stage = [&](int c0, int buf) {
        float *dst = reinterpret_cast<float *>(s_box[buf]);
        if (c0 + s_cc < a.C) {
            const float *p = xb + (long long)(c0 + s_cc) * a.xs_c + (long long)by0 * a.xs_h + bx0;
            for (int r = wy; r < bh; r += WT_Y) {
                const float *q = p + (long long)r * a.xs_h;
#pragma unroll
                for (int col0 = 0; col0 < WT_BW; col0 += 8) {
                    const int col = col0 + s_col;
                    if (col < bw) cp_async4(dst + (r * WT_BW + col) * 4 + s_cc, q + col);
                }
            }
        }
        cp_async_commit();
    };
    stage(0, 0);
    int buf = 0;
    for (int c0 = 0; c0 < a.C; c0 += WT_CC, buf ^= 1) {
        const int nc = min(WT_CC, a.C - c0);
        const bool more = (c0 + WT_CC < a.C);
        if (more) stage(c0 + WT_CC, buf ^ 1);              // next pass streams in while this one is consumed
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        if (inside) {
            const float4 *sb = s_box[buf];
            float4 acc;
            if (INTERP == 1) {
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < NT; r++) {
                    const float4 v0 = sb[off[r][0]], v1 = sb[off[r][1]], v2 = sb[off[r][NT - 2]], v3 = sb[off[r][NT - 1]];
                    acc.x += (v0.x * cx[0] + v1.x * cx[1] + v2.x * cx[NT - 2] + v3.x * cx[NT - 1]) * cy[r];
                    acc.y += (v0.y * cx[0] + v1.y * cx[1] + v2.y * cx[NT - 2] + v3.y * cx[NT - 1]) * cy[r];
                    acc.z += (v0.z * cx[0] + v1.z * cx[1] + v2.z * cx[NT - 2] + v3.z * cx[NT - 1]) * cy[r];
                    acc.w += (v0.w * cx[0] + v1.w * cx[1] + v2.w * cx[NT - 2] + v3.w * cx[NT - 1]) * cy[r];
                }
            } else {
                const float4 v00 = sb[off[0][0]], v01 = sb[off[0][NT - 1]], v10 = sb[off[NT - 1][0]], v11 = sb[off[NT - 1][NT - 1]];
                const float w00 = cx[0] * cy[0], w01 = cx[NT - 1] * cy[0], w10 = cx[0] * cy[NT - 1], w11 = cx[NT - 1] * cy[NT - 1];
                acc.x = v00.x * w00 + v01.x * w01 + v10.x * w10 + v11.x * w11;
                acc.y = v00.y * w00 + v01.y * w01 + v10.y * w10 + v11.y * w11;
                acc.z = v00.z * w00 + v01.z * w01 + v10.z * w10 + v11.z * w11;
                acc.w = v00.w * w00 + v01.w * w01 + v10.w * w10 + v11.w * w11;
            }
            float *o = ob + (long long)c0 * a.os_c;
            o[0] = acc.x;
            if (nc > 1) o[a.os_c] = acc.y;
            if (nc > 2) o[2 * a.os_c] = acc.z;
            if (nc > 3) o[3 * a.os_c] = acc.w;
        }
        __syncthreads();                                   // this buffer is refilled two passes from now
    }
}

// ------------------------------------------------------------------------------------------------ HWC, 4 channels
//
// single_warp / compute_flow_and_warp hand over frames in their on-disk layout: (H, W, 4) packed raw, channel
// innermost (flow_utils.py:105-122, base_dataset.py:174-178).  There the 4 channels of a tap are ONE aligned 16-byte
// load, so the gather needs no staging at all: 16 x LDG.128 per pixel through L1 (a warp's 32 pixels read 512
// contiguous bytes per tap), one 128-bit store.  Same arithmetic as warp_kernel.
template <int INTERP>
__global__ void __launch_bounds__(256) warp_hwc4_kernel(const WarpArgs a)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= a.W || y >= a.H) return;
    float fu, fv;
    const float *fl = a.flow + (long long)b * 2 * a.fh * a.fw;
    if (a.fh == a.H && a.fw == a.W) {
        fu = fl[(long long)y * a.W + x];
        fv = fl[(long long)a.H * a.W + (long long)y * a.W + x];
    } else {
        const float sy = a.H > 1 ? (float)(a.fh - 1) / (float)(a.H - 1) : 0.f;
        const float sx = a.W > 1 ? (float)(a.fw - 1) / (float)(a.W - 1) : 0.f;
        fu = up2_sample(fl, a.fh, a.fw, y, x, sy, sx);
        fv = up2_sample(fl + (long long)a.fh * a.fw, a.fh, a.fw, y, x, sy, sx);
    }
    fu *= a.flow_mul;
    fv *= a.flow_mul;
    const float gxn = 2.0f * ((float)x + fu) / (float)(a.W - 1) - 1.0f;
    const float gyn = 2.0f * ((float)y + fv) / (float)(a.H - 1) - 1.0f;
    if (a.mask)
        a.mask[((long long)b * a.H + y) * a.W + x] = (gxn >= -1.f && gxn <= 1.f && gyn >= -1.f && gyn <= 1.f) ? 1.f : 0.f;
    float ix = ((gxn + 1.f) / 2.f) * (float)(a.W - 1);
    float iy = ((gyn + 1.f) / 2.f) * (float)(a.H - 1);
    const float4 *xb = reinterpret_cast<const float4 *>(a.x + (long long)b * a.xs_b);
    const long long rowq = a.xs_h >> 2;                 // row pitch in float4 units
    float4 acc;
    if (INTERP == 1) {
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        float cx[4], cy[4];
        cubic_coeffs(ix - fx0, cx);
        cubic_coeffs(iy - fy0, cy);
        int ox[4];
        long long oy[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            ox[k] = (int)fminf((float)(a.W - 1), fmaxf(fx0 - 1.f + (float)k, 0.f));
            oy[k] = (long long)(int)fminf((float)(a.H - 1), fmaxf(fy0 - 1.f + (float)k, 0.f)) * rowq;
        }
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const float4 *q = xb + oy[r];
            const float4 v0 = __ldg(q + ox[0]), v1 = __ldg(q + ox[1]), v2 = __ldg(q + ox[2]), v3 = __ldg(q + ox[3]);
            acc.x += (v0.x * cx[0] + v1.x * cx[1] + v2.x * cx[2] + v3.x * cx[3]) * cy[r];
            acc.y += (v0.y * cx[0] + v1.y * cx[1] + v2.y * cx[2] + v3.y * cx[3]) * cy[r];
            acc.z += (v0.z * cx[0] + v1.z * cx[1] + v2.z * cx[2] + v3.z * cx[3]) * cy[r];
            acc.w += (v0.w * cx[0] + v1.w * cx[1] + v2.w * cx[2] + v3.w * cx[3]) * cy[r];
        }
    } else {
        ix = fminf((float)(a.W - 1), fmaxf(ix, 0.f));
        iy = fminf((float)(a.H - 1), fmaxf(iy, 0.f));
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        const int x1 = min(x0 + 1, a.W - 1), y1 = min(y0 + 1, a.H - 1);
        const float lx = ix - fx0, ly = iy - fy0;
        const float wnw = (1.f - lx) * (1.f - ly), wne = lx * (1.f - ly), wsw = (1.f - lx) * ly, wse = lx * ly;
        const float4 v00 = __ldg(xb + y0 * rowq + x0), v01 = __ldg(xb + y0 * rowq + x1);
        const float4 v10 = __ldg(xb + y1 * rowq + x0), v11 = __ldg(xb + y1 * rowq + x1);
        acc.x = v00.x * wnw + v01.x * wne + v10.x * wsw + v11.x * wse;
        acc.y = v00.y * wnw + v01.y * wne + v10.y * wsw + v11.y * wse;
        acc.z = v00.z * wnw + v01.z * wne + v10.z * wsw + v11.z * wse;
        acc.w = v00.w * wnw + v01.w * wne + v10.w * wsw + v11.w * wse;
    }
    float *ob = a.out + (long long)b * a.os_b + (long long)y * a.os_h + (long long)x * a.os_w;
    if (a.os_c == 1 && ((reinterpret_cast<uintptr_t>(ob) & 15) == 0)) {
        *reinterpret_cast<float4 *>(ob) = acc;
    } else {
        ob[0] = acc.x; ob[a.os_c] = acc.y; ob[2 * a.os_c] = acc.z; ob[3 * a.os_c] = acc.w;
    }
}

// Bicubic, TWO vertically adjacent pixels per thread: the arithmetic of warp_hwc4_kernel unchanged, both sample positions
// (flow loads) first, then one pixel after the other.  Measured on 29 x (720, 1280, 4), B200 (profiles/warp_hwc_variants_r02.txt):
// 0.431 -> 0.403 ms (white-noise flow 0.816 -> 0.688 ms).  The kernel is latency-bound and sensitive to the schedule, not
// to the number of L1 look-ups -- every "smarter" variant lost:
//   * SHARING taps between the two pixels (their windows overlap in 3 of 4 rows when the flow is smooth: 20 loads for 32):
//     paired along x the lanes stride 32 bytes, 0.591 ms; paired along y, 0.394 ms on a noise-free flow but 0.493 ms as soon
//     as a few per cent of the lanes leave the shared path (0.430 ms when whole warps must agree);
//   * 3 or 4 pixels per thread: 0.441 / 0.500 ms;  taller tiles (512 or 1024 threads, 16 .. 64 rows): 0.46 .. 0.65 ms;
//   * TMA-staged tiles (below): 0.55 ms.
__device__ __forceinline__ void hwc4_sample_pos(const WarpArgs &a, int b, int x, int y, float &ix, float &iy)
{
    float fu, fv;
    const float *fl = a.flow + (long long)b * 2 * a.fh * a.fw;
    if (a.fh == a.H && a.fw == a.W) {
        fu = fl[(long long)y * a.W + x];
        fv = fl[(long long)a.H * a.W + (long long)y * a.W + x];
    } else {
        const float sy = a.H > 1 ? (float)(a.fh - 1) / (float)(a.H - 1) : 0.f;
        const float sx = a.W > 1 ? (float)(a.fw - 1) / (float)(a.W - 1) : 0.f;
        fu = up2_sample(fl, a.fh, a.fw, y, x, sy, sx);
        fv = up2_sample(fl + (long long)a.fh * a.fw, a.fh, a.fw, y, x, sy, sx);
    }
    fu *= a.flow_mul;
    fv *= a.flow_mul;
    const float gxn = 2.0f * ((float)x + fu) / (float)(a.W - 1) - 1.0f;
    const float gyn = 2.0f * ((float)y + fv) / (float)(a.H - 1) - 1.0f;
    if (a.mask)
        a.mask[((long long)b * a.H + y) * a.W + x] = (gxn >= -1.f && gxn <= 1.f && gyn >= -1.f && gyn <= 1.f) ? 1.f : 0.f;
    ix = ((gxn + 1.f) / 2.f) * (float)(a.W - 1);
    iy = ((gyn + 1.f) / 2.f) * (float)(a.H - 1);
}

__device__ __forceinline__ void hwc4_store(const WarpArgs &a, int b, int x, int y, const float4 &acc)
{
    float *ob = a.out + (long long)b * a.os_b + (long long)y * a.os_h + (long long)x * a.os_w;
    if (a.os_c == 1 && ((reinterpret_cast<uintptr_t>(ob) & 15) == 0)) {
        *reinterpret_cast<float4 *>(ob) = acc;
    } else {
        ob[0] = acc.x; ob[a.os_c] = acc.y; ob[2 * a.os_c] = acc.z; ob[3 * a.os_c] = acc.w;
    }
}

#define HWC4_ROW(ACC, V0, V1, V2, V3, CX, CYR)                                                                  \
    ACC.x += (V0.x * CX[0] + V1.x * CX[1] + V2.x * CX[2] + V3.x * CX[3]) * (CYR);                               \
    ACC.y += (V0.y * CX[0] + V1.y * CX[1] + V2.y * CX[2] + V3.y * CX[3]) * (CYR);                               \
    ACC.z += (V0.z * CX[0] + V1.z * CX[1] + V2.z * CX[2] + V3.z * CX[3]) * (CYR);                               \
    ACC.w += (V0.w * CX[0] + V1.w * CX[1] + V2.w * CX[2] + V3.w * CX[3]) * (CYR);

__device__ __forceinline__ float4 hwc4_bicubic_px(const WarpArgs &a, const float4 *xb, long long rowq, float ix, float iy)
{
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    float cx[4], cy[4];
    cubic_coeffs(ix - fx0, cx);
    cubic_coeffs(iy - fy0, cy);
    int ox[4];
    long long oy[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        ox[k] = (int)fminf((float)(a.W - 1), fmaxf(fx0 - 1.f + (float)k, 0.f));
        oy[k] = (long long)(int)fminf((float)(a.H - 1), fmaxf(fy0 - 1.f + (float)k, 0.f)) * rowq;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const float4 *q = xb + oy[r];
        const float4 v0 = __ldg(q + ox[0]), v1 = __ldg(q + ox[1]), v2 = __ldg(q + ox[2]), v3 = __ldg(q + ox[3]);
        HWC4_ROW(acc, v0, v1, v2, v3, cx, cy[r])
    }
    return acc;
}

__global__ void __launch_bounds__(256) warp_hwc4_2rows_kernel(const WarpArgs a)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = 2 * (blockIdx.y * 8 + (threadIdx.x >> 5));
    const int b = blockIdx.z;
    if (x >= a.W || y >= a.H) return;
    const bool two = y + 1 < a.H;
    float ix0, iy0, ix1 = 0.f, iy1 = 0.f;
    hwc4_sample_pos(a, b, x, y, ix0, iy0);
    if (two) hwc4_sample_pos(a, b, x, y + 1, ix1, iy1);
    const float4 *xb = reinterpret_cast<const float4 *>(a.x + (long long)b * a.xs_b);
    const long long rowq = a.xs_h >> 2;                 // row pitch in float4 units
    hwc4_store(a, b, x, y, hwc4_bicubic_px(a, xb, rowq, ix0, iy0));
    if (two) hwc4_store(a, b, x, y + 1, hwc4_bicubic_px(a, xb, rowq, ix1, iy1));
}

// ------------------------------------------------------------------------------------------------ HWC, 4 channels, TMA-staged
//
// OPT-IN (environment RVDD_WARP_TMA=1): measured SLOWER than the L1 gathers of warp_hwc4_kernel on B200 -- 29 x (720, 1280, 4):
// 0.53 ms against 0.43 ms, also with smaller boxes (40 x 14: 0.53 ms, 36 x 12: 0.57 ms; profiles/stage_kernels_r02.jsonl).
// The per-tile chain flow -> bounding box -> TMA request -> wait -> gather is serial inside a CTA and eight resident CTAs do
// not hide it, while the direct kernel keeps 16 independent 128-bit loads per thread in flight.  Kept as the measured
// alternative (and parity-tested), not as the default.
//
// Same job as warp_hwc4_kernel, but the taps do not come through L1 (16 tag look-ups per pixel, 83 % hits, the misses at
// L2 latency): the bounding box of the 32x8 tile's taps is fetched ONCE by the TMA unit -- cp.async.bulk.tensor.2d over the
// frames viewed as a [B*H rows][W*4 floats] tensor, a box of WH_BW pixels x WH_BH rows landing in shared memory in exactly
// the channel-interleaved layout the gather wants (one pixel = one float4) -- and the 16 taps of a pixel are 16 conflict-free
// 128-bit shared loads.  No thread issues a global load for the frames at all.  A tile whose flow spreads the taps over more
// than the box (or frames that are not densely packed) takes the direct-gather kernel's path.  Same arithmetic.
#ifndef WH_BW
#define WH_BW 48
#endif
#ifndef WH_BH
#define WH_BH 24
#endif

template <int INTERP>
__global__ void __launch_bounds__(256) warp_hwc4_tma_kernel(const WarpArgs a, const __grid_constant__ CUtensorMap tm)
{
    __shared__ __align__(128) float4 s_box[WH_BH * WH_BW];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_lim[4];
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
    const bool inside = (x < a.W && y < a.H);
    if (threadIdx.x == 0) {
        s_lim[0] = 0x7fffffff; s_lim[1] = -0x7fffffff; s_lim[2] = 0x7fffffff; s_lim[3] = -0x7fffffff;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float ix = 0.f, iy = 0.f;
    if (inside) {
        float fu, fv;
        const float *fl = a.flow + (long long)b * 2 * a.fh * a.fw;
        if (a.fh == a.H && a.fw == a.W) {
            fu = fl[(long long)y * a.W + x];
            fv = fl[(long long)a.H * a.W + (long long)y * a.W + x];
        } else {
            const float sy = a.H > 1 ? (float)(a.fh - 1) / (float)(a.H - 1) : 0.f;
            const float sx = a.W > 1 ? (float)(a.fw - 1) / (float)(a.W - 1) : 0.f;
            fu = up2_sample(fl, a.fh, a.fw, y, x, sy, sx);
            fv = up2_sample(fl + (long long)a.fh * a.fw, a.fh, a.fw, y, x, sy, sx);
        }
        fu *= a.flow_mul;
        fv *= a.flow_mul;
        const float gxn = 2.0f * ((float)x + fu) / (float)(a.W - 1) - 1.0f;
        const float gyn = 2.0f * ((float)y + fv) / (float)(a.H - 1) - 1.0f;
        if (a.mask)
            a.mask[((long long)b * a.H + y) * a.W + x] = (gxn >= -1.f && gxn <= 1.f && gyn >= -1.f && gyn <= 1.f) ? 1.f : 0.f;
        ix = ((gxn + 1.f) / 2.f) * (float)(a.W - 1);
        iy = ((gyn + 1.f) / 2.f) * (float)(a.H - 1);
    }
    constexpr int NT = (INTERP == 1) ? 4 : 2;
    float cx[NT], cy[NT];
    int tx[NT], ty[NT];
    if (INTERP == 1) {
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        float wx[4], wyy[4];
        cubic_coeffs(ix - fx0, wx);
        cubic_coeffs(iy - fy0, wyy);
#pragma unroll
        for (int k = 0; k < NT; k++) {
            cx[k] = wx[k];
            cy[k] = wyy[k];
            tx[k] = (int)fminf((float)(a.W - 1), fmaxf(fx0 - 1.f + (float)k, 0.f));
            ty[k] = (int)fminf((float)(a.H - 1), fmaxf(fy0 - 1.f + (float)k, 0.f));
        }
    } else {
        ix = fminf((float)(a.W - 1), fmaxf(ix, 0.f));
        iy = fminf((float)(a.H - 1), fmaxf(iy, 0.f));
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        tx[0] = (int)fx0; ty[0] = (int)fy0;
        tx[NT - 1] = min(tx[0] + 1, a.W - 1); ty[NT - 1] = min(ty[0] + 1, a.H - 1);
        cx[NT - 1] = ix - fx0; cy[NT - 1] = iy - fy0;
        cx[0] = 1.f - cx[NT - 1]; cy[0] = 1.f - cy[NT - 1];
    }
    __syncthreads();
    {
        int lox = inside ? tx[0] : 0x7fffffff, hix = inside ? tx[NT - 1] : -0x7fffffff;
        int loy = inside ? ty[0] : 0x7fffffff, hiy = inside ? ty[NT - 1] : -0x7fffffff;
        lox = __reduce_min_sync(0xffffffffu, lox); hix = __reduce_max_sync(0xffffffffu, hix);
        loy = __reduce_min_sync(0xffffffffu, loy); hiy = __reduce_max_sync(0xffffffffu, hiy);
        if (lane == 0) { atomicMin(&s_lim[0], lox); atomicMax(&s_lim[1], hix); atomicMin(&s_lim[2], loy); atomicMax(&s_lim[3], hiy); }
    }
    __syncthreads();
    const int bx0 = s_lim[0], by0 = s_lim[2], bw = s_lim[1] - s_lim[0] + 1, bh = s_lim[3] - s_lim[2] + 1;
    const bool staged = (bw <= WH_BW && bh <= WH_BH);      // block-uniform
    float *ob = a.out + (long long)b * a.os_b + (long long)y * a.os_h + (long long)x * a.os_w;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (staged) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(WH_BW * WH_BH * 16)) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(s_box)),
                         "l"(&tm), "r"(bx0 * 4), "r"(b * a.H + by0), "r"(bar)
                         : "memory");
        }
        unsigned ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar) : "memory");
        if (!inside) return;
#pragma unroll
        for (int r = 0; r < NT; r++) {
            const float4 *q = s_box + (ty[r] - by0) * WH_BW - bx0;
            float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
            if (INTERP == 1) {
                const float4 v0 = q[tx[0]], v1 = q[tx[1]], v2 = q[tx[NT - 2]], v3 = q[tx[NT - 1]];
                row.x = v0.x * cx[0] + v1.x * cx[1] + v2.x * cx[NT - 2] + v3.x * cx[NT - 1];
                row.y = v0.y * cx[0] + v1.y * cx[1] + v2.y * cx[NT - 2] + v3.y * cx[NT - 1];
                row.z = v0.z * cx[0] + v1.z * cx[1] + v2.z * cx[NT - 2] + v3.z * cx[NT - 1];
                row.w = v0.w * cx[0] + v1.w * cx[1] + v2.w * cx[NT - 2] + v3.w * cx[NT - 1];
                acc.x += row.x * cy[r]; acc.y += row.y * cy[r]; acc.z += row.z * cy[r]; acc.w += row.w * cy[r];
            } else {
                const float4 v0 = q[tx[0]], v1 = q[tx[NT - 1]];
                const float w0 = cx[0] * cy[r], w1 = cx[NT - 1] * cy[r];
                if (r == 0) {
                    acc.x = v0.x * w0 + v1.x * w1; acc.y = v0.y * w0 + v1.y * w1; acc.z = v0.z * w0 + v1.z * w1; acc.w = v0.w * w0 + v1.w * w1;
                } else {
                    acc.x = acc.x + v0.x * w0 + v1.x * w1; acc.y = acc.y + v0.y * w0 + v1.y * w1;
                    acc.z = acc.z + v0.z * w0 + v1.z * w1; acc.w = acc.w + v0.w * w0 + v1.w * w1;
                }
            }
        }
    } else {
        if (!inside) return;
        const float4 *xb = reinterpret_cast<const float4 *>(a.x + (long long)b * a.xs_b);
        const long long rowq = a.xs_h >> 2;
#pragma unroll
        for (int r = 0; r < NT; r++) {
            const float4 *q = xb + (long long)ty[r] * rowq;
            float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NT; k++) {
                const float4 v = __ldg(q + tx[k]);
                row.x += v.x * cx[k]; row.y += v.y * cx[k]; row.z += v.z * cx[k]; row.w += v.w * cx[k];
            }
            acc.x += row.x * cy[r]; acc.y += row.y * cy[r]; acc.z += row.z * cy[r]; acc.w += row.w * cy[r];
        }
    }
    if (a.os_c == 1 && ((reinterpret_cast<uintptr_t>(ob) & 15) == 0)) {
        *reinterpret_cast<float4 *>(ob) = acc;
    } else {
        ob[0] = acc.x; ob[a.os_c] = acc.y; ob[2 * a.os_c] = acc.z; ob[3 * a.os_c] = acc.w;
    }
}

cudaError_t launch_warp(const WarpArgs &a, cudaStream_t st)
{
    if (a.B <= 0 || a.C <= 0 || a.H <= 0 || a.W <= 0) return cudaSuccess;
    dim3 grid((a.W + 31) / 32, (a.H + 7) / 8, a.B);
    if (a.C == 4 && a.xs_c == 1 && a.xs_w == 4 && (a.xs_h & 3) == 0 && (a.xs_b & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(a.x) & 15) == 0) {   // channel-innermost 4-channel frames: 128-bit gathers
        // densely packed frames: the TMA-staged kernel (one tensor map over [B * H rows][W * 4 floats])
        if (a.xs_h == (long long)a.W * 4 && a.xs_b == (long long)a.H * a.W * 4 && a.W * 4LL <= 0x7fffffffLL &&
            (long long)a.B * a.H <= 0x7fffffffLL && getenv("RVDD_WARP_TMA")) {
            CUtensorMap tm;
            if (encode_map_2d(&tm, a.x, (unsigned long long)a.W * 4, (unsigned long long)a.B * a.H, (unsigned long long)a.W * 16,
                              WH_BW * 4, WH_BH) == cudaSuccess) {
                if (a.interp == 1)
                    warp_hwc4_tma_kernel<1><<<grid, 256, 0, st>>>(a, tm);
                else
                    warp_hwc4_tma_kernel<0><<<grid, 256, 0, st>>>(a, tm);
                return cudaGetLastError();
            }
        }
        if (a.interp == 1) {
            const bool one_px = getenv("RVDD_WARP_HWC_1PX") != nullptr;             // A/B switch (read per call): one pixel per thread
            if (one_px) {
                warp_hwc4_kernel<1><<<grid, 256, 0, st>>>(a);
            } else {
                warp_hwc4_2rows_kernel<<<dim3(grid.x, (a.H + 15) / 16, grid.z), 256, 0, st>>>(a);
            }
        } else {
            warp_hwc4_kernel<0><<<grid, 256, 0, st>>>(a);
        }
        return cudaGetLastError();
    }
    if (a.xs_w == 1 && a.C >= 3) {                      // plane-contiguous input: staged gathers
        if (a.interp == 1)
            warp_tile_kernel<1><<<grid, 256, 0, st>>>(a);
        else
            warp_tile_kernel<0><<<grid, 256, 0, st>>>(a);
    } else if (a.interp == 1) {
        warp_kernel<1><<<grid, 256, 0, st>>>(a);
    } else {
        warp_kernel<0><<<grid, 256, 0, st>>>(a);
    }
    return cudaGetLastError();
}

__global__ void upsample2_kernel(const float *__restrict__ in, float *__restrict__ out, int h, int w, float mul)
{
    const int H = 2 * h, W = 2 * w;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const float sy = (float)(h - 1) / (float)(H - 1), sx = (float)(w - 1) / (float)(W - 1);
    const float *p = in + (long long)blockIdx.z * h * w;
    out[((long long)blockIdx.z * H + y) * W + x] = up2_sample(p, h, w, y, x, sy, sx) * mul;
}

cudaError_t launch_upsample2(const float *in, float *out, long long planes, int h, int w, float mul, cudaStream_t st)
{
    if (planes <= 0) return cudaSuccess;
    if (planes > 65535) return cudaErrorInvalidValue;
    dim3 grid((2 * w + 31) / 32, (2 * h + 7) / 8, (unsigned)planes);
    upsample2_kernel<<<grid, 256, 0, st>>>(in, out, h, w, mul);
    return cudaGetLastError();
}

}  // namespace rvdd
