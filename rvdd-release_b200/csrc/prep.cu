// prep.cu -- batched fixed-cost stages in front of the solver (all pairs of a batch per launch):
//   gray conversion (library.py:162-170), joint min/max + normalisation (tvl1flow_lib.c:280-335),
//   separable Gaussian with the reference's asymmetric reflection (mask.c:214-330), bicubic resampling
//   for zoom_out (zoom.c:41-77).
// Everything is rounding-exact with respect to the reference (see exact_math.h).
#include "internal.h"

namespace rvdd {

// ------------------------------------------------------------------------------------------------ helpers

__device__ __forceinline__ int f2ord(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void setup_kernel(int *slots, int npairs, unsigned *bar, int nbar, int *status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npairs) {
        slots[2 * i] = 0x7fffffff;               // running min (ordered-int encoding)
        slots[2 * i + 1] = (int)0x80000000;      // running max
    }
    if (i < nbar) bar[i] = 0u;
    if (i == 0) { status[0] = 0; status[1] = 0; }
}

cudaError_t launch_setup(int *slots, int npairs, unsigned *bar, int nbar, int *status, cudaStream_t st)
{
    const int n = npairs > nbar ? npairs : nbar;
    setup_kernel<<<(n + 255) / 256, 256, 0, st>>>(slots, npairs, bar, nbar, status);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ gray

// c == 4: np.mean(axis=2) of a float32 HWC image = ((a+b)+c)+d, then / 4 (library.py:166, verified against
// numpy 2.3).  c == 3: skimage rgb2gray weights (library.py:163-164).  c == 1: copy.
__global__ void gray_kernel(const float *__restrict__ img, float *__restrict__ gray, long long n, int c)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (c == 4) {
            const float4 v = reinterpret_cast<const float4 *>(img)[i];
            gray[i] = FDIV(FADD(FADD(FADD(v.x, v.y), v.z), v.w), 4.0f);
        } else if (c == 3) {
            const float *p = img + 3 * i;
            gray[i] = FADD(FADD(FMUL(0.2125f, p[0]), FMUL(0.7154f, p[1])), FMUL(0.0721f, p[2]));
        } else {
            gray[i] = img[i];
        }
    }
}

cudaError_t launch_gray(const float *img, float *gray, long long n, int c, cudaStream_t st)
{
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    gray_kernel<<<(unsigned)blocks, 256, 0, st>>>(img, gray, n, c);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ min/max

__global__ void minmax_kernel(const float *const *__restrict__ I0, const float *const *__restrict__ I1, int n,
                              int *__restrict__ slots)
{
    const int k = blockIdx.y;
    const float *a = I0[k], *b = I1[k];
    float lo = a[0], hi = a[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float va = a[i], vb = b[i];
        lo = fminf(lo, fminf(va, vb));
        hi = fmaxf(hi, fmaxf(va, vb));
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&slots[2 * k], f2ord(lo));
        atomicMax(&slots[2 * k + 1], f2ord(hi));
    }
}

cudaError_t launch_minmax(const float *const *I0, const float *const *I1, int n, int npairs, int *slots, cudaStream_t st)
{
    int bx = (n + 256 * 8 - 1) / (256 * 8);
    if (bx > 64) bx = 64;
    if (bx < 1) bx = 1;
    minmax_kernel<<<dim3(bx, npairs), 256, 0, st>>>(I0, I1, n, slots);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ Gaussian

#define GT_X 64
#define GT_Y 32

// The reference accumulates in double (mask.c:279-285, :319-325); float -> double conversions run at a quarter of the
// FP64 rate on this GPU (16 / clk / SM), and a tap-by-tap conversion made them the kernel's bottleneck.  So every pixel
// is converted ONCE when the tile is staged, the row pass reads doubles, and its result -- rounded to float as the
// reference's in-place store does -- is converted back once for the column pass.
//
// DEC == false: dst = gaussian(src), same size.
// DEC == true : zoom_out with factor 0.5 (zoom.c:41-77): the bicubic sample of the blurred image at (2 x, 2 y) has zero
//               fractions and is the blurred pixel itself (bicubic_interpolation.c:100-108 with x = 0), so only the even
//               columns need the row pass and only the even rows / columns the column pass, and the quarter-size level
//               is written directly: dst is (nxx, nyy), the tile covers 2 GT_X x 2 GT_Y input pixels.
template <bool NORM, bool DEC, int RR>
__global__ void __launch_bounds__(256)
gauss_tile_kernel(const float *const *__restrict__ srcs, const float *__restrict__ src_base, long long src_stride,
                  float *__restrict__ dst_base, long long dst_stride, int nx, int ny, int nxx, int nyy, GaussTaps taps,
                  const int *__restrict__ slots, int npairs)
{
    extern __shared__ double smem_d[];
    constexpr int STEP = DEC ? 2 : 1;
    constexpr int TY = DEC ? GT_Y / 2 : GT_Y;      // output rows per tile (the decimating tile stages 2 TY + 2R input rows)
    const int z = blockIdx.z;
    const float *src = srcs ? srcs[z] : src_base + (long long)z * src_stride;
    float *dst = dst_base + (long long)z * dst_stride;
    constexpr int R = RR;                    // taps.size - 1 (mask.c:225), compile-time so the sliding windows stay in registers
    constexpr int GB = 4;                    // outputs per thread and pass
    constexpr int in_w = STEP * GT_X + 2 * R, in_h = STEP * TY + 2 * R;
    constexpr int in_p = in_w | 1;           // odd pitch: lanes walking down a column of doubles hit distinct banks
    constexpr int row_p = GT_X + 1;          // odd pitch again: the row pass writes s_row with lanes walking down a column
    double *s_in = smem_d;                   // [in_h][in_p]
    double *s_row = smem_d + in_h * in_p;    // [in_h][row_p]  (row-pass results at the columns the outputs need)
    const int ox = blockIdx.x * GT_X * STEP, oy = blockIdx.y * TY * STEP;        // tile origin in the input
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;

    float lo = 0.f, den = 0.f;
    if (NORM) {
        const int k = z % npairs;
        lo = ord2f(slots[2 * k]);
        den = FSUB(ord2f(slots[2 * k + 1]), lo);
    }
    // staging: all loads of a thread are issued before the first conversion (fully unrolled, loads into registers first)
    {
        constexpr int NR = (in_h + 7) / 8, NC = (in_w + 31) / 32;
        float v[NR][NC];
#pragma unroll
        for (int a = 0; a < NR; a++) {
            const int r = wrp + 8 * a;
            const float *srow = src + (long long)rvdd_reflect(oy - R + min(r, in_h - 1), ny) * nx;
#pragma unroll
            for (int b = 0; b < NC; b++) v[a][b] = __ldg(srow + rvdd_reflect(ox - R + min(lane + 32 * b, in_w - 1), nx));
        }
#pragma unroll
        for (int a = 0; a < NR; a++) {
            const int r = wrp + 8 * a;
#pragma unroll
            for (int b = 0; b < NC; b++) {
                const int c = lane + 32 * b;
                if (r < in_h && c < in_w) {
                    float t = v[a][b];
                    if (NORM && den > 0.f) t = rvdd_normalize_px(t, lo, den);
                    s_in[r * in_p + c] = (double)t;
                }
            }
        }
    }
    __syncthreads();
    // rows (mask.c:279-285): B[0]*R[i] + sum_j B[j]*(R[i-j]+R[i+j]) in double, stored as float.  A thread owns a run of
    // GB consecutive output columns of one row and slides over the 2R + STEP (GB - 1) + 1 inputs it needs (3x fewer
    // shared-memory reads than output by output); lanes map to rows, the odd row pitch keeps that conflict-free.
    for (int i = threadIdx.x; i < in_h * (GT_X / GB); i += 256) {
        const int r = i % in_h, c0 = (i / in_h) * GB;
        const double *row = s_in + r * in_p + STEP * c0;
        double w[2 * RR + STEP * (GB - 1) + 1];
#pragma unroll
        for (int k = 0; k < 2 * RR + STEP * (GB - 1) + 1; k++) w[k] = row[k];
#pragma unroll
        for (int o = 0; o < GB; o++) {
            double acc = DMUL(taps.B[0], w[STEP * o + RR]);
#pragma unroll
            for (int j = 1; j <= RR; j++) acc = DADD(acc, DMUL(taps.B[j], DADD(w[STEP * o + RR - j], w[STEP * o + RR + j])));
            s_row[r * row_p + c0 + o] = (double)(float)acc;
        }
    }
    __syncthreads();
    // columns (mask.c:319-325): a thread owns GB consecutive output rows of one column, lanes map to columns
    const int onx = DEC ? nxx : nx, ony = DEC ? nyy : ny;
    const int qx = blockIdx.x * GT_X, qy = blockIdx.y * TY;                      // tile origin in the output
    for (int i = threadIdx.x; i < (TY / GB) * GT_X; i += 256) {
        const int c = i % GT_X, r0 = (i / GT_X) * GB;
        const double *col = s_row + (STEP * r0) * row_p + c;
        double w[2 * RR + STEP * (GB - 1) + 1];
#pragma unroll
        for (int k = 0; k < 2 * RR + STEP * (GB - 1) + 1; k++) w[k] = col[k * row_p];
#pragma unroll
        for (int o = 0; o < GB; o++) {
            double acc = DMUL(taps.B[0], w[STEP * o + RR]);
#pragma unroll
            for (int j = 1; j <= RR; j++) acc = DADD(acc, DMUL(taps.B[j], DADD(w[STEP * o + RR - j], w[STEP * o + RR + j])));
            if (qy + r0 + o < ony && qx + c < onx) dst[(long long)(qy + r0 + o) * onx + qx + c] = (float)acc;
        }
    }
}

static size_t gauss_smem(int R, int step)
{
    const int ty = step == 2 ? GT_Y / 2 : GT_Y;
    return sizeof(double) * ((size_t)(step * ty + 2 * R) * ((step * GT_X + 2 * R) | 1) + (size_t)(step * ty + 2 * R) * (GT_X + 1));
}

template <bool NORM, bool DEC, int RR, typename... Args>
static cudaError_t gauss_go(dim3 grid, size_t smem, cudaStream_t st, Args... args)
{
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(gauss_tile_kernel<NORM, DEC, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    gauss_tile_kernel<NORM, DEC, RR><<<grid, 256, smem, st>>>(args...);
    return cudaGetLastError();
}

// The sliding windows need the tap radius R = taps.size - 1 (mask.c:225) at compile time.  The normalising variant is
// only ever used with the presmoothing kernel (sigma 0.8: R = 4), the decimating one with zoom factor 0.5 (sigma 1.04:
// R = 5); the plain variant is instantiated for R = 1 .. 12, which covers zoom factors down to 0.24; anything wider is
// refused (cudaErrorNotSupported -> "kernel too wide").
#define GAUSS_CASE(NORM, DEC, R, ...) case R: return gauss_go<NORM, DEC, R>(__VA_ARGS__);
cudaError_t launch_gauss(const float *const *srcs, const float *src_base, long long src_stride, float *dst_base,
                         long long dst_stride, int nx, int ny, int nimg, const GaussTaps &taps, const int *slots,
                         int npairs, cudaStream_t st)
{
    const int R = taps.size - 1;
    const size_t smem = gauss_smem(R, 1);
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    const dim3 grid((nx + GT_X - 1) / GT_X, (ny + GT_Y - 1) / GT_Y, nimg);
#define ARGS grid, smem, st, srcs, src_base, src_stride, dst_base, dst_stride, nx, ny, nx, ny, taps, slots, npairs
    if (slots) return R == 4 ? gauss_go<true, false, 4>(ARGS) : cudaErrorNotSupported;
    switch (R) {
        GAUSS_CASE(false, false, 1, ARGS) GAUSS_CASE(false, false, 2, ARGS) GAUSS_CASE(false, false, 3, ARGS)
        GAUSS_CASE(false, false, 4, ARGS) GAUSS_CASE(false, false, 5, ARGS) GAUSS_CASE(false, false, 6, ARGS)
        GAUSS_CASE(false, false, 7, ARGS) GAUSS_CASE(false, false, 8, ARGS) GAUSS_CASE(false, false, 9, ARGS)
        GAUSS_CASE(false, false, 10, ARGS) GAUSS_CASE(false, false, 11, ARGS) GAUSS_CASE(false, false, 12, ARGS)
    default: return cudaErrorNotSupported;
    }
#undef ARGS
}

// zoom_out with factor exactly 0.5: Gaussian + pick of the even pixels in one pass; returns cudaErrorNotSupported when the
// tile does not fit in shared memory (the caller then uses launch_gauss + launch_resample).
cudaError_t launch_gauss_decimate(const float *src_base, long long src_stride, float *dst_base, long long dst_stride, int nx,
                                  int ny, int nxx, int nyy, int nimg, const GaussTaps &taps, cudaStream_t st)
{
    const int R = taps.size - 1;
    const size_t smem = gauss_smem(R, 2);
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    const dim3 grid((nxx + GT_X - 1) / GT_X, (nyy + GT_Y / 2 - 1) / (GT_Y / 2), nimg);
    const float *const *nosrcs = nullptr;
    const int *noslots = nullptr;
    if (R != 5) return cudaErrorNotSupported;
    return gauss_go<false, true, 5>(grid, smem, st, nosrcs, src_base, src_stride, dst_base, dst_stride, nx, ny, nxx, nyy, taps, noslots, 1);
}

// ------------------------------------------------------------------------------------------------ resample

// zoom_out's sampling loop (zoom.c:66-74): out(x, y) = bicubic_at(src, x / fx, y / fy), clamped taps.
__global__ void resample_kernel(const float *__restrict__ src_base, long long src_stride, int nx, int ny,
                                float *__restrict__ dst_base, long long dst_stride, int nxx, int nyy, float fx, float fy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nxx || y >= nyy) return;
    const float *src = src_base + (long long)blockIdx.z * src_stride;
    float *dst = dst_base + (long long)blockIdx.z * dst_stride;
    dst[(long long)y * nxx + x] = rvdd_bicubic_clamped(src, FDIV((float)x, fx), FDIV((float)y, fy), nx, ny);
}

cudaError_t launch_resample(const float *src_base, long long src_stride, int nx, int ny, float *dst_base,
                            long long dst_stride, int nxx, int nyy, float fx, float fy, int nimg, cudaStream_t st)
{
    dim3 block(32, 8), grid((nxx + 31) / 32, (nyy + 7) / 8, nimg);
    resample_kernel<<<grid, block, 0, st>>>(src_base, src_stride, nx, ny, dst_base, dst_stride, nxx, nyy, fx, fy);
    return cudaGetLastError();
}

}  // namespace rvdd
