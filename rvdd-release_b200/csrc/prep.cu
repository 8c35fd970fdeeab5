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
    if (i == 0) status[0] = 0;
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

template <bool NORM>
__global__ void __launch_bounds__(256)
gauss_tile_kernel(const float *const *__restrict__ srcs, const float *__restrict__ src_base, long long src_stride,
                  float *__restrict__ dst_base, long long dst_stride, int nx, int ny, GaussTaps taps,
                  const int *__restrict__ slots, int npairs)
{
    extern __shared__ float smem[];
    const int z = blockIdx.z;
    const float *src = srcs ? srcs[z] : src_base + (long long)z * src_stride;
    float *dst = dst_base + (long long)z * dst_stride;
    const int R = taps.size - 1;
    const int in_w = GT_X + 2 * R, in_h = GT_Y + 2 * R;
    float *s_in = smem;                      // [in_h][in_w]
    float *s_row = smem + in_h * in_w;       // [in_h][GT_X]
    const int ox = blockIdx.x * GT_X, oy = blockIdx.y * GT_Y;
    const int tid = threadIdx.x;

    float lo = 0.f, den = 0.f;
    if (NORM) {
        const int k = z % npairs;
        lo = ord2f(slots[2 * k]);
        den = FSUB(ord2f(slots[2 * k + 1]), lo);
    }
    for (int i = tid; i < in_h * in_w; i += 256) {
        const int r = i / in_w, c = i - r * in_w;
        const int gy = rvdd_reflect(oy - R + r, ny), gx = rvdd_reflect(ox - R + c, nx);
        float v = src[(long long)gy * nx + gx];
        if (NORM && den > 0.f) v = rvdd_normalize_px(v, lo, den);
        s_in[i] = v;
    }
    __syncthreads();
    // rows (mask.c:279-285): B[0]*R[i] + sum_j B[j]*(R[i-j]+R[i+j]) in double, stored as float
    for (int i = tid; i < in_h * GT_X; i += 256) {
        const int r = i / GT_X, c = i - r * GT_X;
        const float *row = s_in + r * in_w + c + R;
        double acc = DMUL(taps.B[0], (double)row[0]);
        for (int j = 1; j <= R; j++) acc = DADD(acc, DMUL(taps.B[j], DADD((double)row[-j], (double)row[j])));
        s_row[i] = (float)acc;
    }
    __syncthreads();
    // columns (mask.c:319-325)
    for (int i = tid; i < GT_Y * GT_X; i += 256) {
        const int r = i / GT_X, c = i - r * GT_X;
        if (oy + r >= ny || ox + c >= nx) continue;
        const float *col = s_row + (r + R) * GT_X + c;
        double acc = DMUL(taps.B[0], (double)col[0]);
        for (int j = 1; j <= R; j++)
            acc = DADD(acc, DMUL(taps.B[j], DADD((double)col[-j * GT_X], (double)col[j * GT_X])));
        dst[(long long)(oy + r) * nx + ox + c] = (float)acc;
    }
}

cudaError_t launch_gauss(const float *const *srcs, const float *src_base, long long src_stride, float *dst_base,
                         long long dst_stride, int nx, int ny, int nimg, const GaussTaps &taps, const int *slots,
                         int npairs, cudaStream_t st)
{
    const int R = taps.size - 1;
    const size_t smem = sizeof(float) * ((size_t)(GT_Y + 2 * R) * (GT_X + 2 * R) + (size_t)(GT_Y + 2 * R) * GT_X);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(gauss_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(gauss_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_done = true;
    }
    if (smem > 160 * 1024) return cudaErrorInvalidValue;
    dim3 grid((nx + GT_X - 1) / GT_X, (ny + GT_Y - 1) / GT_Y, nimg);
    if (slots)
        gauss_tile_kernel<true><<<grid, 256, smem, st>>>(srcs, src_base, src_stride, dst_base, dst_stride, nx, ny, taps, slots, npairs);
    else
        gauss_tile_kernel<false><<<grid, 256, smem, st>>>(srcs, src_base, src_stride, dst_base, dst_stride, nx, ny, taps, slots, npairs);
    return cudaGetLastError();
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
