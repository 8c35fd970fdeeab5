// selftest.cu -- test hook: exhaustive-style comparison of the straight-line "fast path" exact operations
// (exact_math.h: rvdd_hypot_fast, rvdd_div_by_rcp) against the reference-exact ones (double sqrt, IEEE division) on
// pseudo-random operands, on the device.  Any accepted (not `bad`) fast result that differs in even one bit is a bug.
#include "../../include/rvdd_bridge.h"
#include "internal.h"

namespace rvdd {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// random float with a random sign, a full random mantissa and an exponent drawn from [emin, emax]
__device__ __forceinline__ float rnd_float(unsigned long long r, int emin, int emax)
{
    const unsigned man = (unsigned)(r & 0x7fffffu);
    const unsigned sgn = (unsigned)((r >> 23) & 1u) << 31;
    const int e = emin + (int)((r >> 24) % (unsigned)(emax - emin + 1));
    return __uint_as_float(sgn | ((unsigned)(e + 127) << 23) | man);
}

// counters: [0] hypot tested, [1] hypot bad (fell back), [2] hypot mismatches, [3] div tested, [4] div rejected,
//           [5] div mismatches
__global__ void selftest_kernel(unsigned long long seed, int iters, unsigned long long *counters)
{
#if defined(__CUDA_ARCH__)      // the fast-path functions only exist in the device pass
    unsigned long long c[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long state = seed + 0x1234567ULL * (blockIdx.x * blockDim.x + threadIdx.x);
    for (int it = 0; it < iters; it++) {
        const unsigned long long r1 = mix64(state++), r2 = mix64(state++), r3 = mix64(state++);
        // ---- hypot: flow differences are small numbers; vary the exponent gap between the two operands, and visit the
        // whole window the fast path accepts (2^-100 < max(|a|, |b|) < 2^40) including its edges
        const int mode = (int)(r3 & 7);
        float a = rnd_float(r1, -30, 8), b = rnd_float(r2, -30, 8);
        if (mode == 5) a = rnd_float(r1, -104, -60), b = rnd_float(r2, -126, -60);
        if (mode == 6) a = rnd_float(r1, 20, 41), b = rnd_float(r2, -10, 41);
        if (mode == 0) b = a * (1.0f + (float)((r3 >> 8) & 0xff) * 1.1920929e-07f);   // nearly equal
        if (mode == 1) b = 0.0f;
        if (mode == 2) a = rnd_float(r1, -3, 3), b = rnd_float(r2, -3, 3);
        if (mode == 3) a = (float)(int)((r1 >> 40) & 0xfff) * 0.125f, b = (float)(int)((r2 >> 40) & 0xfff) * 0.125f;   // exact roots
        bool bad = false;
        const float gf = rvdd_hypot_fast(a, b, bad);
        const float ge = rvdd_hypotf_wide(a, b);
        c[0]++;
        if (bad) c[1]++;
        else if (__float_as_uint(gf) != __float_as_uint(ge)) c[2]++;
        // ---- division as used by the dual update (divisor 1 + taut * g in [1, 2^41)) and by the thresholding step (divisor
        // |grad|^2 in [1e-10, 2^21), or 1): accepted exactly when the callers accept it -- the numerator is 0 or at least 2^-60
        // in magnitude (rvdd_num_key / RVDD_KEY_2M60) -- with numerators down to the threshold and a little below
        const float num = (mode == 4) ? 0.0f : ((mode == 6) ? rnd_float(r2, -64, -56) : rnd_float(r2, -62, 24));
        const float den = (mode & 1) ? (1.0f + fabsf(rnd_float(r1, -20, 40))) : fabsf(rnd_float(r1, -34, 20));
        const float q = rvdd_div_by_rcp(num, den, rvdd_rcp_refined(den));
        c[3]++;
        if (!(den < 2.199023255552e12f) || rvdd_num_key(num) < RVDD_KEY_2M60) c[4]++;          // den < 2^41
        else if (__float_as_uint(q) != __float_as_uint(__fdiv_rn(num, den))) c[5]++;
    }
    for (int k = 0; k < 6; k++) atomicAdd(&counters[k], c[k]);
#endif
}

}  // namespace rvdd

extern "C" RVDD_API int rvdd_selftest_fastmath(unsigned long long seed, int blocks, int iters, unsigned long long *counters_host)
{
    unsigned long long *d = nullptr;
    if (cudaMalloc(&d, 6 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    cudaMemset(d, 0, 6 * sizeof(unsigned long long));
    rvdd::selftest_kernel<<<blocks, 256>>>(seed, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(counters_host, d, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? 0 : (int)e;
}
