// Issue rates of the double-precision instructions the warp-constants phase is made of, per SM per clock, on the GPU at hand:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rates fp64_rates.cu && ./fp64_rates
// Each kernel runs ILP independent chains per thread, `warps` warps per SM on every SM, and reports lane-operations per clock
// per SM from clock64() around the loop (max over the block's warps).
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void rate_kernel(int iters, double seed, float fseed, double *sink, long long *clocks)
{
    double x[ILP];
    float f[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) { x[k] = seed + k + threadIdx.x; f[k] = fseed + k + threadIdx.x; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if (OP == 0) x[k] = __dadd_rn(x[k], seed);
            if (OP == 1) x[k] = __dmul_rn(x[k], seed);
            if (OP == 2) x[k] = __fma_rn(x[k], seed, seed);
            if (OP == 3) { x[k] = (double)f[k]; f[k] = __int_as_float(__float_as_int(f[k]) + (int)__double2hiint(x[k])); }   // F2F.F64.F32 + 1 IADD
            if (OP == 5) { f[k] = __double2float_rn(x[k]); x[k] = __hiloint2double(__float_as_int(f[k]), __double2loint(x[k])); }   // F2F.F32.F64
            if (OP == 6) { x[k] = __dadd_rn(x[k], seed); f[k] = __fadd_rn(f[k], fseed); }                                      // DADD + FADD
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k] + f[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)&clocks[blockIdx.x], (unsigned long long)(t1 - t0));
}

template <int OP>
void run(const char *name, int warps, int sms)
{
    const int ILP = 8, iters = 4096;
    double *sink; long long *clk;
    cudaMalloc(&sink, sizeof(double) * sms * warps * 32);
    cudaMalloc(&clk, sizeof(long long) * sms);
    cudaMemset(clk, 0, sizeof(long long) * sms);
    rate_kernel<OP, ILP><<<sms, warps * 32>>>(iters, 1.0000001, 1.5f, sink, clk);
    cudaMemset(clk, 0, sizeof(long long) * sms);
    rate_kernel<OP, ILP><<<sms, warps * 32>>>(iters, 1.0000001, 1.5f, sink, clk);
    long long h[256];
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-34s warps/SM %2d  lane-ops/clk/SM %7.2f  (%s)\n", name, warps, (double)warps * 32 * ILP * iters / (double)mx, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink); cudaFree(clk);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {4, 12, 32}) {
        run<0>("DADD", w, p.multiProcessorCount);
        run<1>("DMUL", w, p.multiProcessorCount);
        run<2>("DFMA", w, p.multiProcessorCount);
        run<3>("F2F.F64.F32 (+IADD)", w, p.multiProcessorCount);
        run<5>("F2F.F32.F64 (+mov)", w, p.multiProcessorCount);
        run<6>("DADD + FADD pairs (DADD count)", w, p.multiProcessorCount);
    }
    return 0;
}
