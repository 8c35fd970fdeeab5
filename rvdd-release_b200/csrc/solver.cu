// solver.cu -- the persistent TV-L1 solver: everything Dual_TVL1_optic_flow_multiscale does after the pyramid
// is built (tvl1flow_lib.c:403-453 and :91-273), for a whole batch of frame pairs, in ONE cooperative launch.
//
// The grid is split into `ngroups` groups of `ctas_per_group` CTAs.  A group owns one frame pair at a time and
// walks it through coarse-to-fine scales, warps and primal-dual iterations; groups never talk to each other, so
// while one group sits in the latency-bound coarse scales the others keep HBM busy with fine-scale iterations.
// Inside a group the phases (centred gradient, bicubic warp constants, one iteration, flow upsampling) are
// separated by a group barrier (monotonic counter in global memory, release/acquire, same structure as a
// cooperative-groups grid sync).  The barrier after an iteration also carries the residual reduction, so the
// reference's stopping rule `while (error > eps^2 && n < 300)` (tvl1flow_lib.c:163) is evaluated on the device
// after EVERY iteration, identically by every CTA of the group -- no host round trip, no speculation.
//
// One iteration is a single pass (64 B/pixel of HBM traffic: read u, p, the four per-warp constants, write u, p):
// each warp marches down a strip of rows, 4 pixels per lane; the thresholding step + primal update of the row
// below (needed by the forward differences of the dual update) is computed once and carried in registers to
// the next step, so only the strip's last row is evaluated twice.  u and p are double-buffered (read A, write B).
#include "internal.h"
#include "solver_core.h"

namespace rvdd {

#define SOLVER_THREADS 256
#define SOLVER_WARPS (SOLVER_THREADS / 32)

int solver_threads() { return SOLVER_THREADS; }

// ------------------------------------------------------------------------------------------------ group sync

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct GroupCtx {
    unsigned *bar;          // this group's counter
    unsigned target;        // value the counter reaches when everybody has arrived (thread 0 only)
    int nctas, cta;         // group size, rank inside the group
    double *partials;       // [2][nctas]
    int *status;
    long long spin_limit;
    int slot;
};

// Returns false if the watchdog fired (somebody waited longer than spin_limit): the kernel then unwinds.
__device__ __forceinline__ bool group_sync(GroupCtx &g, int *s_flag)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        g.target += (unsigned)g.nctas;
        __threadfence();
        atomicAdd(g.bar, 1u);
        const long long t0 = clock64();
        int dead = 0;
        unsigned spins = 0;
        while ((int)(ld_acquire_u32(g.bar) - g.target) < 0) {
            if ((++spins & 1023u) == 0) {
                if (*(volatile int *)g.status != 0) { dead = 1; break; }
                if (clock64() - t0 > g.spin_limit) { atomicExch(g.status, 1); dead = 1; break; }
            }
        }
        if (!dead && *(volatile int *)g.status != 0) dead = 1;
        __threadfence();
        *s_flag = dead;
    }
    __syncthreads();
    return *s_flag == 0;
}

// Sum of the per-CTA residual partials of this group, in a fixed order (deterministic, identical in every CTA).
__device__ __forceinline__ double group_sum(const GroupCtx &g, int slot, double *s_val)
{
    if (threadIdx.x < 32) {
        double acc = 0.0;
        const double *p = g.partials + (size_t)slot * g.nctas;
        for (int i = threadIdx.x; i < g.nctas; i += 32) acc += __ldcg(p + i);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) *s_val = acc;
    }
    __syncthreads();
    return *s_val;
}

// Distribute the image over the group's warps: column segments of 32*V pixels, strips of `rows` rows.
template <int V>
__device__ __forceinline__ double iterate_group(const IterPtrs &P, int nx, int ny, const IterConsts &K, int gwarp,
                                                int nwarps_group)
{
    const int lane = threadIdx.x & 31;
    const StripPlan sp = plan_strips<V>(nx, ny, nwarps_group);
    const int segw = 32 * V, ncol = sp.ncol, rows = sp.rows, total = sp.total;
    double err = 0.0;
    for (int w = gwarp; w < total; w += nwarps_group) {
        const int col = w % ncol, strip = w / ncol;
        const int x0 = col * segw + lane * V;
        const int y0 = strip * rows;
        const int y1 = min(ny, y0 + rows);
        if (x0 < nx) err += iterate_strip<V>(P, x0, col * segw, y0, y1, nx, ny, K);
    }
    return err;
}

// ------------------------------------------------------------------------------------------------ the kernel

__global__ void __launch_bounds__(SOLVER_THREADS, 2) solver_kernel(const SolverArgs A)
{
    __shared__ double s_red[SOLVER_WARPS];
    __shared__ double s_val;
    __shared__ int s_flag;

    const int group = blockIdx.x / A.ctas_per_group;
    GroupCtx g;
    g.nctas = A.ctas_per_group;
    g.cta = blockIdx.x - group * A.ctas_per_group;
    g.bar = A.bar + group * 32;
    g.target = 0u;
    g.partials = A.partials + (size_t)group * 2 * A.ctas_per_group;
    g.status = A.status;
    g.spin_limit = A.spin_limit;
    g.slot = 0;
    if (group >= A.ngroups) return;

    const int gthreads = g.nctas * SOLVER_THREADS;
    const int gtid = g.cta * SOLVER_THREADS + threadIdx.x;
    const int gwarp = g.cta * SOLVER_WARPS + (threadIdx.x >> 5);
    const int gwarps = g.nctas * SOLVER_WARPS;

    // per-group scratch planes
    float *S = A.scratch + (long long)group * A.scratch_stride;
    const long long PL = A.plane;
    float *I1x = S, *I1y = S + PL, *gx = S + 2 * PL, *gy = S + 3 * PL, *g2 = S + 4 * PL, *rc = S + 5 * PL;
    // flow and dual variable are double-buffered: plane 6 + 2*buf + comp and 10 + 4*buf + comp (no pointer tables:
    // indexing a local array of pointers would force generic loads and local memory)
#define UB(buf, comp) (S + (6 + 2 * (buf) + (comp)) * PL)
#define PB(buf, comp) (S + (10 + 4 * (buf) + (comp)) * PL)
    IterConsts K;
    K.l_t = A.l_t; K.theta = A.theta; K.taut = A.taut; K.g0f = A.g0f;

    for (int pair = group; pair < A.npairs; pair += A.ngroups) {
        const float *P0 = A.pyr0 + (long long)pair * A.pyr_stride;
        const float *P1 = A.pyr1 + (long long)pair * A.pyr_stride;
        int uc = 0, pc = 0;

        for (int s = A.S - 1; s >= 0; s--) {
            const int nx = A.nx[s], ny = A.ny[s], n = nx * ny;
            const float *I0 = P0 + A.off[s], *I1 = P1 + A.off[s];

            if (s == A.S - 1) {
                // ---- flow = 0 at the coarsest scale (:404-405)
                for (int i = gtid; i < n; i += gthreads) { UB(uc, 0)[i] = 0.f; UB(uc, 1)[i] = 0.f; }
                if (s < A.fscale && !group_sync(g, &s_flag)) return;
            }
            if (s >= A.fscale) {
                // ---- per-scale setup: p = 0 (:134-138), centred gradient of I1 (:131, mask.c:149-206)
                for (int i = gtid; i < n; i += gthreads) {
                    const int y = i / nx, x = i - y * nx;
                    PB(pc, 0)[i] = 0.f; PB(pc, 1)[i] = 0.f; PB(pc, 2)[i] = 0.f; PB(pc, 3)[i] = 0.f;
                    cgrad_px(I1, x, y, nx, ny, &I1x[i], &I1y[i]);
                }
                if (!group_sync(g, &s_flag)) return;

                for (int w = 0; w < A.nwarps; w++) {
                    // ---- warp constants (:143-159): bicubic samples of I1, I1x, I1y at x + u
                    const float *u1 = UB(uc, 0), *u2 = UB(uc, 1);
                    for (int i = gtid; i < n; i += gthreads) {
                        const int y = i / nx, x = i - y * nx;
                        warp_consts_px(I0, I1, I1x, I1y, u1[i], u2[i], x, y, nx, ny, &gx[i], &gy[i], &g2[i], &rc[i]);
                    }
                    if (!group_sync(g, &s_flag)) return;

                    // ---- inner loop (:161-244), stop test after every iteration
                    int it = 0;
                    float err = INFINITY;
                    while (err > A.eps2 && it < RVDD_MAX_ITERATIONS) {
                        it++;
                        IterPtrs P;
                        P.u1 = UB(uc, 0); P.u2 = UB(uc, 1);
                        P.p11 = PB(pc, 0); P.p12 = PB(pc, 1); P.p21 = PB(pc, 2); P.p22 = PB(pc, 3);
                        P.nu1 = UB(uc ^ 1, 0); P.nu2 = UB(uc ^ 1, 1);
                        P.np11 = PB(pc ^ 1, 0); P.np12 = PB(pc ^ 1, 1); P.np21 = PB(pc ^ 1, 2); P.np22 = PB(pc ^ 1, 3);
                        P.gx = gx; P.gy = gy; P.g2 = g2; P.rc = rc;
                        double e = ((nx & 3) == 0) ? iterate_group<4>(P, nx, ny, K, gwarp, gwarps)
                                                   : iterate_group<1>(P, nx, ny, K, gwarp, gwarps);
                        // CTA partial in a fixed order, then the group reduction rides on the barrier
                        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
                        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = e;
                        __syncthreads();
                        if (threadIdx.x == 0) {
                            double t = 0.0;
                            for (int k = 0; k < SOLVER_WARPS; k++) t += s_red[k];
                            g.partials[(size_t)g.slot * g.nctas + g.cta] = t;
                        }
                        if (!group_sync(g, &s_flag)) return;
                        const double tot = group_sum(g, g.slot, &s_val);
                        g.slot ^= 1;
                        err = FDIV((float)tot, (float)n);        // error /= size (:223)
                        uc ^= 1;
                        pc ^= 1;
                    }
                    if (g.cta == 0 && threadIdx.x == 0) {
                        const long long t = ((long long)pair * RVDD_MAX_SCALES + s) * A.nwarps + w;
                        if (A.iters_out) A.iters_out[t] = it;
                        if (A.err_out) A.err_out[t] = err;
                    }
                }
            }

            if (s > 0) {
                // ---- zoom_in to the next finer level and rescale (:425-433, zoom.c:85-109)
                const int fx_n = A.nx[s - 1], fy_n = A.ny[s - 1], fn = fx_n * fy_n;
                const float *c1 = UB(uc, 0), *c2 = UB(uc, 1);
                float *f1 = UB(uc ^ 1, 0), *f2 = UB(uc ^ 1, 1);
                const float zx = A.zfx[s - 1], zy = A.zfy[s - 1];
                for (int i = gtid; i < fn; i += gthreads) {
                    const int y = i / fx_n, x = i - y * fx_n;
                    f1[i] = zoom_in_px(c1, x, y, nx, ny, zx, zy, A.zoom_mul);
                    f2[i] = zoom_in_px(c2, x, y, nx, ny, zx, zy, A.zoom_mul);
                }
                uc ^= 1;
            } else {
                // ---- finest flow -> caller's planar (u, v) buffer (:374-375)
                float *o1 = A.flow_out + (long long)pair * 2 * n, *o2 = o1 + n;
                const float *c1 = UB(uc, 0), *c2 = UB(uc, 1);
                for (int i = gtid; i < n; i += gthreads) { o1[i] = c1[i]; o2[i] = c2[i]; }
            }
            if (!group_sync(g, &s_flag)) return;
        }
    }
}

cudaError_t solver_max_ctas(int *ctas_per_sm, int *sms)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, solver_kernel, SOLVER_THREADS, 0);
}

cudaError_t launch_solver(const SolverArgs &args, cudaStream_t st)
{
    void *params[] = {(void *)&args};
    const dim3 grid(args.ngroups * args.ctas_per_group), block(SOLVER_THREADS);
    return cudaLaunchCooperativeKernel((const void *)solver_kernel, grid, block, params, 0, st);
}

}  // namespace rvdd
