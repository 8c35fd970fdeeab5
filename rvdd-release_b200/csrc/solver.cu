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
// One iteration is a single pass (60 B/pixel of HBM traffic: read u, p and three per-warp constants, write u, p; the
// reference's fourth per-warp array, |grad|^2, is recomputed from I1wx, I1wy instead of being stored):
// each warp marches down a strip of rows, 4 pixels per lane; the thresholding step + primal update of the row
// below (needed by the forward differences of the dual update) is computed once and carried in registers to
// the next step, so only the strip's last row is evaluated twice.  u and p are double-buffered (read A, write B).
#include "internal.h"
#include "solver_core.h"

namespace rvdd {

// 192 threads x 2 CTAs per SM = 12 warps with up to 168 registers each: the 4-pixel lane state (two row states, the
// staged inputs, the address arithmetic) spills at 128 registers, and spill reloads sit on the critical path of every
// row.  Measured on B200 (29 pairs, 1280x720): 256x2 (128 regs) 58 ms, 256x1 (207 regs) 52 ms, 384x1 / 192x2
// (168 regs) 46 ms.
#ifndef SOLVER_THREADS
#define SOLVER_THREADS 192
#endif
#ifndef SOLVER_MIN_CTAS
#define SOLVER_MIN_CTAS 2
#endif
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
    double *partials;       // [2 slots][2 sums][nctas]
    int *status;
    long long spin_limit;
    int slot;
};

// Returns false if the watchdog fired (somebody waited longer than spin_limit): the kernel then unwinds.
__device__ __forceinline__ bool group_sync(GroupCtx &g, int *s_flag)
{
    asm volatile("fence.proxy.async;" ::: "memory");   // this thread's stores may next be read by bulk async copies
    __syncthreads();
    if (threadIdx.x == 0) {
        g.target += (unsigned)g.nctas;
        __threadfence();
        atomicAdd(g.bar, 1u);
        const long long t0 = clock64();
        int dead = 0;
        unsigned spins = 0;
        while ((int)(ld_acquire_u32(g.bar) - g.target) < 0) {
            if ((spins & 1023u) == 0) {      // the clock is looked at on the first unsuccessful poll, then every 1024
                if (spins && *(volatile int *)g.status != 0) { dead = 1; break; }
                if (clock64() - t0 > g.spin_limit) { atomicExch(g.status, 1); dead = 1; break; }
            }
            ++spins;
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

// ------------------------------------------------------------------------------------------------ strip iteration

#ifndef SOLVER_V
#define SOLVER_V 4
#endif
#ifndef WC_BATCH
#define WC_BATCH 2                       // pixels per thread and trip of the warp-constants phase (loads in flight)
#endif

// ---- bulk-copy (TMA) row staging, used when the image width is a multiple of 4 ---------------------------------
//
// By Little's law the direct-load strip loop cannot keep HBM busy: a warp asks for one row (5 KB), waits ~2 us for
// it, computes, and only then asks for the next; with 16 warps per SM (the state of a 4-pixel lane needs ~128
// registers) that is ~30 KB in flight per SM, good for ~2.5 TB/s.  So every warp stages the rows it is about to
// process in shared memory with 1-D bulk asynchronous copies (cp.async.bulk global -> shared, completion on an
// mbarrier): 9 arrays x (128 pixels + a 4-pixel block on either side) per row, two rows per warp in flight.  The
// copy of row y+3 is issued as soon as row y+1 has been consumed, so the loads of the next rows overlap the
// arithmetic of the current one.  The right/left neighbour pixels come out of the same staged row.

#define ST_PAD 4                         // pixels staged on either side of the warp's 128
#define ST_SLOT (128 + 2 * ST_PAD)       // floats per array per row (544 B, a multiple of 16 B)
// One staged row = three TMA boxes, each landing at a 128-byte aligned offset of the stage:
//   [gx gy rc] (3 planes) at float 0, [u1 u2] at float 416, [p11 p12 p21 p22] at float 704
#define ST_OFF_C 0
#define ST_OFF_U 416
#define ST_OFF_P 704
#define ST_ROW 1248                      // floats per stage (4992 B, a multiple of 128 B)
// Ring depth NST (a template parameter of everything below): 2 rows in flight for the single-iteration pass; the pass that
// fuses two iterations keeps rows L-1 and L while row L+1 is in flight and uses 3.
#define ST_WARP_BYTES(NST) ((NST) * ST_ROW * 4 + 128)  // + the mbarriers; keeps every warp's stages 128-byte aligned

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Per-warp ring of NST staged rows.  Rows are issued and consumed strictly in order, so two running counters say
// everything: row number n lives in stage n % NST and completes that stage's mbarrier phase (n / NST) & 1.
template <int NST> struct TmaRing {
    unsigned char *base;                 // stage k at base + k * ST_ROW * 4, mbarrier k at base + NST * ST_ROW * 4 + 16 k
    unsigned issued, taken;
    __device__ __forceinline__ float *stage(unsigned k) const { return reinterpret_cast<float *>(base + (size_t)k * ST_ROW * 4); }
    __device__ __forceinline__ unsigned long long *bar(unsigned k) const
    {
        return reinterpret_cast<unsigned long long *>(base + (size_t)NST * ST_ROW * 4 + 16 * k);
    }
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// One lane of a converged warp.  Unlike `lane == 0`, elect.sync tells the compiler that exactly one thread runs the
// guarded region, so the bulk copies (uniform-datapath instructions) are issued once instead of in a per-lane loop.
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Row coordinates of a strip for the tensor-map copies: c0 = linear float index of (y, warp_x0 - ST_PAD) inside a plane,
// plane rows of the three families in the [ngroups * RVDD_NPLANES][plane] view of the scratch.
struct TmaSrc {
    int c0, row_c, row_u, row_p;
};

__device__ __forceinline__ void tma_box(const CUtensorMap *tm, float *dst, int c0, int c1, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// elected lane: queue the three box copies of the image row `delta` floats below the strip's first row into `stage`.
// Out-of-range coordinates (left of pixel 0 of the plane, beyond its end) are zero-filled by the TMA unit and still
// count towards the transaction bytes, so the byte count is a constant.
template <int NST>
__device__ __forceinline__ void tma_issue_row(const SolverArgs &A, const TmaSrc &Q, int delta, const TmaRing<NST> &T, unsigned n)
{
    float *stage = T.stage(n % NST);
    unsigned long long *bar = T.bar(n % NST);
    mbar_expect_tx(bar, 9u * ST_SLOT * 4u);
    tma_box(&A.tm3, stage + ST_OFF_C, Q.c0 + delta, Q.row_c, bar);
    tma_box(&A.tm2, stage + ST_OFF_U, Q.c0 + delta, Q.row_u, bar);
    tma_box(&A.tm4, stage + ST_OFF_P, Q.c0 + delta, Q.row_p, bar);
}

// all lanes: wait for stage `st`, then evaluate the staged row straight out of shared memory.  The row is consumed in
// two halves (dual variable -> divergence, then flow + constants -> primal update) so that at most half of its 47
// input values are live in registers at any time.
template <int NST>
__device__ __forceinline__ void tma_eval_row(TmaRing<NST> &T, int lane, const LaneEdges &E, bool first, bool last,
                                             const IterConsts &K, const float (&up12)[5], const float (&up22)[5],
                                             RowState<4> &R, int *status)
{
    const unsigned st = T.taken % NST, parity = (T.taken / NST) & 1u;
    T.taken++;
    unsigned long long *bar = T.bar(st);
    bool ok = mbar_try_wait(bar, parity);
    for (unsigned spins = 0; !ok; ++spins) {
        ok = mbar_try_wait(bar, parity);
        if (!ok && spins > (1u << 22)) {            // ~ seconds: something is badly wrong, do not hang the GPU
            atomicExch(status, 2);
            break;
        }
    }
    const float *s = T.stage(st) + ST_PAD + 4 * lane;
    RowIn<4> I;
    float4 v;
    // The right neighbour column (index 4) is taken as staged even where no such column exists (last lane of the image row:
    // it then holds the next row's first pixel); whatever is computed from it is never used (finish_row zeroes the forward
    // difference on the last column).  Likewise the left neighbour of column 0 is the previous row's last column of p11 / p21,
    // which is exactly zero (see eval_div, ZB), or the zero fill of the TMA unit above the first row.
#define TAKE(o, dst)                                                \
    v = *reinterpret_cast<const float4 *>(s + (o));                 \
    I.dst[0] = v.x; I.dst[1] = v.y; I.dst[2] = v.z; I.dst[3] = v.w; \
    I.dst[4] = s[(o) + 4];
    TAKE(ST_OFF_P + 1 * ST_SLOT, p12) TAKE(ST_OFF_P + 3 * ST_SLOT, p22) TAKE(ST_OFF_P, a11) TAKE(ST_OFF_P + 2 * ST_SLOT, a21)
    I.l11 = s[ST_OFF_P - 1];
    I.l21 = s[ST_OFF_P + 2 * ST_SLOT - 1];
    float d1[5], d2[5];
    eval_div<4, true>(I, E, first, last, up12, up22, R, d1, d2);
    asm volatile("" ::: "memory");                  // keep the second half's shared-memory loads below this point
    TAKE(ST_OFF_U, u1) TAKE(ST_OFF_U + ST_SLOT, u2)
    TAKE(ST_OFF_C, gx) TAKE(ST_OFF_C + ST_SLOT, gy) TAKE(ST_OFF_C + 2 * ST_SLOT, rc)
#undef TAKE
    eval_primal<4>(I, K, d1, d2, R);
}

// One warp's strip with staged rows (V = 4): same arithmetic as iterate_strip<4>, different data path.
template <int NST>
__device__ __forceinline__ double iterate_strip_tma(const SolverArgs &SA, int group, const IterPtrs &P, TmaRing<NST> &T, int lane,
                                                    int warp_x0, int y0, int y1, int nx, int ny, const IterConsts &K,
                                                    int *status)
{
    const int x0 = warp_x0 + 4 * lane;
    const bool active = x0 < nx;
    const LaneEdges E = lane_edges<4>(active ? x0 : 0x3fffff00, nx, warp_x0);
    const int yend = (y1 < ny) ? y1 : ny - 1;            // last row to evaluate (the strip's halo row if any)
    const int nr = yend - y0 + 1;
    const long long wrow = (long long)y0 * nx + warp_x0;
    TmaSrc Q;
    Q.c0 = (int)wrow - ST_PAD;
    Q.row_c = group * RVDD_NPLANES + RVDD_PL_C; Q.row_u = group * RVDD_NPLANES + RVDD_PL_U + 2 * P.uc;
    Q.row_p = group * RVDD_NPLANES + RVDD_PL_P + 4 * P.pc;
    double err = 0.0;

    // prologue: the first NST rows of the strip go in flight at once
    __syncwarp();
    const unsigned n0 = T.issued;                        // == T.taken: the ring is empty between strips
    if (elect_one()) {
        fence_proxy_async();                             // other CTAs' stores (generic proxy) -> our bulk reads
#pragma unroll
        for (int k = 0; k < NST; k++)
            if (k < nr) tma_issue_row(SA, Q, k * nx, T, n0 + k);
    }
    T.issued = n0 + (unsigned)min(nr, NST);
    float up12[5], up22[5];
    long long row = (long long)y0 * nx + x0;
    if (active) {
        load_up_row<4>(P, row, nx, y0, E, up12, up22);
    } else {
#pragma unroll
        for (int j = 0; j < 5; j++) up12[j] = up22[j] = 0.f;
    }

    // after a row has been consumed its stage is refilled with the row NST further down
    auto refill = [&](int consumed) {
        __syncwarp();                                    // every lane has pulled its values out of the stage
        if (consumed + NST < nr) {
            if (elect_one()) tma_issue_row(SA, Q, (consumed + NST) * nx, T, T.issued);
            T.issued++;
        }
    };

    RowState<4> A, B;
    tma_eval_row(T, lane, E, y0 == 0, y0 == ny - 1, K, up12, up22, A, status);
    refill(0);

    int y = y0, i = 0;                                   // i = y - y0
    // One copy of the row body; the row state is moved from B to A after every row.  Unrolling by two (the states
    // swapping roles) saves the 32 moves but doubles the hot loop's code, and measured 1.3 % slower (instruction
    // cache: stall_no_inst 6 %).
#pragma unroll 1
    while (true) {
        const bool down = (i + 1 < nr);
        if (down) {
            tma_eval_row(T, lane, E, false, y + 2 == ny, K, A.p12, A.p22, B, status);
            refill(i + 1);
        }
        finish_row<4>(P, row, E, down, K, A, B, err, active);
        row += nx; ++i;
        if (++y >= y1) break;
        A = B;
    }
    return err;
}

// ------------------------------------------------------------------------------------------------ two fused iterations
//
// One pass over the strip performs TWO primal-dual iterations (A, then B) and moves the state through HBM once: 60 B per
// pixel for two iterations instead of 120.  A warp covers 128 columns [c0, c0 + 128) but only lanes 1..30 (120 columns)
// own output pixels; lanes 0 and 31 recompute the halo columns that the second iteration's stencils reach into, so the
// horizontal neighbours travel by warp shuffle (no redundant fifth column, no lane-31 special case).  Vertically the strip
// is a software pipeline over the staged rows: with row L newly staged,
//   S1  primal A of row L      u^A(L)   from u^k(L), consts(L), p^k(L), p^k(L-1)            (rows L-1, L still in the ring)
//   S2  dual   A of row L-1    p^A(L-1) from p^k(L-1), u^A(L-1), u^A(L)
//   S3  primal B of row L-1    u^B(L-1) from u^A(L-1), consts(L-1), p^A(L-1), p^A(L-2)
//   S4  dual   B of row L-2    p^B(L-2) from p^A(L-2), u^B(L-2), u^B(L-1)   -> store u^B(L-2), p^B(L-2)
// so a lane carries u^A(L), p^A(L-1) and u^B(L-1) (32 floats) from one row to the next.  A strip of rows [y0, y1) stages
// rows y0-2 .. y1+1 (three halo rows of iteration A, one of iteration B).  Both iterations' residual sums are produced
// (own pixels only), so the reference's stopping rule is still evaluated after EVERY iteration: if iteration A already
// met it, the caller replays ONE single iteration from the untouched input buffers (solver_kernel).  Same arithmetic,
// operation for operation, as the single-iteration path; only the data path differs.

#define F2_OUT 120                       // columns owned by one warp (lanes 1..30)

struct F2Edges {
    bool left, last_own, edge_seg;       // lane's first pixel is column 0 / last pixel is column nx-1 / warp touches a border
};

__device__ __forceinline__ void ld4s(const float *s, float (&d)[4])
{
    const float4 t = *reinterpret_cast<const float4 *>(s);
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
}

// thresholding + primal update of 4 pixels of one row (eval_div + eval_primal of solver_core.h without the fifth column):
// u = flow before this iteration, (gx, gy, rc) = per-warp constants, (a11, a21, b12, b22) = dual variable of the row,
// (l11, l21) = its left neighbours (0 on the first column), (up12, up22) = p12 / p22 of the row above (0 on the first row).
__device__ __forceinline__ void f2_primal(const float (&u1)[4], const float (&u2)[4], const float (&gx)[4], const float (&gy)[4],
                                          const float (&rc)[4], const float (&a11)[4], const float (&a21)[4],
                                          const float (&b12)[4], const float (&b22)[4], float l11, float l21,
                                          const float (&up12)[4], const float (&up22)[4], bool first, bool last,
                                          const F2Edges &E, const IterConsts &K, float (&n1)[4], float (&n2)[4], float (&res)[4])
{
#if defined(__CUDA_ARCH__)      // the straight-line exact operations only exist in the device pass
    float d1[4], d2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool lastcol = (j == 3) && E.last_own;
        const float a1 = lastcol ? 0.f : a11[j], a2 = lastcol ? 0.f : a21[j];
        const float c1 = last ? 0.f : b12[j], c2 = last ? 0.f : b22[j];
        d1[j] = rvdd_div_inner(a1, j ? a11[j ? j - 1 : 0] : l11, c1, up12[j]);
        d2[j] = rvdd_div_inner(a2, j ? a21[j ? j - 1 : 0] : l21, c2, up22[j]);
    }
    if (E.edge_seg && !first && !last) {         // first / last column of a middle row: (s + b) - bu (mask.c:80-81)
        if (E.left) {
            d1[0] = rvdd_div_edge(a11[0], b12[0], up12[0]);
            d2[0] = rvdd_div_edge(a21[0], b22[0], up22[0]);
        }
        if (E.last_own) {
            d1[3] = rvdd_div_edge(-a11[2], b12[3], up12[3]);
            d2[3] = rvdd_div_edge(-a21[2], b22[3], up22[3]);
        }
    }
    unsigned tiny = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 4; j++)
        rvdd_primal_px_fast(u1[j], u2[j], gx[j], gy[j], rvdd_grad2(gx[j], gy[j]), rc[j], d1[j], d2[j], K.l_t, K.theta, K.g0f, &n1[j],
                            &n2[j], tiny);
    if (tiny < RVDD_KEY_2M60) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 n = rvdd_primal_px_slow(u1[j], u2[j], gx[j], gy[j], rvdd_grad2(gx[j], gy[j]), rc[j], d1[j], d2[j], K.l_t,
                                                 K.theta, K.g0f);
            n1[j] = n.x;
            n2[j] = n.y;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) res[j] = rvdd_residual_px(n1[j], u1[j], n2[j], u2[j]);
#endif
}

// dual update of 4 pixels of one row (finish_row of solver_core.h): n = NEW flow of the row, (r1, r2) = its right neighbour
// (lane + 1), (m1, m2) = NEW flow of the row below; p in: old dual variable, out: new.
__device__ __forceinline__ void f2_dual(const float (&n1)[4], const float (&n2)[4], float r1, float r2, const float (&m1)[4],
                                        const float (&m2)[4], bool down, const F2Edges &E, const IterConsts &K, float (&p11)[4],
                                        float (&p12)[4], float (&p21)[4], float (&p22)[4])
{
#if defined(__CUDA_ARCH__)
    float u1x[4], u2x[4], u1y[4], u2y[4], o11[4], o12[4], o21[4], o22[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool lastcol = (j == 3) && E.last_own;
        u1x[j] = lastcol ? 0.f : FSUB(j < 3 ? n1[j < 3 ? j + 1 : 3] : r1, n1[j]);
        u2x[j] = lastcol ? 0.f : FSUB(j < 3 ? n2[j < 3 ? j + 1 : 3] : r2, n2[j]);
        u1y[j] = down ? FSUB(m1[j], n1[j]) : 0.f;
        u2y[j] = down ? FSUB(m2[j], n2[j]) : 0.f;
        o11[j] = p11[j]; o12[j] = p12[j]; o21[j] = p21[j]; o22[j] = p22[j];
    }
    bool bad = false;
    unsigned tiny = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        rvdd_dual_px_fast(&o11[j], &o12[j], u1x[j], u1y[j], K.taut, bad, tiny);
        rvdd_dual_px_fast(&o21[j], &o22[j], u2x[j], u2y[j], K.taut, bad, tiny);
    }
    if (bad || tiny < RVDD_KEY_2M60) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 a = rvdd_dual_px_slow(p11[j], p12[j], u1x[j], u1y[j], K.taut);
            const float2 b = rvdd_dual_px_slow(p21[j], p22[j], u2x[j], u2y[j], K.taut);
            o11[j] = a.x; o12[j] = a.y; o21[j] = b.x; o22[j] = b.y;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) { p11[j] = o11[j]; p12[j] = o12[j]; p21[j] = o21[j]; p22[j] = o22[j]; }
#endif
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity, int *status)
{
    bool ok = mbar_try_wait(bar, parity);
    for (unsigned spins = 0; !ok; ++spins) {
        ok = mbar_try_wait(bar, parity);
        if (!ok && spins > (1u << 22)) {            // ~ seconds: something is badly wrong, do not hang the GPU
            atomicExch(status, 2);
            break;
        }
    }
}

// One warp's strip of the fused pass: columns [c0, c0 + 128) (lanes 1..30 own c0 + 4 .. c0 + 123), output rows [y0, y1).
template <int NST>
__device__ __forceinline__ void iterate2_strip_tma(const SolverArgs &SA, int group, const IterPtrs &P, TmaRing<NST> &T, int lane, int c0,
                                                   int y0, int y1, int nx, int ny, const IterConsts &K, int *status, double &errA,
                                                   double &errB)
{
    const int x0 = c0 + 4 * lane;
    const bool in_img = (x0 >= 0) && (x0 < nx);
    const bool owner = in_img && lane >= 1 && lane <= 30;
    F2Edges E;
    E.left = (x0 == 0);
    E.last_own = (x0 + 4 == nx);
    E.edge_seg = (c0 <= 0) || (c0 + 128 >= nx);
    const int R0 = max(y0 - 2, 0), R1 = min(y1 + 1, ny - 1), Lstart = max(y0 - 1, 0);
    const int nrows = R1 - R0 + 1;
    TmaSrc Q;
    Q.c0 = R0 * nx + c0 - ST_PAD;
    Q.row_c = group * RVDD_NPLANES + RVDD_PL_C; Q.row_u = group * RVDD_NPLANES + RVDD_PL_U + 2 * P.uc;
    Q.row_p = group * RVDD_NPLANES + RVDD_PL_P + 4 * P.pc;

    __syncwarp();
    const unsigned n0 = T.issued;                        // the ring is empty between strips
    if (elect_one()) {
        fence_proxy_async();                             // other CTAs' stores (generic proxy) -> our bulk reads
#pragma unroll
        for (int k = 0; k < NST; k++)
            if (k < nrows) tma_issue_row(SA, Q, k * nx, T, n0 + k);
    }
    T.issued = n0 + (unsigned)min(nrows, NST);

    // carried between rows: u^A of the last staged row, p^A and u^B of the row before it.  The pipeline fills and drains
    // by simply running every stage in every step: what a stage computes before its inputs exist (zeros at first, then
    // finite values of neighbouring rows) is neither stored nor accumulated and never reaches a stage whose output counts.
    float uA1[4], uA2[4], pA11[4], pA12[4], pA21[4], pA22[4], uB1[4], uB2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) uA1[j] = uA2[j] = pA11[j] = pA12[j] = pA21[j] = pA22[j] = uB1[j] = uB2[j] = 0.f;
    const long long colq = x0;                           // float offset of the lane's pixels inside a row

#pragma unroll 1
    for (int L = R0; L <= y1 + 1; ++L) {
        const bool haveL = (L <= R1);
        const unsigned iL = n0 + (unsigned)(L - R0);
        const float *sL = T.stage(iL % NST) + ST_PAD + 4 * lane;                       // row L
        const float *sM = T.stage((iL + NST - 1) % NST) + ST_PAD + 4 * lane;     // row L - 1
        if (haveL) mbar_wait(T.bar(iL % NST), (iL / NST) & 1u, status);
        const int r = L - 1, q = L - 2;

        // ---- S1: primal A of row L (a row below the image recomputes the stage's stale contents: unused, `down` is false)
        float nA1[4], nA2[4];
        {
            float u1[4], u2[4], gx[4], gy[4], rc[4], a11[4], a21[4], b12[4], b22[4], up12[4], up22[4], res[4];
            ld4s(sL + ST_OFF_P, a11); ld4s(sL + ST_OFF_P + 2 * ST_SLOT, a21);
            ld4s(sL + ST_OFF_P + ST_SLOT, b12); ld4s(sL + ST_OFF_P + 3 * ST_SLOT, b22);
            const float l11 = sL[ST_OFF_P - 1], l21 = sL[ST_OFF_P + 2 * ST_SLOT - 1];       // column -1: exactly zero (ZB)
            ld4s(sM + ST_OFF_P + ST_SLOT, up12); ld4s(sM + ST_OFF_P + 3 * ST_SLOT, up22);
            if (L == 0) {
#pragma unroll
                for (int j = 0; j < 4; j++) up12[j] = up22[j] = 0.f;
            }
            ld4s(sL + ST_OFF_U, u1); ld4s(sL + ST_OFF_U + ST_SLOT, u2);
            ld4s(sL + ST_OFF_C, gx); ld4s(sL + ST_OFF_C + ST_SLOT, gy); ld4s(sL + ST_OFF_C + 2 * ST_SLOT, rc);
            f2_primal(u1, u2, gx, gy, rc, a11, a21, b12, b22, l11, l21, up12, up22, L == 0, L == ny - 1, E, K, nA1, nA2, res);
            if (owner && L >= y0 && L < y1) {
#pragma unroll
                for (int j = 0; j < 4; j++) errA += (double)res[j];
            }
        }

        // ---- S2: dual A of row L - 1 (its forward differences need u^A of rows L - 1 and L)
        float qA11[4], qA12[4], qA21[4], qA22[4];
        {
            ld4s(sM + ST_OFF_P, qA11); ld4s(sM + ST_OFF_P + ST_SLOT, qA12);
            ld4s(sM + ST_OFF_P + 2 * ST_SLOT, qA21); ld4s(sM + ST_OFF_P + 3 * ST_SLOT, qA22);
            const float r1 = __shfl_down_sync(0xffffffffu, uA1[0], 1), r2 = __shfl_down_sync(0xffffffffu, uA2[0], 1);
            f2_dual(uA1, uA2, r1, r2, nA1, nA2, haveL, E, K, qA11, qA12, qA21, qA22);
        }

        // ---- S3: primal B of row L - 1
        float nB1[4], nB2[4];
        {
            float gx[4], gy[4], rc[4], up12[4], up22[4], res[4];
            ld4s(sM + ST_OFF_C, gx); ld4s(sM + ST_OFF_C + ST_SLOT, gy); ld4s(sM + ST_OFF_C + 2 * ST_SLOT, rc);
            // left neighbours of p^A: lane - 1's last pixel; on column 0 that is lane 0's own (shfl_up) or a halo lane's value,
            // and must read as the zero the reference's divergence uses there
            const float t11 = __shfl_up_sync(0xffffffffu, qA11[3], 1), t21 = __shfl_up_sync(0xffffffffu, qA21[3], 1);
            const float l11 = E.left ? 0.f : t11, l21 = E.left ? 0.f : t21;
#pragma unroll
            for (int j = 0; j < 4; j++) { up12[j] = r > 0 ? pA12[j] : 0.f; up22[j] = r > 0 ? pA22[j] : 0.f; }
            f2_primal(uA1, uA2, gx, gy, rc, qA11, qA21, qA12, qA22, l11, l21, up12, up22, r == 0, r == ny - 1, E, K, nB1, nB2, res);
            if (owner && r >= y0 && r < y1) {
#pragma unroll
                for (int j = 0; j < 4; j++) errB += (double)res[j];
            }
        }

        // ---- S4: dual B of row L - 2, stores
        {
            const float r1 = __shfl_down_sync(0xffffffffu, uB1[0], 1), r2 = __shfl_down_sync(0xffffffffu, uB2[0], 1);
            f2_dual(uB1, uB2, r1, r2, nB1, nB2, q + 1 <= ny - 1, E, K, pA11, pA12, pA21, pA22);
            if (owner && q >= y0 && q < y1) {
                const long long o = (long long)q * nx + colq;
                Vec<4>::st(P.nu1() + o, uB1);
                Vec<4>::st(P.nu2() + o, uB2);
                Vec<4>::st(P.np11() + o, pA11);
                Vec<4>::st(P.np12() + o, pA12);
                Vec<4>::st(P.np21() + o, pA21);
                Vec<4>::st(P.np22() + o, pA22);
            }
        }

        // ---- rotate the carried rows
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uA1[j] = nA1[j]; uA2[j] = nA2[j];
            pA11[j] = qA11[j]; pA12[j] = qA12[j]; pA21[j] = qA21[j]; pA22[j] = qA22[j];
            uB1[j] = nB1[j]; uB2[j] = nB2[j];
        }

        // ---- the stage of row L - 1 is free now: refill it with row L + 2
        __syncwarp();
        if (L - 1 >= R0 && L + 2 <= R1) {
            if (elect_one()) tma_issue_row(SA, Q, (L + 2 - R0) * nx, T, T.issued);
            T.issued++;
        }
    }
    T.taken = T.issued;
}

// Work split of the fused pass: column segments of F2_OUT pixels (the first one starting one halo lane left of column 0), and
// the ncol * ny segment-rows, taken column after column, dealt out evenly -- every warp gets the same number of rows, in one
// strip or, where its share crosses from one column segment into the next, in two (1280 x 720 on 60 warps: 11 segments x 720
// rows = 132 rows each, instead of 5 strips of 144 rows per segment with 5 of the 60 warps idle).
template <int NST>
__device__ __forceinline__ void iterate2_group(const SolverArgs &A, int group, const IterPtrs &P, TmaRing<NST> &T, int nx, int ny,
                                               const IterConsts &K, int gwarp, int nwarps_group, int *status, double &errA,
                                               double &errB)
{
    const int lane = threadIdx.x & 31;
    const int ncol = (nx + F2_OUT - 1) / F2_OUT;
    const int total = ncol * ny;
    int share = (total + nwarps_group - 1) / nwarps_group;
    if (share < 4) share = 4;                            // tiny levels: a strip of fewer rows is all halo
    int pos = gwarp * share;
    const int end = min(total, pos + share);
    while (pos < end) {
        const int col = pos / ny, y0 = pos - col * ny;
        const int y1 = min(ny, y0 + (end - pos));
        iterate2_strip_tma(A, group, P, T, lane, col * F2_OUT - 4, y0, y1, nx, ny, K, status, errA, errB);
        pos += y1 - y0;
    }
}

// Distribute the image over the group's warps: column segments of 32*V pixels, strips of `rows` rows
// (iterate_strip, the direct-load version, is in solver_core.h and shared with the host-compiled unit tests).
template <int V, int NST>
__device__ __forceinline__ double iterate_group(const SolverArgs &A, int group, const IterPtrs &P, TmaRing<NST> &T, int nx, int ny,
                                                const IterConsts &K, int gwarp, int nwarps_group, int *status)
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
        if (V == 4) {
            err += iterate_strip_tma(A, group, P, T, lane, col * segw, y0, y1, nx, ny, K, status);
        } else {
            if (x0 < nx) err += iterate_strip<V>(P, x0, col * segw, y0, y1, nx, ny, K);
        }
    }
    return err;
}

// ------------------------------------------------------------------------------------------------ warp constants

// The per-warp constants of a whole level (tvl1flow_lib.c:143-159).  The image is cut into tiles of 32 columns x TR
// rows; a warp marches down its tile WC_BATCH rows at a time, so three of the four tap rows of every bicubic window
// were touched by the same warp one trip earlier and come out of L1, and the flow / I0 values of the next trip are
// requested before the double-precision arithmetic of the current one starts.
__device__ __forceinline__ void warp_consts_group(const float *I0, const float *I1, const float *I1x, const float *I1y,
                                                  const float *u1, const float *u2, float *gx, float *gy, float *rc, int nx,
                                                  int ny, int gwarp, int gwarps)
{
    const int lane = threadIdx.x & 31;
    const int ncolt = (nx + 31) >> 5;
    // tile height: the candidate in [16, 32] that needs the fewest row-trips per warp; small levels get one tile per warp
    int TR = ((ny * ncolt + gwarps - 1) / gwarps + WC_BATCH - 1) / WC_BATCH * WC_BATCH;
    if (TR > 16) {
        int best = 1 << 30;
        for (int c = 16; c <= 32; c += WC_BATCH) {
            const int tiles = ncolt * ((ny + c - 1) / c), cost = ((tiles + gwarps - 1) / gwarps) * c;
            if (cost < best) { best = cost; TR = c; }
        }
    }
    const int ntr = (ny + TR - 1) / TR, ntiles = ncolt * ntr;
    for (int t = gwarp; t < ntiles; t += gwarps) {
        const int tyi = t / ncolt, cx = t - tyi * ncolt;
        const int x = cx * 32 + lane, y0 = tyi * TR, y1 = min(ny, y0 + TR);
        const bool lane_on = x < nx;
        float a[WC_BATCH], b[WC_BATCH], i0[WC_BATCH], na[WC_BATCH], nb[WC_BATCH], ni0[WC_BATCH];
        int px[WC_BATCH], py[WC_BATCH];
        bool act[WC_BATCH];
#pragma unroll
        for (int k = 0; k < WC_BATCH; k++) {
            px[k] = x;
            const bool on = lane_on && y0 + k < y1;
            const long long i = on ? (long long)(y0 + k) * nx + x : 0;
            na[k] = u1[i]; nb[k] = u2[i]; ni0[k] = I0[i];
        }
        for (int y = y0; y < y1; y += WC_BATCH) {
#pragma unroll
            for (int k = 0; k < WC_BATCH; k++) {
                a[k] = na[k]; b[k] = nb[k]; i0[k] = ni0[k];
                py[k] = y + k;
                act[k] = lane_on && y + k < y1;
            }
#pragma unroll
            for (int k = 0; k < WC_BATCH; k++) {          // next trip's inputs: in flight during this trip's arithmetic
                const bool on = lane_on && y + WC_BATCH + k < y1;
                const long long i = on ? (long long)(y + WC_BATCH + k) * nx + x : 0;
                na[k] = u1[i]; nb[k] = u2[i]; ni0[k] = I0[i];
            }
            warp_consts_eval<WC_BATCH>(I1, I1x, I1y, a, b, i0, px, py, act, nx, ny, gx, gy, rc);
        }
    }
}

// ------------------------------------------------------------------------------------------------ the kernel

// F2: the instantiation whose big levels may run two iterations per pass (iterate2_strip_tma).  It is a separate kernel, not
// a flag: compiled into one function, the two row loops cost each other 15-20 % (register allocation), and the host picks the
// instantiation per launch (bridge.cu: by the iteration counts of the previous launch).  Both give the same bits.
template <bool F2>
__global__ void __launch_bounds__(SOLVER_THREADS, SOLVER_MIN_CTAS) solver_kernel(const __grid_constant__ SolverArgs A)
{
    constexpr int NST = F2 ? 3 : 2;
    __shared__ double s_red[SOLVER_WARPS], s_red2[SOLVER_WARPS];
    __shared__ double s_val, s_val2;
    __shared__ int s_flag;
    extern __shared__ __align__(128) unsigned char s_dyn[];

    // per-warp staging ring for the bulk-copy row pipeline
    TmaRing<NST> T;
    {
        T.base = s_dyn + (size_t)(threadIdx.x >> 5) * ST_WARP_BYTES(NST);
        T.issued = T.taken = 0u;
#pragma unroll
        for (int k = 0; k < NST; k++)
            if ((threadIdx.x & 31) == 0) mbar_init(T.bar(k), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    const int group = blockIdx.x / A.ctas_per_group;
    GroupCtx g;
    g.nctas = A.ctas_per_group;
    g.cta = blockIdx.x - group * A.ctas_per_group;
    g.bar = A.bar + group * 32;
    g.target = 0u;
    g.partials = A.partials + (size_t)group * 4 * A.ctas_per_group;
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
    float *I1x = S, *I1y = S + PL, *gx = S + RVDD_PL_C * PL, *gy = S + (RVDD_PL_C + 1) * PL, *rc = S + (RVDD_PL_C + 2) * PL;
    // flow and dual variable are double-buffered: plane 5 + 2*buf + comp and 9 + 4*buf + comp (no pointer tables:
    // indexing a local array of pointers would force generic loads and local memory)
#define UB(buf, comp) (S + (RVDD_PL_U + 2 * (buf) + (comp)) * PL)
#define PB(buf, comp) (S + (RVDD_PL_P + 4 * (buf) + (comp)) * PL)
    IterConsts K;
    K.l_t = A.l_t; K.theta = A.theta; K.taut = A.taut; K.g0f = A.g0f;

    // profiling: the group's leader thread accumulates the time a pair spends in the warp-constants phases and in the
    // iteration loops of every level (read back by rvdd_profile_phases)
    const bool stamping = A.scale_ns && g.cta == 0 && threadIdx.x == 0;
    auto now_ns = []() {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    };

    for (int pair = group; pair < A.npairs; pair += A.ngroups) {
        const float *P0 = A.pyr0 + (long long)pair * A.pyr_stride;
        const float *P1 = A.pyr1 + (long long)pair * A.pyr_stride;
        int uc = 0, pc = 0;

        for (int s = A.S - 1; s >= 0; s--) {
            if (A.scale_ns && g.cta == 0 && threadIdx.x == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                A.scale_ns[(long long)pair * (RVDD_MAX_SCALES + 1) + s] = t;
                if (s == 0) A.scale_ns[(long long)pair * (RVDD_MAX_SCALES + 1) + RVDD_MAX_SCALES] = 0ULL;
            }
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

                unsigned long long ns_consts = 0ULL, ns_iter = 0ULL;
                int prev_iters = A.fuse_hint;            // iterations of this level's previous warp (fused-pass policy); for the
                                                         // first warp: the finest-level average of the previous launch
                for (int w = 0; w < A.nwarps; w++) {
                    const unsigned long long tp0 = stamping ? now_ns() : 0ULL;
                    // ---- warp constants (:143-159): bicubic samples of I1, I1x, I1y at x + u
                    const float *u1 = UB(uc, 0), *u2 = UB(uc, 1);
                    warp_consts_group(I0, I1, I1x, I1y, u1, u2, gx, gy, rc, nx, ny, gwarp, gwarps);
                    if (!group_sync(g, &s_flag)) return;
                    const unsigned long long tp1 = stamping ? now_ns() : 0ULL;

                    // ---- inner loop (:161-244), stop test after every iteration.  Big levels run the first iteration alone
                    // (many loops stop right there) and then TWO iterations per pass (iterate2_strip_tma): both residuals
                    // come back, and if the first of the two already met the stopping rule that iteration is replayed
                    // alone from the input buffers, which the fused pass leaves untouched.
                    int it = 0;
                    float err = INFINITY, err_before = INFINITY;
                    // the fused pass needs long strips to amortise its halo rows and its deeper pipeline: measured on 1280x720, it
                    // wins with 132 rows per warp (29 pairs in flight) and loses with 35 (8 pairs)
                    const bool can_fuse = F2 && ((nx & 3) == 0) && n >= A.fuse_min_px &&
                                          ((nx + F2_OUT - 1) / F2_OUT) * ny >= A.fuse_min_rows * gwarps;
                    while (err > A.eps2 && it < RVDD_MAX_ITERATIONS) {
                        IterPtrs P;
                        P.S = S; P.PL = PL; P.uc = uc; P.pc = pc;
                        // Two iterations in one pass only when the loop is not expected to stop after the first of them (a
                        // stop there costs the pass plus a replay): at the start of a loop, if the previous warp of this
                        // level needed at least three iterations; later, if the residual extrapolated with its last decay
                        // ratio stays above eps^2 for one more iteration.  Efficiency only -- every CTA takes the same decision
                        // from the same numbers, and the iteration the loop stops at does not depend on it.
                        bool fused = can_fuse && it >= A.fuse_first && it + 2 <= RVDD_MAX_ITERATIONS;
                        if (fused) fused = (it == 0) ? (prev_iters >= 3) : (err * (err / err_before) > A.eps2);
                        double e = 0.0, e2 = 0.0;
                        if (F2 && fused)
                            iterate2_group(A, group, P, T, nx, ny, K, gwarp, gwarps, A.status, e, e2);
                        else
                            e = ((nx & 3) == 0) ? iterate_group<4>(A, group, P, T, nx, ny, K, gwarp, gwarps, A.status)
                                                : iterate_group<1>(A, group, P, T, nx, ny, K, gwarp, gwarps, A.status);
                        // CTA partials in a fixed order, then the group reduction rides on the barrier
                        for (int o = 16; o > 0; o >>= 1) {
                            e += __shfl_xor_sync(0xffffffffu, e, o);
                            if (fused) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
                        }
                        if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = e; s_red2[threadIdx.x >> 5] = e2; }
                        __syncthreads();
                        if (threadIdx.x == 0) {
                            double t = 0.0, t2 = 0.0;
                            for (int k = 0; k < SOLVER_WARPS; k++) { t += s_red[k]; t2 += s_red2[k]; }
                            g.partials[(size_t)(2 * g.slot) * g.nctas + g.cta] = t;
                            g.partials[(size_t)(2 * g.slot + 1) * g.nctas + g.cta] = t2;
                        }
                        if (!group_sync(g, &s_flag)) return;
                        const double tot = group_sum(g, 2 * g.slot, &s_val);
                        const double tot2 = fused ? group_sum(g, 2 * g.slot + 1, &s_val2) : 0.0;
                        g.slot ^= 1;
                        err_before = err;
                        err = FDIV((float)tot, (float)n);        // error /= size (:223)
                        if (!fused) {
                            it++;
                        } else if (err > A.eps2) {               // the loop continues past iteration A: B is the state
                            it += 2;
                            err_before = err;
                            err = FDIV((float)tot2, (float)n);
                        } else {                                 // the reference stops after iteration A: replay it alone
                            iterate_group<4>(A, group, P, T, nx, ny, K, gwarp, gwarps, A.status);
                            if (!group_sync(g, &s_flag)) return;
                            it++;
                        }
                        uc ^= 1;
                        pc ^= 1;
                    }
                    prev_iters = it;
                    if (g.cta == 0 && threadIdx.x == 0) {
                        const long long t = ((long long)pair * RVDD_MAX_SCALES + s) * A.nwarps + w;
                        if (A.iters_out) A.iters_out[t] = it;
                        if (A.err_out) A.err_out[t] = err;
                        if (s == 0) atomicAdd(A.status + 1, it);     // finest-level iterations of the launch (kernel choice)
                    }
                    if (stamping) {
                        const unsigned long long tp2 = now_ns();
                        ns_consts += tp1 - tp0;
                        ns_iter += tp2 - tp1;
                    }
                }
                if (stamping) {
                    unsigned long long *ph = A.scale_ns + (long long)A.npairs * (RVDD_MAX_SCALES + 1) +
                                             ((long long)pair * RVDD_MAX_SCALES + s) * 2;
                    ph[0] = ns_consts;
                    ph[1] = ns_iter;
                }
            }

            if (s > 0) {
                // ---- zoom_in to the next finer level and rescale (:425-433, zoom.c:85-109)
                const int fx_n = A.nx[s - 1], fy_n = A.ny[s - 1], fn = fx_n * fy_n;
                const float *c1 = UB(uc, 0), *c2 = UB(uc, 1);
                float *f1 = UB(uc ^ 1, 0), *f2 = UB(uc ^ 1, 1);
                const float zx = A.zfx[s - 1], zy = A.zfy[s - 1];
                if (fx_n == 2 * nx && fy_n == 2 * ny) {
                    // exact factor 2: one thread per coarse pixel writes its 2x2 fine block (shared taps, zero fractions)
                    for (int i = gtid; i < n; i += gthreads) {
                        const int Y = i / nx, X = i - Y * nx;
                        float b1[2][2], b2[2][2];
                        zoom_in_2x_block(c1, X, Y, nx, ny, A.zoom_mul, b1);
                        zoom_in_2x_block(c2, X, Y, nx, ny, A.zoom_mul, b2);
                        const long long o = (long long)(2 * Y) * fx_n + 2 * X;
                        *reinterpret_cast<float2 *>(f1 + o) = make_float2(b1[0][0], b1[0][1]);
                        *reinterpret_cast<float2 *>(f1 + o + fx_n) = make_float2(b1[1][0], b1[1][1]);
                        *reinterpret_cast<float2 *>(f2 + o) = make_float2(b2[0][0], b2[0][1]);
                        *reinterpret_cast<float2 *>(f2 + o + fx_n) = make_float2(b2[1][0], b2[1][1]);
                    }
                } else {
                    for (int i = gtid; i < fn; i += gthreads) {
                        const int y = i / fx_n, x = i - y * fx_n;
                        f1[i] = zoom_in_px(c1, x, y, nx, ny, zx, zy, A.zoom_mul);
                        f2[i] = zoom_in_px(c2, x, y, nx, ny, zx, zy, A.zoom_mul);
                    }
                }
                uc ^= 1;
            } else {
                // ---- finest flow -> caller's planar (u, v) buffer (:374-375)
                float *o1 = A.flow_out + (long long)pair * 2 * n, *o2 = o1 + n;
                const float *c1 = UB(uc, 0), *c2 = UB(uc, 1);
                for (int i = gtid; i < n; i += gthreads) { o1[i] = c1[i]; o2[i] = c2[i]; }
                if (A.scale_ns && g.cta == 0 && threadIdx.x == 0) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    A.scale_ns[(long long)pair * (RVDD_MAX_SCALES + 1) + RVDD_MAX_SCALES] = t;
                }
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
    // both instantiations must fit the same grid (they share the workspace layout): take the smaller occupancy
    int occ1 = 0, occ2 = 0;
    e = cudaFuncSetAttribute(solver_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SOLVER_WARPS * ST_WARP_BYTES(2));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(solver_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SOLVER_WARPS * ST_WARP_BYTES(3));
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, solver_kernel<false>, SOLVER_THREADS, SOLVER_WARPS * ST_WARP_BYTES(2));
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, solver_kernel<true>, SOLVER_THREADS, SOLVER_WARPS * ST_WARP_BYTES(3));
    if (e != cudaSuccess) return e;
    *ctas_per_sm = occ1 < occ2 ? occ1 : occ2;
    return cudaSuccess;
}

cudaError_t launch_solver(const SolverArgs &args, bool fused_kernel, cudaStream_t st)
{
    void *params[] = {(void *)&args};
    const dim3 grid(args.ngroups * args.ctas_per_group), block(SOLVER_THREADS);
    if (fused_kernel)
        return cudaLaunchCooperativeKernel((const void *)solver_kernel<true>, grid, block, params, SOLVER_WARPS * ST_WARP_BYTES(3), st);
    return cudaLaunchCooperativeKernel((const void *)solver_kernel<false>, grid, block, params, SOLVER_WARPS * ST_WARP_BYTES(2), st);
}

}  // namespace rvdd
