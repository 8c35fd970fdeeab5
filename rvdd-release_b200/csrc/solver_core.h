// solver_core.h -- the per-lane math of the persistent solver (solver.cu), written as host/device code so the
// very same functions can be compiled by g++ and checked bit for bit against the oracle on a machine without
// a GPU (tests/hostsim/).  On the device RVDD_HD is __host__ __device__ __forceinline__.
#pragma once
#include "exact_math.h"

namespace rvdd {

// ------------------------------------------------------------------------------------------------ vectors

template <int V> struct Vec;
template <> struct Vec<4> {
    RVDD_HDM void ld(const float *p, float (&v)[4])
    {
#if defined(__CUDA_ARCH__)
        const float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
#else
        v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
#endif
    }
    RVDD_HDM void st(float *p, const float (&v)[4])
    {
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
#else
        p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; p[3] = v[3];
#endif
    }
};
template <> struct Vec<1> {
    RVDD_HDM void ld(const float *p, float (&v)[1]) { v[0] = *p; }
    RVDD_HDM void st(float *p, const float (&v)[1]) { *p = v[0]; }
};

// All arrays of one iteration are planes of the group's scratch block: planes 0, 1 = centred gradient of I1, planes
// 2..4 = the per-warp constants I1wx, I1wy, rho_c (|grad|^2 is NOT stored: the iteration recomputes it from I1wx, I1wy
// with the reference's own three float operations, tvl1flow_lib.c:155, which is cheaper than reading 4 more bytes per
// pixel per iteration), plane 5 + 2*buf + c = flow component c of buffer buf, plane 9 + 4*buf + c = dual variable.
// Buffer (uc, pc) is read, the other one written.  Addresses are formed on demand from (S, PL) instead of keeping 15
// pointers in registers.
struct IterPtrs {
    float *S;
    long long PL;
    int uc, pc;
    RVDD_HDX const float *gx() const { return S + RVDD_PL_C * PL; }
    RVDD_HDX const float *gy() const { return S + (RVDD_PL_C + 1) * PL; }
    RVDD_HDX const float *rc() const { return S + (RVDD_PL_C + 2) * PL; }
    RVDD_HDX const float *u1() const { return S + (RVDD_PL_U + 2 * uc) * PL; }
    RVDD_HDX const float *u2() const { return S + (RVDD_PL_U + 1 + 2 * uc) * PL; }
    RVDD_HDX const float *p11() const { return S + (RVDD_PL_P + 4 * pc) * PL; }
    RVDD_HDX const float *p12() const { return S + (RVDD_PL_P + 1 + 4 * pc) * PL; }
    RVDD_HDX const float *p21() const { return S + (RVDD_PL_P + 2 + 4 * pc) * PL; }
    RVDD_HDX const float *p22() const { return S + (RVDD_PL_P + 3 + 4 * pc) * PL; }
    RVDD_HDX float *nu1() const { return S + (RVDD_PL_U + 2 * (uc ^ 1)) * PL; }
    RVDD_HDX float *nu2() const { return S + (RVDD_PL_U + 1 + 2 * (uc ^ 1)) * PL; }
    RVDD_HDX float *np11() const { return S + (RVDD_PL_P + 4 * (pc ^ 1)) * PL; }
    RVDD_HDX float *np12() const { return S + (RVDD_PL_P + 1 + 4 * (pc ^ 1)) * PL; }
    RVDD_HDX float *np21() const { return S + (RVDD_PL_P + 2 + 4 * (pc ^ 1)) * PL; }
    RVDD_HDX float *np22() const { return S + (RVDD_PL_P + 3 + 4 * (pc ^ 1)) * PL; }
};

struct IterConsts {
    float l_t, theta, taut, g0f;
};

// What one lane carries for one image row: the NEW flow at its V pixels plus the right neighbour, the OLD dual
// variable, and the residual terms.
template <int V> struct RowState {
    float n1[V + 1], n2[V + 1];
    float p11[V], p21[V];
    float p12[V + 1], p22[V + 1];
    float res[V];
};

// Where a lane's pixels sit relative to the image border; fixed for a whole strip.
struct LaneEdges {
    bool left;       // x0 == 0: the lane's first pixel is the first column
    bool right;      // a right neighbour column x0 + V exists
    bool last_own;   // x0 + V == nx: the lane's last pixel is the last column
    bool last_nb;    // x0 + V == nx - 1: the right neighbour is the last column (only possible when V == 1)
    bool edge_warp;  // some lane of this warp touches the first or last column (warp-uniform)
};

template <int V> RVDD_HD LaneEdges lane_edges(int x0, int nx, int warp_x0)
{
    LaneEdges e;
    e.left = (x0 == 0);
    e.right = (x0 + V < nx);
    e.last_own = (x0 + V == nx);
    e.last_nb = (V == 1) && (x0 + V == nx - 1);
    e.edge_warp = (warp_x0 == 0) || (warp_x0 + 32 * V + 1 >= nx);
    return e;
}

// The inputs of one row for one lane: V own pixels plus the right neighbour (index V) of the flow, the per-warp
// constants and the dual variable, and the left neighbour of p11 / p21.
template <int V> struct RowIn {
    float u1[V + 1], u2[V + 1], gx[V + 1], gy[V + 1], rc[V + 1], a11[V + 1], a21[V + 1];
    float p12[V + 1], p22[V + 1];
    float l11, l21;
};

// Load row `row` (element offset of (y, x0)) straight from global memory.
template <int V> RVDD_HD void load_row_global(const IterPtrs &P, long long row, const LaneEdges &E, RowIn<V> &I)
{
    float t[V];
    Vec<V>::ld(P.u1() + row, t);
#pragma unroll
    for (int j = 0; j < V; j++) I.u1[j] = t[j];
    Vec<V>::ld(P.u2() + row, t);
#pragma unroll
    for (int j = 0; j < V; j++) I.u2[j] = t[j];
    Vec<V>::ld(P.gx() + row, t);
#pragma unroll
    for (int j = 0; j < V; j++) I.gx[j] = t[j];
    Vec<V>::ld(P.gy() + row, t);
#pragma unroll
    for (int j = 0; j < V; j++) I.gy[j] = t[j];
    Vec<V>::ld(P.rc() + row, t);
#pragma unroll
    for (int j = 0; j < V; j++) I.rc[j] = t[j];
    Vec<V>::ld(P.p11() + row, t);
#pragma unroll
    for (int j = 0; j < V; j++) I.a11[j] = t[j];
    Vec<V>::ld(P.p21() + row, t);
#pragma unroll
    for (int j = 0; j < V; j++) I.a21[j] = t[j];
    Vec<V>::ld(P.p12() + row, t);
#pragma unroll
    for (int j = 0; j < V; j++) I.p12[j] = t[j];
    Vec<V>::ld(P.p22() + row, t);
#pragma unroll
    for (int j = 0; j < V; j++) I.p22[j] = t[j];
    I.u1[V] = I.u2[V] = I.gx[V] = I.gy[V] = I.rc[V] = I.a11[V] = I.a21[V] = I.p12[V] = I.p22[V] = 0.f;
    if (E.right) {
        const long long q = row + V;
        I.u1[V] = P.u1()[q]; I.u2[V] = P.u2()[q]; I.gx[V] = P.gx()[q]; I.gy[V] = P.gy()[q]; I.rc[V] = P.rc()[q];
        I.a11[V] = P.p11()[q]; I.a21[V] = P.p21()[q]; I.p12[V] = P.p12()[q]; I.p22[V] = P.p22()[q];
    }
    I.l11 = I.l21 = 0.f;
    if (!E.left) { I.l11 = P.p11()[row - 1]; I.l21 = P.p21()[row - 1]; }
}

// TH + primal update of one loaded row.  up12 / up22 hold p12 / p22 of the row above at the same V+1 columns
// (zeros when y == 0).  first/last: y == 0 / y == ny-1.
// Part 1 needs only the dual variable of the row (a11, a21, p12, p22, l11, l21): divergence of p.
// ZB ("zero border"): the caller relies on the invariant that the dual variable is EXACTLY zero where the divergence would
// zero it -- p11 / p21 on the last column and p12 / p22 on the last row never leave their initial 0, because the forward
// difference they are updated with is forced to 0 there (finish_row) and (0 + taut * 0) / ng = 0 -- so the selects that zero
// those operands are skipped.  Only valid for rows read from the solver's own buffers (the staged path of solver.cu).
template <int V, bool ZB = false>
RVDD_HD void eval_div(const RowIn<V> &I, const LaneEdges &E, bool first, bool last, const float (&up12)[V + 1],
                      const float (&up22)[V + 1], RowState<V> &R, float (&d1)[V + 1], float (&d2)[V + 1])
{
#pragma unroll
    for (int j = 0; j < V; j++) {
        R.p11[j] = I.a11[j];
        R.p21[j] = I.a21[j];
    }
#pragma unroll
    for (int j = 0; j <= V; j++) {
        R.p12[j] = I.p12[j];
        R.p22[j] = I.p22[j];
    }
    // divergence of p: operands zeroed where the stencil leaves the image, see rvdd_div_inner
#pragma unroll
    for (int j = 0; j <= V; j++) {
        const bool lastcol = (j == V - 1) ? E.last_own : ((j == V) ? E.last_nb : false);
        const float a1 = (!ZB && lastcol) ? 0.f : I.a11[j], a2 = (!ZB && lastcol) ? 0.f : I.a21[j];
        const float b1 = (!ZB && last) ? 0.f : I.p12[j], b2 = (!ZB && last) ? 0.f : I.p22[j];
        d1[j] = rvdd_div_inner(a1, j ? I.a11[j ? j - 1 : 0] : I.l11, b1, up12[j]);
        d2[j] = rvdd_div_inner(a2, j ? I.a21[j ? j - 1 : 0] : I.l21, b2, up22[j]);
    }
    if (E.edge_warp && !first && !last) {
        // first / last column of a middle row: the reference associates (s + b) - bu (mask.c:80-81)
        if (E.left) {
            d1[0] = rvdd_div_edge(I.a11[0], I.p12[0], up12[0]);
            d2[0] = rvdd_div_edge(I.a21[0], I.p22[0], up22[0]);
        }
        if (E.last_own) {
            d1[V - 1] = rvdd_div_edge(-(V > 1 ? I.a11[V > 1 ? V - 2 : 0] : I.l11), I.p12[V - 1], up12[V - 1]);
            d2[V - 1] = rvdd_div_edge(-(V > 1 ? I.a21[V > 1 ? V - 2 : 0] : I.l21), I.p22[V - 1], up22[V - 1]);
        }
        if (E.last_nb) {
            d1[V] = rvdd_div_edge(-I.a11[V - 1], I.p12[V], up12[V]);
            d2[V] = rvdd_div_edge(-I.a21[V - 1], I.p22[V], up22[V]);
        }
    }
}

// Part 2 needs the flow and the per-warp constants (u1, u2, gx, gy, rc): thresholding + primal update + residual.
// grad = I1wx^2 + I1wy^2 (tvl1flow_lib.c:155) is formed here from the same two floats the reference squares.
template <int V>
RVDD_HD void eval_primal(const RowIn<V> &I, const IterConsts &K, const float (&d1)[V + 1], const float (&d2)[V + 1],
                         RowState<V> &R)
{
#if defined(__CUDA_ARCH__)
    unsigned tiny = 0xffffffffu;
#pragma unroll
    for (int j = 0; j <= V; j++)
        rvdd_primal_px_fast(I.u1[j], I.u2[j], I.gx[j], I.gy[j], rvdd_grad2(I.gx[j], I.gy[j]), I.rc[j], d1[j], d2[j], K.l_t, K.theta, K.g0f,
                            &R.n1[j], &R.n2[j], tiny);
    if (tiny < RVDD_KEY_2M60) {      // rare: some quotient could not be proven exact -> the reference-exact routine for this row
#pragma unroll
        for (int j = 0; j <= V; j++) {
            const float2 n = rvdd_primal_px_slow(I.u1[j], I.u2[j], I.gx[j], I.gy[j], rvdd_grad2(I.gx[j], I.gy[j]), I.rc[j], d1[j], d2[j],
                                                 K.l_t, K.theta, K.g0f);
            R.n1[j] = n.x;
            R.n2[j] = n.y;
        }
    }
#else
#pragma unroll
    for (int j = 0; j <= V; j++)
        rvdd_primal_px(I.u1[j], I.u2[j], I.gx[j], I.gy[j], rvdd_grad2(I.gx[j], I.gy[j]), I.rc[j], d1[j], d2[j], K.l_t, K.theta, K.g0f,
                       &R.n1[j], &R.n2[j]);
#endif
#pragma unroll
    for (int j = 0; j < V; j++) R.res[j] = rvdd_residual_px(R.n1[j], I.u1[j], R.n2[j], I.u2[j]);
}

template <int V>
RVDD_HD void eval_loaded(const RowIn<V> &I, const LaneEdges &E, bool first, bool last, const IterConsts &K,
                         const float (&up12)[V + 1], const float (&up22)[V + 1], RowState<V> &R)
{
    float d1[V + 1], d2[V + 1];
    eval_div<V>(I, E, first, last, up12, up22, R, d1, d2);
    eval_primal<V>(I, K, d1, d2, R);
}

template <int V>
RVDD_HD void eval_row(const IterPtrs &P, long long row, const LaneEdges &E, bool first, bool last, const IterConsts &K,
                      const float (&up12)[V + 1], const float (&up22)[V + 1], RowState<V> &R)
{
    RowIn<V> I;
    load_row_global<V>(P, row, E, I);
    eval_loaded<V>(I, E, first, last, K, up12, up22, R);
}

// Dual update of row `cur` (its forward differences need the new flow of the row below, `nxt`), stores of the new
// flow and dual variable, residual accumulation.
template <int V>
RVDD_HD void finish_row(const IterPtrs &P, long long row, const LaneEdges &E, bool down, const IterConsts &K,
                        const RowState<V> &cur, const RowState<V> &nxt, double &err, bool active = true)
{
    float o11[V], o12[V], o21[V], o22[V], o1[V], o2[V];
    float u1x[V], u2x[V], u1y[V], u2y[V];
#pragma unroll
    for (int j = 0; j < V; j++) {
        // forward differences of the NEW flow (mask.c:98-141): zero on the last column / row
        const bool lastcol = (j == V - 1) && E.last_own;
        u1x[j] = lastcol ? 0.f : FSUB(cur.n1[j + 1], cur.n1[j]);
        u2x[j] = lastcol ? 0.f : FSUB(cur.n2[j + 1], cur.n2[j]);
        u1y[j] = down ? FSUB(nxt.n1[j], cur.n1[j]) : 0.f;
        u2y[j] = down ? FSUB(nxt.n2[j], cur.n2[j]) : 0.f;
        o11[j] = cur.p11[j]; o12[j] = cur.p12[j]; o21[j] = cur.p21[j]; o22[j] = cur.p22[j];
        o1[j] = cur.n1[j]; o2[j] = cur.n2[j];
        if (active) err += (double)cur.res[j];
    }
#if defined(__CUDA_ARCH__)
    bool bad = false;
    unsigned tiny = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < V; j++) {
        rvdd_dual_px_fast(&o11[j], &o12[j], u1x[j], u1y[j], K.taut, bad, tiny);
        rvdd_dual_px_fast(&o21[j], &o22[j], u2x[j], u2y[j], K.taut, bad, tiny);
    }
    if (bad || tiny < RVDD_KEY_2M60) {      // rare: recompute the row's dual update with the reference-exact routine
#pragma unroll
        for (int j = 0; j < V; j++) {
            const float2 a = rvdd_dual_px_slow(cur.p11[j], cur.p12[j], u1x[j], u1y[j], K.taut);
            const float2 b = rvdd_dual_px_slow(cur.p21[j], cur.p22[j], u2x[j], u2y[j], K.taut);
            o11[j] = a.x; o12[j] = a.y; o21[j] = b.x; o22[j] = b.y;
        }
    }
#else
#pragma unroll
    for (int j = 0; j < V; j++) {
        rvdd_dual_px(&o11[j], &o12[j], u1x[j], u1y[j], K.taut);
        rvdd_dual_px(&o21[j], &o22[j], u2x[j], u2y[j], K.taut);
    }
#endif
    if (!active) return;
    Vec<V>::st(P.nu1() + row, o1);
    Vec<V>::st(P.nu2() + row, o2);
    Vec<V>::st(P.np11() + row, o11);
    Vec<V>::st(P.np12() + row, o12);
    Vec<V>::st(P.np21() + row, o21);
    Vec<V>::st(P.np22() + row, o22);
}

// p12 / p22 of the row above the strip (zeros above the image)
template <int V>
RVDD_HD void load_up_row(const IterPtrs &P, long long row, int nx, int y0, const LaneEdges &E, float (&up12)[V + 1],
                         float (&up22)[V + 1])
{
#pragma unroll
    for (int j = 0; j <= V; j++) up12[j] = up22[j] = 0.f;
    if (y0 > 0) {
        const long long b = row - nx;
        float t[V];
        Vec<V>::ld(P.p12() + b, t);
#pragma unroll
        for (int j = 0; j < V; j++) up12[j] = t[j];
        Vec<V>::ld(P.p22() + b, t);
#pragma unroll
        for (int j = 0; j < V; j++) up22[j] = t[j];
        if (E.right) { up12[V] = P.p12()[b + V]; up22[V] = P.p22()[b + V]; }
    }
}

// One lane's strip: columns x0..x0+V-1, rows y0..y1-1.  Returns the residual sum of those pixels.  The row loop is
// unrolled by two so the "current" and "next" row states swap roles without being copied.
template <int V>
RVDD_HD double iterate_strip(const IterPtrs &P, int x0, int warp_x0, int y0, int y1, int nx, int ny, const IterConsts &K)
{
    double err = 0.0;
    const LaneEdges E = lane_edges<V>(x0, nx, warp_x0);

    float up12[V + 1], up22[V + 1];
    long long row = (long long)y0 * nx + x0;
    load_up_row<V>(P, row, nx, y0, E, up12, up22);
    RowState<V> A, B;
    eval_row<V>(P, row, E, y0 == 0, y0 == ny - 1, K, up12, up22, A);
    int y = y0;
    while (true) {
        bool down = (y + 1 < ny);
        if (down) eval_row<V>(P, row + nx, E, false, y + 2 == ny, K, A.p12, A.p22, B);
        finish_row<V>(P, row, E, down, K, A, B, err);
        row += nx;
        if (++y >= y1) break;
        down = (y + 1 < ny);
        if (down) eval_row<V>(P, row + nx, E, false, y + 2 == ny, K, B.p12, B.p22, A);
        finish_row<V>(P, row, E, down, K, B, A, err);
        row += nx;
        if (++y >= y1) break;
    }
    return err;
}

// ------------------------------------------------------------------------------------------------ per-pixel phases

// centred gradient of I1 at pixel i = (x, y) (tvl1flow_lib.c:131, mask.c:149-206)
RVDD_HD void cgrad_px(const float *I1, int x, int y, int nx, int ny, float *dx, float *dy)
{
    const long long i = (long long)y * nx + x;
    const float c = I1[i];
    const float xl = x > 0 ? I1[i - 1] : c, xr = x < nx - 1 ? I1[i + 1] : c;
    const float yu = y > 0 ? I1[i - nx] : c, yd = y < ny - 1 ? I1[i + nx] : c;
    *dx = rvdd_half_diff(xr, xl);
    *dy = rvdd_half_diff(yd, yu);
}

// per-warp constants (tvl1flow_lib.c:143-159) for U pixels at once: bicubic samples of I1, I1x, I1y at (x+u1, y+u2) with
// border_out = true, then the constant part of rho (|grad|^2, :155, is recomputed by the iteration: rvdd_grad2).
// The phase is a gather whose addresses depend on a load (the flow), so it is latency-bound unless many loads are in
// flight: the routine is branch-free -- a sample that leaves the image (bicubic_interpolation.c:195-196 returns 0)
// reads element 0 sixteen times through zero strides and is zeroed by a select -- so that the compiler can issue the
// flow loads of all U pixels, then the 16 * U taps of each image, before the double-precision arithmetic starts.
// idx[k] < 0 marks an unused slot.
template <int U>
RVDD_HD void warp_consts_eval(const float *I1, const float *I1x, const float *I1y, const float (&a)[U], const float (&b)[U],
                              const float (&i0)[U], const int (&px)[U], const int (&py)[U], const bool (&act)[U], int nx,
                              int ny, float *gx, float *gy, float *rc)
{
    float tx[U], ty[U];
    long long o[U];
    int rs[U], cs[U];
    bool in[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
        const float uu = FADD((float)px[k], a[k]), vv = FADD((float)py[k], b[k]);
        in[k] = act[k] && rvdd_inside_strict(uu, vv, nx, ny);
        const int bx = in[k] ? (int)uu : 1, by = in[k] ? (int)vv : 1;
        tx[k] = FSUB(uu, (float)bx);
        ty[k] = FSUB(vv, (float)by);
        rs[k] = in[k] ? nx : 0;
        cs[k] = in[k] ? 1 : 0;
        o[k] = in[k] ? (long long)(by - 1) * nx + (bx - 1) : 0;
    }
    float w0[U], wx[U], wy[U];
#pragma unroll
    for (int img = 0; img < 3; img++) {
        const float *src = img == 0 ? I1 : (img == 1 ? I1x : I1y);
        float v[U][4][4];
#pragma unroll
        for (int k = 0; k < U; k++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) v[k][c][r] = src[o[k] + r * rs[k] + c * cs[k]];
#pragma unroll
        for (int k = 0; k < U; k++) {
            const float w = in[k] ? rvdd_bicubic_cell(v[k], tx[k], ty[k]) : 0.f;
            if (img == 0) w0[k] = w;
            else if (img == 1) wx[k] = w;
            else wy[k] = w;
        }
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
        if (!act[k]) continue;
        const long long i = (long long)py[k] * nx + px[k];
        gx[i] = wx[k];
        gy[i] = wy[k];
        rc[i] = FSUB(FSUB(FSUB(w0[k], FMUL(wx[k], a[k])), FMUL(wy[k], b[k])), i0[k]);
    }
}

// The same for U linear pixel indices (idx[k] < 0 marks an unused slot), loads included.
template <int U>
RVDD_HD void warp_consts_batch(const float *I0, const float *I1, const float *I1x, const float *I1y, const float *u1,
                               const float *u2, const int (&idx)[U], int nx, int ny, float *gx, float *gy, float *rc)
{
    float a[U], b[U], i0[U];
    int px[U], py[U];
    bool act[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
        act[k] = idx[k] >= 0;
        const int i = act[k] ? idx[k] : 0;
        a[k] = u1[i];
        b[k] = u2[i];
        i0[k] = I0[i];
        py[k] = i / nx;
        px[k] = i - py[k] * nx;
    }
    warp_consts_eval<U>(I1, I1x, I1y, a, b, i0, px, py, act, nx, ny, gx, gy, rc);
}

// single-pixel form (host-side tests)
RVDD_HD void warp_consts_px(const float *I0, const float *I1, const float *I1x, const float *I1y, const float *u1,
                            const float *u2, int i, int nx, int ny, float *gx, float *gy, float *rc)
{
    const int idx[1] = {i};
    warp_consts_batch<1>(I0, I1, I1x, I1y, u1, u2, idx, nx, ny, gx, gy, rc);
}

// flow upsampling to the next finer level at fine pixel (x, y) (zoom.c:85-109 + tvl1flow_lib.c:431-432)
RVDD_HD float zoom_in_px(const float *coarse, int x, int y, int nx, int ny, float zx, float zy, float mul)
{
    return FMUL(rvdd_bicubic_clamped(coarse, FDIV((float)x, zx), FDIV((float)y, zy), nx, ny), mul);
}

// The same upsampling when the fine level is EXACTLY twice the coarse one in both directions (every level of an even-sized
// pyramid, e.g. 1280x720 -> 640x360 -> 320x180): the sample positions x / 2.0f are integers or integers + 0.5, and the
// four fine pixels (2X + i, 2Y + j) of coarse pixel (X, Y) share their 16 taps.  With a zero fraction the Keys cubic
// returns its second sample (v1 + 0.5 * 0 * (...) = v1), so per block only six of the twenty cubics are evaluated
// (108 double operations instead of 360) and the taps are loaded and converted once.  Identical values (up to the sign
// of a zero).  out[j][i] = fine pixel (2X + i, 2Y + j).
RVDD_HD void zoom_in_2x_block(const float *coarse, int X, int Y, int nx, int ny, float mul, float (&out)[2][2])
{
    const int xi[4] = {rvdd_clampi(X - 1, nx), X, rvdd_clampi(X + 1, nx), rvdd_clampi(X + 2, nx)};
    const int yi[4] = {rvdd_clampi(Y - 1, ny), Y, rvdd_clampi(Y + 1, ny), rvdd_clampi(Y + 2, ny)};
    double d[4][4];                                        // d[column][row]
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int r = 0; r < 4; r++) d[c][r] = (double)coarse[xi[c] + (long long)nx * yi[r]];
    double colh[4];                                        // columns interpolated at ty = 0.5
#pragma unroll
    for (int c = 0; c < 4; c++) colh[c] = rvdd_keys_half(d[c][0], d[c][1], d[c][2], d[c][3], 0.5);
    out[0][0] = FMUL((float)d[1][1], mul);
    out[0][1] = FMUL((float)rvdd_keys_half(d[0][1], d[1][1], d[2][1], d[3][1], 0.5), mul);
    out[1][0] = FMUL((float)colh[1], mul);
    out[1][1] = FMUL((float)rvdd_keys_half(colh[0], colh[1], colh[2], colh[3], 0.5), mul);
}

// strip decomposition of an nx*ny image over `nwarps` warps with V pixels per lane: column segments of 32*V
// pixels, `rows` rows per strip.
struct StripPlan {
    int ncol, rows, nstrips, total;
};
template <int V> RVDD_HD StripPlan plan_strips(int nx, int ny, int nwarps)
{
    StripPlan p;
    const int segw = 32 * V;
    p.ncol = (nx + segw - 1) / segw;
    int per_col = nwarps / p.ncol;
    if (per_col < 1) per_col = 1;
    p.rows = (ny + per_col - 1) / per_col;
    if (p.rows < 1) p.rows = 1;
    p.nstrips = (ny + p.rows - 1) / p.rows;
    p.total = p.ncol * p.nstrips;
    return p;
}

}  // namespace rvdd
