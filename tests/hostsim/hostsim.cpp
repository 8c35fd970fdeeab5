// hostsim.cpp -- TEST INFRASTRUCTURE.  Compiles the product's host/device math headers (csrc/exact_math.h,
// csrc/solver_core.h) with g++ and drives them with plain loops that mirror the control flow of the CUDA
// kernels (prep.cu tiles -> per-pixel loops, solver.cu warps/lanes -> nested loops over strips and lanes).
// This lets the CPU-only test-suite check, bit for bit against the oracle, the exact arithmetic, the border
// handling and the strip-marching logic of the iteration kernel -- everything except the GPU plumbing itself.
// It is never loaded by the product path.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../rvdd-release_b200/csrc/solver_core.h"

using namespace rvdd;

static void taps_for(double sigma, std::vector<double> &B)
{
    const double den = 2 * sigma * sigma;
    const int size = (int)(5 * sigma) + 1;
    B.resize(size);
    for (int i = 0; i < size; i++) B[i] = 1 / (sigma * sqrt(2.0 * 3.1415926)) * exp(-i * i / den);
    double norm = 0;
    for (int i = 0; i < size; i++) norm += B[i];
    norm *= 2;
    norm -= B[0];
    for (int i = 0; i < size; i++) B[i] /= norm;
}

// what gauss_tile_kernel computes, without the tiling
static void gauss(const float *src, float *dst, int nx, int ny, double sigma, bool norm, float lo, float den)
{
    std::vector<double> B;
    taps_for(sigma, B);
    const int R = (int)B.size() - 1;
    std::vector<float> in((size_t)nx * ny), row((size_t)nx * ny);
    for (size_t i = 0; i < in.size(); i++) in[i] = (norm && den > 0.f) ? rvdd_normalize_px(src[i], lo, den) : src[i];
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
            const float *r = &in[(size_t)y * nx];
            double acc = DMUL(B[0], (double)r[x]);
            for (int j = 1; j <= R; j++)
                acc = DADD(acc, DMUL(B[j], DADD((double)r[rvdd_reflect(x - j, nx)], (double)r[rvdd_reflect(x + j, nx)])));
            row[(size_t)y * nx + x] = (float)acc;
        }
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
            double acc = DMUL(B[0], (double)row[(size_t)y * nx + x]);
            for (int j = 1; j <= R; j++)
                acc = DADD(acc, DMUL(B[j], DADD((double)row[(size_t)rvdd_reflect(y - j, ny) * nx + x],
                                                 (double)row[(size_t)rvdd_reflect(y + j, ny) * nx + x])));
            dst[(size_t)y * nx + x] = (float)acc;
        }
}

extern "C" void hs_gauss(const float *src, float *dst, int nx, int ny, double sigma) { gauss(src, dst, nx, ny, sigma, false, 0, 0); }

extern "C" void hs_resample(const float *src, int nx, int ny, float *dst, int nxx, int nyy, float fx, float fy)
{
    for (int y = 0; y < nyy; y++)
        for (int x = 0; x < nxx; x++)
            dst[(size_t)y * nxx + x] = rvdd_bicubic_clamped(src, FDIV((float)x, fx), FDIV((float)y, fy), nx, ny);
}

template <int V>
static double iterate_image(const IterPtrs &P, int nx, int ny, const IterConsts &K, int nwarps)
{
    const StripPlan sp = plan_strips<V>(nx, ny, nwarps);
    double err = 0.0;
    for (int w = 0; w < sp.total; w++) {
        const int col = w % sp.ncol, strip = w / sp.ncol;
        const int y0 = strip * sp.rows, y1 = (y0 + sp.rows < ny) ? y0 + sp.rows : ny;
        for (int lane = 0; lane < 32; lane++) {
            const int x0 = col * 32 * V + lane * V;
            if (x0 < nx) err += iterate_strip<V>(P, x0, col * 32 * V, y0, y1, nx, ny, K);
        }
    }
    return err;
}

// One call = what the persistent solver does for one pair (solver.cu), given the two pyramids.
// nwarps_group only changes how the image is cut into strips; results must not depend on it.
extern "C" int hs_tvl1flow(const float *I0, const float *I1, float *u, int nx0, int ny0, int nwarps_group, int *iters,
                           int force_scalar)
{
    // parameters and pyramid geometry exactly as bridge.cu derives them
    const float tau = 0.25, lambda = 0.15, theta = 0.3, zfactor = 0.5, epsilon = 0.01;
    const int nwarps = 5;
    int nscales = 100;
    const float N = 1 + log(hypot(nx0, ny0) / 16.0) / log(1 / zfactor);
    if (N < nscales) nscales = N;
    const int S = nscales;
    std::vector<int> nx(S), ny(S);
    nx[0] = nx0; ny[0] = ny0;
    for (int s = 1; s < S; s++) {
        nx[s] = (int)((float)nx[s - 1] * zfactor + 0.5);
        ny[s] = (int)((float)ny[s - 1] * zfactor + 0.5);
    }
    const float l_t = lambda * theta, taut = tau / theta, eps2 = epsilon * epsilon, zoom_mul = (float)1.0 / zfactor;
    const float zsigma = RVDD_ZOOM_SIGMA_ZERO * sqrt(1.0 / (double)(zfactor * zfactor) - 1.0);

    // prep: min/max, normalise + presmooth, zoom_out
    const size_t n0 = (size_t)nx0 * ny0;
    float lo = I0[0], hi = I0[0];
    for (size_t i = 0; i < n0; i++) { lo = fminf(lo, fminf(I0[i], I1[i])); hi = fmaxf(hi, fmaxf(I0[i], I1[i])); }
    const float den = FSUB(hi, lo);
    std::vector<std::vector<float>> A(S), B(S);
    A[0].resize(n0); B[0].resize(n0);
    gauss(I0, A[0].data(), nx0, ny0, RVDD_PRESMOOTH_SIGMA, true, lo, den);
    gauss(I1, B[0].data(), nx0, ny0, RVDD_PRESMOOTH_SIGMA, true, lo, den);
    for (int s = 1; s < S; s++) {
        std::vector<float> tmp((size_t)nx[s - 1] * ny[s - 1]);
        A[s].resize((size_t)nx[s] * ny[s]); B[s].resize((size_t)nx[s] * ny[s]);
        gauss(A[s - 1].data(), tmp.data(), nx[s - 1], ny[s - 1], (double)zsigma, false, 0, 0);
        hs_resample(tmp.data(), nx[s - 1], ny[s - 1], A[s].data(), nx[s], ny[s], zfactor, zfactor);
        gauss(B[s - 1].data(), tmp.data(), nx[s - 1], ny[s - 1], (double)zsigma, false, 0, 0);
        hs_resample(tmp.data(), nx[s - 1], ny[s - 1], B[s].data(), nx[s], ny[s], zfactor, zfactor);
    }

    // solver
    const size_t PL = (n0 + 3) & ~(size_t)3;
    std::vector<float> Sbuf(RVDD_NPLANES * PL, 0.f);
    float *Sp = Sbuf.data();
    float *I1x = Sp, *I1y = Sp + PL, *gx = Sp + RVDD_PL_C * PL, *gy = Sp + (RVDD_PL_C + 1) * PL, *rc = Sp + (RVDD_PL_C + 2) * PL;
    float *ub[2][2] = {{Sp + RVDD_PL_U * PL, Sp + (RVDD_PL_U + 1) * PL}, {Sp + (RVDD_PL_U + 2) * PL, Sp + (RVDD_PL_U + 3) * PL}};
    float *pb[2][4] = {{Sp + RVDD_PL_P * PL, Sp + (RVDD_PL_P + 1) * PL, Sp + (RVDD_PL_P + 2) * PL, Sp + (RVDD_PL_P + 3) * PL},
                       {Sp + (RVDD_PL_P + 4) * PL, Sp + (RVDD_PL_P + 5) * PL, Sp + (RVDD_PL_P + 6) * PL, Sp + (RVDD_PL_P + 7) * PL}};
    int uc = 0, pc = 0;
    for (int s = S - 1; s >= 0; s--) {
        const int w_ = nx[s], h_ = ny[s], n = w_ * h_;
        const float *J0 = A[s].data(), *J1 = B[s].data();
        if (s == S - 1) for (int i = 0; i < n; i++) ub[uc][0][i] = ub[uc][1][i] = 0.f;
        for (int i = 0; i < n; i++) {
            pb[pc][0][i] = pb[pc][1][i] = pb[pc][2][i] = pb[pc][3][i] = 0.f;
            cgrad_px(J1, i % w_, i / w_, w_, h_, &I1x[i], &I1y[i]);
        }
        for (int w = 0; w < nwarps; w++) {
            // the device walks the image with a stride (several pixels per thread); mimic that with a stride of 7
            for (int i = 0; i < n; i += 14) {
                int idx[2] = {i, i + 7 < n ? i + 7 : -1};
                for (int k = 0; k < 7 && idx[0] < n && idx[0] < i + 7; k++) {
                    warp_consts_batch<2>(J0, J1, I1x, I1y, ub[uc][0], ub[uc][1], idx, w_, h_, gx, gy, rc);
                    idx[0]++;
                    idx[1] = (idx[1] >= 0 && idx[1] + 1 < n && idx[1] + 1 < i + 14) ? idx[1] + 1 : -1;
                }
            }
            int it = 0;
            float err = INFINITY;
            while (err > eps2 && it < RVDD_MAX_ITERATIONS) {
                it++;
                IterPtrs P;
                P.S = Sp; P.PL = (long long)PL; P.uc = uc; P.pc = pc;
                IterConsts K;
                K.l_t = l_t; K.theta = theta; K.taut = taut; K.g0f = rvdd_grad_zero_f32();
                const double tot = ((w_ & 3) == 0 && !force_scalar) ? iterate_image<4>(P, w_, h_, K, nwarps_group)
                                                                    : iterate_image<1>(P, w_, h_, K, nwarps_group);
                err = FDIV((float)tot, (float)n);
                uc ^= 1; pc ^= 1;
            }
            if (iters) iters[s * nwarps + w] = it;
        }
        if (s > 0) {
            const int fw = nx[s - 1], fh = ny[s - 1];
            const float zx = ((float)fw / w_), zy = ((float)fh / h_);
            if (fw == 2 * w_ && fh == 2 * h_) {           // the device's exact-factor-2 path: 2x2 blocks with shared taps
                for (int i = 0; i < n; i++)
                    for (int comp = 0; comp < 2; comp++) {
                        float blk[2][2];
                        zoom_in_2x_block(ub[uc][comp], i % w_, i / w_, w_, h_, zoom_mul, blk);
                        for (int j = 0; j < 2; j++)
                            for (int k = 0; k < 2; k++) ub[uc ^ 1][comp][(size_t)(2 * (i / w_) + j) * fw + 2 * (i % w_) + k] = blk[j][k];
                    }
            } else
            for (int i = 0; i < fw * fh; i++) {
                ub[uc ^ 1][0][i] = zoom_in_px(ub[uc][0], i % fw, i / fw, w_, h_, zx, zy, zoom_mul);
                ub[uc ^ 1][1][i] = zoom_in_px(ub[uc][1], i % fw, i / fw, w_, h_, zx, zy, zoom_mul);
            }
            uc ^= 1;
        } else {
            memcpy(u, ub[uc][0], n0 * sizeof(float));
            memcpy(u + n0, ub[uc][1], n0 * sizeof(float));
        }
    }
    return S;
}
