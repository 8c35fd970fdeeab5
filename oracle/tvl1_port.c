/*
 * oracle/tvl1_port.c -- CPU restatement of the reference dual TV-L1 optical-flow path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rvdd-release_b200/, libBridge.so) may include, link,
 * load or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do, and there only as the checker / baseline.
 *
 * Parity pin: this restatement is checked BIT-FOR-BIT against the unmodified reference sources compiled into
 * oracle/_ref/libref_serial.so (oracle/Makefile), stage by stage and end to end (tests/test_oracle_pin.py).
 * The reference itself ships no golden vectors or tests for this path.
 *
 * Every function names the reference lines it follows (paths relative to /root/reference).  The arithmetic
 * types are reproduced exactly: where the reference mixes float and double (bicubic polynomial, Gaussian taps,
 * normalisation, hypot) this file does the same, and the file is built with -ffp-contract=off because the
 * reference objects are built for baseline x86-64 (no FMA).
 *
 * On top of the reference behaviour the port can (a) report the inner-iteration count of every (scale, warp),
 * (b) record both the float-sequential and the wide (double) residual sums of every iteration, and (c) take
 * its stop decision from either of them -- this is how the accumulation-order sensitivity of the stopping rule
 * is measured ("exact stop").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PORT_MAX_ITER 300        /* tvl1flow_lib.c:22 */
#define PORT_PRESMOOTH 0.8       /* tvl1flow_lib.c:23 */
#define PORT_GRAD_ZERO 1E-10     /* tvl1flow_lib.c:24 */
#define PORT_ZOOM_SIGMA0 0.6     /* zoom.c:15 */

typedef struct {
	float tau, lambda, theta;   /* libBridge.cpp:28-30 */
	int nscales, fscale;        /* libBridge.cpp:31-32 */
	float zfactor;              /* libBridge.cpp:33 */
	int nwarps;                 /* libBridge.cpp:34 */
	float epsilon;              /* libBridge.cpp:35 */
	int err_mode;               /* 0: float, pixel order (reference serial build); 1: double sum */
} port_params;

/* optional trace sinks (all may be NULL) */
typedef struct {
	int *iters;        /* [nscales * nwarps], index s * nwarps + w */
	float *last_err;   /* [nscales * nwarps] error value at loop exit */
	float *err_f32;    /* per-iteration float-sequential error, capacity err_cap */
	double *err_f64;   /* per-iteration double residual sum / size, capacity err_cap */
	int err_cap, err_len;
} port_trace;

static void *grab(size_t n)
{
	void *p = malloc(n ? n : 1);
	if (!p) abort();              /* xmalloc.c:13-19 exits on OOM */
	return p;
}

/* ---------------------------------------------------------------- parameters */

void port_default_params(port_params *P)
{
	/* libBridge.cpp:27-36 and the fauxval=-1 selection at :47-57 */
	P->tau = 0.25; P->lambda = 0.15; P->theta = 0.3;
	P->nscales = 100; P->fscale = 0; P->zfactor = 0.5;
	P->nwarps = 5; P->epsilon = 0.01; P->err_mode = 0;
}

/* libBridge.cpp:131-138: N is held in a float, then truncated */
int port_clamp_nscales(int nx, int ny, float zfactor, int nscales)
{
	const float N = 1 + log(hypot(nx, ny) / 16.0) / log(1 / zfactor);
	if (N < nscales) nscales = N;
	return nscales;
}

/* ---------------------------------------------------------------- stencils (mask.c) */

/* mask.c:40-89 -- backward-difference divergence, explicit border/corner formulas */
void port_divergence(const float *a, const float *b, float *out, int nx, int ny)
{
	for (int y = 0; y < ny; y++)
		for (int x = 0; x < nx; x++) {
			const int p = y * nx + x;
			float dx, dy;
			if (x == 0) dx = a[p];
			else if (x == nx - 1) dx = -a[p - 1];
			else dx = a[p] - a[p - 1];
			if (y == 0) dy = b[p];
			else if (y == ny - 1) dy = -b[p - nx];
			else dy = b[p] - b[p - nx];
			/* interior/row/col formulas all evaluate left to right: (v1x) + (v2y) with the
			 * two-term differences formed first only in the interior (mask.c:58-61); on
			 * borders the source writes e.g. v1[p]-v1[p-1]+v2[j] (mask.c:70), which is the
			 * same float sequence ((a-b)+c). First column: v1[p1] + v2[p1] - v2[p1-nx] is
			 * ((a+b)-c) (mask.c:80) -- NOT a+(b-c); handle that ordering explicitly. */
			if (x == 0 && y > 0 && y < ny - 1)
				out[p] = (a[p] + b[p]) - b[p - nx];
			else if (x == nx - 1 && y > 0 && y < ny - 1)
				out[p] = (-a[p - 1] + b[p]) - b[p - nx];
			else
				out[p] = dx + dy;
		}
}

/* mask.c:98-141 -- forward differences, zero on the last column / row */
void port_forward_gradient(const float *f, float *fx, float *fy, int nx, int ny)
{
	for (int y = 0; y < ny; y++)
		for (int x = 0; x < nx; x++) {
			const int p = y * nx + x;
			fx[p] = (x < nx - 1) ? f[p + 1] - f[p] : 0.0f;
			fy[p] = (y < ny - 1) ? f[p + nx] - f[p] : 0.0f;
		}
}

/* mask.c:149-206 -- centred differences; one-sided on the border but still scaled by 0.5 */
void port_centered_gradient(const float *in, float *dx, float *dy, int nx, int ny)
{
	for (int y = 0; y < ny; y++)
		for (int x = 0; x < nx; x++) {
			const int p = y * nx + x;
			const int xl = x > 0 ? p - 1 : p, xr = x < nx - 1 ? p + 1 : p;
			const int yu = y > 0 ? p - nx : p, yd = y < ny - 1 ? p + nx : p;
			dx[p] = 0.5 * (in[xr] - in[xl]);
			dy[p] = 0.5 * (in[yd] - in[yu]);
		}
}

/* mask.c:214-330 -- in-place separable blur; taps in double with pi=3.1415926 (:237); rows first,
 * rounded to float in place (:284), then columns; asymmetric reflection (:264-268, :305-308). */
int port_gauss_radius(double sigma) { return (int)(5 * sigma) + 1; }   /* mask.c:222,225 */

void port_gauss_taps(double sigma, double *B, int size)
{
	const double den = 2 * sigma * sigma;
	for (int i = 0; i < size; i++)
		B[i] = 1 / (sigma * sqrt(2.0 * 3.1415926)) * exp(-i * i / den);
	double norm = 0;
	for (int i = 0; i < size; i++) norm += B[i];
	norm *= 2;
	norm -= B[0];
	for (int i = 0; i < size; i++) B[i] /= norm;
}

static void blur_line(const float *src, int stride, int n, const double *B, int size, double *pad, float *dst)
{
	/* pad[size + i] = src[i]; left mirror about sample 0 (edge not repeated), right mirror with
	 * the edge repeated */
	for (int i = 0; i < n; i++) pad[size + i] = src[(size_t)i * stride];
	for (int i = 0; i < size; i++) {
		pad[i] = src[(size_t)(size - i) * stride];
		pad[size + n + i] = src[(size_t)(n - 1 - i) * stride];
	}
	for (int i = size; i < size + n; i++) {
		double acc = B[0] * pad[i];
		for (int j = 1; j < size; j++) acc += B[j] * (pad[i - j] + pad[i + j]);
		dst[(size_t)(i - size) * stride] = acc;
	}
}

void port_gaussian(float *I, int nx, int ny, double sigma)
{
	const int size = port_gauss_radius(sigma);
	if (size > nx) abort();                      /* mask.c:229-232 */
	double *B = grab(size * sizeof *B);
	port_gauss_taps(sigma, B, size);
	double *pad = grab((size_t)(2 * size + (nx > ny ? nx : ny)) * sizeof *pad);
	for (int y = 0; y < ny; y++) blur_line(I + (size_t)y * nx, 1, nx, B, size, pad, I + (size_t)y * nx);
	for (int x = 0; x < nx; x++) blur_line(I + x, nx, ny, B, size, pad, I + x);
	free(pad);
	free(B);
}

/* ---------------------------------------------------------------- bicubic (bicubic_interpolation.c) */

static int clampi(int v, int n, int *hit)      /* neumann_bc, bicubic_interpolation.c:23-37 */
{
	if (v < 0) { *hit = 1; return 0; }
	if (v >= n) { *hit = 1; return n - 1; }
	return v;
}

static double keys_half(const double v[4], double t)   /* cubic_interpolation_cell, :100-108 */
{
	return v[1] + 0.5 * t * (v[2] - v[0] +
		t * (2.0 * v[0] - 5.0 * v[1] + 4.0 * v[2] - v[3] +
		t * (3.0 * (v[1] - v[2]) + v[3] - v[0])));
}

/* bicubic_interpolation_at, :136-232.  Note :157 -- the "minus" row uses sx, not sy. */
float port_bicubic_at(const float *img, float uu, float vv, int nx, int ny, int border_out)
{
	const int sx = uu < 0 ? -1 : 1, sy = vv < 0 ? -1 : 1;
	int hit = 0;
	const int xi[4] = { clampi((int)uu - sx, nx, &hit), clampi((int)uu, nx, &hit),
	                    clampi((int)uu + sx, nx, &hit), clampi((int)uu + 2 * sx, nx, &hit) };
	const int yi[4] = { clampi((int)vv - sx, ny, &hit), clampi((int)vv, ny, &hit),
	                    clampi((int)vv + sy, ny, &hit), clampi((int)vv + 2 * sy, ny, &hit) };
	if (hit && border_out) return 0.0f;
	const double tx = uu - xi[1], ty = vv - yi[1];     /* float subtraction, then widened (:230) */
	double col[4];
	for (int c = 0; c < 4; c++) {
		double v[4];
		for (int r = 0; r < 4; r++) v[r] = img[xi[c] + (size_t)nx * yi[r]];
		col[c] = keys_half(v, ty);                 /* along y first (:123-126) */
	}
	return keys_half(col, tx);                         /* then along x (:127) */
}

/* bicubic_interpolation_warp, :240-262 */
void port_bicubic_warp(const float *img, const float *u, const float *v, float *out, int nx, int ny, int border_out)
{
#pragma omp parallel for
	for (int y = 0; y < ny; y++)
		for (int x = 0; x < nx; x++) {
			const int p = y * nx + x;
			const float uu = (float)(x + u[p]), vv = (float)(y + v[p]);
			out[p] = port_bicubic_at(img, uu, vv, nx, ny, border_out);
		}
}

/* ---------------------------------------------------------------- pyramid (zoom.c) */

void port_zoom_size(int nx, int ny, int *nxx, int *nyy, float factor)   /* zoom.c:22-34 */
{
	*nxx = (int)((float)nx * factor + 0.5);
	*nyy = (int)((float)ny * factor + 0.5);
}

float port_zoom_sigma(float factor)                                     /* zoom.c:59, kept in float */
{
	const float sigma = PORT_ZOOM_SIGMA0 * sqrt(1.0 / (factor * factor) - 1.0);
	return sigma;
}

void port_zoom_out(const float *I, float *out, int nx, int ny, float factor)   /* zoom.c:41-77 */
{
	float *tmp = grab((size_t)nx * ny * sizeof *tmp);
	memcpy(tmp, I, (size_t)nx * ny * sizeof *tmp);
	int nxx, nyy;
	port_zoom_size(nx, ny, &nxx, &nyy, factor);
	port_gaussian(tmp, nx, ny, port_zoom_sigma(factor));
#pragma omp parallel for
	for (int y = 0; y < nyy; y++)
		for (int x = 0; x < nxx; x++)
			out[y * nxx + x] = port_bicubic_at(tmp, (float)x / factor, (float)y / factor, nx, ny, 0);
	free(tmp);
}

void port_zoom_in(const float *I, float *out, int nx, int ny, int nxx, int nyy)   /* zoom.c:85-109 */
{
	const float fx = ((float)nxx / nx), fy = ((float)nyy / ny);
#pragma omp parallel for
	for (int y = 0; y < nyy; y++)
		for (int x = 0; x < nxx; x++)
			out[y * nxx + x] = port_bicubic_at(I, (float)x / fx, (float)y / fy, nx, ny, 0);
}

/* ---------------------------------------------------------------- normalisation (tvl1flow_lib.c:280-335) */

void port_normalize(const float *a, const float *b, float *an, float *bn, int n)
{
	float lo = a[0], hi = a[0];
	for (int i = 0; i < n; i++) { if (a[i] < lo) lo = a[i]; if (a[i] > hi) hi = a[i]; }
	for (int i = 0; i < n; i++) { if (b[i] < lo) lo = b[i]; if (b[i] > hi) hi = b[i]; }
	const float den = hi - lo;
	if (den > 0)
		for (int i = 0; i < n; i++) {
			an[i] = 255.0 * (a[i] - lo) / den;
			bn[i] = 255.0 * (b[i] - lo) / den;
		}
	else {
		memcpy(an, a, (size_t)n * sizeof *an);
		memcpy(bn, b, (size_t)n * sizeof *bn);
	}
}

/* ---------------------------------------------------------------- one scale (tvl1flow_lib.c:91-273) */

/* One warp's constants (:143-159): warped gradients, |grad|^2 and the constant part of rho. */
void port_warp_constants(const float *I0, const float *I1, const float *I1x, const float *I1y,
                         const float *u1, const float *u2, float *gx, float *gy, float *g2, float *rc,
                         int nx, int ny)
{
	const int n = nx * ny;
	float *w = grab((size_t)n * sizeof *w);
	port_bicubic_warp(I1, u1, u2, w, nx, ny, 1);
	port_bicubic_warp(I1x, u1, u2, gx, nx, ny, 1);
	port_bicubic_warp(I1y, u1, u2, gy, nx, ny, 1);
	for (int i = 0; i < n; i++) {
		const float a = gx[i] * gx[i], b = gy[i] * gy[i];
		g2[i] = (a + b);
		rc[i] = (w[i] - gx[i] * u1[i] - gy[i] * u2[i] - I0[i]);
	}
	free(w);
}

/* One primal-dual iteration (:165-243).  u and p are updated in place exactly as the reference does
 * (v -> div p -> u -> grad u -> p are full-image passes there, so in-place is safe with the scratch
 * arrays below).  Returns the float-sequential error; *wide receives the raw double residual sum. */
float port_iteration(float *u1, float *u2, float *p11, float *p12, float *p21, float *p22,
                     const float *gx, const float *gy, const float *g2, const float *rc,
                     int nx, int ny, float theta, float l_t, float taut, float *scratch, double *wide)
{
	const int n = nx * ny;
	float *d1 = scratch, *d2 = scratch + n, *term = scratch + 2 * (size_t)n;
	port_divergence(p11, p12, d1, nx, ny);      /* :206-207 (p is not touched by the TH step) */
	port_divergence(p21, p22, d2, nx, ny);
#pragma omp parallel for
	for (int i = 0; i < n; i++) {
		/* thresholding step TH (:169-203) */
		const float rho = rc[i] + (gx[i] * u1[i] + gy[i] * u2[i]);
		float e1, e2;
		if (rho < -l_t * g2[i]) { e1 = l_t * gx[i]; e2 = l_t * gy[i]; }
		else if (rho > l_t * g2[i]) { e1 = -l_t * gx[i]; e2 = -l_t * gy[i]; }
		else if (g2[i] < PORT_GRAD_ZERO) e1 = e2 = 0;
		else { const float fi = -rho / g2[i]; e1 = fi * gx[i]; e2 = fi * gy[i]; }
		const float v1 = u1[i] + e1, v2 = u2[i] + e2;
		/* primal update and residual (:212-222) */
		const float o1 = u1[i], o2 = u2[i];
		u1[i] = v1 + theta * d1[i];
		u2[i] = v2 + theta * d2[i];
		term[i] = (u1[i] - o1) * (u1[i] - o1) + (u2[i] - o2) * (u2[i] - o2);
	}
	float ef = 0.0;
	double ed = 0.0;
	for (int i = 0; i < n; i++) { ef += term[i]; ed += term[i]; }
	ef /= n;                                     /* :223 */
	if (wide) *wide = ed;
	/* dual update (:226-243): forward differences of the NEW u, double hypot, 1.0 + float product */
#pragma omp parallel for
	for (int y = 0; y < ny; y++)
		for (int x = 0; x < nx; x++) {
			const int i = y * nx + x;
			const float u1x = x < nx - 1 ? u1[i + 1] - u1[i] : 0.0f, u1y = y < ny - 1 ? u1[i + nx] - u1[i] : 0.0f;
			const float u2x = x < nx - 1 ? u2[i + 1] - u2[i] : 0.0f, u2y = y < ny - 1 ? u2[i + nx] - u2[i] : 0.0f;
			const float g1 = hypot(u1x, u1y), g2n = hypot(u2x, u2y);
			const float n1 = 1.0 + taut * g1, n2 = 1.0 + taut * g2n;
			p11[i] = (p11[i] + taut * u1x) / n1;
			p12[i] = (p12[i] + taut * u1y) / n1;
			p21[i] = (p21[i] + taut * u2x) / n2;
			p22[i] = (p22[i] + taut * u2y) / n2;
		}
	return ef;
}

void port_solve_scale(const float *I0, const float *I1, float *u1, float *u2, int nx, int ny,
                      const port_params *P, int scale, port_trace *T)
{
	const int n = nx * ny;
	const float l_t = P->lambda * P->theta;       /* :107 */
	const float taut = P->tau / P->theta;         /* :233 */
	const float eps2 = P->epsilon * P->epsilon;   /* :163 */
	float *buf = grab((size_t)13 * n * sizeof *buf);
	float *I1x = buf, *I1y = buf + n, *gx = buf + 2 * (size_t)n, *gy = buf + 3 * (size_t)n, *g2 = buf + 4 * (size_t)n,
	      *rc = buf + 5 * (size_t)n, *p11 = buf + 6 * (size_t)n, *p12 = buf + 7 * (size_t)n, *p21 = buf + 8 * (size_t)n,
	      *p22 = buf + 9 * (size_t)n, *scratch = buf + 10 * (size_t)n;
	port_centered_gradient(I1, I1x, I1y, nx, ny);             /* :131 */
	memset(p11, 0, (size_t)4 * n * sizeof *p11);              /* :134-138, once per scale */
	for (int w = 0; w < P->nwarps; w++) {
		port_warp_constants(I0, I1, I1x, I1y, u1, u2, gx, gy, g2, rc, nx, ny);
		int it = 0;
		float err = INFINITY;
		while (err > eps2 && it < PORT_MAX_ITER) {        /* :163 */
			it++;
			double wide;
			const float ef = port_iteration(u1, u2, p11, p12, p21, p22, gx, gy, g2, rc, nx, ny,
			                                P->theta, l_t, taut, scratch, &wide);
			if (T && T->err_f32 && T->err_len < T->err_cap) {
				T->err_f32[T->err_len] = ef;
				if (T->err_f64) T->err_f64[T->err_len] = wide / n;
				T->err_len++;
			}
			/* the wide variant reproduces what the CUDA path does: double sum -> float -> / (float)n */
			err = P->err_mode ? (float)wide / n : ef;
		}
		if (T && T->iters) T->iters[scale * P->nwarps + w] = it;
		if (T && T->last_err) T->last_err[scale * P->nwarps + w] = err;
	}
	free(buf);
}

/* ---------------------------------------------------------------- pyramid driver (tvl1flow_lib.c:343-472) */

/* sizes of every level (tvl1flow_lib.c:387-390) */
void port_pyramid_sizes(int nx, int ny, float zfactor, int nscales, int *nxs, int *nys)
{
	nxs[0] = nx; nys[0] = ny;
	for (int s = 1; s < nscales; s++) port_zoom_size(nxs[s - 1], nys[s - 1], &nxs[s], &nys[s], zfactor);
}

void port_multiscale(const float *I0, const float *I1, float *u1, float *u2, int nx, int ny,
                     const port_params *P, port_trace *T)
{
	const int S = P->nscales, n = nx * ny;
	float **a = grab(S * sizeof *a), **b = grab(S * sizeof *b), **f1 = grab(S * sizeof *f1), **f2 = grab(S * sizeof *f2);
	int *nxs = grab(S * sizeof *nxs), *nys = grab(S * sizeof *nys);
	port_pyramid_sizes(nx, ny, P->zfactor, S, nxs, nys);
	a[0] = grab((size_t)n * sizeof(float));
	b[0] = grab((size_t)n * sizeof(float));
	f1[0] = u1; f2[0] = u2;                                   /* :374-375 */
	port_normalize(I0, I1, a[0], b[0], n);                    /* :380 */
	port_gaussian(a[0], nx, ny, PORT_PRESMOOTH);              /* :383-384 */
	port_gaussian(b[0], nx, ny, PORT_PRESMOOTH);
	for (int s = 1; s < S; s++) {                             /* :387-401 */
		const size_t m = (size_t)nxs[s] * nys[s];
		a[s] = grab(m * sizeof(float)); b[s] = grab(m * sizeof(float));
		f1[s] = grab(m * sizeof(float)); f2[s] = grab(m * sizeof(float));
		port_zoom_out(a[s - 1], a[s], nxs[s - 1], nys[s - 1], P->zfactor);
		port_zoom_out(b[s - 1], b[s], nxs[s - 1], nys[s - 1], P->zfactor);
	}
	for (int i = 0; i < nxs[S - 1] * nys[S - 1]; i++) f1[S - 1][i] = f2[S - 1][i] = 0.0f;   /* :404-405 */
	for (int s = S - 1; s >= 0; s--) {                        /* :408-453 (solve only for s >= fscale) */
		if (s >= P->fscale) port_solve_scale(a[s], b[s], f1[s], f2[s], nxs[s], nys[s], P, s, T);
		if (!s) break;
		port_zoom_in(f1[s], f1[s - 1], nxs[s], nys[s], nxs[s - 1], nys[s - 1]);
		port_zoom_in(f2[s], f2[s - 1], nxs[s], nys[s], nxs[s - 1], nys[s - 1]);
		const float up = (float)1.0 / P->zfactor;         /* :431-432 */
		for (int i = 0; i < nxs[s - 1] * nys[s - 1]; i++) { f1[s - 1][i] *= up; f2[s - 1][i] *= up; }
	}
	for (int s = 1; s < S; s++) { free(a[s]); free(b[s]); free(f1[s]); free(f2[s]); }
	free(a[0]); free(b[0]); free(a); free(b); free(f1); free(f2); free(nxs); free(nys);
}

/* Same contract as the reference bridge symbol (libBridge.cpp:44-163): planar u then v in `u`. */
void port_tvl1flow(const float *I0, const float *I1, float *u, int nx, int ny)
{
	port_params P;
	port_default_params(&P);
	P.nscales = port_clamp_nscales(nx, ny, P.zfactor, P.nscales);
	if (P.nscales < P.fscale) P.fscale = P.nscales;
	port_multiscale(I0, I1, u, u + (size_t)nx * ny, nx, ny, &P, NULL);
}

/* Traced variant used by the tests and the bench: err_mode as in port_params; iters has room for
 * 32 * nwarps entries.  Returns the number of scales. */
int port_tvl1flow_traced(const float *I0, const float *I1, float *u, int nx, int ny, int err_mode,
                         int *iters, float *last_err, float *err_f32, double *err_f64, int err_cap, int *err_len)
{
	port_params P;
	port_default_params(&P);
	P.err_mode = err_mode;
	P.nscales = port_clamp_nscales(nx, ny, P.zfactor, P.nscales);
	port_trace T = { iters, last_err, err_f32, err_f64, err_cap, 0 };
	port_multiscale(I0, I1, u, u + (size_t)nx * ny, nx, ny, &P, &T);
	if (err_len) *err_len = T.err_len;
	return P.nscales;
}
