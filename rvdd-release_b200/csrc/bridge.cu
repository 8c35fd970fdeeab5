// bridge.cu -- the C ABI of libBridge.so (include/rvdd_bridge.h): context / workspace arena, the host-side
// pyramid driver that queues the kernels of prep.cu, solver.cu and warp.cu on a stream, the drop-in
// `tvl1flow` symbol (libBridge.cpp:44) and the host-buffer end-to-end entry point.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rvdd_bridge.h"
#include "internal.h"

using namespace rvdd;

// ------------------------------------------------------------------------------------------------ errors

static thread_local std::string g_err;

static int fail(const char *what, cudaError_t e = cudaSuccess)
{
    g_err = what;
    if (e != cudaSuccess) {
        g_err += ": ";
        g_err += cudaGetErrorString(e);
    }
    return e != cudaSuccess ? (int)e : -1;
}

#define CK(call)                                         \
    do {                                                 \
        cudaError_t e_ = (call);                         \
        if (e_ != cudaSuccess) return fail(#call, e_);   \
    } while (0)

extern "C" const char *rvdd_last_error(void) { return g_err.c_str(); }
extern "C" int rvdd_abi_version(void) { return RVDD_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------ parameters

extern "C" void rvdd_default_params(rvdd_tvl1_params *p)
{
    // libBridge.cpp:27-36
    p->tau = 0.25;
    p->lambda = 0.15;
    p->theta = 0.3;
    p->nscales = 100;
    p->fscale = 0;
    p->zfactor = 0.5;
    p->nwarps = 5;
    p->epsilon = 0.01;
}

// parameter checks of libBridge.cpp:60-123 (out-of-range values fall back to the defaults)
static rvdd_tvl1_params sanitize(const rvdd_tvl1_params *in)
{
    rvdd_tvl1_params d, p;
    rvdd_default_params(&d);
    p = in ? *in : d;
    if (p.tau <= 0 || p.tau > 0.25) p.tau = d.tau;
    if (p.lambda <= 0) p.lambda = d.lambda;
    if (p.theta <= 0) p.theta = d.theta;
    if (p.nscales <= 0) p.nscales = d.nscales;
    if (p.zfactor <= 0 || p.zfactor >= 1) p.zfactor = d.zfactor;
    if (p.nwarps <= 0) p.nwarps = d.nwarps;
    if (p.epsilon <= 0) p.epsilon = d.epsilon;
    return p;
}

// libBridge.cpp:131-138 (same C++ expression, so the same float/double overloads are picked) + zoom.c:22-34
// Returns false when the reference would use more scales than the workspace tables hold (RVDD_MAX_SCALES; only
// reachable with a zoom factor close to 1): the caller reports an error instead of silently computing another pyramid.
static bool build_pyramid(int nx, int ny, rvdd_tvl1_params &p, Pyramid &P)
{
    float zfactor = p.zfactor;
    int nscales = p.nscales;
    const float N = 1 + log(hypot(nx, ny) / 16.0) / log(1 / zfactor);
    if (N < nscales) nscales = N;
    if (nscales < p.fscale) p.fscale = nscales;
    if (nscales > RVDD_MAX_SCALES) return false;
    if (nscales < 1) nscales = 1;
    p.nscales = nscales;
    P.S = nscales;
    P.nx[0] = nx;
    P.ny[0] = ny;
    long long off = 0;
    for (int s = 0; s < nscales; s++) {
        if (s > 0) {
            P.nx[s] = (int)((float)P.nx[s - 1] * zfactor + 0.5);
            P.ny[s] = (int)((float)P.ny[s - 1] * zfactor + 0.5);
        }
        P.off[s] = off;
        off += ((long long)P.nx[s] * P.ny[s] + 3) & ~3LL;
    }
    P.total = off;
    return true;
}

extern "C" int rvdd_pyramid(int nx, int ny, const rvdd_tvl1_params *params, int *nxs, int *nys)
{
    rvdd_tvl1_params p = sanitize(params);
    Pyramid P;
    if (!build_pyramid(nx, ny, p, P)) {
        fail("rvdd_pyramid: more than RVDD_TRACE_SCALES scales (zfactor too close to 1)");
        return -1;
    }
    for (int s = 0; s < P.S; s++) {
        if (nxs) nxs[s] = P.nx[s];
        if (nys) nys[s] = P.ny[s];
    }
    return P.S;
}

// mask.c:224-246
static int make_taps(double sigma, GaussTaps &t)
{
    const double den = 2 * sigma * sigma;
    const int size = (int)(5 * sigma) + 1;
    if (size > RVDD_MAX_TAPS) return -1;
    t.size = size;
    for (int i = 0; i < size; i++) t.B[i] = 1 / (sigma * sqrt(2.0 * 3.1415926)) * exp(-i * i / den);
    double norm = 0;
    for (int i = 0; i < size; i++) norm += t.B[i];
    norm *= 2;
    norm -= t.B[0];
    for (int i = 0; i < size; i++) t.B[i] /= norm;
    return 0;
}

// ------------------------------------------------------------------------------------------------ context

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

#define RING 4

struct rvdd_ctx {
    int device = 0, sms = 0, ctas_per_sm = 0;
    int req_groups = 0;
    DevBuf pyr, tmp, scratch, small, table, stamps;
    // pinned staging ring for the pointer tables
    void *ring_host[RING] = {nullptr, nullptr, nullptr, nullptr};
    size_t ring_cap[RING] = {0, 0, 0, 0};
    cudaEvent_t ring_ev[RING] = {nullptr, nullptr, nullptr, nullptr};
    int ring_next = 0;
    int *status_dev = nullptr;
    Pyramid last_pyr;                   // geometry of the last rvdd_tvl1_flow_dev call (for rvdd_debug_level_dev)
    int last_pairs = 0;
    // end-to-end staging (rvdd_flow_and_warp_host, tvl1flow)
    DevBuf e_frames, e_gray, e_flow, e_hw2, e_warp, e_iters;    // slot 0 (also used by tvl1flow)
    DevBuf f_frames, f_gray, f_flow, f_hw2, f_warp, f_iters;    // slot 1 (pipelined submissions)
    cudaEvent_t slot_in[2] = {nullptr, nullptr}, slot_compute[2] = {nullptr, nullptr}, slot_done[2] = {nullptr, nullptr};
    int *slot_status_host[2] = {nullptr, nullptr};               // pinned: solver watchdog word of the slot's launch
    bool slot_busy[2] = {false, false};
    cudaStream_t st_compute = nullptr, st_in = nullptr, st_out = nullptr;
    std::vector<cudaEvent_t> events;
    // optional timing of the solver launches (rvdd_profile / rvdd_profile_read)
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;   // begin/end pairs
    int prof_n = 0;
    // The TV-L1 workspace (pyramids, solver scratch, barrier words, pointer table) is shared by every entry point: a call
    // queued on another stream than the previous one first waits for that one's kernels (ws_ev).
    cudaEvent_t ws_ev = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_used = false;
    long long spin_limit = 4000000000LL;        // solver watchdog in clock64 ticks (~2 s at 2 GHz); rvdd_set_watchdog
    // Two iterations per pass (solver.cu, iterate2_strip_tma).  fuse_mode: 0 = auto -- the solver instantiation with the fused
    // pass is launched when the previous launch on this context averaged at least 4 inner iterations per warp on the finest
    // level (noisy frames: ~19; clean frames: ~1.2, where the plain instantiation is faster) --, 1 = never, 2 = always.
    // Levels with fewer than fuse_min_px pixels always iterate one at a time.  Both kernels produce the same bits.
    int fuse_mode = 0, fuse_min_px = 600000, fuse_first = 0, fuse_min_rows = 64;
    int *stat_host = nullptr;                   // pinned + mapped: 16 x (finest-level iterations per pair and warp) of the last launch
    int *stat_dev = nullptr;                    // its device alias
    int last_fused = 0;                         // instantiation of the last launch (rvdd_last_solver_fused)
};

static int create_resources(rvdd_ctx *c)
{
    CK(cudaGetDevice(&c->device));
    int coop = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
    if (!coop) return fail("rvdd_create: device lacks cooperative launch");
    const cudaError_t e = solver_max_ctas(&c->ctas_per_sm, &c->sms);
    if (e != cudaSuccess || c->ctas_per_sm < 1) return fail("rvdd_create: solver kernel does not fit on this device", e);
    for (int i = 0; i < RING; i++) CK(cudaEventCreateWithFlags(&c->ring_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        CK(cudaEventCreateWithFlags(&c->slot_in[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->slot_compute[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->slot_done[i], cudaEventDisableTiming));
        CK(cudaMallocHost((void **)&c->slot_status_host[i], sizeof(int)));
        *c->slot_status_host[i] = 0;
    }
    CK(cudaEventCreateWithFlags(&c->ws_ev, cudaEventDisableTiming));
    if (const char *env = getenv("RVDD_WATCHDOG_TICKS")) {
        const long long v = atoll(env);
        if (v > 0) c->spin_limit = v;
    }
    if (const char *env = getenv("RVDD_FUSE")) c->fuse_mode = (env[0] == 'a') ? 0 : (atoi(env) ? 2 : 1);   // auto | 0 | 1
    if (const char *env = getenv("RVDD_FUSE_MIN_PX")) c->fuse_min_px = atoi(env);     // tuning / A-B runs only
    if (const char *env = getenv("RVDD_FUSE_FIRST")) c->fuse_first = atoi(env);
    if (const char *env = getenv("RVDD_FUSE_MIN_ROWS")) c->fuse_min_rows = atoi(env);
    CK(cudaHostAlloc((void **)&c->stat_host, sizeof(int), cudaHostAllocMapped));
    *c->stat_host = 0;
    CK(cudaHostGetDevicePointer((void **)&c->stat_dev, c->stat_host, 0));
    CK(cudaStreamCreateWithFlags(&c->st_compute, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
    return 0;
}

extern "C" int rvdd_destroy(rvdd_ctx *c);

extern "C" int rvdd_create(rvdd_ctx **out)
{
    if (!out) return fail("rvdd_create: null out");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail("rvdd_create: no CUDA device (there is no CPU fallback)", e);
    rvdd_ctx *c = new rvdd_ctx();
    const int rc = create_resources(c);
    if (rc) {
        const std::string why = g_err;          // rvdd_destroy may overwrite the message
        rvdd_destroy(c);                        // releases whatever was created
        g_err = why;
        return rc;
    }
    *out = c;
    return 0;
}

extern "C" int rvdd_destroy(rvdd_ctx *c)
{
    if (!c) return 0;
    cudaDeviceSynchronize();
    c->stamps.release(); c->pyr.release(); c->tmp.release(); c->scratch.release(); c->small.release(); c->table.release();
    c->e_frames.release(); c->e_gray.release(); c->e_flow.release(); c->e_hw2.release(); c->e_warp.release();
    c->e_iters.release();
    c->f_frames.release(); c->f_gray.release(); c->f_flow.release(); c->f_hw2.release(); c->f_warp.release(); c->f_iters.release();
    for (int i = 0; i < 2; i++) {
        if (c->slot_in[i]) cudaEventDestroy(c->slot_in[i]);
        if (c->slot_compute[i]) cudaEventDestroy(c->slot_compute[i]);
        if (c->slot_done[i]) cudaEventDestroy(c->slot_done[i]);
        if (c->slot_status_host[i]) cudaFreeHost(c->slot_status_host[i]);
    }
    for (int i = 0; i < RING; i++) {
        if (c->ring_host[i]) cudaFreeHost(c->ring_host[i]);
        if (c->ring_ev[i]) cudaEventDestroy(c->ring_ev[i]);
    }
    if (c->ws_ev) cudaEventDestroy(c->ws_ev);
    if (c->stat_host) cudaFreeHost(c->stat_host);
    for (cudaEvent_t ev : c->events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->prof_ev) cudaEventDestroy(ev);
    if (c->st_compute) cudaStreamDestroy(c->st_compute);
    if (c->st_in) cudaStreamDestroy(c->st_in);
    if (c->st_out) cudaStreamDestroy(c->st_out);
    delete c;
    return 0;
}

extern "C" int rvdd_set_groups(rvdd_ctx *c, int n)
{
    if (!c) return fail("rvdd_set_groups: null context");
    c->req_groups = n < 0 ? 0 : n;
    return 0;
}

extern "C" int rvdd_set_fuse(rvdd_ctx *c, int mode, int min_px)
{
    if (!c) return fail("rvdd_set_fuse: null context");
    if (mode < 0 || mode > 2) return fail("rvdd_set_fuse: mode must be 0 (auto), 1 (never) or 2 (always)");
    c->fuse_mode = mode;
    if (min_px >= 0) c->fuse_min_px = min_px;
    return 0;
}

extern "C" int rvdd_last_solver_fused(rvdd_ctx *c) { return c ? c->last_fused : -1; }

extern "C" int rvdd_set_watchdog(rvdd_ctx *c, long long ticks)
{
    if (!c) return fail("rvdd_set_watchdog: null context");
    if (ticks <= 0) return fail("rvdd_set_watchdog: ticks must be positive");
    c->spin_limit = ticks;
    return 0;
}

// A solver launch whose watchdog fired unwinds early and leaves the flow buffer partly written: make that visible without
// a host synchronisation by overwriting the whole result with NaN (every consumer then fails loudly instead of using
// stale values).  One tiny launch per solver call; the threads of a healthy launch read one word and return.
__global__ void poison_on_failure_kernel(const int *__restrict__ status, float *__restrict__ flow, long long n, int *stat_host,
                                         int stat_den)
{
    // (also: the launch's finest-level iteration average, x16, goes to the host's mapped word for the next launch's kernel choice)
    if (blockIdx.x == 0 && threadIdx.x == 0) *stat_host = (int)(((long long)status[1] * 16) / stat_den);
    if (*status == 0) return;
    const float nan = __int_as_float(0x7fc00000);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) flow[i] = nan;
}

// ------------------------------------------------------------------------------------------------ TMA descriptors

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D float32 tensor map: dims (d0 innermost, d1), row pitch in bytes, box (b0, b1); out-of-range elements read as zero
namespace rvdd {
cudaError_t encode_map_2d(CUtensorMap *tm, const float *base, unsigned long long d0, unsigned long long d1,
                          unsigned long long pitch_bytes, unsigned b0, unsigned b1)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            return cudaErrorNotSupported;
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
}  // namespace rvdd

// scratch viewed as [nplanes][plane] floats; box = `box_planes` adjacent planes x 136 consecutive floats
static int encode_scratch_map(CUtensorMap *tm, float *scratch, long long plane, long long nplanes, int box_planes)
{
    const cudaError_t e = encode_map_2d(tm, scratch, (unsigned long long)plane, (unsigned long long)nplanes,
                                        (unsigned long long)plane * sizeof(float), 136u, (unsigned)box_planes);
    if (e == cudaErrorNotSupported) return fail("cuTensorMapEncodeTiled not available from this driver");
    if (e != cudaSuccess) return fail("cuTensorMapEncodeTiled failed");
    return 0;
}

// ------------------------------------------------------------------------------------------------ TV-L1 driver

extern "C" int rvdd_tvl1_flow_dev(rvdd_ctx *c, const float *gray, int nframes, int nx, int ny, const int *src,
                                  const int *tgt, int npairs, const rvdd_tvl1_params *params, float *flow,
                                  int *iters, void *stream)
{
    if (!c) return fail("rvdd_tvl1_flow_dev: null context");
    if (npairs <= 0) return 0;
    if (!gray || !src || !tgt || !flow) return fail("rvdd_tvl1_flow_dev: null argument");
    if (nx < 4 || ny < 4 || (long long)nx * ny > (1LL << 30)) return fail("rvdd_tvl1_flow_dev: unsupported image size");
    if (npairs > 32767) return fail("rvdd_tvl1_flow_dev: at most 32767 pairs per call (split the batch)");
    cudaStream_t st = (cudaStream_t)stream;
    rvdd_tvl1_params p = sanitize(params);
    Pyramid P;
    if (!build_pyramid(nx, ny, p, P))
        return fail("rvdd_tvl1_flow_dev: the reference would use more than 16 scales for this zfactor (unsupported)");
    const int S = P.S, K = npairs;
    for (int k = 0; k < K; k++)
        if (src[k] < 0 || src[k] >= nframes || tgt[k] < 0 || tgt[k] >= nframes)
            return fail("rvdd_tvl1_flow_dev: pair index out of range");

    GaussTaps pre, zoom;
    if (make_taps(RVDD_PRESMOOTH_SIGMA, pre)) return fail("presmoothing kernel too wide");
    const float zsigma = RVDD_ZOOM_SIGMA_ZERO * sqrt(1.0 / (double)(p.zfactor * p.zfactor) - 1.0);   // zoom.c:59
    if (S > 1 && make_taps((double)zsigma, zoom)) return fail("zoom kernel too wide (zfactor too small)");
    if (pre.size > nx || pre.size > ny) return fail("image smaller than the presmoothing kernel (mask.c:229)");
    for (int s = 0; s + 1 < S; s++)
        if (zoom.size > P.nx[s] || zoom.size > P.ny[s]) return fail("pyramid level smaller than the zoom kernel (mask.c:229)");

    // ---- the shared workspace: order this call after the previous one if that ran on another stream
    if (c->ws_used && c->ws_stream != st) CK(cudaStreamWaitEvent(st, c->ws_ev, 0));

    // ---- groups and workspace
    const int total_ctas = c->sms * c->ctas_per_sm;
    // one group per pair while a group keeps a minimum of CTAs -- 8 (48 warps) for 1280x720 and larger, 4 / 2 for smaller
    // frames, whose levels have too few row strips to feed more (measured at 640x360: 148 groups of 2 CTAs are 7 %
    // faster than 37 groups of 8); more pairs than that queue up behind the groups
    const long long npx = (long long)nx * ny;
    const int min_ctas = npx >= 600000 ? 8 : (npx >= 300000 ? 4 : 2);
    int G = c->req_groups > 0 ? c->req_groups : (K < total_ctas / min_ctas ? K : total_ctas / min_ctas);
    if (G < 1) G = 1;
    if (G > K) G = K;
    if (G > total_ctas) G = total_ctas;
    const int C = total_ctas / G;
    const long long plane = ((long long)nx * ny + 3) & ~3LL;
    const long long scratch_stride = RVDD_NPLANES * plane;
    CK(c->pyr.ensure(sizeof(float) * (size_t)(2 * K) * P.total));
    CK(c->tmp.ensure(sizeof(float) * (size_t)(2 * K) * plane));
    CK(c->scratch.ensure(sizeof(float) * (size_t)G * scratch_stride));
    // small: [slots 2K ints][status 32 ints][bar G*32 uints][partials G*4*C doubles]
    const size_t off_status = ((size_t)2 * K * sizeof(int) + 255) & ~(size_t)255;
    const size_t off_bar = off_status + 256;
    const size_t nbar = (size_t)G * 32;
    const size_t off_part = (off_bar + nbar * sizeof(unsigned) + 255) & ~(size_t)255;
    CK(c->small.ensure(off_part + sizeof(double) * (size_t)G * 4 * C));
    CK(c->table.ensure(sizeof(void *) * (size_t)2 * K));
    char *small = (char *)c->small.p;
    int *slots = (int *)small;
    int *status = (int *)(small + off_status);
    unsigned *bar = (unsigned *)(small + off_bar);
    double *partials = (double *)(small + off_part);
    c->status_dev = status;
    c->last_pyr = P;
    c->last_pairs = K;

    // ---- pointer table (I0 of every pair, then I1 of every pair) through a pinned ring slot
    const int slot = c->ring_next;
    c->ring_next = (c->ring_next + 1) % RING;
    const size_t tbytes = sizeof(void *) * (size_t)2 * K;
    CK(cudaEventSynchronize(c->ring_ev[slot]));
    if (c->ring_cap[slot] < tbytes) {
        if (c->ring_host[slot]) cudaFreeHost(c->ring_host[slot]);
        c->ring_host[slot] = nullptr;
        c->ring_cap[slot] = 0;
        CK(cudaMallocHost(&c->ring_host[slot], tbytes));
        c->ring_cap[slot] = tbytes;
    }
    const float **tab = (const float **)c->ring_host[slot];
    for (int k = 0; k < K; k++) {
        tab[k] = gray + (long long)tgt[k] * nx * ny;          // I0 = target frame (flow_utils.py:149)
        tab[K + k] = gray + (long long)src[k] * nx * ny;      // I1 = source frame
    }
    const float *const *dtab = (const float *const *)c->table.p;
    CK(cudaMemcpyAsync(c->table.p, tab, tbytes, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(c->ring_ev[slot], st));

    // ---- pyramid: normalise + presmooth (tvl1flow_lib.c:380-384), then zoom_out per level (:387-401)
    float *pyr = (float *)c->pyr.p, *tmp = (float *)c->tmp.p;
    CK(launch_setup(slots, K, bar, (int)nbar, status, st));
    CK(launch_minmax(dtab, dtab + K, nx * ny, K, slots, st));
    CK(launch_gauss(dtab, nullptr, 0, pyr, P.total, nx, ny, 2 * K, pre, slots, K, st));
    for (int s = 1; s < S; s++) {
        if (p.zfactor == 0.5f) {      // the default: blur + pick of the even pixels fused, no full-size temporary
            const cudaError_t e = launch_gauss_decimate(pyr + P.off[s - 1], P.total, pyr + P.off[s], P.total, P.nx[s - 1],
                                                        P.ny[s - 1], P.nx[s], P.ny[s], 2 * K, zoom, st);
            if (e == cudaSuccess) continue;
            if (e != cudaErrorNotSupported) return fail("launch_gauss_decimate", e);
        }
        CK(launch_gauss(nullptr, pyr + P.off[s - 1], P.total, tmp, plane, P.nx[s - 1], P.ny[s - 1], 2 * K, zoom, nullptr, K, st));
        CK(launch_resample(tmp, plane, P.nx[s - 1], P.ny[s - 1], pyr + P.off[s], P.total, P.nx[s], P.ny[s], p.zfactor,
                           p.zfactor, 2 * K, st));
    }

    // ---- persistent solver
    SolverArgs A;
    memset(&A, 0, sizeof A);
    A.npairs = K; A.S = S; A.fscale = p.fscale; A.nwarps = p.nwarps;
    for (int s = 0; s < S; s++) {
        A.nx[s] = P.nx[s]; A.ny[s] = P.ny[s]; A.off[s] = P.off[s];
        if (s + 1 < S) {
            A.zfx[s] = ((float)P.nx[s] / P.nx[s + 1]);      // zoom.c:95-96
            A.zfy[s] = ((float)P.ny[s] / P.ny[s + 1]);
        }
    }
    A.l_t = p.lambda * p.theta;                              // tvl1flow_lib.c:107
    A.theta = p.theta;
    A.taut = p.tau / p.theta;                                // :233
    A.eps2 = p.epsilon * p.epsilon;                          // :163
    A.zoom_mul = (float)1.0 / p.zfactor;                     // :431
    A.g0f = rvdd_grad_zero_f32();
    A.pyr0 = pyr; A.pyr1 = pyr + (long long)K * P.total; A.pyr_stride = P.total;
    A.flow_out = flow;
    A.scratch = (float *)c->scratch.p; A.scratch_stride = scratch_stride; A.plane = plane;
    if (encode_scratch_map(&A.tm4, A.scratch, plane, (long long)RVDD_NPLANES * G, 4)) return -1;
    if (encode_scratch_map(&A.tm3, A.scratch, plane, (long long)RVDD_NPLANES * G, 3)) return -1;
    if (encode_scratch_map(&A.tm2, A.scratch, plane, (long long)RVDD_NPLANES * G, 2)) return -1;
    A.iters_out = iters; A.err_out = nullptr;
    A.scale_ns = nullptr;
    if (c->prof) {
        CK(c->stamps.ensure(sizeof(unsigned long long) * (size_t)K * (RVDD_MAX_SCALES + 1 + 2 * RVDD_MAX_SCALES)));
        CK(cudaMemsetAsync(c->stamps.p, 0, sizeof(unsigned long long) * (size_t)K * (RVDD_MAX_SCALES + 1 + 2 * RVDD_MAX_SCALES), st));
        A.scale_ns = (unsigned long long *)c->stamps.p;
    }
    A.bar = bar; A.partials = partials; A.status = status;
    A.ngroups = G; A.ctas_per_group = C;
    A.spin_limit = c->spin_limit;
    A.fuse_min_px = c->fuse_min_px;
    A.fuse_first = c->fuse_first;
    A.fuse_min_rows = c->fuse_min_rows;
    A.fuse_hint = *(volatile int *)c->stat_host / 16;
    if (iters) CK(cudaMemsetAsync(iters, 0, sizeof(int) * (size_t)K * RVDD_TRACE_SCALES * p.nwarps, st));
    if (c->prof) {
        if ((int)c->prof_ev.size() < 2 * (c->prof_n + 1)) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            c->prof_ev.push_back(a);
            c->prof_ev.push_back(b);
        }
        CK(cudaEventRecord(c->prof_ev[2 * c->prof_n], st));
    }
    // auto: the fused instantiation only if the finest level can actually be fused with this group size (the kernel's own
    // criterion), and the previous launch's inner loops were long
    const bool fusable = (nx & 3) == 0 && (long long)nx * ny >= c->fuse_min_px &&
                         (long long)((nx + 119) / 120) * ny >= (long long)c->fuse_min_rows * C * solver_threads() / 32;
    const bool fused_kernel = c->fuse_mode == 2 || (c->fuse_mode == 0 && fusable && *(volatile int *)c->stat_host >= 4 * 16);
    CK(launch_solver(A, fused_kernel, st));
    c->last_fused = fused_kernel ? 1 : 0;
    if (c->prof) {
        CK(cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], st));
        c->prof_n++;
    }
    poison_on_failure_kernel<<<c->sms, 256, 0, st>>>(status, flow, (long long)K * 2 * nx * ny, c->stat_dev, K * p.nwarps);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ws_ev, st));
    c->ws_stream = st;
    c->ws_used = true;
    return 0;
}

extern "C" int rvdd_profile(rvdd_ctx *c, int enable)
{
    if (!c) return fail("rvdd_profile: null context");
    c->prof = enable != 0;
    c->prof_n = 0;
    return 0;
}

// Per-scale wall time of the last profiled solver launch, averaged over its pairs: ms[s] = time a pair spent at
// pyramid level s (index 0 = finest).  Returns the number of scales written, negative on error.
extern "C" int rvdd_profile_scales(rvdd_ctx *c, float *ms, int cap)
{
    if (!c || !c->stamps.p || c->last_pairs <= 0) return -1;
    const int K = c->last_pairs, S = c->last_pyr.S, W = RVDD_MAX_SCALES + 1;
    std::vector<unsigned long long> h((size_t)K * W);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(h.data(), c->stamps.p, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (int s = 0; s < S && s < cap; s++) {
        double acc = 0;
        for (int k = 0; k < K; k++) {
            const unsigned long long t0 = h[(size_t)k * W + s], t1 = s ? h[(size_t)k * W + s - 1] : h[(size_t)k * W + RVDD_MAX_SCALES];
            acc += (double)(t1 - t0) * 1e-6;
        }
        ms[s] = (float)(acc / K);
    }
    return S < cap ? S : cap;
}

// Same launch, split by phase: ms[2 * s] = time a pair spent in the warp-constants phases of level s (bicubic warps of
// I1 and its gradient, tvl1flow_lib.c:143-159), ms[2 * s + 1] = in its primal-dual iteration loops (:161-244).
extern "C" int rvdd_profile_phases(rvdd_ctx *c, float *ms, int cap)
{
    if (!c || !c->stamps.p || c->last_pairs <= 0) return -1;
    const int K = c->last_pairs, S = c->last_pyr.S;
    std::vector<unsigned long long> h((size_t)K * 2 * RVDD_MAX_SCALES);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(h.data(), (const unsigned long long *)c->stamps.p + (size_t)K * (RVDD_MAX_SCALES + 1),
                   sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    for (int s = 0; s < S && 2 * s + 1 < cap; s++)
        for (int ph = 0; ph < 2; ph++) {
            double acc = 0;
            for (int k = 0; k < K; k++) acc += (double)h[((size_t)k * RVDD_MAX_SCALES + s) * 2 + ph] * 1e-6;
            ms[2 * s + ph] = (float)(acc / K);
        }
    return S;
}

extern "C" int rvdd_profile_read(rvdd_ctx *c, float *solver_ms, int cap)
{
    if (!c) return -1;
    int n = c->prof_n < cap ? c->prof_n : cap;
    for (int i = 0; i < n; i++) {
        if (cudaEventSynchronize(c->prof_ev[2 * i + 1]) != cudaSuccess) return -1;
        if (cudaEventElapsedTime(&solver_ms[i], c->prof_ev[2 * i], c->prof_ev[2 * i + 1]) != cudaSuccess) return -1;
    }
    c->prof_n = 0;
    return n;
}

extern "C" int rvdd_solver_status(rvdd_ctx *c, void *stream)
{
    if (!c) return fail("rvdd_solver_status: null context");
    if (!c->status_dev) return 0;
    int v = 0;
    CK(cudaMemcpyAsync(&v, c->status_dev, sizeof v, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    if (v) fail("solver watchdog fired: a group barrier timed out, results are invalid");
    return v;
}

extern "C" int rvdd_debug_level_dev(rvdd_ctx *c, int pair, int which, int level, float *dst, void *stream)
{
    if (!c || !c->pyr.p) return fail("rvdd_debug_level_dev: no pyramid in the workspace");
    const Pyramid &P = c->last_pyr;
    if (pair < 0 || pair >= c->last_pairs || level < 0 || level >= P.S || (which != 0 && which != 1))
        return fail("rvdd_debug_level_dev: index out of range");
    const float *src = (const float *)c->pyr.p + ((long long)which * c->last_pairs + pair) * P.total + P.off[level];
    CK(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)P.nx[level] * P.ny[level], cudaMemcpyDeviceToDevice,
                       (cudaStream_t)stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------ thin wrappers

extern "C" int rvdd_gray_dev(const float *img, float *gray, int nimg, int h, int w, int ch, void *stream)
{
    if (!img || !gray) return fail("rvdd_gray_dev: null argument");
    if (ch != 1 && ch != 3 && ch != 4) return fail("rvdd_gray_dev: channels must be 1, 3 or 4 (library.py:162-170)");
    CK(launch_gray(img, gray, (long long)nimg * h * w, ch, (cudaStream_t)stream));
    return 0;
}

extern "C" int rvdd_warp_dev(const float *x, const float *flow, float *out, float *mask, int B, int C, int H, int W,
                             long long xs_b, long long xs_c, long long xs_h, long long xs_w, long long os_b,
                             long long os_c, long long os_h, long long os_w, int fh, int fw, float flow_mul, int interp,
                             void *stream)
{
    if (!x || !flow || !out) return fail("rvdd_warp_dev: null argument");
    if (interp != 0 && interp != 1) return fail("rvdd_warp_dev: interp must be 0 (bilinear) or 1 (bicubic)");
    if (!((fh == H && fw == W) || (2 * fh == H && 2 * fw == W))) return fail("rvdd_warp_dev: flow grid must be (H,W) or (H/2,W/2)");
    if (B > 65535) return fail("rvdd_warp_dev: batch too large");
    WarpArgs a;
    a.x = x; a.flow = flow; a.out = out; a.mask = mask;
    a.B = B; a.C = C; a.H = H; a.W = W;
    a.xs_b = xs_b; a.xs_c = xs_c; a.xs_h = xs_h; a.xs_w = xs_w;
    a.os_b = os_b; a.os_c = os_c; a.os_h = os_h; a.os_w = os_w;
    a.fh = fh; a.fw = fw; a.flow_mul = flow_mul; a.interp = interp;
    CK(launch_warp(a, (cudaStream_t)stream));
    return 0;
}

extern "C" int rvdd_upsample2_dev(const float *in, float *out, long long planes, int h, int w, float mul, void *stream)
{
    if (!in || !out) return fail("rvdd_upsample2_dev: null argument");
    CK(launch_upsample2(in, out, planes, h, w, mul, (cudaStream_t)stream));
    return 0;
}

// Bayer pattern string ('grbg', 'rggb', 'gbrg', 'bggr': colour of cell positions (0,0), (0,1), (1,0), (1,1), as in
// Hamilton_Adam_demo.py:201-224) -> positions of the red and blue samples
static int parse_pattern(const char *pattern, int *ry, int *rx, int *by, int *bx)
{
    if (!pattern || strlen(pattern) != 4) return -1;
    int nr = 0, nb = 0, ng = 0;
    for (int k = 0; k < 4; k++) {
        if (pattern[k] == 'r') { *ry = k >> 1; *rx = k & 1; nr++; }
        else if (pattern[k] == 'b') { *by = k >> 1; *bx = k & 1; nb++; }
        else if (pattern[k] == 'g') ng++;
    }
    return (nr == 1 && nb == 1 && ng == 2 && *ry != *by && *rx != *bx) ? 0 : -1;
}

extern "C" int rvdd_demosaic_ha_dev(const float *x, float *y, int B, int H, int W, const char *pattern, void *stream)
{
    if (B == 0) return 0;
    if (!x || !y) return fail("rvdd_demosaic_ha_dev: null argument");
    if (B < 0 || H < 1 || W < 1 || B > 65535) return fail("rvdd_demosaic_ha_dev: bad geometry");
    int ry = 0, rx = 0, by = 0, bx = 0;
    if (parse_pattern(pattern, &ry, &rx, &by, &bx)) return fail("rvdd_demosaic_ha_dev: pattern must be grbg, rggb, gbrg or bggr");
    if (B == 0) return 0;
    CK(launch_demosaic_ha(x, y, B, H, W, ry, rx, by, bx, (cudaStream_t)stream));
    return 0;
}

extern "C" int rvdd_remosaick_gray_dev(const float *rgb, float *gray, int B, int H, int W, const char *pattern, float add,
                                       float mul, void *stream)
{
    if (B == 0) return 0;
    if (!rgb || !gray) return fail("rvdd_remosaick_gray_dev: null argument");
    if (B < 0 || H < 1 || W < 1 || B > 65535 || H > 65535) return fail("rvdd_remosaick_gray_dev: bad geometry");
    int ry = 0, rx = 0, by = 0, bx = 0;
    if (parse_pattern(pattern, &ry, &rx, &by, &bx)) return fail("rvdd_remosaick_gray_dev: pattern must be grbg, rggb, gbrg or bggr");
    if (B == 0) return 0;
    CK(launch_remosaick_gray(rgb, gray, B, H, W, ry, rx, by, bx, add, mul, (cudaStream_t)stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------ host entry points

// planar [2][n] -> chunky [n][2] (library.py:175 returns the (h, w, 2) view; base_dataset.py:180 writes it)
__global__ void interleave_kernel(const float *__restrict__ planar, float2 *__restrict__ hw2, long long n)
{
    const long long k = blockIdx.y;
    const float *p = planar + k * 2 * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        hw2[k * n + i] = make_float2(p[i], p[n + i]);
}

// Three-stage pipeline over two staging slots: uploads on st_in, kernels on st_compute, downloads on st_out, chained
// with events.  While slot s computes, the other slot can already upload its frames and the previous submission of
// slot s^1 can still be downloading -- for a stream of sequences (the offline precompute) the copies disappear behind
// the solver.  A slot must be waited for before it is submitted again.
extern "C" int rvdd_flow_and_warp_host_submit_ex(rvdd_ctx *c, int slot, const float *frames, int nframes, int h, int w, int ch,
                                                 const int *src, const int *tgt, int npairs, const rvdd_tvl1_params *params,
                                                 float *flow_host, float *warped_host, int *iters_host, int warp_mode);

extern "C" int rvdd_flow_and_warp_host_submit(rvdd_ctx *c, int slot, const float *frames, int nframes, int h, int w, int ch,
                                              const int *src, const int *tgt, int npairs, const rvdd_tvl1_params *params,
                                              float *flow_host, float *warped_host, int *iters_host)
{
    return rvdd_flow_and_warp_host_submit_ex(c, slot, frames, nframes, h, w, ch, src, tgt, npairs, params, flow_host, warped_host,
                                             iters_host, warped_host ? RVDD_WARP_DOWNLOAD : RVDD_WARP_SKIP);
}

extern "C" int rvdd_flow_and_warp_host_submit_ex(rvdd_ctx *c, int slot, const float *frames, int nframes, int h, int w, int ch,
                                                 const int *src, const int *tgt, int npairs, const rvdd_tvl1_params *params,
                                                 float *flow_host, float *warped_host, int *iters_host, int warp_mode)
{
    if (warp_mode != RVDD_WARP_SKIP && warp_mode != RVDD_WARP_DOWNLOAD && warp_mode != RVDD_WARP_DISCARD)
        return fail("rvdd_flow_and_warp_host_submit_ex: bad warp_mode");
    if (warp_mode == RVDD_WARP_DOWNLOAD && !warped_host) return fail("rvdd_flow_and_warp_host_submit_ex: warped_host is null");
    const bool do_warp = warp_mode != RVDD_WARP_SKIP, get_warp = warp_mode == RVDD_WARP_DOWNLOAD;
    if (!c) return fail("rvdd_flow_and_warp_host_submit: null context");
    if (slot != 0 && slot != 1) return fail("rvdd_flow_and_warp_host_submit: slot must be 0 or 1");
    if (c->slot_busy[slot]) return fail("rvdd_flow_and_warp_host_submit: slot still in flight (call ..._wait first)");
    if (npairs <= 0) return 0;
    if (!frames || !src || !tgt || !flow_host) return fail("rvdd_flow_and_warp_host_submit: null argument");
    if (ch != 1 && ch != 3 && ch != 4) return fail("rvdd_flow_and_warp_host_submit: channels must be 1, 3 or 4");
    const long long n = (long long)h * w;
    const rvdd_tvl1_params p = sanitize(params);
    DevBuf &b_frames = slot ? c->f_frames : c->e_frames, &b_gray = slot ? c->f_gray : c->e_gray;
    DevBuf &b_flow = slot ? c->f_flow : c->e_flow, &b_hw2 = slot ? c->f_hw2 : c->e_hw2;
    DevBuf &b_warp = slot ? c->f_warp : c->e_warp, &b_iters = slot ? c->f_iters : c->e_iters;
    CK(b_frames.ensure(sizeof(float) * (size_t)nframes * n * ch));
    CK(b_gray.ensure(sizeof(float) * (size_t)nframes * n));
    CK(b_flow.ensure(sizeof(float) * (size_t)npairs * 2 * n));
    CK(b_hw2.ensure(sizeof(float) * (size_t)npairs * 2 * n));
    if (do_warp) CK(b_warp.ensure(sizeof(float) * (size_t)npairs * n * ch));
    if (iters_host) CK(b_iters.ensure(sizeof(int) * (size_t)npairs * RVDD_TRACE_SCALES * p.nwarps));
    float *d_frames = (float *)b_frames.p, *d_gray = (float *)b_gray.p, *d_flow = (float *)b_flow.p;
    float *d_hw2 = (float *)b_hw2.p, *d_warp = (float *)b_warp.p;
    int *d_iters = iters_host ? (int *)b_iters.p : nullptr;
    cudaStream_t st = c->st_compute;

    // stage 1: upload
    CK(cudaMemcpyAsync(d_frames, frames, sizeof(float) * (size_t)nframes * n * ch, cudaMemcpyHostToDevice, c->st_in));
    CK(cudaEventRecord(c->slot_in[slot], c->st_in));
    // stage 2: gray, TV-L1, (h, w, 2) interleave, warp of the source frames
    CK(cudaStreamWaitEvent(st, c->slot_in[slot], 0));
    CK(launch_gray(d_frames, d_gray, (long long)nframes * n, ch, st));
    int rc = rvdd_tvl1_flow_dev(c, d_gray, nframes, w, h, src, tgt, npairs, &p, d_flow, d_iters, st);
    if (rc) return rc;
    {
        unsigned bx = (unsigned)((n + 255) / 256);
        if (bx > 592) bx = 592;
        interleave_kernel<<<dim3(bx, npairs), 256, 0, st>>>(d_flow, (float2 *)d_hw2, n);
        CK(cudaGetLastError());
    }
    if (do_warp) {
        // single_warp(img1 = source frame, flow) in the frames' own HWC layout (flow_utils.py:105-122, :154); runs of
        // consecutive source frames (a video: sources t-1 = 0, 1, 2, ...) go out as one batched launch
        for (int k = 0; k < npairs;) {
            int run = 1;
            while (k + run < npairs && src[k + run] == src[k] + run && run < 65535) run++;
            WarpArgs a;
            a.x = d_frames + (long long)src[k] * n * ch;
            a.flow = d_flow + (long long)k * 2 * n;
            a.out = d_warp + (long long)k * n * ch;
            a.mask = nullptr;
            a.B = run; a.C = ch; a.H = h; a.W = w;
            a.xs_b = a.os_b = n * ch; a.xs_c = a.os_c = 1; a.xs_h = a.os_h = (long long)w * ch; a.xs_w = a.os_w = ch;
            a.fh = h; a.fw = w; a.flow_mul = 1.0f; a.interp = 1;
            CK(launch_warp(a, st));
            k += run;
        }
    }
    CK(cudaMemcpyAsync(c->slot_status_host[slot], c->status_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c->slot_compute[slot], st));
    // stage 3: download
    CK(cudaStreamWaitEvent(c->st_out, c->slot_compute[slot], 0));
    CK(cudaMemcpyAsync(flow_host, d_hw2, sizeof(float) * (size_t)npairs * 2 * n, cudaMemcpyDeviceToHost, c->st_out));
    if (get_warp)
        CK(cudaMemcpyAsync(warped_host, d_warp, sizeof(float) * (size_t)npairs * n * ch, cudaMemcpyDeviceToHost, c->st_out));
    if (iters_host)
        CK(cudaMemcpyAsync(iters_host, d_iters, sizeof(int) * (size_t)npairs * RVDD_TRACE_SCALES * p.nwarps,
                           cudaMemcpyDeviceToHost, c->st_out));
    CK(cudaEventRecord(c->slot_done[slot], c->st_out));
    // the next submission of the OTHER slot may overwrite the shared TV-L1 workspace only after this one's kernels:
    // all kernels run on st_compute, so stream order already guarantees it.
    c->slot_busy[slot] = true;
    return 0;
}

extern "C" int rvdd_flow_and_warp_host_wait(rvdd_ctx *c, int slot)
{
    if (!c) return fail("rvdd_flow_and_warp_host_wait: null context");
    if (slot != 0 && slot != 1) return fail("rvdd_flow_and_warp_host_wait: slot must be 0 or 1");
    if (!c->slot_busy[slot]) return 0;
    CK(cudaEventSynchronize(c->slot_done[slot]));
    c->slot_busy[slot] = false;
    if (*c->slot_status_host[slot]) return fail("solver watchdog fired: a group barrier timed out, results are invalid");
    return 0;
}

extern "C" int rvdd_flow_and_warp_host(rvdd_ctx *c, const float *frames, int nframes, int h, int w, int ch,
                                       const int *src, const int *tgt, int npairs, const rvdd_tvl1_params *params,
                                       float *flow_host, float *warped_host, int *iters_host)
{
    if (!c) return fail("rvdd_flow_and_warp_host: null context");
    int rc = rvdd_flow_and_warp_host_wait(c, 0);
    if (rc) return rc;
    rc = rvdd_flow_and_warp_host_submit(c, 0, frames, nframes, h, w, ch, src, tgt, npairs, params, flow_host, warped_host,
                                        iters_host);
    if (rc) return rc;
    return rvdd_flow_and_warp_host_wait(c, 0);
}

// The reference symbol (libBridge.cpp:44): host float buffers, default parameters, planar (u, v) result.
// One lazily created context per CUDA device: the call runs on whichever device is current in the calling thread.
static std::mutex g_mu;
static std::map<int, rvdd_ctx *> g_ctxs;

extern "C" void tvl1flow(float *I0, float *I1, float *u, int nx, int ny)
{
    std::lock_guard<std::mutex> lock(g_mu);
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        fprintf(stderr, "libBridge(tvl1flow): no CUDA device (there is no CPU fallback)\n");
        return;
    }
    rvdd_ctx *&slot = g_ctxs[dev];
    if (!slot && rvdd_create(&slot)) {
        fprintf(stderr, "libBridge(tvl1flow): %s\n", rvdd_last_error());
        slot = nullptr;
        return;
    }
    rvdd_ctx *c = slot;
    const size_t n = (size_t)nx * ny;
    cudaStream_t st = c->st_compute;
    auto bail = [&](const char *what, cudaError_t e) {
        fail(what, e);
        fprintf(stderr, "libBridge(tvl1flow): %s\n", rvdd_last_error());
    };
    cudaError_t e;
    if ((e = c->e_gray.ensure(sizeof(float) * 2 * n)) != cudaSuccess) return bail("alloc", e);
    if ((e = c->e_flow.ensure(sizeof(float) * 2 * n)) != cudaSuccess) return bail("alloc", e);
    float *d_gray = (float *)c->e_gray.p, *d_flow = (float *)c->e_flow.p;
    if ((e = cudaMemcpyAsync(d_gray, I0, sizeof(float) * n, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D", e);
    if ((e = cudaMemcpyAsync(d_gray + n, I1, sizeof(float) * n, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail("H2D", e);
    const int src = 1, tgt = 0;
    if (rvdd_tvl1_flow_dev(c, d_gray, 2, nx, ny, &src, &tgt, 1, nullptr, d_flow, nullptr, st)) {
        fprintf(stderr, "libBridge(tvl1flow): %s\n", rvdd_last_error());
        return;
    }
    if (rvdd_solver_status(c, st)) {
        fprintf(stderr, "libBridge(tvl1flow): %s\n", rvdd_last_error());
        return;
    }
    if ((e = cudaMemcpyAsync(u, d_flow, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return bail("D2H", e);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return bail("sync", e);
}
