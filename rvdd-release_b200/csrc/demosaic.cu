// demosaic.cu -- Hamilton-Adams demosaicking of packed Bayer raw, the step right before the warp in the inference
// loop (models/recurrent_model.py:126 -> util/Hamilton_Adam_demo.py:249-289), and its inverse packing `remosaick`
// fused with the mean-of-4 gray conversion for the online-flow path (validate.py:29-33).
//
// The reference expresses the algorithm as three fixed-weight convolutions with replication padding plus ~40
// element-wise tensor ops over masks (about 45 kernel launches and 30 full-resolution temporaries per frame).  Here
// it is ONE kernel: a CTA stages the CFA tile (+3 pixel halo, coordinates clamped = ReplicationPad2d) in shared
// memory, interpolates green on the tile (+1 halo, algorithm 1), then red and blue from the colour differences
// (algorithm 2), and writes the three output planes.  HBM traffic is the minimum: 4 B read + 12 B written per
// full-resolution pixel.
//
// Every arithmetic expression keeps the reference's operand order (conv taps accumulate in kernel-memory order, the
// 0.5 / 0.25 / 2 weights are exact scalings), so results agree with the torch CPU reference to the last bit or two.
#include "internal.h"

namespace rvdd {

#define DM_TW 64                      // output tile width
#define DM_TH 16                      // output tile height
#define DM_PW (DM_TW + 6)             // CFA tile with a 3 pixel halo
#define DM_PH (DM_TH + 6)
#define DM_GW (DM_TW + 2)             // green tile with a 1 pixel halo
#define DM_GH (DM_TH + 2)

struct DemosaicArgs {
    const float *x;                   // [B][4][H][W] packed raw
    float *y;                         // [B][3][2H][2W]
    int B, H, W;
    int ry, rx, by, bx;               // position of the red / blue sample inside the 2x2 cell
};

__device__ __forceinline__ float sgn(float v) { return (float)((v > 0.f) - (v < 0.f)); }     // torch.sign

__global__ void __launch_bounds__(256) demosaic_ha_kernel(const DemosaicArgs a)
{
    __shared__ float P[DM_PH][DM_PW + 1];
    __shared__ float G[DM_GH][DM_GW + 1];
    const int H2 = 2 * a.H, W2 = 2 * a.W;
    const int X0 = blockIdx.x * DM_TW, Y0 = blockIdx.y * DM_TH;
    const float *xb = a.x + (long long)blockIdx.z * 4 * a.H * a.W;
    const long long plane = (long long)a.H * a.W;

    // CFA tile: pack_in_one (Hamilton_Adam_demo.py:226-234), clamped coordinates = ReplicationPad2d
    for (int i = threadIdx.x; i < DM_PH * DM_PW; i += 256) {
        const int ty = i / DM_PW, tx = i - ty * DM_PW;
        const int yy = min(max(Y0 + ty - 3, 0), H2 - 1), xx = min(max(X0 + tx - 3, 0), W2 - 1);
        P[ty][tx] = xb[(long long)((yy & 1) * 2 + (xx & 1)) * plane + (long long)(yy >> 1) * a.W + (xx >> 1)];
    }
    __syncthreads();

    // green on the tile + 1 pixel halo (algorithm 1, :123-142).  The halo positions are clamped to the image, which is
    // what the replication padding in front of conv_algo2_green sees.
    for (int i = threadIdx.x; i < DM_GH * DM_GW; i += 256) {
        const int gy = i / DM_GW, gx = i - gy * DM_GW;
        const int yy = min(max(Y0 + gy - 1, 0), H2 - 1), xx = min(max(X0 + gx - 1, 0), W2 - 1);
        // neighbours of (yy, xx) in the CFA, clamped (ReplicationPad2d((2, 2, 2, 2)))
        auto cfa = [&](int dy, int dx) {
            const int y2 = min(max(yy + dy, 0), H2 - 1), x2 = min(max(xx + dx, 0), W2 - 1);
            return P[y2 - Y0 + 3][x2 - X0 + 3];
        };
        const float c = cfa(0, 0);
        const bool is_r = ((yy & 1) == a.ry) && ((xx & 1) == a.rx), is_b = ((yy & 1) == a.by) && ((xx & 1) == a.bx);
        float g = c;
        if (is_r || is_b) {
            const float l1 = cfa(0, -1), r1 = cfa(0, 1), u1 = cfa(-1, 0), d1 = cfa(1, 0);
            const float kh = 0.5f * l1 + 0.5f * r1, kv = 0.5f * u1 + 0.5f * d1;
            const float dh = (cfa(0, -2) + -2.f * c) + cfa(0, 2), dv = (cfa(-2, 0) + -2.f * c) + cfa(2, 0);
            const float rawh = kh - dh / 4.f, rawv = kv - dv / 4.f;
            const float clh = fabsf(l1 - r1) + fabsf(dh), clv = fabsf(u1 - d1) + fabsf(dv);
            const float s = sgn(clh - clv);
            g = (1.f + s) * rawv / 2.f + (1.f - s) * rawh / 2.f;
        }
        G[gy][gx] = g;
    }
    __syncthreads();

    // red and blue (algorithm 2, :145-172) and the three output planes
    float *yb = a.y + (long long)blockIdx.z * 3 * H2 * W2;
    for (int i = threadIdx.x; i < DM_TH * DM_TW; i += 256) {
        const int ty = i / DM_TW, tx = i - ty * DM_TW;
        const int yy = Y0 + ty, xx = X0 + tx;
        if (yy >= H2 || xx >= W2) continue;
        const float g = G[ty + 1][tx + 1];
        // green second differences (conv_algo2_green): clamped neighbours are already in the halo
        auto gr = [&](int dy, int dx) { return G[ty + 1 + dy][tx + 1 + dx]; };
        const float gdh = (0.25f * gr(0, -1) + -0.5f * g) + 0.25f * gr(0, 1);
        const float gdv = (0.25f * gr(-1, 0) + -0.5f * g) + 0.25f * gr(1, 0);
        const float gdp = (gr(-1, -1) + -2.f * g) + gr(1, 1);
        const float gdn = (gr(-1, 1) + -2.f * g) + gr(1, -1);
        const int py = yy & 1, px = xx & 1;
        float out[2];
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
            // ch 0: red (mode 1), ch 1: blue (mode 2, Gr / Gb swapped, the "other" channel is red)
            const int cy = ch ? a.by : a.ry, cx = ch ? a.bx : a.rx;      // where this colour is sampled
            const int oy = ch ? a.ry : a.by, ox = ch ? a.rx : a.bx;      // where the other colour is sampled
            // masked CFA of this colour with replication padding: value at the clamped position if that position
            // carries this colour, else 0
            auto m = [&](int dy, int dx) {
                const int y2 = min(max(yy + dy, 0), H2 - 1), x2 = min(max(xx + dx, 0), W2 - 1);
                return (((y2 & 1) == cy) && ((x2 & 1) == cx)) ? P[y2 - Y0 + 3][x2 - X0 + 3] : 0.f;
            };
            const bool on_row = (py == cy) && (px != cx);       // green pixel on this colour's row   (maskGr for red)
            const bool on_col = (py != cy) && (px == cx);       // green pixel on this colour's column (maskGb for red)
            const bool on_other = (py == oy) && (px == ox);     // pixel of the other colour          (mask_ochan)
            float v = 0.f;
            if (on_other) {
                const float a00 = m(-1, -1), a22 = m(1, 1), a02 = m(-1, 1), a20 = m(1, -1);
                const float cp = (0.5f * a00 + 0.5f * a22) - gdp / 4.f, cn = (0.5f * a02 + 0.5f * a20) - gdn / 4.f;
                const float clp = fabsf(-a00 + a22) + fabsf(gdp), cln = fabsf(-a02 + a20) + fabsf(gdn);
                const float s = sgn(clp - cln);
                v = (1.f + s) * cn / 2.f + (1.f - s) * cp / 2.f;
            }
            const float chh = on_row ? (0.5f * m(0, -1) + 0.5f * m(0, 1)) - gdh : 0.f;
            const float cvv = on_col ? (0.5f * m(-1, 0) + 0.5f * m(1, 0)) - gdv : 0.f;
            out[ch] = ((v + chh) + cvv) + m(0, 0);
        }
        const long long o = (long long)yy * W2 + xx;
        yb[o] = out[0];
        yb[(long long)H2 * W2 + o] = g;
        yb[2LL * H2 * W2 + o] = out[1];
    }
}

cudaError_t launch_demosaic_ha(const float *x, float *y, int B, int H, int W, int ry, int rx, int by, int bx, cudaStream_t st)
{
    DemosaicArgs a;
    a.x = x; a.y = y; a.B = B; a.H = H; a.W = W;
    a.ry = ry; a.rx = rx; a.by = by; a.bx = bx;
    const dim3 grid((2 * W + DM_TW - 1) / DM_TW, (2 * H + DM_TH - 1) / DM_TH, B);
    demosaic_ha_kernel<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// remosaick (Hamilton_Adam_demo.py:237-246) + singleiT (library.py:67: (x + 1) / 2) + mean of the 4 packed channels
// (library.py:165-167) in one pass: rgb [B][3][2H][2W] -> gray [B][H][W], the image the online-flow path hands to TV-L1.
__global__ void remosaick_gray_kernel(const float *__restrict__ rgb, float *__restrict__ gray, int H, int W, int ry, int rx,
                                      int by, int bx, float add, float mul)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const long long W2 = 2LL * W, pl = 4LL * H * W;
    const float *b = rgb + (long long)blockIdx.z * 3 * pl;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int py = k >> 1, px = k & 1;
        const int c = (py == ry && px == rx) ? 0 : ((py == by && px == bx) ? 2 : 1);
        v[k] = (b[c * pl + (2LL * y + py) * W2 + 2 * x + px] + add) * mul;
    }
    gray[((long long)blockIdx.z * H + y) * W + x] = (((v[0] + v[1]) + v[2]) + v[3]) * 0.25f;
}

cudaError_t launch_remosaick_gray(const float *rgb, float *gray, int B, int H, int W, int ry, int rx, int by, int bx,
                                  float add, float mul, cudaStream_t st)
{
    const dim3 grid((W + 127) / 128, H, B);
    remosaick_gray_kernel<<<grid, 128, 0, st>>>(rgb, gray, H, W, ry, rx, by, bx, add, mul);
    return cudaGetLastError();
}

}  // namespace rvdd
