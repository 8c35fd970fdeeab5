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

#define DM_TW 64                      // output tile: 64 x 16 pixels = 32 x 8 Bayer cells, one cell per thread
#define DM_TH 16
#define DM_PW (DM_TW + 8)             // CFA tile: 4 columns of halo on each side (3 used; 4 keep the cells 16-byte aligned)
#define DM_PH (DM_TH + 6)             //           3 rows of halo
#define DM_PX 4                       // tile column of image column X0
#define DM_GW (DM_TW + 4)             // green tile: 2 columns of halo on each side (1 used), 1 row
#define DM_GH (DM_TH + 2)
#define DM_GX 2
#define DM_GC (DM_TW / 2 + 1)           // compact green tile (red / blue positions only): columns per row

struct DemosaicArgs {
    const float *x;                   // [B][4][H][W] packed raw
    float *y;                         // [B][3][2H][2W]
    int B, H, W;
    int force_general;                // test hook (RVDD_DEMOSAIC_GENERAL=1): every tile takes the border path
};

__device__ __forceinline__ float sgn(float v) { return (float)((v > 0.f) - (v < 0.f)); }     // torch.sign

// green at a red / blue pixel (algorithm 1, Hamilton_Adam_demo.py:123-142).  p points at the pixel inside the staged
// CFA tile, whose halo already holds the replicated border (ReplicationPad2d((2, 2, 2, 2))); PW = tile pitch.
template <int PW> __device__ __forceinline__ float green_at(const float *p)
{
    const float c = p[0], l1 = p[-1], r1 = p[1], u1 = p[-PW], d1 = p[PW];
    const float kh = 0.5f * l1 + 0.5f * r1, kv = 0.5f * u1 + 0.5f * d1;
    const float dh = (p[-2] + -2.f * c) + p[2], dv = (p[-2 * PW] + -2.f * c) + p[2 * PW];
    const float rawh = kh - dh / 4.f, rawv = kv - dv / 4.f;
    const float clh = fabsf(l1 - r1) + fabsf(dh), clv = fabsf(u1 - d1) + fabsf(dv);
    const float s = sgn(clh - clv);
    return (1.f + s) * rawv / 2.f + (1.f - s) * rawh / 2.f;
}

// Phase 3 for the cell whose top-left pixel sits at p (CFA tile) / g (green tile): the missing colours from the colour
// differences (algorithm 2, :145-172).  The reference convolves the CFA MASKED to one colour, replication-padded: a
// neighbour contributes its CFA value if it lies inside the image and zero if the padding replicated a pixel of another
// colour -- at a pixel whose neighbours in that direction carry the colour, "outside the image" is exactly that case.
// EDGE = false: the tile lies inside the image with its halo, every neighbour exists and the selects disappear.
// COMPACT: the green tile holds only the red / blue positions (the only ones ever read), DM_GC per row, column
// (c + 1) >> 1 for tile column c: the lanes of a warp then read consecutive words instead of every other one (2-way bank
// conflicts); g0 points at row 0 / compact column `lane` of the cell.  Otherwise g0 points at the cell's pixel (0, 0) in
// the full-resolution green tile.
template <int RY, int RX, bool EDGE, bool COMPACT>
__device__ __forceinline__ void dm_cell(const float *p0, const float *g0, bool up, bool left, bool down, bool right,
                                        float (&red)[2][2], float (&grn)[2][2], float (&blu)[2][2])
{
    constexpr int BY = 1 - RY, BX = 1 - RX;
    // green at cell-relative (row, col): col in -1 .. 2
    auto G_ = [&](int row, int col) -> float {
        return COMPACT ? g0[row * DM_GC + ((col + 1) >> 1)] : g0[row * DM_GW + col];
    };
    // the pixel that carries colour A sees colour B on its four diagonals: B there from the diagonal colour differences
    auto diag = [&](int py, int px, float &outB) {
        const float *p = p0 + py * DM_PW + px;
        const bool t = (!EDGE || py) ? true : up, bm = (EDGE && py) ? down : true;
        const bool l = (!EDGE || px) ? true : left, r = (EDGE && px) ? right : true;
        const float a00 = (t && l) ? p[-DM_PW - 1] : 0.f, a22 = (bm && r) ? p[DM_PW + 1] : 0.f;
        const float a02 = (t && r) ? p[-DM_PW + 1] : 0.f, a20 = (bm && l) ? p[DM_PW - 1] : 0.f;
        const float gc = G_(py, px);
        const float gdp = (G_(py - 1, px - 1) + -2.f * gc) + G_(py + 1, px + 1);
        const float gdn = (G_(py - 1, px + 1) + -2.f * gc) + G_(py + 1, px - 1);
        const float cp = (0.5f * a00 + 0.5f * a22) - gdp / 4.f, cn = (0.5f * a02 + 0.5f * a20) - gdn / 4.f;
        const float clp = fabsf(-a00 + a22) + fabsf(gdp), cln = fabsf(-a02 + a20) + fabsf(gdn);
        const float s = sgn(clp - cln);
        outB = (1.f + s) * cn / 2.f + (1.f - s) * cp / 2.f;
        grn[py][px] = gc;
    };
    // a green pixel: the colour of its row from the horizontal neighbours, the colour of its column from the vertical ones
    auto cross = [&](int py, int px, float &outRow, float &outCol) {
        const float *p = p0 + py * DM_PW + px;
        const bool t = (!EDGE || py) ? true : up, bm = (EDGE && py) ? down : true;
        const bool l = (!EDGE || px) ? true : left, r = (EDGE && px) ? right : true;
        const float gc = p[0];                                   // green sample of the CFA
        // at the image border the replicated neighbour is this very pixel: its green is the sample itself
        const float gl = l ? G_(py, px - 1) : gc, grr = r ? G_(py, px + 1) : gc;
        const float gu = t ? G_(py - 1, px) : gc, gd = bm ? G_(py + 1, px) : gc;
        const float gdh = (0.25f * gl + -0.5f * gc) + 0.25f * grr;
        const float gdv = (0.25f * gu + -0.5f * gc) + 0.25f * gd;
        outRow = (0.5f * (l ? p[-1] : 0.f) + 0.5f * (r ? p[1] : 0.f)) - gdh;
        outCol = (0.5f * (t ? p[-DM_PW] : 0.f) + 0.5f * (bm ? p[DM_PW] : 0.f)) - gdv;
        grn[py][px] = gc;
    };
    // red pixel (ry, rx): blue from the diagonals; blue pixel (by, bx): red from the diagonals
    red[RY][RX] = p0[RY * DM_PW + RX];
    blu[BY][BX] = p0[BY * DM_PW + BX];
    diag(RY, RX, blu[RY][RX]);
    diag(BY, BX, red[BY][BX]);
    // green on the red row (ry, bx): red along the row, blue along the column; green on the blue row (by, rx): the reverse
    cross(RY, BX, red[RY][BX], blu[RY][BX]);
    cross(BY, RX, blu[BY][RX], red[BY][RX]);
}

// One CTA = one 64 x 16 output tile; one thread = one 2x2 Bayer cell.
//   phase 1: CFA tile + halo into shared memory (pack_in_one, :226-234; clamped coordinates = replication)
//   phase 2: green at every red / blue position of the tile + 1 pixel halo (positions outside the image take the green
//            of the clamped position, which is what the replication padding in front of conv_algo2_green produces)
//   phase 3: per cell, dm_cell and the output.
// The kernel is bound by instruction issue (profiles/demosaic_kernel_r01.txt: issue 80 %, ALU pipe 60 %, DRAM 34 %), so
// the 93 % of the tiles that lie inside the image with their halo take a path without clamps, bounds tests and index
// divisions: 16-byte interleaving stores in phase 1 (two packed planes -> four CFA columns), each thread the two greens of
// its own cell plus a 164-position ring in phase 2, select-free colour differences in phase 3.  Tiles on the image border
// (and frames whose rows are not 8-byte aligned) keep the general per-element path.  Same expressions on both.
// RY, RX: position of the red sample inside the 2x2 cell (blue sits on the opposite corner); compile-time so that the
// per-cell colour arrays stay in registers.
template <int RY, int RX>
__global__ void __launch_bounds__(256, 8) demosaic_ha_kernel(const DemosaicArgs a)
{
    constexpr int BY = 1 - RY, BX = 1 - RX;
    constexpr int rb = RY ^ RX;
    __shared__ __align__(16) float P[DM_PH][DM_PW];
    __shared__ __align__(16) float G[DM_GH][DM_GW];
    const int H2 = 2 * a.H, W2 = 2 * a.W;
    const int X0 = blockIdx.x * DM_TW, Y0 = blockIdx.y * DM_TH;
    const float *xb = a.x + (long long)blockIdx.z * 4 * a.H * a.W;
    const long long plane = (long long)a.H * a.W;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const bool inner = X0 >= DM_PX && Y0 >= 3 && X0 + DM_TW + DM_PX <= W2 && Y0 + DM_TH + 3 <= H2 && a.force_general == 0;
    const bool vec = (a.W & 1) == 0 && (reinterpret_cast<uintptr_t>(xb) & 7) == 0;
    const float *p0 = &P[2 * wrp + 3][2 * lane + DM_PX];       // this thread's cell
    float *g0 = &G[2 * wrp + 1][2 * lane + DM_GX];
    float red[2][2], grn[2][2], blu[2][2];

    if (inner) {
        // ---- phase 1: row yy of the CFA interleaves packed planes (yy & 1) * 2 and (yy & 1) * 2 + 1
        if (vec) {
            for (int i = threadIdx.x; i < DM_PH * (DM_PW / 4); i += 256) {
                const int ty = i / (DM_PW / 4), q = i - ty * (DM_PW / 4);
                const int yy = Y0 + ty - 3;
                const float *row = xb + (long long)((yy & 1) * 2) * plane + (long long)(yy >> 1) * a.W + ((X0 - DM_PX) >> 1) + 2 * q;
                const float2 e = __ldg(reinterpret_cast<const float2 *>(row));
                const float2 o = __ldg(reinterpret_cast<const float2 *>(row + plane));
                *reinterpret_cast<float4 *>(&P[ty][4 * q]) = make_float4(e.x, o.x, e.y, o.y);
            }
        } else {
            for (int i = threadIdx.x; i < DM_PH * (DM_PW / 2); i += 256) {
                const int ty = i / (DM_PW / 2), q = i - ty * (DM_PW / 2);
                const int yy = Y0 + ty - 3;
                const float *row = xb + (long long)((yy & 1) * 2) * plane + (long long)(yy >> 1) * a.W + ((X0 - DM_PX) >> 1) + q;
                *reinterpret_cast<float2 *>(&P[ty][2 * q]) = make_float2(__ldg(row), __ldg(row + plane));
            }
        }
        __syncthreads();
        // ---- phase 2: the two red / blue pixels of the own cell, then the ring around the tile; compact green tile
        float *gc0 = &G[0][0] + (2 * wrp + 1) * DM_GC + lane;     // tile row 0 of the cell, compact column of tile column 2 * lane - 1
        gc0[RY * DM_GC + ((RX + 1) >> 1)] = green_at<DM_PW>(p0 + RY * DM_PW + RX);
        gc0[BY * DM_GC + ((BX + 1) >> 1)] = green_at<DM_PW>(p0 + BY * DM_PW + BX);
        if (threadIdx.x < 2 * (DM_TW + 2) + 2 * DM_TH) {
            int ry, rx;                                          // tile-relative position on the ring
            const int t = threadIdx.x;
            if (t < DM_TW + 2) { ry = -1; rx = t - 1; }
            else if (t < 2 * (DM_TW + 2)) { ry = DM_TH; rx = t - (DM_TW + 2) - 1; }
            else if (t < 2 * (DM_TW + 2) + DM_TH) { ry = t - 2 * (DM_TW + 2); rx = -1; }
            else { ry = t - 2 * (DM_TW + 2) - DM_TH; rx = DM_TW; }
            if (((ry ^ rx) & 1) == rb)                           // X0, Y0 even: the parity inside the tile is the parity in the image
                (&G[0][0])[(ry + 1) * DM_GC + ((rx + 1) >> 1)] = green_at<DM_PW>(&P[ry + 3][rx + DM_PX]);
        }
        __syncthreads();
        dm_cell<RY, RX, false, true>(p0, gc0, true, true, true, true, red, grn, blu);
    } else {
        // ---- phase 1, general: one element at a time, coordinates clamped
        for (int ty = wrp; ty < DM_PH; ty += 8) {
            const int yy = min(max(Y0 + ty - 3, 0), H2 - 1);
            const float *row = xb + (long long)((yy & 1) * 2) * plane + (long long)(yy >> 1) * a.W;
#pragma unroll
            for (int tx = lane; tx < DM_PW; tx += 32) {
                const int xx = min(max(X0 + tx - DM_PX, 0), W2 - 1);
                P[ty][tx] = __ldg(row + (long long)(xx & 1) * plane + (xx >> 1));
            }
        }
        __syncthreads();
        // ---- phase 2, general: the red / blue positions sit where (y ^ x) & 1 == (ry ^ rx); (DM_TW + 2) / 2 per tile row
        constexpr int GWU = DM_TW + 2;                           // used columns of the green tile
        for (int i = threadIdx.x; i < DM_GH * (GWU / 2); i += 256) {
            const int gy = i / (GWU / 2), k = i - gy * (GWU / 2);
            const int gx = 2 * k + (((Y0 + gy - 1) ^ (X0 - 1) ^ rb) & 1);      // X0, GWU even: parity of column gx
            const int yy = Y0 + gy - 1, xx = X0 + gx - 1;
            if (yy >= 0 && yy < H2 && xx >= 0 && xx < W2) G[gy][gx + DM_GX - 1] = green_at<DM_PW>(&P[gy + 2][gx + DM_PX - 1]);
        }
        // halo positions outside the image: green of the clamped position, any colour
        for (int i = threadIdx.x; i < DM_GH * GWU; i += 256) {
            const int gy = i / GWU, gx = i - gy * GWU;
            const int yy = Y0 + gy - 1, xx = X0 + gx - 1;
            if (yy >= 0 && yy < H2 && xx >= 0 && xx < W2) continue;
            const int yc = min(max(yy, 0), H2 - 1), xc = min(max(xx, 0), W2 - 1);
            const float *p = &P[yc - Y0 + 3][xc - X0 + DM_PX];
            G[gy][gx + DM_GX - 1] = (((yc ^ xc) & 1) == rb) ? green_at<DM_PW>(p) : p[0];
        }
        __syncthreads();
        const int cyy = Y0 + 2 * wrp, cxx = X0 + 2 * lane;
        if (cyy >= H2 || cxx >= W2) return;
        dm_cell<RY, RX, true, false>(p0, g0, cyy > 0, cxx > 0, cyy + 2 < H2, cxx + 2 < W2, red, grn, blu);
    }

    // ---- output
    const int cyy = Y0 + 2 * wrp, cxx = X0 + 2 * lane;
    float *yb = a.y + (long long)blockIdx.z * 3 * H2 * W2 + (long long)cyy * W2 + cxx;
    const long long pl = (long long)H2 * W2;
#pragma unroll
    for (int py = 0; py < 2; py++) {
        *reinterpret_cast<float2 *>(yb + (long long)py * W2) = make_float2(red[py][0], red[py][1]);
        *reinterpret_cast<float2 *>(yb + pl + (long long)py * W2) = make_float2(grn[py][0], grn[py][1]);
        *reinterpret_cast<float2 *>(yb + 2 * pl + (long long)py * W2) = make_float2(blu[py][0], blu[py][1]);
    }
}

cudaError_t launch_demosaic_ha(const float *x, float *y, int B, int H, int W, int ry, int rx, int by, int bx, cudaStream_t st)
{
    DemosaicArgs a;
    a.x = x; a.y = y; a.B = B; a.H = H; a.W = W;
    a.force_general = getenv("RVDD_DEMOSAIC_GENERAL") != nullptr;       // read per call: tests flip it
    if (by != 1 - ry || bx != 1 - rx) return cudaErrorInvalidValue;        // red and blue sit on a diagonal of the cell
    const dim3 grid((2 * W + DM_TW - 1) / DM_TW, (2 * H + DM_TH - 1) / DM_TH, B);
    switch (ry * 2 + rx) {
    case 0: demosaic_ha_kernel<0, 0><<<grid, 256, 0, st>>>(a); break;
    case 1: demosaic_ha_kernel<0, 1><<<grid, 256, 0, st>>>(a); break;
    case 2: demosaic_ha_kernel<1, 0><<<grid, 256, 0, st>>>(a); break;
    default: demosaic_ha_kernel<1, 1><<<grid, 256, 0, st>>>(a); break;
    }
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
