// tile.cuh -- shared-memory staging of a 2-D tile of cells plus its one-cell
// halo, used by the prediction sweep (K1) and the psi6 kernel (K4).
//
// A CTA owns a tile of TX x TY cells (run-time sizes, chosen by the host from
// the mean cell occupancy so that a tile holds ~100 particles for 128 threads).
// In the cell-ordered arrays every row of the halo box is (left halo cell |
// main cells | right halo cell); the main cells are one contiguous run, the two
// halo cells are contiguous with it except at the periodic x edge.  Each row is
// copied to shared memory as one run in that order, which is exactly the
// reference's column order k = -1, 0, 1 (src/EDMD.c:2959-2965), and the rows
// are stored in the reference's row order j = -1, 0, 1.  Per halo cell a local
// start offset is kept, so the candidate list of a particle in local cell
// (r, xl) is three contiguous shared-memory ranges
//     [coff[r+j][xl-1], coff[r+j][xl+2])   j = -1, 0, 1.
//
// When the halo box touches no periodic edge, the grid is at least 12 cells
// wide and every staged particle sits within one cell width of the cell it is
// filed under, no pair in the tile can need a minimum-image shift
// (|d| <= 5 cells < L/2), and `fast` is set: the PBC compare/adjust
// (src/EDMD.c:5896-5913) is then the identity and is skipped.
#pragma once

#include "edmd_internal.cuh"

constexpr int kTileThreads = 128;
constexpr int kTileMaxTX = 32;
constexpr int kTileMaxTY = 8;
constexpr int kTileRows = kTileMaxTY + 2;
constexpr int kTileCols = kTileMaxTX + 2;
constexpr int kTileCap = 384;  // staged particles per CTA

struct TileShared {
    double4 xv[kTileCap];
    double rad[kTileCap];
    int id[kTileCap];
    int cell[kTileCap];
    unsigned short coff[kTileRows][kTileCols + 2];
    int seg_lo[kTileRows][3];
    int seg_n[kTileRows][3];
    int row_off[kTileRows + 1];
    int own_cum[kTileRows + 1];
    int staged, own, fast, overflow;
};

struct TileInfo {
    int x0, y0, txe, tye, rows, cols;
};

__device__ __forceinline__ int tile_wrap(int a, int n)
{
    if (a < 0) return a + n;
    if (a >= n) return a - n;
    return a;
}

// Stage the tile `blockIdx.x` into `s`.  Returns the tile geometry.  All
// threads of the CTA must call it; ends with a __syncthreads().
__device__ __forceinline__ TileInfo tile_stage(TileShared &s, const edmd_dev_box &b, int tx, int ty,
                                               int tiles_x, const double4 *__restrict__ sxv,
                                               const double *__restrict__ srad,
                                               const int32_t *__restrict__ sid,
                                               const int32_t *__restrict__ scid,
                                               const int32_t *__restrict__ start)
{
    const int t = threadIdx.x;
    TileInfo ti;
    const int by = blockIdx.x / tiles_x;
    const int bx = blockIdx.x - by * tiles_x;
    ti.x0 = bx * tx;
    ti.y0 = by * ty;
    ti.txe = min(tx, b.nx - ti.x0);
    ti.tye = min(ty, b.ny - ti.y0);
    ti.rows = ti.tye + 2;
    ti.cols = ti.txe + 2;

    if (t < ti.rows) {
        const int base = tile_wrap(ti.y0 - 1 + t, b.ny) * b.nx;
        const int xl = base + tile_wrap(ti.x0 - 1, b.nx);
        const int xr = base + tile_wrap(ti.x0 + ti.txe, b.nx);
        const int m0 = start[base + ti.x0];
        const int m1 = start[base + ti.x0 + ti.txe];
        const int l0 = start[xl], l1 = start[xl + 1];
        const int r0 = start[xr], r1 = start[xr + 1];
        s.seg_lo[t][0] = l0; s.seg_n[t][0] = l1 - l0;
        s.seg_lo[t][1] = m0; s.seg_n[t][1] = m1 - m0;
        s.seg_lo[t][2] = r0; s.seg_n[t][2] = r1 - r0;
    }
    __syncthreads();
    if (t == 0) {
        int acc = 0, own = 0;
        for (int r = 0; r < ti.rows; r++) {
            s.row_off[r] = acc;
            acc += s.seg_n[r][0] + s.seg_n[r][1] + s.seg_n[r][2];
        }
        s.row_off[ti.rows] = acc;
        s.own_cum[0] = 0;
        for (int r = 1; r <= ti.tye; r++) {
            own += s.seg_n[r][1];
            s.own_cum[r] = own;
        }
        s.staged = acc;
        s.own = own;
        s.overflow = acc > kTileCap;
        s.fast = (b.nx >= 12) && (b.ny >= 12) && (ti.x0 >= 1) && (ti.x0 + ti.txe < b.nx) &&
                 (ti.y0 >= 1) && (ti.y0 + ti.tye < b.ny);
    }
    __syncthreads();
    if (s.overflow) return ti;  // caller falls back to the global-memory path

    // local start offset of every halo cell (and the row end)
    const int per_row = ti.cols + 1;
    for (int idx = t; idx < ti.rows * per_row; idx += kTileThreads) {
        const int r = idx / per_row;
        const int xl = idx - r * per_row;
        const int nl = s.seg_n[r][0], nm = s.seg_n[r][1], nr = s.seg_n[r][2];
        int off;
        if (xl == 0) off = 0;
        else if (xl <= ti.txe) {
            const int base = tile_wrap(ti.y0 - 1 + r, b.ny) * b.nx;
            off = nl + (start[base + ti.x0 + xl - 1] - s.seg_lo[r][1]);
        } else if (xl == ti.txe + 1) off = nl + nm;
        else off = nl + nm + nr;
        s.coff[r][xl] = (unsigned short)(s.row_off[r] + off);
    }
    // the particles themselves
    const int staged = s.staged;
    const bool check = s.fast != 0;
    for (int m = t; m < staged; m += kTileThreads) {
        int r = 0;
        while (m >= s.row_off[r + 1]) r++;
        int loc = m - s.row_off[r];
        int g;
        const int n0 = s.seg_n[r][0], n1 = s.seg_n[r][1];
        if (loc < n0) g = s.seg_lo[r][0] + loc;
        else if (loc < n0 + n1) g = s.seg_lo[r][1] + (loc - n0);
        else g = s.seg_lo[r][2] + (loc - n0 - n1);
        const double4 p = sxv[g];
        const int c = scid[g];
        s.xv[m] = p;
        s.rad[m] = srad[g];
        s.id[m] = sid[g];
        s.cell[m] = c;
        if (check) {
            // filed cell (X, Y); the particle must lie within [X-1, X+2) x [Y-1, Y+2) cells
            const int Y = ti.y0 - 1 + r;  // no wrap on the fast path
            const int X = c - Y * b.nx;
            const double cx = ((double)X + 0.5) * b.csx, cy = ((double)Y + 0.5) * b.csy;
            if (!(fabs(p.x - cx) <= 1.5 * b.csx) || !(fabs(p.y - cy) <= 1.5 * b.csy)) s.fast = 0;
        }
    }
    __syncthreads();
    return ti;
}

// k-th own particle of the tile -> halo row r (1..tye) and shared index q
__device__ __forceinline__ void tile_own(const TileShared &s, const TileInfo &ti, int k, int &r, int &q)
{
    r = 1;
    while (k >= s.own_cum[r]) r++;
    q = s.row_off[r] + s.seg_n[r][0] + (k - s.own_cum[r - 1]);
}
