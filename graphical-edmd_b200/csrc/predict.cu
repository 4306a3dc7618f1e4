// predict.cu -- K1: the whole-system prediction sweep.
//
// One thread per particle, in cell order.  For particle i this evaluates what
// the reference's `crossingEvent(i); collisionEvent(i);` evaluate
//   crossingEventNormal / crossingEventGrow      src/EDMD.c:2405-2482, 2343-2403
//   collisionEventNormal / collisionEventGrow    src/EDMD.c:2829-3102, 3104-3230
//   collisionTimeNormal / collisionTimeGrow      src/EDMD.c:2661-2723, 2598-2659
//   PBC, PBCinsideCellX/Y                        src/EDMD.c:5896-5936
// on a synchronous snapshot (lat2 == 0), in the reference's evaluation order
// and WITHOUT fused multiply-adds: every product and sum that the reference's
// (FMA-less x86-64) build rounds separately is an explicit __dmul_rn /
// __dadd_rn / __dsub_rn here; division and square root are IEEE
// correctly-rounded on both sides.  That makes event times bit-identical, not
// merely close.
//
// Scan order = reference order: rows j = -1..1, columns k = -1..1, inside a
// cell descending particle id (the cell index keeps that order), running
// minimum with a strict `>`; NaN candidates lose; no candidate => partner 0 at
// t + 1e26 (1e7 while growing).
//
// Two kernels:
//   k_predict_tile  (NORMAL mode) -- a CTA stages a 2-D tile of cells + halo in
//       shared memory (tile.cuh) with coalesced loads, then every thread scans
//       its particle's 3x3 cells out of shared memory.  Candidate selection is
//       two-phase: phase 1 computes b, |dv|^2, c, det EXACTLY (they decide the
//       reference's `b > 0` / `det < 0` / overlap branches) and ranks the
//       survivors by an APPROXIMATE time c / (sqrt~(det) - b) built from the
//       MUFU rsqrt/rcp seeds (rel. error ~2^-20); phase 2 evaluates the
//       reference's formula (-b - sqrt(det)) / v2 with IEEE sqrt and division
//       for the winner only.  If any other survivor lies within 2^-13 relative
//       of the winner, or anything looks ill-conditioned (non-positive or
//       non-finite estimate, catastrophic cancellation in -b - sqrt(det)), the
//       particle is re-done by the plain exact loop, so the result is always
//       the reference's, bit for bit.
//   k_predict_generic -- the plain exact loop straight from global memory;
//       used for GROW mode (once per run) and as the overflow fallback.
#include "edmd_internal.cuh"
#include "tile.cuh"

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ double min_image(double d, double half, double len)
{
    // `if (d >= half) d -= L; else if (d < -half) d += L;`
    if (d >= half) return __dsub_rn(d, len);
    if (d < -half) return __dadd_rn(d, len);
    return d;
}

__device__ __forceinline__ int wrap_cell(int a, int n)
{
    if (a < 0) return a + n;
    if (a >= n) return a - n;
    return a;
}

// collisionTimeNormal, lat2 == 0.  Returns the candidate time (may be NaN).
__device__ __forceinline__ double pair_time_normal(const edmd_dev_box &b,
                                                   const double4 &p1, double r1,
                                                   const double4 &p2, double r2,
                                                   bool &overlap)
{
    double dvx = __dsub_rn(p2.z, p1.z);
    double dvy = __dsub_rn(p2.w, p1.w);
    double dx = min_image(__dsub_rn(p2.x, p1.x), b.half_lx, b.lx);
    double dy = min_image(__dsub_rn(p2.y, p1.y), b.half_ly, b.ly);
    double bb = __dadd_rn(__dmul_rn(dx, dvx), __dmul_rn(dy, dvy));
    if (bb > 0) return EDMD_NEVER;
    double v2 = __dadd_rn(__dmul_rn(dvx, dvx), __dmul_rn(dvy, dvy));
    double c = __dsub_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)),
                         __dmul_rn(__dmul_rn(4.0, r1), r2));
    double det = __dsub_rn(__dmul_rn(bb, bb), __dmul_rn(v2, c));
    if (c < -0.01) overlap = true;
    if (det < 0) return EDMD_NEVER;
    return __ddiv_rn(__dsub_rn(-bb, __dsqrt_rn(det)), v2);
}

// collisionTimeGrow, lat2 == 0.
__device__ __forceinline__ double pair_time_grow(const edmd_dev_box &b,
                                                 const double4 &p1, double r1,
                                                 double vr1, const double4 &p2,
                                                 double r2, double vr2,
                                                 bool &overlap)
{
    double dvx = __dsub_rn(p2.z, p1.z);
    double dvy = __dsub_rn(p2.w, p1.w);
    double dvr = __dadd_rn(vr1, vr2);
    double dx = __dsub_rn(p2.x, p1.x);
    double dy = __dsub_rn(p2.y, p1.y);
    double dr = __dsqrt_rn(__dmul_rn(__dmul_rn(4.0, r1), r2));
    dx = min_image(dx, b.half_lx, b.lx);
    dy = min_image(dy, b.half_ly, b.ly);
    double bb = __dsub_rn(__dadd_rn(__dmul_rn(dx, dvx), __dmul_rn(dy, dvy)),
                          __dmul_rn(dvr, dr));
    double v2 = __dadd_rn(__dmul_rn(dvx, dvx), __dmul_rn(dvy, dvy));
    double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    double a = __dsub_rn(v2, __dmul_rn(dvr, dvr));
    double gap = __dsub_rn(d2, __dmul_rn(dr, dr));
    double det = __dsub_rn(__dmul_rn(bb, bb), __dmul_rn(a, gap));
    if (det < 0) return EDMD_NEVER;
    double sq = __dsqrt_rn(det);
    double plus = __ddiv_rn(__dadd_rn(-bb, sq), a);
    double minus = __ddiv_rn(__dsub_rn(-bb, sq), a);
    if (((minus > 0) && (plus > 0) && (minus < plus)) || ((minus > 0) && (plus < 0)))
        return minus;
    else if (((minus > 0) && (plus > 0.000000001) && (plus < minus)) ||
             ((minus < 0) && (plus > 0.000000001)))
        return plus;
    if (gap < -0.01) overlap = true;
    return EDMD_NEVER;
}

// Everything the sweep reads and writes (kernel argument block).
struct SweepArgs {
    int n;
    edmd_dev_box b;
    double t;
    const double4 *sxv;
    const double *srad;
    const double *svr;
    const int32_t *sid;
    const int32_t *scid;
    const int32_t *start;
    double *t_cross;
    uint8_t *dir;
    double *t_coll;
    int32_t *partner;
    uint8_t *ctype;
    unsigned long long *overlap_key;
    unsigned int *stats;  // [0] particles resolved by the exact re-scan
    int tx, ty, tiles_x;
};

// crossingEventNormal / crossingEventGrow for one particle, exact.
template <bool WRAP>
__device__ __forceinline__ void crossing_exact(const edmd_dev_box &b, const double4 &p1, int X, int Y,
                                               double &dt, int &d)
{
    double ax = __dsub_rn(__dmul_rn((double)(p1.z < 0 ? X : 1 + X), b.csx), p1.x);
    double ay = __dsub_rn(__dmul_rn((double)(p1.w < 0 ? Y : 1 + Y), b.csy), p1.y);
    if (WRAP) {
        ax = min_image(ax, b.half_lx, b.lx);
        ay = min_image(ay, b.half_ly, b.ly);
    }
    const double tx = __ddiv_rn(ax, p1.z);
    const double ty = __ddiv_rn(ay, p1.w);
    const bool takex = tx < ty;  // strict: ties go to y
    dt = takex ? tx : ty;
    d = takex ? (p1.z < 0 ? 1 : 2) : (p1.w < 0 ? 3 : 4);
}

// One particle straight from the cell-ordered global arrays: the reference's
// loops as written.  Used by k_predict_generic and as the tile overflow path.
template <bool GROW>
__device__ void predict_one_global(const SweepArgs &a, int s)
{
    const edmd_dev_box &b = a.b;
    const double4 p1 = a.sxv[s];
    const double r1 = a.srad[s];
    const double vr1 = GROW ? a.svr[s] : 0.0;
    const int id = a.sid[s];
    const int c = a.scid[s];
    const int Y = c / b.nx;
    const int X = c - Y * b.nx;

    double dtc;
    int d;
    crossing_exact<true>(b, p1, X, Y, dtc, d);
    a.t_cross[id] = __dadd_rn(a.t, dtc);
    a.dir[id] = (uint8_t)d;

    double best = GROW ? 10000000.0 : EDMD_NEVER;
    int best_slot = -1;
    int first_overlap = -1;
    const bool interior = (X >= 1) && (X + 1 < b.nx);
#pragma unroll 1
    for (int j = -1; j <= 1; j++) {
        const int rowbase = wrap_cell(Y + j, b.ny) * b.nx;
        const int nseg = interior ? 1 : 3;
#pragma unroll 1
        for (int k = 0; k < nseg; k++) {
            int lo, hi;
            if (interior) {
                lo = a.start[rowbase + X - 1];
                hi = a.start[rowbase + X + 2];
            } else {
                int cc = rowbase + wrap_cell(X + k - 1, b.nx);
                lo = a.start[cc];
                hi = a.start[cc + 1];
            }
#pragma unroll 1
            for (int p = lo; p < hi; p++) {
                if (p == s) continue;
                const double4 p2 = a.sxv[p];
                const double r2 = a.srad[p];
                bool ov = false;
                double dt;
                if (GROW)
                    dt = pair_time_grow(b, p1, r1, vr1, p2, r2, a.svr[p], ov);
                else
                    dt = pair_time_normal(b, p1, r1, p2, r2, ov);
                if (ov && first_overlap < 0) first_overlap = p;
                if (best > dt) {
                    best = dt;
                    best_slot = p;
                }
            }
        }
    }
    a.t_coll[id] = __dadd_rn(a.t, best);
    a.partner[id] = best_slot >= 0 ? a.sid[best_slot] : 0;
    a.ctype[id] = EDMD_EV_COLLISION;
    if (first_overlap >= 0) {
        unsigned long long key = ((unsigned long long)(uint32_t)id << 32) |
                                 (uint32_t)a.sid[first_overlap];
        atomicMin(a.overlap_key, key);
    }
}

template <bool GROW>
__global__ void __launch_bounds__(kThreads)
k_predict_generic(const __grid_constant__ SweepArgs a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < a.n) predict_one_global<GROW>(a, s);
}

// ---- tiled kernel ---------------------------------------------------------
__device__ __forceinline__ double rsqrt_seed(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

__device__ __forceinline__ double rcp_seed(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// hi word of a positive normal finite double, else "suspicious"
__device__ __forceinline__ bool hi_suspicious(int h)
{
    return (unsigned)(h - 0x00100000) >= (unsigned)(0x7ff00000 - 0x00100000);
}

constexpr int kBandHi = 128;  // 128 * 2^-20 = 2^-13 relative

// exact b, v2, c, det of the pair (shared index q -> p), reference order
template <bool WRAP>
__device__ __forceinline__ void pair_terms(const edmd_dev_box &b, const double4 &p1, double four_r1,
                                           const double4 &p2, double r2, double &bb, double &v2,
                                           double &c, double &b2, double &vc)
{
    const double dvx = __dsub_rn(p2.z, p1.z);
    const double dvy = __dsub_rn(p2.w, p1.w);
    double dx = __dsub_rn(p2.x, p1.x);
    double dy = __dsub_rn(p2.y, p1.y);
    if (WRAP) {
        dx = min_image(dx, b.half_lx, b.lx);
        dy = min_image(dy, b.half_ly, b.ly);
    }
    bb = __dadd_rn(__dmul_rn(dx, dvx), __dmul_rn(dy, dvy));
    v2 = __dadd_rn(__dmul_rn(dvx, dvx), __dmul_rn(dvy, dvy));
    c = __dsub_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(four_r1, r2));
    b2 = __dmul_rn(bb, bb);
    vc = __dmul_rn(v2, c);
}

template <bool WRAP>
__device__ __forceinline__ void predict_one_tile(const SweepArgs &a, const TileShared &s,
                                                 const TileInfo &ti, int r, int q)
{
    const edmd_dev_box &b = a.b;
    const double4 p1 = s.xv[q];
    const double r1 = s.rad[q];
    const double four_r1 = __dmul_rn(4.0, r1);
    const int id = s.id[q];
    const int Y = tile_wrap(ti.y0 - 1 + r, b.ny);
    const int X = s.cell[q] - Y * b.nx;
    const int xl = X - ti.x0 + 1;

    // ---- crossing: rank the two axes by seeds, divide once -----------------
    {
        double ax = __dsub_rn(__dmul_rn((double)(p1.z < 0 ? X : 1 + X), b.csx), p1.x);
        double ay = __dsub_rn(__dmul_rn((double)(p1.w < 0 ? Y : 1 + Y), b.csy), p1.y);
        if (WRAP) {
            ax = min_image(ax, b.half_lx, b.lx);
            ay = min_image(ay, b.half_ly, b.ly);
        }
        const double qx = ax * rcp_seed(p1.z);
        const double qy = ay * rcp_seed(p1.w);
        const int hx = __double2hiint(qx), hy = __double2hiint(qy);
        double dtc;
        int d;
        if (hi_suspicious(hx) || hi_suspicious(hy) || abs(hx - hy) <= kBandHi) {
            const double tx = __ddiv_rn(ax, p1.z);
            const double ty = __ddiv_rn(ay, p1.w);
            const bool takex = tx < ty;
            dtc = takex ? tx : ty;
            d = takex ? (p1.z < 0 ? 1 : 2) : (p1.w < 0 ? 3 : 4);
        } else {
            const bool takex = qx < qy;
            dtc = __ddiv_rn(takex ? ax : ay, takex ? p1.z : p1.w);
            d = takex ? (p1.z < 0 ? 1 : 2) : (p1.w < 0 ? 3 : 4);
        }
        a.t_cross[id] = __dadd_rn(a.t, dtc);
        a.dir[id] = (uint8_t)d;
    }

    // ---- collision, phase 1: exact filter + seed ranking --------------------
    double m1 = EDMD_NEVER;
    int hm = __double2hiint(EDMD_NEVER);
    int ibest = -1;
    int first_overlap = -1;
    bool amb = false;
#pragma unroll 1
    for (int rr = r - 1; rr <= r + 1; rr++) {
        const int lo = s.coff[rr][xl - 1];
        const int hi = s.coff[rr][xl + 2];
#pragma unroll 1
        for (int p = lo; p < hi; p++) {
            const double4 p2 = s.xv[p];
            const double r2 = s.rad[p];
            double bb, v2, c, b2, vc;
            pair_terms<WRAP>(b, p1, four_r1, p2, r2, bb, v2, c, b2, vc);
            const double det = __dsub_rn(b2, vc);
            // reference: `if (b > 0) never` ... overlap check ... `if (det < 0) never`
            const bool approaching = !(bb > 0) && (p != q);
            if (approaching && (c < -0.01) && first_overlap < 0) first_overlap = p;
            if (approaching && (det >= 0)) {
                const double sq = det * rsqrt_seed(det);   // ~sqrt(det); NaN when det == 0
                const double qd = c * rcp_seed(sq - bb);   // ~ c / (sqrt(det) - b)
                const int hq = __double2hiint(qd);
                amb |= hi_suspicious(hq) || (abs(hq - hm) <= kBandHi) ||
                       (vc < 1e-10 * b2);                  // cancellation in -b - sqrt(det)
                if (qd < m1) {
                    m1 = qd;
                    hm = hq;
                    ibest = p;
                }
            }
        }
    }

    // ---- phase 2: the reference's formula for the winner --------------------
    double best = EDMD_NEVER;
    int best_slot = -1;
    if (!amb) {
        if (ibest >= 0) {
            double bb, v2, c, b2, vc;
            pair_terms<WRAP>(b, p1, four_r1, s.xv[ibest], s.rad[ibest], bb, v2, c, b2, vc);
            const double det = __dsub_rn(b2, vc);
            const double dt = __ddiv_rn(__dsub_rn(-bb, __dsqrt_rn(det)), v2);
            if (best > dt) {
                best = dt;
                best_slot = ibest;
            }
        }
    } else {
        // exact re-scan in reference order (strict >, first minimum wins)
        atomicAdd(a.stats, 1u);
#pragma unroll 1
        for (int rr = r - 1; rr <= r + 1; rr++) {
            const int lo = s.coff[rr][xl - 1];
            const int hi = s.coff[rr][xl + 2];
#pragma unroll 1
            for (int p = lo; p < hi; p++) {
                if (p == q) continue;
                double bb, v2, c, b2, vc;
                pair_terms<WRAP>(b, p1, four_r1, s.xv[p], s.rad[p], bb, v2, c, b2, vc);
                if (bb > 0) continue;
                const double det = __dsub_rn(b2, vc);
                if (det < 0) continue;
                const double dt = __ddiv_rn(__dsub_rn(-bb, __dsqrt_rn(det)), v2);
                if (best > dt) {
                    best = dt;
                    best_slot = p;
                }
            }
        }
    }
    a.t_coll[id] = __dadd_rn(a.t, best);
    a.partner[id] = best_slot >= 0 ? s.id[best_slot] : 0;
    a.ctype[id] = EDMD_EV_COLLISION;
    if (first_overlap >= 0) {
        unsigned long long key = ((unsigned long long)(uint32_t)id << 32) |
                                 (uint32_t)s.id[first_overlap];
        atomicMin(a.overlap_key, key);
    }
}

__global__ void __launch_bounds__(kTileThreads)
k_predict_tile(const __grid_constant__ SweepArgs a)
{
    __shared__ TileShared s;
    const TileInfo ti = tile_stage(s, a.b, a.tx, a.ty, a.tiles_x, a.sxv, a.srad, a.sid, a.scid, a.start);
    const int own = s.own;
    if (s.overflow) {
        // more particles than the staging buffer holds: global-memory path
        for (int k = threadIdx.x; k < own; k += kTileThreads) {
            int r = 1;
            while (k >= s.own_cum[r]) r++;
            predict_one_global<false>(a, s.seg_lo[r][1] + (k - s.own_cum[r - 1]));
        }
        return;
    }
    const bool fast = s.fast != 0;
    for (int k = threadIdx.x; k < own; k += kTileThreads) {
        int r, q;
        tile_own(s, ti, k, r, q);
        if (fast) predict_one_tile<false>(a, s, ti, r, q);
        else predict_one_tile<true>(a, s, ti, r, q);
    }
}

// ---- K2: batched free flight (freeFlyNormal / freeFlyGrow) -----------------
template <bool GROW>
__global__ void __launch_bounds__(256)
k_free_fly(int n, edmd_dev_box b, double dt, double4 *__restrict__ xv,
           double *__restrict__ rad, const double *__restrict__ vr)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 p = xv[i];
    double x = __dadd_rn(p.x, __dmul_rn(dt, p.z));
    double y = __dadd_rn(p.y, __dmul_rn(dt, p.w));
    if (GROW) rad[i] = __dadd_rn(rad[i], __dmul_rn(dt, vr[i]));
    // PBCpostX / PBCpostY: a single +-L
    if (x < 0) x = __dadd_rn(x, b.lx);
    else if (x >= b.lx) x = __dsub_rn(x, b.lx);
    if (y < 0) y = __dadd_rn(y, b.ly);
    else if (y >= b.ly) y = __dsub_rn(y, b.ly);
    reinterpret_cast<double2 *>(xv)[2 * (size_t)i] = make_double2(x, y);
}

}  // namespace

void edmd_tile_dims(const edmd_ctx *c, int *tx, int *ty)
{
    // aim at ~100 own particles per 128-thread CTA
    double per_cell = c->dbox.nc > 0 ? (double)c->n / (double)c->dbox.nc : 1.0;
    if (per_cell < 0.05) per_cell = 0.05;
    int y = kTileMaxTY;
    if (y > c->dbox.ny) y = c->dbox.ny;
    int x = (int)(100.0 / (per_cell * y) + 0.5);
    if (x < 2) x = 2;
    if (x > kTileMaxTX) x = kTileMaxTX;
    if (x > c->dbox.nx) x = c->dbox.nx;
    *tx = x;
    *ty = y;
}

int edmd_launch_predict(edmd_ctx *c, int mode)
{
    int n = c->n;
    if (n == 0) return 0;
    SweepArgs a;
    a.n = n;
    a.b = c->dbox;
    a.t = c->t;
    a.sxv = c->sxv;
    a.srad = c->srad;
    a.svr = c->svr;
    a.sid = c->sid;
    a.scid = c->scid;
    a.start = c->cell_start;
    a.t_cross = c->t_cross;
    a.dir = c->dir;
    a.t_coll = c->t_coll;
    a.partner = c->partner;
    a.ctype = c->ctype;
    a.overlap_key = c->overlap_key;
    a.stats = reinterpret_cast<unsigned int *>(c->flags + 1);
    edmd_tile_dims(c, &a.tx, &a.ty);
    a.tiles_x = (c->dbox.nx + a.tx - 1) / a.tx;
    if (mode == EDMD_MODE_GROW || c->force_generic) {
        int blocks = (n + kThreads - 1) / kThreads;
        if (mode == EDMD_MODE_GROW)
            k_predict_generic<true><<<blocks, kThreads, 0, c->stream>>>(a);
        else
            k_predict_generic<false><<<blocks, kThreads, 0, c->stream>>>(a);
        return 1;
    }
    int tiles_y = (c->dbox.ny + a.ty - 1) / a.ty;
    k_predict_tile<<<a.tiles_x * tiles_y, kTileThreads, 0, c->stream>>>(a);
    return 1;
}

int edmd_launch_free_fly(edmd_ctx *c, int mode, double dt)
{
    int n = c->n;
    if (n == 0) return 0;
    int blocks = (n + 255) / 256;
    if (mode == EDMD_MODE_GROW)
        k_free_fly<true><<<blocks, 256, 0, c->stream>>>(n, c->dbox, dt, c->xv,
                                                        c->rad, c->vr);
    else
        k_free_fly<false><<<blocks, 256, 0, c->stream>>>(n, c->dbox, dt, c->xv,
                                                         c->rad, c->vr);
    return 1;
}
