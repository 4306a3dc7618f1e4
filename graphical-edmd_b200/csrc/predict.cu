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
// Away from the x edges the three cells of a row are one contiguous range of
// the cell-ordered arrays, so a row is one coalesced streak per warp; the 9x
// re-use of every neighbour record is served by L1.
#include "edmd_internal.cuh"

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

template <bool GROW>
__global__ void __launch_bounds__(kThreads)
k_predict(int n, edmd_dev_box b, double t, const double4 *__restrict__ sxv,
          const double *__restrict__ srad, const double *__restrict__ svr,
          const int32_t *__restrict__ sid, const int32_t *__restrict__ scid,
          const int32_t *__restrict__ start, double *__restrict__ t_cross,
          uint8_t *__restrict__ dir, double *__restrict__ t_coll,
          int32_t *__restrict__ partner, uint8_t *__restrict__ ctype,
          unsigned long long *__restrict__ overlap_key)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double4 p1 = sxv[s];
    const double r1 = srad[s];
    const double vr1 = GROW ? svr[s] : 0.0;
    const int id = sid[s];
    const int c = scid[s];
    const int Y = c / b.nx;
    const int X = c - Y * b.nx;

    // ---- cell crossing ----------------------------------------------------
    double tx, ty;
    int xx, yy;
    if (p1.z < 0) {
        tx = __ddiv_rn(min_image(__dsub_rn(__dmul_rn((double)X, b.csx), p1.x),
                                 b.half_lx, b.lx), p1.z);
        xx = 1;
    } else {
        tx = __ddiv_rn(min_image(__dsub_rn(__dmul_rn((double)(1 + X), b.csx), p1.x),
                                 b.half_lx, b.lx), p1.z);
        xx = 2;
    }
    if (p1.w < 0) {
        ty = __ddiv_rn(min_image(__dsub_rn(__dmul_rn((double)Y, b.csy), p1.y),
                                 b.half_ly, b.ly), p1.w);
        yy = 3;
    } else {
        ty = __ddiv_rn(min_image(__dsub_rn(__dmul_rn((double)(1 + Y), b.csy), p1.y),
                                 b.half_ly, b.ly), p1.w);
        yy = 4;
    }
    const bool takex = tx < ty;  // strict: ties go to y
    t_cross[id] = __dadd_rn(t, takex ? tx : ty);
    dir[id] = (uint8_t)(takex ? xx : yy);

    // ---- collision: 3x3 cells, reference scan order -------------------------
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
                lo = start[rowbase + X - 1];
                hi = start[rowbase + X + 2];
            } else {
                int cc = rowbase + wrap_cell(X + k - 1, b.nx);
                lo = start[cc];
                hi = start[cc + 1];
            }
#pragma unroll 1
            for (int p = lo; p < hi; p++) {
                if (p == s) continue;
                const double4 p2 = sxv[p];
                const double r2 = srad[p];
                bool ov = false;
                double dt;
                if (GROW)
                    dt = pair_time_grow(b, p1, r1, vr1, p2, r2, svr[p], ov);
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
    t_coll[id] = __dadd_rn(t, best);
    partner[id] = best_slot >= 0 ? sid[best_slot] : 0;
    ctype[id] = EDMD_EV_COLLISION;
    if (first_overlap >= 0) {
        unsigned long long key = ((unsigned long long)(uint32_t)id << 32) |
                                 (uint32_t)sid[first_overlap];
        atomicMin(overlap_key, key);
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

int edmd_launch_predict(edmd_ctx *c, int mode)
{
    int n = c->n;
    if (n == 0) return 0;
    int blocks = (n + kThreads - 1) / kThreads;
    if (mode == EDMD_MODE_GROW)
        k_predict<true><<<blocks, kThreads, 0, c->stream>>>(
            n, c->dbox, c->t, c->sxv, c->srad, c->svr, c->sid, c->scid,
            c->cell_start, c->t_cross, c->dir, c->t_coll, c->partner, c->ctype,
            c->overlap_key);
    else
        k_predict<false><<<blocks, kThreads, 0, c->stream>>>(
            n, c->dbox, c->t, c->sxv, c->srad, c->svr, c->sid, c->scid,
            c->cell_start, c->t_cross, c->dir, c->t_coll, c->partner, c->ctype,
            c->overlap_key);
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
