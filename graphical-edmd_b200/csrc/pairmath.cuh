// pairmath.cuh -- the reference's pair and crossing arithmetic in FP64, shared by
// the sweep kernels (predict.cu, predict_lean.cu).
//   collisionTimeNormal   src/EDMD.c:2661-2723      PBC   src/EDMD.c:5896-5913
//   crossingEventNormal   src/EDMD.c:2405-2482      PBCinsideCellX/Y  src/EDMD.c:5922-5936
// Reference evaluation order, no fused multiply-adds (explicit __d*_rn), IEEE
// sqrt and division: results are bit-identical to the reference's x86-64 build.
#pragma once

#include "edmd_internal.cuh"

__device__ __forceinline__ double min_image(double d, double half, double len)
{
    // `if (d >= half) d -= L; else if (d < -half) d += L;`
    if (d >= half) return __dsub_rn(d, len);
    if (d < -half) return __dadd_rn(d, len);
    return d;
}

// exact b, |dv|^2, c and the two products of det, reference order
// (collisionTimeNormal with lat2 == 0)
template <bool WRAP>
__device__ __forceinline__ void pair_terms(const edmd_dev_box &b, const SRec &p1, double four_r1,
                                           const SRec &p2, double &bb, double &v2, double &c,
                                           double &b2, double &vc)
{
    const double dvx = __dsub_rn(p2.vx, p1.vx);
    const double dvy = __dsub_rn(p2.vy, p1.vy);
    double dx = __dsub_rn(p2.x, p1.x);
    double dy = __dsub_rn(p2.y, p1.y);
    if (WRAP) {
        dx = min_image(dx, b.half_lx, b.lx);
        dy = min_image(dy, b.half_ly, b.ly);
    }
    bb = __dadd_rn(__dmul_rn(dx, dvx), __dmul_rn(dy, dvy));
    v2 = __dadd_rn(__dmul_rn(dvx, dvx), __dmul_rn(dvy, dvy));
    c = __dsub_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(four_r1, p2.rad));
    b2 = __dmul_rn(bb, bb);
    vc = __dmul_rn(v2, c);
}

// collisionTimeNormal: candidate time (NaN possible), sets overlap
template <bool WRAP>
__device__ __forceinline__ double pair_time_normal(const edmd_dev_box &b, const SRec &p1,
                                                   double four_r1, const SRec &p2, bool &overlap)
{
    double bb, v2, c, b2, vc;
    pair_terms<WRAP>(b, p1, four_r1, p2, bb, v2, c, b2, vc);
    if (bb > 0) return EDMD_NEVER;
    const double det = __dsub_rn(b2, vc);
    if (c < -0.01) overlap = true;
    if (det < 0) return EDMD_NEVER;
    return __ddiv_rn(__dsub_rn(-bb, __dsqrt_rn(det)), v2);
}

// crossingEventNormal / crossingEventGrow, exact
template <bool WRAP>
__device__ __forceinline__ void crossing_exact(const edmd_dev_box &b, const SRec &p1, int X, int Y,
                                               double &dt, int &d)
{
    double ax = __dsub_rn(__dmul_rn((double)(p1.vx < 0 ? X : 1 + X), b.csx), p1.x);
    double ay = __dsub_rn(__dmul_rn((double)(p1.vy < 0 ? Y : 1 + Y), b.csy), p1.y);
    if (WRAP) {
        ax = min_image(ax, b.half_lx, b.lx);
        ay = min_image(ay, b.half_ly, b.ly);
    }
    const double tx = __ddiv_rn(ax, p1.vx);
    const double ty = __ddiv_rn(ay, p1.vy);
    const bool takex = tx < ty;  // strict: ties go to y
    dt = takex ? tx : ty;
    d = takex ? (p1.vx < 0 ? 1 : 2) : (p1.vy < 0 ? 3 : 4);
}

// ---- reciprocal / rsqrt seeds (MUFU) ---------------------------------------
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

// hi word of anything but a positive, normal, finite double
__device__ __forceinline__ bool hi_suspicious(int h)
{
    return (unsigned)(h - 0x00100000) >= (unsigned)(0x7ff00000 - 0x00100000);
}

constexpr int kBandHi = 128;  // 128 * 2^-20 = 2^-13 relative

// crossingEventNormal with ONE division: the two axes are ranked by reciprocal
// seeds (relative error ~2^-20); only when the two quotients are within 2^-13
// of each other, or either looks odd, both exact quotients are formed and
// compared as the reference does (strict <, ties go to y).
template <bool WRAP>
__device__ __forceinline__ void crossing_fast(const edmd_dev_box &b, const SRec &p1, int X, int Y,
                                              double &dtc, int &d)
{
    double ax = __dsub_rn(__dmul_rn((double)(p1.vx < 0 ? X : 1 + X), b.csx), p1.x);
    double ay = __dsub_rn(__dmul_rn((double)(p1.vy < 0 ? Y : 1 + Y), b.csy), p1.y);
    if (WRAP) {
        ax = min_image(ax, b.half_lx, b.lx);
        ay = min_image(ay, b.half_ly, b.ly);
    }
    const double qx = ax * rcp_seed(p1.vx);
    const double qy = ay * rcp_seed(p1.vy);
    const int hx = __double2hiint(qx), hy = __double2hiint(qy);
    bool takex;
    if (hi_suspicious(hx) || hi_suspicious(hy) || abs(hx - hy) <= kBandHi) {
        const double tx = __ddiv_rn(ax, p1.vx);
        const double ty = __ddiv_rn(ay, p1.vy);
        takex = tx < ty;  // strict: ties go to y
        dtc = takex ? tx : ty;
    } else {
        takex = qx < qy;
        dtc = __ddiv_rn(takex ? ax : ay, takex ? p1.vx : p1.vy);
    }
    d = takex ? (p1.vx < 0 ? 1 : 2) : (p1.vy < 0 ? 3 : 4);
}
