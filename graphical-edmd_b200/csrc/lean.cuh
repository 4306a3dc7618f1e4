// lean.cuh -- data layout of the LEAN cell index and of the certified-FP32
// prediction sweep built on it (lean_index.cu, predict_lean.cu).
//
// The lean path is taken for the re-predict sweep in NORMAL mode of a system with
// ONE or TWO radius classes (every BASELINE.json configuration, including the
// reference's default bidisperse run; general polydisperse systems keep the FP64
// row kernel of predict.cu).  A class is a value and everything within 1e-9 of it
// (edmd_note_radius, edmd_internal.cuh): the reference's growth phase leaves ~10
// distinct radii per species within 1e-15 of each other.  The screening uses the
// class radius inflated by that tolerance; the exact stage reads each disk's own
// FP64 radius unless all radii are exactly equal.  With two classes the class of a particle rides in the
// least significant mantissa bit of its FP32 vy (one more ulp in the error model).
// It moves fewer bytes and issues fewer instructions than the full path:
//
//   K0-lean  count (RED atomics, no ranks) -> ONE index kernel (row scan + chunk
//            plan; every cell row owns a FIXED range of `rowcap` slots, so rows
//            are independent: no scan across rows, no second pass)
//            -> scatter of ONE 32-byte record per particle (= one L2 sector, one
//            256-bit store): a 16-byte FP32 "screening record" + (id, cell) tag;
//            the FP64 state is NOT copied, it is read by particle id where
//            needed (own particle: coalesced; the winner: one gather).
//   K1-lean  per candidate, FP32 arithmetic produces a RIGOROUS LOWER BOUND of
//            the reference's collision time (collisionTimeNormal,
//            src/EDMD.c:2661-2723); the candidate with the smallest bound is
//            evaluated in FP64 exactly as the reference does, and accepted only
//            if its exact time lies below every other candidate's bound.  If
//            not (near ties, near-contact pairs, overlaps) the particle is
//            redone by the exact FP64 loop in the reference's order.  Results
//            are therefore the reference's, bit for bit; FP32 only decides
//            which ONE pair gets the exact evaluation.
//
// Screening record of a particle filed under padded cell column pcx of row Y:
//     rx = (float)(x - (pcx - 1 + 0.5) * csx)     cell-centre relative
//     ry = (float)(y - (Yglobal + 0.5) * csy)
//     vx, vy = (float) velocities
// Ghost copies (padded columns 0 and nx+1) carry the SAME rx, so the
// displacement between a particle in column a and a candidate in column b is
// (rx_b - rx_a) + (b - a) * csx with no periodic image logic at all; rows
// likewise.  (Needs nx, ny >= 12 and every particle within 1.5 cells of the
// centre of the cell it is filed under, which upload / free flight check.)
#pragma once

#include "edmd_internal.cuh"

constexpr int kLeanCap = 144;   // screening records staged per chunk (three row segments, packed)
constexpr int kLeanOffW = 64;   // cell-offset window per row

// one particle in cell order: exactly one 32-byte sector
struct __align__(32) LeanRec {
    float rx, ry, vx, vy;   // screening record (the half that is staged in shared memory)
    int id, pc;             // particle id, padded cell id
    int pad0, pad1;
};
static_assert(sizeof(LeanRec) == 32, "LeanRec must be one sector");

// per 32-slot chunk: x = cell row (-1: nothing to predict), y = first / z = last
// padded column holding active entries of the chunk, w = end slot of the row
typedef int4 LeanChunk;

struct LeanIndex {
    int nx, nl, ps;
    int rowcap;                // slots per cell row (multiple of 32): row Y starts at slot Y * rowcap
    const int32_t *off;        // row-local exclusive scan, [nl * ps]
    const LeanChunk *chunks;
    const int32_t *work;       // ids of the chunks that hold particles, any order
    const LeanRec *rec;        // cell order
};

// constants of the FP32 error analysis (see predict_lean.cu)
struct LeanConsts {
    float csx, csy;
    float inv_rho2, inv_om2;   // Psi = d2 * inv_rho2 + v2 * inv_om2 + 1
    float A;                   // c_lo = d2 * A - Cc
    float Cc00, Cc01, Cc11;    // Cc per pair of radius classes (all equal when there is one radius)
    float Kb, Kdet;            // B_up = Kb * Psi - b ; det_up = det + Kdet * Psi^2
    int ok;                    // 0: velocity scale outside the FP32-safe range
};

// one 32-byte sector with a single 256-bit load (LDG.E.ENL2.256 on sm_100a)
__device__ __forceinline__ double4 ld_sector(const double4 *p)
{
    double4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

#ifdef __CUDACC__
// ---- constants of the certified FP32 screening (error model: top of predict_lean.cu) ----
__device__ __forceinline__ LeanConsts make_consts(const edmd_dev_box &b, double rad0, double rad1, bool two,
                                                   float vmaxf)
{
    LeanConsts K;
    const double V = (double)vmaxf;
    K.ok = (V >= 1e-12 && V <= 1e12) ? 1 : 0;   // NaN fails too
    const double u = 5.9604644775390625e-08;    // 2^-24
    const double cs = fmax(b.csx, b.csy);
    const double Rm = 1.5 * cs, rho = cs, om = 0.5 * V;
    // two classes: the class radii, inflated by the class tolerance (every disk of a class is at most
    // that large: Cc stays an upper bound of 4 r_i r_j); no second class seen: class 1 is empty
    if (two) {
        rad1 = (rad1 > 0.0 ? rad1 : rad0) * (1.0 + kRadClassTol);
        rad0 = rad0 * (1.0 + kRadClassTol);
    }
    const double rmax = two ? fmax(rad0, rad1) : rad0;
    const double s2 = 4.0 * rmax * rmax;        // the error terms take the largest contact distance
    // two radii: the class bit in the last mantissa bit of vy costs one more ulp per particle
    const double ed = 8.0 * Rm, ew = (two ? 8.0 : 4.0) * V;   // e_d / u, e_w / u
    const double kb = 0.7072 * (rho * ew + om * ed) + rho * om;
    const double kv = 1.4143 * om * ew + 2.0 * om * om;
    const double kdet = rho * om * kb + (rho * rho + s2) * kv + om * om * (4.0 * rho * rho + s2);
    const double k4 = 1.5 * u * (1.4143 * ed / rho + 4.0);
    const double k5 = 1.5 * u * (1.4143 * ed * rho) + 4.0 * u * s2 + 1e-10;
    K.csx = __double2float_rn(b.csx);
    K.csy = __double2float_rn(b.csy);
    K.inv_rho2 = __double2float_rn(1.0 / (rho * rho));
    K.inv_om2 = __double2float_rn(1.0 / (om * om));
    K.A = __double2float_rd(1.0 - k4);
    // 4 r1 r2 of the reference (src/EDMD.c:2700) per pair of classes
    const double r1 = two ? rad1 : rad0;
    K.Cc00 = __double2float_ru(4.0 * rad0 * rad0 + k5);
    K.Cc01 = __double2float_ru(4.0 * rad0 * r1 + k5);
    K.Cc11 = __double2float_ru(4.0 * r1 * r1 + k5);
    K.Kb = __double2float_ru(1.5 * u * kb + 1.9073486328125e-06 * rho * om);   // + 2^-19 rho om
    K.Kdet = __double2float_ru(1.5 * u * kdet);
    return K;
}

__device__ __forceinline__ float rsqrt_f32(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float rcp_f32(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#endif
