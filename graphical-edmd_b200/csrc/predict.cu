// predict.cu -- K1: the whole-system prediction sweep, and K2: free flight.
//
// For every particle i this evaluates what the reference's
// `crossingEvent(i); collisionEvent(i);` evaluate
//   crossingEventNormal / crossingEventGrow      src/EDMD.c:2405-2482, 2343-2403
//   collisionEventNormal / collisionEventGrow    src/EDMD.c:2829-3102, 3104-3230
//   collisionTimeNormal / collisionTimeGrow      src/EDMD.c:2661-2723, 2598-2659
//   PBC, PBCinsideCellX/Y                        src/EDMD.c:5896-5936
// on a synchronous snapshot (lat2 == 0), in the reference's evaluation order
// and WITHOUT fused multiply-adds: every product and sum that the reference's
// (FMA-less x86-64) build rounds separately is an explicit __dmul_rn /
// __dadd_rn / __dsub_rn here; division and square root are IEEE
// correctly-rounded on both sides.  Event times are therefore bit-identical.
//
// Selection = the reference's running minimum with a strict `>` over rows
// j = -1..1, columns k = -1..1 and, inside a cell, its linked list (descending
// particle id after cellListInit).  A minimum does not depend on visiting
// order except for exact ties, so records inside a cell may sit in any order;
// exact ties are resolved explicitly by that rule (earlier cell wins, then the
// larger id).  NaN candidates lose; no candidate => partner 0 at t + 1e26
// (1e7 while growing).
//
// k_predict_rows (NORMAL mode): one warp per 32-slot chunk, neighbours staged
// in shared memory (rowstage.cuh).  Two-phase candidate selection: phase 1
// computes b, |dv|^2, c, det EXACTLY (they decide the reference's `b > 0`,
// `det < 0` and overlap branches) and ranks the survivors by an APPROXIMATE
// time c / (sqrt~(det) - b) built from the MUFU rsqrt/rcp seeds (relative error
// ~2^-20); phase 2 evaluates the reference's formula (-b - sqrt(det)) / v2 with
// IEEE sqrt and division for the winner only.  If another survivor lies within
// 2^-13 relative of the winner, or anything looks ill-conditioned (non-positive
// or non-finite estimate, cancellation in -b - sqrt(det)), the particle is
// re-done by the plain exact loop, so the result is always the reference's.
//
// k_predict_generic: the plain exact loop straight from global memory; used for
// GROW mode (once per run), as the overflow path and for cross-checking.
#include "edmd_internal.cuh"
#include "rowstage.cuh"
#include "pairmath.cuh"

namespace {

struct SweepArgs {
    edmd_dev_box b;
    CellIndex g;
    double t;
    int max_chunks;
    double *t_cross;
    uint8_t *dir;
    double *t_coll;
    int32_t *partner;
    uint8_t *ctype;
    unsigned long long *overlap_key;
    unsigned int *stats;  // particles resolved by the exact re-scan
};

// collisionTimeGrow, lat2 == 0
__device__ __forceinline__ double pair_time_grow(const edmd_dev_box &b, const SRec &p1, double vr1,
                                                 const SRec &p2, double vr2, bool &overlap)
{
    const double dvx = __dsub_rn(p2.vx, p1.vx);
    const double dvy = __dsub_rn(p2.vy, p1.vy);
    const double dvr = __dadd_rn(vr1, vr2);
    double dx = __dsub_rn(p2.x, p1.x);
    double dy = __dsub_rn(p2.y, p1.y);
    const double dr = __dsqrt_rn(__dmul_rn(__dmul_rn(4.0, p1.rad), p2.rad));
    dx = min_image(dx, b.half_lx, b.lx);
    dy = min_image(dy, b.half_ly, b.ly);
    const double bb = __dsub_rn(__dadd_rn(__dmul_rn(dx, dvx), __dmul_rn(dy, dvy)),
                                __dmul_rn(dvr, dr));
    const double v2 = __dadd_rn(__dmul_rn(dvx, dvx), __dmul_rn(dvy, dvy));
    const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const double a = __dsub_rn(v2, __dmul_rn(dvr, dvr));
    const double gap = __dsub_rn(d2, __dmul_rn(dr, dr));
    const double det = __dsub_rn(__dmul_rn(bb, bb), __dmul_rn(a, gap));
    if (det < 0) return EDMD_NEVER;
    const double sq = __dsqrt_rn(det);
    const double plus = __ddiv_rn(__dadd_rn(-bb, sq), a);
    const double minus = __ddiv_rn(__dsub_rn(-bb, sq), a);
    if (((minus > 0) && (plus > 0) && (minus < plus)) || ((minus > 0) && (plus < 0)))
        return minus;
    else if (((minus > 0) && (plus > 0.000000001) && (plus < minus)) ||
             ((minus < 0) && (plus > 0.000000001)))
        return plus;
    if (gap < -0.01) overlap = true;
    return EDMD_NEVER;
}

// local particle id -> the id the caller knows (slab contexts hold a subset)
__device__ __forceinline__ int global_id(const CellIndex &g, int id) { return g.gid ? g.gid[id] : id; }

// The reference's first-minimum-in-scan-order rule for a candidate that ties
// the current best exactly: only a larger id in the SAME cell comes earlier.
__device__ __forceinline__ bool tie_wins(const CellIndex &g, const SRec &cand, int best_pc, int best_id)
{
    return cand.pc == best_pc && global_id(g, cand.id) > global_id(g, best_id);
}

__device__ __forceinline__ void emit_collision(const SweepArgs &a, int id, double best, int best_id,
                                               int ov_id)
{
    a.t_coll[id] = __dadd_rn(a.t, best);
    a.partner[id] = best_id >= 0 ? global_id(a.g, best_id) : 0;
    a.ctype[id] = EDMD_EV_COLLISION;
    if (ov_id >= 0) {
        unsigned long long key = ((unsigned long long)(uint32_t)global_id(a.g, id) << 32) |
                                 (uint32_t)global_id(a.g, ov_id);
        atomicMin(a.overlap_key, key);
    }
}

// Exact loop over candidate records [lo, hi) of one row (shared or global).
template <bool GROW, bool WRAP>
__device__ __forceinline__ void exact_scan_range(const CellIndex &g, const edmd_dev_box &b, const SRec &p1, double four_r1,
                                                 double vr1, const SPos *pos, const SAux *aux,
                                                 const double *vrs, int lo, int hi, double &best,
                                                 int &best_id, int &best_pc, int &ov_id, int &ov_pc)
{
#pragma unroll 1
    for (int p = lo; p < hi; p++) {
        const SRec p2 = make_rec(pos[p], aux[p]);
        if (p2.id == p1.id) continue;  // `p1 != p2` is identity (ghost copies included)
        bool ov = false;
        double dt;
        if (GROW)
            dt = pair_time_grow(b, p1, vr1, p2, vrs[p], ov);
        else
            dt = pair_time_normal<WRAP>(b, p1, four_r1, p2, ov);
        if (ov && (ov_id < 0 || (p2.pc == ov_pc && global_id(g, p2.id) > global_id(g, ov_id)))) {
            ov_id = p2.id;
            ov_pc = p2.pc;
        }
        if (best > dt || (best == dt && best_id >= 0 && tie_wins(g, p2, best_pc, best_id))) {
            best = dt;
            best_id = p2.id;
            best_pc = p2.pc;
        }
    }
}

// One particle straight from the global cell-ordered records.
template <bool GROW>
__device__ void predict_one_global(const SweepArgs &a, int s, int Y, int pcx)
{
    const edmd_dev_box &b = a.b;
    const CellIndex &g = a.g;
    const SRec p1 = make_rec(g.spos[s], g.saux[s]);
    if (p1.id >= g.n_owned) return;   // halo copy of a neighbour slab's particle
    const double vr1 = GROW ? g.svr[s] : 0.0;
    const double four_r1 = __dmul_rn(4.0, p1.rad);
    double dtc;
    int d;
    crossing_exact<true>(b, p1, pcx - 1, edmd_global_row(b, Y), dtc, d);
    a.t_cross[p1.id] = __dadd_rn(a.t, dtc);
    a.dir[p1.id] = (uint8_t)d;

    double best = GROW ? 10000000.0 : EDMD_NEVER;
    int best_id = -1, best_pc = -1, ov_id = -1, ov_pc = -1;
#pragma unroll 1
    for (int j = 0; j < 3; j++) {
        const int Yr = row_wrap(Y - 1 + j, g.ny);
        const int rb = g.row_base[Yr];
        const int32_t *o = g.off + (size_t)Yr * g.ps;
        exact_scan_range<GROW, true>(g, b, p1, four_r1, vr1, g.spos, g.saux, g.svr, rb + o[pcx - 1],
                                     rb + o[pcx + 2], best, best_id, best_pc, ov_id, ov_pc);
    }
    emit_collision(a, p1.id, best, best_id, ov_id);
}

template <bool GROW>
__global__ void __launch_bounds__(kStageThreads)
k_predict_generic(const __grid_constant__ SweepArgs a)
{
    edmd_pdl_wait();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = s >> 5;
    if (chunk >= a.max_chunks) return;
    const int Y = a.g.meta[chunk].Y;
    if (Y < 0) return;
    if (s >= a.g.meta[chunk].row_end) return;
    const int pcx = a.g.saux[s].pc - Y * a.g.ps;
    if (pcx < 1 || pcx > a.g.nx) return;  // ghost entry
    predict_one_global<GROW>(a, s, Y, pcx);
}

// ---- the staged two-phase kernel -------------------------------------------
template <bool WRAP>
__device__ __forceinline__ void predict_one_staged(const SweepArgs &a, const StageBuf &w,
                                                   const RowLane &rl)
{
    const edmd_dev_box &b = a.b;
    const int self = kCapW + rl.self;   // flat index into pos[] / aux[]
    const SPos *pos = &w.pos[0][0];
    const SAux *aux = &w.aux[0][0];
    const SRec p1 = make_rec(pos[self], aux[self]);
    if (p1.id >= a.g.n_owned) return;   // halo copy of a neighbour slab's particle
    const double four_r1 = __dmul_rn(4.0, p1.rad);
    const int X = rl.pcx - 1, Y = edmd_global_row(b, rl.Y);

    {
        double dtc;
        int d;
        crossing_fast<WRAP>(b, p1, X, Y, dtc, d);
        a.t_cross[p1.id] = __dadd_rn(a.t, dtc);
        a.dir[p1.id] = (uint8_t)d;
    }

    // ---- collision, phase 1: exact filter + seed ranking ---------------------
    // Branch-free and two candidates per trip, so that two independent FP64
    // dependency chains are in flight per warp.  Estimates are NaN-free for
    // finite inputs (the 1e-300 guards), so anything odd -- an overlapping or
    // touching approaching pair (c <= 0), total cancellation in -b - sqrt(det)
    // (estimate forced to 0) -- surfaces as a non-positive winning estimate and
    // sends the particle to the exact re-scan, which also reports overlaps.
    double m1 = EDMD_NEVER;   // smallest estimate so far
    int hband = __double2hiint(EDMD_NEVER) - kBandHi;   // hi word of m1, minus the band
    int pbest = -1;
    bool amb = false;
    auto rank_one = [&](int pp, const SRec &p2, double bb, double c, double b2, double vc) {
        const double det = __dsub_rn(b2, vc);
        // reference: `if (b > 0) never` ... `if (det < 0) never`; `p1 != p2` is identity
        const bool cand = !(bb > 0) && (det >= 0) && (WRAP ? (p2.id != p1.id) : (pp != self));
        const double sq = det * rsqrt_seed(det + 1e-300);    // ~sqrt(det), 0 when det == 0
        double qd = c * rcp_seed((sq - bb) + 1e-300);        // ~ c / (sqrt(det) - b)
        // v2*c < 2^-33 b^2 (exponent compare): -b - sqrt(det) cancels, distrust the estimate
        if (__double2hiint(b2) - __double2hiint(vc) > (33 << 20)) qd = 0.0;
        // within 2^-13 of the best so far (either side)?
        amb |= cand && ((unsigned)(__double2hiint(qd) - hband) <= 2u * kBandHi);
        if (cand && qd < m1) {
            m1 = qd;
            hband = __double2hiint(qd) - kBandHi;
            pbest = pp;
        }
    };
#pragma unroll
    for (int j = 0; j < 3; j++) {
        int pp = j * kCapW + rl.lo[j];
        const int pe = j * kCapW + rl.hi[j];
#pragma unroll 1
        for (; pp + 1 < pe; pp += 2) {
            SRec pa, pb;   // the id is only needed where cells can repeat (WRAP)
            pa = make_rec(pos[pp], aux[pp]);
            pb = make_rec(pos[pp + 1], aux[pp + 1]);
            double bba, v2a, ca, b2a, vca, bbb, v2b, cb, b2b, vcb;
            pair_terms<WRAP>(b, p1, four_r1, pa, bba, v2a, ca, b2a, vca);
            pair_terms<WRAP>(b, p1, four_r1, pb, bbb, v2b, cb, b2b, vcb);
            rank_one(pp, pa, bba, ca, b2a, vca);
            rank_one(pp + 1, pb, bbb, cb, b2b, vcb);
        }
        if (pp < pe) {
            const SRec pa = make_rec(pos[pp], aux[pp]);
            double bba, v2a, ca, b2a, vca;
            pair_terms<WRAP>(b, p1, four_r1, pa, bba, v2a, ca, b2a, vca);
            rank_one(pp, pa, bba, ca, b2a, vca);
        }
    }
    // the winning estimate must be a positive, normal, finite number
    amb |= (pbest >= 0) && hi_suspicious(__double2hiint(m1));

    double best = EDMD_NEVER;
    int best_id = -1, best_pc = -1, ov_id = -1, ov_pc = -1;
    if (!amb) {
        // ---- phase 2: the reference's formula for the winner -------------------
        if (pbest >= 0) {
            const SRec p2 = make_rec(pos[pbest], aux[pbest]);
            double bb, v2, c, b2, vc;
            pair_terms<WRAP>(b, p1, four_r1, p2, bb, v2, c, b2, vc);
            const double dt = __ddiv_rn(__dsub_rn(-bb, __dsqrt_rn(__dsub_rn(b2, vc))), v2);
            if (best > dt) {
                best = dt;
                best_id = p2.id;
            }
        }
    } else {
        // ---- exact re-scan in reference order ---------------------------------
        atomicAdd(a.stats, 1u);
#pragma unroll
        for (int j = 0; j < 3; j++)
            exact_scan_range<false, WRAP>(a.g, b, p1, four_r1, 0.0, w.pos[j], w.aux[j], nullptr, rl.lo[j],
                                          rl.hi[j], best, best_id, best_pc, ov_id, ov_pc);
    }
    emit_collision(a, p1.id, best, best_id, ov_id);
}

__global__ void __launch_bounds__(kStageThreads, kStageCtasPerSm)
k_predict_rows(const __grid_constant__ SweepArgs a)
{
    edmd_pdl_wait();
    const bool sane = a.g.flags[kFlagInsane] == 0;
    row_pipeline(a.g, a.g.meta, a.max_chunks,
                 [&](const StageBuf &buf, const ChunkMeta &m, const RowLane &rl, int status) {
                     if (!rl.active) return;
                     if (status == 2) {  // segments do not fit the staging window
                         predict_one_global<false>(a, rl.s, rl.Y, rl.pcx);
                         return;
                     }
                     // interior chunk of a wide grid with every particle near its cell:
                     // no periodic image can be involved
                     if (sane && (m.flags & kMetaInterior)) predict_one_staged<false>(a, buf, rl);
                     else predict_one_staged<true>(a, buf, rl);
                 });
}

// ---- K2: batched free flight (freeFlyNormal / freeFlyGrow) -----------------
template <bool GROW>
__global__ void __launch_bounds__(256)
k_free_fly(int n, edmd_dev_box b, int ps, double dt, double4 *__restrict__ xv,
           double *__restrict__ rad, const double *__restrict__ vr,
           const int32_t *__restrict__ cid, int32_t *__restrict__ flags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int insane = 0;
    if (i < n) {
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
        // cells are host state and do not move with the particle: re-check that
        // it is still near the cell it is filed under
        const int pc = cid[i];
        const int l = pc / ps;
        const int X = pc - l * ps - 1;
        const int Y = edmd_global_row(b, l);
        insane = !(fabs(x - ((double)X + 0.5) * b.csx) <= 1.5 * b.csx) ||
                 !(fabs(y - ((double)Y + 0.5) * b.csy) <= 1.5 * b.csy);
    }
    insane = __reduce_add_sync(0xffffffffu, insane);
    if ((threadIdx.x & 31) == 0 && insane) atomicAdd(&flags[kFlagInsane], insane);
}

}  // namespace

int edmd_launch_predict(edmd_ctx *c, int mode)
{
    if (c->n == 0) return 0;
    SweepArgs a;
    a.b = c->dbox;
    a.g = edmd_cell_index(c);
    a.t = c->t;
    a.max_chunks = edmd_chunks_bound(c);
    a.t_cross = c->t_cross;
    a.dir = c->dir;
    a.t_coll = c->t_coll;
    a.partner = c->partner;
    a.ctype = c->ctype;
    a.overlap_key = c->overlap_key;
    a.stats = reinterpret_cast<unsigned int *>(c->flags + kFlagRescans);
    const int blocks = (a.max_chunks + kStageWarps - 1) / kStageWarps;
    if (mode == EDMD_MODE_GROW)
        edmd_launch(k_predict_generic<true>, dim3(blocks), dim3(kStageThreads), 0, c->stream, c->lean_pdl, a);
    else if (c->force_generic)
        edmd_launch(k_predict_generic<false>, dim3(blocks), dim3(kStageThreads), 0, c->stream, c->lean_pdl, a);
    else {
        static unsigned long long attr = 0;   // devices of this process the attributes are set on
        if (edmd_first_on_device(&attr)) {
            cudaFuncSetAttribute(k_predict_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStageSmem);
            cudaFuncSetAttribute(k_predict_rows, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        }
        edmd_launch(k_predict_rows, dim3(min(blocks, edmd_persistent_blocks(c))), dim3(kStageThreads), kStageSmem,
                    c->stream, c->lean_pdl, a);
    }
    return 1;
}

int edmd_launch_free_fly(edmd_ctx *c, int mode, double dt)
{
    int n = c->n;
    if (n == 0) return 0;
    int blocks = (n + 255) / 256;
    cudaMemsetAsync(c->flags + kFlagInsane, 0, sizeof(int32_t), c->stream);
    if (mode == EDMD_MODE_GROW)
        k_free_fly<true><<<blocks, 256, 0, c->stream>>>(n, c->dbox, c->ps, dt, c->xv, c->rad,
                                                        c->vr, c->cid, c->flags);
    else
        k_free_fly<false><<<blocks, 256, 0, c->stream>>>(n, c->dbox, c->ps, dt, c->xv, c->rad,
                                                         c->vr, c->cid, c->flags);
    return 1;
}
