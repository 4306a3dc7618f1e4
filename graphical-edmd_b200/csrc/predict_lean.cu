// predict_lean.cu -- K1-lean: the whole-system prediction sweep (NORMAL mode,
// monodisperse) with FP32 screening certified against the reference's FP64
// arithmetic.  Replaces, like predict.cu,
//   crossingEventNormal     src/EDMD.c:2405-2482
//   collisionEventNormal    src/EDMD.c:2829-3102   (scan 2959-3003, select 2991-2995)
//   collisionTimeNormal     src/EDMD.c:2661-2723
// and produces the same bits.
//
// Why FP32 can be used at all.  The reference keeps, per particle, the strict
// minimum over ~8 candidates of  t = (-b - sqrt(det)) / v2  (= c / (sqrt(det) - b)),
// b = d.dv, v2 = |dv|^2, c = |d|^2 - 4 r1 r2, det = b^2 - v2 c, skipping pairs
// with b > 0 or det < 0.  Only the WINNER's time is stored.  So it suffices to
// find the winner and evaluate that one pair exactly.  For every candidate an
// FP32 computation yields
//     c_lo   <= c,   det_up >= det,   B_up >= -b          (rigorous bounds)
//     t_lo = c_lo / (sqrt(det_up) + B_up)  <=  t           (when c_lo > 0)
// and candidates with B_up <= 0 (surely receding) or det_up <= 0 (surely
// missing) are dropped exactly when the reference drops them.  Let w be the
// candidate with the smallest t_lo and T its time evaluated in FP64 exactly as
// the reference does.  If T is a real collision time, t_lo(w) > 0 and every
// other candidate has t_lo > T, then w is the reference's strict minimum and T
// its time: emitted.  Otherwise (two candidates too close to call, touching or
// overlapping pairs, grazing pairs) the particle is redone by the plain FP64
// loop in the reference's order with its tie rule.  EDMD_STAT_EXACT_RESCANS
// counts those.
//
// Two kernels, so that no warp ever waits on a chain of dependent gathers:
//   k_screen   persistent warps, one 32-slot chunk of the cell-ordered records per
//              trip; the candidates' 16-byte screening halves are staged in shared
//              memory by cp.async (double-buffered, planned three chunks ahead);
//              per slot it leaves (winner slot, second-smallest bound) -- FP32 only
//   k_resolve  one thread per PARTICLE ID: coalesced own state and outputs; gathers
//              the winner, evaluates crossing + the exact pair time, certifies or
//              re-scans
//
// Error model (u = 2^-24, all FP32 operations are explicit round-to-nearest
// mul / add / fma, MUFU rsqrt / rcp with relative error < 2^-21):
//     inputs     |rx - RX| <= u Rm      Rm = 1.5 max(csx, csy)  (upload / free flight check it)
//                |v32 - v| <= u V       V = largest |velocity component|
//     hence      |dx^ - dx| <= e_d = 8 u Rm,   |dv^ - dv| <= e_w = 4 u V
//     with scales rho (distance) and om (velocity), Psi = d2/rho^2 + v2/om^2 + 1 and
//     sqrt(d2) <= rho Psi/2, sqrt(v2) <= om Psi/2, sqrt(d2 v2) <= rho om Psi/2:
//     |b^ - b|     <= u kb Psi,      kb = (rho e_w + om e_d)/(sqrt2 u) + rho om
//     |v2^ - v2|   <= u kv Psi,      kv = sqrt2 om e_w/u + 2 om^2
//     d2 - d2^     <= u [(sqrt2 e_d/(u rho) + 2) d2^ + sqrt2 e_d rho/u]
//     det - det^'  <= u kdet Psi^2,  kdet = rho om kb + (rho^2 + s2) kv + om^2 (4 rho^2 + s2)
//                     (det^' = b^2^ - v2^ c_lo evaluated in FP32, s2 = 4 r^2)
// Every constant carries a further factor 1.5, and B_up an extra 2^-19 rho om Psi
// that absorbs the rounding of sqrt, the sum, the reciprocal and the product.
// The FP64 evaluation error of the reference's own formula (relative
// 2^-50 b^2/(v2 c) from the cancellation in -b - sqrt(det)) is below the slack
// in c_lo by nine orders of magnitude, so "t_lo > T" really orders the
// reference's FP64 values.
#include "lean.cuh"
#include "pairmath.cuh"
#include "rowstage.cuh"

namespace {

constexpr int kLeanThreads = 128;   // 4 independent warps per CTA
constexpr int kLeanWarps = kLeanThreads / 32;
constexpr int kLeanCtas = 8;        // per SM

// one staging buffer of a warp: the three row segments packed back to back
struct __align__(16) LeanBuf {
    float4 scr[kLeanCap];
    int offw[3][kLeanOffW];
    int plan[16];
};
// plan words
enum { kPlStatus = 0, kPlY = 1, kPlRowEnd = 2, kPlWstart = 3, kPlSegLo = 4, kPlBase = 7, kPlCum = 10 };
//   status  0 empty, 1 staged, 2 not staged (segments exceed the buffer)
//   SegLo[j] absolute slot of row j's segment;  Cum[j] its first index in scr[];
//   Base[j]  row-local cell offset o  ->  scr index o + Base[j]

constexpr size_t kLeanSmem = sizeof(LeanBuf) * 2 * kLeanWarps;

// what k_screen leaves per slot: x = winner's slot (kResNone: no candidate can
// collide), y = bits of the second-smallest lower bound (NaN when the smallest
// bound is not positive: never certifiable)
enum { kResNone = -1 };

struct ScreenArgs {
    LeanIndex g;
    edmd_dev_box b;
    double rad0;
    int max_chunks;
    int32_t *flags;
    int2 *res;
};

__device__ __forceinline__ LeanConsts make_consts(const edmd_dev_box &b, double rad0, float vmaxf)
{
    LeanConsts K;
    const double V = (double)vmaxf;
    K.ok = (V >= 1e-12 && V <= 1e12) ? 1 : 0;   // NaN fails too
    const double u = 5.9604644775390625e-08;    // 2^-24
    const double cs = fmax(b.csx, b.csy);
    const double Rm = 1.5 * cs, rho = cs, om = 0.5 * V;
    const double s2 = 4.0 * rad0 * rad0;
    const double ed = 8.0 * Rm, ew = 4.0 * V;   // e_d / u, e_w / u
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
    K.Cc = __double2float_ru(s2 + k5);
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

__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// ---- chunk planning and staging ------------------------------------------------
// A chunk's plan is made in steps so that no load is waited for: its LeanChunk
// word is fetched three trips ahead (lane k < 4 keeps component k), the row
// bases / segment bounds two trips ahead (lane j < 3 keeps row Y-1+j), and the
// copies (cp.async, 16 bytes each: the screening half of the 32-byte records,
// and the windows of cell offsets) are issued one trip ahead by all lanes.
struct RowPlan {
    int rb, sa, sb;   // lane j < 3: row base, first and end offset of the row's segment
};

__device__ __forceinline__ int plan_word(const LeanIndex &g, int c, int nchunks, int lane)
{
    return (c < nchunks && lane < 4) ? reinterpret_cast<const int *>(g.chunks + c)[lane] : -1;
}

__device__ __forceinline__ RowPlan plan_rows(const LeanIndex &g, int mw, int lane)
{
    const int Y = __shfl_sync(0xffffffffu, mw, 0);
    const int ca = __shfl_sync(0xffffffffu, mw, 1);
    const int cb = __shfl_sync(0xffffffffu, mw, 2);
    RowPlan r;
    r.rb = r.sa = r.sb = 0;
    if (Y >= 0 && lane < 3) {
        const int Yr = row_wrap(Y - 1 + lane, g.nl);
        r.rb = g.row_base[Yr];
        const int32_t *o = g.off + (size_t)Yr * g.ps;
        r.sa = o[ca - 1];
        r.sb = o[cb + 2];
    }
    return r;
}

__device__ __forceinline__ void plan_stage(const LeanIndex &g, LeanBuf *buf, int mw, const RowPlan &r,
                                           int lane)
{
    const int Y = __shfl_sync(0xffffffffu, mw, 0);
    const int ca = __shfl_sync(0xffffffffu, mw, 1);
    const int cb = __shfl_sync(0xffffffffu, mw, 2);
    const int row_end = __shfl_sync(0xffffffffu, mw, 3);
    int seg_lo[3], cum[4];
    cum[0] = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int rb = __shfl_sync(0xffffffffu, r.rb, j);
        const int sa = __shfl_sync(0xffffffffu, r.sa, j);
        const int sb = __shfl_sync(0xffffffffu, r.sb, j);
        seg_lo[j] = rb + sa;
        cum[j + 1] = cum[j] + (sb - sa);
        if (lane == 0) {
            buf->plan[kPlSegLo + j] = seg_lo[j];
            buf->plan[kPlBase + j] = cum[j] - sa;
            buf->plan[kPlCum + j] = cum[j];
        }
    }
    const int wstart = (ca - 1) & ~3;
    int wlen = ((cb + 2 - wstart + 1) + 3) & ~3;   // columns wstart .. cb+2, rounded up to 4
    if (wstart + wlen > g.ps) wlen = g.ps - wstart;
    const int status = Y < 0 ? 0 : ((cum[3] > kLeanCap || wlen > kLeanOffW) ? 2 : 1);
    if (lane == 0) {
        buf->plan[kPlStatus] = status;
        buf->plan[kPlY] = Y;
        buf->plan[kPlRowEnd] = row_end;
        buf->plan[kPlWstart] = wstart;
    }
    if (status == 1) {
#pragma unroll
        for (int t = 0; t < (kLeanCap + 31) / 32; t++) {
            const int q = lane + 32 * t;
            if (q < cum[3]) {
                const int j = (q >= cum[1]) + (q >= cum[2]);
                const int src = (j == 0 ? seg_lo[0] : (j == 1 ? seg_lo[1] : seg_lo[2])) -
                                (j == 0 ? 0 : (j == 1 ? cum[1] : cum[2])) + q;
                cp_async16(&buf->scr[q], g.rec + src);
            }
        }
#pragma unroll
        for (int t = 0; t < (3 * kLeanOffW / 4 + 31) / 32; t++) {
            const int q = lane + 32 * t;   // 16-byte granule
            const int j = q / (kLeanOffW / 4), k4 = 4 * (q % (kLeanOffW / 4));
            if (j < 3 && k4 < wlen) {
                const int Yr = row_wrap(Y - 1 + j, g.nl);
                cp_async16(&buf->offw[j][k4], g.off + (size_t)Yr * g.ps + wstart + k4);
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// ---- k_screen: FP32 lower bounds, two smallest per particle ----------------------
// STAGED: candidates and cell offsets come from the warp's shared buffer; else
// (segments exceed the buffer: very dense rows, clustered tiny disks) from
// global memory, same arithmetic.
template <bool STAGED>
__device__ __forceinline__ void screen_chunk(const ScreenArgs &a, const LeanConsts &K, const LeanBuf &w,
                                             int c, int lane, int pc)
{
    const LeanIndex &g = a.g;
    const int Y = w.plan[kPlY];
    const int s = c * 32 + lane;
    const bool valid = s < w.plan[kPlRowEnd];
    const int pcx = pc - Y * g.ps;
    const bool active = valid && pcx >= 1 && pcx <= g.nx;   // ghost entries are copies

    // candidate index ranges: scr[] indices when STAGED, absolute slots otherwise
    int lo[3], t1[3], t2[3], hi[3];
    const int wi = active ? pcx - 1 - w.plan[kPlWstart] : 0;   // window index of column pcx-1
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (STAGED) {
            const int base = w.plan[kPlBase + j];
            lo[j] = active ? w.offw[j][wi] + base : 0;
            t1[j] = w.offw[j][wi + 1] + base;
            t2[j] = w.offw[j][wi + 2] + base;
            hi[j] = active ? w.offw[j][wi + 3] + base : 0;
        } else {
            const int Yr = row_wrap(Y - 1 + j, g.nl);
            const int rb = g.row_base[Yr];
            const int32_t *o = g.off + (size_t)Yr * g.ps + (active ? pcx - 1 : 0);
            lo[j] = active ? rb + o[0] : 0;
            t1[j] = rb + o[1];
            t2[j] = rb + o[2];
            hi[j] = active ? rb + o[3] : 0;
        }
    }
    const float4 *scr = STAGED ? w.scr : nullptr;
    const int self = STAGED ? s - w.plan[kPlSegLo + 1] + w.plan[kPlCum + 1] : s;
    const int selfc = active ? self : (STAGED ? 0 : c * 32);
    const float4 own = STAGED ? scr[selfc] : *reinterpret_cast<const float4 *>(g.rec + selfc);
    // dx = rx_j - (rx_i - k csx), k = column(j) - column(i) in {-1, 0, 1}
    const float pxm = __fadd_rn(own.x, K.csx), px0 = own.x, pxp = __fsub_rn(own.x, K.csx);

    float lo1 = __int_as_float(0x7f800000), lo2 = __int_as_float(0x7f800000);   // two smallest bounds
    int idx = -1;
    const float fnan = __int_as_float(0x7fffffff);

    auto screen = [&](int j, int p, float py) {
        const float4 q = STAGED ? scr[p] : *reinterpret_cast<const float4 *>(g.rec + p);
        const float px = p < t1[j] ? pxm : (p < t2[j] ? px0 : pxp);
        const float dx = __fsub_rn(q.x, px), dy = __fsub_rn(q.y, py);
        const float dvx = __fsub_rn(q.z, own.z), dvy = __fsub_rn(q.w, own.w);
        const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
        const float v2 = __fmaf_rn(dvy, dvy, __fmul_rn(dvx, dvx));
        const float bb = __fmaf_rn(dy, dvy, __fmul_rn(dx, dvx));
        const float psi = __fmaf_rn(d2, K.inv_rho2, __fmaf_rn(v2, K.inv_om2, 1.0f));
        const float clo = __fmaf_rn(d2, K.A, -K.Cc);
        const float det = __fmaf_rn(-v2, clo, __fmul_rn(bb, bb));
        const float detu = __fmaf_rn(__fmul_rn(K.Kdet, psi), psi, det);
        const float bup = __fmaf_rn(K.Kb, psi, -bb);
        const float sq = __fmul_rn(detu, rsqrt_f32(detu));   // NaN when det_up <= 0: dropped below
        const float den = __fadd_rn(sq, bup);
        float tl = __fmul_rn(clo, rcp_f32(den));
        const bool keep = (bup > 0.0f) && !(j == 1 && p == self);
        tl = keep ? tl : fnan;
        // (lo1, lo2) <- two smallest of {lo1, lo2, tl}; min / max drop NaN operands
        idx = tl < lo1 ? p : idx;
        lo2 = fmaxf(lo1, fminf(lo2, tl));
        lo1 = fminf(lo1, tl);
    };
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float py = j == 0 ? __fadd_rn(own.y, K.csy) : (j == 1 ? own.y : __fsub_rn(own.y, K.csy));
        int p = lo[j];
        const int pe = hi[j];
#pragma unroll 1
        for (; p + 1 < pe; p += 2) {
            screen(j, p, py);
            screen(j, p + 1, py);
        }
        if (p < pe) screen(j, p, py);
    }
    if (active) {
        int2 r;
        r.x = idx;
        if (STAGED && idx >= 0) {
            const int wj = (idx >= w.plan[kPlCum + 1]) + (idx >= w.plan[kPlCum + 2]);
            r.x = w.plan[kPlSegLo + wj] + (idx - w.plan[kPlCum + wj]);
        }
        r.y = __float_as_int(lo1 > 0.0f ? lo2 : fnan);
        a.res[s] = r;
    }
}

__global__ void __launch_bounds__(kLeanThreads, kLeanCtas)
k_screen(const __grid_constant__ ScreenArgs a)
{
    extern __shared__ __align__(128) unsigned char lean_smem[];
    // the state must be eligible (the host checked what it knows; halo particles
    // arrive on the device): otherwise decline and let the host take the full path
    const LeanConsts K = make_consts(a.b, a.rad0, __int_as_float(a.flags[kFlagVmax]));
    if (a.flags[kFlagInsane] != 0 || a.flags[kFlagNotMono] != 0 || !K.ok) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.flags[kFlagLeanFail] = 1;
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    LeanBuf *bufs = reinterpret_cast<LeanBuf *>(lean_smem) + 2 * warp;
    const int nchunks = a.max_chunks;
    const int stride = gridDim.x * kLeanWarps;
    int c = blockIdx.x * kLeanWarps + warp;
    if (c >= nchunks) return;
    // prologue: chunk c is staged, c + stride has its rows planned, c + 2 stride its word fetched
    int mw1 = plan_word(a.g, c + stride, nchunks, lane);
    int mw2 = plan_word(a.g, c + 2 * stride, nchunks, lane);
    {
        const int mw0 = plan_word(a.g, c, nchunks, lane);
        const RowPlan r0 = plan_rows(a.g, mw0, lane);
        plan_stage(a.g, &bufs[0], mw0, r0, lane);
    }
    RowPlan r1 = plan_rows(a.g, mw1, lane);
    int pc = a.g.rec[c * 32 + lane].pc;
    for (int k = 0; c < nchunks; c += stride, k ^= 1) {
        // in flight during this trip: copies of chunk c + stride, rows of c + 2 stride,
        // word of c + 3 stride, own cell ids of c + stride
        plan_stage(a.g, &bufs[k ^ 1], mw1, r1, lane);
        const RowPlan r2 = plan_rows(a.g, mw2, lane);
        const int mw3 = plan_word(a.g, c + 3 * stride, nchunks, lane);
        const int cn = c + stride;
        const int pcn = cn < nchunks ? a.g.rec[cn * 32 + lane].pc : 0;
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // this chunk's copies have landed
        __syncwarp();
        const LeanBuf &buf = bufs[k];
        const int status = buf.plan[kPlStatus];
        if (status == 1) screen_chunk<true>(a, K, buf, c, lane, pc);
        else if (status == 2) screen_chunk<false>(a, K, buf, c, lane, pc);
        __syncwarp();   // every lane is done with the buffer before it is refilled
        mw1 = mw2;
        mw2 = mw3;
        r1 = r2;
        pc = pcn;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- k_resolve: exact evaluation of the winner, one thread per particle -------------
struct ResolveArgs {
    edmd_dev_box b;
    LeanIndex g;
    double t, rad0;
    int n_owned;
    const double4 *xv;
    const int32_t *cid;
    const int32_t *slot_of;
    const int2 *res;
    const int32_t *gid;
    int32_t *flags;
    double *t_cross;
    uint8_t *dir;
    double *t_coll;
    int32_t *partner;
    uint8_t *ctype;
    unsigned long long *overlap_key;
};

__device__ __forceinline__ int gid_of(const ResolveArgs &a, int id) { return a.gid ? a.gid[id] : id; }

// The plain FP64 loop over the candidate slots [lo, hi) of one row, straight
// from global memory: the reference's scan with its tie rule (see predict.cu).
__device__ __forceinline__ void exact_scan_lean(const ResolveArgs &a, const SRec &p1, double four_r1,
                                                int lo, int hi, double &best, int &best_id,
                                                int &best_pc, int &ov_id, int &ov_pc)
{
#pragma unroll 1
    for (int p = lo; p < hi; p++) {
        const int2 tag = *reinterpret_cast<const int2 *>(&a.g.rec[p].id);
        if (tag.x == p1.id) continue;  // `p1 != p2` is identity (ghost copies included)
        const double4 q = a.xv[tag.x];
        SRec p2;
        p2.x = q.x; p2.y = q.y; p2.vx = q.z; p2.vy = q.w;
        p2.rad = a.rad0; p2.id = tag.x; p2.pc = tag.y;
        bool ov = false;
        const double dt = pair_time_normal<true>(a.b, p1, four_r1, p2, ov);
        if (ov && (ov_id < 0 || (p2.pc == ov_pc && gid_of(a, p2.id) > gid_of(a, ov_id)))) {
            ov_id = p2.id;
            ov_pc = p2.pc;
        }
        if (best > dt || (best == dt && best_id >= 0 && p2.pc == best_pc &&
                          gid_of(a, p2.id) > gid_of(a, best_id))) {
            best = dt;
            best_id = p2.id;
            best_pc = p2.pc;
        }
    }
}

__global__ void __launch_bounds__(256)
k_resolve(const __grid_constant__ ResolveArgs a)
{
    if (a.flags[kFlagLeanFail] != 0) return;   // k_screen declined
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_owned) return;
    const int pc = a.cid[i];
    const double4 me = a.xv[i];
    const int2 r = a.res[a.slot_of[i]];
    int2 wtag = make_int2(0, 0);
    double4 wq = make_double4(0, 0, 0, 0);
    if (r.x >= 0) {
        wtag = *reinterpret_cast<const int2 *>(&a.g.rec[r.x].id);
        wq = a.xv[wtag.x];
    }
    const int Y = pc / a.g.ps;
    const int pcx = pc - Y * a.g.ps;
    SRec p1;
    p1.x = me.x; p1.y = me.y; p1.vx = me.z; p1.vy = me.w;
    p1.rad = a.rad0; p1.id = i; p1.pc = pc;
    const double four_r1 = __dmul_rn(4.0, p1.rad);
    {
        double dtc;
        int d;
        crossing_fast<true>(a.b, p1, pcx - 1, edmd_global_row(a.b, Y), dtc, d);
        a.t_cross[i] = __dadd_rn(a.t, dtc);
        a.dir[i] = (uint8_t)d;
    }
    double best = EDMD_NEVER;
    int best_id = -1, best_pc = -1, ov_id = -1, ov_pc = -1;
    bool certified = r.x == kResNone;   // no candidate can collide: partner 0 at t + 1e26
    if (r.x >= 0) {
        SRec p2;
        p2.x = wq.x; p2.y = wq.y; p2.vx = wq.z; p2.vy = wq.w;
        p2.rad = a.rad0; p2.id = wtag.x; p2.pc = wtag.y;
        double bb, v2, cc, b2, vc;
        pair_terms<true>(a.b, p1, four_r1, p2, bb, v2, cc, b2, vc);
        const double det = __dsub_rn(b2, vc);
        const double T = __ddiv_rn(__dsub_rn(-bb, __dsqrt_rn(det)), v2);
        // a real collision (the reference's branches) and every other candidate's lower
        // bound above the exact time (a NaN bound or time fails the test)
        certified = !(bb > 0) && (det >= 0) && ((double)__int_as_float(r.y) > T);
        best = T;
        best_id = p2.id;
    }
    if (!certified) {
        atomicAdd(reinterpret_cast<unsigned int *>(a.flags + kFlagRescans), 1u);
        best = EDMD_NEVER;
        best_id = -1;
#pragma unroll 1
        for (int j = 0; j < 3; j++) {
            const int Yr = row_wrap(Y - 1 + j, a.g.nl);
            const int rb = a.g.row_base[Yr];
            const int32_t *o = a.g.off + (size_t)Yr * a.g.ps;
            exact_scan_lean(a, p1, four_r1, rb + o[pcx - 1], rb + o[pcx + 2], best, best_id, best_pc,
                            ov_id, ov_pc);
        }
    }
    a.t_coll[i] = __dadd_rn(a.t, best);
    a.partner[i] = best_id >= 0 ? gid_of(a, best_id) : 0;
    a.ctype[i] = EDMD_EV_COLLISION;
    if (ov_id >= 0) {
        unsigned long long key = ((unsigned long long)(uint32_t)gid_of(a, i) << 32) |
                                 (uint32_t)gid_of(a, ov_id);
        atomicMin(a.overlap_key, key);
    }
}

}  // namespace

bool edmd_lean_eligible(const edmd_ctx *c, int mode)
{
    return mode == EDMD_MODE_NORMAL && !c->force_generic && !c->lean_off && c->lean_ok && c->n > 0 &&
           c->dbox.nx >= 12 && c->dbox.ny >= 12 && c->dbox.nl >= 3;
}

static LeanIndex lean_index_of(const edmd_ctx *c)
{
    LeanIndex g;
    g.nx = c->dbox.nx; g.nl = c->dbox.nl; g.ps = c->ps;
    g.off = c->off; g.row_base = c->row_base; g.chunks = c->lchunks;
    g.rec = c->lrec;
    return g;
}

int edmd_launch_predict_lean(edmd_ctx *c)
{
    if (c->n == 0) return 0;
    ScreenArgs sa;
    sa.g = lean_index_of(c);
    sa.b = c->dbox;
    sa.rad0 = c->rad0;
    sa.max_chunks = edmd_chunks_bound(c);
    sa.flags = c->flags;
    sa.res = c->lres;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_screen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLeanSmem);
        cudaFuncSetAttribute(k_screen, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr = true;
    }
    const int blocks = (sa.max_chunks + kLeanWarps - 1) / kLeanWarps;
    const int grid = (c->sm_count > 0 ? c->sm_count : 148) * kLeanCtas;
    k_screen<<<blocks < grid ? blocks : grid, kLeanThreads, kLeanSmem, c->stream>>>(sa);
    int launched = 1;
    if (c->n_owned > 0) {
        ResolveArgs ra;
        ra.b = c->dbox;
        ra.g = sa.g;
        ra.t = c->t;
        ra.rad0 = c->rad0;
        ra.n_owned = c->n_owned;
        ra.xv = c->xv; ra.cid = c->cid; ra.slot_of = c->rank; ra.res = c->lres;
        ra.gid = c->slab ? c->gid : nullptr;
        ra.flags = c->flags;
        ra.t_cross = c->t_cross; ra.dir = c->dir; ra.t_coll = c->t_coll; ra.partner = c->partner;
        ra.ctype = c->ctype;
        ra.overlap_key = c->overlap_key;
        k_resolve<<<(c->n_owned + 255) / 256, 256, 0, c->stream>>>(ra);
        launched++;
    }
    return launched;
}
