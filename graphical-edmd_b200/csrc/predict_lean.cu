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
//   k_resolve  one thread per PARTICLE ID: coalesced own state, screening result and
//              outputs; ONE gather (the winner's state), crossing + the exact pair
//              time, certifies or re-scans
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
enum { kPlStatus = 0, kPlY = 1, kPlRowEnd = 2, kPlWstart = 3, kPlSegLo = 4, kPlCum1 = 7, kPlBase = 8,
       kPlCum2 = 11 };
//   status  0 empty, 1 staged, 2 not staged (segments exceed the buffer)
//   SegLo[j] absolute slot of row j's segment;  Cum1 / Cum2 first scr[] index of rows 1 / 2;
//   Base[j]  row-local cell offset o  ->  scr index o + Base[j]

constexpr size_t kLeanSmem = sizeof(LeanBuf) * 2 * kLeanWarps;

// what k_screen leaves per PARTICLE ID: x = winner's particle id (kResNone: no
// candidate can collide), y = bits of the second-smallest lower bound (NaN when
// the smallest bound is not positive: never certifiable)
enum { kResNone = -1 };

struct ScreenArgs {
    LeanIndex g;
    edmd_dev_box b;
    double rad0;
    int max_chunks;
    int32_t *flags;
    int2 *res;
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// ---- chunk planning and staging ------------------------------------------------
// All lanes read the six segment bounds of the chunk (uniform addresses: one
// transaction each), then issue the copies one trip ahead:
// cp.async, 16 bytes each -- the screening half of the 32-byte records of the
// three row segments, packed back to back, and the windows of cell offsets.
// Lane 0 leaves the plan in shared memory for the trip that consumes the buffer.
__device__ __forceinline__ void plan_stage(const LeanIndex &g, LeanBuf *buf, const LeanChunk &m, int lane)
{
    const int Y = m.x, ca = m.y, cb = m.z;
    int status = 0;
    if (Y >= 0) {
        int Yr[3], sa[3], sb[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Yr[j] = row_wrap(Y - 1 + j, g.nl);
            const int32_t *o = g.off + (size_t)Yr[j] * g.ps;
            sa[j] = o[ca - 1];
            sb[j] = o[cb + 2];
        }
        const int wstart = (ca - 1) & ~3;
        int wlen = ((cb + 2 - wstart + 1) + 3) & ~3;   // columns wstart .. cb+2, rounded up to 4
        if (wstart + wlen > g.ps) wlen = g.ps - wstart;
        const int cum1 = sb[0] - sa[0], cum2 = cum1 + (sb[1] - sa[1]), cum3 = cum2 + (sb[2] - sa[2]);
        status = (cum3 > kLeanCap || wlen > kLeanOffW) ? 2 : 1;
        const int lo0 = Yr[0] * g.rowcap + sa[0], lo1 = Yr[1] * g.rowcap + sa[1],
                  lo2 = Yr[2] * g.rowcap + sa[2];
        if (lane == 0) {
            int4 *pl = reinterpret_cast<int4 *>(buf->plan);
            pl[0] = make_int4(status, Y, m.w, wstart);
            pl[1] = make_int4(lo0, lo1, lo2, cum1);
            pl[2] = make_int4(-sa[0], cum1 - sa[1], cum2 - sa[2], cum2);
        }
        if (status == 1) {
            // packed index q of scr[] -> record  q + (q < cum1 ? lo0 : q < cum2 ? lo1 - cum1 : lo2 - cum2)
            const int d1 = lo1 - cum1, d2 = lo2 - cum2;
            const uint32_t dst = smem_u32(&buf->scr[lane]);
#pragma unroll
            for (int t = 0; t < (kLeanCap + 31) / 32; t++) {
                const int q = lane + 32 * t;
                const int src = q + (q < cum1 ? lo0 : (q < cum2 ? d1 : d2));
                if (q < cum3)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512u * t),
                                 "l"(g.rec + src)
                                 : "memory");
            }
            // windows: 16-byte granules, rows 0 and 1 by the two half-warps, then row 2
            const int h = lane >> 4, k4 = 4 * (lane & 15);
            if (k4 < wlen) {
                const int32_t *wsrc = g.off + wstart + k4;
                cp_async16(&buf->offw[h][k4], wsrc + (size_t)(h ? Yr[1] : Yr[0]) * g.ps);
                if (h == 0) cp_async16(&buf->offw[2][k4], wsrc + (size_t)Yr[2] * g.ps);
            }
        }
    } else if (lane == 0) {
        buf->plan[kPlStatus] = 0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// ---- k_screen: FP32 lower bounds, two smallest per particle ----------------------
// STAGED: candidates and cell offsets come from the warp's shared buffer; else
// (segments exceed the buffer: very dense rows, clustered tiny disks) from
// global memory, same arithmetic.
// The result (winner's particle id, second bound) is handed back in `out`
// (out.id < 0: nothing to store) with the id gather still in flight: the caller
// stores it one trip later, so nobody waits for that load.
struct Pending {
    int id, wid, y;
};

// Store a particle's screening result (the winner-id gather issued one trip ago
// has landed by now).  Measured: prefetching the FP64 states k_resolve will read
// into L2 from here does not make k_resolve faster.
__device__ __forceinline__ void flush_pending(const ScreenArgs &a, Pending &pend)
{
    if (pend.id >= 0) a.res[pend.id] = make_int2(pend.wid, pend.y);
    pend.id = -1;
}

template <bool STAGED, bool TWO>
__device__ __forceinline__ void screen_chunk(const ScreenArgs &a, const LeanConsts &K, const LeanBuf &w,
                                             int c, int lane, int2 tag, Pending &out)
{
    const LeanIndex &g = a.g;
    const int pc = tag.y;
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
            const int rb = Yr * g.rowcap;
            const int32_t *o = g.off + (size_t)Yr * g.ps + (active ? pcx - 1 : 0);
            lo[j] = active ? rb + o[0] : 0;
            t1[j] = rb + o[1];
            t2[j] = rb + o[2];
            hi[j] = active ? rb + o[3] : 0;
        }
    }
    const float4 *scr = STAGED ? w.scr : nullptr;
    const int self = STAGED ? s - w.plan[kPlSegLo + 1] + w.plan[kPlCum1] : s;
    const int selfc = active ? self : (STAGED ? 0 : c * 32);
    const float4 own = STAGED ? scr[selfc] : *reinterpret_cast<const float4 *>(g.rec + selfc);
    // dx = rx_j - (rx_i - k csx), k = column(j) - column(i) in {-1, 0, 1}
    const float pxm = __fadd_rn(own.x, K.csx), px0 = own.x, pxp = __fsub_rn(own.x, K.csx);
    // contact distance squared (+ slack) against a candidate of class 0 / class 1
    const bool own1 = TWO && (__float_as_int(own.w) & 1);
    const float cc_a = own1 ? K.Cc01 : K.Cc00, cc_b = own1 ? K.Cc11 : K.Cc01;

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
        const float clo = __fmaf_rn(d2, K.A, TWO ? ((__float_as_int(q.w) & 1) ? -cc_b : -cc_a) : -cc_a);
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
    out.id = active ? tag.x : -1;
    out.wid = kResNone;
    out.y = __float_as_int(lo1 > 0.0f ? lo2 : fnan);
    if (active && idx >= 0) {
        int wslot = idx;
        if (STAGED) {
            const int cum1 = w.plan[kPlCum1], cum2 = w.plan[kPlCum2];
            wslot = idx >= cum2 ? w.plan[kPlSegLo + 2] + (idx - cum2)
                                : (idx >= cum1 ? w.plan[kPlSegLo + 1] + (idx - cum1) : w.plan[kPlSegLo] + idx);
        }
        out.wid = g.rec[wslot].id;
    }
}

template <bool TWO>
__device__ __forceinline__ void screen_main(const ScreenArgs &a, const LeanConsts &K, unsigned char *lean_smem)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    LeanBuf *bufs = reinterpret_cast<LeanBuf *>(lean_smem) + 2 * warp;
    const int stride = gridDim.x * kLeanWarps;
    Pending pend;
    pend.id = -1; pend.wid = 0; pend.y = 0;
    const LeanChunk none = make_int4(-1, 0, 0, 0);
    // Work list: the chunks that hold particles, in any order (rows own fixed slot
    // ranges, most chunk slots are empty).  Lane l fetches entry w + (32 t + l) stride
    // of the warp and its plan word; a ballot tells which lanes hold real work.
    const int nwork = a.flags[kFlagWork];
    for (int k0 = blockIdx.x * kLeanWarps + warp; k0 < nwork; k0 += 32 * stride) {
        const int kl = k0 + lane * stride;
        const int cl = kl < nwork ? a.g.work[kl] : 0;
        const LeanChunk ml = kl < nwork ? a.g.chunks[cl] : none;
        unsigned todo = __ballot_sync(0xffffffffu, ml.x >= 0);
        if (todo == 0) continue;
        auto word = [&](int b) {
            return make_int4(__shfl_sync(0xffffffffu, ml.x, b), __shfl_sync(0xffffffffu, ml.y, b),
                             __shfl_sync(0xffffffffu, ml.z, b), __shfl_sync(0xffffffffu, ml.w, b));
        };
        int b = __ffs(todo) - 1;
        todo &= todo - 1;
        int k = 0;
        int c = __shfl_sync(0xffffffffu, cl, b);
        plan_stage(a.g, &bufs[0], word(b), lane);
        int2 tag = *reinterpret_cast<const int2 *>(&a.g.rec[c * 32 + lane].id);
        while (b >= 0) {
            // in flight during this trip: the copies and the own tags of the next chunk
            const int bn = todo ? __ffs(todo) - 1 : -1;
            todo &= todo - 1;
            const int cn = __shfl_sync(0xffffffffu, cl, bn >= 0 ? bn : 0);
            plan_stage(a.g, &bufs[k ^ 1], bn >= 0 ? word(bn) : none, lane);
            const int2 tagn = bn >= 0 ? *reinterpret_cast<const int2 *>(&a.g.rec[cn * 32 + lane].id)
                                      : make_int2(0, 0);
            flush_pending(a, pend);   // last trip's result
            asm volatile("cp.async.wait_group 1;" ::: "memory");   // this chunk's copies have landed
            __syncwarp();
            const LeanBuf &buf = bufs[k];
            const int status = buf.plan[kPlStatus];
            if (status == 1) screen_chunk<true, TWO>(a, K, buf, c, lane, tag, pend);
            else if (status == 2) screen_chunk<false, TWO>(a, K, buf, c, lane, tag, pend);
            __syncwarp();   // every lane is done with the buffer before it is refilled
            tag = tagn;
            b = bn;
            c = cn;
            k ^= 1;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    flush_pending(a, pend);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(kLeanThreads, kLeanCtas)
k_screen(const __grid_constant__ ScreenArgs a)
{
    extern __shared__ __align__(128) unsigned char lean_smem[];
    edmd_pdl_wait();
    // the state must be eligible (the host checked what it knows; halo particles
    // arrive on the device): otherwise decline and let the host take the full path
    const int classes = a.flags[kFlagNotMono];   // 0: one radius, 1: two, more: not eligible
    const double rad1 = __longlong_as_double(*reinterpret_cast<const long long *>(a.flags + kFlagRad1));
    const LeanConsts K = make_consts(a.b, a.rad0, rad1, classes == 1, __int_as_float(a.flags[kFlagVmax]));
    if (a.flags[kFlagLeanFail] != 0) return;   // the index declined (a row denser than its slot range)
    if (a.flags[kFlagInsane] != 0 || classes > 1 || !K.ok) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.flags[kFlagLeanFail] = 1;
        return;
    }
    // the class bit is only looked at when a second class exists: one class with spread radii (a
    // reference-grown monodisperse system) screens like one radius, with the inflated constant
    if (classes == 1 && rad1 > 0.0) screen_main<true>(a, K, lean_smem);
    else screen_main<false>(a, K, lean_smem);
}

// ---- k_resolve: exact evaluation of the winner, one thread per particle -------------
struct ResolveArgs {
    edmd_dev_box b;
    LeanIndex g;
    double t, rad0;
    int n_owned;
    const double4 *xv;
    const double *rad;
    const int32_t *cid;
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
__device__ __forceinline__ void exact_scan_lean(const ResolveArgs &a, bool two, const SRec &p1, double four_r1,
                                                int lo, int hi, double &best, int &best_id,
                                                int &best_pc, int &ov_id, int &ov_pc)
{
#pragma unroll 1
    for (int p = lo; p < hi; p++) {
        const int2 tag = *reinterpret_cast<const int2 *>(&a.g.rec[p].id);
        if (tag.x == p1.id) continue;  // `p1 != p2` is identity (ghost copies included)
        const double4 q = ld_sector(a.xv + tag.x);
        SRec p2;
        p2.x = q.x; p2.y = q.y; p2.vx = q.z; p2.vy = q.w;
        p2.rad = two ? a.rad[tag.x] : a.rad0; p2.id = tag.x; p2.pc = tag.y;
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

constexpr int kResolveThreads = 128;

__global__ void __launch_bounds__(kResolveThreads)
k_resolve(const __grid_constant__ ResolveArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    edmd_pdl_wait();
    if (i >= a.n_owned) return;
    // all first-level loads go out together (the decline flag is one of them)
    const int declined = a.flags[kFlagLeanFail];
    const int2 r = a.res[i];
    const int pc = a.cid[i];
    const double4 me = ld_sector(a.xv + i);
    const bool two = a.flags[kFlagNotMono] == 1;   // two radius classes: radii come from the resident array
    const double rad_i = two ? a.rad[i] : a.rad0;
    if (declined != 0) return;   // k_screen (or the index) declined: res[] is stale
    double4 wq = make_double4(0, 0, 0, 0);
    double rad_w = a.rad0;
    if (r.x >= 0) {
        wq = ld_sector(a.xv + r.x);
        if (two) rad_w = a.rad[r.x];
    }
    const int Y = pc / a.g.ps;
    const int pcx = pc - Y * a.g.ps;
    SRec p1;
    p1.x = me.x; p1.y = me.y; p1.vx = me.z; p1.vy = me.w;
    p1.rad = rad_i; p1.id = i; p1.pc = pc;
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
        p2.rad = rad_w; p2.id = r.x; p2.pc = 0;
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
            const int rb = Yr * a.g.rowcap;
            const int32_t *o = a.g.off + (size_t)Yr * a.g.ps;
            exact_scan_lean(a, two, p1, four_r1, rb + o[pcx - 1], rb + o[pcx + 2], best, best_id, best_pc,
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
           c->dbox.nx >= 12 && c->dbox.ny >= 12 && c->dbox.nl >= 3 && c->rowcap > 0;
}

static LeanIndex lean_index_of(const edmd_ctx *c)
{
    LeanIndex g;
    g.nx = c->dbox.nx; g.nl = c->dbox.nl; g.ps = c->ps;
    g.rowcap = c->rowcap;
    g.off = c->off; g.chunks = c->lchunks; g.work = c->lwork;
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
    sa.max_chunks = c->lean_chunks;
    sa.flags = c->flags;
    sa.res = c->lres;
    static unsigned long long attr = 0;   // devices of this process the attributes are set on
    if (edmd_first_on_device(&attr)) {
        cudaFuncSetAttribute(k_screen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLeanSmem);
        cudaFuncSetAttribute(k_screen, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
    const int blocks = (sa.max_chunks + kLeanWarps - 1) / kLeanWarps;
    const int grid = (c->sm_count > 0 ? c->sm_count : 148) * kLeanCtas;
    edmd_launch(k_screen, dim3(blocks < grid ? blocks : grid), dim3(kLeanThreads), kLeanSmem, c->stream, c->lean_pdl, sa);
    int launched = 1;
    if (c->n_owned > 0) {
        ResolveArgs ra;
        ra.b = c->dbox;
        ra.g = sa.g;
        ra.t = c->t;
        ra.rad0 = c->rad0;
        ra.n_owned = c->n_owned;
        ra.xv = c->xv; ra.rad = c->rad; ra.cid = c->cid; ra.res = c->lres;
        ra.gid = c->slab ? c->gid : nullptr;
        ra.flags = c->flags;
        ra.t_cross = c->t_cross; ra.dir = c->dir; ra.t_coll = c->t_coll; ra.partner = c->partner;
        ra.ctype = c->ctype;
        ra.overlap_key = c->overlap_key;
        edmd_launch(k_resolve, dim3((c->n_owned + kResolveThreads - 1) / kResolveThreads), dim3(kResolveThreads), 0,
                    c->stream, c->lean_pdl, ra);
        launched++;
    }
    return launched;
}
