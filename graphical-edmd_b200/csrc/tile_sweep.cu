// tile_sweep.cu -- the whole-system prediction sweep (NORMAL mode, one or two
// radius classes) in TWO kernels.  Replaces, bit for bit,
//   cellListInit / addToCell   src/EDMD.c:1906-1920, 2053-2078   (the cell index)
//   crossingEventNormal        src/EDMD.c:2405-2482
//   collisionEventNormal       src/EDMD.c:2829-3102   (scan 2959-3003, select 2991-2995)
//   collisionTimeNormal        src/EDMD.c:2661-2723
//
// The cell grid is cut into TILES of kTX x kTY cells.  A tile and the one-cell
// ring around it (its FRAME, (kTX+2) x (kTY+2) cells) hold everything the
// reference's 3 x 3 scan of the tile's particles touches (PBCcellX/Y,
// src/EDMD.c:2110-2124: the ring of an edge tile is the opposite edge of the grid).
//
//   P1  k_tile_partition   one thread per particle (memory order = particle id):
//       the particle's 32-byte record -- FP32 screening half relative to the centre
//       of the cell it is filed under, particle id, cell inside the destination
//       frame -- is appended to the bucket of its tile, and, when it sits in an
//       edge cell of the tile, to the halo region of the neighbouring tiles'
//       buckets (13 % of the particles once, 0.4 % three times).  One returning
//       atomic per append on a per-tile cursor (one 128-byte line per tile: the
//       cursors of ~2000 tiles stay hot in L2, unlike a per-cell histogram), one
//       full-sector 256-bit store per record.  No histogram pass, no scan, no
//       work list.
//   P2  k_tile_sweep       one CTA per tile: the bucket (contiguous, coalesced) is
//       binned by frame cell IN SHARED MEMORY (count -> scan -> place), then every
//       particle of the tile, in cell order, screens its 3 x 3 neighbourhood in
//       FP32 out of shared memory (the certified lower bounds of lean.cuh /
//       predict_lean.cu: same arithmetic, same error model), gathers the winner's
//       FP64 state by particle id, evaluates crossingEventNormal and the winner's
//       collisionTimeNormal exactly as the reference does, certifies the winner
//       against the second-smallest bound or re-scans the neighbourhood in FP64
//       in the reference's order, and writes the five outputs by particle id.
//
// Everything the old five-kernel chain kept in global memory between kernels
// (histogram, offsets, chunk plans, work list, cell-ordered records, per-particle
// screening results) lives in the shared memory of one CTA here.
//
// The state must be eligible exactly as for the lean sweep (lean.cuh); a bucket
// that overflows its fixed capacity (clustered tiny disks) makes the sweep DECLINE
// through kFlagLeanFail and the host re-runs it on the full FP64 path.
#include "lean.cuh"
#include "pairmath.cuh"

namespace {

constexpr int kPartThreads = 256;

struct PartArgs {
    int first, n, ps, slab;
    TileGeom tg;
    edmd_dev_box b;
    const int32_t *cid;
    const double4 *xv;
    const double *rad;
    double rad0;
    int32_t *flags;
    int32_t *tcnt;
    LeanRec *trec;
};

// one full 32-byte sector with a single 256-bit store
__device__ __forceinline__ void put_rec(LeanRec *dst, const float4 &r, int id, int lcell)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst),
                 "r"(__float_as_int(r.x)), "r"(__float_as_int(r.y)), "r"(__float_as_int(r.z)),
                 "r"(__float_as_int(r.w)), "r"(id), "r"(lcell), "r"(0), "r"(0)
                 : "memory");
}

__global__ void __launch_bounds__(kPartThreads)
k_tile_partition(const __grid_constant__ PartArgs a)
{
    const int i = a.first + blockIdx.x * blockDim.x + threadIdx.x;
    edmd_pdl_wait();
    if (i >= a.first + a.n) return;
    const int pc = a.cid[i];
    const double4 p = ld_sector(a.xv + i);
    if (pc < 0) return;   // unused halo slot of a slab context
    const bool two = a.flags[kFlagNotMono] == 1;
    const TileGeom &tg = a.tg;
    const int Yl = pc / a.ps;
    const int X = pc - Yl * a.ps - 1;
    const int tx = X / kTX, ty = Yl / kTY;
    const int lx = X - tx * kTX, ly = Yl - ty * kTY;
    const int w = tx == tg.ntx - 1 ? tg.wlast : kTX, h = ty == tg.nty - 1 ? tg.hlast : kTY;
    // screening record, relative to the centre of the FILED cell (lean.cuh)
    float4 r;
    r.x = __double2float_rn(__dsub_rn(p.x, __dmul_rn((double)X + 0.5, a.b.csx)));
    r.y = __double2float_rn(__dsub_rn(p.y, __dmul_rn((double)edmd_global_row(a.b, Yl) + 0.5, a.b.csy)));
    r.z = __double2float_rn(p.z);
    r.w = __double2float_rn(p.w);
    if (two) r.w = __int_as_float((__float_as_int(r.w) & ~1) | (edmd_same_class(a.rad[i], a.rad0) ? 0 : 1));
    bool fail = false;
    {
        const int t = ty * tg.ntx + tx;
        const int k = atomicAdd(&a.tcnt[(size_t)t * kCntStride], 1);
        if (k < tg.cap_own) put_rec(a.trec + (size_t)t * tg.cap + k, r, i, (ly + 1) * kFW + lx + 1);
        else fail = true;
    }
    // edge cells also belong to the frames of the neighbouring tiles (periodic: the
    // neighbour of the last tile column is the first; a tile one cell wide feeds both sides)
    const bool ex0 = lx == 0, ex1 = lx == w - 1, ey0 = ly == 0, ey1 = ly == h - 1;
    if (ex0 || ex1 || ey0 || ey1) {
#pragma unroll
        for (int sy = -1; sy <= 1; sy++) {
#pragma unroll
            for (int sx = -1; sx <= 1; sx++) {
                if (sx == 0 && sy == 0) continue;
                const bool need = (sx < 0 ? ex0 : (sx > 0 ? ex1 : true)) && (sy < 0 ? ey0 : (sy > 0 ? ey1 : true));
                if (!need) continue;
                int tx2 = tx + sx, ty2 = ty + sy;
                if (tx2 < 0) tx2 = tg.ntx - 1;
                if (tx2 >= tg.ntx) tx2 = 0;
                if (ty2 < 0 || ty2 >= tg.nty) {
                    if (a.slab) continue;   // a slab's rows are not periodic: rows 0 and nl-1 ARE the halo
                    ty2 = ty2 < 0 ? tg.nty - 1 : 0;
                }
                const int w2 = tx2 == tg.ntx - 1 ? tg.wlast : kTX, h2 = ty2 == tg.nty - 1 ? tg.hlast : kTY;
                const int fx = sx < 0 ? w2 + 1 : (sx > 0 ? 0 : lx + 1);
                const int fy = sy < 0 ? h2 + 1 : (sy > 0 ? 0 : ly + 1);
                const int t2 = ty2 * tg.ntx + tx2;
                const int k = atomicAdd(&a.tcnt[(size_t)t2 * kCntStride + kCntHalo], 1);
                if (k < tg.cap_halo) put_rec(a.trec + (size_t)t2 * tg.cap + tg.cap_own + k, r, i, fy * kFW + fx);
                else fail = true;
            }
        }
    }
    if (fail) atomicOr(&a.flags[kFlagLeanFail], 1);   // a bucket is full: the sweep declines
}

// ---- P2 ---------------------------------------------------------------------------
struct SweepArgs {
    TileGeom tg;
    edmd_dev_box b;
    double t, rad0;
    int n_owned;
    const int32_t *tiles;   // tile list (nullptr: blockIdx.x is the tile)
    const double4 *xv;
    const double *rad;
    const int32_t *gid;
    int32_t *flags;
    int32_t *tcnt;
    const LeanRec *trec;
    double *t_cross;
    uint8_t *dir;
    double *t_coll;
    int32_t *partner;
    uint8_t *ctype;
    unsigned long long *overlap_key;
};

__device__ __forceinline__ int gid_of(const SweepArgs &a, int id) { return a.gid ? a.gid[id] : id; }

// shared memory of one CTA: [scr float4 x cap][id int x cap][lc u16 x cap (padded)][off int x kFC+1 ...]
struct TileSmem {
    float4 *scr;
    int *id;
    unsigned short *lc;
    int *off;
    int *rowpre;   // [kTY + 1] owned particles before frame row y+1
    int *misc;     // [0..3] warp sums, [4] n_own, [5] n_halo
};

__device__ __forceinline__ TileSmem carve(unsigned char *base, int cap)
{
    TileSmem s;
    s.scr = reinterpret_cast<float4 *>(base);
    s.id = reinterpret_cast<int *>(base + (size_t)cap * 16);
    s.lc = reinterpret_cast<unsigned short *>(base + (size_t)cap * 20);
    s.off = reinterpret_cast<int *>(base + (size_t)cap * 22 + ((cap & 1) ? 2 : 0));
    s.rowpre = s.off + kFC + 4;
    s.misc = s.rowpre + kTY + 4;
    return s;
}

template <bool TWO>
__device__ __forceinline__ void tile_main(const SweepArgs &a, const LeanConsts &K, const TileSmem &s, int tile,
                                          int n_own, int n_halo)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileGeom &tg = a.tg;
    const LeanRec *bucket = a.trec + (size_t)tile * tg.cap;
    const int R = n_own + n_halo;

    // ---- count: rank of every record inside its frame cell (shared-memory atomics) ----
    int keep[kTileK];   // frame cell | rank << 16
#pragma unroll
    for (int k = 0; k < kTileK; k++) {
        const int r = tid + k * kTileThreads;
        keep[k] = -1;
        if (r < R) {
            const int src = r < n_own ? r : tg.cap_own + (r - n_own);
            int lc = bucket[src].pc;
            lc = min(max(lc, 0), kFC - 1);
            keep[k] = lc | (atomicAdd(&s.off[lc], 1) << 16);
        }
    }
    __syncthreads();
    // ---- exclusive scan of the kFC counters, in place --------------------------------
    {
        constexpr int kCpt = (kFC + 1 + kTileThreads - 1) / kTileThreads;
        const int c0 = tid * kCpt;
        int v[kCpt], sum = 0;
#pragma unroll
        for (int q = 0; q < kCpt; q++) {
            v[q] = c0 + q < kFC ? s.off[c0 + q] : 0;
            sum += v[q];
        }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s.misc[warp] = incl;
        __syncthreads();
        int ex = incl - sum;
#pragma unroll
        for (int w = 0; w < kTileThreads / 32; w++)
            if (w < warp) ex += s.misc[w];
#pragma unroll
        for (int q = 0; q < kCpt; q++) {
            if (c0 + q <= kFC) s.off[c0 + q] = ex;
            ex += v[q];
        }
    }
    __syncthreads();
    // ---- place: records into cell order; owned particles per frame row ------------------
#pragma unroll
    for (int k = 0; k < kTileK; k++) {
        if (keep[k] < 0) continue;
        const int r = tid + k * kTileThreads;
        const int src = r < n_own ? r : tg.cap_own + (r - n_own);
        const int lc = keep[k] & 0xffff;
        const int pos = s.off[lc] + (keep[k] >> 16);
        const float4 q = *reinterpret_cast<const float4 *>(bucket + src);
        s.scr[pos] = q;
        s.id[pos] = bucket[src].id;
        s.lc[pos] = (unsigned short)lc;
    }
    if (warp == 0) {
        // owned particles of frame row y+1 = its cells 1 .. kTX (cells 0 and kTX+1 are the ring)
        int len = 0;
        if (lane < kTY) len = s.off[(lane + 1) * kFW + kTX + 1] - s.off[(lane + 1) * kFW + 1];
        int incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane <= kTY) s.rowpre[lane] = incl - len;   // lane == kTY: the total
    }
    __syncthreads();

    // ---- sweep: every owned particle of the tile, in cell order ---------------------------
    const int nown = s.rowpre[kTY];
    const int tyi = tile / tg.ntx, txi = tile - tyi * tg.ntx;
    const float fnan = __int_as_float(0x7fffffff);
#pragma unroll 1
    for (int k = tid; k < nown; k += kTileThreads) {
        // frame row of the k-th owned particle: largest y with rowpre[y] <= k
        int y = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1)
            if (y + step < kTY && s.rowpre[y + step] <= k) y += step;
        const int self = s.off[(y + 1) * kFW + 1] + (k - s.rowpre[y]);
        const int id = s.id[self];
        if (id >= a.n_owned) continue;   // halo copy from a neighbouring slab: never predicted
        const double4 me = ld_sector(a.xv + id);
        const double rad_i = TWO ? a.rad[id] : a.rad0;
        const int lc = s.lc[self];
        const int lcx = lc - (y + 1) * kFW;   // frame column 1 .. kTX

        int lo[3], t1[3], t2[3], hi[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int *o = s.off + (y + j) * kFW + lcx - 1;
            lo[j] = o[0]; t1[j] = o[1]; t2[j] = o[2]; hi[j] = o[3];
        }
        const float4 own = s.scr[self];
        // dx = rx_j - (rx_i - k csx), k = column(j) - column(i) in {-1, 0, 1}
        const float pxm = __fadd_rn(own.x, K.csx), px0 = own.x, pxp = __fsub_rn(own.x, K.csx);
        const bool own1 = TWO && (__float_as_int(own.w) & 1);
        const float cc_a = own1 ? K.Cc01 : K.Cc00, cc_b = own1 ? K.Cc11 : K.Cc01;
        float lo1 = __int_as_float(0x7f800000), lo2 = __int_as_float(0x7f800000);   // two smallest bounds
        int idx = -1;
        auto screen = [&](int j, int p, float py) {
            const float4 q = s.scr[p];
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
            const bool keepc = (bup > 0.0f) && (p != self);
            tl = keepc ? tl : fnan;
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
        const int wid = idx >= 0 ? s.id[idx] : -1;
        double4 wq = make_double4(0, 0, 0, 0);
        double rad_w = a.rad0;
        if (wid >= 0) {
            wq = ld_sector(a.xv + wid);
            if (TWO) rad_w = a.rad[wid];
        }
        const float second = lo1 > 0.0f ? lo2 : fnan;   // a non-positive smallest bound is never certifiable

        // ---- exact: crossing + the winner's pair time, as the reference computes them ----
        const int X = txi * kTX + lcx - 1, Yl = tyi * kTY + y;
        SRec p1;
        p1.x = me.x; p1.y = me.y; p1.vx = me.z; p1.vy = me.w;
        p1.rad = rad_i; p1.id = id; p1.pc = 0;
        const double four_r1 = __dmul_rn(4.0, p1.rad);
        {
            double dtc;
            int d;
            crossing_fast<true>(a.b, p1, X, edmd_global_row(a.b, Yl), dtc, d);
            a.t_cross[id] = __dadd_rn(a.t, dtc);
            a.dir[id] = (uint8_t)d;
        }
        double best = EDMD_NEVER;
        int best_id = -1, best_pc = -1, ov_id = -1, ov_pc = -1;
        bool certified = wid < 0;   // no candidate can collide: partner 0 at t + 1e26
        if (wid >= 0) {
            SRec p2;
            p2.x = wq.x; p2.y = wq.y; p2.vx = wq.z; p2.vy = wq.w;
            p2.rad = rad_w; p2.id = wid; p2.pc = 0;
            double bb, v2, cc, b2, vc;
            pair_terms<true>(a.b, p1, four_r1, p2, bb, v2, cc, b2, vc);
            const double det = __dsub_rn(b2, vc);
            const double T = __ddiv_rn(__dsub_rn(-bb, __dsqrt_rn(det)), v2);
            // a real collision (the reference's branches) and every other candidate's lower
            // bound above the exact time (a NaN bound or time fails the test)
            certified = !(bb > 0) && (det >= 0) && ((double)second > T);
            best = T;
            best_id = wid;
        }
        if (!certified) {
            // the plain FP64 loop in the reference's order with its tie rule (first in scan
            // order = earlier cell, then larger id = its linked-list order after cellListInit)
            atomicAdd(reinterpret_cast<unsigned int *>(a.flags + kFlagRescans), 1u);
            best = EDMD_NEVER;
            best_id = -1;
#pragma unroll 1
            for (int j = 0; j < 3; j++) {
#pragma unroll 1
                for (int p = lo[j]; p < hi[j]; p++) {
                    const int id2 = s.id[p];
                    if (id2 == id) continue;   // `p1 != p2` is identity (periodic copies included)
                    const double4 q = ld_sector(a.xv + id2);
                    SRec p2;
                    p2.x = q.x; p2.y = q.y; p2.vx = q.z; p2.vy = q.w;
                    p2.rad = TWO ? a.rad[id2] : a.rad0; p2.id = id2;
                    p2.pc = 3 * j + (p < t1[j] ? 0 : (p < t2[j] ? 1 : 2));   // scan-order cell
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
        }
        a.t_coll[id] = __dadd_rn(a.t, best);
        a.partner[id] = best_id >= 0 ? gid_of(a, best_id) : 0;
        a.ctype[id] = EDMD_EV_COLLISION;
        if (ov_id >= 0) {
            unsigned long long key = ((unsigned long long)(uint32_t)gid_of(a, id) << 32) |
                                     (uint32_t)gid_of(a, ov_id);
            atomicMin(a.overlap_key, key);
        }
    }
}

__global__ void __launch_bounds__(kTileThreads, kTileCtas)
k_tile_sweep(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char tile_smem[];
    const TileSmem s = carve(tile_smem, a.tg.cap);
    const int tid = threadIdx.x;
    // zero the cell counters while the previous kernel drains
    for (int c = tid; c <= kFC; c += kTileThreads) s.off[c] = 0;
    edmd_pdl_wait();
    const int tile = a.tiles ? a.tiles[blockIdx.x] : blockIdx.x;
    int32_t *cnt = a.tcnt + (size_t)tile * kCntStride;
    if (tid == 0) {
        s.misc[4] = min(cnt[0], a.tg.cap_own);
        s.misc[5] = min(cnt[kCntHalo], a.tg.cap_halo);
        // cursors back to zero for the next sweep (this CTA is their only reader)
        cnt[0] = 0;
        cnt[kCntHalo] = 0;
    }
    const int classes = a.flags[kFlagNotMono];   // 0: one radius, 1: two classes, more: not eligible
    const double rad1 = __longlong_as_double(*reinterpret_cast<const long long *>(a.flags + kFlagRad1));
    const LeanConsts K = make_consts(a.b, a.rad0, rad1, classes == 1, __int_as_float(a.flags[kFlagVmax]));
    const bool declined = a.flags[kFlagLeanFail] != 0;
    const bool bad = a.flags[kFlagInsane] != 0 || classes > 1 || !K.ok;
    __syncthreads();
    if (declined) return;   // a bucket overflowed: the host re-runs the sweep on the full path
    if (bad) {
        if (blockIdx.x == 0 && tid == 0) a.flags[kFlagLeanFail] = 1;
        return;
    }
    const int n_own = s.misc[4], n_halo = s.misc[5];
    __syncthreads();   // misc[] is reused by the scan
    if (classes == 1 && rad1 > 0.0) tile_main<true>(a, K, s, tile, n_own, n_halo);
    else tile_main<false>(a, K, s, tile, n_own, n_halo);
}

}  // namespace

bool edmd_tile_eligible(const edmd_ctx *c, int mode)
{
    return edmd_lean_eligible(c, mode) && !c->tile_off && c->trec != nullptr;
}

size_t edmd_tile_smem_bytes(const TileGeom &tg)
{
    return (size_t)tg.cap * 22 + 4 + sizeof(int) * (kFC + 4 + kTY + 4 + 8);
}

// geometry + capacities of the tile buckets for a context (nx x nl cells, n particles)
bool edmd_tile_geometry(int nx, int nl, size_t n, TileGeom *out)
{
    TileGeom tg;
    tg.ntx = (nx + kTX - 1) / kTX;
    tg.nty = (nl + kTY - 1) / kTY;
    tg.wlast = nx - (tg.ntx - 1) * kTX;
    tg.hlast = nl - (tg.nty - 1) * kTY;
    const double dens = (double)n / ((double)nx * (double)nl);   // particles per cell
    long long own = (long long)(1.35 * dens * kTX * kTY) + 96;
    long long halo = (long long)(1.6 * dens * (2 * (kTX + kTY) + 4)) + 64;
    own = (own + 31) & ~31ll;
    halo = (halo + 31) & ~31ll;
    if (own + halo > (long long)kTileK * kTileThreads) return false;   // denser than a CTA's registers provide for
    tg.cap_own = (int)own;
    tg.cap_halo = (int)halo;
    tg.cap = (int)(own + halo);
    if ((long long)tg.ntx * tg.nty * tg.cap >= (1ll << 31)) return false;
    *out = tg;
    return true;
}

int edmd_launch_tile_sweep(edmd_ctx *c, cudaEvent_t between)
{
    if (c->n == 0) return 0;
    const TileGeom &tg = c->tgeom;
    PartArgs pa;
    pa.first = 0; pa.n = c->n; pa.ps = c->ps; pa.slab = c->slab ? 1 : 0;
    pa.tg = tg; pa.b = c->dbox;
    pa.cid = c->cid; pa.xv = c->xv; pa.rad = c->rad; pa.rad0 = c->rad0;
    pa.flags = c->flags; pa.tcnt = c->tcnt; pa.trec = c->trec;
    // the first kernel of the chain is launched plainly: whatever precedes it on the stream completes first
    edmd_launch(k_tile_partition, dim3((c->n + kPartThreads - 1) / kPartThreads), dim3(kPartThreads), 0, c->stream,
                false, pa);
    if (between) cudaEventRecord(between, c->stream);
    SweepArgs sa;
    sa.tg = tg; sa.b = c->dbox; sa.t = c->t; sa.rad0 = c->rad0; sa.n_owned = c->n_owned;
    sa.tiles = nullptr;
    sa.xv = c->xv; sa.rad = c->rad; sa.gid = c->slab ? c->gid : nullptr;
    sa.flags = c->flags; sa.tcnt = c->tcnt; sa.trec = c->trec;
    sa.t_cross = c->t_cross; sa.dir = c->dir; sa.t_coll = c->t_coll; sa.partner = c->partner; sa.ctype = c->ctype;
    sa.overlap_key = c->overlap_key;
    const size_t smem = edmd_tile_smem_bytes(tg);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_tile_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(k_tile_sweep, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr = true;
    }
    edmd_launch(k_tile_sweep, dim3(tg.ntx * tg.nty), dim3(kTileThreads), smem, c->stream, c->lean_pdl, sa);
    return 2;
}
