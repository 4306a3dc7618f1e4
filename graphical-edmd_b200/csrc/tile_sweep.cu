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
//       the particle's record -- its FP64 state (x, y, vx, vy: one 32-byte sector, one
//       256-bit store) and a 16-byte tag (radius, particle id, cell inside the
//       destination frame) -- is appended to the bucket of its tile, and, when it sits in an
//       edge cell of the tile, to the buckets of the neighbouring tiles whose ring
//       holds that cell (19 % of the particles once, 0.6 % three times).  A bucket
//       is kFH RUNS, one per frame row, each with its own cursor in its own 32-byte
//       sector: one returning atomic per append, ~30 appends per cursor per sweep
//       (measured: ONE cursor per tile serialises its ~460 atomics at ~60 ns each,
//       31 us for the kernel).  No histogram pass, no scan, no work list.
//   P2  k_tile_sweep       one CTA per tile: the bucket (contiguous, coalesced) is
//       binned by frame cell IN SHARED MEMORY (count -> scan -> place: FP64 state,
//       the FP32 screening record derived from it, id), then every particle of the
//       tile, in cell order, screens its 3 x 3 neighbourhood in FP32 (the certified
//       lower bounds of lean.cuh / predict_lean.cu: same arithmetic, same error
//       model), evaluates crossingEventNormal and the winner's collisionTimeNormal
//       exactly as the reference does from the FP64 states in shared memory,
//       certifies the winner against the second-smallest bound or re-scans the
//       neighbourhood in FP64 in the reference's order, and writes ONE 32-byte event
//       record per particle (edmd_ev32: both event times, partner, direction) by
//       particle id -- the only access of the kernel that is not contiguous.
//       Measured on B200 (profiles/r2b_tile_phases.log): a random 32-byte gather or a
//       scattered partial-sector store costs ~12 us per million, more than the whole
//       screening arithmetic; hence the FP64 state travels with the record and the
//       outputs leave as one full sector.  k_unpack_events turns the records into
//       the ABI's five arrays when a caller fetches them.
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
#include "halo.cuh"

namespace {

constexpr int kPartThreads = 256;

struct PartArgs {
    int first, n, ps, slab, dbg;
    int late_wait;   // fused halo exchange: run BESIDE the preceding kernels of the chain, wait for them at the end
    TileGeom tg;
    edmd_dev_box b;
    const int32_t *cid;
    const double4 *xv;
    const double *rad;
    double rad0;
    int32_t *flags;
    int32_t *tcnt;
    double4 *tst;
    int2 *ttag;
    double *trad;
    unsigned long long *overlap_key;
    unsigned long long *ts;
};

// Append particle i (padded cell id pc, state p, radius rad) to the bucket of its tile and, when it sits
// in an edge cell of the tile, to the buckets of the neighbouring tiles whose ring holds that cell.
__device__ __forceinline__ void partition_one(const PartArgs &a, int i, int pc, const double4 &p, double rad,
                                              const bool radii)
{
    const TileGeom &tg = a.tg;
    const int Yl = pc / a.ps;
    const int X = pc - Yl * a.ps - 1;
    const int tx = X / kTX, ty = Yl / kTY;
    const int lx = X - tx * kTX, ly = Yl - ty * kTY;
    const int w = tx == tg.ntx - 1 ? tg.wlast : kTX, h = ty == tg.nty - 1 ? tg.hlast : kTY;
    // Destinations: the particle's own tile, and -- for a particle in an edge cell of its tile --
    // the neighbouring tiles whose ring holds that cell (periodic: the neighbour of the last tile
    // column is the first).  run = tile * kFH + frame row, cell = frame row * kFW + frame column.
    // All cursors are bumped before the first store: ONE atomic round trip per particle.
    int run[4], cell[4];
    run[0] = (ty * tg.ntx + tx) * kFH + ly + 1;
    cell[0] = (ly + 1) * kFW + lx + 1;
    run[1] = run[2] = run[3] = -1;
    const bool ex0 = lx == 0, ex1 = lx == w - 1, ey0 = ly == 0, ey1 = ly == h - 1;
    const bool degenerate = (ex0 && ex1) || (ey0 && ey1);   // a tile one cell wide / high feeds both sides
    auto target = [&](int sx, int sy, int &r, int &c) {
        int tx2 = tx + sx, ty2 = ty + sy;
        if (tx2 < 0) tx2 = tg.ntx - 1;
        if (tx2 >= tg.ntx) tx2 = 0;
        r = -1;
        if (ty2 < 0 || ty2 >= tg.nty) {
            if (a.slab) return;   // a slab's rows are not periodic: rows 0 and nl-1 ARE the halo
            ty2 = ty2 < 0 ? tg.nty - 1 : 0;
        }
        const int w2 = tx2 == tg.ntx - 1 ? tg.wlast : kTX, h2 = ty2 == tg.nty - 1 ? tg.hlast : kTY;
        const int fx = sx < 0 ? w2 + 1 : (sx > 0 ? 0 : lx + 1);
        const int fy = sy < 0 ? h2 + 1 : (sy > 0 ? 0 : ly + 1);
        r = (ty2 * tg.ntx + tx2) * kFH + fy;
        c = fy * kFW + fx;
    };
    if (!degenerate) {
        const int sx = ex0 ? -1 : 1, sy = ey0 ? -1 : 1;
        if (ex0 || ex1) target(sx, 0, run[1], cell[1]);
        if (ey0 || ey1) target(0, sy, run[2], cell[2]);
        if ((ex0 || ex1) && (ey0 || ey1)) target(sx, sy, run[3], cell[3]);
    }
    int k[4];
#pragma unroll
    for (int d = 0; d < 4; d++)
        k[d] = run[d] >= 0 ? ((a.dbg & 16) ? (i & 31) : atomicAdd(&a.tcnt[(size_t)run[d] * kCurStride], 1)) : 0;
    auto put = [&](int rr, int kk, int cc) {
        const size_t slot = (size_t)rr * kRunCap + kk;
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(a.tst + slot), "d"(p.x), "d"(p.y), "d"(p.z), "d"(p.w)
                     : "memory");
        a.ttag[slot] = make_int2(i, cc);
        if (radii) a.trad[slot] = rad;
    };
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        if (run[d] < 0) continue;
        if (a.dbg & 8) { ok &= k[d] != 12345; continue; }
        if (k[d] < kRunCap) put(run[d], k[d], cell[d]);
        else ok = false;
    }
    if (degenerate) {   // rare: the last tile column / row of the grid is one cell wide / high
        for (int sy = -1; sy <= 1; sy++)
            for (int sx = -1; sx <= 1; sx++) {
                if (sx == 0 && sy == 0) continue;
                if (!((sx < 0 ? ex0 : (sx > 0 ? ex1 : true)) && (sy < 0 ? ey0 : (sy > 0 ? ey1 : true)))) continue;
                int rr, cc;
                target(sx, sy, rr, cc);
                if (rr < 0) continue;
                const int kk = atomicAdd(&a.tcnt[(size_t)rr * kCurStride], 1);
                if (kk < kRunCap) put(rr, kk, cc);
                else ok = false;
            }
    }
    if (!ok) atomicOr(&a.flags[kFlagLeanFail], 1);   // a run is full: the sweep declines
}

__global__ void __launch_bounds__(kPartThreads)
k_tile_partition(const __grid_constant__ PartArgs a)
{
    const int i = a.first + blockIdx.x * blockDim.x + threadIdx.x;
    if (!a.late_wait) edmd_pdl_wait();
    if (i == a.first) edmd_stamp(a.ts, 4);
    if (i == a.first + a.n - 1) edmd_stamp(a.ts, 5);
    if (i == a.first && a.overlap_key) *a.overlap_key = ~0ull;   // the sweep's overlap report starts empty
    if (i < a.first + a.n) {
        const int pc = a.cid[i];
        const double4 p = ld_sector(a.xv + i);
        if (pc >= 0) {   // < 0: unused halo slot of a slab context
            // radii matter only when they are not all exactly rad0 (two classes / spread inside a class)
            const bool radii = a.flags[kFlagNotMono] != 0;
            const double rad = radii ? a.rad[i] : a.rad0;
            partition_one(a, i, pc, p, rad, radii);
        }
    }
    // fused halo exchange: the send and receive kernels run beside this one; "this kernel is complete"
    // must imply "they are" for the sweep kernel that waits on it.  ONE block waits -- the last one
    // dispatched: a grid is complete when all its blocks are.  (Every block waiting kept the first waves
    // resident until the receive kernel was through, 16 us, and the partition made no progress.)
    if (a.late_wait && blockIdx.x == gridDim.x - 1) edmd_pdl_wait();
    if (i == a.first + a.n - 1) edmd_stamp(a.ts, 6);
}

// Slab contexts, peer-to-peer halo: receive + partition in one kernel.  Waits for the neighbours'
// epoch (their k_halo_send wrote the records into my inbox over NVLink), unpacks each record into the
// fixed halo region of the resident arrays -- exactly what k_halo_recv (halo.cu) does -- and appends
// it to the tile buckets right away; acks back.  blockIdx.y = from (0: lower neighbour's records ->
// local row 0, 1: upper -> row nl-1).
struct RecvPartArgs {
    PartArgs p;
    int H, epoch, row[2];
    const char *inbox[2];
    double4 *xv;
    double *rad;
    int32_t *cid, *gid;
    int *peer_ack[2];
    int32_t *done;
};

__global__ void __launch_bounds__(kPartThreads)
k_halo_recv_partition(const __grid_constant__ RecvPartArgs a)
{
    __shared__ bool last;
    const int from = blockIdx.y;
    const InboxHeader *hdr = reinterpret_cast<const InboxHeader *>(a.inbox[from]);
    const HaloRec *rec = reinterpret_cast<const HaloRec *>(a.inbox[from] + sizeof(InboxHeader));
    edmd_pdl_trigger();   // the partition of the owned particles starts beside this kernel
    if (blockIdx.x == 0 && from == 0 && threadIdx.x == 0) edmd_stamp(a.p.ts, 2);
    if (threadIdx.x == 0)
        while (ld_volatile(&hdr->epoch) != a.epoch) __nanosleep(50);
    __syncthreads();
    const int count = ld_volatile(&hdr->count);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < a.H) {
        const int i = a.p.first + from * a.H + k;
        if (k < count) {
            const uint4 *src = reinterpret_cast<const uint4 *>(rec + k);
            HaloRec r;
            uint4 *dst = reinterpret_cast<uint4 *>(&r);
            dst[0] = __ldcv(src); dst[1] = __ldcv(src + 1); dst[2] = __ldcv(src + 2);
            const double4 p = make_double4(r.x, r.y, r.vx, r.vy);
            const int pc = a.row[from] * a.p.ps + r.cell;
            a.xv[i] = p;
            a.rad[i] = r.rad;
            a.gid[i] = r.gid;
            a.cid[i] = pc;
            // keep the sweep's eligibility facts current (lean.cuh)
            edmd_note_radius(a.p.flags, r.rad, a.p.rad0);
            float vm = __double2float_ru(fmax(fabs(r.vx), fabs(r.vy)));
            if (!(vm == vm)) vm = __int_as_float(0x7f800000);
            if (__float_as_int(vm) > a.p.flags[kFlagVmax])
                atomicMax(reinterpret_cast<unsigned int *>(&a.p.flags[kFlagVmax]), (unsigned)__float_as_int(vm));
            // the owned particles were partitioned with the radius facts of the upload; a halo disk of
            // another radius class raises kFlagNotMono now and the sweep kernel declines (rad_smem)
            partition_one(a.p, i, pc, p, r.rad, true);
        } else {
            a.cid[i] = -1;   // unused slot
            a.gid[i] = -1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(a.done + from, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile int *>(a.peer_ack[from]) = a.epoch;   // "consumed", peer store
        __threadfence_system();
        a.done[from] = 0;
        edmd_stamp(a.p.ts, 3);
    }
    // chained behind the send kernel: "this kernel is complete" implies "my own send is"
    edmd_pdl_wait();
}

// ---- P2 ---------------------------------------------------------------------------
struct SweepArgs {
    TileGeom tg;
    edmd_dev_box b;
    double t, rad0;
    int n_owned, dbg, rad_smem;
    const int32_t *tiles;   // tile list (nullptr: blockIdx.x is the tile)
    const int32_t *gid;
    int32_t *flags;
    int32_t *tcnt;
    int32_t *tkeep;   // run lengths of the last partition (kept when the cursors are zeroed)
    const double4 *tst;
    const int2 *ttag;
    const double *trad;
    edmd_ev32 *ev;
    unsigned long long *overlap_key;
    unsigned long long *ts;
};

// Shared memory of one CTA (cap = tg.smem_cap records):
//   xy   double2[cap]  FP64 positions in cell order   (two 16-byte arrays instead of one 32-byte
//   vv   double2[cap]  FP64 velocities in cell order   record: 128-bit accesses stay conflict-free)
//   scr  float4[cap]   screening records in cell order
//   pk   uint2[kFC]    per frame cell c: off[c-1], off[c], off[c+1], off[c+2] as 4 x u16
//   rad  double[cap]   radii in cell order (only when the radii are not all exactly rad0)
//   id   int[cap]      particle ids in cell order (global ids in a slab context)
//   off  int[kFC + 1]  cell counters, then their exclusive scan; dead once pk is built, and
//   own  uint[cap]     ALIASES off: k-th owned particle in cell order -> position | frame cell << 16
//   misc                warp sums, run counts, the screening constants
struct TileSmem {
    double2 *xy, *vv;
    float4 *scr;
    uint2 *pk;
    double *rad;
    int *id;
    int *off;
    unsigned int *own;
    int *wsum;   // [kTileWarps]
    int *cnt;    // [kFH] records per run, [kFH] = owned total, [kFH + 1] = records
    LeanConsts *K;
};

__host__ __device__ inline size_t tile_aliased_bytes(int cap)
{
    const size_t a = (size_t)cap * 4, b = sizeof(int) * (kFC + 4);
    return ((a > b ? a : b) + 15) & ~(size_t)15;
}

__device__ __forceinline__ TileSmem carve(unsigned char *base, int cap, int rad_smem)
{
    TileSmem s;
    s.xy = reinterpret_cast<double2 *>(base);
    base += (size_t)cap * 16;
    s.vv = reinterpret_cast<double2 *>(base);
    base += (size_t)cap * 16;
    s.scr = reinterpret_cast<float4 *>(base);
    base += (size_t)cap * 16;
    s.pk = reinterpret_cast<uint2 *>(base);
    base += sizeof(uint2) * kFC;
    s.rad = reinterpret_cast<double *>(base);
    base += rad_smem ? (size_t)cap * 8 : 0;
    s.id = reinterpret_cast<int *>(base);
    base += (size_t)cap * 4;
    s.off = reinterpret_cast<int *>(base);
    s.own = reinterpret_cast<unsigned int *>(base);
    base += tile_aliased_bytes(cap);
    s.wsum = reinterpret_cast<int *>(base);
    s.cnt = s.wsum + kTileWarps;
    s.K = reinterpret_cast<LeanConsts *>(s.cnt + kFH + 4);
    return s;
}

// the psi6 kernel keeps positions and ids only: half the shared memory, twice the resident CTAs
__device__ __forceinline__ TileSmem carve_boop(unsigned char *base, int cap)
{
    TileSmem s;
    s.xy = reinterpret_cast<double2 *>(base);
    base += (size_t)cap * 16;
    s.vv = nullptr;
    s.scr = nullptr;
    s.rad = nullptr;
    s.K = nullptr;
    s.pk = reinterpret_cast<uint2 *>(base);
    base += sizeof(uint2) * kFC;
    s.id = reinterpret_cast<int *>(base);
    base += (size_t)cap * 4;
    s.off = reinterpret_cast<int *>(base);
    s.own = reinterpret_cast<unsigned int *>(base);
    base += tile_aliased_bytes(cap);
    s.wsum = reinterpret_cast<int *>(base);
    s.cnt = s.wsum + kTileWarps;
    return s;
}
constexpr int kBoopCtas = 6;   // per SM

struct Held {
    double4 st;
    double rad;
    int id, key;   // key = frame cell | rank << 16, -1: none
};

__device__ __forceinline__ void st_ev(edmd_ev32 *dst, double t_cross, double t_coll, int partner, int dir)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst),
                 "r"(__double2loint(t_cross)), "r"(__double2hiint(t_cross)), "r"(__double2loint(t_coll)),
                 "r"(__double2hiint(t_coll)), "r"(partner), "r"(dir | (EDMD_EV_COLLISION << 8)), "r"(0), "r"(0)
                 : "memory");
}

// where a tile sits in the grid
struct TilePos {
    int txi, tyi;   // tile column / row
    int tw, th;     // its size in cells (the last tile column / row of the grid may be smaller)
};
__device__ __forceinline__ TilePos tile_pos(const TileGeom &tg, int tile)
{
    TilePos t;
    t.tyi = tile / tg.ntx;
    t.txi = tile - t.tyi * tg.ntx;
    t.tw = t.txi == tg.ntx - 1 ? tg.wlast : kTX;
    t.th = t.tyi == tg.nty - 1 ? tg.hlast : kTY;
    return t;
}

// Bin the bucket of one tile by frame cell in shared memory: count -> scan -> place.  Leaves
// s.pk (packed cell offsets), s.own (the owned particles in cell order), s.cnt[kFH] (their number);
// `store(pos, lcx, lcy, state, radius, id)` writes one record to its sorted position.  Ends with a
// barrier: everything is visible to the whole CTA.
template <class Store>
__device__ __forceinline__ void tile_bin(const SweepArgs &a, const TileSmem &s, int tile, const TilePos &tp,
                                         const bool radii, Store store)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t slot0 = (size_t)tile * kFH * kRunCap;
    const double4 *bst = a.tst + slot0;
    const int2 *btag = a.ttag + slot0;
    const double *brad = a.trad + slot0;
    const int tw = tp.tw, th = tp.th;

    // ---- count: rank of every record inside its frame cell (shared-memory atomics).  Warp w takes
    // the runs (frame rows) w, w + kTileWarps, ...; all loads go out before the first atomic --------
    Held held[kRunsPerWarp];
    int key2[kRunsPerWarp][2];   // records 32.. of a run longer than 32 (dense systems only)
    {
        int2 tg2[kRunsPerWarp];
#pragma unroll
        for (int q = 0; q < kRunsPerWarp; q++) {
            const int fy = warp + kTileWarps * q;
            const int n = fy < kFH ? s.cnt[fy] : 0;
            tg2[q] = make_int2(0, 0);
            held[q].st = make_double4(0, 0, 0, 0);
            held[q].rad = a.rad0;
            if (lane < n) {
                tg2[q] = btag[fy * kRunCap + lane];
                held[q].st = ld_sector(bst + fy * kRunCap + lane);
                if (radii) held[q].rad = brad[fy * kRunCap + lane];
            }
        }
#pragma unroll
        for (int q = 0; q < kRunsPerWarp; q++) {
            const int fy = warp + kTileWarps * q;
            const int n = fy < kFH ? s.cnt[fy] : 0;
            held[q].key = -1;
            key2[q][0] = key2[q][1] = -1;
            if (lane < n) {
                const int lc = min(max(tg2[q].y, 0), kFC - 1);
                held[q].id = tg2[q].x;
                held[q].key = lc | (atomicAdd(&s.off[lc], 1) << 16);
            }
#pragma unroll
            for (int h = 1; h <= 2; h++)
                if (n > 32 * h && lane + 32 * h < n) {
                    const int lc = min(max(btag[fy * kRunCap + lane + 32 * h].y, 0), kFC - 1);
                    key2[q][h - 1] = lc | (atomicAdd(&s.off[lc], 1) << 16);
                }
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
        if (lane == 31) s.wsum[warp] = incl;
        __syncthreads();
        int ex = incl - sum;
#pragma unroll
        for (int w = 0; w < kTileWarps; w++)
            if (w < warp) ex += s.wsum[w];
#pragma unroll
        for (int q = 0; q < kCpt; q++) {
            if (c0 + q <= kFC) s.off[c0 + q] = ex;
            ex += v[q];
        }
    }
    __syncthreads();
    // ---- tables: packed offsets per cell, owned particles per frame row ------------------
    constexpr int kPkPt = (kFC + kTileThreads - 1) / kTileThreads;
    uint2 mypk[kPkPt];
#pragma unroll
    for (int q = 0; q < kPkPt; q++) {
        const int c = tid + q * kTileThreads;
        if (c < kFC) {
            const unsigned o0 = s.off[max(c - 1, 0)], o1 = s.off[c], o2 = s.off[c + 1], o3 = s.off[min(c + 2, kFC)];
            mypk[q] = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));
        }
    }
    // Owned particles of frame row y = its cells 1 .. tw (cells 0 and tw+1 are the ring; tw x th is this
    // tile's size: the last tile column / row of the grid may be smaller), rows 1 .. th.  Every warp derives
    // the row table for itself (lane y holds row y): rb = sorted start of cell (y, 1) minus the owned
    // particles before row y, so that position - rb = rank among the owned particles.
    int rb;
    {
        int len = 0, start = 0;
        if (lane >= 1 && lane <= th) {
            start = s.off[lane * kFW + 1];
            len = s.off[lane * kFW + tw + 1] - start;
        }
        int incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        rb = start - (incl - len);
        if (tid == 31) s.cnt[kFH] = incl;   // owned particles of the tile
    }
    __syncthreads();   // every read of off[] is done: own[] may overwrite it
#pragma unroll
    for (int q = 0; q < kPkPt; q++) {
        const int c = tid + q * kTileThreads;
        if (c < kFC) s.pk[c] = mypk[q];
    }
    __syncthreads();
    // ---- place: records into cell order; the list of owned particles -----------------------
    auto place = [&](int key, const double4 &st, double rad, int id) -> int {
        const int lc = key & 0xffff;
        const int pos = (int)(s.pk[lc].x >> 16) + (key >> 16);
        const int lcy = lc / kFW, lcx = lc - lcy * kFW;
        store(pos, lcx, lcy, st, rad, id);
        return (lcx >= 1 && lcx <= tw && lcy >= 1 && lcy <= th) ? lcy : -1;
    };
    // (the shuffle that fetches a row's rb runs with the whole warp converged)
    auto place_own = [&](bool active, int key, const double4 &st, double rad, int id) {
        int lcy = -1;
        if (active) lcy = place(key, st, rad, id);
        const int r = __shfl_sync(0xffffffffu, rb, lcy >= 0 ? lcy : 0);
        if (lcy >= 0) {
            const int lc = key & 0xffff;
            const int pos = (int)(s.pk[lc].x >> 16) + (key >> 16);
            s.own[pos - r] = (unsigned)pos | ((unsigned)lc << 16);
        }
    };
#pragma unroll
    for (int q = 0; q < kRunsPerWarp; q++) {
        place_own(held[q].key >= 0, held[q].key, held[q].st, held[q].rad, held[q].id);
#pragma unroll
        for (int h = 1; h <= 2; h++) {
            const int fy = warp + kTileWarps * q;
            if ((fy < kFH ? s.cnt[fy] : 0) <= 32 * h) continue;   // warp-uniform
            const bool act = key2[q][h - 1] >= 0;
            const int slot = fy * kRunCap + lane + 32 * h;
            double4 st2 = make_double4(0, 0, 0, 0);
            double rad2 = a.rad0;
            int id2 = 0;
            if (act) {
                st2 = ld_sector(bst + slot);
                id2 = btag[slot].x;
                if (radii) rad2 = brad[slot];
            }
            place_own(act, key2[q][h - 1], st2, rad2, id2);
        }
    }
    __syncthreads();
}

// ---- sweep: every owned particle of the tile, in cell order ---------------------------
template <bool TWO>
__device__ __forceinline__ void tile_main(const SweepArgs &a, const TileSmem &s, int tile, const bool radii)
{
    const int tid = threadIdx.x;
    const TilePos tp = tile_pos(a.tg, tile);
    const int txi = tp.txi, tyi = tp.tyi;
    const int nx = a.b.nx, nl = a.b.nl;
    tile_bin(a, s, tile, tp, radii, [&](int pos, int lcx, int lcy, const double4 &st, double rad, int id) {
        // the FILED cell of the record (a ring cell at the periodic edge is the opposite edge of the grid)
        int X = txi * kTX + lcx - 1, Yl = tyi * kTY + lcy - 1;
        X = X < 0 ? X + nx : (X >= nx ? X - nx : X);
        Yl = Yl < 0 ? Yl + nl : (Yl >= nl ? Yl - nl : Yl);
        // screening record, relative to the centre of the filed cell (lean.cuh)
        float4 r;
        r.x = __double2float_rn(__dsub_rn(st.x, __dmul_rn((double)X + 0.5, a.b.csx)));
        r.y = __double2float_rn(__dsub_rn(st.y, __dmul_rn((double)edmd_global_row(a.b, Yl) + 0.5, a.b.csy)));
        r.z = __double2float_rn(st.z);
        r.w = __double2float_rn(st.w);
        if (TWO) r.w = __int_as_float((__float_as_int(r.w) & ~1) | (edmd_same_class(rad, a.rad0) ? 0 : 1));
        s.xy[pos] = make_double2(st.x, st.y);
        s.vv[pos] = make_double2(st.z, st.w);
        s.scr[pos] = r;
        s.id[pos] = id;
        if (radii) s.rad[pos] = rad;
    });
    const int nown = (a.dbg & 4) ? 0 : s.cnt[kFH];
    const LeanConsts K = *s.K;
    const float fnan = __int_as_float(0x7fffffff);
#pragma unroll 1
    for (int k = tid; k < nown; k += kTileThreads) {
        const unsigned u = s.own[k];
        const int self = u & 0xffff, lc = u >> 16;
        const int id = s.id[self];
        const float4 own = s.scr[self];
        const uint2 pkr[3] = {s.pk[lc - kFW], s.pk[lc], s.pk[lc + kFW]};
        if (id >= a.n_owned) continue;   // halo copy from a neighbouring slab: never predicted
        const int lcy = lc / kFW, lcx = lc - lcy * kFW;

        int lo[3], t1[3], t2[3], hi[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            lo[j] = pkr[j].x & 0xffff; t1[j] = pkr[j].x >> 16; t2[j] = pkr[j].y & 0xffff; hi[j] = pkr[j].y >> 16;
        }
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
        if (!(a.dbg & 1))
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
        const float second = lo1 > 0.0f ? lo2 : fnan;   // a non-positive smallest bound is never certifiable
        // slab contexts: the partner leaves as the caller's (global) id; fetch the winner's now, the load
        // flies while the exact stage computes
        const int wloc = idx >= 0 ? s.id[idx] : -1;
        const int wglob = (wloc >= 0 && a.gid) ? a.gid[wloc] : wloc;
        const double2 mexy = s.xy[self], mev = s.vv[self];
        const double4 me = make_double4(mexy.x, mexy.y, mev.x, mev.y);
        const double rad_i = radii ? s.rad[self] : a.rad0;
        if (a.dbg & 2) {
            st_ev(a.ev + id, (double)second + me.x, 0.0, idx, 0);
            continue;
        }
        // ---- exact: crossing + the winner's pair time, as the reference computes them ----
        const int X = txi * kTX + lcx - 1, Yl = tyi * kTY + lcy - 1;
        SRec p1;
        p1.x = me.x; p1.y = me.y; p1.vx = me.z; p1.vy = me.w;
        p1.rad = rad_i; p1.id = id; p1.pc = 0;
        const double four_r1 = __dmul_rn(4.0, p1.rad);
        double dtc;
        int dirc;
        crossing_fast<true>(a.b, p1, X, edmd_global_row(a.b, Yl), dtc, dirc);
        double best = EDMD_NEVER;
        int best_id = -1, best_pc = -1, ov_id = -1, ov_pc = -1;
        bool certified = idx < 0;   // no candidate can collide: partner 0 at t + 1e26
        if (idx >= 0) {
            const double2 wxy = s.xy[idx], wv = s.vv[idx];
            const double4 wq = make_double4(wxy.x, wxy.y, wv.x, wv.y);
            SRec p2;
            p2.x = wq.x; p2.y = wq.y; p2.vx = wq.z; p2.vy = wq.w;
            p2.rad = radii ? s.rad[idx] : a.rad0; p2.id = s.id[idx]; p2.pc = 0;
            double bb, v2, cc, b2, vc;
            pair_terms<true>(a.b, p1, four_r1, p2, bb, v2, cc, b2, vc);
            const double det = __dsub_rn(b2, vc);
            const double T = __ddiv_rn(__dsub_rn(-bb, __dsqrt_rn(det)), v2);
            // a real collision (the reference's branches) and every other candidate's lower
            // bound above the exact time (a NaN bound or time fails the test)
            certified = !(bb > 0) && (det >= 0) && ((double)second > T);
            best = T;
            best_id = p2.id;
        }
        auto gidx = [&](int q) { return a.gid ? a.gid[q] : q; };   // the caller's (global) id of a local one
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
                    if (p == self) continue;   // `p1 != p2` is identity
                    const double2 qxy = s.xy[p], qv = s.vv[p];
                    const double4 q = make_double4(qxy.x, qxy.y, qv.x, qv.y);
                    SRec p2;
                    p2.x = q.x; p2.y = q.y; p2.vx = q.z; p2.vy = q.w;
                    p2.rad = radii ? s.rad[p] : a.rad0; p2.id = s.id[p];
                    p2.pc = 3 * j + (p < t1[j] ? 0 : (p < t2[j] ? 1 : 2));   // scan-order cell
                    bool ov = false;
                    const double dt = pair_time_normal<true>(a.b, p1, four_r1, p2, ov);
                    if (ov && (ov_id < 0 || (p2.pc == ov_pc && gidx(p2.id) > gidx(ov_id)))) {
                        ov_id = p2.id;
                        ov_pc = p2.pc;
                    }
                    if (best > dt || (best == dt && best_id >= 0 && p2.pc == best_pc && gidx(p2.id) > gidx(best_id))) {
                        best = dt;
                        best_id = p2.id;
                        best_pc = p2.pc;
                    }
                }
            }
        }
        // ids in shared memory are LOCAL; partners and the overlap report carry the caller's ids
        const int partner = best_id >= 0 ? (best_id == wloc ? wglob : (a.gid ? a.gid[best_id] : best_id)) : 0;
        st_ev(a.ev + id, __dadd_rn(a.t, dtc), __dadd_rn(a.t, best), partner, dirc);
        if (ov_id >= 0) {
            const unsigned long long key = ((unsigned long long)(uint32_t)(a.gid ? a.gid[id] : id) << 32) |
                                           (uint32_t)(a.gid ? a.gid[ov_id] : ov_id);
            atomicMin(a.overlap_key, key);
        }
    }
}

// Common prologue of the tile kernels: the run lengths of the tile (from the live cursors, which go back
// to zero for the next partition -- this CTA is their only reader -- and are kept in tkeep[] for later
// passes over the same buckets; or from tkeep[] when `from_keep`).  Returns the number of records.
__device__ __forceinline__ int tile_prologue(const SweepArgs &a, const TileSmem &s, int tile, bool from_keep)
{
    const int tid = threadIdx.x;
    if (tid < 32) {
        int n = 0;
        if (tid < kFH) {
            const size_t run = (size_t)tile * kFH + tid;
            if (from_keep) {
                n = a.tkeep[run];
            } else {
                int32_t *cur = a.tcnt + run * kCurStride;
                n = min(*cur, kRunCap);
                *cur = 0;
                a.tkeep[run] = n;
            }
            s.cnt[tid] = n;
        }
        const int total = __reduce_add_sync(0xffffffffu, n);
        if (tid == 0) s.cnt[kFH + 1] = total;
    }
    return 0;
}

__global__ void __launch_bounds__(kTileThreads, kTileCtas)
k_tile_sweep(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char tile_smem[];
    const TileSmem s = carve(tile_smem, a.tg.smem_cap, a.rad_smem);
    const int tid = threadIdx.x;
    // zero the cell counters while the previous kernel drains
    for (int c = tid; c <= kFC; c += kTileThreads) s.off[c] = 0;
    edmd_pdl_wait();
    const int tile = a.tiles ? a.tiles[blockIdx.x] : blockIdx.x;
    if (blockIdx.x == 0 && tid == 0) edmd_stamp(a.ts, 7);
    const int classes = a.flags[kFlagNotMono];   // 0: one radius, 1: two classes, more: not eligible
    const double rad1 = __longlong_as_double(*reinterpret_cast<const long long *>(a.flags + kFlagRad1));
    tile_prologue(a, s, tile, false);
    if (tid == 32) *s.K = make_consts(a.b, a.rad0, rad1, classes == 1, __int_as_float(a.flags[kFlagVmax]));
    const bool declined = a.flags[kFlagLeanFail] != 0;
    const bool bad = a.flags[kFlagInsane] != 0 || classes > 1 || (classes == 1 && !a.rad_smem);
    __syncthreads();
    if (declined) return;   // a run overflowed: the host re-runs the sweep on the full path
    if (bad || !s.K->ok || s.cnt[kFH + 1] > a.tg.smem_cap) {
        // not eligible, or more records than this CTA's shared memory holds: decline
        if (tid == 0) atomicOr(&a.flags[kFlagLeanFail], 1);
        return;
    }
    // classes == 1: the radii are spread inside their classes (a reference-grown system): the exact
    // stage takes every disk's own FP64 radius; the class bit is only looked at when a second class exists
    if (classes == 1 && rad1 > 0.0) tile_main<true>(a, s, tile, true);
    else tile_main<false>(a, s, tile, classes == 1);
}

// ---- K4 on the tile buckets: computeBOOPCutoff, src/boop.c:61-107 ---------------------------
// Same binning; positions only.  Per owned particle the reference's 3 x 3 scan with its own
// arithmetic (FP64 differences, PBC, r2 < r_c^2: neighbour counts are integers identical to the
// reference's; the same truncation: cells are ~2.0 wide, r_c = 2.5, neighbours two cells away are
// never seen).  e^{ik theta} by complex powers, sums in 2^-48 fixed point (order-independent), as
// k_boop_rows (analysis.cu).  Works for any radii; a bucket overflow sets kFlagBoopFail and the host
// falls back to the row kernel.
struct BoopTileArgs {
    SweepArgs s;
    int from_keep;
    double rc2;
    double4 *rec;   // two 32-byte sectors per particle id: (q5, q6, q7, q6_arg), (neighbours, 0, 0, 0)
};

__global__ void __launch_bounds__(kTileThreads, kBoopCtas)
k_tile_boop(const __grid_constant__ BoopTileArgs ba)
{
    extern __shared__ __align__(128) unsigned char tile_smem[];
    const SweepArgs &a = ba.s;
    const TileSmem s = carve_boop(tile_smem, a.tg.smem_cap);
    const int tid = threadIdx.x;
    for (int c = tid; c <= kFC; c += kTileThreads) s.off[c] = 0;
    edmd_pdl_wait();
    const int tile = blockIdx.x;
    tile_prologue(a, s, tile, ba.from_keep != 0);
    const bool declined = a.flags[kFlagBoopFail] != 0;
    __syncthreads();
    if (declined) return;
    if (s.cnt[kFH + 1] > a.tg.smem_cap) {
        if (tid == 0) atomicOr(&a.flags[kFlagBoopFail], 1);
        return;
    }
    const TilePos tp = tile_pos(a.tg, tile);
    tile_bin(a, s, tile, tp, false, [&](int pos, int, int, const double4 &st, double, int id) {
        s.xy[pos] = make_double2(st.x, st.y);
        s.id[pos] = id;
    });
    // Records inside a cell sit in arrival order (atomics).  Sort every cell by particle id: the FP64
    // sums below then run in one fixed order (rows, cells, ids) and every output bit is reproducible.
    for (int c = tid; c < kFC; c += kTileThreads) {
        const uint2 pk = s.pk[c];
        const int lo = pk.x >> 16, hi = pk.y & 0xffff;
        for (int i = lo + 1; i < hi; i++) {
            const int idi = s.id[i];
            const double2 xi = s.xy[i];
            int j = i - 1;
            for (; j >= lo && s.id[j] > idi; j--) {
                s.id[j + 1] = s.id[j];
                s.xy[j + 1] = s.xy[j];
            }
            s.id[j + 1] = idi;
            s.xy[j + 1] = xi;
        }
    }
    __syncthreads();
    const int nown = s.cnt[kFH];
    const double half_lx = a.b.half_lx, half_ly = a.b.half_ly, lx = a.b.lx, ly = a.b.ly, rc2 = ba.rc2;
    // a tile away from the edges of the grid never sees the periodic image (every particle lies within
    // 1.5 cells of the cell it is filed under -- kFlagInsane -- so |d| <= 4 cells < L/2): PBC() is the identity
    // (frame rows lo .. hi in local rows; in a slab the global row yoff + l wraps somewhere inside the slab)
    const int fr_lo = tp.tyi * kTY - 1, fr_hi = tp.tyi * kTY + tp.th;
    const bool wrap_y = fr_lo < 0 || fr_hi >= a.b.nl || (a.b.yoff + fr_lo < a.b.ny && a.b.yoff + fr_hi >= a.b.ny);
    const bool wrap = a.flags[kFlagInsane] != 0 || tp.txi == 0 || tp.txi == a.tg.ntx - 1 || wrap_y ||
                      a.b.nx < 12 || a.b.ny < 12;
#pragma unroll 1
    for (int k = tid; k < nown; k += kTileThreads) {
        const unsigned u = s.own[k];
        const int self = u & 0xffff, lc = u >> 16;
        const int id = s.id[self];
        if (id >= a.n_owned) continue;   // halo copy from a neighbouring slab
        const double2 me = s.xy[self];
        const uint2 pkr[3] = {s.pk[lc - kFW], s.pk[lc], s.pk[lc + kFW]};
        double s5r = 0, s5i = 0, s6r = 0, s6i = 0, s7r = 0, s7i = 0;
        int nb = 0;
#pragma unroll 1
        for (int j = 0; j < 3; j++) {
            const int lo = pkr[j].x & 0xffff, hi = pkr[j].y >> 16;
#pragma unroll 1
            for (int p = lo; p < hi; p++) {
                if (p == self) continue;   // `p2->num != p1->num`
                const double2 q = s.xy[p];
                // the reference's own operations decide who is a neighbour (src/boop.c:78-84)
                double dx = __dsub_rn(q.x, me.x), dy = __dsub_rn(q.y, me.y);
                if (wrap) {
                    dx = min_image(dx, half_lx, lx);
                    dy = min_image(dy, half_ly, ly);
                }
                const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (r2 < rc2) {
                    nb++;
                    // e^{ik theta} = ((dx + i dy)/r)^k, k = 5, 6, 7, by complex powers (fused multiply-adds:
                    // not parity-critical, gate 1e-10); 1/r from the MUFU seed + two Newton steps
                    double zr = 1.0, zi = 0.0;   // atan2(0,0) = 0 in the reference
                    if (r2 > 0) {
                        double y = rsqrt_seed(r2);
                        double e = __fma_rn(-__dmul_rn(r2, y), y, 1.0);
                        y = __fma_rn(__dmul_rn(0.5, y), e, y);
                        e = __fma_rn(-__dmul_rn(r2, y), y, 1.0);
                        y = __fma_rn(__dmul_rn(0.5, y), e, y);
                        zr = __dmul_rn(dx, y);
                        zi = __dmul_rn(dy, y);
                    }
                    const double z2r = __fma_rn(zr, zr, -__dmul_rn(zi, zi)), z2i = __dmul_rn(__dadd_rn(zr, zr), zi);
                    const double z4r = __fma_rn(z2r, z2r, -__dmul_rn(z2i, z2i)), z4i = __dmul_rn(__dadd_rn(z2r, z2r), z2i);
                    const double z6r = __fma_rn(z4r, z2r, -__dmul_rn(z4i, z2i)), z6i = __fma_rn(z4r, z2i, __dmul_rn(z4i, z2r));
                    // |z| = 1: z^5 = z^6 conj(z), z^7 = z^6 z
                    s5r += __fma_rn(z6r, zr, __dmul_rn(z6i, zi));
                    s5i += __fma_rn(z6i, zr, -__dmul_rn(z6r, zi));
                    s6r += z6r;
                    s6i += z6i;
                    s7r += __fma_rn(z6r, zr, -__dmul_rn(z6i, zi));
                    s7i += __fma_rn(z6r, zi, __dmul_rn(z6i, zr));
                }
            }
        }
        double q5 = 0.0, q6 = 0.0, q7 = 0.0, arg = 0.0;
        if (nb > 0) {
            const double inv_n = 1.0 / (double)nb;
            auto modulus = [&](double re, double im) {   // |re + i im| / n  (no overflow: |sum| <= n)
                const double m2 = __fma_rn(re, re, __dmul_rn(im, im));
                return __dmul_rn(__dsqrt_rn(m2), inv_n);
            };
            q5 = modulus(s5r, s5i);
            q6 = modulus(s6r, s6i);
            q7 = modulus(s7r, s7i);
            arg = atan2(s6i, s6r);
        }
        // Two FULL-sector stores by particle id.  (Five scattered 8-/4-byte stores cost 60 us at N = 10^6:
        // a partial write to a sector that is not in L2 makes L2 fetch it from DRAM first.)
        double4 *dst = ba.rec + 2 * (size_t)id;
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(q5), "d"(q6), "d"(q7), "d"(arg) : "memory");
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%2,%2,%2,%2,%2,%2};" ::"l"(dst + 1), "r"(nb), "r"(0) : "memory");
    }
}

// the ABI's five psi6 arrays from the records (one coalesced pass)
__global__ void __launch_bounds__(256)
k_unpack_boop(int n, const double4 *__restrict__ rec, double *__restrict__ q5, double *__restrict__ q6,
              double *__restrict__ q7, double *__restrict__ q6arg, int32_t *__restrict__ nbr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 v = ld_sector(rec + 2 * (size_t)i);
    const int nb = *reinterpret_cast<const int *>(rec + 2 * (size_t)i + 1);
    q5[i] = v.x;
    q6[i] = v.y;
    q7[i] = v.z;
    q6arg[i] = v.w;
    nbr[i] = nb;
}

// the ABI's five prediction arrays from the event records (one coalesced pass)
__global__ void __launch_bounds__(256)
k_unpack_events(int n, const edmd_ev32 *__restrict__ ev, double *__restrict__ t_cross, uint8_t *__restrict__ dir,
                double *__restrict__ t_coll, int32_t *__restrict__ partner)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 v = ld_sector(reinterpret_cast<const double4 *>(ev + i));
    t_cross[i] = v.x;
    t_coll[i] = v.y;
    partner[i] = __double2loint(v.z);
    dir[i] = (uint8_t)(__double2hiint(v.z) & 0xff);
}

}  // namespace

bool edmd_tile_eligible(const edmd_ctx *c, int mode)
{
    return edmd_lean_eligible(c, mode) && !c->tile_off && c->tst != nullptr;
}

static size_t tile_smem_bytes(const TileGeom &tg, int rad_smem)
{
    return (size_t)tg.smem_cap * (32 + 16 + 4 + (rad_smem ? 8 : 0)) + sizeof(uint2) * kFC +
           tile_aliased_bytes(tg.smem_cap) + sizeof(int) * (kTileWarps + kFH + 4) + sizeof(LeanConsts) + 16;
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
    long long cap = (long long)(1.25 * dens * kFC) + 96;
    cap = (cap + 31) & ~31ll;
    if (cap > (long long)kFH * kRunCap) cap = (long long)kFH * kRunCap;
    tg.smem_cap = (int)cap;
    if (tile_smem_bytes(tg, 1) > 200 * 1024) return false;   // denser than a CTA's shared memory provides for
    if ((long long)tg.ntx * tg.nty * kFH * kRunCap >= (1ll << 31)) return false;
    *out = tg;
    return true;
}

int edmd_launch_unpack_events(edmd_ctx *c)
{
    const int n = c->n_owned;
    if (n == 0) return 0;
    k_unpack_events<<<(n + 255) / 256, 256, 0, c->stream>>>(n, c->evrec, c->t_cross, c->dir, c->t_coll, c->partner);
    return 1;
}

static PartArgs part_args(edmd_ctx *c, int first, int n)
{
    PartArgs pa;
    pa.first = first; pa.n = n; pa.ps = c->ps; pa.slab = c->slab ? 1 : 0; pa.dbg = c->tile_dbg;
    pa.late_wait = 0;
    pa.tg = c->tgeom; pa.b = c->dbox;
    pa.cid = c->cid; pa.xv = c->xv; pa.rad = c->rad; pa.rad0 = c->rad0;
    pa.flags = c->flags; pa.tcnt = c->tcnt; pa.tst = c->tst; pa.ttag = c->ttag; pa.trad = c->trad;
    pa.overlap_key = c->overlap_key;
    pa.ts = c->tile_dbg & 32 ? c->dbg_ts : nullptr;
    return pa;
}

// P1 alone: (re)build the tile buckets of the resident state (particles [first, first + n))
int edmd_launch_tile_partition_range(edmd_ctx *c, int first, int n, bool beside)
{
    if (n <= 0) return 0;
    PartArgs pa = part_args(c, first, n);
    pa.late_wait = beside ? 1 : 0;
    // normally the first kernel of its chain, launched plainly: whatever precedes it on the stream completes
    // first.  `beside`: behind the halo kernels of the fused exchange, with the programmatic attribute
    edmd_launch(k_tile_partition, dim3((n + kPartThreads - 1) / kPartThreads), dim3(kPartThreads), 0, c->stream,
                beside, pa);
    return 1;
}

int edmd_launch_tile_partition(edmd_ctx *c) { return edmd_launch_tile_partition_range(c, 0, c->n, false); }

// slab contexts with the peer-to-peer halo: receive the neighbours' boundary rows (sent by
// edmd_launch_halo_send) and append them to the tile buckets in the same kernel
int edmd_launch_halo_recv_partition(edmd_ctx *c)
{
    const int H = c->halo_cap;
    const int e = c->halo_epoch, par = e & 1;
    RecvPartArgs ra;
    ra.p = part_args(c, c->n_owned, 2 * H);
    ra.p.overlap_key = nullptr;
    ra.H = H; ra.epoch = e;
    ra.row[0] = 0; ra.row[1] = c->dbox.nl - 1;
    ra.inbox[0] = c->halo_mem + inbox_offset(H, 0, par);
    ra.inbox[1] = c->halo_mem + inbox_offset(H, 1, par);
    ra.xv = c->xv; ra.rad = c->rad; ra.cid = c->cid; ra.gid = c->gid;
    // consuming the lower neighbour's records: I am its UPPER neighbour -> its ack[1]; and vice versa
    ra.peer_ack[0] = reinterpret_cast<int *>(c->peer_mem[0] + ack_offset(H)) + 1;
    ra.peer_ack[1] = reinterpret_cast<int *>(c->peer_mem[1] + ack_offset(H)) + 0;
    ra.done = c->halo_cnt + 4;
    edmd_launch(k_halo_recv_partition, dim3((H + kPartThreads - 1) / kPartThreads, 2), dim3(kPartThreads), 0, c->stream,
                c->lean_pdl, ra);
    return 1;
}

static SweepArgs sweep_args(edmd_ctx *c)
{
    SweepArgs sa;
    sa.tg = c->tgeom; sa.b = c->dbox; sa.t = c->t; sa.rad0 = c->rad0; sa.n_owned = c->n_owned;
    sa.tiles = nullptr;
    sa.dbg = c->tile_dbg;
    sa.rad_smem = c->lean_two ? 1 : 0;   // the host's knowledge; the kernel declines if the device knows better
    sa.gid = c->slab ? c->gid : nullptr;
    sa.flags = c->flags; sa.tcnt = c->tcnt; sa.tkeep = c->tkeep; sa.tst = c->tst; sa.ttag = c->ttag; sa.trad = c->trad;
    sa.ev = c->evrec;
    sa.overlap_key = c->overlap_key;
    sa.ts = c->tile_dbg & 32 ? c->dbg_ts : nullptr;
    return sa;
}

static void tile_attrs()
{
    static bool attr = false;
    if (attr) return;
    cudaFuncSetAttribute(k_tile_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_tile_sweep, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_tile_boop, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_tile_boop, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    attr = true;
}

// the sweep kernel over the buckets (P2)
int edmd_launch_tile_sweep_only(edmd_ctx *c)
{
    const SweepArgs sa = sweep_args(c);
    tile_attrs();
    edmd_launch(k_tile_sweep, dim3(c->tgeom.ntx * c->tgeom.nty), dim3(kTileThreads), tile_smem_bytes(c->tgeom, sa.rad_smem),
                c->stream, c->lean_pdl, sa);
    c->pred_packed = true;
    return 1;
}

int edmd_launch_tile_sweep(edmd_ctx *c, cudaEvent_t between)
{
    if (c->n == 0) return 0;
    edmd_launch_tile_partition(c);
    if (between) cudaEventRecord(between, c->stream);
    const SweepArgs sa = sweep_args(c);
    tile_attrs();
    edmd_launch(k_tile_sweep, dim3(c->tgeom.ntx * c->tgeom.nty), dim3(kTileThreads), tile_smem_bytes(c->tgeom, sa.rad_smem),
                c->stream, c->lean_pdl, sa);
    c->pred_packed = true;
    return 2;
}

// K4 on the tile buckets; from_keep: the buckets were consumed once already (run lengths in tkeep[])
int edmd_launch_tile_boop(edmd_ctx *c, double r_c, bool from_keep)
{
    if (c->n == 0) return 0;
    const size_t N = (size_t)c->n_cap;   // the four psi6 arrays lie n_cap apart
    BoopTileArgs ba;
    ba.s = sweep_args(c);
    ba.from_keep = from_keep ? 1 : 0;
    ba.rc2 = r_c * r_c;   // `r_c*r_c`, a single rounded product
    ba.rec = c->boop_rec;
    tile_attrs();
    const size_t smem = (size_t)c->tgeom.smem_cap * 20 + sizeof(uint2) * kFC + tile_aliased_bytes(c->tgeom.smem_cap) +
                        sizeof(int) * (kTileWarps + kFH + 4) + 16;
    edmd_launch(k_tile_boop, dim3(c->tgeom.ntx * c->tgeom.nty), dim3(kTileThreads), smem, c->stream,
                c->lean_pdl && !from_keep, ba);
    // halo copies of a slab context are not computed: the unpack covers the owned particles
    const int no = c->n_owned;
    k_unpack_boop<<<(no + 255) / 256, 256, 0, c->stream>>>(no, c->boop_rec, c->boop, c->boop + N, c->boop + 2 * N,
                                                           c->boop + 3 * N, c->boop_nb);
    return 2;
}
