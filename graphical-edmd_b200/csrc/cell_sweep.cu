// cell_sweep.cu -- the whole-system prediction sweep (NORMAL mode, one or two
// radius classes) in TWO kernels over FIXED-CAPACITY CELL SLOTS.  Replaces, bit for bit,
//   cellListInit / addToCell   src/EDMD.c:1906-1920, 2053-2078   (the cell index)
//   crossingEventNormal        src/EDMD.c:2405-2482
//   collisionEventNormal       src/EDMD.c:2829-3102   (scan 2959-3003, select 2991-2995)
//   collisionTimeNormal        src/EDMD.c:2661-2723
//
// The reference's cells are at least one disk diameter wide, so a cell holds ONE disk
// nearly always (jittered lattice at phi = 0.70: 12.5 % empty, 85.8 % one, 1.8 % two;
// reference-grown liquid: 15.0 / 80.6 / 4.4 %; phi = 0.85: 2.7 / 86.2 / 11.1 %).  The cell
// index is therefore kSlotK PLANES over the padded cell grid: plane s holds the s-th
// arrival of every cell.  Plane 0 is dense (what the sweep streams), planes 1.. are
// sparse (read only where a cell's counter says so).
//
//   P1  k_cell_partition   one thread per particle (memory order = particle id):
//       slot = atomicAdd(count[cell]); the FP64 state (x, y, vx, vy = one 32-byte sector,
//       one 256-bit store) goes to plane `slot` at the cell's index, the particle id (and
//       the radius when the radii are not all exactly rad0) beside it.  No histogram pass,
//       no scan, no tiles, no copies for neighbouring tiles, no tag: the cell IS the address.
//   P2  k_cell_sweep       one CTA per TILE of kTX x kTY cells.  It streams the counters
//       and plane 0 of its FRAME (the tile + a one-cell ring: everything the reference's
//       3 x 3 scan of the tile's particles touches; the ring of an edge tile is the
//       opposite edge of the grid, PBCcellX/Y src/EDMD.c:2110-2124) into shared memory IN
//       FRAME-CELL ORDER -- no binning, no compaction: an empty cell is a NaN record that the
//       screening drops by itself -- and the few disks of planes 1.. into a short list behind
//       it.  Every particle of the tile then screens its 3 x 3 neighbourhood, fully unrolled
//       (nine plane-0 records at fixed offsets + the listed extras), in FP32 (the certified
//       lower bounds of lean.cuh / predict_lean.cu: same arithmetic, same error model),
//       evaluates crossingEventNormal and the winner's collisionTimeNormal exactly as the
//       reference does from the FP64 states in shared memory, certifies the winner against
//       the second-smallest bound or re-scans the neighbourhood in FP64 in the reference's
//       order, and writes ONE 32-byte event record per particle (edmd_ev32) by particle id.
//       k_unpack_events turns the records into the ABI's five arrays when a caller fetches them.
//
// The counters are double-buffered: a consumer kernel (sweep or psi6) zeroes the OTHER
// buffer's cells of its tile, so the next partition finds zeros without a memset kernel and
// the counters of the current partition stay valid for later passes (psi6 after a sweep).
//
// The state must be eligible exactly as for the lean sweep (lean.cuh); a cell with more
// than kSlotK disks, or a tile with more extras than its shared memory provides for
// (clustered tiny disks), makes the sweep DECLINE through kFlagLeanFail and the host
// re-runs it on the five-kernel lean chain / the full FP64 path.
#include "lean.cuh"
#include "pairmath.cuh"
#include "halo.cuh"
#include "rowstage.cuh"

namespace {

constexpr int kPartThreads = 256;

struct PartArgs {
    int first, n, ncp, dbg, ps, nx;
    int workers;     // CTAs of the sweep kernel that follows: its tile counter starts there
    int late_wait;   // fused halo exchange: run BESIDE the preceding kernels of the chain, wait for them at the end
    const int32_t *cid;
    const double4 *xv;
    const double *rad;
    double rad0;
    int32_t *flags;
    unsigned long long *ccnt;   // cell words of the partition being built (zero on entry): count << 32 | sum of ids
    double4 *pst;    // [kSlotK][ncp] FP64 states
    int32_t *pid;    // [kSlotK][ncp] particle ids
    double *prad;    // [kSlotK][ncp] radii (written only when the radii are not all exactly rad0)
    unsigned long long *overlap_key;
    unsigned long long *ts;
};

// File particle i (padded cell id pc, state p, radius rad) in the next free slot of its cell -- and, when the cell
// is the first / last of its row, of the row's GHOST cell on the other side (padded column nx + 1 / 0, the
// periodic neighbour PBCcellX wraps to, src/EDMD.c:2110-2116): with the ghosts every frame row of the sweep
// is ONE contiguous run of the planes, periodic edge included, and can be fetched by a single bulk copy.
__device__ __forceinline__ void file_one(const PartArgs &a, int i, int pc, const double4 &p, double rad, const bool radii)
{
    // ONE atomic claims the slot AND leaves the id: the cell word is count << 32 | (sum of the ids filed so far,
    // mod 2^32).  A cell with one disk -- nearly all of them -- needs no id store at all (the partition kernel
    // is bound by the number of scattered L2 requests: atomic + state + id were three per particle); the
    // disks of planes 1.. store theirs, the sum gives back plane 0's.
    const int s = (a.dbg & 16) ? 0 : (int)(atomicAdd(&a.ccnt[pc], (1ull << 32) | (unsigned int)i) >> 32);   // (dbg: timing experiments)
    if (a.dbg & 8) return;
    if (s < kSlotK) {
        const size_t slot = (size_t)s * a.ncp + pc;
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(a.pst + slot), "d"(p.x), "d"(p.y), "d"(p.z), "d"(p.w)
                     : "memory");
        if (s > 0) a.pid[slot] = i;
        if (radii) a.prad[slot] = rad;
    } else {
        atomicOr(&a.flags[kFlagLeanFail], 1);   // a cell is full: the sweep declines
    }
}

__device__ __forceinline__ void partition_one(const PartArgs &a, int i, int pc, const double4 &p, double rad,
                                              const bool radii)
{
    file_one(a, i, pc, p, rad, radii);
    const int col = pc - (pc / a.ps) * a.ps;   // padded column 1 .. nx
    if (col == 1) file_one(a, i, pc + a.nx, p, rad, radii);        // cell 0 -> right ghost (column nx + 1)
    else if (col == a.nx) file_one(a, i, pc - a.nx, p, rad, radii);   // cell nx-1 -> left ghost (column 0)
}

__global__ void __launch_bounds__(kPartThreads)
k_cell_partition(const __grid_constant__ PartArgs a)
{
    const int i = a.first + blockIdx.x * blockDim.x + threadIdx.x;
    if (!a.late_wait) edmd_pdl_wait();
    if (i == a.first) edmd_stamp(a.ts, 4);
    if (i == a.first + a.n - 1) edmd_stamp(a.ts, 5);
    if (i == a.first && a.overlap_key) {
        *a.overlap_key = ~0ull;                 // the sweep's overlap report starts empty
        a.flags[kFlagTileCtr] = a.workers;      // ... and its workers take the tiles from here on dynamically
    }
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
// fixed halo region of the resident arrays -- exactly what k_halo_recv (halo.cu) does -- and files
// it in the cell slots right away; acks back.  blockIdx.y = from (0: lower neighbour's records ->
// local row 0, 1: upper -> row nl-1).
struct RecvPartArgs {
    PartArgs p;
    int H, epoch, ps, row[2];
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
            const int pc = a.row[from] * a.ps + r.cell;
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
    int n_owned, dbg, rad_smem, ps, ncp, slab;
    const int32_t *gid;
    int32_t *flags;
    const unsigned long long *ccnt;   // cell words (count << 32 | sum of ids) of the current partition
    unsigned long long *czero;        // the other buffer: this kernel zeroes the cells of its tile
    const double4 *pst;
    const int32_t *pid;
    const double *prad;
    edmd_ev32 *ev;
    unsigned long long *overlap_key;
    unsigned long long *ts;
};

// Shared memory of one CTA.  Index p < kFC = the plane-0 record of frame cell p; p >= kFC = extra
// number p - kFC (a disk of planes 1..).  cap = kFC + tg.ecap.
//   xy    double2[cap]  FP64 positions              (two 16-byte arrays instead of one 32-byte
//   vv    double2[cap]  FP64 velocities              record: 128-bit accesses stay conflict-free)
//   scr   float4[cap]   screening records (NaN: empty cell)
//   rad   double[cap]   radii (only when the radii are not all exactly rad0)
//   id    int[cap]      particle ids, -1: empty (local ids in a slab context)
//   bits  u64[kFH]      per frame row: bit fx set = cell (fx, fy) has extras
//   xinfo u16[kFC]      first extra | (number of extras << 12) of a cell whose bit is set
//   ecell u16[ecap]     home frame cell of every extra
struct CellSmem {
    double2 *xy, *vv;
    float4 *scr;
    double *rad;
    int *id;
    unsigned long long *bits;
    unsigned short *xinfo, *ecell;
    int *misc;   // [0] extras listed, [1] decline
    LeanConsts *K;
};

__host__ __device__ inline size_t cell_smem_bytes(int ecap, int rad_smem, bool boop)
{
    const size_t cap = (size_t)kFC + ecap;
    size_t b = cap * 16;                        // xy
    if (!boop) b += cap * 32;                   // vv, scr
    if (!boop && rad_smem) b += cap * 8;        // rad
    b += sizeof(unsigned long long) * kFH;      // bits
    b += cap * 4;                               // id
    b += 16;                                    // misc
    b += (sizeof(LeanConsts) + 15) & ~(size_t)15;
    b += sizeof(unsigned short) * (kFC + ecap); // xinfo, ecell
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ CellSmem carve(unsigned char *base, int ecap, int rad_smem, bool boop)
{
    const size_t cap = (size_t)kFC + ecap;
    CellSmem s;
    s.xy = reinterpret_cast<double2 *>(base);
    base += cap * 16;
    s.vv = nullptr;
    s.scr = nullptr;
    s.rad = nullptr;
    if (!boop) {
        s.vv = reinterpret_cast<double2 *>(base);
        base += cap * 16;
        s.scr = reinterpret_cast<float4 *>(base);
        base += cap * 16;
        if (rad_smem) {
            s.rad = reinterpret_cast<double *>(base);
            base += cap * 8;
        }
    }
    s.bits = reinterpret_cast<unsigned long long *>(base);
    base += sizeof(unsigned long long) * kFH;
    s.id = reinterpret_cast<int *>(base);
    base += cap * 4;
    s.misc = reinterpret_cast<int *>(base);
    base += 16;
    s.K = reinterpret_cast<LeanConsts *>(base);
    base += (sizeof(LeanConsts) + 15) & ~(size_t)15;
    s.xinfo = reinterpret_cast<unsigned short *>(base);
    s.ecell = s.xinfo + kFC;
    return s;
}

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
    int x0, y0;     // its first cell column / local row
};
__device__ __forceinline__ TilePos tile_pos(const TileGeom &tg, int tile)
{
    TilePos t;
    t.tyi = tile / tg.ntx;
    t.txi = tile - t.tyi * tg.ntx;
    t.tw = t.txi == tg.ntx - 1 ? tg.wlast : kTX;
    t.th = t.tyi == tg.nty - 1 ? tg.hlast : kTY;
    t.x0 = t.txi * kTX;
    t.y0 = t.tyi * kTY;
    return t;
}

// The other counter buffer goes back to zero for the cells of this tile -- and the ghost cells beside the first /
// last tile column -- (nobody reads it: its last consumer completed before the chain of this sweep began).
__device__ __forceinline__ void zero_other_counters(const SweepArgs &a, const TilePos &tp)
{
    const int lo = tp.txi == 0 ? -1 : 0, hi = tp.tw + (tp.txi == a.tg.ntx - 1 ? 1 : 0);
    for (int k = threadIdx.x; k < kFW * kTY; k += kTileThreads) {
        const int r = k / kFW, x = k - r * kFW - 1;
        if (r < tp.th && x >= lo && x < hi) a.czero[(tp.y0 + r) * a.ps + tp.x0 + x + 1] = 0ull;
    }
}

// Stream the frame of one tile into shared memory: counters + plane 0 in frame-cell order, the disks of
// planes 1.. into the extras list.  `store(pos, X, Yl, state, radius, id)` writes one record (X, Yl = the
// cell it is FILED under, a ring cell at the periodic edge being the opposite edge of the grid),
// `store_empty(pos)` marks a cell without a disk.  SORT: the disks of a cell are placed in ascending
// particle id (plane 0 = the smallest), so that sums over a neighbourhood run in one fixed order.
// Ends with a barrier; returns false (for the whole CTA) when the extras do not fit.
template <bool SORT, class Store, class Empty>
__device__ __forceinline__ bool frame_load(const SweepArgs &a, const CellSmem &s, const TilePos &tp, const bool radii,
                                           Store store, Empty store_empty)
{
    const int tid = threadIdx.x;
    const int nx = a.b.nx, nl = a.b.nl;
    constexpr int kLoads = (kFC + kTileThreads - 1) / kTileThreads;
    int pc[kLoads], n[kLoads], id0[kLoads];
    short Xs[kLoads], Ys[kLoads];
    double4 st0[kLoads];
    double rad0[kLoads];
#pragma unroll
    for (int q = 0; q < kLoads; q++) {
        const int fc = tid + q * kTileThreads;
        pc[q] = -1;
        n[q] = 0;
        id0[q] = -1;
        Xs[q] = Ys[q] = 0;
        st0[q] = make_double4(0, 0, 0, 0);
        rad0[q] = a.rad0;
        if (fc < kFC) {
            const int fy = fc / kFW, fx = fc - fy * kFW;
            int X = tp.x0 + fx - 1, Yl = tp.y0 + fy - 1;
            bool valid = fx <= tp.tw + 1 && fy <= tp.th + 1;
            X = X < 0 ? X + nx : (X >= nx ? X - nx : X);
            if (Yl < 0 || Yl >= nl) {
                if (a.slab) valid = false;   // a slab's rows are not periodic: rows 0 and nl-1 ARE the halo
                Yl = Yl < 0 ? Yl + nl : Yl - nl;
            }
            if (valid) {
                const int c = Yl * a.ps + X + 1;
                pc[q] = c;
                Xs[q] = (short)X;
                Ys[q] = (short)Yl;
                const unsigned long long word = a.ccnt[c];
                n[q] = (int)(word >> 32);
                st0[q] = ld_sector(a.pst + c);
                id0[q] = (int)(unsigned int)word;   // one disk: its id; more: the sum of their ids
                if (radii) rad0[q] = a.prad[c];
            }
        }
    }
    // the loads are in flight; the list head, the row bits (zeroed by the caller) and whatever else the
    // caller prepared become visible to the CTA meanwhile
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kLoads; q++) {
        const int fc = tid + q * kTileThreads;
        if (fc >= kFC) continue;
        const int nn = min(n[q], kSlotK);
        if (nn == 0) {
            store_empty(fc);
            continue;
        }
        const int c = pc[q];
        const int X = Xs[q], Yl = Ys[q];
        if (nn == 1) {
            store(fc, X, Yl, st0[q], rad0[q], id0[q]);
            continue;
        }
        // more than one disk in the cell: the others go to the extras list, contiguously
        const int m = nn - 1;
        const int e0 = atomicAdd(&s.misc[0], m);
        if (e0 + m > a.tg.ecap) {
            s.misc[1] = 1;
            store_empty(fc);
            continue;
        }
        const int fy = fc / kFW, fx = fc - fy * kFW;
        s.xinfo[fc] = (unsigned short)(e0 | (m << 12));
        atomicOr(&s.bits[fy], 1ull << fx);
        unsigned int idp0 = (unsigned int)id0[q];   // plane 0's id = the sum - the ids of planes 1..
        for (int k = 1; k < nn; k++) idp0 -= (unsigned int)a.pid[(size_t)k * a.ncp + c];
        for (int k = 0; k < nn; k++) {
            const size_t slot = (size_t)k * a.ncp + c;
            const int idk = k == 0 ? (int)idp0 : a.pid[slot];
            int rank = k;
            if (SORT) {
                rank = 0;
                for (int k2 = 0; k2 < nn; k2++) rank += (k2 == 0 ? (int)idp0 : a.pid[(size_t)k2 * a.ncp + c]) < idk ? 1 : 0;
            }
            const double4 stk = k == 0 ? st0[q] : ld_sector(a.pst + slot);
            const double radk = k == 0 ? rad0[q] : (radii ? a.prad[slot] : a.rad0);
            const int pos = rank == 0 ? fc : kFC + e0 + rank - 1;
            store(pos, X, Yl, stk, radk, idk);
            if (rank > 0) s.ecell[e0 + rank - 1] = (unsigned short)fc;
        }
    }
    __syncthreads();
    return s.misc[1] == 0;
}

// ---- P2, persistent: frames prefetched by TMA bulk copies one tile ahead -------------------------
// Shared memory of one CTA of k_cell_sweep.  Index p < kFC = the plane-0 disk of frame cell p,
// p >= kFC = extra number p - kFC (a disk of planes 1..); cap = kFC + tg.ecap.
//   st[2]   double4[kFC]     FP64 states of plane 0, frame-cell order, DOUBLE-BUFFERED: the bulk copies of
//                            the next tile land while this one is swept (read in place by the exact stage)
//   rrad[2] double[kFC]      radii of plane 0 (only when the radii are not all exactly rad0)
//   rcnt    u64[kFC]         cell words (count << 32 | sum of ids) of the frame as copied (consumed by the convert pass)
//   scr     float4[cap]      screening records (NaN: empty cell)
//   id      int[cap]         particle ids, -1: empty (local ids in a slab context)
//   xst     double4[ecap], xrad double[ecap]   FP64 states / radii of the extras
//   bits, xinfo, ecell       as CellSmem
static_assert((kFW * 8) % 16 == 0, "bulk copies move 16-byte granules");
static_assert(kFW + kFH <= kTileThreads && 32 + kFH <= 60, "frame_tables / the list reset spread their work over the first 64 threads");
static_assert(kFW == 34, "frame_convert divides by 34 with a multiply-shift");
struct SweepSmem {
    double4 *st;    // buffer b at st + b * kFC
    double *rrad;   // buffer b at rrad + b * kFC
    unsigned long long *rcnt;
    float4 *scr;
    int *id;
    double4 *xst;
    double *xrad;
    unsigned long long *bits;   // [2][kFH]: the set of tile k is cleared while tile k+1 uses the other
    unsigned short *xinfo, *ecell;
    unsigned short *items;   // [kTX * kTY + ecap] the particles this tile predicts, compacted: p (see above)
    int *misc;   // per set (4 ints each): [0] extras listed, [1] decline, [2] the next tile, [3] items listed
    double *ctr; // [kFW] x of the centre of the cell frame column fx is FILED under, [kFH] y of frame row fy's
    int *rowl;   // [kFH] local row of frame row fy, -1: none
    LeanConsts *K;
    uint64_t *mbar;   // [2]
};

__host__ __device__ inline size_t sweep_smem_bytes(int ecap, int rad_smem)
{
    const size_t cap = (size_t)kFC + ecap;
    size_t b = 2 * (size_t)kFC * 32;                       // st
    b += (size_t)ecap * 32;                                // xst
    b += cap * 16;                                         // scr
    if (rad_smem) b += 2 * (size_t)kFC * 8 + (size_t)ecap * 8 + 16;   // rrad, xrad
    b += 2 * sizeof(unsigned long long) * kFH + 16;        // bits, mbar
    b += sizeof(double) * (kFW + kFH) + 16;                // ctr
    b += sizeof(int) * kFH + 16;                           // rowl
    b += (size_t)kFC * 8;                                  // rcnt
    b += cap * 4 + 16;                                     // id
    b += 32;                                               // misc
    b += (sizeof(LeanConsts) + 15) & ~(size_t)15;
    b += sizeof(unsigned short) * (kFC + ecap);            // xinfo, ecell
    b += sizeof(unsigned short) * (kTX * kTY + ecap);      // items
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ SweepSmem carve_sweep(unsigned char *base, int ecap, int rad_smem)
{
    const size_t cap = (size_t)kFC + ecap;
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    SweepSmem s;
    s.st = reinterpret_cast<double4 *>(base);
    base += 2 * (size_t)kFC * 32;
    s.xst = reinterpret_cast<double4 *>(base);
    base += (size_t)ecap * 32;
    s.scr = reinterpret_cast<float4 *>(base);
    base += cap * 16;
    s.rrad = s.xrad = nullptr;
    if (rad_smem) {
        s.rrad = reinterpret_cast<double *>(base);
        base += 2 * (size_t)kFC * 8;
        s.xrad = reinterpret_cast<double *>(base);
        base += up16((size_t)ecap * 8);
    }
    s.bits = reinterpret_cast<unsigned long long *>(base);
    base += 2 * sizeof(unsigned long long) * kFH;
    s.mbar = reinterpret_cast<uint64_t *>(base);
    base += 16;
    s.ctr = reinterpret_cast<double *>(base);
    base += up16(sizeof(double) * (kFW + kFH));
    s.rowl = reinterpret_cast<int *>(base);
    base += up16(sizeof(int) * kFH);
    s.rcnt = reinterpret_cast<unsigned long long *>(base);
    base += (size_t)kFC * 8;
    s.id = reinterpret_cast<int *>(base);
    base += up16(cap * 4);
    s.misc = reinterpret_cast<int *>(base);
    base += 32;
    s.K = reinterpret_cast<LeanConsts *>(base);
    base += (sizeof(LeanConsts) + 15) & ~(size_t)15;
    s.xinfo = reinterpret_cast<unsigned short *>(base);
    s.ecell = s.xinfo + kFC;
    s.items = s.ecell + ecap;
    return s;
}

// local row of frame row fy, -1: no such row (beyond the tile's frame, or outside a slab's rows: a slab's
// rows are not periodic, rows 0 and nl-1 ARE the halo)
__device__ __forceinline__ int frame_row(const SweepArgs &a, const TilePos &tp, int fy)
{
    if (fy > tp.th + 1) return -1;
    int Yl = tp.y0 + fy - 1;
    if (Yl < 0 || Yl >= a.b.nl) {
        if (a.slab) return -1;
        Yl = Yl < 0 ? Yl + a.b.nl : Yl - a.b.nl;
    }
    return Yl;
}

// Fire the bulk copies of one tile's frame (warp 0, converged): lane fy copies frame row fy -- with the
// ghost columns of the planes a row is ONE contiguous run starting at padded column x0, periodic edge
// included: kFW states, kFW cell words (and kFW radii).  Completion is counted in bytes on mbar[buf].
__device__ __forceinline__ void frame_issue(const SweepArgs &a, const SweepSmem &s, const TilePos &tp, int buf,
                                            const bool radii)
{
    const int lane = threadIdx.x & 31;
    const int Yl = lane < kFH ? frame_row(a, tp, lane) : -1;
    const int nrows = __popc(__ballot_sync(0xffffffffu, Yl >= 0));
    const uint32_t per_row = kFW * 32u + kFW * 8u + (radii ? kFW * 8u : 0u);
    if (lane == 0) mbar_expect_tx(&s.mbar[buf], (uint32_t)nrows * per_row);
    __syncwarp();
    if (Yl >= 0) {
        const size_t base = (size_t)Yl * a.ps + tp.x0;
        bulk_g2s(s.st + buf * kFC + lane * kFW, a.pst + base, kFW * 32u, &s.mbar[buf]);
        bulk_g2s(s.rcnt + lane * kFW, a.ccnt + base, kFW * 8u, &s.mbar[buf]);
        if (radii) bulk_g2s(s.rrad + buf * kFC + lane * kFW, a.prad + base, kFW * 8u, &s.mbar[buf]);
    }
}

// Per tile, before the convert pass: the centre of the cell every frame column / row is filed under (a ghost
// column is the opposite edge of the grid), the local row of every frame row.  (Once per column instead of
// once per record: the int -> FP64 conversion, the + 0.5 and the product are 3 of the 5 FP64 operations a
// screening coordinate costs.)
__device__ __forceinline__ void frame_tables(const SweepArgs &a, const SweepSmem &s, const TilePos &tp)
{
    const int t = threadIdx.x;
    if (t < kFW) {
        int X = tp.x0 + t - 1;
        X = X < 0 ? X + a.b.nx : (X >= a.b.nx ? X - a.b.nx : X);
        s.ctr[t] = __dmul_rn((double)X + 0.5, a.b.csx);
    } else if (t < kFW + kFH) {
        const int fy = t - kFW;
        const int Yl = frame_row(a, tp, fy);
        s.rowl[fy] = Yl;
        s.ctr[kFW + fy] = __dmul_rn((double)edmd_global_row(a.b, Yl < 0 ? 0 : Yl) + 0.5, a.b.csy);
    }
}

// The frame has landed: derive the screening records of plane 0 (in place order), list the extras (sets
// bits / misc of set `buf`).  Ends with a barrier; returns false (for the whole CTA) when the extras do not fit.
template <bool TWO>
__device__ __forceinline__ bool frame_convert(const SweepArgs &a, const SweepSmem &s, const TilePos &tp, int buf,
                                              const bool radii)
{
    const int tid = threadIdx.x;
    const float fnan = __int_as_float(0x7fffffff);
    const double4 *st0 = s.st + buf * kFC;
    unsigned long long *bits = s.bits + buf * kFH;
    int *misc = s.misc + buf * 4;
    auto record = [&](int fx, int fy, const double2 &xy, const double2 &vv, double rad) {
        // screening record, relative to the centre of the filed cell (lean.cuh)
        float4 r;
        r.x = __double2float_rn(__dsub_rn(xy.x, s.ctr[fx]));
        r.y = __double2float_rn(__dsub_rn(xy.y, s.ctr[kFW + fy]));
        r.z = __double2float_rn(vv.x);
        r.w = __double2float_rn(vv.y);
        if (TWO) r.w = __int_as_float((__float_as_int(r.w) & ~1) | (edmd_same_class(rad, a.rad0) ? 0 : 1));
        return r;
    };
#pragma unroll 1
    for (int fc = tid; fc < kFC; fc += kTileThreads) {
        const int fy = (fc * 241) >> 13, fx = fc - fy * kFW;   // fc / kFW for fc < 4096 (kFW = 34)
        const int Yl = s.rowl[fy];
        const unsigned long long word = s.rcnt[fc];
        const int nn = (Yl >= 0 && fx <= tp.tw + 1) ? min((int)(word >> 32), kSlotK) : 0;
        if (nn == 0) {
            s.scr[fc] = make_float4(fnan, fnan, fnan, fnan);
            s.id[fc] = -1;
            continue;
        }
        const double2 xy = reinterpret_cast<const double2 *>(st0 + fc)[0];
        const double2 vv = reinterpret_cast<const double2 *>(st0 + fc)[1];
        s.scr[fc] = record(fx, fy, xy, vv, radii ? s.rrad[buf * kFC + fc] : a.rad0);
        if (nn == 1) {
            s.id[fc] = (int)(unsigned int)word;   // one disk: the sum of ids is its id
            continue;
        }
        // more than one disk in the cell: the others go to the extras list, contiguously
        const int m = nn - 1;
        const int e0 = atomicAdd(&misc[0], m);
        if (e0 + m > a.tg.ecap) {
            misc[1] = 1;
            continue;
        }
        s.xinfo[fc] = (unsigned short)(e0 | (m << 12));
        atomicOr(&bits[fy], 1ull << fx);
        const int c = Yl * a.ps + tp.x0 + fx;   // padded cell id (ghost columns included)
        unsigned int idp0 = (unsigned int)word;   // plane 0's id = the sum - the ids of planes 1..
        for (int k = 1; k < nn; k++) {
            const size_t slot = (size_t)k * a.ncp + c;
            const double4 stk = ld_sector(a.pst + slot);
            const double radk = radii ? a.prad[slot] : a.rad0;
            const int idk = a.pid[slot];
            const int e = e0 + k - 1;
            s.xst[e] = stk;
            if (radii) s.xrad[e] = radk;
            s.scr[kFC + e] = record(fx, fy, make_double2(stk.x, stk.y), make_double2(stk.z, stk.w), radk);
            s.id[kFC + e] = idk;
            s.ecell[e] = (unsigned short)fc;
            idp0 -= (unsigned int)idk;
        }
        s.id[fc] = (int)idp0;
    }
    zero_other_counters(a, tp);
    __syncthreads();
    // The particles this tile predicts, COMPACTED: 12 % of the cells are empty (and a halo row's disks are never
    // predicted), so a warp that walks a tile row cell by cell runs at 28 of 32 lanes and the extras take a trip of
    // their own; from the list every trip is full (-15 % trips).  Order inside the list is irrelevant.
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int r = warp; r < tp.th; r += kTileWarps) {
            const int c = (r + 1) * kFW + lane + 1;
            const int id = lane < tp.tw ? s.id[c] : -1;
            const bool live = id >= 0 && id < a.n_owned;
            const unsigned m = __ballot_sync(0xffffffffu, live);
            int base = 0;
            if (lane == 0) base = atomicAdd(&misc[3], __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (live) s.items[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)c;
        }
        const int ne = min(misc[0], a.tg.ecap);
        for (int e = tid; e < ne; e += kTileThreads) {
            const int c = s.ecell[e];
            const int fy = (c * 241) >> 13, fx = c - fy * kFW;
            const int id = s.id[kFC + e];
            if (fx >= 1 && fx <= tp.tw && fy >= 1 && fy <= tp.th && id >= 0 && id < a.n_owned)   // (not an extra of the ring)
                s.items[atomicAdd(&misc[3], 1)] = (unsigned short)(kFC + e);
        }
    }
    __syncthreads();
    return misc[1] == 0;
}

// ---- sweep: every particle of the tile ---------------------------------------------
template <bool TWO>
__device__ __forceinline__ void cell_tile(const SweepArgs &a, const SweepSmem &s, const TilePos &tp, int buf,
                                          const bool radii, const LeanConsts &K)
{
    const int tid = threadIdx.x;
    const float fnan = __int_as_float(0x7fffffff);
    const unsigned long long *bits = s.bits + buf * kFH;
    const int nitems = (a.dbg & 4) ? 0 : s.misc[buf * 4 + 3];
    const double4 *st0 = s.st + buf * kFC;
    const double *rad0p = s.rrad + buf * kFC;
    // FP64 state / radius of record p: plane 0 in the copied frame, extras in their list
    auto xy_of = [&](int p) { return reinterpret_cast<const double2 *>(p < kFC ? st0 + p : s.xst + (p - kFC))[0]; };
    auto vv_of = [&](int p) { return reinterpret_cast<const double2 *>(p < kFC ? st0 + p : s.xst + (p - kFC))[1]; };
    auto rad_of = [&](int p) { return radii ? (p < kFC ? rad0p[p] : s.xrad[p - kFC]) : a.rad0; };
#pragma unroll 1
    for (int w = tid; w < nitems; w += kTileThreads) {
        // work item: a particle of the tile (plane 0 of a cell, or an extra), from the compacted list
        const int p = s.items[w];
        const int c = p < kFC ? p : s.ecell[p - kFC];
        const int fy = (c * 241) >> 13, fx = c - fy * kFW;
        const int id = s.id[p];
        const float4 own = s.scr[p];
        // dx = rx_j - (rx_i - k csx), k = column(j) - column(i) in {-1, 0, 1}
        const float pxs[3] = {__fadd_rn(own.x, K.csx), own.x, __fsub_rn(own.x, K.csx)};
        const float pys[3] = {__fadd_rn(own.y, K.csy), own.y, __fsub_rn(own.y, K.csy)};
        const bool own1 = TWO && (__float_as_int(own.w) & 1);
        const float cc_a = own1 ? K.Cc01 : K.Cc00, cc_b = own1 ? K.Cc11 : K.Cc01;
        float lo1 = __int_as_float(0x7f800000), lo2 = __int_as_float(0x7f800000);   // two smallest bounds
        int idx = -1;
        auto screen = [&](int q, float px, float py, bool maybe_self) {
            const float4 qq = s.scr[q];
            const float dx = __fsub_rn(qq.x, px), dy = __fsub_rn(qq.y, py);
            const float dvx = __fsub_rn(qq.z, own.z), dvy = __fsub_rn(qq.w, own.w);
            const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
            const float v2 = __fmaf_rn(dvy, dvy, __fmul_rn(dvx, dvx));
            const float bb = __fmaf_rn(dy, dvy, __fmul_rn(dx, dvx));
            const float psi = __fmaf_rn(d2, K.inv_rho2, __fmaf_rn(v2, K.inv_om2, 1.0f));
            const float clo = __fmaf_rn(d2, K.A, TWO ? ((__float_as_int(qq.w) & 1) ? -cc_b : -cc_a) : -cc_a);
            const float det = __fmaf_rn(-v2, clo, __fmul_rn(bb, bb));
            const float detu = __fmaf_rn(__fmul_rn(K.Kdet, psi), psi, det);
            const float bup = __fmaf_rn(K.Kb, psi, -bb);
            const float sq = __fmul_rn(detu, rsqrt_f32(detu));   // NaN when det_up <= 0: dropped below
            const float den = __fadd_rn(sq, bup);
            float tl = __fmul_rn(clo, rcp_f32(den));
            // an empty cell is a NaN record: bup is NaN, the test fails, the candidate is dropped
            const bool keepc = maybe_self ? ((bup > 0.0f) && (q != p)) : (bup > 0.0f);
            tl = keepc ? tl : fnan;
            // (lo1, lo2) <- two smallest of {lo1, lo2, tl}; min / max drop NaN operands
            idx = tl < lo1 ? q : idx;
            lo2 = fmaxf(lo1, fminf(lo2, tl));
            lo1 = fminf(lo1, tl);
        };
        // extras in the 3 x 3 block: bit 3 j + k
        unsigned xm = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) xm |= ((unsigned)(bits[fy + j - 1] >> (fx - 1)) & 7u) << (3 * j);
        if (!(a.dbg & 1)) {
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int k = 0; k < 3; k++) screen(c + (j - 1) * kFW + (k - 1), pxs[k], pys[j], j == 1 && k == 1);
            unsigned m = xm;
#pragma unroll 1
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int j = (b * 11) >> 5, k = b - 3 * j;
                const unsigned info = s.xinfo[c + (j - 1) * kFW + (k - 1)];
                const int q0 = kFC + (info & 0xfff), qn = info >> 12;
                const float px = k == 0 ? pxs[0] : (k == 1 ? pxs[1] : pxs[2]);
                const float py = j == 0 ? pys[0] : (j == 1 ? pys[1] : pys[2]);
#pragma unroll 1
                for (int q = q0; q < q0 + qn; q++) screen(q, px, py, true);
            }
        }
        const float second = lo1 > 0.0f ? lo2 : fnan;   // a non-positive smallest bound is never certifiable
        // slab contexts: the partner leaves as the caller's (global) id; fetch the winner's now, the load
        // flies while the exact stage computes
        const int wloc = idx >= 0 ? s.id[idx] : -1;
        const int wglob = (wloc >= 0 && a.gid) ? a.gid[wloc] : wloc;
        const double2 mexy = xy_of(p), mev = vv_of(p);
        const double rad_i = rad_of(p);
        if (a.dbg & 2) {
            st_ev(a.ev + id, (double)second + mexy.x, 0.0, idx, 0);
            continue;
        }
        // ---- exact: crossing + the winner's pair time, as the reference computes them ----
        const int X = tp.x0 + fx - 1, Yl = tp.y0 + fy - 1;
        SRec p1;
        p1.x = mexy.x; p1.y = mexy.y; p1.vx = mev.x; p1.vy = mev.y;
        p1.rad = rad_i; p1.id = id; p1.pc = 0;
        const double four_r1 = __dmul_rn(4.0, p1.rad);
        double dtc;
        int dirc;
        crossing_fast<true>(a.b, p1, X, edmd_global_row(a.b, Yl), dtc, dirc);
        double best = EDMD_NEVER;
        int best_id = -1, best_pc = -1, ov_id = -1, ov_pc = -1;
        bool certified = idx < 0;   // no candidate can collide: partner 0 at t + 1e26
        if (idx >= 0) {
            const double2 wxy = xy_of(idx), wv = vv_of(idx);
            SRec p2;
            p2.x = wxy.x; p2.y = wxy.y; p2.vx = wv.x; p2.vy = wv.y;
            p2.rad = rad_of(idx); p2.id = wloc; p2.pc = 0;
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
            auto cand = [&](int q, int scan_cell) {
                SRec p2;
                const double2 qxy = xy_of(q), qv = vv_of(q);
                p2.x = qxy.x; p2.y = qxy.y; p2.vx = qv.x; p2.vy = qv.y;
                p2.rad = rad_of(q); p2.id = s.id[q];
                p2.pc = scan_cell;
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
            };
#pragma unroll 1
            for (int b = 0; b < 9; b++) {
                const int j = (b * 11) >> 5, k = b - 3 * j;
                const int q = c + (j - 1) * kFW + (k - 1);
                if (q != p && s.id[q] >= 0) cand(q, b);   // `p1 != p2` is identity
                if ((xm >> b) & 1u) {
                    const unsigned info = s.xinfo[q];
                    const int q0 = kFC + (info & 0xfff), qn = info >> 12;
#pragma unroll 1
                    for (int e = q0; e < q0 + qn; e++)
                        if (e != p) cand(e, b);
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

// One CTA = a persistent worker.  While it sweeps tile k the bulk copies of tile k+1's frame land in the other
// state buffer, so the memory latency of a frame hides behind the arithmetic of the previous one.  Tiles are
// handed out dynamically (one global counter, reset by the partition kernel; the first gridDim.x tiles are
// taken statically): a static round-robin leaves most SMs idle for the last tile of the slowest workers
// (2232 tiles over 444 workers: six rounds for 5.03 rounds of work, measured +20 %).
template <bool TWO>
__device__ __forceinline__ void cell_worker(const SweepArgs &a, const SweepSmem &s, const bool radii)
{
    const int tid = threadIdx.x;
    const int ntiles = a.tg.ntiles;
    const LeanConsts K = *s.K;
    int buf = 0;
    unsigned phases = 0u;   // bit b = parity of the next completion of mbar[b]
    int tile = blockIdx.x;
    int it = 0;
    const bool stamp = a.ts && blockIdx.x == 0 && tid == 0;   // timing experiments (option 100, bit 32)
#pragma unroll 1
    while (tile < ntiles) {
        const TilePos tp = tile_pos(a.tg, tile);
        if (stamp && it < 6) edmd_stamp(a.ts, 16 + 4 * it);
        // the next tile: asked for now (the round trip hides behind the convert pass), fetched while this
        // one is swept.  ONE tile ahead, not two: what a worker holds when the counter runs out is the tail
        // of the kernel (two tiles held: workers ended over a span of 25 us, measured).  (Also measured: asking
        // for the next tile only AFTER the sweep in the last gridDim.x tiles -- nobody holds a tile at the end, the
        // last fetch exposed -- changes nothing: 61.4 against 59.4-61.5 us; the workers' ends still span ~20 us.)
        int next = ntiles;
        if (tid == 0) next = atomicAdd(a.flags + kFlagTileCtr, 1);
        frame_tables(a, s, tp);
        __syncthreads();
        mbar_wait(&s.mbar[buf], (phases >> buf) & 1u);
        phases ^= 1u << buf;
        if (stamp && it < 6) edmd_stamp(a.ts, 17 + 4 * it);
        if (tid == 0) s.misc[buf * 4 + 2] = next;
        const bool fits = frame_convert<TWO>(a, s, tp, buf, radii);
        if (stamp && it < 6) edmd_stamp(a.ts, 18 + 4 * it);
        if (!fits) {   // more extras than the frame's list holds: decline (no copy is in flight here)
            if (tid == 0) atomicOr(&a.flags[kFlagLeanFail], 1);
            return;
        }
        next = s.misc[buf * 4 + 2];
        if (next < ntiles && tid < 32) frame_issue(a, s, tile_pos(a.tg, next), buf ^ 1, radii);
        // the other set of lists (of the previous tile) goes back to empty for the next one
        if (tid >= 32 && tid < 32 + kFH) s.bits[(buf ^ 1) * kFH + tid - 32] = 0ull;
        if (tid >= 60 && tid < 64) s.misc[(buf ^ 1) * 4 + tid - 60] = 0;
        cell_tile<TWO>(a, s, tp, buf, radii, K);
        if (stamp && it < 6) edmd_stamp(a.ts, 19 + 4 * it);
        it++;
        __syncthreads();   // the records of this tile are dead (the other state buffer has been filling meanwhile)
        tile = next;
        buf ^= 1;
    }
    if (stamp) edmd_stamp(a.ts, 15);
    if (a.ts && tid == 0) {   // timing experiments: the last worker's end, the largest number of tiles per worker
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(a.ts + 14, t);
        atomicMax(a.ts + 13, (unsigned long long)it);
        atomicAdd(a.ts + 12, (unsigned long long)it);
        // histogram of the workers' end times, 2 us bins from the first thread of the partition kernel
        const unsigned long long t4 = *reinterpret_cast<volatile unsigned long long *>(a.ts + 4);
        long long bin = t > t4 ? (long long)((t - t4) / 2000ull) - 16 : 0;
        bin = bin < 0 ? 0 : (bin > 23 ? 23 : bin);
        atomicAdd(a.ts + 40 + bin, 1ull);
    }
}

__global__ void __launch_bounds__(kTileThreads, kTileCtas)
k_cell_sweep(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char cell_smem[];
    const SweepSmem s = carve_sweep(cell_smem, a.tg.ecap, a.rad_smem);
    const int tid = threadIdx.x;
    // before the partition is complete: what does not depend on it
    if (tid < 2 * kFH) s.bits[tid] = 0ull;
    if (tid < 8) s.misc[tid] = 0;
    if (tid == 0) {
        mbar_init(&s.mbar[0], 1);
        mbar_init(&s.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    edmd_pdl_wait();
    if (blockIdx.x == 0 && tid == 0) edmd_stamp(a.ts, 7);
    const int classes = a.flags[kFlagNotMono];   // 0: one radius, 1: two classes, more: not eligible
    const double rad1 = __longlong_as_double(*reinterpret_cast<const long long *>(a.flags + kFlagRad1));
    const bool declined = a.flags[kFlagLeanFail] != 0;
    const bool bad = a.flags[kFlagInsane] != 0 || classes > 1 || (classes == 1 && !a.rad_smem);
    if (blockIdx.x == 0 && tid == 0) edmd_stamp(a.ts, 8);
    if (declined) return;   // a cell overflowed: the host re-runs the sweep on another path
    if (bad) {
        if (tid == 0) atomicOr(&a.flags[kFlagLeanFail], 1);   // not eligible: decline
        return;
    }
    const bool radii = classes == 1;
    // the first frame is on its way while one thread derives the screening constants
    if (tid < 32) frame_issue(a, s, tile_pos(a.tg, blockIdx.x), 0, radii);
    if (tid == kTileThreads - 1) *s.K = make_consts(a.b, a.rad0, rad1, classes == 1, __int_as_float(a.flags[kFlagVmax]));
    __syncthreads();
    if (!s.K->ok) {   // a velocity scale outside the FP32-safe range: decline -- once the copy in flight has landed
        mbar_wait(&s.mbar[0], 0u);
        if (tid == 0) atomicOr(&a.flags[kFlagLeanFail], 1);
        return;
    }
    // classes == 1: the radii are spread inside their classes (a reference-grown system): the exact
    // stage takes every disk's own FP64 radius; the class bit is only looked at when a second class exists
    if (classes == 1 && rad1 > 0.0) cell_worker<true>(a, s, true);
    else cell_worker<false>(a, s, radii);
}

// ---- K4 on the cell slots: computeBOOPCutoff, src/boop.c:61-107 ---------------------------
// Same frame; positions only.  Per particle the reference's 3 x 3 scan with its own arithmetic (FP64
// differences, PBC, r2 < r_c^2: neighbour counts are integers identical to the reference's; the same
// truncation: cells are ~2.0 wide, r_c = 2.5, neighbours two cells away are never seen).
// e^{ik theta} by complex powers; the FP64 sums run in one fixed order (rows, cells, ascending particle
// id inside a cell): every output bit is reproducible.  Works for any radii; a frame whose extras do
// not fit sets kFlagBoopFail and the host falls back to the row kernel.
struct BoopTileArgs {
    SweepArgs s;
    double rc2;
    double4 *rec;   // one 32-byte sector per particle id: (q5, q6, q7, q6_arg | neighbours in the 8 low mantissa bits)
};
// atan2(y, x) to ~1e-14 absolute (the gate on q6_arg is 1e-10; CUDA's atan2 is an IEEE division plus a
// 20-term polynomial): fold into the first octant, ONE rotation by -pi/4 when the angle is above pi/8
// (hi + lo, lo - hi: the common factor sqrt(2) drops out of the quotient), the quotient from the MUFU
// reciprocal seed + two Newton steps, atan(t) = t P(t^2) on |t| <= tan(pi/8) with a degree-8 interpolant
// at Chebyshev nodes (max error 9.5e-15 in FP64 Horner form, tests/test_boop_model.py).
__constant__ double kAtanC[9] = {0x1.fffffffffff0fp-1, -0x1.55555554e5fc5p-2, 0x1.99999911c1573p-3,
                                 -0x1.2492291945813p-3, 0x1.c714d3e720df5p-4, -0x1.73d9cba10d56bp-4,
                                 0x1.35ced8de982a2p-4, -0x1.e13ac280a9accp-5, 0x1.f657ae7e09908p-6};
__device__ __forceinline__ double atan2_gate(double y, double x)
{
    const double ax = fabs(x), ay = fabs(y);
    const double hi = fmax(ax, ay), lo = fmin(ax, ay);
    const bool rot = lo > __dmul_rn(0.41421356237309503, hi);   // tan(pi/8)
    const double h2 = rot ? __dadd_rn(hi, lo) : hi, l2 = rot ? __dsub_rn(lo, hi) : lo;
    double r = rcp_seed(h2);
    r = __fma_rn(r, __fma_rn(-h2, r, 1.0), r);
    r = __fma_rn(r, __fma_rn(-h2, r, 1.0), r);
    const double t = __dmul_rn(l2, r), u = __dmul_rn(t, t);
    double p = kAtanC[8];   // (constant-bank operands of the multiply-adds: no registers, no moves per trip)
#pragma unroll
    for (int k = 7; k >= 0; k--) p = __fma_rn(p, u, kAtanC[k]);
    double a = __dmul_rn(t, p);
    a = rot ? __dadd_rn(a, 0.78539816339744831) : a;
    a = ay > ax ? __dsub_rn(1.5707963267948966, a) : a;
    a = x < 0.0 ? __dsub_rn(3.1415926535897931, a) : a;
    a = h2 > 0.0 ? a : 0.0;   // atan2(0, 0) = 0
    return copysign(a, y);
}

// v when mask == -1, +0.0 when mask == 0
__device__ __forceinline__ double keep_bits(double v, int mask)
{
    return __hiloint2double(__double2hiint(v) & mask, __double2loint(v) & mask);
}

// 1/sqrt(x) for a positive, normal x: MUFU seed (~2^-20) + one third-order step y (1 + e/2 + 3 e^2/8),
// e = 1 - x y^2: relative error ~ (5/16) e^3, below 2^-52 (five operations against eight for two Newton steps)
__device__ __forceinline__ double rsqrt_gate(double x)
{
    const double y = rsqrt_seed(x);
    const double e = __fma_rn(-__dmul_rn(x, y), y, 1.0);
    return __fma_rn(y, __dmul_rn(__fma_rn(0.375, e, 0.5), e), y);
}

// psi6 of every particle of one tile (frame in shared memory).  The nine plane-0 neighbours sit at fixed
// offsets of the frame array: eight unrolled, BRANCH-FREE visits (an empty cell is a position 1e300 away; a
// candidate beyond r_c contributes z = 0), so that the FP64 chains of different candidates overlap -- with a
// branch per candidate the kernel waited on one dependent chain at a time (ncu: `wait` 3.3 stalls per issue,
// FP64 pipe 55 % busy) -- then the listed extras of the block.
template <bool WRAP>
__device__ __forceinline__ void boop_tile(const BoopTileArgs &ba, const CellSmem &s, const TilePos &tp)
{
    const SweepArgs &a = ba.s;
    const int tid = threadIdx.x;
    const int tw = tp.tw, th = tp.th;
    const int ne = s.misc[0];
    const int nrow_items = th * 32;
    const int nitems = nrow_items + ne;
    const double half_lx = a.b.half_lx, half_ly = a.b.half_ly, lx = a.b.lx, ly = a.b.ly, rc2 = ba.rc2;
#pragma unroll 1
    for (int w = tid; w < nitems; w += kTileThreads) {
        int p, c, fx, fy;
        if (w < nrow_items) {
            fy = (w >> 5) + 1;
            fx = (w & 31) + 1;
            c = fy * kFW + fx;
            p = c;
            if (fx > tw) continue;
        } else {
            const int e = w - nrow_items;
            c = s.ecell[e];
            fy = c / kFW;
            fx = c - fy * kFW;
            p = kFC + e;
            if (fx < 1 || fx > tw || fy < 1 || fy > th) continue;
        }
        const int id = s.id[p];
        if (id < 0 || id >= a.n_owned) continue;   // empty cell; halo copy from a neighbouring slab
        const double2 me = s.xy[p];
        // sums: z^6, and A = sum Re(z) z^6, B = sum Im(z) z^6 -- with |z| = 1, z^7 + z^5 = 2 Re(z) z^6 and
        // z^7 - z^5 = 2 i Im(z) z^6: sum7 = A + iB, sum5 = A - iB (four fused multiply-adds instead of two
        // complex products and four additions)
        double s6r = 0, s6i = 0, Ar = 0, Ai = 0, Br = 0, Bi = 0;
        int nb = 0, npos = 0, nvis = 8;
        auto visit = [&](int q) {
            const double2 qq = s.xy[q];
            // the reference's own operations decide who is a neighbour (src/boop.c:78-84)
            double dx = __dsub_rn(qq.x, me.x), dy = __dsub_rn(qq.y, me.y);
            if (WRAP) {
                dx = min_image(dx, half_lx, lx);
                dy = min_image(dy, half_ly, ly);
            }
            const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            const bool in = r2 < rc2;          // (false for an empty cell, a NaN coordinate)
            const bool pos = in && r2 > 0.0;
            nb += in ? 1 : 0;
            npos += pos ? 1 : 0;   // (nb - npos coincident disks: atan2(0, 0) = 0 in the reference, z = 1, added at the end)
            // e^{ik theta} = ((dx + i dy)/r)^k by complex powers (fused multiply-adds: not parity-critical,
            // gate 1e-10).  Everything is computed for every candidate; one out of range gets 1/r = 0, so z = 0.
            // (Masks on the bits, not `pos ? y : 0.0`: the compiler turns the latter into a branch around the powers.)
            const int keep = pos ? -1 : 0;
            const double y = keep_bits(rsqrt_gate(r2), keep);
            if (WRAP) {   // (the variant that also takes states with NaN / infinite coordinates: 0 * NaN)
                dx = keep_bits(dx, keep);
                dy = keep_bits(dy, keep);
            }
            const double zr = __dmul_rn(dx, y), zi = __dmul_rn(dy, y);
            // |z| = 1: z^2 = (2 zr^2 - 1, 2 zr zi), z^4 likewise from z^2 (three operations each instead of four).
            // For z = 0 this chain gives z^6 = -1: sum6 collects -1 per candidate out of range (put back at the
            // end, nvis counts the visits), A and B collect 0 * (-1).
            const double tr = __dadd_rn(zr, zr);
            const double z2r = __fma_rn(tr, zr, -1.0), z2i = __dmul_rn(tr, zi);
            const double t2 = __dadd_rn(z2r, z2r);
            const double z4r = __fma_rn(t2, z2r, -1.0), z4i = __dmul_rn(t2, z2i);
            const double z6r = __fma_rn(z4r, z2r, -__dmul_rn(z4i, z2i));
            const double z6i = __fma_rn(z4r, z2i, __dmul_rn(z4i, z2r));
            s6r = __dadd_rn(s6r, z6r);
            s6i = __dadd_rn(s6i, z6i);
            Ar = __fma_rn(zr, z6r, Ar);
            Ai = __fma_rn(zr, z6i, Ai);
            Br = __fma_rn(zi, z6r, Br);
            Bi = __fma_rn(zi, z6i, Bi);
        };
        // Visit order: the eight plane-0 neighbours at their fixed offsets (unrolled, branch-free), the plane-0 disk of
        // the own cell when the particle is an extra, then the extras of the block by rows and cells in ascending
        // particle id: one fixed order, so every output bit is reproducible (not the reference's order of
        // additions: the gate is 1e-10)
        unsigned xm = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) xm |= ((unsigned)(s.bits[fy + j - 1] >> (fx - 1)) & 7u) << (3 * j);
#pragma unroll
        for (int b = 0; b < 9; b++)
            if (b != 4) visit(c + (b / 3 - 1) * kFW + (b % 3 - 1));
        if (p >= kFC) {   // `p2->num != p1->num`: an extra sees the plane-0 disk of its own cell
            visit(c);
            nvis++;
        }
#pragma unroll 1
        while (xm) {
            const int b = __ffs(xm) - 1;
            xm &= xm - 1;
            const int j = (b * 11) >> 5, k = b - 3 * j;
            const unsigned info = s.xinfo[c + (j - 1) * kFW + (k - 1)];
            const int q0 = kFC + (info & 0xfff), qn = info >> 12;
#pragma unroll 1
            for (int e = q0; e < q0 + qn; e++)
                if (e != p) {
                    visit(e);
                    nvis++;
                }
        }
        double q5 = 0.0, q6 = 0.0, q7 = 0.0, arg = 0.0;
        if (nb > 0) {
            // coincident disks count as z = 1 in all three sums; sum6 gets back the -1 of every visit that was
            // not a neighbour at a distance (nvis - (nb - nzero) of them)
            const int nzero = nb - npos;
            const double nz = (double)nzero;
            s6r = __dadd_rn(s6r, (double)(nvis - npos + nzero));
            const double s5r = __dadd_rn(__dadd_rn(Ar, Bi), nz), s5i = __dsub_rn(Ai, Br);
            const double s7r = __dadd_rn(__dsub_rn(Ar, Bi), nz), s7i = __dadd_rn(Ai, Br);
            // 1/n: exact reciprocals are not needed (gate 1e-10); |sum| = m2 * rsqrt(m2) instead of an IEEE
            // square root (no overflow: |sum| <= n)
            const double nd = (double)nb;
            double inv_n = rcp_seed(nd);
            inv_n = __fma_rn(inv_n, __fma_rn(-nd, inv_n, 1.0), inv_n);
            inv_n = __fma_rn(inv_n, __fma_rn(-nd, inv_n, 1.0), inv_n);
            auto modulus = [&](double re, double im) {
                const double m2 = __fma_rn(re, re, __dmul_rn(im, im));
                const double m = __dmul_rn(__dmul_rn(m2, rsqrt_gate(m2)), inv_n);
                return m2 > 0.0 ? m : 0.0;
            };
            q5 = modulus(s5r, s5i);
            q6 = modulus(s6r, s6i);
            q7 = modulus(s7r, s7i);
            arg = atan2_gate(s6i, s6r);
        }
        // ONE full-sector store by particle id: (q5, q6, q7, q6_arg) with the neighbour count in the eight low
        // mantissa bits of the argument (at most 9 cells x kSlotK disks; k_unpack_boop clears them again: the argument
        // moves by less than 2^-44 relative, gate 1e-10).  (Five scattered 8-/4-byte stores cost 60 us at N = 10^6 -- a
        // partial write to a sector that is not in L2 makes L2 fetch it from DRAM first --, a second sector for the
        // count 32 more bytes per particle here and in the unpack pass.)
        const double argn = __hiloint2double(__double2hiint(arg), (__double2loint(arg) & ~0xff) | nb);
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(ba.rec + id), "d"(q5), "d"(q6), "d"(q7), "d"(argn) : "memory");
    }
}

#ifndef EDMD_BOOP_CTAS
#define EDMD_BOOP_CTAS (1024 / kTileThreads)   // per SM: 32 warps at 64 registers per thread
#endif
constexpr int kBoopCtas = EDMD_BOOP_CTAS;

__global__ void __launch_bounds__(kTileThreads, kBoopCtas)
k_cell_boop(const __grid_constant__ BoopTileArgs ba)
{
    extern __shared__ __align__(128) unsigned char cell_smem[];
    const SweepArgs &a = ba.s;
    const CellSmem s = carve(cell_smem, a.tg.ecap, 0, true);
    const int tid = threadIdx.x;
    const TilePos tp = tile_pos(a.tg, blockIdx.x);
    if (tid < kFH) s.bits[tid] = 0ull;
    if (tid < 2) s.misc[tid] = 0;
    zero_other_counters(a, tp);
    edmd_pdl_wait();
    const bool declined = a.flags[kFlagBoopFail] != 0;
    __syncthreads();
    if (declined) return;
    const double far = 1e300;
    const bool fits = frame_load<true>(
        a, s, tp, false,
        [&](int pos, int, int, const double4 &st, double, int id) {
            s.xy[pos] = make_double2(st.x, st.y);
            s.id[pos] = id;
        },
        [&](int pos) {   // an empty cell is a disk 1e300 away: r2 = inf, `r2 < rc2` fails by itself, no NaN
            s.xy[pos] = make_double2(far, far);
            s.id[pos] = -1;
        });
    if (!fits) {
        if (tid == 0) atomicOr(&a.flags[kFlagBoopFail], 1);
        return;
    }
    // a tile away from the edges of the grid never sees the periodic image (every particle lies within
    // 1.5 cells of the cell it is filed under -- kFlagInsane -- so |d| <= 4 cells < L/2): PBC() is the identity
    // (frame rows lo .. hi in local rows; in a slab the global row yoff + l wraps somewhere inside the slab)
    const int fr_lo = tp.y0 - 1, fr_hi = tp.y0 + tp.th;
    const bool wrap_y = fr_lo < 0 || fr_hi >= a.b.nl || (a.b.yoff + fr_lo < a.b.ny && a.b.yoff + fr_hi >= a.b.ny);
    const bool wrap = a.flags[kFlagInsane] != 0 || tp.txi == 0 || tp.txi == a.tg.ntx - 1 || wrap_y ||
                      a.b.nx < 12 || a.b.ny < 12;
    if (wrap) boop_tile<true>(ba, s, tp);
    else boop_tile<false>(ba, s, tp);
}

// the ABI's five psi6 arrays from the records (one coalesced pass)
__global__ void __launch_bounds__(256)
k_unpack_boop(int n, const double4 *__restrict__ rec, double *__restrict__ q5, double *__restrict__ q6,
              double *__restrict__ q7, double *__restrict__ q6arg, int32_t *__restrict__ nbr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 v = ld_sector(rec + i);
    const int lo = __double2loint(v.w), nb = lo & 0xff;
    q5[i] = v.x;
    q6[i] = v.y;
    q7[i] = v.z;
    q6arg[i] = __hiloint2double(__double2hiint(v.w), lo & ~0xff);
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

// see edmd_preload_exchange_kernels (halo.cu)
void edmd_preload_sweep_kernels()
{
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_cell_partition);
    cudaFuncGetAttributes(&fa, k_halo_recv_partition);
    cudaFuncGetAttributes(&fa, k_cell_sweep);
    cudaFuncGetAttributes(&fa, k_cell_boop);
    cudaFuncGetAttributes(&fa, k_unpack_events);
    cudaFuncGetAttributes(&fa, k_unpack_boop);
}

bool edmd_tile_eligible(const edmd_ctx *c, int mode)
{
    return edmd_lean_eligible(c, mode) && !c->tile_off && c->pst != nullptr;
}

// Geometry of the tiles + the extras a frame may hold, for a context of nx x nl cells and n particles.
// (Measured and not kept: half-height tiles for half of the workers at the start and for everybody's last
// tiles, to break the workers' lock-step and shorten the tail: 63.5 us against 61.5 us for equal tiles.)
bool edmd_tile_geometry(int nx, int nl, size_t n, TileGeom *out)
{
    TileGeom tg;
    tg.ntx = (nx + kTX - 1) / kTX;
    tg.nty = (nl + kTY - 1) / kTY;
    tg.wlast = nx - (tg.ntx - 1) * kTX;
    tg.hlast = nl - (tg.nty - 1) * kTY;
    tg.ntiles = tg.ntx * tg.nty;
    // Extras per cell = mean occupancy - P(occupied).  Hard disks in cells one diameter wide stay far below
    // a Poisson process of the same density (1.8 - 11 % against 30 %), mixtures with small disks come close to it:
    // the Poisson figure + a margin is provided for, and at least 22 % of the frame.
    const double dens = (double)n / ((double)nx * (double)nl);   // particles per cell
    const double poisson = dens - 1.0 + exp(-dens);
    long long ecap = (long long)(kFC * fmax(0.22, 1.15 * poisson)) + 32;
    ecap = (ecap + 31) & ~31ll;
    if (ecap > 4064) return false;   // xinfo holds 12 bits of extras index
    if (nx > 32000 || nl > 32000) return false;   // frame_load keeps cell coordinates in 16 bits
    tg.ecap = (int)ecap;
    if (sweep_smem_bytes(tg.ecap, 1) > 200 * 1024) return false;   // denser than a CTA's shared memory provides for
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

// A new partition goes to the other counter buffer (zeroed by the consumers of the current one).
void edmd_tile_begin_partition(edmd_ctx *c) { c->cbuf ^= 1; }

// Extras a frame of the sweep kernel lists.  tgeom.ecap provides for a Poisson process of the system's density
// (mixtures with small disks come close to it); disks of ONE radius class in cells one diameter wide stay far
// below (1.8 - 11 % of the cells hold a second disk): 23.5 % of the frame then, which lets one more CTA fit an SM.
static int sweep_ecap(const edmd_ctx *c)
{
    const int mono = ((int)(0.235 * kFC) + 15) & ~15;
    const bool one_class = !c->lean_two || !(c->rad1 > 0.0);   // (spread radii inside ONE class: a reference-grown system)
    return (one_class && mono < c->tgeom.ecap) ? mono : c->tgeom.ecap;
}

// CTAs of the persistent sweep kernel: as many as are resident at once (asked of the runtime: a worker that
// had to wait for a free slot would begin its first tile when the others are through), at most one per tile
static void tile_attrs();
static int sweep_workers(edmd_ctx *c)
{
    const int ntiles = c->tgeom.ntiles;
    const int ecap = sweep_ecap(c), rad_smem = c->lean_two ? 1 : 0;
    if (c->workers_key != ecap * 2 + rad_smem) {   // (asked once per shape: the query costs microseconds of host time)
        const size_t smem = sweep_smem_bytes(ecap, rad_smem);
        tile_attrs();
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cell_sweep, kTileThreads, smem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        c->workers_per_sm = per_sm;
        c->workers_key = ecap * 2 + rad_smem;
    }
    const int sms = c->sm_count > 0 ? c->sm_count : 148;
    return ntiles < sms * c->workers_per_sm ? ntiles : sms * c->workers_per_sm;
}

static PartArgs part_args(edmd_ctx *c, int first, int n)
{
    PartArgs pa;
    pa.first = first; pa.n = n; pa.ncp = c->ncp; pa.dbg = c->tile_dbg; pa.ps = c->ps; pa.nx = c->dbox.nx;
    pa.workers = sweep_workers(c);
    pa.late_wait = 0;
    pa.cid = c->cid; pa.xv = c->xv; pa.rad = c->rad; pa.rad0 = c->rad0;
    pa.flags = c->flags;
    pa.ccnt = c->ccnt + (size_t)c->cbuf * c->ncp;
    pa.pst = c->pst; pa.pid = c->pid; pa.prad = c->prad;
    pa.overlap_key = c->overlap_key;
    pa.ts = c->tile_dbg & 32 ? c->dbg_ts : nullptr;
    return pa;
}

// P1 alone: file the particles [first, first + n) of the resident state in the cell slots of the
// partition begun with edmd_tile_begin_partition
int edmd_launch_tile_partition_range(edmd_ctx *c, int first, int n, bool beside)
{
    // (n == 0, a slab that owns nothing at the moment: the kernel still runs -- one block -- because its first
    // thread resets the sweep's overlap report and tile counter)
    if (n < 0) n = 0;
    PartArgs pa = part_args(c, first, n);
    pa.late_wait = beside ? 1 : 0;
    // normally the first kernel of its chain, launched plainly: whatever precedes it on the stream completes
    // first.  `beside`: behind the halo kernels of the fused exchange, with the programmatic attribute
    edmd_launch(k_cell_partition, dim3(n > 0 ? (n + kPartThreads - 1) / kPartThreads : 1), dim3(kPartThreads), 0,
                c->stream, beside, pa);
    return 1;
}

int edmd_launch_tile_partition(edmd_ctx *c)
{
    edmd_tile_begin_partition(c);
    return edmd_launch_tile_partition_range(c, 0, c->n, false);
}

// slab contexts with the peer-to-peer halo: receive the neighbours' boundary rows (sent by
// edmd_launch_halo_send) and file them in the cell slots in the same kernel
int edmd_launch_halo_recv_partition(edmd_ctx *c)
{
    const int H = c->halo_cap;
    const int e = c->halo_epoch, par = e & 1;
    RecvPartArgs ra;
    ra.p = part_args(c, c->n_owned, 2 * H);
    ra.p.overlap_key = nullptr;
    ra.H = H; ra.epoch = e; ra.ps = c->ps;
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
    sa.dbg = c->tile_dbg;
    sa.rad_smem = c->lean_two ? 1 : 0;   // the host's knowledge; the kernel declines if the device knows better
    sa.ps = c->ps; sa.ncp = c->ncp; sa.slab = c->slab ? 1 : 0;
    sa.gid = c->slab ? c->gid : nullptr;
    sa.flags = c->flags;
    sa.ccnt = c->ccnt + (size_t)c->cbuf * c->ncp;
    sa.czero = c->ccnt + (size_t)(c->cbuf ^ 1) * c->ncp;
    sa.pst = c->pst; sa.pid = c->pid; sa.prad = c->prad;
    sa.ev = c->evrec;
    sa.overlap_key = c->overlap_key;
    sa.ts = c->tile_dbg & 32 ? c->dbg_ts : nullptr;
    return sa;
}

static void tile_attrs()
{
    static unsigned long long attr = 0;   // devices of this process the attributes are set on
    if (!edmd_first_on_device(&attr)) return;
    cudaFuncSetAttribute(k_cell_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_cell_sweep, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_cell_boop, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_cell_boop, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

// the sweep kernel over the cell slots (P2): persistent CTAs, kTileCtas per SM
int edmd_launch_tile_sweep_only(edmd_ctx *c)
{
    SweepArgs sa = sweep_args(c);
    tile_attrs();
    sa.tg.ecap = sweep_ecap(c);
    const size_t smem = sweep_smem_bytes(sa.tg.ecap, sa.rad_smem);
    const int grid = sweep_workers(c);
    edmd_launch(k_cell_sweep, dim3(grid), dim3(kTileThreads), smem, c->stream, c->lean_pdl, sa);
    c->pred_packed = true;
    return 1;
}

int edmd_launch_tile_sweep(edmd_ctx *c, cudaEvent_t between)
{
    if (c->n == 0) return 0;
    edmd_launch_tile_partition(c);
    if (between) cudaEventRecord(between, c->stream);
    return 1 + edmd_launch_tile_sweep_only(c);
}

// K4 on the cell slots; chained: launched right behind a partition (programmatic launch)
int edmd_launch_tile_boop(edmd_ctx *c, double r_c, bool from_keep)
{
    if (c->n == 0) return 0;
    const size_t N = (size_t)c->n_cap;   // the four psi6 arrays lie n_cap apart
    BoopTileArgs ba;
    ba.s = sweep_args(c);
    ba.rc2 = r_c * r_c;   // `r_c*r_c`, a single rounded product
    ba.rec = c->boop_rec;
    tile_attrs();
    edmd_launch(k_cell_boop, dim3(c->tgeom.ntiles), dim3(kTileThreads),
                cell_smem_bytes(c->tgeom.ecap, 0, true), c->stream, c->lean_pdl && !from_keep, ba);
    // halo copies of a slab context are not computed: the unpack covers the owned particles
    const int no = c->n_owned;
    k_unpack_boop<<<(no + 255) / 256, 256, 0, c->stream>>>(no, c->boop_rec, c->boop, c->boop + N, c->boop + 2 * N,
                                                           c->boop + 3 * N, c->boop_nb);
    return 2;
}
