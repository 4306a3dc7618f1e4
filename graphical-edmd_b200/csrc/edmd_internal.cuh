// edmd_internal.cuh -- device-side data layout and kernel launchers shared by
// the translation units of libedmd_cuda.so (sm_100a only).
//
// HBM layout (N particles, cell grid nx x ny), all allocated once per context:
//
//   resident state (original particle order, written by upload / free_fly)
//     xv    double4[N]   (x, y, vx, vy)   one 32-byte sector per particle
//     rad   double [N]
//     vr    double [N]   growth rates (GROW mode only)
//     cid   int32  [N]   PADDED cell id  Y*PS + X + 1,  PS = nx + 3 rounded up to 4
//
//   cell index of the DEFAULT sweep (one or two radius classes, cell_sweep.cu): slot planes over the padded
//   grid, rebuilt by every sweep -- plane s = the s-th arrival of every cell:
//     pst   double4[kSlotK][ny*PS]   FP64 states        pid  int32[kSlotK][ny*PS]   ids (planes 1.. only)
//     prad  double [kSlotK][ny*PS]   radii (only when not all equal)
//     ccnt  uint64 [2][ny*PS]        cell words: count << 32 | sum of ids, double-buffered
//     evrec edmd_ev32[N]             one 32-byte event record per particle id
//
//   cell index of the GENERAL path, rebuilt by every sweep (K0).  The reference's intrusive linked
//   cell list (src/EDMD.c:1906-1920, 2053-2078) becomes a counting sort over a
//   PADDED grid: every row of cells carries a left ghost cell (copy of cell
//   nx-1), the nx real cells, a right ghost cell (copy of cell 0) and 1..4 empty
//   sentinels (PS is a multiple of 4 so every row of off[] is 16-byte aligned
//   and windows of it can be fetched with TMA bulk copies).  With the ghosts the three cells X-1, X, X+1 the reference scans
//   (PBCcellX, src/EDMD.c:2110-2116) are ALWAYS one contiguous run of the
//   cell-ordered array, periodic edge included; with the sentinel, off[] of a
//   row ends in the row total.  Rows start on 32-slot boundaries so a warp of
//   the sweep never straddles two rows.
//     cell_cnt  int32[ny*PS]    histogram (self-cleaning: zeroed by the row scan)
//     rank      int32[N]        arrival rank of particle i inside its cell
//     off       int32[ny*PS]    exclusive scan of the padded row (row-local)
//     cstart    int32[ny*PS]    row_base + off: absolute first slot of every cell
//     row_total int32[ny]       entries in the row (ghosts included)
//     row_base  int32[ny+1]     first slot of each row (multiple of 32)
//     meta      ChunkMeta[cap/32]  per 32-slot chunk: its row, the three staged
//                               row segments and the cell-offset window (Y = -1: empty)
//     spos/saux SPos/SAux[cap]  32 + 16-byte records in cell order (arbitrary order
//                               inside a cell; consumers are order-independent
//                               and break exact ties by the reference's rule:
//                               descending particle id = its linked-list order)
//     svr       double[cap]     growth rates in cell order (GROW only)
//
//   sweep outputs (original particle order)
//     t_cross f64[N], t_coll f64[N], partner i32[N], dir u8[N], ctype u8[N]
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "edmd_cuda.h"
#include "edmd_cuda_bench.h"

struct edmd_dev_box {
    int n, nx, ny, nc;     // ny = rows of the GLOBAL cell grid
    int nl, yoff;          // rows held by this context and the global row of local row 0
                           // (whole system: nl = ny, yoff = 0; slab: owned rows + one halo row each side)
    double lx, ly, half_lx, half_ly;
    double csx, csy, fx, fy;
};

// global cell row of a local row
__host__ __device__ inline int edmd_global_row(const edmd_dev_box &b, int l)
{
    int y = l + b.yoff;
    return y >= b.ny ? y - b.ny : y;
}

// One particle in cell order = a 32-byte kinematic record + a 16-byte tag, in
// two parallel arrays.  Measured on B200 (profiles/microbench/scatter_stores.cu,
// 1M records to random slots, cold L2): one 256-bit + one 128-bit store per
// particle 30.7 us, three 128-bit stores into 48-byte records 54.1 us -- the
// 32-byte record is exactly one L2 sector and is written with a single
// STG.256, so no partial-sector writes.
struct __align__(32) SPos { double x, y, vx, vy; };
struct __align__(16) SAux {
    double rad;
    int id;   // original particle id
    int pc;   // padded cell id
};
static_assert(sizeof(SPos) == 32 && sizeof(SAux) == 16, "record layout");
// register view of one particle
struct SRec {
    double x, y, vx, vy, rad;
    int id, pc;
};
__host__ __device__ inline SRec make_rec(const SPos &p, const SAux &a)
{
    SRec r;
    r.x = p.x; r.y = p.y; r.vx = p.vx; r.vy = p.vy;
    r.rad = a.rad; r.id = a.id; r.pc = a.pc;
    return r;
}

// Per 32-slot chunk of the record array (all slots in one cell row Y): what a
// warp of K1 / K4 stages.  Active (non-ghost) entries of the chunk lie in cell
// columns cfirst .. cfirst+ncells-1; row j = -1,0,1 contributes the records
// [seg_lo[j], seg_lo[j]+seg_len[j]) = columns cfirst-1 .. cfirst+ncells; a
// row-local cell offset o maps to staged index o + delta[j].
enum { kMetaOverflow = 1, kMetaInterior = 2 };
struct __align__(16) ChunkMeta {
    int Y, cfirst, ncells, row_end;
    int seg_lo[3];
    int seg_len[3];
    int delta[3];
    int flags;
    int wstart;   // first column of the staged off[] window (multiple of 4, <= cfirst-1)
    int wlen;     // its length in ints (multiple of 4)
};
static_assert(sizeof(ChunkMeta) == 64, "ChunkMeta must be 64 bytes");

// everything a consumer of the cell index needs (kernel argument block)
struct CellIndex {
    int nx, ny, ps;           // ny = LOCAL rows; ps = nx + 3 rounded up to a multiple of 4
    const int32_t *off;
    const int32_t *row_total;
    const int32_t *row_base;
    const ChunkMeta *meta;
    const SPos *spos;
    const SAux *saux;
    const double *svr;
    const int32_t *flags;     // [3] != 0: some particle is far from its filed cell
    const int32_t *gid;       // global particle ids (slab mode), nullptr = identity
    int n_owned;              // local ids >= n_owned are halo copies: never predicted
};

// device flag words
enum {
    kFlagBadCell = 0, kFlagRescans = 1, kFlagGhosts = 2, kFlagInsane = 3,
    kFlagVmax = 4,       // bits of the largest |velocity component| (float, rounded up)
    kFlagNotMono = 5,    // radii: 0 all EXACTLY rad0, 1 at most two classes (rad0's and kFlagRad1's, see edmd_note_radius), >= 2 more
    kFlagLeanFail = 6,   // the lean sweep declined (state not eligible): redo with the full path
    kFlagWork = 7,       // number of entries in the lean work list (chunks that hold particles)
    kFlagTileCtr = 11,   // next tile of the persistent sweep kernel (reset by the partition kernel)
    kFlagBoopFail = 10,  // the tile psi6 kernel declined (a bucket beyond a CTA's shared memory)
    kFlagRad1 = 8,       // (two words, 8-byte aligned) bits of the first radius seen outside rad0's class, 0 = none
    kFlagCount = 16
};

// ---- cell-slot sweep (cell_sweep.cu): kSlotK planes over the padded cell grid, tiles of kTX x kTY cells ----
#ifndef EDMD_TILE_ROWS
#define EDMD_TILE_ROWS 8
#endif
constexpr int kTX = 32, kTY = EDMD_TILE_ROWS;                    // cells per tile
constexpr int kFW = kTX + 2, kFH = kTY + 2, kFC = kFW * kFH;     // a tile's frame: the tile + a one-cell ring
constexpr int kTileThreads = 16 * kTY;                           // one warp per two tile rows
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileCtas = 768 / kTileThreads;                    // per SM: 24 warps at <= 85 registers per thread
constexpr int kSlotK = 8;                                        // disks one cell can hold (plane s = s-th arrival)
struct TileGeom {
    int ntx, nty;            // tiles per row of tiles / rows of tiles
    int wlast, hlast;        // width of the last tile column, height of the last tile row
    int ecap;                // extras (disks of planes 1..) one frame can list in shared memory
    int ntiles;
};

struct edmd_ctx {
    int device;
    int n;               // particles currently held (slab: owned + halo)
    int n_cap;           // capacity
    int n_owned;         // particles this context predicts (== n unless slab)
    bool slab;
    edmd_box box;
    edmd_dev_box dbox;
    cudaStream_t stream;
    cudaEvent_t ev[4];
    char err[512];
    uint64_t launches;
    bool force_generic;  // EDMD_OPT_FORCE_GENERIC: global-memory exact kernel only

    bool have_state;     // upload done
    bool have_rad;       // resident radii are valid (an upload carried them)
    bool have_pred;      // device predictions valid
    bool have_index;     // cell index matches resident state
    bool index_has_vr;   // ... and carries growth rates
    bool index_lean;     // ... but is the LEAN index (lean.cuh): no spos / saux records
    bool lean_ok;        // resident state is eligible for the lean sweep (monodisperse, sane speeds)
    bool radii_dirty;    // a GROW free flight changed the resident radii since the classes were derived
    bool lean_off;       // EDMD_OPT_NO_LEAN
    bool lean_pdl;       // launch the lean chain with programmatic dependent launch (default on)
    double rad0;         // radius of the first particle = radius class 0 of the lean sweep
    double rad1;         // radius class 1 (valid when lean_two: exactly two radii in the system)
    bool lean_two;
    float vmax;          // largest |velocity component| of the upload
    int pred_mode;       // mode of the last sweep
    bool lean_pending;   // the last sweep was launched on the lean path and not yet confirmed
    uint64_t lean_sweeps; // sweeps confirmed on the lean path
    int lean_declines;   // sweeps the device declined
    bool have_vr;
    double t;            // time of the resident snapshot
    int nghost;          // ghost entries of the current upload
    int sm_count;

    // upload staging (device SoA mirror of the host arrays)
    double *in_soa;      // 5*N doubles: x | y | vx | vy | rad
    int32_t *in_cell;    // 2*N
    // pinned host staging
    void *h_pin;
    size_t h_pin_bytes;

    // resident state
    double4 *xv;
    double *rad, *vr;
    int32_t *cid;
    int32_t *gid;        // slab mode: global particle id of every local particle

    // slab mode: peer-to-peer halo over NVLink (see halo.cu)
    int halo_cap;        // records per inbox
    int halo_epoch;
    char *halo_mem;      // my inboxes + acks (exported through CUDA IPC)
    char *peer_mem[2];   // lower / upper neighbour's halo_mem (IPC mapped, or my own)
    bool peer_opened[2];
    int32_t *halo_cnt;   // [8] device counters: records listed per side, finished blocks
    bool halo_list_at_pack;   // the next pack kernel fills halo_list (peer-to-peer halo configured)
    bool halo_list_valid;     // halo_list matches the resident cell ids (filled by the last upload)
    int32_t *halo_list;  // [2][halo_cap] indices of the owned particles in the two boundary rows
    int nghost_extra;    // upper bound of ghost entries contributed by halo particles

    // cell index
    int ps;              // padded row stride: nx + 3 rounded up to a multiple of 4
    int ncp;             // ny * ps
    size_t cap;          // slots in srec
    int max_chunks;
    int32_t *cell_cnt, *rank, *off, *cstart, *row_total, *row_base;
    ChunkMeta *meta;
    SPos *spos;
    SAux *saux;
    double *svr;
    // lean index (lean.cuh)
    struct LeanRec *lrec;            // 32-byte records in cell order, rowcap slots per cell row
    int rowcap;                      // multiple of 32
    int lean_chunks;                 // nl * rowcap / 32
    int32_t *lwork;                  // chunk ids that hold particles, any order (kFlagWork entries)
    int4 *lchunks;
    int2 *lres;                      // k_screen -> k_resolve: (winner slot, second bound) per slot
    // cell-slot sweep (cell_sweep.cu)
    TileGeom tgeom;
    double4 *pst;                    // [kSlotK][ncp] FP64 states: plane s = the s-th arrival of every padded cell
    int32_t *pid;                    // ... their particle ids
    double *prad;                    // ... and radii (written only when the radii are not all exactly rad0)
    unsigned long long *ccnt;        // [2][ncp] cell words (count << 32 | sum of ids), double-buffered (consumers zero the other buffer)
    int cbuf;                        // the buffer of the current partition
    int workers_key, workers_per_sm; // resident CTAs of the sweep kernel per SM for (extras capacity, radii): asked of the runtime once
    bool boop_tile_off;              // EDMD_OPT_NO_TILE_BOOP
    double4 *boop_rec;               // psi6 records of the tile kernel: one sector per particle id
    edmd_ev32 *evrec;                 // event records of the last tile sweep, by particle id
    bool pred_packed;                // the predictions of the last sweep are in ev[], not yet in the five arrays
    bool tile_off;                   // EDMD_OPT_NO_TILE
    unsigned long long *dbg_ts;      // [16] %globaltimer stamps of the fused exchange chain (timing experiments)
    int tile_dbg;                    // timing experiments (internal option 100): skip parts of k_tile_sweep
    bool index_tile;                 // the last sweep ran on the tile path (no global cell index exists)

    // outputs
    double *t_cross, *t_coll;
    int32_t *partner;
    uint8_t *dir, *ctype;
    unsigned long long *overlap_key;  // min over (i<<32 | j), ~0ull = none
    int32_t *flags;                   // kFlagCount words

    // calendar ingest plan (calendar.cu)
    int32_t *cal_mem;                 // scratch + outputs
    size_t cal_ints;                  // its size
    int cal_tree;                     // events of the last plan that go to the BST
    bool cal_declined;                // a bucket held too many events

    // analysis scratch
    unsigned long long *pcf_counts;   // capacity pcf_cap bins
    char *pcfs_mem;                   // sorted-tile g(r) scratch (analysis_pcf_sorted.cu)
    size_t pcfs_bytes;
    unsigned long long *pcfs_stats;   // [0] pairs binned by the exact path, [1] tile pairs skipped
    int pcf_groups;                   // EDMD_OPT_PCF_GROUPS: 0 automatic, else 1 / 2 / 4
    int pcf_mode;                     // EDMD_OPT_PCF_LEGACY: 0 FP32-decided sorted tiles, 1 plain kernel, 2 FP64-certified sorted tiles
    unsigned long long *pcf_wsum;     // weighted sums (2^-32 fixed point), capacity pcf_wcap bins
    int pcf_wcap;
    int pcf_cap;
    double *thermo_mem;               // kinetic sums: block partials + results (thermostat.cu)
    char *vor_mem;                    // Voronoi grid scratch (analysis_voronoi.cu) + psi6 + area/perimeter
    size_t vor_bytes;
    double *boop;                     // 4*N doubles q5|q6|q7|arg
    int32_t *boop_nb;                 // N
    double *red_partial;              // block partials for deterministic sums
    int red_cap;
    char *flush_buf;                  // L2 flush scratch (bench only)
    size_t flush_cap;
};

inline CellIndex edmd_cell_index(const edmd_ctx *c)
{
    CellIndex g;
    g.nx = c->dbox.nx;
    g.ny = c->dbox.nl;
    g.ps = c->ps;
    g.off = c->off;
    g.row_total = c->row_total;
    g.row_base = c->row_base;
    g.meta = c->meta;
    g.spos = c->spos;
    g.saux = c->saux;
    g.svr = c->svr;
    g.flags = c->flags;
    g.gid = c->slab ? c->gid : nullptr;
    g.n_owned = c->n_owned;
    return g;
}

// number of 32-slot chunks the current index can occupy (host-side bound)
inline int edmd_chunks_bound(const edmd_ctx *c)
{
    long long slots = (long long)c->n + c->nghost + c->nghost_extra + 32ll * c->dbox.nl;
    long long ch = (slots + 31) / 32;
    return (int)(ch < c->max_chunks ? ch : c->max_chunks);
}

// grid of the persistent row-pipeline kernels: every SM full, warps loop over chunks
int edmd_persistent_blocks(const edmd_ctx *c);

// ---- launchers (each returns the number of kernels it launched) ----------
int edmd_launch_pack(edmd_ctx *c, bool have_cells, int first, int count, bool keep_rad);
int edmd_launch_halo_pack(edmd_ctx *c, int side, void *out, int cap, int32_t *count_dev);
int edmd_launch_halo_append_row(edmd_ctx *c, const void *in, int count, int row);
size_t edmd_halo_mem_bytes(int halo_cap);
int edmd_launch_halo_p2p(edmd_ctx *c);
void edmd_preload_exchange_kernels();
int edmd_halo_connect_direct(edmd_ctx *c, edmd_ctx *lower, edmd_ctx *upper);
int edmd_launch_halo_send(edmd_ctx *c, bool chained);
int edmd_launch_halo_recv(edmd_ctx *c);
int edmd_launch_halo_recv_partition(edmd_ctx *c);
void edmd_tile_begin_partition(edmd_ctx *c);
int edmd_launch_tile_partition_range(edmd_ctx *c, int first, int n, bool early);
int edmd_launch_tile_sweep_only(edmd_ctx *c);
int edmd_launch_cell_index(edmd_ctx *c, int mode);
int edmd_launch_predict(edmd_ctx *c, int mode);
int edmd_launch_free_fly(edmd_ctx *c, int mode, double dt);
int edmd_launch_calendar_plan(edmd_ctx *c, double paul_time, double dt_paul, int paul_n, int actual,
                              int32_t *scratch, int32_t *bucket, int32_t *next, int32_t *prev,
                              int32_t *head, int *n_overflow_host_sync);
int edmd_launch_lean_index(edmd_ctx *c);
int edmd_launch_predict_lean(edmd_ctx *c);
bool edmd_lean_eligible(const edmd_ctx *c, int mode);
bool edmd_tile_eligible(const edmd_ctx *c, int mode);
bool edmd_tile_geometry(int nx, int nl, size_t n, TileGeom *out);
int edmd_launch_tile_sweep(edmd_ctx *c, cudaEvent_t between);
int edmd_launch_unpack_events(edmd_ctx *c);
int edmd_launch_tile_partition(edmd_ctx *c);
int edmd_launch_tile_boop(edmd_ctx *c, double r_c, bool from_keep);
int edmd_launch_boop(edmd_ctx *c, double r_c);
int edmd_launch_mean(edmd_ctx *c, const double *v, int n, double *out_dev);
int edmd_launch_pcf_bond_order(edmd_ctx *c, double dr, double max_r, int num_bins, double kx, double ky,
                               double2 *psi, bool phase, unsigned long long *counts, unsigned long long *wsum);
double edmd_pcf_s_threshold(double max_r);   // smallest s whose correctly rounded sqrt reaches max_r (analysis_pcf_sorted.cu)
int edmd_launch_structure_factor(edmd_ctx *c, int velocity, int nqx, int nqy, const double *qx, const double *qy,
                                 double *re, double *im);
size_t edmd_voronoi_scratch_bytes(const edmd_ctx *c, int *gx_out, int *gy_out);
int edmd_launch_voronoi(edmd_ctx *c, char *scratch, int boop, double *q5, double *q6, double *q7, double *q6arg,
                        int32_t *nbr, double *area, double *perim, int32_t *fail_dev, cudaEvent_t before_cells = nullptr);
int edmd_launch_psi6(edmd_ctx *c, const double *q6, const double *arg, double2 *psi);
size_t edmd_thermostat_scratch_doubles();
double *edmd_launch_kinetic(edmd_ctx *c, double T, double *scratch, int *launched);
int edmd_launch_rescale(edmd_ctx *c, const double *red);
int edmd_launch_shift_scale(edmd_ctx *c, double dvx, double dvy, double divisor, const double *red, bool div_from_red,
                            int n_total, double *sums);
double *edmd_launch_kinetic_final(edmd_ctx *c, double T, double *scratch, int n_total);
int edmd_launch_langevin(edmd_ctx *c, double T, double gamma, double dtnoise, unsigned int seed, unsigned int tick);
int edmd_launch_bragg(edmd_ctx *c, int nkx, int nky, const double *kx, const double *ky, const unsigned char *mask,
                      double *re, double *im, double *best_s, long long *best_i);
int edmd_launch_rsqrt_selftest(edmd_ctx *c, unsigned long long *worst_bits_dev);
int edmd_launch_pcf_sorted(edmd_ctx *c, double dr, double max_r, int num_bins, const double *xy, int stride,
                           int n, int part, int nparts, unsigned long long *counts);
int edmd_launch_pcf(edmd_ctx *c, double dr, double max_r, int num_bins, const double *xy, int stride,
                    int n, int part, int nparts, unsigned long long *counts);

// Radius classes of the lean sweep: class 0 = rad0 (the first particle's radius),
// class 1 = one other value.  Called for every particle whose radius enters the
// resident state (upload, halo).  Plain reads first: after the first claim lands
// nobody issues atomics any more.
#ifdef __CUDACC__
// A radius CLASS is a value and everything within kRadClassTol of it (relative): the reference's
// growth phase leaves every disk at vr * t with its own rounding (src/EDMD.c:4992-5007, stopGrow
// :4740 does not reset it), so its "monodisperse" and "bidisperse" runs hold ~10 distinct radii per
// species within 1e-15 of each other.  Level 0 = every radius EXACTLY rad0 (the exact stage may use
// the constant); level 1 = at most two classes (rad0's, and kFlagRad1's if one was seen): the
// screening takes the class radius inflated by the tolerance, the exact stage reads each disk's own
// FP64 radius.
constexpr double kRadClassTol = 1e-9;
__device__ __forceinline__ bool edmd_same_class(double r, double rc) { return fabs(r - rc) <= kRadClassTol * rc; }
__device__ __forceinline__ void edmd_note_radius(int32_t *flags, double r, double rad0)
{
    if (r == rad0) return;
    unsigned long long *slot = reinterpret_cast<unsigned long long *>(flags + kFlagRad1);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(r);
    int level = 2;
    if (r > 0) {
        if (edmd_same_class(r, rad0)) {
            level = 1;
        } else {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(slot);
            if (cur == 0ull) cur = atomicCAS(slot, 0ull, bits);
            if (cur == 0ull || cur == bits || edmd_same_class(r, __longlong_as_double((long long)cur))) level = 1;
        }
    }
    if ((*reinterpret_cast<volatile int32_t *>(flags + kFlagNotMono) & level) != level)
        atomicOr(&flags[kFlagNotMono], level);
}
#endif

// Function attributes (dynamic shared memory size, carveout) are PER DEVICE: a launcher sets them the first
// time it runs on each device of the process (one process may drive several: multi_gpu.cu).
// `done` = the call site's own static mask, one bit per device ordinal.
inline bool edmd_first_on_device(unsigned long long *done)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
    const unsigned long long bit = 1ull << dev;
    if (*done & bit) return false;
    *done |= bit;
    return true;
}

// ---- programmatic dependent launch -----------------------------------------------
// The kernels of a sweep form a dependent chain on one stream.  Each is launched
// with the programmatic-stream-serialization attribute and begins with
// griddepcontrol.wait (nothing reads predecessor data before it): the launch and
// scheduling latency of kernel k+1 then overlaps the tail of kernel k.  Measured at
// N = 10^6 (lean sweep): 110 -> 103 us per step.  Triggering the dependents EARLY
// (griddepcontrol.launch_dependents at kernel start) was slower, 117 us: the waiting
// CTAs of the next kernel take SM resources from the running one.
__device__ __forceinline__ void edmd_pdl_wait()
{
#if __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

__device__ __forceinline__ void edmd_stamp(unsigned long long *ts, int k)
{
    if (!ts) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    ts[k] = t;
}

// lets the NEXT kernel of the chain (launched with the programmatic attribute) start now
__device__ __forceinline__ void edmd_pdl_trigger()
{
#if __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
inline cudaError_t edmd_launch(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                               bool pdl, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, static_cast<KArgs>(args)...);
}
