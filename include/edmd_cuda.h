/*
 * edmd_cuda.h -- C ABI of the B200 (sm_100a) implementation of
 * Graphical-EDMD's data-parallel hot path.
 *
 * The reference has no plugin/FFI layer.  Its seams for this path are the
 * batch loops that call the per-particle function pointers
 *     crossingEvent(int) / collisionEvent(int) / freeFly(particle*)
 *                                   (src/EDMD.c:317-321)
 * for ALL particles:
 *     eventListInit   src/EDMD.c:2007-2012   (setup)
 *     stopGrow        src/EDMD.c:4772-4779   (end of growth)
 *     addNoise        src/EDMD.c:4909-4915   (every thermostat tick)
 *     doUmbrella      src/EDMD.c:4519-4536   (umbrella roll-back)
 *     takeAScreenshot src/EDMD.c:4659-4661   (free-fly re-sync)
 * and the analysis signatures
 *     calculate_pcf(particle*, N, dr, max_r, Lx, Ly)          src/pcf.h:33
 *     computeBOOPCutoff(particle*, N, r_c, cellList, Nxcells) src/boop.h:13
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions: plain C types only; every function returns 0 on success, a
 * negative value for a CUDA/runtime failure and a positive EDMD_E* value for a
 * semantic error; edmd_cuda_last_error() gives the text.  A context is driven
 * by one host thread (the reference is single-threaded).  Host buffers are
 * caller-owned and may be pageable.  Nothing here ever calls exit().
 * There is no CPU fallback: without a CUDA device create() fails.
 */
#ifndef EDMD_CUDA_H
#define EDMD_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct edmd_ctx edmd_ctx;

/* semantic error codes (positive) */
#define EDMD_EINVAL 1    /* bad argument */
#define EDMD_ESTATE 2    /* call out of order (e.g. predict before upload) */
#define EDMD_EOVERLAP 3  /* predict found c < -0.01: reference would exit(3) */
#define EDMD_ECELL 4     /* a cell id outside the grid */
#define EDMD_ENOMEM 5
#define EDMD_EPLAN 6     /* calendar_plan declined (a bucket too full): ingest event by event */
#define EDMD_EVORONOI 7  /* a Voronoi cell could not be built locally (tiny or very inhomogeneous system) */

/* prediction mode: which set of reference function pointers is emulated
 * (src/EDMD.c:1977-1989) */
#define EDMD_MODE_NORMAL 0 /* crossingEventNormal + collisionEventNormal */
#define EDMD_MODE_GROW 1   /* crossingEventGrow   + collisionEventGrow   */

/* enum event values the sweep emits (order of src/EDMD.h:19-34) */
#define EDMD_EV_CELLCROSS 0
#define EDMD_EV_COLLISION 1

#define EDMD_NEVER 100000000000000000000000000.0 /* src/EDMD.h:8 */

/* Box constants exactly as boxConstantHelper derives them, src/EDMD.c:679-715 */
typedef struct {
	int32_t n;
	int32_t nxcells, nycells;
	double lx, ly, half_lx, half_ly;
	double cellx_size, celly_size;
	double cellx_fac, celly_fac;
	double dt_paul; /* PAUL_DT_SCALE / N, src/EDMD.c:712 */
} edmd_box;

/* ---- lifetime -------------------------------------------------------- */

/* Replaces: constantInit's box set-up + boxConstantHelper (src/EDMD.c:679-715,
 * 1150-1217), particles calloc (:1358) and cellList calloc (:1909).
 * Allocates the device SoA store, the cell index and pinned staging. */
int edmd_cuda_create(int device, int n, double lx, double ly, edmd_ctx **out);
void edmd_cuda_destroy(edmd_ctx *ctx);
const char *edmd_cuda_last_error(const edmd_ctx *ctx);
int edmd_cuda_get_box(const edmd_ctx *ctx, edmd_box *out);
/* number of kernels this context has launched so far (bench accounting) */
uint64_t edmd_cuda_launch_count(const edmd_ctx *ctx);

/* Options.  EDMD_OPT_FORCE_GENERIC = 1 runs the sweep with the plain
 * global-memory kernel (the reference's loops as written) instead of the tiled
 * two-phase kernel; results are identical, it exists for cross-checking. */
#define EDMD_OPT_FORCE_GENERIC 1
/* EDMD_OPT_NO_LEAN = 1 keeps monodisperse NORMAL-mode sweeps on the full FP64
 * row kernel instead of the lean path (FP32 screening certified against the
 * exact FP64 evaluation, graphical-edmd_b200/csrc/lean.cuh); results are
 * identical, it exists for cross-checking and timing. */
#define EDMD_OPT_NO_LEAN 2
/* EDMD_OPT_NO_PDL = 1 launches the lean sweep's kernel chain without programmatic
 * dependent launch (plain stream order); for timing comparisons. */
#define EDMD_OPT_NO_PDL 3
/* EDMD_OPT_PCF_LEGACY selects the g(r) kernel: 0 (default) = spatially sorted tiles, the
 * bin of a pair decided in FP32 under a rigorous error bound and the pairs FP32 cannot
 * settle redone with the reference's FP64 operations; 1 = the plain tile kernel (IEEE
 * sqrt and division per pair, id-ordered tiles); 2 = sorted tiles with every bin
 * certified in FP64.  Same integer counts; for cross-checking and timing. */
#define EDMD_OPT_PCF_LEGACY 4
/* EDMD_OPT_PCF_GROUPS: the default g(r) kernel runs CTAs of 1, 2 or 4 groups of 256 threads
 * that share one shared-memory histogram; 0 (default) lets the library pick the count with the
 * most resident warps, 1 / 2 / 4 force it (cross-checking and timing). */
#define EDMD_OPT_PCF_GROUPS 5
/* EDMD_OPT_NO_TILE = 1 keeps eligible sweeps off the two-kernel tile sweep (tile_sweep.cu)
 * and on the older five-kernel lean chain (cross-checking and timing). */
#define EDMD_OPT_NO_TILE 6
/* EDMD_OPT_NO_TILE_BOOP = 1 keeps psi6 (edmd_cuda_boop_cutoff) on the row kernel over the full cell index. */
#define EDMD_OPT_NO_TILE_BOOP 7
/* One particle's two scheduled events as the tile sweep leaves them on the device: exactly what
 * addCrossingEvent / addCollisionEvent store into eventList[i] / eventList[N+i]
 * (src/EDMD.c:2487-2496, 3286-3297), one 32-byte record per particle. */
typedef struct edmd_ev32 {
    double t_cross, t_coll;   /* absolute event times */
    int32_t partner;          /* collision partner (caller's particle id) */
    uint8_t dir;              /* crossing direction 1..4 */
    uint8_t ctype;            /* EDMD_EV_COLLISION */
    uint8_t pad[10];
} edmd_ev32;
int edmd_cuda_set_option(edmd_ctx *ctx, int option, int value);
/* Counters: EDMD_STAT_EXACT_RESCANS = particles the tiled sweep had to resolve
 * with the exact re-scan (near-ties, ill-conditioned pairs) since create. */
#define EDMD_STAT_EXACT_RESCANS 1
/* EDMD_STAT_LEAN_SWEEPS = sweeps that completed on the lean path (counted when
 * their results are fetched or planned from; declined ones are not). */
#define EDMD_STAT_LEAN_SWEEPS 2
/* g(r), sorted-tile kernel, since create: pairs whose bin needed the exact sqrt + division;
 * tile pairs skipped because every pair in them is beyond max_r. */
/* EDMD_STAT_LEAN_DECLINES = lean sweeps the device declined (re-run on the full path);
 * EDMD_STAT_LEAN_ELIGIBLE = 1 when the next NORMAL-mode sweep would try the lean path. */
#define EDMD_STAT_LEAN_DECLINES 5
#define EDMD_STAT_LEAN_ELIGIBLE 6
#define EDMD_STAT_PCF_EXACT_PAIRS 3
#define EDMD_STAT_PCF_SKIPPED_TILE_PAIRS 4
int edmd_cuda_get_stat(edmd_ctx *ctx, int stat, uint64_t *value);

/* Self-test of an assumption the default g(r) kernel's error bound rests on: runs the
 * hardware's approximate reciprocal square root over every float in [2^-100, 2^64) and
 * returns the largest relative error found (the kernel budgets 1.28e-7; PTX documents
 * 2^-22.9).  No reference counterpart. */
int edmd_cuda_selftest_rsqrt(edmd_ctx *ctx, double *max_rel_err);

/* Page-locked host memory for the caller's particle arrays: uploads and
 * downloads from/to such memory go straight to the copy engines (pageable
 * memory is bounced through an internal pinned buffer instead). */
int edmd_cuda_host_alloc(void **ptr, size_t bytes);
void edmd_cuda_host_free(void *ptr);

/* ---- state upload ---------------------------------------------------- */

/* Synchronous snapshot: every particle already advanced to time t (the state
 * the reference has after its freeFly loop, src/EDMD.c:4838-4886, :4744).
 * cell_xy (interleaved X,Y = particle.cell[0..1], src/EDMD.h:57) may be NULL:
 * the device then computes (int)(x*cellxFac) like coordToCell (:2098-2107).
 * Mid-run callers MUST pass the host's cell ids.
 * rad may be NULL on any upload after the first one of the same particles:
 * the resident radii are kept (NORMAL mode never changes them), which saves 8
 * of the 40 bytes per particle that cross PCIe at a thermostat tick. */
int edmd_cuda_upload(edmd_ctx *ctx, const double *x, const double *y,
                     const double *vx, const double *vy, const double *rad,
                     const int32_t *cell_xy, double t);

/* Same, reading straight out of an array of records such as the reference's
 * particles[] (struct particle, src/EDMD.h:42-64, 144-byte stride): byte
 * offsets of the double fields x,y,vx,vy,rad and of int cell[2]
 * (off_cell == (size_t)-1 => compute cells on the device). */
int edmd_cuda_upload_aos(edmd_ctx *ctx, const void *base, size_t stride,
                         size_t off_x, size_t off_y, size_t off_vx,
                         size_t off_vy, size_t off_rad, size_t off_cell,
                         double t);

/* ---- the prediction sweep ------------------------------------------- */

/* Replaces the loop bodies `crossingEvent(i); collisionEvent(i);` over all i
 * (src/EDMD.c:2007-2012, 4772-4779, 4909-4915, 4530-4536).  Outputs are indexed
 * by ORIGINAL particle id and carry what addCrossingEvent / addCollisionEvent
 * (src/EDMD.c:2487-2496, 3286-3297) store into eventList[i] / eventList[N+i]:
 *   t_cross[i]  absolute crossing time  (node.t)
 *   dir[i]      1..4                    (node.j)
 *   t_coll[i]   absolute collision time, t + 1e26 when no candidate
 *   partner[i]  partner id, 0 when no candidate (the reference's default)
 *   ctype[i]    EDMD_EV_COLLISION
 * vr: growth rates, required for EDMD_MODE_GROW, ignored otherwise.
 * overlap_pair[2] (nullable): first (i,j) in sweep order with c < -0.01, else
 * -1,-1; the call then returns EDMD_EOVERLAP with all outputs still filled
 * (the host mirrors the reference's exit(3), src/EDMD.c:2708-2717).
 * Blocks until the results are in host memory. */
int edmd_cuda_predict_all(edmd_ctx *ctx, int mode, const double *vr,
                          double *t_cross, uint8_t *dir, double *t_coll,
                          int32_t *partner, uint8_t *ctype,
                          int32_t *overlap_pair);

/* The two halves of predict_all for callers that keep state resident:
 * run K0+K1 on the uploaded state (asynchronous, results stay in HBM) ... */
int edmd_cuda_predict_device(edmd_ctx *ctx, int mode);
/* ... and copy the last device results out (blocking). */
int edmd_cuda_fetch_predictions(edmd_ctx *ctx, double *t_cross, uint8_t *dir,
                                double *t_coll, int32_t *partner,
                                uint8_t *ctype, int32_t *overlap_pair);
int edmd_cuda_set_growth(edmd_ctx *ctx, const double *vr);

/* ---- multi-GPU: row slabs of the cell grid ---------------------------- */

/* One context per GPU/rank owning the cell rows [row_lo, row_hi) of the GLOBAL
 * grid (row-major cell index Y*Nxcells + X, src/EDMD.c:2071, makes a slab one
 * contiguous cell range; periodic in y like PBCcellY :2118-2124).  It holds its
 * owned particles plus copies of the neighbouring slabs' boundary rows (a
 * one-cell-row halo) and predicts only the owned ones.  n_capacity bounds
 * owned + halo particles.  The box is the global one. */
int edmd_cuda_create_slab(int device, int n_capacity, double lx, double ly, int row_lo, int row_hi,
                          edmd_ctx **out);
/* Owned particles of this slab (all in rows [row_lo,row_hi)); global_id[i] is
 * the id the caller knows the particle by: partner[] outputs carry global ids,
 * outputs are indexed by the LOCAL index 0..n_owned-1.  rad == NULL / global_id == NULL
 * keep the resident radii / ids (a thermostat tick: the same n_owned particles with new
 * positions and velocities). */
int edmd_cuda_upload_owned(edmd_ctx *ctx, int n_owned, const double *x, const double *y,
                           const double *vx, const double *vy, const double *rad,
                           const int32_t *cell_xy, const int32_t *global_id, double t);
/* Halo exchange, device buffers of 48-byte records owned by the caller (e.g.
 * NCCL send/recv buffers): pack the first (side 0, for the lower neighbour) or
 * last (side 1, for the upper neighbour) owned row; append what the lower
 * (side 0) / upper (side 1) neighbour sent.  The transport between ranks is
 * the caller's (ncclSend/ncclRecv, or P2P copies). */
#define EDMD_HALO_RECORD_BYTES 48
int edmd_cuda_halo_pack(edmd_ctx *ctx, int side, void *dev_records, int capacity, int *count);
int edmd_cuda_halo_append(edmd_ctx *ctx, int side, const void *dev_records, int count);
/* The same exchange done by the GPUs themselves over NVLink peer memory (no
 * NCCL call, no host synchronisation; see csrc/halo.cu): every rank exports
 * its inbox (a 64-byte CUDA IPC handle), the caller ships the handles to the
 * neighbours by any means, connect() maps them (NULL = the neighbour is this
 * context itself), and halo_exchange() enqueues pack-and-peer-store + unpack
 * kernels on the context's stream.  halo_capacity = records per boundary row
 * buffer; n_capacity must hold n_owned + 2 * halo_capacity. */
int edmd_cuda_halo_export(edmd_ctx *ctx, int halo_capacity, void *handle64);
int edmd_cuda_halo_connect(edmd_ctx *ctx, const void *lower_handle64, const void *upper_handle64);
int edmd_cuda_halo_exchange(edmd_ctx *ctx);
/* Halo exchange + sweep of a slab context as ONE stream-ordered sequence that hides the transfer:
 * send (peer stores over NVLink) -> partition of the owned particles -> receive + partition of the
 * neighbours' boundary rows -> sweep kernel.  Asynchronous like edmd_cuda_predict_device; every rank
 * of the decomposition must call it the same number of times (the exchange counts epochs).  Replaces
 * edmd_cuda_halo_exchange + edmd_cuda_predict_device; results through edmd_cuda_fetch_predictions. */
int edmd_cuda_exchange_predict_device(edmd_ctx *ctx, int mode);
int edmd_cuda_get_counts(const edmd_ctx *ctx, int *n_owned, int *n_total);
/* g(r) share of one rank: positions of ALL particles as (x,y) pairs in device
 * memory (e.g. after an all-gather), tile pairs part (mod nparts); ADDS into
 * counts_dev[num_bins] (device, caller zeroes it and all-reduces it). */
int edmd_cuda_pcf_device(edmd_ctx *ctx, const double *xy_dev, int n_total, double dr, double max_r,
                         int part, int nparts, uint64_t *counts_dev, int *num_bins);

/* ---- several GPUs in ONE process (csrc/multi_gpu.cu) -------------------------------------
 * The single-GPU interface over `ndev` devices: an edmd_mg owns one slab context per device
 * (row slabs as above, halo by peer stores over NVLink -- peer access between the devices of the
 * process, no IPC) and takes / returns WHOLE-SYSTEM host arrays indexed by particle id.  The same
 * device may be listed more than once (several slabs on one GPU: how the tests run without a second
 * GPU).  Needs Nycells >= 3 * ndev; ndev == 1 is an ordinary whole-system context behind the same calls.
 * One host thread per edmd_mg, like an edmd_ctx. */
typedef struct edmd_mg edmd_mg;
int edmd_cuda_create_mg(int ndev, const int *devices, int n, double lx, double ly, edmd_mg **out);
void edmd_cuda_destroy_mg(edmd_mg *mg);
const char *edmd_cuda_mg_last_error(const edmd_mg *mg);
int edmd_cuda_mg_get_box(const edmd_mg *mg, edmd_box *box);
/* particles currently owned by every slab, n_owned[ndev] */
int edmd_cuda_mg_slab_sizes(const edmd_mg *mg, int *n_owned);
/* As edmd_cuda_upload (cell_xy nullable): the particles are dealt to the slabs by the cell row they
 * are filed under -- particles migrate between slabs during a run, so every upload carries radii and ids. */
int edmd_cuda_mg_upload(edmd_mg *mg, const double *x, const double *y, const double *vx, const double *vy,
                        const double *rad, const int32_t *cell_xy, double t);
/* As edmd_cuda_predict_all, mode EDMD_MODE_NORMAL: every slab runs the fused halo exchange + sweep, all
 * devices side by side; the slabs' outputs are disjoint (no collective) and land in the caller's arrays. */
int edmd_cuda_mg_predict_all(edmd_mg *mg, int mode, double *t_cross, uint8_t *dir, double *t_coll,
                             int32_t *partner, uint8_t *ctype, int32_t *overlap_pair);
/* As edmd_cuda_boop_cutoff; mean_q6 = (sum of the slabs' q6 sums) / N, the thermo column (src/EDMD.c:5521-5536). */
int edmd_cuda_mg_boop_cutoff(edmd_mg *mg, double r_c, double *q5, double *q6, double *q7, double *q6_arg,
                             int32_t *neighbors, double *mean_q6);
/* As edmd_cuda_pcf on the positions given (all pairs interact: slabs do not help): every device gets the
 * positions and bins the tile pairs w = k (mod ndev); the integer histograms are added. */
int edmd_cuda_mg_pcf(edmd_mg *mg, const double *x, const double *y, double dr, double max_r, uint64_t *counts,
                     double *g_r, int *num_bins);

/* ---- weighted g(r) family -------------------------------------------------- */

/* Replaces calculate_bond_order_pcf (src/pcf.c:77-167; caller save_pcf_boop
 * :338-403 with dr = 2, max_r = min(Lx,Ly)/2): g(r) and the average of
 * cos(k_vector . d) over the pairs of each bin, on the resident positions.
 * counts[b] = unordered pairs in bin b (the reference's g_r[b] before
 * normalisation is 2 counts[b]); g_r is normalised as the reference does; g6_r
 * = per-bin average (0 for empty bins).  Any output may be NULL. */
int edmd_cuda_pcf_bond_order(edmd_ctx *ctx, double dr, double max_r, const double *k_vector,
                             uint64_t *counts, double *g_r, double *g6_r, int *num_bins);
/* Replaces find_max_structure_factor_bragg (src/pcf.c:405-467): the wave vector
 * k = (2 pi i / Lx, 2 pi j / Ly) with the largest S(k) = |sum_j e^{i k.r_j}|^2 / N
 * among those with |k| >= 1.5 inside the wedge |arg k - pi/2| <= pi/5 and the
 * square |k_x|, |k_y| <= expected_bragg + 0.8; the first maximum in the
 * reference's loop order wins.  s_max (optional) = that S(k). */
int edmd_cuda_bragg_peak(edmd_ctx *ctx, double expected_bragg, double *k_out, double *s_max);

/* Replaces compute_g6_correlation (src/pcf.c:169-230; caller save_g6_correlation
 * :297-336): per-bin average of Re(conj(psi6_i) psi6_j) over the unordered pairs
 * with r < max_r (g6_corr, 0 for empty bins) and the pair counts.  psi_re / psi_im
 * [N] = the bond-orientational field to correlate; both NULL = the reference's
 * choice, psi6_i = q6_i e^{i q6_arg_i} of the VORONOI neighbours (:182-186),
 * computed on the device by the kernels behind edmd_cuda_boop_voronoi. */
int edmd_cuda_g6_correlation(edmd_ctx *ctx, double dr, double max_r, const double *psi_re,
                             const double *psi_im, uint64_t *counts, double *g6_corr, int *num_bins);
/* Replaces initStructureFactor's wave-vector grid (src/struc.c:328-345; qx[i] =
 * (2 pi/Lx)(i - (nqx-1)/2), nqx = (int)(2 q_max/(2 pi/Lx) + 1), same in y) and
 * computeStructureFactor (velocity = 0) / computeVelocityStructureFactor
 * (velocity = 1) (src/struc.c:364-408; caller saveStructureFactor :409-425) on the
 * resident state: s[i*nqy + j] = |sum_n w_n e^{i q.r_n}|^2 / N; re / im = the sums
 * themselves (the reference's structFactorComplex, doFQT).  With s, re and im all
 * NULL only the grid is returned (qx / qy may be NULL too: sizes only). */
int edmd_cuda_structure_factor(edmd_ctx *ctx, double q_max, int velocity, int *nqx, int *nqy,
                               double *qx, double *qy, double *s, double *re, double *im);

/* ---- calendar ingest plan ----------------------------------------------- */

/* For the 2N events of the last sweep (event e = i: crossing of particle i at
 * t_cross[i]; e = N + i: its collision at t_coll[i] -- the indices of the
 * reference's eventList[], src/EDMD.c:1923-1934) compute what 2N calls of
 * addEventToQueue (src/EDMD.c:2144-2170) in the batch loops' order (crossing
 * 0, collision 0, crossing 1, ...; :2007-2012, :4909-4915) would do to an EMPTY
 * calendar with the given geometry (paulTime, dtPaul, paulListN,
 * actualPaulList):
 *   bucket[e]  -1: the event belongs in the BST (dt < dtPaul), the host adds it
 *              itself; else the Paul list it lands in (paul_n = overflow list)
 *   next[e] / prev[e]   its ->rgt / ->lft neighbour in that list (-1 = NULL)
 *   head[k]    eventPaul[k] for k = 0 .. paul_n  (-1 = NULL)
 *   n_tree     number of BST events
 * so the host can fill its nodes in one streaming pass instead of 2N divisions
 * and pointer-chasing insertions.  Same FP64 operations as the reference, hence
 * the same buckets bit for bit.  Arrays are caller-owned: bucket/next/prev
 * [2N], head [paul_n + 1].  Returns EDMD_EPLAN (outputs unusable) if some list
 * would hold more than 128 of the events; the caller then inserts one by one. */
int edmd_cuda_calendar_plan(edmd_ctx *ctx, double paul_time, double dt_paul, int paul_n,
                            int actual_paul, int32_t *bucket, int32_t *next, int32_t *prev,
                            int32_t *head, int32_t *n_tree);

/* ---- free flight ------------------------------------------------------ */

/* Replaces `for i<N freeFly(particles+i)` (takeAScreenshot, src/EDMD.c:
 * 4659-4661; freeFlyNormal :4954-4990 / freeFlyGrow :4992-5007) on the
 * resident state: x += dt*vx, y += dt*vy (rad += dt*vr in GROW mode), one
 * +-L wrap (PBCpostX/Y :5948-5959), t <- t_new.  Cell ids are NOT changed
 * (in the reference only crossing events change them). */
int edmd_cuda_free_fly(edmd_ctx *ctx, int mode, double t_new);
int edmd_cuda_download_state(edmd_ctx *ctx, double *x, double *y, double *vx,
                             double *vy, double *rad);

/* ---- thermostat tick on the resident state -------------------------------- */

/* Replaces physicalQ's sums (src/EDMD.c:5968-5997) over the resident velocities,
 * unit masses (the reference's default initial conditions, :1467-1548):
 * E = sum 1/2 (vx^2 + vy^2), px = sum vx, py = sum vy.  Parallel, reproducible
 * (fixed summation tree); agrees with the reference's sequential sum to ~1e-15
 * relative, not bit for bit. */
int edmd_cuda_kinetic(edmd_ctx *ctx, double *E, double *px, double *py);
/* Replaces the velocity-rescale branch of addNoise (src/EDMD.c:4830-4832 +
 * :4899-4902): `physicalQ(); ... p->vx /= sqrt(E/N/T); p->vy /= sqrt(E/N/T);` on
 * the resident velocities.  E_before / divisor (nullable) = E and sqrt(E/N/T).
 * A whole tick without the state crossing PCIe: edmd_cuda_free_fly(t) (the
 * freeFly(p) of the same loop, :4893), this call, edmd_cuda_predict_device()
 * (the re-predict loop :4909-4915). */
int edmd_cuda_rescale_velocities(edmd_ctx *ctx, double T, double *E_before, double *divisor);
/* Replaces the Langevin branch of addNoise (noise == 1: randomGaussian, src/EDMD.c:5802-5826,
 * the non-Euler form without damping, unit masses) on the resident velocities:
 *     c = exp(-gamma dtnoise);  v <- sqrt(T (1 - c^2)) a (cos b, sin b) + c v,
 * (a, b) the Box-Muller pair of two uniforms per particle.  The reference draws them from its
 * sequential MT19937 stream (genrand_real3), which a parallel kernel cannot follow; here they come
 * from a counter-based generator keyed on (seed, tick, particle id) -- the one of
 * graphical-edmd_b200/synth.py, slab contexts key on the global id -- so a kick is reproducible
 * and independent of the decomposition.  Parity with the reference is therefore STATISTICAL
 * (<v^2> -> c^2 <v^2> + (1 - c^2) T per component); against a numpy restatement of the same
 * generator and formula it is exact to rounding (tests).  Pass a new `tick` at every call. */
int edmd_cuda_langevin_kick(edmd_ctx *ctx, double T, double gamma, double dtnoise, uint32_t seed,
                            uint32_t tick);

/* Replaces normalizePhysicalQ (src/EDMD.c:5723-5764, the branch without a circular wall, unit masses;
 * callers: the initial conditions :1556 and stopGrow :4745) on the RESIDENT state:
 *   physicalQ();  v -= p/(N*m);  physicalQ();  v /= sqrt(E/N/Einit)
 * -- the centre-of-mass velocity removed, the kinetic energy per particle set to Einit (the reference's
 * option of that name, default 1).  The parallel sums agree with the reference's sequential ones to
 * ~1e-15 relative (inside the 1e-12 bar; reproducible run to run).  Outputs (nullable): the momentum
 * before, the energy of the shifted velocities, the divisor sqrt(E/N/Einit).  Whole-system contexts. */
int edmd_cuda_normalize_velocities(edmd_ctx *ctx, double Einit, double *px_before, double *py_before,
                                   double *E_shifted, double *divisor);

/* v <- (v - (dvx, dvy)) / divisor on the particles the context owns: the two loops of normalizePhysicalQ
 * / the rescale of addNoise (:4899-4902) with sums the CALLER supplies.  For slab contexts: all-reduce
 * edmd_cuda_kinetic's (E, px, py) over the ranks, then call this with the whole system's values. */
int edmd_cuda_shift_scale_velocities(edmd_ctx *ctx, double dvx, double dvy, double divisor);

/* ---- per-frame structure analysis ----------------------------------- */

/* Replaces calculate_pcf (src/pcf.c:16-75; caller save_pcf, src/EDMD.c:
 * 5636-5637 with dr = 0.1, max_r = min(Lx,Ly)/2).  counts[b] = number of
 * UNORDERED pairs in bin b (the reference adds 2.0 per pair); g_r (nullable)
 * gets the reference's normalisation (:56-72).  *num_bins = (int)(max_r/dr);
 * counts/g_r must hold that many entries (query with counts == NULL). */
int edmd_cuda_pcf(edmd_ctx *ctx, double dr, double max_r, uint64_t *counts,
                  double *g_r, int *num_bins);

/* Replaces computeBOOPCutoff (src/boop.c:61-107; callers saveTXT src/EDMD.c:
 * 5057-5060 and saveThermo :5521-5536 with r_c = 2.5).  SoA mirror of
 * boop_data (src/boop.h:6-10).  Like the reference it scans only the 3x3 cell
 * block.  mean_q6 (nullable) = sum(q6)/N as saveThermo prints it. */
int edmd_cuda_boop_cutoff(edmd_ctx *ctx, double r_c, double *q5, double *q6,
                          double *q7, double *q6_arg, int32_t *neighbors,
                          double *mean_q6);

/* Replaces computeBOOPVoronoi (src/boop.c:15-59; callers saveTXT src/EDMD.c:5053-5056
 * and saveThermo :5523-5525 with boopThermo == 1): psi_5,6,7 over the VORONOI
 * neighbours of every particle in the periodic box.  The reference runs Fortune's
 * sweep (jc_voronoi.h) over the particles plus the periodic images within 6.0 of
 * the edges (get_particle_voronoi, src/voronoi_edmd.c:33-121); here every particle
 * builds its own cell by clipping against the bisectors of its surroundings --
 * the same diagram wherever it is unique.  Returns EDMD_EVORONOI when some cell
 * cannot be closed locally (fewer than ~3 x 3 grid cells of particles, huge voids). */
int edmd_cuda_boop_voronoi(edmd_ctx *ctx, double *q5, double *q6, double *q7, double *q6_arg,
                           int32_t *neighbors, double *mean_q6);
/* Replaces get_particle_voronoi_area / get_particle_voronoi_perimeter
 * (src/voronoi_edmd.c:123-149; caller saveTXT src/EDMD.c:5061-5064 for the local
 * packing fraction): area and perimeter of every particle's Voronoi cell, and
 * its number of edges.  Any output may be NULL. */
int edmd_cuda_voronoi_cells(edmd_ctx *ctx, double *area, double *perimeter, int32_t *neighbors);

#ifdef __cplusplus
}
#endif
#endif /* EDMD_CUDA_H */
