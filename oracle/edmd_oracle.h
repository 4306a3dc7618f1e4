/*
 * oracle/edmd_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, strict IEEE, no FMA contraction) of the reference's
 * hot path: whole-system prediction sweep + g(r) + psi6.  It exists to CHECK
 * the CUDA library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product
 * (libedmd_cuda.so, edmd_host) never links or calls it.
 *
 * Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
 * section 4), so this restatement is pinned against the reference ITSELF,
 * compiled unmodified into oracle/_ref/libedmd_ref.so (see ref_shim.c), by
 * tests/test_oracle_vs_ref.py (run wherever _ref exists) and by the committed
 * fixtures in tests/golden/ that tests/golden/make_golden.py generated from
 * _ref in the build container.
 */
#ifndef EDMD_ORACLE_H
#define EDMD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int n;            /* number of particles (enters dtPaul only) */
	int nx, ny;       /* Nxcells, Nycells     src/EDMD.c:694-695 */
	double lx, ly;    /* box */
	double half_lx, half_ly;
	double csx, csy;  /* cellxSize, cellySize src/EDMD.c:701-702 */
	double fx, fy;    /* cellxFac, cellyFac   src/EDMD.c:705-706 */
} oracle_box;

#define ORACLE_MODE_NORMAL 0
#define ORACLE_MODE_GROW 1

/* enum event order of src/EDMD.h:19-34 */
#define ORACLE_EV_CELLCROSS 0
#define ORACLE_EV_COLLISION 1

void oracle_box_init(oracle_box *b, int n, double lx, double ly);

/* coordToCell for every particle (src/EDMD.c:2098-2107); cell_xy interleaved X,Y */
void oracle_cells_from_coords(const oracle_box *b, int n, const double *x,
                              const double *y, int32_t *cell_xy);

/* Full sweep (src/EDMD.c:2007-2012 / 4909-4915) on a synchronous snapshot.
 * cell_xy NULL => coordToCell.  vr only read in GROW mode.
 * overlap_pair[2] = first (i, j) in sweep order with c < -0.01, else -1,-1.
 * Returns 0, or 1 when an overlap was flagged. */
int oracle_predict_all(const oracle_box *b, int n, double t, const double *x,
                       const double *y, const double *vx, const double *vy,
                       const double *rad, const double *vr,
                       const int32_t *cell_xy, int mode, double *t_cross,
                       uint8_t *dir, double *t_coll, int32_t *partner,
                       uint8_t *ctype, int32_t *overlap_pair);

/* freeFlyNormal / freeFlyGrow over all particles (src/EDMD.c:4954-5007);
 * tp = per-particle local times (NULL => all at t_old). In place. */
void oracle_free_fly(const oracle_box *b, int n, int mode, double t_old,
                     const double *tp, double t_new, double *x, double *y,
                     const double *vx, const double *vy, double *rad,
                     const double *vr);

/* calculate_pcf (src/pcf.c:16-75).  counts[bin] = unordered pairs (reference
 * adds 2.0 per pair).  g_r, r may be NULL.  Uses all OpenMP threads (integer
 * counts, order-independent).  Returns num_bins. */
int oracle_pcf(const oracle_box *b, int n, const double *x, const double *y,
               double dr, double max_r, uint64_t *counts, double *g_r,
               double *r);
int oracle_pcf_num_bins(double dr, double max_r);

/* computeBOOPCutoff (src/boop.c:61-107). cell_xy NULL => coordToCell. */
void oracle_boop_cutoff(const oracle_box *b, int n, const double *x,
                        const double *y, const int32_t *cell_xy, double r_c,
                        double *q5, double *q6, double *q7, double *q6_arg,
                        int32_t *neighbors);

#ifdef __cplusplus
}
#endif
/* weighted g(r) family (src/pcf.c:77-167, 405-467) */
int oracle_bond_order_pcf(const oracle_box *b, int n, const double *x, const double *y,
                          double dr, double max_r, double kx, double ky, uint64_t *counts,
                          double *g_r, double *g6_r);
void oracle_bragg_peak(int n, const double *x, const double *y, double lx, double ly,
                       double expected_bragg, double *k_out, double *s_out);

/* pair loop of compute_g6_correlation (src/pcf.c:189-228): per-bin average of
 * Re(conj(psi6_i) psi6_j) over the unordered pairs with r < max_r; psi6 given
 * (the reference takes it from computeBOOPVoronoi, :182-186).  counts as integers.
 * Returns num_bins. */
int oracle_g6_correlation(const oracle_box *b, int n, const double *x, const double *y,
                          const double *psi_re, const double *psi_im, double dr, double max_r,
                          uint64_t *counts, double *g6_corr);
/* initStructureFactor's wave-vector grid (src/struc.c:328-345); qx / qy may be
 * NULL to query the sizes. */
void oracle_sq_grid(double q_max, double lx, double ly, int *nqx, int *nqy, double *qx, double *qy);
/* computeStructureFactor / computeVelocityStructureFactor (src/struc.c:364-408):
 * s[i*nqy + j] = |sum_n w_n e^{i q.r_n}|^2 / N. */
void oracle_structure_factor(int n, const double *x, const double *y, const double *vx,
                             const double *vy, int nqx, const double *qx, int nqy,
                             const double *qy, int velocity, double *s);

#endif
