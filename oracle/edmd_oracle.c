/*
 * oracle/edmd_oracle.c -- TEST INFRASTRUCTURE ONLY (see edmd_oracle.h).
 *
 * Plain-C restatement of the reference's hot path, written from the
 * reference's algorithm (each function cites the reference file:line it
 * follows), compiled strict-IEEE with -ffp-contract=off so no multiply-add is
 * fused: the reference's default x86-64 build has no FMA either (SURVEY.md
 * 7.2 #1).  "Parity pinned" against oracle/_ref (the unmodified reference) by
 * tests/test_oracle_vs_ref.py and the fixtures under tests/golden/.
 */
#include "edmd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NEVER 100000000000000000000000000.0 /* src/EDMD.h:8 */

/* boxConstantHelper, src/EDMD.c:679-715 (addWell == 0 branch) */
void oracle_box_init(oracle_box *b, int n, double lx, double ly)
{
	b->n = n;
	b->lx = lx;
	b->ly = ly;
	b->half_lx = lx / 2;
	b->half_ly = ly / 2;
	b->ny = (int)(b->half_ly);
	b->nx = (int)(b->half_lx);
	b->csx = lx / b->nx;
	b->csy = ly / b->ny;
	b->fx = 1 / b->csx;
	b->fy = 1 / b->csy;
}

/* coordToCell, src/EDMD.c:2098-2107: multiply by the reciprocal, truncate */
void oracle_cells_from_coords(const oracle_box *b, int n, const double *x,
                              const double *y, int32_t *cell_xy)
{
	for (int i = 0; i < n; i++) {
		cell_xy[2 * i] = (int)(x[i] * b->fx);
		cell_xy[2 * i + 1] = (int)(y[i] * b->fy);
	}
}

/* PBCcellX / PBCcellY, src/EDMD.c:2110-2124 */
static inline int wrap_cell(int a, int n)
{
	if (a < 0)
		return a + n;
	else if (a >= n)
		return a - n;
	return a;
}

/* PBC / PBCinsideCellX, src/EDMD.c:5896-5913, 5922-5936 */
static inline double min_image(double d, double half, double len)
{
	if (d >= half)
		return d - len;
	else if (d < -half)
		return d + len;
	return d;
}

/* Index-based cell list with the reference's link order: cellListInit
 * (src/EDMD.c:1906-1920) inserts i = 0..N-1 at the head (addToCell :2071-2072),
 * so a cell's list runs in DESCENDING particle index. */
typedef struct {
	int32_t *head; /* per cell, -1 = empty */
	int32_t *next; /* per particle */
} cell_list;

static int cell_list_build(cell_list *cl, const oracle_box *b, int n,
                           const int32_t *cell_xy)
{
	size_t nc = (size_t)b->nx * b->ny;
	cl->head = (int32_t *)malloc(nc * sizeof(int32_t));
	cl->next = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
	if (!cl->head || !cl->next)
		return -1;
	memset(cl->head, 0xff, nc * sizeof(int32_t));
	for (int i = 0; i < n; i++) {
		size_t c = (size_t)cell_xy[2 * i + 1] * b->nx + cell_xy[2 * i];
		cl->next[i] = cl->head[c];
		cl->head[c] = i;
	}
	return 0;
}

static void cell_list_free(cell_list *cl)
{
	free(cl->head);
	free(cl->next);
}

/* crossingEventNormal == crossingEventGrow for the 2-D, field-free build
 * (src/EDMD.c:2405-2482, 2343-2403; logTime is the identity when damping == 0,
 * :5709-5721).  Emits direction 1..4 and the absolute time. */
static void crossing(const oracle_box *b, double t, double x, double y,
                     double vx, double vy, int X, int Y, double *t_out,
                     uint8_t *dir_out)
{
	double tx, ty;
	int xx, yy;
	if (vx < 0) {
		tx = min_image(X * b->csx - x, b->half_lx, b->lx) / vx;
		xx = 1;
	} else {
		tx = min_image((1 + X) * b->csx - x, b->half_lx, b->lx) / vx;
		xx = 2;
	}
	if (vy < 0) {
		ty = min_image(Y * b->csy - y, b->half_ly, b->ly) / vy;
		yy = 3;
	} else {
		ty = min_image((1 + Y) * b->csy - y, b->half_ly, b->ly) / vy;
		yy = 4;
	}
	if (tx < ty) { /* strict: ties go to y, src/EDMD.c:2477-2480 */
		*dir_out = (uint8_t)xx;
		*t_out = t + tx;
	} else {
		*dir_out = (uint8_t)yy;
		*t_out = t + ty;
	}
}

/* collisionTimeNormal with lat2 == 0, src/EDMD.c:2661-2723 */
static double pair_time_normal(const oracle_box *b, double x1, double y1,
                               double vx1, double vy1, double r1, double x2,
                               double y2, double vx2, double vy2, double r2,
                               int *overlap)
{
	double dvx = vx2 - vx1;
	double dvy = vy2 - vy1;
	double dx = x2 - x1;
	double dy = y2 - y1;
	dx = min_image(dx, b->half_lx, b->lx);
	dy = min_image(dy, b->half_ly, b->ly);
	double bb = dx * dvx + dy * dvy;
	if (bb > 0)
		return NEVER;
	double v2 = dvx * dvx + dvy * dvy;
	double c = dx * dx + dy * dy - (4 * r1 * r2); /* DIST2, src/EDMD.c:35-39 */
	double det = bb * bb - v2 * c;
	if (c < -0.01)
		*overlap = 1; /* reference prints, dumps and exit(3)s, :2708-2717 */
	if (det < 0)
		return NEVER;
	return (-bb - sqrt(det)) / v2;
}

/* collisionTimeGrow with lat2 == 0, src/EDMD.c:2598-2659 */
static double pair_time_grow(const oracle_box *b, double x1, double y1,
                             double vx1, double vy1, double r1, double vr1,
                             double x2, double y2, double vx2, double vy2,
                             double r2, double vr2, int *overlap)
{
	double dvx = vx2 - vx1;
	double dvy = vy2 - vy1;
	double dvr = vr1 + vr2;
	double dx = x2 - x1;
	double dy = y2 - y1;
	double dr = sqrt(4 * r1 * r2); /* DIST */
	dx = min_image(dx, b->half_lx, b->lx);
	dy = min_image(dy, b->half_ly, b->ly);
	double bb = dx * dvx + dy * dvy - dvr * dr;
	double v2 = dvx * dvx + dvy * dvy;
	double d2 = dx * dx + dy * dy;
	double a = v2 - dvr * dvr;
	double det = bb * bb - a * (d2 - dr * dr);
	if (det < 0)
		return NEVER;
	double plus = (-bb + sqrt(det)) / a;
	double minus = (-bb - sqrt(det)) / a;
	if (((minus > 0) && (plus > 0) && (minus < plus)) ||
	    ((minus > 0) && (plus < 0)))
		return minus;
	else if (((minus > 0) && (plus > 0.000000001) && (plus < minus)) ||
	         ((minus < 0) && (plus > 0.000000001)))
		return plus;
	if (d2 - dr * dr < -0.01)
		*overlap = 1; /* reference exit(3)s, :2650-2653 */
	return NEVER;
}

int oracle_predict_all(const oracle_box *b, int n, double t, const double *x,
                       const double *y, const double *vx, const double *vy,
                       const double *rad, const double *vr,
                       const int32_t *cell_xy_in, int mode, double *t_cross,
                       uint8_t *dir, double *t_coll, int32_t *partner,
                       uint8_t *ctype, int32_t *overlap_pair)
{
	int32_t *cell_xy = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * 2 * sizeof(int32_t));
	if (cell_xy_in)
		memcpy(cell_xy, cell_xy_in, (size_t)n * 2 * sizeof(int32_t));
	else
		oracle_cells_from_coords(b, n, x, y, cell_xy);
	cell_list cl;
	if (cell_list_build(&cl, b, n, cell_xy) != 0) {
		free(cell_xy);
		return -1;
	}
	int found = 0;
	if (overlap_pair)
		overlap_pair[0] = overlap_pair[1] = -1;

	for (int i = 0; i < n; i++) {
		int X = cell_xy[2 * i], Y = cell_xy[2 * i + 1];
		crossing(b, t, x[i], y[i], vx[i], vy[i], X, Y, &t_cross[i], &dir[i]);

		/* collisionEventNormal :2829-2832,2959-3003 / Grow :3104-3106,3199-3220 */
		int best_j = 0;
		double best = (mode == ORACLE_MODE_GROW) ? 10000000 : NEVER;
		for (int j = -1; j <= 1; j++) {
			for (int k = -1; k <= 1; k++) {
				size_t c = (size_t)wrap_cell(Y + j, b->ny) * b->nx +
				           wrap_cell(X + k, b->nx);
				for (int p2 = cl.head[c]; p2 >= 0; p2 = cl.next[p2]) {
					if (p2 == i)
						continue;
					int ov = 0;
					double dt;
					if (mode == ORACLE_MODE_GROW)
						dt = pair_time_grow(b, x[i], y[i], vx[i], vy[i], rad[i],
						                    vr[i], x[p2], y[p2], vx[p2], vy[p2],
						                    rad[p2], vr[p2], &ov);
					else
						dt = pair_time_normal(b, x[i], y[i], vx[i], vy[i],
						                      rad[i], x[p2], y[p2], vx[p2],
						                      vy[p2], rad[p2], &ov);
					if (ov && !found) {
						found = 1;
						if (overlap_pair) {
							overlap_pair[0] = i;
							overlap_pair[1] = p2;
						}
					}
					if (best > dt) { /* strict >, first minimum wins, :2991 */
						best_j = p2;
						best = dt;
					}
				}
			}
		}
		t_coll[i] = t + best; /* addCollisionEvent(i, partner, t + dt) :3091 */
		partner[i] = best_j;
		if (ctype)
			ctype[i] = ORACLE_EV_COLLISION;
	}
	cell_list_free(&cl);
	free(cell_xy);
	return found;
}

/* PBCpostX/Y, src/EDMD.c:5948-5959 (a single +-L) */
static inline double wrap_pos(double v, double len)
{
	if (v < 0)
		v += len;
	else if (v >= len)
		v -= len;
	return v;
}

/* freeFlyNormal :4954-4990 (no field, no damping) / freeFlyGrow :4992-5007 */
void oracle_free_fly(const oracle_box *b, int n, int mode, double t_old,
                     const double *tp, double t_new, double *x, double *y,
                     const double *vx, const double *vy, double *rad,
                     const double *vr)
{
	for (int i = 0; i < n; i++) {
		double dt = t_new - (tp ? tp[i] : t_old);
		x[i] += dt * vx[i];
		y[i] += dt * vy[i];
		if (mode == ORACLE_MODE_GROW)
			rad[i] += dt * vr[i];
		x[i] = wrap_pos(x[i], b->lx);
		y[i] = wrap_pos(y[i], b->ly);
	}
}

int oracle_pcf_num_bins(double dr, double max_r)
{
	return (int)(max_r / dr); /* src/pcf.c:21 */
}

/* calculate_pcf, src/pcf.c:16-75.  The reference accumulates +2.0 per
 * unordered pair into a double; here the pairs are counted as integers
 * (exact, thread-order independent) and doubled at normalisation. */
int oracle_pcf(const oracle_box *b, int n, const double *x, const double *y,
               double dr, double max_r, uint64_t *counts, double *g_r,
               double *r_out)
{
	int num_bins = oracle_pcf_num_bins(dr, max_r);
	double bin_width = dr;
	if (num_bins <= 0)
		return num_bins;
	memset(counts, 0, (size_t)num_bins * sizeof(uint64_t));

#pragma omp parallel
	{
		uint64_t *loc = (uint64_t *)calloc((size_t)num_bins, sizeof(uint64_t));
#pragma omp for schedule(dynamic, 64)
		for (int i = 0; i < n; i++) {
			for (int j = i + 1; j < n; j++) {
				double dx = x[j] - x[i];
				double dy = y[j] - y[i];
				dx = min_image(dx, b->half_lx, b->lx);
				dy = min_image(dy, b->half_ly, b->ly);
				double r = sqrt(dx * dx + dy * dy);
				if (r < max_r) {
					int bin = (int)(r / bin_width);
					if (bin < num_bins)
						loc[bin] += 1;
				}
			}
		}
#pragma omp critical
		for (int k = 0; k < num_bins; k++)
			counts[k] += loc[k];
		free(loc);
	}

	/* normalisation, src/pcf.c:56-72 */
	double volume = b->lx * b->ly;
	double density = n / volume;
	for (int i = 0; i < num_bins; i++) {
		double r = (i + 0.5) * bin_width;
		if (r_out)
			r_out[i] = r;
		if (g_r) {
			double shell_volume = 2 * M_PI * r * bin_width;
			double norm = shell_volume * density * n;
			double g = 2.0 * (double)counts[i];
			g_r[i] = (norm > 0) ? g / norm : 0.0;
		}
	}
	return num_bins;
}

/* computeBOOPCutoff, src/boop.c:61-107.  cexp(k*theta*I) is (cos, sin) of the
 * rounded product k*theta; cabs is hypot, carg is atan2(im, re). */
void oracle_boop_cutoff(const oracle_box *b, int n, const double *x,
                        const double *y, const int32_t *cell_xy_in, double r_c,
                        double *q5, double *q6, double *q7, double *q6_arg,
                        int32_t *neighbors)
{
	int32_t *cell_xy = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * 2 * sizeof(int32_t));
	if (cell_xy_in)
		memcpy(cell_xy, cell_xy_in, (size_t)n * 2 * sizeof(int32_t));
	else
		oracle_cells_from_coords(b, n, x, y, cell_xy);
	cell_list cl;
	cell_list_build(&cl, b, n, cell_xy);

	for (int i = 0; i < n; i++) {
		int X = cell_xy[2 * i], Y = cell_xy[2 * i + 1];
		double s5r = 0, s5i = 0, s6r = 0, s6i = 0, s7r = 0, s7i = 0;
		int nb = 0;
		for (int j = -1; j <= 1; j++) {
			for (int k = -1; k <= 1; k++) {
				size_t c = (size_t)wrap_cell(Y + j, b->ny) * b->nx +
				           wrap_cell(X + k, b->nx);
				for (int p2 = cl.head[c]; p2 >= 0; p2 = cl.next[p2]) {
					if (p2 == i)
						continue;
					double dx = x[p2] - x[i];
					double dy = y[p2] - y[i];
					dx = min_image(dx, b->half_lx, b->lx);
					dy = min_image(dy, b->half_ly, b->ly);
					double r2 = dx * dx + dy * dy;
					if (r2 < r_c * r_c) {
						double theta = atan2(dy, dx);
						s5r += cos(5 * theta);
						s5i += sin(5 * theta);
						s6r += cos(6 * theta);
						s6i += sin(6 * theta);
						s7r += cos(7 * theta);
						s7i += sin(7 * theta);
						nb++;
					}
				}
			}
		}
		neighbors[i] = nb;
		if (nb > 0) {
			q5[i] = hypot(s5r, s5i) / nb;
			q6[i] = hypot(s6r, s6i) / nb;
			q7[i] = hypot(s7r, s7i) / nb;
			q6_arg[i] = atan2(s6i, s6r);
		} else {
			q5[i] = q6[i] = q7[i] = q6_arg[i] = 0.0;
		}
	}
	cell_list_free(&cl);
	free(cell_xy);
}

/* calculate_bond_order_pcf, src/pcf.c:77-167.  The reference walks ORDERED pairs
 * (i != j) and adds 1 and cos(k.d) per pair; cos is even and d_ji = -d_ij, so
 * both sums are twice the sums over unordered pairs, which is what is walked
 * here (counts as integers).  g6_r = sum / count per bin (0 where empty), g_r
 * normalised like calculate_pcf with the ordered count. */
int oracle_bond_order_pcf(const oracle_box *b, int n, const double *x, const double *y,
                          double dr, double max_r, double kx, double ky, uint64_t *counts,
                          double *g_r, double *g6_r)
{
	int num_bins = oracle_pcf_num_bins(dr, max_r);
	double bin_width = dr;
	if (num_bins <= 0)
		return num_bins;
	memset(counts, 0, (size_t)num_bins * sizeof(uint64_t));
	double *sum = (double *)calloc((size_t)num_bins, sizeof(double));
	for (int i = 0; i < n; i++) {
		for (int j = i + 1; j < n; j++) {
			double dx = x[j] - x[i];
			double dy = y[j] - y[i];
			dx = min_image(dx, b->half_lx, b->lx);
			dy = min_image(dy, b->half_ly, b->ly);
			double r = sqrt(dx * dx + dy * dy);
			if (r < max_r) {
				int bin = (int)(r / bin_width);
				if (bin < num_bins) {
					counts[bin] += 1;
					sum[bin] += cos(kx * dx + ky * dy); /* src/pcf.c:123 */
				}
			}
		}
	}
	double volume = b->lx * b->ly;
	double density = n / volume;
	for (int i = 0; i < num_bins; i++) {
		double r = (i + 0.5) * bin_width;
		g6_r[i] = counts[i] ? sum[i] / (double)counts[i] : 0.0; /* :147-153 */
		double expected = 2 * M_PI * r * bin_width * density * n;   /* :158-166 */
		g_r[i] = expected > 0 ? 2.0 * (double)counts[i] / expected : 0.0;
	}
	free(sum);
	return num_bins;
}

/* find_max_structure_factor_bragg, src/pcf.c:405-467: S(k) = |sum_j exp(i k.r_j)|^2 / N
 * on the reciprocal grid inside the wedge the reference searches; the first
 * maximum in its (iky, ikx) loop order wins.  s_out (optional) = that maximum. */
void oracle_bragg_peak(int n, const double *x, const double *y, double lx, double ly,
                       double expected_bragg, double *k_out, double *s_out)
{
	double max_S = -1, best[2] = {0.0, 0.0};
	double dkx = 2 * M_PI / lx, dky = 2 * M_PI / ly;
	double kx_min = floor((-expected_bragg - 0.8) / dkx) * dkx;
	double kx_max = ceil((expected_bragg + 0.8) / dkx) * dkx;
	double ky_min = floor((-expected_bragg - 0.8) / dky) * dky;
	double ky_max = ceil((expected_bragg + 0.8) / dky) * dky;
	for (int iky = 0; iky <= (int)((ky_max - ky_min) / dky); iky++) {
		for (int ikx = 0; ikx <= (int)((kx_max - kx_min) / dkx); ikx++) {
			double kx = kx_min + ikx * dkx;
			double ky = ky_max - iky * dky;
			double k_abs = sqrt(kx * kx + ky * ky);
			double theta = atan2(ky, kx);
			if (k_abs < 1.5 || theta < M_PI / 2 - M_PI / 5 || theta > M_PI / 2 + M_PI / 5)
				continue;
			double re = 0.0, im = 0.0;
			for (int j = 0; j < n; j++) {
				double phase = kx * x[j] + ky * y[j];
				re += cos(phase);
				im += sin(phase);
			}
			double S = (re * re + im * im) / n;
			if (S > max_S) {
				max_S = S;
				best[0] = kx;
				best[1] = ky;
			}
		}
	}
	k_out[0] = best[0];
	k_out[1] = best[1];
	if (s_out)
		*s_out = max_S;
}

/* pair loop of compute_g6_correlation, src/pcf.c:189-228 (same pair walk, min
 * image and bin as calculate_pcf; `creal(conj(psi6[i]) * psi6[j])` :204). */
int oracle_g6_correlation(const oracle_box *b, int n, const double *x, const double *y,
                          const double *psi_re, const double *psi_im, double dr, double max_r,
                          uint64_t *counts, double *g6_corr)
{
	int num_bins = oracle_pcf_num_bins(dr, max_r);
	if (num_bins <= 0)
		return num_bins;
	memset(counts, 0, (size_t)num_bins * sizeof(uint64_t));
	memset(g6_corr, 0, (size_t)num_bins * sizeof(double));
	for (int i = 0; i < n; i++) {
		for (int j = i + 1; j < n; j++) {
			double dx = x[j] - x[i];
			double dy = y[j] - y[i];
			dx = min_image(dx, b->half_lx, b->lx);
			dy = min_image(dy, b->half_ly, b->ly);
			double r = sqrt(dx * dx + dy * dy);
			if (r < max_r) {
				int bin = (int)(r / dr);
				if (bin >= 0 && bin < num_bins) {
					g6_corr[bin] += psi_re[i] * psi_re[j] + psi_im[i] * psi_im[j];
					counts[bin] += 1;
				}
			}
		}
	}
	for (int i = 0; i < num_bins; i++) /* :219-225 */
		g6_corr[i] = counts[i] ? g6_corr[i] / (double)counts[i] : 0.0;
	return num_bins;
}

/* initStructureFactor, src/struc.c:328-345 (`nqx = 2*qmax/xx + 1` truncates) */
void oracle_sq_grid(double q_max, double lx, double ly, int *nqx, int *nqy, double *qx, double *qy)
{
	double xx = 2 * M_PI / lx;
	double yy = 2 * M_PI / ly;
	*nqx = (int)(2 * q_max / xx + 1);
	*nqy = (int)(2 * q_max / yy + 1);
	if (qx)
		for (int i = 0; i < *nqx; i++)
			qx[i] = xx * (i - (*nqx - 1) / 2);
	if (qy)
		for (int i = 0; i < *nqy; i++)
			qy[i] = yy * (i - (*nqy - 1) / 2);
}

/* computeStructureFactor :364-384 / computeVelocityStructureFactor :386-408 */
void oracle_structure_factor(int n, const double *x, const double *y, const double *vx,
                             const double *vy, int nqx, const double *qx, int nqy,
                             const double *qy, int velocity, double *s)
{
#pragma omp parallel for collapse(2)
	for (int i = 0; i < nqx; i++) {
		for (int j = 0; j < nqy; j++) {
			double im = 0, re = 0;
			for (int k = 0; k < n; k++) {
				double qr = qx[i] * x[k] + qy[j] * y[k];
				if (velocity) {
					re += vx[k] * cos(qr) + vy[k] * sin(qr);
					im += vx[k] * sin(qr) - vy[k] * cos(qr);
				} else {
					re += cos(qr);
					im += sin(qr);
				}
			}
			s[i * nqy + j] = (re * re + im * im) / n;
		}
	}
}
