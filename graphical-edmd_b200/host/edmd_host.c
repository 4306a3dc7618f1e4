/*
 * edmd_host.c -- sequential C host of the B200 hot path.
 *
 * A from-scratch hard-disk event-driven MD driver that keeps what a user of
 * the reference touches -- its command line (-N/--number, -p/--phi, -x/--xs,
 * -q/--sizeratio, -a/--aspect, -t/--time, -D/--dt, -o/--dtimeThermo,
 * -T/--temperature, -v/--version = seed; src/EDMD.c:750-792), its event API (a
 * calendar of bucketed lists feeding a binary search tree with a cached
 * minimum, two event slots per particle; src/EDMD.c:2144-2327, 1923-1934) and
 * its ovito-readable LAMMPS text dump (src/EDMD.c:5052-5134) -- and calls CUDA
 * through the C ABI of include/edmd_cuda.h wherever the reference loops over
 * ALL particles:
 *     setup sweep          (eventListInit,  src/EDMD.c:2007-2012)
 *     thermostat tick      (addNoise,       src/EDMD.c:4828-4923)
 *     psi6 columns of a frame (saveTXT,     src/EDMD.c:5057-5060)
 *     g(r) at thermo time  (save_pcf,       src/EDMD.c:5636-5637)
 * The per-event predictions (1-2 particles, asynchronous positions, lat2 != 0)
 * stay here on the host next to the calendar, as in the reference
 * (doTheCollisionNormal src/EDMD.c:3542, doTheCrossing :2501).
 *
 * Unlike the reference's CLI build (where `noise` is a compile-time 0,
 * src/EDMD.c:289) the thermostat is a run-time switch: --noise 2 --dtnoise dt.
 * Start configuration: --init grow (default, the reference's default start,
 * src/EDMD.c:1690-1829: random points of radius 0 grown Lubachevsky-Stillinger
 * style at rate vr until t = 1/vr, then stopGrow :4740-4779; the setup sweep is
 * then a GROW-mode GPU sweep and stopGrow a NORMAL-mode one) or --init lattice
 * (jittered triangular lattice, no growth phase; needs phi below close packing
 * of the large disks).
 * --verify checks the first GPU sweep against this file's own per-particle
 * predictors (bit-exact) -- a check, not a fallback: without the CUDA library
 * nothing runs.
 */
#define _GNU_SOURCE
#include <getopt.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "edmd_cuda.h"

#define NEVER EDMD_NEVER
enum { EV_CELLCROSS = 0, EV_COLLISION = 1, EV_SCREENSHOT = 2, EV_THERMO = 3, EV_NOISE = 4, EV_GROWSTOP = 9 };

/* ---- options (names and defaults of src/EDMD.c:80-260 where they exist) --- */
static int N = 500;
static double phi = 0.38, sizeratio = 0.4, fractionSmallN = 0.3, aspectRatio = 1.0;
static double tmax = 4000, dtime = 100, dtimeThermo = 100, firstScreen = 1, firstThermo = 1;
static double T = 1.0, dtnoise = 0.5, gamm = 0.02;   /* gamm: the reference's default, src/EDMD.c:382 */
static unsigned noise_tick = 0;
static int noise = 0, seed = 1, boopThermo = 0, pcfThermo = 0, verify = 0, quiet = 0;
/* more of the reference's analysis switches (src/EDMD.c:223-231): areaThermo = Voronoi packing-fraction
 * column, strucThermo = S(q) mode (1 positions, 2.. = 0 velocity as saveStructureFactor's `mode`), pcfg6Thermo */
static int areaThermo = 0, strucThermo = 0, pcfg6Thermo = 0;
static double qmax = 0.3; /* src/EDMD.c:225 */
static int init_grow = 1, growing = 0;
static int bulk_ingest = 1;   /* --ingest bulk|seq: calendar rebuilt from the GPU's ingest plan / event by event */
static double vr = 0.1;
static const char *outdir = "dump";

/* ---- state ------------------------------------------------------------------ */
static double Lx, Ly, halfLx, halfLy, csx, csy;
static int Nx, Ny;
static double t = 0;
static double *px, *py, *pvx, *pvy, *prad, *pt;   /* pinned SoA (pt = local time) */
static double *pvr;                               /* growth rates (growth phase) */
static int32_t *pcell;                            /* interleaved X,Y */
static unsigned long *pcoll;
static int *ptype, *cnext, *cprev, *chead;
static unsigned long ncol = 0, ncross = 0;
static double collTermX = 0, collTermY = 0, collTermXY = 0, collTermYX = 0, lastThermoT = 0;
static double Einit = 1;   /* the reference's --initial-energy / -E (src/EDMD.c:369, :847): E/N after normalizePhysicalQ */

typedef struct node {
	struct node *lft, *rgt, *top;
	int i, j, q, type;
	double t;
	unsigned long collActual;
} node;
static node *events, *root, *treeMin, **paul;
static int paulN, actualPaul = 0;
static double paulTime = 0, dtPaul;

static edmd_ctx *gpu;
static double *g_tcross, *g_tcoll;
static uint8_t *g_dir, *g_type;
static int32_t *g_partner;
static int32_t *g_bucket, *g_next, *g_prev, *g_head;   /* edmd_cuda_calendar_plan outputs (pinned) */
static double gpu_sweep_seconds = 0, ingest_seconds = 0;
static int gpu_sweeps = 0, bulk_sweeps = 0;
#define NSPECIAL 10
static int special_queued[NSPECIAL];

static double now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void die_gpu(int rc, const char *what)
{
	fprintf(stderr, "edmd_host: %s failed (%d): %s\n", what, rc, gpu ? edmd_cuda_last_error(gpu) : "no context");
	exit(2);
}

/* ---- calendar: bucketed lists + BST (same contract as src/EDMD.c:2144-2327) -- */
static void tree_add(node *e)
{
	node *w = root;
	for (;;) {
		if (e->t < w->t) {
			if (!w->lft) { w->lft = e; break; }
			w = w->lft;
		} else {
			if (!w->rgt) { w->rgt = e; break; }
			w = w->rgt;
		}
	}
	e->lft = e->rgt = NULL;
	e->top = w;
	if (!treeMin || e->t < treeMin->t) treeMin = e;
}

static void tree_remove(node *e)
{
	node *top = e->top, *n;
	if (e == treeMin) {
		if (e->rgt) {
			treeMin = e->rgt;
			while (treeMin->lft) treeMin = treeMin->lft;
		} else
			treeMin = (top == root) ? NULL : top;
	}
	if (!e->lft && !e->rgt)
		n = NULL;
	else if (!e->lft) { n = e->rgt; n->top = top; }
	else if (!e->rgt) { n = e->lft; n->top = top; }
	else {
		n = e->rgt;
		while (n->lft) n = n->lft;
		if (n->top != e) {
			n->top->lft = n->rgt;
			if (n->rgt) n->rgt->top = n->top;
			e->rgt->top = n;
			n->rgt = e->rgt;
		}
		e->lft->top = n;
		n->lft = e->lft;
		n->top = top;
	}
	if (top->lft == e) top->lft = n;
	else top->rgt = n;
}

static void queue_add(node *e)
{
	double dt = e->t - paulTime;
	if (dt < dtPaul) {
		e->q = actualPaul;
		tree_add(e);
		return;
	}
	int k;
	if (dt >= dtPaul * paulN)
		k = paulN; /* overflow bucket: "never" events live here */
	else {
		k = actualPaul + (int)(dt / dtPaul);
		if (k >= paulN) k -= paulN;
	}
	e->q = k;
	e->lft = NULL;
	e->rgt = paul[k];
	if (e->rgt) e->rgt->lft = e;
	paul[k] = e;
}

static void queue_remove(node *e)
{
	if (e->q != actualPaul) {
		if (e->rgt) e->rgt->lft = e->lft;
		if (e->lft) e->lft->rgt = e->rgt;
		else paul[e->q] = e->rgt;
	} else
		tree_remove(e);
}

static node *queue_next(void)
{
	while (!treeMin) {
		actualPaul++;
		paulTime += dtPaul;
		if (actualPaul == paulN) {
			actualPaul = 0;
			node *p = paul[paulN];
			paul[paulN] = NULL;
			while (p) { node *nx = p->rgt; queue_add(p); p = nx; }
		}
		node *p = paul[actualPaul];
		while (p) { node *nx = p->rgt; tree_add(p); p = nx; }
		paul[actualPaul] = NULL;
	}
	return treeMin;
}

/* ---- geometry helpers (PBC :5896, PBCpost :5948, free flight :4954) ---------- */
static inline double min_image(double d, double half, double len)
{
	if (d >= half) return d - len;
	else if (d < -half) return d + len;
	return d;
}

static inline void free_fly(int i)
{
	double dt = t - pt[i];
	pt[i] = t;
	px[i] += dt * pvx[i];
	py[i] += dt * pvy[i];
	if (growing) prad[i] += dt * pvr[i]; /* freeFlyGrow :4992-5007 */
	if (px[i] < 0) px[i] += Lx; else if (px[i] >= Lx) px[i] -= Lx;
	if (py[i] < 0) py[i] += Ly; else if (py[i] >= Ly) py[i] -= Ly;
}

static inline int wrapc(int a, int n) { return a < 0 ? a + n : (a >= n ? a - n : a); }

static void cell_insert(int i)
{
	int c = pcell[2 * i + 1] * Nx + pcell[2 * i];
	cprev[i] = -1;
	cnext[i] = chead[c];
	if (chead[c] >= 0) cprev[chead[c]] = i;
	chead[c] = i;
}

static void cell_remove(int i)
{
	int c = pcell[2 * i + 1] * Nx + pcell[2 * i];
	if (cprev[i] < 0) chead[c] = cnext[i];
	else cnext[cprev[i]] = cnext[i];
	if (cnext[i] >= 0) cprev[cnext[i]] = cprev[i];
}

/* ---- per-event predictors (asynchronous form, lat2 = t - t_j) ---------------- */
static void predict_crossing(int i, double *tc, int *dir)
{
	double tx, ty;
	int xx, yy;
	int X = pcell[2 * i], Y = pcell[2 * i + 1];
	if (pvx[i] < 0) { tx = min_image(X * csx - px[i], halfLx, Lx) / pvx[i]; xx = 1; }
	else { tx = min_image((1 + X) * csx - px[i], halfLx, Lx) / pvx[i]; xx = 2; }
	if (pvy[i] < 0) { ty = min_image(Y * csy - py[i], halfLy, Ly) / pvy[i]; yy = 3; }
	else { ty = min_image((1 + Y) * csy - py[i], halfLy, Ly) / pvy[i]; yy = 4; }
	if (tx < ty) { *tc = t + tx; *dir = xx; }
	else { *tc = t + ty; *dir = yy; }
}

static void overlap_abort(int i, int j)
{
	fprintf(stderr, "\nERROR: Overlaps detected during SIMULATION between particle %d and particle %d !\n", i, j);
	exit(3);
}

/* collisionTimeGrow, src/EDMD.c:2598-2659 */
static double pair_time_grow(int i, int p2)
{
	double lat2 = t - pt[p2];
	double dvx = pvx[p2] - pvx[i], dvy = pvy[p2] - pvy[i], dvr = pvr[i] + pvr[p2];
	double dx = (px[p2] + lat2 * pvx[p2]) - px[i], dy = (py[p2] + lat2 * pvy[p2]) - py[i];
	double dr = sqrt(4 * prad[i] * (prad[p2] + lat2 * pvr[p2]));
	dx = min_image(dx, halfLx, Lx);
	dy = min_image(dy, halfLy, Ly);
	double b = dx * dvx + dy * dvy - dvr * dr;
	double v2 = dvx * dvx + dvy * dvy, d2 = dx * dx + dy * dy;
	double a = v2 - dvr * dvr;
	double det = b * b - a * (d2 - dr * dr);
	if (det < 0) return NEVER;
	double plus = (-b + sqrt(det)) / a, minus = (-b - sqrt(det)) / a;
	if (((minus > 0) && (plus > 0) && (minus < plus)) || ((minus > 0) && (plus < 0))) return minus;
	else if (((minus > 0) && (plus > 0.000000001) && (plus < minus)) || ((minus < 0) && (plus > 0.000000001))) return plus;
	if (d2 - dr * dr < -0.01) {
		fprintf(stderr, "\nERROR: Overlaps detected during GROWTH between particle %d and particle %d !\n", i, p2);
		exit(3);
	}
	return NEVER;
}

static void predict_collision(int i, double *tcoll, int *partner)
{
	double best = growing ? 10000000 : NEVER;
	int bj = 0;
	int X = pcell[2 * i], Y = pcell[2 * i + 1];
	for (int j = -1; j <= 1; j++)
		for (int k = -1; k <= 1; k++) {
			int c = wrapc(Y + j, Ny) * Nx + wrapc(X + k, Nx);
			for (int p2 = chead[c]; p2 >= 0; p2 = cnext[p2]) {
				if (p2 == i) continue;
				if (growing) {
					double dtg = pair_time_grow(i, p2);
					if (best > dtg) { best = dtg; bj = p2; }
					continue;
				}
				double lat2 = t - pt[p2];
				double dvx = pvx[p2] - pvx[i], dvy = pvy[p2] - pvy[i];
				double dx = min_image((px[p2] + lat2 * pvx[p2]) - px[i], halfLx, Lx);
				double dy = min_image((py[p2] + lat2 * pvy[p2]) - py[i], halfLy, Ly);
				double b = dx * dvx + dy * dvy;
				if (b > 0) continue;
				double v2 = dvx * dvx + dvy * dvy;
				double cc = dx * dx + dy * dy - (4 * prad[i] * prad[p2]);
				double det = b * b - v2 * cc;
				if (cc < -0.01) overlap_abort(i, p2);
				if (det < 0) continue;
				double dt = (-b - sqrt(det)) / v2;
				if (best > dt) { best = dt; bj = p2; }
			}
		}
	*tcoll = t + best;
	*partner = bj;
}

static void schedule_crossing(int i)
{
	double tc; int d;
	predict_crossing(i, &tc, &d);
	node *e = &events[i];
	e->j = d; e->type = EV_CELLCROSS; e->t = tc;
	queue_add(e);
}

static void schedule_collision(int i)
{
	double tc; int j;
	predict_collision(i, &tc, &j);
	node *e = &events[N + i];
	e->j = j; e->type = EV_COLLISION; e->t = tc; e->collActual = pcoll[j];
	queue_add(e);
}

static void schedule_special(int slot, int type, double when)
{
	node *e = &events[2 * N + slot];
	e->type = type; e->t = when; e->j = 0; e->i = -1;
	queue_add(e);
	special_queued[slot] = 1;
}

/* ---- the whole-system sweep on the GPU -------------------------------------- */
static edmd_mg *gpu_mg;   /* --gpus N > 1: the NORMAL-mode re-predict sweeps and the thermo's mean q6 run on N GPUs */

static void gpu_upload(void)
{
	int rc = edmd_cuda_upload(gpu, px, py, pvx, pvy, prad, pcell, t);
	if (rc) die_gpu(rc, "edmd_cuda_upload");
}

/* Whole-calendar rebuild from the device's ingest plan (edmd_cuda_calendar_plan):
 * every particle event is being replaced, so instead of 2N removals and 2N
 * insertions (one division and two cache misses each) the calendar is emptied,
 * the nodes are filled in one streaming pass from the plan -- bucket, list
 * neighbours, list heads, exactly what the sequential head insertions would
 * have produced -- and only the handful of special events and BST events go
 * through queue_add / tree_add.  Returns 0 (calendar emptied, only the special
 * events re-inserted) when the device declines the plan; the caller then
 * inserts the particle events one by one. */
static int ingest_from_plan(void)
{
	/* empty the calendar */
	memset(paul, 0, sizeof(node *) * (size_t)(paulN + 1));
	root->lft = root->rgt = NULL;
	treeMin = NULL;
	int32_t n_tree = 0;
	int rc = edmd_cuda_calendar_plan(gpu, paulTime, dtPaul, paulN, actualPaul, g_bucket, g_next, g_prev,
	                                 g_head, &n_tree);
	if (rc && rc != EDMD_EPLAN) die_gpu(rc, "edmd_cuda_calendar_plan");
	if (rc == 0) {
		/* streaming passes over independent nodes: split over the host cores */
#pragma omp parallel for schedule(static)
		for (int k = 0; k <= paulN; k++) paul[k] = g_head[k] >= 0 ? &events[g_head[k]] : NULL;
#pragma omp parallel for schedule(static)
		for (int i = 0; i < N; i++) {
			node *ev = &events[i];
			ev->j = g_dir[i]; ev->type = EV_CELLCROSS; ev->t = g_tcross[i];
			int b = g_bucket[i];
			if (b >= 0) {
				ev->q = b;
				ev->lft = g_prev[i] >= 0 ? &events[g_prev[i]] : NULL;
				ev->rgt = g_next[i] >= 0 ? &events[g_next[i]] : NULL;
			}
			ev = &events[N + i];
			ev->j = g_partner[i]; ev->type = EV_COLLISION; ev->t = g_tcoll[i];
			ev->collActual = pcoll[g_partner[i]];
			b = g_bucket[N + i];
			if (b >= 0) {
				ev->q = b;
				ev->lft = g_prev[N + i] >= 0 ? &events[g_prev[N + i]] : NULL;
				ev->rgt = g_next[N + i] >= 0 ? &events[g_next[N + i]] : NULL;
			}
		}
		/* the few events of the current bucket go into the BST, in the reference's order */
		if (n_tree > 0)
			for (int i = 0; i < N; i++) {
				if (g_bucket[i] < 0) { events[i].q = actualPaul; tree_add(&events[i]); }
				if (g_bucket[N + i] < 0) { events[N + i].q = actualPaul; tree_add(&events[N + i]); }
			}
	}
	/* the special events (a handful) go back in the ordinary way */
	for (int s = 0; s < NSPECIAL; s++)
		if (special_queued[s]) queue_add(&events[2 * N + s]);
	return rc == 0;
}

/* every particle must already be at time t.  remove_first: events are in the
 * calendar (thermostat tick, stopGrow) or not (setup). */
static void gpu_predict_all(int remove_first)
{
	double t0 = now();
	int32_t ov[2];
	int rc;
	if (gpu_mg && !growing) {
		/* row slabs over the GPUs (edmd_cuda_create_mg): same arrays in, same arrays out */
		rc = edmd_cuda_mg_upload(gpu_mg, px, py, pvx, pvy, prad, pcell, t);
		if (rc) { fprintf(stderr, "edmd_host: edmd_cuda_mg_upload failed (%d): %s\n", rc, edmd_cuda_mg_last_error(gpu_mg)); exit(2); }
		rc = edmd_cuda_mg_predict_all(gpu_mg, EDMD_MODE_NORMAL, g_tcross, g_dir, g_tcoll, g_partner, g_type, ov);
		if (rc == EDMD_EOVERLAP) overlap_abort(ov[0], ov[1]);
		if (rc) { fprintf(stderr, "edmd_host: edmd_cuda_mg_predict_all failed (%d): %s\n", rc, edmd_cuda_mg_last_error(gpu_mg)); exit(2); }
	} else {
		gpu_upload();
		rc = edmd_cuda_predict_all(gpu, growing ? EDMD_MODE_GROW : EDMD_MODE_NORMAL, growing ? pvr : NULL,
		                           g_tcross, g_dir, g_tcoll, g_partner, g_type, ov);
		if (rc == EDMD_EOVERLAP) overlap_abort(ov[0], ov[1]);
		if (rc) die_gpu(rc, "edmd_cuda_predict_all");
	}
	gpu_sweep_seconds += now() - t0;
	gpu_sweeps++;
	double t1 = now();
	if (bulk_ingest && !(gpu_mg && !growing)) {   /* the ingest plan is made from ONE device's predictions */
		if (ingest_from_plan()) {
			ingest_seconds += now() - t1;
			bulk_sweeps++;
			return;
		}
		remove_first = 0;   /* declined: the calendar is empty, insert one by one */
	}
	/* sequential calendar ingest in the reference's order: crossing, then collision */
	for (int i = 0; i < N; i++) {
		node *e = &events[i];
		if (remove_first) queue_remove(e);
		e->j = g_dir[i]; e->type = EV_CELLCROSS; e->t = g_tcross[i];
		queue_add(e);
		e = &events[N + i];
		if (remove_first) queue_remove(e);
		e->j = g_partner[i]; e->type = EV_COLLISION; e->t = g_tcoll[i];
		e->collActual = pcoll[g_partner[i]];
		queue_add(e);
	}
	ingest_seconds += now() - t1;
}

static void verify_first_sweep(void)
{
	long bad = 0;
	for (int i = 0; i < N; i++) {
		double tc, tl; int d, j;
		predict_crossing(i, &tc, &d);
		predict_collision(i, &tl, &j);
		if (tc != g_tcross[i] || d != g_dir[i] || tl != g_tcoll[i] || j != g_partner[i]) {
			if (bad < 5)
				fprintf(stderr, "verify: particle %d host (%.17g,%d,%.17g,%d) gpu (%.17g,%d,%.17g,%d)\n", i, tc, d,
				        tl, j, g_tcross[i], g_dir[i], g_tcoll[i], g_partner[i]);
			bad++;
		}
	}
	printf("verify: first GPU sweep vs host per-particle predictors: %ld mismatches of %d (bit-exact compare)\n", bad, N);
	if (bad) exit(5);
}

/* ---- event handlers ---------------------------------------------------------- */
static void do_collision(node *ev)
{
	int i = ev->i, j = ev->j;
	free_fly(i);
	if (ev->collActual != pcoll[j]) { /* stale partner: re-predict i only (:3553-3556) */
		schedule_collision(i);
		return;
	}
	ncol++;
	free_fly(j);
	pcoll[i]++;
	pcoll[j]++;
	double dx = min_image(px[j] - px[i], halfLx, Lx), dy = min_image(py[j] - py[i], halfLy, Ly);
	double dvx = pvx[j] - pvx[i], dvy = pvy[j] - pvy[i];
	if (growing) {
		/* doTheCollisionGrow, src/EDMD.c:3465-3540 (unit masses, res = 1) */
		double dist = sqrt(dx * dx + dy * dy), dxr = dx / dist, dyr = dy / dist;
		double k = dxr * dvx + dyr * dvy - (pvr[i] + pvr[j]);
		pvx[i] += k * dxr; pvy[i] += k * dyr;
		pvx[j] -= k * dxr; pvy[j] -= k * dyr;
	} else {
		/* elastic, unit masses: invMass*(1+res) = 1 (:3583-3605, 3802-3827) */
		double f = (dx * dvx + dy * dvy) / (4 * prad[i] * prad[j]);
		collTermX += f * dx * dx;
		collTermY += f * dy * dy;
		collTermXY += f * dx * dy;   /* :3807-3813 */
		collTermYX += f * dy * dx;
		pvx[i] += f * dx; pvy[i] += f * dy;
		pvx[j] -= f * dx; pvy[j] -= f * dy;
	}
	queue_remove(&events[i]);
	queue_remove(&events[j]);
	schedule_crossing(i);
	schedule_crossing(j);
	queue_remove(&events[N + j]);
	schedule_collision(i);
	schedule_collision(j);
}

static void do_crossing(node *ev)
{
	int i = ev->i;
	ncross++;
	free_fly(i);
	cell_remove(i);
	switch (ev->j) {
	case 1: pcell[2 * i] = pcell[2 * i] == 0 ? Nx - 1 : pcell[2 * i] - 1; break;
	case 2: pcell[2 * i] = pcell[2 * i] == Nx - 1 ? 0 : pcell[2 * i] + 1; break;
	case 3: pcell[2 * i + 1] = pcell[2 * i + 1] == 0 ? Ny - 1 : pcell[2 * i + 1] - 1; break;
	default: pcell[2 * i + 1] = pcell[2 * i + 1] == Ny - 1 ? 0 : pcell[2 * i + 1] + 1; break;
	}
	cell_insert(i);
	schedule_crossing(i);
	queue_remove(&events[N + i]);
	schedule_collision(i);
}

static double kinetic_energy(void)
{
	double E = 0;
	for (int i = 0; i < N; i++) E += 0.5 * (pvx[i] * pvx[i] + pvy[i] * pvy[i]);
	return E;
}

static double kinetic_energy(void);

/* stopGrow, src/EDMD.c:4740-4779 */
static void do_growstop(void)
{
	for (int i = 0; i < N; i++) { free_fly(i); pcoll[i] = 0; }
	ncol = ncross = 0;
	growing = 0;
	/* normalizePhysicalQ :5723-5764 on the device (edmd_cuda_normalize_velocities): the centre-of-mass
	 * velocity removed, E/N set to Einit; the new velocities come back for the event loop.
	 * (stopGrow re-inserts every particle's collision event before its crossing event, :4772-4778; the
	 * sweep below inserts crossing first like the other three sites -- which of two events with EXACTLY
	 * equal times pops first is the only thing that could differ.) */
	gpu_upload();
	int rc = edmd_cuda_normalize_velocities(gpu, Einit, NULL, NULL, NULL, NULL);
	if (rc) die_gpu(rc, "edmd_cuda_normalize_velocities");
	rc = edmd_cuda_download_state(gpu, NULL, NULL, pvx, pvy, NULL);
	if (rc) die_gpu(rc, "edmd_cuda_download_state");
	for (int i = 0; i < N; i++) pcoll[i]++;
	collTermX = collTermY = collTermXY = collTermYX = 0;
	lastThermoT = t;
	gpu_predict_all(1);
}

/* thermostat tick = addNoise with noise == 2 (velocity rescale, :4899-4902) or noise == 1
 * (Langevin kick, randomGaussian :5802-5826: done on the device with its counter-based
 * generator, the new velocities come back for the event loop) */
static void do_noise(void)
{
	if (noise == 1) {
		for (int i = 0; i < N; i++) {
			free_fly(i);
			pcoll[i]++;
		}
		gpu_upload();
		int rc = edmd_cuda_langevin_kick(gpu, T, gamm, dtnoise, (uint32_t)seed, noise_tick++);
		if (rc) die_gpu(rc, "edmd_cuda_langevin_kick");
		rc = edmd_cuda_download_state(gpu, NULL, NULL, pvx, pvy, NULL);
		if (rc) die_gpu(rc, "edmd_cuda_download_state");
		gpu_predict_all(1);
		schedule_special(2, EV_NOISE, t + dtnoise);
		return;
	}
	double E = kinetic_energy();
	double s = sqrt(E / N / T);
	for (int i = 0; i < N; i++) {
		free_fly(i);
		pcoll[i]++;
		pvx[i] /= s;
		pvy[i] /= s;
	}
	gpu_predict_all(1);
	schedule_special(2, EV_NOISE, t + dtnoise);
}

static FILE *fdump, *fthermo, *fpcf, *fstruc;
static char pcfg6Name[600];

static void do_screenshot(void)
{
	for (int i = 0; i < N; i++) free_fly(i);
	schedule_special(1, EV_SCREENSHOT, t + dtime);
	double *q5 = NULL, *q6 = NULL, *q7 = NULL, *qa = NULL;
	int32_t *nb = NULL;
	double *area = NULL;
	if (boopThermo || areaThermo) gpu_upload();
	if (boopThermo) {
		q5 = malloc(sizeof(double) * N); q6 = malloc(sizeof(double) * N); q7 = malloc(sizeof(double) * N);
		qa = malloc(sizeof(double) * N); nb = malloc(sizeof(int32_t) * N);
		/* `boopThermo == 1` Voronoi neighbours, `== 2` cutoff 2.5, src/EDMD.c:5053-5060 */
		int rc = boopThermo == 1 ? edmd_cuda_boop_voronoi(gpu, q5, q6, q7, qa, nb, NULL)
		                         : edmd_cuda_boop_cutoff(gpu, 2.5, q5, q6, q7, qa, nb, NULL);
		if (rc) die_gpu(rc, boopThermo == 1 ? "edmd_cuda_boop_voronoi" : "edmd_cuda_boop_cutoff");
	}
	if (areaThermo) { /* get_particle_voronoi_area, src/EDMD.c:5061-5064 */
		area = malloc(sizeof(double) * N);
		int rc = edmd_cuda_voronoi_cells(gpu, area, NULL, NULL);
		if (rc) die_gpu(rc, "edmd_cuda_voronoi_cells");
	}
	/* byte format of saveTXT, src/EDMD.c:5052-5069, row :5124-5134 */
	fprintf(fdump, "ITEM: TIMESTEP\n%lf\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n0 %lf\n0 %lf\n0 0\nITEM: ATOMS id type x y vx vy radius m coll",
	        t, N, Lx, Ly);
	if (boopThermo) fprintf(fdump, " q5 q6 q7 argq6 neighbors");
	if (areaThermo) fprintf(fdump, " packingFraction");
	fprintf(fdump, "\n");
	for (int i = 0; i < N; i++) {
		fprintf(fdump, "%d %d %.3lf %lf %lf %lf %lf %lf %d", i, ptype[i], px[i], py[i], pvx[i], pvy[i], prad[i], 1.0, ptype[i]);
		if (boopThermo) fprintf(fdump, " %.2lf %.2lf %.2lf %.2lf %d", q5[i], q6[i], q7[i], qa[i], nb[i]);
		if (areaThermo) fprintf(fdump, " %lf", M_PI * prad[i] * prad[i] / area[i]); /* :5131-5133 */
		fprintf(fdump, "\n");
	}
	fflush(fdump);
	free(q5); free(q6); free(q7); free(qa); free(nb); free(area);
	if (!quiet) printf("t = %-10.3lf collisions = %-12lu E/N = %.6lf\n", t, ncol, kinetic_energy() / N);
}

static double last_pressure = 0;

/* saveThermo, src/EDMD.c:5232-5609, for the reference's CLI build (Nthermo = 1: every call prints the
 * interval since the last one): the columns and formats of :5407-5577 --
 *   t Ncol E p px py pxy pyx [q6] a2
 * virial pressure tensor :5299-5328, mean q6 :5521-5536 (on the device), Sonine coefficient :5377-5403. */
static void do_thermo(void)
{
	schedule_special(3, EV_THERMO, t + dtimeThermo);
	double E = kinetic_energy();
	double area = Lx * Ly, dtm = t - lastThermoT;
	double eX = 0, eY = 0, eXY = 0, eYX = 0;
	for (int i = 0; i < N; i++) {
		eX += pvx[i] * pvx[i];
		eY += pvy[i] * pvy[i];
		eXY += pvx[i] * pvy[i];
		eYX += pvy[i] * pvx[i];
	}
	double inv = dtm > 0 ? 1 / (area * dtm) : 0;
	double pX = -1 * collTermX * inv + eX / area, pY = -1 * collTermY * inv + eY / area;
	double pXY = -1 * collTermXY * inv + eXY / area, pYX = -1 * collTermYX * inv + eYX / area;
	double p = dtm > 0 ? (-1 / dtm) * (collTermX + collTermY) / (2 * area) + E / area : E / area;
	last_pressure = p;
	collTermX = collTermY = collTermXY = collTermYX = 0;
	lastThermoT = t;
	/* a2 = <dvx^4> / (3 <dvx^2>^2) - 1 */
	double vav = 0, v2 = 0, v4 = 0;
	for (int i = 0; i < N; i++) vav += pvx[i];
	vav /= N;
	for (int i = 0; i < N; i++) {
		double d = pvx[i] - vav, d2 = d * d;
		v2 += d2;
		v4 += d2 * d2;
	}
	v2 /= N; v4 /= N;
	double a2 = v2 > 0.0 ? v4 / (3.0 * v2 * v2) - 1.0 : 0.0;
	if (t != 0) {
		fprintf(fthermo, "%lf %ld %lf %lf %lf %lf %.10lf %.10lf ", t, (long)ncol, E / N, p, pX, pY, pXY, pYX);
		if (boopThermo) { /* mean q6 of the current configuration, computeBOOPVoronoi / computeBOOPCutoff(2.5) */
			for (int i = 0; i < N; i++) free_fly(i);
			double q6 = 0;
			int rc;
			if (gpu_mg && boopThermo == 2) {
				rc = edmd_cuda_mg_upload(gpu_mg, px, py, pvx, pvy, prad, pcell, t);
				if (!rc) rc = edmd_cuda_mg_boop_cutoff(gpu_mg, 2.5, NULL, NULL, NULL, NULL, NULL, &q6);
				if (rc) { fprintf(stderr, "edmd_host: psi6 on the slabs failed (%d): %s\n", rc, edmd_cuda_mg_last_error(gpu_mg)); exit(2); }
			} else {
				gpu_upload();
				rc = boopThermo == 1 ? edmd_cuda_boop_voronoi(gpu, NULL, NULL, NULL, NULL, NULL, &q6)
				                     : edmd_cuda_boop_cutoff(gpu, 2.5, NULL, NULL, NULL, NULL, NULL, &q6);
				if (rc) die_gpu(rc, "edmd_cuda_boop (thermo)");
			}
			fprintf(fthermo, "%lf ", q6);
		}
		fprintf(fthermo, "%lf \n", a2);
		fflush(fthermo);
	}
	if (pcfThermo) {
		for (int i = 0; i < N; i++) free_fly(i);
		gpu_upload();
		double dr = 0.1, max_r = (Lx < Ly ? Lx : Ly) / 2;
		int nbins = 0;
		edmd_cuda_pcf(gpu, dr, max_r, NULL, NULL, &nbins);
		uint64_t *cnt = malloc(sizeof(uint64_t) * (nbins > 0 ? nbins : 1));
		double *g = malloc(sizeof(double) * (nbins > 0 ? nbins : 1));
		int rc = edmd_cuda_pcf(gpu, dr, max_r, cnt, g, &nbins);
		if (rc) die_gpu(rc, "edmd_cuda_pcf");
		for (int b = 0; b < nbins; b++) fprintf(fpcf, "%lf %lf %lf\n", t, (b + 0.5) * dr, g[b]);
		fflush(fpcf);
		free(cnt); free(g);
	}
	if (strucThermo || pcfg6Thermo) {
		for (int i = 0; i < N; i++) free_fly(i);
		gpu_upload();
	}
	if (strucThermo) { /* saveStructureFactor(particles, strucThermo), src/struc.c:409-425: mode 0 = velocity */
		int nqx = 0, nqy = 0;
		edmd_cuda_structure_factor(gpu, qmax, 0, &nqx, &nqy, NULL, NULL, NULL, NULL, NULL);
		double *sq = malloc(sizeof(double) * (size_t)(nqx * nqy > 0 ? nqx * nqy : 1));
		int rc = edmd_cuda_structure_factor(gpu, qmax, strucThermo == 2, &nqx, &nqy, NULL, NULL, sq, NULL, NULL);
		if (rc) die_gpu(rc, "edmd_cuda_structure_factor");
		for (int i = 0; i < nqx; i++) {
			for (int j = 0; j < nqy; j++) fprintf(fstruc, "%g ", sq[i * nqy + j]);
			fprintf(fstruc, "\n");
		}
		fflush(fstruc);
		free(sq);
	}
	if (pcfg6Thermo) { /* save_pcf_g6(name, particles, N, 2, fmin(Lx, Ly)/2, ...), src/EDMD.c:5639-5641, src/pcf.c:297-336 */
		double dr = 2, max_r = (Lx < Ly ? Lx : Ly) / 2;
		int nbins = 0;
		edmd_cuda_g6_correlation(gpu, dr, max_r, NULL, NULL, NULL, NULL, &nbins);
		double *g6 = malloc(sizeof(double) * (nbins > 0 ? nbins : 1));
		int rc = edmd_cuda_g6_correlation(gpu, dr, max_r, NULL, NULL, NULL, g6, &nbins);
		if (rc) die_gpu(rc, "edmd_cuda_g6_correlation");
		FILE *chk = fopen(pcfg6Name, "r");
		FILE *f = fopen(pcfg6Name, chk ? "a" : "w");
		if (!chk) {
			for (int b = 0; b < nbins; b++) fprintf(f, "%lf ", (b + 0.5) * dr);
			fprintf(f, "\n");
		} else fclose(chk);
		for (int b = 0; b < nbins; b++) fprintf(f, "%lf ", g6[b]);
		fprintf(f, "\n");
		fclose(f);
		free(g6);
	}
}

/* ---- set-up ------------------------------------------------------------------ */
static uint64_t rng_state;
static double urand(void)
{
	uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	return (z >> 11) * (1.0 / 9007199254740992.0);
}

static void *pinned(size_t bytes)
{
	void *p = NULL;
	int rc = edmd_cuda_host_alloc(&p, bytes);
	if (rc) { fprintf(stderr, "edmd_host: pinned allocation failed (%d)\n", rc); exit(2); }
	return p;
}

static void particles_init(void)
{
	/* box from N, phi as constantInit does (src/EDMD.c:1121-1152) */
	double r2 = 1 + fractionSmallN * (sizeratio * sizeratio - 1);
	/* jittered triangular lattice in a near-square box (Hex formulas :1026-1033) */
	int ny = (int)floor(sqrt((double)N));
	ny -= ny % 2;
	if (ny < 2) ny = 2;
	int nx = N / ny;
	N = nx * ny;
	Ly = sqrt(sqrt(3.0) / 2.0 * ny / nx * M_PI * N / phi * r2) * sqrt(aspectRatio);
	Lx = 2.0 / sqrt(3.0) * nx / ny * Ly / aspectRatio;
	double dx = Lx / nx, dy = Ly / ny;
	if (!init_grow && dx <= 2.0) { fprintf(stderr, "edmd_host: phi too large for --init lattice\n"); exit(1); }
	/* optimizeGrowConstant, src/EDMD.c:5766-5799 */
	vr = phi < 0.8 ? 0.1 : 0.1 * pow(0.8 / phi, 30);
	if (phi > 0.7) vr /= 8;
	px = pinned(sizeof(double) * N); py = pinned(sizeof(double) * N);
	pvx = pinned(sizeof(double) * N); pvy = pinned(sizeof(double) * N);
	prad = pinned(sizeof(double) * N); pcell = pinned(sizeof(int32_t) * 2 * N);
	pt = malloc(sizeof(double) * N);
	pvr = pinned(sizeof(double) * N);
	pcoll = calloc(N, sizeof(unsigned long));
	ptype = malloc(sizeof(int) * N);
	double amp = 0.3 * (dx - 2.0), sx = 0, sy = 0;
	for (int i = 0; i < N; i++) {
		int a = i % nx, b = i / nx;
		px[i] = a * dx + (b % 2) * dx / 2 + amp * (2 * urand() - 1);
		py[i] = b * dy + amp * (2 * urand() - 1);
		if (px[i] < 0) px[i] += Lx; else if (px[i] >= Lx) px[i] -= Lx;
		if (py[i] < 0) py[i] += Ly; else if (py[i] >= Ly) py[i] -= Ly;
		double u1 = 1 - urand(), u2 = urand(), rr = sqrt(-2 * log(u1));
		pvx[i] = rr * cos(2 * M_PI * u2);
		pvy[i] = rr * sin(2 * M_PI * u2);
		sx += pvx[i]; sy += pvy[i];
		ptype[i] = urand() < fractionSmallN ? 0 : 1;
		prad[i] = ptype[i] ? 1.0 : sizeratio;
		pvr[i] = 0;
		pt[i] = 0;
		if (init_grow) { /* random points, radius 0, growing to the target radius at t = 1/vr */
			px[i] = urand() * Lx;
			py[i] = urand() * Ly;
			pvr[i] = vr * prad[i];
			prad[i] = 0;
		}
	}
	for (int i = 0; i < N; i++) { pvx[i] -= sx / N; pvy[i] -= sy / N; }
	/* growth runs at E/N = 0.05 (EinitGrow :1692); stopGrow's normalizePhysicalQ rescales to Einit.  A lattice
	 * start has no stopGrow: it begins at the thermostat's temperature when there is one, else at Einit */
	double s = sqrt(kinetic_energy() / N / (init_grow ? 0.05 : (noise ? T : Einit)));
	for (int i = 0; i < N; i++) { pvx[i] /= s; pvy[i] /= s; }
	growing = init_grow;
}

int main(int argc, char **argv)
{
	static struct option longopt[] = {
		{"number", required_argument, NULL, 'N'}, {"phi", required_argument, NULL, 'p'},
		{"xs", required_argument, NULL, 'x'}, {"sizeratio", required_argument, NULL, 'q'},
		{"aspect", required_argument, NULL, 'a'}, {"time", required_argument, NULL, 't'},
		{"dt", required_argument, NULL, 'D'}, {"dtimeThermo", required_argument, NULL, 'o'},
		{"temperature", required_argument, NULL, 'T'}, {"version", required_argument, NULL, 'v'},
		{"noise", required_argument, NULL, 1001}, {"dtnoise", required_argument, NULL, 1002},
		{"boop", no_argument, NULL, 1003}, {"pcf", no_argument, NULL, 1004},
		{"verify", no_argument, NULL, 1005}, {"outdir", required_argument, NULL, 1006},
		{"quiet", no_argument, NULL, 1007}, {"device", required_argument, NULL, 1008},
		{"init", required_argument, NULL, 1009}, {"ingest", required_argument, NULL, 1010},
		{"boop-voronoi", no_argument, NULL, 1011}, {"area", no_argument, NULL, 1012},
		{"struc", required_argument, NULL, 1013}, {"qmax", required_argument, NULL, 1014},
		{"pcfg6", no_argument, NULL, 1015}, {"gamma", required_argument, NULL, 1016},
		{"initial-energy", required_argument, NULL, 'E'}, {"gpus", required_argument, NULL, 1017},
		{"slabs", required_argument, NULL, 1018},
		{NULL, 0, NULL, 0}};
	int c, device = 0, ngpus = 1, nslabs = 0;
	while ((c = getopt_long(argc, argv, "N:p:x:q:a:t:D:o:T:v:E:", longopt, NULL)) != -1) {
		switch (c) {
		case 'N': N = atoi(optarg); break;
		case 'p': phi = atof(optarg); break;
		case 'x': fractionSmallN = atof(optarg); break;
		case 'q': sizeratio = atof(optarg); break;
		case 'a': aspectRatio = atof(optarg); break;
		case 't': tmax = atof(optarg); break;
		case 'D': dtime = atof(optarg); break;
		case 'o': dtimeThermo = atof(optarg); break;
		case 'T': T = atof(optarg); break;
		case 'v': seed = atoi(optarg); break;
		case 'E': Einit = atof(optarg); break;
		case 1001: noise = atoi(optarg); break;
		case 1002: dtnoise = atof(optarg); break;
		case 1003: boopThermo = 2; break;
		case 1004: pcfThermo = 1; break;
		case 1005: verify = 1; break;
		case 1006: outdir = optarg; break;
		case 1007: quiet = 1; break;
		case 1008: device = atoi(optarg); break;
		case 1009: init_grow = strcmp(optarg, "lattice") != 0; break;
		case 1010: bulk_ingest = strcmp(optarg, "seq") != 0; break;
		case 1011: boopThermo = 1; break;
		case 1012: areaThermo = 1; break;
		case 1013: strucThermo = atoi(optarg) == 0 ? 2 : 1; break; /* reference: mode 0 = velocity S(q), else positions */
		case 1014: qmax = atof(optarg); break;
		case 1015: pcfg6Thermo = 1; break;
		case 1016: gamm = atof(optarg); break;
		case 1017: ngpus = atoi(optarg); break;
		case 1018: nslabs = atoi(optarg); break;   /* slabs dealt round-robin to the GPUs (default: one per GPU) */
		default: fprintf(stderr, "usage: edmd_host -N n --phi f [-x xs -q ratio -a aspect -t tmax -D dt -o dtThermo -T temp -v seed -E Einit]\n"
		                         "       [--init grow|lattice] [--ingest bulk|seq] [--noise 1|2 --dtnoise dt --gamma g] [--boop | --boop-voronoi] [--area] [--pcf] [--pcfg6] [--struc mode --qmax q] [--verify] [--outdir dir] [--quiet] [--gpus n [--slabs m]]\n");
			return 2;
		}
	}
	if (noise != 0 && noise != 1 && noise != 2) { fprintf(stderr, "edmd_host: --noise 0 (none), 1 (Langevin kick, --gamma) and 2 (velocity rescale) are implemented\n"); return 2; }
	rng_state = 0x1234567ull * (uint64_t)(seed + 1);

	/* the pinned allocator needs the CUDA runtime: touch the library first */
	particles_init();
	int rc = edmd_cuda_create(device, N, Lx, Ly, &gpu);
	if (rc) die_gpu(rc, "edmd_cuda_create");
	if (nslabs < ngpus) nslabs = ngpus;
	if (nslabs > 1) {
		int devs[64];
		if (nslabs > 64 || ngpus < 1) { fprintf(stderr, "edmd_host: --gpus / --slabs out of range\n"); return 2; }
		for (int k = 0; k < nslabs; k++) devs[k] = device + k % ngpus;
		rc = edmd_cuda_create_mg(nslabs, devs, N, Lx, Ly, &gpu_mg);
		if (rc) { fprintf(stderr, "edmd_host: edmd_cuda_create_mg(%d slabs on %d GPUs) failed (%d)\n", nslabs, ngpus, rc); return 2; }
	}
	edmd_box box;
	edmd_cuda_get_box(gpu, &box);
	Nx = box.nxcells; Ny = box.nycells; csx = box.cellx_size; csy = box.celly_size;
	halfLx = box.half_lx; halfLy = box.half_ly;
	dtPaul = box.dt_paul;
	paulN = N;

	chead = malloc(sizeof(int) * Nx * Ny); cnext = malloc(sizeof(int) * N); cprev = malloc(sizeof(int) * N);
	for (int k = 0; k < Nx * Ny; k++) chead[k] = -1;
	for (int i = 0; i < N; i++) {
		pcell[2 * i] = (int)(px[i] * box.cellx_fac);     /* coordToCell :2098-2107 */
		pcell[2 * i + 1] = (int)(py[i] * box.celly_fac);
		cell_insert(i);
	}
	events = calloc(2 * N + 10, sizeof(node));
	paul = calloc(paulN + 1, sizeof(node *));
	root = &events[2 * N];
	root->t = NEVER + 1;
	for (int i = 0; i < N; i++) events[i].i = events[N + i].i = i;
	g_tcross = pinned(sizeof(double) * N); g_tcoll = pinned(sizeof(double) * N);
	g_dir = pinned(N); g_type = pinned(N); g_partner = pinned(sizeof(int32_t) * N);
	g_bucket = pinned(sizeof(int32_t) * 2 * N); g_next = pinned(sizeof(int32_t) * 2 * N);
	g_prev = pinned(sizeof(int32_t) * 2 * N); g_head = pinned(sizeof(int32_t) * ((size_t)paulN + 1));

	/* customName, src/EDMD.c:6015-6116: the reference's file names (under --outdir instead of dump/),
	 * the first version number that does not exist yet */
	mkdir(outdir, 0777);
	char name[768], base[512];
	{
		int nsmall = 0;
		for (int i = 0; i < N; i++) nsmall += ptype[i] == 0;
		int len = snprintf(base, sizeof base, "%s/N_%dres_%.3lfphi_%.6lfq_%.3lfrat_%.3lfLx_%.3lfLy_%.3lf", outdir, N, 1.0,
		                   phi, (double)nsmall / N, sizeratio, Lx, Ly);
		if (noise) len += snprintf(base + len, sizeof base - len, "gamma_%.8lfT_%.3lfdt_%.3lf", gamm, T, dtnoise);
		len += snprintf(base + len, sizeof base - len, "v_");
		int version = 0;
		for (;; version++) {
			snprintf(name, sizeof name, "%s%d.dump", base, version);
			if (access(name, F_OK) != 0) break;
		}
		snprintf(base + len, sizeof base - len, "%d", version);
	}
	snprintf(name, sizeof name, "%s.dump", base); fdump = fopen(name, "w");
	snprintf(name, sizeof name, "%s.thermo", base); fthermo = fopen(name, "w");
	if (pcfThermo) { snprintf(name, sizeof name, "%s.pcf", base); fpcf = fopen(name, "w"); }
	if (strucThermo) {
		snprintf(name, sizeof name, "%s.struc", base); fstruc = fopen(name, "w");
		if (fstruc) { /* header of initStructureFactor, src/struc.c:347-355 */
			int nqx = 0, nqy = 0;
			edmd_cuda_structure_factor(gpu, qmax, 0, &nqx, &nqy, NULL, NULL, NULL, NULL, NULL);
			double *qx = malloc(sizeof(double) * (nqx + 1)), *qy = malloc(sizeof(double) * (nqy + 1));
			edmd_cuda_structure_factor(gpu, qmax, 0, &nqx, &nqy, qx, qy, NULL, NULL, NULL);
			fprintf(fstruc, "-%lf \n", 0.0);
			for (int i = 0; i < nqx; i++) fprintf(fstruc, "%g ", qx[i]);
			fprintf(fstruc, "\n");
			for (int j = 0; j < nqy; j++) fprintf(fstruc, "%g ", qy[j]);
			fprintf(fstruc, "\n");
			free(qx); free(qy);
		}
	}
	if (pcfg6Thermo) { snprintf(pcfg6Name, sizeof pcfg6Name, "%s.pcfg6", base); remove(pcfg6Name); }
	if (!fdump || !fthermo || (pcfThermo && !fpcf) || (strucThermo && !fstruc)) { perror("edmd_host: output files"); return 1; }
	fprintf(fthermo, "t Ncol E p px py pxy pyx %sa2 \n", boopThermo ? "q6 " : "");   /* header, src/EDMD.c:1229-1285 */

	if (!quiet)
		printf("edmd_host: N = %d  phi = %g  Lx = %.3f  Ly = %.3f  cells = %d x %d  noise = %d\n", N, phi, Lx, Ly, Nx, Ny, noise);
	double wall0 = now();
	double t_grow = init_grow ? 1 / vr : 0;   /* eventListInit shifts everything by 1/vr :1958-1976 */
	tmax += t_grow;
	schedule_special(3, EV_THERMO, t_grow + firstThermo);
	schedule_special(1, EV_SCREENSHOT, t_grow + firstScreen);
	if (noise) schedule_special(2, EV_NOISE, t_grow + dtnoise);
	if (init_grow) schedule_special(5, EV_GROWSTOP, t_grow);
	gpu_predict_all(0); /* setup sweep */
	if (verify) verify_first_sweep();
	double wall_setup = now() - wall0;

	double wall1 = now();
	while (t <= tmax) {
		node *ev = queue_next();
		t = ev->t;
		queue_remove(ev);
		if (ev - events >= 2 * N) special_queued[ev - events - 2 * N] = 0;
		switch (ev->type) {
		case EV_COLLISION: do_collision(ev); break;
		case EV_CELLCROSS: do_crossing(ev); break;
		case EV_NOISE: do_noise(); break;
		case EV_GROWSTOP: do_growstop(); wall1 = now(); break;
		case EV_SCREENSHOT: do_screenshot(); break;
		case EV_THERMO: do_thermo(); break;
		}
	}
	double wall = now() - wall1;
	printf("edmd_host: whole run %.3f s (setup + growth phase + event loop)\n", now() - wall0);
	printf("edmd_host: %lu collisions, %lu crossings in %.3f s => %.4g coll/s ; setup %.3f s ; "
	       "%d GPU sweeps, %.3f ms each (upload + K0 + K1 + download) ; calendar ingest %.3f ms each "
	       "(%d from the device plan) ; E/N = %.6f ; p = %.6f\n",
	       ncol, ncross, wall, ncol / wall, wall_setup, gpu_sweeps,
	       gpu_sweeps ? 1e3 * gpu_sweep_seconds / gpu_sweeps : 0.0,
	       gpu_sweeps ? 1e3 * ingest_seconds / gpu_sweeps : 0.0, bulk_sweeps, kinetic_energy() / N, last_pressure);
	fclose(fdump); fclose(fthermo);
	if (fpcf) fclose(fpcf);
	if (fstruc) fclose(fstruc);
	if (gpu_mg) edmd_cuda_destroy_mg(gpu_mg);
	edmd_cuda_destroy(gpu);
	return 0;
}
