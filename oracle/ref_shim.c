/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin driver around the UNMODIFIED reference (Syrocco/Graphical-EDMD).  The
 * reference sources are not copied into this repository: this translation unit
 * textually includes `EDMD.c` from the mounted reference tree at build time
 * (oracle/Makefile passes -I<REF>/src), renaming its main(), exactly the unity
 * build the reference's own Makefile does (Makefile:31,60).  The result is
 * oracle/_ref/libedmd_ref.so (git-ignored, travels to the GPU box).
 *
 * Must be compiled as C (the unity TU relies on tentative definitions).
 *
 * Every ref_* entry point below fills the reference's globals from plain
 * arrays, then runs the reference's own functions:
 *   boxConstantHelper   src/EDMD.c:679
 *   cellListInit        src/EDMD.c:1906
 *   crossingEvent{Normal,Grow}, collisionEvent{Normal,Grow}
 *                       src/EDMD.c:2343,2405,2829,3104
 *   addNoise            src/EDMD.c:4828   (the recurring full re-predict)
 *   freeFlyNormal/Grow  src/EDMD.c:4954,4992
 *   calculate_pcf       src/pcf.c:16
 *   computeBOOPCutoff   src/boop.c:61
 */
/* The reference's main() ends with freeArrays() (src/EDMD.c:659, 2022-2040).  To take
 * the final state of a whole reference run in memory (ref_run_main / ref_export_state
 * below: the equilibrated liquid behind tests/golden/liquid_*.npz) the allocator call
 * -- not the reference's source -- is interposed inside this test shim: while a run is
 * being kept, free(particles) is deferred. */
#include <stdlib.h>
static void shim_free(void *p);
#define free(p) shim_free((void *)(p))
#define main edmd_reference_main
#include "EDMD.c"
#undef main
#undef free

#include <stdint.h>

static int shim_keep_particles = 0;
static void *shim_kept = NULL;

static void shim_free(void *p)
{
	if (shim_keep_particles && p && p == (void *)particles) {
		shim_kept = p;
		return;
	}
	free(p);
}

/* The unmodified reference's own main() (growth start, event loop, thermostat ... as
 * its command line says), keeping particles[] alive past its freeArrays(). */
int ref_run_main(int argc, char **argv)
{
	shim_keep_particles = 1;
	shim_kept = NULL;
	int r = edmd_reference_main(argc, argv);
	shim_keep_particles = 0;
	return r;
}

/* Exact counters of the run edmd_reference_main() / ref_run_main() just finished: the reference's globals
 * ncol (collisions since stopGrow zeroed it, src/EDMD.c:4746) and t. */
unsigned long ref_last_ncol(void) { return ncol; }
double ref_last_time(void) { return t; }

/* State of the run ref_run_main() just finished, every particle brought to the final
 * time by the reference's freeFly (as takeAScreenshot does, src/EDMD.c:4659-4661);
 * box[0..2] = Lx, Ly, t.  Frees the kept array.  Returns N, or < 0. */
int ref_export_state(int cap, double *x, double *y, double *vx, double *vy, double *rad, double *box)
{
	if (!shim_kept || (void *)particles != shim_kept)
		return -1;
	if (N > cap) {
		free(particles);
		particles = NULL;
		shim_kept = NULL;
		return -2;
	}
	for (int i = 0; i < N; i++) {
		particle *p = particles + i;
		freeFly(p);
		x[i] = p->x;
		y[i] = p->y;
		vx[i] = p->vx;
		vy[i] = p->vy;
		rad[i] = p->rad;
	}
	box[0] = Lx;
	box[1] = Ly;
	box[2] = t;
	free(particles);
	particles = NULL;
	shim_kept = NULL;
	return N;
}

static int shim_live = 0;
static int shim_total = 0;

void ref_teardown(void)
{
	if (!shim_live)
		return;
	free(particles);
	free(cellList);
	if (eventList) {
		for (int i = 0; i < shim_total; i++)
			free(eventList[i]);
		free(eventList);
	}
	free(eventPaul);
	particles = NULL;
	cellList = NULL;
	eventList = NULL;
	eventPaul = NULL;
	shim_live = 0;
}

/* Load a synchronous snapshot.  cell_xy == NULL: cells come from the
 * reference's cellListInit (coordToCell).  cell_xy != NULL: host-owned cell
 * ids (interleaved X,Y) are linked in ascending particle order with the same
 * head insertion addToCell does (src/EDMD.c:2071-2077).  vr may be NULL. */
int ref_setup(int n, double lx, double ly, double t0,
              const double *x, const double *y, const double *vx,
              const double *vy, const double *rad, const double *vr_in,
              const int *cell_xy)
{
	ref_teardown();
	N = n;
	Lx = lx;
	Ly = ly;
	t = t0;
	boxConstantHelper();
	particles = calloc(N, sizeof(particle));
	if (!particles)
		return -1;
	for (int i = 0; i < N; i++) {
		particle *p = particles + i;
		p->num = i;
		p->x = x[i];
		p->y = y[i];
		p->vx = vx[i];
		p->vy = vy[i];
		p->rad = rad[i];
		p->vr = vr_in ? vr_in[i] : 0.0;
		p->m = 1;
		p->t = t0;
		p->type = 1;
		p->coll = 0;
	}
	if (!cell_xy) {
		cellListInit();
	} else {
		cellList = calloc((size_t)Nxcells * Nycells, sizeof(particle *));
		for (int i = 0; i < N; i++) {
			particle *p = particles + i;
			p->cell[0] = cell_xy[2 * i];
			p->cell[1] = cell_xy[2 * i + 1];
			p->prv = NULL;
			p->nxt = cellList[p->cell[1] * Nxcells + p->cell[0]];
			cellList[p->cell[1] * Nxcells + p->cell[0]] = p;
			if (p->nxt)
				p->nxt->prv = p;
		}
	}

	/* calendar storage as in eventListInit (src/EDMD.c:1937-1953), without
	 * scheduling the THERMO / SCREENSHOT / GROWSTOP bookkeeping events */
	shim_total = 2 * N + suppSizeEvent;
	eventList = calloc(shim_total, sizeof(node *));
	eventPaul = calloc(paulListN + 1, sizeof(node *));
	for (int i = 0; i < shim_total; i++)
		eventList[i] = calloc(1, sizeof(node));
	root = eventList[2 * N];
	root->t = never + 1;
	root->rgt = root->lft = root->top = NULL;
	treeMin = NULL;
	paulTime = t0;
	actualPaulList = 0;
	for (int i = 0; i < N; i++) {
		eventList[i]->i = i;
		eventList[N + i]->i = i;
	}
	shim_live = 1;
	return 0;
}

static void shim_set_mode(int grow)
{
	if (grow) {
		collisionEvent = &collisionEventGrow;
		doTheCollision = &doTheCollisionGrow;
		freeFly = &freeFlyGrow;
		doTheWall = &doTheWallGrow;
		crossingEvent = &crossingEventGrow;
	} else {
		collisionEvent = &collisionEventNormal;
		doTheCollision = &doTheCollisionNormal;
		freeFly = &freeFlyNormal;
		doTheWall = &doTheWallNormal;
		crossingEvent = &crossingEventNormal;
	}
}

void ref_get_box(int *nx, int *ny, double *csx, double *csy, double *fx,
                 double *fy)
{
	*nx = Nxcells;
	*ny = Nycells;
	*csx = cellxSize;
	*csy = cellySize;
	*fx = cellxFac;
	*fy = cellyFac;
}

void ref_get_cells(int *cell_xy)
{
	for (int i = 0; i < N; i++) {
		cell_xy[2 * i] = particles[i].cell[0];
		cell_xy[2 * i + 1] = particles[i].cell[1];
	}
}

static void shim_copy_out(double *t_cross, uint8_t *dir, double *t_coll,
                          int32_t *partner, uint8_t *ctype)
{
	for (int i = 0; i < N; i++) {
		if (t_cross)
			t_cross[i] = eventList[i]->t;
		if (dir)
			dir[i] = (uint8_t)eventList[i]->j;
		if (t_coll)
			t_coll[i] = eventList[N + i]->t;
		if (partner)
			partner[i] = eventList[N + i]->j;
		if (ctype)
			ctype[i] = (uint8_t)eventList[N + i]->type;
	}
}

static double shim_now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* First sweep, the eventListInit loop (src/EDMD.c:2007-2012).  Must be the
 * first predict after ref_setup (events are not in the calendar yet).
 * Returns seconds spent in the loop. */
double ref_predict_first(int grow, double *t_cross, uint8_t *dir,
                         double *t_coll, int32_t *partner, uint8_t *ctype)
{
	shim_set_mode(grow);
	double t0 = shim_now();
	for (int i = 0; i < N; i++) {
		crossingEvent(i);
		collisionEvent(i);
	}
	double t1 = shim_now();
	shim_copy_out(t_cross, dir, t_coll, partner, ctype);
	return t1 - t0;
}

/* Recurring full re-predict: the reference's own thermostat tick addNoise()
 * (src/EDMD.c:4828-4923).  In the CLI build `noise` is const 0, so the tick
 * takes the velocity-rescale branch `v /= sqrt(E/N/T)` (:4899-4902); E and T
 * are set so that factor is exactly 1 and the tick reduces to
 * freeFly(dt=0) + coll++ + the remove/predict/insert sweep (:4909-4915).
 * Requires a previous ref_predict_first.  Returns seconds. */
double ref_repredict(double *t_cross, uint8_t *dir, double *t_coll,
                     int32_t *partner, uint8_t *ctype)
{
	shim_set_mode(0);
	T = 1.0;
	E = (double)N;
	double t0 = shim_now();
	addNoise();
	double t1 = shim_now();
	removeEventFromQueue(eventList[2 * N + 2]); /* the re-armed NOISE event */
	shim_copy_out(t_cross, dir, t_coll, partner, ctype);
	return t1 - t0;
}

/* Calendar-only cost: remove + re-insert all 2N events unchanged
 * (src/EDMD.c:2144-2221); the host residual after GPU offload. */
double ref_calendar_only(void)
{
	double t0 = shim_now();
	for (int j = 0; j < N; j++) {
		removeEventFromQueue(eventList[j]);
		addEventToQueue(eventList[j]);
		removeEventFromQueue(eventList[N + j]);
		addEventToQueue(eventList[N + j]);
	}
	return shim_now() - t0;
}

/* Batched free flight to t_new (takeAScreenshot loop, src/EDMD.c:4659-4661). */
double ref_free_fly(int grow, double t_new, double *x, double *y, double *rad)
{
	shim_set_mode(grow);
	t = t_new;
	double t0 = shim_now();
	for (int i = 0; i < N; i++)
		freeFly(particles + i);
	double t1 = shim_now();
	for (int i = 0; i < N; i++) {
		if (x)
			x[i] = particles[i].x;
		if (y)
			y[i] = particles[i].y;
		if (rad)
			rad[i] = particles[i].rad;
	}
	return t1 - t0;
}

/* Per-particle local times for the asynchronous form (p->t != t). */
void ref_set_particle_times(const double *tp)
{
	for (int i = 0; i < N; i++)
		particles[i].t = tp[i];
}

/* calculate_pcf (src/pcf.c:16-75).  g_r / r hold at least (int)(max_r/dr)
 * entries; first n_sub particles only when n_sub > 0 (bounded CPU sample with
 * the same box, for timing).  Returns seconds. */
double ref_pcf(double dr, double max_r, int n_sub, double *g_r, double *r,
               int *num_bins)
{
	int n = (n_sub > 0 && n_sub < N) ? n_sub : N;
	double t0 = shim_now();
	pcf_data *d = calculate_pcf(particles, n, dr, max_r, Lx, Ly);
	double t1 = shim_now();
	*num_bins = d->num_bins;
	for (int i = 0; i < d->num_bins; i++) {
		if (g_r)
			g_r[i] = d->g_r[i];
		if (r)
			r[i] = d->r[i];
	}
	free_pcf_data(d);
	return t1 - t0;
}

/* computeBOOPCutoff (src/boop.c:61-107). */
double ref_boop_cutoff(double r_c, double *q5, double *q6, double *q7,
                       double *q6_arg, int32_t *neighbors)
{
	double t0 = shim_now();
	boop_data *b = computeBOOPCutoff(particles, N, r_c, cellList, Nxcells);
	double t1 = shim_now();
	for (int i = 0; i < N; i++) {
		if (q5)
			q5[i] = b[i].q5;
		if (q6)
			q6[i] = b[i].q6;
		if (q7)
			q7[i] = b[i].q7;
		if (q6_arg)
			q6_arg[i] = b[i].q6_arg;
		if (neighbors)
			neighbors[i] = b[i].neighbors;
	}
	free(b);
	return t1 - t0;
}

/* calculate_bond_order_pcf (src/pcf.c:77-167): g(r) and the cos(k.r)-weighted
 * average per bin, for a given wave vector.  Returns seconds. */
double ref_bond_order_pcf(double dr, double max_r, double kx, double ky, double *g_r,
                          double *g6_r, int *num_bins)
{
	double k[2] = {kx, ky};
	double t0 = shim_now();
	bond_order_pcf_data *d = calculate_bond_order_pcf(particles, N, dr, max_r, k, Lx, Ly);
	double t1 = shim_now();
	*num_bins = d->num_bins;
	for (int i = 0; i < d->num_bins; i++) {
		if (g_r)
			g_r[i] = d->g_r[i];
		if (g6_r)
			g6_r[i] = d->g6_r[i];
	}
	free_bond_order_pcf_data(d);
	return t1 - t0;
}

/* find_max_structure_factor_bragg (src/pcf.c:405-467).  Returns seconds. */
double ref_bragg_peak(double expected_bragg, double *k_out)
{
	double t0 = shim_now();
	find_max_structure_factor_bragg(particles, N, Lx, Ly, expected_bragg, k_out);
	return shim_now() - t0;
}

int ref_num_particles(void) { return N; }

/* computeBOOPVoronoi (src/boop.c:15-59): psi_k over the Voronoi neighbours
 * (jc_voronoi on the particles plus the periodic images within 6.0 of the box
 * edges, src/voronoi_edmd.c:33-121).  Returns seconds. */
double ref_boop_voronoi(double *q5, double *q6, double *q7, double *q6_arg,
                        int32_t *neighbors)
{
	double t0 = shim_now();
	boop_data *b = computeBOOPVoronoi(particles, N, Lx, Ly);
	double t1 = shim_now();
	for (int i = 0; i < N; i++) {
		if (q5)
			q5[i] = b[i].q5;
		if (q6)
			q6[i] = b[i].q6;
		if (q7)
			q7[i] = b[i].q7;
		if (q6_arg)
			q6_arg[i] = b[i].q6_arg;
		if (neighbors)
			neighbors[i] = b[i].neighbors;
	}
	free(b);
	return t1 - t0;
}

/* get_particle_voronoi_area / _perimeter (src/voronoi_edmd.c:123-149). */
double ref_voronoi_area(double *area, double *perimeter)
{
	double t0 = shim_now();
	if (area) {
		double *a = get_particle_voronoi_area(particles, N, Lx, Ly);
		memcpy(area, a, (size_t)N * sizeof(double));
		free(a);
	}
	if (perimeter) {
		double *p = get_particle_voronoi_perimeter(particles, N, Lx, Ly);
		memcpy(perimeter, p, (size_t)N * sizeof(double));
		free(p);
	}
	return shim_now() - t0;
}

/* compute_g6_correlation (src/pcf.c:169-230): <Re psi6_i^* psi6_j>(r) with the
 * Voronoi psi6.  Returns seconds. */
double ref_g6_correlation(double dr, double max_r, double *g6_corr, int32_t *counts,
                          int *num_bins)
{
	double t0 = shim_now();
	g6corr_data *d = compute_g6_correlation(particles, N, dr, max_r, Lx, Ly);
	double t1 = shim_now();
	*num_bins = d->num_bins;
	for (int i = 0; i < d->num_bins; i++) {
		if (g6_corr)
			g6_corr[i] = d->g6_corr[i];
		if (counts)
			counts[i] = d->counts[i];
	}
	free(d->r);
	free(d->g6_corr);
	free(d->counts);
	free(d);
	return t1 - t0;
}

/* initStructureFactor's wave-vector grid (src/struc.c:328-345) +
 * computeStructureFactor / computeVelocityStructureFactor (:364-408).
 * s[i*nqy + j]; qx_out / qy_out hold at least 2*qmax/(2 pi/L) + 2 entries.
 * Returns seconds of the compute call. */
double ref_structure_factor(double q_max, int velocity, int *nqx_out, int *nqy_out,
                            double *qx_out, double *qy_out, double *s)
{
	FILE *sink = fopen("/dev/null", "w");
	structFactorComplex = NULL;
	initStructureFactor(q_max, Lx, Ly, N, sink, 0);
	*nqx_out = nqx;
	*nqy_out = nqy;
	for (int i = 0; i < nqx; i++)
		qx_out[i] = qx[i];
	for (int j = 0; j < nqy; j++)
		qy_out[j] = qy[j];
	double t0 = shim_now();
	if (s) {
		if (velocity)
			computeVelocityStructureFactor(particles);
		else
			computeStructureFactor(particles);
	}
	double t1 = shim_now();
	if (s)
		memcpy(s, structFactor, (size_t)nqx * nqy * sizeof(double));
	free(qx);
	free(qy);
	free(structFactor);
	qx = qy = structFactor = NULL;
	fclose(sink);
	return t1 - t0;
}

/* normalizePhysicalQ() (src/EDMD.c:5723-5764) on the loaded system with the reference's option
 * Einit: centre-of-mass velocity removed, E/N set to Einit.  Outputs the new velocities and the
 * reference's globals px, py (before) and E (of the shifted velocities, as its second physicalQ left it). */
void ref_normalize(double e_init, double *vx, double *vy, double *px_before, double *py_before, double *E_shifted)
{
	Einit = e_init;
	physicalQ();
	if (px_before)
		*px_before = px;
	if (py_before)
		*py_before = py;
	normalizePhysicalQ();
	if (E_shifted)
		*E_shifted = E;
	for (int i = 0; i < N; i++) {
		vx[i] = particles[i].vx;
		vy[i] = particles[i].vy;
	}
}

/* A real thermostat tick with the velocity-rescale branch at time t_new:
 * physicalQ() (src/EDMD.c:5968-5997) for E, then addNoise() (:4828-4923): every
 * particle is free-flown to t_new, its velocity divided by sqrt(E/N/T), and the
 * whole system re-predicted.  (In the CLI build `noise` is const 0, which skips
 * the physicalQ() call inside addNoise; it is made here, as the noise == 2 build
 * does at :4830-4832.)  Outputs: E before the tick, the new state, the new events.
 * Requires a previous ref_predict_first.  Returns seconds. */
double ref_tick_rescale(double t_new, double T_target, double *E_out, double *x, double *y,
                        double *vx, double *vy, double *t_cross, uint8_t *dir, double *t_coll,
                        int32_t *partner, uint8_t *ctype)
{
	shim_set_mode(0);
	T = T_target;
	t = t_new;
	double t0 = shim_now();
	physicalQ();
	if (E_out)
		*E_out = E;
	addNoise();
	double t1 = shim_now();
	removeEventFromQueue(eventList[2 * N + 2]); /* the re-armed NOISE event */
	for (int i = 0; i < N; i++) {
		x[i] = particles[i].x;
		y[i] = particles[i].y;
		vx[i] = particles[i].vx;
		vy[i] = particles[i].vy;
	}
	shim_copy_out(t_cross, dir, t_coll, partner, ctype);
	return t1 - t0;
}
