/*
 * oracle/ref_harness.c -- ctypes-callable harness around the UNMODIFIED reference.
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from this file.
 *
 * The reference sources are compiled where they lie under /root/reference (see
 * oracle/Makefile, target _ref/libpft_ref.so); nothing is copied into this repo.
 * This translation unit textually includes the reference driver
 * (apps/intertrack-hybrid-S-freezing/intertrack.c, which itself includes model.c
 * and equation.c) with its main() renamed, so that the reference's file-static
 * globals (n1..n3, N1..N3, param[], solution, calc_mode, ...) can be set the same
 * way main() sets them, and then calls the reference's own PrecalculateData(),
 * RHS meta-pointer and RK_MPI_SA_solve().
 *
 * What each entry point mirrors in the reference driver:
 *   pftref_setup : intertrack.c:1776-1800 (block geometry), :1809-1822 (allocation),
 *                  :2149-2160 (chunk table), :2163-2192 (eqSystem, RK_MPI_SA_init),
 *                  :2208 (check_mem), :2218 (PrecalculateData)
 *   pftref_rhs   : what RK_MPI_SA_solve does for K1 (RK_MPI_SAsolver_hybrid2.c:351,373):
 *                  f(t,x,K) called by all threads of one parallel region
 *   pftref_solve : intertrack.c:2283
 */
#define main pftref_unused_reference_main
#include "intertrack.c"
#undef main

static int              H_ready = 0;
static int             *H_chunk_start = NULL, *H_chunk_size = NULL;
static FLOAT           *H_chunk_eps_mult = NULL;
static RK_MEM_DIST      H_dist;
static RK_MPI_S_SOLUTION H_sys;
static char             H_spheres_path[4096];

/* accepted-step recorder (Service_Callback) */
static double *H_rec = NULL;
static long    H_rec_n = 0, H_rec_cap = 0;
static long    H_break_after = -1;   /* break (return 1) once steps_total >= this; <0: never */

static int H_service(FLOAT final_time, const RK_MPI_S_SOLUTION * const self)
{
	(void)final_time;
	if (H_rec_n + 1 > H_rec_cap) {
		long cap = H_rec_cap ? 2 * H_rec_cap : 4096;
		double *p = (double *)realloc(H_rec, (size_t)cap * 2 * sizeof(double));
		if (p) { H_rec = p; H_rec_cap = cap; }
	}
	if (H_rec_n < H_rec_cap) { H_rec[2 * H_rec_n] = self->t; H_rec[2 * H_rec_n + 1] = self->h; H_rec_n++; }
	if (H_break_after >= 0 && self->steps_total >= H_break_after) return 1;
	return 0;
}

int pftref_cleanup(void)
{
	if (!H_ready) return 0;
	RK_MPI_SA_cleanup();
	FreePrecalcData();
	free(solution); solution = NULL;
	free(H_chunk_start); free(H_chunk_size); free(H_chunk_eps_mult);
	H_chunk_start = H_chunk_size = NULL; H_chunk_eps_mult = NULL;
	free(MPIrankmap); MPIrankmap = NULL;
	free(H_rec); H_rec = NULL; H_rec_n = H_rec_cap = 0;
	H_ready = 0;
	return 0;
}

/*
 * u0,p0,gl0: interior initial fields, C order [n3][n2][n1] (what the formula pass at
 * intertrack.c:1934-1968 produces).  The sphere pack is max-ed into gl by the
 * reference's own PrecalculateData (equation.c:508-529) afterwards.
 */
int pftref_setup(int n1_, int n2_, int total_n3_, double L1_, double L2_, double L3_,
                 const double *params30, int calc_mode_, const char *spheres_path,
                 const double *u0, const double *p0, const double *gl0,
                 double t0, double tau, double tau_min, double delta, int use_callback)
{
	int q, i, j, k, c, rc;
	const double *ic[VAR_COUNT];

	if (H_ready) pftref_cleanup();

	MPIrank = 0; MPIprocs = 1; MPImaster = 0;
	MPIrankmap = (int *)malloc(sizeof(int)); MPIrankmap[0] = 0;
	logfile = NULL;

	L1 = L1_; L2 = L2_; L3 = L3_;
	n1 = n1_; n2 = n2_; total_n3 = total_n3_;
	calc_mode = calc_mode_;
	for (q = 0; q < PARAM_COUNT; q++) model_parameters[q] = params30[q];

	/* intertrack.c:1776-1800 with MPIprocs == 1 */
	N1 = n1 + 2 * bcond_thickness;
	N2 = n2 + 2 * bcond_thickness;
	total_N3 = total_n3 + 2 * bcond_thickness;
	n3 = total_n3; first_row = 0;
	N3 = n3 + 2 * bcond_thickness;
	rowsize = N1 * N2;
	bcond_size = bcond_thickness * rowsize;
	subgridSIZE = rowsize * N3;
	subgridSize = rowsize * n3;
	subgridsize = n1 * n2 * n3;

	c = VAR_COUNT * n2 * n3;
	H_chunk_start = (int *)malloc((size_t)c * sizeof(int));
	H_chunk_size = (int *)malloc((size_t)c * sizeof(int));
	H_chunk_eps_mult = (FLOAT *)malloc((size_t)c * sizeof(FLOAT));
	solution = (FLOAT *)calloc((size_t)VAR_COUNT * subgridSIZE, sizeof(FLOAT));
	if (!H_chunk_start || !H_chunk_size || !H_chunk_eps_mult || !solution) return -101;
	if (AllocPrecalcData()) return -102;

	ic[temperature_field] = u0; ic[phase_field] = p0; ic[glass_field] = gl0;
	for (q = 0; q < VAR_COUNT; q++)
		for (k = 0; k < n3; k++)
			for (j = 0; j < n2; j++)
				for (i = 0; i < n1; i++)
					VAR(solution, q)[(k + bcond_thickness) * rowsize + (j + bcond_thickness) * N1 + bcond_thickness + i]
						= ic[q][((size_t)k * n2 + j) * n1 + i];

	/* intertrack.c:2149-2160 */
	c = 0;
	for (q = 0; q < VAR_COUNT; q++)
		for (k = 0; k < n3; k++)
			for (j = 0; j < n2; j++) {
				H_chunk_start[c] = q * subgridSIZE + (k + bcond_thickness) * rowsize + (j + bcond_thickness) * N1 + bcond_thickness;
				H_chunk_size[c] = n1;
				H_chunk_eps_mult[c] = 1.0;
				c++;
			}
	H_dist.n_chunks = c; H_dist.chunk_start = H_chunk_start; H_dist.chunk_size = H_chunk_size; H_dist.chunk_eps_mult = H_chunk_eps_mult;

	memset(&H_sys, 0, sizeof H_sys);
	H_sys.n = &H_dist;
	H_sys.t = t0;
	H_sys.x = solution;
	H_sys.meta_f = mf_single;
	H_sys.h = tau;
	H_sys.h_min = tau_min;
	H_sys.delta = delta;
	H_sys.delta_mode = DELTA_GLOBAL;
	H_sys.DDLBF_Rearrange = NULL;
	H_sys.Service_Callback = use_callback ? H_service : NULL;
	H_sys.steps = 0L; H_sys.steps_total = 0L;
	H_rec_n = 0; H_break_after = -1;

	rc = RK_MPI_SA_init(VAR_COUNT * subgridSIZE, MPI_COMM_WORLD, MPImaster);
	if (rc) return rc;
	rc = RK_MPI_SA_check_mem(&H_dist);
	if (rc) return rc;

	strncpy(H_spheres_path, spheres_path ? spheres_path : "", sizeof H_spheres_path - 1);
	ball_positions_file = H_spheres_path;
	if (PrecalculateData(H_chunk_eps_mult)) return -103;

	H_ready = 1;
	return 0;
}

/* out: N1,N2,N3,n1,n2,n3,subgridSIZE,rowsize */
int pftref_dims(int *out)
{
	if (!H_ready) return -1;
	out[0] = N1; out[1] = N2; out[2] = N3; out[3] = n1; out[4] = n2; out[5] = n3; out[6] = subgridSIZE; out[7] = rowsize;
	return 0;
}

double *pftref_solution(void) { return H_ready ? solution : NULL; }

/* dw = f(t, w) with the reference's RHS for the current calc_mode; w == NULL -> the solution array.
 * NOTE the reference RHS writes the ghost shell of w (bcond_setup, equation.c:624). */
int pftref_rhs(double t, double *w, double *dw)
{
	RK_RightHandSide f;
	if (!H_ready) return -1;
	f = mf_single();
	if (!w) w = solution;
	#pragma omp parallel default(shared)
	{
		f(t, w, dw);
	}
	return 0;
}

int pftref_bcond(double t, double *w)
{
	if (!H_ready) return -1;
	if (!w) w = solution;
	#pragma omp parallel default(shared)
	{
		bcond_setup(t, w);
	}
	return 0;
}

int pftref_solve(double final_time_)
{
	if (!H_ready) return -1000;
	return RK_MPI_SA_solve(final_time_, &H_sys);
}

/* out: t, h, steps, steps_total */
int pftref_state(double *out)
{
	if (!H_ready) return -1;
	out[0] = H_sys.t; out[1] = H_sys.h; out[2] = (double)H_sys.steps; out[3] = (double)H_sys.steps_total;
	return 0;
}

int pftref_set_state(double t, double h, double h_min, double delta)
{
	if (!H_ready) return -1;
	H_sys.t = t; H_sys.h = h; H_sys.h_min = h_min; H_sys.delta = delta;
	return 0;
}

void pftref_handle_nan(int on) { RK_MPI_SA_handle_NAN(on); }
void pftref_break_after(long attempts) { H_break_after = attempts; }
long pftref_record_count(void) { return H_rec_n; }
void pftref_record_clear(void) { H_rec_n = 0; }
long pftref_record_get(double *out, long max_pairs)
{
	long n = H_rec_n < max_pairs ? H_rec_n : max_pairs;
	if (n > 0) memcpy(out, H_rec, (size_t)n * 2 * sizeof(double));
	return n;
}
