/*
 * rk_mpi_sa_b200.c -- drop-in replacement for the reference's solver module
 * modules/RK_MPI_SAsolver_hybrid2/RK_MPI_SAsolver_hybrid2.c (and its twin
 * modules/RK_MPI_SAsolver_hybrid/RK_MPI_SAsolver_hybrid.c): same six entry points, same return
 * codes, same ownership rules (include/RK_MPI_SAsolver.h of the reference), selected the way
 * the reference selects a solver: `RKSOLVER = RK_MPI_SAsolver_b200` in the application
 * Makefile (apps/intertrack-hybrid-S-freezing/Makefile:15).
 *
 * Host code in C; everything numerical happens behind the C ABI of libpft_b200.so:
 *   RK_MPI_SA_solve = pft_upload_fields(system->x) + pft_solve(...) + pft_download_fields(system->x)
 * (or the pft_group_* forms when this one process drives several GPUs, PFT_NGPUS).  Of system->x only
 * the temperature and phase fields travel per call; the glass field is static (dgl/dt == 0,
 * equation.c:731) and is uploaded once per configuration.  The caller's array is page-locked on first
 * use so that the copies run at PCIe speed.
 * The right-hand side returned by system->meta_f() must be one of the device right-hand sides of
 * the B200 model plug-in (pft_f_device_model01/2); a foreign host function is refused with -2 --
 * there is no CPU path.
 */
#ifdef PFT_USE_REFERENCE_HEADERS
#include "common.h"
#include "RK_MPI_SAsolver.h"       /* the reference's own header: ABI identity by construction */
#else
#include "RK_MPI_SAsolver_b200.h"
#endif
#include "pft_equation_b200.h"

#include <stdio.h>
#include <stdlib.h>

static int max_n = 0;               /* RK_MPI_SAsolver_hybrid2.c:61 */
static int handle_NAN = 0;          /* :63 */
static int last_NAN = 0;            /* :64 */
static MPI_Comm RKcomm;
static int RKrank = 0, RKprocs = 1, RKmaster = 0;
static int gl_generation = -1;      /* configuration whose glass field is resident on the device(s) */
static void *pinned_x = NULL;       /* the caller's solution array, page-locked by this module */

static int env_on(const char *name)
{
	const char *e = getenv(name);
	return e && *e && *e != '0';
}

/* RK_MPI_SAsolver_hybrid2.c:72-120.  No host scratch is allocated: K1,K3,K4 and the stage
 * buffers live in HBM and belong to the pft_solver object. */
int RK_MPI_SA_init(int max_block_size, MPI_Comm comm, int master_rank)
{
	if (MPI_Comm_rank(comm, &RKrank) != MPI_SUCCESS || MPI_Comm_size(comm, &RKprocs) != MPI_SUCCESS) return -4;
	if (max_n) return -3;
	if (max_block_size <= 0) return -2;
	max_n = max_block_size;
	last_NAN = 0;
	RKcomm = comm;
	RKmaster = master_rank;
	return 0;
}

/* :122-136 */
int RK_MPI_SA_cleanup(void)
{
	if (max_n == 0) return -3;
	max_n = 0;
	if (pinned_x) { pft_host_unregister(pinned_x); pinned_x = NULL; }
	gl_generation = -1;
	return 0;
}

/* :138-166 */
void RK_MPI_SA_handle_NAN(int hN) { handle_NAN = (hN == 0) ? 0 : 1; }

/* :168-177 */
int RK_MPI_SA_check_NAN() { return last_NAN; }

/* :179-212 */
int RK_MPI_SA_check_mem(RK_MEM_DIST *n)
{
	int i, end = 0;
	if (max_n == 0) return -3;
	if (n->n_chunks <= 0) return -7;
	for (i = 0; i < n->n_chunks; i++) {
		if (n->chunk_start[i] < end) return -6;             /* out of order or overlapping */
		if (n->chunk_size[i] <= 0) return -6;
		end = n->chunk_start[i] + n->chunk_size[i];
	}
	if (end > max_n) return -5;
	return 0;
}

/* does the chunk table describe exactly the interior x-rows of the three variables of the
 * configured block (apps/intertrack-hybrid-S-freezing/intertrack.c:2149-2156)? */
static int chunks_match_block(const RK_MEM_DIST *n, const pft_case *c)
{
	const int N1 = c->n1 + 2 * PFT_BT, N2 = c->n2 + 2 * PFT_BT, N3 = c->n3 + 2 * PFT_BT;
	const long rowsize = (long)N1 * N2, SIZE = rowsize * N3;
	long i = 0;
	int q, k, j;
	if (n->n_chunks != PFT_NVAR * c->n2 * c->n3) return 0;
	for (q = 0; q < PFT_NVAR; q++)
		for (k = 0; k < c->n3; k++)
			for (j = 0; j < c->n2; j++, i++) {
				if (n->chunk_size[i] != c->n1) return 0;
				if (n->chunk_start[i] != q * SIZE + (k + PFT_BT) * rowsize + (long)(j + PFT_BT) * N1 + PFT_BT) return 0;
				if (n->chunk_eps_mult[i] != 1.0) return 0;   /* the device error norm has no per-chunk weights */
			}
	return 1;
}

typedef struct { RK_MPI_S_SOLUTION *system; } service_ctx;

/* RK_MPI_SAsolver_hybrid2.c:676-685: the callback sees the current t and h in the structure.  It is installed
 * in the master rank only (intertrack.c:2173); the reference re-broadcasts the command word after it (:690)
 * so that a break stops every rank after the same step -- here the master's verdict is broadcast. */
static int service_trampoline(double final_time, double t, double h, long steps, long steps_total, void *ctx)
{
	RK_MPI_S_SOLUTION *system = ((service_ctx *)ctx)->system;
	int stop = 0;
	system->t = t;
	system->h = h;
	system->steps = steps;
	system->steps_total = steps_total;
	if (system->Service_Callback) stop = system->Service_Callback(final_time, system) ? 1 : 0;
	if (RKprocs > 1) MPI_Bcast(&stop, 1, MPI_INT, RKmaster, RKcomm);
	return stop;
}

/* RK_MPI_SAsolver_hybrid2.c:215-770 */
int RK_MPI_SA_solve(FLOAT final_time, RK_MPI_S_SOLUTION *system)
{
	int error_code = 0, errcode_from_others = 0, rc;
	RK_RightHandSide f = NULL;
	pft_solver *solver;
	pft_group *group;
	const pft_case *blk;
	pft_system ps;
	service_ctx sctx;
	unsigned up_mask;
	int have_callback;

	if (system == NULL) return -2;                                  /* :249 */
	if (system->meta_f != NULL) f = system->meta_f();               /* :268 */

	/* :276-286, same priority of codes */
	if (system->n == NULL) error_code = -5;
	else if (system->n->chunk_start[system->n->n_chunks - 1] + system->n->chunk_size[system->n->n_chunks - 1] > max_n) error_code = -5;
	if (max_n == 0) error_code = -3;
	if (system->x == NULL || system->meta_f == NULL) error_code = -2;
	if (RKrank == RKmaster && system->delta <= 0) error_code = -2;

	solver = pft_equation_solver();
	group = pft_equation_group();
	blk = pft_equation_case();
	if (!error_code) {
		if (f == NULL || !pft_equation_is_device_rhs(f)) {
			fprintf(stderr, "RK_MPI_SA_solve (B200): meta_f() returned a host right-hand side; only the device right-hand "
			                "sides of the B200 model plug-in are supported (no CPU path)\n");
			error_code = -2;
		} else if ((solver == NULL && group == NULL) || blk == NULL) {
			fprintf(stderr, "RK_MPI_SA_solve (B200): the model plug-in has not configured a device solver (PrecalculateData)\n");
			error_code = -3;
		} else if (system->DDLBF_Rearrange != NULL) {
			fprintf(stderr, "RK_MPI_SA_solve (B200): DDLBF block rearrangement is not supported\n");
			error_code = -2;
		} else if (!chunks_match_block(system->n, blk)) {
			fprintf(stderr, "RK_MPI_SA_solve (B200): the chunk table is not the interior of the configured %dx%dx%d block\n",
			        blk->n1, blk->n2, blk->n3);
			error_code = -2;
		}
	}

	/* :294-299 */
	MPI_Allreduce(&error_code, &errcode_from_others, 1, MPI_INT, MPI_MIN, RKcomm);
	if (error_code) return error_code;
	if (errcode_from_others < 0) return -6;

	/* :328-336: the master's scalars rule */
	MPI_Bcast(&handle_NAN, 1, MPI_INT, RKmaster, RKcomm);
	MPI_Bcast(&final_time, 1, MPI_DOUBLE, RKmaster, RKcomm);

	ps.t = system->t;
	ps.h = system->h;
	ps.h_min = system->h_min;
	ps.delta = system->delta;
	ps.delta_local = (system->delta_mode == DELTA_LOCAL);
	MPI_Bcast(&ps.t, 1, MPI_DOUBLE, RKmaster, RKcomm);
	MPI_Bcast(&ps.h, 1, MPI_DOUBLE, RKmaster, RKcomm);
	MPI_Bcast(&ps.h_min, 1, MPI_DOUBLE, RKmaster, RKcomm);
	MPI_Bcast(&ps.delta, 1, MPI_DOUBLE, RKmaster, RKcomm);
	MPI_Bcast(&ps.delta_local, 1, MPI_INT, RKmaster, RKcomm);
	ps.steps = system->steps;
	ps.steps_total = system->steps_total;
	ps.handle_nan = handle_NAN;
	ps.last_nan = 0;
	ps.max_attempts = 0;
	ps.rhs_evals = 0;
	sctx.system = system;
	/* every rank has to stop after the same accepted step: if the master has a callback, all ranks take part */
	have_callback = (RKrank == RKmaster && system->Service_Callback) ? 1 : 0;
	MPI_Bcast(&have_callback, 1, MPI_INT, RKmaster, RKcomm);
	ps.service = have_callback ? service_trampoline : NULL;
	ps.service_ctx = &sctx;

	/* page-lock the caller's array once (pageable memory halves the PCIe rate); failure is harmless */
	if (pinned_x != (void *)system->x && !env_on("PFT_NO_PIN")) {
		const int last = system->n->n_chunks - 1;
		const unsigned long long bytes = sizeof(FLOAT) * (unsigned long long)max_n;
		(void)last;
		if (pinned_x) { pft_host_unregister(pinned_x); pinned_x = NULL; }
		if (pft_host_register(system->x, bytes) == 0) pinned_x = (void *)system->x;
	}

	/* the glass field is static: it goes up once per configuration of the plug-in */
	up_mask = PFT_FIELD_U | PFT_FIELD_P;
	if (gl_generation != pft_equation_generation() || env_on("PFT_UPLOAD_GL_ALWAYS")) up_mask |= PFT_FIELD_GL;
	rc = group ? pft_group_upload(group, system->x, up_mask) : pft_upload_fields(solver, system->x, up_mask);
	if (rc) { fprintf(stderr, "RK_MPI_SA_solve (B200): %s\n", pft_last_error()); return -6; }
	gl_generation = pft_equation_generation();

	rc = group ? pft_group_solve(group, &ps, final_time) : pft_solve(solver, &ps, final_time);
	if (rc < 0 && rc != PFT_SOLVE_NAN && rc != PFT_SOLVE_BADSYSTEM) {
		fprintf(stderr, "RK_MPI_SA_solve (B200): %s\n", pft_last_error());
		return -6;
	}

	last_NAN = ps.last_nan;
	system->steps = ps.steps;
	system->steps_total = ps.steps_total;
	system->t = ps.t;                                                /* :768 / :700 / :634 */
	/* system->h follows the reference: written at NEXTFINISH (:754), on a callback break (:701)
	   and before every callback (:679); otherwise left alone */
	system->h = ps.h;

	if (group ? pft_group_download(group, system->x, PFT_FIELD_U | PFT_FIELD_P) : pft_download_fields(solver, system->x, PFT_FIELD_U | PFT_FIELD_P)) {
		fprintf(stderr, "RK_MPI_SA_solve (B200): %s\n", pft_last_error());
		return -6;
	}
	return rc;
}
