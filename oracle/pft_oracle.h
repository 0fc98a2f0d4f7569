/*
 * oracle/pft_oracle.h -- CPU restatement of the PorousFreezeThaw hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load libpft_oracle.so; the product
 * (porousfreezethaw_b200/csrc) never links, loads or calls it.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py checks every function below
 * bit-for-bit against the unmodified reference compiled in oracle/_ref/ (and against
 * the golden vectors in tests/golden/ that were generated from that reference build).
 *
 * Memory layout = the reference's (intertrack.c:431,1776-1800):
 *   N_d = n_d + 2*PFO_BT, rowsize = N1*N2, subgridSIZE = rowsize*N3,
 *   node (i,j,k) of variable q at w[q*subgridSIZE + (k+2)*rowsize + (j+2)*N1 + (i+2)].
 */
#ifndef PFT_ORACLE_H
#define PFT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define PFO_BT          2      /* bcond_thickness, equation.c:38 */
#define PFO_NVAR        3      /* u, p, gl -- model.c:32-40 */
#define PFO_NPARAM      30     /* model.c:44-59 */
#define PFO_MAX_BALLS   1000   /* equation.c:34 (lifted by pfo_rasterise's explicit count) */

/* parameter indices, same order as the enum at model.c:44-59 */
enum {
	PFO_u_star, PFO_L, PFO_xi, PFO_a, PFO_b, PFO_alpha, PFO_mu,
	PFO_beads_scaling, PFO_beads_offset_x, PFO_beads_offset_y, PFO_beads_offset_z,
	PFO_xi_gl, PFO_zeta,
	PFO_p_eps0, PFO_p_eps1,
	PFO_gamma,
	PFO_water_cp, PFO_ice_cp, PFO_glass_cp,
	PFO_water_lambda, PFO_ice_lambda, PFO_glass_lambda,
	PFO_water_rho, PFO_ice_rho, PFO_glass_rho,
	PFO_top_temp1, PFO_top_temp2, PFO_phase_switch_time,
	PFO_u_noise_amp,
	PFO_ball_radius
};

typedef struct {
	int n1, n2, n3;            /* interior nodes of THIS block (n3 = local slab depth) */
	int total_n3;              /* interior nodes of the whole grid along z */
	int first_row;             /* global z index of local row 0 (intertrack.c:1780-1787) */
	int rank, nranks;          /* slab position: rank 0 owns the z-front, nranks-1 the z-rear */
	int calc_mode;             /* 0,1,2,10,11 */
	double L1, L2, L3;
	double param[PFO_NPARAM];
} pfo_case;

/* z-halo exchange hook (the role of sync_solution, equation.c:290-326): called once per RHS
 * evaluation on every rank with nranks > 1, after the physical boundary conditions are set. */
typedef void (*pfo_halo_fn)(double *w, void *ctx);
/* global max hook (the role of MPI_Allreduce(MAX), RK_MPI_SAsolver_hybrid2.c:572) */
typedef double (*pfo_allmax_fn)(double local_eps, void *ctx);
/* service callback (RK_MPI_SAsolver.h:266-279): nonzero return interrupts the solve */
typedef int (*pfo_service_fn)(double final_time, double t, double h, long steps, long steps_total, void *ctx);

typedef struct {
	double t, h, h_min, delta;
	int    delta_local;        /* 1 = DELTA_LOCAL, 0 = DELTA_GLOBAL (the app uses GLOBAL, intertrack.c:2171) */
	long   steps, steps_total;
	pfo_halo_fn    halo;    void *halo_ctx;
	pfo_allmax_fn  allmax;  void *allmax_ctx;
	pfo_service_fn service; void *service_ctx;
	int    handle_nan, last_nan;
	long   max_attempts;       /* test/bench extension: stop after this many attempts (<=0: unlimited);
	                              returns 2 and leaves (t,h) resumable.  Not in the reference. */
	long   rhs_evals;          /* number of RHS evaluations performed (5 per attempt in the reference) */
} pfo_system;

/* sizes: out = {N1, N2, N3, rowsize, bcond_size, subgridSIZE, n_chunks, 0} */
void pfo_sizes(const pfo_case *c, int *out);

/* slab partition of total_n3 rows over nranks (intertrack.c:1780-1787) */
void pfo_partition(int total_n3, int nranks, int rank, int *n3, int *first_row);

/* sphere pack: read up to max_count "x y z" rows (equation.c:476-484); returns count or -1 */
int  pfo_read_spheres(const char *path, double *xyz, int max_count);
/* gl := max(gl, 0.5*(1 - tanh(0.5/xi_gl*(|x-c|+1e-10 - R)))) over all spheres (equation.c:479-482,508-529) */
void pfo_rasterise(const pfo_case *c, const double *xyz_raw, int count, double *w);

/* ghost shell: equation.c:96-284 (x/y mirror, z-front mirror on rank 0, z-rear Dirichlet(u)/mirror(p,gl) on last rank) */
void pfo_bcond(const pfo_case *c, double t, double *w);

/* right-hand side for the case's calc_mode: equation.c:566-741 (modes 0,1,10,11) / :745-884 (mode 2).
 * Like the reference it rewrites the ghost shell of w. halo may be NULL when nranks == 1. */
void pfo_rhs(const pfo_case *c, double t, double *w, double *dw, pfo_halo_fn halo, void *halo_ctx);

/* Runge-Kutta-Merson adaptive integration to final_time: RK_MPI_SAsolver_hybrid2.c:215-770.
 * work: 5 arrays of 3*subgridSIZE doubles (K1,K3,K4,K5,aux) laid out back to back.
 * returns 0 done, 1 interrupted by service callback, 2 attempt budget exhausted, -2 bad arguments, -4 NaN abort */
int  pfo_solve(const pfo_case *c, pfo_system *s, double *x, double *work, double final_time);

/* mean(p > 0.5) over the interior: what scripts/avg.sh:9-20 computes with NCO */
double pfo_ice_fraction(const pfo_case *c, const double *w);

/* out[i] = pow(x[i], y) with the C library's pow (what RK_MPI_SAsolver_hybrid2.c:580 calls) */
void pfo_pow_array(const double *x, double y, double *out, long n);

#ifdef __cplusplus
}
#endif
#endif
