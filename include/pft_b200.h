/*
 * pft_b200.h -- C ABI of libpft_b200.so: the B200-native (sm_100a, float64) implementation of
 * PorousFreezeThaw's hot path, i.e. the Runge-Kutta-Merson adaptive integrator driving the
 * phase-field / heat right-hand side on a z-slab of the grid that lives in this GPU's HBM.
 *
 * Plain C, no torch / C++ types.  Every entry point returns 0 on success or a negative
 * PFT_E* code (pft_last_error() gives the text); nothing throws across the boundary.  All
 * host pointers are caller-owned.  One pft_solver is driven by one host thread.
 *
 * There is NO CPU fallback: pft_create fails with PFT_ENODEVICE when no CUDA device is usable.
 *
 * What each entry point replaces in the reference (paths relative to the reference root):
 *
 *   pft_create / pft_destroy   RK_MPI_SA_init / RK_MPI_SA_cleanup (scratch K1,K3,K4,K5,aux:
 *                              modules/RK_MPI_SAsolver_hybrid2/RK_MPI_SAsolver_hybrid2.c:72-136)
 *                              + the block geometry of apps/intertrack-hybrid-S-freezing/
 *                              intertrack.c:1776-1800 + the derived constants of
 *                              equation.c:442-447,605-612
 *   pft_upload / pft_download  the in-place host array `system->x` of RK_MPI_S_SOLUTION
 *                              (include/RK_MPI_SAsolver.h:196-289) in the padded layout of
 *                              intertrack.c:431,2149-2156
 *   pft_rhs_host               RK_RightHandSide f(t, w, dw) = f_generic_model01 / f_generic_model2
 *                              (equation.c:566-741, :745-884) incl. bcond_setup (:266-284)
 *   pft_solve                  RK_MPI_SA_solve (RK_MPI_SAsolver_hybrid2.c:215-770; twin
 *                              modules/RK_MPI_SAsolver_hybrid/RK_MPI_SAsolver_hybrid.c)
 *   pft_group_*                the same for one host process that owns several GPUs: the z-slab partition of
 *                              intertrack.c:1776-1800 applied inside the solver module
 *   pft_comm_*                 the z-halo exchange sync_solution (equation.c:290-326) and
 *                              MPI_Allreduce(MAX) (RK_MPI_SAsolver_hybrid2.c:572)
 *   pft_rasterise              the sphere rasteriser of PrecalculateData (equation.c:474-529)
 *   pft_ice_fraction           scripts/avg.sh:9-20 (mean(p > 0.5))
 */
#ifndef PFT_B200_H
#define PFT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PFT_NPARAM 30            /* apps/intertrack-hybrid-S-freezing/model.c:44-59 */
#define PFT_NVAR   3             /* u, p, gl -- model.c:32-40 */
#define PFT_BT     2             /* bcond_thickness of the HOST layout, equation.c:38 */

/* error codes (pft_solve additionally returns the reference's own 0 / 1 / -2 / -4 codes) */
#define PFT_OK           0
#define PFT_EINVAL     (-22)
#define PFT_ENOMEM     (-12)
#define PFT_ENODEVICE  (-19)
#define PFT_ECUDA      (-5)
#define PFT_ENCCL      (-71)
#define PFT_ESTATE     (-1)

/* pft_solve return codes */
#define PFT_SOLVE_DONE         0   /* RK_MPI_SA_solve: 0 */
#define PFT_SOLVE_INTERRUPTED  1   /* RK_MPI_SA_solve: 1 (service callback asked to stop) */
#define PFT_SOLVE_BUDGET       2   /* extension: max_attempts reached, state resumable */
#define PFT_SOLVE_BADSYSTEM  (-2)  /* RK_MPI_SA_solve: -2 */
#define PFT_SOLVE_NAN        (-4)  /* RK_MPI_SA_solve: -4 */

/* flags for pft_create */
#define PFT_FLAG_NO_RUNAHEAD      2u   /* enqueue one attempt at a time and wait for its outcome (default: the host
                                          enqueues attempts ahead of the device, which decides by itself) */
#define PFT_FLAG_RECOMPUTE_K1     4u   /* recompute K1 after a rejected attempt, like the reference */
#define PFT_FLAG_LEGACY_LOADS     8u   /* per-thread global loads instead of TMA tile staging (debug / A-B) */
#define PFT_FLAG_HOST_POW        32u   /* take pow(delta/eps, 0.2) of the step-size controller from the host C library
                                          once per attempt instead of the device twin of it (checker for the twin) */

/* field masks for pft_upload_fields / pft_download_fields */
#define PFT_FIELD_U   1u
#define PFT_FIELD_P   2u
#define PFT_FIELD_GL  4u
#define PFT_FIELD_ALL 7u

typedef struct pft_solver pft_solver;
typedef struct pft_group pft_group;      /* several z-slabs of one grid on several GPUs, driven by ONE host process */

/* one block (z-slab) of the grid; field meaning as in intertrack.c */
typedef struct {
	int n1, n2, n3;          /* interior nodes of this block (n3 = slab depth) */
	int total_n3;            /* interior nodes of the whole grid along z */
	int first_row;           /* global z index of local row 0 (intertrack.c:1780-1787) */
	int rank, nranks;        /* slab position; rank 0 owns the z-front, nranks-1 the z-rear */
	int calc_mode;           /* 0 GradP, 1 SigmaP1-P, 2 temperature model, 10/11 = 0/1 without heat equation */
	double L1, L2, L3;       /* domain size */
	double param[PFT_NPARAM];/* model_parameters[], model.c:44-59 order */
	int max_n3;              /* deepest slab of the run; 0 = the reference's partition rule (total_n3/nranks, remainder to
	                            the lowest ranks).  Results do not depend on how the planes are dealt out, so a caller may
	                            give slower slabs fewer planes (pft_group_* does that by itself); every rank passes the
	                            same max_n3 >= its n3. */
} pft_case;

/* service callback: RK_MPI_S_SOLUTION.Service_Callback (RK_MPI_SAsolver.h:266-279) */
typedef int (*pft_service_fn)(double final_time, double t, double h, long steps, long steps_total, void *ctx);

/* integration state: the scalar members of RK_MPI_S_SOLUTION (RK_MPI_SAsolver.h:196-289) */
typedef struct {
	double t;                /* in/out: current time level */
	double h;                /* in/out: time step (same hand-over rules as the reference) */
	double h_min;            /* steps shorter than this are always accepted */
	double delta;            /* error tolerance */
	int    delta_local;      /* 1 = DELTA_LOCAL, 0 = DELTA_GLOBAL */
	long   steps;            /* += accepted steps */
	long   steps_total;      /* += step attempts */
	pft_service_fn service;  /* may be NULL */
	void  *service_ctx;
	int    handle_nan;       /* RK_MPI_SA_handle_NAN */
	int    last_nan;         /* RK_MPI_SA_check_NAN */
	long   max_attempts;     /* extension: stop after this many attempts (<= 0: unlimited) */
	long   rhs_evals;        /* += right-hand-side evaluations actually performed on the device */
} pft_system;

typedef struct {
	long   kernel_launches;  /* kernels of this library launched so far */
	long   rhs_evals;
	long   attempts, accepted;
	double device_ms_total;  /* device time of the last pft_solve (CUDA events on the solver stream) */
	double stage_ms[5];      /* device time per Merson stage kernel, summed since pft_reset_stats like
	                            stage_launches (only collected when PFT_PROFILE_STAGES=1) */
	long   stage_launches[5];
	int    tile_x, tile_y, z_chunk, threads, grid_blocks;
	long   bytes_per_field;  /* device bytes of one field array incl. ghost planes */
	double host_wait_ms;     /* host time spent waiting (asleep, not spinning) for the device to free a slot of
	                            the launch queue */
	double host_launch_ms;   /* host time spent issuing launches */
	int    device_pow;       /* 1: the step-size controller runs on the device with the twin of the host library's
	                            pow(); 0: the host supplies pow() once per attempt (twin unavailable or PFT_FLAG_HOST_POW) */
	int    runahead;         /* attempts the host keeps enqueued ahead of the device */
	long   noop_launches;    /* stage launches that found the integration already finished (run-ahead tail) */
	int    kernel_generation;/* of the dominant stage kernel: 0 merson_stage, 1 merson_stage_tma, 2 merson_stage_v2 */
	int    rows_per_thread;  /* cells per thread of that kernel */
	int    speculative_stage1; /* 1: attempts run as [S2 S3 S4 S5 S1'] -- K1 of the next attempt is evaluated from stage 5's
	                            candidate while the step decision is still being taken (PFT_SPEC=0 turns it off) */
	int    reserved_;
} pft_stats;

const char *pft_last_error(void);
const char *pft_version(void);

int pft_create(const pft_case *c, int device, unsigned flags, pft_solver **out);
int pft_destroy(pft_solver *s);

/* host array in the reference's padded layout: 3 * N1*N2*N3 doubles, N_d = n_d + 2*PFT_BT */
int pft_upload(pft_solver *s, const double *x_padded);
int pft_download(pft_solver *s, double *x_padded);          /* writes interior nodes only */
/* the same for a subset of the variables (PFT_FIELD_*): the glass field is static (dgl/dt == 0, equation.c:731),
 * so a caller that keeps one host array across solve calls uploads it once and never downloads it */
int pft_upload_fields(pft_solver *s, const double *x_padded, unsigned mask);
int pft_download_fields(pft_solver *s, double *x_padded, unsigned mask);
/* page-lock a caller-owned host array so that the copies above run at full PCIe speed (optional) */
int pft_host_register(void *ptr, unsigned long long bytes);
int pft_host_unregister(void *ptr);
/* compact interior arrays [n3][n2][n1]; any pointer may be NULL to skip that field */
int pft_upload_interior(pft_solver *s, const double *u, const double *p, const double *gl);
int pft_download_interior(pft_solver *s, double *u, double *p, double *gl);

/* gl := max(gl, 0.5*(1 - tanh(0.5/xi_gl*(|x-c|+1e-10 - R)))) on the device (device tanh: <= 2 ulp from libm) */
int pft_rasterise(pft_solver *s, const double *xyz_raw, int count);

/* dw = f(t, w): upload w, one device RHS evaluation, download; both in the padded layout */
int pft_rhs_host(pft_solver *s, double t, const double *w_padded, double *dw_padded);

/* integrate the device-resident state to final_time; see RK_MPI_SA_solve for the semantics */
int pft_solve(pft_solver *s, pft_system *sys, double final_time);

/* mean(p > 0.5) over this block's interior, computed on the device */
int pft_ice_fraction(pft_solver *s, double *fraction);

/* multi-GPU (one process per GPU): create a 128-byte communicator id on one rank, distribute it
 * by any means (MPI_Bcast / torch.distributed), then call pft_comm_init on every rank. */
int pft_comm_unique_id(void *id128);
int pft_comm_init(pft_solver *s, const void *id128);

/* multi-GPU, ONE host process driving ngpus devices (one worker thread per device inside pft_group_solve):
 * what the reference does with M MPI ranks inside one application run (intertrack.c:1776-1800 slab partition,
 * equation.c:290-326 halo exchange, RK_MPI_SAsolver_hybrid2.c:572 error maximum).  `whole` describes the WHOLE
 * block this process owns (rank 0 of 1); it is cut into ngpus z-slabs by the rule of intertrack.c:1780-1787.
 * The slabs reach each other through peer access (NVLink), no NCCL.  x_padded is the one padded host array of
 * the whole block; the service callback is invoked by slab 0's worker only. */
int pft_group_create(const pft_case *whole, int ngpus, const int *devices, unsigned flags, pft_group **out);
/* the same with a given number of planes per slab (sum == whole->n3, every slab >= 2 planes); planes == NULL: the rule above */
int pft_group_create_planes(const pft_case *whole, int ngpus, const int *devices, const int *planes, unsigned flags, pft_group **out);
/* Load balance: slabs differ in cost per plane (interior slabs push two halo planes, the division's cost depends on the data),
 * and every attempt ends with all slabs waiting for the slowest.  pft_group_balance runs `attempts` attempts with per-stage
 * device timing, and if the slowest slab is more than 3 % above the mean it deals the planes out again in inverse proportion
 * to the measured cost per plane (state carried over through a host copy).  Returns 1 if the partition changed, 0 if not.
 * pft_group_solve calls it once by itself after the first 30 attempts unless PFT_BALANCE=0.  The planes in use: */
int pft_group_balance(pft_group *g, pft_system *sys, double final_time, int attempts);
int pft_group_planes(const pft_group *g, int *planes);
int pft_group_destroy(pft_group *g);
int pft_group_size(const pft_group *g);
pft_solver *pft_group_member(pft_group *g, int r);
int pft_group_upload(pft_group *g, const double *x_padded, unsigned mask);
int pft_group_download(pft_group *g, double *x_padded, unsigned mask);
int pft_group_solve(pft_group *g, pft_system *sys, double final_time);

/* test hook: n values x[i]^y of the DEVICE instance of the step-size controller's pow twin (pft_pow.h) */
int pft_debug_pow(pft_solver *s, const double *x, double y, double *out, int n);

int pft_get_stats(pft_solver *s, pft_stats *out);
int pft_reset_stats(pft_solver *s);

/* debug/test hook: copy an internal array pair (u,p components) to compact host arrays.
 * which: 0 x, 1 K1, 2 aux3 (stage-3 input), 3 K3, 4 K4, 5 x_new (last attempt's candidate) */
int pft_debug_fetch(pft_solver *s, int which, double *u, double *p);

#ifdef __cplusplus
}
#endif
#endif /* PFT_B200_H */
