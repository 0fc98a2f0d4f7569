/*
 * RK_MPI_SAsolver_b200.h -- the solver-module ABI of the reference, as implemented by the
 * B200 drop-in module (porousfreezethaw_b200/csrc/rk_mpi_sa_b200.c).
 *
 * This header mirrors, declaration for declaration, the interface the reference selects with
 * `RKSOLVER = ...` in apps/intertrack-hybrid-S-freezing/Makefile:15, i.e.
 * include/RK_MPI_SAsolver.h of the reference:
 *     RKA_CMD_*            :30-35        RK_RightHandSide     :173
 *     RK_MEM_DIST          :184-194      RK_MPI_S_SOLUTION    :196-289
 *     RK_MPI_SA_init ... RK_MPI_SA_solve prototypes            :291-392
 * Struct layouts are identical (tests/test_abi.py compares sizeof/offsetof against the
 * reference header where /root/reference is present), so an application compiled against the
 * reference's own header links against the B200 module unchanged.  When the B200 module is
 * built INTO the reference tree it includes the reference's header instead of this one
 * (-DPFT_USE_REFERENCE_HEADERS).
 */
#ifndef RK_MPI_SASOLVER_B200_H
#define RK_MPI_SASOLVER_B200_H

#ifdef PFT_HAVE_MPI
#include <mpi.h>
#else
#include "mpi_single/mpi.h"   /* single-rank token types when no MPI is installed */
#endif

#ifndef PFT_FLOAT_DEFINED
#define PFT_FLOAT_DEFINED
typedef double FLOAT;   /* include/common.h:26,64 of the reference: _DEFAULT_FP_PRECISION == FP_DOUBLE */
#endif

/* commands of the master rank (bit flags) */
#define RKA_CMD_h_TOO_SMALL   1
#define RKA_CMD_NAN           2
#define RKA_CMD_UPDATE        4
#define RKA_CMD_FINISHED      8
#define RKA_CMD_NEXTFINISH    16
#define RKA_CMD_BREAK         32

#ifdef __cplusplus
extern "C" {
#endif

/* dx/dt = f(t, x): f(t, x, dx) */
typedef void (*RK_RightHandSide)(FLOAT, const FLOAT *, FLOAT *);

/* the "chunks" of the array x that the solver integrates (everything else is ghost shell) */
typedef struct {
	int n_chunks;
	int *chunk_start;            /* offsets into x, ascending, non-overlapping */
	int *chunk_size;
	FLOAT *chunk_eps_mult;       /* per-chunk multiplier of the error estimate */
} RK_MEM_DIST;

typedef struct __struct_RK_MPI_S_SOLUTION {
	RK_MEM_DIST *n;
	FLOAT t;                     /* current time level */
	FLOAT *x;                    /* the solution, updated in place */
	RK_RightHandSide (*meta_f)();/* returns the right-hand side to use */
	FLOAT h;                     /* current time step */
	FLOAT h_min;                 /* steps shorter than this are always accepted */
	FLOAT delta;                 /* tolerance */
	enum { DELTA_LOCAL, DELTA_GLOBAL } delta_mode;
	RK_MEM_DIST *(*DDLBF_Rearrange)(RK_MEM_DIST *);
	int (*Service_Callback)(FLOAT, const struct __struct_RK_MPI_S_SOLUTION *const);
	long steps;                  /* accepted steps */
	long steps_total;            /* step attempts */
} RK_MPI_S_SOLUTION;

/* 0 ok, -1 no memory, -2 bad dimension, -3 already initialised, -4 MPI not initialised */
int RK_MPI_SA_init(int max_block_size, MPI_Comm comm, int master_rank);
/* 0 ok, -3 not initialised */
int RK_MPI_SA_cleanup(void);
void RK_MPI_SA_handle_NAN(int hN);
int RK_MPI_SA_check_NAN();
/* 0 ok, -3 not initialised, -5 chunks exceed max_block_size, -6 bad chunk order/size, -7 no chunks */
int RK_MPI_SA_check_mem(RK_MEM_DIST *n);
/* 0 done, 1 interrupted by Service_Callback, -2 bad system, -3 not initialised, -4 NaN abort,
 * -5 chunks exceed memory, -6 error on another rank */
int RK_MPI_SA_solve(FLOAT final_time, RK_MPI_S_SOLUTION *system);

#ifdef __cplusplus
}
#endif
#endif /* RK_MPI_SASOLVER_B200_H */
