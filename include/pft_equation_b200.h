/*
 * pft_equation_b200.h -- the model plug-in seam of the reference, B200 side.
 *
 * In the reference the model (boundary conditions, sphere rasteriser, right-hand sides,
 * RHS meta-pointers) is the file apps/intertrack-hybrid-S-freezing/equation.c, textually
 * included by the driver (intertrack.c:633).  The names the driver uses from it are
 *   bcond_thickness (:1555,1776), bcond_setup (:2511), AllocPrecalcData (:1814),
 *   FreePrecalcData (:2729), PrecalculateData (:646), mf_single/top/middle/bottom (:2132-2135).
 * porousfreezethaw_b200/csrc/equation_plugin_b200.c provides exactly those names for inclusion
 * in place of equation.c; it forwards to the functions declared here, which live in
 * librk_b200.so and own the device-side solver object.
 */
#ifndef PFT_EQUATION_B200_H
#define PFT_EQUATION_B200_H

#include "pft_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Create (or re-create) the device solver for this rank's block.  Called from PrecalculateData
 * once the block geometry and model parameters are known.  Returns 0 or a PFT_E* code.
 * Environment: PFT_NGPUS=n (> 1, single-rank runs only): this ONE host process cuts its block into n
 * z-slabs (intertrack.c:1780-1787) on the devices PFT_DEVICES="0,1,.." (default 0..n-1);
 * otherwise one GPU per process: PFT_DEVICE / LOCAL_RANK / OMPI_COMM_WORLD_LOCAL_RANK. */
int pft_equation_configure(const pft_case *c);

/* What RK_MPI_SA_solve drives: the solver object configured above, or the multi-GPU group (exactly one
 * of the two is non-NULL after a successful configuration). */
pft_solver *pft_equation_solver(void);
pft_group *pft_equation_group(void);
const pft_case *pft_equation_case(void);
/* counts configurations: the static glass field is uploaded once per configuration */
int pft_equation_generation(void);

/* Destroy the device solver (FreePrecalcData). */
void pft_equation_release(void);

/* The right-hand sides handed to the solver module through meta_f().  RK_MPI_SA_solve
 * recognises them and runs the fused device kernels instead of calling them; called directly
 * they evaluate f(t, w) on the device with host buffers (upload, one stage-1 kernel, download),
 * which is what the parity tests use.  Same signature as the reference's f_generic_model01 /
 * f_generic_model2 (equation.c:566,745). */
void pft_f_device_model01(double t, const double *w, double *dw_dt);
void pft_f_device_model2(double t, const double *w, double *dw_dt);
int  pft_equation_is_device_rhs(void (*f)(double, const double *, double *));

/* Sphere pack on the host with libm tanh (bit-identical to equation.c:474-529), OpenMP over
 * z-planes.  xyz_raw: count rows of unit coordinates as read from data/spheres_positions.txt.
 * Writes gl into the padded host array `solution` (3 * N1*N2*N3 doubles). */
int pft_equation_rasterise_host(const pft_case *c, const double *xyz_raw, int count, double *solution);
/* reads "x y z" rows (equation.c:476-484); returns the count (<= max_count) or -1 */
int pft_equation_read_spheres(const char *path, double *xyz_raw, int max_count);

/* ghost shell of a padded host array: equation.c:96-284 (needed by the driver only when it
 * exports the full grid, intertrack.c:2511) */
void pft_equation_bcond_host(const pft_case *c, double t, double *w);

#ifdef __cplusplus
}
#endif
#endif
