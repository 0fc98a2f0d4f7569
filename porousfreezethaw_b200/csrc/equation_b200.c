/*
 * equation_b200.c -- host side of the model plug-in (include/pft_equation_b200.h), plain C.
 *
 * Owns the one device solver object of this process (the reference's model plug-in and solver
 * module are single-instance too: file-static state in equation.c and
 * RK_MPI_SAsolver_hybrid2.c:59-69) and provides the host-only pieces of the plug-in that are
 * NOT on the hot path: sphere-file reader, host rasteriser (libm tanh, so that gl is
 * bit-identical to the reference's), ghost-shell fill for full-grid exports.
 */
#include "pft_equation_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* parameter indices: apps/intertrack-hybrid-S-freezing/model.c:44-59 */
enum {
	Q_u_star, Q_L, Q_xi, Q_a, Q_b, Q_alpha, Q_mu,
	Q_beads_scaling, Q_beads_offset_x, Q_beads_offset_y, Q_beads_offset_z,
	Q_xi_gl, Q_zeta, Q_p_eps0, Q_p_eps1, Q_gamma,
	Q_water_cp, Q_ice_cp, Q_glass_cp, Q_water_lambda, Q_ice_lambda, Q_glass_lambda,
	Q_water_rho, Q_ice_rho, Q_glass_rho, Q_top_temp1, Q_top_temp2, Q_phase_switch_time,
	Q_u_noise_amp, Q_ball_radius
};

static pft_solver *g_solver = NULL;      /* one GPU per host process ... */
static pft_group *g_group = NULL;        /* ... or one host process driving PFT_NGPUS GPUs */
static pft_case g_case;
static int g_generation = 0;             /* bumped by every (re)configuration: the glass field has to be uploaded again */

/* PFT_DEVICES="0,2,4,6" or 0..n-1 */
static void device_list(int n, int *dev)
{
	const char *e = getenv("PFT_DEVICES");
	int i;
	for (i = 0; i < n; i++) dev[i] = i;
	for (i = 0; e && *e && i < n; i++) {
		dev[i] = atoi(e);
		e = strchr(e, ',');
		if (e) e++;
	}
}

int pft_equation_configure(const pft_case *c)
{
	int dev = 0, rc, ngpus = 1;
	const char *e;
	if (!c) return PFT_EINVAL;
	pft_equation_release();
	if ((e = getenv("PFT_NGPUS")) && *e) ngpus = atoi(e);
	if (ngpus > 1 && c->nranks == 1) {
		/* ONE host process (a single MPI rank, or no MPI at all) owns the whole block and cuts it into
		   PFT_NGPUS z-slabs itself, by the reference's own rule (intertrack.c:1780-1787) */
		int devs[64];
		if (ngpus > 64) ngpus = 64;
		device_list(ngpus, devs);
		rc = pft_group_create(c, ngpus, devs, 0u, &g_group);
		if (rc) {
			fprintf(stderr, "pft_equation_configure: %s\n", pft_last_error());
			g_group = NULL;
			return rc;
		}
	} else {
		/* one process per GPU: LOCAL_RANK (torchrun) / OMPI_COMM_WORLD_LOCAL_RANK / PFT_DEVICE pick the device */
		if ((e = getenv("PFT_DEVICE")) && *e) dev = atoi(e);
		else if ((e = getenv("LOCAL_RANK")) && *e) dev = atoi(e);
		else if ((e = getenv("OMPI_COMM_WORLD_LOCAL_RANK")) && *e) dev = atoi(e);
		else if ((e = getenv("MV2_COMM_WORLD_LOCAL_RANK")) && *e) dev = atoi(e);
		rc = pft_create(c, dev, 0u, &g_solver);
		if (rc) {
			fprintf(stderr, "pft_equation_configure: %s\n", pft_last_error());
			g_solver = NULL;
			return rc;
		}
	}
	g_case = *c;
	g_generation++;
	return 0;
}

pft_solver *pft_equation_solver(void) { return g_solver; }
pft_group *pft_equation_group(void) { return g_group; }
const pft_case *pft_equation_case(void) { return (g_solver || g_group) ? &g_case : NULL; }
int pft_equation_generation(void) { return g_generation; }

void pft_equation_release(void)
{
	if (g_solver) { pft_destroy(g_solver); g_solver = NULL; }
	if (g_group) { pft_group_destroy(g_group); g_group = NULL; }
}

/* ---- the right-hand sides handed out by the meta-pointers ---------------------------------- */

static void device_rhs(double t, const double *w, double *dw)
{
	if (g_group) {
		fprintf(stderr, "porousfreezethaw_b200: a single right-hand-side evaluation with host buffers is not offered for a multi-GPU group (PFT_NGPUS > 1)\n");
		abort();
	}
	if (!g_solver) {
		fprintf(stderr, "porousfreezethaw_b200: right-hand side called before PrecalculateData configured the device solver\n");
		abort();     /* there is no CPU path to fall back to */
	}
	if (pft_rhs_host(g_solver, t, w, dw)) {
		fprintf(stderr, "porousfreezethaw_b200: device right-hand side failed: %s\n", pft_last_error());
		abort();
	}
}

void pft_f_device_model01(double t, const double *w, double *dw_dt) { device_rhs(t, w, dw_dt); }
void pft_f_device_model2(double t, const double *w, double *dw_dt) { device_rhs(t, w, dw_dt); }

int pft_equation_is_device_rhs(void (*f)(double, const double *, double *))
{
	return f == pft_f_device_model01 || f == pft_f_device_model2;
}

/* ---- sphere pack ----------------------------------------------------------------------------- */

/* equation.c:476-484 */
int pft_equation_read_spheres(const char *path, double *xyz_raw, int max_count)
{
	FILE *f = fopen(path, "r");
	int n;
	if (!f) return -1;
	for (n = 0; n < max_count; n++)
		if (fscanf(f, "%lg %lg %lg", xyz_raw + 3 * n, xyz_raw + 3 * n + 1, xyz_raw + 3 * n + 2) < 3) break;
	fclose(f);
	return n;
}

/* equation.c:479-482 (scale/offset), :508-529 (max of tanh profiles); node coordinates
 * x = L1*(0.5+i)/n1, z = L3*(0.5+k+first_row)/total_n3 */
int pft_equation_rasterise_host(const pft_case *c, const double *xyz_raw, int count, double *solution)
{
	const int N1 = c->n1 + 2 * PFT_BT, N2 = c->n2 + 2 * PFT_BT, N3 = c->n3 + 2 * PFT_BT;
	const size_t rowsize = (size_t)N1 * N2, SIZE = rowsize * N3;
	const double *P = c->param;
	double *gl = solution + 2 * SIZE;
	double *b;
	int q, k;
	if (count <= 0) return 0;
	b = (double *)malloc(sizeof(double) * 3 * (size_t)count);
	if (!b) return PFT_ENOMEM;
	for (q = 0; q < count; q++) {
		b[3 * q]     = xyz_raw[3 * q]     * P[Q_beads_scaling] + P[Q_beads_offset_x];
		b[3 * q + 1] = xyz_raw[3 * q + 1] * P[Q_beads_scaling] + P[Q_beads_offset_y];
		b[3 * q + 2] = xyz_raw[3 * q + 2] * P[Q_beads_scaling] + P[Q_beads_offset_z];
	}
	#pragma omp parallel for schedule(static)
	for (k = 0; k < c->n3; k++) {
		const double z = c->L3 * (0.5 + k + c->first_row) / c->total_n3;
		int i, j, s;
		for (j = 0; j < c->n2; j++) {
			const double y = c->L2 * (0.5 + j) / c->n2;
			double *row = gl + (size_t)(k + PFT_BT) * rowsize + (size_t)(j + PFT_BT) * N1 + PFT_BT;
			for (i = 0; i < c->n1; i++) {
				const double x = c->L1 * (0.5 + i) / c->n1;
				double v = row[i];
				for (s = 0; s < count; s++) {
					const double dx = x - b[3 * s], dy = y - b[3 * s + 1], dz = z - b[3 * s + 2];
					const double dist = sqrt(dx * dx + dy * dy + dz * dz) + 1E-10;
					const double f = 0.5 * (1.0 - tanh(0.5 / P[Q_xi_gl] * (dist - P[Q_ball_radius])));
					if (v < f) v = f;
				}
				row[i] = v;
			}
		}
	}
	free(b);
	return 0;
}

/* ---- ghost shell of a padded host array (full-grid export only) ------------------------------ */

static void mirror_plane_xy(const pft_case *c, int N1, double *plane)
{
	int i, j, m;
	for (j = 0; j < c->n2; j++) {
		double *row = plane + (size_t)(PFT_BT + j) * N1;
		for (m = 0; m < PFT_BT; m++) {
			row[PFT_BT - 1 - m] = row[PFT_BT + m];
			row[PFT_BT + c->n1 + m] = row[PFT_BT + c->n1 - 1 - m];
		}
	}
	for (m = 0; m < PFT_BT; m++) {
		const double *lo = plane + (size_t)(PFT_BT + m) * N1, *hi = plane + (size_t)(PFT_BT + c->n2 - 1 - m) * N1;
		double *glo = plane + (size_t)(PFT_BT - 1 - m) * N1, *ghi = plane + (size_t)(PFT_BT + c->n2 + m) * N1;
		for (i = 0; i < N1; i++) { glo[i] = lo[i]; ghi[i] = hi[i]; }
	}
}

/* equation.c:113-263: x/y mirror on every interior plane, z-front mirror on rank 0, z-rear
 * Dirichlet (u) or mirror (p, gl) on the last rank */
void pft_equation_bcond_host(const pft_case *c, double t, double *w)
{
	const int N1 = c->n1 + 2 * PFT_BT, N2 = c->n2 + 2 * PFT_BT, N3 = c->n3 + 2 * PFT_BT;
	const size_t rowsize = (size_t)N1 * N2, SIZE = rowsize * N3;
	const double top = t < c->param[Q_phase_switch_time] ? c->param[Q_top_temp1] : c->param[Q_top_temp2];
	int q, k, m;
	size_t i;
	for (q = 0; q < PFT_NVAR; q++) {
		double *v = w + q * SIZE;
		for (k = 0; k < c->n3; k++) mirror_plane_xy(c, N1, v + (size_t)(PFT_BT + k) * rowsize);
		if (c->rank == 0)
			for (m = 0; m < PFT_BT; m++)
				memcpy(v + (size_t)(PFT_BT - 1 - m) * rowsize, v + (size_t)(PFT_BT + m) * rowsize, sizeof(double) * rowsize);
		if (c->rank == c->nranks - 1)
			for (m = 0; m < PFT_BT; m++) {
				double *dst = v + (size_t)(PFT_BT + c->n3 + m) * rowsize;
				if (q == 0) for (i = 0; i < rowsize; i++) dst[i] = top;
				else memcpy(dst, v + (size_t)(PFT_BT + c->n3 - 1 - m) * rowsize, sizeof(double) * rowsize);
			}
	}
}
