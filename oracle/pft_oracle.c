/*
 * oracle/pft_oracle.c -- CPU restatement of the PorousFreezeThaw hot path (plain C99).
 *
 * TEST INFRASTRUCTURE ONLY -- see pft_oracle.h.  Every function cites the reference
 * file:line it restates (paths relative to /root/reference).  Arithmetic is strict
 * IEEE-754 double in the reference's association order; build with
 * -ffp-contract=off (the reference is built with gcc -O1, which never forms FMAs).
 *
 * This is a restatement, not a copy: loops are organised around an index helper and
 * a per-cell function instead of the reference's pointer walking, and the OpenMP
 * decomposition is a plain collapse over (k,j).  Results are bit-identical because
 * every floating-point expression keeps the reference's operand order.
 */
#include "pft_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define P(name) (c->param[PFO_##name])

typedef struct {
	int N1, N2, N3, rowsize, bcond_size, SIZE;
} geom;

static geom make_geom(const pfo_case *c)
{
	geom g;
	g.N1 = c->n1 + 2 * PFO_BT;
	g.N2 = c->n2 + 2 * PFO_BT;
	g.N3 = c->n3 + 2 * PFO_BT;
	g.rowsize = g.N1 * g.N2;
	g.bcond_size = PFO_BT * g.rowsize;
	g.SIZE = g.rowsize * g.N3;
	return g;
}

/* intertrack.c:1776-1800 */
void pfo_sizes(const pfo_case *c, int *out)
{
	geom g = make_geom(c);
	out[0] = g.N1; out[1] = g.N2; out[2] = g.N3; out[3] = g.rowsize;
	out[4] = g.bcond_size; out[5] = g.SIZE; out[6] = PFO_NVAR * c->n2 * c->n3; out[7] = 0;
}

/* intertrack.c:1780-1787: total_n3/nranks rows each, the remainder one by one to the lowest ranks */
void pfo_partition(int total_n3, int nranks, int rank, int *n3, int *first_row)
{
	int q = total_n3 / nranks, r = total_n3 % nranks;
	*n3 = q;
	*first_row = rank * q;
	if (rank < r) { (*n3)++; *first_row += rank; }
	else *first_row += r;
}

/* ------------------------------------------------------------------------- */
/* sphere pack                                                               */

/* equation.c:476-484 (file read part; scaling is applied in pfo_rasterise) */
int pfo_read_spheres(const char *path, double *xyz, int max_count)
{
	FILE *f = fopen(path, "r");
	int n;
	if (!f) return -1;
	for (n = 0; n < max_count; n++)
		if (fscanf(f, "%lg %lg %lg", xyz + 3 * n, xyz + 3 * n + 1, xyz + 3 * n + 2) < 3) break;
	fclose(f);
	return n;
}

/* equation.c:332-336 */
static inline double euclidean_norm(double v1, double v2, double v3)
{
	return sqrt(v1 * v1 + v2 * v2 + v3 * v3) + 1E-10;
}

/* equation.c:479-482 (scale + offset) and :508-529 (max of tanh profiles, node coordinates) */
void pfo_rasterise(const pfo_case *c, const double *xyz_raw, int count, double *w)
{
	geom g = make_geom(c);
	double *bx = (double *)malloc(sizeof(double) * (size_t)(count > 0 ? count : 1) * 3);
	double *by = bx + (count > 0 ? count : 1), *bz = by + (count > 0 ? count : 1);
	double *gl = w + (size_t)2 * g.SIZE;
	int q, k;

	for (q = 0; q < count; q++) {
		bx[q] = xyz_raw[3 * q]     * P(beads_scaling) + P(beads_offset_x);
		by[q] = xyz_raw[3 * q + 1] * P(beads_scaling) + P(beads_offset_y);
		bz[q] = xyz_raw[3 * q + 2] * P(beads_scaling) + P(beads_offset_z);
	}

	#pragma omp parallel for schedule(static)
	for (k = 0; k < c->n3; k++) {
		int i, j, qq;
		double z = c->L3 * (0.5 + k + c->first_row) / c->total_n3;
		for (j = 0; j < c->n2; j++) {
			double y = c->L2 * (0.5 + j) / c->n2;
			double *ptr = gl + (size_t)(k + PFO_BT) * g.rowsize + (size_t)(j + PFO_BT) * g.N1 + PFO_BT;
			for (i = 0; i < c->n1; i++) {
				double x = c->L1 * (0.5 + i) / c->n1;
				double v = ptr[i];
				for (qq = 0; qq < count; qq++) {
					double glass_phf = 0.5 * (1.0 - tanh(0.5 / P(xi_gl) * (euclidean_norm(x - bx[qq], y - by[qq], z - bz[qq]) - P(ball_radius))));
					if (v < glass_phf) v = glass_phf;
				}
				ptr[i] = v;
			}
		}
	}
	free(bx);
}

/* ------------------------------------------------------------------------- */
/* boundary conditions                                                       */

/* in-plane mirror of one z-plane: equation.c:142-160 (identical at :218-236) */
static void mirror_xy_plane(const pfo_case *c, const geom *g, double *plane)
{
	int i, j, m;
	for (j = 0; j < c->n2; j++) {
		double *row = plane + (size_t)(PFO_BT + j) * g->N1;
		for (m = 0; m < PFO_BT; m++) {
			row[PFO_BT - 1 - m] = row[PFO_BT + m];                          /* left edge  */
			row[PFO_BT + c->n1 + m] = row[PFO_BT + c->n1 - 1 - m];          /* right edge */
		}
	}
	for (m = 0; m < PFO_BT; m++) {
		double *src_b = plane + (size_t)(PFO_BT + m) * g->N1,              *dst_b = plane + (size_t)(PFO_BT - m - 1) * g->N1;
		double *src_t = plane + (size_t)(PFO_BT + c->n2 - m - 1) * g->N1,  *dst_t = plane + (size_t)(PFO_BT + c->n2 + m) * g->N1;
		for (i = 0; i < g->N1; i++) { dst_b[i] = src_b[i]; dst_t[i] = src_t[i]; }
	}
}

/* equation.c:96-111 */
static inline double temperature_dirichlet(const pfo_case *c, double t)
{
	return t < P(phase_switch_time) ? P(top_temp1) : P(top_temp2);
}

/* one variable: equation.c:113-185 (dirichlet_top = 1, used for u) / :187-263 (pure Neumann, p and gl) */
static void bcond_var(const pfo_case *c, const geom *g, double t, double *v, int dirichlet_top)
{
	int k, m;
	#pragma omp parallel for schedule(static)
	for (k = 0; k < c->n3; k++)
		mirror_xy_plane(c, g, v + (size_t)(PFO_BT + k) * g->rowsize);

	if (c->rank == 0)                      /* z-front: mirror, whole planes incl. xy ghosts (:164-174) */
		for (m = 0; m < PFO_BT; m++)
			memcpy(v + (size_t)(PFO_BT - 1 - m) * g->rowsize, v + (size_t)(PFO_BT + m) * g->rowsize, sizeof(double) * (size_t)g->rowsize);

	if (c->rank == c->nranks - 1) {        /* z-rear (:175-183 Dirichlet, :251-261 mirror) */
		for (m = 0; m < PFO_BT; m++) {
			double *dst = v + (size_t)(PFO_BT + c->n3 + m) * g->rowsize;
			if (dirichlet_top) {
				double val = temperature_dirichlet(c, t);
				int i;
				for (i = 0; i < g->rowsize; i++) dst[i] = val;
			} else
				memcpy(dst, v + (size_t)(PFO_BT + c->n3 - 1 - m) * g->rowsize, sizeof(double) * (size_t)g->rowsize);
		}
	}
}

/* equation.c:266-284 */
void pfo_bcond(const pfo_case *c, double t, double *w)
{
	geom g = make_geom(c);
	bcond_var(c, &g, t, w, 1);
	bcond_var(c, &g, t, w + (size_t)g.SIZE, 0);
	bcond_var(c, &g, t, w + (size_t)2 * g.SIZE, 0);
}

/* ------------------------------------------------------------------------- */
/* constitutive laws                                                         */

typedef struct {
	double h1_2, h2_2, h3_2, h1_div_2, h2_div_2, h3_div_2;
	double xi_2_inv_a, xi_inv_b_sqrt_a2, eps2_3, eps3_2;
} consts;

/* equation.c:442-447 and :605-612 (h3 uses total_n3) */
static consts make_consts(const pfo_case *c)
{
	consts k;
	double h1 = ((double)c->n1) / c->L1;
	double h2 = ((double)c->n2) / c->L2;
	double h3 = ((double)c->total_n3) / c->L3;
	k.h1_2 = h1 * h1; k.h1_div_2 = 0.5 * h1;
	k.h2_2 = h2 * h2; k.h2_div_2 = 0.5 * h2;
	k.h3_2 = h3 * h3; k.h3_div_2 = 0.5 * h3;
	k.xi_2_inv_a = P(a) / (P(xi) * P(xi));
	k.xi_inv_b_sqrt_a2 = P(b) * sqrt(0.5 * P(a)) / P(xi);
	k.eps2_3 = 3.0 / ((P(p_eps1) - P(p_eps0)) * (P(p_eps1) - P(p_eps0)));
	k.eps3_2 = 2.0 / ((P(p_eps1) - P(p_eps0)) * (P(p_eps1) - P(p_eps0)) * (P(p_eps1) - P(p_eps0)));
	return k;
}

/* equation.c:341-357: g*X_glass + (1-g)*(p*X_ice + (1-p)*X_water) */
static inline double mix(double p, double gl, double x_glass, double x_ice, double x_water)
{
	return gl * x_glass + (1.0 - gl) * (p * x_ice + (1.0 - p) * x_water);
}
#define RHO(p, gl)    mix(p, gl, P(glass_rho), P(ice_rho), P(water_rho))
#define CP(p, gl)     mix(p, gl, P(glass_cp), P(ice_cp), P(water_cp))
#define LAMBDA(p, gl) mix(p, gl, P(glass_lambda), P(ice_lambda), P(water_lambda))

/* equation.c:359-363 */
static inline double water_indicator(const pfo_case *c, double gl) { return fmax(0.0, 1.0 - P(zeta) * gl); }

/* equation.c:367-371 */
static inline double f_GradP(const pfo_case *c, const consts *k, double u, double p, double gradp_norm)
{
	return k->xi_2_inv_a * p * (1.0 - p) * (p - 0.5) - P(b) * P(alpha) * P(mu) * gradp_norm * (u - P(u_star));
}

/* equation.c:375-382 */
static inline double Sshape(const pfo_case *c, const consts *k, double x)
{
	if (x <= P(p_eps0)) return 0.0;
	if (x >= P(p_eps1)) return 1.0;
	x -= P(p_eps0);
	return x * x * (k->eps2_3 - k->eps3_2 * x);
}

/* equation.c:384-388 */
static inline double f_SigmaP1_P(const pfo_case *c, const consts *k, double u, double p)
{
	return k->xi_2_inv_a * p * (1.0 - p) * (p - 0.5)
	     - k->xi_inv_b_sqrt_a2 * P(alpha) * P(mu) * Sshape(c, k, p) * Sshape(c, k, 1.0 - p) * fmax(p * (1.0 - p), 0.0) * (u - P(u_star));
}

/* equation.c:410-421 */
static inline double dphf_du(const pfo_case *c, double u)
{
	double aux = cosh(P(gamma) * (u - P(u_star)));
	return -0.5 * P(gamma) / (aux * aux);
}

/* div(lambda grad u) numerator shared by both models: equation.c:711-720 == :850-859 */
static inline double heat_flux_sum(const pfo_case *c, const consts *k, const double *u, const double *p, const double *gl, int sx, int sy, int sz)
{
	(void)sx;
	return
	    k->h1_2 * ( - LAMBDA(0.5 * (p[-1] + p[0]),  0.5 * (gl[-1] + gl[0]))  * (-u[-1] + u[0])
	                + LAMBDA(0.5 * (p[0] + p[1]),   0.5 * (gl[0] + gl[1]))   * (-u[0] + u[1]) )
	  + k->h2_2 * ( - LAMBDA(0.5 * (p[-sy] + p[0]), 0.5 * (gl[-sy] + gl[0])) * (-u[-sy] + u[0])
	                + LAMBDA(0.5 * (p[0] + p[sy]),  0.5 * (gl[0] + gl[sy]))  * (-u[0] + u[sy]) )
	  + k->h3_2 * ( - LAMBDA(0.5 * (p[-sz] + p[0]), 0.5 * (gl[-sz] + gl[0])) * (-u[-sz] + u[0])
	                + LAMBDA(0.5 * (p[0] + p[sz]),  0.5 * (gl[0] + gl[sz]))  * (-u[0] + u[sz]) );
}

/* one node of f_generic_model01: equation.c:650-731 */
static inline void cell_model01(const pfo_case *c, const consts *k, const double *u, const double *p, const double *gl,
                                int sy, int sz, double *du_dt, double *dp_dt)
{
	double this_rho = RHO(p[0], gl[0]);
	double this_cp = CP(p[0], gl[0]);
	double dp;

	dp = ( k->h1_2 * ( - ( - p[-1]  + p[0] ) + ( - p[0] + p[1]  ) )
	     + k->h2_2 * ( - ( - p[-sy] + p[0] ) + ( - p[0] + p[sy] ) )
	     + k->h3_2 * ( - ( - p[-sz] + p[0] ) + ( - p[0] + p[sz] ) ) );

	switch (c->calc_mode) {
	case 0: case 10:
		dp += f_GradP(c, k, u[0], p[0],
		              euclidean_norm(k->h1_div_2 * ( - p[-1]  + p[1]  ),
		                             k->h2_div_2 * ( - p[-sy] + p[sy] ),
		                             k->h3_div_2 * ( - p[-sz] + p[sz] )));
		break;
	case 1: case 11:
		dp += f_SigmaP1_P(c, k, u[0], p[0]);
		break;
	}
	dp /= P(alpha);
	dp *= water_indicator(c, gl[0]);
	*dp_dt = dp;

	switch (c->calc_mode) {
	case 10: case 11:
		*du_dt = 0.0;
		break;
	default:
		*du_dt = ( heat_flux_sum(c, k, u, p, gl, 1, sy, sz) / this_rho + P(L) * dp ) / this_cp;
	}
}

/* one node of f_generic_model2: equation.c:835-874 */
static inline void cell_model2(const pfo_case *c, const consts *k, const double *u, const double *p, const double *gl,
                               int sy, int sz, double *du_dt, double *dp_dt)
{
	double this_rho = RHO(p[0], gl[0]);
	double this_cp = CP(p[0], gl[0]);
	double dp_du = dphf_du(c, u[0]) * water_indicator(c, gl[0]);
	double du = heat_flux_sum(c, k, u, p, gl, 1, sy, sz) / ( this_rho * (this_cp - P(L) * dp_du) );
	*du_dt = du;
	*dp_dt = dp_du * du;
}

/* equation.c:566-741 / :745-884 */
void pfo_rhs(const pfo_case *c, double t, double *w, double *dw, pfo_halo_fn halo, void *halo_ctx)
{
	geom g = make_geom(c);
	consts k = make_consts(c);
	const int sy = g.N1, sz = g.rowsize;
	int kk;

	pfo_bcond(c, t, w);                                   /* :624 */
	if (c->nranks > 1 && halo) halo(w, halo_ctx);         /* :626 */

	#pragma omp parallel for collapse(2) schedule(static)
	for (kk = 0; kk < c->n3; kk++)
		for (int j = 0; j < c->n2; j++) {
			size_t off = (size_t)(kk + PFO_BT) * g.rowsize + (size_t)(j + PFO_BT) * g.N1 + PFO_BT;
			const double *u = w + off, *p = w + g.SIZE + off, *gl = w + (size_t)2 * g.SIZE + off;
			double *du = dw + off, *dp = dw + g.SIZE + off, *dgl = dw + (size_t)2 * g.SIZE + off;
			int i;
			if (c->calc_mode == 2)
				for (i = 0; i < c->n1; i++) { cell_model2(c, &k, u + i, p + i, gl + i, sy, sz, du + i, dp + i); dgl[i] = 0.0; }
			else
				for (i = 0; i < c->n1; i++) { cell_model01(c, &k, u + i, p + i, gl + i, sy, sz, du + i, dp + i); dgl[i] = 0.0; }
		}
}

/* ------------------------------------------------------------------------- */
/* Runge-Kutta-Merson                                                        */

#define CMD_h_TOO_SMALL 1
#define CMD_NAN         2
#define CMD_UPDATE      4
#define CMD_FINISHED    8
#define CMD_NEXTFINISH  16
#define CMD_BREAK       32

/* the interior rows of all variables = the chunk table of intertrack.c:2149-2156 */
#define FOR_CHUNKS(body) \
	_Pragma("omp parallel for collapse(2) schedule(static)") \
	for (int q_ = 0; q_ < PFO_NVAR; q_++) \
		for (int r_ = 0; r_ < c->n2 * c->n3; r_++) { \
			size_t o_ = (size_t)q_ * g.SIZE + (size_t)(r_ / c->n2 + PFO_BT) * g.rowsize + (size_t)(r_ % c->n2 + PFO_BT) * g.N1 + PFO_BT; \
			for (int i = 0; i < c->n1; i++) { size_t I = o_ + i; body; } \
		}

/* RK_MPI_SAsolver_hybrid2.c:215-770 (single thread of control; collectives through the hooks) */
int pfo_solve(const pfo_case *c, pfo_system *s, double *x, double *work, double final_time)
{
	geom g = make_geom(c);
	const size_t n = (size_t)PFO_NVAR * g.SIZE;
	double *K1 = work, *K3 = work + n, *K4 = work + 2 * n, *K5 = work + 3 * n, *aux = work + 4 * n;
	double *K2 = K3;                                      /* :301 */
	double t, h, new_h = 0.0, h_min, delta, h2, h3, h6, h8, eps, max_eps;
	int command = 0;
	long attempts = 0;

	if (!s || !x || !work) return -2;
	t = s->t; h = s->h; h_min = s->h_min; delta = s->delta;
	if (delta <= 0) return -2;                            /* :286 */
	s->last_nan = 0;                                      /* :316 */

	/* :319-325 */
	if ((final_time > t && h < 0) || (final_time < t && h > 0)) h *= -1;
	if (h == 0 || fabs(final_time - t) <= fabs(h)) { h = final_time - t; command |= CMD_FINISHED; }

	for (;;) {
		int nan_local = 0;
		h2 = h / 2.0; h3 = h / 3.0; h6 = h / 6.0; h8 = h / 8.0;          /* :355 */

		pfo_rhs(c, t, x, K1, s->halo, s->halo_ctx);                       /* :373 */
		FOR_CHUNKS(aux[I] = K1[I] * h3 + x[I]);                           /* :378-389 */
		pfo_rhs(c, t + h3, aux, K2, s->halo, s->halo_ctx);                /* :392 */
		FOR_CHUNKS(aux[I] = (K1[I] + K2[I]) * h6 + x[I]);                 /* :397-409 */
		pfo_rhs(c, t + h3, aux, K3, s->halo, s->halo_ctx);                /* :412 */
		FOR_CHUNKS(aux[I] = (K1[I] + 3.0 * K3[I]) * h8 + x[I]);           /* :417-429 */
		pfo_rhs(c, t + h2, aux, K4, s->halo, s->halo_ctx);                /* :432 */
		FOR_CHUNKS(aux[I] = (0.5 * K1[I] - 1.5 * K3[I] + 2.0 * K4[I]) * h + x[I]);   /* :437-450 */
		pfo_rhs(c, t + h, aux, K5, s->halo, s->halo_ctx);                 /* :453 */
		s->rhs_evals += 5;

		s->steps_total++;                                                  /* :460 */
		attempts++;
		eps = 0.0;

		if (s->handle_nan) {                                               /* :465-490, sequential scan with early exit */
			for (int q_ = 0; q_ < PFO_NVAR && !nan_local; q_++)
				for (int r_ = 0; r_ < c->n2 * c->n3 && !nan_local; r_++) {
					size_t o_ = (size_t)q_ * g.SIZE + (size_t)(r_ / c->n2 + PFO_BT) * g.rowsize + (size_t)(r_ % c->n2 + PFO_BT) * g.N1 + PFO_BT;
					for (int i = 0; i < c->n1; i++) {
						size_t I = o_ + i;
						double e = 1.0 * fabs(0.2 * K1[I] - 0.9 * K3[I] + 0.8 * K4[I] - 0.1 * K5[I]);
						if (!isfinite(e)) { nan_local = 1; break; }
						if (e > eps) eps = e;
					}
				}
		} else {                                                           /* :508-524 */
			double m = 0.0;
			#pragma omp parallel for collapse(2) schedule(static) reduction(max:m)
			for (int q_ = 0; q_ < PFO_NVAR; q_++)
				for (int r_ = 0; r_ < c->n2 * c->n3; r_++) {
					size_t o_ = (size_t)q_ * g.SIZE + (size_t)(r_ / c->n2 + PFO_BT) * g.rowsize + (size_t)(r_ % c->n2 + PFO_BT) * g.N1 + PFO_BT;
					for (int i = 0; i < c->n1; i++) {
						size_t I = o_ + i;
						double e = 1.0 * fabs(0.2 * K1[I] - 0.9 * K3[I] + 0.8 * K4[I] - 0.1 * K5[I]);
						if (e > m) m = e;
					}
				}
			eps = m;
		}
		if (nan_local) {                                                   /* :493-503; single rank: local == global */
			command |= CMD_NAN;
			if (h / (final_time - t) < 1e-11) command |= CMD_h_TOO_SMALL;
		}

		max_eps = s->allmax ? s->allmax(eps, s->allmax_ctx) : eps;        /* :572 */
		if (s->delta_local) max_eps *= fabs(h3);                          /* :578 */
		new_h = ((max_eps > 0.0) ? pow((delta / max_eps), 0.2) * 0.8 : 2.0) * h;   /* :580 */

		if (max_eps < delta || fabs(h) < h_min) {                         /* :599-611 */
			command |= CMD_UPDATE;
			if (fabs(final_time - (t + h)) <= fabs(new_h)) command |= CMD_NEXTFINISH;
		}

		if (command & CMD_NAN) {                                           /* :626-646 */
			s->last_nan = 1;
			if (command & CMD_h_TOO_SMALL) { s->t = t; return -4; }
			h /= 10;
			command = 0;
		} else {
			if (command & CMD_UPDATE) {                                    /* :651-734 */
				t += h;
				FOR_CHUNKS(x[I] += h3 * (0.5 * (K1[I] + K5[I]) + 2.0 * K4[I]));
				s->steps++;
				if (s->service) {
					s->t = t; s->h = h;
					if (s->service(final_time, t, h, s->steps, s->steps_total, s->service_ctx)) command |= CMD_BREAK;
				}
				if (command & CMD_FINISHED) break;
				if (command & CMD_BREAK) { s->t = t; s->h = new_h; return 1; }
			}
			if (command & CMD_NEXTFINISH) {                                /* :745-760 */
				s->h = new_h;
				h = final_time - t;
				command = CMD_FINISHED;
			} else {
				command = 0;
				h = new_h;
			}
		}

		if (s->max_attempts > 0 && attempts >= s->max_attempts) {
			/* extension (not in the reference): leave a state from which a later call
			   re-derives exactly the (h, command) pair held here */
			s->t = t;
			if (!(command & CMD_FINISHED)) s->h = h;
			return 2;
		}
	}

	s->t = t;                                                              /* :768 */
	return 0;
}

/* scripts/avg.sh:9-20: mean over all nodes of (p > 0.5) */
double pfo_ice_fraction(const pfo_case *c, const double *w)
{
	geom g = make_geom(c);
	const double *p = w + g.SIZE;
	long cnt = 0;
	int k, j, i;
	for (k = 0; k < c->n3; k++)
		for (j = 0; j < c->n2; j++)
			for (i = 0; i < c->n1; i++)
				if (p[(size_t)(k + PFO_BT) * g.rowsize + (size_t)(j + PFO_BT) * g.N1 + PFO_BT + i] > 0.5) cnt++;
	return (double)cnt / ((double)c->n1 * c->n2 * c->n3);
}

/* the C library's pow() on an array: the reference's step-size controller calls exactly this function
 * (RK_MPI_SAsolver_hybrid2.c:580); the tests hold the device twin (csrc/pft_pow.h) against it */
void pfo_pow_array(const double *x, double y, double *out, long n)
{
	long i;
	#pragma omp parallel for schedule(static)
	for (i = 0; i < n; i++) out[i] = pow(x[i], y);
}
