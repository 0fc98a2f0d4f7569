/*
 * equation_plugin_b200.c -- the B200 model plug-in, to be textually included by the reference
 * driver IN PLACE OF apps/intertrack-hybrid-S-freezing/equation.c (intertrack.c:633
 * `#include "equation.c"`).  It is not compiled on its own: like equation.c it uses the
 * driver's file-static globals (n1..n3, N1..N3, total_n3, first_row, rowsize, bcond_size,
 * subgridSIZE, solution, param[], calc_mode, L1..L3, MPIrank, MPIprocs, MPIrankmap, logfile).
 *
 * It defines exactly the names the driver takes from equation.c:
 *   bcond_thickness, bcond_setup, AllocPrecalcData, FreePrecalcData, PrecalculateData,
 *   mf_single, mf_top, mf_middle, mf_bottom
 * and forwards everything to librk_b200.so / libpft_b200.so (include/pft_equation_b200.h).
 * The right-hand sides it hands to the solver are the device right-hand sides; the B200 solver
 * module (rk_mpi_sa_b200.c) recognises them and runs the fused sm_100a stage kernels.
 */
#include "pft_equation_b200.h"

#define MAX_BALLS_COUNT 100000
_string_ ball_positions_file = "data/spheres_positions.txt";      /* equation.c:35 */

/* the HOST array keeps the reference's 2-node ghost shell (equation.c:38); the device layout
 * carries no x/y ghosts at all */
static int bcond_thickness = 2;

static void pft_plugin_fill_case(pft_case *c)
{
	int q;
	c->n1 = n1; c->n2 = n2; c->n3 = n3;
	c->total_n3 = total_n3; c->first_row = first_row;
	c->rank = MPIrank; c->nranks = MPIprocs;
	c->calc_mode = calc_mode;
	c->L1 = L1; c->L2 = L2; c->L3 = L3;
	c->max_n3 = 0;               /* the application's own slab partition (intertrack.c:1780-1787) */
	for (q = 0; q < PARAM_COUNT && q < PFT_NPARAM; q++) c->param[q] = param[q];
}

/* equation.c:266-284 -- only reached when the driver exports the full grid (intertrack.c:2511) */
void bcond_setup(FLOAT t, FLOAT *w)
{
	pft_case c;
	pft_plugin_fill_case(&c);
	pft_equation_bcond_host(&c, t, w);
}

/* equation.c:427-437: the reference allocates the noise field here; the device solver needs none */
int AllocPrecalcData(void) { return 0; }

void FreePrecalcData(void) { pft_equation_release(); }

/* equation.c:439-558 */
int PrecalculateData(FLOAT *var_eps_mult)
{
	char *Glass_balls_errors[] = { "Reading glass balls positions failed." };
	char *Device_errors[] = { "Creating the B200 device solver failed (see stderr)." };
	double *xyz = NULL;
	int ball_count = 0, error_code = 0;
	pft_case c;
	(void)var_eps_mult;

	if (MPIrank == 0) {
		xyz = (double *)malloc(sizeof(double) * 3 * MAX_BALLS_COUNT);
		ball_count = xyz ? pft_equation_read_spheres(ball_positions_file, xyz, MAX_BALLS_COUNT) : -1;
		if (ball_count >= 0) Mmprintf(logfile, "Successfully read coordinates of %d glass balls.\n\n", ball_count);
		else {
			Mmprintf(logfile, "ERROR: Could not read glass balls coordinates from: %s\n", ball_positions_file);
			error_code = 1;
		}
	}
	CheckErrorAcrossRanks(error_code, 1, Glass_balls_errors);

	MPI_Bcast(&ball_count, 1, MPI_INT, MPIrankmap[0], MPI_COMM_WORLD);
	if (MPIrank != 0) xyz = (double *)malloc(sizeof(double) * 3 * (size_t)(ball_count > 0 ? ball_count : 1));
	MPI_Bcast(xyz, 3 * ball_count, MPI_DOUBLE, MPIrankmap[0], MPI_COMM_WORLD);

	pft_plugin_fill_case(&c);
	/* host rasteriser with libm tanh: gl bit-identical to the reference's (it is uploaded with the
	   rest of `solution` at the first RK_MPI_SA_solve) */
	pft_equation_rasterise_host(&c, xyz, ball_count, solution);
	free(xyz);

	if (MPIrank == 0) {
		_string_ models[] = { "Phase field / GradP", "Phase field / SigmaP1-P", "Heat equation with latent heat release focusing" };
		switch (calc_mode) {
		case 0: case 1: case 2:   Mmprintf(logfile, "\nSolidification model: %s\n\n", models[calc_mode]); break;
		case 10: case 11:         Mmprintf(logfile, "\nSolidification model: %s\n\n", models[calc_mode - 10]); break;
		default:
			Mmprintf(logfile, "\nError : invalid calc_mode value %d\n\n", calc_mode);
			return 1;
		}
	}

	error_code = pft_equation_configure(&c) ? 1 : 0;
	CheckErrorAcrossRanks(error_code, 1, Device_errors);

	/* several ranks, one GPU each: the slabs find each other (NCCL bootstrap, then peer memory over NVLink;
	   include/pft_b200.h).  The 128-byte id travels the way the reference moves its setup data: MPI_Bcast. */
	if (MPIprocs > 1) {
		unsigned char comm_id[128];
		error_code = 0;
		if (MPIrank == 0) error_code = pft_comm_unique_id(comm_id) ? 1 : 0;
		CheckErrorAcrossRanks(error_code, 1, Device_errors);
		MPI_Bcast(comm_id, 128, MPI_BYTE, MPIrankmap[0], MPI_COMM_WORLD);
		error_code = pft_comm_init(pft_equation_solver(), comm_id) ? 1 : 0;
		if (error_code) Mmprintf(logfile, "ERROR: %s\n", pft_last_error());
		CheckErrorAcrossRanks(error_code, 1, Device_errors);
	}
	return 0;
}

/* equation.c:945-973: one right-hand side for every rank position; the device kernels take the
 * slab position from the pft_case */
static RK_RightHandSide pft_plugin_mf(void)
{
	return (calc_mode == 2) ? (RK_RightHandSide)pft_f_device_model2 : (RK_RightHandSide)pft_f_device_model01;
}
RK_RightHandSide mf_single() { return pft_plugin_mf(); }
RK_RightHandSide mf_top()    { return pft_plugin_mf(); }
RK_RightHandSide mf_middle() { return pft_plugin_mf(); }
RK_RightHandSide mf_bottom() { return pft_plugin_mf(); }
