/* oracle/shim/mpi.h -- TEST INFRASTRUCTURE: the unmodified reference sources are compiled against the
 * single-rank <mpi.h> stand-in that the product ships for hosts without MPI. */
#include "../../include/mpi_single/mpi.h"
