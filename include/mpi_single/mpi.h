/*
 * include/mpi_single/mpi.h -- single-rank stand-in for <mpi.h>, for hosts without an MPI installation.
 *
 * The B200 drop-in host objects (csrc/rk_mpi_sa_b200.c, the solver-module ABI of
 * include/RK_MPI_SAsolver.h names MPI_Comm) are compiled against it when real MPI is not available
 * (PFT_HAVE_MPI undefined): MPI_Comm is then only a token, every collective degenerates to a local
 * copy, and one host process drives all GPUs of the box (PFT_NGPUS).
 * Only the MPI-1 entry points the reference application actually calls are declared; the
 * point-to-point calls are never reached with one rank (they return an error code if they ever are).
 *
 * The checker tree uses the same header (oracle/shim/mpi.h includes it) to compile the unmodified
 * reference sources in this image, which has no mpicc / mpi.h.
 */
#ifndef PFT_ORACLE_SHIM_MPI_H
#define PFT_ORACLE_SHIM_MPI_H

#include <string.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;   /* value == element size in bytes */
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD        0
#define MPI_SUCCESS           0
#define MPI_ERR_OTHER         15

#define MPI_CHAR              ((MPI_Datatype)1)
#define MPI_BYTE              ((MPI_Datatype)1)
#define MPI_INT               ((MPI_Datatype)(0x100 | 4))
#define MPI_FLOAT             ((MPI_Datatype)(0x200 | 4))
#define MPI_DOUBLE            ((MPI_Datatype)8)
#define MPI_LONG_DOUBLE       ((MPI_Datatype)16)

#define MPI_MIN               1
#define MPI_MAX               2
#define MPI_BOR               3

#define MPI_THREAD_SINGLE     0
#define MPI_THREAD_FUNNELED   1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE   3

static inline size_t pft_shim_mpi_bytes(int count, MPI_Datatype t) { return (size_t)count * (size_t)(t & 0xff); }

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Init_thread(int *argc, char ***argv, int required, int *provided)
{ (void)argc; (void)argv; if (provided) *provided = required; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = 1; return MPI_SUCCESS; }
static inline int MPI_Get_processor_name(char *name, int *len) { strcpy(name, "single-rank-shim"); *len = (int)strlen(name); return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }

/* point-to-point: unreachable with one rank */
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; return MPI_ERR_OTHER; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s; return MPI_ERR_OTHER; }
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; (void)r; return MPI_ERR_OTHER; }
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)r; return MPI_ERR_OTHER; }
static inline int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) { (void)n; (void)r; (void)s; return MPI_SUCCESS; }

/* collectives over a communicator of one */
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS; }
static inline int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{ (void)rn; (void)rt; (void)root; (void)c; if (r != s) memmove(r, s, pft_shim_mpi_bytes(n, t)); return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)op; (void)c; if (r != s) memmove(r, s, pft_shim_mpi_bytes(n, t)); return MPI_SUCCESS; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void)op; (void)root; (void)c; if (r != s) memmove(r, s, pft_shim_mpi_bytes(n, t)); return MPI_SUCCESS; }

/* defined in shim_rt.c (needs POSIX.1b clocks, which the reference's feature macros hide) */
double MPI_Wtime(void);

#endif /* PFT_ORACLE_SHIM_MPI_H */
