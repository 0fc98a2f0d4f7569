/*
 * oracle/shim/shim_rt.c -- runtime for the mpi.h / netcdf.h stand-ins.
 * TEST INFRASTRUCTURE ONLY (linked into oracle/_ref/ binaries, never into the product).
 *
 * "PFTSNAP 1" container (what the nc_* shim reads and writes):
 *
 *   PFTSNAP 1\n
 *   dim  <name> <len>\n                      one per nc_def_dim, ids in order
 *   var  <name> <ndims> <dimid> ...\n        one per nc_def_var, ids in order
 *   attd <name> <%.17g>\n                    global double attribute
 *   atti <name> <int>\n                      global int attribute
 *   attt <name> <nbytes>\n<raw bytes>\n      global text attribute
 *   end\n
 *   <float64 little-endian payload: the variables back to back, in definition order,
 *    each prod(dims) values in C order>
 *
 * This mirrors the dataset the reference writes with libnetcdf
 * (apps/intertrack-hybrid-S-freezing/intertrack.c:2331-2548): dims n3,n2,n1;
 * coordinate variables n3,n2,n1; u,p,gl as double[n3][n2][n1]; global attributes
 * L1..L3, every model parameter, calc_mode, delta, tau, t, final_time, snapshot,
 * total_snapshots, title.  One dataset may be open at a time (the reference never
 * opens two).
 */
#define _POSIX_C_SOURCE 200809L
#include "netcdf.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

double MPI_Wtime(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------- */

#define MAXDIMS 8
#define MAXVARS 16
#define MAXATTS 96
#define NAMELEN 64

typedef struct { char name[NAMELEN]; size_t len; } DIM;
typedef struct { char name[NAMELEN]; int ndims; int dimid[MAXDIMS]; size_t count; long offset; } VAR;
typedef struct { char name[NAMELEN]; char kind; double d; int i; char *text; size_t tlen; } ATT;

static struct {
	FILE *f;
	int writing;
	int defining;
	int ndims, nvars, natts;
	DIM dim[MAXDIMS];
	VAR var[MAXVARS];
	ATT att[MAXATTS];
	long payload;      /* file offset of the binary payload */
} D;

static void reset(void)
{
	int i;
	for (i = 0; i < D.natts; i++) free(D.att[i].text);
	memset(&D, 0, sizeof D);
}

static void layout(void)
{
	int v, k;
	long off = 0;
	for (v = 0; v < D.nvars; v++) {
		size_t c = 1;
		for (k = 0; k < D.var[v].ndims; k++) c *= D.dim[D.var[v].dimid[k]].len;
		D.var[v].count = c;
		D.var[v].offset = off;
		off += (long)(c * sizeof(double));
	}
}

int nc_set_default_format(int format, int *old_format) { (void)format; if (old_format) *old_format = NC_FORMAT_NETCDF4; return NC_NOERR; }

const char *nc_strerror(int err)
{
	switch (err) {
	case NC_NOERR: return "no error";
	case NC_EPFT_IO: return "PFTSNAP shim: I/O error";
	case NC_EPFT_NOTFOUND: return "PFTSNAP shim: name not found";
	case NC_EPFT_STATE: return "PFTSNAP shim: operation not allowed in this state";
	}
	return "PFTSNAP shim: unknown error";
}

int nc_create(const char *path, int mode, int *ncid)
{
	(void)mode;
	if (D.f) return NC_EPFT_STATE;
	reset();
	D.f = fopen(path, "wb");
	if (!D.f) return NC_EPFT_IO;
	D.writing = 1; D.defining = 1;
	*ncid = 1;
	return NC_NOERR;
}

int nc_def_dim(int ncid, const char *name, size_t len, int *dimid)
{
	(void)ncid;
	if (!D.f || !D.defining || D.ndims >= MAXDIMS) return NC_EPFT_STATE;
	strncpy(D.dim[D.ndims].name, name, NAMELEN - 1);
	D.dim[D.ndims].len = len;
	*dimid = D.ndims++;
	return NC_NOERR;
}

int nc_def_var(int ncid, const char *name, int type, int ndims, const int *dimids, int *varid)
{
	int k;
	(void)ncid; (void)type;
	if (!D.f || !D.defining || D.nvars >= MAXVARS || ndims > MAXDIMS) return NC_EPFT_STATE;
	strncpy(D.var[D.nvars].name, name, NAMELEN - 1);
	D.var[D.nvars].ndims = ndims;
	for (k = 0; k < ndims; k++) D.var[D.nvars].dimid[k] = dimids[k];
	*varid = D.nvars++;
	return NC_NOERR;
}

static ATT *new_att(const char *name, char kind)
{
	ATT *a;
	if (!D.f || !D.defining || D.natts >= MAXATTS) return NULL;
	a = &D.att[D.natts++];
	memset(a, 0, sizeof *a);
	strncpy(a->name, name, NAMELEN - 1);
	a->kind = kind;
	return a;
}

int nc_put_att_double(int ncid, int varid, const char *name, int type, size_t len, const double *v)
{
	ATT *a = new_att(name, 'd');
	(void)ncid; (void)varid; (void)type; (void)len;
	if (!a) return NC_EPFT_STATE;
	a->d = *v;
	return NC_NOERR;
}

int nc_put_att_int(int ncid, int varid, const char *name, int type, size_t len, const int *v)
{
	ATT *a = new_att(name, 'i');
	(void)ncid; (void)varid; (void)type; (void)len;
	if (!a) return NC_EPFT_STATE;
	a->i = *v;
	return NC_NOERR;
}

int nc_put_att_text(int ncid, int varid, const char *name, size_t len, const char *v)
{
	ATT *a = new_att(name, 't');
	(void)ncid; (void)varid;
	if (!a) return NC_EPFT_STATE;
	a->text = (char *)malloc(len + 1);
	if (!a->text) return NC_EPFT_IO;
	memcpy(a->text, v, len); a->text[len] = 0; a->tlen = len;
	return NC_NOERR;
}

int nc_enddef(int ncid)
{
	int i, k;
	(void)ncid;
	if (!D.f || !D.writing || !D.defining) return NC_EPFT_STATE;
	fprintf(D.f, "PFTSNAP 1\n");
	for (i = 0; i < D.ndims; i++) fprintf(D.f, "dim %s %zu\n", D.dim[i].name, D.dim[i].len);
	for (i = 0; i < D.nvars; i++) {
		fprintf(D.f, "var %s %d", D.var[i].name, D.var[i].ndims);
		for (k = 0; k < D.var[i].ndims; k++) fprintf(D.f, " %d", D.var[i].dimid[k]);
		fprintf(D.f, "\n");
	}
	for (i = 0; i < D.natts; i++) {
		ATT *a = &D.att[i];
		if (a->kind == 'd') fprintf(D.f, "attd %s %.17g\n", a->name, a->d);
		else if (a->kind == 'i') fprintf(D.f, "atti %s %d\n", a->name, a->i);
		else { fprintf(D.f, "attt %s %zu\n", a->name, a->tlen); fwrite(a->text, 1, a->tlen, D.f); fputc('\n', D.f); }
	}
	fprintf(D.f, "end\n");
	D.payload = ftell(D.f);
	D.defining = 0;
	layout();
	/* reserve the payload so that partially written datasets have a defined size */
	if (D.nvars) {
		long total = D.var[D.nvars - 1].offset + (long)(D.var[D.nvars - 1].count * sizeof(double));
		if (total > 0) { fseek(D.f, D.payload + total - 1, SEEK_SET); fputc(0, D.f); }
	}
	return ferror(D.f) ? NC_EPFT_IO : NC_NOERR;
}

/* hyperslab transfer (row-major, arbitrary rank) */
static int slab_io(int varid, const size_t *start, const size_t *count, double *rbuf, const double *wbuf)
{
	VAR *v;
	size_t idx[MAXDIMS] = {0}, stride[MAXDIMS];
	size_t inner, nrows = 1, r;
	int nd, k;
	if (!D.f || D.defining || varid < 0 || varid >= D.nvars) return NC_EPFT_STATE;
	v = &D.var[varid];
	nd = v->ndims;
	stride[nd - 1] = 1;
	for (k = nd - 2; k >= 0; k--) stride[k] = stride[k + 1] * D.dim[v->dimid[k + 1]].len;
	inner = count[nd - 1];
	for (k = 0; k < nd - 1; k++) nrows *= count[k];
	for (r = 0; r < nrows; r++) {
		size_t lin = start[nd - 1];
		for (k = 0; k < nd - 1; k++) lin += (start[k] + idx[k]) * stride[k];
		if (fseek(D.f, D.payload + v->offset + (long)(lin * sizeof(double)), SEEK_SET)) return NC_EPFT_IO;
		if (wbuf) { if (fwrite(wbuf + r * inner, sizeof(double), inner, D.f) != inner) return NC_EPFT_IO; }
		else      { if (fread(rbuf + r * inner, sizeof(double), inner, D.f) != inner) return NC_EPFT_IO; }
		for (k = nd - 2; k >= 0; k--) { if (++idx[k] < count[k]) break; idx[k] = 0; }
	}
	return NC_NOERR;
}

int nc_put_vara_double(int ncid, int varid, const size_t *start, const size_t *count, const double *v)
{
	(void)ncid;
	if (!D.writing) return NC_EPFT_STATE;
	return slab_io(varid, start, count, NULL, v);
}

int nc_put_var_double(int ncid, int varid, const double *v)
{
	size_t start[MAXDIMS] = {0}, count[MAXDIMS];
	int k;
	(void)ncid;
	if (!D.f || D.defining || !D.writing || varid < 0 || varid >= D.nvars) return NC_EPFT_STATE;
	for (k = 0; k < D.var[varid].ndims; k++) count[k] = D.dim[D.var[varid].dimid[k]].len;
	return slab_io(varid, start, count, NULL, v);
}

int nc_close(int ncid)
{
	(void)ncid;
	if (!D.f) return NC_EPFT_STATE;
	if (D.writing && D.defining) nc_enddef(1);
	fclose(D.f);
	D.f = NULL;
	reset();
	return NC_NOERR;
}

/* ------------------------------- read side -------------------------------- */

int nc_open(const char *path, int mode, int *ncid)
{
	char line[512], kw[16];
	(void)mode;
	if (D.f) return NC_EPFT_STATE;
	reset();
	D.f = fopen(path, "rb");
	if (!D.f) return NC_EPFT_IO;
	if (!fgets(line, sizeof line, D.f) || strncmp(line, "PFTSNAP 1", 9)) { fclose(D.f); D.f = NULL; return NC_EPFT_IO; }
	while (fgets(line, sizeof line, D.f)) {
		if (sscanf(line, "%15s", kw) != 1) continue;
		if (!strcmp(kw, "end")) break;
		if (!strcmp(kw, "dim") && D.ndims < MAXDIMS) {
			sscanf(line, "%*s %63s %zu", D.dim[D.ndims].name, &D.dim[D.ndims].len); D.ndims++;
		} else if (!strcmp(kw, "var") && D.nvars < MAXVARS) {
			VAR *v = &D.var[D.nvars];
			int pos = 0, k;
			sscanf(line, "%*s %63s %d%n", v->name, &v->ndims, &pos);
			for (k = 0; k < v->ndims && k < MAXDIMS; k++) { int adv = 0; sscanf(line + pos, "%d%n", &v->dimid[k], &adv); pos += adv; }
			D.nvars++;
		} else if (!strcmp(kw, "attd") && D.natts < MAXATTS) {
			ATT *a = &D.att[D.natts++]; a->kind = 'd'; sscanf(line, "%*s %63s %lf", a->name, &a->d);
		} else if (!strcmp(kw, "atti") && D.natts < MAXATTS) {
			ATT *a = &D.att[D.natts++]; a->kind = 'i'; sscanf(line, "%*s %63s %d", a->name, &a->i);
		} else if (!strcmp(kw, "attt")) {
			char nm[NAMELEN]; size_t n = 0;
			sscanf(line, "%*s %63s %zu", nm, &n);
			fseek(D.f, (long)n + 1, SEEK_CUR);
		}
	}
	D.payload = ftell(D.f);
	layout();
	*ncid = 1;
	return NC_NOERR;
}

int nc_inq_varid(int ncid, const char *name, int *varid)
{
	int i; (void)ncid;
	for (i = 0; i < D.nvars; i++) if (!strcmp(D.var[i].name, name)) { *varid = i; return NC_NOERR; }
	return NC_EPFT_NOTFOUND;
}

int nc_inq_dimid(int ncid, const char *name, int *dimid)
{
	int i; (void)ncid;
	for (i = 0; i < D.ndims; i++) if (!strcmp(D.dim[i].name, name)) { *dimid = i; return NC_NOERR; }
	return NC_EPFT_NOTFOUND;
}

int nc_inq_dimlen(int ncid, int dimid, size_t *len)
{
	(void)ncid;
	if (dimid < 0 || dimid >= D.ndims) return NC_EPFT_NOTFOUND;
	*len = D.dim[dimid].len;
	return NC_NOERR;
}

static ATT *find_att(const char *name, char kind)
{
	int i;
	for (i = 0; i < D.natts; i++) if (D.att[i].kind == kind && !strcmp(D.att[i].name, name)) return &D.att[i];
	return NULL;
}

int nc_get_att_int(int ncid, int varid, const char *name, int *v)
{
	ATT *a = find_att(name, 'i'); (void)ncid; (void)varid;
	if (!a) return NC_EPFT_NOTFOUND;
	*v = a->i; return NC_NOERR;
}

int nc_get_att_double(int ncid, int varid, const char *name, double *v)
{
	ATT *a = find_att(name, 'd'); (void)ncid; (void)varid;
	if (!a) return NC_EPFT_NOTFOUND;
	*v = a->d; return NC_NOERR;
}

int nc_get_vara_double(int ncid, int varid, const size_t *start, const size_t *count, double *v)
{
	(void)ncid;
	return slab_io(varid, start, count, v, NULL);
}
