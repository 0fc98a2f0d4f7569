/*
 * oracle/shim/netcdf.h -- stand-in for <netcdf.h>.  TEST INFRASTRUCTURE ONLY.
 *
 * libnetcdf is not installed in the image, so the unmodified reference driver
 * (intertrack.c) is compiled against this header.  Only the nc_* entry points
 * that intertrack.c calls are provided (see the call sites at
 * apps/intertrack-hybrid-S-freezing/intertrack.c:1313,1603-1671,2040-2084,
 * 2331-2548).  The implementation (shim_rt.c) stores a dataset in the small
 * self-describing "PFTSNAP 1" container documented in shim_rt.c, which keeps the
 * reference's dataset layout (dims n3,n2,n1; coordinate variables; u,p,gl as
 * double[n3][n2][n1]; global attributes) and supports the read side used by
 * `set icond_file` / `continue_series`.
 */
#ifndef PFT_ORACLE_SHIM_NETCDF_H
#define PFT_ORACLE_SHIM_NETCDF_H

#include <stddef.h>

#define NC_NOERR            0
#define NC_EPFT_IO        (-31)
#define NC_EPFT_NOTFOUND  (-49)
#define NC_EPFT_STATE     (-39)

#define NC_NOWRITE          0
#define NC_CLOBBER          0
#define NC_FORMAT_NETCDF4   3
#define NC_INT              4
#define NC_DOUBLE           6
#define NC_GLOBAL         (-1)

int nc_set_default_format(int format, int *old_format);
const char *nc_strerror(int err);

int nc_create(const char *path, int mode, int *ncid);
int nc_open(const char *path, int mode, int *ncid);
int nc_close(int ncid);
int nc_enddef(int ncid);

int nc_def_dim(int ncid, const char *name, size_t len, int *dimid);
int nc_def_var(int ncid, const char *name, int type, int ndims, const int *dimids, int *varid);
int nc_put_att_double(int ncid, int varid, const char *name, int type, size_t len, const double *v);
int nc_put_att_int(int ncid, int varid, const char *name, int type, size_t len, const int *v);
int nc_put_att_text(int ncid, int varid, const char *name, size_t len, const char *v);
int nc_put_var_double(int ncid, int varid, const double *v);
int nc_put_vara_double(int ncid, int varid, const size_t *start, const size_t *count, const double *v);

int nc_inq_varid(int ncid, const char *name, int *varid);
int nc_inq_dimid(int ncid, const char *name, int *dimid);
int nc_inq_dimlen(int ncid, int dimid, size_t *len);
int nc_get_att_int(int ncid, int varid, const char *name, int *v);
int nc_get_att_double(int ncid, int varid, const char *name, double *v);
int nc_get_vara_double(int ncid, int varid, const size_t *start, const size_t *count, double *v);

#endif /* PFT_ORACLE_SHIM_NETCDF_H */
