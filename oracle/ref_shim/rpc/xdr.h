/*
 * Stand-in for <rpc/xdr.h>: the handful of XDR-over-stdio calls the reference uses for "standard"
 * (big-endian) Tipsy and checkpoint I/O.  Test infrastructure only (libtirpc is absent from this image).
 */
#ifndef SHIM_RPC_XDR_H
#define SHIM_RPC_XDR_H
#include "types.h"
enum xdr_op { XDR_ENCODE = 0, XDR_DECODE = 1, XDR_FREE = 2 };
typedef struct {
    enum xdr_op x_op;
    FILE *fp;
} XDR;
typedef long long longlong_t;
typedef unsigned long long u_longlong_t;
void xdrstdio_create(XDR *xdrs, FILE *fp, enum xdr_op op);
#define xdr_destroy(xdrs) ((void)0)
bool_t xdr_int(XDR *xdrs, int *ip);
bool_t xdr_u_int(XDR *xdrs, unsigned int *ip);
bool_t xdr_float(XDR *xdrs, float *fp);
bool_t xdr_double(XDR *xdrs, double *dp);
bool_t xdr_longlong_t(XDR *xdrs, longlong_t *llp);
bool_t xdr_setpos(XDR *xdrs, unsigned int pos);
unsigned int xdr_getpos(XDR *xdrs);
#endif
