/* Big-endian stdio XDR stand-in (see rpc/xdr.h). Test infrastructure only. */
#include <string.h>
#include <stdint.h>
#include "rpc/xdr.h"

void xdrstdio_create(XDR *xdrs, FILE *fp, enum xdr_op op) {
    xdrs->x_op = op;
    xdrs->fp = fp;
}

static bool_t xfer(XDR *xdrs, void *p, int n) {
    unsigned char b[8], *q = (unsigned char *)p;
    int i;
    if (xdrs->x_op == XDR_ENCODE) {
        for (i = 0; i < n; ++i) b[i] = q[n - 1 - i];
        return fwrite(b, 1, n, xdrs->fp) == (size_t)n;
    }
    if (xdrs->x_op == XDR_DECODE) {
        if (fread(b, 1, n, xdrs->fp) != (size_t)n) return FALSE;
        for (i = 0; i < n; ++i) q[i] = b[n - 1 - i];
        return TRUE;
    }
    return TRUE;
}
bool_t xdr_int(XDR *x, int *p) { return xfer(x, p, 4); }
bool_t xdr_u_int(XDR *x, unsigned int *p) { return xfer(x, p, 4); }
bool_t xdr_float(XDR *x, float *p) { return xfer(x, p, 4); }
bool_t xdr_double(XDR *x, double *p) { return xfer(x, p, 8); }
bool_t xdr_longlong_t(XDR *x, longlong_t *p) { return xfer(x, p, 8); }
bool_t xdr_setpos(XDR *x, unsigned int pos) { return fseek(x->fp, (long)pos, SEEK_SET) == 0; }
unsigned int xdr_getpos(XDR *x) { return (unsigned int)ftell(x->fp); }
