/* Stand-in for <rpc/types.h> (libtirpc is absent from this image). Test infrastructure only. */
#ifndef SHIM_RPC_TYPES_H
#define SHIM_RPC_TYPES_H
#include <stdio.h>
#include <sys/types.h>
typedef int bool_t;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#endif
