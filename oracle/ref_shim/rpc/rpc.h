#ifndef SHIM_RPC_RPC_H
#define SHIM_RPC_RPC_H
#include "types.h"
#include "xdr.h"
#endif
