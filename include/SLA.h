/* Drop-in shim: programs written against the reference's src/include/public/SLA.h
 * compile unchanged against libsla_b200.so.  Everything is declared in sla_b200.h. */
#ifndef SLAB200_SHIM_SLA_H
#define SLAB200_SHIM_SLA_H
#include "sla_b200.h"
#endif
