/* Drop-in shim for the reference's src/include/public/SLAStdint.h. */
#ifndef SLAB200_SHIM_STDINT_H
#define SLAB200_SHIM_STDINT_H
#include <stdint.h>
#endif
