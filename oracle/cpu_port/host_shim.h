/* Lets g++ compile a generated pointwise-integrand header (echemfem_b200/compiler.py) for the host.
 * TEST / BASELINE INFRASTRUCTURE ONLY (oracle/): never linked into the product. */
#ifndef ECH_CPU_PORT_SHIM_H
#define ECH_CPU_PORT_SHIM_H
#ifndef _GNU_SOURCE
#define _GNU_SOURCE 1 /* sincos */
#endif
#include <math.h>
#include <cmath>
#include <stdint.h>
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __constant__ static const
using std::fabs;
using std::fmax;
#endif
