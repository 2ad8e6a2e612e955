/*
 * oracle/clshim.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A keyword shim that lets gcc compile OpenCL C 1.1 device sources as plain
 * C99, so that the reference's own device code (ad.cl and its example kernels
 * under /root/reference) can be executed work-item by work-item on the host.
 * This is the "Oracle B" recipe of SURVEY.md section 8(c).
 *
 * Work-items are driven by a loop in the TU that includes this header; the
 * loop variable is published through clshim_gid[] (thread-local so that the
 * OpenMP baseline can run work-items on several host threads).
 */
#ifndef AD4CL_ORACLE_CLSHIM_H
#define AD4CL_ORACLE_CLSHIM_H

#include <math.h>
#include <stddef.h>

/* address-space / function qualifiers */
#ifndef __kernel
#define __kernel
#endif
#define __global
#define __local
#define __private
#define __constant const

/* the reference enables fp64 through this feature macro (ad.cl:16-22) */
#define cl_khr_fp64 1

/* vector types only appear in typedefs (ad.cl:28-34); never used */
typedef struct { double s[2];  } double2;
typedef struct { double s[3];  } double3;
typedef struct { double s[4];  } double4;
typedef struct { double s[8];  } double8;
typedef struct { double s[16]; } double16;

/* NDRange ids: set by the driving loop */
static __thread int clshim_gid[3];
static __thread int clshim_gsize[3];
#define get_global_id(d)   (clshim_gid[(d)])
#define get_global_size(d) (clshim_gsize[(d)])

/* OpenCL 1.1 32-bit atomics return the OLD value */
#define atomic_inc(p)    __sync_fetch_and_add((p), 1)
#define atomic_dec(p)    __sync_fetch_and_sub((p), 1)
#define atomic_add(p, v) __sync_fetch_and_add((p), (v))

#define CLK_LOCAL_MEM_FENCE  1
#define CLK_GLOBAL_MEM_FENCE 2
/* barrier(): only legal in this serial emulation when it is a no-op */
#define barrier(flags) ((void) 0)

#endif
