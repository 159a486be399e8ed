/*
 * cl_kernel_shim.h -- ORACLE SUPPORT (test infrastructure, NOT product code).
 *
 * Lets gcc compile the reference's OpenCL C kernel file, *where it lies*
 * (/root/reference/tau_kernel.cl), as a plain C translation unit:
 *     gcc -x c -include cl_kernel_shim.h /root/reference/tau_kernel.cl
 * Nothing from the reference is copied; this header only supplies what the
 * OpenCL C language/runtime would: address-space qualifiers, work-item
 * functions, barrier(), and type-overloaded math built-ins (OpenCL resolves
 * sqrt((float)x) to the float overload; C would promote to double).
 * Float built-ins map to glibc's correctly-rounded-or-1ulp libm, an OpenCL
 * conformant choice (spec allows <=3-4 ulp for these functions).
 */
#ifndef CL_KERNEL_SHIM_H
#define CL_KERNEL_SHIM_H
#include <math.h>

typedef unsigned long ulong;
typedef unsigned int uint;

#define __kernel
#define __global
#define __local
#define __constant const
#define CLK_GLOBAL_MEM_FENCE 2
#define CLK_LOCAL_MEM_FENCE 1

/* work-item context, provided by ref_driver.c */
extern long sq_ref_cur_gid;
extern long sq_ref_gsize;
void sq_ref_barrier(int flags);
static inline int get_global_id(int d) { (void)d; return (int)sq_ref_cur_gid; }
static inline int get_global_size(int d) { (void)d; return (int)sq_ref_gsize; }
static inline int get_local_id(int d) { (void)d; return (int)sq_ref_cur_gid; }
#define barrier(f) sq_ref_barrier(f)

/* the kernel file defines its own `random`; keep it away from stdlib's */
#define random sq_ref_random

/* OpenCL-style overloading of the math built-ins the kernel uses */
#undef isinf
#undef isnan
#define isinf(x) (__builtin_isinf(x) != 0)
#define isnan(x) (__builtin_isnan(x) != 0)
#define sqrt(x) _Generic((x), float: sqrtf, default: sqrt)(x)
#define cos(x) _Generic((x), float: cosf, default: cos)(x)
#define log(x) _Generic((x), float: logf, default: log)(x)
#define tanh(x) _Generic((x), float: tanhf, default: tanh)(x)
#define exp(x) _Generic((x), float: expf, default: exp)(x)
#define pow(x, y) _Generic((x), float: powf, default: pow)(x, y)
/* pown(float,int): the kernel only uses n=2 (a square: one rounded multiply)
 * and exact powers of two */
static inline float sq_ref_pown(float x, int n)
{
    if (n == 2) return x * x;
    return powf(x, (float)n);
}
#define pown(x, n) sq_ref_pown((x), (n))

#endif
