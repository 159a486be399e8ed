// sq_noise.cuh -- Box-Muller (cos branch) of tau_kernel.cl:274-277 from the two LCG outputs.
//
//   v1 = (double)(t1>>16) / 2^32 ; v2 = (double)(t2>>16) / 2^32
//   r  = (double)cos((float)(2.*3.1415*v2)) * (double)sqrt((float)(-2.*(double)log((float)v1)))
//
// ACCURATE keeps every cast of the reference and uses CUDA's cosf/logf/sqrtf
// (<=2 ulp, the OpenCL spec allows <=4).  FAST stays in fp32 and uses the SFU:
// the integer inputs u1=t1>>16, u2=t2>>16 are still the bit-exact stream.
#pragma once
#include "sq_lcg.cuh"

namespace sq {

// reference-literal, returns the double `result` of :277
__device__ __forceinline__ double noise_accurate(u64 t1, u64 t2) {
    const double v1 = (double)(unsigned)(t1 >> 16) * (1.0 / 4294967296.0);  // exact: /2^32
    const double v2 = (double)(unsigned)(t2 >> 16) * (1.0 / 4294967296.0);
    const float ang = (float)__dmul_rn(__dmul_rn(2., 3.1415), v2);  // 2.*3.1415 folds to one double
    const float lg = logf((float)v1);
    const float rad = (float)__dmul_rn(-2., (double)lg);
    return __dmul_rn((double)cosf(ang), (double)sqrtf(rad));
}

// u32 -> f32 with round-to-nearest, on the FMA/ALU pipes instead of the conversion unit:
// hi*65536+lo is exact in real arithmetic, the single FFMA rounds once == cvt.rn.f32.u32.
__device__ __forceinline__ float u32_to_f32_rn(unsigned u) {
    const float hi = __uint_as_float(0x4B000000u | (u >> 16)) - 8388608.0f;
    const float lo = __uint_as_float(0x4B000000u | (u & 0xFFFFu)) - 8388608.0f;
    return __fmaf_rn(hi, 65536.0f, lo);
}

// FAST: r = cos(2*3.1415*v2) * sqrt(-2 ln v1), fp32 + SFU.
//   ln v1 = ln2 * log2(v1), v1 = (float)u1 * 2^-32 (exact scaling): MUFU.LG2 directly on v1 keeps
//   its 2^-22 absolute error where v1 -> 1 (no "32 - log2(u1)" cancellation).
//   cos(theta), theta = 2*3.1415*v2 in [0, 6.283): MUFU.COS wants |x| <= pi for its
//   2^-21.4 abs error, so evaluate -cos(theta - pi).
// Error vs the reference-literal value: |dr| <~ 1e-6 typically; for draws with v1 > 1-2^-12
// (p = 2.4e-4) the sqrt of a small argument amplifies the SFU's absolute log error: |dr| <= ~5e-5,
// worst case ~1e-3 when v1 = 1-2^-32.
__device__ __forceinline__ float noise_fast(u64 t1, u64 t2) {
    const unsigned u1 = (unsigned)(t1 >> 16), u2 = (unsigned)(t2 >> 16);
    const float v1 = u32_to_f32_rn(u1) * 2.3283064365386963e-10f;  // == (float)v1 of :274
    const float f2 = u32_to_f32_rn(u2);
    // clamp at 0: v1 may round to 1.0f and the SFU may return a tiny positive log there
    const float m2l = fmaxf(0.0f, __log2f(v1) * -1.3862943611198906f);  // -2 ln2 log2 v1
    const float rad = __fsqrt_rn(m2l);
    const float th = __fmaf_rn(f2, (float)(2. * 3.1415 / 4294967296.0), -3.14159265358979f);
    return -__cosf(th) * rad;
}

// ACCURATE noise of one site, scaled and rounded to fp32, OUT OF LINE: the expression inlines to ~150 instructions (libdevice
// cosf / logf with their slow paths); eight to sixteen copies per loop body pushed the fp32 lattice kernels past the
// instruction cache (on-chip kernel: 84 -> 201 G site-updates/s when it became a call).  The streaming kernels keep it inline:
// with their ~70 live registers around the call the out-of-line form was measured slower (64^4: 98 -> 108 us per step).
static __device__ __noinline__ float noise_accurate_dw(unsigned u1, unsigned u2, double nscale) {
    return (float)__dmul_rn(nscale, noise_accurate((u64)u1 << 16, (u64)u2 << 16));
}

}  // namespace sq
