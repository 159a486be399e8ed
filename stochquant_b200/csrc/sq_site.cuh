// sq_site.cuh -- the per-site inner pipeline shared by the lattice kernels, written for issue
// slots: the fused Langevin step is instruction-bound on B200 (ncu, profiles/), not HBM-bound.
//
//   RNG   (tau_kernel.cl:273-281)  two 48-bit LCG steps as 32-bit limbs: 12 integer ops / site
//   noise (tau_kernel.cl:274-277)  FAST: 4 SFU ops + ~8 FMA/ALU ops;  ACCURATE: reference casts
//   events                         detected for ~2 ops / site, resolved on a cold path
//
// The 48-bit seed s is carried as (sl, sh); only bits 0..47 are meaningful, higher bits of sh are
// garbage that never reaches a result (every consumer takes bits <= 47).
#pragma once
#include "sq_lcg.cuh"
#include "sq_noise.cuh"

namespace sq {

constexpr unsigned A_LO = 0xDEECE66Du, A_HI = 0x5u;

struct Seed32 {
    unsigned lo, hi;
};
__device__ __forceinline__ Seed32 seed_split(u64 s) { return Seed32{(unsigned)s, (unsigned)(s >> 32)}; }
__device__ __forceinline__ u64 seed_join(Seed32 s) { return (((u64)s.hi << 32) | s.lo) & LCG_MASK; }

// x*A + c  (mod 2^64 in the low word, bits 32..47 valid in the high word); c stays a native
// 64-bit value so the IMAD.WIDE addend needs no register-pair packing.  Inline PTX on purpose:
// left to the optimiser, a chain of these is re-associated into independent polynomials in the
// seed with a dozen loop-carried constants -- twice the integer instructions (profiles/).
__device__ __forceinline__ void mad48(unsigned xl, unsigned xh, u64 c, unsigned &rl, unsigned &rh) {
    u64 p;
    unsigned ph, t;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(xl), "r"(A_LO), "l"(c));  // IMAD.WIDE.U32
    asm("mov.b64 {%0, %1}, %2;" : "=r"(rl), "=r"(ph) : "l"(p));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(xl), "r"(A_HI), "r"(ph));     // IMAD
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(rh) : "r"(xh), "r"(A_LO), "r"(t));     // IMAD
}

// One draw at the site whose constant is c = gid*A + B.  Returns u1 = t1>>16, u2 = t2>>16 (32 bits
// each) and advances the seed to t2 - 2^31 (event-free path, :281).
__device__ __forceinline__ void site_draw(Seed32 &s, u64 c, unsigned &u1, unsigned &u2) {
    unsigned t1l, t1h, t2l, t2h;
    mad48(s.lo, s.hi, c, t1l, t1h);           // t1 = (s+g)A + B = sA + c
    u1 = __funnelshift_r(t1l, t1h, 16);
    mad48(t1l, t1h, c, t2l, t2h);             // t2 = (t1+g)A + B = t1 A + c
    u2 = __funnelshift_r(t2l, t2h, 16);
    // t2 - 2^31 = t2 + 2^31 - 2^32: carry chain, 2 instructions
    asm("add.cc.u32 %0, %2, 0x80000000;\n\taddc.u32 %1, %3, 0xffffffff;" : "=r"(s.lo), "=r"(s.hi) : "r"(t2l), "r"(t2h));
}
// c += A (next site's constant), kept in IMAD.WIDE form so that c lives in an aligned register
// pair and feeds the next mad.wide addend without moves: c + A_LO (carry into the high word), +5.
// The same two functions left to the compiler (the form the resident kernel's short 2-site chains
// compile best from: 13 integer instructions per site there, against 25 with the PTX form above,
// while a 4-site chain in this form is expanded into polynomials by the optimiser -- measured both ways).
__device__ __forceinline__ void mad48_c(unsigned xl, unsigned xh, u64 c, unsigned &rl, unsigned &rh) {
    const u64 p = (u64)xl * A_LO + c;                  // IMAD.WIDE.U32 with 64-bit addend
    rl = (unsigned)p;
    rh = (unsigned)(p >> 32) + xl * A_HI + xh * A_LO;  // 2 x IMAD
}
__device__ __forceinline__ void site_draw_c(Seed32 &s, u64 c, unsigned &u1, unsigned &u2) {
    unsigned t1l, t1h, t2l, t2h;
    mad48_c(s.lo, s.hi, c, t1l, t1h);
    u1 = __funnelshift_r(t1l, t1h, 16);
    mad48_c(t1l, t1h, c, t2l, t2h);
    u2 = __funnelshift_r(t2l, t2h, 16);
    s.lo = t2l + 0x80000000u;                 // t2 - 2^31
    s.hi = t2h - (t2l < 0x80000000u ? 1u : 0u);
}

// `one` must be opaque to ptxas (see opaque_one): a literal 1 is strength-reduced to a carry-chain add
// whose result is NOT an aligned pair, and every following mad.wide then splits into 4-5 instructions.
__device__ __forceinline__ unsigned opaque_one() { return blockDim.x >> 8; }  // ONLY for kernels launched with 256..511 threads
__device__ __forceinline__ u64 site_const_next(u64 c, unsigned one) {
    u64 r;
    unsigned lo, hi;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(one), "r"(A_LO), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(r));
    hi += A_HI;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ u64 site_const(u64 gid) { return gid * LCG_A + LCG_B; }

// Cheap necessary condition for an RNG event at this draw, given what the hot path has anyway:
//   inf-retry  <=> u1 == 0
//   `seed+=`    => t2 < 2^31 <=> u2 < 2^15
__device__ __forceinline__ bool site_maybe_event(unsigned u1, unsigned u2) { return (u1 == 0u) | (u2 < 32768u); }

// FAST noise, scalar form of what the packed kernels compute (sq_rowres.cu, sq_tile.cu, sq_march.cu) -- the SAME operations
// and roundings, so a FAST fp32 run does not depend on which kernel took which step:
//   v1 = (float)u1 * 2^-32 (exact scaling of the RN conversion), lg = lg2.approx(v1), rad = sqrt.approx(|lg * k2|)
//   (k2 = 2 ln2 * scale^2: the amplitude folded under the square root), theta - pi = fma((float)u2, 2*3.1415*2^-32, -pi),
//   dw = -RN(cos.approx(theta - pi) * rad)          [cos(theta) = -cos(theta - pi) keeps MUFU.COS in [-pi, pi)]
// The caller adds dw with one more rounding.
__device__ __forceinline__ float site_noise_fast(unsigned u1, unsigned u2, float k2) {
    const float v1 = __fmul_rn(__uint2float_rn(u1), 2.3283064365386963e-10f);
    float lg;                                           // MUFU.LG2 (v1 >= 2^-32: never denormal)
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(v1));
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(fabsf(__fmul_rn(lg, k2))));
    const float th = __fmaf_rn(__uint2float_rn(u2), (float)(2.0 * 3.1415 / 4294967296.0), -3.14159265358979f);
    return -__fmul_rn(__cosf(th), rad);
}

}  // namespace sq
