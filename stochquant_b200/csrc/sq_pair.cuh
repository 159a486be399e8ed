// sq_pair.cuh -- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2) and the SFU wrappers of
// the FAST noise path.  Each lane of a packed instruction rounds on its own (.rn), so a packed
// sequence is bit-identical to the scalar sequence the oracle defines; what it
// saves is issue slots: one instruction per two sites.
#pragma once

namespace sq {
namespace {

typedef unsigned long long pair_t;  // two fp32 in one 64-bit register pair

__device__ __forceinline__ pair_t pk(float lo, float hi) {
    pair_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(pair_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pair_t add2(pair_t a, pair_t b) {
    pair_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pair_t mul2(pair_t a, pair_t b) {
    pair_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pair_t fma2(pair_t a, pair_t b, pair_t c) {
    pair_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}


}  // namespace
}  // namespace sq
