// sq_strip_slow.cuh -- rare paths shared by the fp32 streaming kernels (sq_march.cu, sq_tile.cu): a strip that
// contains a replay entry's first site or its overridden site, the position of a strip relative to the step's
// replay entries, the clamp (tau_kernel.cl:122-132) on values the hot path could not prove in range.
#pragma once
#include "sq_lattice_common.cuh"
#include "sq_pair.cuh"

namespace sq {
namespace {

struct Draws {
    unsigned u1[4], u2[4];
    u64 s_after;  // seed behind the four draws (48 bits)
};

// generic draws of one strip under the step's replay entries (a strip that contains an entry's
// gid_start or its overridden site): literal 64-bit chain with overrides, as in lattice_step_kernel
__device__ __noinline__ Draws draws_slow(const RebaseEntry *rebase, int n_rebase, u64 *event_key_ptr, int step_index, int chain,
                                         u64 s, u64 g0) {
    Draws d;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const u64 g = g0 + e;
        u64 t1, t2;
        bool overridden = false;
        for (int j = 0; j < n_rebase; ++j)
            if (rebase[j].chain == chain && rebase[j].gid_start == g) s = rebase[j].seed;
        lcg_draw(s, g, t1, t2);
        for (int j = 0; j < n_rebase; ++j)
            if (rebase[j].chain == chain && rebase[j].ov_gid == g) {
                t1 = rebase[j].ov_t1;
                t2 = rebase[j].ov_t2;
                overridden = true;
            }
        if (!overridden && lcg_event(s & LCG_MASK, t1, t2))
            atomicMin((unsigned long long *)event_key_ptr, event_key(step_index, chain, g));
        s = lcg_next_seed(t2) & LCG_MASK;
        d.u1[e] = (unsigned)(t1 >> 16);
        d.u2[e] = (unsigned)(t2 >> 16);
    }
    d.s_after = s;
    return d;
}

// A whole strip on the rare path (a replay entry's first site or its overridden site lies inside it):
// scalar code, same operations and roundings per site as the packed hot path.
struct SlowIn {
    const float *cur, *tm, *tp;
    float *dst, *push0, *push1;
    unsigned o, o_up1, o_dn1, o_up2, o_dn2, o_left, o_right;  // offsets inside the slice
    u64 s, g0;
    int chain, step_index, n_rebase;
    const RebaseEntry *rebase;
    u64 *event_key;
    float c_lap, c_dt, m2, lam, k2;
    double nscale;
};
struct SlowOut {
    float a1, a2;
    unsigned nclamp;
    u64 s_after;  // seed behind the strip's four draws: a wider strip continues from here
};
template <int MATH, int NDIM, int POT>
__device__ __noinline__ SlowOut strip_slow(const SlowIn I) {
    const Draws d = draws_slow(I.rebase, I.n_rebase, I.event_key, I.step_index, I.chain, I.s, I.g0);
    const float4 c = *reinterpret_cast<const float4 *>(I.cur + I.o);
    const float4 u1 = *reinterpret_cast<const float4 *>(I.cur + I.o_up1), d1 = *reinterpret_cast<const float4 *>(I.cur + I.o_dn1);
    float4 u2 = make_float4(0, 0, 0, 0), d2 = u2;
    if (NDIM >= 4) {
        u2 = *reinterpret_cast<const float4 *>(I.cur + I.o_up2);
        d2 = *reinterpret_cast<const float4 *>(I.cur + I.o_dn2);
    }
    const float4 tp = *reinterpret_cast<const float4 *>(I.tp + I.o), tm = *reinterpret_cast<const float4 *>(I.tm + I.o);
    const float left = I.cur[I.o_left], right = I.cur[I.o_right];
    const float cc[4] = {c.x, c.y, c.z, c.w}, xp[4] = {c.y, c.z, c.w, right}, xm[4] = {left, c.x, c.y, c.z};
    const float nu1[4] = {u1.x, u1.y, u1.z, u1.w}, nd1[4] = {d1.x, d1.y, d1.z, d1.w}, nu2[4] = {u2.x, u2.y, u2.z, u2.w},
                nd2[4] = {d2.x, d2.y, d2.z, d2.w}, ntp[4] = {tp.x, tp.y, tp.z, tp.w}, ntm[4] = {tm.x, tm.y, tm.z, tm.w};
    const float kth = (float)(2.0 * 3.1415 / 4294967296.0);
    SlowOut r;
    r.s_after = d.s_after;
    r.a1 = 0.f;
    r.a2 = 0.f;
    r.nclamp = 0;
    float out[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float phi = cc[e];
        float sum = __fadd_rn(xp[e], xm[e]);
        sum = __fadd_rn(sum, nu1[e]);
        sum = __fadd_rn(sum, nd1[e]);
        if (NDIM >= 4) {
            sum = __fadd_rn(sum, nu2[e]);
            sum = __fadd_rn(sum, nd2[e]);
        }
        sum = __fadd_rn(sum, ntp[e]);
        sum = __fadd_rn(sum, ntm[e]);
        float v = __fmaf_rn(I.c_lap, __fmaf_rn(-(float)(2 * NDIM), phi, sum), phi);
        if (POT == 4) v = __fmaf_rn(-I.c_dt, __fmul_rn(phi, __fmaf_rn(I.lam, __fmul_rn(phi, phi), I.m2)), v);
        else v = __fmaf_rn(-2.0f * I.c_dt, phi, v);
        if (MATH == 1) {
            const float a = __fmul_rn(__uint2float_rn(d.u1[e]), 2.3283064365386963e-10f);
            const float t = __fmul_rn(lg2_approx(a), I.k2);
            const float th = __fmaf_rn(__uint2float_rn(d.u2[e]), kth, -3.14159265358979f);
            v = __fsub_rn(v, __fmul_rn(__cosf(th), sqrt_approx(fabsf(t))));  // dw = -RN(cos rad), then v + dw
        } else {
            v = __fadd_rn(v, (float)__dmul_rn(I.nscale, noise_accurate((u64)d.u1[e] << 16, (u64)d.u2[e] << 16)));
        }
        r.nclamp += (fabsf(v) <= 1000.0f) ? 0u : 1u;
        out[e] = (v < 1000.0f) ? ((v > -1000.0f) ? v : -1000.0f) : 1000.0f;
        r.a1 = __fadd_rn(r.a1, phi);
        r.a2 = __fmaf_rn(phi, phi, r.a2);
    }
    const float4 res = make_float4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<float4 *>(I.dst + I.o) = res;
    if (I.push0) *reinterpret_cast<float4 *>(I.push0 + I.o) = res;
    if (I.push1) *reinterpret_cast<float4 *>(I.push1 + I.o) = res;
    return r;
}

// Replay entries (REBASE): where does a strip stand relative to the step's (sorted) entries?  Every
// entry carries a virtual step-start seed (RebaseEntry::vseed) under which the kernel's ordinary
// gid-0-based jump tables and row recurrence stay valid behind it, so the hot loop only watches the
// distance to the thread's NEXT entry; this runs once per thread and when an entry is reached.
struct Rebased {
    u64 S_eff;       // step-start seed whose event-free chain is valid at this strip
    unsigned cnt;    // entries at or before the strip
    unsigned nxt32;  // slice-relative offset of the next entry behind this strip (0x7FFFFFFF: none in this slice)
    bool slow;       // an entry's gid_start or overridden site (= gid_start - 1) lies inside the strip
};
__device__ __forceinline__ Rebased rebase_eval(const RebaseEntry *rebase, int n_rebase, int chain, u64 S, u64 g0, u64 gslice,
                                               unsigned vs, unsigned w = 4u /* sites per strip */) {
    Rebased r;
    r.S_eff = S;
    r.cnt = 0;
    r.slow = false;
    u64 nxt = ~0ULL, bg = 0;
    for (int j = 0; j < n_rebase; ++j) {
        const u64 gs = rebase[j].gid_start;
        if (rebase[j].chain != chain) continue;
        if (gs <= g0) {
            r.cnt++;
            if (gs >= bg) { bg = gs; r.S_eff = rebase[j].vseed; }
        }
        r.slow |= (gs - g0 <= (u64)w);
        if (gs > g0 + w && gs < nxt) nxt = gs;
    }
    r.nxt32 = (nxt - gslice < (u64)vs) ? (unsigned)(nxt - gslice) : 0x7FFFFFFFu;
    return r;
}

struct Clamped {
    float v[4];
    unsigned n;
};
// tau_kernel.cl:122-132 on the cold path: clamp to [-1000, 1000], inf/NaN -> +1000, count the hits
__device__ __noinline__ Clamped clamp_cold(float a, float b, float c, float d) {
    Clamped r;
    const float in[4] = {a, b, c, d};
    r.n = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float v = in[e];
        r.n += (fabsf(v) <= 1000.0f) ? 0u : 1u;  // NaN counts
        r.v[e] = (v < 1000.0f) ? ((v > -1000.0f) ? v : -1000.0f) : 1000.0f;
    }
    return r;
}

}  // namespace
}  // namespace sq
