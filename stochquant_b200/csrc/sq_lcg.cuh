// sq_lcg.cuh -- the reference's shared-seed 48-bit LCG as a counter-based generator.
//
// Reference: /root/reference/tau_kernel.cl:269-284 (`random`).  Every work-item of a
// tau-step read-modify-writes ONE global seed; under the canonical as-if-sequential
// order (gid 0,1,...,N within a step) a draw at gid g maps the seed
//     t1 = (A(s+g)+B) mod 2^48,  t2 = (A(t1+g)+B) mod 2^48,  s' = t2 - 2^31
// i.e. affinely:  s' = ALPHA*s + BETA*g + GAMMA  (mod 2^48), except for two rare
// data-dependent events: the inf-retry (t1>>16 == 0, :282) and the `*seed += temp`
// branch (s < 2^31 && t2 < 2^31, :278-279).  D consecutive draws from gid g0:
//     s_D = ALPHA^D s + (BETA g0 + GAMMA) G0(D) + BETA G1(D)
// which gives O(1) jump-ahead from small tables.  Events are detected on the device
// and replayed literally on the host (sq_api.cu), so the integer stream is bit-exact.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SQ_HD __host__ __device__ __forceinline__
#else
#define SQ_HD inline
#endif

namespace sq {

typedef unsigned long long u64;

constexpr u64 LCG_A = 0x5DEECE66DULL;
constexpr u64 LCG_B = 0xBULL;
constexpr u64 LCG_MASK = (1ULL << 48) - 1;
constexpr u64 TWO31 = 1ULL << 31;
constexpr u64 LCG_ALPHA = (LCG_A * LCG_A) & LCG_MASK;                          // A^2
constexpr u64 LCG_BETA = (LCG_A * LCG_A + LCG_A) & LCG_MASK;                   // A^2 + A
constexpr u64 LCG_GAMMA = (LCG_A * LCG_B + LCG_B - TWO31) & LCG_MASK;          // AB + B - 2^31

// one table entry = the jump over D draws: (ALPHA^D, G0(D), BETA*G1(D)) mod 2^64
struct JumpEntry {
    u64 a, g0, bg1;
};
constexpr int JUMP_LEVELS = 5;  // D < 2^40
constexpr int JUMP_RADIX = 256;
constexpr int JUMP_TABLE_ENTRIES = JUMP_LEVELS * JUMP_RADIX;

// first and second LCG output of the draw at gid g with seed s (tau_kernel.cl:273,275)
SQ_HD void lcg_draw(u64 s, u64 g, u64 &t1, u64 &t2) {
    t1 = ((s + g) * LCG_A + LCG_B) & LCG_MASK;
    t2 = ((t1 + g) * LCG_A + LCG_B) & LCG_MASK;
}
// seed after an event-free draw (:281); only its low 48 bits matter downstream
SQ_HD u64 lcg_next_seed(u64 t2) { return t2 - TWO31; }

// necessary condition for either event; sm = seed before the draw, masked to 48 bits
SQ_HD bool lcg_event(u64 sm, u64 t1, u64 t2) {
    return ((t1 >> 16) == 0) || (sm < TWO31 && t2 < TWO31);
}

// apply one table entry: D draws starting at gid g with seed s
SQ_HD u64 lcg_apply(const JumpEntry &e, u64 s, u64 g) {
    return e.a * s + (LCG_BETA * g + LCG_GAMMA) * e.g0 + e.bg1;
}

// seed before the draw at gid g+D given seed s before the draw at gid g (no events)
SQ_HD u64 lcg_seed_at(u64 s, u64 g, u64 D, const JumpEntry *__restrict__ tab) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int L = 0; L < JUMP_LEVELS; ++L) {
        const unsigned j = (unsigned)(D >> (8 * L)) & 255u;
        if (j) {
            const JumpEntry e = tab[L * JUMP_RADIX + j];
            s = lcg_apply(e, s, g);
            g += (u64)j << (8 * L);
        }
    }
    return s & LCG_MASK;
}

// ---- host-side table construction (independent of the oracle's implementation) ----
struct JumpTriple {
    u64 a, g0, g1, d;
};
inline JumpTriple jump_compose(const JumpTriple &p, const JumpTriple &q) {  // p first, then q
    JumpTriple r;
    r.a = p.a * q.a;
    r.g0 = q.a * p.g0 + q.g0;
    r.g1 = q.a * p.g1 + p.d * q.g0 + q.g1;
    r.d = p.d + q.d;
    return r;
}
inline JumpTriple jump_triple(u64 D) {
    JumpTriple res{1, 0, 0, 0}, pw{LCG_ALPHA, 1, 0, 1};
    while (D) {
        if (D & 1) res = jump_compose(res, pw);
        pw = jump_compose(pw, pw);
        D >>= 1;
    }
    return res;
}
inline JumpEntry jump_entry(u64 D) {
    JumpTriple t = jump_triple(D);
    return JumpEntry{t.a, t.g0, LCG_BETA * t.g1};
}
inline void build_jump_table(JumpEntry *tab /* JUMP_TABLE_ENTRIES */) {
    for (int L = 0; L < JUMP_LEVELS; ++L) {
        const JumpTriple unit = jump_triple(1ULL << (8 * L));
        JumpTriple acc{1, 0, 0, 0};
        for (int j = 0; j < JUMP_RADIX; ++j) {
            tab[L * JUMP_RADIX + j] = JumpEntry{acc.a, acc.g0, LCG_BETA * acc.g1};
            acc = jump_compose(acc, unit);
        }
    }
}

// process-wide host copy of the table (C++11 magic static: built once, thread-safe)
inline const JumpEntry *host_jump_table() {
    struct Holder {
        JumpEntry t[JUMP_TABLE_ENTRIES];
        Holder() { build_jump_table(t); }
    };
    static const Holder h;
    return h.t;
}

// start seed S' at gid 0 whose event-free chain passes through seed `bs` before the draw at gid bg
inline u64 virtual_start_seed(u64 bs, u64 bg, const JumpEntry *tab) {
    const u64 c0 = lcg_seed_at(0, 0, bg, tab);
    const u64 a = (lcg_seed_at(1, 0, bg, tab) - c0) & LCG_MASK;  // alpha^bg, odd
    u64 inv = a;                                                 // Newton: 3 -> 6 -> ... -> 96 bits
    for (int i = 0; i < 5; ++i) inv *= 2 - a * inv;
    return ((bs - c0) * inv) & LCG_MASK;
}

// literal host replay of tau_kernel.cl:269-284 without the floating-point part:
// the do/while repeats exactly when t1>>16 == 0 (v1 == 0 -> log = -inf -> result inf).
struct HostDraw {
    u64 t1, t2, seed_after;
    int ndraws, plus;
};
inline HostDraw host_draw_literal(u64 seed_full, u64 gid) {
    HostDraw h{0, 0, seed_full, 0, 0};
    u64 temp;
    do {
        temp = ((h.seed_after + gid) * LCG_A + LCG_B) & LCG_MASK;
        h.t1 = temp;
        temp = ((temp + gid) * LCG_A + LCG_B) & LCG_MASK;
        h.t2 = temp;
        if (h.seed_after < TWO31 && temp < TWO31) {
            h.seed_after += temp;
            h.plus = 1;
        } else {
            h.seed_after = temp - TWO31;
            h.plus = 0;
        }
        h.ndraws++;
    } while ((h.t1 >> 16) == 0);
    return h;
}

}  // namespace sq
