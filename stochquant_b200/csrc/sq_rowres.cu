// sq_rowres.cu -- persistent, on-chip-resident Langevin kernel for 2-D lattices that fit in the register
// files of one B200 (configs[1]: 1024^2 fp32 = 4 MiB over 148 SMs), second design.
//
// Same update and the same integer stream as every other lattice kernel (oracle
// sqo_lattice_step; generalisation of tau_kernel.cl:64-173).  What changed against the first resident
// kernel (round 1: a thread owned 2 columns x all rows of the CTA's band, 4 warps per scheduler, 57
// warp-instructions per site, issue-active 65 %) is the decomposition, chosen for ISSUE SLOTS:
//   * a thread owns EIGHT CONSECUTIVE sites of ONE row (four where the row length is not a multiple of
//     256); the rows of a CTA's band are different threads (7 rows x 128 threads = 28 warps per SM for 1024^2).  Everything that is paid per strip -- chain
//     state, left/right neighbours, the event / clamp tests, row-sum partials -- is paid once per 8 sites
//     instead of once per 2;
//   * the reference's LCG (tau_kernel.cl:269-284) is walked in its t2 form: with T = seed + 2^31,
//         t1_e = A T_{e-1} + c1_e ,   T_e = A^2 T_{e-1} + c2_e ,   u1 = t1 >> 16, u2 = T >> 16,
//     two INDEPENDENT 48-bit multiply-adds per site off the previous T (the literal form chains them);
//   * a step is split into a noise phase (draws + Box-Muller: ~2/3 of the instructions, no field data) and
//     a stencil phase (neighbours from shared memory, update, store).  The CTA barrier between steps is an
//     mbarrier used split-phase: a warp ARRIVES after its stencil phase, runs the noise phase of the NEXT
//     step, and only then WAITS -- barrier skew and the halo round trip through L2 hide behind the
//     noise phase;
//   * band edges travel between CTAs as {float, step tag} 64-bit words (no fence, no flag), read by the
//     boundary rows' threads for their own columns, requested before the barrier wait;
//   * what a strip needs once per step -- its T, the step-advance constant, the two site constants -- lives
//     in shared memory (32 bytes per thread), not in registers: the noise phase loads it, the stencil phase
//     has the registers for its neighbours.
// RNG events (inf-retry / `seed+=`) and the checkpoint / resume protocol are those of round 1: the first
// (step, gid) is raised with atomicMin, every CTA leaves, the host resumes from the last checkpoint every
// CTA has written (sq_api.cu).  Waits are bounded and abort-aware.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "sq_kernels.h"
#include "sq_site.cuh"
#include "sq_pair.cuh"

namespace sq {
namespace {

constexpr unsigned ALPHA_LO32 = (unsigned)LCG_ALPHA, ALPHA_HI32 = (unsigned)(LCG_ALPHA >> 32);

// x*M + c mod 2^48 in 32-bit limbs (bits above 47 of the high word are garbage that never reaches a result).
// Inline PTX: the multiply-add chain must not be re-associated by the optimiser (sq_site.cuh).
__device__ __forceinline__ void mad48k(unsigned xl, unsigned xh, unsigned ml, unsigned mh, u64 c, unsigned &rl, unsigned &rh) {
    u64 p;
    unsigned ph, t;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(xl), "r"(ml), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(rl), "=r"(ph) : "l"(p));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(xl), "r"(mh), "r"(ph));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(rh) : "r"(xh), "r"(ml), "r"(t));
}

__device__ __forceinline__ ulonglong2 ld_words(const unsigned long long *p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_words(unsigned long long *p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ unsigned long long word_of(float v, unsigned tag) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(__float_as_uint(v)), "r"(tag));
    return r;
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    // try_wait suspends the warp in hardware until the phase completes or a time limit passes
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// shared memory through 32-bit shared-window addresses: one register + immediate per access
__device__ __forceinline__ float lds_f32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ ulonglong2 lds_pairs(unsigned a) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_pairs(unsigned a, pair_t x, pair_t y) {
    asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(a), "l"(x), "l"(y) : "memory");
}
__device__ __forceinline__ unsigned lds_f32_bits(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ u64 lds_u64(unsigned a) {
    u64 v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u64(unsigned a, u64 v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// (ACCURATE noise is the out-of-line noise_accurate_dw of sq_noise.cuh: eight inlined copies of logf / cosf / sqrtf and the
// fp64 products per role loop made the kernel 12 k instructions -- no longer resident in the instruction caches.)

// cold: exact event test of the w draws of a strip from its start seed (literal replay is the host's job)
__device__ __noinline__ void stripw_events_cold(u64 *event_key_ptr, int step, u64 sm, u64 g0, int w) {
    for (int e = 0; e < w; ++e) {
        u64 t1, t2;
        lcg_draw(sm, g0 + e, t1, t2);
        if (lcg_event(sm, t1, t2)) atomicMin((unsigned long long *)event_key_ptr, event_key(step, 0, g0 + e));
        sm = lcg_next_seed(t2) & LCG_MASK;
    }
}

// cold: clamp to [-1000, 1000], inf/NaN -> +1000 (tau_kernel.cl:122-132); returns the number of hits
struct Pair2 {
    pair_t a, b;
    unsigned n;
};
__device__ __noinline__ Pair2 clamp4_cold(pair_t a, pair_t b) {
    float v[4];
    upk(a, v[0], v[1]);
    upk(b, v[2], v[3]);
    Pair2 r;
    r.n = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        r.n += (fabsf(v[e]) <= 1000.0f) ? 0u : 1u;  // NaN counts
        v[e] = (v[e] < 1000.0f) ? ((v[e] > -1000.0f) ? v[e] : -1000.0f) : 1000.0f;
    }
    r.a = pk(v[0], v[1]);
    r.b = pk(v[2], v[3]);
    return r;
}

// reducer warps when a CTA has fewer warps than rows (small lattices): the general walk.  rs_par: the parity's array of
// {sum phi, sum phi^2} pairs, one per thread.
__device__ __noinline__ void reduce_rows_general(double *hist_rows, double *hist_p2, int L1, unsigned rs_par, int TPR, int first, int nw,
                                                 int nr, int r0, int step, int lane) {
    for (int row = first; row < nr; row += nw) {
        const unsigned src = rs_par + (unsigned)(row * TPR) * 8u;
        float a = 0.f, q = 0.f;
        for (int i = 2 * lane; i < TPR; i += 64) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + (unsigned)i * 8u));
            a += v.x + v.z;
            q += v.y + v.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) {
            hist_rows[(size_t)step * L1 + r0 + row] = (double)a;
            hist_p2[(size_t)step * L1 + r0 + row] = (double)q;
        }
    }
}

__device__ __noinline__ void checkpoint4_cold(float *dst, pair_t a, pair_t b) {
    *reinterpret_cast<ulonglong2 *>(dst) = make_ulonglong2(a, b);
}

}  // namespace

// NP: packed pairs per strip (4: eight sites per thread, rows of a multiple of 256 sites; 2: four sites).
// Rows are warp-aligned (L0 / (2 NP) is a multiple of 32), so a warp has ONE role for the whole launch and the
// step loop is instantiated per role -- a plain warp's loop carries no halo, reduction or service code:
//   EDGE     0 interior row | 1 first row of the band | 2 last row (a band has at least two rows): the edge rows
//            talk to the neighbour CTAs through the halo words
//   REDUCER  one warp per row turns the row's {sum phi, sum phi^2} partials into the per-step history; the first of them also
//            polls the event word and (CTA 0) draws for the omega work-item
constexpr int RES_MAX_STEPS_K = 2048;  // = RES_MAX_STEPS (sq_ctx.h): hist_p2 = hist_rows + RES_MAX_STEPS_K * L1
constexpr int ROWRES_THREADS = 896;  // 7 rows x 128 strips for 1024^2; 72 registers per thread

struct RowGeo {
    int r0, nr, k, j;
    unsigned smem_base, mbar, flags_a, chain_a;
};

template <int MATH, int POT, int NP, int EDGE, bool REDUCER>
__device__ __forceinline__ void rowres_steps(const ResidentArgs &A, pair_t (&PH)[NP], const RowGeo G) {
    constexpr int W = 2 * NP, NH = NP / 2;
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x, lane = tid & 31;
    const int L0 = A.L0, TPR = L0 / W, NRM = A.rows_max;
    const int r0 = G.r0, nr = G.nr, k = G.k, j = G.j;
    const unsigned mbar = G.mbar, flags_a = G.flags_a;

    // shared-memory geometry (byte addresses in the shared window).  A row is stored float4-interleaved:
    // float4 (h, j) at (h TPR + j), so the 128-bit accesses of a warp are contiguous also for 8-site strips.
    const unsigned rowb = (unsigned)L0 * 4u, halfb = (unsigned)TPR * 16u, parb = (unsigned)NRM * rowb;
    const unsigned a_own = G.smem_base + (unsigned)k * rowb + (unsigned)j * 16u;
    const int jl = (j == 0) ? TPR - 1 : j - 1, jr = (j + 1 == TPR) ? 0 : j + 1;
    const unsigned d_left = (unsigned)(((NH - 1) * TPR + jl) * 16 + 12) - (unsigned)j * 16u;  // relative to a_own (wraps mod 2^32)
    const unsigned d_right = (unsigned)jr * 16u - (unsigned)j * 16u;
    const unsigned rs_base = G.smem_base + 2u * parb;                   // rs[2 parity][NRM][TPR] pairs {sum phi, sum phi^2}
    const unsigned rsp = (unsigned)(NRM * TPR) * 8u;                    // bytes between the two parities
    const unsigned a_rs = rs_base + (unsigned)(k * TPR + j) * 8u;
    const unsigned a_chain = G.chain_a + (unsigned)tid * 16u;           // {T, K2} ; + 16 blockDim: {c1_0, c2_0} ; + 32 blockDim: role slot
    const unsigned chain_pl = (unsigned)blockDim.x * 16u;

    // halo words: [parity][CTA][first | last][L0]; only the edge rows' warps touch them
    // (32-bit word offsets off the base: one uniform base + offset per access instead of 64-bit pointers in registers)
    const unsigned hpar = (unsigned)nb * 2u * (unsigned)L0;
    const unsigned pub_off = ((unsigned)b * 2u + (EDGE == 2 ? 1u : 0u)) * (unsigned)L0 + (unsigned)(W * j);
    const unsigned src_off =   // the row across the band edge: neighbour below's LAST / neighbour above's FIRST row
        (EDGE == 1 ? ((unsigned)((b == 0) ? nb - 1 : b - 1) * 2u + 1u) : ((unsigned)((b + 1 == nb) ? 0 : b + 1) * 2u)) * (unsigned)L0 + (unsigned)(W * j);

    // constants of the packed pipeline
    const pair_t K_m4 = pk(-4.0f, -4.0f), K_clap = pk(A.c_lap_f, A.c_lap_f), K_m2cdt = pk(-A.c_2dt_f, -A.c_2dt_f);
    const pair_t K_mcdt = pk(-A.c_dt_f, -A.c_dt_f), K_lam = pk(A.lam_f, A.lam_f), K_m2 = pk(A.m2_f, A.m2_f);
    const pair_t K_k2 = pk(A.k2_f, A.k2_f), K_2m32 = pk(2.3283064365386963e-10f, 2.3283064365386963e-10f);
    const float kth = (float)(2.0 * 3.1415 / 4294967296.0);  // theta - pi = 2*3.1415 * u2 * 2^-32 - pi
    const pair_t K_th = pk(kth, kth), K_mpi = pk(-3.14159265358979f, -3.14159265358979f), K_m1 = pk(-1.0f, -1.0f);
    (void)K_m2cdt; (void)K_mcdt; (void)K_lam; (void)K_m2; (void)K_k2; (void)K_2m32; (void)K_th; (void)K_mpi;

    // ---- noise phase: the W draws of tau-step `step` and their Box-Muller values ------------------------
    // NZ[q] = -(noise of sites 2q, 2q+1) as a packed pair (the update subtracts it); advances T to the next step
    auto noise_phase = [&](pair_t (&NZ)[NP], int step) {
        const ulonglong2 tk = lds_pairs(a_chain), cc = lds_pairs(a_chain + chain_pl);
        const unsigned T0l = (unsigned)tk.x, T0h = (unsigned)(tk.x >> 32);
        unsigned tl = T0l, th = T0h, um = 0xFFFFFFFFu;
        u64 c1 = cc.x, c2 = cc.y;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            unsigned u1a, u2a, u1b, u2b, al, ah, bl, bh;
            mad48k(tl, th, A_LO, A_HI, c1, al, ah);              // t1 of site 2q
            mad48k(tl, th, ALPHA_LO32, ALPHA_HI32, c2, bl, bh);  // T  of site 2q
            u1a = __funnelshift_r(al, ah, 16);
            u2a = __funnelshift_r(bl, bh, 16);
            c1 += LCG_A;
            c2 += LCG_BETA;
            mad48k(bl, bh, A_LO, A_HI, c1, al, ah);              // site 2q+1
            mad48k(bl, bh, ALPHA_LO32, ALPHA_HI32, c2, tl, th);
            u1b = __funnelshift_r(al, ah, 16);
            u2b = __funnelshift_r(tl, th, 16);
            c1 += LCG_A;
            c2 += LCG_BETA;
            // inf-retry <=> u1 == 0 ; `seed+=` => u2 < 2^15: one 3-input min per site
            um = min(min(um, u1a), u2a);
            um = min(min(um, u1b), u2b);
            if (MATH == 1) {
                // r = cos(2*3.1415 v2) sqrt(-2 ln v1) with the amplitude folded under the root (k2 = 2 ln2 nscale^2):
                // v1 = (float)u1 * 2^-32 (exact scaling of the RN conversion), MUFU.LG2, MUFU.SQRT;
                // cos(theta) = -cos(theta - pi) keeps MUFU.COS in [-pi, pi)
                float l1a, l1b, ta, tb, tha, thb;
                upk(mul2(pk(__uint2float_rn(u1a), __uint2float_rn(u1b)), K_2m32), l1a, l1b);
                upk(mul2(pk(lg2_approx(l1a), lg2_approx(l1b)), K_k2), ta, tb);
                upk(fma2(pk(__uint2float_rn(u2a), __uint2float_rn(u2b)), K_th, K_mpi), tha, thb);
                NZ[q] = mul2(pk(__cosf(tha), __cosf(thb)), pk(sqrt_approx(fabsf(ta)), sqrt_approx(fabsf(tb))));
            } else {
                const float da = noise_accurate_dw(u1a, u2a, A.nscale), db = noise_accurate_dw(u1b, u2b, A.nscale);
                NZ[q] = pk(-da, -db);
            }
        }
        if (__builtin_expect(um < 32768u, 0)) {
            const u64 z0 = (tk.x - TWO31) & LCG_MASK;
            stripw_events_cold(A.event_key, step, z0, (u64)(r0 + k) * L0 + W * j, W);
        }
        unsigned nl, nh;
        mad48k(T0l, T0h, (unsigned)A.P, (unsigned)(A.P >> 32), tk.y, nl, nh);  // the strip's T one whole step later
        sts_u64(a_chain, ((u64)nh << 32) | nl);
    };

    // ---- reducer duties ------------------------------------------------------------------------------
    // A reducer warp reduces ONE row per step (warp w: row w; fewer warps than rows: the general walk): the row's TPR
    // pairs {sum phi, sum phi^2} -> two history entries.  A lane covers 4 pairs per 128 threads of the row (TPR <= 256).
    // What it needs every step -- its shared-memory source (parity 0), the word offset of its history entries, the mask of
    // lanes beyond the row -- sits in the thread's role slot in shared memory: registers are what this loop is short of,
    // and ptxas would otherwise re-derive all of it from %tid every step.
    const int nw = (nr * TPR) >> 5;
    if (REDUCER) {
        const int wid = tid >> 5;
        const bool red_lane = 4 * lane < TPR;
        const unsigned red_src = rs_base + (unsigned)(wid * TPR + (red_lane ? 4 * lane : 0)) * 8u;
        const unsigned red_off = (unsigned)(r0 + wid);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a_chain + 2u * chain_pl), "r"(red_src), "r"(red_off),
                     "r"(__float_as_uint(red_lane ? 1.0f : 0.0f)), "r"((unsigned)wid + (unsigned)nw) : "memory");
    }
    auto reduce_rows = [&](int step, unsigned parity) {
        unsigned red_src, red_off, red_mask, red_next;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(red_src), "=r"(red_off), "=r"(red_mask), "=r"(red_next) : "r"(a_chain + 2u * chain_pl));
        const ulonglong2 v0 = lds_pairs(red_src + parity * rsp), v1 = lds_pairs(red_src + parity * rsp + 16u);
        const pair_t M = pk(__uint_as_float(red_mask), __uint_as_float(red_mask));
        pair_t acc = mul2(add2(add2(v0.x, v0.y), add2(v1.x, v1.y)), M);  // {sum phi, sum phi^2} of 4 threads
        if (TPR > 128) {  // (uniform) rows of more than 128 strips: a second chunk per lane
            const ulonglong2 w0 = lds_pairs(red_src + parity * rsp + 1024u), w1 = lds_pairs(red_src + parity * rsp + 1040u);
            if (4 * lane + 128 < TPR) acc = add2(acc, add2(add2(w0.x, w0.y), add2(w1.x, w1.y)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = add2(acc, (pair_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)acc, o));
        if (lane == 0) {
            float a, q;
            upk(acc, a, q);
            const unsigned idx = red_off + (unsigned)step * (unsigned)A.L1;
            A.hist_rows[idx] = (double)a;
            A.hist_p2[idx] = (double)q;
        }
        if (nw < nr)  // (uniform) fewer warps than rows
            reduce_rows_general(A.hist_rows, A.hist_p2, A.L1, rs_base + parity * rsp, TPR, (int)red_next, nw, nr, r0, step, lane);
    };
    // service lane (first reducer warp): event-word poll; CTA 0: the omega work-item's draws, whose running
    // step-start seed lives in shared memory too.
    const bool svc_lane = REDUCER && tid == 0, omega_lane = svc_lane && b == 0;
    const unsigned a_svc = G.flags_a + 8u;
    if (svc_lane) sts_u64(a_svc, omega_lane ? A.seed_in[0] : 0ull);

    unsigned myclamp = 0, failed = 0;
    pair_t NZ[NP];
    // Step order of a warp:  noise(n) | wait B(n-1) | stencil(n) | arrive B(n).  The noise phase fills the time the
    // barrier needs; an edge row's halo words are requested between the two and travel during the wait.  (Letting every
    // other warp of a scheduler run one noise phase out of step with its siblings, so that the integer / SFU work of one
    // overlaps the shared-memory / packed-FP work of the other, was measured: 491 against 506 G site-updates/s -- the
    // warps de-phase by themselves.)

    // first barrier phase: everybody's initial row is in rowbuf[0]
    __syncwarp();
    if (lane == 0) mbar_arrive(mbar);

    ulonglong2 pre[EDGE ? NP : 1];  // edge rows: halo words of the field in hand, requested before the barrier wait
    int n = 0, voided = 0;
    for (; n < A.nsteps; ++n) {
        const unsigned par = (unsigned)n & 1u;
        noise_phase(NZ, A.step_index0 + n);
        if (EDGE && n > 0) {  // request the neighbour CTA's words now: the answer travels while this warp waits for the CTA
            const unsigned long long *ps = A.halo_ll + (src_off + par * hpar);
#pragma unroll
            for (int i = 0; i < NP; ++i) pre[i] = ld_words(ps + 2 * i);
        }
        mbar_wait(mbar, par);  // phase n: B(n-1), the initial phase for n = 0
        if (REDUCER && n > 0) reduce_rows(n - 1, par ^ 1u);  // per-row sums of the field before step n-1
        // ---- rare: leave (flags written before the arrivals of the phase just waited for: the same answer in
        //      every thread of the CTA), checkpoint of the field before step n -----------------------------------
        const unsigned fl = lds_f32_bits(flags_a + par * 4u);
        if (__builtin_expect((fl != 0u) | (((unsigned)n & (RES_CKPT - 1)) == 0u), 0)) {
            voided = (int)(fl >> 1);
            if (n > 0 && !voided && ((unsigned)n & (RES_CKPT - 1)) == 0u) {
                float *dst = A.ckpt + (size_t)((n / RES_CKPT) % RES_NCKPT) * (size_t)A.V + (size_t)(r0 + k) * L0 + W * j;
#pragma unroll
                for (int h = 0; h < NH; ++h) checkpoint4_cold(dst + 4 * h, PH[2 * h], PH[2 * h + 1]);
                if (myclamp) atomicAdd(A.nclamp_slots + (n / RES_CKPT - 1), (unsigned long long)myclamp);
                myclamp = 0;
            }
            if (fl) break;
        }
        u64 ek = NO_EVENT;  // service lane: the event word, requested here and looked at after the stencil phase
        if (REDUCER && svc_lane) ek = *((volatile const u64 *)A.event_key);

        // ---- stencil phase: neighbours of the field before step n -----------------------------------------
        const unsigned a_cur = a_own + par * parb, a_nxt = a_own + (par ^ 1u) * parb;
        const float left = lds_f32(a_cur + d_left), right = lds_f32(a_cur + d_right);
        pair_t UP[NP], DN[NP];
        if (EDGE != 2) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const ulonglong2 a = lds_pairs(a_cur + rowb + h * halfb);
                UP[2 * h] = a.x;
                UP[2 * h + 1] = a.y;
            }
        }
        if (EDGE != 1) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const ulonglong2 a = lds_pairs(a_cur - rowb + h * halfb);
                DN[2 * h] = a.x;
                DN[2 * h + 1] = a.y;
            }
        }
        if (EDGE) {
            pair_t(&OUT)[NP] = (EDGE == 1) ? DN : UP;
            if (n > 0) {
                // {value, tag} words of the neighbour CTA's boundary row of THIS field (tag = step0 + n)
                const unsigned want = A.step0 + (unsigned)n;
                const unsigned long long *s = A.halo_ll + (src_off + par * hpar);
                unsigned spins = 0;
                for (;;) {
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < NP; ++i)
                        ok &= ((unsigned)(pre[i].x >> 32) == want) & ((unsigned)(pre[i].y >> 32) == want);
                    if (__builtin_expect(ok, 1)) break;
                    ++spins;
                    // the launch is being abandoned (an RNG event must be replayed) or the neighbour is lost (never
                    // hang the GPU): stop waiting; what this step computes from here on is void
                    const bool lost = spins > (1u << 20);
                    if (lost || ((spins & 15u) == 0 && *((volatile const u64 *)A.event_key) != NO_EVENT)) {
                        if (lost) failed = 1;
                        sts_u32(flags_a + (par ^ 1u) * 4u, 3u);  // leave, and the step in flight is void
                        break;
                    }
#pragma unroll
                    for (int i = 0; i < NP; ++i) pre[i] = ld_words(s + 2 * i);
                }
#pragma unroll
                for (int i = 0; i < NP; ++i) OUT[i] = pk(__uint_as_float((unsigned)pre[i].x), __uint_as_float((unsigned)pre[i].y));
            } else {  // the neighbour's row of the initial field comes straight from the input buffer
                const int rr = (EDGE == 1) ? ((r0 == 0) ? A.L1 - 1 : r0 - 1) : ((r0 + nr == A.L1) ? 0 : r0 + nr);
                const float *src = A.in + (size_t)rr * L0 + W * j;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(src + 4 * h);
                    OUT[2 * h] = a.x;
                    OUT[2 * h + 1] = a.y;
                }
            }
        }
        // ---- observables of the pre-update field: per-thread partials, reduced after the barrier ----------
        {
            pair_t s = PH[0], p2 = mul2(PH[0], PH[0]);
#pragma unroll
            for (int q = 1; q < NP; ++q) {
                s = add2(s, PH[q]);
                p2 = fma2(PH[q], PH[q], p2);
            }
            float sl, sh, pl, ph;
            upk(s, sl, sh);
            upk(p2, pl, ph);
            sts_u64(a_rs + par * rsp, pk(__fadd_rn(sl, sh), __fadd_rn(pl, ph)));
        }
        // ---- update: s = phi(+0) + phi(-0); s += phi(+1); s += phi(-1)   -----------------
        float p[W];
#pragma unroll
        for (int q = 0; q < NP; ++q) upk(PH[q], p[2 * q], p[2 * q + 1]);
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const float xm0 = (q == 0) ? left : p[(2 * q - 1 + W) % W], xp1 = (q == NP - 1) ? right : p[(2 * q + 2) % W];
            pair_t S = pk(__fadd_rn(p[2 * q + 1], xm0), __fadd_rn(xp1, p[2 * q]));
            S = add2(S, UP[q]);
            S = add2(S, DN[q]);
            pair_t v = fma2(K_clap, fma2(K_m4, PH[q], S), PH[q]);
            if (POT == 4) v = fma2(K_mcdt, mul2(PH[q], fma2(K_lam, mul2(PH[q], PH[q]), K_m2)), v);
            else v = fma2(K_m2cdt, PH[q], v);  // (-c_dt)(2 phi) == (-2 c_dt) phi exactly
            v = fma2(K_m1, NZ[q], v);          // v + dw, one rounding
            PH[q] = v;
            float a0, a1;
            upk(v, a0, a1);
            m = fmaxf(fmaxf(fabsf(a0), fabsf(a1)), m);
        }
        // clamp (tau_kernel.cl:122-132): values at or beyond +-1000 leave through one test per strip.  (A NaN cannot
        // arise without an inf in the noise, i.e. without the inf-retry event of this strip's draws -- flagged in the
        // noise phase: the step is replayed anyway.)
        if (__builtin_expect(!(m < 1000.0f), 0)) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const Pair2 c = clamp4_cold(PH[2 * h], PH[2 * h + 1]);
                PH[2 * h] = c.a;
                PH[2 * h + 1] = c.b;
                myclamp += c.n;
            }
        }
        // ---- hand the new row over: shared memory for the band, {value, tag} words for the neighbour CTA ------
#pragma unroll
        for (int h = 0; h < NH; ++h) sts_pairs(a_nxt + h * halfb, PH[2 * h], PH[2 * h + 1]);
        const bool more = n + 1 < A.nsteps;
        if (EDGE && more) {
            const unsigned tag = A.step0 + (unsigned)n + 1u;
            unsigned long long *ho = A.halo_ll + (pub_off + (par ^ 1u) * hpar);
            float a[W];
#pragma unroll
            for (int q = 0; q < NP; ++q) upk(PH[q], a[2 * q], a[2 * q + 1]);
#pragma unroll
            for (int i = 0; i < NP; ++i) st_words(ho + 2 * i, word_of(a[2 * i], tag), word_of(a[2 * i + 1], tag));
        }
        if (REDUCER && svc_lane && ek != NO_EVENT) sts_u32(flags_a + (par ^ 1u) * 4u, 1u);  // everybody leaves at the top of the next step
        __syncwarp();
        if (lane == 0) mbar_arrive(mbar);

        // ---- independent of every other thread: the omega work-item's draw, the next step's noise --------------
        if (REDUCER && omega_lane) {  // gid = V, tau_kernel.cl:103-110
            const u64 sV = lcg_apply(A.vol_jump, lds_u64(a_svc), 0) & LCG_MASK;
            u64 t1, t2;
            lcg_draw(sV, (u64)A.V, t1, t2);
            if (lcg_event(sV, t1, t2)) atomicMin((unsigned long long *)A.event_key, event_key(A.step_index0 + n, 0, (u64)A.V));
            sts_u64(a_svc, lcg_next_seed(t2));
            if (!more) A.seed_out[0] = lcg_next_seed(t2);
        }
    }
    if (n == A.nsteps) {  // went through: the last step's sums, and was the last step void?
        mbar_wait(mbar, (unsigned)n & 1u);
        const unsigned fl = lds_f32_bits(flags_a + ((unsigned)n & 1u) * 4u);
        voided = (int)(fl >> 1);
        if (REDUCER && n > 0) reduce_rows(n - 1, ((unsigned)n & 1u) ^ 1u);
    }

    // n = number of steps this CTA went through; the last one is void if a halo wait was abandoned in it
    const int valid = n - (voided ? 1 : 0);
    if (tid == 0) A.progress[b] = (unsigned)(valid < 0 ? 0 : valid);
    if (failed) atomicExch(A.error_flag, 1u);
    if (n < A.nsteps || voided) return;  // left early: the output buffer is not needed
    // ---- write the band back ---------------------------------------------------------------------------
    {
        float *dst = A.out + (size_t)(r0 + k) * L0 + W * j;
#pragma unroll
        for (int h = 0; h < NH; ++h) *reinterpret_cast<ulonglong2 *>(dst + 4 * h) = make_ulonglong2(PH[2 * h], PH[2 * h + 1]);
    }
    if (myclamp) atomicAdd(A.nclamp_slots + (A.nsteps - 1) / RES_CKPT, (unsigned long long)myclamp);
}

template <int MATH, int POT, int NP>
__global__ void __launch_bounds__(ROWRES_THREADS, 1) rowres_kernel(const ResidentArgs A) {
    constexpr int W = 2 * NP, NH = NP / 2;  // sites per strip, float4 per strip
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    const int L0 = A.L0, TPR = L0 / W, NRM = A.rows_max;
    RowGeo G;
    G.r0 = (int)(((long long)b * A.L1) / nb);
    G.nr = (int)(((long long)(b + 1) * A.L1) / nb) - G.r0;  // >= 2 (sq_api.cu picks the grid)
    // thread -> (row, strip).  The two edge rows of the band get the HIGHEST warp ids (the warp arbiter favours
    // them): they carry the halo round trip on top of the common work.
    const int kk = tid / TPR;
    G.j = tid - kk * TPR;
    G.k = (kk >= G.nr) ? kk : ((kk == G.nr - 1) ? G.nr - 1 : ((kk == G.nr - 2) ? 0 : kk + 1));

    // shared: rowbuf[2][NRM][L0] | rs[2][2][NRM][TPR] | chain[3][threads] x 16 B | mbarrier | flags[2] | service state
    G.smem_base = (unsigned)__cvta_generic_to_shared(smem_raw);
    G.chain_a = G.smem_base + (unsigned)(2 * NRM * L0 + 4 * NRM * TPR) * 4u;
    G.mbar = G.chain_a + (unsigned)blockDim.x * 48u;
    G.flags_a = G.mbar + 8u;
    // flags[step parity]: bit 0 leave, bit 1 the step in flight is void; written during step n for step n+1, read at
    // the top of a step after the barrier phase that orders them -- every thread of the CTA takes the same decision
    if (tid == 0) {
        mbar_init(G.mbar, (unsigned)((G.nr * TPR) >> 5));
        // an earlier launch flagged an event: this one will be replayed (one thread decides for the CTA)
        sts_u32(G.flags_a, (*((volatile const u64 *)A.event_key) != NO_EVENT) ? 4u : 0u);
        sts_u32(G.flags_a + 4u, 0u);
    }
    __syncthreads();
    if (kk >= G.nr || lds_f32_bits(G.flags_a) == 4u) return;

    // ---- the strip: field, chain state -------------------------------------------------------------
    pair_t PH[NP];
    {
        const float *src = A.in + (size_t)(G.r0 + G.k) * L0 + W * G.j;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(src + 4 * h);
            PH[2 * h] = a.x;
            PH[2 * h + 1] = a.y;
            sts_pairs(G.smem_base + (unsigned)(G.k * L0) * 4u + (unsigned)G.j * 16u + h * (unsigned)TPR * 16u, a.x, a.y);
        }
    }
    {   // T = (seed before the strip's first draw) + 2^31;  T(n+1) = P T(n) + K2: the strip keeps its gids, one
        // affine map per step;  t1_e = A T + c1_0 + e A,  T_e = A^2 T + c2_0 + e (A^2 + A)
        const u64 g0 = (u64)(G.r0 + G.k) * L0 + W * G.j;
        const u64 S0 = A.seed_in[0];
        const u64 S1 = (A.P * S0 + A.Q) & LCG_MASK;  // predicted seed after one whole step
        const u64 s0 = lcg_seed_at(S0, 0, g0, A.jump), s1 = lcg_seed_at(S1, 0, g0, A.jump);
        const u64 K2 = (s1 - A.P * s0) - A.P * TWO31 + TWO31;
        const u64 c1_0 = site_const(g0) - LCG_A * TWO31, c2_0 = (LCG_A + 1) * site_const(g0) - LCG_ALPHA * TWO31;
        sts_pairs(G.chain_a + (unsigned)tid * 16u, s0 + TWO31, K2);
        sts_pairs(G.chain_a + (unsigned)(blockDim.x + tid) * 16u, c1_0, c2_0);
    }
    const int edge = (G.k == 0) ? 1 : ((G.k == G.nr - 1) ? 2 : 0);
    const bool reducer = (tid >> 5) < G.nr;
#define SQ_ROLE(E, R) rowres_steps<MATH, POT, NP, E, R>(A, PH, G)
    if (edge == 0) { if (reducer) SQ_ROLE(0, true); else SQ_ROLE(0, false); }
    else if (edge == 1) { if (reducer) SQ_ROLE(1, true); else SQ_ROLE(1, false); }
    else { if (reducer) SQ_ROLE(2, true); else SQ_ROLE(2, false); }
#undef SQ_ROLE
}

// ---- history of a launch -> running means -----------------------------------------------------------------
// per-step global sums of the history: one warp per step -> step_sums[n] = (sum phi, sum phi^2)
__global__ void __launch_bounds__(256) history_sums_kernel(const WelfordArgs A, double *step_sums) {
    if (*((volatile const u64 *)A.event_key) != NO_EVENT) return;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (n >= A.nsteps) return;
    double s1 = 0, s2 = 0;
    for (int k = l; k < A.nt; k += 32) s1 += A.hist_rows[(size_t)n * A.nt + k];
    for (int k = l; k < A.np2; k += 32) s2 += A.hist_p2[(size_t)n * A.np2 + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (l == 0) {
        step_sums[2 * n] = s1;
        step_sums[2 * n + 1] = s2;
    }
}

// history -> running means (tau_kernel.cl:144-145 per time slice).  The reference's update
// x <- x + (v - x)/(runs+j+1) is the running mean, so n more samples give, in closed form,
//     x' = x + (sum_j v_j - n x) / (runs + n):
// two sums per slice over the launch's steps instead of a sequential recurrence (fp64 rounding differs
// from the step-by-step form at the 1e-15 level; the parity tolerance on these observables is 1e-3).
// A block = 32 slices x 8 step ranges: thread (ts, sg) sums its eighth of the launch's steps for slice ts (a warp's loads of
// one step are 256 consecutive bytes), the eight partial sums meet in shared memory in fixed order.  One thread per slice
// over all 1000 steps -- 63 dependent batches of loads -- took 46 us of a 1.9 ms frame; this takes a fifth of it.  One more
// block sums the per-step global sums the same way.
constexpr int WELFORD_SG = 8;
__global__ void __launch_bounds__(256) welford_history_kernel(const WelfordArgs A, const double *step_sums) {
    if (*((volatile const u64 *)A.event_key) != NO_EVENT) return;
    __shared__ double sh1[WELFORD_SG][32], sh2[WELFORD_SG][32];
    const int ts = threadIdx.x & 31, sg = threadIdx.x >> 5;
    const double n = (double)A.nsteps, den = (double)(A.runs + A.nsteps);
    const int chunk = (A.nsteps + WELFORD_SG - 1) / WELFORD_SG;
    if ((int)blockIdx.x < (A.nt + 31) / 32) {
        const int t = blockIdx.x * 32 + ts;
        const int n_lo = min(A.nsteps, sg * chunk), n_hi = min(A.nsteps, n_lo + chunk);
        double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
        if (t < A.nt) {
            constexpr int B = 8;
            for (int n0 = n_lo; n0 < n_hi; n0 += B) {
                double h[B], hm[B];
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    const int k = min(n0 + j, n_hi - 1);
                    h[j] = A.hist_rows[(size_t)k * A.nt + t];
                    hm[j] = A.hist_rows[(size_t)k * A.nt + A.tmid];
                }
#pragma unroll
                for (int j = 0; j < B; ++j)
                    if (n0 + j < n_hi) {
                        s1[j & 3] += h[j];
                        s2[j & 3] = fma(h[j], hm[j], s2[j & 3]);
                    }
            }
        }
        sh1[sg][ts] = (s1[0] + s1[1]) + (s1[2] + s1[3]);
        sh2[sg][ts] = (s2[0] + s2[1]) + (s2[2] + s2[3]);
        __syncthreads();
        if (sg == 0 && t < A.nt) {
            const double inv_vs = 1.0 / (double)A.vslice;
            double a1 = 0, a2 = 0;
#pragma unroll
            for (int g = 0; g < WELFORD_SG; ++g) {
                a1 += sh1[g][ts];
                a2 += sh2[g][ts];
            }
            const double SP = a1 * inv_vs, SPP = a2 * inv_vs * inv_vs;
            const double x = A.slice_x[t], xx0 = A.slice_xx0[t];
            A.slice_x[t] = x + (SP - n * x) / den;
            A.slice_xx0[t] = xx0 + (SPP - n * xx0) / den;
            A.slice_sum[t] = A.hist_rows[(size_t)(A.nsteps - 1) * A.nt + t];
        }
    } else {  // the extra block: running means of <phi>, <phi^2> from the per-step global sums
        double a1 = 0, a2 = 0;
        for (int k = threadIdx.x; k < A.nsteps; k += 256) {
            a1 += step_sums[2 * k];
            a2 += step_sums[2 * k + 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (ts == 0) { sh1[sg][0] = a1; sh2[sg][0] = a2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            a1 = 0;
            a2 = 0;
            for (int g = 0; g < WELFORD_SG; ++g) {
                a1 += sh1[g][0];
                a2 += sh2[g][0];
            }
            const double inv_vol = 1.0 / ((double)A.vslice * (double)A.nt);
            A.sums[0] = step_sums[2 * (A.nsteps - 1)];
            A.sums[1] = step_sums[2 * (A.nsteps - 1) + 1];
            A.sums_mean[0] += (a1 * inv_vol - n * A.sums_mean[0]) / den;
            A.sums_mean[1] += (a2 * inv_vol - n * A.sums_mean[1]) / den;
        }
    }
}

cudaError_t launch_welford_history(const WelfordArgs &A, double *step_sums, cudaStream_t stream) {
    history_sums_kernel<<<(A.nsteps + 7) / 8, 256, 0, stream>>>(A, step_sums);
    welford_history_kernel<<<(A.nt + 31) / 32 + 1, 256, 0, stream>>>(A, step_sums);
    return cudaGetLastError();
}

template <int MATH, int POT, int NP>
static cudaError_t launch_rowres_np(const ResidentArgs &A, int nblocks, int threads, size_t smem, cudaStream_t st) {
    void *args[] = {(void *)&A};
    const void *fn = (const void *)rowres_kernel<MATH, POT, NP>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaLaunchCooperativeKernel(fn, dim3(nblocks), dim3(threads), args, smem, st);
}
template <int MATH, int POT>
static cudaError_t launch_rowres_mp(const ResidentArgs &A, int nblocks, cudaStream_t st) {
    const int w = rowres_strip(A.L0, A.rows_max);
    if (!w || A.L1 < 2 * nblocks) return cudaErrorInvalidValue;
    const int tpr = A.L0 / w, threads = A.rows_max * tpr;
    const size_t smem = sizeof(float) * ((size_t)2 * A.rows_max * A.L0 + (size_t)4 * A.rows_max * tpr) + (size_t)threads * 48 + 64;
    return w == 8 ? launch_rowres_np<MATH, POT, 4>(A, nblocks, threads, smem, st) : launch_rowres_np<MATH, POT, 2>(A, nblocks, threads, smem, st);
}

// sites per thread for rows of L0 sites and rows_max rows per CTA: 8 where rows stay warp-aligned, else 4; 0: the
// shape does not fit (more than ROWRES_THREADS threads)
int rowres_strip(int L0, int rows_max) {
    if (L0 % 128 != 0 || L0 > 1024 || rows_max < 2) return 0;
    // (16 sites per thread -- 14 warps per SM with 127 registers -- was measured at 188 G site-updates/s against 470:
    // the code of six role loops no longer fits the instruction caches and half the warps hide half the latency)
    if (L0 % 256 == 0 && rows_max * (L0 / 8) <= ROWRES_THREADS) return 8;
    if (rows_max * (L0 / 4) <= ROWRES_THREADS) return 4;
    return 0;
}

// rows_max = ceil(L1 / nblocks) rows per CTA; every CTA owns at least two rows (L1 >= 2 nblocks)
cudaError_t launch_rowres(const ResidentArgs &A, int math, int nblocks, cudaStream_t st) {
    if (A.pot == 4) return math ? launch_rowres_mp<1, 4>(A, nblocks, st) : launch_rowres_mp<0, 4>(A, nblocks, st);
    return math ? launch_rowres_mp<1, 0>(A, nblocks, st) : launch_rowres_mp<0, 0>(A, nblocks, st);
}

}  // namespace sq
