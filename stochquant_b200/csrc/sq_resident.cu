// sq_resident.cu -- persistent, on-chip-resident Langevin kernel for 2-D lattices that fit in the
// register files of one B200 (configs[1]: 1024^2 fp32 = 4 MiB over 148 SMs).
//
// Why: at 1024^2 one tau-step is ~1.3 us of work; a launch per step costs more than the step.
// Here ONE cooperative launch advances `nsteps` steps:
//   * CTA b owns a band of consecutive rows (time slices); thread t owns the 4 columns 4t..4t+3
//     of every row of the band: the field lives in registers for the whole launch, up/down
//     neighbours are the thread's own registers, left/right neighbours cross threads through a
//     double-buffered shared-memory edge array (one __syncthreads per step for it);
//   * band edges: each CTA publishes its first/last row of step n+1 to a global (L2) halo buffer
//     right after computing them (boundary rows first), then raises a flag; neighbours prefetch
//     those rows while they finish their interior rows -- point-to-point flags, no grid barrier;
//   * RNG: the reference's shared-seed chain (tau_kernel.cl:269-284) per strip of 4 sites; a strip
//     keeps the same gids every step, so its seed advances by ONE affine map per step:
//     s(n+1) = P s(n) + K(strip)   (P = ALPHA^(V+1), see sq_lcg.cuh);
//   * observables: per-row sums of the pre-update field (tau_kernel.cl:144-145 at slice
//     granularity) -> shared-memory transpose -> one warp per row -> fp64 history[step][row];
//     welford_history_kernel turns the history into the running means after the launch;
//   * RNG events (inf-retry / `seed+=`): flagged with atomicMin(step,gid); the host rolls the
//     launch back and replays (sq_api.cu).  Waits are bounded: a lost neighbour raises an error
//     flag instead of hanging the GPU.
#include <cooperative_groups.h>

#include "sq_kernels.h"
#include "sq_site.cuh"
#include "sq_pair.cuh"

namespace sq {

namespace {

// cold path: exact event test of the 4 draws of a strip (literal replay is the host's job)
__device__ __noinline__ void strip_events_cold(u64 *event_key_ptr, int step, u64 sm, u64 g0, int w) {
    for (int e = 0; e < w; ++e) {
        u64 t1, t2;
        lcg_draw(sm, g0 + e, t1, t2);
        if (lcg_event(sm, t1, t2)) atomicMin((unsigned long long *)event_key_ptr, event_key(step, 0, g0 + e));
        sm = lcg_next_seed(t2) & LCG_MASK;
    }
}

// by value: taking the address of a register array would force it into local memory
__device__ __noinline__ unsigned count_clamped_cold(float a, float b, float c, float d) {
    return (fabsf(a) >= 1000.0f ? 1u : 0u) + (fabsf(b) >= 1000.0f ? 1u : 0u) + (fabsf(c) >= 1000.0f ? 1u : 0u) +
           (fabsf(d) >= 1000.0f ? 1u : 0u);
}

struct SiteCoef {
    float c_lap;   // (float)(m dtau / a2f)
    float c_dt;    // (float)dtau
    float c_2dt;   // 2*c_dt: pot 0 has F = 2 phi, and (-c_dt)(2 phi) == (-2 c_dt) phi exactly
    float m2, lam; // pot 4
    float k2;      // 2 ln2 * nscale^2 (FAST: noise amplitude folded under the square root)
    double nscale_d;
    int pot;
};

// new value of one site.  nsum = sum of the 4 neighbours in the oracle's order (+0,-0,+1,-1).
// Operation order is part of the definition of the fp32 lattice update ("lattice update").
template <int MATH, int POT>
__device__ __forceinline__ float site_update(float phi, float nsum, unsigned u1, unsigned u2, const SiteCoef &C) {
    const float lap = __fmaf_rn(-4.0f, phi, nsum);
    float v = __fmaf_rn(C.c_lap, lap, phi);
    if (POT == 4) {
        const float F = __fmul_rn(phi, __fmaf_rn(C.lam, __fmul_rn(phi, phi), C.m2));
        v = __fmaf_rn(-C.c_dt, F, v);
    } else {
        v = __fmaf_rn(-C.c_2dt, phi, v);
    }
    float dw;
    if (MATH == 1) dw = site_noise_fast(u1, u2, C.k2);
    else dw = (float)__dmul_rn(C.nscale_d, noise_accurate((u64)u1 << 16, (u64)u2 << 16));
    v = __fadd_rn(v, dw);
    return fmaxf(fminf(v, 1000.0f), -1000.0f);  // NaN -> +1000 (tau_kernel.cl:122-132)
}

}  // namespace

// 16-byte load of two {float bits, tag} words; each 64-bit element is a single-copy-atomic scalar
__device__ __forceinline__ ulonglong2 ld_relaxed_ll(const unsigned long long *p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ll_pack(float v, unsigned tag) {
    return ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
}

// Halo protocol: every boundary value travels as one 64-bit word {float, step tag}.  A word whose
// tag equals the expected step IS the data of that step (64-bit scalar accesses are single-copy
// atomic), so the producer needs no fence and no separate flag, the consumer no acquire:
// one L2 write + one L2 read of latency.  Tags never repeat within a context (step0 is monotonic,
// also across replayed launches).  Each thread prefetches the words above/below its own columns
// into registers in the middle of a step (non-blocking) and only checks the tags when it needs
// the values at the start of the next step; a miss re-polls (bounded, abort-aware).

// ---- compute warps.  NR = rows owned by this CTA (compile-time: the band stays in registers) ---
// W  = sites per thread and row (4: 128-bit strips, 8 warps per 1024-site row; 2: 16 warps -- more
//      thread-level parallelism for the dependent integer/SFU chains, the kernel's real limiter)
// TT = compile-time threads per row (0: runtime) so shared-memory offsets become immediates.
template <int W>
struct alignas(4 * W) RowPack {
    float v[W];
};

// checkpoint of a thread's band, out of line (by value: the hot loop's registers stay untouched)
template <int N>
struct Band {
    float v[N];
};
template <int NR, int W>
__device__ __noinline__ void checkpoint_cold(float *ckpt, int n, int L1, int L0, int r0, int t, const Band<NR * W> bd) {
    float *dst = ckpt + (size_t)((n / RES_CKPT) % 3) * (size_t)L1 * L0 + (size_t)r0 * L0 + W * t;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        RowPack<W> v;
#pragma unroll
        for (int e = 0; e < W; ++e) v.v[e] = bd.v[k * W + e];
        *reinterpret_cast<RowPack<W> *>(dst + (size_t)k * L0) = v;
    }
}

template <int NR, int MATH, int POT, int W, int TT>
__device__ __forceinline__ void resident_run(const ResidentArgs &A, int r0, float *smem) {
    const int T = TT ? TT : (int)blockDim.x;
    const int t = threadIdx.x, b = blockIdx.x, nb = gridDim.x;
    const int L0 = A.L0;
    const int bup = (b + 1 == nb) ? 0 : b + 1, bdn = (b == 0) ? nb - 1 : b - 1;
    // shared: edges [2 buffers][NR rows][2 (first,last)][T] ; row-sum transpose [2][NR+1][T] ;
    // per-strip chain state [NR][T] x {seed lo, seed hi, K lo, K hi} (kept out of the register file)
    float *edge = smem;
    float *rs_all = smem + 2 * NR * 2 * T;
    uint4 *chain = reinterpret_cast<uint4 *>(rs_all + 2 * (NR + 1) * T);
    SiteCoef C;  // float constants are cast on the host: plain constant-bank operands, no F2F here
    C.c_lap = A.c_lap_f;
    C.pot = A.pot;
    C.c_dt = A.c_dt_f;
    C.c_2dt = A.c_2dt_f;
    C.m2 = A.m2_f;
    C.lam = A.lam_f;
    C.nscale_d = A.nscale;
    C.k2 = A.k2_f;

    // ---- load the band, set up the per-strip chain state -----------------------------------
    float phi[NR][W];
    const u64 S0 = A.seed_in[0];
    const u64 S1 = (A.P * S0 + A.Q) & LCG_MASK;  // predicted seed after one whole step
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const RowPack<W> v = *reinterpret_cast<const RowPack<W> *>(A.in + (size_t)(r0 + k) * L0 + W * t);
#pragma unroll
        for (int e = 0; e < W; ++e) phi[k][e] = v.v[e];
        const u64 g = (u64)(r0 + k) * L0 + W * t;
        const u64 s0 = lcg_seed_at(S0, 0, g, A.jump);
        const u64 s1 = lcg_seed_at(S1, 0, g, A.jump);
        const u64 K = (s1 - A.P * s0) & LCG_MASK;
        chain[k * T + t] = make_uint4((unsigned)s0, (unsigned)(s0 >> 32), (unsigned)K, (unsigned)(K >> 32));
    }
    const unsigned Pl = (unsigned)A.P, Ph = (unsigned)(A.P >> 32);
    // c = gid*A + B of the first site of row 0's strip; rows advance it by L0*A
    // (opaque: otherwise the 64-bit multiply gid*A+B is re-derived for every row of every step)
    u64 c0 = site_const((u64)r0 * L0 + W * t);
    asm volatile("" : "+l"(c0));
    // where this thread publishes its boundary-row words (parity 0 / 1) and reads its neighbours'
    unsigned long long *const pub0 = A.halo_ll + ((size_t)0 * nb + b) * 2 * L0 + W * t;
    unsigned long long *const pub1 = A.halo_ll + ((size_t)1 * nb + b) * 2 * L0 + W * t;

    // edges of the initial field
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        edge[((0 * NR + k) * 2 + 0) * T + t] = phi[k][0];
        edge[((0 * NR + k) * 2 + 1) * T + t] = phi[k][W - 1];
    }
    __syncthreads();

    // neighbours' rows of the initial field, straight from the input buffer
    float hdn[W], hup[W];
    {
        const int rup = (r0 + NR == A.L1) ? 0 : r0 + NR, rdn = (r0 == 0) ? A.L1 - 1 : r0 - 1;
        const RowPack<W> d = *reinterpret_cast<const RowPack<W> *>(A.in + (size_t)rdn * L0 + W * t);
        const RowPack<W> u = *reinterpret_cast<const RowPack<W> *>(A.in + (size_t)rup * L0 + W * t);
#pragma unroll
        for (int e = 0; e < W; ++e) { hdn[e] = d.v[e]; hup[e] = u.v[e]; }
    }
    ulonglong2 pre_dn[W / 2], pre_up[W / 2];  // prefetched {value, tag} words of the next step's halo
    Seed32 Som = seed_split(S0);              // (b==0, t==0): the step-start seed, for the omega draw
    unsigned failed = 0;

    const int tl = (t == 0) ? T - 1 : t - 1, tr = (t + 1 == T) ? 0 : t + 1;
    unsigned myclamp = 0;

    // constants of the packed pipeline (uniform operands of FFMA2 / FMUL2)
    const pair_t K_m4 = pk(-4.0f, -4.0f), K_clap = pk(C.c_lap, C.c_lap), K_m2cdt = pk(-C.c_2dt, -C.c_2dt), K_mcdt = pk(-C.c_dt, -C.c_dt);
    const pair_t K_lam = pk(C.lam, C.lam), K_m2 = pk(C.m2, C.m2), K_k2 = pk(C.k2, C.k2);
    const pair_t K_2m32 = pk(2.3283064365386963e-10f, 2.3283064365386963e-10f);
    const float kth = (float)(2.0 * 3.1415 / 4294967296.0);  // theta - pi = 2*3.1415 * u2 * 2^-32 - pi
    const pair_t K_th = pk(kth, kth), K_mpi = pk(-3.14159265358979f, -3.14159265358979f);
    const unsigned one = A.one;
    (void)K_m4; (void)K_clap; (void)K_m2cdt; (void)K_mcdt; (void)K_lam; (void)K_m2; (void)K_k2; (void)K_2m32; (void)K_th; (void)K_mpi; (void)one;
    __shared__ unsigned s_abort;
    if (t == 0) s_abort = 0;
    u64 ek = NO_EVENT;  // thread 0: the event word as read one step ago (the load is never waited for)
    int n = 0;
    for (; n < A.nsteps; ++n) {
        const int eb = n & 1;
        // ---- event recovery: checkpoint, early exit -------------------------------------------
        if (__builtin_expect((n & (RES_CKPT - 1)) == 0 && n > 0, 0)) {
            Band<NR * W> bd;
#pragma unroll
            for (int k = 0; k < NR; ++k)
#pragma unroll
                for (int e = 0; e < W; ++e) bd.v[k * W + e] = phi[k][e];
            checkpoint_cold<NR, W>(A.ckpt, n, A.L1, L0, r0, t, bd);
        }
        if (t == 0) {
            if (ek != NO_EVENT) s_abort = 1;
            ek = *((volatile const u64 *)A.event_key);
        }
        float *rs = rs_all + (size_t)eb * (NR + 1) * T;  // double-buffered: last step's is being reduced
        float p2 = 0.f;
        pair_t P2 = 0;  // packed path: (sum phi^2 over even sites, over odd sites)
        // ---- one row: draws + update; returns the new values in out[] -----------------------
        auto do_row = [&](int k, const float *cur, const float *up, const float *dn, float *out) {
            const float left = edge[((eb * NR + k) * 2 + 1) * T + tl];
            const float right = edge[((eb * NR + k) * 2 + 0) * T + tr];
            u64 c = c0 + A.row_const[k];  // constant-bank operand: 2 instructions
            const uint4 st = chain[k * T + t];
            Seed32 s{st.x, st.y};
            const Seed32 s_before = s;
            float a = 0.f;
            bool maybe = false;
            if constexpr (W == 2 && MATH == 1) {
                // packed pipeline (sq_pair.cuh): both sites of the strip per instruction; unclamped values
                // are stored once proved in range, RNG events / clamp hits leave through one cold test
                unsigned u1a, u2a, u1b, u2b;
                site_draw_c(s, c, u1a, u2a);
                c += LCG_A;
                site_draw_c(s, c, u1b, u2b);
                const unsigned um = min(min(min(u1a, u2a), u1b), u2b);  // u1 == 0 or u2 < 2^15 => um < 2^15
                const float ca = cur[0], cb = cur[1];
                const pair_t Cp = pk(ca, cb);
                pair_t S = pk(__fadd_rn(cb, left), __fadd_rn(right, ca));  // phi(+0) + phi(-0)
                S = add2(S, pk(up[0], up[1]));
                S = add2(S, pk(dn[0], dn[1]));
                pair_t V = fma2(K_clap, fma2(K_m4, Cp, S), Cp);
                if (POT == 4) V = fma2(K_mcdt, mul2(Cp, fma2(K_lam, mul2(Cp, Cp), K_m2)), V);
                else V = fma2(K_m2cdt, Cp, V);
                float va, vb, l1a, l1b, ta, tb, tha, thb;
                upk(V, va, vb);
                upk(mul2(pk(__uint2float_rn(u1a), __uint2float_rn(u1b)), K_2m32), l1a, l1b);
                upk(mul2(pk(lg2_approx(l1a), lg2_approx(l1b)), K_k2), ta, tb);
                upk(fma2(pk(__uint2float_rn(u2a), __uint2float_rn(u2b)), K_th, K_mpi), tha, thb);
                out[0] = __fmaf_rn(-__cosf(tha), sqrt_approx(fabsf(ta)), va);
                out[1] = __fmaf_rn(-__cosf(thb), sqrt_approx(fabsf(tb)), vb);
                a = __fadd_rn(ca, cb);
                P2 = fma2(Cp, Cp, P2);
                const float amax = fmaxf(fabsf(out[0]), fabsf(out[1]));
                if (__builtin_expect((um < 32768u) | !(amax < 1000.0f), 0)) {
                    maybe = um < 32768u;
                    myclamp += (fabsf(out[0]) <= 1000.0f ? 0u : 1u) + (fabsf(out[1]) <= 1000.0f ? 0u : 1u);
                    out[0] = (out[0] < 1000.0f) ? ((out[0] > -1000.0f) ? out[0] : -1000.0f) : 1000.0f;  // NaN -> +1000
                    out[1] = (out[1] < 1000.0f) ? ((out[1] > -1000.0f) ? out[1] : -1000.0f) : 1000.0f;
                }
            } else {
            float amax = 0.f;
#pragma unroll
            for (int e = 0; e < W; ++e) {
                unsigned u1, u2;
                site_draw_c(s, c, u1, u2);
                c += LCG_A;
                maybe |= site_maybe_event(u1, u2);
                const float p = cur[e];
                const float xp = (e < W - 1) ? cur[(e + 1) % W] : right;
                const float xm = (e > 0) ? cur[(e + W - 1) % W] : left;
                const float nsum = __fadd_rn(__fadd_rn(__fadd_rn(xp, xm), up[e]), dn[e]);
                out[e] = site_update<MATH, POT>(p, nsum, u1, u2, C);
                a = __fadd_rn(a, p);
                p2 = __fmaf_rn(p, p, p2);
                amax = fmaxf(amax, fabsf(out[e]));
            }
            // clamp hits (tau_kernel.cl:122-132) are counted on a cold path
            if (__builtin_expect(amax >= 1000.0f, 0))
                myclamp += count_clamped_cold(out[0], out[1], W > 2 ? out[W - 2] : 0.f, W > 2 ? out[W - 1] : 0.f);
            }
            rs[k * T + t] = a;
            if (__builtin_expect(maybe, 0))
                strip_events_cold(A.event_key, A.step_index0 + n, seed_join(s_before), (u64)(r0 + k) * L0 + W * t, W);
            // the strip keeps its gids: next step's seed by one affine map (5 integer ops)
            const u64 p = (u64)st.x * Pl + (((u64)st.w << 32) | st.z);
            const unsigned nh = (unsigned)(p >> 32) + st.x * Ph + st.y * Pl;
            *reinterpret_cast<uint2 *>(&chain[k * T + t]) = make_uint2((unsigned)p, nh);
        };

        // ---- neighbours' rows of the current field: the words prefetched during the last step ---
        if (n > 0) {
            const unsigned want = A.step0 + (unsigned)n;
            const unsigned long long *hb = A.halo_ll + (size_t)eb * nb * 2 * L0 + W * t;
            const unsigned long long *src_dn = hb + ((size_t)bdn * 2 + 1) * L0;  // neighbour below: its LAST row
            const unsigned long long *src_up = hb + ((size_t)bup * 2 + 0) * L0;  // neighbour above: its FIRST row
            unsigned spins = 0;
            for (;;) {
                bool ok = true;
#pragma unroll
                for (int j = 0; j < W / 2; ++j)
                    ok &= ((unsigned)(pre_dn[j].x >> 32) == want) & ((unsigned)(pre_dn[j].y >> 32) == want) &
                          ((unsigned)(pre_up[j].x >> 32) == want) & ((unsigned)(pre_up[j].y >> 32) == want);
                if (__builtin_expect(ok, 1)) break;
                ++spins;
                // the launch is being abandoned (an RNG event must be replayed): stop waiting
                if ((spins & 15u) == 0 && *((volatile const u64 *)A.event_key) != NO_EVENT) break;
                if (spins > (1u << 20)) { failed = 1; break; }  // never hang the GPU
#pragma unroll
                for (int j = 0; j < W / 2; ++j) {
                    pre_dn[j] = ld_relaxed_ll(src_dn + 2 * j);
                    pre_up[j] = ld_relaxed_ll(src_up + 2 * j);
                }
            }
#pragma unroll
            for (int j = 0; j < W / 2; ++j) {
                hdn[2 * j] = __uint_as_float((unsigned)pre_dn[j].x);
                hdn[2 * j + 1] = __uint_as_float((unsigned)pre_dn[j].y);
                hup[2 * j] = __uint_as_float((unsigned)pre_up[j].x);
                hup[2 * j + 1] = __uint_as_float((unsigned)pre_up[j].y);
            }
        }

        // ---- phase A: the two boundary rows first, publish them (fire and forget) -----------
        float first[W], last[W];
        if (NR == 1) {
            do_row(0, phi[0], hup, hdn, first);
#pragma unroll
            for (int e = 0; e < W; ++e) last[e] = first[e];
        } else {
            do_row(0, phi[0], phi[1], hdn, first);
            do_row(NR - 1, phi[NR - 1], hup, phi[NR - 2], last);
        }
        if (n + 1 < A.nsteps) {  // {value, tag} words: no fence, no flag
            const unsigned tag = A.step0 + (unsigned)n + 1u;
            unsigned long long *ho = eb ? pub0 : pub1;  // parity (n+1)&1
#pragma unroll
            for (int e = 0; e < W; e += 2) {
                *reinterpret_cast<ulonglong2 *>(ho + e) = make_ulonglong2(ll_pack(first[e], tag), ll_pack(first[e + 1], tag));
                *reinterpret_cast<ulonglong2 *>(ho + L0 + e) = make_ulonglong2(ll_pack(last[e], tag), ll_pack(last[e + 1], tag));
            }
        }

        // ---- phase B: interior rows, in place (old copies of the rows still needed below) ----
        // after interior row KPRE: issue (do not wait for) the loads of the neighbours' new rows
        constexpr int KPRE = (NR >= 5) ? NR - 4 : ((NR >= 3) ? 1 : 0);
        auto prefetch_halo = [&]() {
            if (n + 1 < A.nsteps) {
                const unsigned long long *hb = A.halo_ll + (size_t)(eb ^ 1) * nb * 2 * L0 + W * t;
#pragma unroll
                for (int j = 0; j < W / 2; ++j) {
                    pre_dn[j] = ld_relaxed_ll(hb + ((size_t)bdn * 2 + 1) * L0 + 2 * j);
                    pre_up[j] = ld_relaxed_ll(hb + ((size_t)bup * 2 + 0) * L0 + 2 * j);
                }
            }
        };
        float prev[W];  // old values of row k-1
#pragma unroll
        for (int e = 0; e < W; ++e) prev[e] = phi[0][e];
#pragma unroll
        for (int k = 1; k < NR - 1; ++k) {
            float out[W];
            do_row(k, phi[k], phi[k + 1], prev, out);  // phi[k+1] is still old (row NR-1 lives in `last`)
#pragma unroll
            for (int e = 0; e < W; ++e) { prev[e] = phi[k][e]; phi[k][e] = out[e]; }
            if (k == KPRE) prefetch_halo();
        }
        if (KPRE < 1) prefetch_halo();
#pragma unroll
        for (int e = 0; e < W; ++e) {
            phi[0][e] = first[e];
            if (NR > 1) phi[NR - 1][e] = last[e];
        }

        // ---- hand-over: edges, row sums -------------------------------------------------------
        const int nbuf = eb ^ 1;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            edge[((nbuf * NR + k) * 2 + 0) * T + t] = phi[k][0];
            edge[((nbuf * NR + k) * 2 + 1) * T + t] = phi[k][W - 1];
        }
        if constexpr (W == 2 && MATH == 1) {
            float lo, hi;
            upk(P2, lo, hi);
            p2 = __fadd_rn(lo, hi);
        }
        rs[NR * T + t] = p2;
        __syncthreads();
        if (s_abort) break;  // an RNG event somewhere: this launch will be resumed from a checkpoint

        // ---- per-row sums: warp w reduces row w (and the phi^2 column) ---------------------
        {
            const int w = t >> 5, l = t & 31, nwarp = T >> 5;
            for (int row = w; row <= NR; row += nwarp) {
                float a = 0.f;
                for (int j = l; j < T; j += 32) a += rs[row * T + j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (l == 0) {
                    if (row < NR) A.hist_rows[(size_t)n * A.L1 + r0 + row] = (double)a;
                    else A.hist_p2[(size_t)n * nb + b] = (double)a;
                }
            }
        }

        // ---- the omega work-item's draw (gid = V), tau_kernel.cl:103-110 ----------------------
        if (b == 0 && t == 0) {
            const u64 S = seed_join(Som);
            const u64 sV = lcg_apply(A.vol_jump, S, 0) & LCG_MASK;
            u64 t1, t2;
            lcg_draw(sV, (u64)A.V, t1, t2);
            if (lcg_event(sV, t1, t2))
                atomicMin((unsigned long long *)A.event_key, event_key(A.step_index0 + n, 0, (u64)A.V));
            Som = seed_split(lcg_next_seed(t2));
            if (n == A.nsteps - 1) A.seed_out[0] = lcg_next_seed(t2);
        }
    }

    if (t == 0) A.progress[b] = (unsigned)n;
    if (n < A.nsteps) return;  // aborted: the output buffer is not needed
    // ---- write the band back ------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        RowPack<W> v;
#pragma unroll
        for (int e = 0; e < W; ++e) v.v[e] = phi[k][e];
        *reinterpret_cast<RowPack<W> *>(A.out + (size_t)(r0 + k) * L0 + W * t) = v;
    }
    if (myclamp) atomicAdd(A.nclamped, (unsigned long long)myclamp);
    if (failed) atomicExch(A.error_flag, 1u);
}

template <int ROWS, int MATH, int POT, int W, int TT>
__global__ void __launch_bounds__(W == 2 ? 512 : 256, 1) resident2d_kernel(const ResidentArgs A) {
    extern __shared__ float smem_f[];
    // an earlier launch flagged an event: this one will be replayed.  (If the flag rises while
    // the grid is still starting, late CTAs leave here and their neighbours' waits give up on it.)
    if (*((volatile const u64 *)A.event_key) != NO_EVENT) return;
    const int b = blockIdx.x, nb = gridDim.x;
    const int r0 = (int)(((long long)b * A.L1) / nb), r1 = (int)(((long long)(b + 1) * A.L1) / nb);
    if (r1 - r0 == ROWS) resident_run<ROWS, MATH, POT, W, TT>(A, r0, smem_f);
    else resident_run<(ROWS > 1 ? ROWS - 1 : 1), MATH, POT, W, TT>(A, r0, smem_f);
}

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
// One thread per slice, loads batched so their latencies overlap; fixed summation order.  (Splitting the
// step range over four warps per slice block was measured slower.)
__global__ void __launch_bounds__(128) welford_history_kernel(const WelfordArgs A, const double *step_sums) {
    if (*((volatile const u64 *)A.event_key) != NO_EVENT) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const double inv_vs = 1.0 / (double)A.vslice;
    const double n = (double)A.nsteps, den = (double)(A.runs + A.nsteps);
    if (t < A.nt) {
        constexpr int B = 16;
        double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, last = 0;
        for (int n0 = 0; n0 < A.nsteps; n0 += B) {
            double h[B], hm[B];
#pragma unroll
            for (int j = 0; j < B; ++j) {
                const int k = min(n0 + j, A.nsteps - 1);
                h[j] = A.hist_rows[(size_t)k * A.nt + t];
                hm[j] = A.hist_rows[(size_t)k * A.nt + A.tmid];
            }
#pragma unroll
            for (int j = 0; j < B; ++j)
                if (n0 + j < A.nsteps) {
                    s1[j & 3] += h[j];
                    s2[j & 3] = fma(h[j], hm[j], s2[j & 3]);
                    last = h[j];
                }
        }
        const double SP = ((s1[0] + s1[1]) + (s1[2] + s1[3])) * inv_vs;
        const double SPP = ((s2[0] + s2[1]) + (s2[2] + s2[3])) * inv_vs * inv_vs;
        const double x = A.slice_x[t], xx0 = A.slice_xx0[t];
        A.slice_x[t] = x + (SP - n * x) / den;
        A.slice_xx0[t] = xx0 + (SPP - n * xx0) / den;
        A.slice_sum[t] = last;
    }
    if (t == A.nt) {  // one spare thread: running means of <phi>, <phi^2>
        const double inv_vol = 1.0 / ((double)A.vslice * (double)A.nt);
        double a1 = 0, a2 = 0, s1 = 0, s2 = 0;
        for (int k = 0; k < A.nsteps; ++k) {
            s1 = step_sums[2 * k];
            s2 = step_sums[2 * k + 1];
            a1 += s1;
            a2 += s2;
        }
        A.sums[0] = s1;
        A.sums[1] = s2;
        A.sums_mean[0] += (a1 * inv_vol - n * A.sums_mean[0]) / den;
        A.sums_mean[1] += (a2 * inv_vol - n * A.sums_mean[1]) / den;
    }
}

cudaError_t launch_welford_history(const WelfordArgs &A, double *step_sums, cudaStream_t stream) {
    history_sums_kernel<<<(A.nsteps + 7) / 8, 256, 0, stream>>>(A, step_sums);
    welford_history_kernel<<<(A.nt + 1 + 31) / 32, 32, 0, stream>>>(A, step_sums);
    return cudaGetLastError();
}

template <int ROWS, int W, int TT>
static cudaError_t launch_rows(const ResidentArgs &A, int math, int nblocks, cudaStream_t st) {
    const int threads = A.L0 / W;
    const size_t smem = sizeof(float) * ((size_t)2 * ROWS * 2 * threads + (size_t)2 * (ROWS + 1) * threads) +
                        sizeof(uint4) * (size_t)ROWS * threads;
    void *args[] = {(void *)&A};
    const void *fn;
    if (A.pot == 4) fn = math ? (const void *)resident2d_kernel<ROWS, 1, 4, W, TT> : (const void *)resident2d_kernel<ROWS, 0, 4, W, TT>;
    else fn = math ? (const void *)resident2d_kernel<ROWS, 1, 0, W, TT> : (const void *)resident2d_kernel<ROWS, 0, 0, W, TT>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaLaunchCooperativeKernel(fn, dim3(nblocks), dim3(threads), args, smem, st);
}

// rows_max = ceil(L1 / nblocks).  strip_w: 2 or 4 sites per thread and row (0 = default).
cudaError_t launch_resident2d(const ResidentArgs &A, int math, int nblocks, int rows_max, int strip_w, cudaStream_t st) {
    if (A.L0 == 1024) {  // rows of 1024 sites (configs[1]): fully specialised
        if (strip_w != 4) {
            switch (rows_max) {
                case 4: return launch_rows<4, 2, 512>(A, math, nblocks, st);
                case 7: return launch_rows<7, 2, 512>(A, math, nblocks, st);
                case 8: return launch_rows<8, 2, 512>(A, math, nblocks, st);
            }
        } else {
            switch (rows_max) {
                case 4: return launch_rows<4, 4, 256>(A, math, nblocks, st);
                case 7: return launch_rows<7, 4, 256>(A, math, nblocks, st);
                case 8: return launch_rows<8, 4, 256>(A, math, nblocks, st);
            }
        }
    }
    switch (rows_max) {
        case 1: return launch_rows<1, 4, 0>(A, math, nblocks, st);
        case 2: return launch_rows<2, 4, 0>(A, math, nblocks, st);
        case 3: return launch_rows<3, 4, 0>(A, math, nblocks, st);
        case 4: return launch_rows<4, 4, 0>(A, math, nblocks, st);
        case 5: return launch_rows<5, 4, 0>(A, math, nblocks, st);
        case 6: return launch_rows<6, 4, 0>(A, math, nblocks, st);
        case 7: return launch_rows<7, 4, 0>(A, math, nblocks, st);
        case 8: return launch_rows<8, 4, 0>(A, math, nblocks, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sq
