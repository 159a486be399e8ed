// sq_march.cu -- row-marching streaming Langevin step, fp32, d = 3 and 4.
//
// Same update as lattice_step_kernel (sq_lattice.cu; generalisation of tau_kernel.cl:64-173 defined
// in DESIGN.md section 4 / oracle sqo_lattice_step) and the same integer stream, re-organised for
// ISSUE SLOTS: ncu showed the generic kernel spending 105-154 warp-instructions per site on a path
// whose arithmetic needs ~30 (profiles/), most of it index decomposition (runtime divisions), 64-bit
// jump-ahead per strip and per-load address arithmetic.  Here
//   * a thread owns one 16-byte strip position x0 and marches over R consecutive rows (x1) of one
//     (x2, t) plane: nothing is divided per strip, every address is base + a running 32-bit offset;
//   * the strip-to-strip seed advance is one affine map with a FIXED stride (one row = L0 draws):
//     s' = alpha^L0 s + c_k in 32-bit limbs (3 IMAD) with c_k itself a running sum (2 IADD);
//   * the floating-point work of the 4 sites runs on packed fp32x2 instructions (FADD2/FFMA2,
//     each lane rounded separately: bit-identical to the scalar sequence the oracle defines);
//   * RNG events and clamp hits are detected with one 3-input min / max per site (VIMNMX3 / FMNMX3)
//     and resolved on cold paths; the hot path stores unclamped values it has proved in range.
// Per-slice observables, the omega draw, replay entries (REBASE), L2 chunking and the slab ring's
// halo protocol are those of the generic kernel.
#include "sq_strip_slow.cuh"

namespace sq {

#ifndef MARCH_MINB
#define MARCH_MINB 4
#endif
#ifndef MARCH_MINB_RB
#define MARCH_MINB_RB 3
#endif
template <int MATH, int NDIM, int POT, bool REBASE>
__global__ void __launch_bounds__(256, REBASE ? MARCH_MINB_RB : MARCH_MINB) lattice_march_kernel(const LatticeArgs A) {
    if (*((volatile const u64 *)A.event_key) != NO_EVENT) return;  // an earlier launch must be replayed
    const int chain = blockIdx.z;
    int tl;
    unsigned bx;
    cta_slice_position(A, tl, bx);
    const bool edge_lo = A.slab_on && tl == 0, edge_hi = A.slab_on && tl == A.nt - 1;
    if (edge_lo | edge_hi) {
        if (threadIdx.x == 0) {
            if (edge_lo) slab_wait(A.wait_flag[0], A.wait_tag, A.slab_error);
            if (edge_hi) slab_wait(A.wait_flag[1], A.wait_tag, A.slab_error);
        }
        __syncthreads();
    }
    // ---- geometry: thread (tx, ty) owns strip position x0 = 4 tx of rows r_start .. r_start + R - 1 ----
    const unsigned L0 = (unsigned)A.dim[0], L1 = (unsigned)A.dim[1];
    const unsigned L2 = (NDIM >= 4) ? (unsigned)A.dim[2] : 1u;
    const unsigned tx = threadIdx.x & ((1u << A.m_tpr_log) - 1u), ty = threadIdx.x >> A.m_tpr_log;
    const unsigned R = (unsigned)A.m_R;
    const unsigned rows_per_cta = (256u >> A.m_tpr_log) * R;
    const unsigned r_start = bx * rows_per_cta + ty * R;
    unsigned x1s = r_start, x2 = 0;
    if (NDIM >= 4) {
        x2 = r_start / L1;
        x1s = r_start - x2 * L1;
    }
    const unsigned x0 = tx * 4u;
    const long long vs = A.vslice;
    const float *in = (const float *)A.in + (long long)chain * A.chain_stride;
    float *out = (float *)A.out + (long long)chain * A.chain_stride;
    const float *cur = in + (long long)tl * vs;
    const float *tm = (tl > 0) ? cur - vs : (A.wrap_time ? in + (long long)(A.nt - 1) * vs : (const float *)A.ghost_lo);
    const float *tp = (tl < A.nt - 1) ? cur + vs : (A.wrap_time ? in : (const float *)A.ghost_hi);
    float *dst = out + (long long)tl * vs;
    // opaque loop invariants: otherwise the register-capped compiler re-derives the bases from the
    // kernel arguments inside the loop and folds them into 64-bit element arithmetic per load
    asm volatile("" : "+l"(cur), "+l"(tm), "+l"(tp), "+l"(dst));
    __builtin_assume(__isGlobal(cur));
    __builtin_assume(__isGlobal(tm));
    __builtin_assume(__isGlobal(tp));
    __builtin_assume(__isGlobal(dst));
    const bool push_lo = edge_lo && A.push_tag, push_hi = edge_hi && A.push_tag;
    // loop-invariant neighbour offsets (in reals, relative to the strip's own offset o)
    // (unsigned wrap-around arithmetic: o + d is always a valid non-negative offset, so every
    // address is one IMAD.WIDE.U32 off a uniform base)
    const unsigned d_left = (x0 == 0) ? L0 - 1u : 0u - 1u;
    const unsigned d_right = (x0 + 4 == L0) ? 4u - L0 : 4u;
    const unsigned plane = L0 * L1;
    const unsigned d_up2 = (NDIM >= 4) ? ((x2 + 1 == L2) ? 0u - (L2 - 1) * plane : plane) : 0u;
    const unsigned d_dn2 = (NDIM >= 4) ? ((x2 == 0) ? (L2 - 1) * plane : 0u - plane) : 0u;
    const unsigned row_wrap = (L1 - 1) * L0;

    // ---- chain state: seed before the thread's first draw, per-row affine advance ----------------
    const u64 gslice = (u64)(A.slab_t0 + tl) * (u64)vs;
    const u64 S = A.seed_in[chain];
    unsigned o = r_start * L0 + x0;  // offset of the current strip inside the slice
    u64 g0 = gslice + o;
    unsigned cnt_prev = 0, nxt32 = 0x7FFFFFFFu;
    u64 S_eff = S;
    if (REBASE) {
        const Rebased rb = rebase_eval(A.rebase, A.n_rebase, chain, S, g0, gslice, (unsigned)vs);
        S_eff = rb.S_eff;
        cnt_prev = rb.cnt;
        nxt32 = rb.slow ? o : rb.nxt32;  // an entry inside the first strip: take the rare path at k == 0
    }
    // three precomputed jumps from gid 0: slice start, CTA's first row, this thread's first strip
    const u64 g_cta = gslice + (u64)bx * rows_per_cta * L0;
    u64 s;
    {
        const u64 s_sl = lcg_apply(A.slice_jump[tl], S_eff, 0) & LCG_MASK;
        const u64 s_cta = lcg_apply(A.cta_jump[bx], s_sl, gslice) & LCG_MASK;
        s = lcg_apply(A.thr_jump[threadIdx.x], s_cta, g_cta) & LCG_MASK;
    }
    const unsigned aDl = (unsigned)A.row_jump.a, aDh = (unsigned)(A.row_jump.a >> 32);
    u64 ck = (LCG_BETA * g0 + LCG_GAMMA) * A.row_jump.g0 + A.row_jump.bg1;  // s(next row) = alpha^L0 s + ck
    const u64 dck = LCG_BETA * (u64)L0 * A.row_jump.g0;
    u64 cg = site_const(g0);
    u64 dcg = (u64)L0 * LCG_A;
    u64 dck_ = dck;
    asm volatile("" : "+l"(dcg), "+l"(dck_));
    Seed32 s32 = seed_split(s);

    // ---- constants as fp32 pairs -----------------------------------------------------------------
    const float c_lap = (float)A.c_lap, c_dt = (float)A.c_dt;
    const float m2 = (float)(A.m2_chain ? A.m2_chain[chain] : A.m2);
    const float lam = (float)(A.lam_chain ? A.lam_chain[chain] : A.lam);
    const pair_t K_m2d = pk(-(float)(2 * NDIM), -(float)(2 * NDIM)), K_clap = pk(c_lap, c_lap);
    const pair_t K_m2cdt = pk(-2.0f * c_dt, -2.0f * c_dt), K_mcdt = pk(-c_dt, -c_dt);
    const pair_t K_lam = pk(lam, lam), K_m2 = pk(m2, m2);
    const pair_t K_2m32 = pk(2.3283064365386963e-10f, 2.3283064365386963e-10f), K_k2 = pk(A.k2_f, A.k2_f);
    // theta - pi = 2*3.1415*v2 - pi with v2 = (float)u2 * 2^-32
    const float kth = (float)(2.0 * 3.1415 / 4294967296.0);
    const pair_t K_th = pk(kth, kth), K_mpi = pk(-3.14159265358979f, -3.14159265358979f);

    const unsigned one = opaque_one();
    pair_t ACC1 = 0, ACC2 = 0;  // (+0.0f, +0.0f)
    unsigned nclamp = 0;

    unsigned k = 0;
    for (;;) {
    // The centre row is register-rolled: row k+1 is fetched one iteration early, serves as the x1+1
    // neighbour now and as the centre next time (and row k-1, last iteration's centre, as the x1-1
    // neighbour).  Two loads per strip less, and the thread's main first-touch stream -- otherwise
    // consumed by the very first add of the chain -- gets a whole iteration to arrive.
    ulonglong2 Cprev, Ccur;
    if (k < R) {
        const unsigned x1 = x1s + k;
        Cprev = *reinterpret_cast<const ulonglong2 *>(cur + (unsigned)(o + ((x1 == 0) ? row_wrap : 0u - L0)));
        Ccur = *reinterpret_cast<const ulonglong2 *>(cur + o);
    }
    // ---- hot inner loop: strips until the thread is done or (REBASE) reaches a replay entry -----
    for (; k < R && (!REBASE || (int)(nxt32 - o) > 4); ++k) {
        const unsigned x1 = x1s + k;
        // ---- loads: everything is cur/tm/tp + 32-bit offset ----------------------------------------
        const unsigned d_up1 = (x1 + 1 == L1) ? 0u - row_wrap : L0;
        const ulonglong2 Cnext = *reinterpret_cast<const ulonglong2 *>(cur + (unsigned)(o + d_up1));
        const ulonglong2 C = Ccur, U1 = Cnext, D1 = Cprev;
        ulonglong2 U2, D2;
        if (NDIM >= 4) {
            U2 = *reinterpret_cast<const ulonglong2 *>(cur + (unsigned)(o + d_up2));
            D2 = *reinterpret_cast<const ulonglong2 *>(cur + (unsigned)(o + d_dn2));
        }
        const ulonglong2 TP = *reinterpret_cast<const ulonglong2 *>(tp + o);
        const ulonglong2 TM = *reinterpret_cast<const ulonglong2 *>(tm + o);
        const float left = cur[(unsigned)(o + d_left)];
        const float right = cur[(unsigned)(o + d_right)];

        const Seed32 s_strip = s32;

        // ---- draws -----------------------------------------------------------------------------------
        unsigned u1[4], u2[4];
        unsigned umin = 0xFFFFFFFFu;
        {
            u64 c = cg;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                site_draw(s32, c, u1[e], u2[e]);
                if (e < 3) c = site_const_next(c, one);
                umin = min(min(umin, u1[e]), u2[e]);  // u1 == 0 (retry) or u2 < 2^15 (`seed+=`) => umin < 2^15
            }
        }

        // ---- deterministic part, packed: sites (0,1) and (2,3) -----------------------------------
        float c0, c1, c2, c3;
        upk(C.x, c0, c1);
        upk(C.y, c2, c3);
        pair_t S01 = pk(__fadd_rn(c1, left), __fadd_rn(c2, c0));  // phi(+0) + phi(-0)
        pair_t S23 = pk(__fadd_rn(c3, c1), __fadd_rn(right, c2));
        S01 = add2(S01, U1.x);
        S23 = add2(S23, U1.y);
        S01 = add2(S01, D1.x);
        S23 = add2(S23, D1.y);
        if (NDIM >= 4) {
            S01 = add2(S01, U2.x);
            S23 = add2(S23, U2.y);
            S01 = add2(S01, D2.x);
            S23 = add2(S23, D2.y);
        }
        S01 = add2(S01, TP.x);
        S23 = add2(S23, TP.y);
        S01 = add2(S01, TM.x);
        S23 = add2(S23, TM.y);
        pair_t V01 = fma2(K_clap, fma2(K_m2d, C.x, S01), C.x);
        pair_t V23 = fma2(K_clap, fma2(K_m2d, C.y, S23), C.y);
        if (POT == 4) {
            V01 = fma2(K_mcdt, mul2(C.x, fma2(K_lam, mul2(C.x, C.x), K_m2)), V01);
            V23 = fma2(K_mcdt, mul2(C.y, fma2(K_lam, mul2(C.y, C.y), K_m2)), V23);
        } else {
            V01 = fma2(K_m2cdt, C.x, V01);  // (-c_dt)(2 phi) == (-2 c_dt) phi exactly
            V23 = fma2(K_m2cdt, C.y, V23);
        }
        float v[4];
        upk(V01, v[0], v[1]);
        upk(V23, v[2], v[3]);

        // ---- noise ---------------------------------------------------------------------------------
        if (MATH == 1) {
            // r = cos(2*3.1415 v2) sqrt(-2 ln v1), amplitude folded under the root (k2 = 2 ln2 nscale^2):
            //   v1 = (float)u1 * 2^-32 exact scaling of the RN conversion, MUFU.LG2, MUFU.SQRT;
            //   cos(theta) = -cos(theta - pi) keeps MUFU.COS in [-pi, pi); the final multiply-add is fused
            const pair_t F1a = pk(__uint2float_rn(u1[0]), __uint2float_rn(u1[1]));
            const pair_t F1b = pk(__uint2float_rn(u1[2]), __uint2float_rn(u1[3]));
            const pair_t F2a = pk(__uint2float_rn(u2[0]), __uint2float_rn(u2[1]));
            const pair_t F2b = pk(__uint2float_rn(u2[2]), __uint2float_rn(u2[3]));
            float a[4], th[4];
            upk(mul2(F1a, K_2m32), a[0], a[1]);
            upk(mul2(F1b, K_2m32), a[2], a[3]);
            const pair_t Ta = mul2(pk(lg2_approx(a[0]), lg2_approx(a[1])), K_k2);
            const pair_t Tb = mul2(pk(lg2_approx(a[2]), lg2_approx(a[3])), K_k2);
            float t[4];
            upk(Ta, t[0], t[1]);
            upk(Tb, t[2], t[3]);
            upk(fma2(F2a, K_th, K_mpi), th[0], th[1]);
            upk(fma2(F2b, K_th, K_mpi), th[2], th[3]);
#pragma unroll
            // dw = -RN(cos(theta - pi) rad), then v + dw: the two roundings every lattice kernel uses
            for (int e = 0; e < 4; ++e) v[e] = __fsub_rn(v[e], __fmul_rn(__cosf(th[e]), sqrt_approx(fabsf(t[e]))));
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                v[e] = __fadd_rn(v[e], (float)__dmul_rn(A.nscale, noise_accurate((u64)u1[e] << 16, (u64)u2[e] << 16)));
        }

        // ---- rare paths: possible RNG event in this strip, values at or beyond the clamp ---------
        const float amax = fmaxf(fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fabsf(v[2])), fabsf(v[3]));
        if (__builtin_expect((umin < 32768u) | !(amax < 1000.0f), 0)) {
            bool replayed = false;
            if (umin < 32768u) replayed = strip_events_cold(A.event_key, A.step_index, chain, seed_join(s_strip), g0, 4);
            const Clamped cl = clamp_cold(v[0], v[1], v[2], v[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = cl.v[e];
            if (!replayed) nclamp += cl.n;
        }

        // ---- observables of the pre-update field, store ------------------------------------------
        ACC1 = add2(ACC1, C.x);
        ACC1 = add2(ACC1, C.y);
        ACC2 = fma2(C.x, C.x, ACC2);
        ACC2 = fma2(C.y, C.y, ACC2);
        const float4 res = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(dst + o) = res;
        if (__builtin_expect(push_lo | push_hi, 0)) {  // CTA-uniform: boundary slices of a slab ring only
            if (push_lo) *reinterpret_cast<float4 *>((float *)A.push_ghost[0] + o) = res;
            if (push_hi) *reinterpret_cast<float4 *>((float *)A.push_ghost[1] + o) = res;
        }

        Cprev = Ccur;
        Ccur = Cnext;
        // ---- next row: same x0, L0 draws further ---------------------------------------------------
        {
            const u64 p = (u64)s_strip.lo * aDl + ck;
            s32.lo = (unsigned)p;
            s32.hi = (unsigned)(p >> 32) + s_strip.lo * aDh + s_strip.hi * aDl;
        }
        ck += dck_;
        cg += dcg;
        g0 += L0;
        o += L0;
    }
    if (!REBASE || k >= R) break;
        // ---- seed of this strip under replay entries ---------------------------------------------
        // Entries are sorted by gid and a thread visits its strips in increasing gid, so it only has
        // to watch the distance to the NEXT entry (32-bit, relative to the slice): two instructions
        // per strip.  Reaching one (rare) re-evaluates the base with the full 64-bit logic.
        if (REBASE) {
            {   // an entry lies at or before strip k
                const unsigned x1 = x1s + k;
                // (everything it needs is re-derived here rather than kept live across the hot loop)
                const u64 gsl = (u64)(A.slab_t0 + tl) * (u64)A.vslice;
                const Rebased rb = rebase_eval(A.rebase, A.n_rebase, chain, A.seed_in[chain], g0, gsl, (unsigned)A.vslice);
                nxt32 = rb.nxt32;
                if (rb.cnt != cnt_prev) {  // new base: the thread's first strip under the new start seed, k rows down
                    cnt_prev = rb.cnt;
                    const u64 gc = gsl + (u64)bx * rows_per_cta * L0;
                    const u64 s_sl = lcg_apply(A.slice_jump[tl], rb.S_eff, 0) & LCG_MASK;
                    const u64 s_cta = lcg_apply(A.cta_jump[bx], s_sl, gsl) & LCG_MASK;
                    Seed32 t = seed_split(lcg_apply(A.thr_jump[threadIdx.x], s_cta, gc) & LCG_MASK);
                    u64 cj = (LCG_BETA * (g0 - (u64)k * L0) + LCG_GAMMA) * A.row_jump.g0 + A.row_jump.bg1;
                    for (unsigned j = 0; j < k; ++j) {
                        const u64 p = (u64)t.lo * aDl + cj;
                        t.hi = (unsigned)(p >> 32) + t.lo * aDh + t.hi * aDl;
                        t.lo = (unsigned)p;
                        cj += dck;
                    }
                    s32 = t;
                }
                if (rb.slow) {  // the whole strip out of line; the row recurrence is void behind an entry
                    SlowIn I;
                    I.cur = cur; I.tm = tm; I.tp = tp; I.dst = dst;
                    I.push0 = push_lo ? (float *)A.push_ghost[0] : nullptr;
                    I.push1 = push_hi ? (float *)A.push_ghost[1] : nullptr;
                    I.o = o;
                    I.o_up1 = o + ((x1 + 1 == L1) ? 0u - row_wrap : L0);
                    I.o_dn1 = o + ((x1 == 0) ? row_wrap : 0u - L0);
                    I.o_up2 = o + d_up2; I.o_dn2 = o + d_dn2; I.o_left = o + d_left; I.o_right = o + d_right;
                    I.s = seed_join(s32); I.g0 = g0;
                    I.chain = chain; I.step_index = A.step_index; I.n_rebase = A.n_rebase;
                    I.rebase = A.rebase; I.event_key = A.event_key;
                    I.c_lap = c_lap; I.c_dt = c_dt; I.m2 = m2; I.lam = lam; I.k2 = A.k2_f; I.nscale = A.nscale;
                    const SlowOut so = strip_slow<MATH, NDIM, POT>(I);
                    ACC1 = add2(ACC1, pk(so.a1, 0.f));
                    ACC2 = add2(ACC2, pk(so.a2, 0.f));
                    nclamp += so.nclamp;
                    nxt32 = o + L0;  // re-evaluate at the next strip (new base)
                    ck += dck_;
                    cg += dcg;
                    g0 += L0;
                    o += L0;
                    ++k;
                }
            }
        }
    }

    if (push_lo || push_hi) {  // last CTA of the slice: everything is out, raise the neighbour's flag
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (push_lo && atomicAdd(A.push_count + 0, 1u) == gridDim.x - 1) {
                A.push_count[0] = 0;
                __threadfence_system();
                st_release_sys_u32(A.push_flag[0], A.push_tag);
            }
            if (push_hi && atomicAdd(A.push_count + 1, 1u) == gridDim.x - 1) {
                A.push_count[1] = 0;
                __threadfence_system();
                st_release_sys_u32(A.push_flag[1], A.push_tag);
            }
        }
    }

    // ---- the omega work-item's draw (gid = V) and the step's final seed -------------------------
    if (bx == 0 && tl == 0 && threadIdx.x == 0) {
        const u64 Vg = (u64)A.V;
        u64 sv, t1, t2;
        bool overridden = false;
        u64 next = 0;
        if (REBASE) {
            u64 bg, bs;
            rebase_lookup(A, chain, S, Vg, bg, bs);
            sv = lcg_seed_at(bs, bg, Vg - bg, A.jump);
            for (int j = 0; j < A.n_rebase; ++j)
                if (A.rebase[j].chain == chain && A.rebase[j].ov_gid == Vg) {
                    overridden = true;
                    next = A.rebase[j].seed;  // entry with gid_start == V+1
                }
        } else {
            sv = lcg_apply(A.vol_jump, S, 0) & LCG_MASK;
        }
        lcg_draw(sv, Vg, t1, t2);
        if (!overridden) {
            if (lcg_event(sv & LCG_MASK, t1, t2))
                atomicMin((unsigned long long *)A.event_key, event_key(A.step_index, chain, Vg));
            next = lcg_next_seed(t2);
        }
        A.seed_out[chain] = next;
    }

    // ---- per-CTA observable partial ---------------------------------------------------------------
    if (A.partials) {
        __shared__ double red[2][8];
        float a1l, a1h, a2l, a2h;
        upk(ACC1, a1l, a1h);
        upk(ACC2, a2l, a2h);
        double a1 = warp_sum((double)a1l + (double)a1h), a2 = warp_sum((double)a2l + (double)a2h);
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) { red[0][w] = a1; red[1][w] = a2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s1 = 0, s2 = 0;
            for (int q = 0; q < 8; ++q) { s1 += red[0][q]; s2 += red[1][q]; }
            double *p = A.partials + (((long long)chain * A.nt + tl) * gridDim.x + bx) * 2;
            p[0] = s1;
            p[1] = s2;
        }
    }
    if (nclamp) atomicAdd(A.nclamped, (unsigned long long)nclamp);
}

template <int MATH, int NDIM, int POT>
static cudaError_t march_rb(const LatticeArgs &A, dim3 grid, cudaStream_t st) {
    if (A.n_rebase > 0) lattice_march_kernel<MATH, NDIM, POT, true><<<grid, 256, 0, st>>>(A);
    else lattice_march_kernel<MATH, NDIM, POT, false><<<grid, 256, 0, st>>>(A);
    return cudaGetLastError();
}
template <int MATH, int NDIM>
static cudaError_t march_pot(const LatticeArgs &A, dim3 grid, cudaStream_t st) {
    return A.pot == 4 ? march_rb<MATH, NDIM, 4>(A, grid, st) : march_rb<MATH, NDIM, 0>(A, grid, st);
}

cudaError_t launch_lattice_march(const LatticeArgs &A, int math, int ctas_per_slice, cudaStream_t stream) {
    dim3 grid((unsigned)ctas_per_slice, (unsigned)A.nt, (unsigned)A.nchains);
    if (A.ndim == 3) return math ? march_pot<1, 3>(A, grid, stream) : march_pot<0, 3>(A, grid, stream);
    if (A.ndim == 4) return math ? march_pot<1, 4>(A, grid, stream) : march_pot<0, 4>(A, grid, stream);
    return cudaErrorInvalidValue;
}

}  // namespace sq
