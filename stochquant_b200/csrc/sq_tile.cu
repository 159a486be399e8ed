// sq_tile.cu -- streaming Langevin step, fp32, d = 3 and 4: tile staging through shared memory by bulk
// asynchronous copies, second design of the marching kernel (sq_march.cu).
//
// Same update as lattice_step_kernel / lattice_march_kernel (generalisation of tau_kernel.cl:64-173 defined in
// DESIGN.md section 4 / oracle sqo_lattice_step) and the same integer stream.  The marching kernel was
// instruction-bound (58 warp-instructions per site, profiles/r01_ring_c4s.txt): 18 of every strip's
// instructions computed load addresses, 8 rolled the centre row through registers, and the two first-touch
// FADD2s of the neighbour sum held a quarter of all stall samples.  Here
//   * a CTA's tile -- rows_per_cta consecutive rows of one time slice -- is fetched ONCE into shared memory by
//     cp.async.bulk copies (one per run of rows inside an (x2, t) plane, plus that run's two x1-halo rows, with
//     the periodic wrap resolved by the copy's source address), completion on an mbarrier.  The centre strip,
//     its x1+-1 neighbours and the two x0 neighbours then are LDS at base + immediate: no address arithmetic,
//     no L1 round trip, no register rolling;
//   * the four streams without reuse (t+-1, x2+-1) stay 128-bit global loads, issued at the top of a pass and
//     consumed after the pass's noise phase: the draws and the Box-Muller arithmetic (2/3 of the instructions,
//     no field data) run while they are in flight.  Their addresses are per-thread 64-bit bases + an immediate
//     (the row stride is a template parameter for rows of 32 / 64 / 256 sites);
//   * the LCG is walked in its t2 form (see sq_rowres.cu): two independent 48-bit multiply-adds per site;
//   * a thread's strip is four consecutive sites (the template also takes eight: 124 registers, two CTAs per SM, measured
//     slower and not instantiated);
//   * what does not depend on the thread is computed once per CTA (TileHdr), what does not depend on the CTA once per
//     context (TileThread table): a thread's prologue is ~135 instructions instead of ~455;
//   * the tile's share of the four direct streams is asked into L2 (cp.async.bulk.prefetch.L2) when its copies are issued.
// CTA -> tile mapping, jump tables, per-tile observable partials, clamp slots, L2 chunking and the slab ring's halo
// protocol are those of the marching kernel, so the host side is shared; steps that carry replay entries run THERE.
// Second kernel in this file: lattice_rows_kernel, the row-block staging form (every operand through shared memory,
// producer warp, tiles claimed from a counter), behind SQ_FLAG_ROWBLOCK_KERNEL -- see its header comment below.
#include <stdio.h>
#include <stdlib.h>

#include "sq_strip_slow.cuh"

namespace sq {

namespace {

constexpr unsigned ALPHA_LO32T = (unsigned)LCG_ALPHA, ALPHA_HI32T = (unsigned)(LCG_ALPHA >> 32);

__device__ __forceinline__ void mad48t(unsigned xl, unsigned xh, unsigned ml, unsigned mh, u64 c, unsigned &rl, unsigned &rh) {
    u64 p;
    unsigned ph, t;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(xl), "r"(ml), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(rl), "=r"(ph) : "l"(p));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(xl), "r"(mh), "r"(ph));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(rh) : "r"(xh), "r"(ml), "r"(t));
}
__device__ __forceinline__ ulonglong2 lds128(unsigned a) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ ulonglong2 ldg128(const char *p) {
    ulonglong2 v;
    asm volatile("ld.global.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void tile_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tile_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA engine, no tensor map: a linear run of bytes); bytes % 16 == 0, both sides 16-byte aligned
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// ask L2 for a linear run of bytes ahead of the loads that will read it (no destination, no completion to wait for)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tile_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

}  // namespace

// What thread 0 works out once for the CTA (the other 255 threads used to repeat all of it: 2/3 of the ~450 instructions a
// thread spent before its first strip, 14 per site on 64^4 -- profiles/r02_c3_tile.txt).
struct TileHdr {
    int skip, tl;          // an earlier launch flagged an event ; the CTA's time slice
    unsigned bx, x2;       // tile index inside the slice ; x2 of the tile's first row
    u64 s_cta, scg;        // seed before the tile's first draw ; BETA g_cta + GAMMA (middle term of a table jump from there)
    u64 c1, c2, ck;        // site constants and row-advance constant of the tile's FIRST site
    u64 g_cta;             // its gid
    const char *cur, *tp, *tm;  // the tile's first site in the slice, the slice above, the slice below
    char *dst;
    const float *cur_slice;
    unsigned o_cta, pad;   // offset of the tile's first site inside the slice (reals)
};

// L0T: row length known at compile time (0: runtime) -- the byte stride between a thread's passes becomes an immediate
// NP: packed pairs per strip (2: four sites, 4: eight sites per thread and pass)
// (steps with replay entries are launched on the marching kernel: this one carries none of that code)
#ifndef TILE_MINB
#define TILE_MINB 3
#endif
template <int MATH, int NDIM, int POT, int L0T, int NP>
__global__ void __launch_bounds__(256, NP == 4 ? 2 : TILE_MINB) lattice_tile_kernel(const LatticeArgs A) {
    constexpr unsigned W = 2u * NP, NH = NP / 2u;  // sites per strip, float4 per strip
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ TileHdr H;
    const int chain = blockIdx.z;
    const unsigned L0 = L0T ? (unsigned)L0T : (unsigned)A.dim[0], L1 = (unsigned)A.dim[1];
    const unsigned L2 = (NDIM >= 4) ? (unsigned)A.dim[2] : 1u;
    const unsigned ROWB = L0 * 4u;  // bytes per row
    const unsigned R = (unsigned)A.m_R;
    const unsigned rows_per_cta = (256u >> A.m_tpr_log) * R;
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(tile_smem);
    const unsigned bar = smem0, data0 = smem0 + 128u;

    // ---- the thread's place inside a tile does not depend on the CTA: one 64-byte table entry (host-built), requested
    // now, used behind the barrier ----
    const char *te = (const char *)(A.tile_thr + threadIdx.x);
    const ulonglong2 e0 = ldg128(te), e1 = ldg128(te + 16), e2 = ldg128(te + 32);
    unsigned e_plane = 0;
    if (NDIM >= 4) e_plane = A.tile_thr[threadIdx.x].plane;

    // ---- the CTA's place: thread 0 ---------------------------------------------------------------------------
    // Programmatic dependent launch (launch_lattice_tile sets the attribute): the NEXT step's CTAs may take the slots this
    // grid's last wave leaves empty and run up to here -- table entry requested, position worked out -- while this grid
    // finishes; everything below reads what the previous launch wrote (event word, seed, field) and waits for it.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int tl0 = 0;
    unsigned bx0 = 0;
    if (threadIdx.x == 0) {
        cta_slice_position(A, tl0, bx0);
        tile_mbar_init(bar, 1);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) {
        const int tl = tl0;
        const unsigned bx = bx0;
        // an earlier launch flagged an event: this one will be replayed.  ONE thread decides for the CTA -- the word may
        // rise between two threads' reads, and a CTA that has lost the warp which issues its copies would wait forever.
        H.skip = *((volatile const u64 *)A.event_key) != NO_EVENT;
        if (A.slab_on) {  // slab ring: the neighbour's boundary slice of this step's input must have landed
            if (tl == 0) slab_wait(A.wait_flag[0], A.wait_tag, A.slab_error);
            if (tl == A.nt - 1) slab_wait(A.wait_flag[1], A.wait_tag, A.slab_error);
        }
        const long long vs = A.vslice;
        const u64 gslice = (u64)(A.slab_t0 + tl) * (u64)vs;
        const unsigned o_cta = bx * rows_per_cta * L0;
        const u64 g_cta = gslice + o_cta;
        // two precomputed jumps from gid 0: slice start, tile start (the third, to the thread's strip, is per thread)
        const u64 s_sl = lcg_apply(A.slice_jump[tl], A.seed_in[chain], 0) & LCG_MASK;
        H.s_cta = lcg_apply(A.cta_jump[bx], s_sl, gslice) & LCG_MASK;
        const u64 scg = LCG_BETA * g_cta + LCG_GAMMA;
        H.scg = scg;
        // T(next row) = alpha^L0 T + ck ; ck itself is a running sum
        H.ck = scg * A.row_jump.g0 + A.row_jump.bg1 - A.row_jump.a * TWO31 + TWO31;
        // t1 = A T + c1, T' = A^2 T + c2 ; +A / +(A^2+A) per site
        H.c1 = site_const(g_cta) - LCG_A * TWO31;
        H.c2 = (LCG_A + 1) * site_const(g_cta) - LCG_ALPHA * TWO31;
        H.g_cta = g_cta;
        const float *in = (const float *)A.in + (long long)chain * A.chain_stride;
        const float *cur = in + (long long)tl * vs;
        const float *tm = (tl > 0) ? cur - vs : (A.wrap_time ? in + (long long)(A.nt - 1) * vs : (const float *)A.ghost_lo);
        const float *tp = (tl < A.nt - 1) ? cur + vs : (A.wrap_time ? in : (const float *)A.ghost_hi);
        H.cur_slice = cur;
        H.cur = (const char *)(cur + o_cta);
        H.tp = (const char *)(tp + o_cta);
        H.tm = (const char *)(tm + o_cta);
        H.dst = (char *)((float *)A.out + (long long)chain * A.chain_stride + (long long)tl * vs + o_cta);
        H.tl = tl;
        H.bx = bx;
        H.x2 = (NDIM >= 4) ? (bx * rows_per_cta) / L1 : 0u;
        H.o_cta = o_cta;
    }
    __syncthreads();
    if (H.skip) return;
    const int tl = H.tl;

    // ---- stage the tile: segments = runs of rows inside one plane, each stored as [halo_lo | rows | halo_hi] ---
    if (threadIdx.x < 32) {
        const unsigned seg_rows = rows_per_cta < L1 ? rows_per_cta : L1, nseg = rows_per_cta / seg_rows;
        const unsigned seg_bytes = (seg_rows + 2u) * ROWB;
        const unsigned tile_row0 = H.bx * rows_per_cta;
        const float *cur = H.cur_slice;
        if (threadIdx.x == 0) tile_expect_tx(bar, nseg * seg_bytes);
        __syncwarp();
        for (unsigned c = threadIdx.x; c < 3u * nseg; c += 32u) {
            const unsigned sg = c / 3u, part = c - 3u * sg;            // 0 rows, 1 halo below, 2 halo above
            const unsigned row0 = tile_row0 + sg * seg_rows;            // first row of the run (row index inside the slice)
            const unsigned x1a = row0 % L1, prow0 = row0 - x1a;         // its x1, first row of its plane
            unsigned src_row, dst_off, bytes;
            if (part == 0) { src_row = row0; dst_off = ROWB; bytes = seg_rows * ROWB; }
            else if (part == 1) { src_row = (x1a == 0) ? prow0 + L1 - 1u : row0 - 1u; dst_off = 0; bytes = ROWB; }
            else { src_row = (x1a + seg_rows == L1) ? prow0 : row0 + seg_rows; dst_off = (seg_rows + 1u) * ROWB; bytes = ROWB; }
            bulk_g2s(data0 + sg * seg_bytes + dst_off, cur + (size_t)src_row * L0, bytes, bar);
        }
#ifndef TILE_NO_PREFETCH
        // the four streams without reuse are read straight from global memory, a pass at a time: ask L2 for the tile's
        // share of them now (the slice above / below: one run; the planes above / below: one run per segment)
        if (threadIdx.x == 30) bulk_prefetch_l2(H.tp, rows_per_cta * ROWB);
        if (threadIdx.x == 31) bulk_prefetch_l2(H.tm, rows_per_cta * ROWB);
        if (NDIM >= 4) {
            const long long planeR = (long long)L0 * L1, wrapR = (long long)(L2 - 1) * planeR;
            for (unsigned c = threadIdx.x; c < 2u * nseg; c += 32u) {
                const unsigned sg = c >> 1, row0 = tile_row0 + sg * seg_rows, x2s = row0 / L1;
                const float *base = cur + (size_t)row0 * L0;
                const float *q = (c & 1) ? base + ((x2s == 0) ? wrapR : -planeR) : base + ((x2s + 1 == L2) ? -wrapR : planeR);
                bulk_prefetch_l2(q, seg_rows * ROWB);
            }
        }
#endif
    }

    // ---- the thread's first strip: stream bases, staged-tile addresses, chain state ---------------------------
    const unsigned thr_off = (unsigned)e2.x;   // first site relative to the tile's first site (reals)
    const u64 bo = (u64)thr_off * 4u;
    // per-thread stream bases (bytes); pass k adds k * ROWB
    const char *p_tp = H.tp + bo, *p_tm = H.tm + bo;
    const char *p_u2 = nullptr, *p_d2 = nullptr;
    if (NDIM >= 4) {
        const unsigned x2 = H.x2 + e_plane;
        const long long planeB = (long long)L0 * L1 * 4, wrapB = (long long)(L2 - 1) * planeB;
        p_u2 = H.cur + bo + ((x2 + 1 == L2) ? -wrapB : planeB);
        p_d2 = H.cur + bo + ((x2 == 0) ? wrapB : -planeB);
    }
    char *p_dst = H.dst + bo;
    // shared-memory addresses of the thread's first strip and its x0 neighbours
    unsigned s_c = smem0 + (unsigned)(e2.x >> 32), s_left = smem0 + (unsigned)e2.y, s_right = smem0 + (unsigned)(e2.y >> 32);
    const bool push_lo = A.slab_on && tl == 0 && A.push_tag, push_hi = A.slab_on && tl == A.nt - 1 && A.push_tag;

    // ---- chain state: T = seed before the thread's first draw + 2^31, per-row affine advance ------------------
    unsigned Tl, Th;
    {
        const u64 s = (e0.x * H.s_cta + H.scg * e0.y + e1.x) & LCG_MASK;  // the table jump over thr_off draws
        const u64 T = s + TWO31;
        Tl = (unsigned)T;
        Th = (unsigned)(T >> 32);
    }
    const unsigned aDl = (unsigned)A.row_jump.a, aDh = (unsigned)(A.row_jump.a >> 32);
    u64 ck = H.ck + e1.y;               // + BETA thr_off row_jump.g0
    u64 c1 = H.c1 + LCG_A * thr_off;    // site constants of the strip's first site
    u64 c2 = H.c2 + LCG_BETA * thr_off;
    // (the per-row increments A.t_dck, A.t_dc1 = (L0 - (W-1)) A, A.t_dc2 = (L0 - (W-1))(A^2 + A) are kernel parameters:
    // constant-bank operands of the adds, not registers)

    // ---- constants as fp32 pairs -------------------------------------------------------------------------
    const float c_lap = (float)A.c_lap, c_dt = (float)A.c_dt;
    const float m2 = (float)(A.m2_chain ? A.m2_chain[chain] : A.m2);
    const float lam = (float)(A.lam_chain ? A.lam_chain[chain] : A.lam);
    const pair_t K_m2d = pk(-(float)(2 * NDIM), -(float)(2 * NDIM)), K_clap = pk(c_lap, c_lap);
    const pair_t K_m2cdt = pk(-2.0f * c_dt, -2.0f * c_dt), K_mcdt = pk(-c_dt, -c_dt);
    const pair_t K_lam = pk(lam, lam), K_m2 = pk(m2, m2), K_m1 = pk(-1.0f, -1.0f);
    const pair_t K_2m32 = pk(2.3283064365386963e-10f, 2.3283064365386963e-10f), K_k2 = pk(A.k2_f, A.k2_f);
    const float kth = (float)(2.0 * 3.1415 / 4294967296.0);  // theta - pi = 2*3.1415 v2 - pi, v2 = (float)u2 * 2^-32
    const pair_t K_th = pk(kth, kth), K_mpi = pk(-3.14159265358979f, -3.14159265358979f);
    (void)K_m2cdt; (void)K_mcdt; (void)K_lam; (void)K_m2; (void)K_2m32; (void)K_k2; (void)K_th; (void)K_mpi;

    pair_t ACC1 = 0, ACC2 = 0;  // (+0.0f, +0.0f)
    unsigned nclamp = 0;

    // ---- one pass = one strip of W sites; `kb` = byte offset of the pass inside the thread's streams, k its row ----
    auto pass = [&](const unsigned kb, const unsigned k) {
        // t+-1 and x2+-1 have no reuse: straight from global memory, requested now, used after the noise phase
        ulonglong2 U2[NH], D2[NH], TP[NH], TM[NH];
#pragma unroll
        for (unsigned h = 0; h < NH; ++h) {
            if (NDIM >= 4) {
                U2[h] = ldg128(p_u2 + kb + 16u * h);
                D2[h] = ldg128(p_d2 + kb + 16u * h);
            }
            TP[h] = ldg128(p_tp + kb + 16u * h);
            TM[h] = ldg128(p_tm + kb + 16u * h);
        }

        // ---- noise phase: W draws in t2 form, Box-Muller ------------------------------------------------
        const unsigned T0l = Tl, T0h = Th;
        unsigned tl_ = Tl, th_ = Th, um = 0xFFFFFFFFu;
        pair_t NZ[NP];
#pragma unroll
        for (unsigned q = 0; q < (unsigned)NP; ++q) {
            unsigned u1a, u2a, u1b, u2b, al, ah, bl, bh;
            mad48t(tl_, th_, A_LO, A_HI, c1, al, ah);
            mad48t(tl_, th_, ALPHA_LO32T, ALPHA_HI32T, c2, bl, bh);
            u1a = __funnelshift_r(al, ah, 16);
            u2a = __funnelshift_r(bl, bh, 16);
            c1 += LCG_A;
            c2 += LCG_BETA;
            mad48t(bl, bh, A_LO, A_HI, c1, al, ah);
            mad48t(bl, bh, ALPHA_LO32T, ALPHA_HI32T, c2, tl_, th_);
            u1b = __funnelshift_r(al, ah, 16);
            u2b = __funnelshift_r(tl_, th_, 16);
            if (q + 1 < (unsigned)NP) {
                c1 += LCG_A;
                c2 += LCG_BETA;
            } else {  // on to the first site of the next row
                c1 += A.t_dc1;
                c2 += A.t_dc2;
            }
            um = min(min(um, u1a), u2a);  // u1 == 0 (retry) or u2 < 2^15 (`seed+=`) => um < 2^15
            um = min(min(um, u1b), u2b);
            if (MATH == 1) {
                float l1a, l1b, ta, tb, tha, thb;
                upk(mul2(pk(__uint2float_rn(u1a), __uint2float_rn(u1b)), K_2m32), l1a, l1b);
                upk(mul2(pk(lg2_approx(l1a), lg2_approx(l1b)), K_k2), ta, tb);
                upk(fma2(pk(__uint2float_rn(u2a), __uint2float_rn(u2b)), K_th, K_mpi), tha, thb);
                NZ[q] = mul2(pk(__cosf(tha), __cosf(thb)), pk(sqrt_approx(fabsf(ta)), sqrt_approx(fabsf(tb))));
            } else {
                const float da = (float)__dmul_rn(A.nscale, noise_accurate((u64)u1a << 16, (u64)u2a << 16));
                const float db = (float)__dmul_rn(A.nscale, noise_accurate((u64)u1b << 16, (u64)u2b << 16));
                NZ[q] = pk(-da, -db);
            }
        }
        {   // next row: same x0, L0 draws further
            mad48t(T0l, T0h, aDl, aDh, ck, Tl, Th);
            ck += A.t_dck;
        }

        // ---- stencil phase ------------------------------------------------------------------------------------
        pair_t C[NP], U1[NP], D1[NP];
#pragma unroll
        for (unsigned h = 0; h < NH; ++h) {
            const ulonglong2 c = lds128(s_c + kb + 16u * h), u = lds128(s_c + kb + ROWB + 16u * h), d = lds128(s_c + kb - ROWB + 16u * h);
            C[2 * h] = c.x; C[2 * h + 1] = c.y;
            U1[2 * h] = u.x; U1[2 * h + 1] = u.y;
            D1[2 * h] = d.x; D1[2 * h + 1] = d.y;
        }
        const float left = lds32(s_left + kb), right = lds32(s_right + kb);
        float p[W];
#pragma unroll
        for (unsigned q = 0; q < (unsigned)NP; ++q) upk(C[q], p[2 * q], p[2 * q + 1]);
        pair_t V[NP];
        float amax = 0.f;
#pragma unroll
        for (unsigned q = 0; q < (unsigned)NP; ++q) {
            const float xm0 = (q == 0) ? left : p[(2 * q + W - 1) % W], xp1 = (q == (unsigned)NP - 1) ? right : p[(2 * q + 2) % W];
            pair_t Sq = pk(__fadd_rn(p[2 * q + 1], xm0), __fadd_rn(xp1, p[2 * q]));  // phi(+0) + phi(-0)
            Sq = add2(Sq, U1[q]);
            Sq = add2(Sq, D1[q]);
            if (NDIM >= 4) {
                Sq = add2(Sq, (q & 1) ? U2[q / 2].y : U2[q / 2].x);
                Sq = add2(Sq, (q & 1) ? D2[q / 2].y : D2[q / 2].x);
            }
            Sq = add2(Sq, (q & 1) ? TP[q / 2].y : TP[q / 2].x);
            Sq = add2(Sq, (q & 1) ? TM[q / 2].y : TM[q / 2].x);
            pair_t v = fma2(K_clap, fma2(K_m2d, C[q], Sq), C[q]);
            if (POT == 4) v = fma2(K_mcdt, mul2(C[q], fma2(K_lam, mul2(C[q], C[q]), K_m2)), v);
            else v = fma2(K_m2cdt, C[q], v);  // (-c_dt)(2 phi) == (-2 c_dt) phi exactly
            v = fma2(K_m1, NZ[q], v);          // v + dw, one rounding
            V[q] = v;
            float a0, a1;
            upk(v, a0, a1);
            amax = fmaxf(fmaxf(fabsf(a0), fabsf(a1)), amax);
            ACC1 = add2(ACC1, C[q]);           // observables of the pre-update field
            ACC2 = fma2(C[q], C[q], ACC2);
        }
        // clamp (tau_kernel.cl:122-132) and RNG events: one test per strip for both rare cases
        if (__builtin_expect(!(amax < 1000.0f) | (um < 32768u), 0)) {
            bool replayed = false;  // an event in this strip: the launch is redone, its clamp hits are not counted
            if (um < 32768u) {
                const u64 z0 = ((((u64)T0h << 32) | T0l) - TWO31) & LCG_MASK;
                replayed = strip_events_cold(A.event_key, A.step_index, chain, z0, H.g_cta + thr_off + k * L0, (int)W);
            }
#pragma unroll
            for (unsigned h = 0; h < NH; ++h) {
                float v0, v1, v2, v3;
                upk(V[2 * h], v0, v1);
                upk(V[2 * h + 1], v2, v3);
                const Clamped cl = clamp_cold(v0, v1, v2, v3);
                V[2 * h] = pk(cl.v[0], cl.v[1]);
                V[2 * h + 1] = pk(cl.v[2], cl.v[3]);
                if (!replayed) nclamp += cl.n;
            }
        }
#pragma unroll
        for (unsigned h = 0; h < NH; ++h) *reinterpret_cast<ulonglong2 *>(p_dst + kb + 16u * h) = make_ulonglong2(V[2 * h], V[2 * h + 1]);
        if (__builtin_expect(push_lo | push_hi, 0)) {  // CTA-uniform: boundary slices of a slab ring only
            const size_t oo = (size_t)(H.o_cta + thr_off + k * L0) * 4u;
#pragma unroll
            for (unsigned h = 0; h < NH; ++h) {
                if (push_lo) *reinterpret_cast<ulonglong2 *>((char *)A.push_ghost[0] + oo + 16u * h) = make_ulonglong2(V[2 * h], V[2 * h + 1]);
                if (push_hi) *reinterpret_cast<ulonglong2 *>((char *)A.push_ghost[1] + oo + 16u * h) = make_ulonglong2(V[2 * h], V[2 * h + 1]);
            }
        }
    };

    tile_mbar_wait(bar, 0);  // the tile has landed (other CTAs of the SM computed meanwhile)
    // ---- the thread's R strips: U passes per trip (pass offsets are immediates), then the stream bases move on --------
    // (R is a multiple of 4 -- tile_shape_ok -- so there is no tail test)
    constexpr unsigned U = NP == 4 ? 2u : 4u;
    for (unsigned k0 = 0; k0 < R; k0 += U) {
#pragma unroll
        for (unsigned u = 0; u < U; ++u) pass(u * ROWB, k0 + u);
        p_tp += U * ROWB;
        p_tm += U * ROWB;
        if (NDIM >= 4) {
            p_u2 += U * ROWB;
            p_d2 += U * ROWB;
        }
        p_dst += U * ROWB;
        // (opaque: otherwise the four streams are rewritten as ONE running offset plus four bases, and every load pays a
        // 64-bit add again instead of using base + immediate)
        asm volatile("" : "+l"(p_tp), "+l"(p_tm), "+l"(p_u2), "+l"(p_d2), "+l"(p_dst));
        s_c += U * ROWB;
        s_left += U * ROWB;
        s_right += U * ROWB;
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

    // ---- the omega work-item's draw (gid = V) and the step's final seed -------------------------------------
    if (threadIdx.x == 0 && H.bx == 0 && tl == 0) {
        const u64 Vg = (u64)A.V;
        u64 t1, t2;
        const u64 sv = lcg_apply(A.vol_jump, A.seed_in[chain], 0) & LCG_MASK;
        lcg_draw(sv, Vg, t1, t2);
        if (lcg_event(sv & LCG_MASK, t1, t2)) atomicMin((unsigned long long *)A.event_key, event_key(A.step_index, chain, Vg));
        A.seed_out[chain] = lcg_next_seed(t2);
    }

    // ---- per-CTA observable partial ---------------------------------------------------------------------------
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
            double *p = A.partials + (((long long)chain * A.nt + tl) * gridDim.x + H.bx) * 2;
            p[0] = s1;
            p[1] = s2;
        }
    }
    if (nclamp) atomicAdd(A.nclamped, (unsigned long long)nclamp);
}

// =====================================================================================================================
// Row-block streaming form (the default): every operand of the stencil comes from shared memory, staged a PASS at a time.
//
// A pass = the 1024 sites all 256 computing threads update at once = RPP = 1024 / L0 consecutive rows of a slice (thread
// (tx, ty): row ty of the block, sites 4 tx .. 4 tx + 3).  What a pass reads is seven linear runs of global memory:
//   the RPP rows themselves + the row before + the row after (x1 -+ 1, wrapped inside the plane), and the same RPP rows of
//   the slice above, the slice below, the plane above and the plane below (t +- 1, x2 +- 1)
// -- seven cp.async.bulk copies into one STAGE of (5 RPP + 2) rows, about 21 KB.  Three stages per CTA, three CTAs per SM:
// ~190 KB of loads in flight per SM, against the 48 KB the 24 warps' own 128-bit loads could keep in flight (four per
// warp and pass) -- that was the one-tile kernel's limit: a third of its stall samples sat on the first use of those loads
// (profiles/r02_slab_tile.txt), Little's law at ~1 us of loaded latency.
//   * full[s] / empty[s] mbarriers per stage; a ninth warp produces: waits for empty[s] (count 8: lane 0 of every
//     computing warp), arms full[s] with the stage's byte count, seven lanes issue one copy each;
//   * the eight computing warps run the noise phase of a pass (2/3 of its instructions, no field data) BEFORE they wait
//     for its stage; then seven 128-bit and two 32-bit shared-memory loads, the update, one 128-bit global store;
//   * tiles (R passes = what the marching kernel calls a CTA's rows: partials, jump tables and clamp slots are shared with
//     it) are CLAIMED from a counter, in order, by the producer -- what the hardware does with CTAs -- so the tiles in
//     flight stay one contiguous window of the L2 chunk order (cta_slice_position).  The chain state is set from the
//     tile's header (written by the producer before it arms the first stage of the tile) and walked affinely from pass
//     to pass; no CTA-wide barrier after the first.
// Measured dead ends on the way (256^3 x 32 slices; one-tile kernel = 384 G site-updates/s):
//   - whole tiles in two stages, FIXED stride (tile = blockIdx.x + i gridDim.x): 209.  The CTAs drift apart over their
//     ~150 tiles, the L2 chunk order then means nothing and every site comes from HBM three times (6.6 GB read per step
//     instead of 2.6, ncu);
//   - the producer role rotating over the eight computing warps instead of a ninth warp: 219;
//   - whole tiles in two stages, claimed dynamically, t+-1 / x2+-1 still direct loads: 347 -- hiding the CTA start-up
//     buys nothing while the direct loads bound the passes;
//   - this kernel + cp.async.bulk.prefetch.L2 of the slice above / plane above 2..8 passes ahead of their copies: 308
//     against 355 without.
#ifndef ROWS_STAGES
#define ROWS_STAGES 2
#endif
constexpr unsigned ROWS_D = ROWS_STAGES;  // stages (the computing warps' pass loop is unrolled by it: a stage's address is an immediate)
#ifndef ROWS_WAIT_NS
#define ROWS_WAIT_NS 0      // > 0: waits park the warp for up to this many ns per try (suspend-time hint)
#endif
#if ROWS_WAIT_NS > 0
#define ROWS_WAIT(bar, par) tile_mbar_wait_hint((bar), (par), ROWS_WAIT_NS)
#else
#define ROWS_WAIT(bar, par) tile_mbar_wait((bar), (par))
#endif
#ifndef ROWS_ARRIVE_ALL
#define ROWS_ARRIVE_ALL 0   // 1: every lane arrives on empty[s] (count 256, no __syncwarp) ; 0: lane 0 of each warp (count 8)
#endif
struct RowsShared {
    TileHdr H[3];          // by tile number mod 3 (the producer writes tile t+1's while t is computed and t-1 retired)
    double red[2][2][8];   // by tile parity
    int chain[3];
    int end[3];            // no tile: the grid has run out of them
    int skip;
};

// wait that parks the warp in hardware for up to `ns` per try instead of spinning through the issue slots
__device__ __forceinline__ void tile_mbar_wait_hint(unsigned bar, unsigned parity, unsigned ns) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity), "r"(ns) : "memory");
}
__device__ __forceinline__ void tile_mbar_arrive(unsigned bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}

template <int MATH, int NDIM, int POT, int L0T>
__global__ void __launch_bounds__(288, 3) lattice_rows_kernel(const LatticeArgs A, const unsigned ntiles, const unsigned cps, const unsigned stage_stride) {
    constexpr unsigned NP = 2, W = 4;
    constexpr unsigned NSTREAM = (NDIM >= 4) ? 5u : 3u;
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ RowsShared S;
    const unsigned L0 = L0T ? (unsigned)L0T : (unsigned)A.dim[0], L1 = (unsigned)A.dim[1];
    const unsigned L2 = (NDIM >= 4) ? (unsigned)A.dim[2] : 1u;
    const unsigned ROWB = L0 * 4u;
    const unsigned RPP = 256u >> A.m_tpr_log;   // rows per pass
    const unsigned R = (unsigned)A.m_R;         // passes per tile
    constexpr unsigned PASSB = 4096u;           // RPP * ROWB: bytes of one stream in a stage
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(tile_smem);
    // mbarriers: full[s] at smem0 + 8 s, empty[s] at smem0 + 32 + 8 s ; stage s at smem0 + 128 + s * stage_stride:
    //   [x1-1 halo row | RPP rows | x1+1 halo row | t+1 rows | t-1 rows | x2+1 rows | x2-1 rows]
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    if (threadIdx.x == 0) {
        // an earlier launch flagged an event: this one will be replayed
        S.skip = *((volatile const u64 *)A.event_key) != NO_EVENT;
        for (unsigned s = 0; s < ROWS_D; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem0 + 8u * s) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem0 + 32u + 8u * s), "r"(ROWS_ARRIVE_ALL ? 256u : 8u) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (S.skip) return;

    if (warp == 8) {  // ================= the producer warp =========================================================
        // ---- claim the next tile and write its header into slot hs (lane 0); false: none left ----
        auto claim = [&](const unsigned hs) -> bool {
            unsigned lin = 0;
            if (lane == 0) lin = atomicAdd(A.tile_ctr, 1u);
            lin = __shfl_sync(0xFFFFFFFFu, lin, 0);
            if (lin >= ntiles) return false;
            if (lane == 0) {
                TileHdr &H = S.H[hs];
                S.end[hs] = 0;
                const unsigned per_chain = cps * (unsigned)A.nt;
                const unsigned chain = lin / per_chain;
                lin -= chain * per_chain;
                int tl = 0;
                unsigned bx = 0;
                {   // cta_slice_position for a linear tile index
                    unsigned t_first = 0, nt_sweep = (unsigned)A.nt;
                    bool done = false;
                    if (A.slab_on && A.nt > 1) {
                        if (lin < 2 * cps) {
                            tl = (lin < cps) ? 0 : A.nt - 1;
                            bx = (lin < cps) ? lin : lin - cps;
                            done = true;
                        } else {
                            lin -= 2 * cps;
                            t_first = 1;
                            nt_sweep = (unsigned)A.nt - 2;
                        }
                    }
                    if (!done) {
                        const unsigned cpc = (unsigned)A.ctas_per_chunk;
                        const unsigned per_chunk = cpc * nt_sweep;
                        const unsigned chunk = lin / per_chunk;
                        const unsigned rem = lin - chunk * per_chunk;
                        const unsigned width = (chunk * cpc + cpc <= cps) ? cpc : cps - chunk * cpc;  // the last chunk may be narrower
                        const unsigned ty = rem / width;
                        tl = (int)(t_first + ty);
                        bx = chunk * cpc + (rem - ty * width);
                    }
                }
                if (A.slab_on) {  // slab ring: the neighbour's boundary slice of this step's input must have landed
                    if (tl == 0) slab_wait(A.wait_flag[0], A.wait_tag, A.slab_error);
                    if (tl == A.nt - 1) slab_wait(A.wait_flag[1], A.wait_tag, A.slab_error);
                }
                const long long vs = A.vslice;
                const u64 gslice = (u64)(A.slab_t0 + tl) * (u64)vs;
                const unsigned o_cta = bx * (RPP * R) * L0;
                const u64 g_cta = gslice + o_cta;
                const u64 s_sl = lcg_apply(A.slice_jump[tl], A.seed_in[chain], 0) & LCG_MASK;
                H.s_cta = lcg_apply(A.cta_jump[bx], s_sl, gslice) & LCG_MASK;
                const u64 scg = LCG_BETA * g_cta + LCG_GAMMA;
                H.scg = scg;
                // T(next pass) = alpha^1024 T + ck ; ck itself is a running sum
                H.ck = scg * A.prow_jump.g0 + A.prow_jump.bg1 - A.prow_jump.a * TWO31 + TWO31;
                H.c1 = site_const(g_cta) - LCG_A * TWO31;
                H.c2 = (LCG_A + 1) * site_const(g_cta) - LCG_ALPHA * TWO31;
                H.g_cta = g_cta;
                const float *in = (const float *)A.in + (long long)chain * A.chain_stride;
                const float *cur = in + (long long)tl * vs;
                const float *tm = (tl > 0) ? cur - vs : (A.wrap_time ? in + (long long)(A.nt - 1) * vs : (const float *)A.ghost_lo);
                const float *tp = (tl < A.nt - 1) ? cur + vs : (A.wrap_time ? in : (const float *)A.ghost_hi);
                H.cur_slice = cur;
                H.cur = (const char *)cur;   // (slice bases: the producer adds the row)
                H.tp = (const char *)tp;
                H.tm = (const char *)tm;
                H.dst = (char *)((float *)A.out + (long long)chain * A.chain_stride + (long long)tl * vs + o_cta);
                H.tl = tl;
                H.bx = bx;
                H.x2 = 0;
                H.o_cta = o_cta;
                S.chain[hs] = (int)chain;
            }
            __syncwarp();
            return true;
        };
        // ---- everybody has left the tile in header slot hs: its observable partial, and a slab ring's flag (lane 0) ----
        auto retire = [&](const unsigned hs, const unsigned rs) {
            if (lane == 0) {
                const TileHdr &H = S.H[hs];
                const int tl = H.tl, chain = S.chain[hs];
                if (A.partials) {
                    double s1 = 0, s2 = 0;
                    for (int q = 0; q < 8; ++q) { s1 += S.red[rs][0][q]; s2 += S.red[rs][1][q]; }
                    double *p = A.partials + (((long long)chain * A.nt + tl) * cps + H.bx) * 2;
                    p[0] = s1;
                    p[1] = s2;
                }
                const bool push_lo = A.slab_on && tl == 0 && A.push_tag, push_hi = A.slab_on && tl == A.nt - 1 && A.push_tag;
                if (push_lo || push_hi) {  // last tile of the slice: everything is out, raise the neighbour's flag
                    __threadfence_system();
                    if (push_lo && atomicAdd(A.push_count + 0, 1u) == cps - 1) {
                        A.push_count[0] = 0;
                        __threadfence_system();
                        st_release_sys_u32(A.push_flag[0], A.push_tag);
                    }
                    if (push_hi && atomicAdd(A.push_count + 1, 1u) == cps - 1) {
                        A.push_count[1] = 0;
                        __threadfence_system();
                        st_release_sys_u32(A.push_flag[1], A.push_tag);
                    }
                }
            }
            __syncwarp();
        };
        const unsigned stage_bytes = (NSTREAM * RPP + 2u) * ROWB;
        unsigned stage = 0, phase = 0, g = 0, t = 0;  // stage / phase of pass g ; t: tiles started
        unsigned hs = 0;                              // header slot of tile t (t mod 3)
        bool more = claim(0);
        if (!more && lane == 0) {
            S.end[0] = 1;
            tile_mbar_arrive(smem0);
        }
#pragma unroll 1
        while (more) {
            const TileHdr &H = S.H[hs];
            const unsigned hs_next = hs == 2 ? 0u : hs + 1u, hs_prev = hs == 0 ? 2u : hs - 1u;
            // lane roles: 0 rows, 1 row before, 2 row after, 3 slice above, 4 slice below, 5 plane above, 6 plane below
            const char *base = (lane == 3) ? H.tp : (lane == 4 ? H.tm : H.cur);
            const unsigned row0 = H.bx * (RPP * R);
            unsigned x1a = row0 % L1, x2 = (NDIM >= 4) ? row0 / L1 : 0u;
            const unsigned dst_row = lane == 0 ? 1u : (lane == 1 ? 0u : (lane == 2 ? RPP + 1u : (lane - 2u) * RPP + 2u));
            const unsigned bytes = (lane == 1 || lane == 2) ? ROWB : PASSB;
            const bool active = lane < NSTREAM + 2u;
            bool next = false;
#pragma unroll 1
            for (unsigned p = 0; p < R; ++p) {
                if (g >= ROWS_D) {
                    ROWS_WAIT(smem0 + 32u + 8u * stage, phase ^ 1u);  // everybody has left pass g - D
                    // ... if that was the first pass of this tile, every warp has written its share of the tile before's partial
                    if (p == ROWS_D && t >= 1) retire(hs_prev, (t - 1u) & 1u);
                }
                long long off = 0;  // rows, relative to the block's first row
                if (lane == 1) off = (x1a == 0) ? (long long)L1 - 1 : -1;
                if (lane == 2) off = (x1a + RPP == L1) ? (long long)RPP - (long long)L1 : (long long)RPP;
                if (NDIM >= 4) {
                    if (lane == 5) off = (x2 + 1 == L2) ? -(long long)(L2 - 1) * L1 : (long long)L1;
                    if (lane == 6) off = (x2 == 0) ? (long long)(L2 - 1) * L1 : -(long long)L1;
                }
                const unsigned bar = smem0 + 8u * stage;
                if (lane == 0) tile_expect_tx(bar, stage_bytes);  // (release: a new tile's header is visible to whoever sees the phase end)
                __syncwarp();
                if (active) bulk_g2s(smem0 + 128u + stage * stage_stride + dst_row * ROWB, base + ((long long)(row0 + p * RPP) + off) * (long long)ROWB, bytes, bar);
                x1a += RPP;
                if (x1a == L1) {
                    x1a = 0;
                    ++x2;
                }
                ++g;
                if (++stage == ROWS_D) {
                    stage = 0;
                    phase ^= 1u;
                }
                // the NEXT tile is claimed and its header worked out while this one streams (the counter's round trip and the
                // header arithmetic would otherwise stall the stages at every tile edge: 9 % of the stall samples)
                if (p == 0) next = claim(hs_next);
            }
            ++t;
            hs = hs_next;
            more = next;
            if (!more) {  // tell the computing warps: the next pass's stage carries `end` instead of data
                if (g >= ROWS_D) tile_mbar_wait(smem0 + 32u + 8u * stage, phase ^ 1u);
                if (lane == 0) {
                    S.end[hs] = 1;
                    tile_mbar_arrive(smem0 + 8u * stage);
                }
            }
        }
        if (t >= 1) {  // the last tile's partial and flag: the computing warps arrive once more when they see `end`
            tile_mbar_wait(smem0 + 32u + 8u * stage, phase);
            retire(hs == 0 ? 2u : hs - 1u, (t - 1u) & 1u);
        }
        // the last CTA out rearms the counters for the next launch on the stream
        if (lane == 0 && atomicAdd(A.tile_ctr + 1, 1u) == gridDim.x - 1) {
            A.tile_ctr[0] = 0;
            A.tile_ctr[1] = 0;
        }
        return;
    }

    // ================================ the eight computing warps ===================================================
    unsigned nclamp = 0;
    const unsigned aDl = (unsigned)A.prow_jump.a, aDh = (unsigned)(A.prow_jump.a >> 32);
    const float c_lap = (float)A.c_lap, c_dt = (float)A.c_dt;
    const pair_t K_m2d = pk(-(float)(2 * NDIM), -(float)(2 * NDIM)), K_clap = pk(c_lap, c_lap);
    const pair_t K_m2cdt = pk(-2.0f * c_dt, -2.0f * c_dt), K_mcdt = pk(-c_dt, -c_dt), K_m1 = pk(-1.0f, -1.0f);
    const pair_t K_2m32 = pk(2.3283064365386963e-10f, 2.3283064365386963e-10f), K_k2 = pk(A.k2_f, A.k2_f);
    const float kth = (float)(2.0 * 3.1415 / 4294967296.0);  // theta - pi = 2*3.1415 v2 - pi, v2 = (float)u2 * 2^-32
    const pair_t K_th = pk(kth, kth), K_mpi = pk(-3.14159265358979f, -3.14159265358979f);
    (void)K_m2cdt; (void)K_mcdt; (void)K_2m32; (void)K_k2; (void)K_th; (void)K_mpi;
    // the thread's place inside a pass: one table entry, for the whole kernel
    const TileThread &E = A.rows_thr[threadIdx.x];
    const unsigned thr_off = E.thr_off;                       // ty L0 + 4 tx
    const unsigned o_c = 128u + E.s_c, o_left = 128u + E.s_left, o_right = 128u + E.s_right;  // byte offsets inside a stage (+ the barrier line)
    const u64 e_a = E.a, e_g0 = E.g0, e_bg1 = E.bg1, e_ck = E.ck_off;
    unsigned phase = 0, hs = 0;  // (R is a multiple of the stage count: every tile starts in stage 0)
    auto leave = [&](const unsigned st) {  // the thread / warp has read what it needs from stage st
        if (ROWS_ARRIVE_ALL) {
            tile_mbar_arrive(smem0 + 32u + 8u * st);
        } else {
            __syncwarp();
            if (lane == 0) tile_mbar_arrive(smem0 + 32u + 8u * st);
        }
    };

#pragma unroll 1
    for (unsigned t = 0;; ++t) {
        ROWS_WAIT(smem0, phase);  // the tile's first stage: header written (and the rows landed)
        if (S.end[hs]) {  // (one more arrival: the producer then knows the last tile's partial is complete)
            leave(0);
            break;
        }
        const TileHdr &H = S.H[hs];
        const int chain = S.chain[hs];
        const int tl = H.tl;
        char *p_dst = H.dst + (u64)thr_off * 4u;
        const bool push_lo = A.slab_on && tl == 0 && A.push_tag, push_hi = A.slab_on && tl == A.nt - 1 && A.push_tag;
        unsigned Tl, Th;
        {
            const u64 s = (e_a * H.s_cta + H.scg * e_g0 + e_bg1) & LCG_MASK;  // the table jump over thr_off draws
            const u64 T = s + TWO31;
            Tl = (unsigned)T;
            Th = (unsigned)(T >> 32);
        }
        u64 ck = H.ck + e_ck;
        u64 c1 = H.c1 + LCG_A * thr_off;
        u64 c2 = H.c2 + LCG_BETA * thr_off;
        const float m2 = (float)(A.m2_chain ? A.m2_chain[chain] : A.m2);
        const float lam = (float)(A.lam_chain ? A.lam_chain[chain] : A.lam);
        const pair_t K_lam = pk(lam, lam), K_m2 = pk(m2, m2);
        (void)K_lam; (void)K_m2;
        pair_t ACC1 = 0, ACC2 = 0;

        auto pass = [&](const unsigned p, const unsigned stage) {
            // ---- noise phase: W draws in t2 form, Box-Muller (no field data: runs ahead of the stage) ---------------
            const unsigned T0l = Tl, T0h = Th;
            unsigned tl_ = Tl, th_ = Th, um = 0xFFFFFFFFu;
            pair_t NZ[NP];
#pragma unroll
            for (unsigned q = 0; q < NP; ++q) {
                unsigned u1a, u2a, u1b, u2b, al, ah, bl, bh;
                mad48t(tl_, th_, A_LO, A_HI, c1, al, ah);
                mad48t(tl_, th_, ALPHA_LO32T, ALPHA_HI32T, c2, bl, bh);
                u1a = __funnelshift_r(al, ah, 16);
                u2a = __funnelshift_r(bl, bh, 16);
                c1 += LCG_A;
                c2 += LCG_BETA;
                mad48t(bl, bh, A_LO, A_HI, c1, al, ah);
                mad48t(bl, bh, ALPHA_LO32T, ALPHA_HI32T, c2, tl_, th_);
                u1b = __funnelshift_r(al, ah, 16);
                u2b = __funnelshift_r(tl_, th_, 16);
                if (q + 1 < NP) {
                    c1 += LCG_A;
                    c2 += LCG_BETA;
                } else {  // on to the strip's place in the next pass
                    c1 += A.p_dc1;
                    c2 += A.p_dc2;
                }
                um = min(min(um, u1a), u2a);  // u1 == 0 (retry) or u2 < 2^15 (`seed+=`) => um < 2^15
                um = min(min(um, u1b), u2b);
                if (MATH == 1) {
                    float l1a, l1b, ta, tb, tha, thb;
                    upk(mul2(pk(__uint2float_rn(u1a), __uint2float_rn(u1b)), K_2m32), l1a, l1b);
                    upk(mul2(pk(lg2_approx(l1a), lg2_approx(l1b)), K_k2), ta, tb);
                    upk(fma2(pk(__uint2float_rn(u2a), __uint2float_rn(u2b)), K_th, K_mpi), tha, thb);
                    NZ[q] = mul2(pk(__cosf(tha), __cosf(thb)), pk(sqrt_approx(fabsf(ta)), sqrt_approx(fabsf(tb))));
                } else {
                    const float da = (float)__dmul_rn(A.nscale, noise_accurate((u64)u1a << 16, (u64)u2a << 16));
                    const float db = (float)__dmul_rn(A.nscale, noise_accurate((u64)u1b << 16, (u64)u2b << 16));
                    NZ[q] = pk(-da, -db);
                }
            }
            {   // next pass: same place in the block, 1024 draws further
                mad48t(T0l, T0h, aDl, aDh, ck, Tl, Th);
                ck += A.p_dck;
            }
            // ---- stencil phase: everything from the stage ------------------------------------------------------------
            ROWS_WAIT(smem0 + 8u * stage, phase);
            const unsigned sb = smem0 + stage * stage_stride;
            pair_t C[NP], U1[NP], D1[NP];
            ulonglong2 U2, D2, TP, TM;
            {
                const ulonglong2 c = lds128(sb + o_c), u = lds128(sb + o_c + ROWB), d = lds128(sb + o_c - ROWB);
                C[0] = c.x; C[1] = c.y;
                U1[0] = u.x; U1[1] = u.y;
                D1[0] = d.x; D1[1] = d.y;
            }
            const float left = lds32(sb + o_left), right = lds32(sb + o_right);
            TP = lds128(sb + o_c + ROWB + PASSB);
            TM = lds128(sb + o_c + ROWB + 2u * PASSB);
            if (NDIM >= 4) {
                U2 = lds128(sb + o_c + ROWB + 3u * PASSB);
                D2 = lds128(sb + o_c + ROWB + 4u * PASSB);
            }
            float pp[W];
#pragma unroll
            for (unsigned q = 0; q < NP; ++q) upk(C[q], pp[2 * q], pp[2 * q + 1]);
            pair_t V[NP];
            float amax = 0.f;
#pragma unroll
            for (unsigned q = 0; q < NP; ++q) {
                const float xm0 = (q == 0) ? left : pp[(2 * q + W - 1) % W], xp1 = (q == NP - 1) ? right : pp[(2 * q + 2) % W];
                pair_t Sq = pk(__fadd_rn(pp[2 * q + 1], xm0), __fadd_rn(xp1, pp[2 * q]));  // phi(+0) + phi(-0)
                Sq = add2(Sq, U1[q]);
                Sq = add2(Sq, D1[q]);
                if (NDIM >= 4) {
                    Sq = add2(Sq, (q & 1) ? U2.y : U2.x);
                    Sq = add2(Sq, (q & 1) ? D2.y : D2.x);
                }
                Sq = add2(Sq, (q & 1) ? TP.y : TP.x);
                Sq = add2(Sq, (q & 1) ? TM.y : TM.x);
                pair_t v = fma2(K_clap, fma2(K_m2d, C[q], Sq), C[q]);
                if (POT == 4) v = fma2(K_mcdt, mul2(C[q], fma2(K_lam, mul2(C[q], C[q]), K_m2)), v);
                else v = fma2(K_m2cdt, C[q], v);  // (-c_dt)(2 phi) == (-2 c_dt) phi exactly
                v = fma2(K_m1, NZ[q], v);          // v + dw, one rounding
                V[q] = v;
                float a0, a1;
                upk(v, a0, a1);
                amax = fmaxf(fmaxf(fabsf(a0), fabsf(a1)), amax);
                ACC1 = add2(ACC1, C[q]);           // observables of the pre-update field
                ACC2 = fma2(C[q], C[q], ACC2);
            }
            leave(stage);
            // clamp (tau_kernel.cl:122-132) and RNG events: one test per strip for both rare cases
            if (__builtin_expect(!(amax < 1000.0f) | (um < 32768u), 0)) {
                bool replayed = false;  // an event in this strip: the launch is redone, its clamp hits are not counted
                if (um < 32768u) {
                    const u64 z0 = ((((u64)T0h << 32) | T0l) - TWO31) & LCG_MASK;
                    replayed = strip_events_cold(A.event_key, A.step_index, chain, z0, H.g_cta + thr_off + p * 1024u, (int)W);
                }
                float v0, v1, v2, v3;
                upk(V[0], v0, v1);
                upk(V[1], v2, v3);
                const Clamped cl = clamp_cold(v0, v1, v2, v3);
                V[0] = pk(cl.v[0], cl.v[1]);
                V[1] = pk(cl.v[2], cl.v[3]);
                if (!replayed) nclamp += cl.n;
            }
            *reinterpret_cast<ulonglong2 *>(p_dst) = make_ulonglong2(V[0], V[1]);
            if (__builtin_expect(push_lo | push_hi, 0)) {  // CTA-uniform: boundary slices of a slab ring only
                const size_t oo = (size_t)(H.o_cta + thr_off + p * 1024u) * 4u;
                if (push_lo) *reinterpret_cast<ulonglong2 *>((char *)A.push_ghost[0] + oo) = make_ulonglong2(V[0], V[1]);
                if (push_hi) *reinterpret_cast<ulonglong2 *>((char *)A.push_ghost[1] + oo) = make_ulonglong2(V[0], V[1]);
            }
            p_dst += PASSB;
        };
#pragma unroll 1
        for (unsigned p = 0; p < R; p += ROWS_D) {
#pragma unroll
            for (unsigned u = 0; u < ROWS_D; ++u) pass(p + u, u);
            phase ^= 1u;
        }

        // ---- the omega work-item's draw (gid = V) and the step's final seed -------------------------------------
        if (threadIdx.x == 0 && H.bx == 0 && tl == 0) {
            const u64 Vg = (u64)A.V;
            u64 t1, t2;
            const u64 sv = lcg_apply(A.vol_jump, A.seed_in[chain], 0) & LCG_MASK;
            lcg_draw(sv, Vg, t1, t2);
            if (lcg_event(sv & LCG_MASK, t1, t2)) atomicMin((unsigned long long *)A.event_key, event_key(A.step_index, chain, Vg));
            A.seed_out[chain] = lcg_next_seed(t2);
        }
        // ---- the warp's share of the tile's observable partial ---------------------------------------------------------
        if (A.partials) {
            float a1l, a1h, a2l, a2h;
            upk(ACC1, a1l, a1h);
            upk(ACC2, a2l, a2h);
            const double a1 = warp_sum((double)a1l + (double)a1h), a2 = warp_sum((double)a2l + (double)a2h);
            if (lane == 0) { S.red[t & 1u][0][warp] = a1; S.red[t & 1u][1][warp] = a2; }
        }
        hs = hs == 2 ? 0u : hs + 1u;
    }
    if (nclamp) atomicAdd(A.nclamped, (unsigned long long)nclamp);
}

// shared memory a CTA needs.  Tile kernel: a line for the mbarrier + the tile as [halo | rows | halo] per run of rows
// inside a plane.  Row-block kernel: the barrier line + its stages of (5 RPP + 2) rows (3 RPP + 2 in three dimensions).
size_t tile_smem_bytes(int ndim, int L0, int L1, int tpr_log, int R, bool rows) {
    const unsigned rpp = 256u >> tpr_log, rows_per_cta = rpp * (unsigned)R;
    if (rows) {
        const size_t stage = ((size_t)((ndim >= 4 ? 5u : 3u) * rpp + 2u) * (size_t)L0 * 4u + 127) / 128 * 128;
        return 128 + ROWS_D * stage;
    }
    const unsigned seg_rows = rows_per_cta < (unsigned)L1 ? rows_per_cta : (unsigned)L1, nseg = rows_per_cta / seg_rows;
    return 128 + (size_t)nseg * (seg_rows + 2u) * (size_t)L0 * 4u;
}
// the tile must be a whole number of planes or divide one (the rows of a thread never straddle a plane edge); tpr_log = log2
// of the threads per row (row length / 4).  Row-block kernel: a pass (1024 / L0 rows) must divide a plane, and the passes of a
// tile a multiple of the stage count.  Three CTAs per SM: 72 KB each.
bool tile_shape_ok(int ndim, int L0, int L1, int tpr_log, int R, bool rows) {
    const unsigned rpp = 256u >> tpr_log, rows_per_cta = rpp * (unsigned)R;
    if (L1 % R != 0 || R % 4 != 0) return false;
    if (!(rows_per_cta % (unsigned)L1 == 0 || (unsigned)L1 % rows_per_cta == 0)) return false;
    if (rows && ((unsigned)L1 % rpp != 0 || R % (int)ROWS_D != 0)) return false;
    return tile_smem_bytes(ndim, L0, L1, tpr_log, R, rows) <= 72 * 1024;
}

// Instantiated: 4-site strips, event-free steps.  (8-site strips -- 124 registers, two CTAs per SM -- were measured at
// 328 G site-updates/s on 256^3 slices against 368 for 4-site strips; steps with replay entries go to the marching kernel,
// which shares this kernel's tiles and jump tables.)
template <int MATH, int NDIM, int POT, int L0T>
static cudaError_t tile_go(const LatticeArgs &A, dim3 grid, size_t smem, cudaStream_t st) {
    if (A.m_on == 3) {
        // CTAs the device holds at once, per device ordinal and shared-memory size (asked once: the answer does not change;
        // two host threads asking at the same time write the same numbers)
        static unsigned slots_of[64] = {0};
        static size_t smem_of[64] = {0};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
        auto kern = lattice_rows_kernel<MATH, NDIM, POT, L0T>;
        if (slots_of[dev] == 0 || smem_of[dev] != smem) {
            int sms = 0, per_sm = 0;
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024)) != cudaSuccess) return e;
            if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 288, smem)) != cudaSuccess) return e;
            if (per_sm < 1) return cudaErrorLaunchOutOfResources;
            smem_of[dev] = smem;
            slots_of[dev] = (unsigned)(sms * per_sm);
        }
        const unsigned long long nt64 = (unsigned long long)grid.x * grid.y * grid.z;
        if (nt64 >= (1ull << 31) || !A.tile_ctr || !A.rows_thr || A.m_R < 4 || A.m_R % (int)ROWS_D != 0) return cudaErrorInvalidValue;
        const unsigned ntiles = (unsigned)nt64, slots = slots_of[dev];
        const unsigned nb = ntiles < slots ? ntiles : slots;
        if (getenv("SQ_DEBUG")) fprintf(stderr, "rows: slots %u ntiles %u grid %u smem %zu\n", slots, ntiles, nb, smem);
        kern<<<nb, 288, smem, st>>>(A, ntiles, grid.x, (unsigned)((smem - 128) / ROWS_D));
        return cudaGetLastError();
    }
    if (smem > 48 * 1024) {  // (idempotent; the attribute is per function)
        cudaError_t e = cudaFuncSetAttribute(lattice_tile_kernel<MATH, NDIM, POT, L0T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
        if (e != cudaSuccess) return e;
    }
    // Programmatic dependent launch pays only where nothing sits between two steps' update kernels in the stream.  With
    // observables on, the events that order the finalize stream used to be recorded there (64^4: 56.9 -> 56.7 us per step
    // with observables, 54.3 -> 50.7 with SQ_FLAG_NO_OBSERVABLES); sq_enqueue_step now hands the finalizes over in groups
    // of fin_batch steps, so the update kernels inside a group are neighbours in the stream.  Everything this kernel reads
    // or writes that another launch touches (field, seed, event word, partials) sits behind its griddepcontrol.wait.
    static const bool pdl_env = !(getenv("SQ_PDL") && atoi(getenv("SQ_PDL")) == 0);  // A/B knob
    const bool pdl = pdl_env && !A.slab_on;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, lattice_tile_kernel<MATH, NDIM, POT, L0T, 2>, A);
}
template <int MATH, int NDIM, int POT>
static cudaError_t tile_l0(const LatticeArgs &A, dim3 grid, size_t smem, cudaStream_t st) {
    switch (A.dim[0]) {
        case 32: return tile_go<MATH, NDIM, POT, 32>(A, grid, smem, st);
        case 64: return tile_go<MATH, NDIM, POT, 64>(A, grid, smem, st);
        case 256: return tile_go<MATH, NDIM, POT, 256>(A, grid, smem, st);
    }
    return tile_go<MATH, NDIM, POT, 0>(A, grid, smem, st);
}
template <int MATH, int NDIM>
static cudaError_t tile_pot(const LatticeArgs &A, dim3 grid, size_t smem, cudaStream_t st) {
    return A.pot == 4 ? tile_l0<MATH, NDIM, 4>(A, grid, smem, st) : tile_l0<MATH, NDIM, 0>(A, grid, smem, st);
}

cudaError_t launch_lattice_tile(const LatticeArgs &A, int math, int ctas_per_slice, cudaStream_t stream) {
    dim3 grid((unsigned)ctas_per_slice, (unsigned)A.nt, (unsigned)A.nchains);
    const size_t smem = tile_smem_bytes(A.ndim, (int)A.dim[0], (int)A.dim[1], A.m_tpr_log, A.m_R, A.m_on == 3);
    if (A.m_w != 4 || A.n_rebase != 0) return cudaErrorInvalidValue;
    if (A.ndim == 3) return math ? tile_pot<1, 3>(A, grid, smem, stream) : tile_pot<0, 3>(A, grid, smem, stream);
    if (A.ndim == 4) return math ? tile_pot<1, 4>(A, grid, smem, stream) : tile_pot<0, 4>(A, grid, smem, stream);
    return cudaErrorInvalidValue;
}

}  // namespace sq
