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

namespace sq {

namespace {

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_cg_f4(const float *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// cold path: exact event test of the 4 draws of a strip (literal replay is the host's job)
__device__ __noinline__ void strip_events_cold(u64 *event_key_ptr, int step, u64 sm, u64 g0) {
    for (int e = 0; e < 4; ++e) {
        u64 t1, t2;
        lcg_draw(sm, g0 + e, t1, t2);
        if (lcg_event(sm, t1, t2)) atomicMin((unsigned long long *)event_key_ptr, event_key(step, 0, g0 + e));
        sm = lcg_next_seed(t2) & LCG_MASK;
    }
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
template <int MATH>
__device__ __forceinline__ float site_update(float phi, float nsum, unsigned u1, unsigned u2, const SiteCoef &C) {
    const float lap = __fmaf_rn(-4.0f, phi, nsum);
    float v = __fmaf_rn(C.c_lap, lap, phi);
    if (C.pot == 4) {
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

// NR = rows actually owned by this CTA (compile-time so the band stays in registers)
template <int NR, int MATH>
__device__ __forceinline__ void resident_run(const ResidentArgs &A, int r0, float *smem) {
    const int T = blockDim.x, t = threadIdx.x, b = blockIdx.x, nb = gridDim.x;
    const int L0 = A.L0;
    const int bup = (b + 1 == nb) ? 0 : b + 1, bdn = (b == 0) ? nb - 1 : b - 1;
    // shared: edges [2 buffers][NR rows][2 (first,last)][T] ; row-sum transpose [NR+1][T]
    float *edge = smem;
    float *rs = smem + 2 * NR * 2 * T;
    SiteCoef C;
    C.c_lap = (float)A.c_lap;
    C.pot = A.pot;
    C.c_dt = (float)A.c_dt;
    C.c_2dt = 2.0f * C.c_dt;
    C.m2 = (float)A.m2;
    C.lam = (float)A.lam;
    C.nscale_d = A.nscale;
    C.k2 = (float)(2.0 * 0.6931471805599453 * A.nscale * A.nscale);

    // ---- load the band, set up the per-strip chain state -----------------------------------
    float phi[NR][4];
    Seed32 sd[NR];
    unsigned Kl[NR], Kh[NR];
    const u64 S0 = A.seed_in[0];
    const u64 S1 = (A.P * S0 + A.Q) & LCG_MASK;  // predicted seed after one whole step
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const float4 v = *reinterpret_cast<const float4 *>(A.in + (size_t)(r0 + k) * L0 + 4 * t);
        phi[k][0] = v.x; phi[k][1] = v.y; phi[k][2] = v.z; phi[k][3] = v.w;
        const u64 g = (u64)(r0 + k) * L0 + 4 * t;
        const u64 s0 = lcg_seed_at(S0, 0, g, A.jump);
        const u64 s1 = lcg_seed_at(S1, 0, g, A.jump);
        const u64 K = (s1 - A.P * s0) & LCG_MASK;
        sd[k] = seed_split(s0);
        Kl[k] = (unsigned)K;
        Kh[k] = (unsigned)(K >> 32);
    }
    const unsigned Pl = (unsigned)A.P, Ph = (unsigned)(A.P >> 32);
    // c = gid*A + B of the first site of row 0's strip; rows advance it by L0*A
    unsigned c0l, c0h;
    site_const((u64)r0 * L0 + 4 * t, c0l, c0h);
    const u64 rowA = (u64)L0 * LCG_A;

    // halo rows of the current field: from the neighbours' input rows
    float hup[4], hdn[4];
    {
        const int rup = (r0 + NR == A.L1) ? 0 : r0 + NR, rdn = (r0 == 0) ? A.L1 - 1 : r0 - 1;
        const float4 u = *reinterpret_cast<const float4 *>(A.in + (size_t)rup * L0 + 4 * t);
        const float4 d = *reinterpret_cast<const float4 *>(A.in + (size_t)rdn * L0 + 4 * t);
        hup[0] = u.x; hup[1] = u.y; hup[2] = u.z; hup[3] = u.w;
        hdn[0] = d.x; hdn[1] = d.y; hdn[2] = d.z; hdn[3] = d.w;
    }
    // edges of the initial field
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        edge[((0 * NR + k) * 2 + 0) * T + t] = phi[k][0];
        edge[((0 * NR + k) * 2 + 1) * T + t] = phi[k][3];
    }
    __syncthreads();

    const int tl = (t == 0) ? T - 1 : t - 1, tr = (t + 1 == T) ? 0 : t + 1;
    Seed32 Som = seed_split(S0);  // only used by (b==0,t==0): the step-start seed
    unsigned myclamp = 0;
    unsigned failed = 0;

    for (int n = 0; n < A.nsteps; ++n) {
        const int eb = n & 1;
        float psum[NR];   // per-thread sums of the pre-update row values
        float p2 = 0.f;
        float nw[NR][4];  // new values
        // ---- one row: draws + update -----------------------------------------------------
        auto do_row = [&](int k, const float *up, const float *dn) {
            const float left = edge[((eb * NR + k) * 2 + 1) * T + tl];
            const float right = edge[((eb * NR + k) * 2 + 0) * T + tr];
            unsigned cl, ch;
            {
                const u64 c = (((u64)c0h << 32) | c0l) + (u64)k * rowA;
                cl = (unsigned)c;
                ch = (unsigned)(c >> 32);
            }
            Seed32 s = sd[k];
            const Seed32 s_before = s;
            float a = 0.f;
            bool maybe = false;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                unsigned u1, u2;
                site_draw(s, cl, ch, u1, u2);
                site_const_next(cl, ch);
                maybe |= site_maybe_event(u1, u2);
                const float p = phi[k][e];
                const float xp = (e < 3) ? phi[k][(e + 1) & 3] : right;
                const float xm = (e > 0) ? phi[k][(e + 3) & 3] : left;
                const float nsum = __fadd_rn(__fadd_rn(__fadd_rn(xp, xm), up[e]), dn[e]);
                nw[k][e] = site_update<MATH>(p, nsum, u1, u2, C);
                a = __fadd_rn(a, p);
                p2 = __fmaf_rn(p, p, p2);
                if (fabsf(nw[k][e]) >= 1000.0f) ++myclamp;
            }
            psum[k] = a;
            if (__builtin_expect(maybe, 0))
                strip_events_cold(A.event_key, A.step_index0 + n, seed_join(s_before), (u64)(r0 + k) * L0 + 4 * t);
            // the strip keeps its gids: next step's seed by one affine map (5 integer ops)
            unsigned nl, nh;
            {
                const u64 p = (u64)sd[k].lo * Pl + (((u64)Kh[k] << 32) | Kl[k]);
                nl = (unsigned)p;
                nh = (unsigned)(p >> 32) + sd[k].lo * Ph + sd[k].hi * Pl;
            }
            sd[k].lo = nl;
            sd[k].hi = nh;
        };

        // ---- phase A: the two boundary rows first, publish them --------------------------
        if (NR == 1) {
            do_row(0, hup, hdn);
        } else {
            do_row(0, phi[1], hdn);
            do_row(NR - 1, hup, phi[NR - 2]);
        }
        {
            float *ho = A.halo + ((size_t)((n + 1) & 1) * nb + b) * 2 * L0;
            *reinterpret_cast<float4 *>(ho + 4 * t) = make_float4(nw[0][0], nw[0][1], nw[0][2], nw[0][3]);
            *reinterpret_cast<float4 *>(ho + L0 + 4 * t) =
                make_float4(nw[NR - 1][0], nw[NR - 1][1], nw[NR - 1][2], nw[NR - 1][3]);
            __threadfence();
        }
        __syncthreads();
        if (t == 0) st_release_u32(A.flags + b, A.step0 + (unsigned)n + 1u);

        // ---- phase B: interior rows; prefetch the neighbours' new boundary rows -----------
        float nup[4], ndn[4];
        bool fetched = false;
        auto fetch_halo = [&]() {
            const unsigned want = A.step0 + (unsigned)n + 1u;
            unsigned spins = 0;
            while ((int)(ld_volatile_u32(A.flags + bup) - want) < 0 || (int)(ld_volatile_u32(A.flags + bdn) - want) < 0) {
                ++spins;
                // the launch is being abandoned (an RNG event must be replayed): stop waiting
                if ((spins & 63u) == 0 && *((volatile const u64 *)A.event_key) != NO_EVENT) break;
                if (spins > (1u << 21)) { failed = 1; break; }  // ~1 s: never hang the GPU
            }
            const float *hb = A.halo + (size_t)((n + 1) & 1) * nb * 2 * L0;
            const float4 u = ld_cg_f4(hb + (size_t)bup * 2 * L0 + 4 * t);        // neighbour above: its FIRST row
            const float4 d = ld_cg_f4(hb + (size_t)bdn * 2 * L0 + L0 + 4 * t);   // neighbour below: its LAST row
            nup[0] = u.x; nup[1] = u.y; nup[2] = u.z; nup[3] = u.w;
            ndn[0] = d.x; ndn[1] = d.y; ndn[2] = d.z; ndn[3] = d.w;
            fetched = true;
        };
#pragma unroll
        for (int k = 1; k < NR - 1; ++k) {
            do_row(k, phi[k + 1], phi[k - 1]);
            if (k == (NR - 1) / 2) fetch_halo();
        }
        if (!fetched) fetch_halo();

        // ---- hand-over: registers, edges, row sums ----------------------------------------
        const int nbuf = eb ^ 1;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) phi[k][e] = nw[k][e];
            edge[((nbuf * NR + k) * 2 + 0) * T + t] = nw[k][0];
            edge[((nbuf * NR + k) * 2 + 1) * T + t] = nw[k][3];
            rs[k * T + t] = psum[k];
        }
        rs[NR * T + t] = p2;
#pragma unroll
        for (int e = 0; e < 4; ++e) { hup[e] = nup[e]; hdn[e] = ndn[e]; }
        __syncthreads();

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

        // ---- the omega work-item's draw (gid = V) ------------------------------------------
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

    // ---- write the band back ------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < NR; ++k)
        *reinterpret_cast<float4 *>(A.out + (size_t)(r0 + k) * L0 + 4 * t) =
            make_float4(phi[k][0], phi[k][1], phi[k][2], phi[k][3]);
    if (myclamp) atomicAdd(A.nclamped, (unsigned long long)myclamp);
    if (failed) atomicExch(A.error_flag, 1u);
}

template <int ROWS, int MATH>
__global__ void __launch_bounds__(256, 1) resident2d_kernel(const ResidentArgs A) {
    extern __shared__ float smem_f[];
    // an earlier launch flagged an event: this one will be replayed.  (If the flag rises while
    // the grid is still starting, late CTAs leave here and their neighbours' waits give up on it.)
    if (*((volatile const u64 *)A.event_key) != NO_EVENT) return;
    const int b = blockIdx.x, nb = gridDim.x;
    const int r0 = (int)(((long long)b * A.L1) / nb), r1 = (int)(((long long)(b + 1) * A.L1) / nb);
    if (r1 - r0 == ROWS) resident_run<ROWS, MATH>(A, r0, smem_f);
    else resident_run<(ROWS > 1 ? ROWS - 1 : 1), MATH>(A, r0, smem_f);
}

// history -> running means (tau_kernel.cl:144-145 per time slice), one thread per slice
__global__ void __launch_bounds__(1024) welford_history_kernel(const WelfordArgs A) {
    if (*((volatile const u64 *)A.event_key) != NO_EVENT) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < A.nt) {
        double x = A.slice_x[t], xx0 = A.slice_xx0[t], last = 0;
        for (int n = 0; n < A.nsteps; ++n) {
            const double cnt = (double)(A.runs + n + 1);
            const double P = A.hist_rows[(size_t)n * A.nt + t] / (double)A.vslice;
            const double Pm = A.hist_rows[(size_t)n * A.nt + A.tmid] / (double)A.vslice;
            xx0 = xx0 + (P * Pm - xx0) / cnt;
            x = x + (P - x) / cnt;
            last = A.hist_rows[(size_t)n * A.nt + t];
        }
        A.slice_x[t] = x;
        A.slice_xx0[t] = xx0;
        A.slice_sum[t] = last;
    }
    if (t == 0) {
        double m1 = A.sums_mean[0], m2 = A.sums_mean[1], s1 = 0, s2 = 0;
        const double vol = (double)A.vslice * (double)A.nt;
        for (int n = 0; n < A.nsteps; ++n) {
            s1 = 0;
            s2 = 0;
            for (int k = 0; k < A.nt; ++k) s1 += A.hist_rows[(size_t)n * A.nt + k];
            for (int k = 0; k < A.np2; ++k) s2 += A.hist_p2[(size_t)n * A.np2 + k];
            const double cnt = (double)(A.runs + n + 1);
            m1 += (s1 / vol - m1) / cnt;
            m2 += (s2 / vol - m2) / cnt;
        }
        A.sums[0] = s1;
        A.sums[1] = s2;
        A.sums_mean[0] = m1;
        A.sums_mean[1] = m2;
    }
}

cudaError_t launch_welford_history(const WelfordArgs &A, cudaStream_t stream) {
    welford_history_kernel<<<(A.nt + 1023) / 1024, 1024, 0, stream>>>(A);
    return cudaGetLastError();
}

template <int ROWS>
static cudaError_t launch_rows(const ResidentArgs &A, int math, int nblocks, int threads, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)2 * ROWS * 2 * threads + (size_t)(ROWS + 1) * threads);
    void *args[] = {(void *)&A};
    const void *fn = math ? (const void *)resident2d_kernel<ROWS, 1> : (const void *)resident2d_kernel<ROWS, 0>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaLaunchCooperativeKernel(fn, dim3(nblocks), dim3(threads), args, smem, st);
}

// rows_max = ceil(L1 / nblocks)
cudaError_t launch_resident2d(const ResidentArgs &A, int math, int nblocks, int rows_max, cudaStream_t st) {
    const int threads = A.L0 / 4;
    switch (rows_max) {
        case 1: return launch_rows<1>(A, math, nblocks, threads, st);
        case 2: return launch_rows<2>(A, math, nblocks, threads, st);
        case 3: return launch_rows<3>(A, math, nblocks, threads, st);
        case 4: return launch_rows<4>(A, math, nblocks, threads, st);
        case 5: return launch_rows<5>(A, math, nblocks, threads, st);
        case 6: return launch_rows<6>(A, math, nblocks, threads, st);
        case 7: return launch_rows<7>(A, math, nblocks, threads, st);
        case 8: return launch_rows<8>(A, math, nblocks, threads, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sq
