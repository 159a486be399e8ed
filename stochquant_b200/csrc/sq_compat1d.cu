// sq_compat1d.cu -- the reference's 1-D Langevin kernel, re-designed for one B200 CTA.
//
// Replaces __kernel time_dev (/root/reference/tau_kernel.cl:25-175) in the canonical
// {chain RNG, Jacobi field} semantics of SURVEY.md 8(a):
//   * rows S/P/N/C: fp64 update with the reference's operand order and float casts,
//     every product/sum rounded once (explicit __dmul_rn/__dadd_rn: no FMA contraction);
//   * row R: the shared-seed LCG chain evaluated per work-item through affine
//     jump-ahead (s_i(n+1) = P s_i(n) + K_i); a step in which any draw hits an
//     inf-retry or `seed+=` event is re-drawn literally by one thread;
//   * row T: the racy lrgEl/lrgVl/stable update as an exact as-if-sequential scan
//     done by one warp (two-pass: chunk maxima, shuffle prefix, literal replay);
//   * row O: per-site Welford means of x and x*x_mid, pre-update field;
//   * row B: field/accumulators stay in shared memory and registers for all `Loops`
//     steps of a frame -- no copy-back traffic, two block barriers per step;
//   * frame commit / rollback (tauhost.c:506-554) happens in the kernel epilogue.
#include "sq_kernels.h"
#include "sq_noise.cuh"

namespace sq {

namespace {

constexpr double ETA = .8;  // tau_kernel.cl:19
constexpr double V0 = 2.;   // tau_kernel.cl:21

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
// a / b, correctly rounded, for a divisor whose correctly rounded reciprocal rb = __drcp_rn(b) is at
// hand (every division of time_dev has a loop-invariant divisor: eta, eta^2, dt2, the running-mean
// counter).  Markstein's correction: q = RN(a rb), r = a - b q (exact in one fma), q' = RN(q + r rb)
// is RN(a/b) whenever nothing overflows and b's significand is not all ones -- three dependent FMAs
// where __ddiv_rn is ~130 instructions on a GPU without a hardware fp64 divider.  Bit-identical
// results; non-finite intermediates fall back to the library division.
__device__ __noinline__ double ddiv_cold(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double ddiv_by(double a, double b, double rb) {
    const double q = __dmul_rn(a, rb);
    if (__builtin_expect(!(fabs(q) < 1e300), 0)) return ddiv_cold(a, b);  // a real branch: a select would evaluate both
    const double r = __fma_rn(-q, b, a);
    return __fma_rn(r, rb, q);
}
__device__ __forceinline__ double absol(double a) { return (a <= 0) ? -a : a; }  // :259-267

// clas(a, w, pot), tau_kernel.cl:215-226 / :184-189 / :201-205
__device__ __forceinline__ double clas(double a, double w, int pot) {
    if (pot == 3) {
        // eta * (double)tanh((float)((double)sqrt((float)(2.*V0/m))*(t-t0)/eta)); sqrtf(4.f) == 2
        const double arg = ddiv_by(dmul(2.0, dsub(a, w)), ETA, __drcp_rn(ETA));  // (constant-folded)
        return dmul(ETA, (double)tanhf((float)arg));
    }
    return 0.;
}
// ddPot(a, pot), :227-236 / :190-195 / :206-209
__device__ __forceinline__ double ddPot(double a, int pot) {
    if (pot == 3) {
        constexpr double ee = ETA * ETA;  // eta*eta, folded with one rounding like any C compiler
        const double ree = __drcp_rn(ee);  // (constant-folded)
        return ddiv_by(dsub(ddiv_by(dmul(dmul(12. * V0, a), a), ee, ree), 4. * V0), ee, ree);
    }
    return 2.;
}

struct ScanIn {
    double v, a, d;
};

}  // namespace

template <int IPT>
__global__ void __launch_bounds__(1024, 1) compat1d_frame_kernel(Compat1DArgs A) {
    extern __shared__ double smem[];
    const int N = A.N;
    const int tid = threadIdx.x;
    // shared arrays
    double *f_s = smem;            // [N]   Jacobi field (step-start values)
    double *nfp_s = f_s + N;       // [N]   the reference's persistent `newf` buffer
    double *v_s = nfp_s + N;       // [N]   newf+cl
    double *a_s = v_s + N;         // [N]   |newf+cl|
    double *d_s = a_s + N;         // [N]   |newf-f-dw|
    u64 *t_s = (u64 *)(d_s + N);   // [2*(N+1)] literal (t1,t2) of a replayed step
    __shared__ double sh_om[2];    // omega (current / next)
    __shared__ double sh_lrgVl;
    __shared__ int sh_lrgEl, sh_unstable;
    __shared__ u64 sh_seed;        // full-u64 seed at the start of the current step
    // block-parallel stability scan (IPT == 1): warp totals of the three prefix maxima, warp results
    __shared__ double sc_tot[3][32];
    __shared__ int sc_res[2][32];
    __shared__ double sh_R0;

    const int pot = A.potential;
    const int midpt = N / 2;
    // controller mode: this frame's step size and counter come from the device block; the two noise
    // scales are the host's expressions (sq_api.cu: enqueue_compat) evaluated here -- sqrtf is IEEE
    // correctly rounded on both sides, the casts are the reference's (tau_kernel.cl:105,112)
    const double dt = A.dt;
    const double dtau = A.ctl ? A.ctl->dtau : A.dtau;
    const long long runs0 = A.ctl ? A.ctl->runs : A.runs;
    const double nscale_site = A.ctl ? dmul(A.noise_c, (double)sqrtf((float)ddiv_cold(dmul(2., dtau), dt))) : A.nscale_site;
    const double nscale_omega = A.ctl ? dmul(A.noise_c, (double)sqrtf((float)dmul(2., dtau))) : A.nscale_omega;

    for (int i = tid; i < N; i += blockDim.x) {
        f_s[i] = A.f[i];
        nfp_s[i] = A.newf[i];
    }
    if (tid == 0) {
        sh_om[0] = *A.omega;
        sh_lrgVl = *A.lrgVl;
        sh_lrgEl = *A.lrgEl;
        sh_unstable = 0;
        sh_seed = *A.seed;
    }
    // per-item registers: items [tid*IPT, tid*IPT+IPT)
    double x_r[IPT], xx0_r[IPT];
    u64 s_r[IPT], K_r[IPT];
    __syncthreads();
    {
        const u64 S = sh_seed;
        const u64 Snext = (A.P * S + A.Q) & LCG_MASK;  // predicted seed after N+1 draws
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int i = tid * IPT + k;
            if (i < N) {
                x_r[k] = A.x[i];
                xx0_r[k] = A.xx0[i];
            } else {
                x_r[k] = xx0_r[k] = 0.;
            }
            if (i <= N) {
                s_r[k] = lcg_seed_at(S, 0, (u64)i, A.jump);
                const u64 s1 = lcg_seed_at(Snext, 0, (u64)i, A.jump);
                K_r[k] = (s1 - A.P * s_r[k]) & LCG_MASK;
            } else {
                s_r[k] = K_r[k] = 0;
            }
        }
    }
    const double r_dt2 = __drcp_rn(A.dt2);
    int cur = 0;  // sh_om index holding the current omega
    int j = 0;
    int unstable = 0;
    // IPT == 1: every thread derives the scan's results itself and carries them in registers
    int E0c = sh_lrgEl;
    double Vlc = sh_lrgVl;
    for (; j < A.loops; ++j) {
        const double om = sh_om[cur];
        const int E0 = (IPT == 1) ? E0c : sh_lrgEl;
        const double Vl0 = (IPT == 1) ? Vlc : sh_lrgVl;
        const double stale = nfp_s[E0];
        const double n_inv_den = (double)(runs0 + j + 1);  // (double)(*runs+j+1), :144
        const double r_n = __drcp_rn(n_inv_den);
        const double fmid = f_s[midpt];
        const double clmid = clas(dmul((double)midpt, dt), om, pot);
        // the scan's reference value newf[lrgEl]+cl (:135) is known now: one thread prepares it while
        // the others draw (visible after the barrier of phase 1)
        if (IPT == 1 && tid == (int)blockDim.x - 1) sh_R0 = dadd(stale, clas(dmul((double)E0, dt), om, pot));

        // ---- phase 1: draws (speculative affine chain) -------------------------------
        u64 t1_r[IPT], t2_r[IPT];
        int ev = 0;
        double my_v = -INFINITY, my_a = -INFINITY, my_d = 0.;  // IPT == 1: this thread's site for the scan
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int i = tid * IPT + k;
            if (i <= N) {
                lcg_draw(s_r[k], (u64)i, t1_r[k], t2_r[k]);
                ev |= lcg_event(s_r[k] & LCG_MASK, t1_r[k], t2_r[k]) ? 1 : 0;
            } else {
                t1_r[k] = t2_r[k] = 0;
            }
        }
        const int any_ev = __syncthreads_or(ev);
        if (any_ev) {
            // literal sequential replay of this step's N+1 draws (rare: p ~ (N+1) 2^-32 per step)
            if (tid == 0) {
                u64 sd = sh_seed;
                for (int i = 0; i <= N; ++i) {
                    u64 t1, t2;
                    do {  // tau_kernel.cl:272-282
                        lcg_draw(sd, (u64)i, t1, t2);
                        if (sd < TWO31 && t2 < TWO31) sd += t2;
                        else sd = t2 - TWO31;
                    } while ((t1 >> 16) == 0);
                    t_s[2 * i] = t1;
                    t_s[2 * i + 1] = t2;
                }
                sh_seed = sd;
                atomicAdd(A.nevents, 1ULL);
            }
            __syncthreads();
            const u64 S = sh_seed;
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const int i = tid * IPT + k;
                if (i <= N) {
                    t1_r[k] = t_s[2 * i];
                    t2_r[k] = t_s[2 * i + 1];
                    s_r[k] = lcg_seed_at(S, 0, (u64)i, A.jump);  // seeds for the NEXT step
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const int i = tid * IPT + k;
                if (i == N) sh_seed = lcg_next_seed(t2_r[k]);  // full-u64 value of :281
                s_r[k] = (A.P * s_r[k] + K_r[k]) & LCG_MASK;
            }
        }

        // ---- phase 2: update, observables, scan inputs ---------------------------------
        double nf_r[IPT];
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int i = tid * IPT + k;
            nf_r[k] = 0.;
            if (i > N) continue;
            const double r = noise_accurate(t1_r[k], t2_r[k]);
            if (i == N) {  // the omega work-item, :103-110, :155-167
                const double dw = dmul(nscale_omega, r);
                const double newomega = dadd(om, dmul(A.intconst, dw));
                const double top = dmul((double)(N - 1), dt);
                double o;
                if (newomega > top) o = dsub(dmul(dmul(2., (double)(N - 1)), dt), newomega);
                else if (newomega < 0) o = -newomega;
                else o = newomega;
                sh_om[cur ^ 1] = o;
                continue;
            }
            const double dw = dmul(nscale_site, r);
            const double fi = f_s[i];
            const double cl = clas(dmul((double)i, dt), om, pot);
            double inner;
            if (i == 0) {  // :74   f[1]+boundary(-1)-clas(-dt)-2f[0]
                inner = dsub(dsub(dadd(f_s[1], -ETA), clas(dmul(-1., dt), om, pot)), dmul(2., fi));
            } else if (i == N - 1) {  // :92   f[N-2]+boundary(1)-clas(N dt)-2f[N-1]
                inner = dsub(dsub(dadd(f_s[N - 2], ETA), clas(dmul((double)N, dt), om, pot)), dmul(2., fi));
            } else {  // :114
                inner = dsub(dadd(f_s[i + 1], f_s[i - 1]), dmul(2., fi));
            }
            // f + m*dtau*inner/dt2 - ddPot(cl)*f*dtau + dw      (m = 1: m*dtau == dtau)
            double nf = dadd(dsub(dadd(fi, ddiv_by(dmul(dtau, inner), A.dt2, r_dt2)),
                                  dmul(dmul(ddPot(cl, pot), fi), dtau)),
                             dw);
            if (nf > 1000.) nf = 1000.;  // :122-132
            if (nf < -1000.) nf = -1000.;
            if (isinf((float)nf) || isnan((float)nf)) nf = 1000.;
            nf_r[k] = nf;
            const double v = dadd(nf, cl);
            v_s[i] = v;
            a_s[i] = absol(v);
            d_s[i] = absol(dsub(dsub(nf, fi), dw));
            if (IPT == 1) {
                my_v = v;
                my_a = absol(v);
                my_d = absol(dsub(dsub(nf, fi), dw));
            }
            // :144-145, pre-update field
            const double path = dadd(fi, cl);
            xx0_r[k] = dadd(xx0_r[k], ddiv_by(dsub(dmul(path, dadd(fmid, clmid)), xx0_r[k]), n_inv_den, r_n));
            x_r[k] = dadd(x_r[k], ddiv_by(dsub(path, x_r[k]), n_inv_den, r_n));
        }
        // ---- phase 3: field hand-over (all) + stability scan, :135-143 ---------------------
        // IPT == 1 (N <= 1023, the reference's default 200): every thread holds its site's (v, a, d) in
        // registers and the as-if-sequential scan is a block-wide exclusive prefix maximum:
        //   record_i  <=>  v_i > T_i,   T_i = max(R0, max_{j<i} v_j)              if a record precedes lrgEl (case A)
        //                              T_i = R0 (i < lrgEl) | max_{lrgEl<=j<i} v_j (i > lrgEl); i == lrgEl never records
        //   unstable  <=>  some record has d_i > max(lrgVl, max_{j<i} a_j)
        // (the per-site rule of the chunked scan below, which stays for IPT > 1).  The warp-level part
        // runs BEFORE the barrier that ends phase 2, so no warp waits for another one's serial loop.
        int caseA_blk = 0;
        double pv = my_v, pu = -INFINITY, pa = my_a;
        if (IPT == 1) {
            const int lane = tid & 31, w = tid >> 5;
            pu = (tid >= E0) ? my_v : -INFINITY;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double tv = __shfl_up_sync(0xffffffffu, pv, o);
                const double tu = __shfl_up_sync(0xffffffffu, pu, o);
                const double ta = __shfl_up_sync(0xffffffffu, pa, o);
                if (lane >= o) {
                    pv = fmax(pv, tv);
                    pu = fmax(pu, tu);
                    pa = fmax(pa, ta);
                }
            }
            if (lane == 31) {
                sc_tot[0][w] = pv;
                sc_tot[1][w] = pu;
                sc_tot[2][w] = pa;
            }
            // (sh_R0 was published by the barrier of phase 1)
            caseA_blk = __syncthreads_or((tid < N && tid < E0 && my_v > sh_R0) ? 1 : 0);  // also ends phase 2
        } else {
            __syncthreads();
        }

#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int i = tid * IPT + k;
            if (i < N) {
                f_s[i] = nf_r[k];
                nfp_s[i] = nf_r[k];
            }
        }
        if (IPT == 1) {
            const int lane = tid & 31, w = tid >> 5, nw = (int)blockDim.x >> 5;
            const double NEG = -INFINITY, R0 = sh_R0;
            double ov = NEG, ou = NEG, oa = NEG, alla = NEG;  // maxima over the warps before this one / all
            for (int q = 0; q < nw; ++q) {
                const double tv = sc_tot[0][q], tu = sc_tot[1][q], ta = sc_tot[2][q];
                if (q < w) {
                    ov = fmax(ov, tv);
                    ou = fmax(ou, tu);
                    oa = fmax(oa, ta);
                }
                alla = fmax(alla, ta);
            }
            double ev_ = __shfl_up_sync(0xffffffffu, pv, 1), eu_ = __shfl_up_sync(0xffffffffu, pu, 1),
                   ea_ = __shfl_up_sync(0xffffffffu, pa, 1);
            if (lane == 0) ev_ = eu_ = ea_ = NEG;
            ev_ = fmax(ev_, ov);
            eu_ = fmax(eu_, ou);
            ea_ = fmax(ea_, oa);
            const double T = caseA_blk ? fmax(R0, ev_) : ((tid > E0) ? eu_ : R0);
            const bool rec = (tid < N) && (caseA_blk || tid != E0) && (my_v > T);
            const int unst_i = (rec && my_d > fmax(Vl0, ea_)) ? 1 : 0;
            const int lr = __reduce_max_sync(0xffffffffu, rec ? tid : -1);
            const unsigned ub = __ballot_sync(0xffffffffu, unst_i);
            if (lane == 0) {
                sc_res[0][w] = lr;
                sc_res[1][w] = ub != 0u;
            }
            __syncthreads();
            int lastrec = -1, unst = 0;
            for (int q = 0; q < nw; ++q) {
                lastrec = max(lastrec, sc_res[0][q]);
                unst |= sc_res[1][q];
            }
            if (lastrec >= 0) E0c = lastrec;
            Vlc = fmax(Vl0, alla);
            unstable = unst;
            if (tid == 0) {  // for the epilogue
                sh_lrgEl = E0c;
                sh_lrgVl = Vlc;
                if (unst) sh_unstable = 1;
            }
        } else if (tid < 32) {
            const int lane = tid;
            const int chunk = (N + 31) / 32;
            const int b = lane * chunk, e = min(N, b + chunk);
            const double R0 = dadd(stale, clas(dmul((double)E0, dt), om, pot));
            const double NEG = -INFINITY;
            // pass 1: chunk maxima
            double mv = NEG, mu = NEG, ma = NEG;
            int c_lane = 0;
            for (int i = b; i < e; ++i) {
                const double v = v_s[i];
                mv = fmax(mv, v);
                if (i >= E0) mu = fmax(mu, v);
                ma = fmax(ma, a_s[i]);
                if (i < E0 && v > R0) c_lane = 1;
            }
            const bool caseA = __any_sync(0xffffffffu, c_lane);  // a record precedes E0
            // exclusive prefix maxima across lanes
            double pv = mv, pu = mu, pa = ma;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double tv = __shfl_up_sync(0xffffffffu, pv, o);
                const double tu = __shfl_up_sync(0xffffffffu, pu, o);
                const double ta = __shfl_up_sync(0xffffffffu, pa, o);
                if (lane >= o) {
                    pv = fmax(pv, tv);
                    pu = fmax(pu, tu);
                    pa = fmax(pa, ta);
                }
            }
            const double allA = __shfl_sync(0xffffffffu, pa, 31);
            pv = __shfl_up_sync(0xffffffffu, pv, 1);
            pu = __shfl_up_sync(0xffffffffu, pu, 1);
            pa = __shfl_up_sync(0xffffffffu, pa, 1);
            if (lane == 0) pv = pu = pa = NEG;
            // pass 2: literal replay of the chunk with the incoming state
            double T = caseA ? fmax(R0, pv) : ((b > E0) ? pu : R0);
            double Vl = fmax(Vl0, pa);
            int lastrec = -1, unst = 0;
            for (int i = b; i < e; ++i) {
                const double v = v_s[i];
                if (!caseA && i == E0) {
                    T = v;  // newf[lrgEl] was just overwritten: compares v > v, then tracks it
                } else if (v > T) {
                    lastrec = i;
                    if (d_s[i] > Vl) unst = 1;
                    T = v;
                }
                if (a_s[i] > Vl) Vl = a_s[i];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lastrec = max(lastrec, __shfl_xor_sync(0xffffffffu, lastrec, o));
            unst = __any_sync(0xffffffffu, unst);
            if (lane == 0) {
                if (lastrec >= 0) sh_lrgEl = lastrec;
                sh_lrgVl = fmax(Vl0, allA);
                if (unst) sh_unstable = 1;
            }
        }
        if (IPT != 1) {
            __syncthreads();
            unstable = sh_unstable;
        }
        cur ^= 1;
        if (unstable) {  // :169-171
            ++j;
            break;
        }
    }

    // ---- epilogue: the reference's buffers + the host's commit / rollback --------------
    for (int i = tid; i < N; i += blockDim.x) A.newf[i] = nfp_s[i];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const int i = tid * IPT + k;
        if (i < N) {
            A.newx[i] = x_r[k];
            A.newxx0[i] = xx0_r[k];
            if (!unstable) {  // tauhost.c:508-513 + :550-552
                A.f[i] = nfp_s[i];
                A.x[i] = x_r[k];
                A.xx0[i] = xx0_r[k];
            }
        }
    }
    if (tid == 0) {
        if (!unstable) *A.omega = sh_om[cur];  // else keep the pre-frame omega (:553)
        *A.seed = sh_seed;                     // never rolled back
        *A.lrgEl = sh_lrgEl;
        *A.lrgVl = sh_lrgVl;
        *A.stable = unstable ? 0 : 1;
        *A.steps_done = j;
    }
    // ---- controller mode: log the frame, adapt the step size (tauhost.c:517-545) ---------------
    if (A.ctl) {
        const int fr = A.ctl->frame;  // (read by every thread before thread 0 updates it below)
        const int slot = fr % A.log_cap;
        if (!unstable) {
            __syncthreads();  // x_r / xx0_r of the mid point were stored to A.x / A.xx0 above
            const double xm = A.x[midpt];
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const int i = tid * IPT + k;
                if (i < N) A.log_xavg[(size_t)slot * N + i] = dsub(xx0_r[k], dmul(x_r[k], xm));
            }
        }
        __syncthreads();
        if (tid == 0) {
            A.log_rec[slot].dtau = dtau;
            A.log_rec[slot].stable = unstable ? 0 : 1;
            A.log_rec[slot].steps = j;
            double nd = dtau;
            int sc = A.ctl->stab_cnt;
            if (!unstable) {
                if (sc > 10) {  // :523-528
                    sc = 0;
                    nd = ddiv_cold(dtau, 0.950);
                }
                ++sc;
                A.ctl->runs = runs0 + A.loops;
            } else {  // :533-545
                nd = dmul(dtau, 0.950);
                sc = 0;
                *A.stable = 1;  // the host re-arms the flag (:542-544)
            }
            A.ctl->dtau = nd;
            A.ctl->stab_cnt = sc;
            A.ctl->frame = fr + 1;
        }
    }
}

// <x>, <x^2> of the committed path f+cl and the host's xavg (tauhost.c:519-521), one CTA
__global__ void __launch_bounds__(256) compat_reduce_kernel(const double *f, const double *x, const double *xx0,
                                                            const double *omega, int N, double dt, int pot,
                                                            double *out) {
    const double om = *omega;
    const double xmid = x[N / 2];
    double a1 = 0, a2 = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double path = dadd(f[i], clas(dmul((double)i, dt), om, pot));
        a1 += path;
        a2 = fma(path, path, a2);
        out[8 + i] = dsub(xx0[i], dmul(x[i], xmid));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    __shared__ double red[2][8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = a1; red[1][w] = a2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 8; ++k) { s1 += red[0][k]; s2 += red[1][k]; }
        out[0] = s1;
        out[1] = s2;
    }
}
cudaError_t launch_compat_reduce(const double *f, const double *x, const double *xx0, const double *omega,
                                 int N, double dt, int pot, double *out, cudaStream_t stream) {
    compat_reduce_kernel<<<1, 256, 0, stream>>>(f, x, xx0, omega, N, dt, pot, out);
    return cudaGetLastError();
}

cudaError_t launch_compat1d(const Compat1DArgs &A, cudaStream_t stream) {
    const int items = A.N + 1;
    int ipt = 1;
    while ((items + ipt - 1) / ipt > 1024) ipt *= 2;
    if (ipt > 8) return cudaErrorInvalidValue;
    int threads = (items + ipt - 1) / ipt;
    threads = (threads + 31) / 32 * 32;
    const size_t smem = sizeof(double) * 5 * (size_t)A.N + sizeof(u64) * 2 * (size_t)(A.N + 1);
    cudaError_t e = cudaSuccess;
#define SQ_LAUNCH_IPT(K)                                                                          \
    case K:                                                                                       \
        e = cudaFuncSetAttribute(compat1d_frame_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                           \
        compat1d_frame_kernel<K><<<1, threads, smem, stream>>>(A);                                \
        break;
    switch (ipt) {
        SQ_LAUNCH_IPT(1)
        SQ_LAUNCH_IPT(2)
        SQ_LAUNCH_IPT(4)
        SQ_LAUNCH_IPT(8)
    }
#undef SQ_LAUNCH_IPT
    return cudaGetLastError();
}

}  // namespace sq
