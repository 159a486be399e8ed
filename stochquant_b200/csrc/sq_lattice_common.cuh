// sq_lattice_common.cuh -- device helpers shared by the streaming lattice kernels
// (sq_lattice.cu: generic strips; sq_march.cu: row-marching fp32 kernel).
#pragma once
#include "sq_kernels.h"
#include "sq_site.cuh"

namespace sq {
namespace {

template <typename real>
struct alignas(16) Pack {
    real v[16 / sizeof(real)];
};

template <typename real> struct Ops;
template <> struct Ops<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
};
template <> struct Ops<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// base (gid, seed) applicable to draws starting at gid g under the step's rebase list
__device__ __forceinline__ void rebase_lookup(const LatticeArgs &A, int chain, u64 S, u64 g,
                                              u64 &bg, u64 &bs) {
    bg = 0;
    bs = S;
    for (int j = 0; j < A.n_rebase; ++j) {
        const RebaseEntry e = A.rebase[j];
        if (e.chain == chain && e.gid_start <= g && e.gid_start >= bg) {
            bg = e.gid_start;
            bs = e.seed;
        }
    }
}

// cold paths, arguments by value (taking the address of a register array would spill it)
// returns true if the strip holds an event: the launch will be replayed, so nothing this strip computed
// (in particular a clamp hit caused by the +-inf of the draw that is about to be retried) may be counted
__device__ __noinline__ bool strip_events_cold(u64 *event_key_ptr, int step, int chain, u64 sm, u64 g0, int w) {
    bool any = false;
    for (int e = 0; e < w; ++e) {
        u64 t1, t2;
        lcg_draw(sm, g0 + e, t1, t2);
        if (lcg_event(sm, t1, t2)) {
            atomicMin((unsigned long long *)event_key_ptr, event_key(step, chain, g0 + e));
            any = true;
        }
        sm = lcg_next_seed(t2) & LCG_MASK;
    }
    return any;
}

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// slab ring: wait until the neighbour's boundary slice of this step's input field has landed in the
// local ghost buffer (tags are monotonic, so ">=" in wrap-around arithmetic).  Bounded.
__device__ __noinline__ void slab_wait(const unsigned *flag, unsigned want, unsigned *err) {
    for (unsigned spins = 0; (int)(ld_acquire_sys_u32(flag) - want) < 0; ++spins) {
        __nanosleep(64);
        if (spins > (1u << 24)) {  // ~ seconds: the neighbour is gone
            atomicExch(err, 1u);
            return;
        }
    }
}


// CTA -> (time slice tl, position bx inside the slice).  CTAs are dispatched in linear order
// (x fastest): a plain (x = position, y = slice) grid streams whole slices, and with slices
// larger than ~L2/4 every site is fetched from HBM three times (as t-1, t, t+1).  Instead a
// chunk of `ctas_per_chunk` positions is swept through all time slices before the next chunk
// starts, so the three time levels of a chunk are still in L2 when they are needed again.
// Slab ring: the two boundary slices come first, whole, so that their output reaches the
// neighbours while the interior is still being computed.
__device__ __forceinline__ void cta_slice_position(const LatticeArgs &A, int &tl, unsigned &bx) {
    const unsigned cps = gridDim.x;
    unsigned lin = blockIdx.y * cps + blockIdx.x;
    unsigned t_first = 0, nt_sweep = (unsigned)A.nt;
    if (A.slab_on && A.nt > 1) {
        if (lin < 2 * cps) {
            tl = (lin < cps) ? 0 : A.nt - 1;
            bx = (lin < cps) ? lin : lin - cps;
            return;
        }
        lin -= 2 * cps;
        t_first = 1;
        nt_sweep = (unsigned)A.nt - 2;
    }
    const unsigned cpc = (unsigned)A.ctas_per_chunk;
    const unsigned per_chunk = cpc * nt_sweep;
    const unsigned chunk = lin / per_chunk;
    const unsigned rem = lin - chunk * per_chunk;
    const unsigned width = (chunk * cpc + cpc <= cps) ? cpc : cps - chunk * cpc;  // the last chunk may be narrower
    const unsigned ty = rem / width;
    tl = (int)(t_first + ty);
    bx = chunk * cpc + (rem - ty * width);
}

}  // namespace
}  // namespace sq
