// sq_slab.cu -- one lattice cut into time slabs over the GPUs of one box (north_star (4),
// SURVEY.md 8(e); the reference itself is single-device, tauhost.c:249-252).
//
// Data plane.  Every rank owns nt consecutive time slices plus a halo arena in its own HBM:
//     ghost[parity][dir][vslice] reals,  flag[parity][dir] u32      (dir 0 = below, 1 = above)
// The update kernel (sq_lattice.cu) computes the two boundary slices FIRST and stores them both
// into the local field and -- posted peer stores over NVLink -- into the neighbours' ghost buffers
// of the next parity; the last CTA of a boundary slice raises the neighbour's arrival flag with
// the step's tag.  The neighbour's boundary CTAs of the next step wait on that flag (it has long
// arrived: it was sent before the interior was computed).  No host, no NCCL, no extra kernel in
// the loop; double buffering by tag parity is sufficient because a writer of ghost[p][d] is gated
// by the completion of that buffer's readers of two steps ago.
//
// RNG.  All ranks draw from the ONE shared-seed chain (tau_kernel.cl:269-284) at global gids, by
// jump-ahead, so no noise is communicated.  The chain's data-dependent events (inf-retry :282,
// `seed+=` :278-279; ~1.25 per step at 256^4) change every later seed, so they must be known to
// all ranks BEFORE a step's update runs: an integer-only finder kernel scans each rank's slab of
// the step's draws, the ranks agree on the first event through the session (a few bytes), every
// rank replays that draw literally (deterministic), and only the part of the stream behind the
// event is scanned again.  The finder runs on its own stream, one step ahead of the update, so
// the host round trips hide behind the previous step's update kernel.
//
// Observables.  Slice sums stay per rank; the running means of Phi(t)Phi(t_mid)
// (tau_kernel.cl:145 at slice granularity) need the mid slice's sum of the same step: per-step
// histories are kept on the device and combined through the session at sq_sync.
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <new>
#include <vector>

#include "sq_ctx.h"
#include "sq_session.h"
#include "sq_site.cuh"

using namespace sq;

namespace sq {

constexpr int FW = 16;         // sites per finder strip
constexpr int SERIES_CHUNK = SESSION_SERIES_MAX / 3;

struct SlabState {
    sq_session *sess = nullptr;
    int rank = 0, nranks = 1;
    char *arena = nullptr;
    size_t ghost_bytes = 0, flags_off = 0, arena_bytes = 0;
    char *peer[2] = {nullptr, nullptr};  // lower / upper neighbour's arena, mapped on this device
    bool peer_ipc[2] = {false, false};
    unsigned *d_count = nullptr, *d_error = nullptr;
    unsigned tag = 1;  // next unused tag, identical on every rank (collective calls)
    cudaStream_t fstream = nullptr;
    u64 *d_found = nullptr, *h_found = nullptr;
    u64 seed = 0;      // host mirror of the step-start seed (full u64)
    int fgrid = 0;
    JumpEntry fstride{};
    double *d_hist = nullptr;
    size_t hist_cap = 0;  // steps
    int tmid_owner = 0, tmid_local = -1;
    // sequence in flight
    int seq_steps = 0;
    int64_t seq_runs0 = 0;
    // host mirrors of the running means (this slab's slices) and of the global sums
    std::vector<double> slice_x, slice_xx0, slice_sum;
    double sums[2] = {0, 0}, sums_mean[2] = {0, 0};
    uint64_t scans = 0, rounds = 0;

    char *ghost(char *base, int par, int dir) const { return base + (size_t)(par * 2 + dir) * ghost_bytes; }
    unsigned *flag(char *base, int par, int dir) const { return (unsigned *)(base + flags_off) + (par * 2 + dir); }
};

namespace {

__device__ __noinline__ void finder_cold(u64 *found, u64 sm, u64 g0, int w) {
    for (int e = 0; e < w; ++e) {
        u64 t1, t2;
        lcg_draw(sm, g0 + e, t1, t2);
        if (lcg_event(sm, t1, t2)) atomicMin((unsigned long long *)found, g0 + e);
        sm = lcg_next_seed(t2) & LCG_MASK;
    }
}

// Integer-only scan of the draws at gids [g_from, g_hi): the first gid whose draw meets the
// necessary condition of an event (lcg_event) under the base (bg, bs) -- "draws at gid >= bg
// chain from seed bs".  No memory traffic.  Both events are decidable from the SEED sequence z
// (z' = ALPHA z + BETA g + GAMMA: one 48-bit multiply-add per site instead of the two of a draw):
//   inf-retry  t1 < 2^16        => (t1 mod 2^32) < 2^16, and the low 32 bits of t1 = z A + c are one IMAD;
//   `seed+=`   z < 2^31 && ...  => bits 32..47 of z are zero.
// One 3-input min per site collects both filters (threshold 2^16); candidates (p ~ 2^-15 per site)
// are decided exactly on a cold path.  The strip-to-strip advance of a thread is one affine map with
// a fixed stride in 32-bit limbs, as in sq_march.cu.
constexpr unsigned ALPHA_LO = (unsigned)LCG_ALPHA, ALPHA_HI = (unsigned)(LCG_ALPHA >> 32);

__global__ void __launch_bounds__(256) find_events_kernel(u64 bs, u64 bg, u64 g_from, u64 g_hi, const JumpEntry *jump,
                                                          JumpEntry stride_jump, u64 *found) {
    const u64 nthreads = (u64)gridDim.x * blockDim.x;
    u64 g = g_from + ((u64)blockIdx.x * blockDim.x + threadIdx.x) * FW;
    if (g >= g_hi) return;
    const u64 stride = nthreads * FW;
    Seed32 z = seed_split(lcg_seed_at(bs, bg, g - bg, jump));
    const unsigned aDl = (unsigned)stride_jump.a, aDh = (unsigned)(stride_jump.a >> 32);
    u64 ck = (LCG_BETA * g + LCG_GAMMA) * stride_jump.g0 + stride_jump.bg1;  // z(next strip) = alpha^D z + ck
    u64 dck = LCG_BETA * stride * stride_jump.g0;
    u64 K = LCG_BETA * g + LCG_GAMMA;                     // z(next site) = ALPHA z + K
    u64 dK = LCG_BETA * stride;
    unsigned clo = (unsigned)site_const(g);               // low limb of gid*A + B
    unsigned dclo = (unsigned)(stride * LCG_A);
    asm volatile("" : "+l"(dck), "+l"(dK), "+r"(dclo));
    for (;;) {
        const Seed32 z0 = z;
        unsigned m = 0xFFFFFFFFu;
        u64 Ke = K;
        unsigned ce = clo;
#pragma unroll
        for (int e = 0; e < FW; ++e) {
            const unsigned t1l = z.lo * A_LO + ce;                  // t1 mod 2^32
            // z < 2^31 => bits 32..47 of z are zero => (z.hi << 16) == 0; a plain shift (the assembler
            // may put it on the FMA pipe as IMAD.SHL: this kernel is ALU-pipe bound) instead of a funnel shift
            const unsigned zs = z.hi << 16;
            m = min(min(m, t1l), zs);
            u64 p;
            unsigned pl, ph, t;
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(z.lo), "r"(ALPHA_LO), "l"(Ke));
            asm("mov.b64 {%0, %1}, %2;" : "=r"(pl), "=r"(ph) : "l"(p));
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(z.lo), "r"(ALPHA_HI), "r"(ph));
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(z.hi) : "r"(z.hi), "r"(ALPHA_LO), "r"(t));
            z.lo = pl;
            Ke += LCG_BETA;
            ce += A_LO;
        }
        if (__builtin_expect(m < 65536u, 0))
            finder_cold(found, seed_join(z0), g, (int)((g_hi - g < (u64)FW) ? g_hi - g : (u64)FW));
        const u64 gn = g + stride;
        if (gn >= g_hi) break;
        {   // next strip of this thread: stride draws further
            const u64 p = (u64)z0.lo * aDl + ck;
            z.lo = (unsigned)p;
            z.hi = (unsigned)(p >> 32) + z0.lo * aDh + z0.hi * aDl;
        }
        ck += dck;
        K += dK;
        clo += dclo;
        g = gn;
    }
}

__global__ void slab_flags_kernel(unsigned *f0, unsigned *f1, unsigned tag) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f0), "r"(tag) : "memory");
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f1), "r"(tag) : "memory");
}

}  // namespace
}  // namespace sq

// ------------------------------------------------------------------ host utility ------------
extern "C" int sq_rng_resolve(uint64_t step_seed, const sq_rng_entry *entries, int n, uint64_t gid, sq_rng_entry *out,
                              int *ndraws, int *plus) {
    if (!out || n < 0 || (n > 0 && !entries)) return SQ_ERR_INVALID;
    const JumpEntry *tab = host_jump_table();
    u64 bg = 0, bs = step_seed;
    for (int k = 0; k < n; ++k)
        if (entries[k].gid_start <= gid && entries[k].gid_start >= bg) {
            bg = entries[k].gid_start;
            bs = entries[k].seed;
        }
    u64 sfull = bs;
    if (gid != bg) {  // the draw at gid-1 was event-free: seed = t2(gid-1) - 2^31 as a wrapping u64
        const u64 sp = (gid - 1 == bg) ? bs : lcg_seed_at(bs, bg, gid - 1 - bg, tab);
        u64 t1, t2;
        lcg_draw(sp, gid - 1, t1, t2);
        sfull = lcg_next_seed(t2);
    }
    const HostDraw h = host_draw_literal(sfull, gid);
    out->gid_start = gid + 1;
    out->seed = h.seed_after;
    out->ov_gid = gid;
    out->ov_t1 = h.t1;
    out->ov_t2 = h.t2;
    if (ndraws) *ndraws = h.ndraws;
    if (plus) *plus = h.plus;
    return SQ_OK;
}

// ------------------------------------------------------------------ join / destroy ----------
void sq_slab_destroy(sq_ctx *c) {
    SlabState *sl = c->slab;
    if (!sl) return;
    cudaSetDevice(c->p.device);
    if (sl->fstream) cudaStreamSynchronize(sl->fstream);
    for (int d = 0; d < 2; ++d)
        if (sl->peer_ipc[d] && sl->peer[d] && !(d == 1 && sl->peer[1] == sl->peer[0])) cudaIpcCloseMemHandle(sl->peer[d]);
    if (sl->arena) cudaFree(sl->arena);
    if (sl->d_count) cudaFree(sl->d_count);
    if (sl->d_error) cudaFree(sl->d_error);
    if (sl->d_found) cudaFree(sl->d_found);
    if (sl->h_found) cudaFreeHost(sl->h_found);
    if (sl->d_hist) cudaFree(sl->d_hist);
    if (sl->fstream) cudaStreamDestroy(sl->fstream);
    delete sl;
    c->slab = nullptr;
}

static int slab_fail(sq_ctx *c, int rc) {
    if (c->slab && c->slab->sess) sq_session_abort(c->slab->sess);
    sq_slab_destroy(c);
    return rc;
}

extern "C" int sq_slab_join(sq_ctx *c, sq_session *s) {
    if (!c || !s || !s->shm || c->p.kernel != SQ_KERNEL_LATTICE || c->pending || c->slab) return SQ_ERR_INVALID;
    if (c->p.nchains != 1) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    const sq_params &p = c->p;
    const int R = s->nranks, me = s->rank;
    // ---- the slabs must tile [0, Lt) in rank order; lattice and seed must agree ----------------
    u64 seed0 = 0;
    CK(cudaMemcpy(&seed0, c->l_seeds[c->cur], sizeof(u64), cudaMemcpyDeviceToHost));
    uint64_t mine[6] = {(uint64_t)p.slab_t0, (uint64_t)p.slab_nt, seed0, (uint64_t)c->V,
                        (uint64_t)(p.real * 16 + p.ndim), (uint64_t)c->runs};
    std::vector<uint64_t> all((size_t)R * 6);
    if ((rc = sq_session_allgather_u64(s, mine, 6, all.data()))) return rc;
    uint64_t t = 0;
    for (int r = 0; r < R; ++r) {
        const uint64_t *a = &all[(size_t)r * 6];
        if (a[0] != t || a[2] != seed0 || a[3] != (uint64_t)c->V || a[4] != mine[4] || a[5] != mine[5]) return SQ_ERR_INVALID;
        t += a[1];
    }
    if (t != (uint64_t)p.dims[p.ndim - 1]) return SQ_ERR_INVALID;

    SlabState *sl = new (std::nothrow) SlabState();
    if (!sl) return SQ_ERR_NOMEM;
    c->slab = sl;
    sl->sess = s;
    sl->rank = me;
    sl->nranks = R;
    sl->seed = seed0;
    const int64_t tmid = p.dims[p.ndim - 1] / 2;
    for (int r = 0; r < R; ++r)
        if ((uint64_t)tmid >= all[(size_t)r * 6] && (uint64_t)tmid < all[(size_t)r * 6] + all[(size_t)r * 6 + 1]) sl->tmid_owner = r;
    sl->tmid_local = (sl->tmid_owner == me) ? (int)(tmid - p.slab_t0) : -1;
    sl->slice_x.assign((size_t)c->nt, 0.);
    sl->slice_xx0.assign((size_t)c->nt, 0.);
    sl->slice_sum.assign((size_t)c->nt, 0.);

    // ---- device side ---------------------------------------------------------------------------
    sl->ghost_bytes = ((size_t)c->vslice * c->rsz + 255) & ~(size_t)255;
    sl->flags_off = 4 * sl->ghost_bytes;
    sl->arena_bytes = sl->flags_off + 256;
    cudaError_t e = cudaMalloc((void **)&sl->arena, sl->arena_bytes);
    if (e == cudaSuccess) e = cudaMemset(sl->arena, 0, sl->arena_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&sl->d_count, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(sl->d_count, 0, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc((void **)&sl->d_error, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(sl->d_error, 0, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc((void **)&sl->d_found, sizeof(u64));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&sl->h_found, sizeof(u64));
    int lo_pri = 0, hi_pri = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&sl->fstream, cudaStreamNonBlocking, hi_pri);
    int sms = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p.device);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        snprintf(cuda_err_buf(), 512, "sq_slab_join: %s", cudaGetErrorString(e));
        return slab_fail(c, SQ_ERR_CUDA);
    }
    sl->fgrid = std::max(1, sms * 4);
    sl->fstride = jump_entry((u64)sl->fgrid * 256 * FW);

    // ---- exchange the arenas: CUDA IPC between processes, plain pointers inside one -----------
    SessionRankSlot &slot = s->shm->slot[me];
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, sl->arena);
    if (e != cudaSuccess) {
        snprintf(cuda_err_buf(), 512, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return slab_fail(c, SQ_ERR_CUDA);
    }
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(slot.ipc), "ipc handle size");
    memcpy(slot.ipc, &h, sizeof h);
    slot.pid = (uint64_t)getpid();
    slot.raw_ptr = (uint64_t)(uintptr_t)sl->arena;
    slot.device = p.device;
    if ((rc = sq_session_barrier(s))) return slab_fail(c, rc);
    const int nb[2] = {(me + R - 1) % R, (me + 1) % R};
    for (int d = 0; d < 2; ++d) {
        if (d == 1 && nb[1] == nb[0]) {  // ring of 1 or 2: both neighbours are the same rank
            sl->peer[1] = sl->peer[0];
            sl->peer_ipc[1] = sl->peer_ipc[0];
            continue;
        }
        const SessionRankSlot &o = s->shm->slot[nb[d]];
        if (o.pid == (uint64_t)getpid()) {
            sl->peer[d] = (char *)(uintptr_t)o.raw_ptr;
            if (o.device != p.device) {
                e = cudaDeviceEnablePeerAccess(o.device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            }
        } else {
            cudaIpcMemHandle_t oh;
            memcpy(&oh, o.ipc, sizeof oh);
            e = cudaIpcOpenMemHandle((void **)&sl->peer[d], oh, cudaIpcMemLazyEnablePeerAccess);
            sl->peer_ipc[d] = (e == cudaSuccess);
        }
        if (e != cudaSuccess) {
            snprintf(cuda_err_buf(), 512, "mapping the neighbour's halo arena: %s", cudaGetErrorString(e));
            return slab_fail(c, SQ_ERR_CUDA);
        }
    }
    if ((rc = sq_session_barrier(s))) return slab_fail(c, rc);
    return SQ_OK;
}

// ------------------------------------------------------------------ per-step event agreement --
// Entries of the step whose start seed is S; S_next = seed after the step's V+1 draws.
static int resolve_step(sq_ctx *c, u64 S, std::vector<RebaseEntry> &entries, u64 &S_next) {
    SlabState *sl = c->slab;
    entries.clear();
    // The scan needs no field data, so every round's remaining range [from, V) is split evenly over
    // ALL ranks (not by slab): the ranks below an event would otherwise idle while those above rescan.
    const u64 Vg = (u64)c->V;
    u64 bg = 0, bs = S, from = 0;
    std::vector<uint64_t> all((size_t)sl->nranks);
    for (;;) {
        uint64_t local = NO_EVENT;
        const u64 per = ((Vg - from + (u64)sl->nranks - 1) / (u64)sl->nranks + FW - 1) / FW * FW;
        const u64 lo = std::min(Vg, from + (u64)sl->rank * per), hi = std::min(Vg, lo + per);
        if (lo < hi) {
            *sl->h_found = NO_EVENT;
            CK(cudaMemcpyAsync(sl->d_found, sl->h_found, sizeof(u64), cudaMemcpyHostToDevice, sl->fstream));
            find_events_kernel<<<sl->fgrid, 256, 0, sl->fstream>>>(bs, bg, lo, hi, c->d_jump, sl->fstride, sl->d_found);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(sl->h_found, sl->d_found, sizeof(u64), cudaMemcpyDeviceToHost, sl->fstream));
            CK(cudaStreamSynchronize(sl->fstream));
            local = *sl->h_found;
            c->launches++;
            sl->scans++;
        }
        sl->rounds++;
        int rc = sq_session_allgather_u64(sl->sess, &local, 1, all.data());
        if (rc) return rc;
        const uint64_t g = *std::min_element(all.begin(), all.end());
        if (g == NO_EVENT) break;
        if ((int)entries.size() >= MAX_REBASE) return SQ_ERR_INTERNAL;
        // every rank replays the same draw from the same (S, entries): identical entries everywhere
        const u64 sfull = sq_host_seed_before(c, entries, 0, S, g);
        const HostDraw h = host_draw_literal(sfull, g);
        RebaseEntry e{};
        e.gid_start = g + 1;
        e.seed = h.seed_after;
        e.ov_gid = g;
        e.ov_t1 = h.t1;
        e.ov_t2 = h.t2;
        e.chain = 0;
        e.vseed = virtual_start_seed(e.seed, e.gid_start, c->h_jump.data());
        entries.push_back(e);
        bg = g + 1;
        bs = h.seed_after;
        from = g + 1;  // everything up to the event is final; the rest is scanned again under the new base
    }
    // the omega work-item's draw at gid V (tau_kernel.cl:103-110)
    const u64 sV = sq_host_seed_before(c, entries, 0, S, Vg);
    u64 t1, t2;
    lcg_draw(sV, Vg, t1, t2);
    const HostDraw h = host_draw_literal(sV, Vg);
    if (lcg_event(sV & LCG_MASK, t1, t2)) {
        if ((int)entries.size() >= MAX_REBASE) return SQ_ERR_INTERNAL;
        RebaseEntry e{};
        e.gid_start = Vg + 1;
        e.seed = h.seed_after;
        e.ov_gid = Vg;
        e.ov_t1 = h.t1;
        e.ov_t2 = h.t2;
        e.chain = 0;
        e.vseed = virtual_start_seed(e.seed, e.gid_start, c->h_jump.data());
        entries.push_back(e);
    }
    S_next = h.seed_after;
    return SQ_OK;
}

// ------------------------------------------------------------------ sequence ----------------
int sq_slab_enqueue(sq_ctx *c, double dtau, int nsteps, int64_t runs0) {
    SlabState *sl = c->slab;
    const sq_params &p = c->p;
    sl->seq_steps = nsteps;
    sl->seq_runs0 = runs0;
    if (nsteps == 0) return SQ_OK;
    const int nt = c->nt;
    if ((size_t)nsteps > sl->hist_cap) {
        if (sl->d_hist) CK(cudaFree(sl->d_hist));
        sl->d_hist = nullptr;
        sl->hist_cap = 0;
        CK(cudaMalloc((void **)&sl->d_hist, sizeof(double) * (size_t)nsteps * (size_t)(nt + 2)));
        sl->hist_cap = (size_t)nsteps;
    }
    // ---- publish the current boundary slices under a fresh tag (the field may have been
    //      uploaded since the last sequence) ----------------------------------------------------
    const unsigned T0 = sl->tag;
    sl->tag += (unsigned)nsteps;
    {
        const int par = (int)(T0 & 1u);
        const char *f = (const char *)c->l_field[c->cur];
        const size_t sb = (size_t)c->vslice * c->rsz;
        CK(cudaMemcpyAsync(sl->ghost(sl->peer[0], par, 1), f, sb, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(sl->ghost(sl->peer[1], par, 0), f + (size_t)(nt - 1) * sb, sb, cudaMemcpyDefault, c->stream));
        slab_flags_kernel<<<1, 1, 0, c->stream>>>(sl->flag(sl->peer[0], par, 1), sl->flag(sl->peer[1], par, 0), T0);
        CK(cudaGetLastError());
        c->launches++;
    }
    const int Lt = (int)p.dims[p.ndim - 1];
    (void)Lt;
    u64 S = sl->seed;
    std::vector<RebaseEntry> entries;
    for (int k = 0; k < nsteps; ++k) {
        u64 S_next = 0;
        int rc = resolve_step(c, S, entries, S_next);  // blocks on the finder stream only
        if (rc) return rc;
        c->nevents += entries.size();
        if (!entries.empty())
            CK(cudaMemcpyAsync(c->l_rebase, entries.data(), sizeof(RebaseEntry) * entries.size(), cudaMemcpyHostToDevice, c->stream));
        LatticeArgs A = sq_lattice_args(c, dtau, k);
        A.n_rebase = (int)entries.size();
        sq_fill_rebase_inline(A, entries.data(), A.n_rebase);
        A.nclamped = c->l_nclamp_step + k;
        A.wrap_time = 0;
        const unsigned Tw = T0 + (unsigned)k, Tp = Tw + 1u;
        const int pw = (int)(Tw & 1u), pp = (int)(Tp & 1u);
        A.slab_on = 1;
        A.ghost_lo = sl->ghost(sl->arena, pw, 0);
        A.ghost_hi = sl->ghost(sl->arena, pw, 1);
        A.wait_tag = Tw;
        A.wait_flag[0] = sl->flag(sl->arena, pw, 0);
        A.wait_flag[1] = sl->flag(sl->arena, pw, 1);
        A.push_tag = (k + 1 < nsteps) ? Tp : 0u;
        A.push_ghost[0] = sl->ghost(sl->peer[0], pp, 1);
        A.push_ghost[1] = sl->ghost(sl->peer[1], pp, 0);
        A.push_flag[0] = sl->flag(sl->peer[0], pp, 1);
        A.push_flag[1] = sl->flag(sl->peer[1], pp, 0);
        A.push_count = sl->d_count;
        A.slab_error = sl->d_error;
        FinalizeArgs F{};
        F.nt = nt;
        F.nchains = 1;
        F.ctas_per_slice = c->ctas_per_slice;
        F.tmid_local = -1;
        F.vslice = c->vslice;
        F.runs = runs0 + k;
        F.partials = c->l_partials;
        F.slice_sum = c->l_slice_sum;
        F.slice_x = c->l_slice_x;
        F.slice_xx0 = c->l_slice_xx0;
        F.sums = c->l_sums;
        F.sums_mean = c->l_sums_mean;
        F.history = sl->d_hist + (size_t)k * (size_t)(nt + 2);
        F.event_key = c->l_event;
        F.step_index = k;
        if ((rc = sq_enqueue_step(c, A, F, k))) return rc;
        S = S_next;
    }
    sl->seed = S;
    CK(launch_commit_clamps(c->l_nclamp_step, nsteps, nsteps, c->l_nclamped, nullptr, c->stream));
    c->launches++;
    return sq_join_finalize(c);
}

int sq_slab_finish(sq_ctx *c) {
    SlabState *sl = c->slab;
    const int nsteps = sl->seq_steps, nt = c->nt;
    CK(cudaStreamSynchronize(c->stream));
    if (c->timing) {
        int rt = sq_timing_collect(c, (size_t)-1);
        if (rt) return rt;
    }
    u64 key = NO_EVENT, dseed = 0;
    unsigned err = 0;
    CK(cudaMemcpy(&key, c->l_event, sizeof key, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&err, sl->d_error, sizeof err, cudaMemcpyDeviceToHost));
    if (err) {
        sq_session_abort(sl->sess);
        return SQ_ERR_TIMEOUT;
    }
    c->cur = (c->cur + nsteps) & 1;
    if (nsteps > 0) {
        CK(cudaMemcpy(&dseed, c->l_seeds[c->cur], sizeof dseed, cudaMemcpyDeviceToHost));
        // the update kernels must never meet an event the finder has not announced, and must
        // have walked the chain to the same seed as the host replay
        if (key != NO_EVENT || dseed != sl->seed) {
            snprintf(cuda_err_buf(), 512, "slab: device chain diverged (event key %llx, seed %llx vs %llx)",
                     (unsigned long long)key, (unsigned long long)dseed, (unsigned long long)sl->seed);
            sq_session_abort(sl->sess);
            return SQ_ERR_INTERNAL;
        }
    }
    // ---- observables: per-step histories -> running means (tau_kernel.cl:144-145 per slice) ----
    if (nsteps > 0 && !(c->p.flags & SQ_FLAG_NO_OBSERVABLES)) {
        std::vector<double> hist((size_t)nsteps * (size_t)(nt + 2));
        CK(cudaMemcpy(hist.data(), sl->d_hist, sizeof(double) * hist.size(), cudaMemcpyDeviceToHost));
        std::vector<double> mine((size_t)3 * SERIES_CHUNK), all((size_t)sl->nranks * 3 * SERIES_CHUNK);
        const double vs = (double)c->vslice, vol = (double)c->V;
        for (int k0 = 0; k0 < nsteps; k0 += SERIES_CHUNK) {
            const int n = std::min(SERIES_CHUNK, nsteps - k0);
            for (int k = 0; k < n; ++k) {
                const double *h = &hist[(size_t)(k0 + k) * (size_t)(nt + 2)];
                mine[3 * k] = sl->tmid_local >= 0 ? h[sl->tmid_local] : 0.;
                mine[3 * k + 1] = h[nt];
                mine[3 * k + 2] = h[nt + 1];
            }
            int rc = sq_session_allgather_f64(sl->sess, mine.data(), 3 * n, all.data());
            if (rc) return rc;
            for (int k = 0; k < n; ++k) {
                const double *h = &hist[(size_t)(k0 + k) * (size_t)(nt + 2)];
                const double cnt = (double)(sl->seq_runs0 + k0 + k + 1);
                const double pmid = all[(size_t)sl->tmid_owner * 3 * n + 3 * k] / vs;
                double s1 = 0, s2 = 0;
                for (int r = 0; r < sl->nranks; ++r) {
                    s1 += all[(size_t)r * 3 * n + 3 * k + 1];
                    s2 += all[(size_t)r * 3 * n + 3 * k + 2];
                }
                for (int t = 0; t < nt; ++t) {
                    const double P = h[t] / vs;
                    sl->slice_xx0[t] = sl->slice_xx0[t] + (P * pmid - sl->slice_xx0[t]) / cnt;
                    sl->slice_x[t] = sl->slice_x[t] + (P - sl->slice_x[t]) / cnt;
                    sl->slice_sum[t] = h[t];
                }
                sl->sums[0] = s1;
                sl->sums[1] = s2;
                sl->sums_mean[0] += (s1 / vol - sl->sums_mean[0]) / cnt;
                sl->sums_mean[1] += (s2 / vol - sl->sums_mean[1]) / cnt;
            }
        }
    } else {
        int rc = sq_session_barrier(sl->sess);
        if (rc) return rc;
    }
    c->runs = sl->seq_runs0 + nsteps;
    c->last_stable = 1;
    c->last_steps = nsteps;
    return SQ_OK;
}

// slab part of sq_measure: the running means live on the host in this mode
void sq_slab_measure(sq_ctx *c, sq_obs *o) {
    SlabState *sl = c->slab;
    const int nt = c->nt;
    o->seed = sl->seed;
    if (o->slice_x) memcpy(o->slice_x, sl->slice_x.data(), sizeof(double) * nt);
    if (o->slice_xx0) memcpy(o->slice_xx0, sl->slice_xx0.data(), sizeof(double) * nt);
    if (o->corr) {  // corr[t] = xx0[t] - x[t] x[t_mid]: the mid slice's running mean lives on its owner
        double xm = sl->tmid_local >= 0 ? sl->slice_x[(size_t)sl->tmid_local] : 0.;
        std::vector<double> all((size_t)sl->nranks);
        if (sq_session_allgather_f64(sl->sess, &xm, 1, all.data()) == SQ_OK) xm = all[(size_t)sl->tmid_owner];
        for (int t = 0; t < nt; ++t) o->corr[t] = sl->slice_xx0[t] - sl->slice_x[t] * xm;
    }
}

void sq_slab_stats_impl(sq_ctx *c, uint64_t *scans, uint64_t *rounds) {
    if (scans) *scans = c->slab ? c->slab->scans : 0;
    if (rounds) *rounds = c->slab ? c->slab->rounds : 0;
}
