// sq_ctx.h -- the context behind the opaque sq_ctx handle of include/sq.h, shared by sq_api.cu
// (single-GPU sequencing, event replay) and sq_slab.cu (multi-GPU slab decomposition).
#pragma once
#include <stdio.h>

#include <vector>

#include "../../include/sq.h"
#include "sq_kernels.h"

namespace sq {
char *cuda_err_buf();  // thread-local text of the last CUDA failure (512 bytes)
struct SlabState;      // sq_slab.cu
}

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            snprintf(sq::cuda_err_buf(), 512, "%s at %s:%d: %s", #call, __FILE__, __LINE__,    \
                     cudaGetErrorString(e__));                                                 \
            return SQ_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

constexpr int MAX_SEQ_STEPS = 32768;  // step field of the event key has 16 bits
constexpr int MAX_REBASE = 64;
constexpr int RES_MAX_STEPS = 2048;  // tau-steps per resident launch (history buffer)

struct sq_ctx {
    using JumpEntry = sq::JumpEntry;
    using RebaseEntry = sq::RebaseEntry;
    using u64 = sq::u64;
    sq_params p{};
    cudaStream_t stream = nullptr;
    JumpEntry *d_jump = nullptr;
    std::vector<JumpEntry> h_jump;
    int64_t launches = 0;
    int64_t runs = 0;
    int last_stable = 1;
    int64_t last_steps = 0;
    uint64_t nevents = 0;
    void *h_pin = nullptr;  // pinned scratch (4 KB)
    void *h_pin2 = nullptr; // pinned scratch of sq_measure (lattice), grown on demand
    size_t h_pin2_bytes = 0;

    // ---- compat 1-D ----
    double *c_f = nullptr, *c_x = nullptr, *c_xx0 = nullptr, *c_newf = nullptr, *c_newx = nullptr,
           *c_newxx0 = nullptr, *c_omega = nullptr, *c_lrgVl = nullptr, *c_red = nullptr;
    u64 *c_seed = nullptr, *c_nevents = nullptr;
    int *c_stable = nullptr, *c_lrgEl = nullptr, *c_steps = nullptr;
    sq::Compat1DCtl *c_ctl = nullptr;          // device-side frame controller (f-3)
    sq::Compat1DFrameRec *c_log_rec = nullptr;
    double *c_log_xavg = nullptr;
    int64_t c_frames_done = 0;                 // frames run through the controller (log slot = frame % cap)

    // ---- lattice ----
    void *l_field[2] = {nullptr, nullptr};
    void *l_ghost[2] = {nullptr, nullptr};  // local halo buffers (slab mode without P2P)
    u64 *l_seeds[2] = {nullptr, nullptr};
    u64 *l_event = nullptr;
    RebaseEntry *l_rebase = nullptr;
    JumpEntry *l_slice_jump = nullptr, *l_strip_jump = nullptr;
    double *l_partials = nullptr, *l_slice_sum = nullptr, *l_slice_x = nullptr, *l_slice_xx0 = nullptr,
           *l_sums = nullptr, *l_sums_mean = nullptr, *l_m2 = nullptr, *l_lam = nullptr, *l_redbuf = nullptr;
    unsigned long long *l_nclamped = nullptr;
    unsigned long long *l_nclamp_step = nullptr;  // [MAX_SEQ_STEPS] clamp hits per step of the sequence in flight
    // observables off the critical path: the finalizes of a GROUP of fin_batch steps run on a side stream while the next
    // group's updates run; l_partials is a ring of 2 * fin_batch per-step buffers of npart doubles, events order producer
    // and consumer per group (sq_api.cu: sq_enqueue_step).  Inside a group nothing sits between two update kernels in
    // `stream`, so the tile kernel's programmatic dependent launch overlaps them.
    static constexpr int FIN_BATCH_MAX = 8;
    int fin_batch = 1;
    size_t npart = 0;
    sq::FinalizeArgs fin_queue[FIN_BATCH_MAX];
    int fin_queued = 0;
    cudaStream_t fin_stream = nullptr;
    // sq_frame_host: the field's device -> host copy is enqueued on its own stream right behind the frame's (first) batch of
    // update kernels -- no host round trip before it starts, and it overlaps the observable kernels; redone if an RNG event
    // made the batch stop early
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr;
    void *spec_want = nullptr;       // sq_frame_host's destination while its frame is pending
    void *spec_host = nullptr;       // armed for the batch that completes the frame
    const void *spec_src = nullptr;  // the device buffer that copy read
    unsigned batch_seq = 0, spec_batch = 0, ok_batch = 0;  // batches enqueued ; the one the copy followed ; the last one that ran to its end
    cudaEvent_t ev_upd[2] = {nullptr, nullptr}, ev_fin[2] = {nullptr, nullptr};
    int fin_pending = 0;  // finalize GROUPS of the current sequence handed to the side stream and not yet joined into `stream`
    int cur = 0;
    int nt = 0, ctas_per_slice = 1;
    int64_t vslice = 0, V = 0, vlocal = 0;
    size_t rsz = 4;
    bool per_chain_coupling = false;
    // row-marching kernel (sq_march.cu): geometry + jump tables, when the shape qualifies
    bool march_ok = false;
    bool rows_ok = false;   // ... through the row-block staging kernel (SQ_FLAG_ROWBLOCK_KERNEL / SQ_ROWS=1): lattice_rows_kernel
    bool tile_ok = false;   // ... and its tiles can be staged in shared memory: lattice_tile_kernel (sq_tile.cu)
    int m_R = 0, m_tpr_log = 0, m_w = 4;
    JumpEntry *l_cta_jump = nullptr, *l_thr_jump = nullptr;
    unsigned *l_tile_ctr = nullptr;    // [2] persistent tile kernel's claim counters
    sq::TileThread *l_tile_thr = nullptr, *l_rows_thr = nullptr;  // [256] tile kernel: a thread's place inside any tile
    // resident 2-D path (sq_resident.cu)
    bool res_ok = false;
    int res_nb = 0, res_rows = 0;
    unsigned long long *r_nclamp_slots = nullptr;  // [RES_SLOTS] clamp hits per checkpoint interval of a resident launch
    unsigned long long *r_halo = nullptr;
    unsigned *r_error = nullptr, *r_progress = nullptr;
    float *r_ckpt = nullptr;  // [RES_NCKPT][V] checkpoints of the resident kernel (RNG-event recovery)
    unsigned r_tag = 1;     // monotonic halo tag base (never reused, also across replays)
    double *r_hist_rows = nullptr, *r_hist_p2 = nullptr, *r_step_sums = nullptr;
    int res_limit = 0;      // >0: the next resident batch must stop after this many steps
    int force_stream = 0;   // >0: this many steps must go through the streaming kernel
    int pend_kind = 0;      // 0 streaming, 1 resident
    // pending sequence
    bool pending = false;
    double pend_dtau = 0;
    int pend_nsteps = 0;  // steps currently enqueued
    int pend_total = 0;   // steps the caller asked for
    int64_t pend_runs0 = 0;
    std::vector<RebaseEntry> entries;  // replay entries valid for the first step of the sequence
    std::vector<RebaseEntry> deferred; // on-chip kernel: the entry of the event step, resolved when the launch was abandoned,
                                       // installed when the re-run has reached that step
    // optional per-launch timing of the update kernel
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    double timing_ms = 0;
    int64_t timing_launches = 0;
    // multi-GPU slab ring (sq_slab.cu), null unless sq_slab_join succeeded
    sq::SlabState *slab = nullptr;
};

// ---- internals shared between sq_api.cu and sq_slab.cu -------------------------------------------
int sq_set_dev(sq_ctx *c);
int sq_timing_mark(sq_ctx *c);
int sq_timing_collect(sq_ctx *c, size_t valid);
sq::LatticeArgs sq_lattice_args(sq_ctx *c, double dtau, int k);
int sq_launch_update(sq_ctx *c, const sq::LatticeArgs &A);
// one step = update on `stream` + finalize on the side stream (k = step index within the sequence)
int sq_enqueue_step(sq_ctx *c, sq::LatticeArgs &A, sq::FinalizeArgs &F, int k);
int sq_join_finalize(sq_ctx *c);  // make `stream` wait for the side stream's outstanding finalizes  // generic or marching kernel
void sq_fill_rebase_inline(sq::LatticeArgs &A, const sq::RebaseEntry *e, int n);
// seed (full u64) before the draw at gid g of the step whose start seed is S, under `entries`
sq::u64 sq_host_seed_before(const sq_ctx *c, const std::vector<sq::RebaseEntry> &entries, int chain, sq::u64 S, sq::u64 g);
// sq_slab.cu
int sq_slab_enqueue(sq_ctx *c, double dtau, int nsteps, int64_t runs0);
int sq_slab_finish(sq_ctx *c);
void sq_slab_measure(sq_ctx *c, sq_obs *o);
void sq_slab_destroy(sq_ctx *c);
void sq_slab_stats_impl(sq_ctx *c, uint64_t *scans, uint64_t *rounds);
