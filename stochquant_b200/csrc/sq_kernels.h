// sq_kernels.h -- argument blocks and launchers shared by the kernels and sq_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sq_lcg.cuh"

namespace sq {

constexpr u64 NO_EVENT = ~0ULL;
// event key = step_in_sequence (16 bits) | chain (14 bits) | gid (34 bits); the minimum over
// all detections of a launch sequence is the first genuine event (everything before it was
// drawn from correct seeds).
constexpr int KEY_STEP_SHIFT = 48, KEY_CHAIN_SHIFT = 34;
__host__ __device__ inline u64 event_key(int step, int chain, u64 gid) {
    return ((u64)step << KEY_STEP_SHIFT) | ((u64)chain << KEY_CHAIN_SHIFT) | gid;
}

// ---------------------------------------------------------------- compat 1-D ----
// Device-side frame controller (the host logic of tauhost.c:504-545 moved next to
// the kernel): when Compat1DArgs::ctl is set, the frame takes dtau / runs from this block, and its
// epilogue applies the reference's accept / reject + step-size adaptation and logs what the host prints.
struct Compat1DCtl {
    double dtau;        // step size of the NEXT frame
    long long runs;     // tau-steps accumulated into the running means (tauhost.c:554 `runs`)
    int stab_cnt;       // consecutive stable frames (tauhost.c:523-528)
    int frame;          // frames executed through the controller so far
};
struct Compat1DFrameRec {
    double dtau;        // step size the frame ran with (the value its stdout line shows)
    int stable, steps;  // frame accepted? ; tau-steps executed
};
struct Compat1DArgs {
    int N, loops, potential;
    long long runs;
    double dt, dtau;
    double dt2;           // (double)pown((float)deltat,2), tau_kernel.cl:114
    double nscale_site;   // c*(double)sqrt((float)(2.*deltatau/deltat)), :112
    double nscale_omega;  // c*(double)sqrt((float)(2.*deltatau)), :105
    double intconst;      // intConst(potID), :237-246 (float expression, host-evaluated)
    u64 P, Q;             // whole-step affine seed advance over N+1 draws
    const JumpEntry *jump;
    // device state: the reference's buffers (committed f/x/xx0/omega + persistent new*)
    double *f, *x, *xx0, *newf, *newx, *newxx0, *omega;
    u64 *seed;
    int *stable, *lrgEl, *steps_done;
    double *lrgVl;
    u64 *nevents;
    // controller mode (null: the host passes dtau / runs and reads `stable` back every frame)
    Compat1DCtl *ctl;
    Compat1DFrameRec *log_rec;  // [log_cap]
    double *log_xavg;           // [log_cap][N]  xx0[i] - x[i] x[mid] after an accepted frame (tauhost.c:519-521)
    int log_cap;
    double noise_c;             // `C`: the noise scales depend on dtau and are derived in the kernel
};
cudaError_t launch_compat1d(const Compat1DArgs &A, cudaStream_t stream);

// ---------------------------------------------------------------- lattice --------
// one replayed RNG event of the current step (host-resolved, see sq_api.cu)
struct RebaseEntry {
    u64 gid_start;  // draws at gid >= gid_start chain from `seed`
    u64 seed;       // full-u64 seed before the draw at gid_start
    u64 ov_gid;     // site whose (t1,t2) are overridden (the event site), or NO_EVENT
    u64 ov_t1, ov_t2;
    int chain;
    int pad;
    // virtual step-start seed: the event-free chain from vseed at gid 0 reproduces every seed at
    // gid >= gid_start (alpha is odd, so  alpha^g vseed + c(0,g) = seed  can be solved for vseed).
    // Lets a kernel reuse its gid-0-based jump tables behind the entry.  Filled by sq_fill_vseed.
    u64 vseed;
};

// lattice_tile_kernel: where a thread's first strip lies inside ANY tile (tile geometry does not depend on the CTA), built on
// the host once per context.  64 bytes: the kernel fetches it with three 128-bit loads (+ `plane` for d = 4).
struct alignas(64) TileThread {
    u64 a, g0;                 // table jump over thr_off draws (JumpEntry a, g0, bg1) ...
    u64 bg1, ck_off;           // ... and BETA thr_off row_jump.g0, the thread's share of the row-advance constant
    unsigned thr_off, s_c;     // first site relative to the tile's first site (ty R L0 + tx w) ; byte offset of that strip ...
    unsigned s_left, s_right;  // ... and of its two x0 neighbours inside the staged tile (from the start of dynamic smem)
    unsigned plane, row;       // x2 planes / x1 rows between the tile's first row and the thread's first row
    unsigned pad[2];
};

constexpr int RB_INLINE = 4;
struct LatticeArgs {
    int ndim;            // 2..4
    int pot;             // 0 | 4
    int nt;              // local time slices
    int wrap_time;       // 1: this context owns the whole time extent (periodic wrap)
    int nchains;
    int step_index;      // position in the current launch sequence (event key)
    int n_rebase;        // entries valid for this step
    int strips_per_cta_iter;  // blockDim.x * gridDim.x
    int ctas_per_chunk;  // L2 blocking: CTAs sweep a chunk of this many CTAs' sites through ALL time slices
                         // before moving to the next chunk (chunk x 4 slices must fit in L2)
    long long dim[4];    // extents, dim[ndim-1] = global Lt
    long long vslice;    // sites per time slice
    long long V;         // global volume (= draws per step - 1)
    long long slab_t0;   // first global slice owned
    long long chain_stride;  // reals between chains (local volume)
    const void *in;      // [nchains][nt*vslice]
    void *out;
    const void *ghost_lo, *ghost_hi;  // [vslice] slices below / above (slab mode), else unused
    double c_lap, c_dt, nscale;       // m*dtau/a2f ; dtau ; C*sqrtf(2 dtau/a^d)
    double m2, lam;                   // used when m2_chain == nullptr
    const double *m2_chain, *lam_chain;
    const u64 *seed_in;  // [nchains] full-u64 seed at step start
    u64 *seed_out;       // [nchains] seed after the step's V+1 draws
    float k2_f;              // 2 ln2 nscale^2 (FAST noise amplitude folded under the square root)
    const JumpEntry *slice_jump;  // [nt]   jump from gid 0 to the start of local slice t
    const JumpEntry *strip_jump;  // [strips_per_cta_iter] jump over q*VEC draws (first strip of a thread)
    JumpEntry stride_jump;   // jump over strips_per_cta_iter*VEC draws
    JumpEntry vol_jump;      // jump over V draws from gid 0 (to the omega draw)
    const JumpEntry *jump;
    const RebaseEntry *rebase;
    // the entries' gid_start (ascending, padded with ~0) and chains by value: strips test them
    // from the constant bank; more than RB_INLINE entries take the generic path
    u64 rb_gid[4];
    int rb_chain[4];
    u64 *event_key;      // atomicMin target; != NO_EVENT also aborts later launches
    double *partials;    // [nchains][nt][ctas_per_slice][2] (sum phi, sum phi^2), or null
    unsigned long long *nclamped;
    // ---- row-marching kernel (sq_march.cu), fp32 d = 3,4 with dims[0]/4 a power of two <= 256 ----
    int m_on;                    // 1: launch lattice_march_kernel (gridDim.x = its own CTAs per slice), 2: lattice_tile_kernel, 3: lattice_rows_kernel
    int m_R;                     // consecutive rows (x1) per thread; divides dims[1]
    int m_w;                     // sites per strip: 4 (marching kernel, tile kernel) or 8 (tile kernel)
    int m_tpr_log;               // log2(threads per row) = log2(dims[0] / m_w)
    const JumpEntry *cta_jump;   // [ctas per slice] jump over bx * rows_per_cta * L0 draws
    const JumpEntry *thr_jump;   // [256] jump over (ty * R * L0 + tx * m_w) draws
    const TileThread *tile_thr;  // [256] tile kernel (null unless m_on == 2)
    unsigned *tile_ctr;          // [2] row-block kernel: next tile to claim, CTAs that have finished (both 0 between launches)
    const TileThread *rows_thr;  // [256] row-block kernel: thread (tx, ty) = row ty of a pass, sites 4 tx ..
    JumpEntry prow_jump;         // jump over 1024 draws (one pass down at the same place in the block)
    u64 p_dck, p_dc1, p_dc2;     // row-block kernel: per-pass increments of the affine constant and of the two site constants
    JumpEntry row_jump;          // jump over L0 draws (one row down at fixed x0)
    u64 t_dck, t_dc1, t_dc2;     // tile kernel: per-row increments of the affine constant and of the two site constants
    // ---- multi-GPU slab ring (sq_slab.cu); slab_on == 0: everything below is unused ----------
    // direction 0 = the slice below local slice 0, 1 = the slice above local slice nt-1.
    int slab_on;
    unsigned wait_tag;             // ghost_lo/ghost_hi are valid once *wait_flag[d] has reached this tag
    const unsigned *wait_flag[2];  // local arrival flags of this step's ghost buffers
    unsigned push_tag;             // tag of the field this step produces; 0: do not push (last step)
    void *push_ghost[2];           // [0]: lower neighbour's "above" ghost, [1]: upper neighbour's "below" ghost
    unsigned *push_flag[2];        // the neighbours' arrival flags for those buffers
    unsigned *push_count;          // local [2]: CTAs of the boundary slice that have finished
    unsigned *slab_error;          // bounded waits: raised instead of hanging the GPU
};
cudaError_t launch_lattice_step(const LatticeArgs &A, int real, int math, int ctas_per_slice,
                                cudaStream_t stream);
cudaError_t preload_lattice_step(int real, int math, int ndim);
cudaError_t launch_lattice_march(const LatticeArgs &A, int math, int ctas_per_slice, cudaStream_t stream);
// sq_tile.cu: the same tiles staged through shared memory by bulk asynchronous copies (LatticeArgs::m_on == 2)
cudaError_t launch_lattice_tile(const LatticeArgs &A, int math, int ctas_per_slice, cudaStream_t stream);
bool tile_shape_ok(int ndim, int L0, int L1, int tpr_log, int R, bool rows);
size_t tile_smem_bytes(int ndim, int L0, int L1, int tpr_log, int R, bool rows);

struct FinalizeArgs {
    int nt, nchains, ctas_per_slice;
    int tmid_local;      // local index of the global mid slice, or -1 if not owned
    long long vslice;
    long long runs;      // running-mean counter before this step
    const double *partials;
    double *slice_sum;   // [nchains][nt]   last step's slice sums
    double *slice_x;     // [nchains][nt]   running mean of Phi(t)
    double *slice_xx0;   // [nchains][nt]   running mean of Phi(t) Phi(t_mid)
    double *sums;        // [nchains][2]    last step's global sums (phi, phi^2)
    double *sums_mean;   // [nchains][2]    running means of <phi>, <phi^2>
    double *history;     // slab mode: this step's [nt] slice sums + (sum phi, sum phi^2), else null
    const u64 *event_key;
    int step_index;      // position of the step whose partials this reduces in the launch sequence
};
constexpr int FINALIZE_MAX_NT = 14000;  // 16 bytes of shared memory per local time slice
cudaError_t launch_finalize(const FinalizeArgs &A, cudaStream_t stream);
// Clamp hits are counted per step of a launch sequence (LatticeArgs::nclamped points at the step's slot):
// a step that is replayed after an RNG event must not be counted twice.  total += sum of the first
// nvalid slots; all ntotal slots are cleared for the next sequence.
cudaError_t launch_commit_clamps(unsigned long long *slots, int nvalid, int ntotal, unsigned long long *total,
                                 const u64 *event_key /* null, or: do nothing while an event is flagged */, cudaStream_t stream);

cudaError_t launch_debug_draws(u64 seed, u64 gid0, u64 n, const JumpEntry *jump, u64 *t1, u64 *t2,
                               cudaStream_t stream);
// ---------------------------------------------------------------- resident 2-D ---
struct ResidentArgs {
    int L0, L1, nsteps, pot;
    int step_index0;     // sequence index of the launch's first step (event key)
    unsigned step0;      // tag base: tags step0+1.. are unique within the context
    long long V;
    const float *in;
    float *out;
    unsigned long long *halo_ll;  // [2][nblocks][2][L0] words {float bits, step tag}
    double c_lap, c_dt, nscale, m2, lam;
    float c_lap_f, c_dt_f, c_2dt_f, m2_f, lam_f, k2_f;  // host-side casts of the above (+ 2 ln2 nscale^2)
    const u64 *seed_in;
    u64 *seed_out;
    u64 P, Q;            // whole-step affine seed advance over V+1 draws
    JumpEntry vol_jump;  // jump over V draws from gid 0
    const JumpEntry *jump;
    u64 *event_key;
    double *hist_rows;   // [nsteps][L1] slice sums of the pre-update field
    double *hist_p2;     // [nsteps][nblocks] partial sums of phi^2
    unsigned long long *nclamped;
    unsigned *error_flag;
    // RNG-event recovery: every RES_CKPT steps each CTA drops its band (the field BEFORE step n,
    // n a multiple of RES_CKPT) into ckpt[(n / RES_CKPT) % RES_NCKPT]; a CTA that sees the event word raised
    // stops and records how far it got.  The host resumes from the last checkpoint every CTA has
    // written instead of from the start of the launch (CTA skew is bounded by the halo dependency:
    // at most nblocks/2 = 74 steps, far less than (RES_NCKPT - 1) * RES_CKPT = 224).
    u64 row_const[8];     // k * L0 * A: the site constant gid*A+B of row k relative to the band's first row
    unsigned one;         // 1, from the host: a multiplier ptxas cannot fold (sq_site.cuh: site_const_next)
    float *ckpt;          // [RES_NCKPT][L1][L0]
    unsigned *progress;   // [nblocks] steps completed by each CTA when it left
    // ---- row-parallel kernel (sq_rowres.cu) ----
    int rows_max;                       // rows per CTA (ceil(L1 / nblocks))
    unsigned long long *nclamp_slots;   // [RES_SLOTS] clamp hits per interval of RES_CKPT steps (committed for the
                                        // intervals that stand: an abandoned launch must not count twice)
};
constexpr int RES_CKPT = 32;    // steps between checkpoints: an RNG event costs a re-run of 16 steps on average (16-step
                                // checkpoints were measured: -15 us per event, +20 us per event-free 1000-step frame)
constexpr int RES_NCKPT = 8;    // checkpoint ring
constexpr int RES_SLOTS = 64;   // RES_MAX_STEPS / RES_CKPT
int rowres_strip(int L0, int rows_max);  // sites per thread (8 | 4), 0: shape not eligible
cudaError_t launch_rowres(const ResidentArgs &A, int math, int nblocks, cudaStream_t st);

struct WelfordArgs {
    int nt, nsteps, tmid, np2;
    long long vslice, runs;
    const double *hist_rows, *hist_p2;
    double *slice_x, *slice_xx0, *slice_sum, *sums, *sums_mean;
    const u64 *event_key;
};
cudaError_t launch_welford_history(const WelfordArgs &A, double *step_sums /* [2*nsteps] scratch */, cudaStream_t stream);

// current-configuration reductions: partial sums [nchains][nblocks][2] (phi, phi^2), fixed order
constexpr int REDUCE_BLOCKS = 256;
cudaError_t launch_reduce_field(const void *field, int real, long long nper_chain, int nchains,
                                double *partials, cudaStream_t stream);
// compat 1-D: sums of the path f+cl and its square, and corr[i] = xx0[i]-x[i]*x[mid]
// out: [0]=sum path, [1]=sum path^2, [8..8+N) = corr
cudaError_t launch_compat_reduce(const double *f, const double *x, const double *xx0, const double *omega,
                                 int N, double dt, int pot, double *out, cudaStream_t stream);
cudaError_t launch_convert(const void *src, int src_real, void *dst, int dst_real, long long n,
                           cudaStream_t stream);

}  // namespace sq
