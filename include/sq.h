/*
 * sq.h -- C-ABI of libsq, the B200-native (sm_100a) replacement for StochQuant's
 * device side: the OpenCL kernel `time_dev` (/root/reference/tau_kernel.cl:25-175,
 * helpers :184-284) and the OpenCL plumbing that drives it from the host
 * (/root/reference/tauhost.c:187-481 setup, :481-554 per-frame launch/transfers,
 * :587-612 teardown).
 *
 * The reference exposes no library API: its boundary is the `tauhost.o` process
 * (argv, stdout, start/end files).  `host/tauhost.c` in this repo keeps that
 * boundary byte-compatible and calls the four entry points BASELINE.json names:
 * sq_init / sq_step / sq_measure / sq_free.  Everything here is extern "C",
 * plain pointers and sizes; the caller owns host memory, the library owns device
 * memory and its stream.  All functions return SQ_OK (0) or a negative error
 * code; nothing throws or aborts across the boundary.  A context is not
 * thread-safe; use one per host thread.  There is NO CPU fallback: without a
 * usable CUDA device sq_init fails with SQ_ERR_NODEVICE.
 */
#ifndef SQ_H
#define SQ_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQ_API_VERSION 3

/* error codes */
#define SQ_OK 0
#define SQ_ERR_INVALID (-1)     /* bad argument / unsupported combination        */
#define SQ_ERR_CUDA (-2)        /* CUDA runtime error (see sq_last_cuda_error)   */
#define SQ_ERR_NOMEM (-3)
#define SQ_ERR_UNSUPPORTED (-4) /* e.g. potential ids 1,2: no kernel in reference */
#define SQ_ERR_NODEVICE (-5)
#define SQ_ERR_TIMEOUT (-6)     /* bounded wait expired (halo flags, session barrier) */
#define SQ_ERR_INTERNAL (-7)    /* an invariant of the library was violated          */

/* which kernel family */
enum {
    /* reference-faithful 1-D kernel: fp64, potID 0|3, ghost-cell boundary, omega
     * random walk, stability heuristic, per-site running means -- every row of
     * SURVEY.md 8(a).  One CTA, state on chip for the whole frame. */
    SQ_KERNEL_COMPAT1D = 0,
    /* d-dimensional periodic lattice generalisation of SURVEY.md 8(d)
     * (no reference code): fp32/fp64, potID 0 (W=2) or 4 (phi^4). */
    SQ_KERNEL_LATTICE = 1
};
enum { SQ_REAL_F32 = 0, SQ_REAL_F64 = 1 };
/* transcendental flavour of the Box-Muller step (tau_kernel.cl:277).  The
 * integer stream (t1,t2,seed) is bit-exact in both; ACCURATE uses CUDA's
 * <=2-ulp cosf/logf/sqrtf with the reference's float/double casts, FAST uses
 * the SFU (MUFU) approximations: |dr| <~ 2e-6 on the N(0,1) draw. */
enum { SQ_MATH_ACCURATE = 0, SQ_MATH_FAST = 1 };

#define SQ_POT_HARMONIC 0   /* tau_kernel.cl:201-212 */
#define SQ_POT_DOUBLEWELL 3 /* tau_kernel.cl:184-200 (COMPAT1D only) */
#define SQ_POT_PHI4 4       /* F = m2*phi + lambda*phi^3 (LATTICE only, not in reference) */

typedef struct sq_ctx sq_ctx;

/* Replaces the argv-derived constants uploaded at tauhost.c:361-377 */
typedef struct sq_params {
    uint32_t struct_size; /* = sizeof(sq_params) */
    int32_t kernel;       /* SQ_KERNEL_*                                          */
    int32_t real;         /* SQ_REAL_* (COMPAT1D: must be F64)                    */
    int32_t math;         /* SQ_MATH_*                                            */
    int32_t potential;    /* `potential` kernel arg, tau_kernel.cl:41             */
    int32_t ndim;         /* 1..4 (COMPAT1D: 1)                                   */
    int64_t dims[4];      /* dims[0] fastest; dims[ndim-1] = Euclidean time.
                             COMPAT1D: dims[0] = LIST_SIZE (tau_kernel.cl:38)     */
    double spacing;       /* `deltaT`, tau_kernel.cl:39                           */
    double noise_c;       /* `C`, tau_kernel.cl:42                                */
    double m2, lambda;    /* SQ_POT_PHI4 only                                     */
    int32_t device;       /* CUDA ordinal; tauhost maps argv[7] (an OpenCL platform
                             index, tauhost.c:205) with dev % deviceCount         */
    int32_t nchains;      /* LATTICE: independent chains in one context (>=1),
                             chain k starts from seed+k unless sq_set_chain is used */
    /* slab decomposition along the time axis (LATTICE): this context owns global
     * time slices [slab_t0, slab_t0+slab_nt).  0,0 = the whole lattice.          */
    int64_t slab_t0, slab_nt;
    int32_t reserved;         /* 0.  (Was steps_per_launch in API 2, never read: the library chooses the fusion itself --
                                 a whole frame per launch on chip for 2-D lattices that fit, one step per launch for the
                                 streaming kernels, which are instruction-bound, not HBM-bound: DESIGN.md section 8.) */
    int32_t flags;            /* SQ_FLAG_*                                        */
} sq_params;

#define SQ_FLAG_NO_OBSERVABLES 1 /* LATTICE: skip the per-step reductions          */
#define SQ_FLAG_FORCE_STREAMING 2 /* LATTICE: never use the on-chip resident kernel */
#define SQ_FLAG_GENERIC_KERNEL 4  /* LATTICE: never use the row-marching fp32 kernel    */
#define SQ_FLAG_ROWBLOCK_KERNEL 8 /* LATTICE d>=3 fp32: the row-block staging kernel instead of the tile kernel */

/* What sq_measure copies out.  Pointer members are caller-allocated (or NULL to
 * skip).  Replaces the per-frame blocking reads tauhost.c:504-515. */
typedef struct sq_obs {
    uint32_t struct_size;
    /* COMPAT1D: committed f, x, xx0 -- N doubles each (tauhost.c:508-513)        */
    double *f, *x, *xx0;
    double omega;     /* tauhost.c:514 */
    uint64_t seed;    /* device RNG seed `rand1` (never read back by the reference) */
    int32_t lrgEl;    /* tau_kernel.cl:35 */
    int32_t stable;   /* result of the last sq_step */
    double lrgVl;     /* tau_kernel.cl:36 */
    int64_t runs;     /* tau-steps accumulated into the running means             */
    /* reductions of the current configuration (COMPAT1D: of the path f+cl)       */
    double mean_phi, mean_phi2;
    /* LATTICE: per-time-slice running means (Lt doubles each, chain 0, this slab's
     * slices) of Phi(t) and Phi(t)Phi(t_mid); corr[t] = xx0[t]-x[t]*x[t_mid]
     * (the host's xavg, tauhost.c:519-521)                                       */
    double *slice_x, *slice_xx0, *corr;
    int64_t nclamped; /* sites that hit the +-1000 clamp (tau_kernel.cl:122-132)  */
    uint64_t nevents; /* RNG chain events replayed (inf-retry / `seed+=`)         */
    int64_t steps_done; /* tau-steps executed by the last sq_step                 */
} sq_obs;

/* ---- the four entry points named by BASELINE.json ---------------------------- */

/* Replaces tauhost.c:196-435 (platform/device/context/queue, 19 buffers, 19 uploads,
 * JIT build, 19 clSetKernelArg).  f0/x0/xx0_0: initial state, N (=volume) doubles
 * each or NULL for zeros (LATTICE ignores x0/xx0_0; use sq_upload_field for native
 * fp32 data).  omega0: tauhost.c:84-89.  seed: `rand1`, tauhost.c:185. */
int sq_init(sq_ctx **out, const sq_params *p, const double *f0, const double *x0,
            const double *xx0_0, double omega0, uint64_t seed);

/* One reference frame: replaces clEnqueueNDRangeKernel+clFinish (tauhost.c:481-483),
 * the `stable` read (:504), and the commit / rollback of :506-554.  Runs `nsteps`
 * (= `Loops`) tau-steps at step size dtau with the running means' counter starting
 * at runs0 (`runs`, tauhost.c:554).  *stable=1: state committed.  *stable=0
 * (COMPAT1D): the frame is rolled back to the pre-frame f/x/xx0/omega while RNG
 * seed, lrgEl, lrgVl keep their new values -- exactly what the reference host
 * does by not reading back.  Synchronous. */
int sq_step(sq_ctx *ctx, double dtau, int nsteps, int64_t runs0, int *stable);

/* Replaces the blocking read-backs tauhost.c:508-515 (+ reductions). */
int sq_measure(sq_ctx *ctx, sq_obs *out);

/* Replaces tauhost.c:587-612. */
void sq_free(sq_ctx *ctx);

/* ---- frame controller on the device (COMPAT1D) ---------------------------
 * The reference host decides after EVERY frame whether to keep it and how to change the step size
 * (tauhost.c:504-545): one blocking read of `stable`, four read-backs, five re-uploads per frame.
 * Here that logic runs in the frame kernel's epilogue, so a batch of frames needs no host round trip:
 *   stable:   commit; every 11th consecutive stable frame dtau /= 0.95 (:523-528); runs += nsteps
 *   unstable: roll back (seed, lrgEl, lrgVl keep their new values, :533-554); dtau *= 0.95
 * and each frame logs what the host prints for it.  Results are bit-identical to driving sq_step
 * frame by frame with the same rules (tests/test_gpu_compat1d.py::test_device_controller_*). */
typedef struct sq_frame_rec {
    double dtau;      /* step size the frame ran with (the value its stdout line shows, :495)  */
    int32_t stable;   /* 1: committed, 0: rolled back                                          */
    int32_t steps;    /* tau-steps executed (an unstable frame stops early, tau_kernel.cl:169) */
} sq_frame_rec;
#define SQ_FRAMES_MAX 64 /* frames per sq_frames call */
/* state of the controller: dtau of the next frame, `runs` (tauhost.c:554), consecutive stable frames */
int sq_controller_set(sq_ctx *ctx, double dtau, int64_t runs, int stab_cnt);
int sq_controller_get(sq_ctx *ctx, double *dtau, int64_t *runs, int *stab_cnt);
/* Run nframes (<= SQ_FRAMES_MAX) frames of nsteps tau-steps back to back.  recs: [nframes].
 * xavg (NULL to skip): [nframes][N], row k = xx0[i]-x[i]*x[mid] after frame k if it was stable
 * (tauhost.c:519-521), else untouched.  Synchronous. */
int sq_frames(sq_ctx *ctx, int nframes, int nsteps, sq_frame_rec *recs, double *xavg);

/* ---- exact resume (COMPAT1D) ----------------------------------------------
 * The reference's end file (tauhost.c:562-581) does not hold everything a run needs to continue
 * bit-exactly: the device seed `rand1`, lrgEl / lrgVl (which survive frames and rollbacks,
 * tauhost.c:533-554), omega (written but ignored on read, :122-124) and the one stale value
 * newf[lrgEl] the stability scan compares against after a rejected frame (tau_kernel.cl:135) are
 * lost, so a restarted reference run re-randomises them.  These two entry points let the host carry
 * them in an extended trailer (host/tauhost_io.h) behind the reference's three trailer lines. */
typedef struct sq_compat_state {
    uint32_t struct_size;
    int32_t lrgEl;
    uint64_t seed;
    double lrgVl, omega, newf_lrgEl;
} sq_compat_state;
int sq_compat_get_state(sq_ctx *ctx, sq_compat_state *out);
int sq_compat_set_state(sq_ctx *ctx, const sq_compat_state *in);

/* ---- support ----------------------------------------------------------------- */
const char *sq_strerror(int code);
const char *sq_last_cuda_error(void); /* text of the last CUDA failure, this thread */
int sq_api_version(void);
int sq_device_count(void); /* <0 on error, 0 if none */

/* Asynchronous flavour of sq_step for callers that time on the device: enqueues
 * the launches on the context's stream and returns.  sq_sync finishes the frame
 * (event replay, commit) and reports stable. */
int sq_step_async(sq_ctx *ctx, double dtau, int nsteps, int64_t runs0);
int sq_sync(sq_ctx *ctx, int *stable);
void *sq_stream(sq_ctx *ctx); /* the cudaStream_t the kernels are launched on */
int64_t sq_launch_count(sq_ctx *ctx); /* kernels launched by this context so far */

/* Per-launch device timing of the dominant (update) kernel: when enabled, every update-kernel
 * launch is bracketed by CUDA events on the context's stream; sq_kernel_time returns the summed
 * elapsed milliseconds and the number of launches since it was enabled.  Off by default (the
 * event records sit between launches). */
int sq_kernel_timing(sq_ctx *ctx, int enable);
int sq_kernel_time(sq_ctx *ctx, double *ms_total, int64_t *launches);

/* LATTICE field transfer in the native or another real type (SQ_REAL_*).
 * chain: which chain of the batch.  Layout: lexicographic, dims[0] fastest, this
 * slab's slices only. */
int sq_upload_field(sq_ctx *ctx, int chain, const void *host, int real);
int sq_download_field(sq_ctx *ctx, int chain, void *host, int real);
/* per-chain seed and phi^4 couplings for batched ensembles */
int sq_set_chain(sq_ctx *ctx, int chain, uint64_t seed, double m2, double lambda);
/* per-chain observables of a batch: [nchains] each (NULL to skip) */
int sq_measure_chains(sq_ctx *ctx, double *mean_phi, double *mean_phi2, uint64_t *seeds);

/* End-to-end frame through host buffers (the reference's per-frame traffic,
 * tauhost.c:550-554 uploads + :508-515 read-backs): H2D of `host_in` (volume reals),
 * nsteps tau-steps, D2H of the field into `host_out` and of the observables.
 * Synchronous: `host_out` is complete on return (internally the read-back of a small
 * field follows the last update kernel on a stream of its own; pin the buffers). */
int sq_frame_host(sq_ctx *ctx, const void *host_in, void *host_out, int real, double dtau,
                  int nsteps, int64_t runs0, sq_obs *obs, int *stable);

/* Debug / parity hook: the (t1,t2) 48-bit LCG outputs every site of the *next*
 * tau-step would use (tau_kernel.cl:273,275), computed on the device with the
 * same jump-ahead code path as the update kernels.  n = number of sites from gid0. */
int sq_debug_draws(sq_ctx *ctx, int chain, uint64_t gid0, uint64_t n, uint64_t *t1, uint64_t *t2);

/* Host-side utility (no GPU needed): seed before the draw at gid0+ndraws given the seed
 * before the draw at gid0, assuming no retry / `seed+=` event in between (low 48 bits).
 * Lets a caller position the shared chain (tau_kernel.cl:269-284) for a slab or a resume. */
uint64_t sq_lcg_jump(uint64_t seed, uint64_t gid0, uint64_t ndraws);

/* ---- multi-GPU slab decomposition (north_star (4)) --------------------
 * The reference is single-device (tauhost.c:249-252: one context, one in-order queue); this is
 * new work behind the same four entry points.  One lattice is cut along its time axis into slabs,
 * one context per GPU (sq_params.slab_t0/slab_nt), one process or host thread per context, all on
 * ONE box.  The ranks meet in a session (POSIX shared memory: barrier + small all-gathers; no GPU
 * needed); sq_slab_join then maps the ring neighbours' halo arenas (CUDA IPC, or plain peer
 * pointers inside one process) and from then on
 *   - the boundary slices travel GPU-to-GPU over NVLink inside the update kernel itself
 *     (boundary slices first, posted peer stores, arrival flags; no host in the loop);
 *   - every rank draws from the ONE shared-seed chain of tau_kernel.cl:269-284 at its global
 *     gids; the chain's rare data-dependent events (inf-retry :282, `seed+=` :278-279) are
 *     found ahead of the update by an integer-only scan of each slab and agreed on through the
 *     session, so all ranks apply the same corrections and the stream stays bit-exact;
 *   - the running means of Phi(t)Phi(t_mid) use the mid slice's owner's per-step sums.
 * With a joined context sq_step / sq_step_async / sq_sync / sq_measure are COLLECTIVE: every
 * rank of the ring must make the same calls with the same dtau / nsteps / runs0.
 * One rank per GPU: several ranks of a ring on the SAME device work for small lattices (tests) but
 * can starve each other once a rank's boundary CTAs alone fill the device. */
typedef struct sq_session sq_session;
/* name: unique per ring (no '/'); the segment is unlinked once every rank has attached */
int sq_session_open(sq_session **out, const char *name, int rank, int nranks);
int sq_session_barrier(sq_session *s);
/* n <= 8 words per rank; out: [nranks][n] */
int sq_session_allgather_u64(sq_session *s, const uint64_t *in, int n, uint64_t *out);
/* n <= 3072 doubles per rank; out: [nranks][n] */
int sq_session_allgather_f64(sq_session *s, const double *in, int n, double *out);
void sq_session_abort(sq_session *s); /* wake every waiter with SQ_ERR_TIMEOUT */
int sq_session_rank(const sq_session *s);
int sq_session_size(const sq_session *s);
void sq_session_close(sq_session *s);

/* Collective.  The contexts' slabs must tile [0, Lt) in rank order; seeds must agree.  The
 * session must outlive the context (close it after sq_free). */
int sq_slab_join(sq_ctx *ctx, sq_session *s);
/* statistics of a joined context: finder kernel launches and agreement rounds so far */
int sq_slab_stats(sq_ctx *ctx, uint64_t *finder_scans, uint64_t *agree_rounds);

/* One resolved event of the shared-seed chain within a tau-step: draws at gid >= gid_start
 * continue from `seed`; the draw at ov_gid used (ov_t1, ov_t2). */
typedef struct sq_rng_entry {
    uint64_t gid_start, seed, ov_gid, ov_t1, ov_t2;
} sq_rng_entry;
/* Host-side utility (no GPU needed): replay the draw at `gid` literally (tau_kernel.cl:269-284,
 * including the do/while retry) given the step-start seed and the step's earlier entries
 * (ascending gid).  Every rank of a ring calls this with the same arguments and gets the same
 * entry.  *ndraws = LCG double-draws consumed (1 = no retry), *plus = the `seed+=` branch taken. */
int sq_rng_resolve(uint64_t step_seed, const sq_rng_entry *entries, int n, uint64_t gid,
                   sq_rng_entry *out, int *ndraws, int *plus);

#ifdef __cplusplus
}
#endif
#endif /* SQ_H */
