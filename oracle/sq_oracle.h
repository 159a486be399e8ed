/*
 * sq_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the StochQuant hot path:
 *   /root/reference/tau_kernel.cl:25-175   __kernel time_dev
 *   /root/reference/tau_kernel.cl:184-284  clas/ddPot/intConst/boundary/absol/random
 *   /root/reference/tauhost.c:84-102,185   initial state (unseeded glibc rand())
 *   /root/reference/tauhost.c:479-560      frame loop + dtau controller
 *   /root/reference/tauhost.c:485-501      stdout frame line
 *   /root/reference/tauhost.c:103-173,562-581  start / end file
 * plus the d-dimensional generalisation of SURVEY.md section 8(d), which has no
 * reference code and is *defined* here (d=1 fp64 reduces to the reference update
 * up to floating-point association).
 *
 * PARITY PINNING: the reference ships no golden vectors and cannot be built with
 * its own toolchain here (no CL/cl.h, no OpenCL runtime).  The restatement is
 * pinned instead against the reference's *kernel source itself*, compiled as C
 * through oracle/ref_shim (gcc, serial work-item schedule) into oracle/_ref/ --
 * see oracle/Makefile and tests/test_oracle_vs_ref.py.  Where oracle/_ref is
 * absent the KATs in tests/golden/ (generated with it) stand in.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may use anything under oracle/.  The product (libsq.so,
 * tauhost.o) never links or calls it.
 */
#ifndef SQ_ORACLE_H
#define SQ_ORACLE_H
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* execution-order semantics for the racy reference */
enum { SQO_RNG_CHAIN = 0, SQO_RNG_SHARED = 1 };
enum { SQO_FIELD_JACOBI = 0, SQO_FIELD_INPLACE = 1 };

typedef struct sqo_draw {
    uint64_t t1, t2;      /* the two 48-bit LCG outputs finally used        */
    uint64_t seed_after;  /* *seed after the call (full u64, wraps)          */
    int      ndraws;      /* 1 + number of inf-retries                       */
    int      plus_branch; /* 1 if the last iteration took the `*seed+=temp` */
} sqo_draw;

/* tau_kernel.cl:269-284 -- literal. rec may be NULL. */
double sqo_random(uint64_t *seed, uint64_t gid, sqo_draw *rec);

/* tau_kernel.cl:184-267 -- literal model functions */
double sqo_clas(double a, double w, int pot);
double sqo_ddPot(double a, int pot);
double sqo_intConst(int pot);
double sqo_boundary(int rl, int pot);
double sqo_absol(double a);

/* ---- 1-D reference kernel state: mirrors the 19 kernel arguments -------- */
typedef struct sqo_state {
    int     N;            /* LIST_SIZE                                       */
    double  deltaT;
    double  deltaTau;
    double  C;
    int     potential;
    double *f, *x, *xx0, *newf, *newx, *newxx0;  /* caller-owned, N each     */
    double  omega;
    uint64_t rand1;
    int     stable;
    int     lrgEl;
    double  lrgVl;
    int     runs;
} sqo_state;

/* One launch of time_dev with Loops=loops (tau_kernel.cl:64-173).
 * Returns the number of tau-steps actually executed (break on unstable). */
int sqo_time_dev(sqo_state *s, int loops, int rng_mode, int field_mode);

/* Optional per-draw trace of the last executed step (N+1 entries), for tests. */
void sqo_set_trace(sqo_draw *buf, int capacity);

/* ---- host restatement ---------------------------------------------------- */
/* tauhost.c:84-102,185: omega0, cold-start f[], rand1 from unseeded rand().
 * Calls srand(1) first so that the result equals a fresh process. */
void sqo_host_init(int N, double deltat, double deltatau, int cold_start,
                   double *f, double *omega, uint64_t *rand1);

/* tauhost.c:485-501: one stdout frame line (returns bytes written to fp). */
int sqo_print_frame(FILE *fp, int N, const double *xavg, double dtau,
                    int j, int frames);

/* tauhost.c:562-581 */
int sqo_write_endfile(const char *path, int N, int acc, const double *xavg,
                      const double *xx0, const double *x, const double *f,
                      double omega, int runs_plus_rec, double dtau);

/* tauhost.c:103-173 (returns 0 ok, 1 cannot open) */
int sqo_read_startfile(const char *path, int N, double deltatau,
                       double *xavg, double *xx0, double *x, double *f,
                       int *recSimlgth, double *dtautmp);

/* tauhost.c:29-621 with the OpenCL launch replaced by sqo_time_dev:
 * full reference program on the CPU.  argv as tauhost.c:31-43.
 * out = stdout stream.  Returns exit code. */
int sqo_tauhost_main(int argc, char **argv, FILE *out, int rng_mode, int field_mode);

int sqo_print_frame_path(const char *path, int N, const double *xavg, double dtau, int j, int frames);
int sqo_tauhost_main_path(int argc, char **argv, const char *outpath, int rng_mode, int field_mode);

/* ---- d-dimensional generalisation ---------------------- */
enum { SQO_F32 = 0, SQO_F64 = 1 };

typedef struct sqo_lattice {
    int      ndim;         /* 1..4                                           */
    int64_t  dims[4];      /* dims[0] fastest; dims[ndim-1] = Euclidean time */
    int      real;         /* SQO_F32 / SQO_F64 storage + arithmetic         */
    int      potential;    /* 0: F=2 phi ; 4: F = phi*(m2 + lambda phi^2)    */
    double   a;            /* lattice spacing (deltat)                       */
    double   C;
    double   m2, lambda;
    void    *phi;          /* V reals, caller-owned                          */
    void    *phi_new;      /* V reals scratch, caller-owned                  */
    uint64_t seed;         /* shared chain seed (full u64)                   */
    int64_t  runs;         /* steps accumulated into the running means       */
    double  *slice_x;      /* [Lt] running mean of Phi(t)                    */
    double  *slice_xx0;    /* [Lt] running mean of Phi(t) Phi(t_mid)         */
    double  *slice_sum;    /* [Lt] last step's slice sums of phi (pre-update)*/
    double   sum_phi, sum_phi2; /* last step's global sums (pre-update)      */
    int64_t  nclamped;     /* sites clamped so far                           */
    uint64_t nevents;      /* retry / += events seen so far                  */
} sqo_lattice;

/* One Langevin step, serial chain in gid order (the definition). */
void sqo_lattice_step(sqo_lattice *L, double dtau);
/* Same result computed with OpenMP + affine jump-ahead (CPU baseline). */
void sqo_lattice_step_omp(sqo_lattice *L, double dtau);
int sqo_set_threads(int n); /* n > 0: set the OpenMP thread count; returns the count in force */

/* chain jump-ahead (independent of the product's implementation):
 * seed before the draw at gid g0+D given seed s before the draw at gid g0,
 * assuming no retry / += event in between (masked to 48 bits). */
uint64_t sqo_jump(uint64_t s, uint64_t g0, uint64_t D);

/* dump (t1,t2) used by each site of a hypothetical step: tests of the stream */
void sqo_lattice_draws(uint64_t seed, uint64_t V, uint64_t *t1, uint64_t *t2,
                       uint64_t *seed_after);

#ifdef __cplusplus
}
#endif
#endif
