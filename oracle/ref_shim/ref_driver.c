/*
 * ref_driver.c -- ORACLE SUPPORT (test infrastructure, NOT product code).
 *
 * Minimal single-work-group executor for the reference kernel `time_dev`
 * compiled from /root/reference/tau_kernel.cl through cl_kernel_shim.h.
 * Work-items are coroutines (ucontext) run in gid order; barrier() yields to
 * the scheduler, so between two barriers every work-item runs to completion
 * before the next one starts -- the schedule a CPU OpenCL runtime (work-item
 * loop per work-group) produces.  With Loops==1 per launch this equals the
 * canonical {chain RNG, Jacobi field} semantics of SURVEY.md 8(a); with
 * Loops>1 it equals {chain, in-place}.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <ucontext.h>

long sq_ref_cur_gid = 0;
long sq_ref_gsize = 0;

/* the kernel, from the reference's own source file */
void time_dev(double *f, double *x, double *xx0, double *newf, double *newx, double *newxx0,
              double *omega, unsigned long *rand1, int *stable, const double *deltaTau,
              int *lrgEl, double *lrgVl, int *initRun, const int *LIST_SIZE,
              const double *deltaT, int *runs, const int *potential, const double *C,
              const int *Loops);

typedef struct {
    double *f, *x, *xx0, *newf, *newx, *newxx0, *omega;
    unsigned long *rand1;
    int *stable;
    const double *deltaTau;
    int *lrgEl;
    double *lrgVl;
    int *initRun;
    const int *LIST_SIZE;
    const double *deltaT;
    int *runs;
    const int *potential;
    const double *C;
    const int *Loops;
} kargs;

static ucontext_t sched_ctx;
static ucontext_t *item_ctx;
static int *item_done;
static kargs g_args;
static long g_running = -1;

void sq_ref_barrier(int flags)
{
    (void)flags;
    swapcontext(&item_ctx[g_running], &sched_ctx);
}

static void item_main(void)
{
    kargs *a = &g_args;
    time_dev(a->f, a->x, a->xx0, a->newf, a->newx, a->newxx0, a->omega, a->rand1, a->stable,
             a->deltaTau, a->lrgEl, a->lrgVl, a->initRun, a->LIST_SIZE, a->deltaT, a->runs,
             a->potential, a->C, a->Loops);
    item_done[g_running] = 1;
    swapcontext(&item_ctx[g_running], &sched_ctx);
}

/* clEnqueueNDRangeKernel(global = N+1, one work-group) + clFinish */
void sq_ref_launch(double *f, double *x, double *xx0, double *newf, double *newx,
                   double *newxx0, double *omega, unsigned long *rand1, int *stable,
                   double *deltaTau, int *lrgEl, double *lrgVl, int *initRun, int *LIST_SIZE,
                   double *deltaT, int *runs, int *potential, double *C, int *Loops)
{
    const long G = (long)*LIST_SIZE + 1;
    const size_t STK = 256 * 1024;
    kargs a = {f, x, xx0, newf, newx, newxx0, omega, rand1, stable, deltaTau, lrgEl, lrgVl,
               initRun, LIST_SIZE, deltaT, runs, potential, C, Loops};
    g_args = a;
    sq_ref_gsize = G;
    item_ctx = (ucontext_t *)calloc((size_t)G, sizeof(ucontext_t));
    item_done = (int *)calloc((size_t)G, sizeof(int));
    char *stacks = (char *)malloc(STK * (size_t)G);
    for (long i = 0; i < G; i++) {
        getcontext(&item_ctx[i]);
        item_ctx[i].uc_stack.ss_sp = stacks + STK * (size_t)i;
        item_ctx[i].uc_stack.ss_size = STK;
        item_ctx[i].uc_link = &sched_ctx;
        makecontext(&item_ctx[i], item_main, 0);
    }
    long remaining = G;
    while (remaining > 0) { /* one pass = one barrier phase */
        for (long i = 0; i < G; i++) {
            if (item_done[i] == 2) continue;
            g_running = i;
            sq_ref_cur_gid = i;
            swapcontext(&sched_ctx, &item_ctx[i]);
            if (item_done[i] == 1) {
                item_done[i] = 2;
                remaining--;
            }
        }
    }
    free(stacks);
    free(item_done);
    free(item_ctx);
}
