/*
 * tauhost.c -- drop-in replacement for the reference's host program
 * (/root/reference/tauhost.c): same positional argv (:31-43), same stdout frame stream
 * (:485-501), same start/end files (:103-173, :562-581), same exit codes and stderr
 * messages -- so /root/reference/taumain.py (:132 Popen) drives it unchanged.
 *
 * What changed underneath: the OpenCL context/queue/buffer/kernel-launch code
 * (:187-481, :504-554, :587-612) is gone; the frame loop calls the C-ABI of libsq
 * (include/sq.h): sq_init / sq_step(_async) / sq_measure / sq_free.  No tau_kernel.cl is
 * read from the working directory.  There is no CPU fallback: without a CUDA device the
 * program prints the libsq error and exits non-zero.
 *
 * Build: make tauhost.o   (an executable despite its name, as in the reference README)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sq.h"
#include "tauhost_io.h"

static int fail_sq(const char *what, int rc)
{
    fprintf(stderr, "%s: %s (%s)\n", what, sq_strerror(rc), sq_last_cuda_error());
    return 3;
}

int main(int argc, char **argv)
{
    th_args a;
    char msg[600];
    if (th_parse_args(argc, argv, &a, msg, sizeof msg) != 0) {
        fputs(msg, stderr);
        return 2;
    }
    if (a.pot_id != SQ_POT_HARMONIC && a.pot_id != SQ_POT_DOUBLEWELL) {
        /* taumain_windows.py:111-129 lists potID 1 and 2; tau_kernel.cl has no code for them */
        fprintf(stderr, "potID %d is not supported: the reference kernel only defines potID 0 and 3.\n", a.pot_id);
        return 1;
    }
    const int n = a.n;
    const int mid = n / 2;
    double *f = (double *)calloc((size_t)n, sizeof(double));
    double *x = (double *)calloc((size_t)n, sizeof(double));
    double *xx0 = (double *)calloc((size_t)n, sizeof(double));
    double *xavg = (double *)calloc((size_t)n, sizeof(double));
    if (!f || !x || !xx0 || !xavg) return 3;

    const int cold = strcmp(a.start_file, "0") == 0;
    double omega = 0, dtau = a.deltatau;
    unsigned long rand1 = 0;
    int rec_sim_length = 0;
    th_initial_state(n, a.deltat, a.deltatau, cold, f, &omega, &rand1);
    if (!cold && th_read_start_file(a.start_file, n, a.deltatau, xavg, xx0, x, f, &rec_sim_length, &dtau)) {
        fprintf(stderr, "Failed to read Input.\n");
        return 1;
    }

    const int ndev = sq_device_count();
    if (ndev <= 0) return fail_sq("tauhost", SQ_ERR_NODEVICE);
    sq_params p;
    memset(&p, 0, sizeof p);
    p.struct_size = sizeof p;
    p.kernel = SQ_KERNEL_COMPAT1D;
    p.real = SQ_REAL_F64;
    p.math = SQ_MATH_ACCURATE;
    p.potential = a.pot_id;
    p.ndim = 1;
    p.dims[0] = n;
    p.spacing = a.deltat;
    p.noise_c = a.c;
    /* argv[7] is an OpenCL *platform* index in the reference (tauhost.c:205; taumain.py passes 2):
     * fold it onto the CUDA ordinals that exist */
    p.device = ((a.dev % ndev) + ndev) % ndev;
    p.nchains = 1;
    sq_ctx *ctx = NULL;
    int rc = sq_init(&ctx, &p, f, x, xx0, omega, (uint64_t)rand1);
    if (rc != SQ_OK) return fail_sq("sq_init", rc);

    sq_obs obs;
    memset(&obs, 0, sizeof obs);
    obs.struct_size = sizeof obs;
    obs.f = f;
    obs.x = x;
    obs.xx0 = xx0;

    int runs = rec_sim_length, stab_cnt = 0;
    for (int j = 0; j < a.frames; ++j) {
        /* launch the frame, then format the previous frame's line while the GPU works:
         * the reference prints after clFinish but before its read-backs (:483-501), i.e. the
         * same (one frame old) xavg */
        if ((rc = sq_step_async(ctx, dtau, a.loops, runs)) != SQ_OK) return fail_sq("sq_step", rc);
        if (j % a.fps == 0) th_print_frame(stdout, n, xavg, dtau, j, a.frames);
        int stable = 1;
        if ((rc = sq_sync(ctx, &stable)) != SQ_OK) return fail_sq("sq_sync", rc);
        if (stable == 1) {
            if ((rc = sq_measure(ctx, &obs)) != SQ_OK) return fail_sq("sq_measure", rc);
            omega = obs.omega;
            for (int i = 0; i < n; ++i) xavg[i] = xx0[i] - x[i] * x[mid];
            if (stab_cnt > 10) { /* :523-528 */
                stab_cnt = 0;
                dtau /= 0.950;
            }
            ++stab_cnt;
            runs += a.loops;
        } else { /* :533-545: shrink the step; libsq already rolled the frame back */
            dtau *= 0.950;
            stab_cnt = 0;
        }
        fflush(stdout);
    }

    int status = 0;
    if (strcmp(a.end_file, "0") != 0) {
        if (th_write_end_file(a.end_file, n, a.end_accuracy, xavg, xx0, x, f, omega, runs + rec_sim_length, dtau)) {
            fprintf(stderr, "Failed to write to Output.\n");
            status = 1;
        }
    }
    sq_free(ctx);
    free(f); free(x); free(xx0); free(xavg);
    return status;
}
