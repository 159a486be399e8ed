/*
 * tauhost.c -- drop-in replacement for the reference's host program
 * (/root/reference/tauhost.c): same positional argv (:31-43), same stdout frame stream
 * (:485-501), same start/end files (:103-173, :562-581), same exit codes and stderr
 * messages -- so /root/reference/taumain.py (:132 Popen) drives it unchanged.
 *
 * What changed underneath: the OpenCL context/queue/buffer/kernel-launch code
 * (:187-481, :504-554, :587-612) is gone; the frame loop calls the C-ABI of libsq
 * (include/sq.h): sq_init / sq_frames (= sq_step + the controller of :504-545, on the device) /
 * sq_measure / sq_free.  No tau_kernel.cl is
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
    /* TAUHOST_EXT_TRAILER=1 (not a reference feature): write the extended trailer behind the end
     * file's three trailer lines and honour one found in the start file -- a bit-exact resume.  Unset: the
     * reference's behaviour to the letter (restart re-randomises omega and the seed, `N` double-counts). */
    const char *ext_env = getenv("TAUHOST_EXT_TRAILER");
    const int use_ext = ext_env && atoi(ext_env) == 1;
    th_ext ext;
    int has_ext = 0;
    memset(&ext, 0, sizeof ext);
    if (!cold && th_read_start_file_ext(a.start_file, n, a.deltatau, xavg, xx0, x, f, &rec_sim_length, &dtau, &ext, &has_ext)) {
        fprintf(stderr, "Failed to read Input.\n");
        return 1;
    }
    has_ext = has_ext && use_ext;
    if (has_ext) { /* continue as if the run had never stopped */
        omega = ext.omega;
        rand1 = (unsigned long)ext.seed;
        dtau = ext.dtau;
        rec_sim_length = 0;
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

    /* The frame loop of tauhost.c:479-560 with its controller (:504-545) on the device: frames run in
     * batches with no host round trip in between (sq_frames); the host formats the lines afterwards.
     * Line j shows the xavg of the last accepted frame before j and the step size frame j ran with,
     * exactly what the reference prints between its clFinish and its read-backs (:483-501). */
    int64_t runs = has_ext ? (int64_t)ext.runs : rec_sim_length;
    int stab_cnt = has_ext ? ext.stab_cnt : 0;
    if (has_ext) {
        sq_compat_state st;
        memset(&st, 0, sizeof st);
        st.struct_size = sizeof st;
        st.seed = ext.seed; st.lrgEl = ext.lrgEl; st.lrgVl = ext.lrgVl; st.omega = ext.omega; st.newf_lrgEl = ext.newf_lrgEl;
        if ((rc = sq_compat_set_state(ctx, &st)) != SQ_OK) return fail_sq("sq_compat_set_state", rc);
    }
    if ((rc = sq_controller_set(ctx, dtau, runs, stab_cnt)) != SQ_OK) return fail_sq("sq_controller_set", rc);
    sq_frame_rec recs[SQ_FRAMES_MAX];
    double *xlog = (double *)malloc(sizeof(double) * (size_t)SQ_FRAMES_MAX * (size_t)n);
    if (!xlog) return 3;
    /* small batches first: the front-end shows progress while the step size is still settling */
    for (int j = 0; j < a.frames;) {
        int nb = a.frames - j;
        const int cap = j < 64 ? 8 : SQ_FRAMES_MAX;
        if (nb > cap) nb = cap;
        if ((rc = sq_frames(ctx, nb, a.loops, recs, xlog)) != SQ_OK) return fail_sq("sq_frames", rc);
        for (int k = 0; k < nb; ++k, ++j) {
            if (j % a.fps == 0) th_print_frame(stdout, n, xavg, recs[k].dtau, j, a.frames);
            if (recs[k].stable == 1) memcpy(xavg, xlog + (size_t)k * n, sizeof(double) * (size_t)n);
            fflush(stdout); /* every frame, as the reference does (tauhost.c:558) */
        }
    }
    free(xlog);
    if ((rc = sq_measure(ctx, &obs)) != SQ_OK) return fail_sq("sq_measure", rc); /* f, x, xx0, omega for the end file */
    omega = obs.omega;
    if ((rc = sq_controller_get(ctx, &dtau, &runs, &stab_cnt)) != SQ_OK) return fail_sq("sq_controller_get", rc);

    int status = 0;
    if (strcmp(a.end_file, "0") != 0) {
        th_ext out;
        memset(&out, 0, sizeof out);
        if (use_ext) {
            sq_compat_state st;
            memset(&st, 0, sizeof st);
            st.struct_size = sizeof st;
            if ((rc = sq_compat_get_state(ctx, &st)) != SQ_OK) return fail_sq("sq_compat_get_state", rc);
            out.seed = st.seed; out.lrgEl = st.lrgEl; out.lrgVl = st.lrgVl; out.omega = st.omega; out.newf_lrgEl = st.newf_lrgEl;
            out.dtau = dtau; out.stab_cnt = stab_cnt; out.runs = (long long)runs;
        }
        if (th_write_end_file_ext(a.end_file, n, a.end_accuracy, xavg, xx0, x, f, omega, (int)runs + rec_sim_length, dtau,
                                  use_ext ? &out : NULL)) {
            fprintf(stderr, "Failed to write to Output.\n");
            status = 1;
        }
    }
    sq_free(ctx);
    free(f); free(x); free(xx0); free(xavg);
    return status;
}
