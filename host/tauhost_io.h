/*
 * tauhost_io.h -- host-side pieces of the drop-in `tauhost.o` that do not touch the GPU:
 * initial state, stdout frame line, start/end file codec.  Behaviour follows
 * /root/reference/tauhost.c (line numbers cited per function); the code is new.
 * Built twice: into ./tauhost.o and into host/libtauhost_io.so for the CPU tests.
 */
#ifndef TAUHOST_IO_H
#define TAUHOST_IO_H
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Parsed command line.  13-argument form: tauhost.c:31-43 (what taumain.py:132 passes);
 * 15-argument form: taumain_windows.py:163 (adds h, parisi after deltatau). */
typedef struct th_args {
    int n;             /* LIST_SIZE */
    double deltat, deltatau;
    int frames, pot_id;
    double c;
    int dev, fps, in_time, loops;
    const char *start_file, *end_file;
    int end_accuracy;
    double h;          /* 15-arg form only (ignored) */
    int parisi;        /* 15-arg form only (must be 0) */
} th_args;

/* returns 0, or -1 with a message in err (argc as passed to main) */
int th_parse_args(int argc, char **argv, th_args *a, char *err, size_t errlen);

/* tauhost.c:84-102 and :185 -- omega0, cold-start f[] and the device seed from the
 * process's (unseeded) glibc rand() stream, consumed in the reference's order. */
void th_initial_state(int n, double deltat, double deltatau, int cold_start, double *f,
                      double *omega, unsigned long *rand1);

/* tauhost.c:485-501 -- one stdout line: n-1 fields " % -.20f |" of log|xavg[i]|, then
 * "% -.20f | " dtau and "% -.2f\n" percent.  Written with one fwrite. */
int th_print_frame(FILE *out, int n, const double *xavg, double dtau, int frame, int frames);

/* tauhost.c:562-581 -- returns 0, or 1 if the file cannot be opened */
int th_write_end_file(const char *path, int n, int width, const double *xavg, const double *xx0,
                      const double *x, const double *f, double omega, int runs_field, double dtau);

/* tauhost.c:103-173 -- site lines -> xavg,xx0,x,f; omega line skipped; line n+1 ->
 * rec_sim_length; line n+2 -> dtau capped at deltatau.  returns 0 / 1 (cannot open) */
int th_read_start_file(const char *path, int n, double deltatau, double *xavg, double *xx0,
                       double *x, double *f, int *rec_sim_length, double *dtau);

/* ---- extended trailer: what the reference's end file lacks for a bit-exact resume.
 * Written BEHIND the reference's three trailer lines, one "value|name" line each, first line "1|sqext":
 * the reference's reader (tauhost.c:116-168) acts on lines 0..N+2 only, so such a file still restarts the
 * reference (which then re-randomises as it always does).  A restart that honours the trailer continues the
 * run as if it had never stopped: same seed, lrgEl / lrgVl, omega, the stale newf[lrgEl], the step size to the
 * last bit (the reference prints it with 7 digits), the controller's counter and the true tau-step count (the
 * reference's `N` line double-counts across restarts, tauhost.c:477,577). */
typedef struct th_ext {
    unsigned long long seed;   /* device RNG seed `rand1` */
    int lrgEl, stab_cnt;
    long long runs;            /* tau-steps in the running means */
    double lrgVl, omega, newf_lrgEl, dtau;
} th_ext;
/* th_write_end_file + the extended trailer (ext == NULL: exactly th_write_end_file) */
int th_write_end_file_ext(const char *path, int n, int width, const double *xavg, const double *xx0,
                          const double *x, const double *f, double omega, int runs_field, double dtau, const th_ext *ext);
/* th_read_start_file; *has_ext = 1 and *ext filled when the file carries a complete extended trailer */
int th_read_start_file_ext(const char *path, int n, double deltatau, double *xavg, double *xx0,
                           double *x, double *f, int *rec_sim_length, double *dtau, th_ext *ext, int *has_ext);

#ifdef __cplusplus
}
#endif
#endif
