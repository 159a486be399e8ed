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

#ifdef __cplusplus
}
#endif
#endif
