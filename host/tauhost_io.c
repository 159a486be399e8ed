/* tauhost_io.c -- see tauhost_io.h.  No GPU code here. */
#define _GNU_SOURCE
#include "tauhost_io.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

int th_parse_args(int argc, char **argv, th_args *a, char *err, size_t errlen)
{
    memset(a, 0, sizeof *a);
    if (argc != 14 && argc != 16) {
        snprintf(err, errlen,
                 "usage: %s N deltat deltatau frames potID C dev fps inTime loops startFile endFile endAccuracy\n"
                 "   or: %s N deltat deltatau h parisi frames potID C dev fps inTime loops startFile endFile endAccuracy\n",
                 argv[0], argv[0]);
        return -1;
    }
    int k = 1;
    a->n = atoi(argv[k++]);
    a->deltat = atof(argv[k++]);
    a->deltatau = atof(argv[k++]);
    if (argc == 16) {
        a->h = atof(argv[k++]);
        a->parisi = atoi(argv[k++]);
    }
    a->frames = atoi(argv[k++]);
    a->pot_id = atoi(argv[k++]);
    a->c = atof(argv[k++]);
    a->dev = atoi(argv[k++]);
    a->fps = atoi(argv[k++]);
    a->in_time = atoi(argv[k++]);
    a->loops = atoi(argv[k++]);
    a->start_file = argv[k++];
    a->end_file = argv[k++];
    a->end_accuracy = atoi(argv[k++]);
    if (a->parisi != 0) {
        snprintf(err, errlen, "parisi != 0 is not supported: the reference has no kernel for it.\n");
        return -1;
    }
    if (a->n < 3 || a->frames < 0 || a->loops < 1 || a->fps < 1 || !(a->deltat > 0) || !(a->deltatau > 0)) {
        snprintf(err, errlen, "invalid arguments (need N>=3, loops>=1, fps>=1, deltat>0, deltatau>0).\n");
        return -1;
    }
    return 0;
}

static double unit_draw(void)
{ /* (rand()+1)/(RAND_MAX+1) in (0,1] */
    return (double)(rand() + 1.) / ((double)(RAND_MAX) + 1.);
}

void th_initial_state(int n, double deltat, double deltatau, int cold_start, double *f,
                      double *omega, unsigned long *rand1)
{
    const double amp = sqrt(2. * deltatau);
    double u1 = unit_draw();
    double u2 = unit_draw();
    /* Box-Muller sine branch around the lattice midpoint, "2 pi" = 2*3.14 as in the reference */
    double w = amp * sin(2. * 3.14 * u2) * sqrt(-2. * log(u1)) + deltat * (double)(n / 2);
    while (w > n * deltat) w -= deltat;
    if (cold_start) {
        for (int i = 0; i < n; ++i) {
            u1 = unit_draw();
            u2 = unit_draw();
            f[i] = amp * cos(2. * 3.14 * u2) * sqrt(-2. * log(u1));
        }
    }
    *omega = w;
    *rand1 = (unsigned long)abs(rand());
}

int th_print_frame(FILE *out, int n, const double *xavg, double dtau, int frame, int frames)
{
    /* worst case per field: sign + 310 digits + '.' + 20 decimals; log|x| is < 1e3 in magnitude,
     * so 48 bytes per field is ample -- still, grow on demand */
    size_t cap = (size_t)n * 48 + 128, len = 0;
    char *buf = (char *)malloc(cap);
    if (!buf) return -1;
    for (int i = 1; i < n; ++i) {
        if (cap - len < 400) {
            cap *= 2;
            char *nb = (char *)realloc(buf, cap);
            if (!nb) { free(buf); return -1; }
            buf = nb;
        }
        const double a = xavg[i] <= 0 ? -xavg[i] : xavg[i];
        len += (size_t)snprintf(buf + len, cap - len, " % -.20f |", log(a));
    }
    if (n > 1) {
        len += (size_t)snprintf(buf + len, cap - len, "% -.20f | ", dtau);
        len += (size_t)snprintf(buf + len, cap - len, "% -.2f\n", 100. * ((double)frame + 1) / (double)frames);
    }
    fwrite(buf, 1, len, out);
    free(buf);
    return (int)len;
}

int th_write_end_file(const char *path, int n, int width, const double *xavg, const double *xx0,
                      const double *x, const double *f, double omega, int runs_field, double dtau)
{
    return th_write_end_file_ext(path, n, width, xavg, xx0, x, f, omega, runs_field, dtau, NULL);
}

int th_write_end_file_ext(const char *path, int n, int width, const double *xavg, const double *xx0,
                          const double *x, const double *f, double omega, int runs_field, double dtau, const th_ext *ext)
{
    FILE *fp = fopen(path, "w");
    if (!fp) return 1;
    for (int i = 0; i < n; ++i)
        fprintf(fp, "% -*a| % -*a| % -*a| % -*a\n", width, xavg[i], width, xx0[i], width, x[i], width, f[i]);
    fprintf(fp, "% -*a|omega\n", width, omega);
    fprintf(fp, "%*d|N\n", width, runs_field);
    fprintf(fp, "% -*e|deltaTau\n", width, dtau);
    if (ext) {
        fprintf(fp, "1|sqext\n");
        fprintf(fp, "%llu|rand1\n", ext->seed);
        fprintf(fp, "%d|lrgEl\n", ext->lrgEl);
        fprintf(fp, "%a|lrgVl\n", ext->lrgVl);
        fprintf(fp, "%a|omega\n", ext->omega);
        fprintf(fp, "%a|newf_lrgEl\n", ext->newf_lrgEl);
        fprintf(fp, "%a|deltaTau\n", ext->dtau);
        fprintf(fp, "%d|stabCnt\n", ext->stab_cnt);
        fprintf(fp, "%lld|runs\n", ext->runs);
    }
    fclose(fp);
    return 0;
}

/* next '|'-separated token of *cursor as a C string (in place); NULL when exhausted.
 * Like strtok: empty tokens are skipped. */
static char *next_field(char **cursor)
{
    char *p = *cursor;
    if (!p) return NULL;
    while (*p == '|') ++p;
    if (!*p) { *cursor = NULL; return NULL; }
    char *start = p;
    while (*p && *p != '|') ++p;
    if (*p) { *p = 0; *cursor = p + 1; } else { *cursor = NULL; }
    return start;
}

int th_read_start_file(const char *path, int n, double deltatau, double *xavg, double *xx0,
                       double *x, double *f, int *rec_sim_length, double *dtau)
{
    return th_read_start_file_ext(path, n, deltatau, xavg, xx0, x, f, rec_sim_length, dtau, NULL, NULL);
}

int th_read_start_file_ext(const char *path, int n, double deltatau, double *xavg, double *xx0,
                           double *x, double *f, int *rec_sim_length, double *dtau, th_ext *ext, int *has_ext)
{
    int ext_lines = 0;
    if (has_ext) *has_ext = 0;
    FILE *fp = fopen(path, "r");
    if (!fp) return 1;
    char *line = NULL;
    size_t cap = 0;
    ssize_t got;
    int row = 0;
    while ((got = getline(&line, &cap, fp)) >= 0) {
        if (got == 0 || line[got - 1] != '\n') break; /* the reference only acts on '\n'-terminated lines */
        line[got - 1] = 0;
        char *cur = line, *tok;
        if (row < n) {
            double v[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4 && (tok = next_field(&cur)); ++k) v[k] = atof(tok);
            xavg[row] = v[0];
            xx0[row] = v[1];
            x[row] = v[2];
            f[row] = v[3];
        } else if (row == n + 1) {
            if ((tok = next_field(&cur))) *rec_sim_length = atoi(tok);
        } else if (row == n + 2) {
            if ((tok = next_field(&cur))) {
                double d = atof(tok);
                *dtau = d > deltatau ? deltatau : d;
            }
        } else if (row > n + 2 && ext) { /* extended trailer: "value|name" lines behind the reference's three */
            char *val = next_field(&cur), *name = next_field(&cur);
            if (val && name) {
                if (row == n + 3) { if (strcmp(name, "sqext") == 0 && atoi(val) == 1) ext_lines = 1; }
                else if (ext_lines >= 1) {
                    if (!strcmp(name, "rand1")) { ext->seed = strtoull(val, NULL, 10); ++ext_lines; }
                    else if (!strcmp(name, "lrgEl")) { ext->lrgEl = atoi(val); ++ext_lines; }
                    else if (!strcmp(name, "lrgVl")) { ext->lrgVl = strtod(val, NULL); ++ext_lines; }
                    else if (!strcmp(name, "omega")) { ext->omega = strtod(val, NULL); ++ext_lines; }
                    else if (!strcmp(name, "newf_lrgEl")) { ext->newf_lrgEl = strtod(val, NULL); ++ext_lines; }
                    else if (!strcmp(name, "deltaTau")) { ext->dtau = strtod(val, NULL); ++ext_lines; }
                    else if (!strcmp(name, "stabCnt")) { ext->stab_cnt = atoi(val); ++ext_lines; }
                    else if (!strcmp(name, "runs")) { ext->runs = atoll(val); ++ext_lines; }
                }
            }
        } /* row == n: the omega line is ignored (tauhost.c:122-124) */
        ++row;
    }
    if (has_ext && ext_lines == 9) *has_ext = 1;
    free(line);
    fclose(fp);
    return 0;
}
