/*
 * sq_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See sq_oracle.h for scope, citations and the parity-pinning statement.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off is part of the definition: every +,-,*,/ below rounds once,
 * fused operations are written explicitly as fma()/fmaf().
 */
#define _GNU_SOURCE
#include "sq_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* tau_kernel.cl:19-22 */
static const double eta = .8;
static const double V0 = 2.;
static const double m = 1.;

#define LCG_A 0x5DEECE66DULL
#define LCG_B 0xBULL
#define LCG_MASK ((1ULL << 48) - 1)
#define TWO31 2147483648ULL /* (ulong)pown((float)2,31) */

/* ------------------------------------------------------------------ RNG -- */
/* tau_kernel.cl:269-284.  pown((float)2,32) == 4294967296.0f exactly. */
double sqo_random(uint64_t *seed, uint64_t gid, sqo_draw *rec)
{
    double result;
    uint64_t temp, t1;
    int n = 0, plus = 0;
    do {
        temp = ((*seed + gid) * LCG_A + LCG_B) & LCG_MASK;             /* :273 */
        t1 = temp;
        double v1 = (double)(temp >> 16) / (double)4294967296.0f;       /* :274 */
        temp = ((temp + gid) * LCG_A + LCG_B) & LCG_MASK;               /* :275 */
        double v2 = (double)(temp >> 16) / (double)4294967296.0f;       /* :276 */
        result = (double)cosf((float)(2. * 3.1415 * v2)) *
                 (double)sqrtf((float)(-2. * (double)logf((float)v1))); /* :277 */
        if (*seed < TWO31 && temp < TWO31) {                            /* :278 */
            *seed += temp;
            plus = 1;
        } else {
            *seed = temp - TWO31;                                       /* :281 */
            plus = 0;
        }
        n++;
    } while (isinf((float)result));                                     /* :282 */
    if (rec) {
        rec->t1 = t1;
        rec->t2 = temp;
        rec->seed_after = *seed;
        rec->ndraws = n;
        rec->plus_branch = plus;
    }
    return result;
}

/* ------------------------------------------------------- model functions -- */
static double doubleWellSol(double t, double t0)
{ /* :184-189 */
    return eta * (double)tanhf((float)((double)sqrtf((float)(2. * V0 / m)) * (t - t0) / eta));
}
static double doubleWellPot(double a)
{ /* :190-195 */
    return (12. * V0 * a * a / (eta * eta) - 4. * V0) / (eta * eta);
}
static double doubleWellConst(void)
{ /* :196-200 -- all-float expression, widened at the end */
    return (double)(sqrtf((float)3.) * powf((float)2., (float)(-5. / 4.)) *
                    powf((float)V0, (float)(-1. / 4.)) / sqrtf((float)eta));
}
double sqo_clas(double a, double w, int pot)
{ /* :215-226 */
    if (pot == 0) return 0.; /* harmOscSol :201-205 */
    if (pot == 3) return doubleWellSol(a, w);
    return 0.; /* reference falls off the end (UB) */
}
double sqo_ddPot(double a, int pot)
{ /* :227-236 */
    if (pot == 0) return 2.; /* harmOscPot :206-209 */
    if (pot == 3) return doubleWellPot(a);
    return 0.;
}
double sqo_intConst(int pot)
{ /* :237-246 */
    if (pot == 0) return 0.;
    if (pot == 3) return doubleWellConst();
    return 0.;
}
double sqo_boundary(int rl, int pot)
{ /* :247-256 */
    (void)pot;
    if (rl == 1) return eta;
    if (rl == -1) return -eta;
    return 0.;
}
double sqo_absol(double a)
{ /* :259-267 */
    if (a <= 0) return -a;
    return a;
}

/* ------------------------------------------------------------- time_dev -- */
static sqo_draw *g_trace = NULL;
static int g_trace_cap = 0;
void sqo_set_trace(sqo_draw *buf, int capacity)
{
    g_trace = buf;
    g_trace_cap = capacity;
}

int sqo_time_dev(sqo_state *s, int loops, int rng_mode, int field_mode)
{
    const int N = s->N;
    const double deltatau = s->deltaTau;
    const double deltat = s->deltaT;
    const int potID = s->potential;
    const double c = s->C;
    const double max = 1000;
    const int midpt = N / 2;
    double *f = s->f, *x = s->x, *xx0 = s->xx0;
    double *newf = s->newf, *newx = s->newx, *newxx0 = s->newxx0;
    double *fold = (double *)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    int executed = 0;

    for (int j = 0; j < loops; j++) {
        const double om = s->omega; /* :65 every item reads it before item N writes */
        double newomega = om;
        const uint64_t seed0 = s->rand1;
        /* Jacobi: reads of f during this step see the step-start field */
        memcpy(fold, f, sizeof(double) * (size_t)N);
        const double *fr = (field_mode == SQO_FIELD_JACOBI) ? fold : f;

        for (int i = 0; i <= N; i++) { /* work-items in gid order */
            double dw = 0;
            uint64_t myseed = seed0;
            uint64_t *sp = (rng_mode == SQO_RNG_CHAIN) ? &s->rand1 : &myseed;
            sqo_draw rec;
            if (i == 0) { /* :68-85, boundaryConditions==1 */
                dw = c * (double)sqrtf((float)(2. * deltatau / deltat)) * sqo_random(sp, (uint64_t)i, &rec);
                newf[i] = fr[0] + m * deltatau * (fr[1] + sqo_boundary(-1, potID) - sqo_clas(-1. * deltat, om, potID) - 2 * fr[0]) / (double)((float)deltat * (float)deltat) - sqo_ddPot(sqo_clas((double)i * deltat, om, potID), potID) * fr[0] * deltatau + dw;
            }
            if (i == N - 1) { /* :86-102 */
                dw = c * (double)sqrtf((float)(2. * deltatau / deltat)) * sqo_random(sp, (uint64_t)i, &rec);
                newf[i] = fr[N - 1] + m * deltatau * (fr[N - 2] + sqo_boundary(1, potID) - sqo_clas((double)N * deltat, om, potID) - 2 * fr[N - 1]) / (double)((float)deltat * (float)deltat) - sqo_ddPot(sqo_clas((double)i * deltat, om, potID), potID) * fr[N - 1] * deltatau + dw;
            }
            if (i == N) { /* :103-110 */
                dw = c * (double)sqrtf((float)(2. * deltatau)) * sqo_random(sp, (uint64_t)i, &rec);
                newomega = om + sqo_intConst(potID) * dw;
            }
            if (i < N - 1 && i > 0) { /* :111-117 */
                dw = c * (double)sqrtf((float)(2. * deltatau / deltat)) * sqo_random(sp, (uint64_t)i, &rec);
                newf[i] = fr[i] + m * deltatau * (fr[i + 1] + fr[i - 1] - 2 * fr[i]) / (double)((float)deltat * (float)deltat) - sqo_ddPot(sqo_clas((double)i * deltat, om, potID), potID) * fr[i] * deltatau + dw;
            }
            if (g_trace && i < g_trace_cap) g_trace[i] = rec;
            if (rng_mode == SQO_RNG_SHARED && i == N) s->rand1 = myseed; /* last writer wins */

            if (i < N) { /* :119-154 */
                if (newf[i] > max) newf[i] = max;
                if (newf[i] < -max) newf[i] = -max;
                if (isinf((float)newf[i]) || isnan((float)newf[i])) newf[i] = max;

                if (newf[i] + sqo_clas((double)i * deltat, om, potID) >
                    newf[s->lrgEl] + sqo_clas((double)s->lrgEl * deltat, om, potID)) { /* :135 */
                    s->lrgEl = i;
                    if (sqo_absol(newf[i] - fr[i] - dw) > s->lrgVl) s->stable = 0;
                }
                if (sqo_absol(newf[i] + sqo_clas((double)i * deltat, om, potID)) > s->lrgVl)
                    s->lrgVl = sqo_absol(newf[i] + sqo_clas((double)i * deltat, om, potID));
                newxx0[i] = xx0[i] + ((fr[i] + sqo_clas((double)i * deltat, om, potID)) * (fr[midpt] + sqo_clas((double)midpt * deltat, om, potID)) - xx0[i]) / ((double)(s->runs + j + 1)); /* :144 */
                newx[i] = x[i] + ((fr[i] + sqo_clas((double)i * deltat, om, potID)) - x[i]) / ((double)(s->runs + j + 1)); /* :145 */
                if (j < loops - 1) { /* :147-151 */
                    f[i] = newf[i];
                    xx0[i] = newxx0[i];
                    x[i] = newx[i];
                }
            } else { /* :155-167 */
                if (newomega > (double)(N - 1) * deltat) {
                    s->omega = 2 * (double)(N - 1) * deltat - newomega;
                } else if (newomega < 0) {
                    s->omega = -newomega;
                } else {
                    s->omega = newomega;
                }
            }
        }
        executed++;
        if (s->stable != 1) break; /* :168-171 */
    }
    free(fold);
    return executed;
}

/* ------------------------------------------------------ host restatement -- */
void sqo_host_init(int N, double deltat, double deltatau, int cold_start,
                   double *f, double *omega, uint64_t *rand1)
{
    double v1, v2, om;
    srand(1); /* == state of a fresh process: tauhost.c never calls srand */
    v1 = (double)(rand() + 1.) / ((double)(RAND_MAX) + 1.); /* tauhost.c:84 */
    v2 = (double)(rand() + 1.) / ((double)(RAND_MAX) + 1.);
    om = sqrt(2. * deltatau) * sin(2. * 3.14 * v2) * sqrt(-2. * log(v1)) + deltat * (double)(N / 2);
    while (om > N * deltat) om -= deltat; /* :87-89 */
    if (cold_start) {
        for (int i = 0; i < N; i++) { /* :92-100 */
            v1 = (double)(rand() + 1.) / ((double)(RAND_MAX) + 1.);
            v2 = (double)(rand() + 1.) / ((double)(RAND_MAX) + 1.);
            f[i] = sqrt(2. * deltatau) * cos(2. * 3.14 * v2) * sqrt(-2. * log(v1));
        }
    }
    *omega = om;
    *rand1 = (uint64_t)(unsigned long)abs(rand()); /* :185 */
}

int sqo_print_frame(FILE *fp, int N, const double *xavg, double dtau, int j, int frames)
{
    int n = 0;
    for (int i = 0; i < N; i++) { /* tauhost.c:485-501 (fps test done by caller) */
        if (i != 0) {
            n += fprintf(fp, " % -.20f |", log(sqo_absol(xavg[i])));
            if (i == N - 1) {
                n += fprintf(fp, "% -.20f | ", dtau);
                n += fprintf(fp, "% -.2f\n", 100. * ((double)j + 1) / (double)frames);
            }
        }
    }
    return n;
}

int sqo_write_endfile(const char *path, int N, int acc, const double *xavg,
                      const double *xx0, const double *x, const double *f,
                      double omega, int runs_plus_rec, double dtau)
{
    FILE *fp = fopen(path, "w"); /* tauhost.c:563-580 */
    if (!fp) return 1;
    for (int i = 0; i < N; i++) {
        fprintf(fp, "% -*a| % -*a| % -*a| % -*a", acc, xavg[i], acc, xx0[i], acc, x[i], acc, f[i]);
        fprintf(fp, "\n");
    }
    fprintf(fp, "% -*a|omega\n", acc, omega);
    fprintf(fp, "%*d|N\n", acc, runs_plus_rec);
    fprintf(fp, "% -*e|deltaTau\n", acc, dtau);
    fclose(fp);
    return 0;
}

int sqo_read_startfile(const char *path, int N, double deltatau,
                       double *xavg, double *xx0, double *x, double *f,
                       int *recSimlgth, double *dtautmp)
{
    /* tauhost.c:104-171: char-by-char line assembly, split on '|', atof/atoi.
     * The reference's litstr is not NUL-terminated (:119-121); terminated here. */
    FILE *fp = fopen(path, "r");
    if (!fp) return 1;
    size_t cap = 256, len = 0;
    char *buf = (char *)malloc(cap);
    int i = 0, ch;
    while ((ch = fgetc(fp)) != EOF) {
        if (ch == '\n') {
            buf[len] = 0;
            char *token;
            if (i == N + 1) {
                token = strtok(buf, "|");
                if (token) *recSimlgth = atoi(token);
            }
            if (i == N + 2) {
                token = strtok(buf, "|");
                if (token) {
                    *dtautmp = atof(token);
                    if (*dtautmp > deltatau) *dtautmp = deltatau;
                }
            }
            if (i < N) {
                token = strtok(buf, "|");
                xavg[i] = token ? atof(token) : 0.;
                token = strtok(NULL, "|");
                xx0[i] = token ? atof(token) : 0.;
                token = strtok(NULL, "|");
                x[i] = token ? atof(token) : 0.;
                token = strtok(NULL, "|");
                f[i] = token ? atof(token) : 0.;
            }
            len = 0;
            i++;
        } else {
            if (len + 2 > cap) {
                cap *= 2;
                buf = (char *)realloc(buf, cap);
            }
            buf[len++] = (char)ch;
        }
    }
    free(buf);
    fclose(fp);
    return 0;
}

int sqo_tauhost_main(int argc, char **argv, FILE *out, int rng_mode, int field_mode)
{
    if (argc < 14) return 2;
    const int LIST_SIZE = atoi(argv[1]); /* tauhost.c:31-43 */
    const double deltat = atof(argv[2]);
    const double deltatau = atof(argv[3]);
    const int frames = atoi(argv[4]);
    const int potID = atoi(argv[5]);
    const double C = atof(argv[6]);
    const int fps = atoi(argv[8]);
    const int loops = atoi(argv[10]);
    const char *startFile = argv[11];
    const char *endFile = argv[12];
    const int endAccuracy = atoi(argv[13]);
    const int N = LIST_SIZE;
    int recSimlgth = 0;
    const int midpt = N / 2;
    double omega;
    uint64_t rand1;
    double dtautmp = deltatau;

    /* host arrays (x, xx0 zero-initialised: SURVEY appendix B) */
    double *f = calloc((size_t)N, sizeof(double)), *x = calloc((size_t)N, sizeof(double));
    double *xx0 = calloc((size_t)N, sizeof(double)), *xavg = calloc((size_t)N, sizeof(double));
    const int cold = strcmp(startFile, "0") == 0;
    sqo_host_init(N, deltat, deltatau, cold, f, &omega, &rand1);
    if (!cold) {
        if (sqo_read_startfile(startFile, N, deltatau, xavg, xx0, x, f, &recSimlgth, &dtautmp)) {
            fprintf(stderr, "Failed to read Input.\n");
            return 1;
        }
    }
    /* "device" buffers */
    sqo_state d;
    d.N = N;
    d.deltaT = deltat;
    d.deltaTau = dtautmp;
    d.C = C;
    d.potential = potID;
    d.f = malloc(sizeof(double) * (size_t)N);
    d.x = malloc(sizeof(double) * (size_t)N);
    d.xx0 = malloc(sizeof(double) * (size_t)N);
    d.newf = malloc(sizeof(double) * (size_t)N);
    d.newx = malloc(sizeof(double) * (size_t)N);
    d.newxx0 = malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < N; i++) { /* tauhost.c:177-183, 319-334 */
        d.f[i] = d.newf[i] = f[i];
        d.x[i] = d.newx[i] = x[i];
        d.xx0[i] = d.newxx0[i] = xx0[i];
    }
    d.omega = omega;
    d.rand1 = rand1;
    d.stable = 1;
    d.lrgEl = 0;
    d.lrgVl = 0;
    d.runs = recSimlgth;

    int stable, stabCnt = 0, runs = recSimlgth;
    for (int j = 0; j < frames; j++) { /* tauhost.c:479-560 */
        sqo_time_dev(&d, loops, rng_mode, field_mode);
        if (j % fps == 0) sqo_print_frame(out, N, xavg, dtautmp, j, frames);
        stable = d.stable;
        if (stable == 1) {
            memcpy(f, d.newf, sizeof(double) * (size_t)N);
            memcpy(x, d.newx, sizeof(double) * (size_t)N);
            memcpy(xx0, d.newxx0, sizeof(double) * (size_t)N);
            omega = d.omega;
            for (int i = 0; i < N; i++) xavg[i] = (xx0[i] - x[i] * x[midpt]);
            if (stabCnt > 10) {
                stabCnt = 0;
                dtautmp /= 0.950;
                d.deltaTau = dtautmp;
            }
            stabCnt++;
            runs += loops;
        } else {
            dtautmp = d.deltaTau;
            dtautmp *= 0.950;
            stabCnt = 0;
            d.deltaTau = dtautmp;
            stable = 1;
            d.stable = stable;
        }
        memcpy(d.f, f, sizeof(double) * (size_t)N);
        memcpy(d.x, x, sizeof(double) * (size_t)N);
        memcpy(d.xx0, xx0, sizeof(double) * (size_t)N);
        d.omega = omega;
        d.runs = runs;
        fflush(out);
    }
    int rc = 0;
    if (strcmp(endFile, "0") != 0) {
        if (sqo_write_endfile(endFile, N, endAccuracy, xavg, xx0, x, f, omega, runs + recSimlgth, dtautmp)) {
            fprintf(stderr, "Failed to write to Output.\n");
            rc = 1;
        }
    }
    free(d.f); free(d.x); free(d.xx0); free(d.newf); free(d.newx); free(d.newxx0);
    free(f); free(x); free(xx0); free(xavg);
    return rc;
}

int sqo_print_frame_path(const char *path, int N, const double *xavg, double dtau, int j, int frames)
{
    FILE *fp = fopen(path, "w");
    if (!fp) return -1;
    int n = sqo_print_frame(fp, N, xavg, dtau, j, frames);
    fclose(fp);
    return n;
}

int sqo_tauhost_main_path(int argc, char **argv, const char *outpath, int rng_mode, int field_mode)
{
    FILE *fp = fopen(outpath, "w");
    if (!fp) return -1;
    int rc = sqo_tauhost_main(argc, argv, fp, rng_mode, field_mode);
    fclose(fp);
    return rc;
}

/* ------------------------------------------------ chain jump-ahead (oracle) -- */
/* Outside retry / += events one draw at gid g maps the seed affinely mod 2^48:
 *   t1 = A(s+g)+B ; t2 = A(t1+g)+B ; s' = t2 - 2^31
 *   => s' = al*s + be*g + ga,  al=A^2, be=A^2+A, ga=A*B+B-2^31.
 * D consecutive draws starting at gid g0:
 *   s_D = al^D s + (be*g0+ga) G0(D) + be G1(D),
 *   G0(D)=sum_{j<D} al^(D-1-j),  G1(D)=sum_{j<D} j al^(D-1-j).
 * Square-and-multiply on the triple (al^D, G0, G1); arithmetic mod 2^64,
 * masked at the end (2^48 | 2^64). */
typedef struct { uint64_t a, g0, g1, d; } jtrip;
static jtrip jt_compose(jtrip p, jtrip q)
{ /* first p (D1 draws) then q (D2 draws) */
    jtrip r;
    r.a = p.a * q.a;
    r.g0 = q.a * p.g0 + q.g0;
    r.g1 = q.a * p.g1 + p.d * q.g0 + q.g1;
    r.d = p.d + q.d;
    return r;
}
static jtrip jt_for(uint64_t D)
{
    const uint64_t al = LCG_A * LCG_A;
    jtrip res = {1, 0, 0, 0}, pw = {al, 1, 0, 1};
    while (D) {
        if (D & 1) res = jt_compose(res, pw);
        pw = jt_compose(pw, pw);
        D >>= 1;
    }
    return res;
}
uint64_t sqo_jump(uint64_t s, uint64_t g0, uint64_t D)
{
    const uint64_t al = LCG_A * LCG_A, be = al + LCG_A, ga = LCG_A * LCG_B + LCG_B - TWO31;
    (void)al;
    jtrip t = jt_for(D);
    return (t.a * s + (be * g0 + ga) * t.g0 + be * t.g1) & LCG_MASK;
}

void sqo_lattice_draws(uint64_t seed, uint64_t V, uint64_t *t1, uint64_t *t2, uint64_t *seed_after)
{
    sqo_draw rec;
    for (uint64_t g = 0; g <= V; g++) {
        sqo_random(&seed, g, &rec);
        if (g < V) {
            if (t1) t1[g] = rec.t1;
            if (t2) t2[g] = rec.t2;
        }
    }
    *seed_after = seed;
}

/* ------------------------------------------- d-dim lattice generalisation -- */
static int64_t lat_volume(const sqo_lattice *L)
{
    int64_t v = 1;
    for (int k = 0; k < L->ndim; k++) v *= L->dims[k];
    return v;
}
static double lat_noise_scale(const sqo_lattice *L, double dtau)
{
    double ad = 1.;
    for (int k = 0; k < L->ndim; k++) ad *= L->a;
    return L->C * (double)sqrtf((float)(2. * dtau / ad));
}

#define CLAMP_MAX 1000

/* update of one site; `REAL`, FMA(), nb[] = 2*ndim neighbour values in the order
 * +0,-0,+1,-1,... ; returns new value, *clamped set if the clamp fired */
#define DEFINE_SITE_UPDATE(NAME, REAL, FMA)                                          \
    static inline REAL NAME(REAL phi, const REAL *nb, int ndim, int pot, REAL c_lap, \
                            REAL c_dt, REAL m2, REAL lam, REAL dw, int *clamped)     \
    {                                                                                \
        REAL s = nb[0] + nb[1];                                                      \
        for (int k = 1; k < ndim; k++) {                                             \
            s = s + nb[2 * k];                                                       \
            s = s + nb[2 * k + 1];                                                   \
        }                                                                            \
        REAL lap = FMA(-(REAL)(2 * ndim), phi, s);                                   \
        REAL F;                                                                      \
        if (pot == 4) {                                                              \
            REAL p2 = phi * phi;                                                     \
            F = phi * FMA(lam, p2, m2);                                              \
        } else {                                                                     \
            F = (REAL)2 * phi;                                                       \
        }                                                                            \
        REAL v = FMA(c_lap, lap, phi);                                               \
        v = FMA(-c_dt, F, v);                                                        \
        v = v + dw;                                                                  \
        if (v > (REAL)CLAMP_MAX) { v = (REAL)CLAMP_MAX; *clamped = 1; }              \
        if (v < -(REAL)CLAMP_MAX) { v = -(REAL)CLAMP_MAX; *clamped = 1; }            \
        if (isinf((float)v) || isnan((float)v)) { v = (REAL)CLAMP_MAX; *clamped = 1; } \
        return v;                                                                    \
    }
DEFINE_SITE_UPDATE(site_update_f32, float, fmaf)
DEFINE_SITE_UPDATE(site_update_f64, double, fma)

/* process gids [g_begin, g_end) with the chain starting at *seed (seed before
 * the draw at g_begin).  Returns number of events (retry or += branch). */
static uint64_t lat_range(const sqo_lattice *L, double dtau, int64_t g_begin, int64_t g_end,
                          uint64_t *seed, double *slice_sum, double *sum1, double *sum2,
                          int64_t *nclamped)
{
    const int d = L->ndim;
    int64_t stride[4], dim[4];
    int64_t st = 1;
    for (int k = 0; k < d; k++) { stride[k] = st; dim[k] = L->dims[k]; st *= L->dims[k]; }
    const int64_t vslice = st / dim[d - 1];
    const double a2f = (double)((float)L->a * (float)L->a);
    const double nscale = lat_noise_scale(L, dtau);
    const double c_lap_d = (m * dtau) / a2f;
    uint64_t nev = 0;
    const float *p32 = (const float *)L->phi;
    const double *p64 = (const double *)L->phi;
    float *n32 = (float *)L->phi_new;
    double *n64 = (double *)L->phi_new;
    int64_t c[4] = {0, 0, 0, 0};
    { int64_t r = g_begin; for (int k = 0; k < d; k++) { c[k] = r % dim[k]; r /= dim[k]; } }
    for (int64_t g = g_begin; g < g_end; g++) {
        sqo_draw rec;
        double r = sqo_random(seed, (uint64_t)g, &rec);
        if (rec.ndraws > 1 || rec.plus_branch) nev++;
        int64_t nbi[8];
        for (int k = 0; k < d; k++) {
            nbi[2 * k] = (c[k] + 1 == dim[k]) ? g - (dim[k] - 1) * stride[k] : g + stride[k];
            nbi[2 * k + 1] = (c[k] == 0) ? g + (dim[k] - 1) * stride[k] : g - stride[k];
        }
        int clamped = 0;
        double phi_d;
        if (L->real == SQO_F32) {
            float nb[8];
            for (int k = 0; k < 2 * d; k++) nb[k] = p32[nbi[k]];
            float phi = p32[g];
            float dw = (float)(nscale * r);
            n32[g] = site_update_f32(phi, nb, d, L->potential, (float)c_lap_d, (float)dtau,
                                     (float)L->m2, (float)L->lambda, dw, &clamped);
            phi_d = (double)phi;
        } else {
            double nb[8];
            for (int k = 0; k < 2 * d; k++) nb[k] = p64[nbi[k]];
            double phi = p64[g];
            double dw = nscale * r;
            n64[g] = site_update_f64(phi, nb, d, L->potential, c_lap_d, dtau, L->m2, L->lambda,
                                     dw, &clamped);
            phi_d = phi;
        }
        *nclamped += clamped;
        slice_sum[g / vslice] += phi_d;
        *sum1 += phi_d;
        *sum2 += phi_d * phi_d;
        for (int k = 0; k < d; k++) { if (++c[k] < dim[k]) break; c[k] = 0; }
    }
    return nev;
}

static void lat_finish_step(sqo_lattice *L, const double *slice_sum, double sum1, double sum2)
{
    const int d = L->ndim;
    const int64_t Lt = L->dims[d - 1];
    const int64_t vslice = lat_volume(L) / Lt;
    const int64_t tmid = Lt / 2;
    const double n = (double)(L->runs + 1);
    const double phimid = slice_sum[tmid] / (double)vslice;
    for (int64_t t = 0; t < Lt; t++) {
        const double P = slice_sum[t] / (double)vslice;
        L->slice_sum[t] = slice_sum[t];
        L->slice_xx0[t] = L->slice_xx0[t] + (P * phimid - L->slice_xx0[t]) / n;
        L->slice_x[t] = L->slice_x[t] + (P - L->slice_x[t]) / n;
    }
    L->sum_phi = sum1;
    L->sum_phi2 = sum2;
    L->runs += 1;
    void *tmp = L->phi; L->phi = L->phi_new; L->phi_new = tmp;
}

void sqo_lattice_step(sqo_lattice *L, double dtau)
{
    const int64_t V = lat_volume(L);
    const int64_t Lt = L->dims[L->ndim - 1];
    double *ss = (double *)calloc((size_t)Lt, sizeof(double));
    double s1 = 0, s2 = 0;
    L->nevents += lat_range(L, dtau, 0, V, &L->seed, ss, &s1, &s2, &L->nclamped);
    sqo_draw rec;
    sqo_random(&L->seed, (uint64_t)V, &rec); /* the omega item's draw: consumed */
    if (rec.ndraws > 1 || rec.plus_branch) L->nevents++;
    lat_finish_step(L, ss, s1, s2);
    free(ss);
}

/* Thread count of sqo_lattice_step_omp.  torchrun exports OMP_NUM_THREADS=1 to its workers, so the
 * timing legs of bench.py set the count explicitly (n <= 0: leave as is) and report what they got. */
int sqo_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

void sqo_lattice_step_omp(sqo_lattice *L, double dtau)
{
    const int64_t V = lat_volume(L);
    const int64_t Lt = L->dims[L->ndim - 1];
    const int64_t vslice = V / Lt;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    /* chunks are whole multiples of a time-slice fraction so that per-chunk slice
     * sums stay simple: chunk = contiguous gid range */
    int nchunks = nthreads * 4;
    if (nchunks > V) nchunks = (int)V;
    double *ss = (double *)calloc((size_t)Lt * (size_t)nchunks, sizeof(double));
    double *s1 = (double *)calloc((size_t)nchunks, sizeof(double));
    double *s2 = (double *)calloc((size_t)nchunks, sizeof(double));
    int64_t *ncl = (int64_t *)calloc((size_t)nchunks, sizeof(int64_t));
    uint64_t *nev = (uint64_t *)calloc((size_t)nchunks, sizeof(uint64_t));
    const uint64_t base = L->seed;
    /* the first draw of the step sees the full u64 seed (:278 test); later chunk
     * starts are t2-2^31 values whose mod-2^48 image is equivalent */
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nchunks; k++) {
        const int64_t gb = V * k / nchunks, ge = V * (k + 1) / nchunks;
        uint64_t sd = (gb == 0) ? base : sqo_jump(base, 0, (uint64_t)gb);
        nev[k] = lat_range(L, dtau, gb, ge, &sd, ss + (size_t)k * Lt, &s1[k], &s2[k], &ncl[k]);
    }
    uint64_t events = 0;
    for (int k = 0; k < nchunks; k++) events += nev[k];
    uint64_t sd_end = sqo_jump(base, 0, (uint64_t)V);
    sqo_draw rec;
    sqo_random(&sd_end, (uint64_t)V, &rec);
    if (rec.ndraws > 1 || rec.plus_branch) events++;
    (void)vslice;
    if (events) {
        /* an event invalidates every prediction after it: redo serially (rare) */
        free(ss); free(s1); free(s2); free(ncl); free(nev);
        sqo_lattice_step(L, dtau);
        return;
    }
    double *tot = (double *)calloc((size_t)Lt, sizeof(double));
    double t1 = 0, t2 = 0;
    for (int k = 0; k < nchunks; k++) {
        for (int64_t t = 0; t < Lt; t++) tot[t] += ss[(size_t)k * Lt + t];
        t1 += s1[k];
        t2 += s2[k];
        L->nclamped += ncl[k];
    }
    L->seed = sd_end;
    lat_finish_step(L, tot, t1, t2);
    free(tot); free(ss); free(s1); free(s2); free(ncl); free(nev);
}
