// sq_api.cu -- the C-ABI of include/sq.h: context, stream and memory ownership, launch
// sequencing, RNG event replay, commit/rollback.  Host side of what tauhost.c:187-481,
// :504-554 and :587-612 did through OpenCL.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "sq_ctx.h"

#ifndef SQ_TILE_DEFAULT
#define SQ_TILE_DEFAULT 4  // streaming fp32 d >= 3, event-free steps: 4 = tile kernel (sq_tile.cu), 0 = marching kernel only
#endif

using namespace sq;

static thread_local char g_cuda_err_tls[512] = "";
namespace sq {
char *cuda_err_buf() { return g_cuda_err_tls; }
}
#define g_cuda_err (sq::cuda_err_buf())


int sq_timing_mark(sq_ctx *c) {  // record the next pooled event on the stream
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        c->ev_pool.push_back(e);
    }
    CK(cudaEventRecord(c->ev_pool[c->ev_used++], c->stream));
    return SQ_OK;
}
// after a stream sync: sum the first `valid` (start,stop) pairs -- launches that were aborted
// because an earlier one flagged an RNG event are not update-kernel time
int sq_timing_collect(sq_ctx *c, size_t valid) {
    for (size_t i = 0; i + 1 < c->ev_used && i / 2 < valid; i += 2) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev_pool[i], c->ev_pool[i + 1]));
        c->timing_ms += ms;
        c->timing_launches++;
    }
    c->ev_used = 0;
    return SQ_OK;
}

// ------------------------------------------------------------------- helpers --------------
int sq_set_dev(sq_ctx *c) {
    CK(cudaSetDevice(c->p.device));
    return SQ_OK;
}

template <typename T>
static int dalloc(T **p, size_t n) {
    CK(cudaMalloc((void **)p, n * sizeof(T)));
    CK(cudaMemset(*p, 0, n * sizeof(T)));
    return SQ_OK;
}

extern "C" const char *sq_strerror(int code) {
    switch (code) {
        case SQ_OK: return "ok";
        case SQ_ERR_INVALID: return "invalid argument";
        case SQ_ERR_CUDA: return "CUDA error";
        case SQ_ERR_NOMEM: return "out of memory";
        case SQ_ERR_UNSUPPORTED: return "unsupported (no kernel for this potential / mode in the reference)";
        case SQ_ERR_NODEVICE: return "no usable CUDA device (libsq has no CPU fallback)";
        case SQ_ERR_TIMEOUT: return "bounded wait timed out (halo flag or session barrier)";
        case SQ_ERR_INTERNAL: return "internal invariant violated";
    }
    return "unknown error";
}
extern "C" const char *sq_last_cuda_error(void) { return g_cuda_err; }
extern "C" int sq_api_version(void) { return SQ_API_VERSION; }
extern "C" int sq_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        snprintf(g_cuda_err, 512, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? 0 : -1;
    }
    return n;
}
extern "C" void *sq_stream(sq_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int64_t sq_launch_count(sq_ctx *c) { return c ? c->launches : 0; }

// intConst(potID), tau_kernel.cl:196-200 / :237-246: an all-float expression widened at the end
static double host_intconst(int pot) {
    if (pot != 3) return 0.;
    const float V0f = (float)2., etaf = (float).8;
    return (double)(sqrtf((float)3.) * powf((float)2., (float)(-5. / 4.)) * powf(V0f, (float)(-1. / 4.)) /
                    sqrtf(etaf));
}

// ------------------------------------------------------------------- init / free ----------
extern "C" void sq_free(sq_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->p.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    sq_slab_destroy(c);
    void *ptrs[] = {c->d_jump, c->c_f, c->c_x, c->c_xx0, c->c_newf, c->c_newx, c->c_newxx0, c->c_omega,
                    c->c_lrgVl, c->c_red, c->c_seed, c->c_nevents, c->c_stable, c->c_lrgEl, c->c_steps, c->c_ctl, c->c_log_rec, c->c_log_xavg,
                    c->l_field[0], c->l_field[1], c->l_ghost[0], c->l_ghost[1], c->l_seeds[0], c->l_seeds[1],
                    c->l_event, c->l_rebase, c->l_partials, c->l_slice_sum, c->l_slice_x, c->l_slice_xx0,
                    c->l_sums, c->l_sums_mean, c->l_m2, c->l_lam, c->l_redbuf, c->l_nclamped, c->l_nclamp_step, c->l_slice_jump, c->l_strip_jump, c->l_cta_jump, c->l_thr_jump, c->l_tile_thr, c->l_rows_thr, c->l_tile_ctr,
                    c->r_halo, c->r_error, c->r_progress, c->r_ckpt, c->r_hist_rows, c->r_step_sums, c->r_nclamp_slots};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_upd[i]) cudaEventDestroy(c->ev_upd[i]);
        if (c->ev_fin[i]) cudaEventDestroy(c->ev_fin[i]);
    }
    if (c->fin_stream) { cudaStreamSynchronize(c->fin_stream); cudaStreamDestroy(c->fin_stream); }
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->h_pin2) cudaFreeHost(c->h_pin2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int init_compat(sq_ctx *c, const double *f0, const double *x0, const double *xx0_0, double omega0,
                       uint64_t seed) {
    const sq_params &p = c->p;
    if (p.ndim != 1 || p.real != SQ_REAL_F64) return SQ_ERR_INVALID;
    if (p.potential != SQ_POT_HARMONIC && p.potential != SQ_POT_DOUBLEWELL) return SQ_ERR_UNSUPPORTED;
    const int64_t N = p.dims[0];
    // the frame kernel keeps its whole state on chip: 56 N + 16 bytes of shared memory (sq_compat1d.cu)
    if (N < 3 || N > 4096) return SQ_ERR_INVALID;
    int rc;
    double **arrs[] = {&c->c_f, &c->c_x, &c->c_xx0, &c->c_newf, &c->c_newx, &c->c_newxx0};
    for (auto a : arrs)
        if ((rc = dalloc(a, (size_t)N))) return rc;
    if ((rc = dalloc(&c->c_omega, 1))) return rc;
    if ((rc = dalloc(&c->c_lrgVl, 1))) return rc;
    if ((rc = dalloc(&c->c_red, (size_t)N + 8))) return rc;
    if ((rc = dalloc(&c->c_seed, 1))) return rc;
    if ((rc = dalloc(&c->c_nevents, 1))) return rc;
    if ((rc = dalloc(&c->c_stable, 1))) return rc;
    if ((rc = dalloc(&c->c_lrgEl, 1))) return rc;
    if ((rc = dalloc(&c->c_steps, 1))) return rc;
    if ((rc = dalloc(&c->c_ctl, 1))) return rc;
    if ((rc = dalloc(&c->c_log_rec, SQ_FRAMES_MAX))) return rc;
    if ((rc = dalloc(&c->c_log_xavg, (size_t)SQ_FRAMES_MAX * (size_t)N))) return rc;
    const size_t nb = sizeof(double) * (size_t)N;
    // tauhost.c:177-183 + :319-334: new* start as copies of f/x/xx0
    if (f0) {
        CK(cudaMemcpy(c->c_f, f0, nb, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->c_newf, f0, nb, cudaMemcpyHostToDevice));
    }
    if (x0) {
        CK(cudaMemcpy(c->c_x, x0, nb, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->c_newx, x0, nb, cudaMemcpyHostToDevice));
    }
    if (xx0_0) {
        CK(cudaMemcpy(c->c_xx0, xx0_0, nb, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->c_newxx0, xx0_0, nb, cudaMemcpyHostToDevice));
    }
    const int one = 1;
    CK(cudaMemcpy(c->c_omega, &omega0, sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->c_seed, &seed, sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->c_stable, &one, sizeof(int), cudaMemcpyHostToDevice));
    return SQ_OK;
}

static int init_lattice(sq_ctx *c, const double *f0, uint64_t seed) {
    sq_params &p = c->p;
    if (p.ndim < 2 || p.ndim > 4) return SQ_ERR_INVALID;
    if (p.potential != SQ_POT_HARMONIC && p.potential != SQ_POT_PHI4) return SQ_ERR_UNSUPPORTED;
    if (p.real != SQ_REAL_F32 && p.real != SQ_REAL_F64) return SQ_ERR_INVALID;
    if (p.nchains < 1) p.nchains = 1;
    if (p.nchains > 16383) return SQ_ERR_INVALID;
    c->rsz = p.real == SQ_REAL_F32 ? 4 : 8;
    const int vec = (int)(16 / c->rsz);
    int64_t V = 1;
    for (int k = 0; k < p.ndim; ++k) {
        if (p.dims[k] < 2) return SQ_ERR_INVALID;
        V *= p.dims[k];
    }
    if (p.dims[0] % vec != 0) return SQ_ERR_INVALID;  // strips never straddle rows
    const int64_t Lt = p.dims[p.ndim - 1];
    c->V = V;
    c->vslice = V / Lt;
    if (c->vslice >= (1LL << 31) || V >= (1LL << 34)) return SQ_ERR_INVALID;
    if (p.slab_nt == 0) {
        p.slab_t0 = 0;
        p.slab_nt = Lt;
    }
    // (grid.y = local slices; finalize_kernel reduces them in 16 nt bytes of shared memory)
    if (p.slab_t0 < 0 || p.slab_nt < 1 || p.slab_t0 + p.slab_nt > Lt || p.slab_nt > FINALIZE_MAX_NT) return SQ_ERR_INVALID;
    if (p.slab_nt != Lt && p.nchains != 1) return SQ_ERR_INVALID;
    c->nt = (int)p.slab_nt;
    c->vlocal = c->vslice * c->nt;
    const int64_t nstrips = c->vslice / vec;
    // a few strips per thread amortise the table jump; keep >= ~4 CTAs per SM in flight
    int64_t cps = (nstrips + 256 * 4 - 1) / (256 * 4);
    if (cps < 1) cps = 1;
    c->ctas_per_slice = (int)std::min<int64_t>(cps, 65535);
    // fp32, d >= 3: the row-marching kernel (sq_march.cu; 4-site strips, dims[0]/4 threads per row a power of two) or the tile
    // kernel (sq_tile.cu; 8- or 4-site strips, tiles of 4 / 8 / 16 rows per thread that are whole planes or divide one, at
    // most 72 KB of shared memory)
    if (p.real == SQ_REAL_F32 && p.ndim >= 3 && !(p.flags & SQ_FLAG_GENERIC_KERNEL)) {
        const int64_t L0 = p.dims[0], L1 = p.dims[1], nrows = c->vslice / L0;
        static const int force_R = getenv("SQ_MARCH_R") ? atoi(getenv("SQ_MARCH_R")) : 0;    // tuning knob
        static const int tile_w = getenv("SQ_TILE") ? (atoi(getenv("SQ_TILE")) ? 4 : 0) : SQ_TILE_DEFAULT;  // 0: marching kernel only, 4: tile kernel
        static const bool env_rows = getenv("SQ_ROWS") && atoi(getenv("SQ_ROWS")) == 1;   // A/B knob
        const bool want_rows = env_rows || (p.flags & SQ_FLAG_ROWBLOCK_KERNEL);
        int best = 0, best_w = 4;
        for (int attempt = 0; attempt < 3 && !best; ++attempt) {
            // attempt 0: tile kernel with the requested strip width; 1: tile kernel with 4-site strips; 2: marching kernel
            const bool tile = attempt < 2;
            const int w = attempt == 0 ? tile_w : 4;
            if (tile && (tile_w == 0 || (attempt == 1 && tile_w == 4))) continue;
            if (w != 4 && w != 8) continue;
            const int64_t tpr = L0 / w;
            if (L0 % w != 0 || tpr < 1 || tpr > 256 || (tpr & (tpr - 1)) != 0) continue;
            const int64_t rg = 256 / tpr;
            int tlog = 0;
            while ((1 << tlog) < tpr) tlog++;
            for (int R = force_R ? force_R : 16; R >= 1; R >>= 1) {
                if (L1 % R != 0 || nrows % (rg * R) != 0 || nrows / (rg * R) > 65535) continue;
                if (tile && !tile_shape_ok(p.ndim, (int)L0, (int)L1, tlog, R, want_rows)) continue;
                const int64_t ctas = nrows / (rg * R) * c->nt * p.nchains;
                best = R;  // the largest that fits, unless a smaller one is needed to fill the GPU
                best_w = w;
                if (ctas >= 148 * 12) break;
            }
            if (best) {
                c->march_ok = true;
                c->tile_ok = tile;
                c->rows_ok = tile && want_rows;
                c->m_R = best;
                c->m_w = best_w;
                c->m_tpr_log = tlog;
                c->ctas_per_slice = (int)(nrows / (rg * best));
            }
        }
    }

    const size_t fbytes = (size_t)c->vlocal * c->rsz * (size_t)p.nchains;
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&c->l_field[b], fbytes));
        CK(cudaMemset(c->l_field[b], 0, fbytes));
        CK(cudaMalloc(&c->l_ghost[b], (size_t)c->vslice * c->rsz));
        CK(cudaMemset(c->l_ghost[b], 0, (size_t)c->vslice * c->rsz));
    }
    int rc;
    if ((rc = dalloc(&c->l_seeds[0], (size_t)p.nchains))) return rc;
    if ((rc = dalloc(&c->l_seeds[1], (size_t)p.nchains))) return rc;
    if ((rc = dalloc(&c->l_event, 1))) return rc;
    if ((rc = dalloc(&c->l_rebase, MAX_REBASE))) return rc;
    const size_t npart = (size_t)p.nchains * c->nt * c->ctas_per_slice * 2;
    {   // finalizes are handed to the side stream in groups of fin_batch steps (sq_enqueue_step); SQ_FIN_BATCH=1: per step
        static const int env_batch = getenv("SQ_FIN_BATCH") ? atoi(getenv("SQ_FIN_BATCH")) : 4;  // A/B knob
        c->fin_batch = std::min(std::max(env_batch, 1), (int)sq_ctx::FIN_BATCH_MAX);
        c->npart = npart;
    }
    if ((rc = dalloc(&c->l_partials, npart * 2 * (size_t)c->fin_batch))) return rc;
    CK(cudaStreamCreateWithFlags(&c->fin_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&c->ev_upd[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_fin[i], cudaEventDisableTiming));
    }
    const size_t nsl = (size_t)p.nchains * c->nt;
    if ((rc = dalloc(&c->l_slice_sum, nsl))) return rc;
    if ((rc = dalloc(&c->l_slice_x, nsl))) return rc;
    if ((rc = dalloc(&c->l_slice_xx0, nsl))) return rc;
    if ((rc = dalloc(&c->l_sums, (size_t)p.nchains * 2))) return rc;
    if ((rc = dalloc(&c->l_sums_mean, (size_t)p.nchains * 2))) return rc;
    if ((rc = dalloc(&c->l_m2, (size_t)p.nchains))) return rc;
    if ((rc = dalloc(&c->l_lam, (size_t)p.nchains))) return rc;
    if ((rc = dalloc(&c->l_redbuf, 2 * 1024))) return rc;
    if ((rc = dalloc(&c->l_nclamped, 1))) return rc;
    if ((rc = dalloc(&c->l_nclamp_step, MAX_SEQ_STEPS))) return rc;
    {  // jump coefficients that do not depend on the seed: slice starts and first strips
        std::vector<JumpEntry> sj((size_t)c->nt), qj((size_t)256 * c->ctas_per_slice);
        for (int t = 0; t < c->nt; ++t) sj[t] = jump_entry((u64)(p.slab_t0 + t) * (u64)c->vslice);
        for (size_t q = 0; q < qj.size(); ++q) qj[q] = jump_entry((u64)q * (u64)vec);
        CK(cudaMalloc((void **)&c->l_slice_jump, sizeof(JumpEntry) * sj.size()));
        CK(cudaMalloc((void **)&c->l_strip_jump, sizeof(JumpEntry) * qj.size()));
        CK(cudaMemcpy(c->l_slice_jump, sj.data(), sizeof(JumpEntry) * sj.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->l_strip_jump, qj.data(), sizeof(JumpEntry) * qj.size(), cudaMemcpyHostToDevice));
        if (c->march_ok) {
            const u64 L0 = (u64)p.dims[0], tpr = L0 / (u64)c->m_w, rows_per_cta = (256 / tpr) * (u64)c->m_R;
            std::vector<JumpEntry> cj((size_t)c->ctas_per_slice), tj(256);
            for (size_t b = 0; b < cj.size(); ++b) cj[b] = jump_entry((u64)b * rows_per_cta * L0);
            for (u64 t = 0; t < 256; ++t) tj[t] = jump_entry((t / tpr) * (u64)c->m_R * L0 + (t % tpr) * (u64)c->m_w);
            CK(cudaMalloc((void **)&c->l_cta_jump, sizeof(JumpEntry) * cj.size()));
            CK(cudaMalloc((void **)&c->l_thr_jump, sizeof(JumpEntry) * tj.size()));
            CK(cudaMemcpy(c->l_cta_jump, cj.data(), sizeof(JumpEntry) * cj.size(), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(c->l_thr_jump, tj.data(), sizeof(JumpEntry) * tj.size(), cudaMemcpyHostToDevice));
            if (c->tile_ok) {  // the tile kernel's per-thread table: same jump + where the strip lies in the staged tile
                const u64 L1 = (u64)p.dims[1], R = (u64)c->m_R, w = (u64)c->m_w, rowb = L0 * 4;
                const u64 seg_rows = rows_per_cta < L1 ? rows_per_cta : L1, seg_bytes = (seg_rows + 2) * rowb;
                const JumpEntry rowj = jump_entry(L0);
                std::vector<TileThread> tt(256);
                for (u64 t = 0; t < 256; ++t) {
                    const u64 tx = t % tpr, rt = (t / tpr) * R, x0 = tx * w, off = rt * L0 + x0;
                    const u64 sg = rt / seg_rows, s_row = 128 + sg * seg_bytes + (1 + rt - sg * seg_rows) * rowb;
                    TileThread &e = tt[t];
                    e = TileThread{};
                    e.a = tj[t].a; e.g0 = tj[t].g0; e.bg1 = tj[t].bg1;
                    e.ck_off = LCG_BETA * off * rowj.g0;
                    e.thr_off = (unsigned)off;
                    e.s_c = (unsigned)(s_row + x0 * 4);
                    e.s_left = (unsigned)(s_row + (x0 == 0 ? (L0 - 1) * 4 : x0 * 4 - 4));
                    e.s_right = (unsigned)(s_row + (x0 + w == L0 ? 0 : (x0 + w) * 4));
                    e.plane = (unsigned)(rt / L1);
                    e.row = (unsigned)rt;
                }
                CK(cudaMalloc((void **)&c->l_tile_ctr, 2 * sizeof(unsigned)));
                CK(cudaMemset(c->l_tile_ctr, 0, 2 * sizeof(unsigned)));
                CK(cudaMalloc((void **)&c->l_tile_thr, sizeof(TileThread) * tt.size()));
                CK(cudaMemcpy(c->l_tile_thr, tt.data(), sizeof(TileThread) * tt.size(), cudaMemcpyHostToDevice));
                // row-block kernel: thread (tx, ty) owns row ty of every pass; a stage is [halo | RPP rows | halo | t+1 | t-1 | x2+1 | x2-1]
                const JumpEntry passj = jump_entry(1024);
                for (u64 t = 0; t < 256; ++t) {
                    const u64 tx = t % tpr, ty = t / tpr, x0 = tx * w, off = ty * L0 + x0, s_row = (1 + ty) * rowb;
                    const JumpEntry j = jump_entry(off);
                    TileThread &e = tt[t];
                    e = TileThread{};
                    e.a = j.a; e.g0 = j.g0; e.bg1 = j.bg1;
                    e.ck_off = LCG_BETA * off * passj.g0;
                    e.thr_off = (unsigned)off;
                    e.s_c = (unsigned)(s_row + x0 * 4);
                    e.s_left = (unsigned)(s_row + (x0 == 0 ? (L0 - 1) * 4 : x0 * 4 - 4));
                    e.s_right = (unsigned)(s_row + (x0 + w == L0 ? 0 : (x0 + w) * 4));
                    e.row = (unsigned)ty;
                }
                CK(cudaMalloc((void **)&c->l_rows_thr, sizeof(TileThread) * tt.size()));
                CK(cudaMemcpy(c->l_rows_thr, tt.data(), sizeof(TileThread) * tt.size(), cudaMemcpyHostToDevice));
            }

        }
    }
    std::vector<u64> seeds((size_t)p.nchains);
    std::vector<double> m2((size_t)p.nchains, p.m2), lam((size_t)p.nchains, p.lambda);
    for (int k = 0; k < p.nchains; ++k) seeds[k] = seed + (u64)k;
    CK(cudaMemcpy(c->l_seeds[0], seeds.data(), sizeof(u64) * seeds.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->l_m2, m2.data(), sizeof(double) * m2.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->l_lam, lam.data(), sizeof(double) * lam.size(), cudaMemcpyHostToDevice));
    const u64 none = NO_EVENT;
    CK(cudaMemcpy(c->l_event, &none, sizeof(u64), cudaMemcpyHostToDevice));
    if (f0) {
        rc = sq_upload_field(c, 0, f0, SQ_REAL_F64);
        if (rc) return rc;
    }
    // the on-chip resident kernel: 2-D fp32 single chain, whole lattice, rows of 128..1024 sites
    {
        int coop = 0, sms = 0;
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p.device));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p.device));
        const int64_t L0 = p.dims[0], L1 = p.dims[1];
        const bool shape_ok = p.ndim == 2 && p.real == SQ_REAL_F32 && p.nchains == 1 && c->nt == Lt &&
                              L0 % 128 == 0 && L0 <= 1024 && !(p.flags & (SQ_FLAG_FORCE_STREAMING | SQ_FLAG_NO_OBSERVABLES));
        if (coop && shape_ok && sms > 0) {
            // bands of at least two rows (each edge row has ONE neighbour CTA), at most 896 threads per CTA
            const int nb = (int)std::min<int64_t>(sms, L1 / 2);
            const int rows = nb > 0 ? (int)((L1 + nb - 1) / nb) : 0;
            if (nb > 0 && rowres_strip((int)L0, rows) != 0) {
                CK(preload_lattice_step(p.real, p.math, p.ndim));  // the event-recovery path: see sq_lattice.cu
                c->res_ok = true;
                c->res_nb = nb;
                c->res_rows = rows;
                if ((rc = dalloc(&c->r_halo, 2 * (size_t)nb * 2 * (size_t)L0))) return rc;
                if ((rc = dalloc(&c->r_error, 1))) return rc;
                if ((rc = dalloc(&c->r_progress, (size_t)nb))) return rc;
                if ((rc = dalloc(&c->r_ckpt, (size_t)RES_NCKPT * (size_t)c->V))) return rc;
                // one allocation: hist_rows[RES_MAX_STEPS][L1], then hist_p2[RES_MAX_STEPS][L1]
                if ((rc = dalloc(&c->r_hist_rows, 2 * (size_t)RES_MAX_STEPS * (size_t)L1))) return rc;
                c->r_hist_p2 = c->r_hist_rows + (size_t)RES_MAX_STEPS * (size_t)L1;
                if ((rc = dalloc(&c->r_nclamp_slots, (size_t)RES_SLOTS))) return rc;
                if ((rc = dalloc(&c->r_step_sums, (size_t)RES_MAX_STEPS * 2))) return rc;
            }
        }
    }
    return SQ_OK;
}

extern "C" int sq_init(sq_ctx **out, const sq_params *p, const double *f0, const double *x0,
                       const double *xx0_0, double omega0, uint64_t seed) {
    if (!out || !p) return SQ_ERR_INVALID;
    *out = nullptr;
    if (p->struct_size != sizeof(sq_params)) return SQ_ERR_INVALID;
    const int ndev = sq_device_count();
    if (ndev <= 0) return SQ_ERR_NODEVICE;
    if (p->device < 0 || p->device >= ndev) return SQ_ERR_INVALID;
    sq_ctx *c = new (std::nothrow) sq_ctx();
    if (!c) return SQ_ERR_NOMEM;
    c->p = *p;
    int rc = SQ_OK;
    do {
        if ((rc = sq_set_dev(c))) break;
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            snprintf(g_cuda_err, 512, "cudaStreamCreate: %s", cudaGetErrorString(e));
            rc = SQ_ERR_CUDA;
            break;
        }
        e = cudaMallocHost(&c->h_pin, 4096);
        if (e != cudaSuccess) { rc = SQ_ERR_NOMEM; break; }
        c->h_jump.resize(JUMP_TABLE_ENTRIES);
        build_jump_table(c->h_jump.data());
        e = cudaMalloc((void **)&c->d_jump, sizeof(JumpEntry) * JUMP_TABLE_ENTRIES);
        if (e != cudaSuccess) { rc = SQ_ERR_NOMEM; break; }
        e = cudaMemcpy(c->d_jump, c->h_jump.data(), sizeof(JumpEntry) * JUMP_TABLE_ENTRIES, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { rc = SQ_ERR_CUDA; break; }
        if (p->kernel == SQ_KERNEL_COMPAT1D) rc = init_compat(c, f0, x0, xx0_0, omega0, seed);
        else if (p->kernel == SQ_KERNEL_LATTICE) rc = init_lattice(c, f0, seed);
        else rc = SQ_ERR_INVALID;
    } while (0);
    if (rc != SQ_OK) {
        sq_free(c);
        return rc;
    }
    *out = c;
    return SQ_OK;
}

// ------------------------------------------------------------------- stepping -------------
static int enqueue_compat(sq_ctx *c, double dtau, int nsteps, int64_t runs0, bool controller = false) {
    const sq_params &p = c->p;
    Compat1DArgs A{};
    if (controller) {
        A.ctl = c->c_ctl;
        A.log_rec = c->c_log_rec;
        A.log_xavg = c->c_log_xavg;
        A.log_cap = SQ_FRAMES_MAX;
        A.noise_c = p.noise_c;
    }
    A.N = (int)p.dims[0];
    A.loops = nsteps;
    A.potential = p.potential;
    A.runs = runs0;
    A.dt = p.spacing;
    A.dtau = dtau;
    A.dt2 = (double)((float)p.spacing * (float)p.spacing);
    A.nscale_site = p.noise_c * (double)sqrtf((float)(2. * dtau / p.spacing));
    A.nscale_omega = p.noise_c * (double)sqrtf((float)(2. * dtau));
    A.intconst = host_intconst(p.potential);
    const JumpEntry e = jump_entry((u64)A.N + 1);
    A.P = e.a & LCG_MASK;
    A.Q = (LCG_GAMMA * e.g0 + e.bg1) & LCG_MASK;
    A.jump = c->d_jump;
    A.f = c->c_f; A.x = c->c_x; A.xx0 = c->c_xx0;
    A.newf = c->c_newf; A.newx = c->c_newx; A.newxx0 = c->c_newxx0;
    A.omega = c->c_omega; A.seed = c->c_seed; A.stable = c->c_stable;
    A.lrgEl = c->c_lrgEl; A.lrgVl = c->c_lrgVl; A.steps_done = c->c_steps; A.nevents = c->c_nevents;
    if (c->timing) { int rt = sq_timing_mark(c); if (rt) return rt; }
    CK(launch_compat1d(A, c->stream));
    if (c->timing) { int rt = sq_timing_mark(c); if (rt) return rt; }
    c->launches++;
    return SQ_OK;
}

void sq_fill_rebase_inline(LatticeArgs &A, const RebaseEntry *e, int n) {
    for (int j = 0; j < RB_INLINE; ++j) {
        A.rb_gid[j] = (j < n && n <= RB_INLINE) ? e[j].gid_start : ~0ULL;
        A.rb_chain[j] = (j < n && n <= RB_INLINE) ? e[j].chain : -1;
    }
}

LatticeArgs sq_lattice_args(sq_ctx *c, double dtau, int k /* step in sequence */) {
    const sq_params &p = c->p;
    LatticeArgs A{};
    const int vec = (int)(16 / c->rsz);
    A.ndim = p.ndim;
    A.pot = p.potential;
    A.nt = c->nt;
    A.wrap_time = (c->nt == p.dims[p.ndim - 1]) ? 1 : 0;
    A.nchains = p.nchains;
    A.step_index = k;
    A.n_rebase = 0;
    A.strips_per_cta_iter = 256 * c->ctas_per_slice;
    {   // L2 blocking (sq_lattice.cu): one chunk = ~6 MB of each time level
        const double bytes_per_cta = (double)c->vslice * (double)c->rsz / (double)c->ctas_per_slice;
        static const double chunk_mb = getenv("SQ_CHUNK_MB") ? atof(getenv("SQ_CHUNK_MB")) : 6.0;  // tuning knob
        long long cpc = (long long)(chunk_mb * 1048576.0 / bytes_per_cta);
        if (cpc < 1) cpc = 1;
        if (cpc > c->ctas_per_slice) cpc = c->ctas_per_slice;
        A.ctas_per_chunk = (int)cpc;
    }
    for (int i = 0; i < 4; ++i) A.dim[i] = i < p.ndim ? p.dims[i] : 1;
    A.vslice = c->vslice;
    A.V = c->V;
    A.slab_t0 = p.slab_t0;
    A.chain_stride = c->vlocal;
    const int b = (c->cur + k) & 1;
    A.in = c->l_field[b];
    A.out = c->l_field[b ^ 1];
    A.ghost_lo = c->l_ghost[0];
    A.ghost_hi = c->l_ghost[1];
    const double a2f = (double)((float)p.spacing * (float)p.spacing);
    double ad = 1.;
    for (int i = 0; i < p.ndim; ++i) ad *= p.spacing;
    A.c_lap = (1. * dtau) / a2f;
    A.c_dt = dtau;
    A.nscale = p.noise_c * (double)sqrtf((float)(2. * dtau / ad));
    A.m2 = p.m2;
    A.lam = p.lambda;
    A.m2_chain = c->l_m2;
    A.lam_chain = c->l_lam;
    A.seed_in = c->l_seeds[b];
    A.seed_out = c->l_seeds[b ^ 1];
    A.k2_f = (float)(2.0 * 0.6931471805599453 * A.nscale * A.nscale);
    A.slice_jump = c->l_slice_jump;
    A.strip_jump = c->l_strip_jump;
    A.stride_jump = jump_entry((u64)A.strips_per_cta_iter * (u64)vec);
    A.vol_jump = jump_entry((u64)c->V);
    A.jump = c->d_jump;
    A.m_on = c->march_ok ? (c->tile_ok ? (c->rows_ok ? 3 : 2) : 1) : 0;
    A.t_dck = LCG_BETA * (u64)p.dims[0] * jump_entry((u64)p.dims[0]).g0;
    A.m_w = c->m_w;
    A.t_dc1 = (u64)(p.dims[0] - (c->m_w - 1)) * LCG_A;   // from a strip's last site to the next row's first
    A.t_dc2 = (u64)(p.dims[0] - (c->m_w - 1)) * LCG_BETA;
    A.m_R = c->m_R;
    A.m_tpr_log = c->m_tpr_log;
    A.cta_jump = c->l_cta_jump;
    A.thr_jump = c->l_thr_jump;
    A.tile_thr = c->l_tile_thr;
    A.tile_ctr = c->l_tile_ctr;
    A.rows_thr = c->l_rows_thr;
    A.prow_jump = jump_entry(1024);
    A.p_dck = LCG_BETA * 1024ULL * A.prow_jump.g0;
    A.p_dc1 = (1024ULL - 3) * LCG_A;
    A.p_dc2 = (1024ULL - 3) * LCG_BETA;
    A.row_jump = jump_entry((u64)p.dims[0]);
    A.rebase = c->l_rebase;
    sq_fill_rebase_inline(A, nullptr, 0);
    A.event_key = c->l_event;
    A.partials = (p.flags & SQ_FLAG_NO_OBSERVABLES) ? nullptr : c->l_partials;
    A.nclamped = c->l_nclamped;
    return A;
}

int sq_launch_update(sq_ctx *c, const LatticeArgs &A) {
    if (A.m_on >= 2 && A.n_rebase == 0) CK(launch_lattice_tile(A, c->p.math, c->ctas_per_slice, c->stream));  // (entries: marching kernel)
    else if (A.m_on) CK(launch_lattice_march(A, c->p.math, c->ctas_per_slice, c->stream));
    else CK(launch_lattice_step(A, c->p.real, c->p.math, c->ctas_per_slice, c->stream));
    return SQ_OK;
}

// One step of a launch sequence (k = 0, 1, ... since the last sq_join_finalize).  The update kernels of a group of M steps
// follow each other in `stream` with nothing in between (the tile kernel's programmatic dependent launch needs that);
// each writes its per-tile partials into its own buffer of the ring, and the group's finalizes are handed to the side
// stream together behind the group's last update.  Group g reuses the buffers of group g - 2 and waits for its finalizes.
static int flush_finalize_group(sq_ctx *c) {
    if (!c->fin_queued) return SQ_OK;
    const int pb = c->fin_pending & 1;
    CK(cudaEventRecord(c->ev_upd[pb], c->stream));
    CK(cudaStreamWaitEvent(c->fin_stream, c->ev_upd[pb], 0));
    for (int i = 0; i < c->fin_queued; ++i) {
        CK(launch_finalize(c->fin_queue[i], c->fin_stream));
        c->launches++;
    }
    CK(cudaEventRecord(c->ev_fin[pb], c->fin_stream));
    c->fin_queued = 0;
    c->fin_pending++;
    return SQ_OK;
}

int sq_enqueue_step(sq_ctx *c, LatticeArgs &A, FinalizeArgs &F, int k) {
    const int M = A.slab_on ? 1 : c->fin_batch;  // (the ring's finder and history need every step's sums at once)
    if (A.partials) {
        A.partials = c->l_partials + (size_t)(k % (2 * M)) * c->npart;
        F.partials = A.partials;
        // the finalizes that read this group's buffers two groups ago must be done before they are overwritten
        if (c->fin_queued == 0 && c->fin_pending > 1) CK(cudaStreamWaitEvent(c->stream, c->ev_fin[c->fin_pending & 1], 0));
    }
    if (c->timing) { int rt = sq_timing_mark(c); if (rt) return rt; }
    { int rl = sq_launch_update(c, A); if (rl) return rl; }
    if (c->timing) { int rt = sq_timing_mark(c); if (rt) return rt; }
    c->launches++;
    if (A.partials) {
        c->fin_queue[c->fin_queued++] = F;
        if (c->fin_queued >= M) return flush_finalize_group(c);
    }
    return SQ_OK;
}
int sq_join_finalize(sq_ctx *c) {
    { int rf = flush_finalize_group(c); if (rf) return rf; }
    if (c->fin_pending > 0) {
        CK(cudaStreamWaitEvent(c->stream, c->ev_fin[0], 0));
        if (c->fin_pending > 1) CK(cudaStreamWaitEvent(c->stream, c->ev_fin[1], 0));
        c->fin_pending = 0;
    }
    return SQ_OK;
}

// sq_frame_host's early read-back: `src` is what the batch's last update kernel (just enqueued on c->stream) writes
static int enqueue_spec_copy(sq_ctx *c, const void *src) {
    if (!c->spec_host) return SQ_OK;
    const size_t nb = (size_t)c->vlocal * c->rsz * (size_t)c->p.nchains;
    CK(cudaEventRecord(c->ev_copy, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
    CK(cudaMemcpyAsync(c->spec_host, src, nb, cudaMemcpyDeviceToHost, c->copy_stream));
    c->spec_src = src;
    c->spec_batch = c->batch_seq;
    c->spec_host = nullptr;  // this batch only
    return SQ_OK;
}

static int enqueue_lattice(sq_ctx *c, double dtau, int nsteps, int64_t runs0) {
    const sq_params &p = c->p;
    LatticeArgs A = sq_lattice_args(c, dtau, 0);
    const int Lt = (int)p.dims[p.ndim - 1];
    const int tmid = Lt / 2;
    for (int k = 0; k < nsteps; ++k) {
        const int b = (c->cur + k) & 1;
        A.step_index = k;
        A.in = c->l_field[b];
        A.out = c->l_field[b ^ 1];
        A.seed_in = c->l_seeds[b];
        A.seed_out = c->l_seeds[b ^ 1];
        A.n_rebase = (k == 0) ? (int)c->entries.size() : 0;
        sq_fill_rebase_inline(A, c->entries.data(), A.n_rebase);
        A.nclamped = c->l_nclamp_step + k;
        FinalizeArgs F{};
        F.nt = c->nt;
        F.nchains = p.nchains;
        F.ctas_per_slice = c->ctas_per_slice;
        F.tmid_local = (tmid >= p.slab_t0 && tmid < p.slab_t0 + c->nt) ? (int)(tmid - p.slab_t0) : -1;
        F.vslice = c->vslice;
        F.runs = runs0 + k;
        F.partials = c->l_partials;
        F.slice_sum = c->l_slice_sum;
        F.slice_x = c->l_slice_x;
        F.slice_xx0 = c->l_slice_xx0;
        F.sums = c->l_sums;
        F.sums_mean = c->l_sums_mean;
        F.history = nullptr;
        F.event_key = c->l_event;
        F.step_index = k;
        { int rs = sq_enqueue_step(c, A, F, k); if (rs) return rs; }
    }
    { int rs = enqueue_spec_copy(c, c->l_field[(c->cur + nsteps) & 1]); if (rs) return rs; }
    return sq_join_finalize(c);
}

static int enqueue_resident_welford(sq_ctx *c, int nsteps, int64_t runs0);
static int enqueue_resident(sq_ctx *c, double dtau, int nsteps, int64_t runs0) {
    const sq_params &p = c->p;
    const LatticeArgs L = sq_lattice_args(c, dtau, 0);
    ResidentArgs A{};
    A.L0 = (int)p.dims[0];
    A.L1 = (int)p.dims[1];
    A.nsteps = nsteps;
    A.pot = p.potential;
    A.step_index0 = 0;
    A.step0 = c->r_tag;
    c->r_tag += (unsigned)nsteps + 1u;
    A.V = c->V;
    A.in = (const float *)c->l_field[c->cur];
    A.out = (float *)c->l_field[c->cur ^ 1];
    A.halo_ll = c->r_halo;
    A.c_lap = L.c_lap;
    A.c_dt = L.c_dt;
    A.nscale = L.nscale;
    A.m2 = p.m2;
    A.lam = p.lambda;
    A.c_lap_f = (float)A.c_lap;
    A.c_dt_f = (float)A.c_dt;
    A.c_2dt_f = 2.0f * A.c_dt_f;
    A.m2_f = (float)A.m2;
    A.lam_f = (float)A.lam;
    A.k2_f = (float)(2.0 * 0.6931471805599453 * A.nscale * A.nscale);
    A.seed_in = c->l_seeds[c->cur];
    A.seed_out = c->l_seeds[c->cur ^ 1];
    const JumpEntry e = jump_entry((u64)c->V + 1);
    A.P = e.a & LCG_MASK;
    A.Q = (LCG_GAMMA * e.g0 + e.bg1) & LCG_MASK;
    A.vol_jump = L.vol_jump;
    A.jump = c->d_jump;
    A.event_key = c->l_event;
    A.hist_rows = c->r_hist_rows;
    A.hist_p2 = c->r_hist_p2;
    A.nclamped = c->l_nclamped;
    A.error_flag = c->r_error;
    A.one = 1u;
    for (int k = 0; k < 8; ++k) A.row_const[k] = (u64)k * (u64)p.dims[0] * LCG_A;
    A.ckpt = c->r_ckpt;
    A.progress = c->r_progress;
    // per-chain couplings live in device arrays for the streaming kernel; the resident kernel is
    // single-chain and takes them by value: keep both in sync through sq_set_chain (host mirror)
    if (c->timing) { int rt = sq_timing_mark(c); if (rt) return rt; }
    A.rows_max = c->res_rows;
    A.nclamp_slots = c->r_nclamp_slots;
    CK(launch_rowres(A, p.math, c->res_nb, c->stream));
    if (c->timing) { int rt = sq_timing_mark(c); if (rt) return rt; }
    c->launches++;
    { int rs = enqueue_spec_copy(c, c->l_field[c->cur ^ 1]); if (rs) return rs; }
    return enqueue_resident_welford(c, nsteps, runs0);
}

// history of a resident launch -> running means, for its first `nsteps` steps
static int enqueue_resident_welford(sq_ctx *c, int nsteps, int64_t runs0) {
    const sq_params &p = c->p;
    WelfordArgs W{};
    W.nt = c->nt;
    W.nsteps = nsteps;
    W.tmid = (int)(p.dims[1] / 2);
    W.np2 = (int)p.dims[1];  // partial sums of phi^2: per row
    W.vslice = c->vslice;
    W.runs = runs0;
    W.hist_rows = c->r_hist_rows;
    W.hist_p2 = c->r_hist_p2;
    W.slice_x = c->l_slice_x;
    W.slice_xx0 = c->l_slice_xx0;
    W.slice_sum = c->l_slice_sum;
    W.sums = c->l_sums;
    W.sums_mean = c->l_sums_mean;
    W.event_key = c->l_event;
    CK(launch_welford_history(W, c->r_step_sums, c->stream));
    c->launches += 2;
    {   // clamp hits of the checkpoint intervals that stand (none of them while an event is flagged)
        CK(launch_commit_clamps(c->r_nclamp_slots, (nsteps + RES_CKPT - 1) / RES_CKPT, RES_SLOTS, c->l_nclamped, c->l_event, c->stream));
        c->launches++;
    }
    return SQ_OK;
}

// enqueue the next batch of the pending sequence; sets pend_kind / pend_nsteps
static int enqueue_batch(sq_ctx *c, int remaining, int64_t runs0) {
    ++c->batch_seq;
    if (c->res_ok && c->entries.empty() && c->force_stream == 0) {
        int n = std::min(remaining, RES_MAX_STEPS);
        if (c->res_limit > 0) n = std::min(n, c->res_limit);
        c->spec_host = (n == remaining) ? c->spec_want : nullptr;  // the batch that completes the frame carries the read-back
        int rc = enqueue_resident(c, c->pend_dtau, n, runs0);
        if (rc) return rc;
        c->pend_kind = 1;
        c->pend_nsteps = n;
    } else {
        int n = std::min(remaining, MAX_SEQ_STEPS);
        if (c->force_stream > 0) n = std::min(n, c->force_stream);
        c->spec_host = (n == remaining) ? c->spec_want : nullptr;
        int rc = enqueue_lattice(c, c->pend_dtau, n, runs0);
        if (rc) return rc;
        c->pend_kind = 0;
        c->pend_nsteps = n;
    }
    c->spec_host = nullptr;
    return SQ_OK;
}

extern "C" int sq_step_async(sq_ctx *c, double dtau, int nsteps, int64_t runs0) {
    if (!c || nsteps < 0 || !(dtau > 0)) return SQ_ERR_INVALID;
    if (c->pending) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    c->pend_dtau = dtau;
    c->pend_nsteps = nsteps;
    c->pend_total = nsteps;
    c->pend_runs0 = runs0;
    if (c->p.kernel == SQ_KERNEL_COMPAT1D) {
        if (nsteps > 0 && (rc = enqueue_compat(c, dtau, nsteps, runs0))) return rc;
    } else if (c->slab) {
        if ((rc = sq_slab_enqueue(c, dtau, nsteps, runs0))) return rc;
    } else if (nsteps > 0) {
        if ((rc = enqueue_batch(c, nsteps, runs0))) return rc;
    }
    c->pending = true;
    return SQ_OK;
}

// seed (full u64) before the draw at gid g of the step whose start seed is S, under `entries`
u64 sq_host_seed_before(const sq_ctx *c, const std::vector<RebaseEntry> &entries, int chain, u64 S, u64 g) {
    u64 bg = 0, bs = S;
    for (const RebaseEntry &e : entries)
        if (e.chain == chain && e.gid_start <= g && e.gid_start >= bg) { bg = e.gid_start; bs = e.seed; }
    if (g == bg) return bs;
    // the draw at g-1 was event-free: seed = t2(g-1) - 2^31 as a full (wrapping) u64
    const u64 sp = (g - 1 == bg) ? bs : lcg_seed_at(bs, bg, g - 1 - bg, c->h_jump.data());
    u64 t1, t2;
    lcg_draw(sp, g - 1, t1, t2);
    return lcg_next_seed(t2);
}

// Finish the pending lattice sequence.  RNG events (inf-retry / `seed+=`, tau_kernel.cl:278-282)
// are speculated away on the device and replayed here:
//   streaming batch, event at step k: steps < k stand; the host replays the event draw literally,
//     appends a rebase entry and the sequence resumes at step k with the entry list;
//   resident batch (one launch = many steps, output written at the end): nothing stands; the
//     batch is re-run up to the event step, then ONE streaming step takes the event.
// the on-chip path's deferred entry becomes the entry list of the streaming step that is enqueued next
static int install_deferred(sq_ctx *c) {
    if (c->deferred.empty()) return SQ_OK;
    c->entries = c->deferred;
    c->deferred.clear();
    c->nevents += c->entries.size();
    CK(cudaMemcpy(c->l_rebase, c->entries.data(), sizeof(RebaseEntry) * c->entries.size(), cudaMemcpyHostToDevice));
    return SQ_OK;
}

static int sync_lattice(sq_ctx *c) {
    const int total = c->pend_total;
    int done = 0;
    int64_t runs0 = c->pend_runs0;
    while (total > 0) {
        // the event word (and the resident kernel's error flag) ride behind the kernels into pinned
        // memory: one synchronisation, no blocking pageable copies on the per-frame path
        u64 *pin_key = (u64 *)c->h_pin;
        unsigned *pin_err = (unsigned *)((char *)c->h_pin + 64);
        CK(cudaMemcpyAsync(pin_key, c->l_event, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        if (c->pend_kind == 1) CK(cudaMemcpyAsync(pin_err, c->r_error, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        u64 key = *pin_key;
        const int n = c->pend_nsteps;
        if (c->timing) {
            size_t valid = (size_t)-1;  // streaming: launches up to and including the event step ran in full
            if (key != NO_EVENT) valid = c->pend_kind == 1 ? 0 : (size_t)(key >> KEY_STEP_SHIFT) + 1;
            int rt = sq_timing_collect(c, valid);
            if (rt) return rt;
        }
        if (c->pend_kind == 1) {
            const unsigned err = *pin_err;
            if (err) return SQ_ERR_TIMEOUT;
            if (key == NO_EVENT) {
                c->cur ^= 1;  // one launch, one buffer flip
                c->ok_batch = c->batch_seq;
                done += n;
                runs0 += n;
                if (c->res_limit > 0) {  // the re-run has reached the event step: it is a streaming step, entry in hand
                    c->res_limit = 0;
                    c->force_stream = 1;
                    int ri = install_deferred(c);
                    if (ri) return ri;
                }
            } else {
                // The launch stopped early (sq_resident.cu): keep everything up to the last checkpoint
                // that EVERY CTA has written and that lies before the event, redo only the rest.
                int k = (int)(key >> KEY_STEP_SHIFT);
                std::vector<unsigned> prog((size_t)c->res_nb);
                CK(cudaMemcpy(prog.data(), c->r_progress, sizeof(unsigned) * prog.size(), cudaMemcpyDeviceToHost));
                const int reached = (int)*std::min_element(prog.begin(), prog.end());
                const int c0 = std::min(reached, k) / RES_CKPT * RES_CKPT;
                // The event's replay entry is resolved NOW: steps 0..k-1 of the launch are event-free as far as anyone has
                // looked (the key is the minimum over everything detected), so the start seed of step k follows from the
                // launch's by k whole-step advances, and the draw at the key's gid is replayed literally.  The streaming
                // step that takes the event then runs WITH its entry instead of being launched once just to find the
                // event again (one launch + one synchronisation less per event).  Should the re-run of c0..k-1 meet an
                // EARLIER event (in rows a CTA had not reached when it left), this branch runs again and overwrites the
                // entry; an earlier event inside step k itself is caught by the streaming step, which then drops it.
                u64 S0;  // the launch's start seed
                CK(cudaMemcpy(&S0, c->l_seeds[c->cur], sizeof(u64), cudaMemcpyDeviceToHost));
                const JumpEntry vj = jump_entry((u64)c->V);
                u64 S = S0, Sc0 = S0;
                for (int i = 0; i < k; ++i) {  // what the kernel's omega thread does, step by step
                    if (i == c0) Sc0 = S;
                    u64 t1, t2;
                    lcg_draw(lcg_apply(vj, S, 0) & LCG_MASK, (u64)c->V, t1, t2);
                    S = lcg_next_seed(t2);
                }
                if (c0 == k) Sc0 = S;
                {
                    const u64 g = key & ((1ULL << KEY_CHAIN_SHIFT) - 1);
                    const std::vector<RebaseEntry> none_yet;
                    const HostDraw h = host_draw_literal(sq_host_seed_before(c, none_yet, 0, S, g), g);
                    RebaseEntry e{};
                    e.gid_start = g + 1;
                    e.seed = h.seed_after;
                    e.ov_gid = g;
                    e.ov_t1 = h.t1;
                    e.ov_t2 = h.t2;
                    e.chain = 0;
                    e.vseed = virtual_start_seed(e.seed, e.gid_start, c->h_jump.data());
                    c->deferred.assign(1, e);
                }
                if (c0 > 0) {
                    const u64 none = NO_EVENT;
                    CK(cudaMemcpy(c->l_event, &none, sizeof(u64), cudaMemcpyHostToDevice));
                    key = NO_EVENT;  // (already cleared: skip the reset below)
                    CK(cudaMemcpyAsync(c->l_field[c->cur], c->r_ckpt + (size_t)((c0 / RES_CKPT) % RES_NCKPT) * (size_t)c->V,
                                       sizeof(float) * (size_t)c->V, cudaMemcpyDeviceToDevice, c->stream));
                    CK(cudaMemcpy(c->l_seeds[c->cur], &Sc0, sizeof(u64), cudaMemcpyHostToDevice));  // the seed c0 steps on
                    int rw = enqueue_resident_welford(c, c0, runs0);
                    if (rw) return rw;
                    done += c0;
                    runs0 += c0;
                    k -= c0;
                } else {  // nothing stands: drop the abandoned launch's clamp counts
                    CK(launch_commit_clamps(c->r_nclamp_slots, 0, RES_SLOTS, c->l_nclamped, nullptr, c->stream));
                    c->launches++;
                }
                if (k > 0) c->res_limit = k;
                else {
                    c->force_stream = 1;
                    int ri = install_deferred(c);
                    if (ri) return ri;
                }
            }
        } else {
            int ok = n;
            if (key != NO_EVENT) ok = (int)(key >> KEY_STEP_SHIFT);
            else c->ok_batch = c->batch_seq;
            // clamp hits of the steps that stand (the event step is redone and counted then)
            CK(launch_commit_clamps(c->l_nclamp_step, ok, n, c->l_nclamped, nullptr, c->stream));
            c->launches++;
            if (ok > 0) c->entries.clear();  // entries belonged to the batch's first step
            c->cur = (c->cur + ok) & 1;
            done += ok;
            runs0 += ok;
            if (c->force_stream > 0) c->force_stream = std::max(0, c->force_stream - ok);
            if (key != NO_EVENT) {
                const int chain = (int)((key >> KEY_CHAIN_SHIFT) & 0x3FFF);
                const u64 g = key & ((1ULL << KEY_CHAIN_SHIFT) - 1);
                u64 S;
                CK(cudaMemcpy(&S, c->l_seeds[c->cur] + chain, sizeof(u64), cudaMemcpyDeviceToHost));
                // (entries behind the new event were derived from a chain that this event changes: an entry installed ahead of
                // its step by the on-chip path can be overtaken by an earlier event of the same step)
                {
                    const size_t before = c->entries.size();
                    c->entries.erase(std::remove_if(c->entries.begin(), c->entries.end(),
                                                    [&](const RebaseEntry &x) { return x.chain == chain && x.gid_start > g; }),
                                     c->entries.end());
                    c->nevents -= before - c->entries.size();  // (they never happened in the chain that stands)
                }
                const u64 sfull = sq_host_seed_before(c, c->entries, chain, S, g);
                const HostDraw h = host_draw_literal(sfull, g);
                RebaseEntry e{};
                e.gid_start = g + 1;
                e.seed = h.seed_after;
                e.ov_gid = g;
                e.ov_t1 = h.t1;
                e.ov_t2 = h.t2;
                e.chain = chain;
                e.vseed = virtual_start_seed(e.seed, e.gid_start, c->h_jump.data());
                if ((int)c->entries.size() >= MAX_REBASE) return SQ_ERR_INVALID;
                c->entries.push_back(e);
                c->nevents++;
                CK(cudaMemcpy(c->l_rebase, c->entries.data(), sizeof(RebaseEntry) * c->entries.size(),
                              cudaMemcpyHostToDevice));
                if (c->force_stream == 0) c->force_stream = 1;  // the step with entries is a streaming step
            }
        }
        if (key != NO_EVENT) {
            const u64 none = NO_EVENT;
            CK(cudaMemcpy(c->l_event, &none, sizeof(u64), cudaMemcpyHostToDevice));
        }
        if (done >= total) break;
        int rc = enqueue_batch(c, total - done, runs0);
        if (rc) return rc;
    }
    c->entries.clear();
    c->deferred.clear();
    c->force_stream = 0;
    c->res_limit = 0;
    c->runs = runs0;
    c->last_stable = 1;
    c->last_steps = done;
    return SQ_OK;
}

extern "C" int sq_sync(sq_ctx *c, int *stable) {
    if (!c) return SQ_ERR_INVALID;
    if (!c->pending) {
        if (stable) *stable = c->last_stable;
        return SQ_OK;
    }
    int rc = sq_set_dev(c);
    if (rc) return rc;
    if (c->p.kernel == SQ_KERNEL_COMPAT1D) {
        CK(cudaStreamSynchronize(c->stream));
        if (c->timing) { int rt = sq_timing_collect(c, (size_t)-1); if (rt) return rt; }
        if (c->pend_nsteps > 0) {
            int st = 1, steps = 0;
            CK(cudaMemcpy(&st, c->c_stable, sizeof(int), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&steps, c->c_steps, sizeof(int), cudaMemcpyDeviceToHost));
            c->last_stable = st;
            c->last_steps = steps;
            if (st == 1) {
                c->runs = c->pend_runs0 + c->pend_nsteps;
            } else {
                const int one = 1;  // tauhost.c:542-544: the host re-arms the flag
                CK(cudaMemcpy(c->c_stable, &one, sizeof(int), cudaMemcpyHostToDevice));
            }
        }
    } else {
        rc = c->slab ? sq_slab_finish(c) : sync_lattice(c);
        if (rc) { c->pending = false; return rc; }
    }
    c->pending = false;
    if (stable) *stable = c->last_stable;
    return SQ_OK;
}

extern "C" int sq_step(sq_ctx *c, double dtau, int nsteps, int64_t runs0, int *stable) {
    int rc = sq_step_async(c, dtau, nsteps, runs0);
    if (rc) return rc;
    return sq_sync(c, stable);
}

// ------------------------------------------------------------------- measurement ----------
extern "C" int sq_measure(sq_ctx *c, sq_obs *o) {
    if (!c || !o || o->struct_size != sizeof(sq_obs)) return SQ_ERR_INVALID;
    if (c->pending) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    const sq_params &p = c->p;
    o->runs = c->runs;
    o->stable = c->last_stable;
    o->steps_done = c->last_steps;
    if (p.kernel == SQ_KERNEL_COMPAT1D) {
        const int N = (int)p.dims[0];
        const size_t nb = sizeof(double) * (size_t)N;
        CK(launch_compat_reduce(c->c_f, c->c_x, c->c_xx0, c->c_omega, N, p.spacing, p.potential, c->c_red, c->stream));
        c->launches++;
        if (o->f) CK(cudaMemcpyAsync(o->f, c->c_f, nb, cudaMemcpyDeviceToHost, c->stream));
        if (o->x) CK(cudaMemcpyAsync(o->x, c->c_x, nb, cudaMemcpyDeviceToHost, c->stream));
        if (o->xx0) CK(cudaMemcpyAsync(o->xx0, c->c_xx0, nb, cudaMemcpyDeviceToHost, c->stream));
        if (o->corr) CK(cudaMemcpyAsync(o->corr, c->c_red + 8, nb, cudaMemcpyDeviceToHost, c->stream));
        double *pin = (double *)c->h_pin;
        CK(cudaMemcpyAsync(pin, c->c_red, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(pin + 2, c->c_omega, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(pin + 3, c->c_lrgVl, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(pin + 4, c->c_seed, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(pin + 5, c->c_lrgEl, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(pin + 6, c->c_nevents, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        o->mean_phi = pin[0] / N;
        o->mean_phi2 = pin[1] / N;
        o->omega = pin[2];
        o->lrgVl = pin[3];
        memcpy(&o->seed, pin + 4, sizeof(u64));
        memcpy(&o->lrgEl, pin + 5, sizeof(int));
        memcpy(&o->nevents, pin + 6, sizeof(u64));
        o->nclamped = 0;
        if (o->slice_x && o->x) memcpy(o->slice_x, o->x, nb);
        if (o->slice_xx0 && o->xx0) memcpy(o->slice_xx0, o->xx0, nb);
        return SQ_OK;
    }
    // lattice
    const int nt = c->nt;
    CK(launch_reduce_field(c->l_field[c->cur], p.real, c->vlocal, p.nchains, c->l_redbuf, c->stream));
    c->launches++;
    // all read-backs go through one pinned scratch buffer: truly asynchronous copies, one synchronisation
    // (pageable destinations would make each of the five copies a blocking staged transfer)
    const size_t need = sizeof(double) * ((size_t)REDUCE_BLOCKS * 2 + 2 * (size_t)nt + 2);
    if (need > c->h_pin2_bytes) {
        if (c->h_pin2) CK(cudaFreeHost(c->h_pin2));
        c->h_pin2 = nullptr;
        c->h_pin2_bytes = 0;
        CK(cudaMallocHost(&c->h_pin2, need));
        c->h_pin2_bytes = need;
    }
    double *part = (double *)c->h_pin2, *sxp = part + (size_t)REDUCE_BLOCKS * 2, *sxxp = sxp + nt;
    unsigned long long *tail = (unsigned long long *)(sxxp + nt);  // [0] = nclamped, [1] = seed
    CK(cudaMemcpyAsync(part, c->l_redbuf, sizeof(double) * (size_t)REDUCE_BLOCKS * 2, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sxp, c->l_slice_x, sizeof(double) * nt, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sxxp, c->l_slice_xx0, sizeof(double) * nt, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(tail, c->l_nclamped, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(tail + 1, c->l_seeds[c->cur], sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const unsigned long long ncl = tail[0];
    const u64 seed0 = tail[1];
    const std::vector<double> sx(sxp, sxp + nt), sxx(sxxp, sxxp + nt);
    double s1 = 0, s2 = 0;
    for (int k = 0; k < REDUCE_BLOCKS; ++k) { s1 += part[2 * k]; s2 += part[2 * k + 1]; }
    o->mean_phi = s1 / (double)c->vlocal;
    o->mean_phi2 = s2 / (double)c->vlocal;
    o->seed = seed0;
    o->nclamped = (int64_t)ncl;
    o->nevents = c->nevents;
    o->omega = 0;
    o->lrgEl = 0;
    o->lrgVl = 0;
    if (o->slice_x) memcpy(o->slice_x, sx.data(), sizeof(double) * nt);
    if (o->slice_xx0) memcpy(o->slice_xx0, sxx.data(), sizeof(double) * nt);
    if (o->corr) {
        const int64_t Lt = p.dims[p.ndim - 1];
        const int64_t tm = Lt / 2 - p.slab_t0;  // the host's xavg, tauhost.c:519-521
        const double xm = (tm >= 0 && tm < nt) ? sx[(size_t)tm] : 0.;
        for (int t = 0; t < nt; ++t) o->corr[t] = sxx[t] - sx[t] * xm;
    }
    if (c->slab) sq_slab_measure(c, o);  // joined ring: running means and the seed live on the host
    return SQ_OK;
}

extern "C" int sq_measure_chains(sq_ctx *c, double *mean_phi, double *mean_phi2, uint64_t *seeds) {
    if (!c || c->p.kernel != SQ_KERNEL_LATTICE || c->pending) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    const int nc = c->p.nchains;
    double *buf = nullptr;
    CK(cudaMalloc((void **)&buf, sizeof(double) * (size_t)nc * REDUCE_BLOCKS * 2));
    cudaError_t e = launch_reduce_field(c->l_field[c->cur], c->p.real, c->vlocal, nc, buf, c->stream);
    c->launches++;
    std::vector<double> part((size_t)nc * REDUCE_BLOCKS * 2);
    if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), buf, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && seeds) e = cudaMemcpyAsync(seeds, c->l_seeds[c->cur], sizeof(u64) * nc, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(buf);
    CK(e);
    for (int k = 0; k < nc; ++k) {
        double s1 = 0, s2 = 0;
        for (int b = 0; b < REDUCE_BLOCKS; ++b) {
            s1 += part[((size_t)k * REDUCE_BLOCKS + b) * 2];
            s2 += part[((size_t)k * REDUCE_BLOCKS + b) * 2 + 1];
        }
        if (mean_phi) mean_phi[k] = s1 / (double)c->vlocal;
        if (mean_phi2) mean_phi2[k] = s2 / (double)c->vlocal;
    }
    return SQ_OK;
}

// ------------------------------------------------------------------- field transfer -------
static int field_xfer(sq_ctx *c, int chain, void *host, int real, bool upload) {
    if (!c || !host || c->p.kernel != SQ_KERNEL_LATTICE || c->pending) return SQ_ERR_INVALID;
    if (chain < 0 || chain >= c->p.nchains || (real != SQ_REAL_F32 && real != SQ_REAL_F64)) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    char *dev = (char *)c->l_field[c->cur] + (size_t)chain * c->vlocal * c->rsz;
    const size_t hsz = real == SQ_REAL_F32 ? 4 : 8;
    if (real == c->p.real) {
        if (upload) CK(cudaMemcpyAsync(dev, host, (size_t)c->vlocal * hsz, cudaMemcpyHostToDevice, c->stream));
        else CK(cudaMemcpyAsync(host, dev, (size_t)c->vlocal * hsz, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return SQ_OK;
    }
    void *tmp = nullptr;
    CK(cudaMalloc(&tmp, (size_t)c->vlocal * hsz));
    cudaError_t e;
    if (upload) {
        e = cudaMemcpyAsync(tmp, host, (size_t)c->vlocal * hsz, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = launch_convert(tmp, real, dev, c->p.real, c->vlocal, c->stream);
    } else {
        e = launch_convert(dev, c->p.real, tmp, real, c->vlocal, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(host, tmp, (size_t)c->vlocal * hsz, cudaMemcpyDeviceToHost, c->stream);
    }
    c->launches++;
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    CK(e);
    return SQ_OK;
}
extern "C" int sq_upload_field(sq_ctx *c, int chain, const void *host, int real) {
    return field_xfer(c, chain, const_cast<void *>(host), real, true);
}
extern "C" int sq_download_field(sq_ctx *c, int chain, void *host, int real) {
    return field_xfer(c, chain, host, real, false);
}

extern "C" int sq_set_chain(sq_ctx *c, int chain, uint64_t seed, double m2, double lambda) {
    if (!c || c->p.kernel != SQ_KERNEL_LATTICE || c->pending) return SQ_ERR_INVALID;
    if (chain < 0 || chain >= c->p.nchains) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    CK(cudaMemcpy(c->l_seeds[c->cur] + chain, &seed, sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->l_m2 + chain, &m2, sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->l_lam + chain, &lambda, sizeof(double), cudaMemcpyHostToDevice));
    if (chain == 0) {  // host mirror: the resident kernel takes the couplings by value
        c->p.m2 = m2;
        c->p.lambda = lambda;
    }
    return SQ_OK;
}

// ------------------------------------------------------------------- end-to-end frame -----
extern "C" int sq_frame_host(sq_ctx *c, const void *host_in, void *host_out, int real, double dtau, int nsteps,
                             int64_t runs0, sq_obs *obs, int *stable) {
    if (!c || c->pending) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    if (c->p.kernel == SQ_KERNEL_COMPAT1D) {
        if (real != SQ_REAL_F64) return SQ_ERR_INVALID;
        const size_t nb = sizeof(double) * (size_t)c->p.dims[0];
        // tauhost.c:550: the host re-uploads f every frame
        if (host_in) CK(cudaMemcpyAsync(c->c_f, host_in, nb, cudaMemcpyHostToDevice, c->stream));
        if ((rc = sq_step(c, dtau, nsteps, runs0, stable))) return rc;
        if (host_out) {
            CK(cudaMemcpyAsync(host_out, c->c_f, nb, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
    } else {
        if (real != c->p.real) return SQ_ERR_INVALID;
        const size_t nb = (size_t)c->vlocal * c->rsz * (size_t)c->p.nchains;
        if (host_in) CK(cudaMemcpyAsync(c->l_field[c->cur], host_in, nb, cudaMemcpyHostToDevice, c->stream));
        // the read-back rides on its own stream behind the update kernels of the batch that completes the frame
        // (enqueue_batch / enqueue_spec_copy); it stands if that batch ran to its end
        c->spec_src = nullptr;
        // (small fields only: when an RNG event makes the batch stop early the copy already enqueued is wasted and the real one
        // queues behind it -- 80 us for 1024^2, but 1.3 ms for 64^4 and 43 ms for 256^3 x 32, whose frames meet events often:
        // measured 196 -> 177 and 57 -> 41 G site-updates/s end to end with the early copy, 465 -> 478 for 1024^2)
        c->spec_want = (host_out && !c->slab && nsteps > 0 && nb <= (size_t)8 << 20) ? host_out : nullptr;
        rc = sq_step_async(c, dtau, nsteps, runs0);
        if (!rc) rc = sq_sync(c, stable);
        c->spec_want = nullptr;
        CK(cudaStreamSynchronize(c->copy_stream));
        if (rc) return rc;
        if (host_out && !(c->spec_src == (const void *)c->l_field[c->cur] && c->spec_batch == c->ok_batch)) {
            CK(cudaMemcpyAsync(host_out, c->l_field[c->cur], nb, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
    }
    if (obs) return sq_measure(c, obs);
    return SQ_OK;
}

// ------------------------------------------------------------------- parity hook ----------
extern "C" int sq_debug_draws(sq_ctx *c, int chain, uint64_t gid0, uint64_t n, uint64_t *t1, uint64_t *t2) {
    if (!c || c->pending || !t1 || !t2) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    u64 seed = 0;
    if (c->p.kernel == SQ_KERNEL_COMPAT1D) {
        CK(cudaMemcpy(&seed, c->c_seed, sizeof(u64), cudaMemcpyDeviceToHost));
    } else {
        if (chain < 0 || chain >= c->p.nchains) return SQ_ERR_INVALID;
        CK(cudaMemcpy(&seed, c->l_seeds[c->cur] + chain, sizeof(u64), cudaMemcpyDeviceToHost));
    }
    u64 *d1 = nullptr, *d2 = nullptr;
    CK(cudaMalloc((void **)&d1, sizeof(u64) * n));
    cudaError_t e = cudaMalloc((void **)&d2, sizeof(u64) * n);
    if (e == cudaSuccess) e = launch_debug_draws(seed, gid0, n, c->d_jump, d1, d2, c->stream);
    c->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(t1, d1, sizeof(u64) * n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t2, d2, sizeof(u64) * n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d1);
    if (d2) cudaFree(d2);
    CK(e);
    return SQ_OK;
}

extern "C" int sq_kernel_timing(sq_ctx *c, int enable) {
    if (!c || c->pending) return SQ_ERR_INVALID;
    c->timing = enable != 0;
    c->timing_ms = 0;
    c->timing_launches = 0;
    c->ev_used = 0;
    return SQ_OK;
}
extern "C" int sq_kernel_time(sq_ctx *c, double *ms_total, int64_t *launches) {
    if (!c) return SQ_ERR_INVALID;
    if (ms_total) *ms_total = c->timing_ms;
    if (launches) *launches = c->timing_launches;
    return SQ_OK;
}

extern "C" uint64_t sq_lcg_jump(uint64_t seed, uint64_t gid0, uint64_t ndraws) {
    return lcg_seed_at(seed, gid0, ndraws, host_jump_table());
}

extern "C" int sq_slab_stats(sq_ctx *c, uint64_t *finder_scans, uint64_t *agree_rounds) {
    if (!c) return SQ_ERR_INVALID;
    sq_slab_stats_impl(c, finder_scans, agree_rounds);
    return SQ_OK;
}

// ------------------------------------------------------------------- exact resume (f-1) ------
extern "C" int sq_compat_get_state(sq_ctx *c, sq_compat_state *o) {
    if (!c || !o || o->struct_size != sizeof(sq_compat_state) || c->p.kernel != SQ_KERNEL_COMPAT1D || c->pending) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(&o->seed, c->c_seed, sizeof(u64), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&o->lrgEl, c->c_lrgEl, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&o->lrgVl, c->c_lrgVl, sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&o->omega, c->c_omega, sizeof(double), cudaMemcpyDeviceToHost));
    if (o->lrgEl < 0 || o->lrgEl >= c->p.dims[0]) return SQ_ERR_INTERNAL;
    CK(cudaMemcpy(&o->newf_lrgEl, c->c_newf + o->lrgEl, sizeof(double), cudaMemcpyDeviceToHost));
    return SQ_OK;
}
extern "C" int sq_compat_set_state(sq_ctx *c, const sq_compat_state *s) {
    if (!c || !s || s->struct_size != sizeof(sq_compat_state) || c->p.kernel != SQ_KERNEL_COMPAT1D || c->pending) return SQ_ERR_INVALID;
    if (s->lrgEl < 0 || s->lrgEl >= c->p.dims[0]) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(c->c_seed, &s->seed, sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->c_lrgEl, &s->lrgEl, sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->c_lrgVl, &s->lrgVl, sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->c_omega, &s->omega, sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->c_newf + s->lrgEl, &s->newf_lrgEl, sizeof(double), cudaMemcpyHostToDevice));
    return SQ_OK;
}

// ------------------------------------------------------------------- frame controller (f-3) --
extern "C" int sq_controller_set(sq_ctx *c, double dtau, int64_t runs, int stab_cnt) {
    if (!c || c->p.kernel != SQ_KERNEL_COMPAT1D || c->pending || !(dtau > 0)) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    Compat1DCtl h{};
    h.dtau = dtau;
    h.runs = runs;
    h.stab_cnt = stab_cnt;
    h.frame = 0;
    c->c_frames_done = 0;
    CK(cudaMemcpy(c->c_ctl, &h, sizeof h, cudaMemcpyHostToDevice));
    return SQ_OK;
}
extern "C" int sq_controller_get(sq_ctx *c, double *dtau, int64_t *runs, int *stab_cnt) {
    if (!c || c->p.kernel != SQ_KERNEL_COMPAT1D || c->pending) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    Compat1DCtl h{};
    CK(cudaMemcpy(&h, c->c_ctl, sizeof h, cudaMemcpyDeviceToHost));
    if (dtau) *dtau = h.dtau;
    if (runs) *runs = h.runs;
    if (stab_cnt) *stab_cnt = h.stab_cnt;
    return SQ_OK;
}
extern "C" int sq_frames(sq_ctx *c, int nframes, int nsteps, sq_frame_rec *recs, double *xavg) {
    if (!c || c->p.kernel != SQ_KERNEL_COMPAT1D || c->pending || !recs) return SQ_ERR_INVALID;
    if (nframes < 0 || nframes > SQ_FRAMES_MAX || nsteps < 1) return SQ_ERR_INVALID;
    int rc = sq_set_dev(c);
    if (rc) return rc;
    const int N = (int)c->p.dims[0];
    for (int k = 0; k < nframes; ++k)
        if ((rc = enqueue_compat(c, 1.0 /* unused: the controller block holds dtau */, nsteps, 0, true))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    if (c->timing && (rc = sq_timing_collect(c, (size_t)-1))) return rc;
    std::vector<Compat1DFrameRec> log((size_t)SQ_FRAMES_MAX);
    CK(cudaMemcpy(log.data(), c->c_log_rec, sizeof(Compat1DFrameRec) * log.size(), cudaMemcpyDeviceToHost));
    std::vector<double> xlog;
    if (xavg && nframes > 0) {  // the whole log in one copy
        xlog.resize((size_t)SQ_FRAMES_MAX * (size_t)N);
        CK(cudaMemcpy(xlog.data(), c->c_log_xavg, sizeof(double) * xlog.size(), cudaMemcpyDeviceToHost));
    }
    for (int k = 0; k < nframes; ++k) {
        const int slot = (int)((c->c_frames_done + k) % SQ_FRAMES_MAX);
        recs[k].dtau = log[(size_t)slot].dtau;
        recs[k].stable = log[(size_t)slot].stable;
        recs[k].steps = log[(size_t)slot].steps;
        if (xavg && log[(size_t)slot].stable)
            memcpy(xavg + (size_t)k * N, xlog.data() + (size_t)slot * N, sizeof(double) * (size_t)N);
    }
    c->c_frames_done += nframes;
    Compat1DCtl h{};
    CK(cudaMemcpy(&h, c->c_ctl, sizeof h, cudaMemcpyDeviceToHost));
    c->runs = h.runs;
    if (nframes > 0) {
        c->last_stable = recs[nframes - 1].stable;
        c->last_steps = recs[nframes - 1].steps;
    }
    return SQ_OK;
}
