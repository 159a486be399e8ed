// sq_session.cu -- shared-memory rendezvous of the ranks of one slab ring (see sq_session.h).
// Host code only; it lives in a .cu file because the whole library is built by one nvcc rule.
#include "sq_session.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <new>

#include "../../include/sq.h"

using namespace sq;

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

extern "C" int sq_session_open(sq_session **out, const char *name, int rank, int nranks) {
    if (!out || !name || nranks < 1 || nranks > SESSION_MAX_RANKS || rank < 0 || rank >= nranks) return SQ_ERR_INVALID;
    *out = nullptr;
    if (strlen(name) == 0 || strlen(name) > 100 || strchr(name, '/')) return SQ_ERR_INVALID;
    sq_session *s = new (std::nothrow) sq_session();
    if (!s) return SQ_ERR_NOMEM;
    snprintf(s->name, sizeof s->name, "/sq_%s", name);
    s->rank = rank;
    s->nranks = nranks;
    if (const char *t = getenv("SQ_SESSION_TIMEOUT")) {  // seconds a barrier waits for the other ranks
        const double v = atof(t);
        if (v > 0) s->timeout_s = v;
    }
    // every rank may be the creator: a fresh segment is zero-filled, which is the initial state
    s->fd = shm_open(s->name, O_CREAT | O_RDWR, 0600);
    if (s->fd < 0 || ftruncate(s->fd, (off_t)sizeof(SessionShm)) != 0) {
        if (s->fd >= 0) close(s->fd);
        delete s;
        return SQ_ERR_NOMEM;
    }
    void *p = mmap(nullptr, sizeof(SessionShm), PROT_READ | PROT_WRITE, MAP_SHARED, s->fd, 0);
    if (p == MAP_FAILED) {
        close(s->fd);
        delete s;
        return SQ_ERR_NOMEM;
    }
    s->shm = (SessionShm *)p;
    uint32_t expect = 0;
    timespec rt;
    clock_gettime(CLOCK_REALTIME, &rt);
    const uint64_t now_ns = (uint64_t)rt.tv_sec * 1000000000ull + (uint64_t)rt.tv_nsec;
    bool bad = false;
    if (s->shm->nranks.compare_exchange_strong(expect, (uint32_t)nranks)) {
        s->shm->created_ns.store(now_ns, std::memory_order_release);  // this rank created the ring
    } else {
        bad = expect != (uint32_t)nranks;  // ranks disagree on the ring size
        // A live ring's name disappears as soon as every rank has attached.  A segment that a rank has
        // already left, that was aborted, or that is older than a barrier would wait is the debris of a
        // crashed open: its barrier counters cannot be trusted.
        uint64_t born = 0;
        for (int spin = 0; spin < 2000 && !(born = s->shm->created_ns.load(std::memory_order_acquire)); ++spin) usleep(100);
        bad |= s->shm->closed.load(std::memory_order_acquire) != 0 || s->shm->abort_flag.load(std::memory_order_acquire) != 0;
        bad |= born == 0 || (now_ns > born && (double)(now_ns - born) * 1e-9 > s->timeout_s);
    }
    if (bad) {
        munmap(p, sizeof(SessionShm));
        close(s->fd);
        delete s;
        return SQ_ERR_INVALID;
    }
    *out = s;
    // everybody is attached after this barrier: the name can go, the mapping stays
    int rc = sq_session_barrier(s);
    if (rc == SQ_OK && rank == 0) {
        shm_unlink(s->name);
        s->unlinked = true;
    }
    if (rc != SQ_OK) {
        sq_session_close(s);
        *out = nullptr;
    }
    return rc;
}

extern "C" int sq_session_barrier(sq_session *s) {
    if (!s || !s->shm) return SQ_ERR_INVALID;
    SessionShm *m = s->shm;
    if (m->abort_flag.load(std::memory_order_acquire)) return SQ_ERR_TIMEOUT;
    const uint32_t gen = m->generation.load(std::memory_order_acquire);
    if (m->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)s->nranks) {
        m->arrived.store(0, std::memory_order_relaxed);
        m->generation.store(gen + 1, std::memory_order_release);
        return SQ_OK;
    }
    const double t0 = now_s();
    unsigned spins = 0;
    while (m->generation.load(std::memory_order_acquire) == gen) {
        if (m->abort_flag.load(std::memory_order_acquire)) return SQ_ERR_TIMEOUT;
        if ((++spins & 1023u) == 0) {
            sched_yield();
            if (now_s() - t0 > s->timeout_s) {  // a rank died: never hang the others
                m->abort_flag.store(1, std::memory_order_release);
                return SQ_ERR_TIMEOUT;
            }
        }
    }
    return SQ_OK;
}

extern "C" int sq_session_allgather_u64(sq_session *s, const uint64_t *in, int n, uint64_t *out) {
    if (!s || !s->shm || !in || !out || n < 1 || n > SESSION_WORDS) return SQ_ERR_INVALID;
    const unsigned par = s->xchg++ & 1u;  // a rank can be at most one exchange ahead of the slowest
    for (int k = 0; k < n; ++k) s->shm->slot[s->rank].words[par][k] = in[k];
    int rc = sq_session_barrier(s);
    if (rc) return rc;
    for (int r = 0; r < s->nranks; ++r)
        for (int k = 0; k < n; ++k) out[(size_t)r * n + k] = s->shm->slot[r].words[par][k];
    return SQ_OK;
}

extern "C" int sq_session_allgather_f64(sq_session *s, const double *in, int n, double *out) {
    if (!s || !s->shm || !in || !out || n < 1 || n > SESSION_SERIES_MAX) return SQ_ERR_INVALID;
    memcpy(s->shm->slot[s->rank].series, in, sizeof(double) * (size_t)n);
    int rc = sq_session_barrier(s);
    if (rc) return rc;
    for (int r = 0; r < s->nranks; ++r) memcpy(out + (size_t)r * n, s->shm->slot[r].series, sizeof(double) * (size_t)n);
    return sq_session_barrier(s);  // the single series buffer may be rewritten after this
}

extern "C" void sq_session_abort(sq_session *s) {
    if (s && s->shm) s->shm->abort_flag.store(1, std::memory_order_release);
}

extern "C" int sq_session_rank(const sq_session *s) { return s ? s->rank : -1; }
extern "C" int sq_session_size(const sq_session *s) { return s ? s->nranks : 0; }

extern "C" void sq_session_close(sq_session *s) {
    if (!s) return;
    if (s->shm) {
        s->shm->closed.fetch_add(1, std::memory_order_acq_rel);
        munmap(s->shm, sizeof(SessionShm));
    }
    if (s->fd >= 0) close(s->fd);
    // only when open failed before the unlink: afterwards the name may belong to a NEW ring
    if (s->rank == 0 && !s->unlinked) shm_unlink(s->name);
    delete s;
}
