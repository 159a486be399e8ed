// sq_session.h -- host-side rendezvous of the ranks (one process or thread per GPU) that share one
// slab-decomposed lattice on ONE box: a POSIX shared-memory segment with a sense-reversing barrier,
// a double-buffered mailbox for small all-gathers, and per-rank records for the CUDA IPC handles.
// No GPU is needed for any of it (CPU tests drive it with two processes).
//
// The reference has no multi-device code (single context, single queue,
// tauhost.c:249-252); this is the control plane of the decomposition SURVEY.md 8(e) asks for.  The
// data plane (halo slices) never goes through here: it moves GPU-to-GPU over NVLink (sq_slab.cu).
#pragma once
#include <stdint.h>

#include <atomic>

namespace sq {

constexpr int SESSION_MAX_RANKS = 64;
constexpr int SESSION_SERIES_MAX = 3 * 1024;  // doubles per rank and exchange
constexpr int SESSION_WORDS = 8;              // u64 per rank and mailbox

struct alignas(64) SessionRankSlot {
    uint64_t words[2][SESSION_WORDS];  // mailbox, double-buffered by exchange parity
    unsigned char ipc[64];             // cudaIpcMemHandle_t of the rank's halo arena
    uint64_t pid, raw_ptr;             // same-process rings pass the pointer itself
    int32_t device, pad;
    double series[SESSION_SERIES_MAX];
};

struct SessionShm {
    std::atomic<uint32_t> nranks;      // 0 until the first rank arrives
    std::atomic<uint32_t> arrived;     // barrier: ranks that reached the current generation
    std::atomic<uint32_t> generation;  // barrier sense
    std::atomic<uint32_t> abort_flag;  // raised by any rank that fails: wakes every waiter
    std::atomic<uint32_t> closed;      // ranks that left
    uint32_t pad0;
    std::atomic<uint64_t> created_ns;  // CLOCK_REALTIME of the creating rank: a segment older than the barrier
                                       // timeout that is still linked was left behind by a crashed open
    uint32_t pad[8];
    SessionRankSlot slot[SESSION_MAX_RANKS];
};

}  // namespace sq

// the opaque handle of include/sq.h
struct sq_session {
    sq::SessionShm *shm = nullptr;
    int rank = 0, nranks = 1;
    int fd = -1;
    unsigned xchg = 0;  // exchange counter (mailbox parity)
    bool unlinked = false;  // rank 0: the name was removed after the open barrier
    char name[128] = "";
    double timeout_s = 120.0;
};
