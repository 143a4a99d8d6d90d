// qb_comm_shm.cu -- "ranks share a GPU" transport: POSIX shared memory control plane + CUDA IPC data plane.
// See qb_comm_shm.cuh for why it exists.  Every wait is bounded (QUEST_B200_SHM_TIMEOUT_S, default 300 s) and a rank
// that fails raises an abort word in the segment, so the surviving ranks report an error instead of hanging.
// Throughput is irrelevant here (it is a correctness vehicle); what matters is that it is the SAME C ABI the NCCL
// transport serves, so the sharding shim above cannot tell the difference.
#include "qb_comm_shm.cuh"
#include <atomic>
#include <fcntl.h>
#include <sched.h>
#include <string.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>

#define SHM_MAX_RANKS 64
#define SHM_SLOT_BYTES ((size_t)1 << 20)          // per-rank host mail slot
#define SHM_STAGE_AMPS ((qindex)1 << 20)          // per-rank device staging buffer (amplitudes)
static const char SHM_MAGIC[8] = {'Q', 'B', 'S', 'H', 'M', 'v', '1', 0};

typedef std::atomic<unsigned long long> au64;

struct ShmHeader {
    au64 arrived;                                  // ranks that have mapped the segment
    au64 barrierCount, barrierGen;
    au64 abortWord;
    au64 seq[SHM_MAX_RANKS][SHM_MAX_RANKS];        // seq[writer][reader]: rendezvous sequence numbers
    cudaIpcMemHandle_t stage[SHM_MAX_RANKS];       // device staging buffers, exported once
    int stageDevice[SHM_MAX_RANKS];                // -1: that rank has no device (host collectives only)
};

static ShmHeader* s_hdr = nullptr;
static char* s_slots = nullptr;
static size_t s_mapBytes = 0;
static int s_me = 0, s_P = 1;
static unsigned long long s_epoch[SHM_MAX_RANKS] = {0};
static cplx* s_stage = nullptr;                    // mine
static cplx* s_peerStage[SHM_MAX_RANKS] = {nullptr};
static double s_timeout = 300.0;
static_assert(sizeof(au64) == 8, "lock-free 64-bit atomics expected");

static double now_s() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static char* slot(int r) { return s_slots + (size_t)r * SHM_SLOT_BYTES; }

// spin politely until pred() holds; fails on timeout or when any rank raised the abort word
template <class Pred>
static int wait_until(Pred pred, const char* what) {
    double t0 = 0;
    for (unsigned long long it = 0; ; it++) {
        if (pred()) return 0;
        if (s_hdr->abortWord.load(std::memory_order_acquire)) return qb_set_error(-2, "shared-memory transport: another rank aborted", __FILE__, __LINE__);
        if (it < 2000) continue;
        sched_yield();
        if ((it & 1023) == 0) {
            if (t0 == 0) t0 = now_s();
            else if (now_s() - t0 > s_timeout) {
                s_hdr->abortWord.store(1, std::memory_order_release);
                return qb_set_error(-3, what, __FILE__, __LINE__);
            }
        }
    }
}

int shm_make_id(char* id, int idBytes) {
    memset(id, 0, idBytes);
    memcpy(id, SHM_MAGIC, 8);
    unsigned long long r[2] = {(unsigned long long)getpid(), 0};
    timespec t; clock_gettime(CLOCK_REALTIME, &t);
    r[1] = (unsigned long long)t.tv_nsec ^ ((unsigned long long)t.tv_sec << 20);
    int fd = open("/dev/urandom", O_RDONLY);
    if (fd >= 0) { unsigned long long x[2]; if (read(fd, x, sizeof x) == (ssize_t)sizeof x) { r[0] ^= x[0]; r[1] ^= x[1]; } close(fd); }
    snprintf(id + 8, idBytes - 8, "/qb200_%016llx%016llx", r[0], r[1]);
    return 0;
}

bool shm_is_id(const char* id) { return memcmp(id, SHM_MAGIC, 8) == 0; }

int shm_init(int rank, int numRanks, const char* id) {
    QB_REQUIRE(numRanks <= SHM_MAX_RANKS, "shared-memory transport: too many ranks");
    if (const char* e = getenv("QUEST_B200_SHM_TIMEOUT_S")) { double v = atof(e); if (v > 0) s_timeout = v; }
    s_me = rank; s_P = numRanks;
    const char* name = id + 8;
    size_t hdrBytes = (sizeof(ShmHeader) + 4095) & ~(size_t)4095;
    s_mapBytes = hdrBytes + (size_t)numRanks * SHM_SLOT_BYTES;
    int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    QB_REQUIRE(fd >= 0, "shared-memory transport: shm_open failed");
    if (ftruncate(fd, (off_t)s_mapBytes) != 0) { close(fd); return qb_set_error(-1, "shared-memory transport: ftruncate failed", __FILE__, __LINE__); }
    void* p = mmap(nullptr, s_mapBytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    QB_REQUIRE(p != MAP_FAILED, "shared-memory transport: mmap failed");
    s_hdr = (ShmHeader*)p;                         // a fresh segment is zero-filled: every counter starts at 0
    s_slots = (char*)p + hdrBytes;
    memset(s_epoch, 0, sizeof s_epoch);

    // export this rank's device staging buffer (ranks without a device still get the host collectives)
    s_hdr->stageDevice[rank] = -1;
    if (g_qb.device >= 0) {
        QB_CUDA(cudaMalloc(&s_stage, (size_t)SHM_STAGE_AMPS * sizeof(cplx)));
        QB_CUDA(cudaIpcGetMemHandle(&s_hdr->stage[rank], s_stage));
        s_hdr->stageDevice[rank] = g_qb.device;
    }
    s_hdr->arrived.fetch_add(1, std::memory_order_acq_rel);
    int r = wait_until([&] { return s_hdr->arrived.load(std::memory_order_acquire) >= (unsigned long long)s_P; },
                       "shared-memory transport: timed out waiting for every rank to join");
    if (r) { shm_unlink(name); return r; }
    for (int q = 0; q < s_P; q++) {
        s_peerStage[q] = nullptr;
        if (q == rank) { s_peerStage[q] = s_stage; continue; }
        if (g_qb.device < 0 || s_hdr->stageDevice[q] < 0) continue;
        void* ptr = nullptr;
        QB_CUDA(cudaIpcOpenMemHandle(&ptr, s_hdr->stage[q], cudaIpcMemLazyEnablePeerAccess));
        s_peerStage[q] = (cplx*)ptr;
    }
    r = shm_barrier(); if (r) return r;
    if (rank == 0) shm_unlink(name);               // everyone has it mapped: the name may go, the memory lives on
    return 0;
}

int shm_end() {
    if (!s_hdr) return 0;
    shm_barrier();
    for (int q = 0; q < s_P; q++) if (q != s_me && s_peerStage[q]) cudaIpcCloseMemHandle(s_peerStage[q]);
    shm_barrier();                                 // nobody frees a buffer a peer still has mapped
    if (s_stage) cudaFree(s_stage);
    s_stage = nullptr;
    munmap((void*)s_hdr, s_mapBytes);
    s_hdr = nullptr; s_slots = nullptr;
    return 0;
}

int shm_barrier() {
    unsigned long long gen = s_hdr->barrierGen.load(std::memory_order_acquire);
    if (s_hdr->barrierCount.fetch_add(1, std::memory_order_acq_rel) + 1 == (unsigned long long)s_P) {
        s_hdr->barrierCount.store(0, std::memory_order_relaxed);
        s_hdr->barrierGen.store(gen + 1, std::memory_order_release);
        return 0;
    }
    return wait_until([&] { return s_hdr->barrierGen.load(std::memory_order_acquire) != gen; },
                      "shared-memory transport: barrier timed out (a rank died or diverged)");
}

static void post(int r) { s_hdr->seq[s_me][r].store(++s_epoch[r], std::memory_order_release); }
static int await(int r) {
    const unsigned long long e = s_epoch[r];
    return wait_until([&] { return s_hdr->seq[r][s_me].load(std::memory_order_acquire) >= e; },
                      "shared-memory transport: pair rendezvous timed out (the partner died or diverged)");
}

int shm_pair_sync(int pairRank) { post(pairRank); return await(pairRank); }

int shm_sync_with(const int* ranks, int numRanks) {
    for (int i = 0; i < numRanks; i++) post(ranks[i]);            // post to everyone first: no ordering deadlock
    for (int i = 0; i < numRanks; i++) { int r = await(ranks[i]); if (r) return r; }
    return 0;
}

int shm_allreduce_sum(double* v, qindex n) {
    const qindex per = (qindex)(SHM_SLOT_BYTES / sizeof(double));
    for (qindex off = 0; off < n; off += per) {
        const qindex m = std::min(per, n - off);
        memcpy(slot(s_me), v + off, (size_t)m * sizeof(double));
        int r = shm_barrier(); if (r) return r;
        for (qindex i = 0; i < m; i++) {                          // rank order: every rank gets the identical sum
            double s = 0;
            for (int q = 0; q < s_P; q++) s += ((const double*)slot(q))[i];
            v[off + i] = s;
        }
        r = shm_barrier(); if (r) return r;
    }
    return 0;
}

int shm_broadcast(void* buf, size_t bytes, int root) {
    for (size_t off = 0; off < bytes; off += SHM_SLOT_BYTES) {
        const size_t m = std::min(SHM_SLOT_BYTES, bytes - off);
        if (s_me == root) memcpy(slot(root), (char*)buf + off, m);
        int r = shm_barrier(); if (r) return r;
        if (s_me != root) memcpy((char*)buf + off, slot(root), m);
        r = shm_barrier(); if (r) return r;
    }
    return 0;
}

int shm_allgather_host(const void* send, void* recvAll, size_t bytesPerRank) {
    for (size_t off = 0; off < bytesPerRank; off += SHM_SLOT_BYTES) {
        const size_t m = std::min(SHM_SLOT_BYTES, bytesPerRank - off);
        memcpy(slot(s_me), (const char*)send + off, m);
        int r = shm_barrier(); if (r) return r;
        if (recvAll) for (int q = 0; q < s_P; q++) memcpy((char*)recvAll + (size_t)q * bytesPerRank + off, slot(q), m);
        r = shm_barrier(); if (r) return r;
    }
    return 0;
}

int shm_sendrecv_host(const void* send, void* recv, size_t bytes, int pairRank) {
    for (size_t off = 0; off < bytes; off += SHM_SLOT_BYTES) {
        const size_t m = std::min(SHM_SLOT_BYTES, bytes - off);
        memcpy(slot(s_me), (const char*)send + off, m);
        int r = shm_pair_sync(pairRank); if (r) return r;
        memcpy((char*)recv + off, slot(pairRank), m);
        r = shm_pair_sync(pairRank); if (r) return r;
    }
    return 0;
}

int shm_send_host(const void* send, size_t bytes, int toRank) {
    for (size_t off = 0; off < bytes; off += SHM_SLOT_BYTES) {
        const size_t m = std::min(SHM_SLOT_BYTES, bytes - off);
        memcpy(slot(s_me), (const char*)send + off, m);
        int r = shm_pair_sync(toRank); if (r) return r;           // data is in my slot
        r = shm_pair_sync(toRank); if (r) return r;               // the receiver has taken it
    }
    return 0;
}

int shm_recv_host(void* recv, size_t bytes, int fromRank) {
    for (size_t off = 0; off < bytes; off += SHM_SLOT_BYTES) {
        const size_t m = std::min(SHM_SLOT_BYTES, bytes - off);
        int r = shm_pair_sync(fromRank); if (r) return r;
        memcpy((char*)recv + off, slot(fromRank), m);
        r = shm_pair_sync(fromRank); if (r) return r;
    }
    return 0;
}

// ---- device data plane: through the exported staging buffers -------------------------------------------------
static int need_stage(int pairRank) {
    QB_REQUIRE(s_stage && s_peerStage[pairRank], "shared-memory transport: no device staging buffer (rank without a GPU?)");
    return 0;
}

static int stage_out(const cplx* devSend, qindex m) {
    QB_CUDA(cudaMemcpyAsync(s_stage, devSend, (size_t)m * sizeof(cplx), cudaMemcpyDeviceToDevice, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    return 0;
}
static int stage_in(cplx* devRecv, int fromRank, qindex m) {
    QB_CUDA(cudaMemcpyAsync(devRecv, s_peerStage[fromRank], (size_t)m * sizeof(cplx), cudaMemcpyDeviceToDevice, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    return 0;
}

int shm_exchange(const cplx* devSend, cplx* devRecv, qindex numAmps, int pairRank) {
    int r = need_stage(pairRank); if (r) return r;
    for (qindex off = 0; off < numAmps; off += SHM_STAGE_AMPS) {
        const qindex m = std::min(SHM_STAGE_AMPS, numAmps - off);
        r = stage_out(devSend + off, m); if (r) return r;
        r = shm_pair_sync(pairRank); if (r) return r;             // both staging buffers are filled
        r = stage_in(devRecv + off, pairRank, m); if (r) return r;
        r = shm_pair_sync(pairRank); if (r) return r;             // both have been read: they may be refilled
    }
    return 0;
}

int shm_send(const cplx* devSend, qindex numAmps, int pairRank) {
    int r = need_stage(pairRank); if (r) return r;
    for (qindex off = 0; off < numAmps; off += SHM_STAGE_AMPS) {
        const qindex m = std::min(SHM_STAGE_AMPS, numAmps - off);
        r = stage_out(devSend + off, m); if (r) return r;
        r = shm_pair_sync(pairRank); if (r) return r;
        r = shm_pair_sync(pairRank); if (r) return r;
    }
    return 0;
}

int shm_recv(cplx* devRecv, qindex numAmps, int pairRank) {
    int r = need_stage(pairRank); if (r) return r;
    for (qindex off = 0; off < numAmps; off += SHM_STAGE_AMPS) {
        const qindex m = std::min(SHM_STAGE_AMPS, numAmps - off);
        r = shm_pair_sync(pairRank); if (r) return r;
        r = stage_in(devRecv + off, pairRank, m); if (r) return r;
        r = shm_pair_sync(pairRank); if (r) return r;
    }
    return 0;
}

int shm_allgather(const cplx* devSend, cplx* devRecv, qindex numAmpsPerRank) {
    for (int q = 0; q < s_P; q++) { int r = need_stage(q); if (r) return r; }
    for (qindex off = 0; off < numAmpsPerRank; off += SHM_STAGE_AMPS) {
        const qindex m = std::min(SHM_STAGE_AMPS, numAmpsPerRank - off);
        int r = stage_out(devSend + off, m); if (r) return r;
        r = shm_barrier(); if (r) return r;
        for (int q = 0; q < s_P; q++)
            QB_CUDA(cudaMemcpyAsync(devRecv + (qindex)q * numAmpsPerRank + off, s_peerStage[q], (size_t)m * sizeof(cplx),
                                    cudaMemcpyDeviceToDevice, g_qb.stream));
        QB_CUDA(cudaStreamSynchronize(g_qb.stream));
        r = shm_barrier(); if (r) return r;
    }
    return 0;
}
