// qb_p2p.cu -- fused compute + exchange over NVLink peer memory.
//
// One process per GPU.  Every amplitude allocation made through qb_alloc can be exported with CUDA IPC; the
// first time a pair of ranks needs each other's amplitudes they trade IPC handles (over the NCCL control
// plane) and map the partner's allocation with cudaIpcOpenMemHandle(..., LazyEnablePeerAccess).  After that,
// a gate on a prefix (rank-bit) qubit is ONE kernel per GPU: each GPU owns half of the amplitude pairs, loads
// its own element from local HBM and the partner's element straight over NVLink (coalesced 128-bit peer loads),
// applies the 2x2 matrix and stores both results -- one locally, one through NVLink.  Compared with the
// reference's "exchange the whole shard into the buffer, then combine" (core/localiser.cpp:941-953):
//   * no communication buffer traffic (saves writing + re-reading B*N bytes of HBM per GPU),
//   * no separate combine pass: the NVLink transfer overlaps the arithmetic element by element,
//   * the link carries the same B*N bytes per direction, so the pass is purely NVLink-bound.
// Cross-GPU ordering: system-scope epoch flags in an IPC-shared page per rank; a one-thread kernel publishes
// "my stream reached this point" to the partner and waits for the partner's flag, before and after the fused
// kernel (so neither GPU reads amplitudes the other is still producing, nor runs ahead of remote writes).
#include "qb_common.cuh"
#include "qb_kernels.cuh"
#include "qb_tile.cuh"
#include <map>
#include <vector>
#include <string.h>
#include <stdlib.h>

int qb_comm_internal_allgather_host(const void* send, void* recvAll, size_t bytesPerRank);
int qb_comm_internal_sendrecv_host(const void* send, void* recv, size_t bytes, int pairRank);
int qb_comm_internal_sync_with(const int* ranks, int numRanks);
bool qb_comm_internal_is_shm();
int qb_comm_internal_pair_sync_host(int pairRank, cudaStream_t stream);

#define QB_P2P_MAX_RANKS 64

struct Alloc { size_t bytes; bool exported; unsigned long long seq; };      // seq: n-th allocation of this process (SPMD: the same on every rank)
static std::map<uintptr_t, Alloc> s_allocs;                          // local allocations made by qb_alloc
static std::map<std::pair<uintptr_t,int>, void*> s_peerBase;         // (local base, pair rank) -> mapped partner base

static bool s_ready = false, s_enabled = true, s_triedInit = false;
static unsigned long long* s_myFlags = nullptr;                       // [QB_P2P_MAX_RANKS] written by peers
static unsigned long long* s_peerFlags[QB_P2P_MAX_RANKS] = {nullptr}; // mapped flag pages of every peer
static unsigned long long s_epoch[QB_P2P_MAX_RANKS] = {0};
static unsigned long long s_numExchanges = 0, s_linkBytesPerDir = 0;    // statistics (qb_p2p_stats)

static unsigned long long s_allocSeq = 0;
void qb_p2p_note_alloc(void* base, size_t bytes) { s_allocs[(uintptr_t)base] = Alloc{bytes, false, ++s_allocSeq}; }

// called by qb_free before cudaFree: unmap what we imported for this logical allocation, and -- if the memory
// was exported -- make sure every importer has unmapped it before it is released (collective, like destroyQureg)
int qb_p2p_note_free(void* base) {
    auto it = s_allocs.find((uintptr_t)base);
    if (it == s_allocs.end()) return 0;
    // handle exchange is symmetric, so the ranks we imported from are exactly the ranks that imported ours:
    // unmap theirs, then rendezvous with each of them so that nobody frees memory a partner still has mapped
    std::vector<int> partners;
    for (auto p = s_peerBase.begin(); p != s_peerBase.end(); ) {
        if (p->first.first == (uintptr_t)base) { cudaIpcCloseMemHandle(p->second); partners.push_back(p->first.second); p = s_peerBase.erase(p); }
        else ++p;
    }
    s_allocs.erase(it);
    if (!partners.empty() && qb_comm_is_init()) return qb_comm_internal_sync_with(partners.data(), (int)partners.size());
    return 0;
}

static int p2p_init() {
    if (s_triedInit) return 0;
    s_triedInit = true;
    if (!qb_comm_is_init() || qb_comm_num_ranks() < 2 || qb_comm_num_ranks() > QB_P2P_MAX_RANKS) return 0;
    const int P = qb_comm_num_ranks(), me = qb_comm_rank();
    QB_CUDA(cudaMalloc(&s_myFlags, sizeof(unsigned long long) * QB_P2P_MAX_RANKS));
    QB_CUDA(cudaMemset(s_myFlags, 0, sizeof(unsigned long long) * QB_P2P_MAX_RANKS));
    cudaIpcMemHandle_t mine;
    cudaError_t e = cudaIpcGetMemHandle(&mine, s_myFlags);
    int ok = (e == cudaSuccess);
    if (!ok) cudaGetLastError();
    std::vector<cudaIpcMemHandle_t> all(P);
    int r = qb_comm_internal_allgather_host(&mine, all.data(), sizeof mine); if (r) return r;
    for (int p = 0; p < P && ok; p++) {
        if (p == me) { s_peerFlags[p] = s_myFlags; continue; }
        void* ptr = nullptr;
        e = cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        s_peerFlags[p] = (unsigned long long*)ptr;
    }
    // every rank must agree, otherwise one side would wait for flags that never come
    double bad = ok ? 0.0 : 1.0;
    r = qb_comm_allreduce_sum(&bad, 1); if (r) return r;
    s_ready = (bad == 0.0);
    return 0;
}

static int peer_pointer(const void* localPtr, int pairRank, void** peerPtr) {
    uintptr_t p = (uintptr_t)localPtr;
    auto it = s_allocs.upper_bound(p);
    QB_REQUIRE(it != s_allocs.begin(), "p2p: pointer was not allocated by qb_alloc");
    --it;
    QB_REQUIRE(p < it->first + it->second.bytes, "p2p: pointer was not allocated by qb_alloc");
    uintptr_t base = it->first;
    unsigned long long offset = p - base;
    auto key = std::make_pair(base, pairRank);
    auto found = s_peerBase.find(key);
    if (found == s_peerBase.end()) {
        struct Msg { cudaIpcMemHandle_t h; unsigned long long offset, bytes, seq; } mine, theirs;
        memset(&mine, 0, sizeof mine);
        QB_CUDA(cudaIpcGetMemHandle(&mine.h, (void*)base));
        mine.offset = offset; mine.bytes = it->second.bytes; mine.seq = it->second.seq;
        int r = qb_comm_internal_sendrecv_host(&mine, &theirs, sizeof mine, pairRank); if (r) return r;
        // the partner must be talking about the SAME logical allocation (two Quregs of equal size would pass a size check)
        QB_REQUIRE(theirs.offset == offset && theirs.bytes == mine.bytes && theirs.seq == mine.seq, "p2p: partner's allocation does not mirror this rank's");
        void* mapped = nullptr;
        QB_CUDA(cudaIpcOpenMemHandle(&mapped, theirs.h, cudaIpcMemLazyEnablePeerAccess));
        it->second.exported = true;
        found = s_peerBase.emplace(key, mapped).first;
    }
    *peerPtr = (char*)found->second + offset;
    return 0;
}

// The spin is bounded: if the partner never arrives (its process died, or the ranks diverged) the kernel traps after
// `timeoutNs`, which poisons the context -- the next runtime call of this rank fails and is reported through
// error_cudaCallFailed -- instead of leaving an unkillable spinning kernel behind.
__global__ void k_pair_barrier(unsigned long long* peerSlot, volatile unsigned long long* mySlot, unsigned long long epoch, unsigned long long timeoutNs) {
    __threadfence_system();
    *(volatile unsigned long long*)peerSlot = epoch;        // "my stream has reached epoch"
    __threadfence_system();
    unsigned long long t0 = 0;
    for (unsigned long long it = 0; *mySlot < epoch; it++) {
        if ((it & 0xFFF) == 0xFFF) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeoutNs) { printf("quest_b200: pair barrier timed out waiting for the partner GPU (epoch %llu)\n", epoch); __trap(); }
        }
    }
    __threadfence_system();
}

static unsigned long long pair_timeout_ns() {
    static unsigned long long ns = 0;
    if (!ns) { const char* e = getenv("QUEST_B200_PAIR_TIMEOUT_S"); double s = e ? atof(e) : 120.0; if (s <= 0) s = 120.0; ns = (unsigned long long)(s * 1e9); }
    return ns;
}

// The same barrier without the diagnostic printf: <= 32 registers, so that its single warp still finds room on an SM
// whose register file a persistent k_tile_pass CTA has all but filled (384 x 168 of 65536) -- it runs on the exchange
// stream WHILE fused passes run on the compute stream (qb_p2p_swapHalvesOverlapped)
__global__ void __launch_bounds__(32) k_pair_barrier_lean(unsigned long long* peerSlot, volatile unsigned long long* mySlot, unsigned long long epoch, unsigned long long timeoutNs) {
    if (threadIdx.x) return;
    __threadfence_system();
    *(volatile unsigned long long*)peerSlot = epoch;
    __threadfence_system();
    unsigned long long t0 = 0;
    for (unsigned it = 0; *mySlot < epoch; it++) {
        if ((it & 0xFFF) == 0xFFF) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeoutNs) __trap();
        }
    }
    __threadfence_system();
}

static int pair_barrier_on(cudaStream_t stream, int pairRank, bool lean) {
    // ranks sharing one device rendezvous on the host (qb_comm_shm.cu): a spinning kernel would only hold the GPU
    // against the very kernel it is waiting for
    if (qb_comm_internal_is_shm()) return qb_comm_internal_pair_sync_host(pairRank, stream);
    unsigned long long epoch = ++s_epoch[pairRank];
    if (lean) k_pair_barrier_lean<<<1, 32, 0, stream>>>(s_peerFlags[pairRank] + qb_comm_rank(), s_myFlags + pairRank, epoch, pair_timeout_ns());
    else k_pair_barrier<<<1, 1, 0, stream>>>(s_peerFlags[pairRank] + qb_comm_rank(), s_myFlags + pairRank, epoch, pair_timeout_ns());
    QB_LAUNCH_CHECK();
    return 0;
}
static int pair_barrier(int pairRank) { return pair_barrier_on(g_qb.stream, pairRank, false); }

// MODE 0: 2x2 dense gate across the pair; MODE 1: swap.  Each item = one amplitude of mine + one of the partner's.
struct P2POp { BitIns ins; qindex peerXor; int bit; cplx m00, m01, m10, m11; };

template <int MODE, int ITEMS>
__global__ void __launch_bounds__(QB_BLOCK) k_p2p_pair(cplx* __restrict__ mine, cplx* __restrict__ peer, qindex first, qindex count, const P2POp op) {
    qindex idx[ITEMS];
    cplx a[ITEMS], b[ITEMS];
    const qindex base = (qindex)blockIdx.x * (QB_BLOCK * ITEMS) + threadIdx.x;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const qindex n = base + (qindex)j * QB_BLOCK;
        if (ITEMS == 1 && n >= count) return;
        idx[j] = op.ins(first + n);
        a[j] = mine[idx[j]];
        b[j] = peer[idx[j] ^ op.peerXor];
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        cplx outMine, outPeer;
        if (MODE == 0) {
            cplx a0 = op.bit ? b[j] : a[j], a1 = op.bit ? a[j] : b[j];
            cplx n0 = cfma(op.m01, a1, cmul(op.m00, a0));
            cplx n1 = cfma(op.m11, a1, cmul(op.m10, a0));
            outMine = op.bit ? n1 : n0; outPeer = op.bit ? n0 : n1;
        } else { outMine = b[j]; outPeer = a[j]; }
        mine[idx[j]] = outMine;
        peer[idx[j] ^ op.peerXor] = outPeer;
    }
    __threadfence_system();
}

template <int MODE>
static int launch_pair(cplx* mine, cplx* peer, qindex first, qindex count, const P2POp& op) {
    if (count <= 0) return 0;
    constexpr int ITEMS = 4;
    if (count >= (qindex)QB_BLOCK * ITEMS && count % (QB_BLOCK * ITEMS) == 0)
        k_p2p_pair<MODE, ITEMS><<<qb_grid(count, ITEMS), QB_BLOCK, 0, g_qb.stream>>>(mine, peer, first, count, op);
    else
        k_p2p_pair<MODE, 1><<<qb_grid(count, 1), QB_BLOCK, 0, g_qb.stream>>>(mine, peer, first, count, op);
    QB_LAUNCH_CHECK();
    return 0;
}

// splits `total` work items between the two ranks of a pair: the rank with the smaller id takes the first half
static void my_share(qindex total, int pairRank, qindex* first, qindex* count) {
    bool low = qb_comm_rank() < pairRank;
    qindex half = total / 2;
    if (total == 1) { *first = 0; *count = low ? 1 : 0; return; }
    *first = low ? 0 : half;
    *count = low ? half : total - half;
}

// ---- half-shard swap through the partner's communication buffer -------------------------------------------
// mode 1 (kernel push):  every GPU WRITES its outgoing half straight into the partner's buffer over NVLink (posted
//   writes, no read round trips), the pair synchronises, and a local HBM-speed pass moves the received half from the
//   own buffer into place.  mode 2 (copy-engine push): the same two steps issued as cudaMemcpy(2D)Async.
// mode 0: the in-place exchange kernel (each GPU reads and writes half of the pairs remotely; no buffer needed).
static int s_swapMode = -1;
static bool s_swapModeIsDefault = true;
static int swap_mode() {
    if (s_swapMode < 0) { const char* e = getenv("QUEST_B200_SWAP_MODE"); s_swapModeIsDefault = (e == nullptr); s_swapMode = e ? atoi(e) : 0; if (s_swapMode < 0 || s_swapMode > 2) s_swapMode = 0; }
    return s_swapMode;
}

template <int ITEMS, bool GATHER>      // GATHER: dst[n] = src[ins(n)] (pack + push);  else dst[ins(n)] = src[n] (unpack in place)
__global__ void __launch_bounds__(QB_BLOCK) k_half_copy(cplx* __restrict__ dst, const cplx* __restrict__ src, qindex count, const BitIns ins) {
    const qindex base = (qindex)blockIdx.x * (QB_BLOCK * ITEMS) + threadIdx.x;
    cplx v[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const qindex n = base + (qindex)j * QB_BLOCK;
        if (n < count) v[j] = GATHER ? src[ins(n)] : src[n];
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const qindex n = base + (qindex)j * QB_BLOCK;
        if (n < count) { if (GATHER) dst[n] = v[j]; else dst[ins(n)] = v[j]; }
    }
    if (GATHER) __threadfence_system();
}

// copies between the strided half "suffix bit s == bitVal" of `shard` and a compact array, with the copy engines
static int dma_half_copy(cplx* compact, cplx* shard, qindex numAmps, int s, int bitVal, bool toCompact, cudaStream_t stream = nullptr) {
    if (!stream) stream = g_qb.stream;
    const qindex run = (qindex)1 << s;                     // contiguous amplitudes per row
    const qindex rows = numAmps / (2 * run);
    cplx* strided = shard + (bitVal ? run : 0);
    const size_t width = (size_t)run * sizeof(cplx), pitch = 2 * width;
    if (rows <= 16 || pitch > ((size_t)1 << 30)) {
        for (qindex r = 0; r < rows; r++) {
            cplx* a = compact + r * run; cplx* b = strided + 2 * r * run;
            QB_CUDA(cudaMemcpyAsync(toCompact ? a : b, toCompact ? b : a, width, cudaMemcpyDeviceToDevice, stream));
        }
    } else {
        if (toCompact) QB_CUDA(cudaMemcpy2DAsync(compact, width, strided, pitch, width, (size_t)rows, cudaMemcpyDeviceToDevice, stream));
        else           QB_CUDA(cudaMemcpy2DAsync(strided, pitch, compact, width, width, (size_t)rows, cudaMemcpyDeviceToDevice, stream));
    }
    g_qb.launches += 1;
    return 0;
}

extern "C" {

int qb_p2p_is_available(void) {
    if (!s_enabled) return 0;
    if (!s_triedInit) { if (p2p_init()) return 0; }
    return s_ready ? 1 : 0;
}

int qb_p2p_set_enabled(int enabled) { s_enabled = enabled != 0; return 0; }

int qb_p2p_anyCtrlOneTargDenseMatr(const qb_state* q, const int* ctrls, const int* cs, int nc, int pairRank, int rankBit, const qb_cplx m[4]) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(ctrls, nc, q);
    QB_REQUIRE(qb_p2p_is_available(), "p2p path is not available");
    void* peer = nullptr;
    int r = peer_pointer(q->amps, pairRank, &peer); if (r) return r;
    P2POp op; op.ins = qb_make_ins(ctrls, cs, nc, nullptr, nullptr, 0); op.peerXor = 0; op.bit = rankBit;
    op.m00 = mk(m[0]); op.m01 = mk(m[1]); op.m10 = mk(m[2]); op.m11 = mk(m[3]);
    qindex first, count;
    my_share(q->numAmpsPerNode >> nc, pairRank, &first, &count);
    s_numExchanges++; s_linkBytesPerDir += (unsigned long long)(q->numAmpsPerNode >> nc) * sizeof(cplx);
    r = pair_barrier(pairRank); if (r) return r;
    r = launch_pair<0>((cplx*)q->amps, (cplx*)peer, first, count, op); if (r) return r;
    return pair_barrier(pairRank);
}

static int swap_halves(const qb_state* q, int suffixTarg, int pairRank);

int qb_p2p_swapHalves(const qb_state* q, int suffixTarg, int pairRank) {
    QB_READY(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(&suffixTarg, 1, q);
    return swap_halves(q, suffixTarg, pairRank);
}

// The same swap, but allowed to OVERTAKE the gates still deferred in the queue: it runs now, they run later.  That
// is only legal when every queued gate commutes with the swap -- none touches `suffixTarg` (checked here; otherwise
// the queue is flushed first) and none depends on the rank bit being swapped (the caller's promise: it resolves
// rank-bit controls and diagonal sites before they reach this library, so only it can know).
int qb_p2p_swapHalvesDeferred(const qb_state* q, int suffixTarg, int pairRank) {
    QB_READY_NOFLUSH(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(&suffixTarg, 1, q);
    unsigned long long touched = 0;
    int queued = qb_queue_info(q, &touched, nullptr);
    if (queued == 0 || ((touched >> suffixTarg) & 1)) QB_FLUSH();      // (a queue of another state is flushed too: cheap and safe)
    return swap_halves(q, suffixTarg, pairRank);
}

// The swap OVERLAPPED with the deferred gates.  The queued gates commute with the swap when none of them involves
// `suffixTarg` (nor the rank bit: the caller's promise), and then they act independently on the half of the shard that
// stays (suffix bit == this rank's bit) and on the half that leaves.  So the leaving half is sent UNPROCESSED, right
// away, by the copy engines on a second stream into the partner's communication buffer, while the fused passes of
// the queued gates run on the staying half; when the partner's half has landed it is moved into place and the same
// passes run on it.  Every amplitude meets every gate exactly once -- on whichever GPU holds it when the gate runs.
// Unlike swapHalvesDeferred this drains the queue (in two halves) on every rank, and both ranks of the pair must call
// it (the data path differs from the in-place exchange kernel's): the caller decides from rank-independent state.
// Reference behaviour replaced: core/localiser.cpp:997-1040 (swap in, apply, swap back -- serial, twice the traffic),
// comm/comm_routines.cpp:384-407 (device-wide sync before every exchange).
static unsigned long long s_numOverlapped = 0, s_numOverlappedWithGates = 0;
extern "C" unsigned long long qb_p2p_overlapped_count(int withQueuedGatesOnly) { return withQueuedGatesOnly ? s_numOverlappedWithGates : s_numOverlapped; }
static cudaStream_t s_xStream = nullptr;
static cudaEvent_t s_evReady = nullptr, s_evArrived = nullptr;

int qb_p2p_swapHalvesOverlapped(const qb_state* q, int suffixTarg, int pairRank) {
    QB_READY_NOFLUSH(); QB_CHECK_STATE(q); QB_CHECK_SUFFIX(&suffixTarg, 1, q);
    QB_REQUIRE(qb_p2p_is_available(), "p2p path is not available");
    QB_REQUIRE(q->buffer != nullptr && q->numAmpsPerNode >= 2, "overlapped swap: the state has no communication buffer");
    unsigned long long touched = 0;
    int queued = qb_queue_info(q, &touched, nullptr);
    if ((touched >> suffixTarg) & 1) { QB_FLUSH(); queued = 0; }          // cannot overlap: plain buffered exchange below
    if (!s_xStream) {
        QB_CUDA(cudaStreamCreateWithFlags(&s_xStream, cudaStreamNonBlocking));
        QB_CUDA(cudaEventCreateWithFlags(&s_evReady, cudaEventDisableTiming));
        QB_CUDA(cudaEventCreateWithFlags(&s_evArrived, cudaEventDisableTiming));
    }
    s_numExchanges++; s_linkBytesPerDir += (unsigned long long)q->numAmpsPerNode / 2 * sizeof(cplx);
    s_numOverlapped++; if (queued) s_numOverlappedWithGates++;
    const int myBit = qb_comm_rank() > pairRank ? 1 : 0, st = !myBit;     // the half with suffix bit == st leaves
    void* peerBuf = nullptr;
    int r = peer_pointer(q->buffer, pairRank, &peerBuf); if (r) return r;
    const unsigned long long bitMask = 1ULL << suffixTarg;

    // optional timeline of one call (QUEST_B200_OVERLAP_TRACE=1): synchronises, so for diagnosis only
    static int trace = -1;
    if (trace < 0) { const char* e = getenv("QUEST_B200_OVERLAP_TRACE"); trace = (e && e[0] == '1') ? 1 : 0; }
    cudaEvent_t tv[7];
    if (trace) for (int i = 0; i < 7; i++) cudaEventCreate(&tv[i]);
#define TRACE(i, stream) do { if (trace) cudaEventRecord(tv[i], stream); } while (0)

    // exchange stream: after everything issued so far (the leaving half must be final, the buffers free) ...
    TRACE(0, g_qb.stream);
    QB_CUDA(cudaEventRecord(s_evReady, g_qb.stream));
    QB_CUDA(cudaStreamWaitEvent(s_xStream, s_evReady, 0));
    r = pair_barrier_on(s_xStream, pairRank, true); if (r) return r;     // ... on BOTH GPUs
    TRACE(1, s_xStream);
    r = dma_half_copy((cplx*)peerBuf, (cplx*)q->amps, q->numAmpsPerNode, suffixTarg, st, true, s_xStream); if (r) return r;
    TRACE(2, s_xStream);
    r = pair_barrier_on(s_xStream, pairRank, true); if (r) return r;     // both halves have landed
    TRACE(3, s_xStream);
    QB_CUDA(cudaEventRecord(s_evArrived, s_xStream));

    // compute stream, meanwhile: the queued gates on the half that stays
    if (queued) { r = qb_tile_flush_restricted(q, bitMask, myBit ? bitMask : 0, true); if (r) return r; }
    TRACE(4, g_qb.stream);

    // the arrived half goes where the departed one was, then meets the same gates
    QB_CUDA(cudaStreamWaitEvent(g_qb.stream, s_evArrived, 0));
    const qindex half = q->numAmpsPerNode / 2;
    const BitIns ins = qb_make_ins(&suffixTarg, &st, 1, nullptr, nullptr, 0);
    k_half_copy<4, false><<<qb_grid(half, 4), QB_BLOCK, 0, g_qb.stream>>>((cplx*)q->amps, (const cplx*)q->buffer, half, ins);
    QB_LAUNCH_CHECK();
    TRACE(5, g_qb.stream);
    if (queued) { r = qb_tile_flush_restricted(q, bitMask, st ? bitMask : 0, false); if (r) return r; }
    TRACE(6, g_qb.stream);
#undef TRACE
    if (trace) {
        cudaDeviceSynchronize();
        float t[7]; for (int i = 1; i < 7; i++) cudaEventElapsedTime(&t[i], tv[0], tv[i]);
        fprintf(stderr, "[overlap rank %d] victim %d queued %d | xs: barrier1 %.2f copy-done %.2f barrier2 %.2f | compute: stay-half done %.2f unpack done %.2f arrived-half done %.2f ms\n",
                qb_comm_rank(), suffixTarg, queued, t[1], t[2], t[3], t[4], t[5], t[6]);
        for (int i = 0; i < 7; i++) cudaEventDestroy(tv[i]);
    }
    return 0;
}

static int swap_halves(const qb_state* q, int suffixTarg, int pairRank) {
    s_numExchanges++; s_linkBytesPerDir += (unsigned long long)q->numAmpsPerNode / 2 * sizeof(cplx);
    QB_REQUIRE(qb_p2p_is_available(), "p2p path is not available");
    void* peer = nullptr;
    // my amplitudes with suffix bit == !myBit trade places with the partner's amplitudes with suffix bit == myBit,
    // where myBit is this rank's value of the prefix qubit: myBit = 1 iff rank > pairRank (they differ in that bit only)
    int myBit = qb_comm_rank() > pairRank ? 1 : 0;
    int st = !myBit;
    // default (QUEST_B200_SWAP_MODE unset): the copy engines push through the buffers when the rows are long (>= 64 KiB):
    // 8 GiB in 11.05 ms = 777 GB/s per direction + 2.5 ms local unpack, against 14.7 ms for the in-place kernel
    // (profiles/r2_overlap_trace_2gpu.txt); short rows and buffer-less states use the in-place exchange kernel
    int mode = (q->buffer != nullptr && q->numAmpsPerNode >= 2) ? swap_mode() : 0;
    if (s_swapModeIsDefault && q->buffer != nullptr && suffixTarg >= 12) mode = 2;
    int r;
    if (mode == 0) {
        r = peer_pointer(q->amps, pairRank, &peer); if (r) return r;
        P2POp op; op.ins = qb_make_ins(&suffixTarg, &st, 1, nullptr, nullptr, 0); op.peerXor = pow2(suffixTarg); op.bit = myBit;
        op.m00 = op.m01 = op.m10 = op.m11 = mk(0, 0);
        qindex first, count;
        my_share(q->numAmpsPerNode / 2, pairRank, &first, &count);
        r = pair_barrier(pairRank); if (r) return r;
        r = launch_pair<1>((cplx*)q->amps, (cplx*)peer, first, count, op); if (r) return r;
        return pair_barrier(pairRank);
    }
    // through the buffers: element n of my outgoing half lands at element n of the partner's buffer, and the partner's
    // element n (its half with suffix bit == myBit) belongs at my index ins(n) -- the place my own element n came from
    r = peer_pointer(q->buffer, pairRank, &peer); if (r) return r;
    const qindex half = q->numAmpsPerNode / 2;
    const BitIns ins = qb_make_ins(&suffixTarg, &st, 1, nullptr, nullptr, 0);
    constexpr int ITEMS = 4;
    r = pair_barrier(pairRank); if (r) return r;                 // the partner's buffer is free, its stream is here too
    if (mode == 2) { r = dma_half_copy((cplx*)peer, (cplx*)q->amps, q->numAmpsPerNode, suffixTarg, st, true); if (r) return r; }
    else { k_half_copy<ITEMS, true><<<qb_grid(half, ITEMS), QB_BLOCK, 0, g_qb.stream>>>((cplx*)peer, (const cplx*)q->amps, half, ins); QB_LAUNCH_CHECK(); }
    r = pair_barrier(pairRank); if (r) return r;                 // both halves have landed
    // (the local unpack is an HBM-speed kernel in both modes: 2.5 ms for 8 GiB; the copy engines are slower at it)
    { k_half_copy<ITEMS, false><<<qb_grid(half, ITEMS), QB_BLOCK, 0, g_qb.stream>>>((cplx*)q->amps, (const cplx*)q->buffer, half, ins); QB_LAUNCH_CHECK(); }
    return 0;
}

extern "C" int qb_p2p_stats(unsigned long long* numExchanges, unsigned long long* linkBytesPerDir) {
    if (numExchanges) *numExchanges = s_numExchanges;
    if (linkBytesPerDir) *linkBytesPerDir = s_linkBytesPerDir;
    return 0;
}

extern "C" int qb_p2p_set_swap_mode(int mode) { s_swapMode = (mode < 0 || mode > 2) ? 0 : mode; s_swapModeIsDefault = false; return 0; }

} // extern "C"
