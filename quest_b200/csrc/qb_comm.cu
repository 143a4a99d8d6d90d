// qb_comm.cu -- inter-GPU communication over NCCL (NVLink 5 / NVSwitch inside one 8xB200 box).
// Replaces quest/src/comm/comm_config.cpp (MPI_Init/rank/size/barrier, :59-187) and
// quest/src/comm/comm_routines.cpp (chunked Isend/Irecv exchange :209-232, Ibcast all-gather :292-349,
// Allreduce :714-760, Bcast :632-695, Gather :786-812).  One process per GPU; ranks pair up as
// rank ^ mask, and NVSwitch gives every pairing full bandwidth simultaneously, so no chunk-size
// juggling against MPI's 2^31-element message limit (comm_routines.cpp:95) is needed.
//
// Amplitude traffic goes GPU-to-GPU directly: device pointers are handed to ncclSend/ncclRecv on the
// library's compute stream, so an exchange is ordered after the kernels that produced its input and
// before the kernels that consume its output without any host synchronisation (the reference
// cudaDeviceSynchronize()s before every exchange, comm_routines.cpp:390).
//
// Second transport, for boxes with fewer GPUs than ranks (qb_comm_shm.cu): NCCL refuses two ranks on one device, so
// when the communicator id was made by the shared-memory transport every entry point below routes to it instead --
// same ABI, same semantics, host-synchronous.  It exists so that the sharding logic can be verified with 2/4/8 ranks
// on a single GPU (the reference's CI shares GPUs between MPI ranks for the same reason, CMakeLists.txt:235-240).
#include "qb_common.cuh"
#include "qb_comm_shm.cuh"
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

static ncclComm_t s_comm = nullptr;
static int s_rank = 0, s_numRanks = 1;
static bool s_init = false;
static bool s_shm = false;                // this communicator runs over the shared-memory / CUDA-IPC transport
static int s_wantTransport = -1;          // -1: from QUEST_B200_TRANSPORT (default nccl); 0: nccl; 1: shm
static char* s_devScratch = nullptr;      // small device staging area for host-side collectives
static char* s_hostScratch = nullptr;     // pinned
static size_t s_scratchBytes = 0;

static int nccl_error(ncclResult_t r, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s -> %s", what, ncclGetErrorString(r));
    return qb_set_error(-1000 - (int)r, buf, file, line);
}
#define QB_NCCL(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return nccl_error(r__, #call, __FILE__, __LINE__); } while (0)
#define QB_COMM_READY() do { QB_READY(); QB_REQUIRE(s_init, "communicator not initialised (qb_comm_init)"); } while (0)
// host-side collectives of the shared-memory transport also serve a rank that has no device (control-plane tests)
#define QB_COMM_READY_HOST() do { if (!(s_shm && g_qb.device < 0)) QB_READY(); QB_REQUIRE(s_init, "communicator not initialised (qb_comm_init)"); } while (0)

static bool want_shm() {
    if (s_wantTransport >= 0) return s_wantTransport == 1;
    const char* e = getenv("QUEST_B200_TRANSPORT");
    return e && strcmp(e, "shm") == 0;
}

static int ensureScratch(size_t bytes) {
    if (bytes <= s_scratchBytes) return 0;
    size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    if (s_devScratch) cudaFree(s_devScratch);
    if (s_hostScratch) cudaFreeHost(s_hostScratch);
    s_devScratch = nullptr; s_hostScratch = nullptr; s_scratchBytes = 0;
    QB_CUDA(cudaMalloc(&s_devScratch, want));
    QB_CUDA(cudaMallocHost(&s_hostScratch, want));
    s_scratchBytes = want;
    return 0;
}

extern "C" {

int qb_comm_get_unique_id(char id[QB_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) <= QB_COMM_ID_BYTES, "ncclUniqueId larger than QB_COMM_ID_BYTES");
    if (want_shm()) return shm_make_id(id, QB_COMM_ID_BYTES);
    ncclUniqueId uid;
    QB_NCCL(ncclGetUniqueId(&uid));
    memset(id, 0, QB_COMM_ID_BYTES);
    memcpy(id, &uid, sizeof uid);
    return 0;
}

int qb_comm_init(int rank, int numRanks, const char id[QB_COMM_ID_BYTES]) {
    QB_REQUIRE(!s_init, "communicator already initialised");
    QB_REQUIRE(numRanks >= 1 && rank >= 0 && rank < numRanks, "qb_comm_init: bad rank / size");
    QB_REQUIRE((numRanks & (numRanks - 1)) == 0, "qb_comm_init: number of ranks must be a power of two");
    // bind this process to "its" GPU before creating the communicator (gpu_config.cpp:332-353)
    int nd = qb_num_devices();
    if (shm_is_id(id)) {
        // ranks may share a device; a box without any device still gets the host-side collectives
        if (nd > 0 && g_qb.device < 0) { int r = qb_bind_device(rank % nd); if (r) return r; }
        int r = shm_init(rank, numRanks, id); if (r) return r;
        s_shm = true; s_rank = rank; s_numRanks = numRanks; s_init = true;
        return 0;
    }
    QB_REQUIRE(nd > 0, "qb_comm_init: no CUDA device");
    QB_REQUIRE(numRanks <= nd || numRanks == 1, "qb_comm_init: more ranks than GPUs needs the shared-memory transport (qb_comm_set_transport(1) / QUEST_B200_TRANSPORT=shm before the id is made)");
    if (g_qb.device < 0) { int r = qb_bind_device(rank % nd); if (r) return r; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    QB_NCCL(ncclCommInitRank(&s_comm, numRanks, uid, rank));
    s_rank = rank; s_numRanks = numRanks; s_init = true;
    return ensureScratch(1u << 20);
}

int qb_comm_end(void) {
    if (!s_init) return 0;
    if (g_qb.device >= 0) cudaStreamSynchronize(g_qb.stream);
    if (s_shm) shm_end(); else ncclCommDestroy(s_comm);
    s_shm = false;
    s_comm = nullptr; s_init = false; s_rank = 0; s_numRanks = 1;
    return 0;
}

int qb_comm_is_init(void) { return s_init ? 1 : 0; }
int qb_comm_rank(void) { return s_rank; }
int qb_comm_num_ranks(void) { return s_numRanks; }
int qb_comm_set_transport(int transport) { s_wantTransport = (transport == 1) ? 1 : 0; return 0; }
int qb_comm_transport(void) { return s_init && s_shm ? 1 : 0; }

int qb_comm_allreduce_sum(double* hostValues, qb_index n) {
    QB_COMM_READY_HOST();
    if (n <= 0) return 0;
    if (s_shm) return shm_allreduce_sum(hostValues, n);
    size_t bytes = sizeof(double) * (size_t)n;
    int r = ensureScratch(bytes); if (r) return r;
    memcpy(s_hostScratch, hostValues, bytes);
    QB_CUDA(cudaMemcpyAsync(s_devScratch, s_hostScratch, bytes, cudaMemcpyHostToDevice, g_qb.stream));
    QB_NCCL(ncclAllReduce(s_devScratch, s_devScratch, (size_t)n, ncclDouble, ncclSum, s_comm, g_qb.stream));
    QB_CUDA(cudaMemcpyAsync(s_hostScratch, s_devScratch, bytes, cudaMemcpyDeviceToHost, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    memcpy(hostValues, s_hostScratch, bytes);
    return 0;
}

int qb_comm_barrier(void) {
    QB_COMM_READY_HOST();
    if (g_qb.device >= 0) QB_CUDA(cudaDeviceSynchronize());
    if (s_shm) return shm_barrier();
    double x = 0;
    return qb_comm_allreduce_sum(&x, 1);
}

int qb_comm_allreduce_and(int* hostFlag) {
    QB_COMM_READY_HOST();
    double v = *hostFlag ? 0.0 : 1.0;          // count the ranks on which the flag is false
    int r = qb_comm_allreduce_sum(&v, 1); if (r) return r;
    *hostFlag = (v == 0.0);
    return 0;
}

int qb_comm_exchange(const qb_cplx* devSend, qb_cplx* devRecv, qb_index numAmps, int pairRank) {
    QB_COMM_READY();
    QB_REQUIRE(pairRank >= 0 && pairRank < s_numRanks && pairRank != s_rank, "exchange: bad pair rank");
    if (numAmps <= 0) return 0;
    if (s_shm) return shm_exchange((const cplx*)devSend, (cplx*)devRecv, numAmps, pairRank);
    QB_NCCL(ncclGroupStart());
    QB_NCCL(ncclSend(devSend, (size_t)numAmps * sizeof(cplx), ncclChar, pairRank, s_comm, g_qb.stream));
    QB_NCCL(ncclRecv(devRecv, (size_t)numAmps * sizeof(cplx), ncclChar, pairRank, s_comm, g_qb.stream));
    QB_NCCL(ncclGroupEnd());
    return 0;
}

int qb_comm_send(const qb_cplx* devSend, qb_index numAmps, int pairRank) {
    QB_COMM_READY();
    QB_REQUIRE(pairRank >= 0 && pairRank < s_numRanks && pairRank != s_rank, "send: bad pair rank");
    if (numAmps <= 0) return 0;
    if (s_shm) return shm_send((const cplx*)devSend, numAmps, pairRank);
    QB_NCCL(ncclSend(devSend, (size_t)numAmps * sizeof(cplx), ncclChar, pairRank, s_comm, g_qb.stream));
    return 0;
}

int qb_comm_recv(qb_cplx* devRecv, qb_index numAmps, int pairRank) {
    QB_COMM_READY();
    QB_REQUIRE(pairRank >= 0 && pairRank < s_numRanks && pairRank != s_rank, "recv: bad pair rank");
    if (numAmps <= 0) return 0;
    if (s_shm) return shm_recv((cplx*)devRecv, numAmps, pairRank);
    QB_NCCL(ncclRecv(devRecv, (size_t)numAmps * sizeof(cplx), ncclChar, pairRank, s_comm, g_qb.stream));
    return 0;
}

int qb_comm_allgather(const qb_cplx* devSend, qb_cplx* devRecv, qb_index numAmpsPerRank) {
    QB_COMM_READY();
    if (numAmpsPerRank <= 0) return 0;
    if (s_shm) return shm_allgather((const cplx*)devSend, (cplx*)devRecv, numAmpsPerRank);
    QB_NCCL(ncclAllGather(devSend, devRecv, (size_t)numAmpsPerRank * sizeof(cplx), ncclChar, s_comm, g_qb.stream));
    return 0;
}

int qb_comm_broadcast_bytes(void* hostBuf, size_t numBytes, int root) {
    QB_COMM_READY_HOST();
    if (numBytes == 0) return 0;
    if (s_shm) return shm_broadcast(hostBuf, numBytes, root);
    int r = ensureScratch(numBytes); if (r) return r;
    if (s_rank == root) memcpy(s_hostScratch, hostBuf, numBytes);
    QB_CUDA(cudaMemcpyAsync(s_devScratch, s_hostScratch, numBytes, cudaMemcpyHostToDevice, g_qb.stream));
    QB_NCCL(ncclBroadcast(s_devScratch, s_devScratch, numBytes, ncclChar, root, s_comm, g_qb.stream));
    QB_CUDA(cudaMemcpyAsync(s_hostScratch, s_devScratch, numBytes, cudaMemcpyDeviceToHost, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    memcpy(hostBuf, s_hostScratch, numBytes);
    return 0;
}

int qb_comm_gather_bytes(const void* hostSend, void* hostRecvOnRoot, size_t numBytesPerRank, int root) {
    QB_COMM_READY_HOST();
    if (numBytesPerRank == 0) return 0;
    if (s_shm) return shm_allgather_host(hostSend, (s_rank == root) ? hostRecvOnRoot : nullptr, numBytesPerRank);
    size_t total = numBytesPerRank * (size_t)(s_numRanks + 1);
    int r = ensureScratch(total); if (r) return r;
    // layout: [own contribution][gathered numRanks contributions]
    memcpy(s_hostScratch, hostSend, numBytesPerRank);
    QB_CUDA(cudaMemcpyAsync(s_devScratch, s_hostScratch, numBytesPerRank, cudaMemcpyHostToDevice, g_qb.stream));
    QB_NCCL(ncclAllGather(s_devScratch, s_devScratch + numBytesPerRank, numBytesPerRank, ncclChar, s_comm, g_qb.stream));
    QB_CUDA(cudaMemcpyAsync(s_hostScratch + numBytesPerRank, s_devScratch + numBytesPerRank,
                            numBytesPerRank * s_numRanks, cudaMemcpyDeviceToHost, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    if (s_rank == root && hostRecvOnRoot) memcpy(hostRecvOnRoot, s_hostScratch + numBytesPerRank, numBytesPerRank * s_numRanks);
    return 0;
}

int qb_comm_sendrecv_host(const qb_cplx* hostSend, qb_cplx* hostRecv, qb_index numAmps, int sendRank, int recvRank) {
    // comm_sendAmpsToRoot (comm_routines.cpp:643-671): sendRank's host array -> recvRank's host array
    QB_COMM_READY_HOST();
    if (numAmps <= 0 || sendRank == recvRank) return 0;
    if (s_rank != sendRank && s_rank != recvRank) return 0;
    size_t bytes = sizeof(qb_cplx) * (size_t)numAmps;
    if (s_shm) return s_rank == sendRank ? shm_send_host(hostSend, bytes, recvRank) : shm_recv_host(hostRecv, bytes, sendRank);
    int r = ensureScratch(bytes); if (r) return r;
    if (s_rank == sendRank) {
        memcpy(s_hostScratch, hostSend, bytes);
        QB_CUDA(cudaMemcpyAsync(s_devScratch, s_hostScratch, bytes, cudaMemcpyHostToDevice, g_qb.stream));
        QB_NCCL(ncclSend(s_devScratch, (size_t)numAmps * sizeof(cplx), ncclChar, recvRank, s_comm, g_qb.stream));
        QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    } else {
        QB_NCCL(ncclRecv(s_devScratch, (size_t)numAmps * sizeof(cplx), ncclChar, sendRank, s_comm, g_qb.stream));
        QB_CUDA(cudaMemcpyAsync(s_hostScratch, s_devScratch, bytes, cudaMemcpyDeviceToHost, g_qb.stream));
        QB_CUDA(cudaStreamSynchronize(g_qb.stream));
        memcpy(hostRecv, s_hostScratch, bytes);
    }
    return 0;
}

} // extern "C"

// ------------------------------------------------------------------------------------------
// internal host-side helpers for the peer-memory layer (qb_p2p.cu)
// ------------------------------------------------------------------------------------------
int qb_comm_internal_allgather_host(const void* send, void* recvAll, size_t bytesPerRank) {
    QB_COMM_READY();
    if (s_shm) return shm_allgather_host(send, recvAll, bytesPerRank);
    size_t total = bytesPerRank * (size_t)(s_numRanks + 1);
    int r = ensureScratch(total); if (r) return r;
    memcpy(s_hostScratch, send, bytesPerRank);
    QB_CUDA(cudaMemcpyAsync(s_devScratch, s_hostScratch, bytesPerRank, cudaMemcpyHostToDevice, g_qb.stream));
    QB_NCCL(ncclAllGather(s_devScratch, s_devScratch + bytesPerRank, bytesPerRank, ncclChar, s_comm, g_qb.stream));
    QB_CUDA(cudaMemcpyAsync(s_hostScratch + bytesPerRank, s_devScratch + bytesPerRank, bytesPerRank * s_numRanks,
                            cudaMemcpyDeviceToHost, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    memcpy(recvAll, s_hostScratch + bytesPerRank, bytesPerRank * s_numRanks);
    return 0;
}

int qb_comm_internal_sendrecv_host(const void* send, void* recv, size_t bytes, int pairRank) {
    QB_COMM_READY();
    if (s_shm) return shm_sendrecv_host(send, recv, bytes, pairRank);
    int r = ensureScratch(2 * bytes); if (r) return r;
    memcpy(s_hostScratch, send, bytes);
    QB_CUDA(cudaMemcpyAsync(s_devScratch, s_hostScratch, bytes, cudaMemcpyHostToDevice, g_qb.stream));
    QB_NCCL(ncclGroupStart());
    QB_NCCL(ncclSend(s_devScratch, bytes, ncclChar, pairRank, s_comm, g_qb.stream));
    QB_NCCL(ncclRecv(s_devScratch + bytes, bytes, ncclChar, pairRank, s_comm, g_qb.stream));
    QB_NCCL(ncclGroupEnd());
    QB_CUDA(cudaMemcpyAsync(s_hostScratch + bytes, s_devScratch + bytes, bytes, cudaMemcpyDeviceToHost, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    memcpy(recv, s_hostScratch + bytes, bytes);
    return 0;
}

// blocking rendezvous with a set of ranks (each of which names this rank in its own set): one grouped
// NCCL send/recv of a byte per partner, so the order in which ranks list their partners cannot deadlock
int qb_comm_internal_sync_with(const int* ranks, int numRanks) {
    QB_COMM_READY();
    if (numRanks <= 0) return 0;
    if (s_shm) { QB_CUDA(cudaStreamSynchronize(g_qb.stream)); return shm_sync_with(ranks, numRanks); }
    int r = ensureScratch(2 * (size_t)s_numRanks + 16); if (r) return r;
    QB_NCCL(ncclGroupStart());
    for (int i = 0; i < numRanks; i++) {
        QB_NCCL(ncclSend(s_devScratch + ranks[i], 1, ncclChar, ranks[i], s_comm, g_qb.stream));
        QB_NCCL(ncclRecv(s_devScratch + s_numRanks + ranks[i], 1, ncclChar, ranks[i], s_comm, g_qb.stream));
    }
    QB_NCCL(ncclGroupEnd());
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    return 0;
}

// host-side rendezvous of a pair when the ranks share a device (qb_p2p.cu: a spinning flag kernel of one process could
// starve the partner's kernel of the same GPU; the host can simply wait for its own stream and meet the partner)
bool qb_comm_internal_is_shm() { return s_init && s_shm; }
int qb_comm_internal_pair_sync_host(int pairRank, cudaStream_t stream) {
    QB_CUDA(cudaStreamSynchronize(stream));
    return shm_pair_sync(pairRank);
}
