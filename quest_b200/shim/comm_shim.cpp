/* comm_shim.cpp -- defines the reference's comm_* symbols (quest/src/comm/comm_config.hpp:15-27,
 * quest/src/comm/comm_routines.hpp:29-77) over NCCL via the quest_b200 C ABI (qb_comm_*).
 *
 * It replaces quest/src/comm/comm_config.cpp and quest/src/comm/comm_routines.cpp for ONE 8xB200 box:
 *   - one process per GPU; rank / world size come from the launcher's environment (torchrun's RANK,
 *     WORLD_SIZE, LOCAL_RANK; or QUEST_B200_RANK / QUEST_B200_WORLD_SIZE), not from mpirun;
 *   - the 128-byte ncclUniqueId is shared either through QUEST_B200_NCCL_ID (hex, set by a host program
 *     that already has a control plane, e.g. bench.py via torch.distributed) or through a rendezvous file
 *     written by rank 0 (QUEST_B200_ID_FILE, default /tmp/quest_b200_nccl_<parent pid>_<MASTER_PORT>);
 *   - amplitudes always travel GPU-to-GPU (NVLink); there is no host-staged path and therefore no
 *     distributed CPU-only Quregs in this build (they fail loudly);
 *   - host-side scalars (seeds, flags, reductions of a few doubles) ride NCCL through a device scratch word.
 * Ordering: every exchange is enqueued on the library's compute stream, so it is ordered after the kernels
 * that produced its input and before the kernels that consume its output; the reference's
 * cudaDeviceSynchronize() before every exchange (comm_routines.cpp:390,418,446) is not needed.
 */
#include "quest/include/types.h"
#include "quest/include/qureg.h"
#include "quest/include/matrices.h"

#include "quest/src/core/errors.hpp"
#include "quest/src/comm/comm_config.hpp"
#include "quest/src/comm/comm_routines.hpp"
#include "quest/src/comm/comm_indices.hpp"

#include "quest_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

using std::string;
using std::vector;

#define QB_CHECK(call) qbCommCheck((call), #call, __func__, __FILE__, __LINE__)

static void qbCommCheck(int status, const char* call, const char* caller, const char* file, int line) {
    if (status != 0)
        error_cudaCallFailed(qb_error_string(), call, caller, file, line);
}

static qb_cplx* qp(qcomp* p) { return reinterpret_cast<qb_cplx*>(p); }

static bool s_commInit = false;
static bool s_commEnded = false;
static int s_rank = 0;
static int s_numRanks = 1;


/*
 * BOOTSTRAP (replaces MPI_Init & friends, comm_config.cpp:59-187)
 */

static int envInt(const char* a, const char* b, int fallback) {
    const char* v = std::getenv(a);
    if (!v) v = std::getenv(b);
    return v ? std::atoi(v) : fallback;
}

static string idFilePath() {
    if (const char* f = std::getenv("QUEST_B200_ID_FILE"))
        return f;
    const char* port = std::getenv("MASTER_PORT");
    return "/tmp/quest_b200_nccl_" + std::to_string((long) getppid()) + "_" + (port ? port : "0");
}

static void obtainUniqueId(char id[QB_COMM_ID_BYTES]) {

    // (1) an id handed over by a host program that already has a control plane
    if (const char* hex = std::getenv("QUEST_B200_NCCL_ID")) {
        if (std::strlen(hex) != 2 * QB_COMM_ID_BYTES)
            error_commButEnvNotDistributed();
        for (int i = 0; i < QB_COMM_ID_BYTES; i++) {
            unsigned v = 0;
            std::sscanf(hex + 2 * i, "%2x", &v);
            id[i] = (char) v;
        }
        return;
    }

    // (2) rendezvous file: rank 0 publishes atomically (write + rename), the others poll
    string path = idFilePath();
    if (s_rank == 0) {
        QB_CHECK( qb_comm_get_unique_id(id) );
        string tmp = path + ".tmp";
        std::ofstream(tmp, std::ios::binary).write(id, QB_COMM_ID_BYTES);
        std::rename(tmp.c_str(), path.c_str());
        return;
    }
    for (int attempt = 0; attempt < 6000; attempt++) {          // up to 10 minutes
        std::ifstream in(path, std::ios::binary);
        if (in && in.read(id, QB_COMM_ID_BYTES) && in.gcount() == QB_COMM_ID_BYTES)
            return;
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    error_commButEnvNotDistributed();
}

bool comm_isMpiCompiled() {
    // "is a multi-process backend compiled": yes, NCCL. Gates env auto-deployment (core/autodeployer.cpp:29-31)
    return true;
}

bool comm_isMpiGpuAware() {
    // device pointers are always communicated directly
    return true;
}

bool comm_isInit() {
    return s_commInit;
}

void comm_init() {
    if (s_commInit || s_commEnded)
        error_commAlreadyInit();

    s_rank = envInt("QUEST_B200_RANK", "RANK", 0);
    s_numRanks = envInt("QUEST_B200_WORLD_SIZE", "WORLD_SIZE", 1);

    // a single process is a legal (trivial) distributed environment, exactly like `mpirun -n 1`
    if (s_numRanks > 1) {

        // NCCL needs this process bound to its GPU before the communicator exists
        int numGpus = qb_num_devices();
        int local = envInt("QUEST_B200_LOCAL_RANK", "LOCAL_RANK", s_rank);
        if (numGpus > 0)
            QB_CHECK( qb_bind_device(local % numGpus) );

        // fewer GPUs than ranks: ranks share devices, which NCCL refuses -- use the shared-memory / CUDA-IPC transport
        // (the reference permits the same sharing for its tests: PERMIT_NODES_TO_SHARE_GPU, api/environment.cpp:110)
        if (numGpus > 0 && s_numRanks > numGpus)
            QB_CHECK( qb_comm_set_transport(1) );

        char id[QB_COMM_ID_BYTES];
        obtainUniqueId(id);
        QB_CHECK( qb_comm_init(s_rank, s_numRanks, id) );
        QB_CHECK( qb_comm_barrier() );

        // map every peer's flag page now, while all ranks are in lock-step (a collective): later uses of the
        // NVLink peer-memory path are pair-wise and may be skipped by ranks whose control bits exclude them
        (void) qb_p2p_is_available();

        if (s_rank == 0 && !std::getenv("QUEST_B200_NCCL_ID"))
            std::remove(idFilePath().c_str());
    }

    s_commInit = true;
}

void comm_end() {
    if (!s_commInit)
        return;
    if (s_numRanks > 1) {
        QB_CHECK( qb_comm_barrier() );
        QB_CHECK( qb_comm_end() );
    }
    s_commInit = false;
    s_commEnded = true;
}

int comm_getRank() {
    return s_commInit ? s_rank : ROOT_RANK;
}

int comm_getNumNodes() {
    return s_commInit ? s_numRanks : 1;
}

bool comm_isRootNode(int rank) {
    return rank == ROOT_RANK;
}

bool comm_isRootNode() {
    return comm_isRootNode(comm_getRank());
}

void comm_sync() {
    if (!s_commInit || s_numRanks == 1)
        return;
    QB_CHECK( qb_comm_barrier() );
}


/*
 * STATE EXCHANGE (comm_routines.cpp:493-623)
 */

static void assertGpuDistributed(Qureg qureg) {
    if (!qureg.isDistributed)
        error_commButQuregNotDistributed();
    if (!qureg.isGpuAccelerated || qureg.gpuAmps == nullptr)
        error_cudaCallFailed("quest_b200 distributes GPU-accelerated Quregs only (no host-staged communication path)",
            "comm", __func__, __FILE__, __LINE__);
}

void comm_exchangeAmpsToBuffers(Qureg qureg, qindex sendInd, qindex recvInd, qindex numAmps, int pairRank) {
    assertGpuDistributed(qureg);
    if (pairRank == qureg.rank)
        error_commWithSameRank();
    QB_CHECK( qb_comm_exchange(qp(&qureg.gpuAmps[sendInd]), qp(&qureg.gpuCommBuffer[recvInd]), numAmps, pairRank) );
}

void comm_exchangeAmpsToBuffers(Qureg qureg, int pairRank) {
    comm_exchangeAmpsToBuffers(qureg, 0, 0, qureg.numAmpsPerNode, pairRank);
}

void comm_exchangeSubBuffers(Qureg qureg, qindex numAmps, int pairRank) {
    assertGpuDistributed(qureg);
    auto [sendInd, recvInd] = getSubBufferSendRecvInds(qureg);
    QB_CHECK( qb_comm_exchange(qp(&qureg.gpuCommBuffer[sendInd]), qp(&qureg.gpuCommBuffer[recvInd]), numAmps, pairRank) );
}

void comm_asynchSendSubBuffer(Qureg qureg, qindex numElems, int pairRank) {
    assertGpuDistributed(qureg);
    auto [sendInd, recvInd] = getSubBufferSendRecvInds(qureg);
    (void) recvInd;
    QB_CHECK( qb_comm_send(qp(&qureg.gpuCommBuffer[sendInd]), numElems, pairRank) );
}

void comm_receiveArrayToBuffer(Qureg qureg, qindex numElems, int pairRank) {
    assertGpuDistributed(qureg);
    auto [sendInd, recvInd] = getSubBufferSendRecvInds(qureg);
    (void) sendInd;
    QB_CHECK( qb_comm_recv(qp(&qureg.gpuCommBuffer[recvInd]), numElems, pairRank) );
}

void comm_combineAmpsIntoBuffer(Qureg receiver, Qureg sender) {
    assertGpuDistributed(receiver);
    assertGpuDistributed(sender);
    QB_CHECK( qb_comm_allgather(qp(sender.gpuAmps), qp(receiver.gpuCommBuffer), sender.numAmpsPerNode) );
}

void comm_combineElemsIntoBuffer(Qureg receiver, FullStateDiagMatr sender) {
    assertGpuDistributed(receiver);
    if (!sender.isDistributed || !sender.isGpuAccelerated || sender.gpuElems == nullptr)
        error_cudaCallFailed("quest_b200 needs a GPU-accelerated distributed FullStateDiagMatr here",
            "comm", __func__, __FILE__, __LINE__);
    QB_CHECK( qb_comm_allgather(qp(sender.gpuElems), qp(receiver.gpuCommBuffer), sender.numElemsPerNode) );
}


/*
 * MISC COMMUNICATION (comm_routines.cpp:632-705): small host-side messages
 */

static void assertDistributedEnv() {
    if (!s_commInit)
        error_commButEnvNotDistributed();
}

void comm_broadcastAmp(int sendRank, qcomp* sendAmp) {
    assertDistributedEnv();
    if (s_numRanks == 1) return;
    QB_CHECK( qb_comm_broadcast_bytes(sendAmp, sizeof(qcomp), sendRank) );
}

void comm_sendAmpsToRoot(int sendRank, qcomp* send, qcomp* recv, qindex numAmps) {
    assertDistributedEnv();
    if (s_numRanks == 1) return;
    QB_CHECK( qb_comm_sendrecv_host(qp(send), qp(recv), numAmps, sendRank, ROOT_RANK) );
}

void comm_broadcastIntsFromRoot(int* arr, qindex length) {
    assertDistributedEnv();
    if (s_numRanks == 1) return;
    QB_CHECK( qb_comm_broadcast_bytes(arr, sizeof(int) * length, ROOT_RANK) );
}

void comm_broadcastUnsignedsFromRoot(unsigned* arr, qindex length) {
    assertDistributedEnv();
    if (s_numRanks == 1) return;
    QB_CHECK( qb_comm_broadcast_bytes(arr, sizeof(unsigned) * length, ROOT_RANK) );
}

void comm_combineSubArrays(qcomp* recv, vector<qindex> globalRecvInds, vector<qindex> localSendInds, vector<qindex> numAmpsPerRank) {
    assertDistributedEnv();
    (void) localSendInds; // each rank has already placed its own contribution inside recv
    if (s_numRanks == 1) return;
    if (globalRecvInds.size() != (size_t) s_numRanks)
        error_commGivenInconsistentNumSubArraysANodes();

    // every contributing rank broadcasts its slice of the (duplicated) host array, in rank order
    for (int sendRank = 0; sendRank < s_numRanks; sendRank++)
        if (numAmpsPerRank[sendRank] > 0)
            QB_CHECK( qb_comm_broadcast_bytes(&recv[globalRecvInds[sendRank]], sizeof(qcomp) * numAmpsPerRank[sendRank], sendRank) );
}


/*
 * REDUCTIONS (comm_routines.cpp:714-777)
 */

void comm_reduceAmp(qcomp* localAmp) {
    assertDistributedEnv();
    if (s_numRanks == 1) return;
    double v[2] = {(double) std::real(*localAmp), (double) std::imag(*localAmp)};
    QB_CHECK( qb_comm_allreduce_sum(v, 2) );
    *localAmp = qcomp((qreal) v[0], (qreal) v[1]);
}

void comm_reduceReal(qreal* localReal) {
    assertDistributedEnv();
    if (s_numRanks == 1) return;
    double v = (double) *localReal;
    QB_CHECK( qb_comm_allreduce_sum(&v, 1) );
    *localReal = (qreal) v;
}

void comm_reduceReals(qreal* localReals, qindex numLocalReals) {
    assertDistributedEnv();
    if (s_numRanks == 1) return;
    vector<double> v(localReals, localReals + numLocalReals);
    QB_CHECK( qb_comm_allreduce_sum(v.data(), numLocalReals) );
    for (qindex i = 0; i < numLocalReals; i++)
        localReals[i] = (qreal) v[i];
}

bool comm_isTrueOnAllNodes(bool val) {
    assertDistributedEnv();
    if (s_numRanks == 1) return val;
    int flag = (int) val;
    QB_CHECK( qb_comm_allreduce_and(&flag) );
    return (bool) flag;
}

bool comm_isTrueOnRootNode(bool val) {
    assertDistributedEnv();
    unsigned out = (unsigned) val;
    comm_broadcastUnsignedsFromRoot(&out, 1);
    return (bool) out;
}


/*
 * GATHER (comm_routines.cpp:786-812)
 */

vector<string> comm_gatherStringsToRoot(char* localChars, int maxNumLocalChars) {
    assertDistributedEnv();
    vector<char> allChars((size_t) maxNumLocalChars * s_numRanks);
    if (s_numRanks == 1)
        std::memcpy(allChars.data(), localChars, maxNumLocalChars);
    else
        QB_CHECK( qb_comm_gather_bytes(localChars, allChars.data(), maxNumLocalChars, ROOT_RANK) );

    vector<string> out(s_numRanks);
    for (int r = 0; r < s_numRanks; r++)
        out[r] = (s_rank == ROOT_RANK || s_numRanks == 1)? string(&allChars[(size_t) r * maxNumLocalChars]) : string();
    return out;
}
