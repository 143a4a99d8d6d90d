// qb_comm_shm.cuh -- the "ranks share a GPU" transport of the communication layer (qb_comm_shm.cu).
//
// NCCL refuses two ranks on one device, yet the sharding logic (prefix / suffix case analysis, lazy relabelling,
// exchanges, reductions) must be testable on a 1-GPU box with 2, 4 or 8 ranks -- the reference's CI does the same by
// letting several MPI ranks bind to one GPU (CMakeLists.txt:235-240 PERMIT_NODES_TO_SHARE_GPU, api/environment.cpp:110).
// This transport keeps the process model (one process per rank, SPMD) and replaces only the wires:
//   control plane: a POSIX shared-memory segment (sequence-number rendezvous, per-rank mail slots)
//   data plane:    CUDA IPC -- every rank exports one device staging buffer, peers map it and copy device-to-device
// It is selected by the communicator id (ids made by shm_make_id carry a magic prefix), so every rank picks the same
// transport without any further agreement.
#pragma once
#include "qb_common.cuh"

int  shm_make_id(char* id, int idBytes);
bool shm_is_id(const char* id);
int  shm_init(int rank, int numRanks, const char* id);
int  shm_end();
int  shm_barrier();
int  shm_pair_sync(int pairRank);                                     // rendezvous of exactly two ranks
int  shm_sync_with(const int* ranks, int numRanks);                   // rendezvous with several ranks at once
int  shm_allreduce_sum(double* hostValues, qindex n);
int  shm_broadcast(void* hostBuf, size_t numBytes, int root);
int  shm_allgather_host(const void* send, void* recvAll, size_t bytesPerRank);
int  shm_sendrecv_host(const void* send, void* recv, size_t bytes, int pairRank);      // symmetric, both ranks call
int  shm_send_host(const void* send, size_t bytes, int toRank);
int  shm_recv_host(void* recv, size_t bytes, int fromRank);
int  shm_exchange(const cplx* devSend, cplx* devRecv, qindex numAmps, int pairRank);
int  shm_send(const cplx* devSend, qindex numAmps, int pairRank);
int  shm_recv(cplx* devRecv, qindex numAmps, int pairRank);
int  shm_allgather(const cplx* devSend, cplx* devRecv, qindex numAmpsPerRank);
