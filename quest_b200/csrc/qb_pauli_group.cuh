// qb_pauli_group.cuh -- interface of the coset-blocked Pauli kernels (qb_pauli_group.cu)
#pragma once
#include "qb_common.cuh"

#define PG_K 5                       // independent X/Y masks per pass (2^PG_K amplitudes per thread: 32 = 128 registers)
#define PG_AMPS (1 << PG_K)
#define PG_K_GADGET 4                // gadgets per pass chosen by the planner: ncu, 28 qubits: K=4 1.303 ms per pass (6.55 TB/s, 0.326 ms per gadget),
                                     // K=5 1.895 ms (4.5 TB/s at 192 registers per thread, 0.379 ms per gadget)
#define PG_K_EXPEC 4                 // expectation terms per pass: measured faster than 5 (32.7 vs 45.0 ms for 200 terms at 28 qubits: occupancy)
#define PG_MAX_OPS 12                // gadgets per pass: PG_K non-diagonal ones plus diagonal ones riding along

// one control-free op: xy != 0: a <- c a + f (-1)^{popc(j & yz)} a_j, j = n ^ xy (f carries i^numY);
//                      xy == 0: a <- (parity(n & yz) ? f : c) a
struct PGOp { unsigned long long xy, yz; cplx c, f; };

int pg_rank(const unsigned long long* masks, int k, int* pivots);
int qb_pauli_group_apply(const qb_state* q, const PGOp* ops, int numOps, unsigned long long restrictMask = 0, unsigned long long restrictVals = 0);
int qb_pauli_group_expec(const qb_state* q, const unsigned long long* masks, int k, double* devOut);
