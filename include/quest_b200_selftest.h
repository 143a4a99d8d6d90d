/* quest_b200_selftest.h -- host-only self-tests of the backend's index algebra and of the tile engine's planner.
 *
 * NOT part of the product: these entry points exist only in quest_b200/lib/libquest_b200_selftest.so, which the Makefile
 * builds from the same sources with -DQB_SELFTEST for the CPU test-suite (tests/test_abi_cpu.py).  The product library
 * libquest_b200.so is compiled without them.
 */
#ifndef QUEST_B200_SELFTEST_H
#define QUEST_B200_SELFTEST_H
#include "quest_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int         qb_selftest_bitins(const int* qubits, const int* states, int n, qb_index item, qb_index* out); /* host-only index-algebra check */
int         qb_selftest_planner(int numQubits, int numOps, unsigned seed, int reorder, double* maxErr, int* numPasses, int* numRounds, int* numOpsPlanned); /* host-only: random gate list applied in program order vs in the tile planner's order (absorbed, merged, re-ordered) on a small host state */
int         qb_selftest_tile_emulation(int numQubits, int numOps, unsigned seed, int reorder, double* maxErr, int* numTilePasses, int* numDirectOps); /* host-only: planner + emit_pass descriptors + the kernel's own round driver and gate bodies (compiled for the host) vs gate-by-gate application */
int         qb_selftest_restricted_flush(int numQubits, int numOps, unsigned seed, int bit, double* maxErr, int* numTilePasses, int* numOpsPlanned); /* host-only: the half-shard (restricted) flush behind the exchange/compute overlap == plain application */
int         qb_selftest_pauli_group(int numQubits, int numOps, unsigned seed, int restrictBit, double* maxErr, int* numPasses); /* host-only: the coset kernel's body + descriptors vs the pair-by-pair definition (restrictBit < 0: whole state) */
#ifdef __cplusplus
}
#endif
#endif
