// qb_tile.cuh -- interface of the TMA tile engine (qb_tile.cu) to the per-gate entry points.
// Each qb_tile_try_* returns 1 when the tile engine executed (or failed executing) the operation, in
// which case qb_tile_status() holds the C-ABI status to return, or 0 when the operation is outside the
// engine's envelope (tiny states, engine disabled) and the direct kernel must run instead.
#pragma once
#include "qb_common.cuh"

int qb_tile_status();
int qb_tile_try_dense(const qb_state* q, const int* ctrls, const int* cs, int nc, const int* targs, int nt, const qb_cplx* hostMatr);
int qb_tile_try_diag(const qb_state* q, const int* ctrls, const int* cs, int nc, const int* targs, int nt, const qb_cplx* hostElems);
int qb_tile_try_pauli(const qb_state* q, const int* ctrls, const int* cs, int nc, unsigned long long maskXY, unsigned long long maskYZ, cplx ampFac, cplx pairFac);
int qb_tile_try_phase(const qb_state* q, const int* ctrls, const int* cs, int nc, unsigned long long targMask, cplx f0, cplx f1);
int qb_tile_try_swap(const qb_state* q, const int* ctrls, const int* cs, int nc, int t1, int t2);

// direct Pauli kernel from raw masks (pairFac already multiplied by i^numY); qb_gates.cu
int qb_pauli_raw(const qb_state* q, const int* ctrls, const int* cs, int nc, unsigned long long maskXY, unsigned long long maskYZ, qb_cplx ampFac, qb_cplx pairFac);

// applies the deferred gates of q to the amplitudes whose index bits `mask` hold `vals` only (see qb_tile.cu)
int qb_tile_flush_restricted(const qb_state* q, unsigned long long mask, unsigned long long vals, bool keepQueue);
