// qb_tile.cu -- TMA tile engine (placeholder: every operation is declined, the direct kernels run).
#include "qb_tile.cuh"
static int s_status = 0;
int qb_tile_status() { return s_status; }
int qb_tile_try_dense(const qb_state*, const int*, const int*, int, const int*, int, const qb_cplx*) { return 0; }
int qb_tile_try_denseK(const qb_state*, const int*, const int*, int, const int*, int, const qb_cplx*, int) { return 0; }
int qb_tile_try_diag(const qb_state*, const int*, const int*, int, const int*, int, const qb_cplx*) { return 0; }
int qb_tile_try_pauli(const qb_state*, const int*, const int*, int, unsigned long long, unsigned long long, cplx, cplx) { return 0; }
int qb_tile_try_phase(const qb_state*, const int*, const int*, int, unsigned long long, cplx, cplx) { return 0; }
int qb_tile_try_swap(const qb_state*, const int*, const int*, int, int, int) { return 0; }
