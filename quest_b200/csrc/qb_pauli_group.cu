// qb_pauli_group.cu -- several Pauli gadgets / Pauli-string expectation terms in ONE pass over the state.
//
// A Pauli tensor or gadget with X/Y mask m mixes the amplitude pairs (n, n ^ m) (cpu_subroutines.cpp:804-905); the
// reference -- CPU and GPU alike -- spends one full pass over the state per gadget (gpu_kernels.cuh:598-634) and one
// reduction pass per term of a Pauli-string sum (localiser.cpp:2097-2112, gpu_thrust.cuh:918-950).  Strings on many
// high qubits (BASELINE cfg 5: ~9 X/Y sites spread over 28 qubits) cannot enter a shared-memory tile either.
//
// But K gadgets with linearly independent masks m_1..m_K only ever mix amplitudes WITHIN the cosets  n ^ span{m_i}
// of 2^K elements.  So one thread takes a whole coset into registers (32 amplitudes for K = 5), applies the K gadgets
// one after the other there -- gadget t pairs register g with register g ^ (1 << t) -- and writes the coset back:
// K gadgets for one read and one write of the state.  Coset representatives are the indices whose K pivot bits (of
// the masks' echelon form over GF(2)) are zero, enumerated by bit insertion like every other kernel here; consecutive
// threads take consecutive representatives, and since XOR with a constant maps an aligned 512-byte block onto an
// aligned 512-byte block, every one of the 16 loads / stores of a warp is a whole block.  Diagonal gadgets (Z strings,
// phase gadgets) ride along for free.  The same cosets serve the expectation values: the K terms' sums
// sum_n (-1)^{popc(j & maskYZ)} conj(a_n) a_j, j = n ^ m_t, are all evaluated from the 16 registers.
// Algorithmic bytes: K x 2*16*N for gadgets, K x 16*N for expectation terms; physical: 2*16*N, 16*N.
#include "qb_common.cuh"
#include "qb_reduce.cuh"
#include "qb_pauli_group.cuh"
#include <string.h>

struct PGDevOp { int slot; unsigned sgn; qindex yz; cplx c, f; };      // slot >= 0: pairs (g, g ^ (1 << slot)); -1: diagonal
struct PGDev {
    int k, numOps;
    BitIns ins;
    qindex xr[PG_AMPS];
    PGDevOp ops[PG_MAX_OPS];
};

// gadget on register bit T: a_g' = c a_g + f s(idx_g1) a_g1,  a_g1' = c a_g1 + f s(idx_g) a_g   (OpPauliA, qb_gates.cu)
template <int A, int T>
__device__ __forceinline__ void pg_pair(cplx (&v)[A], cplx c, cplx f, unsigned par) {
#pragma unroll
    for (int g = 0; g < A; g++) {
        if (g & (1 << T)) continue;
        const int g1 = g | (1 << T);
        const double s0 = 1.0 - 2.0 * (double)((par >> g) & 1u), s1 = 1.0 - 2.0 * (double)((par >> g1) & 1u);
        const cplx a = v[g], b = v[g1];
        v[g] = cfma(f, cscale(s1, b), cmul(c, a));
        v[g1] = cfma(f, cscale(s0, a), cmul(c, b));
    }
}

template <int K>
__global__ void __launch_bounds__(128) k_pauli_group(cplx* __restrict__ amps, qindex numGroups, const PGDev* __restrict__ gp) {
    constexpr int A = 1 << K;
    __shared__ PGDev p;
    for (int i = threadIdx.x; i < (int)(sizeof(PGDev) / 4); i += blockDim.x) ((int*)&p)[i] = ((const int*)gp)[i];
    __syncthreads();
    const qindex n = (qindex)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= numGroups) return;
    const qindex base = p.ins(n);
    cplx v[A];
#pragma unroll
    for (int g = 0; g < A; g++) v[g] = ld_stream(amps + (base ^ p.xr[g]));
    const int numOps = p.numOps;
    for (int o = 0; o < numOps; o++) {
        const PGDevOp op = p.ops[o];
        // parity of (index & yz) per register: that of the representative, flipped where the register's offset says so
        const unsigned par = ((__popcll((unsigned long long)(base & op.yz)) & 1) ? ~op.sgn : op.sgn);
        switch (op.slot) {
        case 0: pg_pair<A, 0>(v, op.c, op.f, par); break;
        case 1: if (K > 1) pg_pair<A, (K > 1 ? 1 : 0)>(v, op.c, op.f, par); break;
        case 2: if (K > 2) pg_pair<A, (K > 2 ? 2 : 0)>(v, op.c, op.f, par); break;
        case 3: if (K > 3) pg_pair<A, (K > 3 ? 3 : 0)>(v, op.c, op.f, par); break;
        case 4: if (K > 4) pg_pair<A, (K > 4 ? 4 : 0)>(v, op.c, op.f, par); break;
        default:                                           // diagonal: v_g *= (parity ? f : c)
#pragma unroll
            for (int g = 0; g < A; g++) v[g] = cmul(v[g], ((par >> g) & 1u) ? op.f : op.c);
            break;
        }
    }
#pragma unroll
    for (int g = 0; g < A; g++) st_stream(amps + (base ^ p.xr[g]), v[g]);
}

// expectation terms: term t pairs register g with g ^ (1 << t); raw sums as qb_statevec_calcExpecPauliStrBatch_subA defines them
template <int K>
__global__ void __launch_bounds__(QB_BLOCK) k_pauli_group_expec(const cplx* __restrict__ amps, qindex numGroups, const PGDev* __restrict__ gp,
                                                               double* partials, unsigned int* ticket, double* out) {
    constexpr int A = 1 << K;
    __shared__ PGDev p;
    __shared__ double sm[2 * (QB_BLOCK / 32)];
    __shared__ bool isLast;
    for (int i = threadIdx.x; i < (int)(sizeof(PGDev) / 4); i += blockDim.x) ((int*)&p)[i] = ((const int*)gp)[i];
    __syncthreads();
    double re[K], im[K];
#pragma unroll
    for (int t = 0; t < K; t++) { re[t] = 0; im[t] = 0; }
    const qindex stride = (qindex)gridDim.x * QB_BLOCK;
    for (qindex n = (qindex)blockIdx.x * QB_BLOCK + threadIdx.x; n < numGroups; n += stride) {
        const qindex base = p.ins(n);
        cplx v[A];
#pragma unroll
        for (int g = 0; g < A; g++) v[g] = ld_stream(amps + (base ^ p.xr[g]));
#pragma unroll
        for (int t = 0; t < K; t++) {
            const unsigned par = ((__popcll((unsigned long long)(base & p.ops[t].yz)) & 1) ? ~p.ops[t].sgn : p.ops[t].sgn);
#pragma unroll
            for (int g = 0; g < A; g++) {
                if (g & (1 << t)) continue;
                const int g1 = g | (1 << t);
                const double s0 = 1.0 - 2.0 * (double)((par >> g) & 1u), s1 = 1.0 - 2.0 * (double)((par >> g1) & 1u);
                const cplx a = v[g], b = v[g1];
                // n = idx_g: conj(a) s(idx_g1) b ;  n = idx_g1: conj(b) s(idx_g) a
                const double zr = a.x * b.x + a.y * b.y, zi = a.x * b.y - a.y * b.x;      // conj(a) b
                re[t] += (s1 + s0) * zr;
                im[t] += (s1 - s0) * zi;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < K; t++) {
        double r = re[t], i = im[t];
        block_sum2(r, i, sm);
        if (threadIdx.x == 0) {
            partials[(2 * t) * QB_RED_MAX_BLOCKS + blockIdx.x] = r;
            partials[(2 * t + 1) * QB_RED_MAX_BLOCKS + blockIdx.x] = i;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int tk = atomicInc(ticket, gridDim.x - 1);
        isLast = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    for (int t = 0; t < K; t++) {
        double r = 0, i = 0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += QB_BLOCK) {
            r += __ldcg(&partials[(2 * t) * QB_RED_MAX_BLOCKS + b]);
            i += __ldcg(&partials[(2 * t + 1) * QB_RED_MAX_BLOCKS + b]);
        }
        block_sum2(r, i, sm);
        if (threadIdx.x == 0) { out[2 * t] = r; out[2 * t + 1] = i; }
    }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
// echelon form over GF(2): returns the rank of masks[0..k) and their pivot bits (each the highest set bit of the vector
// left after clearing the earlier pivots).  Every coset of the span then holds exactly one index whose pivot bits are 0.
int pg_rank(const unsigned long long* masks, int k, int* pivots) {
    unsigned long long basis[64]; int piv[64]; int r = 0;
    for (int i = 0; i < k; i++) {
        unsigned long long m = masks[i];
        for (int j = 0; j < r; j++) if ((m >> piv[j]) & 1ULL) m ^= basis[j];
        if (!m) continue;
        piv[r] = 63 - __builtin_clzll(m); basis[r] = m;
        if (pivots) pivots[r] = piv[r];
        r++;
    }
    return r;
}

static char* s_pgDev = nullptr; static size_t s_pgSlots = 0, s_pgNext = 0;
// descriptors live in a small ring in device memory: a slot is re-used only after PG_RING launches, all on one stream
#define PG_RING 256
static int pg_upload(const PGDev& h, const PGDev** dev) {
    if (!s_pgDev) { QB_CUDA(cudaMalloc(&s_pgDev, sizeof(PGDev) * PG_RING)); s_pgSlots = PG_RING; }
    char* slot = s_pgDev + sizeof(PGDev) * (s_pgNext++ % s_pgSlots);
    QB_CUDA(cudaMemcpyAsync(slot, &h, sizeof(PGDev), cudaMemcpyHostToDevice, g_qb.stream));     // pageable source: staged before return
    *dev = (const PGDev*)slot;
    return 0;
}

static int pg_build(const qb_state* q, const unsigned long long* xy, int k, PGDev& h, unsigned long long restrictMask = 0, unsigned long long restrictVals = 0) {
    memset(&h, 0, sizeof h);
    int piv[PG_K], zero[PG_K] = {0};
    int rb[64], rv[64], nr = 0;          // index bits held fixed (restricted pass): extra inserted bits of the representative
    for (int b = 0; b < 64; b++) if ((restrictMask >> b) & 1) { rb[nr] = b; rv[nr++] = (int)((restrictVals >> b) & 1); }
    QB_REQUIRE(k >= 1 && k <= PG_K && pg_rank(xy, k, piv) == k, "pauli group: masks must be linearly independent");
    for (int i = 0; i < k; i++) QB_REQUIRE(xy[i] < (unsigned long long)q->numAmpsPerNode, "pauli group: X/Y mask reaches prefix qubits");
    h.k = k;
    for (int i = 0; i < k; i++) QB_REQUIRE(!(xy[i] & restrictMask), "pauli group: a mask touches a restricted bit");
    h.ins = qb_make_ins(piv, zero, k, rb, rv, nr);
    for (int g = 0; g < (1 << k); g++) {
        unsigned long long x = 0;
        for (int i = 0; i < k; i++) if ((g >> i) & 1) x ^= xy[i];
        h.xr[g] = (qindex)x;
    }
    return 0;
}

static unsigned pg_sign_bits(const PGDev& h, unsigned long long yz) {
    unsigned s = 0;
    for (int g = 0; g < (1 << h.k); g++) if (__builtin_parityll((unsigned long long)h.xr[g] & yz)) s |= 1u << g;
    return s;
}

// applies `numOps` control-free Pauli gadgets / tensors (xy != 0) and parity phase gadgets (xy == 0, yz = target mask) in
// order, in one pass.  The non-zero xy masks must be linearly independent (at most PG_K of them).
int qb_pauli_group_apply(const qb_state* q, const PGOp* ops, int numOps, unsigned long long restrictMask, unsigned long long restrictVals) {
    QB_REQUIRE(numOps >= 1 && numOps <= PG_MAX_OPS, "pauli group: bad op count");
    unsigned long long xy[PG_K]; int k = 0;
    for (int i = 0; i < numOps; i++) if (ops[i].xy) { QB_REQUIRE(k < PG_K, "pauli group: too many X/Y masks"); xy[k++] = ops[i].xy; }
    QB_REQUIRE(k >= 1, "pauli group: needs at least one non-diagonal op");
    PGDev h; int r = pg_build(q, xy, k, h, restrictMask, restrictVals); if (r) return r;
    h.numOps = numOps;
    int slot = 0;
    for (int i = 0; i < numOps; i++) {
        PGDevOp& d = h.ops[i];
        d.slot = ops[i].xy ? slot++ : -1;
        d.yz = (qindex)ops[i].yz; d.sgn = pg_sign_bits(h, ops[i].yz); d.c = ops[i].c; d.f = ops[i].f;
    }
    const PGDev* dev; r = pg_upload(h, &dev); if (r) return r;
    const qindex groups = q->numAmpsPerNode >> (k + __builtin_popcountll(restrictMask));
    const unsigned grid = (unsigned)((groups + 127) / 128);
    switch (k) {
    case 1: k_pauli_group<1><<<grid, 128, 0, g_qb.stream>>>((cplx*)q->amps, groups, dev); break;
    case 2: k_pauli_group<2><<<grid, 128, 0, g_qb.stream>>>((cplx*)q->amps, groups, dev); break;
    case 3: k_pauli_group<3><<<grid, 128, 0, g_qb.stream>>>((cplx*)q->amps, groups, dev); break;
    case 4: k_pauli_group<4><<<grid, 128, 0, g_qb.stream>>>((cplx*)q->amps, groups, dev); break;
    default: k_pauli_group<5><<<grid, 128, 0, g_qb.stream>>>((cplx*)q->amps, groups, dev); break;
    }
    QB_LAUNCH_CHECK();
    return 0;
}

// raw sums of k (<= PG_K) expectation terms with linearly independent, non-zero X/Y masks; results land in
// devOut[0 .. 2k) (device, re/im interleaved) -- the caller copies them back (it batches several groups per sync)
int qb_pauli_group_expec(const qb_state* q, const unsigned long long* masks, int k, double* devOut) {
    unsigned long long xy[PG_K];
    for (int i = 0; i < k && i < PG_K; i++) xy[i] = masks[2 * i];
    PGDev h; int r = pg_build(q, xy, k, h); if (r) return r;
    h.numOps = k;
    for (int i = 0; i < k; i++) { h.ops[i].slot = i; h.ops[i].yz = (qindex)masks[2 * i + 1]; h.ops[i].sgn = pg_sign_bits(h, masks[2 * i + 1]); }
    const PGDev* dev; r = pg_upload(h, &dev); if (r) return r;
    const qindex groups = q->numAmpsPerNode >> k;
    qindex blocks = (groups + QB_BLOCK - 1) / QB_BLOCK, maxBlocks = (qindex)g_qb.numSMs * 4;
    if (maxBlocks > QB_RED_MAX_BLOCKS) maxBlocks = QB_RED_MAX_BLOCKS;
    if (blocks > maxBlocks) blocks = maxBlocks;
    if (blocks < 1) blocks = 1;
    switch (k) {
    case 1: k_pauli_group_expec<1><<<(unsigned)blocks, QB_BLOCK, 0, g_qb.stream>>>((const cplx*)q->amps, groups, dev, g_qb.redPartials, g_qb.redTicket, devOut); break;
    case 2: k_pauli_group_expec<2><<<(unsigned)blocks, QB_BLOCK, 0, g_qb.stream>>>((const cplx*)q->amps, groups, dev, g_qb.redPartials, g_qb.redTicket, devOut); break;
    case 3: k_pauli_group_expec<3><<<(unsigned)blocks, QB_BLOCK, 0, g_qb.stream>>>((const cplx*)q->amps, groups, dev, g_qb.redPartials, g_qb.redTicket, devOut); break;
    case 4: k_pauli_group_expec<4><<<(unsigned)blocks, QB_BLOCK, 0, g_qb.stream>>>((const cplx*)q->amps, groups, dev, g_qb.redPartials, g_qb.redTicket, devOut); break;
    default: k_pauli_group_expec<5><<<(unsigned)blocks, QB_BLOCK, 0, g_qb.stream>>>((const cplx*)q->amps, groups, dev, g_qb.redPartials, g_qb.redTicket, devOut); break;
    }
    QB_LAUNCH_CHECK();
    return 0;
}
