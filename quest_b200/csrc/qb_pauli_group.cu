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
// one coset: the body of a thread of k_pauli_group, compiled for the host as well (qb_selftest_pauli_group emulates the
// kernel representative by representative on the CPU)
#ifdef __CUDA_ARCH__
#define PG_LD(p) ld_stream(p)
#define PG_ST(p, v) st_stream(p, v)
#define PG_POPC(x) __popcll(x)
#else
#define PG_LD(p) (*(p))
#define PG_ST(p, v) (*(p) = (v))
#define PG_POPC(x) __builtin_popcountll(x)
#endif
template <int A, int T>
__host__ __device__ __forceinline__ void pg_pair_hd(cplx (&v)[A], cplx c, cplx f, unsigned par) {
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
__host__ __device__ __forceinline__ void pg_coset(cplx* __restrict__ amps, qindex n, const PGDev& p) {
    constexpr int A = 1 << K;
    const qindex base = p.ins(n);
    cplx v[A];
#pragma unroll
    for (int g = 0; g < A; g++) v[g] = PG_LD(amps + (base ^ p.xr[g]));
    const int numOps = p.numOps;
    for (int o = 0; o < numOps; o++) {
        const PGDevOp op = p.ops[o];
        // parity of (index & yz) per register: that of the representative, flipped where the register's offset says so
        const unsigned par = ((PG_POPC((unsigned long long)(base & op.yz)) & 1) ? ~op.sgn : op.sgn);
        switch (op.slot) {
        case 0: pg_pair_hd<A, 0>(v, op.c, op.f, par); break;
        case 1: if (K > 1) pg_pair_hd<A, (K > 1 ? 1 : 0)>(v, op.c, op.f, par); break;
        case 2: if (K > 2) pg_pair_hd<A, (K > 2 ? 2 : 0)>(v, op.c, op.f, par); break;
        case 3: if (K > 3) pg_pair_hd<A, (K > 3 ? 3 : 0)>(v, op.c, op.f, par); break;
        case 4: if (K > 4) pg_pair_hd<A, (K > 4 ? 4 : 0)>(v, op.c, op.f, par); break;
        default:                                           // diagonal: v_g *= (parity ? f : c)
#pragma unroll
            for (int g = 0; g < A; g++) v[g] = cmul(v[g], ((par >> g) & 1u) ? op.f : op.c);
            break;
        }
    }
#pragma unroll
    for (int g = 0; g < A; g++) PG_ST(amps + (base ^ p.xr[g]), v[g]);
}

template <int K>
__global__ void __launch_bounds__(128) k_pauli_group(cplx* __restrict__ amps, qindex numGroups, const PGDev* __restrict__ gp) {
    __shared__ PGDev p;
    for (int i = threadIdx.x; i < (int)(sizeof(PGDev) / 4); i += blockDim.x) ((int*)&p)[i] = ((const int*)gp)[i];
    __syncthreads();
    const qindex n = (qindex)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= numGroups) return;
    pg_coset<K>(amps, n, p);
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
// host: the pass descriptor of a gadget group (pivots, coset offsets, per-op sign bits); returns k through *kOut
static int pg_describe(const qb_state* q, const PGOp* ops, int numOps, unsigned long long restrictMask, unsigned long long restrictVals, PGDev& h, int* kOut) {
    QB_REQUIRE(numOps >= 1 && numOps <= PG_MAX_OPS, "pauli group: bad op count");
    unsigned long long xy[PG_K]; int k = 0;
    for (int i = 0; i < numOps; i++) if (ops[i].xy) { QB_REQUIRE(k < PG_K, "pauli group: too many X/Y masks"); xy[k++] = ops[i].xy; }
    QB_REQUIRE(k >= 1, "pauli group: needs at least one non-diagonal op");
    int r = pg_build(q, xy, k, h, restrictMask, restrictVals); if (r) return r;
    h.numOps = numOps;
    int slot = 0;
    for (int i = 0; i < numOps; i++) {
        PGDevOp& d = h.ops[i];
        d.slot = ops[i].xy ? slot++ : -1;
        d.yz = (qindex)ops[i].yz; d.sgn = pg_sign_bits(h, ops[i].yz); d.c = ops[i].c; d.f = ops[i].f;
    }
    *kOut = k;
    return 0;
}

int qb_pauli_group_apply(const qb_state* q, const PGOp* ops, int numOps, unsigned long long restrictMask, unsigned long long restrictVals) {
    PGDev h; int k = 0;
    int r = pg_describe(q, ops, numOps, restrictMask, restrictVals, h, &k); if (r) return r;
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

#ifdef QB_SELFTEST
// ------------------------------------------------------------------------------------------
// host emulation of the coset kernel (no CUDA): random control-free Pauli gadgets and parity gadgets are grouped as the
// planner groups them (linearly independent X/Y masks, at most PG_K_GADGET per pass), each group's descriptor is built by
// the product's own pg_describe and run coset by coset through the kernel's own body (pg_coset, compiled for the host), on
// the whole state or restricted to half of it; the result must equal pair-by-pair application of the definition
// (cpu_subroutines.cpp:804-905).  tests/test_abi_cpu.py
// ------------------------------------------------------------------------------------------
#include <complex>
#include <random>
#include <vector>
#include "../../include/quest_b200_selftest.h"
extern "C" int qb_selftest_pauli_group(int numQubits, int numOps, unsigned seed, int restrictBit, double* maxErr, int* numPasses) {
    if (numQubits < 3 || numQubits > 20 || numOps < 1 || restrictBit >= numQubits) return -1;
    typedef std::complex<double> hc;
    const int n = numQubits;
    std::mt19937_64 rng(seed);
    auto unif = [&]() { return (double)(rng() >> 11) * (1.0 / 9007199254740992.0); };
    const unsigned long long all = ((1ULL << n) - 1) & ~(restrictBit >= 0 ? (1ULL << restrictBit) : 0ULL);
    std::vector<PGOp> ops;
    for (int i = 0; i < numOps; i++) {
        PGOp o; o.xy = (rng() & rng()) & all; o.yz = (rng() & rng()) & all;          // ~1/4 of the sites each
        if (rng() % 5 == 0) o.xy = 0;                                                // a diagonal (parity) gadget
        if (rng() % 7 == 0 && !ops.empty()) o.xy = ops.back().xy;                    // a repeated string: dependent mask
        o.c = mk(2 * unif() - 1, 2 * unif() - 1); o.f = mk(2 * unif() - 1, 2 * unif() - 1);
        ops.push_back(o);
    }
    std::vector<hc> ref((size_t)1 << n);
    for (auto& v : ref) v = hc(2 * unif() - 1, 2 * unif() - 1);
    std::vector<cplx> state(ref.size());
    for (size_t i = 0; i < ref.size(); i++) state[i] = mk(ref[i].real(), ref[i].imag());
    auto inHalf = [&](unsigned long long i, int half) { return restrictBit < 0 || (int)((i >> restrictBit) & 1ULL) == half; };
    // the definition, on the (possibly restricted) index set; both halves in turn when restricted
    for (const PGOp& o : ops) {
        const hc c(o.c.x, o.c.y), f(o.f.x, o.f.y);
        if (!o.xy) { for (size_t i = 0; i < ref.size(); i++) ref[i] *= __builtin_parityll(i & o.yz) ? f : c; continue; }
        const int h = 63 - __builtin_clzll(o.xy);
        for (unsigned long long i = 0; i < ref.size(); i++) if (!((i >> h) & 1)) {
            const unsigned long long w = i ^ o.xy;
            const double si = __builtin_parityll(i & o.yz) ? -1.0 : 1.0, sw = __builtin_parityll(w & o.yz) ? -1.0 : 1.0;
            const hc x = ref[i], y = ref[w];
            ref[i] = c * x + f * sw * y; ref[w] = c * y + f * si * x;
        }
    }
    qb_state q; memset(&q, 0, sizeof q); q.numAmpsPerNode = 1LL << n; q.logNumAmpsPerNode = n; q.numQubits = n;
    int passes = 0;
    for (int half = 0; half < (restrictBit >= 0 ? 2 : 1); half++) {
        size_t i = 0;
        while (i < ops.size()) {
            // the planner's grouping rule (qb_tile.cu plan_passes): program order, independent masks, <= PG_K_GADGET strings
            unsigned long long masks[PG_K]; int nm = 0; size_t take = 0;
            for (; i + take < ops.size() && take < PG_MAX_OPS; take++) {
                const PGOp& o = ops[i + take];
                if (o.xy) {
                    if (nm == PG_K_GADGET) break;
                    masks[nm] = o.xy;
                    if (pg_rank(masks, nm + 1, nullptr) != nm + 1) break;
                    nm++;
                }
            }
            if (nm == 0) {       // only diagonal gadgets left in this stretch: apply them directly
                for (size_t j = 0; j < take; j++) for (size_t a = 0; a < state.size(); a++) if (inHalf(a, half)) state[a] = cmul(state[a], __builtin_parityll(a & ops[i + j].yz) ? ops[i + j].f : ops[i + j].c);
                i += take; continue;
            }
            PGDev d; int k = 0;
            const unsigned long long rm = restrictBit >= 0 ? (1ULL << restrictBit) : 0ULL;
            int r = pg_describe(&q, &ops[i], (int)take, rm, half ? rm : 0ULL, d, &k); if (r) return -2;
            const qindex groups = q.numAmpsPerNode >> (k + (restrictBit >= 0 ? 1 : 0));
            for (qindex g = 0; g < groups; g++) {
                switch (k) {
                case 1: pg_coset<1>(state.data(), g, d); break;
                case 2: pg_coset<2>(state.data(), g, d); break;
                case 3: pg_coset<3>(state.data(), g, d); break;
                case 4: pg_coset<4>(state.data(), g, d); break;
                default: pg_coset<5>(state.data(), g, d); break;
                }
            }
            passes++; i += take;
        }
    }
    double err = 0, norm = 0;
    for (size_t a = 0; a < ref.size(); a++) { err = std::max(err, std::abs(ref[a] - hc(state[a].x, state[a].y))); norm = std::max(norm, std::abs(ref[a])); }
    if (maxErr) *maxErr = norm > 0 ? err / norm : err;
    if (numPasses) *numPasses = passes;
    return 0;
}
#endif
