// qb_runtime.cu -- device binding, memory, copies, stream and scratch management.
// Replaces quest/src/gpu/gpu_config.cpp (hardware queries :124-283, binding :332-379,
// sync :382, alloc :399-431, copies :449-623, cache :635-687) behind the C ABI.
#include "qb_common.cuh"
#include "qb_reduce.cuh"
#ifdef QB_SELFTEST
#include "../../include/quest_b200_selftest.h"
#endif
#include <string.h>
#include <string>
#include <algorithm>

QbRuntime g_qb;

static thread_local std::string t_lastError = "no error";

int qb_set_error(int code, const char* what, const char* file, int line) {
    char buf[1024];
    const char* desc = (code > 0) ? cudaGetErrorString((cudaError_t)code) : "invalid argument / precondition";
    snprintf(buf, sizeof buf, "quest_b200: %s failed (%d: %s) at %s:%d", what, code, desc, file, line);
    t_lastError = buf;
    if (code > 0) cudaGetLastError();  // clear sticky-less errors so later calls report their own
    return code == 0 ? -1 : code;
}

int qb_ensure_ready() {
    if (g_qb.device >= 0) return 0;
    return qb_bind_device(0);
}

// ------------------------------------------------------------------------------------------
// host helpers for qubit lists
// ------------------------------------------------------------------------------------------
BitIns qb_make_ins(const int* a, const int* aStates, int na, const int* b, const int* bStates, int nb) {
    // == util_getSorted(a, b) + util_getBitMask(a, aStates, b, bStates)  (core/utilities.cpp:188-216)
    BitIns ins;
    memset(&ins, 0, sizeof ins);
    ins.n = na + nb;
    int tmp[2 * QB_MAX_QUBITS + 2];
    for (int i = 0; i < na; i++) { tmp[i] = a[i]; if (aStates && aStates[i]) ins.mask |= 1ULL << a[i]; }
    for (int i = 0; i < nb; i++) { tmp[na + i] = b[i]; if (bStates && bStates[i]) ins.mask |= 1ULL << b[i]; }
    std::sort(tmp, tmp + ins.n);
    unsigned long long inserted = 0;
    for (int i = 0; i < ins.n; i++) inserted |= 1ULL << tmp[i];
    ins.p0 = ins.n > 0 ? tmp[0] : 0; ins.p1 = ins.n > 1 ? tmp[1] : 0;
    ins.p2 = ins.n > 2 ? tmp[2] : 0; ins.p3 = ins.n > 3 ? tmp[3] : 0;
    ins.keep = ~inserted;
    // move masks of the expand operation (Hacker's Delight, fig. 7-12, widened to 64 bits)
    unsigned long long m = ins.keep, mk = ~m << 1;
    for (int i = 0; i < 6; i++) {
        unsigned long long mp = mk ^ (mk << 1);
        mp ^= mp << 2; mp ^= mp << 4; mp ^= mp << 8; mp ^= mp << 16; mp ^= mp << 32;
        unsigned long long mv = mp & m;
        ins.mv[i] = mv;
        m = (m ^ mv) | (mv >> (1 << i));
        mk &= ~mp;
    }
    return ins;
}

BitList qb_make_list(const int* q, int n) {
    BitList l;
    l.n = n;
    for (int i = 0; i < n; i++) l.q[i] = (unsigned char)q[i];
    return l;
}

unsigned long long qb_make_mask(const int* q, int n) {
    unsigned long long m = 0;
    for (int i = 0; i < n; i++) m |= 1ULL << q[i];
    return m;
}

int qb_check_qubits(const int* q, int n, int limit) {
    if (n < 0 || n > QB_MAX_QUBITS) return 0;
    if (n > 0 && !q) return 0;
    for (int i = 0; i < n; i++) if (q[i] < 0 || q[i] >= limit) return 0;
    return 1;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int qb_abi_version(void) { return 1; }
int qb_precision(void) { return QB_PRECISION; }

const char* qb_error_string(void) { return t_lastError.c_str(); }

int qb_num_devices(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int qb_is_device_available(void) {
    int n = qb_num_devices();
    for (int d = 0; d < n; d++) {
        int major = 9999;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major != 9999)
            return 1;
    }
    cudaGetLastError();
    return 0;
}

int qb_bind_device(int dev) {
    int n = qb_num_devices();
    if (n <= 0) return qb_set_error((int)cudaErrorNoDevice, "qb_bind_device (no CUDA device; this backend has no CPU fallback)", __FILE__, __LINE__);
    QB_REQUIRE(dev >= 0 && dev < n, "qb_bind_device: device index out of range");
    QB_CUDA(cudaSetDevice(dev));
    if (g_qb.device == dev) return 0;
    g_qb.device = dev;
    QB_CUDA(cudaDeviceGetAttribute(&g_qb.numSMs, cudaDevAttrMultiProcessorCount, dev));
    // reduction scratch
    QB_CUDA(cudaMalloc(&g_qb.redPartials, sizeof(double) * QB_RED_SCRATCH_DOUBLES));
    QB_CUDA(cudaMalloc(&g_qb.redTicket, sizeof(unsigned int)));
    QB_CUDA(cudaMemset(g_qb.redTicket, 0, sizeof(unsigned int)));
    QB_CUDA(cudaMalloc(&g_qb.redOutDev, sizeof(double) * QB_RED_MAX_OUT));
    QB_CUDA(cudaMallocHost(&g_qb.redOutHost, sizeof(double) * QB_RED_MAX_OUT));
    return 0;
}

int qb_bound_device(void) { return g_qb.device; }

int qb_compute_capability(void) {
    if (qb_ensure_ready()) return -1;
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, g_qb.device);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, g_qb.device);
    return major * 10 + minor;
}

int qb_mem_info(size_t* freeBytes, size_t* totalBytes) {
    QB_READY();
    size_t f, t;
    QB_CUDA(cudaMemGetInfo(&f, &t));
    if (freeBytes) *freeBytes = f;
    if (totalBytes) *totalBytes = t;
    return 0;
}

int qb_supports_mem_pools(void) {
    if (qb_ensure_ready()) return 0;
    int s = 0;
    cudaDeviceGetAttribute(&s, cudaDevAttrMemoryPoolsSupported, g_qb.device);
    return s;
}

qb_index qb_max_concurrent_threads(void) {
    if (qb_ensure_ready()) return -1;
    int perBlock = 0;
    cudaDeviceGetAttribute(&perBlock, cudaDevAttrMaxThreadsPerBlock, g_qb.device);
    return (qb_index)perBlock * g_qb.numSMs;
}

int qb_device_uuid(char out16[16]) {
    QB_READY();
    cudaDeviceProp prop;
    QB_CUDA(cudaGetDeviceProperties(&prop, g_qb.device));
    memcpy(out16, prop.uuid.bytes, 16);
    return 0;
}

int qb_flush(void) {
    if (g_qb.device < 0) return 0;
    QB_FLUSH();
    return 0;
}

int qb_sync(void) {
    QB_READY();
    // gpu_sync() is a full device sync in the reference (gpu_config.cpp:382-390); the library issues
    // work on g_qb.stream and helper streams, so synchronise the whole device here too.
    QB_CUDA(cudaDeviceSynchronize());
    return 0;
}

void* qb_get_stream(void) { return (void*)g_qb.stream; }

int qb_set_stream(void* s) {
    QB_READY();
    g_qb.stream = (cudaStream_t)s;
    return 0;
}

qb_cplx* qb_alloc(qb_index numAmps, int* status) {
    if (status) *status = 0;
    int r = qb_ensure_ready();
    if (r) { if (status) *status = r; return nullptr; }
    if (numAmps <= 0) return nullptr;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)numAmps * sizeof(qb_cplx));
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return nullptr; }  // soft failure, like gpu_config.cpp:406-413
    if (e != cudaSuccess) { int c = qb_set_error((int)e, "cudaMalloc", __FILE__, __LINE__); if (status) *status = c; return nullptr; }
    qb_p2p_note_alloc(p, (size_t)numAmps * sizeof(qb_cplx));
    return (qb_cplx*)p;
}

int qb_free(qb_cplx* p) {
    if (!p) return 0;
    QB_READY();
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    qb_tile_forget(p);
    { int r = qb_p2p_note_free(p); if (r) return r; }
    QB_CUDA(cudaFree(p));
    return 0;
}

int qb_copy_h2d(qb_cplx* dev, const qb_cplx* host, qb_index n) {
    QB_READY();
    if (n <= 0) return 0;
    QB_CUDA(cudaMemcpyAsync(dev, host, (size_t)n * sizeof(qb_cplx), cudaMemcpyHostToDevice, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    return 0;
}

int qb_copy_d2h(qb_cplx* host, const qb_cplx* dev, qb_index n) {
    QB_READY();
    if (n <= 0) return 0;
    QB_CUDA(cudaMemcpyAsync(host, dev, (size_t)n * sizeof(qb_cplx), cudaMemcpyDeviceToHost, g_qb.stream));
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    return 0;
}

int qb_copy_d2d(qb_cplx* dst, const qb_cplx* src, qb_index n) {
    QB_READY();
    if (n <= 0) return 0;
    QB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(qb_cplx), cudaMemcpyDeviceToDevice, g_qb.stream));
    return 0;
}

qb_cplx* qb_get_cache(qb_index numElems, int* status) {
    if (status) *status = 0;
    int r = qb_ensure_ready();
    if (r) { if (status) *status = r; return nullptr; }
    if (numElems <= g_qb.cacheLen) return (qb_cplx*)g_qb.cache;
    cudaStreamSynchronize(g_qb.stream);
    if (g_qb.cache) cudaFree(g_qb.cache);
    g_qb.cache = nullptr; g_qb.cacheLen = 0;
    cudaError_t e = cudaMalloc(&g_qb.cache, (size_t)numElems * sizeof(cplx));
    if (e != cudaSuccess) { int c = qb_set_error((int)e, "cudaMalloc(cache)", __FILE__, __LINE__); if (status) *status = c; return nullptr; }
    g_qb.cacheLen = numElems;
    return (qb_cplx*)g_qb.cache;
}

int qb_clear_cache(void) {
    if (g_qb.device < 0) return 0;
    QB_CUDA(cudaStreamSynchronize(g_qb.stream));
    if (g_qb.cache) QB_CUDA(cudaFree(g_qb.cache));
    g_qb.cache = nullptr; g_qb.cacheLen = 0;
    return 0;
}

size_t qb_cache_bytes(void) { return (size_t)g_qb.cacheLen * sizeof(cplx); }

unsigned long long qb_launch_count(void) { return g_qb.launches; }

int qb_set_tile_engine(int enabled) {
    if (g_qb.device >= 0) QB_FLUSH();
    g_qb.tileEngine = (enabled < 0 || enabled > 2) ? 1 : enabled;
    return 0;
}

int qb_statevec_getAmp_sub(const qb_state* q, qb_index ind, qb_cplx* out) {
    QB_REQUIRE(q && q->amps && out && ind >= 0 && ind < q->numAmpsPerNode, "getAmp: bad arguments");
    return qb_copy_d2h(out, q->amps + ind, 1);
}

#ifdef QB_SELFTEST
// host-side self test of the index algebra used by every kernel: BitIns against the literal
// one-bit-at-a-time definition (core/bitwise.hpp:99-105,164-171,206-210). Returns #mismatches.
int qb_selftest_bitins(const int* qubits, const int* states, int n, qb_index item, qb_index* out) {
    BitIns ins = qb_make_ins(qubits, states, n, nullptr, nullptr, 0);
    qindex got = ins(item);
    int sorted[QB_MAX_QUBITS + 1];
    for (int i = 0; i < n; i++) sorted[i] = qubits[i];
    std::sort(sorted, sorted + n);
    qindex want = item;
    for (int i = 0; i < n; i++) want = insertZeroBit(want, sorted[i]);
    for (int i = 0; i < n; i++) if (states && states[i]) want |= pow2(qubits[i]);
    if (out) *out = got;
    return got != want;
}

#endif  // QB_SELFTEST

} // extern "C"
