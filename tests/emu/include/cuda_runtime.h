// cuda_runtime.h -- TEST INFRASTRUCTURE, not product code.
//
// A host-side SIMT emulation shim: with `-I tests/emu/include` g++ compiles the UNMODIFIED kernel sources
// (manusim_b200/csrc/msim_kernel.cu, msim_api.cu, philox.cuh) into tests/emu/_build/libmsim_emu.so, in which every
// CUDA thread is a fiber and every warp/block primitive (__shfl_sync, __reduce_*_sync, __match_any_sync,
// __ballot_sync, __syncwarp, __syncthreads) is a rendezvous of the participating fibers.  It exists for two jobs that
// need no GPU:
//   * run the kernel source under AddressSanitizer / UBSan (compute-sanitizer is closed on the GPU pool);
//   * check event-order logic changes bit for bit against the reference goldens and the CPU oracle before GPU
//     minutes are spent on them.
// Nothing in manusim_b200/ loads this library; the product path is libmsim.so on a CUDA device or an exception.
//
// Fidelity: integer work, IEEE float/double add/sub/mul/div/rint and float<->double conversion are exact, so REPLAY
// mode (which touches no transcendental) is bit-identical to the GPU.  Philox mode uses libm instead of the CUDA
// fast-math intrinsics (__logf, __powf, cospif) and no FMA contraction, so its variates differ from the GPU's in the
// last bits; the record -> oracle-replay tests are self-consistent on either side.
#pragma once

// system headers first: the qualifier macros below (__noinline__ in particular) must not be visible to them
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <string>
#include <vector>

#ifndef __CUDACC__
#define __CUDACC__ 1   // philox.cuh guards its device-only code with it
#endif
#define MSIM_EMU 1

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline)) inline   // vague linkage: device functions live in headers shared by several .cu files
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local   // one block at a time per host thread

// ------------------------------------------------------------------------------------------------ vector types
struct uint4 { uint32_t x, y, z, w; } __attribute__((aligned(16)));
struct float2 { float x, y; } __attribute__((aligned(8)));
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct emu_dim3 { unsigned x, y, z; };

// ------------------------------------------------------------------------------------------------ fiber runtime
namespace emu {
struct Fiber;
extern thread_local Fiber* g_cur;
emu_dim3 tid();
emu_dim3 bid();
emu_dim3 bdim();
unsigned char* dyn_smem();
// every participating lane contributes `v`; returns a pointer to the 32 contributions of the warp (valid until the
// caller's next collective).  `kind` names the primitive: lanes meeting at different primitives is an error.
enum Kind { K_SYNCWARP = 1, K_SHFL, K_SHFL_XOR, K_BALLOT, K_MATCH, K_REDUX_MIN, K_REDUX_ADD, K_REDUX_MAX };
const uint64_t* warp_gather(unsigned mask, uint64_t v, int kind);
void block_barrier();
// run `body` once per thread of a grid x block launch (blocks spread over host threads)
void launch(long long grid, int block, size_t smem, const std::function<void()>& body);
int last_error();
}  // namespace emu

#define threadIdx (emu::tid())
#define blockIdx (emu::bid())
#define blockDim (emu::bdim())

#define MSIM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((long long)(grid), (int)(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })
#define MSIM_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()

// ------------------------------------------------------------------------------------------------ warp primitives
template <typename T>
static inline uint64_t emu_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
static inline T emu_unbits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    (void)width;
    const uint64_t* all = emu::warp_gather(mask, emu_bits(v), emu::K_SHFL);
    return emu_unbits<T>(all[src & 31]);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    (void)width;
    const int lane = (int)(emu::tid().x & 31u);
    const uint64_t* all = emu::warp_gather(mask, emu_bits(v), emu::K_SHFL_XOR);
    return emu_unbits<T>(all[(lane ^ lane_mask) & 31]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    const uint64_t* all = emu::warp_gather(mask, pred ? 1u : 0u, emu::K_BALLOT);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && all[i]) r |= 1u << i;
    return r;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
    const uint64_t* all = emu::warp_gather(mask, v, emu::K_MATCH);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && (unsigned)all[i] == v) r |= 1u << i;
    return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    const uint64_t* all = emu::warp_gather(mask, v, emu::K_REDUX_MIN);
    unsigned r = 0xFFFFFFFFu;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r = std::min(r, (unsigned)all[i]);
    return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    const uint64_t* all = emu::warp_gather(mask, v, emu::K_REDUX_MAX);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r = std::max(r, (unsigned)all[i]);
    return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    const uint64_t* all = emu::warp_gather(mask, v, emu::K_REDUX_ADD);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r += (unsigned)all[i];
    return r;
}
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { (void)emu::warp_gather(mask, 0, emu::K_SYNCWARP); }
static inline void __syncthreads() { emu::block_barrier(); }

// ------------------------------------------------------------------------------------------------ scalar intrinsics
// (compile with -ffp-contract=off: the *_rn intrinsics are single IEEE operations)
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline uint32_t __float_as_uint(float f) { return emu_unbits<uint32_t>(emu_bits(f)); }
static inline float __uint_as_float(uint32_t u) { return emu_unbits<float>(emu_bits(u)); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
#define __logf(x) logf(x)      // (glibc declares these two names itself, hence macros)
#define __powf(x, y) powf(x, y)
static inline float cospif(float x) { return (float)cos(3.14159265358979323846 * (double)x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }

// ------------------------------------------------------------------------------------------------ runtime API stubs
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; };
struct cudaFuncAttributes { int numRegs; };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated launch failed"; }
static inline cudaError_t cudaGetLastError() { return emu::last_error() ? cudaErrorEmu : cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    const char* s = getenv("MSIM_EMU_SMS");
    p->multiProcessorCount = s ? atoi(s) : 2;   // few "SMs": a small persistent grid that still exercises the work counter
    p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void* pp, size_t n) {
    void* p = malloc(n ? n : 1);
    *reinterpret_cast<void**>(pp) = p;
    return p ? cudaSuccess : cudaErrorEmu;
}
template <typename T>
static inline cudaError_t cudaMalloc(T** pp, size_t n) { return cudaMalloc(reinterpret_cast<void*>(pp), n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    if (n) memcpy(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <typename F>
static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { a->numRegs = 0; return cudaSuccess; }
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int block, size_t smem) {
    // shared-memory bound (228 KB per SM, 1 KB reserved per CTA), 2048 threads and 32 CTAs per SM.
    // MSIM_EMU_WPB=w admits only CTAs of w warps: w=1 gives every replication slab its own allocation, so that
    // AddressSanitizer sees an overflow from one slab into the next.
    const char* force = getenv("MSIM_EMU_WPB");
    if (force && atoi(force) * 32 != block) { *n = 0; return cudaSuccess; }
    int by_smem = (int)((228 * 1024) / (smem + 1024));
    int by_thr = 2048 / (block > 0 ? block : 1);
    *n = std::min(std::min(by_smem, by_thr), 32);
    return cudaSuccess;
}
