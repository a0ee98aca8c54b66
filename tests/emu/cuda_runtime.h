// tests/emu/cuda_runtime.h — TEST INFRASTRUCTURE, never part of the product.
//
// A warp-synchronous CPU emulator of the small CUDA subset libvcfx_cuda uses, so that the
// *unchanged kernel source* (vcfx_b200/csrc/vcfx_kernels.cuh + vcfx_api.cu) can be compiled with
// g++ into tests/emu/_build/libvcfx_emu.so and checked against the oracle by the "not gpu" tests
// (tests/test_emu_parity.py) before any GPU time is spent.  Every CUDA thread is a fiber; a fiber
// runs until it reaches a warp collective (__ballot_sync, __shfl_*_sync, __any_sync,
// __reduce_add_sync, __syncwarp) or __syncthreads, the collective is resolved when all 32 lanes
// (all threads of the CTA) have arrived, and a warp whose lanes arrive at DIFFERENT collectives, or
// exit while others wait, is reported as an error (that is a bug on the GPU as well).  CTAs run one
// after the other.  Nothing under vcfx_b200/ loads this library: the product has no CPU path.
#pragma once
#ifndef VCFX_EMU
#error "tests/emu/cuda_runtime.h is only for the emulator build (-DVCFX_EMU)"
#endif
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
// every system header the product sources use comes in BEFORE the CUDA keywords become macros
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = {x, y}; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }

namespace emu {
enum Op { OP_NONE, OP_BALLOT, OP_ANY, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_REDUCE_ADD, OP_REDUCE_OR, OP_SYNCWARP, OP_SYNCTHREADS };
struct Fiber {
    void *sp = nullptr; char *stack = nullptr;
    int state = 0;              // 0 runnable, 1 waiting (warp), 2 waiting (cta), 3 done
    int op = OP_NONE; unsigned long long in = 0, out = 0; int arg = 0;
    dim3 tid;
};
extern dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
extern Fiber *g_cur;
extern void *g_dyn_smem;
unsigned long long collective(int op, unsigned long long v, int arg);
void launch(void (*entry)(void *), void *args, size_t arg_bytes, dim3 grid, dim3 block, size_t smem);
inline void *dyn_smem() { return g_dyn_smem; }
}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

// ---- warp collectives (mask is always FULL in this code base; anything else is rejected)
static inline void emu_mask(unsigned m) { if (m != 0xFFFFFFFFu) { fprintf(stderr, "emu: partial warp mask %08x not supported\n", m); abort(); } }
static inline unsigned __ballot_sync(unsigned m, int p) { emu_mask(m); return (unsigned)emu::collective(emu::OP_BALLOT, p ? 1 : 0, 0); }
static inline int __any_sync(unsigned m, int p) { emu_mask(m); return (int)emu::collective(emu::OP_ANY, p ? 1 : 0, 0); }
static inline int __all_sync(unsigned m, int p) { emu_mask(m); return !(int)emu::collective(emu::OP_ANY, p ? 0 : 1, 0); }
static inline void __syncwarp(unsigned m = 0xFFFFFFFFu) { emu_mask(m); emu::collective(emu::OP_SYNCWARP, 0, 0); }
static inline void __syncthreads() { emu::collective(emu::OP_SYNCTHREADS, 0, 0); }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { emu_mask(m); return (unsigned)emu::collective(emu::OP_REDUCE_ADD, v, 0); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { emu_mask(m); return (unsigned)emu::collective(emu::OP_REDUCE_OR, v, 0); }
template <class T> static inline T emu_shfl(int op, T v, int arg) {
    static_assert(sizeof(T) <= 8, "shfl of at most 64 bits");
    unsigned long long u = 0; memcpy(&u, &v, sizeof(T));
    u = emu::collective(op, u, arg);
    T r; memcpy(&r, &u, sizeof(T)); return r;
}
template <class T> static inline T __shfl_sync(unsigned m, T v, int src) { emu_mask(m); return emu_shfl(emu::OP_SHFL, v, src & 31); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) { emu_mask(m); return emu_shfl(emu::OP_SHFL_UP, v, (int)d); }
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) { emu_mask(m); return emu_shfl(emu::OP_SHFL_DOWN, v, (int)d); }

// ---- scalar intrinsics
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; s = s > 32 ? 32 : s; return (unsigned)(v >> s); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32); }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; s = s > 32 ? 32 : s; return (unsigned)((s == 32 ? (v << 16 << 16) : (v << s)) >> 32); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    unsigned long long pool = ((unsigned long long)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; ++i) { unsigned sel = (s >> (4 * i)) & 0xF; unsigned byte = (unsigned)(pool >> (8 * (sel & 7))) & 0xFF; if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0; r |= byte << (8 * i); }
    return r;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b) { typedef typename std::common_type<A, B>::type C; return (C)a < (C)b ? (C)a : (C)b; }
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b) { typedef typename std::common_type<A, B>::type C; return (C)a > (C)b ? (C)a : (C)b; }

// ---- atomics: fibers never run concurrently
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }

// ---- runtime API subset
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
typedef struct emu_stream_t { int id; } *cudaStream_t;
typedef struct emu_event_t { double t_ms; } *cudaEvent_t;
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaHostAllocWriteCombined = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline int emu_env_int(const char *k, int d) { const char *v = getenv(k); return v && *v ? atoi(v) : d; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = emu_env_int("VCFX_EMU_DEVICES", 1); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof *p); p->multiProcessorCount = emu_env_int("VCFX_EMU_SMS", 2); strcpy(p->name, "vcfx warp emulator"); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = emu_env_int("VCFX_EMU_CTAS_PER_SM", 1); return cudaSuccess; }
// device allocations carry a poisoned guard band on both sides so that stray writes show up
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
    char *raw = (char *)aligned_alloc(256, ((n + 255) & ~(size_t)255) + 512);
    if (!raw) return cudaErrorMemoryAllocation;
    memset(raw, 0xA5, ((n + 255) & ~(size_t)255) + 512);
    *p = (T *)(raw + 256); return cudaSuccess;
}
static inline cudaError_t cudaFree(void *p) { if (p) free((char *)p - 256); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)aligned_alloc(4096, (n + 4095) & ~(size_t)4095); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emu_stream_t{1}; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new emu_stream_t{1}; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event_t{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event_t{0}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    e->t_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }

// ---- kernel launch: the product source launches through VCFX_LAUNCH (kern<<<...>>>(args) under nvcc)
template <class P> struct emu_call { void (*fn)(const P); P p; };
template <class P> static void emu_entry(void *a) { emu_call<P> *c = (emu_call<P> *)a; c->fn(c->p); }
template <class P, class A> static inline void emu_launch(void (*fn)(const P), dim3 grid, dim3 block, size_t smem, cudaStream_t, const A &arg) {
    emu_call<P> c = {fn, (P)arg};
    emu::launch(emu_entry<P>, &c, sizeof c, grid, block, smem);
}
#define VCFX_LAUNCH(kern, grid, block, smem, stream, arg) emu_launch((kern), dim3(grid), dim3(block), (smem), (stream), (arg))
