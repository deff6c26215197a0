// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A small SIMT interpreter that lets the CPU test suite execute the product's
// CUDA sources (simd-gaussian-ray-tracing_b200/csrc/vrt_cuda.cu + *.cuh) without a GPU, so that `pytest -m "not gpu"`
// exercises the real kernel and launch logic (indexing, list building, queues, bands, split cells) on small frames.
//
// It is NOT a CPU fallback of the product: nothing under simd-gaussian-ray-tracing_b200/ includes, links or loads it;
// tests/emu/build_emu.py rewrites a COPY of the sources (kernel launches -> emu::launch, inline PTX -> emu:: helpers,
// <cuda_runtime.h> -> this header) into tests/emu/_build/ and compiles that copy with g++ into libvrt_cuda_emu.so,
// which only tests/test_emu.py loads.  The product library stays libvrt_cuda.so and fails without a GPU.
//
// Execution model: one CTA at a time, every CUDA thread a fiber (own stack, cooperative switch).  A fiber runs until it
// reaches a warp collective (__shfl_sync, __ballot_sync, __any_sync, ... -- all built on one 32-lane all-gather) or
// __syncthreads, where it waits for the other live lanes of its warp / threads of its CTA.  The interpreter aborts with
// a message when lanes of a warp meet in DIFFERENT collectives, when a collective can never complete (deadlock), on
// __trap, and on writes outside a cudaMalloc'ed block (guard bands, checked after every launch).
// Arithmetic is IEEE fp32 with the GPU's flush-to-zero where the product asks for it (ex2.approx.ftz, rcp.approx.ftz are
// exact-rounded here, not the MUFU approximations), so results agree with the GPU to a few ulp, not bit for bit.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// ---- qualifiers --------------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__ static
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types ------------------------------------------------------------------------------------------------------
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3
{
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// ---- runtime (cuda_emu.cpp) -------------------------------------------------------------------------------------------
namespace emu
{
struct ThreadCtx
{
    uint3 tid, bid;
    dim3 bdim, gdim;
};
ThreadCtx &cur();
[[noreturn]] void fatal(const char *fmt, ...);
// 32-lane all-gather of one 64-bit value per lane; `kind` names the collective (lanes must agree on it)
void warp_gather(unsigned mask, uint64_t v, uint64_t out[32], int kind);
unsigned lane_id();
void block_barrier();
unsigned char *dynamic_smem();
void run_grid(dim3 grid, dim3 block, size_t smem, void (*body)(void *), void *arg);
template <class F>
static inline void launch(F &&f, dim3 grid, dim3 block, size_t smem = 0, void * /*stream*/ = nullptr)
{
    run_grid(grid, block, smem, [](void *p) { (*static_cast<F *>(p))(); }, &f);
}
// shared-memory addresses as 32-bit handles (the product keeps mbarrier / TMA addresses in uint32_t)
size_t to_shared(const void *p);
void *from_shared(uint32_t h);
void mbar_init(uint32_t bar, uint32_t count);
void mbar_expect_tx(uint32_t bar, uint32_t bytes);
uint32_t mbar_try_wait(uint32_t bar, uint32_t parity);
void bulk_copy(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar);
static inline float ftz(float x) { return std::fabs(x) < 1.17549435e-38f ? std::copysign(0.f, x) : x; }
static inline float ex2_ftz(float x) { return ftz(std::exp2(ftz(x))); }
static inline float rcp_ftz(float x) { return ftz(1.0f / ftz(x)); }
template <class T>
static inline uint64_t bits(T v)
{
    static_assert(sizeof(T) <= 8, "shuffle operand wider than 64 bits");
    uint64_t u = 0;
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T>
static inline T unbits(uint64_t u)
{
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}
// process-wide recursive lock taken by every C-ABI entry of the translated library (see build_emu.py)
struct ApiLock
{
    ApiLock();
    ~ApiLock();
};
enum { K_BALLOT = 1, K_ANY, K_ALL, K_SHFL, K_SHFL_XOR, K_SHFL_UP, K_SHFL_DOWN, K_REDUCE_MAX, K_REDUCE_MIN, K_SYNCWARP };
} // namespace emu

#define threadIdx (emu::cur().tid)
#define blockIdx (emu::cur().bid)
#define blockDim (emu::cur().bdim)
#define gridDim (emu::cur().gdim)

// ---- warp collectives ----------------------------------------------------------------------------------------------------
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    uint64_t o[32];
    emu::warp_gather(mask, pred ? 1u : 0u, o, emu::K_BALLOT);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= o[i] ? (1u << i) : 0u;
    return r & mask;
}
static inline int __any_sync(unsigned mask, int pred)
{
    uint64_t o[32];
    emu::warp_gather(mask, pred ? 1u : 0u, o, emu::K_ANY);
    for (int i = 0; i < 32; ++i)
        if (o[i]) return 1;
    return 0;
}
// lanes that already exited count as "true" (they do not take part)
static inline int __all_sync(unsigned mask, int pred)
{
    uint64_t o[32];
    emu::warp_gather(mask, pred ? 0u : 1u, o, emu::K_ALL);
    for (int i = 0; i < 32; ++i)
        if (o[i]) return 0;
    return 1;
}
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src)
{
    uint64_t o[32];
    emu::warp_gather(mask, emu::bits(v), o, emu::K_SHFL);
    return emu::unbits<T>(o[src & 31]);
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask)
{
    uint64_t o[32];
    emu::warp_gather(mask, emu::bits(v), o, emu::K_SHFL_XOR);
    return emu::unbits<T>(o[(emu::lane_id() ^ (unsigned)lane_mask) & 31]);
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta)
{
    uint64_t o[32];
    emu::warp_gather(mask, emu::bits(v), o, emu::K_SHFL_UP);
    const unsigned l = emu::lane_id();
    return l >= delta ? emu::unbits<T>(o[l - delta]) : v;
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta)
{
    uint64_t o[32];
    emu::warp_gather(mask, emu::bits(v), o, emu::K_SHFL_DOWN);
    const unsigned l = emu::lane_id();
    return l + delta < 32 ? emu::unbits<T>(o[l + delta]) : v;
}
static inline int __reduce_max_sync(unsigned mask, int v)
{
    uint64_t o[32];
    emu::warp_gather(mask, emu::bits(v) | (1ull << 40), o, emu::K_REDUCE_MAX);
    int r = INT32_MIN;
    for (int i = 0; i < 32; ++i)
        if (o[i] >> 40) r = std::max(r, emu::unbits<int>(o[i] & 0xffffffffull));
    return r;
}
static inline int __reduce_min_sync(unsigned mask, int v)
{
    uint64_t o[32];
    emu::warp_gather(mask, emu::bits(v) | (1ull << 40), o, emu::K_REDUCE_MIN);
    int r = INT32_MAX;
    for (int i = 0; i < 32; ++i)
        if (o[i] >> 40) r = std::min(r, emu::unbits<int>(o[i] & 0xffffffffull));
    return r;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu)
{
    uint64_t o[32];
    emu::warp_gather(mask, 0, o, emu::K_SYNCWARP);
}
static inline void __syncthreads() { emu::block_barrier(); }
#define __trap() emu::fatal("__trap() at %s:%d", __FILE__, __LINE__)

// ---- integer / float intrinsics ------------------------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline unsigned __float_as_uint(float f) { return emu::unbits<unsigned>(emu::bits(f)); }
static inline float __uint_as_float(unsigned u) { return emu::unbits<float>(u); }
static inline int __float_as_int(float f) { return emu::unbits<int>(emu::bits(f)); }
static inline float __int_as_float(int i) { return emu::unbits<float>(emu::bits(i)); }
static inline int __float2int_rn(float x) { return (int)std::lrintf(x); } // round to nearest even (default rounding mode)
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline size_t __cvta_generic_to_shared(const void *p) { return emu::to_shared(p); }

// CUDA's mixed-type min / max overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }

// atomics: one CTA runs at a time on one host thread, so plain read-modify-write is atomic by construction
template <class T, class U>
static inline T atomicAdd(T *p, U v)
{
    const T old = *p;
    *p = old + (T)v;
    return old;
}
template <class T, class U>
static inline T atomicMax(T *p, U v)
{
    const T old = *p;
    if ((T)v > old) *p = (T)v;
    return old;
}

template <class T, class U>
static inline T atomicOr(T *p, U v)
{
    const T old = *p;
    *p = old | (T)v;
    return old;
}

// ---- the slice of the runtime API the product uses --------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct emuStream;
struct emuEvent { std::chrono::steady_clock::time_point t; };
typedef emuStream *cudaStream_t;
typedef emuEvent *cudaEvent_t;
struct cudaDeviceProp
{
    int major, minor, multiProcessorCount;
    char name[64];
};
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaGetLastError();
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemsetAsync(void *dst, int v, size_t n, cudaStream_t s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaEventDestroy(cudaEvent_t e);
// mapped / registered host memory: host and "device" share one address space here, so these are bookkeeping only
enum { cudaErrorNotReady = 600 };
enum { cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2, cudaEventDisableTiming = 2, cudaHostRegisterDefault = 0, cudaHostRegisterPortable = 1, cudaHostRegisterMapped = 2 };
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; };
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = std::calloc(1, n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
// CUDA IPC (peer images): one process, one address space here -- exporting works, importing is refused
enum { cudaErrorNotSupported = 801, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaErrorNotSupported; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; } // kernels complete inside their launch
template <class T>
static inline cudaError_t cudaMemcpyToSymbolAsync(T &symbol, const void *src, size_t n, size_t offset, cudaMemcpyKind, cudaStream_t)
{
    if (offset + n > sizeof(T)) return cudaErrorInvalidValue;
    std::memcpy(reinterpret_cast<char *>(&symbol) + offset, src, n);
    return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaMemcpyToSymbol(T &symbol, const void *src, size_t n)
{
    return cudaMemcpyToSymbolAsync(symbol, src, n, 0, cudaMemcpyHostToDevice, nullptr);
}
template <class F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t)
{
    *n = 1;
    return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int)
{
    return cudaSuccess;
}
