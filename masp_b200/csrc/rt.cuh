// Thin runtime layer: kernel definition / launch macros, device memory and
// error handling.  Compiled two ways:
//   * nvcc, sm_100a: the product (libmasp_b200.so);
//   * g++ with -DMB200_EMU: every kernel body runs as a plain loop over its
//     thread ids.  That build exists only so the CPU test-suite can exercise
//     the launch orchestration and index arithmetic without a GPU
//     (tests/emu/); it is never loaded by the masp_b200 package and is not a
//     fallback of any product path.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/masp_b200.h"  // MB200_E* codes

#ifdef MB200_EMU
#define MB_HD inline
#define MB_D inline
#define MB_COLD static inline
#define MB_UNROLL
#define MB_NOUNROLL
typedef int cudaStream_t;
typedef int cudaError_t;
#else
#include <cuda_runtime.h>
#define MB_HD __host__ __device__ __forceinline__
#define MB_D __device__ __forceinline__
#define MB_COLD static __host__ __device__ __noinline__
#define MB_UNROLL _Pragma("unroll")
#define MB_NOUNROLL _Pragma("unroll 1")
#endif

namespace mb {

struct Error {
    int code;
    char msg[256];
};
inline Error& last_error() {
    static thread_local Error e = {0, {0}};
    return e;
}
struct Exc {
    int code;
};
[[noreturn]] inline void fail(int code, const char* fmt, const char* a = "", long b = 0) {
    Error& e = last_error();
    e.code = code;
    snprintf(e.msg, sizeof e.msg, fmt, a, b);
    throw Exc{code};
}

extern std::atomic<unsigned long long> g_launches;  // kernels launched by this library, all devices (bench "gpu_launches")

#ifndef MB200_EMU
#define MB_CUDA(x)                                                                       \
    do {                                                                                 \
        cudaError_t _e = (x);                                                            \
        if (_e != cudaSuccess) ::mb::fail(MB200_ECUDA, "CUDA: %s (line %ld)", cudaGetErrorString(_e), __LINE__); \
    } while (0)

// cudaFuncSetAttribute is per device: run `f` the first time each device reaches a launch site.
template <class Fn>
inline void once_per_device(std::atomic<unsigned long long>& done, Fn f) {
    int dev = 0;
    MB_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return;
    f();
    done.fetch_or(bit, std::memory_order_release);
}

// One thread per work item; Args carries `size_t nthreads`.  Every kernel is
// declared wherever its header is included and defined in exactly one
// translation unit (the k_*.cu file that sets the header's MB_DEFINE_* macro),
// so the units compile in parallel.
#define MB_KERNEL_DECL(name, Args) void launch_##name(const Args& a, cudaStream_t s);
#define MB_KERNEL_DEF(name, Args, body, BLOCK)                                           \
    __global__ void __launch_bounds__(BLOCK) name(const Args a) {                        \
        size_t tid = (size_t)blockIdx.x * BLOCK + threadIdx.x;                           \
        if (tid < a.nthreads) body(a, tid);                                              \
    }                                                                                    \
    void launch_##name(const Args& a, cudaStream_t s) {                                  \
        if (!a.nthreads) return;                                                         \
        name<<<(unsigned)((a.nthreads + BLOCK - 1) / BLOCK), BLOCK, 0, s>>>(a);          \
        MB_CUDA(cudaGetLastError());                                                     \
        ::mb::g_launches++;                                                              \
    }
// same, with a minimum number of resident blocks per SM (caps the registers per thread)
#define MB_KERNEL_DEF_OCC(name, Args, body, BLOCK, MINB)                                 \
    __global__ void __launch_bounds__(BLOCK, MINB) name(const Args a) {                  \
        size_t tid = (size_t)blockIdx.x * BLOCK + threadIdx.x;                           \
        if (tid < a.nthreads) body(a, tid);                                              \
    }                                                                                    \
    void launch_##name(const Args& a, cudaStream_t s) {                                  \
        if (!a.nthreads) return;                                                         \
        name<<<(unsigned)((a.nthreads + BLOCK - 1) / BLOCK), BLOCK, 0, s>>>(a);          \
        MB_CUDA(cudaGetLastError());                                                     \
        ::mb::g_launches++;                                                              \
    }
template <class T>
__host__ __device__ __forceinline__ T mb_atomic_add(T* p, T v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    T o = *p;  // host instantiation is never executed in the product build
    *p = o + v;
    return o;
#endif
}
#define MB_ATOMIC_ADD(ptr, v) ::mb::mb_atomic_add((ptr), (v))

inline void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) fail(MB200_ENOMEM, "cudaMalloc of %s%ld bytes failed", "", (long)bytes);
    return p;
}
inline void dev_free(void* p) {
    if (p) cudaFree(p);
}
inline void* host_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) fail(MB200_ENOMEM, "cudaMallocHost of %s%ld bytes failed", "", (long)bytes);
    return p;
}
inline void host_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}
inline void dev_memset(void* p, int v, size_t bytes, cudaStream_t s) { MB_CUDA(cudaMemsetAsync(p, v, bytes, s)); }
inline void copy_h2d(void* d, const void* h, size_t bytes, cudaStream_t s) {
    MB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
}
inline void copy_d2h(void* h, const void* d, size_t bytes, cudaStream_t s) {
    MB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
}
inline void copy_d2d(void* d, const void* s_, size_t bytes, cudaStream_t s) {
    MB_CUDA(cudaMemcpyAsync(d, s_, bytes, cudaMemcpyDeviceToDevice, s));
}
inline void stream_sync(cudaStream_t s) { MB_CUDA(cudaStreamSynchronize(s)); }
#else  // ---------------------------------------------------------------- EMU
#define MB_CUDA(x) (void)(x)
#define MB_KERNEL_DECL(name, Args) void launch_##name(const Args& a, cudaStream_t s);
#define MB_KERNEL_DEF(name, Args, body, BLOCK)                                           \
    void launch_##name(const Args& a, cudaStream_t) {                                    \
        for (size_t tid = 0; tid < a.nthreads; ++tid) body(a, tid);                      \
        ::mb::g_launches++;                                                              \
    }
#define MB_KERNEL_DEF_OCC(name, Args, body, BLOCK, MINB) MB_KERNEL_DEF(name, Args, body, BLOCK)
template <class T>
inline T emu_atomic_add(T* p, T v) {
    T o = *p;
    *p = o + v;
    return o;
}
#define MB_ATOMIC_ADD(ptr, v) ::mb::emu_atomic_add((ptr), (v))
inline void* dev_alloc(size_t bytes) {
    void* p = malloc(bytes ? bytes : 16);
    if (!p) fail(MB200_ENOMEM, "malloc of %s%ld bytes failed", "", (long)bytes);
    return p;
}
inline void dev_free(void* p) { free(p); }
inline void* host_alloc_pinned(size_t bytes) { return dev_alloc(bytes); }
inline void host_free_pinned(void* p) { free(p); }
inline void dev_memset(void* p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); }
inline void copy_h2d(void* d, const void* h, size_t bytes, cudaStream_t) { memcpy(d, h, bytes); }
inline void copy_d2h(void* h, const void* d, size_t bytes, cudaStream_t) { memcpy(h, d, bytes); }
inline void copy_d2d(void* d, const void* s_, size_t bytes, cudaStream_t) { memcpy(d, s_, bytes); }
inline void stream_sync(cudaStream_t) {}
#endif

// kernel groups: group X is defined by the unit that sets MB_DEFINE_X
#ifdef MB_DEFINE_MSM_G1
#define MB_K_MSM_G1(name, Args, body, BLOCK) MB_KERNEL_DEF(name, Args, body, BLOCK)
#else
#define MB_K_MSM_G1(name, Args, body, BLOCK) MB_KERNEL_DECL(name, Args)
#endif
#ifdef MB_DEFINE_MSM_G2
#define MB_K_MSM_G2(name, Args, body, BLOCK) MB_KERNEL_DEF(name, Args, body, BLOCK)
#else
#define MB_K_MSM_G2(name, Args, body, BLOCK) MB_KERNEL_DECL(name, Args)
#endif
#ifdef MB_DEFINE_NTT
#define MB_K_NTT(name, Args, body, BLOCK) MB_KERNEL_DEF(name, Args, body, BLOCK)
#else
#define MB_K_NTT(name, Args, body, BLOCK) MB_KERNEL_DECL(name, Args)
#endif
#ifdef MB_DEFINE_G1
#define MB_K_G1(name, Args, body, BLOCK) MB_KERNEL_DEF(name, Args, body, BLOCK)
#else
#define MB_K_G1(name, Args, body, BLOCK) MB_KERNEL_DECL(name, Args)
#endif
#ifdef MB_DEFINE_G2
#define MB_K_G2(name, Args, body, BLOCK) MB_KERNEL_DEF(name, Args, body, BLOCK)
#else
#define MB_K_G2(name, Args, body, BLOCK) MB_KERNEL_DECL(name, Args)
#endif
#ifdef MB_DEFINE_MISC
#define MB_K_MISC(name, Args, body, BLOCK) MB_KERNEL_DEF(name, Args, body, BLOCK)
#else
#define MB_K_MISC(name, Args, body, BLOCK) MB_KERNEL_DECL(name, Args)
#endif
#ifdef MB_DEFINE_RED_G1
#define MB_K_RED_G1(name, Args, body, BLOCK) MB_KERNEL_DEF(name, Args, body, BLOCK)
#else
#define MB_K_RED_G1(name, Args, body, BLOCK) MB_KERNEL_DECL(name, Args)
#endif
#ifdef MB_DEFINE_RED_G2
#define MB_K_RED_G2(name, Args, body, BLOCK) MB_KERNEL_DEF(name, Args, body, BLOCK)
#else
#define MB_K_RED_G2(name, Args, body, BLOCK) MB_KERNEL_DECL(name, Args)
#endif

// RAII device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    explicit DevBuf(size_t n) { alloc(n); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n) { release(); p = dev_alloc(n); bytes = n; }
    void ensure(size_t n) { if (n > bytes) alloc(n); }
    void release() { if (p) dev_free(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return (T*)p; }
};

}  // namespace mb
