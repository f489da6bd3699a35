// common.cuh — error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>
#include <string>

#include "../../include/dashing_b200.h"

namespace db200 {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_kernel_launches;

#define DB200_CUDA(expr)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            ::db200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return (e__ == cudaErrorMemoryAllocation) ? DB200_ENOMEM                             \
                 : (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? DB200_ENODEV : DB200_ECUDA; \
        }                                                                                        \
    } while (0)

#define DB200_TRY(expr)                \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != DB200_OK) return rc__; \
    } while (0)

#define DB200_LAUNCHED() (++::db200::g_kernel_launches)

inline int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no usable CUDA device (%s); libdashing_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return DB200_ENODEV;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range [0,%d)", device, n);
        return DB200_EINVAL;
    }
    DB200_CUDA(cudaSetDevice(device));
    return DB200_OK;
}

// RAII device buffer that grows on demand (cudaMalloc/cudaFree are synchronising: keep them out of hot loops).
struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return DB200_OK;
        if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
        size_t want = (bytes + 255) & ~size_t(255);
        DB200_CUDA(cudaMalloc(&ptr, want));
        cap = want;
        return DB200_OK;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(ptr); }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

} // namespace db200
