// common.cuh — error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
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

// Logical devices.  Normally logical == physical.  DB200_VIRTUAL_DEVICES=N (a test / debugging knob) exposes N logical
// devices mapped round-robin onto the physical ones, so that the multi-device code paths (device = DB200_ALL_DEVICES)
// can be exercised on a single-GPU box; every logical device has its own streams and buffers.
inline int phys_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
inline int logical_device_count() {
    const int n = phys_device_count();
    if (n == 0) return 0;
    const char *e = std::getenv("DB200_VIRTUAL_DEVICES");
    const int v = e ? std::atoi(e) : 0;
    return v > 0 ? (v > 64 ? 64 : v) : n;
}
inline int phys_of(int logical) {
    const int n = phys_device_count();
    return n > 0 ? logical % n : 0;
}

inline int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no usable CUDA device (%s); libdashing_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return DB200_ENODEV;
    }
    const int nl = logical_device_count();
    if (device < 0 || device >= nl) {
        set_error("device %d out of range [0,%d)", device, nl);
        return DB200_EINVAL;
    }
    DB200_CUDA(cudaSetDevice(device % n));
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
