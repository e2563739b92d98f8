// common.h — error plumbing shared by the C-ABI translation units of libsnake_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "snake_b200.h"

namespace snk {

char *err_buf();                       // thread-local, 512 bytes
int fail(int code, const char *fmt, ...);

#define SNK_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return snk::fail(SNK_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,       \
                             cudaGetErrorString(e__));                                         \
    } while (0)

#define SNK_REQUIRE(cond, msg)                                                                 \
    do {                                                                                       \
        if (!(cond)) return snk::fail(SNK_ERR_INVALID, "%s: %s", __func__, msg);               \
    } while (0)

typedef unsigned long long u64;

// Every ABI entry runs on its handle's device and leaves the caller's current device as it found it (a process driving
// several GPUs, e.g. from torch, must not have its device changed behind its back).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != device) prev = cur;
        cudaSetDevice(device);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// The stateless entry points (Gram, centring, masked target, ...) take device pointers only: they run on the device that owns
// their first output / workspace pointer, whatever the caller's current device is.
inline int device_of(const void *p) {
    cudaPointerAttributes a;
    if (p != nullptr && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) return a.device;
    cudaGetLastError();
    int cur = 0;
    cudaGetDevice(&cur);
    return cur;
}

}  // namespace snk
