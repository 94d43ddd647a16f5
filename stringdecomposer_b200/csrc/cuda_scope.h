// cuda_scope.h -- makes a device current for the calling thread and puts the caller's device back on scope exit: a
// host application (or torch, in bench.py) that works on another GPU must not find its current device changed by a
// library call.  The previous device is only restored if its primary context is alive -- a thread that never used
// CUDA reports device 0, and "restoring" that would create a context there for nothing.
#pragma once
#include <cuda_runtime.h>

namespace sdb {

inline bool primary_context_active(int dev)
{
    typedef int (*state_fn)(int, unsigned *, int *);            // CUresult cuDevicePrimaryCtxGetState(CUdevice, unsigned*, int*)
    static const state_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<state_fn>(p);
    }();
    if (!fn) return true;
    unsigned flags = 0;
    int active = 0;
    return fn(dev, &flags, &active) == 0 && active != 0;
}

struct DeviceScope {
    int prev = -1, dev;
    cudaError_t status;
    explicit DeviceScope(int d) : dev(d)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev == dev || (prev >= 0 && !primary_context_active(prev))) prev = -1;
        status = cudaSetDevice(dev);
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope &) = delete;
    DeviceScope &operator=(const DeviceScope &) = delete;
};

} // namespace sdb
