// runtime.h -- process-level plumbing shared by the host-side translation units of libkmg.so: error text, phase trace,
// cached device buffers, the library's own streams.  No compute.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

const char* kmg_rt_last_error();
// KMG_TRACE=1: phase timings of the host entry points on stderr
void kmg_trace(const char* what);
// KMG_ERR_CUDA (with the "no CPU fallback" message) when no device is usable
int kmg_rt_require_device();
// two non-blocking streams per host thread and device (s1 may be null)
int kmg_rt_get_streams(cudaStream_t* s0, cudaStream_t* s1);
// bytes parked in the device-buffer cache (reclaimable), and giving them back to the driver
size_t kmg_rt_cached_bytes();
void kmg_rt_flush_cache();

// A device allocation from the size-bucketed cache: cudaMalloc / cudaFree of multi-GB buffers cost tens of milliseconds
// per host call (cudaFree also synchronises the device); repeated Gram builds (run.py builds nine kernels) reuse the
// buffers instead.  kmg_release() returns everything to the driver.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;  // bucket size actually allocated
    int dev = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release();
    int alloc(size_t n);
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
