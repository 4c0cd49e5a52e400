// runtime.cu -- see runtime.h.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <map>
#include <mutex>

#include "../../include/kmg.h"
#include "kmg_common.cuh"
#include "runtime.h"

// ------------------------------------------------------------------------------------------
// error reporting
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void kmg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// KMG_TRACE=1: phase timings of the host entry points on stderr
void kmg_trace(const char* what) {
    static const bool on = getenv("KMG_TRACE") != nullptr;
    if (!on) return;
    static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[kmg] %-48s +%.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}


namespace {

// Size-bucketed cache of device allocations: cudaMalloc / cudaFree of multi-GB buffers cost tens of
// milliseconds per host call (cudaFree also synchronises the device); repeated Gram builds (run.py
// builds nine kernels) reuse the buffers instead.  kmg_release() returns everything to the driver.
struct DevCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> free_list;  // (device, bucket bytes) -> pointer
    size_t cached_bytes = 0;
    static size_t bucket(size_t n) {
        size_t b = 256;
        while (b < n) b <<= 1;
        const size_t step = b >> 3;  // 8 sub-buckets per power of two: <= 12.5 % slack
        return step ? (n + step - 1) / step * step : b;
    }
    void flush() {
        for (auto& kv : free_list) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
        free_list.clear();
        cached_bytes = 0;
    }
};
DevCache g_cache;

}  // namespace

void DevBuf::release() {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    g_cache.free_list.emplace(std::make_pair(dev, bytes), p);
    g_cache.cached_bytes += bytes;
    p = nullptr;
}
int DevBuf::alloc(size_t n) {
    release();
    if (n == 0) return KMG_OK;
    cudaGetDevice(&dev);
    bytes = DevCache::bucket(n);
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        auto it = g_cache.free_list.find(std::make_pair(dev, bytes));
        if (it != g_cache.free_list.end()) {
            p = it->second;
            g_cache.free_list.erase(it);
            g_cache.cached_bytes -= bytes;
            return KMG_OK;
        }
    }
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {  // give the cached buffers back and retry once
        cudaGetLastError();
        { std::lock_guard<std::mutex> lk(g_cache.mu); g_cache.flush(); cudaSetDevice(dev); }
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        p = nullptr;
        kmg_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return KMG_ERR_NOMEM;
    }
    return KMG_OK;
}

namespace {

struct StreamHolder {
    cudaStream_t s[2] = {nullptr, nullptr};
    int dev = -1;
};
thread_local StreamHolder g_streams;

}  // namespace

int kmg_rt_get_streams(cudaStream_t* s0, cudaStream_t* s1) {
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_streams.dev != dev || g_streams.s[0] == nullptr) {
        for (int i = 0; i < 2; ++i) KMG_CUDA_CHECK(cudaStreamCreateWithFlags(&g_streams.s[i], cudaStreamNonBlocking));
        g_streams.dev = dev;
    }
    *s0 = g_streams.s[0];
    if (s1) *s1 = g_streams.s[1];
    return KMG_OK;
}

int kmg_rt_require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        kmg_set_error("no CUDA device available (%s): libkmg has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return KMG_ERR_CUDA;
    }
    return KMG_OK;
}


const char* kmg_rt_last_error() { return g_err; }

size_t kmg_rt_cached_bytes() {
    std::lock_guard<std::mutex> lk(g_cache.mu);
    return g_cache.cached_bytes;
}

void kmg_rt_flush_cache() {
    std::lock_guard<std::mutex> lk(g_cache.mu);
    int dev = 0;
    cudaGetDevice(&dev);
    g_cache.flush();
    cudaSetDevice(dev);
}
