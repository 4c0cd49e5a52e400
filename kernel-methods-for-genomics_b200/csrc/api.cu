// api.cu -- the C-ABI of libkmg.so (include/kmg.h): argument checking, device-memory plumbing and
// the host<->device orchestration behind the `*_host` entry points.  No compute happens here and
// there is no CPU fallback: every path ends in one of the sm_100a kernels of this directory.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <chrono>
#include <future>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/kmg.h"
#include "elementwise.h"
#include "gram_i8.h"
#include "kmg_common.cuh"
#include "pair_kernels.h"
#include "seq_kernels.h"

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
static void kmg_trace(const char* what) {
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

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;  // bucket size actually allocated
    int dev = 0;
    ~DevBuf() { release(); }
    void release() {
        if (!p) return;
        std::lock_guard<std::mutex> lk(g_cache.mu);
        g_cache.free_list.emplace(std::make_pair(dev, bytes), p);
        g_cache.cached_bytes += bytes;
        p = nullptr;
    }
    int alloc(size_t n) {
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
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

struct StreamHolder {
    cudaStream_t s[2] = {nullptr, nullptr};
    int dev = -1;
};
thread_local StreamHolder g_streams;

int get_streams(cudaStream_t* s0, cudaStream_t* s1) {
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

int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        kmg_set_error("no CUDA device available (%s): libkmg has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return KMG_ERR_CUDA;
    }
    return KMG_OK;
}

// Upload + pack one set of sequences.  planes: n x 8 u32.
int upload_planes(const uint8_t* seqs, int64_t n, int L, int fmt, DevBuf* planes, cudaStream_t s) {
    KMG_REQUIRE(L >= 1 && L <= KMG_MAX_L, KMG_ERR_UNSUPPORTED, "sequence length %d not supported (1..%d)", L, KMG_MAX_L);
    int rc = planes->alloc((size_t)std::max<int64_t>(n, 1) * KMG_SEQ_WORDS * sizeof(uint32_t));
    if (rc) return rc;
    if (n == 0) return KMG_OK;
    DevBuf raw, err;
    if ((rc = raw.alloc((size_t)n * L))) return rc;
    if ((rc = err.alloc(sizeof(int)))) return rc;
    KMG_CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), s));
    KMG_CUDA_CHECK(cudaMemcpyAsync(raw.p, seqs, (size_t)n * L, cudaMemcpyHostToDevice, s));
    if ((rc = kmg_pack_launch(raw.as<uint8_t>(), fmt == KMG_SEQ_ASCII, n, L, planes->as<uint32_t>(), err.as<int>(), s))) return rc;
    int herr = 0;
    KMG_CUDA_CHECK(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    KMG_REQUIRE(herr == 0, KMG_ERR_ALPHABET, "sequence contains a character outside {A,C,G,T}");
    return KMG_OK;
}

// Row-block size so that two output buffers of `rows x cols` doubles fit in a fraction of free memory.
int pick_block_rows(int64_t nr, int64_t nc, int64_t* block_rows) {
    size_t free_b = 0, total_b = 0;
    KMG_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    { std::lock_guard<std::mutex> lk(g_cache.mu); free_b += g_cache.cached_bytes; }  // cached buffers are reclaimable
    double budget = 0.70 * (double)free_b;
    if (const char* v = getenv("KMG_DEVICE_BUDGET_BYTES")) budget = atof(v);  // tests force the streamed path at small n
    int64_t r = (int64_t)(budget / (2.0 * 8.0 * (double)std::max<int64_t>(nc, 1)));
    r = std::min<int64_t>(r, 32768);
    r = (r / 256) * 256;
    if (getenv("KMG_DEVICE_BUDGET_BYTES") && r < 256) r = 256;
    KMG_REQUIRE(r >= 256 || r >= nr, KMG_ERR_NOMEM, "not enough device memory for a 256-row block of %lld columns", (long long)nc);
    *block_rows = std::max<int64_t>(std::min<int64_t>(r, nr), 1);
    return KMG_OK;
}

// ------------------------------------------------------------------------------------------
// Device -> pageable host copy through a ring of pinned staging slots.  A plain cudaMemcpy into
// pageable memory is staged by the driver on one thread (3-4 GB/s measured into freshly
// allocated numpy memory); here the DMA into pinned slots runs at PCIe speed while one host
// thread per slot copies (and first-touches) the caller's pages in parallel.
// ------------------------------------------------------------------------------------------
constexpr int D2H_SLOTS = 16;
constexpr size_t D2H_SLOT_BYTES = 8u << 20;
struct PinnedRing {
    void* buf[D2H_SLOTS] = {};
    cudaEvent_t ev[D2H_SLOTS] = {};
    std::future<int> fut[D2H_SLOTS];  // the copy thread that empties each slot
    int slot = 0;                     // next slot to fill
    int* h_flags = nullptr;           // mapped pinned ints the device raises (u16 overflow of a block), host view
    int* d_flags = nullptr;           // ... device view
    bool ready = false;
    std::mutex mu;
};
constexpr int D2H_FLAGS = 64;
PinnedRing g_ring;

int ring_init() {
    if (g_ring.ready) return KMG_OK;
    for (int i = 0; i < D2H_SLOTS; ++i) {
        KMG_CUDA_CHECK(cudaHostAlloc(&g_ring.buf[i], D2H_SLOT_BYTES, cudaHostAllocDefault));
        KMG_CUDA_CHECK(cudaEventCreateWithFlags(&g_ring.ev[i], cudaEventDisableTiming));
    }
    KMG_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&g_ring.h_flags), D2H_FLAGS * sizeof(int), cudaHostAllocMapped));
    KMG_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_ring.d_flags), g_ring.h_flags, 0));
    g_ring.ready = true;
    return KMG_OK;
}

// int32 -> double widening of one staged row into the caller's memory (exact: every s32 is a double).  Streaming
// stores: the destination is written once and not read back here, so skip the read-for-ownership.
void widen_s32_row(double* __restrict__ dst, const int32_t* __restrict__ src, int64_t n) {
    int64_t j = 0;
#if defined(__SSE2__)
    while (j < n && (reinterpret_cast<uintptr_t>(dst + j) & 15)) { dst[j] = (double)src[j]; ++j; }
    for (; j + 4 <= n; j += 4) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + j));
        _mm_stream_pd(dst + j, _mm_cvtepi32_pd(v));
        _mm_stream_pd(dst + j + 2, _mm_cvtepi32_pd(_mm_shuffle_epi32(v, 0xEE)));
    }
#endif
    for (; j < n; ++j) dst[j] = (double)src[j];
}

void widen_u16_row(double* __restrict__ dst, const uint16_t* __restrict__ src, int64_t n) {
    int64_t j = 0;
#if defined(__SSE2__)
    while (j < n && (reinterpret_cast<uintptr_t>(dst + j) & 15)) { dst[j] = (double)src[j]; ++j; }
    const __m128i zero = _mm_setzero_si128();
    for (; j + 8 <= n; j += 8) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + j));
        const __m128i lo = _mm_unpacklo_epi16(v, zero), hi = _mm_unpackhi_epi16(v, zero);
        _mm_stream_pd(dst + j, _mm_cvtepi32_pd(lo));
        _mm_stream_pd(dst + j + 2, _mm_cvtepi32_pd(_mm_shuffle_epi32(lo, 0xEE)));
        _mm_stream_pd(dst + j + 4, _mm_cvtepi32_pd(hi));
        _mm_stream_pd(dst + j + 6, _mm_cvtepi32_pd(_mm_shuffle_epi32(hi, 0xEE)));
    }
#endif
    for (; j < n; ++j) dst[j] = (double)src[j];
}

// src: device, `rows` x `cols` contiguous, doubles or (src_s32) int32 counts that the copy threads widen to double on
// the way into the caller's buffer -- an unnormalised spectrum Gram is integer valued, so shipping the tensor cores'
// own s32 accumulators halves the PCIe bytes per entry.  dst: host doubles, row stride ldk.
// src_elem: 8 = doubles, 4 = s32 counts, 2 = u16 counts (both widened exactly).
// drain = false: returns once every DMA of the block is enqueued on `s` (the source may be overwritten by later work on
// `s`); the copy threads of the last slots may still be writing `dst` -- call ring_drain() before handing `dst` out.
int ring_drain_locked() {
    int err = KMG_OK;
    for (int i = 0; i < D2H_SLOTS; ++i)
        if (g_ring.fut[i].valid() && g_ring.fut[i].get() != 0) err = KMG_ERR_CUDA;
    if (err) kmg_set_error("device-to-host copy failed");
    return err;
}
int ring_drain() {
    std::lock_guard<std::mutex> lk(g_ring.mu);
    return ring_drain_locked();
}

int d2h_rows(double* dst, int64_t ldk, const void* src_v, int src_elem, int64_t cols, int64_t rows, cudaStream_t s, bool drain = true) {
    const bool src_s32 = src_elem != 8;  // "needs widening"
    if (rows <= 0 || cols <= 0) return KMG_OK;
    std::lock_guard<std::mutex> lk(g_ring.mu);
    int rc = ring_init();
    if (rc) return rc;
    const char* src = static_cast<const char*>(src_v);
    const size_t row_bytes = (size_t)cols * (size_t)src_elem;
    {
        // Freshly allocated numpy memory is first touched by the copy threads below; with transparent huge pages the
        // kernel zero-fills 2 MB at a time instead of taking a fault per 4 KB page.  Advisory: errors are ignored.
        const uintptr_t lo = (reinterpret_cast<uintptr_t>(dst) + 0x1FFFFF) & ~uintptr_t(0x1FFFFF);
        const uintptr_t hi = (reinterpret_cast<uintptr_t>(dst + (rows - 1) * ldk + cols)) & ~uintptr_t(0x1FFFFF);
        if (hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
    }
    if (row_bytes > D2H_SLOT_BYTES) {  // absurdly wide rows (callers never ask for s32 here): let the driver stage it
        KMG_REQUIRE(!src_s32, KMG_ERR_UNSUPPORTED, "rows wider than a staging slot");
        KMG_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)ldk * 8, src, row_bytes, row_bytes, (size_t)rows, cudaMemcpyDeviceToHost, s));
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
        return KMG_OK;
    }
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(D2H_SLOT_BYTES / row_bytes));
    std::future<int>* fut = g_ring.fut;
    int err = KMG_OK;
    for (int64_t r = 0; r < rows; r += chunk_rows, g_ring.slot = (g_ring.slot + 1) % D2H_SLOTS) {
        const int slot = g_ring.slot;
        const int64_t nr = std::min<int64_t>(chunk_rows, rows - r);
        if (fut[slot].valid() && fut[slot].get() != 0) err = KMG_ERR_CUDA;
        if (err) break;
        KMG_CUDA_CHECK(cudaMemcpyAsync(g_ring.buf[slot], src + (size_t)r * row_bytes, (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, s));
        KMG_CUDA_CHECK(cudaEventRecord(g_ring.ev[slot], s));
        const char* stage = reinterpret_cast<const char*>(g_ring.buf[slot]);
        cudaEvent_t ev = g_ring.ev[slot];
        double* d0 = dst + r * ldk;
        fut[slot] = std::async(std::launch::async, [=]() -> int {
            if (cudaEventSynchronize(ev) != cudaSuccess) return 1;
            if (src_s32) {
                for (int64_t i = 0; i < nr; ++i) {
                    if (src_elem == 4) widen_s32_row(d0 + i * ldk, reinterpret_cast<const int32_t*>(stage + (size_t)i * row_bytes), cols);
                    else widen_u16_row(d0 + i * ldk, reinterpret_cast<const uint16_t*>(stage + (size_t)i * row_bytes), cols);
                }
#if defined(__SSE2__)
                _mm_sfence();
#endif
            } else if (ldk == cols) {
                memcpy(d0, stage, (size_t)nr * row_bytes);
            } else {
                for (int64_t i = 0; i < nr; ++i) memcpy(d0 + i * ldk, stage + (size_t)i * row_bytes, row_bytes);
            }
            return 0;
        });
    }
    if (err || drain) {
        const int e2 = ring_drain_locked();
        if (err || e2) { kmg_set_error("device-to-host copy failed"); return err ? err : e2; }
    }
    return KMG_OK;
}

// ------------------------------------------------------------------------------------------
// Recycled host memory for results.  A fresh numpy array of a few GB is mmap'ed untouched, so every byte the copy
// threads write first takes a page fault + kernel zero-fill (measured: ~30 GB/s aggregate over 16 threads, below the
// PCIe rate), and free() munmaps it again.  Blocks handed out here are 2 MB aligned, huge-page advised, and go back to
// a bounded cache on release instead of to the kernel, so the second and later results of a job are written into
// memory that is already mapped.  Pageable memory: nothing is pinned, the cold cost equals plain malloc's.
// ------------------------------------------------------------------------------------------
struct HostPool {
    std::mutex mu;
    std::map<void*, size_t> live;
    std::multimap<size_t, void*> cached;
    size_t cached_bytes = 0;
    size_t cap() const {
        if (const char* v = getenv("KMG_HOST_POOL_BYTES")) return (size_t)atof(v);
        return (size_t)8 << 30;
    }
    void trim(size_t keep) {
        while (cached_bytes > keep && !cached.empty()) {
            auto it = std::prev(cached.end());
            munmap(it->second, it->first);
            cached_bytes -= it->first;
            cached.erase(it);
        }
    }
};
HostPool g_hostpool;

typedef int (*BlockFn)(void* ctx, int64_t r0, int64_t rows, void* d_out, int64_t ldo, int symmetric, cudaStream_t s);

// Build an nr x nc Gram block-row by block-row on the device and copy it to host memory.
// If the whole (square, symmetric) matrix fits it is built in one symmetric launch.
// Second narrowing of a block of s32 counts for the link: if every entry fits 16 bits (checked on the device, exact)
// the block crosses PCIe as u16 -- 2 bytes per Gram entry instead of 8.  Costs one HBM pass (6 B/entry) and a flag read.
int narrow_enabled() { return getenv("KMG_D2H_S32") == nullptr; }

// overflow flags: D2H_FLAGS mapped pinned ints shared by all calls of the process; a call owns the slots it acquired
// (host entry points may run concurrently from several threads: ctypes drops the GIL)
uint64_t g_flag_busy = 0;
int flag_acquire() {
    std::lock_guard<std::mutex> lk(g_ring.mu);
    if (ring_init() != KMG_OK) return -1;
    for (int f = 0; f < D2H_FLAGS; ++f)
        if (!((g_flag_busy >> f) & 1)) { g_flag_busy |= (uint64_t)1 << f; g_ring.h_flags[f] = 0; return f; }
    return -1;  // none free: the caller ships s32
}
void flag_release(int f) {
    if (f < 0) return;
    std::lock_guard<std::mutex> lk(g_ring.mu);
    g_flag_busy &= ~((uint64_t)1 << f);
}

// enqueue the check-and-pack pass of one block on `s`; flag slot `f` is raised on overflow
int narrow_launch(const void* d_s32, int64_t count, void* d_u16, int f, cudaStream_t s) {
    return kmg_ew_narrow_u16(static_cast<const int32_t*>(d_s32), count, static_cast<uint16_t*>(d_u16), g_ring.d_flags + f, s);
}

int try_narrow(const void* d_s32, int64_t count, DevBuf* narrow, int* elem, cudaStream_t s) {
    if (!narrow_enabled()) return KMG_OK;
    int rc;
    if ((rc = narrow->alloc((size_t)count * 2 + 16))) return rc;
    const int f = flag_acquire();
    if (f < 0) return KMG_OK;
    rc = narrow_launch(d_s32, count, narrow->p, f, s);
    if (rc == KMG_OK && cudaStreamSynchronize(s) != cudaSuccess) { kmg_set_error("narrow pass failed"); rc = KMG_ERR_CUDA; }
    if (rc == KMG_OK && g_ring.h_flags[f] == 0) *elem = 2;
    flag_release(f);
    return rc;
}

// out_s32: `fn` writes int32 counts (d2h_rows widens them on the host side of the link).
int build_to_host(int64_t nr, int64_t nc, bool symmetric, BlockFn fn, void* ctx, double* K, int64_t ldk, bool out_s32 = false) {
    if (nr == 0 || nc == 0) return KMG_OK;
    const size_t esz = out_s32 ? sizeof(int32_t) : sizeof(double);
    cudaStream_t s0, s1;
    int rc = get_streams(&s0, &s1);
    if (rc) return rc;
    int64_t br = 0;
    if ((rc = pick_block_rows(nr, nc, &br))) return rc;
    if (br >= nr) {
        // The whole block fits.  A large cross-Gram is still built in a few row chunks, all enqueued up front: the host
        // link (the slow side) starts on chunk 0 while the GPU builds the rest, and the copy threads run across chunk
        // boundaries (d2h_rows does not drain between chunks).
        const int nchunks = (!symmetric && nr >= 1024 && (double)nr * (double)nc >= 64e6 && !getenv("KMG_NO_SPLIT")) ? 4 : 1;
        const int64_t crow = nchunks == 1 ? nr : ((nr + nchunks - 1) / nchunks + 255) / 256 * 256;
        const bool narrowing = out_s32 && narrow_enabled();
        DevBuf out, narrow;
        if ((rc = out.alloc((size_t)nr * nc * esz))) return rc;
        if (narrowing && (rc = narrow.alloc((size_t)nr * nc * 2 + 16))) return rc;
        kmg_trace("build_to_host: output allocated");
        cudaEvent_t ev[4] = {};
        int flag[4] = {-1, -1, -1, -1};
        int used = 0;
        for (int64_t r0 = 0; r0 < nr; r0 += crow, ++used) {
            const int64_t rows = std::min<int64_t>(crow, nr - r0);
            char* o = static_cast<char*>(out.p) + (size_t)r0 * nc * esz;
            if ((rc = fn(ctx, r0, rows, o, nc, symmetric ? 1 : 0, s0))) break;
            if (narrowing && (flag[used] = flag_acquire()) >= 0 &&
                (rc = narrow_launch(o, rows * nc, static_cast<char*>(narrow.p) + (size_t)r0 * nc * 2, flag[used], s0))) break;
            if (cudaEventCreateWithFlags(&ev[used], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(ev[used], s0) != cudaSuccess) {
                kmg_set_error("build_to_host: event failed"); rc = KMG_ERR_CUDA; ++used; break;
            }
        }
        int c = 0;
        for (int64_t r0 = 0; rc == KMG_OK && r0 < nr; r0 += crow, ++c) {
            const int64_t rows = std::min<int64_t>(crow, nr - r0);
            if (cudaEventSynchronize(ev[c]) != cudaSuccess) { kmg_set_error("build_to_host: kernel failed: %s", cudaGetErrorString(cudaGetLastError())); rc = KMG_ERR_CUDA; break; }
            if (c == 0) kmg_trace("build_to_host: first chunk done");
            const int elem = out_s32 ? ((flag[c] >= 0 && g_ring.h_flags[flag[c]] == 0) ? 2 : 4) : 8;
            const char* src = elem == 2 ? static_cast<char*>(narrow.p) + (size_t)r0 * nc * 2 : static_cast<char*>(out.p) + (size_t)r0 * nc * esz;
            rc = d2h_rows(K + r0 * ldk, ldk, src, elem, nc, rows, s1, /*drain=*/false);
        }
        const int rc2 = ring_drain();
        if (rc) cudaStreamSynchronize(s0);  // nothing of this call may still run when its buffers and flags are released
        for (int i = 0; i < 4; ++i) {
            if (ev[i]) cudaEventDestroy(ev[i]);
            flag_release(flag[i]);
        }
        kmg_trace("build_to_host: copied to host");
        return rc ? rc : rc2;
    }
    // streamed: two device buffers; the GPU builds block b+1 while block b drains to the host
    DevBuf buf[2];
    cudaStream_t st[2] = {s0, s1};
    for (int i = 0; i < 2; ++i)
        if ((rc = buf[i].alloc((size_t)br * nc * esz))) return rc;
    const int64_t nblocks = (nr + br - 1) / br;
    if ((rc = fn(ctx, 0, std::min<int64_t>(br, nr), buf[0].p, nc, 0, st[0]))) return rc;
    for (int64_t b = 0; b < nblocks; ++b) {
        const int64_t r0 = b * br, rows = std::min<int64_t>(br, nr - r0);
        if (b + 1 < nblocks) {
            const int64_t r1 = (b + 1) * br;
            if ((rc = fn(ctx, r1, std::min<int64_t>(br, nr - r1), buf[(b + 1) & 1].p, nc, 0, st[(b + 1) & 1]))) return rc;
        }
        int elem = out_s32 ? 4 : 8;
        DevBuf narrow;
        if (out_s32 && (rc = try_narrow(buf[b & 1].p, rows * nc, &narrow, &elem, st[b & 1]))) return rc;
        if ((rc = d2h_rows(K + r0 * ldk, ldk, elem == 2 ? narrow.p : buf[b & 1].p, elem, nc, rows, st[b & 1]))) return rc;
    }
    return KMG_OK;
}

struct SeqPair {
    DevBuf prow, pcol;
    const uint32_t* rows = nullptr;
    const uint32_t* cols = nullptr;
    int64_t nr = 0, nc = 0;
    bool symmetric = false;
};

int upload_pair(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int fmt, SeqPair* sp) {
    KMG_REQUIRE(nr >= 0 && (cols == nullptr || nc >= 0), KMG_ERR_ARG, "negative sequence count");
    KMG_REQUIRE(rows != nullptr || nr == 0, KMG_ERR_ARG, "null sequence pointer");
    cudaStream_t s;
    int rc = get_streams(&s, nullptr);
    if (rc) return rc;
    if ((rc = upload_planes(rows, nr, L, fmt, &sp->prow, s))) return rc;
    sp->rows = sp->prow.as<uint32_t>();
    sp->nr = nr;
    if (cols == nullptr) {
        sp->cols = sp->rows; sp->nc = nr; sp->symmetric = true;
    } else {
        if ((rc = upload_planes(cols, nc, L, fmt, &sp->pcol, s))) return rc;
        sp->cols = sp->pcol.as<uint32_t>(); sp->nc = nc; sp->symmetric = false;
    }
    return KMG_OK;
}

// ---- per-kernel block builders -----------------------------------------------------------
struct SpectrumCtx {
    const int8_t* phi_rows; const int8_t* phi_cols; int64_t nc; int64_t width;
    const double* sd_rows; const double* sd_cols;
    int out_dtype;
};
int spectrum_block(void* c, int64_t r0, int64_t rows, void* out, int64_t ldo, int symmetric, cudaStream_t s) {
    SpectrumCtx* x = (SpectrumCtx*)c;
    GramI8Args a;
    memset(&a, 0, sizeof(a));
    a.phi_rows = x->phi_rows + r0 * x->width; a.phi_cols = x->phi_cols;
    a.rows = rows; a.cols = x->nc; a.Dpad = x->width; a.ld_phi = x->width;
    a.row_index0 = symmetric ? 0 : r0; a.col_index0 = 0;
    a.out = out; a.ldo = ldo; a.out_dtype = x->out_dtype; a.symmetric = symmetric; a.out_t = out; a.ldo_t = ldo;
    a.sd_rows = x->sd_rows ? x->sd_rows + r0 : nullptr; a.sd_cols = x->sd_cols;
    return kmg_gram_i8_launch(&a, s);
}

struct PairCtx {
    const SeqPair* sp; int L; int kind; int k, m, d, smith; double e, dd, beta; const double* sd; bool index_diag;
};
int pair_block(void* c, int64_t r0, int64_t rows, void* out, int64_t ldo, int symmetric, cudaStream_t s) {
    PairCtx* x = (PairCtx*)c;
    PairBlock b;
    memset(&b, 0, sizeof(b));
    b.planes_rows = x->sp->rows + r0 * KMG_SEQ_WORDS; b.planes_cols = x->sp->cols;
    b.rows = rows; b.cols = x->sp->nc;
    // cross-Grams have no diagonal: offset the column indices so that row index never equals column index
    b.row_index0 = r0; b.col_index0 = x->index_diag ? 0 : (int64_t)1 << 40;
    b.L = x->L; b.out = out; b.ldo = ldo; b.out_dtype = KMG_OUT_F64;
    b.symmetric = symmetric; b.out_t = out; b.ldo_t = ldo;
    b.sd_rows = x->sd ? x->sd + r0 : nullptr; b.sd_cols = x->sd;
    if (x->kind == 0) return kmg_mismatch_launch(&b, x->k, x->m, s);
    if (x->kind == 1) return kmg_wd_launch(&b, x->d, s);
    if (x->kind == 3) return kmg_wds_launch(&b, x->d, x->k /* S */, s);
    return kmg_la_launch(&b, x->e, x->dd, x->beta, x->smith, s);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// library
// ------------------------------------------------------------------------------------------
extern "C" {

int kmg_version(void) { return 100; }
const char* kmg_last_error(void) { return g_err; }

int kmg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int kmg_set_device(int device) {
    KMG_CUDA_CHECK(cudaSetDevice(device));
    return KMG_OK;
}

// ---- plain device buffers for callers that keep Grams resident between calls (kmg/resident.py) ----
int kmg_dev_malloc(int64_t bytes, void** ptr) {
    int rc = require_device();
    if (rc) return rc;
    KMG_REQUIRE(bytes >= 0 && ptr != nullptr, KMG_ERR_ARG, "dev_malloc: bad arguments");
    *ptr = nullptr;
    if (bytes == 0) return KMG_OK;
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        { std::lock_guard<std::mutex> lk(g_cache.mu); g_cache.flush(); }
        e = cudaMalloc(ptr, (size_t)bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        kmg_set_error("cudaMalloc(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return KMG_ERR_NOMEM;
    }
    return KMG_OK;
}

int kmg_dev_free(void* ptr) {
    if (ptr) KMG_CUDA_CHECK(cudaFree(ptr));
    return KMG_OK;
}

int kmg_dev_upload(void* d_dst, const void* h_src, int64_t bytes) {
    if (bytes > 0) KMG_CUDA_CHECK(cudaMemcpy(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice));
    return KMG_OK;
}

int kmg_dev_download(void* h_dst, const void* d_src, int64_t bytes) {
    if (bytes > 0) KMG_CUDA_CHECK(cudaMemcpy(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return KMG_OK;
}

// ---- recycled host memory for result arrays (kmg/host.py wraps a block as the numpy array it returns) ----
int kmg_host_alloc(int64_t bytes, void** ptr) {
    KMG_REQUIRE(bytes >= 0 && ptr != nullptr, KMG_ERR_ARG, "host_alloc: bad arguments");
    *ptr = nullptr;
    if (bytes == 0) return KMG_OK;
    const size_t need = ((size_t)bytes + 0x1FFFFF) & ~(size_t)0x1FFFFF;
    std::lock_guard<std::mutex> lk(g_hostpool.mu);
    auto it = g_hostpool.cached.lower_bound(need);
    if (it != g_hostpool.cached.end() && it->first <= need + need / 4) {
        *ptr = it->second;
        g_hostpool.live[it->second] = it->first;
        g_hostpool.cached_bytes -= it->first;
        g_hostpool.cached.erase(it);
        return KMG_OK;
    }
    // over-map by 2 MB and trim so that the block is huge-page aligned
    const size_t span = need + 0x200000;
    void* raw = mmap(nullptr, span, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (raw == MAP_FAILED) {
        g_hostpool.trim(0);
        raw = mmap(nullptr, span, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    }
    KMG_REQUIRE(raw != MAP_FAILED, KMG_ERR_NOMEM, "host_alloc: mmap of %lld bytes failed", (long long)bytes);
    const uintptr_t a = (reinterpret_cast<uintptr_t>(raw) + 0x1FFFFF) & ~uintptr_t(0x1FFFFF);
    if (a > reinterpret_cast<uintptr_t>(raw)) munmap(raw, a - reinterpret_cast<uintptr_t>(raw));
    const uintptr_t end = reinterpret_cast<uintptr_t>(raw) + span;
    if (end > a + need) munmap(reinterpret_cast<void*>(a + need), end - (a + need));
    madvise(reinterpret_cast<void*>(a), need, MADV_HUGEPAGE);  // advisory
    *ptr = reinterpret_cast<void*>(a);
    g_hostpool.live[*ptr] = need;
    return KMG_OK;
}

int kmg_host_free(void* ptr) {
    if (!ptr) return KMG_OK;
    std::lock_guard<std::mutex> lk(g_hostpool.mu);
    auto it = g_hostpool.live.find(ptr);
    KMG_REQUIRE(it != g_hostpool.live.end(), KMG_ERR_ARG, "host_free: pointer was not returned by kmg_host_alloc");
    const size_t sz = it->second;
    g_hostpool.live.erase(it);
    const size_t cap = g_hostpool.cap();
    if (sz > cap) { munmap(ptr, sz); return KMG_OK; }
    g_hostpool.cached.emplace(sz, ptr);
    g_hostpool.cached_bytes += sz;
    if (g_hostpool.cached_bytes > cap) {  // evict the other blocks, largest first, keeping the one just returned
        for (auto c = g_hostpool.cached.end(); g_hostpool.cached_bytes > cap && c != g_hostpool.cached.begin();) {
            --c;
            if (c->second == ptr) continue;
            munmap(c->second, c->first);
            g_hostpool.cached_bytes -= c->first;
            c = g_hostpool.cached.erase(c);
        }
    }
    return KMG_OK;
}

int kmg_release(void) {
    { std::lock_guard<std::mutex> lk(g_hostpool.mu); g_hostpool.trim(0); }
    std::lock_guard<std::mutex> lk(g_cache.mu);
    int dev = 0;
    cudaGetDevice(&dev);
    g_cache.flush();
    cudaSetDevice(dev);
    return KMG_OK;
}

int kmg_mismatch_table_host(int k, int m, int64_t* T) {
    KMG_REQUIRE(k >= 1 && k <= KMG_MAX_L && m >= 0 && T != nullptr, KMG_ERR_ARG, "mismatch_table: bad arguments");
    return kmg_mismatch_table(k, m, T);
}

// ------------------------------------------------------------------------------------------
// host-buffer entry points
// ------------------------------------------------------------------------------------------
int kmg_spectrum_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                      const int* ks, int nk, double* K, int64_t ldk) {
    int rc = require_device();
    if (rc) return rc;
    KMG_REQUIRE(ks != nullptr && nk >= 1 && nk <= KMG_MAX_KS, KMG_ERR_ARG, "spectrum: between 1 and %d values of k", KMG_MAX_KS);
    bool pairwise = false;
    for (int q = 0; q < nk; ++q) {
        KMG_REQUIRE(ks[q] >= 1, KMG_ERR_ARG, "spectrum: k must be >= 1");
        if (ks[q] > KMG_MAX_DENSE_K || ks[q] > L) pairwise = true;
    }
    if (pairwise) {
        // k too large for a dense 4^k feature row: SP(k) == raw MM(k, 0) (kernels.py:161-175 with m=0)
        KMG_REQUIRE(nk == 1, KMG_ERR_UNSUPPORTED, "spectrum: sums over several k need every k <= %d", KMG_MAX_DENSE_K);
        if (ks[0] > L) {  // no window at all: the reference returns zeros
            const int64_t c = cols ? nc : nr;
            for (int64_t i = 0; i < nr; ++i) memset(K + i * ldk, 0, (size_t)c * sizeof(double));
            return KMG_OK;
        }
        return kmg_mismatch_host(rows, nr, cols, nc, L, seq_format, ks[0], 0, 0, KMG_MM_PAIRWISE, K, ldk);
    }
    SeqPair sp;
    kmg_trace("spectrum_host: enter");
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    kmg_trace("spectrum_host: sequences uploaded and packed");
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "spectrum: bad output buffer");
    if (sp.nr == 0 || sp.nc == 0) return KMG_OK;
    cudaStream_t s;
    if ((rc = get_streams(&s, nullptr))) return rc;
    const int64_t width = kmg_spectrum_padded_width(ks, nk, L);
    DevBuf phi_r, phi_c;
    if ((rc = phi_r.alloc((size_t)sp.nr * width))) return rc;
    if ((rc = kmg_spectrum_phi_launch(sp.rows, sp.nr, L, ks, nk, phi_r.as<int8_t>(), width, s))) return rc;
    const int8_t* pc = phi_r.as<int8_t>();
    if (!sp.symmetric) {
        if ((rc = phi_c.alloc((size_t)sp.nc * width))) return rc;
        if ((rc = kmg_spectrum_phi_launch(sp.cols, sp.nc, L, ks, nk, phi_c.as<int8_t>(), width, s))) return rc;
        pc = phi_c.as<int8_t>();
    }
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    kmg_trace("spectrum_host: Phi built");
    // unnormalised counts: ship the s32 accumulators, widen to double on the host side of PCIe
    const bool s32 = (size_t)sp.nc * 4 <= D2H_SLOT_BYTES && !getenv("KMG_D2H_F64");
    SpectrumCtx ctx{phi_r.as<int8_t>(), pc, sp.nc, width, nullptr, nullptr, s32 ? KMG_OUT_S32 : KMG_OUT_F64};
    return build_to_host(sp.nr, sp.nc, sp.symmetric, spectrum_block, &ctx, K, ldk, s32);
}

int kmg_mismatch_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                      int k, int m, int normalize, int algo, double* K, int64_t ldk) {
    int rc = require_device();
    if (rc) return rc;
    KMG_REQUIRE(!(normalize && cols != nullptr), KMG_ERR_ARG, "mismatch: normalisation is defined for the symmetric Gram only");
    KMG_REQUIRE(algo >= KMG_MM_AUTO && algo <= KMG_MM_DENSE, KMG_ERR_ARG, "mismatch: algo must be 0 (auto), 1 (pairwise) or 2 (dense)");
    KMG_REQUIRE(k >= 1 && k <= L, KMG_ERR_ARG, "mismatch: need 1 <= k <= L (k=%d, L=%d)", k, L);
    SeqPair sp;
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "mismatch: bad output buffer");
    if (sp.nr == 0 || sp.nc == 0) return KMG_OK;
    cudaStream_t s;
    if ((rc = get_streams(&s, nullptr))) return rc;
    if (algo == KMG_MM_AUTO) algo = (k <= KMG_MAX_DENSE_K && m <= 3 && L - k + 1 <= 127) ? KMG_MM_DENSE : KMG_MM_PAIRWISE;
    if (algo == KMG_MM_DENSE) {
        // the reference's own structure (kernels.py:206-215): dense phi_km, then Phi Phi^T -- on the tensor cores
        const int64_t width = ((1ll << (2 * k)) + 127) / 128 * 128;
        DevBuf phi_r, phi_c, sd;
        if ((rc = phi_r.alloc((size_t)sp.nr * width))) return rc;
        if ((rc = kmg_mismatch_phi_launch(sp.rows, sp.nr, L, k, m, phi_r.as<int8_t>(), width, s))) return rc;
        const int8_t* pc = phi_r.as<int8_t>();
        if (!sp.symmetric) {
            if ((rc = phi_c.alloc((size_t)sp.nc * width))) return rc;
            if ((rc = kmg_mismatch_phi_launch(sp.cols, sp.nc, L, k, m, phi_c.as<int8_t>(), width, s))) return rc;
            pc = phi_c.as<int8_t>();
        }
        const double* sdp = nullptr;
        if (normalize) {
            if ((rc = sd.alloc((size_t)sp.nr * sizeof(double)))) return rc;
            if ((rc = kmg_phi_diag_sqrt_launch(phi_r.as<int8_t>(), sp.nr, width, width, sd.as<double>(), s))) return rc;
            double sd0 = 0.0;
            KMG_CUDA_CHECK(cudaMemcpyAsync(&sd0, sd.p, sizeof(double), cudaMemcpyDeviceToHost, s));
            KMG_CUDA_CHECK(cudaStreamSynchronize(s));
            if (sd0 != 1.0) sdp = sd.as<double>();  // normalize_K early-out, kernels.py:404
        }
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
        SpectrumCtx ctx{phi_r.as<int8_t>(), pc, sp.nc, width, sdp, sdp, KMG_OUT_F64};
        return build_to_host(sp.nr, sp.nc, sp.symmetric, spectrum_block, &ctx, K, ldk);
    }
    DevBuf sd;
    const double* sdp = nullptr;
    if (normalize) {
        if ((rc = sd.alloc((size_t)sp.nr * sizeof(double)))) return rc;
        if ((rc = kmg_mismatch_diag_launch(sp.rows, sp.nr, L, k, m, sd.as<double>(), s))) return rc;
        // normalize_K early-out (kernels.py:404): a raw K[0,0] of exactly 1 leaves the matrix unnormalised
        double sd0 = 0.0;
        KMG_CUDA_CHECK(cudaMemcpyAsync(&sd0, sd.p, sizeof(double), cudaMemcpyDeviceToHost, s));
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
        if (sd0 != 1.0) sdp = sd.as<double>();
    }
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    PairCtx ctx{&sp, L, 0, k, m, 0, 0, 0.0, 0.0, 0.0, sdp, sp.symmetric};
    return build_to_host(sp.nr, sp.nc, sp.symmetric, pair_block, &ctx, K, ldk);
}

static int phi_host(const uint8_t* seqs, int64_t n, int L, int seq_format, int which, const int* ks, int nk, int k, int m,
                    int8_t* phi, int64_t ld) {
    int rc = require_device();
    if (rc) return rc;
    KMG_REQUIRE(n >= 0 && (seqs != nullptr || n == 0) && (phi != nullptr || n == 0), KMG_ERR_ARG, "phi: bad arguments");
    cudaStream_t s;
    if ((rc = get_streams(&s, nullptr))) return rc;
    DevBuf planes, dphi;
    if ((rc = upload_planes(seqs, n, L, seq_format, &planes, s))) return rc;
    const int64_t width = which == 0 ? kmg_spectrum_padded_width(ks, nk, L) : ((1ll << (2 * k)) + 127) / 128 * 128;
    KMG_REQUIRE(ld >= width, KMG_ERR_ARG, "phi: ld must be >= %lld", (long long)width);
    if (n == 0) return KMG_OK;
    if ((rc = dphi.alloc((size_t)n * width))) return rc;
    rc = which == 0 ? kmg_spectrum_phi_launch(planes.as<uint32_t>(), n, L, ks, nk, dphi.as<int8_t>(), width, s)
                    : kmg_mismatch_phi_launch(planes.as<uint32_t>(), n, L, k, m, dphi.as<int8_t>(), width, s);
    if (rc) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(phi, (size_t)ld, dphi.p, (size_t)width, (size_t)width, (size_t)n, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    return KMG_OK;
}

int kmg_spectrum_phi_host(const uint8_t* seqs, int64_t n, int L, int seq_format, const int* ks, int nk, int8_t* phi, int64_t ld) {
    KMG_REQUIRE(ks != nullptr && nk >= 1, KMG_ERR_ARG, "spectrum_phi: need at least one k");
    return phi_host(seqs, n, L, seq_format, 0, ks, nk, 0, 0, phi, ld);
}

int kmg_mismatch_phi_host(const uint8_t* seqs, int64_t n, int L, int seq_format, int k, int m, int8_t* phi, int64_t ld) {
    return phi_host(seqs, n, L, seq_format, 1, nullptr, 0, k, m, phi, ld);
}

int kmg_wd_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                int d, double* K, int64_t ldk) {
    int rc = require_device();
    if (rc) return rc;
    SeqPair sp;
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "wd: bad output buffer");
    PairCtx ctx{&sp, L, 1, 0, 0, d, 0, 0.0, 0.0, 0.0, nullptr, sp.symmetric};
    return build_to_host(sp.nr, sp.nc, sp.symmetric, pair_block, &ctx, K, ldk);
}

int kmg_wds_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                 int d, int S, double* K, int64_t ldk) {
    int rc = require_device();
    if (rc) return rc;
    SeqPair sp;
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "wds: bad output buffer");
    PairCtx ctx{&sp, L, 3, /*k := S*/ S, 0, d, 0, 0.0, 0.0, 0.0, nullptr, sp.symmetric};
    return build_to_host(sp.nr, sp.nc, sp.symmetric, pair_block, &ctx, K, ldk);
}

int kmg_la_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                double e, double d, double beta, int smith, double* K, int64_t ldk) {
    int rc = require_device();
    if (rc) return rc;
    SeqPair sp;
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "la: bad output buffer");
    PairCtx ctx{&sp, L, 2, 0, 0, 0, smith, e, d, beta, nullptr, sp.symmetric};
    return build_to_host(sp.nr, sp.nc, sp.symmetric, pair_block, &ctx, K, ldk);
}

int kmg_normalize_host(double* K, int64_t n, int64_t ldk) {
    int rc = require_device();
    if (rc) return rc;
    KMG_REQUIRE(n >= 0 && (K != nullptr || n == 0) && ldk >= n, KMG_ERR_ARG, "normalize: bad arguments");
    if (n == 0) return 0;
    if (K[0] == 1.0) return 1;  // kernels.py:404-405
    cudaStream_t s;
    if ((rc = get_streams(&s, nullptr))) return rc;
    DevBuf d, sd;
    if ((rc = d.alloc((size_t)n * n * 8))) return rc;
    if ((rc = sd.alloc((size_t)n * 8))) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(d.p, (size_t)n * 8, K, (size_t)ldk * 8, (size_t)n * 8, (size_t)n, cudaMemcpyHostToDevice, s));
    if ((rc = kmg_ew_diag_sqrt(d.as<double>(), n, n, sd.as<double>(), s))) return rc;
    if ((rc = kmg_ew_normalize(d.as<double>(), n, n, sd.as<double>(), s))) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(K, (size_t)ldk * 8, d.p, (size_t)n * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
}

int kmg_center_host(const double* K, int64_t n, int64_t ldk, double* out, int64_t ldo) {
    int rc = require_device();
    if (rc) return rc;
    KMG_REQUIRE(n >= 0 && ldk >= n && ldo >= n, KMG_ERR_ARG, "center: bad arguments");
    if (n == 0) return KMG_OK;
    cudaStream_t s;
    if ((rc = get_streams(&s, nullptr))) return rc;
    DevBuf d, o, ws;
    if ((rc = d.alloc((size_t)n * n * 8))) return rc;
    if ((rc = o.alloc((size_t)n * n * 8))) return rc;
    if ((rc = ws.alloc((size_t)kmg_ew_center_workspace(n)))) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(d.p, (size_t)n * 8, K, (size_t)ldk * 8, (size_t)n * 8, (size_t)n, cudaMemcpyHostToDevice, s));
    if ((rc = kmg_ew_center(d.as<double>(), n, n, o.as<double>(), n, ws.p, s))) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)ldo * 8, o.p, (size_t)n * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    return KMG_OK;
}

int kmg_combine_host(const double* const* Ks, int p, int64_t n, const double* u, int degree, int normalize, double* out) {
    int rc = require_device();
    if (rc) return rc;
    KMG_REQUIRE(p >= 1 && p <= KMG_MAX_COMBINE && n >= 0 && Ks && u && out, KMG_ERR_ARG, "combine: bad arguments");
    if (n == 0) return KMG_OK;
    cudaStream_t s;
    if ((rc = get_streams(&s, nullptr))) return rc;
    std::vector<DevBuf> bufs(p);
    const double* dptr[KMG_MAX_COMBINE];
    int64_t lds[KMG_MAX_COMBINE];
    for (int m = 0; m < p; ++m) {
        if ((rc = bufs[m].alloc((size_t)n * n * 8))) return rc;
        KMG_CUDA_CHECK(cudaMemcpyAsync(bufs[m].p, Ks[m], (size_t)n * n * 8, cudaMemcpyHostToDevice, s));
        dptr[m] = bufs[m].as<double>();
        lds[m] = n;
    }
    DevBuf o, sd;
    if ((rc = o.alloc((size_t)n * n * 8))) return rc;
    if ((rc = kmg_ew_combine(dptr, lds, u, p, degree, n, n, o.as<double>(), n, s))) return rc;
    if (normalize) {
        double k00 = 0.0;
        KMG_CUDA_CHECK(cudaMemcpyAsync(&k00, o.p, 8, cudaMemcpyDeviceToHost, s));
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
        if (k00 != 1.0) {
            if ((rc = sd.alloc((size_t)n * 8))) return rc;
            if ((rc = kmg_ew_diag_sqrt(o.as<double>(), n, n, sd.as<double>(), s))) return rc;
            if ((rc = kmg_ew_normalize(o.as<double>(), n, n, sd.as<double>(), s))) return rc;
        }
    }
    KMG_CUDA_CHECK(cudaMemcpyAsync(out, o.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    return KMG_OK;
}

int kmg_alignf_stats_host(const double* const* Ks, int p, int64_t n, const int64_t* idx, int64_t nfit, const double* y,
                          double* a, double* M) {
    int rc = require_device();
    if (rc) return rc;
    KMG_REQUIRE(p >= 1 && p <= KMG_MAX_COMBINE && n >= 0 && nfit >= 0 && Ks && idx && y && a && M, KMG_ERR_ARG, "alignf_stats: bad arguments");
    for (int64_t t = 0; t < nfit; ++t) KMG_REQUIRE(idx[t] >= 0 && idx[t] < n, KMG_ERR_ARG, "alignf_stats: index out of range");
    if (nfit == 0) { for (int i = 0; i < p; ++i) { a[i] = 0; for (int j = 0; j < p; ++j) M[i * p + j] = 0; } return KMG_OK; }
    cudaStream_t s;
    if ((rc = get_streams(&s, nullptr))) return rc;
    DevBuf full, sub, didx, dy, ws, part, res;
    std::vector<DevBuf> kc(p);
    if ((rc = full.alloc((size_t)n * n * 8))) return rc;
    if ((rc = sub.alloc((size_t)nfit * nfit * 8))) return rc;
    if ((rc = didx.alloc((size_t)nfit * 8))) return rc;
    if ((rc = dy.alloc((size_t)nfit * 8))) return rc;
    if ((rc = ws.alloc((size_t)kmg_ew_center_workspace(nfit)))) return rc;
    if ((rc = part.alloc((size_t)nfit * 8))) return rc;
    if ((rc = res.alloc((size_t)(p + p * p) * 8))) return rc;
    KMG_CUDA_CHECK(cudaMemcpyAsync(didx.p, idx, (size_t)nfit * 8, cudaMemcpyHostToDevice, s));
    KMG_CUDA_CHECK(cudaMemcpyAsync(dy.p, y, (size_t)nfit * 8, cudaMemcpyHostToDevice, s));
    for (int i = 0; i < p; ++i) {
        if ((rc = kc[i].alloc((size_t)nfit * nfit * 8))) return rc;
        KMG_CUDA_CHECK(cudaMemcpyAsync(full.p, Ks[i], (size_t)n * n * 8, cudaMemcpyHostToDevice, s));
        if ((rc = kmg_ew_gather(full.as<double>(), n, didx.as<int64_t>(), nfit, sub.as<double>(), nfit, s))) return rc;   // ALIGNF.py:28
        if ((rc = kmg_ew_center(sub.as<double>(), nfit, nfit, kc[i].as<double>(), nfit, ws.p, s))) return rc;              // ALIGNF.py:36-41
        // a_i = sum(Kc_i * y y')  (ALIGNF.py:43-48)
        if ((rc = kmg_ew_weighted_dot(kc[i].as<double>(), nfit, nullptr, 0, dy.as<double>(), nfit, part.as<double>(), res.as<double>() + i, s))) return rc;
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));  // `full` is reused by the next kernel's upload
    }
    for (int i = 0; i < p; ++i)
        for (int j = i; j < p; ++j)  // M_ij = sum(Kc_i * Kc_j)  (ALIGNF.py:50-58)
            if ((rc = kmg_ew_weighted_dot(kc[i].as<double>(), nfit, kc[j].as<double>(), nfit, nullptr, nfit, part.as<double>(),
                                          res.as<double>() + p + i * p + j, s))) return rc;
    std::vector<double> h((size_t)(p + p * p), 0.0);
    KMG_CUDA_CHECK(cudaMemcpyAsync(h.data(), res.p, h.size() * 8, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int i = 0; i < p; ++i) a[i] = h[i];
    for (int i = 0; i < p; ++i)
        for (int j = i; j < p; ++j) M[i * p + j] = M[j * p + i] = h[p + i * p + j];
    return KMG_OK;
}

int kmg_nlck_grad_host(const double* const* Ks_fit, int p, int64_t nfit, const double* u, const double* alpha, int degree,
                       double* grad) {
    int rc = require_device();
    if (rc) return rc;
    KMG_REQUIRE(p >= 1 && p <= KMG_MAX_COMBINE && nfit >= 0 && Ks_fit && u && alpha && grad && degree >= 1, KMG_ERR_ARG, "nlck_grad: bad arguments");
    if (nfit == 0) { for (int m = 0; m < p; ++m) grad[m] = 0.0; return KMG_OK; }
    cudaStream_t s;
    if ((rc = get_streams(&s, nullptr))) return rc;
    std::vector<DevBuf> bufs(p);
    const double* dptr[KMG_MAX_COMBINE];
    int64_t lds[KMG_MAX_COMBINE];
    for (int m = 0; m < p; ++m) {
        if ((rc = bufs[m].alloc((size_t)nfit * nfit * 8))) return rc;
        KMG_CUDA_CHECK(cudaMemcpyAsync(bufs[m].p, Ks_fit[m], (size_t)nfit * nfit * 8, cudaMemcpyHostToDevice, s));
        dptr[m] = bufs[m].as<double>();
        lds[m] = nfit;
    }
    DevBuf kt, da, part, res;
    if ((rc = kt.alloc((size_t)nfit * nfit * 8))) return rc;
    if ((rc = da.alloc((size_t)nfit * 8))) return rc;
    if ((rc = part.alloc((size_t)nfit * 8))) return rc;
    if ((rc = res.alloc((size_t)p * 8))) return rc;
    KMG_CUDA_CHECK(cudaMemcpyAsync(da.p, alpha, (size_t)nfit * 8, cudaMemcpyHostToDevice, s));
    // K_t = (sum_m u_m K_m) ** (degree - 1)   (NLCKernels.py:62)
    if ((rc = kmg_ew_combine(dptr, lds, u, p, degree - 1, nfit, nfit, kt.as<double>(), nfit, s))) return rc;
    for (int m = 0; m < p; ++m)  // alpha' (K_t * K_m) alpha  (NLCKernels.py:65)
        if ((rc = kmg_ew_weighted_dot(kt.as<double>(), nfit, dptr[m], nfit, da.as<double>(), nfit, part.as<double>(), res.as<double>() + m, s))) return rc;
    std::vector<double> h((size_t)p);
    KMG_CUDA_CHECK(cudaMemcpyAsync(h.data(), res.p, (size_t)p * 8, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int m = 0; m < p; ++m) grad[m] = -(double)degree * h[m];  // NLCKernels.py:66
    return KMG_OK;
}

// ------------------------------------------------------------------------------------------
// device-pointer entry points
// ------------------------------------------------------------------------------------------
int kmg_pack_dev(const uint8_t* d_seqs, int seq_format, int64_t n, int L, uint32_t* d_planes, int* d_err_flag, void* stream) {
    return kmg_pack_launch(d_seqs, seq_format == KMG_SEQ_ASCII, n, L, d_planes, d_err_flag, (cudaStream_t)stream);
}

int64_t kmg_spectrum_phi_width(const int* ks, int nk) { return kmg_spectrum_padded_width(ks, nk, 0); }

int kmg_spectrum_phi_dev(const uint32_t* d_planes, int64_t n, int L, const int* ks, int nk, int8_t* d_phi, int64_t ld_phi,
                         void* stream) {
    return kmg_spectrum_phi_launch(d_planes, n, L, ks, nk, d_phi, ld_phi, (cudaStream_t)stream);
}

int kmg_gram_i8_dev(const int8_t* d_phi_rows, const int8_t* d_phi_cols, int64_t rows, int64_t cols, int64_t width,
                    int64_t ld_phi, int64_t row_index0, int64_t col_index0, void* d_out, int64_t ldo, int out_dtype,
                    int symmetric, const double* d_sd_rows, const double* d_sd_cols, int m_sub, void* stream) {
    KMG_REQUIRE(out_dtype == KMG_OUT_S32 || out_dtype == KMG_OUT_F64, KMG_ERR_ARG, "gram_i8: bad out_dtype");
    KMG_REQUIRE(!(d_sd_rows && out_dtype != KMG_OUT_F64), KMG_ERR_ARG, "gram_i8: normalisation needs the f64 output");
    KMG_REQUIRE((d_sd_rows == nullptr) == (d_sd_cols == nullptr), KMG_ERR_ARG, "gram_i8: sd_rows and sd_cols go together");
    if (symmetric)
        KMG_REQUIRE(rows == cols && row_index0 == col_index0 && d_phi_rows == d_phi_cols, KMG_ERR_ARG,
                    "gram_i8: symmetric mode needs a square diagonal block of one Phi");
    if (rows == 0 || cols == 0) return KMG_OK;
    GramI8Args a;
    memset(&a, 0, sizeof(a));
    a.phi_rows = d_phi_rows; a.phi_cols = d_phi_cols; a.rows = rows; a.cols = cols; a.Dpad = width; a.ld_phi = ld_phi;
    a.row_index0 = row_index0; a.col_index0 = col_index0; a.out = d_out; a.ldo = ldo; a.out_dtype = out_dtype;
    a.symmetric = symmetric; a.out_t = d_out; a.ldo_t = ldo; a.sd_rows = d_sd_rows; a.sd_cols = d_sd_cols; a.m_sub = m_sub;
    return kmg_gram_i8_launch(&a, (cudaStream_t)stream);
}

// Sharded symmetric spectrum Gram: part `part` of `n_parts` computes its share of the upper-triangle work
// and delivers every tile twice -- into its own block-row and, transposed, into the block-row of the part that owns the
// tile's columns: the mirror of kernels.py:45 is the one exchange step of the path.
//   d_stage == NULL : one launch; the GEMM epilogue stores the transposed tiles straight into the owners' buffers (peer
//                     memory).  Right for buffers on one device; over NVLink the scattered 256-byte stores cap the
//                     kernel (2 GPUs, n = 100 000: 40.5 ms against 30.1 ms with both buffers local).
//   d_stage != NULL : one launch per peer block (cyclic distance 1, 2, ...) whose epilogue writes the transposed block
//                     contiguously into local staging, each followed by ONE pitched peer copy on the copy stream while
//                     the next block's GEMM runs; the diagonal block (local mirror) goes last and hides the final copy.
namespace {
struct SubBlock { int b; int64_t r_lo, r_hi, c_lo, c_hi; };
void sharded_plan(int g, const int64_t* bounds, int a, std::vector<SubBlock>* out) {
    const int64_t a0 = bounds[a], a1 = bounds[a + 1];
    for (int d = 1; d < g; ++d) {
        const int b = (a + d) % g;
        const int64_t b0 = bounds[b], b1 = bounds[b + 1];
        if (2 * d < g) { out->push_back({b, a0, a1, b0, b1}); continue; }
        if (2 * d > g) continue;
        // distance g/2: split by the lower-numbered part's row tiles, as kmg_gram_sharded_takes does
        const int lo = a < b ? a : b;
        const int64_t lon = (bounds[lo + 1] - bounds[lo] + 255) / 256;
        const int64_t split = std::min<int64_t>(bounds[lo] + (lon + 1) / 2 * 256, bounds[lo + 1]);
        if (a < b) { if (split > a0) out->push_back({b, a0, split, b0, b1}); }
        else if (split < b1) out->push_back({b, a0, a1, split, b1});
    }
}
size_t stage_align(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

int kmg_gram_sharded_stage_bytes(int n_parts, const int64_t* part_row0, int part, int out_dtype, int64_t* bytes) {
    KMG_REQUIRE(n_parts >= 1 && n_parts <= KMG_MAX_PARTS && part_row0 && part >= 0 && part < n_parts && bytes, KMG_ERR_ARG,
                "gram_sharded_stage_bytes: bad arguments");
    std::vector<SubBlock> plan;
    sharded_plan(n_parts, part_row0, part, &plan);
    size_t total = 0;
    for (const SubBlock& sb : plan)
        total += stage_align((size_t)(sb.r_hi - sb.r_lo) * (size_t)(sb.c_hi - sb.c_lo) * (out_dtype == KMG_OUT_F64 ? 8 : 4));
    *bytes = (int64_t)total;
    return KMG_OK;
}

int kmg_gram_i8_sharded_dev(const int8_t* d_phi, int64_t n, int64_t width, int64_t ld_phi, int n_parts, int part,
                            const int64_t* part_row0, void* const* part_out, int64_t ldo, int out_dtype, const double* d_sd,
                            void* d_stage, int64_t* computed_entries, void* stream) {
    KMG_REQUIRE(out_dtype == KMG_OUT_S32 || out_dtype == KMG_OUT_F64, KMG_ERR_ARG, "gram_i8_sharded: bad out_dtype");
    KMG_REQUIRE(!(d_sd && out_dtype != KMG_OUT_F64), KMG_ERR_ARG, "gram_i8_sharded: normalisation needs the f64 output");
    KMG_REQUIRE(n_parts >= 1 && n_parts <= KMG_MAX_PARTS && part >= 0 && part < n_parts && part_row0 && part_out, KMG_ERR_ARG,
                "gram_i8_sharded: 1..%d parts", KMG_MAX_PARTS);
    KMG_REQUIRE(ldo >= n, KMG_ERR_ARG, "gram_i8_sharded: ldo < n");
    if (computed_entries) *computed_entries = 0;
    if (n == 0) return KMG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t a0 = part_row0[part], a1 = part_row0[part + 1];
    const int64_t esz = out_dtype == KMG_OUT_F64 ? 8 : 4;
    GramI8Args a;
    if (d_stage == nullptr) {
        memset(&a, 0, sizeof(a));
        a.phi_rows = d_phi + a0 * ld_phi; a.phi_cols = d_phi; a.rows = a1 - a0; a.cols = n; a.Dpad = width; a.ld_phi = ld_phi;
        a.row_index0 = a0; a.col_index0 = 0; a.out = part_out[part]; a.ldo = ldo; a.out_dtype = out_dtype;
        a.sd_rows = d_sd ? d_sd + a0 : nullptr; a.sd_cols = d_sd;
        a.n_parts = n_parts; a.part = part; a.part_row0 = part_row0; a.part_out = part_out; a.computed_entries = computed_entries;
        return kmg_gram_i8_launch(&a, s);
    }
    for (int q = 0; q <= n_parts; ++q)
        KMG_REQUIRE((q == n_parts ? part_row0[q] == n : part_row0[q] % 256 == 0) && (q == 0 ? part_row0[0] == 0 : part_row0[q] > part_row0[q - 1]),
                    KMG_ERR_ARG, "gram_i8_sharded: part boundaries must start at 0, increase in multiples of 256 and end at n");
    int rc;
    cudaStream_t s0, copy;
    if ((rc = get_streams(&s0, &copy))) return rc;
    std::vector<SubBlock> plan;
    sharded_plan(n_parts, part_row0, part, &plan);
    // Order: the half block at distance g/2 first (the first peer copy starts after the shortest launch), then distance
    // 1, 2, ...  At every phase rank a sends to (a + d) mod g: a permutation, so each receiver has exactly one incoming
    // stream at a time.  (Measured on 8 GPUs, n = 200 000: 38.5 ms/step this way, 35.3 ms of it the GEMM phase; splitting
    // the blocks and feeding two peers at once through two copy streams broke the pattern and cost 7 ms; two copy
    // engines on the same block changed nothing -- the copies are not the limit, the doubled epilogue stores and the
    // 35 GB of copy traffic through HBM are: the same launches take 30.1 ms with no peer traffic at all.)
    if (!plan.empty() && n_parts % 2 == 0) std::rotate(plan.begin(), plan.end() - 1, plan.end());
    char* my = static_cast<char*>(part_out[part]);
    char* stage = static_cast<char*>(d_stage);
    int64_t total = 0, got = 0;
    cudaEvent_t ev;
    for (const SubBlock& sb : plan) {
        const int64_t rows = sb.r_hi - sb.r_lo, cols = sb.c_hi - sb.c_lo;
        memset(&a, 0, sizeof(a));
        a.phi_rows = d_phi + sb.r_lo * ld_phi; a.phi_cols = d_phi + sb.c_lo * ld_phi; a.rows = rows; a.cols = cols; a.Dpad = width; a.ld_phi = ld_phi;
        a.row_index0 = sb.r_lo; a.col_index0 = sb.c_lo; a.out = my + ((sb.r_lo - a0) * ldo + sb.c_lo) * esz; a.ldo = ldo; a.out_dtype = out_dtype;
        a.sd_rows = d_sd ? d_sd + sb.r_lo : nullptr; a.sd_cols = d_sd ? d_sd + sb.c_lo : nullptr;
        a.mirror_all = 1; a.out_t = stage; a.ldo_t = rows; a.computed_entries = &got;
        if ((rc = kmg_gram_i8_launch(&a, s))) return rc;
        total += got;
        // the transposed block (cols x rows, contiguous) -> rows [c_lo, c_hi) x columns [r_lo, r_hi) of the owner's block-row
        KMG_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        KMG_CUDA_CHECK(cudaEventRecord(ev, s));
        KMG_CUDA_CHECK(cudaStreamWaitEvent(copy, ev, 0));
        KMG_CUDA_CHECK(cudaEventDestroy(ev));
        char* dst = static_cast<char*>(part_out[sb.b]) + ((sb.c_lo - part_row0[sb.b]) * ldo + sb.r_lo) * esz;
        KMG_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)(ldo * esz), stage, (size_t)(rows * esz), (size_t)(rows * esz), (size_t)cols,
                                         cudaMemcpyDefault, copy));
        stage += stage_align((size_t)(rows * cols * esz));
    }
    // diagonal block: the single-GPU symmetric build on this part's own square
    memset(&a, 0, sizeof(a));
    a.phi_rows = d_phi + a0 * ld_phi; a.phi_cols = a.phi_rows; a.rows = a1 - a0; a.cols = a1 - a0; a.Dpad = width; a.ld_phi = ld_phi;
    a.row_index0 = a0; a.col_index0 = a0; a.out = my + a0 * esz; a.ldo = ldo; a.out_dtype = out_dtype;
    a.symmetric = 1; a.out_t = a.out; a.ldo_t = ldo;
    a.sd_rows = d_sd ? d_sd + a0 : nullptr; a.sd_cols = a.sd_rows; a.computed_entries = &got;
    if ((rc = kmg_gram_i8_launch(&a, s))) return rc;
    total += got;
    // `stream` completes only after the peer copies have
    KMG_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    KMG_CUDA_CHECK(cudaEventRecord(ev, copy));
    KMG_CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));
    KMG_CUDA_CHECK(cudaEventDestroy(ev));
    if (computed_entries) *computed_entries = total;
    return KMG_OK;
}

// Measured int8 tensor-core peak (mma_peak.cu): enqueue `iters` x 4 back-to-back MMAs per CTA pair; the caller times it.
int kmg_mma_peak_i8_dev(int iters, int64_t* ops, void* stream) {
    int rc = require_device();
    if (rc) return rc;
    return kmg_mma_peak_i8_launch(iters, ops, (cudaStream_t)stream);
}

int kmg_gram_sharded_takes_host(int n_parts, const int64_t* part_row0, int a, int b, int64_t I, int64_t J) {
    KMG_REQUIRE(n_parts >= 1 && n_parts <= KMG_MAX_PARTS && part_row0 && a >= 0 && a < n_parts && b >= 0 && b < n_parts, KMG_ERR_ARG,
                "gram_sharded_takes: bad arguments");
    return kmg_gram_sharded_takes(n_parts, part_row0, a, b, I, J);
}

// ---- CUDA IPC: how the ranks of one node see each other's block-row buffers (cudaMalloc'ed by kmg_dev_malloc) ----
int kmg_ipc_export(const void* d_ptr, uint8_t* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    KMG_REQUIRE(d_ptr && handle64, KMG_ERR_ARG, "ipc_export: null argument");
    cudaIpcMemHandle_t h;
    KMG_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle64, &h, 64);
    return KMG_OK;
}

int kmg_ipc_open(const uint8_t* handle64, void** d_ptr) {
    KMG_REQUIRE(d_ptr && handle64, KMG_ERR_ARG, "ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    KMG_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KMG_OK;
}

int kmg_ipc_close(void* d_ptr) {
    if (d_ptr) KMG_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
    return KMG_OK;
}

int kmg_gram_i8_simt_dev(const int8_t* d_phi_rows, const int8_t* d_phi_cols, int64_t rows, int64_t cols, int64_t width,
                         int64_t ld_phi, int32_t* d_out, int64_t ldo, void* stream) {
    return kmg_gram_i8_simt_launch(d_phi_rows, d_phi_cols, ld_phi, rows, cols, width, d_out, ldo, (cudaStream_t)stream);
}

int kmg_mismatch_phi_dev(const uint32_t* d_planes, int64_t n, int L, int k, int m, int8_t* d_phi, int64_t ld_phi, void* stream) {
    return kmg_mismatch_phi_launch(d_planes, n, L, k, m, d_phi, ld_phi, (cudaStream_t)stream);
}

int kmg_phi_diag_sqrt_dev(const int8_t* d_phi, int64_t n, int64_t width, int64_t ld_phi, double* d_sd, void* stream) {
    return kmg_phi_diag_sqrt_launch(d_phi, n, width, ld_phi, d_sd, (cudaStream_t)stream);
}

static PairBlock make_block(const uint32_t* pr, const uint32_t* pc, int64_t rows, int64_t cols, int64_t r0, int64_t c0, int L,
                            void* out, int64_t ldo, int dtype, int symmetric, const double* sdr, const double* sdc) {
    PairBlock b;
    memset(&b, 0, sizeof(b));
    b.planes_rows = pr; b.planes_cols = pc; b.rows = rows; b.cols = cols; b.row_index0 = r0; b.col_index0 = c0; b.L = L;
    b.out = out; b.ldo = ldo; b.out_dtype = dtype; b.symmetric = symmetric; b.out_t = out; b.ldo_t = ldo;
    b.sd_rows = sdr; b.sd_cols = sdc;
    return b;
}

int kmg_mismatch_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols,
                     int64_t row_index0, int64_t col_index0, int L, int k, int m, void* d_out, int64_t ldo, int out_dtype,
                     int symmetric, const double* d_sd_rows, const double* d_sd_cols, void* stream) {
    KMG_REQUIRE(out_dtype == KMG_OUT_S32 || out_dtype == KMG_OUT_F64, KMG_ERR_ARG, "mismatch: bad out_dtype");
    KMG_REQUIRE(!(d_sd_rows && out_dtype != KMG_OUT_F64), KMG_ERR_ARG, "mismatch: normalisation needs the f64 output");
    KMG_REQUIRE((d_sd_rows == nullptr) == (d_sd_cols == nullptr), KMG_ERR_ARG, "mismatch: sd_rows and sd_cols go together");
    PairBlock b = make_block(d_planes_rows, d_planes_cols, rows, cols, row_index0, col_index0, L, d_out, ldo, out_dtype, symmetric,
                             d_sd_rows, d_sd_cols);
    return kmg_mismatch_launch(&b, k, m, (cudaStream_t)stream);
}

int kmg_mismatch_diag_dev(const uint32_t* d_planes, int64_t n, int L, int k, int m, double* d_sd, void* stream) {
    return kmg_mismatch_diag_launch(d_planes, n, L, k, m, d_sd, (cudaStream_t)stream);
}

int kmg_wd_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols, int64_t row_index0,
               int64_t col_index0, int L, int d, double* d_out, int64_t ldo, int symmetric, void* stream) {
    PairBlock b = make_block(d_planes_rows, d_planes_cols, rows, cols, row_index0, col_index0, L, d_out, ldo, KMG_OUT_F64, symmetric,
                             nullptr, nullptr);
    return kmg_wd_launch(&b, d, (cudaStream_t)stream);
}

int kmg_wds_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols, int64_t row_index0,
                int64_t col_index0, int L, int d, int S, double* d_out, int64_t ldo, int symmetric, void* stream) {
    PairBlock b = make_block(d_planes_rows, d_planes_cols, rows, cols, row_index0, col_index0, L, d_out, ldo, KMG_OUT_F64, symmetric,
                             nullptr, nullptr);
    return kmg_wds_launch(&b, d, S, (cudaStream_t)stream);
}

int kmg_la_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols, int64_t row_index0,
               int64_t col_index0, int L, double e, double d, double beta, int smith, double* d_out, int64_t ldo,
               int symmetric, void* stream) {
    PairBlock b = make_block(d_planes_rows, d_planes_cols, rows, cols, row_index0, col_index0, L, d_out, ldo, KMG_OUT_F64, symmetric,
                             nullptr, nullptr);
    return kmg_la_launch(&b, e, d, beta, smith, (cudaStream_t)stream);
}

int kmg_normalize_dev(double* d_K, int64_t n, int64_t ld, double* d_sd_scratch, void* stream) {
    int rc = kmg_ew_diag_sqrt(d_K, n, ld, d_sd_scratch, (cudaStream_t)stream);
    if (rc) return rc;
    return kmg_ew_normalize(d_K, n, ld, d_sd_scratch, (cudaStream_t)stream);
}

int64_t kmg_center_workspace_bytes(int64_t n) { return kmg_ew_center_workspace(n); }

int kmg_center_dev(const double* d_K, int64_t n, int64_t ld, double* d_out, int64_t ldo, void* d_workspace, void* stream) {
    return kmg_ew_center(d_K, n, ld, d_out, ldo, d_workspace, (cudaStream_t)stream);
}

int kmg_row_sums_dev(const double* d_K, int64_t rows, int64_t cols, int64_t ld, double* d_rs, void* stream) {
    return kmg_ew_row_sums(d_K, rows, cols, ld, d_rs, (cudaStream_t)stream);
}

int64_t kmg_col_sums_workspace_bytes(int64_t rows, int64_t cols) { return kmg_ew_col_sums_workspace(rows, cols); }

int kmg_col_sums_dev(const double* d_K, int64_t rows, int64_t cols, int64_t ld, double* d_cs, void* d_workspace, void* stream) {
    return kmg_ew_col_sums(d_K, rows, cols, ld, d_cs, d_workspace, (cudaStream_t)stream);
}

int kmg_center_apply_dev(const double* d_K, int64_t rows, int64_t cols, int64_t n_total, int64_t ld, const double* d_rs,
                         const double* d_cs, const double* d_g, double* d_out, int64_t ldo, void* stream) {
    return kmg_ew_center_apply(d_K, rows, cols, n_total, ld, d_rs, d_cs, d_g, d_out, ldo, (cudaStream_t)stream);
}

int kmg_gather_dev(const double* d_K, int64_t ld, const int64_t* d_idx, int64_t m, double* d_out, int64_t ldo, void* stream) {
    return kmg_ew_gather(d_K, ld, d_idx, m, d_out, ldo, (cudaStream_t)stream);
}

int kmg_combine_dev(const double* const* d_Ks, const int64_t* lds, const double* u, int p, int degree, int64_t rows, int64_t cols,
                    double* d_out, int64_t ldo, void* stream) {
    return kmg_ew_combine(d_Ks, lds, u, p, degree, rows, cols, d_out, ldo, (cudaStream_t)stream);
}

int kmg_weighted_dot_dev(const double* d_A, int64_t lda, const double* d_B, int64_t ldb, const double* d_w, int64_t n,
                         double* d_partial, double* d_result, void* stream) {
    return kmg_ew_weighted_dot(d_A, lda, d_B, ldb, d_w, n, d_partial, d_result, (cudaStream_t)stream);
}

}  // extern "C"
