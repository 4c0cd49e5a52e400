// host_link.cu -- see host_link.h.  Host-side orchestration only: the arithmetic on Gram entries happens in the kernels
// `fn` launches; the copy threads move bytes and widen exact integers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <future>
#include <map>
#include <mutex>

#include "../../include/kmg.h"
#include "elementwise.h"
#include "host_link.h"
#include "kmg_common.cuh"
#include "runtime.h"

namespace {

// Row-block size so that two output buffers of `rows x cols` doubles fit in a fraction of free memory.
int pick_block_rows(int64_t nr, int64_t nc, int64_t* block_rows) {
    size_t free_b = 0, total_b = 0;
    KMG_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    free_b += kmg_rt_cached_bytes();  // cached buffers are reclaimable
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
        // no early return from here on: copy threads of earlier slots may still be writing into `dst`
        if (cudaMemcpyAsync(g_ring.buf[slot], src + (size_t)r * row_bytes, (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaEventRecord(g_ring.ev[slot], s) != cudaSuccess) {
            kmg_set_error("device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            err = KMG_ERR_CUDA;
            break;
        }
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

}  // namespace

// Pageable host -> device copy through the same pinned slots: a plain cudaMemcpyAsync from pageable memory is staged by
// the driver on the calling thread (~10 GB/s: 2 ms for the 20 MB of 200 000 sequences); here up to 16 threads copy 1 MB
// pieces into pinned slots and enqueue their own DMA.  Returns when every piece has landed on the device.
int kmg_hl_h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t s) {
    constexpr size_t PIECE = 1u << 20;
    if (bytes < 4 * PIECE) {
        KMG_CUDA_CHECK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s));
        return KMG_OK;
    }
    std::lock_guard<std::mutex> lk(g_ring.mu);
    int rc = ring_init();
    if (rc) return rc;
    int err = KMG_OK;
    for (size_t off = 0; off < bytes && !err; off += PIECE, g_ring.slot = (g_ring.slot + 1) % D2H_SLOTS) {
        const int slot = g_ring.slot;
        const size_t len = std::min(PIECE, bytes - off);
        if (g_ring.fut[slot].valid() && g_ring.fut[slot].get() != 0) { err = KMG_ERR_CUDA; break; }
        void* stage = g_ring.buf[slot];
        cudaEvent_t ev = g_ring.ev[slot];
        char* dst = static_cast<char*>(d_dst) + off;
        const char* src = static_cast<const char*>(h_src) + off;
        int dev = 0;
        cudaGetDevice(&dev);
        g_ring.fut[slot] = std::async(std::launch::async, [=]() -> int {
            memcpy(stage, src, len);
            if (cudaSetDevice(dev) != cudaSuccess) return 1;
            if (cudaMemcpyAsync(dst, stage, len, cudaMemcpyHostToDevice, s) != cudaSuccess) return 1;
            if (cudaEventRecord(ev, s) != cudaSuccess) return 1;
            return cudaEventSynchronize(ev) == cudaSuccess ? 0 : 1;  // the slot is free again once its DMA is done
        });
    }
    const int e2 = ring_drain_locked();
    if (err || e2) { kmg_set_error("host-to-device copy failed"); return err ? err : e2; }
    return KMG_OK;
}

namespace {

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
    std::map<void*, size_t> registered;  // blocks pinned with cudaHostRegister (DMA / mapped delivery), live or cached
    size_t cached_bytes = 0;
    void unmap(void* p, size_t sz) {
        auto it = registered.find(p);
        if (it != registered.end()) { cudaHostUnregister(p); registered.erase(it); }
        munmap(p, sz);
    }
    size_t cap() const {
        if (const char* v = getenv("KMG_HOST_POOL_BYTES")) return (size_t)atof(v);
        return (size_t)8 << 30;
    }
    void trim(size_t keep) {
        while (cached_bytes > keep && !cached.empty()) {
            auto it = std::prev(cached.end());
            unmap(it->second, it->first);
            cached_bytes -= it->first;
            cached.erase(it);
        }
    }
};
HostPool g_hostpool;


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

// ---- delivery modes --------------------------------------------------------------------------------------------------
//   widen  (0) narrow integer transport over PCIe + copy threads that widen to fp64 while writing the caller's array:
//              2-4 B/entry on the link, but every delivered byte is a CPU store -- bound by the cores a process has
//              (one GPU, 16 threads: ~1.2e10 entries/s; eight processes on one host share the same cores and DRAM).
//   dma    (1) the kernel writes fp64 on the device and the copy engine writes it straight into the caller's array
//              (result blocks from kmg_host_alloc, pinned once with cudaHostRegister): 8 B/entry on the link, no CPU
//              in the data path -- PCIe-bound per GPU (~6.5e9 entries/s) but it scales with the number of GPUs.
//   mapped (2) the kernel's epilogue stores fp64 through the mapped address of the same pinned block (zero copy).
int g_mode = -1;
int current_mode() {
    if (g_mode < 0) {
        const char* v = getenv("KMG_D2H_MODE");
        g_mode = (v && (!strcmp(v, "dma") || !strcmp(v, "1"))) ? 1 : ((v && (!strcmp(v, "mapped") || !strcmp(v, "2"))) ? 2 : 0);
    }
    return g_mode;
}

// Is [K, K + bytes) inside a live result block?  Pins the block on first use.  Returns the device-visible address.
bool pinned_target(const double* K, size_t bytes, void** dev_addr) {
    std::lock_guard<std::mutex> lk(g_hostpool.mu);
    auto it = g_hostpool.live.upper_bound(const_cast<double*>(K));
    if (it == g_hostpool.live.begin()) return false;
    --it;
    char* base = static_cast<char*>(it->first);
    const char* p = reinterpret_cast<const char*>(K);
    if (p < base || p + bytes > base + it->second) return false;
    if (!g_hostpool.registered.count(it->first)) {
        if (cudaHostRegister(it->first, it->second, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        g_hostpool.registered[it->first] = it->second;
    }
    if (dev_addr) {
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, it->first, 0) != cudaSuccess) { cudaGetLastError(); return false; }
        *dev_addr = static_cast<char*>(d) + (p - base);
    }
    return true;
}
}  // namespace

int kmg_hl_set_mode(int mode) {
    KMG_REQUIRE(mode >= 0 && mode <= 2, KMG_ERR_ARG, "d2h mode: 0 widen, 1 dma, 2 mapped");
    if (mode == 0 && g_mode > 0) {
        // back to the copy threads: un-pin the result blocks (measured: the widening threads write registered memory at
        // about half the rate of plain pageable huge pages)
        std::lock_guard<std::mutex> lk(g_hostpool.mu);
        for (auto& kv : g_hostpool.registered) cudaHostUnregister(kv.first);
        g_hostpool.registered.clear();
    }
    g_mode = mode;
    return KMG_OK;
}
int kmg_hl_get_mode() { return current_mode(); }
// true: kmg_hl_build_to_host will deliver fp64 by DMA / mapped stores into K -- the caller then asks its kernel for fp64
bool kmg_hl_direct_fp64(const double* K, int64_t ldk, int64_t nr) {
    if (current_mode() == 0 || nr <= 0) return false;
    return pinned_target(K, (size_t)((nr - 1) * ldk + ldk) * sizeof(double), nullptr);
}

int kmg_hl_build_to_host(int64_t nr, int64_t nc, bool symmetric, BlockFn fn, void* ctx, double* K, int64_t ldk, bool out_s32) {
    if (nr == 0 || nc == 0) return KMG_OK;
    const size_t esz = out_s32 ? sizeof(int32_t) : sizeof(double);
    cudaStream_t s0, s1;
    int rc = kmg_rt_get_streams(&s0, &s1);
    if (rc) return rc;
    int64_t br = 0;
    if ((rc = pick_block_rows(nr, nc, &br))) return rc;
    void* mapped = nullptr;
    const bool direct = !out_s32 && current_mode() != 0 && !(symmetric && br < nr) && pinned_target(K, (size_t)((nr - 1) * ldk + ldk) * sizeof(double), &mapped);
    if (direct && current_mode() == 2) {
        // zero copy: the producing kernel stores through the mapped address of the caller's (pinned) array
        if ((rc = fn(ctx, 0, nr, mapped, ldk, symmetric ? 1 : 0, s0))) { cudaStreamSynchronize(s0); return rc; }
        KMG_CUDA_CHECK(cudaStreamSynchronize(s0));
        return KMG_OK;
    }
    if (direct) {
        // copy engine straight into the caller's (pinned) array: row chunks so the link starts while the GPU still builds
        const int64_t rows_cap = std::min<int64_t>(br, nr);
        const int nchunks = symmetric ? 1 : (int)std::min<int64_t>(8, std::max<int64_t>(1, nr / 256));
        const int64_t crow = symmetric ? nr : std::min<int64_t>(rows_cap, ((nr + nchunks - 1) / nchunks + 255) / 256 * 256);
        DevBuf buf[2];
        struct Quiesce2 {
            cudaStream_t a, b;
            ~Quiesce2() { cudaStreamSynchronize(a); cudaStreamSynchronize(b); }
        } quiesce{s0, s1};
        for (int i = 0; i < 2; ++i)
            if ((rc = buf[i].alloc((size_t)crow * nc * sizeof(double)))) return rc;
        cudaEvent_t built[2] = {}, copied[2] = {};
        for (int i = 0; i < 2; ++i) {
            KMG_CUDA_CHECK(cudaEventCreateWithFlags(&built[i], cudaEventDisableTiming));
            KMG_CUDA_CHECK(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        }
        int c = 0;
        for (int64_t r0 = 0; r0 < nr && rc == KMG_OK; r0 += crow, ++c) {
            const int64_t rows = std::min<int64_t>(crow, nr - r0);
            const int b = c & 1;
            if (c >= 2 && cudaStreamWaitEvent(s0, copied[b], 0) != cudaSuccess) { rc = KMG_ERR_CUDA; break; }  // buffer drained
            if ((rc = fn(ctx, r0, rows, buf[b].p, nc, symmetric ? 1 : 0, s0))) break;
            if (cudaEventRecord(built[b], s0) != cudaSuccess || cudaStreamWaitEvent(s1, built[b], 0) != cudaSuccess ||
                cudaMemcpy2DAsync(K + r0 * ldk, (size_t)ldk * 8, buf[b].p, (size_t)nc * 8, (size_t)nc * 8, (size_t)rows, cudaMemcpyDeviceToHost, s1) != cudaSuccess ||
                cudaEventRecord(copied[b], s1) != cudaSuccess) {
                kmg_set_error("build_to_host (dma): %s", cudaGetErrorString(cudaGetLastError()));
                rc = KMG_ERR_CUDA;
            }
        }
        if (rc == KMG_OK && cudaStreamSynchronize(s1) != cudaSuccess) { kmg_set_error("build_to_host (dma): copy failed"); rc = KMG_ERR_CUDA; }
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(built[i]); cudaEventDestroy(copied[i]); }
        return rc;
    }
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
    // On EVERY exit -- error returns included -- nothing of this call may still be running when the buffers go back to
    // the allocation cache (another call could be handed them while a kernel still writes) or when control returns to
    // a caller that may free K (copy threads still writing it): both streams are synchronised and the ring is drained
    // before `buf` is destroyed (declared after it, so destroyed first).
    struct Quiesce {
        cudaStream_t a, b;
        ~Quiesce() { cudaStreamSynchronize(a); cudaStreamSynchronize(b); ring_drain(); }
    } quiesce{s0, s1};
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


size_t kmg_hl_slot_bytes() { return D2H_SLOT_BYTES; }

// ---- recycled host memory for result arrays (kmg/host.py wraps a block as the numpy array it returns) ----
int kmg_hl_host_alloc(int64_t bytes, void** ptr) {
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

int kmg_hl_host_free(void* ptr) {
    if (!ptr) return KMG_OK;
    std::lock_guard<std::mutex> lk(g_hostpool.mu);
    auto it = g_hostpool.live.find(ptr);
    KMG_REQUIRE(it != g_hostpool.live.end(), KMG_ERR_ARG, "host_free: pointer was not returned by kmg_host_alloc");
    const size_t sz = it->second;
    g_hostpool.live.erase(it);
    const size_t cap = g_hostpool.cap();
    if (sz > cap) { g_hostpool.unmap(ptr, sz); return KMG_OK; }
    g_hostpool.cached.emplace(sz, ptr);
    g_hostpool.cached_bytes += sz;
    if (g_hostpool.cached_bytes > cap) {  // evict the other blocks, largest first, keeping the one just returned
        for (auto c = g_hostpool.cached.end(); g_hostpool.cached_bytes > cap && c != g_hostpool.cached.begin();) {
            --c;
            if (c->second == ptr) continue;
            g_hostpool.unmap(c->second, c->first);
            g_hostpool.cached_bytes -= c->first;
            c = g_hostpool.cached.erase(c);
        }
    }
    return KMG_OK;
}


void kmg_hl_trim_pool() {
    std::lock_guard<std::mutex> lk(g_hostpool.mu);
    g_hostpool.trim(0);
}
