// host_link.h -- device -> host delivery of Gram blocks (host_link.cu): pinned staging ring with copy / widening threads,
// narrow integer transports, recycled result blocks, and the block-row driver behind every `*_host` builder.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// builds rows [r0, r0 + rows) of the Gram into d_out (row stride ldo elements) on stream s
typedef int (*BlockFn)(void* ctx, int64_t r0, int64_t rows, void* d_out, int64_t ldo, int symmetric, cudaStream_t s);

// Build an nr x nc Gram on the device -- whole, in row chunks, or streamed through two buffers when it does not fit --
// and deliver it as doubles into K (row stride ldk).  out_s32: `fn` writes int32 counts, which cross the link as u16
// or s32 and are widened by the copy threads.
int kmg_hl_build_to_host(int64_t nr, int64_t nc, bool symmetric, BlockFn fn, void* ctx, double* K, int64_t ldk, bool out_s32 = false);
// pageable host -> device through the pinned slots with parallel copy threads; complete on return
int kmg_hl_h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t s);
// widest row (in bytes on the link) the staging ring takes
size_t kmg_hl_slot_bytes();
// recycled host blocks behind kmg_host_alloc / kmg_host_free / kmg_release
int kmg_hl_host_alloc(int64_t bytes, void** ptr);
int kmg_hl_host_free(void* ptr);
void kmg_hl_trim_pool();
// delivery mode of the fp64 results (0 widen: narrow transport + copy threads, 1 dma: copy engine into pinned result
// blocks, 2 mapped: kernel stores through the mapped address); default from KMG_D2H_MODE, else widen
int kmg_hl_set_mode(int mode);
int kmg_hl_get_mode();
// true when kmg_hl_build_to_host will deliver fp64 straight into K (dma / mapped): the caller then builds fp64, not s32
bool kmg_hl_direct_fp64(const double* K, int64_t ldk, int64_t nr);
