// p2p_probe.cu -- dev tool (not part of libkmg): what store pattern does an SM-issued NVLink write stream need to reach
// the link rate?  One process, two GPUs with peer access: a kernel on GPU 0 writes `total` bytes into a buffer on GPU 1
// (and, with --bidir, GPU 1 writes into GPU 0 at the same time), with the store patterns the Gram epilogue could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/p2p_probe tools/p2p_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// rows of `row_doubles` doubles, row pitch `ld` doubles.  Work item w = (piece of `piece` doubles in a row, block of
// `box_rows` consecutive rows); a warp takes items round-robin.
struct Shape { int64_t ld; int64_t rows; int64_t row_doubles; };

// mode 0: thread-issued 8-byte stores, a warp stores 256-byte pieces of 16 consecutive rows (the round-1 mirror)
// mode 1: thread-issued 16-byte stores, fully sequential (a warp streams 512 B per instruction along a row)
__global__ void __launch_bounds__(256) lsu_kernel(double* dst, Shape s, int mode, int64_t items) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * 8 + warp, nw = (int64_t)gridDim.x * 8;
    if (mode == 0) {
        const int64_t ppr = s.row_doubles / 32;  // pieces per row
        for (int64_t it = gw; it < items; it += nw) {
            const int64_t rb = it / ppr, pc = it % ppr;
#pragma unroll 4
            for (int r = 0; r < 16; ++r) dst[(rb * 16 + r) * s.ld + pc * 32 + lane] = (double)(it + r);
        }
    } else {
        const int64_t ppr = s.row_doubles / 1024;  // 8 KB per item: 16 instructions of 512 B
        for (int64_t it = gw; it < items; it += nw) {
            const int64_t row = it / ppr, pc = it % ppr;
            double2* p = reinterpret_cast<double2*>(dst + row * s.ld + pc * 1024);
#pragma unroll 4
            for (int q = 0; q < 16; ++q) p[q * 32 + lane] = make_double2((double)it, (double)q);
        }
    }
}

// mode 2: TMA tensor stores, one box per item from a per-warp shared-memory tile (box = bx doubles x by rows <= 4 KB...16 KB)
__global__ void __launch_bounds__(256) tma_kernel(const __grid_constant__ CUtensorMap tm, Shape s, int bx, int by, int64_t items, int depth) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int box_bytes = bx * by * 8;
    uint8_t* tile = smem + (size_t)warp * box_bytes * depth;
    for (int i = lane; i < box_bytes * depth / 8; i += 32) reinterpret_cast<double*>(tile)[i] = (double)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const int64_t gw = (int64_t)blockIdx.x * 8 + warp, nw = (int64_t)gridDim.x * 8;
    const int64_t ppr = s.row_doubles / bx;
    int slot = 0;
    if (lane == 0) {
        for (int64_t it = gw; it < items; it += nw) {
            const int64_t rb = it / ppr, pc = it % ppr;
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(tile + (size_t)slot * box_bytes)), "r"((int)(pc * bx)), "r"((int)(rb * by)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++slot == depth) slot = 0;
            // allow depth-1 groups in flight (source reuse)
            if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const bool local_only = ndev < 2;
    int bidir = 0, ctas = 148;
    for (int i = 1; i < argc; ++i) { if (!strcmp(argv[i], "--bidir")) bidir = 1; if (!strcmp(argv[i], "--ctas") && i + 1 < argc) ctas = atoi(argv[++i]); }
    const int D = local_only ? 1 : 2;
    const int64_t ld = 200000, rows = 6144, row_doubles = 196608;  // 6144 rows x 1.5 MB = 9.66 GB written per run
    Shape s{ld, rows, row_doubles};
    double* buf[2] = {nullptr, nullptr};
    cudaStream_t st[2];
    cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < D; ++d) {
        CK(cudaSetDevice(d));
        if (!local_only) { cudaError_t e = cudaDeviceEnablePeerAccess(1 - d, 0); if (e != cudaSuccess) { printf("peer access %d->%d: %s\n", d, 1 - d, cudaGetErrorString(e)); return 1; } }
        CK(cudaMalloc(&buf[d], (size_t)ld * rows * 8));
        CK(cudaStreamCreate(&st[d]));
        CK(cudaEventCreate(&e0[d]));
        CK(cudaEventCreate(&e1[d]));
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    PFN_encodeTiled enc = (PFN_encodeTiled)fn;
    const double gb = (double)rows * row_doubles * 8 / 1e9;
    struct Cfg { const char* name; int mode, bx, by, depth; } cfgs[] = {
        {"lsu 8B/lane, 256 B pieces of 16 rows (round-1 mirror)", 0, 0, 0, 0},
        {"lsu 16B/lane sequential 8 KB runs", 1, 0, 0, 0},
        {"tma box 32 x 16 (256 B pieces, 4 KB) depth 1", 2, 32, 16, 1},
        {"tma box 32 x 16 (256 B pieces, 4 KB) depth 2", 2, 32, 16, 2},
        {"tma box 128 x 4 (1 KB pieces, 4 KB) depth 2", 2, 128, 4, 2},
        {"tma box 128 x 16 (1 KB pieces, 16 KB) depth 1", 2, 128, 16, 1},
        {"tma box 256 x 8 (2 KB pieces, 16 KB) depth 1", 2, 256, 8, 1},
        {"tma box 256 x 2 (2 KB pieces, 4 KB) depth 4", 2, 256, 2, 4},
    };
    for (int pass = 0; pass < (local_only ? 1 : 2); ++pass) {
        const int remote = local_only ? 0 : 1 - pass;  // pass 0: remote destination, pass 1: local destination (baseline)
        for (const Cfg& c : cfgs) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                const int nsrc = (bidir && remote) ? 2 : 1;
                for (int d = 0; d < nsrc; ++d) {
                    CK(cudaSetDevice(d));
                    double* dst = buf[remote ? 1 - d : d];
                    CK(cudaEventRecord(e0[d], st[d]));
                    if (c.mode < 2) {
                        const int64_t items = c.mode == 0 ? (rows / 16) * (row_doubles / 32) : rows * (row_doubles / 1024);
                        lsu_kernel<<<ctas, 256, 0, st[d]>>>(dst, s, c.mode, items);
                    } else {
                        CUtensorMap tm;
                        cuuint64_t dims[2] = {(cuuint64_t)row_doubles, (cuuint64_t)rows};
                        cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
                        cuuint32_t box[2] = {(cuuint32_t)c.bx, (cuuint32_t)c.by};
                        cuuint32_t es[2] = {1, 1};
                        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
                        const size_t smem = (size_t)8 * c.bx * c.by * 8 * c.depth;
                        CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        const int64_t items = (rows / c.by) * (row_doubles / c.bx);
                        tma_kernel<<<ctas, 256, smem, st[d]>>>(tm, s, c.bx, c.by, items, c.depth);
                    }
                    CK(cudaGetLastError());
                    CK(cudaEventRecord(e1[d], st[d]));
                }
                float worst = 0.f;
                for (int d = 0; d < nsrc; ++d) {
                    CK(cudaSetDevice(d));
                    CK(cudaStreamSynchronize(st[d]));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
                    if (ms > worst) worst = ms;
                }
                if (worst < best) best = worst;
            }
            printf("ctas=%d %-7s%s %-58s %8.2f ms  %7.1f GB/s per direction\n", ctas, remote ? "REMOTE" : "local", (bidir && remote) ? " bidir" : "      ", c.name, best, gb / (best * 1e-3));
            fflush(stdout);
        }
    }
    // copy-engine reference: one 2-D peer copy of the same shape
    if (!local_only) {
        CK(cudaSetDevice(0));
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0[0], st[0]));
            CK(cudaMemcpy2DAsync(buf[1], (size_t)ld * 8, buf[0], (size_t)ld * 8, (size_t)row_doubles * 8, (size_t)rows, cudaMemcpyDefault, st[0]));
            CK(cudaEventRecord(e1[0], st[0]));
            CK(cudaStreamSynchronize(st[0]));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0[0], e1[0]));
            if (ms < best) best = ms;
        }
        printf("REMOTE        copy engine cudaMemcpy2DAsync (1.5 MB rows)                         %8.2f ms  %7.1f GB/s\n", best, gb / (best * 1e-3));
    }
    return 0;
}
