// mma_peak.cu -- measured int8 tensor-core peak of this GPU, the denominator of bench.py's roofline.
//
// MEASURED_PEAKS.json carries a cuBLAS bf16 rate but no int8 entry, and a library GEMM is itself short of the pipe's
// peak.  This kernel issues the same instruction the Gram GEMM issues -- tcgen05.mma.cta_group::2.kind::i8, M = 256,
// N = 256, K = 32, operands in SWIZZLE_128B shared memory, accumulators in TMEM -- back to back from one thread per
// CTA pair, with no loads, no epilogue and no barriers between them: what comes out is the rate the tensor pipe
// sustains under this box's clocks and power cap with operands resident in shared memory.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gram_i8.h"
#include "kmg_common.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int PK_THREADS = 128;
constexpr int PK_STAGES = 4;
constexpr uint32_t PK_STAGE_BYTES = 2 * 128 * 128;  // A: 128 rows x 128 B, B: 128 rows x 128 B per CTA
constexpr uint32_t PK_SMEM = PK_STAGES * PK_STAGE_BYTES + 1024 + 64;

__global__ void __launch_bounds__(PK_THREADS, 1) mma_peak_i8_kernel(int iters, uint32_t seed) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* done = reinterpret_cast<uint64_t*>(smem + PK_STAGES * PK_STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    // operand bytes: small counts like a k-mer count row (the data pattern sets the switching power)
    for (uint32_t i = threadIdx.x; i < PK_STAGES * PK_STAGE_BYTES / 4; i += PK_THREADS) {
        uint32_t h = (i + seed + blockIdx.x * 7919u) * 2654435761u;
        reinterpret_cast<uint32_t*>(smem)[i] = (h >> 7) & 0x03010203u;
    }
    ptx::fence_proxy_async_smem();
    if (warp == 0 && lane == 0) {
        ptx::mbar_init(done, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_2cta(tmem_slot, 512);
        ptx::tmem_relinquish_2cta();
    }
    ptx::tcgen05_fence_before();
    ptx::cluster_sync_all();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (rank == 0 && warp == 0 && lane == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_i8(256, 256);
        for (int it = 0; it < iters; ++it) {
            const uint32_t a_addr = ptx::smem_u32(smem + (it & (PK_STAGES - 1)) * PK_STAGE_BYTES);
            const uint32_t b_addr = a_addr + 128 * 128;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t adesc = ptx::make_smem_desc_kmajor_sw128(a_addr + k4 * 32);
                const uint64_t bdesc = ptx::make_smem_desc_kmajor_sw128(b_addr + k4 * 32);
                ptx::umma_i8_2cta(tmem_base + ((it >> 2) & 1) * 256, adesc, bdesc, idesc, 1u);
            }
        }
        ptx::umma_commit_2cta(done, 3);
    }
    ptx::mbar_wait(done, 0);
    ptx::tcgen05_fence_after();
    ptx::tcgen05_fence_before();
    ptx::cluster_sync_all();
    if (warp == 1) ptx::tmem_dealloc_2cta(tmem_base, 512);
}

}  // namespace

// Launches `iters` x 4 MMAs (256 x 256 x 32 each) on every CTA pair; int8 operations issued = pairs * iters * 4 * 2*256*256*32.
int kmg_mma_peak_i8_launch(int iters, int64_t* ops, cudaStream_t stream) {
    KMG_REQUIRE(iters >= 1 && iters <= (1 << 24), KMG_ERR_ARG, "mma_peak: iters out of range");
    int dev = 0, sms = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    KMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static bool attr_set[64] = {};
    if (!attr_set[dev & 63]) {
        KMG_CUDA_CHECK(cudaFuncSetAttribute(mma_peak_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM));
        attr_set[dev & 63] = true;
    }
    const int pairs = sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(PK_THREADS);
    cfg.dynamicSmemBytes = PK_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    KMG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mma_peak_i8_kernel, iters, 12345u));
    if (ops) *ops = (int64_t)pairs * iters * 4 * 2ll * 256 * 256 * 32;
    return KMG_OK;
}
