// gram_i8_tcgen05.cu -- K = Phi_rows * Phi_cols^T  (int8 x int8 -> int32) on the 5th-gen tensor cores.
//
// Replaces the pair loop `K[i,j] = np.dot(phi_u[i], phi_u[j])` of the reference
// (kernels.py:41-45 spectrum, kernels.py:211-215 mismatch): the dense k-mer count matrix Phi
// (row-major n x Dpad int8, K contiguous) is contracted with itself.
//
// Design (sm_100a):
//   * persistent kernel, one CTA per SM, static round-robin over a host-built tile list (band
//     rasterised so that the 148 concurrent tiles share A/B panels in L2);
//   * warp 0 = TMA producer (cp.async.bulk.tensor, 128-byte swizzle, mbarrier complete_tx),
//     warp 1 = MMA issuer (one thread, tcgen05.mma.cta_group::1.kind::i8, 128 x 256 x 32 per
//     instruction, accumulators in TMEM), warp 2 = TMEM allocator, warps 4..11 = epilogue
//     (tcgen05.ld 32x32b.x32 -> int32 -> {s32 | f64 | cosine-normalised f64} -> global, plus the
//     mirrored store K[j,i] = K[i,j] for tiles strictly above the diagonal, kernels.py:45);
//   * CTA tile = (128*M_SUB) x 256, K-slab 128 int8 per stage. M_SUB=1: 4 smem stages and two TMEM
//     accumulator stages (epilogue overlaps the next tile's main loop; for small D, HBM-write
//     bound). M_SUB=2: 256 x 256 tile, 3 stages, whole TMEM is one accumulator (1/3 less operand
//     traffic from L2 per MAC; for large D, tensor bound).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "epi_ops.cuh"
#include "gram_i8.h"
#include "kmg_common.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int BN = 256;
constexpr int BK = 128;  // int8 elements = bytes = one 128-B swizzle row
constexpr int UMMA_K = 32;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + 32 * NUM_EPI_WARPS;
constexpr int TMEM_COLS = 512;
// Warp roles.  The epilogue warps take the LOW warp ids and the TMA producer / MMA issuer the HIGH ones: the SM
// sub-partition arbiter favours the highest warp id, and an MMA-issuing thread that loses issue slots to two busy
// epilogue warps on its sub-partition starves the tensor pipe (measured: a leaner epilogue made the large-D GEMM 3 %
// slower until the roles were swapped).  Epilogue warp w reads TMEM lanes 32*(w%4) .. +31 (hardware rule).
constexpr int WARP_TMA = NUM_EPI_WARPS;
constexpr int WARP_MMA = NUM_EPI_WARPS + 1;
constexpr int WARP_ALLOC = NUM_EPI_WARPS + 2;

struct KernelParams {
    int64_t rows, cols;
    int64_t row_index0, col_index0;
    int32_t kblocks;
    int32_t ntiles;
    int32_t out_dtype;  // KMG_OUT_S32 / KMG_OUT_F64
    void* out;
    int64_t ldo;
    // mirrored destinations: a tile with flag z stores element (r,c) (block-local indices) to mirror_base[w][c*ldo_t + r];
    // w = 0 for a single symmetric block; in a sharded symmetric build w is the part that owns the tile's columns and
    // mirror_base[w] its block-row buffer (peer device memory), pre-offset on the host so that local indices address it
    void* mirror_base[KMG_MAX_PARTS];
    int64_t ldo_t;
    const double* sd_rows;  // sqrt(diag) per local row / col for cosine normalisation; null = raw
    const double* sd_cols;
    const int4* tiles;  // {row_tile, col_tile, mirror, destination part of the mirror store}
    uint32_t* wave_counter;  // grid-wide arrival counter (zeroed before the launch) or null
    int32_t wave_wait_kb;    // K-slab of a tile before which the producer waits for the wave (0: at the tile start)
    uint64_t hint_a, hint_b;  // L2 eviction policy of the A / B operand loads
};

template <int M_SUB>
struct Cfg {
    static constexpr int BM = 128 * M_SUB;
    static constexpr int STAGES = (M_SUB == 1) ? 4 : 3;
    static constexpr int ACC_STAGES = (M_SUB == 1) ? 2 : 1;
    static constexpr int ACC_COLS = 256 * M_SUB;
    static constexpr uint32_t A_BYTES = BM * BK;
    static constexpr uint32_t B_BYTES = BN * BK;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t EPI_STAGE_WORDS = 32 * 32;  // per epilogue warp: 32 x 32 int32 transpose tile, XOR-swizzled
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + NUM_EPI_WARPS * EPI_STAGE_WORDS * 4;
};

// Epilogue store of one 32-row x 32-column accumulator chunk.  After tcgen05.ld thread t holds row t, so a
// direct store would touch 32 different cache lines per instruction (measured: the LSU wavefronts of that pattern
// cost ~17 us per 256x256 tile).  Instead the chunk is transposed through a per-warp 4 KB shared-memory tile
// (32 rows x 128 B, 16-byte chunks XOR-swizzled by row so that both the 128-bit row writes and the 64-bit
// transposed reads are conflict-free); a half-warp then stores 32 consecutive entries of one row, 16 B per lane.
// int32 -> fp64 uses the 2^52 magic-number add (one DADD) instead of the quarter-rate I2F.F64; accumulators are
// non-negative counts.  The mirrored store K[c][r] needs no transpose: for a fixed register j the 32 lanes are
// 32 consecutive rows.
// Exact int -> double with integer instructions only (FLO + shifts): while the tensor pipe is saturated by the next
// tile's UTCIMMA stream every FP64-pipe instruction of the epilogue (I2F.F64, or the 2^52 magic-number DADD) sits in
// `stall_math` (ncu: a third of the epilogue warps' samples), integer ALU ops do not.  0 <= v < 2^31.
// Where the epilogue runs in series with the main loop (tensor pipe idle) the one-instruction 2^52 magic-number DADD is
// the cheaper conversion; INT_CVT selects per kernel.
template <bool INT_CVT>
__device__ __forceinline__ double u32_to_f64(uint32_t v) {
    if (!INT_CVT) return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
    const int lz = __clz((int)v);                    // 32 for v == 0
    const uint32_t w = (v << (lz & 31)) << 1;        // fraction bits left-aligned, leading one shifted out
    const uint32_t hi = v ? (((1023u + 31u - (uint32_t)lz) << 20) | (w >> 12)) : 0u;
    return __hiloint2double((int)hi, (int)(w << 20));
}

template <bool INT_CVT, bool FUSED = false>
__device__ __forceinline__ void store_chunk(const KernelParams& p, const uint32_t (&v)[32], uint32_t* stage /*4 KB*/,
                                            int lane, int64_t row_base, int64_t col0, void* out_t, const EpiOps* ep = nullptr) {
    const bool mirror = out_t != nullptr;
    const int64_t ncol = (p.cols - col0 < 32) ? (p.cols - col0) : 32;  // warp-uniform, > 0
    int64_t nrow = p.rows - row_base;
    if (nrow > 32) nrow = 32;
    if (nrow <= 0) return;  // warp-uniform
    if (FUSED && (ep->row_sum_partial != nullptr || ep->row_wsum_partial != nullptr)) {
        // Row statistics of the chunk straight from the accumulators: after tcgen05.ld thread t holds row t, so the sums
        // over the chunk's 32 columns need no communication.  Values as they are stored (normalised if sd is given).
        const int64_t row = row_base + lane;
        if (row < p.rows) {
            const double sr = p.sd_rows != nullptr ? p.sd_rows[row] : 1.0;
            double s = 0.0, w = 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < ncol) {
                    double d = u32_to_f64<INT_CVT>(v[j]);
                    if (p.sd_rows != nullptr) {
                        d = __ddiv_rn(d, __dmul_rn(sr, p.sd_cols[col0 + j]));
                        if (p.row_index0 + row == p.col_index0 + col0 + j) d = 1.0;
                    }
                    s += d;
                    if (ep->row_wsum_partial != nullptr) w += __dmul_rn(d, ep->w_cols[col0 + j]);
                }
            }
            if (ep->row_sum_partial != nullptr) ep->row_sum_partial[row * ep->n_chunks + (col0 >> 5)] = s;
            if (ep->row_wsum_partial != nullptr) ep->row_wsum_partial[row * ep->n_chunks + (col0 >> 5)] = w;
        }
    }
    {
        uint4* srow = reinterpret_cast<uint4*>(stage + lane * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) srow[c ^ (lane & 7)] = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    __syncwarp();
    const int half = lane >> 4, l16 = lane & 15;
    const int64_t col = col0 + 2 * l16;  // this lane stores columns col, col+1
    const int ok = (2 * l16 + 1 < ncol) ? 2 : ((2 * l16 < ncol) ? 1 : 0);
    const bool norm = p.sd_rows != nullptr;
    // Fast path for interior chunks (all but the last tile row / column): no per-element bounds or alignment tests and
    // one 64-bit pointer increment per pair of rows -- the general path below spends ~20 instructions per element on
    // index arithmetic and predicates, which is what bounds the write-bound small-D regime (ncu: issue slots spread
    // over the epilogue, no single stall).
    // Only in the kernels whose epilogue is on the critical path (INT_CVT == false: small D, or the serial-epilogue
    // variants): under a long main loop the leaner, burstier store stream measurably slows the MMA pipeline
    // (25k x 200k block-row: 54.3 ms with the general path, 56.5 ms with this one), and the epilogue is hidden anyway.
    if (!FUSED && !INT_CVT && nrow == 32 && ncol == 32 && !norm && (p.ldo & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
        const uint32_t* src = stage + half * 32 + ((l16 & 1) << 1);
        const int cl = l16 >> 1;
        if (p.out_dtype == KMG_OUT_S32) {
            int32_t* dst = reinterpret_cast<int32_t*>(p.out) + (row_base + half) * p.ldo + col;
            const int64_t step = 2 * p.ldo;
#pragma unroll 4
            for (int r2 = 0; r2 < 16; ++r2) {
                const int rr = 2 * r2 + half;
                const uint2 iv = *reinterpret_cast<const uint2*>(src + r2 * 64 + ((cl ^ (rr & 7)) << 2));
                *reinterpret_cast<int2*>(dst) = make_int2((int)iv.x, (int)iv.y);
                dst += step;
            }
        } else {
            double* dst = reinterpret_cast<double*>(p.out) + (row_base + half) * p.ldo + col;
            const int64_t step = 2 * p.ldo;
#pragma unroll 4
            for (int r2 = 0; r2 < 16; ++r2) {
                const int rr = 2 * r2 + half;
                const uint2 iv = *reinterpret_cast<const uint2*>(src + r2 * 64 + ((cl ^ (rr & 7)) << 2));
                *reinterpret_cast<double2*>(dst) = make_double2(u32_to_f64<INT_CVT>(iv.x), u32_to_f64<INT_CVT>(iv.y));
                dst += step;
            }
        }
        if (mirror) {
            const int64_t row_t = row_base + lane;
            if (p.out_dtype == KMG_OUT_S32) {
                int32_t* dt = reinterpret_cast<int32_t*>(out_t) + col0 * p.ldo_t + row_t;
#pragma unroll
                for (int j = 0; j < 32; ++j) { *dt = (int32_t)v[j]; dt += p.ldo_t; }
            } else {
                double* dt = reinterpret_cast<double*>(out_t) + col0 * p.ldo_t + row_t;
#pragma unroll
                for (int j = 0; j < 32; ++j) { *dt = u32_to_f64<INT_CVT>(v[j]); dt += p.ldo_t; }
            }
        }
        __syncwarp();
        return;
    }
    if (p.out_dtype == KMG_OUT_S32) {
        int32_t* base = reinterpret_cast<int32_t*>(p.out);
        const bool vec = ((p.ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(base) & 7) == 0);
#pragma unroll 4
        for (int r2 = 0; r2 < 16; ++r2) {
            const int rr = 2 * r2 + half;
            const uint2 iv = *reinterpret_cast<const uint2*>(stage + rr * 32 + (((l16 >> 1) ^ (rr & 7)) << 2) + ((l16 & 1) << 1));
            if (rr < nrow && ok) {
                int32_t* dst = base + (row_base + rr) * p.ldo + col;
                if (ok == 2 && vec) *reinterpret_cast<int2*>(dst) = make_int2((int)iv.x, (int)iv.y);
                else { dst[0] = (int)iv.x; if (ok == 2) dst[1] = (int)iv.y; }
            }
        }
        const int64_t row_t = row_base + lane;
        if (mirror && row_t < p.rows) {
            int32_t* dt = reinterpret_cast<int32_t*>(out_t) + col0 * p.ldo_t + row_t;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncol) dt[(int64_t)j * p.ldo_t] = (int32_t)v[j];
        }
    } else {
        double* base = reinterpret_cast<double*>(p.out);
        const bool vec = ((p.ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
        double sc0 = 1.0, sc1 = 1.0;
        if (norm) {
            if (ok >= 1) sc0 = p.sd_cols[col];
            if (ok == 2) sc1 = p.sd_cols[col + 1];
        }
#pragma unroll 4
        for (int r2 = 0; r2 < 16; ++r2) {
            const int rr = 2 * r2 + half;
            const uint2 iv = *reinterpret_cast<const uint2*>(stage + rr * 32 + (((l16 >> 1) ^ (rr & 7)) << 2) + ((l16 & 1) << 1));
            if (rr < nrow && ok) {
                double d0 = u32_to_f64<INT_CVT>(iv.x), d1 = u32_to_f64<INT_CVT>(iv.y);
                if (norm) {
                    // normalize_K (kernels.py:408-414): K_ij / (sqrt(K_ii) * sqrt(K_jj)), diagonal := 1.0
                    const double sr = p.sd_rows[row_base + rr];
                    const int64_t grow = p.row_index0 + row_base + rr, gcol = p.col_index0 + col;
                    d0 = __ddiv_rn(d0, __dmul_rn(sr, sc0));
                    d1 = __ddiv_rn(d1, __dmul_rn(sr, sc1));
                    if (grow == gcol) d0 = 1.0;
                    if (grow == gcol + 1) d1 = 1.0;
                }
                double* dst = base + (row_base + rr) * p.ldo + col;
                if (FUSED) {  // sum_m u_m K_m, power, normalisation of the combination (epi_ops.cuh)
                    const int64_t grow = p.row_index0 + row_base + rr, gcol = p.col_index0 + col;
                    const bool nrm = ep->post_sd_rows != nullptr;
                    const double psr = nrm ? ep->post_sd_rows[row_base + rr] : 1.0;
                    const double prev0 = ep->accumulate == 2 ? dst[0] : 0.0;
                    d0 = epi_finish(*ep, d0, prev0, grow == gcol, psr, nrm ? ep->post_sd_cols[col] : 1.0);
                    if (ok == 2) {
                        const double prev1 = ep->accumulate == 2 ? dst[1] : 0.0;
                        d1 = epi_finish(*ep, d1, prev1, grow == gcol + 1, psr, nrm ? ep->post_sd_cols[col + 1] : 1.0);
                    }
                }
                if (ok == 2 && vec) *reinterpret_cast<double2*>(dst) = make_double2(d0, d1);
                else { dst[0] = d0; if (ok == 2) dst[1] = d1; }
            }
        }
        const int64_t row_t = row_base + lane;
        if (mirror && nrow == 32 && ncol == 32 && !norm && (p.ldo_t & 1) == 0 && (reinterpret_cast<uintptr_t>(out_t) & 15) == 0) {
            // Full chunk: neighbouring lanes swap one value per column pair so that every lane stores 16 bytes (rows
            // L, L+1 of one mirrored row): half the store instructions of the scalar path below, 512 B per instruction.
            // What bounds mirror stores into PEER memory is the number of store instructions, not the bytes
            // (2 GPUs, n = 100 000: 40.4 ms with scalar stores, fp64 and s32 output alike).
            const int odd = lane & 1;
            double* dt = reinterpret_cast<double*>(out_t) + (col0 + odd) * p.ldo_t + (row_base + lane - odd);
            const int64_t step = 2 * p.ldo_t;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const uint32_t t = __shfl_xor_sync(0xffffffffu, odd ? v[j] : v[j + 1], 1);
                const uint32_t a = odd ? t : v[j], b = odd ? v[j + 1] : t;  // rows (L - odd, L - odd + 1) of column j + odd
                *reinterpret_cast<double2*>(dt) = make_double2(u32_to_f64<INT_CVT>(a), u32_to_f64<INT_CVT>(b));
                dt += step;
            }
        } else if (mirror && row_t < p.rows) {
            double* dt = reinterpret_cast<double*>(out_t) + col0 * p.ldo_t + row_t;
            const double sr = norm ? p.sd_rows[row_t] : 1.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < ncol) {
                    double val = u32_to_f64<INT_CVT>(v[j]);
                    if (norm) val = __ddiv_rn(val, __dmul_rn(sr, p.sd_cols[col0 + j]));  // mirrored tiles never contain the diagonal
                    dt[(int64_t)j * p.ldo_t] = val;
                }
            }
        }
    }
    __syncwarp();  // the staging tile is reused by the next chunk
}

// Mirror store of one 32 x 32 chunk through the TMA engine: the warp writes the chunk TRANSPOSED (and converted) into its
// 4 KB staging tile -- for a fixed column the 32 lanes are 32 consecutive rows, i.e. 32 consecutive entries of one
// mirrored row: conflict-free 8-byte shared stores -- and one lane issues cp.async.bulk.tensor stores of
// [16 mirrored rows] x [32 entries] (fp64) boxes into the destination's tensor map.  Nothing of the mirror goes through
// the LSU as a global store: the destination may be the block-row buffer of ANOTHER GPU (CUDA IPC mapping, NVLink), where
// thread-issued stores capped the kernel (2 GPUs, n = 100 000: 40.5 ms against 30.1 ms with local buffers).  Rows /
// columns beyond the block are clipped by the tensor map.  The staging tile is shared with the own-row store, so the
// caller waits for the previous bulk group to have been read before it writes the tile again.
template <bool INT_CVT>
__device__ __forceinline__ void tma_mirror_chunk(const KernelParams& p, const uint32_t (&v)[32], uint32_t* stage /*4 KB*/, int lane,
                                                 int64_t row_base, int64_t col0, const CUtensorMap* tmT) {
    if (row_base >= p.rows) return;  // warp-uniform
    if (p.out_dtype == KMG_OUT_S32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[j * 32 + lane] = v[j];
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            ptx::tma_store_2d(tmT, stage, (int32_t)row_base, (int32_t)col0);
            ptx::tma_store_commit();
        }
        return;
    }
    double* sdst = reinterpret_cast<double*>(stage);
    const bool norm = p.sd_rows != nullptr;
    const int64_t row_t = row_base + lane;
    const double sr = (norm && row_t < p.rows) ? p.sd_rows[row_t] : 1.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (col0 + 16 * h >= p.cols) break;  // warp-uniform
        if (h == 1) {
            if (lane == 0) ptx::tma_store_wait_read<0>();
            __syncwarp();
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * h + jj;
            double val = u32_to_f64<INT_CVT>(v[j]);
            if (norm) {  // mirrored tiles never contain the diagonal (kernels.py:408-414)
                const double sc = (col0 + j < p.cols) ? p.sd_cols[col0 + j] : 1.0;
                val = __ddiv_rn(val, __dmul_rn(sr, sc));
            }
            sdst[jj * 32 + lane] = val;
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            ptx::tma_store_2d(tmT, stage, (int32_t)row_base, (int32_t)(col0 + 16 * h));
            ptx::tma_store_commit();
        }
    }
}

template <int M_SUB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_i8_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const KernelParams p) {
    using C = Cfg<M_SUB>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-B alignment by offsetting the __shared__ symbol (keeps the address space known to the compiler: LDS/STS, not generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tfull = empty + C::STAGES;
    uint64_t* tempty = tfull + C::ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ACC_STAGES);
    uint32_t* epi_stage = reinterpret_cast<uint32_t*>(smem + C::STAGES * C::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == WARP_TMA && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
    }
    if (warp == WARP_MMA && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < C::ACC_STAGES; ++a) {
            ptx::mbar_init(&tfull[a], 1);
            ptx::mbar_init(&tempty[a], NUM_EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == WARP_ALLOC) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == WARP_TMA) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, wave = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++wave) {
                const int4 tile = p.tiles[t];
                const int32_t row0 = tile.x * C::BM, col0 = tile.y * BN;
                if (p.wave_counter != nullptr) {
                    // Wave barrier: all CTAs start their w-th tile together.  The 148 tiles of a wave share A/B
                    // panels (band rasterisation); without this the CTAs drift apart over hundreds of waves, the
                    // k-slabs they share leave L2 before the partner arrives and every panel is fetched from HBM
                    // by each CTA separately (measured: 275 GB of DRAM reads per launch instead of ~80 GB).
                    ptx::red_release_gpu_add(p.wave_counter, 1u);
                    const uint32_t done = (wave + 1) * gridDim.x;
                    const uint32_t target = done < (uint32_t)p.ntiles ? done : (uint32_t)p.ntiles;
                    if (ptx::ld_acquire_gpu(p.wave_counter) < target) {
                        const uint64_t t0 = ptx::globaltimer_ns();
                        while (ptx::ld_acquire_gpu(p.wave_counter) < target) {
                            __nanosleep(200);
                            if (ptx::globaltimer_ns() - t0 > 4000000000ull) __trap();
                        }
                    }
                }
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    ptx::tma_load_2d_hint(sa, &tmA, &full[stage], kb * BK, row0, p.hint_a);
                    ptx::tma_load_2d_hint(sa + C::A_BYTES, &tmB, &full[stage], kb * BK, col0, p.hint_b);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_i8(128, BN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
                ptx::tcgen05_fence_after();
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tcgen05_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
                    for (int ms = 0; ms < M_SUB; ++ms) {
#pragma unroll
                        for (int k4 = 0; k4 < BK / UMMA_K; ++k4) {
                            const uint64_t adesc = ptx::make_smem_desc_kmajor_sw128(a_addr + ms * 128 * BK + k4 * UMMA_K);
                            const uint64_t bdesc = ptx::make_smem_desc_kmajor_sw128(b_addr + k4 * UMMA_K);
                            ptx::umma_i8(tmem_base + acc * C::ACC_COLS + ms * 256, adesc, bdesc, idesc,
                                         (kb | k4) != 0 ? 1u : 0u);
                        }
                    }
                    ptx::umma_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tfull[acc]);  // accumulator ready for the epilogue
                if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp < NUM_EPI_WARPS) {
        // ------------------------------------------------------------ epilogue
        const int quarter = warp & 3;        // TMEM lane quarter this warp may access
        const int half = warp >> 2;    // which 128 of the 256 accumulator columns
        uint32_t acc = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const int4 tile = p.tiles[t];
            void* const mirror = tile.z != 0 ? p.mirror_base[tile.w & (KMG_MAX_PARTS - 1)] : nullptr;
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tcgen05_fence_after();
#pragma unroll
            for (int ms = 0; ms < M_SUB; ++ms) {
                const int64_t row_base = (int64_t)tile.x * C::BM + ms * 128 + quarter * 32;
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    const int64_t col0 = (int64_t)tile.y * BN + half * 128 + ch * 32;
                    if (col0 >= p.cols) break;  // warp-uniform
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                           acc * C::ACC_COLS + ms * 256 + half * 128 + ch * 32;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(taddr, v);
                    ptx::tmem_ld_wait();
                    store_chunk<false>(p, v, epi_stage + warp * C::EPI_STAGE_WORDS, lane, row_base, col0, mirror);
                }
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
            if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == WARP_ALLOC) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two SMs of one TPC compute one 256 x 256 tile.  Each CTA holds 128 rows of A and
// 128 of the 256 rows of B per stage (32 KB, so 6 stages fit) and 128 x 256 accumulators in its TMEM -- two
// accumulator stages fit, so the epilogue (bounded by the ~64 B/clk/SM store path: 256 KB per CTA and tile) runs
// under the next tile's main loop instead of in series with it.  The leader CTA's single MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256); both CTAs' TMA loads credit the leader's `full` barrier; the MMA completion
// is multicast to both CTAs' `empty` / `tmem_full` barriers; both CTAs' epilogue warps arrive on the leader's
// `tmem_empty` barrier.
// ---------------------------------------------------------------------------------------------
struct Cfg2 {
    static constexpr int BM = 256;        // pair tile rows (128 per CTA)
    static constexpr int STAGES = 6;
    static constexpr int ACC_STAGES = 2;
    static constexpr int ACC_COLS = 256;
    static constexpr uint32_t A_BYTES = 128 * BK;
    static constexpr uint32_t B_BYTES = 128 * BK;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t EPI_STAGE_WORDS = 32 * 32;
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + NUM_EPI_WARPS * EPI_STAGE_WORDS * 4;
};

template <bool INT_CVT, bool TMA_MIRROR, bool FUSED = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_i8_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmT, const KernelParams p, const __grid_constant__ EpiOps epi) {
    using C = Cfg2;
    extern __shared__ uint8_t smem_raw[];
    // 1024-B alignment by offsetting the __shared__ symbol (keeps the address space known to the compiler: LDS/STS, not generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tfull = empty + C::STAGES;
    uint64_t* tempty = tfull + C::ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ACC_STAGES);
    uint32_t* epi_stage = reinterpret_cast<uint32_t*>(smem + C::STAGES * C::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == WARP_TMA && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        if (TMA_MIRROR) ptx::prefetch_tensormap(&tmT);
    }
    if (warp == WARP_MMA && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);    // leader: its own arrive.expect_tx for the bytes of both CTAs
            ptx::mbar_init(&empty[s], 1);   // multicast tcgen05.commit
        }
        for (int a = 0; a < C::ACC_STAGES; ++a) {
            ptx::mbar_init(&tfull[a], 1);                    // multicast tcgen05.commit
            ptx::mbar_init(&tempty[a], 2 * NUM_EPI_WARPS);   // leader: epilogue warps of both CTAs
        }
        ptx::fence_barrier_init();
    }
    if (warp == WARP_ALLOC) {
        ptx::tmem_alloc_2cta(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish_2cta();
    }
    ptx::tcgen05_fence_before();
    ptx::cluster_sync_all();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == WARP_TMA) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, wave = 0;
            for (int t = cluster_id; t < p.ntiles; t += num_clusters, ++wave) {
                const int4 tile = p.tiles[t];
                const int32_t row0 = tile.x * C::BM + (int)rank * 128, col0 = tile.y * BN + (int)rank * 128;
                // Wave barrier (see gram_i8_tcgen05_kernel): all CTAs work on their w-th tile together.  A CTA announces its
                // arrival at the tile start but waits for the others only before its (STAGES+1)-th K-slab: the slabs that
                // fit the operand pipeline are loaded right away, so the pipeline does not drain at every tile boundary
                // while the slowest CTA of the wave finishes (the early slabs of a wave's panels stay in L2 for the
                // microseconds the stragglers need).
                if (p.wave_counter != nullptr) ptx::red_release_gpu_add(p.wave_counter, 1u);
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    if (p.wave_counter != nullptr && kb == p.wave_wait_kb) {
                        const uint32_t done = (wave + 1) * gridDim.x;
                        const uint32_t target = done < 2u * (uint32_t)p.ntiles ? done : 2u * (uint32_t)p.ntiles;
                        if (ptx::ld_acquire_gpu(p.wave_counter) < target) {
                            const uint64_t t0 = ptx::globaltimer_ns();
                            while (ptx::ld_acquire_gpu(p.wave_counter) < target) {
                                __nanosleep(200);
                                if (ptx::globaltimer_ns() - t0 > 4000000000ull) __trap();
                            }
                        }
                    }
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    ptx::tma_load_2d_2cta(sa, &tmA, &full[stage], kb * BK, row0, p.hint_a);
                    ptx::tma_load_2d_2cta(sa + C::A_BYTES, &tmB, &full[stage], kb * BK, col0, p.hint_b);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_i8(256, BN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int t = cluster_id; t < p.ntiles; t += num_clusters) {
                ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
                ptx::tcgen05_fence_after();
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tcgen05_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < BK / UMMA_K; ++k4) {
                        const uint64_t adesc = ptx::make_smem_desc_kmajor_sw128(a_addr + k4 * UMMA_K);
                        const uint64_t bdesc = ptx::make_smem_desc_kmajor_sw128(b_addr + k4 * UMMA_K);
                        ptx::umma_i8_2cta(tmem_base + acc * C::ACC_COLS, adesc, bdesc, idesc, (kb | k4) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit_2cta(&empty[stage], 3);  // both CTAs may refill this stage
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit_2cta(&tfull[acc], 3);  // both CTAs' epilogues
                if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp < NUM_EPI_WARPS) {
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        const int quarter = warp & 3;
        const int half = warp >> 2;
        uint32_t acc = 0, acc_phase = 0;
        for (int t = cluster_id; t < p.ntiles; t += num_clusters) {
            const int4 tile = p.tiles[t];
            void* const mirror = tile.z != 0 ? p.mirror_base[tile.w & (KMG_MAX_PARTS - 1)] : nullptr;
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tcgen05_fence_after();
            const int64_t row_base = (int64_t)tile.x * C::BM + (int64_t)rank * 128 + quarter * 32;
            // (software-pipelining the tcgen05.ld of chunk c+1 under the stores of chunk c was measured: no gain, and the
            // unrolled copies of the store code blew the kernel up to 250 KB of SASS, which cost 3 % on the large-D shapes)
            uint32_t* st = epi_stage + warp * C::EPI_STAGE_WORDS;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                const int64_t col0 = (int64_t)tile.y * BN + half * 128 + ch * 32;
                if (col0 >= p.cols) break;  // warp-uniform
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * C::ACC_COLS + half * 128 + ch * 32;
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(taddr, v);
                ptx::tmem_ld_wait();
                if (TMA_MIRROR) {
                    // the staging tile may still be the source of the previous chunk's bulk store
                    if (lane == 0) ptx::tma_store_wait_read<0>();
                    __syncwarp();
                    store_chunk<INT_CVT>(p, v, st, lane, row_base, col0, nullptr);
                    if (mirror != nullptr) tma_mirror_chunk<INT_CVT>(p, v, st, lane, row_base, col0, &tmT);
                } else {
                    store_chunk<INT_CVT, FUSED>(p, v, st, lane, row_base, col0, mirror, &epi);
                }
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(&tempty[acc], 0);  // the leader's MMA thread waits on it
            if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        if (TMA_MIRROR && lane == 0) ptx::tma_store_wait<0>();  // every mirror store has landed before the CTA retires
    }
    ptx::tcgen05_fence_before();
    ptx::cluster_sync_all();
    if (warp == WARP_ALLOC) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc_2cta(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// Validation kernel (SIMT, dp4a).  Not on the product path: it exists so the tests can check the
// tcgen05 kernel against an independent on-device evaluation at sizes the CPU oracle cannot reach.
// ---------------------------------------------------------------------------------------------
__global__ void gram_i8_simt_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int64_t ld,
                                    int64_t rows, int64_t cols, int64_t Dpad, int32_t* __restrict__ out, int64_t ldo) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c >= cols || r >= rows) return;
    const int* a = reinterpret_cast<const int*>(A + r * ld);
    const int* b = reinterpret_cast<const int*>(B + c * ld);
    int acc = 0;
    for (int64_t t = 0; t < Dpad / 4; ++t) acc = __dp4a(__ldg(a + t), __ldg(b + t), acc);
    out[r * ldo + c] = acc;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(f);
    });
    return fn;
}

int make_map(CUtensorMap* m, const int8_t* base, int64_t nrows, int64_t Dpad, int64_t ld, int box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    KMG_REQUIRE(enc != nullptr, KMG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)Dpad, (cuuint64_t)nrows};
    cuuint64_t strides[1] = {(cuuint64_t)ld};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KMG_REQUIRE(r == CUDA_SUCCESS, KMG_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return KMG_OK;
}

// Tile list, band-rasterised: bands of G row-tiles, inside a band column-major over the band's rows.
// symmetric: only tiles that intersect {col >= row} (global indices); mirror flag for tiles whose
// transposed image is not covered by a computed tile.
struct TileKey {
    int64_t rows, cols, r0, c0;
    int bm, sym, band;
    int n_parts, part;
    std::array<int64_t, KMG_MAX_PARTS + 1> bounds;  // part_row0 of a sharded build, zeros otherwise
    int mirror_all;
    int interleave;
    bool operator<(const TileKey& o) const {
        return std::tie(rows, cols, r0, c0, bm, sym, band, n_parts, part, bounds, mirror_all, interleave) <
               std::tie(o.rows, o.cols, o.r0, o.c0, o.bm, o.sym, o.band, o.n_parts, o.part, o.bounds, o.mirror_all, o.interleave);
    }
};

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// rotating pool of zero-initialised arrival counters, one per launch in flight
constexpr int COUNTER_SLOTS = 1024;
uint32_t* g_counters[64] = {};
unsigned g_counter_next[64] = {};
std::mutex g_counter_mu;
int get_counter(cudaStream_t stream, uint32_t** out) {
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    dev &= 63;
    std::lock_guard<std::mutex> lk(g_counter_mu);  // host entry points may run from several threads
    if (g_counters[dev] == nullptr) KMG_CUDA_CHECK(cudaMalloc(&g_counters[dev], COUNTER_SLOTS * sizeof(uint32_t)));
    uint32_t* c = g_counters[dev] + (g_counter_next[dev]++ % COUNTER_SLOTS);
    KMG_CUDA_CHECK(cudaMemsetAsync(c, 0, sizeof(uint32_t), stream));
    *out = c;
    return KMG_OK;
}
struct TileList {
    int4* dev = nullptr;
    int32_t n = 0;
    int64_t computed_entries = 0;
    uint64_t last_use = 0;
};
// Tile lists are cached per (device, shape): a long-lived process that sees many shapes keeps at most TILE_CACHE_MAX of
// them (least recently used goes first; cudaFree waits for the kernels that may still read it) and kmg_release() frees all.
constexpr size_t TILE_CACHE_MAX = 48;
std::mutex g_tile_mu;
std::map<std::pair<int, TileKey>, TileList> g_tile_cache;  // per device
uint64_t g_tile_clock = 0;

int get_tiles(const TileKey& key, cudaStream_t stream, TileList* out) {
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tile_mu);
    auto it = g_tile_cache.find({dev, key});
    if (it != g_tile_cache.end()) { it->second.last_use = ++g_tile_clock; *out = it->second; return KMG_OK; }
    while (g_tile_cache.size() >= TILE_CACHE_MAX) {
        auto victim = g_tile_cache.begin();
        for (auto c = g_tile_cache.begin(); c != g_tile_cache.end(); ++c)
            if (c->second.last_use < victim->second.last_use) victim = c;
        if (victim->second.dev) {
            int cur = dev;
            if (victim->first.first != cur) cudaSetDevice(victim->first.first);
            cudaFree(victim->second.dev);
            if (victim->first.first != cur) cudaSetDevice(cur);
        }
        g_tile_cache.erase(victim);
    }
    const int BM = key.bm;
    const int64_t tm_n = (key.rows + BM - 1) / BM, tn_n = (key.cols + BN - 1) / BN;
    const int G = key.band;
    std::vector<int4> tiles;
    tiles.reserve((size_t)(tm_n * tn_n));
    int64_t entries = 0;
    // Column order inside a band.  Sharded single-launch build: the column tiles of the parts are interleaved (position
    // p of part 0, of part 1, ...), so that tiles whose mirror goes over NVLink alternate with tiles of the local
    // diagonal block and the link sees the AVERAGE demand of the launch instead of phases at 100 % / 0 %.
    std::vector<int64_t> col_order;
    col_order.reserve((size_t)tn_n);
    if (key.n_parts > 1 && key.interleave) {
        std::vector<int64_t> first(key.n_parts + 1);
        for (int q = 0; q <= key.n_parts; ++q) first[q] = (key.bounds[q] + BN - 1) / BN;  // boundaries are multiples of 256
        first[key.n_parts] = tn_n;
        int64_t longest = 0;
        for (int q = 0; q < key.n_parts; ++q) longest = std::max(longest, first[q + 1] - first[q]);
        for (int64_t pos = 0; pos < longest; ++pos)
            for (int q = 0; q < key.n_parts; ++q)
                if (first[q] + pos < first[q + 1]) col_order.push_back(first[q] + pos);
    } else {
        for (int64_t tn = 0; tn < tn_n; ++tn) col_order.push_back(tn);
    }
    for (int64_t b0 = 0; b0 < tm_n; b0 += G) {
        const int64_t b1 = (b0 + G < tm_n) ? b0 + G : tm_n;
        for (int64_t tn : col_order) {
            for (int64_t tm = b0; tm < b1; ++tm) {
                int mirror = key.mirror_all, dest = 0;
                if (key.n_parts > 0) {
                    // sharded symmetric build: global 256-grid tile (I, J); b = the part owning columns of J
                    const int64_t I = key.bounds[key.part] / BM + tm, J = tn;  // the part's first row tile on the global grid
                    int b = 0;
                    while (b + 1 < key.n_parts && J * BN >= key.bounds[b + 1]) ++b;
                    if (!kmg_gram_sharded_takes(key.n_parts, key.bounds.data(), key.part, b, I, J)) continue;
                    mirror = (I != J) ? 1 : 0;
                    dest = b;
                } else if (key.sym) {
                    // global index ranges of this tile
                    const int64_t rlo = key.r0 + tm * BM, rhi = rlo + BM - 1;
                    const int64_t clo = key.c0 + tn * BN, chi = clo + BN - 1;
                    if (chi < rlo) continue;        // entirely below the diagonal: produced by a mirror store
                    mirror = (clo > rhi) ? 1 : 0;   // entirely above: its transpose is nobody's tile
                }
                tiles.push_back(make_int4((int)tm, (int)tn, mirror, dest));
                const int64_t rr = (key.rows - tm * BM < BM) ? key.rows - tm * BM : BM;
                const int64_t cc = (key.cols - tn * BN < BN) ? key.cols - tn * BN : BN;
                entries += rr * cc;
            }
        }
    }
    TileList tl;
    tl.n = (int32_t)tiles.size();
    tl.computed_entries = entries;
    if (tl.n > 0) {
        KMG_CUDA_CHECK(cudaMalloc(&tl.dev, sizeof(int4) * tiles.size()));
        // pageable source: the call returns once the list sits in the driver's staging buffer, so `tiles` may go out of
        // scope and the launch that follows on `stream` is ordered after the copy -- no stream synchronisation needed
        KMG_CUDA_CHECK(cudaMemcpyAsync(tl.dev, tiles.data(), sizeof(int4) * tiles.size(), cudaMemcpyHostToDevice, stream));
    }
    tl.last_use = ++g_tile_clock;
    g_tile_cache[{dev, key}] = tl;
    *out = tl;
    return KMG_OK;
}

// KMG_GEMM_COOP=0 drops the cooperative attribute: Nsight Compute refuses cooperative cluster launches ("LaunchFailed"),
// and under the profiler kernels are serialised, so every CTA is resident anyway.  Not for production use.
bool coop_enabled() {
    static const int v = env_int("KMG_GEMM_COOP", 1);
    return v != 0;
}

template <int M_SUB>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const KernelParams& p, int sms, cudaStream_t stream) {
    using C = Cfg<M_SUB>;
    static bool attr_set[64] = {};
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        KMG_CUDA_CHECK(cudaFuncSetAttribute(gram_i8_tcgen05_kernel<M_SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)C::SMEM_BYTES));
        attr_set[dev & 63] = true;
    }
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    // the wave barrier needs every CTA resident: a cooperative launch makes the driver guarantee it (or fail the launch)
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = (p.wave_counter != nullptr && coop_enabled()) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    KMG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gram_i8_tcgen05_kernel<M_SUB>, tmA, tmB, p));
    return KMG_OK;
}

template <bool INT_CVT, bool TMA_MIRROR, bool FUSED = false>
int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmT, const KernelParams& p, int sms, cudaStream_t stream,
                const EpiOps* epi = nullptr) {
    EpiOps e;
    memset(&e, 0, sizeof(e));
    if (epi != nullptr) e = *epi;
    static bool attr_set[64] = {};
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        KMG_CUDA_CHECK(cudaFuncSetAttribute(gram_i8_2cta_kernel<INT_CVT, TMA_MIRROR, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg2::SMEM_BYTES));
        attr_set[dev & 63] = true;
    }
    int clusters = sms / 2;
    if (p.ntiles < clusters) clusters = p.ntiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg2::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    // the wave barrier needs every CTA resident: a cooperative launch makes the driver guarantee it (or fail the launch)
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = (p.wave_counter != nullptr && coop_enabled()) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    KMG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gram_i8_2cta_kernel<INT_CVT, TMA_MIRROR, FUSED>, tmA, tmB, tmT, p, e));
    return KMG_OK;
}

// Tensor map over a MIRROR destination: element (r, c) of the block (block-local indices) lives at base[c * ld + r], so
// the inner (contiguous) coordinate is the block ROW and the outer one the block COLUMN; extents = the block shape, which
// makes the TMA engine clip the ragged last tile row / column.  Box = 32 entries x (16 fp64 | 32 s32) mirrored rows = 4 KB.
int make_mirror_map(CUtensorMap* m, void* base, int out_dtype, int64_t rows, int64_t cols, int64_t ld) {
    PFN_encodeTiled enc = get_encode_fn();
    KMG_REQUIRE(enc != nullptr, KMG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const bool f64 = out_dtype == KMG_OUT_F64;
    const cuuint64_t esz = f64 ? 8 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {32u, f64 ? 16u : 32u};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KMG_REQUIRE(r == CUDA_SUCCESS, KMG_ERR_CUDA, "cuTensorMapEncodeTiled (mirror destination) failed (CUresult %d)", (int)r);
    return KMG_OK;
}

}  // namespace

// Which part computes tile (I, J), I != J, of the symmetric Gram when its rows belong to part a and its columns to part
// b: exactly one of (I, J) [by a] and (J, I) [by b] is computed, the other arrives as the mirror store.
//   a == b            : the upper triangle, J >= I (as in the single-GPU symmetric build)
//   cyclic distance d = (b - a) mod g:  2d < g -> a computes the whole block (a, b);  2d > g -> b does
//   2d == g (g even)  : the block is split between the two parts by row tile so that every part computes g/2 blocks:
//                       the lower-numbered part takes its first half of row tiles, the other one the transposes of the rest
int kmg_gram_sharded_takes(int g, const int64_t* part_row0, int a, int b, int64_t I, int64_t J) {
    if (a == b) return J >= I;
    const int d = ((b - a) % g + g) % g;
    if (2 * d < g) return 1;
    if (2 * d > g) return 0;
    // distance g/2: lo = min(a, b) splits ITS row tiles; lo computes (I, J) for I in its first half, hi computes the
    // transposed tiles (J', I') of the rest, i.e. the tiles whose COLUMN tile lies in lo's second half
    const int lo = a < b ? a : b;
    const int64_t lo0 = part_row0[lo] / 256, lon = (part_row0[lo + 1] - part_row0[lo] + 255) / 256;
    const int64_t half = (lon + 1) / 2;
    return a < b ? (I - lo0) < half : (J - lo0) >= half;
}

void kmg_gram_i8_clear_cache() {
    std::lock_guard<std::mutex> lk(g_tile_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : g_tile_cache) {
        if (!kv.second.dev) continue;
        if (kv.first.first != cur) cudaSetDevice(kv.first.first);
        cudaFree(kv.second.dev);
        if (kv.first.first != cur) cudaSetDevice(cur);
    }
    g_tile_cache.clear();
}

int kmg_gram_i8_launch(const GramI8Args* a, cudaStream_t stream) {
    KMG_REQUIRE(a->rows > 0 && a->cols > 0, KMG_ERR_ARG, "gram_i8: empty block");
    KMG_REQUIRE(a->Dpad > 0 && a->Dpad % BK == 0, KMG_ERR_ARG, "gram_i8: Dpad (%lld) must be a positive multiple of %d",
                (long long)a->Dpad, BK);
    KMG_REQUIRE(a->ld_phi % 16 == 0 && a->ld_phi >= a->Dpad, KMG_ERR_ARG, "gram_i8: ld_phi must be >= Dpad and a multiple of 16");
    KMG_REQUIRE((reinterpret_cast<uintptr_t>(a->phi_rows) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->phi_cols) & 15) == 0,
                KMG_ERR_ARG, "gram_i8: Phi must be 16-byte aligned");
    KMG_REQUIRE(a->rows < (1ll << 31) && a->cols < (1ll << 31), KMG_ERR_ARG, "gram_i8: block too large");
    KMG_REQUIRE(!a->symmetric || a->n_parts > 0 || a->out_t != nullptr, KMG_ERR_ARG, "gram_i8: symmetric needs a mirror destination");
    int m_sub = a->m_sub;
    static const int default_variant = env_int("KMG_GEMM_VARIANT", 0);
    if (m_sub == 0) m_sub = default_variant;
    if (m_sub == 0) m_sub = 3;  // the CTA-pair kernel is the fastest variant at every measured shape
    KMG_REQUIRE(m_sub >= 1 && m_sub <= 3, KMG_ERR_ARG, "gram_i8: m_sub must be 0 (auto), 1, 2 or 3 (CTA pair)");
    const bool pair = m_sub == 3;
    const int BM = pair ? 256 : 128 * m_sub;
    int dev = 0, sms = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    KMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (a->max_ctas > 0 && a->max_ctas < sms) sms = a->max_ctas;

    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, a->phi_rows, a->rows, a->Dpad, a->ld_phi, pair ? 128 : BM);
    if (rc) return rc;
    rc = make_map(&tmB, a->phi_cols, a->cols, a->Dpad, a->ld_phi, pair ? 128 : BN);
    if (rc) return rc;

    static const int band = env_int("KMG_GEMM_BAND", 8);
    static const int sync_waves = env_int("KMG_GEMM_SYNC", 1);
    static const int hint_mode = env_int("KMG_GEMM_HINT", 0);
    const bool sharded = a->n_parts > 0;
    // The tile list of a plain block does not depend on where the block sits in the Gram; a symmetric block's depends on
    // (row_index0 - col_index0) only; a sharded one's on (part, boundaries).  Keeping the origins out of the key lets the
    // streamed block-rows, the row chunks of the host link and the ranks' offsets share one cached list.
    const bool key_sym = a->symmetric && !sharded;
    TileKey key{a->rows, a->cols, key_sym ? a->row_index0 - a->col_index0 : 0, 0, BM, key_sym ? 1 : 0, band > 0 ? band : 8, 0, 0, {}, 0, 0};
    if (a->mirror_all) {
        KMG_REQUIRE(!sharded && !a->symmetric && a->out_t != nullptr, KMG_ERR_ARG, "gram_i8: mirror_all is a plain block with a transposed copy");
        key.mirror_all = 1;
    }
    if (sharded) {
        KMG_REQUIRE(pair, KMG_ERR_ARG, "gram_i8: the sharded symmetric build uses the CTA-pair kernel (m_sub 0 or 3)");
        KMG_REQUIRE(a->n_parts <= KMG_MAX_PARTS && a->part >= 0 && a->part < a->n_parts && a->part_row0 && a->part_out,
                    KMG_ERR_ARG, "gram_i8: bad sharding (1..%d parts)", KMG_MAX_PARTS);
        for (int q = 0; q <= a->n_parts; ++q) {
            KMG_REQUIRE(q == a->n_parts || a->part_row0[q] % 256 == 0, KMG_ERR_ARG, "gram_i8: part boundaries must be multiples of 256");
            KMG_REQUIRE(q == 0 || a->part_row0[q] > a->part_row0[q - 1], KMG_ERR_ARG, "gram_i8: part boundaries must increase");
            key.bounds[q] = a->part_row0[q];
        }
        KMG_REQUIRE(a->part_row0[0] == 0 && a->part_row0[a->n_parts] == a->cols && a->col_index0 == 0 &&
                    a->row_index0 == a->part_row0[a->part] && a->rows == a->part_row0[a->part + 1] - a->part_row0[a->part],
                    KMG_ERR_ARG, "gram_i8: block shape does not match the sharding");
        key.n_parts = a->n_parts; key.part = a->part;
        static const int interleave = env_int("KMG_SHARD_INTERLEAVE", 1);
        key.interleave = interleave;
    }
    TileList tl;
    rc = get_tiles(key, stream, &tl);
    if (rc) return rc;
    if (a->computed_entries) *a->computed_entries = tl.computed_entries;
    if (tl.n == 0) return KMG_OK;

    KernelParams p;
    p.rows = a->rows; p.cols = a->cols;
    p.row_index0 = a->row_index0; p.col_index0 = a->col_index0;
    p.kblocks = (int32_t)(a->Dpad / BK);
    p.ntiles = tl.n;
    p.out_dtype = a->out_dtype;
    p.out = a->out; p.ldo = a->ldo;
    for (int q = 0; q < KMG_MAX_PARTS; ++q) p.mirror_base[q] = nullptr;
    p.ldo_t = a->ldo_t;
    if (sharded) {
        const int64_t esz = a->out_dtype == KMG_OUT_F64 ? 8 : 4;
        p.ldo_t = a->ldo;
        for (int q = 0; q < a->n_parts; ++q)  // element (r_local, c_global) -> part q's buffer [c_global - row0_q][row_index0 + r_local]
            p.mirror_base[q] = static_cast<char*>(a->part_out[q]) + (a->row_index0 - a->part_row0[q] * a->ldo) * esz;
    } else if (a->symmetric || a->mirror_all) {
        p.mirror_base[0] = a->out_t;
    }
    p.sd_rows = a->sd_rows; p.sd_cols = a->sd_cols;
    p.tiles = tl.dev;
    p.wave_counter = nullptr;
    static const int wave_wait_env = env_int("KMG_GEMM_WAVE_WAIT", -1);
    const int wave_wait = wave_wait_env >= 0 ? wave_wait_env : (pair ? Cfg2::STAGES : 0);
    p.wave_wait_kb = wave_wait < p.kblocks ? wave_wait : 0;
    // the wave barrier only pays when the operands do not stay in L2 on their own (126 MB, two partitions)
    // one wave touches ~27 operand panels of 256 rows x Dpad bytes: below ~8 KB of features per row they all sit in L2
    const bool spills_l2 = a->Dpad >= 8192 && (double)(a->rows + a->cols) * (double)a->Dpad > 96e6;
    if (sync_waves && spills_l2 && tl.n > (pair ? sms / 2 : sms)) {
        rc = get_counter(stream, &p.wave_counter);
        if (rc) return rc;
    }
    p.hint_a = (hint_mode & 1) ? ptx::L2_EVICT_LAST : ptx::L2_EVICT_NORMAL;   // bit 0: the band's A panels stay
    p.hint_b = (hint_mode & 2) ? ptx::L2_EVICT_FIRST : ptx::L2_EVICT_NORMAL;  // bit 1: the B panels stream through
    // FP64-pipe instructions stall behind a saturated tensor pipe: integer conversion once the main loop dominates
    if (a->epi != nullptr && epi_active(*a->epi)) {
        // fused ALIGNF / NLCK steps: a separate kernel variant, so the plain variants carry none of this code
        KMG_REQUIRE(pair && a->out_dtype == KMG_OUT_F64 && !a->symmetric && !a->mirror_all && !sharded, KMG_ERR_ARG,
                    "gram_i8: fused epilogue steps need the CTA-pair kernel, fp64 output and a plain block");
        KMG_REQUIRE(!(a->epi->row_sum_partial || a->epi->row_wsum_partial) || a->epi->n_chunks >= (a->cols + 31) / 32, KMG_ERR_ARG,
                    "gram_i8: n_chunks >= ceil(cols / 32)");
        KMG_REQUIRE(!a->epi->row_wsum_partial || a->epi->w_cols, KMG_ERR_ARG, "gram_i8: weighted row partial sums need w_cols");
        return launch_pair<false, false, true>(tmA, tmB, tmA, p, sms, stream, a->epi);
    }
    if (pair) {
        // Mirror stores through the TMA engine (cp.async.bulk.tensor) whenever the destination allows a tensor map: one
        // destination (not the multi-destination single-launch sharded variant), 16-byte aligned base and row pitch.
        static const int tma_mirror = env_int("KMG_GEMM_TMA_MIRROR", 1);
        const int64_t esz = a->out_dtype == KMG_OUT_F64 ? 8 : 4;
        const bool has_mirror = !sharded && (a->symmetric || a->mirror_all);
        const bool use_tma = tma_mirror && has_mirror && (reinterpret_cast<uintptr_t>(p.mirror_base[0]) & 15) == 0 &&
                             (p.ldo_t * esz) % 16 == 0 && p.ldo_t >= a->rows;
        CUtensorMap tmT = tmA;  // placeholder when unused
        if (use_tma) {
            rc = make_mirror_map(&tmT, p.mirror_base[0], a->out_dtype, a->rows, a->cols, p.ldo_t);
            if (rc) return rc;
            return a->Dpad >= 3072 ? launch_pair<true, true>(tmA, tmB, tmT, p, sms, stream) : launch_pair<false, true>(tmA, tmB, tmT, p, sms, stream);
        }
        return a->Dpad >= 3072 ? launch_pair<true, false>(tmA, tmB, tmT, p, sms, stream) : launch_pair<false, false>(tmA, tmB, tmT, p, sms, stream);
    }
    return m_sub == 1 ? launch<1>(tmA, tmB, p, sms, stream) : launch<2>(tmA, tmB, p, sms, stream);
}

int kmg_gram_i8_simt_launch(const int8_t* A, const int8_t* B, int64_t ld, int64_t rows, int64_t cols, int64_t Dpad,
                            int32_t* out, int64_t ldo, cudaStream_t stream) {
    KMG_REQUIRE(Dpad % 4 == 0 && ld % 4 == 0, KMG_ERR_ARG, "gram_i8_simt: Dpad and ld must be multiples of 4");
    if (rows <= 0 || cols <= 0) return KMG_OK;
    dim3 block(128), grid((unsigned)((cols + 127) / 128), (unsigned)rows);
    KMG_REQUIRE(rows <= 65535, KMG_ERR_ARG, "gram_i8_simt: rows <= 65535");
    gram_i8_simt_kernel<<<grid, block, 0, stream>>>(A, B, ld, rows, cols, Dpad, out, ldo);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}
