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

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

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

struct KernelParams {
    int64_t rows, cols;
    int64_t row_index0, col_index0;
    int32_t kblocks;
    int32_t ntiles;
    int32_t out_dtype;  // KMG_OUT_S32 / KMG_OUT_F64
    void* out;
    int64_t ldo;
    void* out_t;  // mirrored destination base (element (r,c) -> out_t[c*ldo_t + r]); may be null
    int64_t ldo_t;
    const double* sd_rows;  // sqrt(diag) per local row / col for cosine normalisation; null = raw
    const double* sd_cols;
    const int4* tiles;  // {row_tile, col_tile, mirror, 0}
    uint32_t* wave_counter;  // grid-wide arrival counter (zeroed before the launch) or null
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
    static constexpr uint32_t EPI_STAGE_WORDS = 32 * 33;  // per epilogue warp: 32 x 32 int32 transpose tile, padded
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + NUM_EPI_WARPS * EPI_STAGE_WORDS * 4;
};

// Epilogue store of one 32-row x 32-column accumulator chunk.  After tcgen05.ld thread t holds row t, so a
// direct store would touch 32 different cache lines per instruction (measured: the LSU wavefronts of that pattern
// cost ~17 us per 256x256 tile).  Instead the chunk is transposed through a per-warp shared-memory tile
// (row stride 33 words: conflict-free both ways) and every warp store writes 32 consecutive entries of one row.
// The mirrored store K[c][r] needs no transpose: for a fixed register j the 32 lanes are 32 consecutive rows.
__device__ __forceinline__ void store_chunk(const KernelParams& p, const uint32_t (&v)[32], uint32_t* stage /*[32][33]*/,
                                            int lane, int64_t row_base, int64_t col0, bool mirror) {
    const int64_t ncol = (p.cols - col0 < 32) ? (p.cols - col0) : 32;  // warp-uniform, > 0
    const int64_t row_t = row_base + lane;                             // the row this thread holds in registers
    const int64_t col = col0 + lane;                                   // the column this thread stores
    const bool col_ok = lane < ncol;
    int64_t nrow = p.rows - row_base;
    if (nrow > 32) nrow = 32;
    if (nrow <= 0) return;  // warp-uniform
#pragma unroll
    for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = v[j];
    __syncwarp();
    if (p.out_dtype == KMG_OUT_S32) {
        int32_t* dst = reinterpret_cast<int32_t*>(p.out) + row_base * p.ldo + col;
        if (col_ok) {
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr)
                if (rr < nrow) dst[(int64_t)rr * p.ldo] = (int32_t)stage[rr * 33 + lane];
        }
        if (mirror && row_t < p.rows) {
            int32_t* dt = reinterpret_cast<int32_t*>(p.out_t) + col0 * p.ldo_t + row_t;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncol) dt[(int64_t)j * p.ldo_t] = (int32_t)v[j];
        }
    } else {
        double* dst = reinterpret_cast<double*>(p.out) + row_base * p.ldo + col;
        const bool norm = p.sd_rows != nullptr;
        const double sc = (norm && col_ok) ? p.sd_cols[col] : 1.0;
        if (col_ok) {
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                if (rr < nrow) {
                    double val = (double)(int32_t)stage[rr * 33 + lane];
                    if (norm) {
                        // normalize_K (kernels.py:408-414): K_ij / (sqrt(K_ii) * sqrt(K_jj)), diagonal := 1.0
                        val = __ddiv_rn(val, __dmul_rn(p.sd_rows[row_base + rr], sc));
                        if (p.row_index0 + row_base + rr == p.col_index0 + col) val = 1.0;
                    }
                    dst[(int64_t)rr * p.ldo] = val;
                }
            }
        }
        if (mirror && row_t < p.rows) {
            double* dt = reinterpret_cast<double*>(p.out_t) + col0 * p.ldo_t + row_t;
            const double sr = norm ? p.sd_rows[row_t] : 1.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < ncol) {
                    double val = (double)(int32_t)v[j];
                    if (norm) val = __ddiv_rn(val, __dmul_rn(sr, p.sd_cols[col0 + j]));  // mirrored tiles never contain the diagonal
                    dt[(int64_t)j * p.ldo_t] = val;
                }
            }
        }
    }
    __syncwarp();  // the staging tile is reused by the next chunk
}

template <int M_SUB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_i8_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const KernelParams p) {
    using C = Cfg<M_SUB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tfull = empty + C::STAGES;
    uint64_t* tempty = tfull + C::ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ACC_STAGES);
    uint32_t* epi_stage = reinterpret_cast<uint32_t*>(smem + C::STAGES * C::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
    }
    if (warp == 1 && lane == 0) {
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
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
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
    } else if (warp == 1) {
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
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int quarter = warp & 3;        // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;    // which 128 of the 256 accumulator columns
        uint32_t acc = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const int4 tile = p.tiles[t];
            const bool mirror = tile.z != 0;
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
                    store_chunk(p, v, epi_stage + (warp - 4) * C::EPI_STAGE_WORDS, lane, row_base, col0, mirror);
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
    if (warp == 2) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
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
    bool operator<(const TileKey& o) const {
        return std::tie(rows, cols, r0, c0, bm, sym, band) < std::tie(o.rows, o.cols, o.r0, o.c0, o.bm, o.sym, o.band);
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
int get_counter(cudaStream_t stream, uint32_t** out) {
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    dev &= 63;
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
};
std::mutex g_tile_mu;
std::map<std::pair<int, TileKey>, TileList> g_tile_cache;  // per device

int get_tiles(const TileKey& key, cudaStream_t stream, TileList* out) {
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tile_mu);
    auto it = g_tile_cache.find({dev, key});
    if (it != g_tile_cache.end()) { *out = it->second; return KMG_OK; }
    const int BM = key.bm;
    const int64_t tm_n = (key.rows + BM - 1) / BM, tn_n = (key.cols + BN - 1) / BN;
    const int G = key.band;
    std::vector<int4> tiles;
    tiles.reserve((size_t)(tm_n * tn_n));
    int64_t entries = 0;
    for (int64_t b0 = 0; b0 < tm_n; b0 += G) {
        const int64_t b1 = (b0 + G < tm_n) ? b0 + G : tm_n;
        for (int64_t tn = 0; tn < tn_n; ++tn) {
            for (int64_t tm = b0; tm < b1; ++tm) {
                int mirror = 0;
                if (key.sym) {
                    // global index ranges of this tile
                    const int64_t rlo = key.r0 + tm * BM, rhi = rlo + BM - 1;
                    const int64_t clo = key.c0 + tn * BN, chi = clo + BN - 1;
                    if (chi < rlo) continue;        // entirely below the diagonal: produced by a mirror store
                    mirror = (clo > rhi) ? 1 : 0;   // entirely above: its transpose is nobody's tile
                }
                tiles.push_back(make_int4((int)tm, (int)tn, mirror, 0));
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
        KMG_CUDA_CHECK(cudaMemcpyAsync(tl.dev, tiles.data(), sizeof(int4) * tiles.size(), cudaMemcpyHostToDevice, stream));
        KMG_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    g_tile_cache[{dev, key}] = tl;
    *out = tl;
    return KMG_OK;
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
    gram_i8_tcgen05_kernel<M_SUB><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, p);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

}  // namespace

int kmg_gram_i8_launch(const GramI8Args* a, cudaStream_t stream) {
    KMG_REQUIRE(a->rows > 0 && a->cols > 0, KMG_ERR_ARG, "gram_i8: empty block");
    KMG_REQUIRE(a->Dpad > 0 && a->Dpad % BK == 0, KMG_ERR_ARG, "gram_i8: Dpad (%lld) must be a positive multiple of %d",
                (long long)a->Dpad, BK);
    KMG_REQUIRE(a->ld_phi % 16 == 0 && a->ld_phi >= a->Dpad, KMG_ERR_ARG, "gram_i8: ld_phi must be >= Dpad and a multiple of 16");
    KMG_REQUIRE((reinterpret_cast<uintptr_t>(a->phi_rows) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->phi_cols) & 15) == 0,
                KMG_ERR_ARG, "gram_i8: Phi must be 16-byte aligned");
    KMG_REQUIRE(a->rows < (1ll << 31) && a->cols < (1ll << 31), KMG_ERR_ARG, "gram_i8: block too large");
    KMG_REQUIRE(!a->symmetric || a->out_t != nullptr, KMG_ERR_ARG, "gram_i8: symmetric needs a mirror destination");
    int m_sub = a->m_sub;
    if (m_sub == 0) m_sub = (a->Dpad >= 2048) ? 2 : 1;
    KMG_REQUIRE(m_sub == 1 || m_sub == 2, KMG_ERR_ARG, "gram_i8: m_sub must be 0, 1 or 2");
    const int BM = 128 * m_sub;
    int dev = 0, sms = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    KMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (a->max_ctas > 0 && a->max_ctas < sms) sms = a->max_ctas;

    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, a->phi_rows, a->rows, a->Dpad, a->ld_phi, BM);
    if (rc) return rc;
    rc = make_map(&tmB, a->phi_cols, a->cols, a->Dpad, a->ld_phi, BN);
    if (rc) return rc;

    static const int band = env_int("KMG_GEMM_BAND", 8);
    static const int sync_waves = env_int("KMG_GEMM_SYNC", 1);
    static const int hint_mode = env_int("KMG_GEMM_HINT", 0);
    TileKey key{a->rows, a->cols, a->row_index0, a->col_index0, BM, a->symmetric ? 1 : 0, band > 0 ? band : 8};
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
    p.out_t = a->symmetric ? a->out_t : nullptr; p.ldo_t = a->ldo_t;
    p.sd_rows = a->sd_rows; p.sd_cols = a->sd_cols;
    p.tiles = tl.dev;
    p.wave_counter = nullptr;
    if (sync_waves && tl.n > sms) {
        rc = get_counter(stream, &p.wave_counter);
        if (rc) return rc;
    }
    p.hint_a = hint_mode == 1 ? ptx::L2_EVICT_LAST : ptx::L2_EVICT_NORMAL;
    p.hint_b = hint_mode == 2 ? ptx::L2_EVICT_FIRST : ptx::L2_EVICT_NORMAL;
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
