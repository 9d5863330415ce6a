// tcgen05 GEMM for every Linear / 1x1-conv / patchify contraction on the path:
//     C[M,N] = A[M,K] . W[N,K]^T   (fp16 operands, fp32 accumulation in TMEM)  + fused epilogue
// Replaces the cuBLAS/cuDNN calls behind nn.Linear / nn.Conv2d(k=1) / Conv2d(k=2,s=2) /
// ConvTranspose2d(k=2,s=2) in ref/src/modules.py (ResBlock :49-55, AttnBlock :71-74, embedding :130-134,
// resamplers :153-156,172-175, clf/out_mapper :179-187) and ref/src/vqgan.py (ResBlock :16-20).
//
// Two persistent, warp-specialised kernels share the epilogues:
//   gemm_f16_cg2_kernel   2-SM tiles (cluster of 2, tcgen05.mma.cta_group::2, M256 x N{128,256}); the default for M > 128
//   gemm_f16_kernel       1-SM tiles (M128 x N{64,128,256}); small M, and the im2col-free convolutions (AMODE 1/2)
// Roles inside a CTA (64 + 32*E threads, E = 8 or 16 epilogue warps):
//   warp 0   : TMA producer  - cp.async.bulk.tensor loads of the 128x64 A tile and of this CTA's share of the W tile,
//              128B-swizzled, into a 4-8 stage shared-memory ring guarded by full/empty mbarriers
//   warp 1   : MMA issuer    - one thread issues tcgen05.mma (K=16) x4 per stage into one of two TMEM accumulators;
//              tcgen05.commit (multicast to both CTAs of a pair) releases the stage / publishes the accumulator
//   warps 2+ : epilogue      - tcgen05.ld 32x32b (lane = row, registers = 32 columns), fused bias / GELU + GRN statistic /
//              residual + FiLM / LayerNorm fold / un-patchify / NCHW transpose; the chunk is transposed inside the warp so
//              that every store instruction writes whole 128-byte lines; overlaps the next tile's MMAs (second accumulator)
// Scheduling: tile width from a tensor / L2-fabric cycle model (gemm_pick_block_n), narrow tail tiles for the leftover
// of the last wave (gemm_tail_block_n), programmatic dependent launch so the prologue overlaps the previous kernel.
#include "gemm.cuh"
#include <type_traits>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>

namespace pb {

// ------------------------------------------------------------------ tensor maps (host)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

int make_tmap_f16_2d(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    PB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    PB_CHECK(((uintptr_t)ptr & 15) == 0, "TMA: base pointer must be 16-byte aligned");
    PB_CHECK((ld * 2) % 16 == 0, "TMA: row stride %lld halves is not a multiple of 16 bytes", (long long)ld);
    PB_CHECK(box_rows >= 1 && box_rows <= 256, "TMA: bad box rows %d", box_rows);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r,
             (long long)rows, (long long)cols, (long long)ld, box_rows);
    return 0;
}

// 2-D fp16 tensor map with an explicit box and swizzle: swizzle_bytes = 128 (box_cols <= 64) or 32 (box_cols <= 16)
int make_tmap_f16_2d_box(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                         int swizzle_bytes) {
    EncodeTiledFn fn = encode_fn();
    PB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    PB_CHECK(((uintptr_t)ptr & 15) == 0, "TMA: base pointer must be 16-byte aligned");
    PB_CHECK((ld * 2) % 16 == 0, "TMA: row stride %lld halves is not a multiple of 16 bytes", (long long)ld);
    PB_CHECK(swizzle_bytes == 128 || swizzle_bytes == 64 || swizzle_bytes == 32 || swizzle_bytes == 0, "TMA: swizzle %d unsupported", swizzle_bytes);
    PB_CHECK(box_rows >= 1 && box_rows <= 256 && box_cols >= 8 && (swizzle_bytes == 0 ? box_cols <= 256 : box_cols * 2 <= swizzle_bytes),
             "TMA: bad box %dx%d for swizzle %d", box_rows, box_cols, swizzle_bytes);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                    : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(box %dx%d, swizzle %d) failed (%d)", box_rows, box_cols, swizzle_bytes, (int)r);
    return 0;
}

// Tensor maps are pure functions of (pointer, shape, box): the executors' workspaces are bump-allocated identically every call
// and the weights never move, so a process-wide cache removes the driver call from all but the first launches.
int cached_tmap_f16_2d(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows, int swizzle_bytes,
                       CUtensorMap* out) {
    using Key = std::tuple<const void*, int64_t, int64_t, int64_t, int, int, int>;
    static std::map<Key, CUtensorMap> cache;
    static std::mutex mu;
    const Key key{ptr, rows, cols, ld, box_cols, box_rows, swizzle_bytes};
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
        if (cache.size() > 8192) cache.clear();
        CUtensorMap tm;
        PB_TRY(make_tmap_f16_2d_box(&tm, ptr, rows, cols, ld, box_cols, box_rows, swizzle_bytes));
        it = cache.emplace(key, tm).first;
    }
    *out = it->second;
    return 0;
}

// rank-N fp16 tensor map (dims/strides innermost first, strides in bytes for dims 1..rank-1), 128-byte swizzle
int make_tmap_f16_nd(CUtensorMap* tm, const void* ptr, int rank, const int64_t* dims, const int64_t* strides_bytes,
                     const int* box) {
    EncodeTiledFn fn = encode_fn();
    PB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    PB_CHECK(rank >= 2 && rank <= 5, "TMA: rank %d unsupported", rank);
    PB_CHECK(((uintptr_t)ptr & 15) == 0, "TMA: base pointer must be 16-byte aligned");
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = (cuuint64_t)dims[i];
        b[i] = (cuuint32_t)box[i];
        e[i] = 1;
        PB_CHECK(box[i] >= 1 && box[i] <= 256, "TMA: bad box[%d]=%d", i, box[i]);
        if (i > 0) {
            PB_CHECK(strides_bytes[i - 1] % 16 == 0, "TMA: stride %lld not a multiple of 16 bytes", (long long)strides_bytes[i - 1]);
            s[i - 1] = (cuuint64_t)strides_bytes[i - 1];
        }
    }
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), d, s, b, e,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rank %d) failed (%d)", rank, (int)r);
    return 0;
}

// ------------------------------------------------------------------ epilogue
// Global operands of a 32-column chunk that do not depend on the accumulator (the fp32 residual row segment).  They
// are fetched BEFORE the wait on the accumulator: out may alias resid (the model updates x in place), so inside the
// store loop the compiler must order every load after the previous store -- 8 dependent L2 round trips per chunk,
// which made the K=1280 out-projection epilogue-bound (73 us for 27 GFLOP).  Preloaded, the latency hides behind
// the tile's main loop.
template <int MODE>
struct EpiPre {};
template <>
struct EpiPre<PB200_EPI_RESID_F32> { float4 r[8]; };
template <>
struct EpiPre<PB200_EPI_F16_LN> { float neg_mean = 0.f, rstd = 0.f; };     // this lane's row, computed once per tile
template <>
struct EpiPre<PB200_EPI_RESID_LN_F32> {
    float4 r[8];
    float ln_s = 0.f, ln_q = 0.f;      // this lane's row: sum / sum of squares over the chunks of the tile done so far
    float shift = 0.f;                 // this lane's row: the shift subtracted before the fp16 copy and the statistics
};

// Coalescing.  tcgen05.ld hands every lane one ROW of the chunk (32 consecutive columns), so a direct 16-byte store
// per lane touches 32 different 128-byte lines per instruction and the LSU serialises them (measured: the fp32
// epilogue of the 8192x1280x1280 out-projection took 2x its main loop).  The chunk is therefore transposed inside the
// warp with xor-butterfly shuffles first:
//   fp32: 8x8 transpose of float4 items in 8-lane groups -> lane (a,b) item i = row 8a+i, columns 4b..4b+3
//         (a store instruction then writes 4 full 128-byte lines);
//   fp16: 4x4 transpose of 8-half items in 4-lane groups -> lane (a,b) item i = row 4a+i, columns 8b..8b+7
//         (8 rows x 64 contiguous bytes per instruction).
__device__ __forceinline__ void transpose4x4_u4(uint32_t (&pk)[16], int lane) {
#pragma unroll
    for (int s = 2; s > 0; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int g0 = 0; g0 < 4; ++g0) {
            if (g0 & s) continue;
            const int g1 = g0 | s;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t send = up ? pk[g0 * 4 + e] : pk[g1 * 4 + e];
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, s);
                if (up) pk[g0 * 4 + e] = recv;
                else pk[g1 * 4 + e] = recv;
            }
        }
    }
}
// fp16 row-per-lane chunk -> transposed, coalesced store.  `orow` is this lane's output row (or -1 if masked).
__device__ __forceinline__ void store_chunk_f16(const float (&v)[32], __half* out, int64_t ldo, int64_t orow, int col0,
                                                int N, int lane) {
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_half2(v[2 * j], v[2 * j + 1]);
    transpose4x4_u4(pk, lane);
    const int col = col0 + (lane & 3) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = __shfl_sync(0xffffffffu, orow, (lane & 28) + i);
        if (r >= 0 && col < N)
            *reinterpret_cast<uint4*>(out + r * ldo + col) = make_uint4(pk[i * 4], pk[i * 4 + 1], pk[i * 4 + 2], pk[i * 4 + 3]);
    }
}

template <int MODE>
__device__ __forceinline__ void epilogue_preload(const pb200_gemm_epilogue& ep, int M, int N, int row, int col0,
                                                 EpiPre<MODE>& pre, int lane, bool first_chunk) {
    if constexpr (MODE == PB200_EPI_F16_LN) {
        if (first_chunk && row < M) {      // row statistics -> (mean, rstd), before the wait on the accumulator
            const longlong2 st = *reinterpret_cast<const longlong2*>(ep.ln_stat + 2 * (int64_t)row);
            const float inv_c = 1.0f / (float)ep.ln_c;
            const float mean = (float)st.x * (1.0f / 1048576.0f) * inv_c;
            const float ex2 = (float)st.y * (1.0f / 65536.0f) * inv_c;
            pre.neg_mean = -mean;
            pre.rstd = 1.0f / sqrtf(fmaxf(ex2 - mean * mean, 0.f) + 1e-6f);
            // the true row mean, for the next producer's shift (every N-tile writes the same value: benign)
            if (ep.ln_mean_out) ep.ln_mean_out[row] = (ep.ln_shift ? ep.ln_shift[row] : 0.f) + mean;
        }
    }
    if constexpr (MODE == PB200_EPI_RESID_LN_F32) {
        if (first_chunk) pre.shift = (ep.ln_shift && row < M) ? ep.ln_shift[row] : 0.f;
    }
    if constexpr (MODE == PB200_EPI_RESID_F32 || MODE == PB200_EPI_RESID_LN_F32) {
        const int col = col0 + (lane & 7) * 4;       // transposed layout: item i = row of lane (lane & 24) + i
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = __shfl_sync(0xffffffffu, row, (lane & 24) + i);
            if (r < M && col < N) pre.r[i] = *reinterpret_cast<const float4*>(ep.resid + (int64_t)r * ep.ldr + col);
        }
    }
}

template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const pb200_gemm_epilogue& ep, int M, int N, int row, int col0,
                                               float (&v)[32], int lane, EpiPre<MODE>& pre) {
    const bool row_ok = row < M;
    // warp-uniform: no column / row of this chunk is out of range -> the hot path carries no per-element predicates
    const bool full_cols = col0 + 32 <= N;
    const bool full = full_cols && __all_sync(0xffffffffu, row_ok);
    // ---- folded LayerNorm of the A rows: acc = x W^T with x un-normalised; LN(x) W^T = rstd (acc - mean rowsum(W))
    if constexpr (MODE == PB200_EPI_F16_LN) {
        const float nm = pre.neg_mean, rstd = pre.rstd;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            if (col0 + g * 4 < N) {
                const float4 ws = __ldg(reinterpret_cast<const float4*>(ep.ln_wsum + col0 + g * 4));
                v[g * 4 + 0] = fmaf(nm, ws.x, v[g * 4 + 0]) * rstd; v[g * 4 + 1] = fmaf(nm, ws.y, v[g * 4 + 1]) * rstd;
                v[g * 4 + 2] = fmaf(nm, ws.z, v[g * 4 + 2]) * rstd; v[g * 4 + 3] = fmaf(nm, ws.w, v[g * 4 + 3]) * rstd;
            }
        }
    }
    // ---- bias (indexed by GEMM column in every mode)
    if (ep.bias) {
        if (full_cols) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + g * 4));
                v[g * 4 + 0] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                if (col0 + g * 4 < N) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + g * 4));
                    v[g * 4 + 0] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
                }
            }
        }
    }
    if (MODE == PB200_EPI_F16 || MODE == PB200_EPI_F32 || MODE == PB200_EPI_F16_LN) {
        int64_t orow = row_ok ? row : -1;
        if (row_ok && ep.remap_in > 0) orow = (int64_t)(row / ep.remap_in) * ep.remap_out + (row % ep.remap_in);
        if (MODE == PB200_EPI_F16 || MODE == PB200_EPI_F16_LN) {
            store_chunk_f16(v, reinterpret_cast<__half*>(ep.out), ep.ldo, orow, col0, N, lane);
        } else {
            transpose8x8_f4(v, lane);
            const int col = col0 + (lane & 7) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t r = __shfl_sync(0xffffffffu, orow, (lane & 24) + i);
                if (r >= 0 && col < N)
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + r * ep.ldo + col) =
                        make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
            }
        }
    } else if (MODE == PB200_EPI_GELU_F16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        store_chunk_f16(v, reinterpret_cast<__half*>(ep.out), ep.ldo, row_ok ? (int64_t)row : -1, col0, N, lane);
        if (ep.sqsum) {   // GlobalResponseNorm statistic: sum over the sample's positions of h^2, per channel.
            // Accumulated in 2^-24 fixed point with 64-bit integer atomics: integer addition is associative, so the
            // result does not depend on the order in which warps/CTAs arrive (float atomics made two runs of the same
            // seed differ in the last bits, which flipped ~2% of the sampled tokens over 8 steps).
            unsigned long long* sq = reinterpret_cast<unsigned long long*>(ep.sqsum);
            auto fx = [](float x) { return (unsigned long long)__float2ull_rn(x * 16777216.0f); };
            const int P = ep.rows_per_sample;
            if (full) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= v[j];
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (row_ok && col0 + j < N) ? v[j] * v[j] : 0.f;
            }
            if ((P & 31) == 0) {
                // transpose-reduce over the warp's 32 rows: lane j ends with the column-(col0+j) sum
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int i = 0; i < o; ++i) {
                        const bool up = (lane & o) != 0;
                        const float send = up ? v[i] : v[i + o];
                        const float keep = up ? v[i + o] : v[i];
                        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
                const int row0 = row - lane;
                if (row0 < M && col0 + lane < N)
                    atomicAdd(sq + (int64_t)(row0 / P) * N + col0 + lane, fx(v[0]));
            } else if (P == 16 || P == 8 || P == 4 || P == 2) {
                // the warp's 32 rows hold 32/P whole samples: the same transpose-reduce inside aligned groups of P
                // lanes, on 32/P sets of P columns -> lane l ends with columns {j*P + (l % P)} of its group's sample
                // (~30 shuffles and 32/P atomics per lane instead of 32 x log2(P) shuffles and 32 atomics)
                auto grouped = [&](auto pc) {
                    constexpr int PP = decltype(pc)::value;
#pragma unroll
                    for (int o = PP / 2; o > 0; o >>= 1) {
                        const bool up = (lane & o) != 0;
#pragma unroll
                        for (int j = 0; j < 32 / PP; ++j)
#pragma unroll
                            for (int i = 0; i < o; ++i) {
                                const float send = up ? v[j * PP + i] : v[j * PP + i + o];
                                const float keep = up ? v[j * PP + i + o] : v[j * PP + i];
                                v[j * PP + i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                            }
                    }
                    const int grow = row - (lane & (PP - 1));          // first row of this lane's sample
                    if (grow < M) {
#pragma unroll
                        for (int j = 0; j < 32 / PP; ++j) {
                            const int col = col0 + j * PP + (lane & (PP - 1));
                            if (col < N) atomicAdd(sq + (int64_t)(grow / PP) * N + col, fx(v[j * PP]));
                        }
                    }
                };
                if (P == 16) grouped(std::integral_constant<int, 16>{});
                else if (P == 8) grouped(std::integral_constant<int, 8>{});
                else if (P == 4) grouped(std::integral_constant<int, 4>{});
                else grouped(std::integral_constant<int, 2>{});
            } else if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < N) atomicAdd(sq + (int64_t)(row / P) * N + col0 + j, fx(v[j]));
            }
        }
    } else if (MODE == PB200_EPI_RESID_F32 || MODE == PB200_EPI_RESID_LN_F32) {
        // (bias was added above in the row-per-lane layout); the rest runs in the transposed, coalesced layout
        transpose8x8_f4(v, lane);
        const int col = col0 + (lane & 7) * 4;
        float* obase = reinterpret_cast<float*>(ep.out);
        float ls[8], lq[8], sh[8];
        if constexpr (MODE == PB200_EPI_RESID_LN_F32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ls[i] = lq[i] = 0.f;
                sh[i] = __shfl_sync(0xffffffffu, pre.shift, (lane & 24) + i);     // item i = the row of lane (lane & 24) + i
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = __shfl_sync(0xffffffffu, row, (lane & 24) + i);
            if (r < M && col < N) {
                float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (MODE == PB200_EPI_RESID_F32 || MODE == PB200_EPI_RESID_LN_F32) rr = pre.r[i];
                float4 y;
                y.x = fmaf(v[i * 4 + 0], ep.alpha, rr.x); y.y = fmaf(v[i * 4 + 1], ep.alpha, rr.y);
                y.z = fmaf(v[i * 4 + 2], ep.alpha, rr.z); y.w = fmaf(v[i * 4 + 3], ep.alpha, rr.w);
                if (ep.film) {
                    const float* fa = ep.film + (int64_t)(r / ep.rows_per_sample) * ep.film_ld + ep.film_off + col;
                    const float4 a = __ldg(reinterpret_cast<const float4*>(fa));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(fa + N));
                    y.x = fmaf(y.x, 1.0f + a.x, b.x); y.y = fmaf(y.y, 1.0f + a.y, b.y);
                    y.z = fmaf(y.z, 1.0f + a.z, b.z); y.w = fmaf(y.w, 1.0f + a.w, b.w);
                }
                *reinterpret_cast<float4*>(obase + (int64_t)r * ep.ldo + col) = y;
                if constexpr (MODE == PB200_EPI_RESID_LN_F32) {
                    y.x -= sh[i]; y.y -= sh[i]; y.z -= sh[i]; y.w -= sh[i];      // (after the fp32 store of the true value)
                    uint2 pk;
                    pk.x = pack_half2(y.x, y.y);
                    pk.y = pack_half2(y.z, y.w);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out16) + (int64_t)r * ep.ldo + col) = pk;
                    ls[i] = (y.x + y.y) + (y.z + y.w);
                    lq[i] = (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
                }
            }
        }
        if constexpr (MODE == PB200_EPI_RESID_LN_F32) {
            // row statistics of the finished rows: item i of lane (a,b) is row 8a+i, columns 4b..4b+3 -> transpose-reduce
            // over the 8 lanes of the group; lane 8a+b ends with the sums of row 8a+b, i.e. of its own `row`
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int i = 0; i < o; ++i) {
                    const float s_send = up ? ls[i] : ls[i + o], s_keep = up ? ls[i + o] : ls[i];
                    const float q_send = up ? lq[i] : lq[i + o], q_keep = up ? lq[i + o] : lq[i];
                    ls[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, o);
                    lq[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, o);
                }
            }
            pre.ln_s += ls[0];          // flushed once per tile by epilogue_finish (one atomic pair per row and warp)
            pre.ln_q += lq[0];
        }
    } else if (MODE == PB200_EPI_UNPATCH_F32) {
        if (!row_ok) return;
        const int hw = ep.up_h * ep.up_w;
        const int b = row / hw, rem = row - b * hw;
        const int y = rem / ep.up_w, x = rem - y * ep.up_w;
        float* obase = reinterpret_cast<float*>(ep.out);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int col = col0 + g * 4;
            if (col < N) {
                const int q = col / ep.up_cout, co = col - q * ep.up_cout;    // q = dy*2+dx
                const int64_t orow = ((int64_t)b * 2 * ep.up_h + 2 * y + (q >> 1)) * (2 * ep.up_w) + 2 * x + (q & 1);
                *reinterpret_cast<float4*>(obase + orow * ep.up_cout + co) =
                    make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
            }
        }
    } else if (MODE == PB200_EPI_NCHW_F32) {
        if (!row_ok) return;
        const int hw = ep.rows_per_sample;
        const int b = row / hw, p = row - b * hw;
        float* o = reinterpret_cast<float*>(ep.out) + ((int64_t)b * N + col0) * hw + p;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j < N) o[(int64_t)j * hw] = v[j];
    }
}

// after the last chunk of a tile: flush what the chunks accumulated per row
template <int MODE>
__device__ __forceinline__ void epilogue_finish(const pb200_gemm_epilogue& ep, int M, int row, EpiPre<MODE>& pre) {
    if constexpr (MODE == PB200_EPI_RESID_LN_F32) {
        if (row < M) {
            unsigned long long* st = reinterpret_cast<unsigned long long*>(ep.ln_stat) + 2 * (int64_t)row;
            atomicAdd(st, (unsigned long long)__float2ll_rn(pre.ln_s * 1048576.0f));
            atomicAdd(st + 1, (unsigned long long)__float2ll_rn(pre.ln_q * 65536.0f));
        }
    }
}

// ------------------------------------------------------------------ kernel
// AMODE 0: A is a [M,K] matrix.  AMODE 1/2: A rows are gathered by TMA straight from an NHWC fp16 activation
// (im2col-free convolution): a 128-row tile is a th x tw patch of the output grid and k-block kb selects a filter
// tap and a 64-channel slice; out-of-image taps are zero-filled by the TMA unit (the conv's zero padding).
template <int BLOCK_N, int MODE, int AMODE>
__global__ void __launch_bounds__(gemm_threads(BLOCK_N), 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const pb200_gemm_epilogue ep, const ConvGeom geom, int M, int N, int K) {
    using L = GemmSmem<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + L::STAGES * L::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (L::STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * L::STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * L::STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * L::STAGES + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int n_tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
    const int n_tiles_m = AMODE == 0 ? (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M : geom.batch * geom.tiles_y * geom.tiles_x;
    const int n_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
    // Optional 2-CTA cluster (launch attribute): the two CTAs take M-tiles 2p and 2p+1 of the same N-tile and each
    // loads only HALF of the W tile, multicasting it into both CTAs' shared memory — the SM<-L2 traffic per FLOP
    // drops by a third (ncu: the single-CTA kernel saturates the L2->SM fabric at ~13 TB/s).
    const uint32_t csize = ptx::cluster_nctarank();
    const uint32_t crank = ptx::cluster_ctarank();
    const int n_units = ((n_tiles_m + (int)csize - 1) / (int)csize) * n_tiles_n;     // work units = (M-tile group, N-tile)
    const int unit0 = blockIdx.x / csize, unit_step = gridDim.x / csize;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < L::STAGES; ++s) {
                ptx::mbar_init(full_bar(s), 1);
                ptx::mbar_init(empty_bar(s), csize);        // every CTA of the cluster must have drained the stage
            }
            for (int s = 0; s < 2; ++s) {
                ptx::mbar_init(tfull_bar(s), 1);
                ptx::mbar_init(tempty_bar(s), gemm_epi_warps(BLOCK_N));       // one arrival per epilogue warp
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, L::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (csize > 1) ptx::cluster_sync();          // peer barriers are initialised before anything can signal them
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                const int mt = (unit / n_tiles_n) * (int)csize + (int)crank;
                const int m_idx = mt * GEMM_BLOCK_M;
                const int n_idx = (unit % n_tiles_n) * BLOCK_N;
                int cb = 0, cy0 = 0, cx0 = 0;
                if (AMODE != 0) {
                    cb = mt / (geom.tiles_y * geom.tiles_x);
                    const int r = mt - cb * geom.tiles_y * geom.tiles_x;
                    cy0 = (r / geom.tiles_x) * geom.th;
                    cx0 = (r % geom.tiles_x) * geom.tw;
                }
                for (int kb = 0; kb < n_kb; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                    ptx::mbar_arrive_expect_tx(full_bar(stage), L::STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                    if (AMODE == 0) {
                        ptx::tma_load_2d(&tm_a, full_bar(stage), sa, kb * GEMM_BLOCK_K, m_idx);
                    } else if (AMODE == 1) {
                        // Conv2d(k=4, s=2, p=1): tap (ky,kx) reads input (2y-1+ky, 2x-1+kx) = parity plane (pa,pb) at
                        // (y+a, x+b) of the [B, H/2, 2, W/2, 2*C] view
                        const int tap = kb / geom.n_cchunk, cc = kb - tap * geom.n_cchunk;
                        const int dy = (tap >> 2) - 1, dx = (tap & 3) - 1;
                        const int a = dy < 0 ? -1 : (dy >> 1), b = dx < 0 ? -1 : (dx >> 1);
                        const int pa = dy - 2 * a, pb = dx - 2 * b;
                        ptx::tma_load_5d(&tm_a, full_bar(stage), sa, pb * geom.cin + cc * 64, cx0 + b, pa, cy0 + a, cb);
                    } else {
                        // ConvTranspose2d(k=4, s=2, p=1), output phase (py,px): 2x2 taps at (y+oy, x+ox)
                        const int tap = kb / geom.n_cchunk, cc = kb - tap * geom.n_cchunk;
                        const int ty = tap >> 1, tx = tap & 1;
                        const int oy = geom.py == 0 ? (ty == 0 ? 0 : -1) : (ty == 0 ? 1 : 0);
                        const int ox = geom.px == 0 ? (tx == 0 ? 0 : -1) : (tx == 0 ? 1 : 0);
                        ptx::tma_load_4d(&tm_a, full_bar(stage), sa, cc * 64, cx0 + ox, cy0 + oy, cb);
                    }
                    if (csize == 1) {          // the W map's box is half a tile (BLOCK_N/2 rows): two loads
                        ptx::tma_load_2d(&tm_b, full_bar(stage), sa + L::A_BYTES, kb * GEMM_BLOCK_K, n_idx);
                        ptx::tma_load_2d(&tm_b, full_bar(stage), sa + L::A_BYTES + L::B_BYTES / 2, kb * GEMM_BLOCK_K,
                                         n_idx + BLOCK_N / 2);
                    } else {   // my half of the W tile, written into both CTAs' stage (same offsets, same barrier slot)
                        ptx::tma_load_2d_mcast(&tm_b, full_bar(stage), sa + L::A_BYTES + crank * (L::B_BYTES / 2),
                                               kb * GEMM_BLOCK_K, n_idx + (int)crank * (BLOCK_N / 2), (uint16_t)0x3);
                    }
                    if (++stage == L::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = ptx::umma_idesc_f16(GEMM_BLOCK_M, BLOCK_N, 0);
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        for (int unit = unit0; unit < n_units; unit += unit_step, ++iter) {
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            ptx::mbar_wait(tempty_bar(as), aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BLOCK_N;
            for (int kb = 0; kb < n_kb; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                    const uint64_t da = ptx::umma_desc_kmajor_sw128(sa);
                    const uint64_t db = ptx::umma_desc_kmajor_sw128(sa + L::A_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
                        // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom: +2 in (addr>>4)
                        ptx::umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if (csize == 1) ptx::umma_commit(empty_bar(stage));  // smem stage reusable once these MMAs retire
                    else ptx::umma_commit_mcast(empty_bar(stage), (uint16_t)0x3);   // ... in both CTAs of the pair
                    if (kb == n_kb - 1) ptx::umma_commit(tfull_bar(as)); // accumulator complete
                }
                __syncwarp();
                if (++stage == L::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;       // which half of the tile's columns (two warps share a quarter)
        constexpr int COLS_PER_WARP = BLOCK_N / (gemm_epi_warps(BLOCK_N) / 4);
        int iter = 0;
        for (int unit = unit0; unit < n_units; unit += unit_step, ++iter) {
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            const int mt = (unit / n_tiles_n) * (int)csize + (int)crank;
            const int m_idx = mt * GEMM_BLOCK_M;
            const int n_idx = (unit % n_tiles_n) * BLOCK_N;
            int row = m_idx + q * 32 + lane;
            if (AMODE != 0) {   // tile row r = (ly, lx) of a th x tw patch -> output pixel row of the NHWC result
                const int cb = mt / (geom.tiles_y * geom.tiles_x);
                const int rr = mt - cb * geom.tiles_y * geom.tiles_x;
                const int r = q * 32 + lane;
                const int gy = (rr / geom.tiles_x) * geom.th + r / geom.tw;
                const int gx = (rr % geom.tiles_x) * geom.tw + r % geom.tw;
                row = (gy < geom.gh && gx < geom.gw)
                          ? ((cb * geom.oh + gy * geom.sy + geom.py) * geom.ow + gx * geom.sx + geom.px)
                          : M;      // M = B*oh*ow: out of range -> masked
            }
            EpiPre<MODE> pre;
            if (n_idx + half * COLS_PER_WARP < N) epilogue_preload<MODE>(ep, M, N, row, n_idx + half * COLS_PER_WARP, pre, lane, true);
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int c = half * COLS_PER_WARP; c < (half + 1) * COLS_PER_WARP; c += 32) {
                if (n_idx + c >= N) break;
                if (c != half * COLS_PER_WARP) epilogue_preload<MODE>(ep, M, N, row, n_idx + c, pre, lane, false);
                float v[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N + c), v);
                epilogue_chunk<MODE>(ep, M, N, row, n_idx + c, v, lane, pre);
            }
            epilogue_finish<MODE>(ep, M, row, pre);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (csize > 1) ptx::cluster_sync();          // the peer may still be multicasting into / signalling this CTA
    if (warp == 1) ptx::tmem_dealloc(tmem_base, L::TMEM_COLS);
}

// ------------------------------------------------------------------ 2-SM kernel (cta_group::2)
// A CTA pair (cluster of 2 on one TPC) computes a 256 x BLOCK_N tile with ONE tcgen05.mma.cta_group::2 stream issued
// by the leader CTA: each CTA stages its own 128 rows of A and only its HALF of the W tile (BLOCK_N/2 rows); the
// tensor cores of both SMs read both halves.  Per SM this halves the W bytes pulled from L2 and the W bytes read
// from shared memory per MMA (ncu on the 1-SM kernel: L2->SM fabric at ~13 TB/s and, for BLOCK_N=128, the
// 128 B/clk shared-memory port are the limiters).  Accumulators: each CTA's TMEM holds its 128 rows x BLOCK_N.
//   full[s]   (leader)  : 2 arrivals (both producers) + the bytes of both CTAs' TMA loads (cta_group::2 loads
//                         complete_tx on the leader's barrier)
//   empty[s]  (each CTA): tcgen05.commit.cta_group::2 multicast from the leader's MMA thread
//   tfull[a]  (each CTA): same commit, when a tile's last k-block has been issued
//   tempty[a] (leader)  : one arrival per epilogue warp of BOTH CTAs (the peer arrives remotely)
// ASCALE (GlobalResponseNorm folded into the A operand, ref/src/modules.py:37-40 + :53): the A tile is multiplied in shared
// memory, between TMA and MMA, by a per-(sample, k) fp16 factor  s[b, k] = 1 + gamma[k] * Nx[b, k]  -- GEMM2 of a ResBlock then
// computes (h * s) W2^T + (W2 beta + b2) = GRN(h) W2^T + b2 without the separate read-modify-write pass over the 4c-wide hidden.
// Four extra warps per CTA (one thread per row) transform every stage: the A tile and the [samples x 64] slice of s land on
// a CTA-local barrier (afull), the transform warps rescale the swizzled rows in place, fence the generic->async proxy and arrive
// on the leader's ready[s]; the MMA thread waits for full[s] (both W halves) and ready[s] (both CTAs' A tiles transformed).
constexpr int GEMM_ASCALE_WARPS = 4;         // one thread per row of the 128-row A tile
constexpr int GEMM_ASCALE_BYTES = 1024;        // up to 8 samples x 64 factors per 128-row tile and k-block
template <int BLOCK_N, int MODE, bool ASCALE>
__global__ void __launch_bounds__(gemm_threads(BLOCK_N) + (ASCALE ? 32 * GEMM_ASCALE_WARPS : 0), 1)
gemm_f16_cg2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const __grid_constant__ CUtensorMap tm_b_tail, const __grid_constant__ CUtensorMap tm_s,
                    const pb200_gemm_epilogue ep, int M, int N, int K, int n_main, int tail_bn) {
    constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
    constexpr int BH_BYTES = (BLOCK_N / 2) * GEMM_BLOCK_K * 2;       // this CTA's half of the W tile
    constexpr int STAGE_BYTES = A_BYTES + BH_BYTES + (ASCALE ? GEMM_ASCALE_BYTES : 0);
    constexpr int STAGES = BLOCK_N >= 256 ? 6 : 8;
    constexpr int TMEM_COLS = BLOCK_N >= 256 ? 512 : 256;
    constexpr int EW = gemm_epi_warps(BLOCK_N);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    auto afull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 6 + s); };          // ASCALE: this CTA's A + factors landed
    auto ready_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + 6 + s); };          // ASCALE (leader): both A tiles rescaled
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t crank = ptx::cluster_ctarank();
    const bool leader = crank == 0;
    const int n_tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
    const int n_pairs_m = (M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
    // Work units.  Units [0, n_main) are full 256 x BLOCK_N pair tiles.  With a tail (tail_bn > 0) the remaining pair
    // tiles -- the ones that would form a mostly idle last wave -- are cut into BLOCK_N / tail_bn narrower tiles of
    // tail_bn columns, so the last wave is short and (nearly) full instead (8192 x 1280: 160 tiles on 74 SM pairs =
    // 2.16 waves ran as 3; with the 12 leftover tiles split 4 ways it runs as 2 + a quarter-width wave).
    const int n_big = n_pairs_m * n_tiles_n;
    const int tail_split = tail_bn > 0 ? BLOCK_N / tail_bn : 1;
    const int n_units = tail_bn > 0 ? n_main + (n_big - n_main) * tail_split : n_big;
    const int unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;
    struct Unit { int m_pair, n0, width; };
    auto decode = [&](int u) {
        Unit r;
        int big = u, sub = 0;
        r.width = BLOCK_N;
        if (u >= n_main && tail_bn > 0) {
            const int v = u - n_main;
            big = n_main + v / tail_split;
            sub = v - (v / tail_split) * tail_split;
            r.width = tail_bn;
        }
        r.m_pair = big / n_tiles_n;
        r.n0 = (big - r.m_pair * n_tiles_n) * BLOCK_N + sub * r.width;
        if (r.n0 >= N) r.width = 0;           // a sub-tile entirely past the last column: every role skips it
        return r;
    };
    const int n_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                ptx::mbar_init(full_bar(s), 2);
                ptx::mbar_init(empty_bar(s), 1);
            }
            for (int s = 0; s < 2; ++s) {
                ptx::mbar_init(tfull_bar(s), 1);
                ptx::mbar_init(tempty_bar(s), 2 * EW);
            }
            if (ASCALE) {
                for (int s = 0; s < STAGES; ++s) {
                    ptx::mbar_init(afull_bar(s), 1);
                    ptx::mbar_init(ready_bar(s), 2 * GEMM_ASCALE_WARPS);
                }
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc_cg2(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish_cg2();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    ptx::cluster_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // everything above overlapped the tail of the previous kernel (programmatic dependent launch); its results are
    // read from here on
    ptx::griddep_launch();
    ptx::griddep_wait();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t leader_full0 = ptx::mapa(full_bar(0), 0);      // leader's full[0] in cluster address space
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                const Unit un = decode(unit);
                if (un.width == 0) continue;
                const int m_idx = un.m_pair * (2 * GEMM_BLOCK_M) + (int)crank * GEMM_BLOCK_M;
                const int n_idx = un.n0 + (int)crank * (un.width / 2);
                const bool narrow = un.width != BLOCK_N;
                const uint32_t tx = 2u * (uint32_t)((ASCALE ? 0 : A_BYTES) + (un.width / 2) * GEMM_BLOCK_K * 2);
                const int P = ep.rows_per_sample;
                const int ns = ASCALE ? (P >= GEMM_BLOCK_M ? 1 : GEMM_BLOCK_M / P) : 0;      // samples in this CTA's 128 rows
                for (int kb = 0; kb < n_kb; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t lfull = leader_full0 + 8u * stage;
                    if (leader) ptx::mbar_arrive_expect_tx(full_bar(stage), tx);
                    else ptx::mbar_arrive_cluster(lfull);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    if (ASCALE) {       // A and its factors complete on this CTA's own barrier: the transform warps wait there
                        ptx::mbar_arrive_expect_tx(afull_bar(stage), (uint32_t)(A_BYTES + ns * GEMM_BLOCK_K * 2));
                        ptx::tma_load_2d(&tm_a, afull_bar(stage), sa, kb * GEMM_BLOCK_K, m_idx);
                        ptx::tma_load_2d(&tm_s, afull_bar(stage), sa + A_BYTES + BH_BYTES, kb * GEMM_BLOCK_K, m_idx / P);
                    } else {
                        ptx::tma_load_2d_cg2(&tm_a, lfull, sa, kb * GEMM_BLOCK_K, m_idx);
                    }
                    ptx::tma_load_2d_cg2(narrow ? &tm_b_tail : &tm_b, lfull, sa + A_BYTES, kb * GEMM_BLOCK_K, n_idx);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            constexpr uint32_t idesc_full = ptx::umma_idesc_f16(2 * GEMM_BLOCK_M, BLOCK_N, 0);
            const uint32_t idesc_tail = ptx::umma_idesc_f16(2 * GEMM_BLOCK_M, tail_bn > 0 ? tail_bn : BLOCK_N, 0);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                const Unit un = decode(unit);
                if (un.width == 0) continue;
                const uint32_t idesc = un.width == BLOCK_N ? idesc_full : idesc_tail;
                const int as = iter & 1;
                const uint32_t aphase = (iter >> 1) & 1;
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BLOCK_N;
                for (int kb = 0; kb < n_kb; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    if (ASCALE) ptx::mbar_wait(ready_bar(stage), phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint32_t sa = smem_base + stage * STAGE_BYTES;
                        const uint64_t da = ptx::umma_desc_kmajor_sw128(sa);
                        const uint64_t db = ptx::umma_desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
                        for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
                            ptx::umma_f16_cg2(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        ptx::umma_commit_cg2_mcast(empty_bar(stage), (uint16_t)0x3);
                        if (kb == n_kb - 1) ptx::umma_commit_cg2_mcast(tfull_bar(as), (uint16_t)0x3);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ++iter;
            }
        }
    } else if (ASCALE && warp >= 2 + EW) {
        // ===================== A-operand transform (both CTAs, own 128 rows) =====================
        const int r = (warp - 2 - EW) * 32 + lane;          // this thread's row of the 128-row tile
        const int P = ep.rows_per_sample;
        const int ns = P >= GEMM_BLOCK_M ? 1 : GEMM_BLOCK_M / P;
        const int sl = P >= GEMM_BLOCK_M ? 0 : min(r / P, ns - 1);          // factor row = sample of this row inside the tile
        const uint32_t leader_ready0 = ptx::mapa(ready_bar(0), 0);
        uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
        int stage = 0;
        uint32_t phase = 0;
        for (int unit = unit0; unit < n_units; unit += unit_step) {
            const Unit un = decode(unit);
            if (un.width == 0) continue;
            for (int kb = 0; kb < n_kb; ++kb) {
                ptx::mbar_wait(afull_bar(stage), phase);
                uint8_t* sa = smem_gen + stage * STAGE_BYTES;
                uint4* arow = reinterpret_cast<uint4*>(sa + r * 128);
                const uint4* srow = reinterpret_cast<const uint4*>(sa + A_BYTES + BH_BYTES + sl * 128);
                uint4 av[8], sv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {                // logical 16-byte chunk j sits at j ^ (r & 7) (128B swizzle)
                    av[j] = arow[j ^ (r & 7)];
                    sv[j] = srow[j];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    __half2* a2 = reinterpret_cast<__half2*>(&av[j]);
                    const __half2* s2 = reinterpret_cast<const __half2*>(&sv[j]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) a2[e] = __hmul2(a2[e], s2[e]);
                    arow[j ^ (r & 7)] = av[j];
                }
                ptx::fence_proxy_async_smem();       // the rescaled tile must be visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(leader_ready0 + 8u * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < 2 + EW) {
        // ===================== epilogue (both CTAs, own 128 rows) =====================
        const int q = warp & 3;
        const int slice = (warp - 2) >> 2;
        constexpr int COLS_PER_WARP = BLOCK_N / (EW / 4);
        const uint32_t leader_tempty0 = ptx::mapa(tempty_bar(0), 0);
        int iter = 0;
        for (int unit = unit0; unit < n_units; unit += unit_step) {
            const Unit un = decode(unit);
            if (un.width == 0) continue;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            ++iter;
            const int m_idx = un.m_pair * (2 * GEMM_BLOCK_M) + (int)crank * GEMM_BLOCK_M;
            const int n_idx = un.n0;
            const int row = m_idx + q * 32 + lane;
            const int c_end = min((slice + 1) * COLS_PER_WARP, un.width);      // a narrow tile leaves the upper slices idle
            EpiPre<MODE> pre;
            if (slice * COLS_PER_WARP < c_end && n_idx + slice * COLS_PER_WARP < N)
                epilogue_preload<MODE>(ep, M, N, row, n_idx + slice * COLS_PER_WARP, pre, lane, true);
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int c = slice * COLS_PER_WARP; c < c_end; c += 32) {
                if (n_idx + c >= N) break;
                if (c != slice * COLS_PER_WARP) epilogue_preload<MODE>(ep, M, N, row, n_idx + c, pre, lane, false);
                float v[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N + c), v);
                epilogue_chunk<MODE>(ep, M, N, row, n_idx + c, v, lane, pre);
            }
            epilogue_finish<MODE>(ep, M, row, pre);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(leader_tempty0 + 8u * as);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 1) ptx::tmem_dealloc_cg2(tmem_base, TMEM_COLS);
}

template <int BLOCK_N, int MODE, bool ASCALE = false>
static int launch_cg2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* tb_tail, int tail_bn,
                      const pb200_gemm_epilogue& ep, int M, int N, int K, cudaStream_t st) {
    constexpr int STAGES = BLOCK_N >= 256 ? 6 : 8;
    constexpr int SMEM = STAGES * (GEMM_BLOCK_M * 128 + (BLOCK_N / 2) * 128 + (ASCALE ? GEMM_ASCALE_BYTES : 0)) + 1024 + 512;
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        PB_CUDA(cudaFuncSetAttribute(gemm_f16_cg2_kernel<BLOCK_N, MODE, ASCALE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    }
    CUtensorMap ts = ta;            // placeholder when unused
    if (ASCALE) {
        const int P = ep.rows_per_sample;
        const int ns = P >= GEMM_BLOCK_M ? 1 : GEMM_BLOCK_M / P;
        const int64_t samples = ((int64_t)M + P - 1) / P;
        PB_TRY(make_tmap_f16_2d_box(&ts, ep.a_scale, samples, K, ep.a_scale_ld, GEMM_BLOCK_K, ns, 0));
    }
    const int n_big = ceil_div(M, 2 * GEMM_BLOCK_M) * ceil_div(N, BLOCK_N);
    const int max_pairs = sm_count() / 2;
    if (!tb_tail || tail_bn <= 0 || tail_bn >= BLOCK_N || BLOCK_N % tail_bn != 0) { tail_bn = 0; tb_tail = &tb; }
    const int n_main = tail_bn ? (n_big / max_pairs) * max_pairs : n_big;      // the complete waves
    const int n_units = n_main + (n_big - n_main) * (tail_bn ? BLOCK_N / tail_bn : 1);
    const int grid = 2 * (n_units < max_pairs ? n_units : max_pairs);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(gemm_threads(BLOCK_N) + (ASCALE ? 32 * GEMM_ASCALE_WARPS : 0));
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    static const bool pdl = getenv("PB200_NO_PDL") == nullptr;      // A/B knob
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    PB_CUDA(cudaLaunchKernelEx(&cfg, gemm_f16_cg2_kernel<BLOCK_N, MODE, ASCALE>, ta, tb, *tb_tail, ts, ep, M, N, K, n_main, tail_bn));
    PB_LAUNCH_CHECK();
    return 0;
}

template <int BLOCK_N>
static int launch_cg2_mode(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* tb_tail, int tail_bn,
                           const pb200_gemm_epilogue& ep, int M, int N, int K, cudaStream_t st) {
    if (ep.a_scale) {       // GlobalResponseNorm folded into the A operand: the two residual epilogues only
        if (ep.mode == PB200_EPI_RESID_F32) return launch_cg2<BLOCK_N, PB200_EPI_RESID_F32, true>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
        if (ep.mode == PB200_EPI_RESID_LN_F32) return launch_cg2<BLOCK_N, PB200_EPI_RESID_LN_F32, true>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
        PB_CHECK(false, "gemm: a_scale is only built for the RESID epilogues (mode %d)", ep.mode);
    }
    switch (ep.mode) {
        case PB200_EPI_F16: return launch_cg2<BLOCK_N, PB200_EPI_F16>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
        case PB200_EPI_F32: return launch_cg2<BLOCK_N, PB200_EPI_F32>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
        case PB200_EPI_GELU_F16: return launch_cg2<BLOCK_N, PB200_EPI_GELU_F16>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
        case PB200_EPI_RESID_F32: return launch_cg2<BLOCK_N, PB200_EPI_RESID_F32>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
        case PB200_EPI_UNPATCH_F32: return launch_cg2<BLOCK_N, PB200_EPI_UNPATCH_F32>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
        case PB200_EPI_NCHW_F32: return launch_cg2<BLOCK_N, PB200_EPI_NCHW_F32>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
        case PB200_EPI_RESID_LN_F32: return launch_cg2<BLOCK_N, PB200_EPI_RESID_LN_F32>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
        case PB200_EPI_F16_LN: return launch_cg2<BLOCK_N, PB200_EPI_F16_LN>(ta, tb, tb_tail, tail_bn, ep, M, N, K, st);
    }
    PB_CHECK(false, "gemm: unknown epilogue mode %d", ep.mode);
    return 1;
}

// ------------------------------------------------------------------ dispatch
template <int BLOCK_N, int MODE, int AMODE = 0>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const pb200_gemm_epilogue& ep, int M, int N, int K,
                      cudaStream_t st, const ConvGeom* geom = nullptr) {
    using L = GemmSmem<BLOCK_N>;
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        PB_CUDA(cudaFuncSetAttribute(gemm_f16_kernel<BLOCK_N, MODE, AMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     L::SMEM_BYTES));
    }
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    if (geom) g = *geom;
    const int tiles_m = AMODE == 0 ? ceil_div(M, GEMM_BLOCK_M) : g.batch * g.tiles_y * g.tiles_x;
    const int tiles_n = ceil_div(N, BLOCK_N);
    // pairs of CTAs (one cluster) share the W tile through TMA multicast when there are >= 2 M-tiles
    // (measured on B200: no gain — the L2 already de-duplicates the two SMs' unicast requests for the same W tile, as
    // B300_MICROARCH.md predicts for clusters <= 4 — so the TMA-multicast pairing is opt-in: PB200_MCAST=1)
    static const bool use_mcast = getenv("PB200_MCAST") != nullptr;
    const int csize = (AMODE == 0 && tiles_m >= 2 && use_mcast) ? 2 : 1;
    const int n_units = ceil_div(tiles_m, csize) * tiles_n;
    const int max_clusters = sm_count() / csize;
    const int grid = csize * (n_units < max_clusters ? n_units : max_clusters);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(gemm_threads(BLOCK_N));
    cfg.dynamicSmemBytes = L::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PB_CUDA(cudaLaunchKernelEx(&cfg, gemm_f16_kernel<BLOCK_N, MODE, AMODE>, ta, tb, ep, g, M, N, K));
    PB_LAUNCH_CHECK();
    return 0;
}

// im2col-free convolution: A gathered from an NHWC fp16 activation (see ConvGeom); fp32 NHWC output (+bias)
int gemm_conv_launch(const CUtensorMap& ta, const CUtensorMap& tb, int block_n, const pb200_gemm_epilogue& ep,
                     const ConvGeom& geom, int64_t N, int64_t K, cudaStream_t st) {
    PB_CHECK(ep.mode == PB200_EPI_F32, "conv gemm: only the fp32 store epilogue is instantiated");
    PB_CHECK(geom.mode == 1 || geom.mode == 2, "conv gemm: bad mode");
    PB_CHECK(geom.tw * geom.th == GEMM_BLOCK_M, "conv gemm: tile must cover 128 positions");
    const int64_t M = (int64_t)geom.batch * geom.oh * geom.ow;     // rows of the output tensor (mask value in-kernel)
    PB_CHECK(M < (1ll << 31) && N % 8 == 0, "conv gemm: problem too large / N not a multiple of 8");
    ProfScope prof(geom.mode == 1 ? "conv_k4s2" : "convT_k4s2", 2.0 * (double)geom.batch * geom.gh * geom.gw * (double)N * (double)K, st);
#define PB_CONV_CASE(BN)                                                                                           \
    case BN:                                                                                                       \
        return geom.mode == 1 ? launch_cfg<BN, PB200_EPI_F32, 1>(ta, tb, ep, (int)M, (int)N, (int)K, st, &geom)    \
                              : launch_cfg<BN, PB200_EPI_F32, 2>(ta, tb, ep, (int)M, (int)N, (int)K, st, &geom);
    switch (block_n) {
        PB_CONV_CASE(64)
        PB_CONV_CASE(128)
        PB_CONV_CASE(256)
    }
#undef PB_CONV_CASE
    PB_CHECK(false, "conv gemm: unsupported BLOCK_N %d", block_n);
    return 1;
}

template <int BLOCK_N>
static int launch_mode(const CUtensorMap& ta, const CUtensorMap& tb, const pb200_gemm_epilogue& ep, int M, int N, int K,
                       cudaStream_t st) {
    switch (ep.mode) {
        case PB200_EPI_F16: return launch_cfg<BLOCK_N, PB200_EPI_F16>(ta, tb, ep, M, N, K, st);
        case PB200_EPI_F32: return launch_cfg<BLOCK_N, PB200_EPI_F32>(ta, tb, ep, M, N, K, st);
        case PB200_EPI_GELU_F16: return launch_cfg<BLOCK_N, PB200_EPI_GELU_F16>(ta, tb, ep, M, N, K, st);
        case PB200_EPI_RESID_F32: return launch_cfg<BLOCK_N, PB200_EPI_RESID_F32>(ta, tb, ep, M, N, K, st);
        case PB200_EPI_UNPATCH_F32: return launch_cfg<BLOCK_N, PB200_EPI_UNPATCH_F32>(ta, tb, ep, M, N, K, st);
        case PB200_EPI_NCHW_F32: return launch_cfg<BLOCK_N, PB200_EPI_NCHW_F32>(ta, tb, ep, M, N, K, st);
        case PB200_EPI_RESID_LN_F32: return launch_cfg<BLOCK_N, PB200_EPI_RESID_LN_F32>(ta, tb, ep, M, N, K, st);
        case PB200_EPI_F16_LN: return launch_cfg<BLOCK_N, PB200_EPI_F16_LN>(ta, tb, ep, M, N, K, st);
    }
    PB_CHECK(false, "gemm: unknown epilogue mode %d", ep.mode);
    return 1;
}

// planning-only override of the multiprocessor count (pb200_gemm_plan); 0 = ask the device
static thread_local int g_plan_sms = 0;
static int plan_sm_count() {
    if (g_plan_sms > 0) return g_plan_sms;
    const int s = sm_count();
    return s > 0 ? s : 148;
}

static bool gemm_use_cg2(int64_t M) {
    // The 2-SM kernel takes every problem with more than one 128-row tile (1243 vs 1095 TFLOP/s on 8192x3840x1280 once
    // its producer stopped issuing MEMBAR.ALL.GPU per k-block, profiles/r01_cg2_gemm_notes.md); PB200_NO_CG2 disables it.
    static const bool off = getenv("PB200_NO_CG2") != nullptr;
    return !off && M > GEMM_BLOCK_M && plan_sm_count() % 2 == 0;
}

bool gemm_can_scale_a(int64_t M, int64_t N, int64_t K, int rows_per_sample) {
    const int P = rows_per_sample;
    if (P <= 0 || !gemm_use_cg2(M) || K % GEMM_BLOCK_K != 0) return false;
    if (!(P >= GEMM_BLOCK_M ? P % GEMM_BLOCK_M == 0 : (GEMM_BLOCK_M % P == 0 && GEMM_BLOCK_M / P <= 8))) return false;
    // 256-wide tiles only: a k-block of a 128-wide tile lasts 256 tensor cycles, less than the rescale of its A tile took with two
    // transform warps (PB200_GRN_FOLD_128=1 lets 128-wide tiles fold too: experiment knob for the four-warp version)
    static const bool allow128 = getenv("PB200_GRN_FOLD_128") != nullptr;
    const int bn = gemm_pick_block_n(M, N, K);
    return bn == 256 || (allow128 && bn == 128);
}

int gemm_pick_block_n(int64_t M, int64_t N, int64_t K, bool allow_cg2) {
    // Cycle model per candidate BLOCK_N (measured anchors on B200, profiles/r01_cg2_gemm_notes.md):
    //   tensor   : waves x k-blocks x 4 MMAs x (BLOCK_N/2 cycles per 128-row MMA)
    //   L2->SM   : all tiles' operand bytes / ~6300 B/clk (the fabric saturates at ~13 TB/s): per k-block a CTA pulls
    //              its 128x64 A tile and BLOCK_N x 64 of W (1-SM kernel) or half of that W (2-SM kernel)
    //   per tile : ~2500 cycles of fill/drain + epilogue tail
    // and pick the minimum.  The 2-SM kernel's work unit is a 256-row pair tile on sm_count/2 pairs.
    static const int force = getenv("PB200_FORCE_BN") ? atoi(getenv("PB200_FORCE_BN")) : 0;   // experiments only
    if (force == 64 || force == 128 || force == 256) return force;
    const int sms = plan_sm_count();
    const bool cg2 = allow_cg2 && gemm_use_cg2(M);
    const long n_kb = K > 0 ? (long)ceil_div(K, GEMM_BLOCK_K) : 16;
    const int cands[3] = {256, 128, 64};
    int best = 128;
    double best_cost = 1e30;
    for (int i = 0; i < (cg2 ? 2 : 3); ++i) {
        const int bn = cands[i];
        const long units = (long)ceil_div(M, cg2 ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M) * ceil_div(N, bn);
        const long workers = cg2 ? sms / 2 : sms;
        const long waves = (units + workers - 1) / workers;
        const double mma = (double)waves * n_kb * 4.0 * (bn / 2.0);
        const double bytes_per_kb = (cg2 ? 2.0 : 1.0) * 16384.0 + bn * 128.0;
        const double l2 = (double)units * n_kb * bytes_per_kb / 6300.0;
        const double cost = (mma > l2 ? mma : l2) + waves * 2500.0;
        if (cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

int gemm_tail_block_n(int64_t M, int64_t N, int block_n) {
    // Worth it when the leftover tiles are a small part of a wave: cut them so that the last wave is full but narrow.
    // A narrow tile costs more per column than a full one (A is re-read per tile: the L2->SM fabric model of
    // gemm_pick_block_n), hence the 0.55 / 0.45 weights; measured on 8192 x 1280 x {1280, 5120}.
    static const int force = getenv("PB200_GEMM_TAIL") ? atoi(getenv("PB200_GEMM_TAIL")) : -1;   // 0 disables, 64/128 force
    if (!gemm_use_cg2(M) || block_n != 256) return 0;
    const long workers = plan_sm_count() / 2;
    const long n_big = (long)ceil_div(M, 2 * GEMM_BLOCK_M) * ceil_div(N, block_n);
    const long rem = n_big % workers;
    if (n_big < workers || rem == 0) return 0;
    if (force >= 0) return (force == 64 || force == 128) ? force : 0;
    double best = 1.0;                 // cost of running the leftover as one more full wave
    int pick = 0;
    const int cand[2] = {128, 64};
    const double weight[2] = {0.55, 0.45};
    for (int i = 0; i < 2; ++i) {
        const long sub = rem * (block_n / cand[i]);
        const double cost = (double)((sub + workers - 1) / workers) * weight[i];
        if (cost < best - 0.15) { best = cost; pick = cand[i]; }
    }
    return pick;
}

int gemm_launch(const CUtensorMap& ta, const CUtensorMap& tb, int block_n, const pb200_gemm_epilogue& ep, int64_t M,
                int64_t N, int64_t K, cudaStream_t st, const GemmTail* tail) {
    PB_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem");
    PB_CHECK(N % 8 == 0, "gemm: N=%lld must be a multiple of 8", (long long)N);
    PB_CHECK(ep.out != nullptr, "gemm: null output");
    if (ep.mode == PB200_EPI_RESID_F32 || ep.mode == PB200_EPI_RESID_LN_F32)
        PB_CHECK(ep.resid != nullptr, "gemm: RESID epilogue without resid");
    if (ep.mode == PB200_EPI_RESID_LN_F32) PB_CHECK(ep.out16 && ep.ln_stat, "gemm: RESID_LN needs out16 and ln_stat");
    if (ep.mode == PB200_EPI_F16_LN)
        PB_CHECK(ep.ln_stat && ep.ln_wsum && ep.ln_c > 0, "gemm: F16_LN needs ln_stat, ln_wsum and ln_c");
    if (ep.mode == PB200_EPI_UNPATCH_F32)
        PB_CHECK(ep.up_cout % 8 == 0 && ep.up_cout * 4 == N && (int64_t)ep.up_h * ep.up_w > 0,
                 "gemm: bad un-patchify geometry");
    if ((ep.mode == PB200_EPI_GELU_F16 && ep.sqsum) || ((ep.mode == PB200_EPI_RESID_F32 || ep.mode == PB200_EPI_RESID_LN_F32) && ep.film) ||
        ep.mode == PB200_EPI_NCHW_F32)
        PB_CHECK(ep.rows_per_sample > 0, "gemm: rows_per_sample required");
    if (ep.a_scale)
        PB_CHECK(gemm_can_scale_a(M, N, K, ep.rows_per_sample) && block_n >= 128 && ep.a_scale_ld % 8 == 0 && ((uintptr_t)ep.a_scale & 15) == 0,
                 "gemm: a_scale needs the 2-SM kernel (M > 128), rows_per_sample dividing or divided by 128, K %% 64 == 0");
    static const char* kTags[8] = {"gemm_f16", "gemm_f32", "gemm_gelu_sqsum", "gemm_resid", "gemm_unpatch", "gemm_nchw",
                                   "gemm_resid", "gemm_f16"};
    ProfScope prof(ep.mode >= 0 && ep.mode < 8 ? kTags[ep.mode] : "gemm", 2.0 * (double)M * (double)N * (double)K, st);
    if (gemm_use_cg2(M) && block_n >= 128) {
        const int tbn = tail ? tail->bn : 0;
        const CUtensorMap* ttb = tail ? tail->tb : nullptr;
        if (block_n == 256) return launch_cg2_mode<256>(ta, tb, ttb, tbn, ep, (int)M, (int)N, (int)K, st);
        return launch_cg2_mode<128>(ta, tb, ttb, tbn, ep, (int)M, (int)N, (int)K, st);
    }
    switch (block_n) {
        case 64: return launch_mode<64>(ta, tb, ep, (int)M, (int)N, (int)K, st);
        case 128: return launch_mode<128>(ta, tb, ep, (int)M, (int)N, (int)K, st);
        case 256: return launch_mode<256>(ta, tb, ep, (int)M, (int)N, (int)K, st);
    }
    PB_CHECK(false, "gemm: unsupported BLOCK_N %d", block_n);
    return 1;
}

int gemm_f16(const void* a, int64_t lda, const void* w, int64_t ldw, int64_t M, int64_t N, int64_t K,
             const pb200_gemm_epilogue& ep, cudaStream_t st) {
    PB_CHECK(K % 8 == 0, "gemm: K=%lld must be a multiple of 8", (long long)K);
    const int bn = gemm_pick_block_n(M, N, K);
    CUtensorMap ta, tb, tbt;
    PB_TRY(make_tmap_f16_2d(&ta, a, M, K, lda, GEMM_BLOCK_M));
    PB_TRY(make_tmap_f16_2d(&tb, w, N, K, ldw, bn / 2));      // W box = half a tile (see the producer)
    GemmTail tail{gemm_tail_block_n(M, N, bn), &tbt};
    if (tail.bn) PB_TRY(make_tmap_f16_2d(&tbt, w, N, K, ldw, tail.bn / 2));
    return gemm_launch(ta, tb, bn, ep, M, N, K, st, tail.bn ? &tail : nullptr);
}

}  // namespace pb

extern "C" int pb200_gemm_plan(int64_t m, int64_t n, int64_t k, int sm_count, int* block_n, int* two_sm, int* tail_block_n) {
    PB_CHECK(m > 0 && n > 0 && k > 0 && block_n && two_sm && tail_block_n, "gemm_plan: bad arguments");
    pb::g_plan_sms = sm_count > 0 ? sm_count : 0;
    const int bn = pb::gemm_pick_block_n(m, n, k);
    *block_n = bn;
    *two_sm = pb::gemm_use_cg2(m) && bn >= 128 ? 1 : 0;
    *tail_block_n = pb::gemm_tail_block_n(m, n, bn);
    pb::g_plan_sms = 0;
    return 0;
}

extern "C" int pb200_gemm_f16(const void* a, int64_t lda, const void* w, int64_t ldw, int64_t m, int64_t n, int64_t k,
                              const pb200_gemm_epilogue* epi, void* stream) {
    PB_CHECK(epi != nullptr, "gemm: null epilogue");
    return pb::gemm_f16(a, lda, w, ldw, m, n, k, *epi, (cudaStream_t)stream);
}
