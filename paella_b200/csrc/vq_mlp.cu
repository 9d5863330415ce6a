// MEASURED SLOWER THAN THE TWO GEMMS IT REPLACES -- opt-in (PB200_VQ_MLP_FUSED=1), kept with its tests as a recorded experiment:
// 17.2 vs 12.7 ms per 64-image round trip (B200, profiles/r02_vqmlp_timeline.md).  The in-kernel timeline shows why: a 64-wide
// hidden chunk makes GEMM1 re-read the whole 128 x C A tile from shared memory for only 64 output columns (24 x 96 KB per tile;
// a plain GEMM tile amortises its A reads over 256 columns), so one chunk moves ~270 KB through the 128 B/clk shared-memory port
// (>= 2 100 cycles) for 1 536 cycles of tensor work, and the N = 64 MMAs themselves issue at ~112 cycles instead of 32.  A wider
// chunk does not fit TMEM next to the C-column output accumulator with double buffering.
//
// Fused MLP of the codec's ResBlock (ref/src/vqgan.py:16-20,40-41):
//     x += alpha * ( GELU(a W1^T + b1) W2^T + b2 ),   a = fp16 [M, C] (LayerNorm'd rows), W1 [4C, C], W2 [C, 4C]
// in ONE kernel: the 4C-wide hidden never leaves the SM.  The two-kernel form (GEMM1 -> fp16 hidden in HBM -> GEMM2) moves
// 16C bytes per position through HBM for 16C^2 FLOP -- memory-bound at C = 384 -- and its first GEMM has K = C
// only, so its GELU epilogue (~20 instructions per element) outlasts its main loop.  Here, per 256-row pair tile (2-SM,
// tcgen05.mma.cta_group::2, the plumbing of gemm_f16_cg2_kernel):
//   A tile (128 rows x C per CTA) resident in shared memory for the whole tile;
//   for each chunk j of 64 hidden units:
//     GEMM1   acc1[j&1] (TMEM, 64 cols)  = A . W1[64j.., :]^T           K = C, W1 k-blocks streamed through a ring
//     GELU    epilogue warps: TMEM -> +b1 -> erf-GELU -> fp16 -> shared H[j&1] (K-major, 128B swizzle)
//     GEMM2   acc2 (TMEM, C cols)       += H[j&1] . W2[:, 64j..]^T      K = 64, N = C (256 + 128 at C = 384)
//   final epilogue: x = x + alpha * (acc2 + b2), coalesced fp32 stores.
// GEMM1(j+1) is issued before GEMM2(j), so the tensor pipe works on the next chunk while the GELU warps convert this one;
// per chunk the tensor work (2 x 768 cycles per SM at C = 384) and the GELU work (~1300 issue cycles) overlap.
// TMEM: acc2 C columns + 2 x 64 for acc1 = 512 at C = 384.  Shared memory per CTA at C = 384: A 96 KB + W1 ring 8 x 4 KB +
// W2 ring 2 x 24 KB + H 2 x 16 KB = 208 KB.
#include "gemm.cuh"

#include <cstring>

namespace pb {
namespace {

constexpr int MLP_HC = 64;          // hidden units per chunk
constexpr int MLP_S1 = 8;           // W1 ring stages ([32 rows x 64 k] = 4 KB per CTA)
constexpr int MLP_S2 = 2;           // W2 ring stages ([C/2 rows x 64 k] per CTA)
constexpr int MLP_EW = 16;          // GELU / epilogue warps
constexpr int MLP_THREADS = 64 + 32 * MLP_EW;

template <int C>
struct MlpSmem {
    static constexpr int KB1 = C / 64;                      // k-blocks of GEMM1
    static constexpr int A_BYTES = KB1 * 128 * 128;         // KB1 atoms of [128 rows x 64]
    static constexpr int W1_STAGE = 32 * 128;               // this CTA's 32 rows of a 64-row chunk, one k-block
    static constexpr int W2_STAGE = (C / 2) * 128;          // this CTA's C/2 rows of W2, one chunk (64 k)
    static constexpr int H_BYTES = 128 * 128;
    static constexpr int OFF_W1 = A_BYTES;
    static constexpr int OFF_W2 = OFF_W1 + MLP_S1 * W1_STAGE;
    static constexpr int OFF_H = OFF_W2 + MLP_S2 * W2_STAGE;
    static constexpr int OFF_BAR = OFF_H + 2 * H_BYTES;
    static constexpr int N_BAR = 2 + 2 * MLP_S1 + 2 * MLP_S2 + 4 + 4 + 2;
    static constexpr int SMEM = OFF_BAR + 8 * N_BAR + 16 + 1024;
    static constexpr int TMEM_COLS = 512;
    static constexpr int N2A = C > 256 ? 256 : C;           // GEMM2 N split (one MMA instruction takes N <= 256)
    static constexpr int N2B = C - N2A;
};

enum { MR_TMA = 0, MR_MMA = 1, MR_GELU = 2 };
enum { ME_A_ISSUE = 0, ME_W1_ISSUE, ME_W2_ISSUE, ME_A_READY, ME_G1_GO, ME_G1_ISSUED, ME_G2_GO, ME_G2_ISSUED, ME_ACC1_READY, ME_GELU_DONE,
       ME_H_FREE, ME_H_WRITTEN, ME_ACC2_READY, ME_TILE_DONE };
const char* const kMlpRoles[TRACE_ROLES] = {"tma", "mma", "gelu", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-"};
const char* const kMlpEvents[] = {"a_issue", "w1_issue", "w2_issue", "a_ready", "g1_go", "g1_issued", "g2_go", "g2_issued", "acc1_ready",
                                  "gelu_done", "h_free", "h_written", "acc2_ready", "tile_done"};

// out[M, C] (fp32, in place on the residual stream) ; M rows, any M
template <int C>
__global__ void __launch_bounds__(MLP_THREADS, 1)
vq_mlp_fused_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w1,
                    const __grid_constant__ CUtensorMap tm_w2a, const __grid_constant__ CUtensorMap tm_w2b, const float* __restrict__ b1,
                    const float* __restrict__ b2, float* __restrict__ x, float alpha, int M, const TraceBuf trace) {
    using L = MlpSmem<C>;
    constexpr int NCH = 4 * C / MLP_HC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar0 = smem_base + L::OFF_BAR;
    int bi = 0;
    const uint32_t a_full = bar0 + 8u * bi++, a_empty = bar0 + 8u * bi++;
    const uint32_t w1_full0 = bar0 + 8u * bi; bi += MLP_S1;
    const uint32_t w1_empty0 = bar0 + 8u * bi; bi += MLP_S1;
    const uint32_t w2_full0 = bar0 + 8u * bi; bi += MLP_S2;
    const uint32_t w2_empty0 = bar0 + 8u * bi; bi += MLP_S2;
    const uint32_t acc1_full0 = bar0 + 8u * bi; bi += 2;
    const uint32_t acc1_empty0 = bar0 + 8u * bi; bi += 2;
    const uint32_t h_full0 = bar0 + 8u * bi; bi += 2;
    const uint32_t h_empty0 = bar0 + 8u * bi; bi += 2;
    const uint32_t acc2_full = bar0 + 8u * bi++, acc2_empty = bar0 + 8u * bi++;
    const uint32_t tmem_slot = bar0 + 8u * bi;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + L::OFF_BAR + 8 * bi);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t crank = ptx::cluster_ctarank();
    const bool leader = crank == 0;
    const int n_tiles = (M + 255) / 256;
    const int tile0 = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a); ptx::prefetch_tensormap(&tm_w1);
        ptx::prefetch_tensormap(&tm_w2a); ptx::prefetch_tensormap(&tm_w2b);
    }
    if (warp == 1) {
        if (lane == 0) {
            ptx::mbar_init(a_full, 2); ptx::mbar_init(a_empty, 1);
            for (int s = 0; s < MLP_S1; ++s) { ptx::mbar_init(w1_full0 + 8u * s, 2); ptx::mbar_init(w1_empty0 + 8u * s, 1); }
            for (int s = 0; s < MLP_S2; ++s) { ptx::mbar_init(w2_full0 + 8u * s, 2); ptx::mbar_init(w2_empty0 + 8u * s, 1); }
            for (int s = 0; s < 2; ++s) {
                ptx::mbar_init(acc1_full0 + 8u * s, 1); ptx::mbar_init(acc1_empty0 + 8u * s, 2 * MLP_EW);
                ptx::mbar_init(h_full0 + 8u * s, 2 * MLP_EW); ptx::mbar_init(h_empty0 + 8u * s, 1);
            }
            ptx::mbar_init(acc2_full, 1); ptx::mbar_init(acc2_empty, 2 * MLP_EW);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc_cg2(tmem_slot, L::TMEM_COLS);
        ptx::tmem_relinquish_cg2();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    ptx::cluster_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;
    ptx::griddep_launch();
    ptx::griddep_wait();

    const uint32_t sA = smem_base, sW1 = smem_base + L::OFF_W1, sW2 = smem_base + L::OFF_W2, sH = smem_base + L::OFF_H;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs): A once per tile, W1 k-blocks and W2 chunks as rings =====================
        if (ptx::elect_one()) {
            const uint32_t l_a_full = ptx::mapa(a_full, 0), l_w1_full0 = ptx::mapa(w1_full0, 0), l_w2_full0 = ptx::mapa(w2_full0, 0);
            int s1 = 0, s2 = 0;
            uint32_t p1 = 0, p2 = 0, pa = 0;
            for (int t = tile0; t < n_tiles; t += tile_step) {
                const int m_idx = t * 256 + (int)crank * 128;
                ptx::mbar_wait(a_empty, pa ^ 1);
                trace_ev(trace, MR_TMA, ME_A_ISSUE, t);
                if (leader) ptx::mbar_arrive_expect_tx(a_full, 2u * L::A_BYTES);
                else ptx::mbar_arrive_cluster(l_a_full);
                for (int kb = 0; kb < L::KB1; ++kb) ptx::tma_load_2d_cg2(&tm_a, l_a_full, sA + kb * 16384, kb * 64, m_idx);
                pa ^= 1;
                for (int j = 0; j < NCH; ++j) {
                    for (int kb = 0; kb < L::KB1; ++kb) {       // W1 rows [64j + 32*crank, +32), k-block kb
                        ptx::mbar_wait(w1_empty0 + 8u * s1, p1 ^ 1);
                        if (kb == 0) trace_ev(trace, MR_TMA, ME_W1_ISSUE, j);
                        const uint32_t lf = l_w1_full0 + 8u * s1;
                        if (leader) ptx::mbar_arrive_expect_tx(w1_full0 + 8u * s1, 2u * L::W1_STAGE);
                        else ptx::mbar_arrive_cluster(lf);
                        ptx::tma_load_2d_cg2(&tm_w1, lf, sW1 + s1 * L::W1_STAGE, kb * 64, j * MLP_HC + (int)crank * 32);
                        if (++s1 == MLP_S1) { s1 = 0; p1 ^= 1; }
                    }
                    // W2 rows: this CTA's half of each N part, k columns [64j, +64)
                    ptx::mbar_wait(w2_empty0 + 8u * s2, p2 ^ 1);
                    trace_ev(trace, MR_TMA, ME_W2_ISSUE, j);
                    const uint32_t lf2 = l_w2_full0 + 8u * s2;
                    if (leader) ptx::mbar_arrive_expect_tx(w2_full0 + 8u * s2, 2u * L::W2_STAGE);
                    else ptx::mbar_arrive_cluster(lf2);
                    ptx::tma_load_2d_cg2(&tm_w2a, lf2, sW2 + s2 * L::W2_STAGE, j * MLP_HC, (int)crank * (L::N2A / 2));
                    if (L::N2B > 0)
                        ptx::tma_load_2d_cg2(&tm_w2b, lf2, sW2 + s2 * L::W2_STAGE + (L::N2A / 2) * 128, j * MLP_HC, L::N2A + (int)crank * (L::N2B / 2));
                    if (++s2 == MLP_S2) { s2 = 0; p2 ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            constexpr uint32_t idesc1 = ptx::umma_idesc_f16(256, MLP_HC, 0);
            constexpr uint32_t idesc2a = ptx::umma_idesc_f16(256, L::N2A, 0);
            constexpr uint32_t idesc2b = ptx::umma_idesc_f16(256, L::N2B > 0 ? L::N2B : 16, 0);
            int s1 = 0, s2 = 0;
            uint32_t p1 = 0, p2 = 0, pa = 0;
            int c1 = 0;            // chunks issued to acc1 so far (buffer = c1 & 1, phase = (c1 >> 1) & 1)
            int c2 = 0;            // chunks consumed from H so far
            int tiles_done = 0;
            auto gemm1 = [&](int) {
                const int ab = c1 & 1;
                ptx::mbar_wait(acc1_empty0 + 8u * ab, (((uint32_t)(c1 >> 1)) & 1u) ^ 1u);
                ptx::tc_fence_after();
                for (int kb = 0; kb < L::KB1; ++kb) {
                    ptx::mbar_wait(w1_full0 + 8u * s1, p1);
                    if (kb == 0 && lane == 0) trace_ev(trace, MR_MMA, ME_G1_GO, c1);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint64_t da = ptx::umma_desc_kmajor_sw128(sA + kb * 16384);
                        const uint64_t db = ptx::umma_desc_kmajor_sw128(sW1 + s1 * L::W1_STAGE);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::umma_f16_cg2(tmem_base + (uint32_t)(C + ab * MLP_HC), da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
                        ptx::umma_commit_cg2_mcast(w1_empty0 + 8u * s1, (uint16_t)0x3);
                        if (kb == L::KB1 - 1) {
                            ptx::umma_commit_cg2_mcast(acc1_full0 + 8u * ab, (uint16_t)0x3);
                            trace_ev(trace, MR_MMA, ME_G1_ISSUED, c1);
                        }
                    }
                    __syncwarp();
                    if (++s1 == MLP_S1) { s1 = 0; p1 ^= 1; }
                }
                ++c1;
            };
            auto gemm2 = [&](bool first, bool last) {
                const int hb = c2 & 1;
                ptx::mbar_wait(h_full0 + 8u * hb, ((uint32_t)(c2 >> 1)) & 1u);
                ptx::mbar_wait(w2_full0 + 8u * s2, p2);
                if (first) ptx::mbar_wait(acc2_empty, (((uint32_t)tiles_done) & 1u) ^ 1u);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    trace_ev(trace, MR_MMA, ME_G2_GO, c2);
                    const uint64_t da = ptx::umma_desc_kmajor_sw128(sH + hb * L::H_BYTES);
                    const uint64_t dba = ptx::umma_desc_kmajor_sw128(sW2 + s2 * L::W2_STAGE);
                    const uint64_t dbb = ptx::umma_desc_kmajor_sw128(sW2 + s2 * L::W2_STAGE + (L::N2A / 2) * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t acc = (first && k == 0) ? 0u : 1u;
                        ptx::umma_f16_cg2(tmem_base, da + 2 * k, dba + 2 * k, idesc2a, acc);
                        if (L::N2B > 0) ptx::umma_f16_cg2(tmem_base + (uint32_t)L::N2A, da + 2 * k, dbb + 2 * k, idesc2b, acc);
                    }
                    ptx::umma_commit_cg2_mcast(h_empty0 + 8u * hb, (uint16_t)0x3);
                    ptx::umma_commit_cg2_mcast(w2_empty0 + 8u * s2, (uint16_t)0x3);
                    if (last) ptx::umma_commit_cg2_mcast(acc2_full, (uint16_t)0x3);
                    trace_ev(trace, MR_MMA, ME_G2_ISSUED, c2);
                }
                __syncwarp();
                if (++s2 == MLP_S2) { s2 = 0; p2 ^= 1; }
                ++c2;
            };
            for (int t = tile0; t < n_tiles; t += tile_step) {
                ptx::mbar_wait(a_full, pa);
                if (lane == 0) trace_ev(trace, MR_MMA, ME_A_READY, t);
                pa ^= 1;
                gemm1(0);
                for (int j = 1; j < NCH; ++j) {
                    gemm1(j);                               // the tensor pipe works on chunk j while the GELU warps convert j-1
                    if (j == NCH - 1 && lane == 0) ptx::umma_commit_cg2_mcast(a_empty, (uint16_t)0x3);     // A is free after the last GEMM1
                    __syncwarp();
                    gemm2(j == 1, false);
                }
                gemm2(NCH == 1, true);
                ++tiles_done;
            }
        }
    } else {
        // ===================== GELU warps (per chunk) and final epilogue (per tile), both CTAs =====================
        const int ew = warp - 2;
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int slice = ew >> 2;              // 0..3: which 16 of the chunk's 64 columns / which quarter of the C output columns
        const int r_tile = q * 32 + lane;       // row inside this CTA's 128
        const uint32_t l_acc1_empty0 = ptx::mapa(acc1_empty0, 0), l_h_full0 = ptx::mapa(h_full0, 0), l_acc2_empty = ptx::mapa(acc2_empty, 0);
        int c = 0;                              // chunk counter
        int tiles_done = 0;
        for (int t = tile0; t < n_tiles; t += tile_step) {
            const int row = t * 256 + (int)crank * 128 + r_tile;
            for (int j = 0; j < NCH; ++j, ++c) {
                const int ab = c & 1;
                const uint32_t ph = ((uint32_t)(c >> 1)) & 1u;
                ptx::mbar_wait(acc1_full0 + 8u * ab, ph);
                const bool tr = ew == 0 && lane == 0;
                if (tr) trace_ev(trace, MR_GELU, ME_ACC1_READY, c);
                ptx::tc_fence_after();
                float v[16];
                ptx::tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(C + ab * MLP_HC + slice * 16), v);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(l_acc1_empty0 + 8u * ab);       // acc1 buffer is free for GEMM1(j+2)
                const float* bp = b1 + j * MLP_HC + slice * 16;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(bp) + g);
                    v[g * 4 + 0] = gelu_erf_fast(v[g * 4 + 0] + bb.x); v[g * 4 + 1] = gelu_erf_fast(v[g * 4 + 1] + bb.y);
                    v[g * 4 + 2] = gelu_erf_fast(v[g * 4 + 2] + bb.z); v[g * 4 + 3] = gelu_erf_fast(v[g * 4 + 3] + bb.w);
                }
                if (tr) trace_ev(trace, MR_GELU, ME_GELU_DONE, c);
                ptx::mbar_wait(h_empty0 + 8u * ab, ph ^ 1);                                 // GEMM2(j-2) is done with this H buffer
                if (tr) trace_ev(trace, MR_GELU, ME_H_FREE, c);
                uint8_t* hrow = smem_gen + L::OFF_H + ab * L::H_BYTES + r_tile * 128;
#pragma unroll
                for (int g = 0; g < 2; ++g) {                // this warp's 16 columns = 16-byte chunks 2*slice, 2*slice + 1 of the row
                    uint4 pk;
                    pk.x = pack_half2(v[g * 8 + 0], v[g * 8 + 1]); pk.y = pack_half2(v[g * 8 + 2], v[g * 8 + 3]);
                    pk.z = pack_half2(v[g * 8 + 4], v[g * 8 + 5]); pk.w = pack_half2(v[g * 8 + 6], v[g * 8 + 7]);
                    *reinterpret_cast<uint4*>(hrow + (((2 * slice + g) ^ (r_tile & 7)) << 4)) = pk;
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(l_h_full0 + 8u * ab);
                if (tr) trace_ev(trace, MR_GELU, ME_H_WRITTEN, c);
            }
            // ---- final epilogue: x = x + alpha * (acc2 + b2)
            ptx::mbar_wait(acc2_full, ((uint32_t)tiles_done) & 1u);
            if (ew == 0 && lane == 0) trace_ev(trace, MR_GELU, ME_ACC2_READY, t);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int c0 = slice * 32; c0 < C; c0 += 128) {       // 32-column chunks, interleaved over the four warps of a lane quarter
                float v[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + c0) + g);
                    v[g * 4 + 0] += bb.x; v[g * 4 + 1] += bb.y; v[g * 4 + 2] += bb.z; v[g * 4 + 3] += bb.w;
                }
                transpose8x8_f4(v, lane);                   // item i of lane (a, b) = row 8a + i, columns 4b .. 4b+3
                const int col = c0 + (lane & 7) * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = __shfl_sync(0xffffffffu, row, (lane & 24) + i);
                    if (r < M) {
                        float4* px = reinterpret_cast<float4*>(x + (int64_t)r * C + col);
                        float4 o = *px;
                        o.x = fmaf(v[i * 4 + 0], alpha, o.x); o.y = fmaf(v[i * 4 + 1], alpha, o.y);
                        o.z = fmaf(v[i * 4 + 2], alpha, o.z); o.w = fmaf(v[i * 4 + 3], alpha, o.w);
                        *px = o;
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(l_acc2_empty);
            if (ew == 0 && lane == 0) trace_ev(trace, MR_GELU, ME_TILE_DONE, t);
            ++tiles_done;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 1) ptx::tmem_dealloc_cg2(tmem_base, L::TMEM_COLS);
}

template <int C>
int launch_mlp(const __half* a16, int64_t M, const __half* w1, const float* b1, const __half* w2, const float* b2, float* x, float alpha,
               cudaStream_t st) {
    using L = MlpSmem<C>;
    CUtensorMap ta, tw1, tw2a, tw2b;
    PB_TRY(cached_tmap_f16_2d(a16, M, C, C, 64, 128, 128, &ta));
    PB_TRY(cached_tmap_f16_2d(w1, 4 * C, C, C, 64, 32, 128, &tw1));
    PB_TRY(cached_tmap_f16_2d(w2, C, 4 * C, 4 * C, 64, L::N2A / 2, 128, &tw2a));
    tw2b = tw2a;
    if (L::N2B > 0) PB_TRY(cached_tmap_f16_2d(w2, C, 4 * C, 4 * C, 64, L::N2B / 2, 128, &tw2b));
    static DeviceOnce once;
    if (once.first()) PB_CUDA(cudaFuncSetAttribute(vq_mlp_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM));
    const int n_tiles = (int)((M + 255) / 256);
    const int pairs = sm_count() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (n_tiles < pairs ? n_tiles : pairs));
    cfg.blockDim = dim3(MLP_THREADS);
    cfg.dynamicSmemBytes = L::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    const TraceBuf trace = C == 384 ? trace_begin("vq_mlp") : TraceBuf{nullptr};
    PB_CUDA(cudaLaunchKernelEx(&cfg, vq_mlp_fused_kernel<C>, ta, tw1, tw2a, tw2b, b1, b2, x, alpha, (int)M, trace));
    PB_LAUNCH_CHECK();
    if (trace.buf) PB_TRY(trace_end("vq_mlp", trace, kMlpRoles, kMlpEvents));
    return 0;
}

}  // namespace

// x[M, C] += alpha * (GELU(a16 W1^T + b1) W2^T + b2).  0 = launched, 1 = error, -1 = shape not handled (caller runs two GEMMs)
int launch_vq_mlp_fused(const __half* a16, int64_t M, int C, const __half* w1, const float* b1, const __half* w2, const float* b2,
                        float* x, float alpha, cudaStream_t st) {
    if (M < 256 || M >= (1ll << 31) || sm_count() % 2 != 0) return -1;
    ProfScope prof("gemm_vq_mlp_fused", 16.0 * (double)M * (double)C * (double)C, st);
    if (C == 384) return launch_mlp<384>(a16, M, w1, b1, w2, b2, x, alpha, st);
    if (C == 192) return launch_mlp<192>(a16, M, w1, b1, w2, b2, x, alpha, st);
    return -1;
}

}  // namespace pb
