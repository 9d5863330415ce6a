// Transposed tcgen05 attention core for the level-1 shape (<= 64 queries per (sample, head), <= 256 keys, head_dim 80).
//   softmax(q k^T / sqrt(hd)) v over keys = [self tokens ; conditioning tokens]      ref/src/modules.py:12-19
//
// Why transposed.  attention_tc.cu computes S = Q K^T with the 64 queries on the accumulator's M axis: tcgen05.mma with
// M = 64 runs at well under half of the M = 128 rate (in-kernel timeline, profiles/r02_attention_timeline.md: 1 500 + 2 850
// cycles of tensor pipe per unit, the limiter of that kernel).  Here the KEYS go on M:
//   MMA1   S^T[keys x 64 q] = K Q^T      M = 128 per instruction (keys in tiles of 128), N = 64, K = 80 = 4 x 16 (128B atom) + 16 (32B atom)
//   softmax  down the TMEM LANES: thread = one key row, its 64 registers = the 64 queries.  Column max by an in-warp
//            transpose-reduce (62 shuffles leave columns 2*lane, 2*lane+1 in each lane) + one shared-memory stage across the
//            warps; P^T = exp2(..) (fp16) is written as one full 128-byte row per thread = the MN-major 128B-swizzle layout
//   MMA2   O^T[128 x 64 q] = V^T P^T     A = V as it lies in memory (keys x head_dim = MN-major): THREE 32-wide 64-byte-swizzle atoms
//            (head dims 0..95) joined by LBO; the instruction's fourth atom (rows 96..127) reads whatever follows in shared memory --
//            an accumulator row depends on its own A row only, so garbage (even NaN) there never reaches a row that is read.
//            B = P^T (MN-major, 128-byte swizzle).  Rows 0..79 of O^T are the head dims; row 80 multiplies a column of ONES the
//            softmax threads plant in the V tile (head-dim slot 80), so the tensor core also delivers the softmax denominator: no
//            second reduction over keys.  (Round-2 history: V as two 64-wide atoms cost 128 B per key and left room for ONE P^T
//            buffer -- the in-kernel timeline showed the loop P^T(u) written -> MMA2(u) -> buffer free -> P^T(u+1) as the
//            critical path, 5 100 cycles per unit; 96 B per key pays for the second buffer.)
//   epilogue O^T / rowsum -> fp16 -> global, one query per store instruction (a warp writes 32 consecutive head dims = 64 B)
// Warp roles (736 threads): 0 TMA producer of K + Q, 2 TMA producer of V, 3 MMA issuer + TMEM allocator (1 idle), 4-19 softmax (two
// groups of 8 warps working on ALTERNATE units, half a period apart; in a group warp w owns key tile w/4, TMEM sub-partition
// w%4, and walks the 64 query columns in two passes of 32), 20-22 epilogue (sub-partitions 0..2 = head dims 0..95).
// Shapes this kernel does not take (attn_weights, > 64 queries, > 256 keys) fall through to attention_tc.cu / attention.cu.
#include "attention.cuh"
#include "gemm.cuh"

namespace pb {
namespace {

constexpr int TT_HD = 80;
constexpr int TT_Q = 64;
constexpr int TT_SM_WARPS = 8;             // softmax warps per group: 2 key tiles x 4 TMEM sub-partitions
constexpr int TT_SM_GROUPS = 2;            // softmax groups: they take alternate units (group g = S^T / P^T / V stage g)
constexpr int TT_QG = 32;                  // query columns per pass (one 32x32b TMEM load)
constexpr int TT_MMA_WARP = 3;             // SM sub-partition 3: the one without a TMA producer or an epilogue warp (warp 1 idles)
constexpr int TT_EPI_WARP0 = 4 + TT_SM_WARPS * TT_SM_GROUPS;      // 20: a multiple of 4 (TMEM sub-partition = warp % 4)
constexpr int TT_THREADS = (TT_EPI_WARP0 + 3) * 32;               // 736

struct TtParams {
    int B, P, nhead, E, S_max;
    int self_rows, sbox, n1, n_mt;
    int nk_st, nv_st, npb;          // ring depths: K(+Q) stages, V stages, P buffers
    const int* kv_len;
    const int* kv_slot;
    float scale_log2;
    __half* out;
    uint32_t off_k, k_bytes;        // nk_st x [Q64 8192 | K64 n1*128 | K16 n1*32 | Q16 2048] (padded to 1024)
    uint32_t off_v, v_bytes;        // nv_st x 3 atoms [V[:, 32a : 32a+32] n1*64]
    uint32_t off_p, p_bytes;        // npb x n1*128
    uint32_t off_red;               // partial column maxima (2 parities), final maxima, 2 x [64] softmax denominators
    uint32_t off_bar;
    TraceBuf trace;                 // PB200_TRACE=attention_tt:<file>
};

enum { TR_TMA = 0, TR_TMAV = 1, TR_MMA = 2, TR_SM = 3, TR_SM1 = 4, TR_EPI = 5 };
enum { TE_K_ISSUE = 0, TE_V_ISSUE, TE_K_READY, TE_S_ISSUED, TE_P_READY, TE_O_ISSUED, TE_S_READY, TE_MAX_DONE, TE_P_FREE, TE_P_WRITTEN,
       TE_O_READY, TE_O_DONE };
const char* const kTtRoles[TRACE_ROLES] = {"tma_kq", "tma_v", "mma", "softmax0", "softmax1", "epilogue", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-"};
const char* const kTtEvents[] = {"k_issue", "v_issue", "k_ready", "s_issued", "p_ready", "o_issued", "s_ready", "max_done", "p_free",
                                 "p_written", "o_ready", "o_done"};

enum { BAR_KF = 0, BAR_KE = 2, BAR_VF = 4, BAR_VE = 6, BAR_SF = 8, BAR_SE = 10, BAR_PF = 12, BAR_PE = 14, BAR_OF = 16, BAR_OE = 18,
       BAR_COUNT = 20 };

// epilogue store of one thread's head dim for `nq` consecutive queries (v[i] = O^T[hd][q0 + i], inv[i] = 1 / softmax denominator):
// one 2-byte store per query row -- a warp covers 32 consecutive head dims = 64 contiguous bytes per instruction.  E_C != 0 makes the
// row pitch a compile-time constant, so every store is [base + immediate] and the 32 multiply / convert / store triples are independent.
template <int E_C>
__device__ __forceinline__ void tt_store_rows(__half* dst, int E, const float (&v)[32], const float* inv, int nq) {
    const int64_t ld = E_C ? E_C : E;
    const float4* lf = reinterpret_cast<const float4*>(inv);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (4 * c < nq) {                                   // nq is a multiple of 16 (warp-uniform)
            const float4 l4 = lf[c];
            dst[(int64_t)(4 * c + 0) * ld] = __float2half_rn(v[4 * c + 0] * l4.x);
            dst[(int64_t)(4 * c + 1) * ld] = __float2half_rn(v[4 * c + 1] * l4.y);
            dst[(int64_t)(4 * c + 2) * ld] = __float2half_rn(v[4 * c + 2] * l4.z);
            dst[(int64_t)(4 * c + 3) * ld] = __float2half_rn(v[4 * c + 3] * l4.w);
        }
    }
}

__global__ void __launch_bounds__(TT_THREADS, 1)
attention_tt_kernel(const __grid_constant__ CUtensorMap tm_q64, const __grid_constant__ CUtensorMap tm_q16,
                    const __grid_constant__ CUtensorMap tm_s64, const __grid_constant__ CUtensorMap tm_s16,
                    const __grid_constant__ CUtensorMap tm_c64, const __grid_constant__ CUtensorMap tm_c16,
                    const __grid_constant__ CUtensorMap tm_sv, const __grid_constant__ CUtensorMap tm_cv, const TtParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar0 = smem_base + p.off_bar;
    auto bar = [&](int slot) { return bar0 + 8u * (uint32_t)slot; };
    const uint32_t tmem_slot = bar0 + 8u * BAR_COUNT;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + p.off_bar + 8 * BAR_COUNT);
    float* colred = reinterpret_cast<float*>(smem_gen + p.off_red);                  // [2 groups][2 slots][8 warps][32]
    float* colfin = colred + 2 * TT_SM_GROUPS * TT_SM_WARPS * TT_QG;                  // [2 groups][8 warps][64]
    float* lsum = colfin + TT_SM_GROUPS * TT_SM_WARPS * TT_Q;                         // [2][64] softmax denominators

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int units = p.B * p.nhead;
    const int s_cols = p.n_mt * TT_Q;                // TMEM columns of one S^T accumulator set

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_q64); ptx::prefetch_tensormap(&tm_q16);
        ptx::prefetch_tensormap(&tm_s64); ptx::prefetch_tensormap(&tm_s16);
        ptx::prefetch_tensormap(&tm_c64); ptx::prefetch_tensormap(&tm_c16);
        ptx::prefetch_tensormap(&tm_sv); ptx::prefetch_tensormap(&tm_cv);
    }
    if (warp == TT_MMA_WARP) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(bar(BAR_KF + i), 1); ptx::mbar_init(bar(BAR_KE + i), 1);
                ptx::mbar_init(bar(BAR_VF + i), 1); ptx::mbar_init(bar(BAR_VE + i), 1);
                ptx::mbar_init(bar(BAR_SF + i), 1); ptx::mbar_init(bar(BAR_SE + i), TT_SM_WARPS);       // buffer i belongs to softmax group i
                ptx::mbar_init(bar(BAR_PF + i), TT_SM_WARPS); ptx::mbar_init(bar(BAR_PE + i), 1);
                ptx::mbar_init(bar(BAR_OF + i), 1); ptx::mbar_init(bar(BAR_OE + i), 3);
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t col_o = (uint32_t)(2 * s_cols);           // two O^T accumulators of 64 columns follow the two S^T sets
    pdl_launch_dependents();

    auto k_stage = [&](int st) { return smem_base + p.off_k + (uint32_t)st * p.k_bytes; };
    auto v_stage = [&](int st) { return smem_base + p.off_v + (uint32_t)st * p.v_bytes; };
    const uint32_t k64_off = TT_Q * 128u, k16_off = k64_off + (uint32_t)p.n1 * 128u, q16_off = k16_off + (uint32_t)p.n1 * 32u;
    const uint32_t v_atom = (uint32_t)p.n1 * 64u;              // one 32-head-dim atom of a V stage

    if (warp == 0) {
        // ===================== TMA producer: K and Q of a unit =====================
        if (ptx::elect_one()) {
            int uc = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x, ++uc) {
                const int b = u / p.nhead, h = u - b * p.nhead;
                const int slot = p.kv_slot ? p.kv_slot[b] : b;
                const int st = uc % p.nk_st;
                ptx::mbar_wait_hint(bar(BAR_KE + st), (((uint32_t)(uc / p.nk_st)) & 1u) ^ 1u);
                const uint32_t fb = bar(BAR_KF + st);
                trace_ev(p.trace, TR_TMA, TE_K_ISSUE, uc);
                ptx::mbar_arrive_expect_tx(fb, (uint32_t)(p.self_rows + p.sbox + TT_Q) * 160u);
                const uint32_t base = k_stage(st);
                const int hc = h * TT_HD;
                if (p.self_rows) {
                    ptx::tma_load_2d(&tm_s64, fb, base + k64_off, p.E + hc, b * p.P);
                    ptx::tma_load_2d(&tm_s16, fb, base + k16_off, p.E + hc + 64, b * p.P);
                }
                ptx::tma_load_2d(&tm_c64, fb, base + k64_off + (uint32_t)p.self_rows * 128u, hc, slot * p.S_max);
                ptx::tma_load_2d(&tm_c16, fb, base + k16_off + (uint32_t)p.self_rows * 32u, hc + 64, slot * p.S_max);
                ptx::tma_load_2d(&tm_q64, fb, base, hc, b * p.P);
                ptx::tma_load_2d(&tm_q16, fb, base + q16_off, hc + 64, b * p.P);
            }
        }
    } else if (warp == 2) {
        // ===================== TMA producer: V of a unit, as three 32-column boxes (the third runs past the head: see header) ======
        if (ptx::elect_one()) {
            int uc = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x, ++uc) {
                const int b = u / p.nhead, h = u - b * p.nhead;
                const int slot = p.kv_slot ? p.kv_slot[b] : b;
                const int st = uc % p.nv_st;
                ptx::mbar_wait_hint(bar(BAR_VE + st), (((uint32_t)(uc / p.nv_st)) & 1u) ^ 1u);
                const uint32_t fb = bar(BAR_VF + st);
                trace_ev(p.trace, TR_TMAV, TE_V_ISSUE, uc);
                ptx::mbar_arrive_expect_tx(fb, (uint32_t)(p.self_rows + p.sbox) * 192u);
                const uint32_t va = v_stage(st);
                const int hc = h * TT_HD;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    if (p.self_rows) ptx::tma_load_2d(&tm_sv, fb, va + (uint32_t)a * v_atom, 2 * p.E + hc + 32 * a, b * p.P);
                    ptx::tma_load_2d(&tm_cv, fb, va + (uint32_t)a * v_atom + (uint32_t)p.self_rows * 64u, p.E + hc + 32 * a, slot * p.S_max);
                }
            }
        }
    } else if (warp == TT_MMA_WARP) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc1 = ptx::umma_idesc_f16_major(128, TT_Q, 0, 0);
        constexpr uint32_t idesc2 = ptx::umma_idesc_f16_major(128, TT_Q, 1, 1);
        auto issue_s = [&](int uc) {
            const int st = uc % p.nk_st, sb = uc & 1;
            ptx::mbar_wait_hint(bar(BAR_KF + st), ((uint32_t)(uc / p.nk_st)) & 1u);
            if (lane == 0) trace_ev(p.trace, TR_MMA, TE_K_READY, uc);
            ptx::mbar_wait_hint(bar(BAR_SE + sb), (((uint32_t)(uc >> 1)) & 1u) ^ 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t base = k_stage(st);
                for (int mt = 0; mt < p.n_mt; ++mt) {
                    const uint32_t d = tmem_base + (uint32_t)(sb * s_cols + mt * TT_Q);
#pragma unroll
                    for (int ks = 0; ks < 5; ++ks) {
                        const uint64_t da = ks < 4 ? ptx::umma_desc_sw128(base + k64_off + (uint32_t)mt * 16384u) + 2 * ks
                                                   : ptx::umma_desc_sw32(base + k16_off + (uint32_t)mt * 4096u);
                        const uint64_t db = ks < 4 ? ptx::umma_desc_sw128(base) + 2 * ks : ptx::umma_desc_sw32(base + q16_off);
                        ptx::umma_f16(d, da, db, idesc1, ks != 0);
                    }
                }
                ptx::umma_commit(bar(BAR_SF + sb));
                ptx::umma_commit(bar(BAR_KE + st));
                trace_ev(p.trace, TR_MMA, TE_S_ISSUED, uc);
            }
            __syncwarp();
        };
        auto issue_o = [&](int uc, int u) {
            const int st = uc % p.nv_st, pb = uc % p.npb, ob = uc & 1;
            const int b = u / p.nhead;
            const int slot = p.kv_slot ? p.kv_slot[b] : b;
            const int nk = p.self_rows + (p.kv_len ? p.kv_len[slot] : p.S_max);
            const int nks = (nk + 15) >> 4;
            ptx::mbar_wait_hint(bar(BAR_VF + st), ((uint32_t)(uc / p.nv_st)) & 1u);
            // the V tile has landed: head-dim slot 80 of every key row := 1 (81..87 := 0), so that row 80 of O^T becomes the softmax
            // denominator (third atom = head dims 64..95, 64-byte rows: slot 80 is 16-byte chunk 2, swizzled with address bits 7..8).
            // Done HERE, by the otherwise idle lanes of the issuing warp, so the softmax groups never wait for a V load.
            for (int r = lane; r < nks * 16; r += 32) {
                uint8_t* vrow = smem_gen + p.off_v + (size_t)st * p.v_bytes + 2 * v_atom + (size_t)r * 64;
                *reinterpret_cast<uint4*>(vrow + ((2 ^ ((r >> 1) & 3)) << 4)) = make_uint4(0x00003C00u, 0u, 0u, 0u);
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            ptx::mbar_wait_hint(bar(BAR_PF + pb), ((uint32_t)(uc / p.npb)) & 1u);
            if (lane == 0) trace_ev(p.trace, TR_MMA, TE_P_READY, uc);
            ptx::mbar_wait_hint(bar(BAR_OE + ob), (((uint32_t)(uc >> 1)) & 1u) ^ 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t va = v_stage(st);
                const uint32_t pt = smem_base + p.off_p + (uint32_t)pb * p.p_bytes;
                const uint32_t d = tmem_base + col_o + (uint32_t)(ob * TT_Q);
                uint64_t da = ptx::umma_desc_mn_sw64(va, v_atom), db = ptx::umma_desc_sw128(pt);
                for (int ks = 0; ks < nks; ++ks, da += 64, db += 128)       // 16 keys = 1024 (V) / 2048 (P^T) bytes further (address field: >> 4)
                    ptx::umma_f16(d, da, db, idesc2, ks != 0);
                ptx::umma_commit(bar(BAR_OF + ob));
                ptx::umma_commit(bar(BAR_PE + pb));
                ptx::umma_commit(bar(BAR_VE + st));
                trace_ev(p.trace, TR_MMA, TE_O_ISSUED, uc);
            }
            __syncwarp();
        };
        // Unit uc's softmax group frees S^T buffer uc&1 and delivers P^T(uc) at the same moment: S^T(uc+2) -- what that group
        // waits for next -- goes to the tensor pipe FIRST, MMA2(uc) behind it (its consumer, the epilogue, has a spare O^T buffer).
        // Both cost ~1 300 cycles here (operand reads at the shared-memory port's limit), so the order is worth one of them
        // per unit on the group's critical path (in-kernel timeline, profiles/r02m_trace_attention_tt.txt).
        int uc = 0, u = blockIdx.x;
        if (u < units) issue_s(0);
        if (u + (int)gridDim.x < units) issue_s(1);
        for (; u < units; u += gridDim.x, ++uc) {
            if (u + 2 * (int)gridDim.x < units) issue_s(uc + 2);
            issue_o(uc, u);
        }
    } else if (warp >= 4 && warp < TT_EPI_WARP0) {
        // ===================== softmax down the lanes: S^T (TMEM) -> P^T (shared, fp16, MN-major 128B swizzle) =====================
        // Two groups of 8 warps take ALTERNATE units (group g: units g, g+2, ... of this CTA = S^T / P^T buffer g), each all 64
        // query columns in two 32-column halves; inside a group warp sw owns key tile sw/4, TMEM sub-partition sw%4.  The groups
        // run half a period apart, so on every SM sub-partition one group's transpose-reduce / exp2 stream fills the other's
        // TMEM-load, shuffle and barrier latencies (the earlier split -- both groups on the SAME unit, 32 columns each -- left all
        // 16 warps stalling in phase: 50 % issue utilisation in ncu, profiles/r02i_*).
        const int grp = (warp - 4) >> 3;
        const int sw = (warp - 4) & 7;
        const int mt = sw >> 2;
        const int j = mt * 128 + (warp & 3) * 32 + lane;             // this thread's key row
        const bool tile_here = mt < p.n_mt;
        const int bar_id = 1 + grp;
        const int nhalf = p.P > 32 ? 2 : 1;                          // 32-query halves with live columns
        float* fin = colfin + (grp * TT_SM_WARPS + sw) * TT_Q;       // this warp's copy of the 64 column maxima (x scale)
        uint32_t syncs = 0;                                          // named-barrier rounds of this group: selects the exchange slot
        const int sb = grp, pb = grp;                                // unit parity = group: it owns these buffers
        for (int u = blockIdx.x + grp * gridDim.x, uc = grp; u < units; u += 2 * gridDim.x, uc += 2) {
            const int b = u / p.nhead;
            const int slot = p.kv_slot ? p.kv_slot[b] : b;
            const int nk = p.self_rows + (p.kv_len ? p.kv_len[slot] : p.S_max);
            const int nk16 = (nk + 15) & ~15;
            const bool tr = sw == 0 && lane == 0;
            const uint32_t ts = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(sb * s_cols + mt * TT_Q);
            const uint32_t par = ((uint32_t)(uc >> 1)) & 1u;
            const bool live = tile_here && j < nk;
            const bool all_live = tile_here && ((j | 31) < nk);      // warp-uniform: no row of this warp needs masking
            float v[32];
            ptx::mbar_wait_hint(bar(BAR_SF + sb), par);
            if (tr) trace_ev(p.trace, TR_SM + grp, TE_S_READY, uc);
            ptx::tc_fence_after();
            for (int hf = 0; hf < nhalf; ++hf) {
                if (tile_here) ptx::tmem_ld_32x32(ts + (uint32_t)(hf * 32), v);      // warp-collective: never under a per-lane condition
                if (!all_live) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = live ? v[i] : -INFINITY;
                }
                // column maxima over this warp's 32 key rows: each round halves the column set a lane holds and exchanges the
                // other half with lane ^ mask; after five rounds lane l holds column l
#define TT_MAX_ROUND(W, MASK)                                                                       \
    _Pragma("unroll") for (int i = 0; i < (W); ++i) {                                                \
        const bool up = (lane & (MASK)) != 0;                                                        \
        const float send = up ? v[i] : v[i + (W)], keep = up ? v[i + (W)] : v[i];                    \
        v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, (MASK)));                              \
    }
                TT_MAX_ROUND(16, 16) TT_MAX_ROUND(8, 8) TT_MAX_ROUND(4, 4) TT_MAX_ROUND(2, 2) TT_MAX_ROUND(1, 1)
#undef TT_MAX_ROUND
                // exchange between the group's 8 warps: two slots, alternating per barrier round (a slot is rewritten two rounds
                // later, after every warp has passed the round in between, i.e. finished reading it)
                float* red = colred + (grp * 2 + (int)(syncs & 1u)) * (TT_SM_WARPS * TT_QG);
                ++syncs;
                red[sw * TT_QG + lane] = v[0];
                ptx::bar_sync(bar_id, TT_SM_WARPS * 32);
                float mx = v[0];
#pragma unroll
                for (int w = 0; w < TT_SM_WARPS; ++w) mx = fmaxf(mx, red[w * TT_QG + lane]);
                if (hf == 0) __syncwarp();              // this warp's reads of `fin` for its previous unit are done
                fin[hf * TT_QG + lane] = mx * p.scale_log2;
            }
            __syncwarp();
            if (tr) trace_ev(p.trace, TR_SM + grp, TE_MAX_DONE, uc);
            ptx::mbar_wait_hint(bar(BAR_PE + pb), par ^ 1u);               // this group's P^T buffer is free again (MMA2 of unit uc-2)
            if (tr) trace_ev(p.trace, TR_SM + grp, TE_P_FREE, uc);
            for (int hf = 0; hf < nhalf; ++hf) {
                if (tile_here) ptx::tmem_ld_32x32(ts + (uint32_t)(hf * 32), v);      // the reduce ran in place: re-read the scores
                if (tile_here && j < nk16) {
                    uint8_t* prow = smem_gen + p.off_p + (size_t)pb * p.p_bytes + (size_t)j * 128;
                    const float4* mf = reinterpret_cast<const float4*>(fin + hf * TT_QG);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {           // 16-byte chunk 4*hf + c = queries 32*hf + 8c .. + 7
                        const float4 m0 = mf[2 * c], m1 = mf[2 * c + 1];
                        uint4 pk;
                        pk.x = pack_half2(ptx::ex2_approx(fmaf(v[8 * c + 0], p.scale_log2, -m0.x)), ptx::ex2_approx(fmaf(v[8 * c + 1], p.scale_log2, -m0.y)));
                        pk.y = pack_half2(ptx::ex2_approx(fmaf(v[8 * c + 2], p.scale_log2, -m0.z)), ptx::ex2_approx(fmaf(v[8 * c + 3], p.scale_log2, -m0.w)));
                        pk.z = pack_half2(ptx::ex2_approx(fmaf(v[8 * c + 4], p.scale_log2, -m1.x)), ptx::ex2_approx(fmaf(v[8 * c + 5], p.scale_log2, -m1.y)));
                        pk.w = pack_half2(ptx::ex2_approx(fmaf(v[8 * c + 6], p.scale_log2, -m1.z)), ptx::ex2_approx(fmaf(v[8 * c + 7], p.scale_log2, -m1.w)));
                        if (!live) pk = make_uint4(0u, 0u, 0u, 0u);          // padding key rows up to the MMA's K step: P^T = 0
                        *reinterpret_cast<uint4*>(prow + (((4 * hf + c) ^ (j & 7)) << 4)) = pk;
                    }
                }
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();              // generic-proxy stores (P^T) -> visible to the tensor core
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(bar(BAR_SE + sb));
                ptx::mbar_arrive(bar(BAR_PF + pb));
            }
            if (tr) trace_ev(p.trace, TR_SM + grp, TE_P_WRITTEN, uc);
        }
    } else if (warp >= TT_EPI_WARP0) {
        // ===================== epilogue: O^T (TMEM) / rowsum -> fp16 -> global =====================
        const int wq = warp & 3;                        // sub-partition = head dims 32*wq ..
        const int hd = wq * 32 + lane;
        const int nhf = p.P > 32 ? 2 : 1;               // 32-query halves of the accumulator that carry live queries
        int uc = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x, ++uc) {
            const int b = u / p.nhead, h = u - b * p.nhead;
            const int ob = uc & 1;
            ptx::mbar_wait_hint(bar(BAR_OF + ob), ((uint32_t)(uc >> 1)) & 1u);
            if (wq == 0 && lane == 0) trace_ev(p.trace, TR_EPI, TE_O_READY, uc);
            ptx::tc_fence_after();
            const uint32_t to = tmem_base + ((uint32_t)(wq * 32) << 16) + col_o + (uint32_t)(ob * TT_Q);
            float* ls = lsum + ob * TT_Q;
            float v[32];
            if (wq == 2) {
                // row 80 (this warp's lane 16) = sum over keys of P^T (the ones column of V); the warp turns the 64 sums into
                // reciprocals once, so the store loop below is multiply / convert / store with no dependent MUFU in between
                for (int hf = 0; hf < nhf; ++hf) {
                    ptx::tmem_ld_32x32(to + (uint32_t)(hf * 32), v);
                    if (hd == TT_HD) {
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            *reinterpret_cast<float4*>(ls + hf * 32 + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    }
                }
                __syncwarp();
                const float i0 = 1.0f / ls[lane], i1 = nhf > 1 ? 1.0f / ls[32 + lane] : 0.f;
                __syncwarp();
                ls[lane] = i0;
                if (nhf > 1) ls[32 + lane] = i1;
            }
            ptx::bar_sync(3, 96);
            __half* dst = p.out + (int64_t)b * p.P * p.E + h * TT_HD + hd;
            for (int hf = 0; hf < nhf; ++hf) {
                ptx::tmem_ld_32x32(to + (uint32_t)(hf * 32), v);
                if (hf == nhf - 1) {                    // the accumulator has been read: MMA2 of unit uc+2 may overwrite it
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(bar(BAR_OE + ob));
                }
                if (hd < TT_HD) {
                    const int nq = p.P - hf * 32 < 32 ? p.P - hf * 32 : 32;
                    if (p.E == 1280) tt_store_rows<1280>(dst + (int64_t)hf * 32 * 1280, 1280, v, ls + hf * 32, nq);
                    else tt_store_rows<0>(dst + (int64_t)hf * 32 * p.E, p.E, v, ls + hf * 32, nq);
                }
            }
            if (wq == 0 && lane == 0) trace_ev(p.trace, TR_EPI, TE_O_DONE, uc);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == TT_MMA_WARP) ptx::tmem_dealloc(tmem_base, 512);
}

int tt_tmap(const void* ptr, int64_t rows, int64_t cols, int box_cols, int box_rows, CUtensorMap* out) {
    return cached_tmap_f16_2d(ptr, rows, cols, cols, box_cols, box_rows, 2 * box_cols, out);      // one swizzle span per box row
}

}  // namespace

// returns 0 = launched, 1 = error, -1 = shape not handled by this kernel (caller tries attention_tc, then the mma.sync kernel)
int launch_attention_tt(const AttnParams& a, cudaStream_t st) {
    static const bool off = getenv("PB200_ATTN_LEGACY") != nullptr || getenv("PB200_ATTN_NO_TT") != nullptr;      // A/B knobs
    if (off) return -1;
    if (a.attn_w != nullptr && a.n_w > 0) return -1;          // post-softmax weights need the un-weighted denominator: attention_tc
    if (a.nhead <= 0 || a.E != a.nhead * TT_HD) return -1;
    if (!(a.P % 16 == 0 && a.P <= TT_Q)) return -1;
    if (a.S_max < 1 || ((uintptr_t)a.qkv & 15) || ((uintptr_t)a.ckv & 15) || ((uintptr_t)a.out & 15)) return -1;
    TtParams p;
    memset(&p, 0, sizeof(p));
    p.B = a.B; p.P = a.P; p.nhead = a.nhead; p.E = a.E; p.S_max = a.S_max;
    p.self_rows = a.self_attn ? a.P : 0;
    p.n1 = (p.self_rows + a.S_max + 15) & ~15;
    p.sbox = p.n1 - p.self_rows;
    if (p.n1 > 256 || p.sbox > 256) return -1;
    p.n_mt = (p.n1 + 127) / 128;
    p.kv_len = a.kv_len; p.kv_slot = a.kv_slot; p.scale_log2 = a.scale_log2; p.out = a.out;
    p.k_bytes = ((uint32_t)(TT_Q * 160 + p.n1 * 160) + 1023u) & ~1023u;
    p.v_bytes = (uint32_t)p.n1 * 192u;
    p.p_bytes = (uint32_t)p.n1 * 128u;
    const uint32_t red_bytes = (2 * TT_SM_GROUPS * TT_SM_WARPS * TT_QG + TT_SM_GROUPS * TT_SM_WARPS * TT_Q + 2 * TT_Q) * 4;
    // MMA1 reads whole 128-row key tiles (the rows past n1 are masked by the softmax) and MMA2 a fourth V atom (rows nobody
    // reads): whatever follows in shared memory.  K stages first, then V, then P^T: every over-read stays inside the allocation
    const uint32_t fixed = red_bytes + 8 * (BAR_COUNT + 1) + 1024 /*alignment slack*/;
    const uint32_t cap = 227 * 1024;
    p.nk_st = 2; p.nv_st = 2; p.npb = 2;      // softmax group g owns stage g of every ring: all rings are two deep
    if (fixed + p.nk_st * p.k_bytes + p.nv_st * p.v_bytes + p.npb * p.p_bytes > cap) return -1;      // more keys: attention_tc
    p.off_k = 0;
    p.off_v = p.nk_st * p.k_bytes;
    p.off_p = p.off_v + p.nv_st * p.v_bytes;
    p.off_red = p.off_p + p.npb * p.p_bytes;
    p.off_bar = p.off_red + red_bytes;
    const size_t smem = (size_t)p.off_bar + 8 * (BAR_COUNT + 1) + 1024;

    if (a.B == 0 || a.P == 0) return 0;
    const int64_t q_rows = (int64_t)a.B * a.P;
    CUtensorMap tq64, tq16, ts64, ts16, tc64, tc16, tsv, tcv;
    PB_TRY(tt_tmap(a.qkv, q_rows, 3 * (int64_t)a.E, 64, TT_Q, &tq64));
    PB_TRY(tt_tmap(a.qkv, q_rows, 3 * (int64_t)a.E, 16, TT_Q, &tq16));
    const int sr = p.self_rows ? p.self_rows : 8;
    PB_TRY(tt_tmap(a.qkv, q_rows, 3 * (int64_t)a.E, 64, sr, &ts64));
    PB_TRY(tt_tmap(a.qkv, q_rows, 3 * (int64_t)a.E, 16, sr, &ts16));
    const int64_t c_rows = (int64_t)(a.n_slots > 0 ? a.n_slots : a.B) * a.S_max;
    PB_TRY(tt_tmap(a.ckv, c_rows, 2 * (int64_t)a.E, 64, p.sbox, &tc64));
    PB_TRY(tt_tmap(a.ckv, c_rows, 2 * (int64_t)a.E, 16, p.sbox, &tc16));
    PB_TRY(tt_tmap(a.qkv, q_rows, 3 * (int64_t)a.E, 32, sr, &tsv));
    PB_TRY(tt_tmap(a.ckv, c_rows, 2 * (int64_t)a.E, 32, p.sbox, &tcv));

    static DeviceOnce attr;
    if (attr.first()) PB_CUDA(cudaFuncSetAttribute(attention_tt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ProfScope prof("attention_tt", 2.0 * ((double)a.B * a.P * 4.0 * a.E + (double)a.B * a.S_max * 2.0 * a.E), st);
    const int units = a.B * a.nhead;
    const int grid = units < sm_count() ? units : sm_count();
    p.trace = a.P == 64 ? trace_begin("attention_tt") : TraceBuf{nullptr};
    attention_tt_kernel<<<grid, TT_THREADS, smem, st>>>(tq64, tq16, ts64, ts16, tc64, tc16, tsv, tcv, p);
    PB_LAUNCH_CHECK();
    if (p.trace.buf) PB_TRY(trace_end("attention_tt", p.trace, kTtRoles, kTtEvents));
    return 0;
}

}  // namespace pb
