// tcgen05.mma issue-to-completion cost by shape / operand layout, one CTA, one issuing thread.  Developer tool behind the
// attention kernels' design notes (profiles/r02_mma_microbench.md):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I paella_b200/csrc -I include -o tools/mma_microbench tools/mma_microbench.cu
// Each case issues `n_mma` MMAs (K = 16) back to back on `n_acc` accumulators in rotation, commits, waits; cycles / MMA =
// (clock after the wait - clock before the first issue) / n_mma, best of 4 repeats.  Operand bytes are zeros (timing only).
#include <cstdio>
#include <cstring>
#include <vector>
#include "ptx.cuh"

using namespace pb;

struct Case {
    int m, n, a_mn, b_mn, a_tmem, n_acc, a_sw32, b_sw32, n_mma;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) bench_kernel(const Case* cases, int n_cases, float* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    // [0, 64K) operand A region, [64K, 128K) operand B region, then barrier + tmem slot
    const uint32_t a_smem = base, b_smem = base + 65536, bar = base + 131072, slot = bar + 8;
    for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0u;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(gen + 131072 + 8);
    if (threadIdx.x == 0) {
        uint32_t phase = 0;
        for (int c = 0; c < n_cases; ++c) {
            const Case k = cases[c];
            const uint32_t idesc = ptx::umma_idesc_f16_major(k.m, k.n, k.a_mn, k.b_mn);
            // descriptors: K-major SW128 (SBO 1024), K-major SW32 (SBO 256), MN-major SW128 (second 64-wide atom 16 KB further)
            const uint64_t da = k.a_sw32 ? ptx::umma_desc_sw32(a_smem) : (k.a_mn ? ptx::umma_desc_mn_sw128(a_smem, 16384) : ptx::umma_desc_sw128(a_smem));
            const uint64_t db = k.b_sw32 ? ptx::umma_desc_sw32(b_smem) : (k.b_mn ? ptx::umma_desc_mn_sw128(b_smem, 16384) : ptx::umma_desc_sw128(b_smem));
            const uint32_t a_t = tmem + 384;             // A-in-TMEM operand: columns 384.. (fp16 pairs)
            const uint32_t acc_cols = (uint32_t)((k.n + 31) & ~31);
            long long best = 1ll << 60;
            for (int rep = 0; rep < 4; ++rep) {
                const uint32_t d0 = tmem, d1 = tmem + (k.n_acc > 1 ? acc_cols : 0u);       // no per-MMA address arithmetic: the loop
                const long long t0 = clock64();                                            // below must not be issue-bound
                if (k.a_tmem) {
#pragma unroll 4
                    for (int i = 0; i < k.n_mma; i += 2) { umma_f16_ts(d0, a_t, db, idesc, 1u); umma_f16_ts(d1, a_t, db, idesc, 1u); }
                } else {
#pragma unroll 4
                    for (int i = 0; i < k.n_mma; i += 2) { ptx::umma_f16(d0, da, db, idesc, 1u); ptx::umma_f16(d1, da, db, idesc, 1u); }
                }
                ptx::umma_commit(bar);
                ptx::mbar_wait(bar, phase);
                phase ^= 1u;
                const long long t1 = clock64();
                if (t1 - t0 < best) best = t1 - t0;
            }
            out[c] = (float)best / (float)k.n_mma;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
    std::vector<Case> cs;
    std::vector<const char*> names;
    auto add = [&](const char* nm, int m, int n, int a_mn, int b_mn, int a_tmem, int n_acc, int a32 = 0, int b32 = 0) {
        cs.push_back(Case{m, n, a_mn, b_mn, a_tmem, n_acc, a32, b32, 256});
        names.push_back(nm);
    };
    add("M128 N256 A:K  B:K  1 acc (GEMM tile, floor 128)", 128, 256, 0, 0, 0, 1);
    add("M128 N128 A:K  B:K  1 acc (floor 64)", 128, 128, 0, 0, 0, 1);
    add("M128 N64  A:K  B:K  1 acc (floor 32)", 128, 64, 0, 0, 0, 1);
    add("M128 N64  A:K  B:K  2 acc", 128, 64, 0, 0, 0, 2);
    add("M128 N64  A:K32 B:K32 1 acc (32-byte-swizzle tails)", 128, 64, 0, 0, 0, 1, 1, 1);
    add("M128 N64  A:MN B:MN 1 acc (attention_tt P V)", 128, 64, 1, 1, 0, 1);
    add("M128 N64  A:MN B:MN 2 acc", 128, 64, 1, 1, 0, 2);
    add("M128 N64  A:K  B:MN 1 acc", 128, 64, 0, 1, 0, 1);
    add("M128 N64  A:MN B:K  1 acc", 128, 64, 1, 0, 0, 1);
    add("M128 N128 A:MN B:K  1 acc", 128, 128, 1, 0, 0, 1);
    add("M128 N128 A:K  B:MN 1 acc", 128, 128, 0, 1, 0, 1);
    add("M128 N256 A:MN B:MN 1 acc", 128, 256, 1, 1, 0, 1);
    add("M64  N208 A:K  B:K  1 acc (attention_tc S, floor 104)", 64, 208, 0, 0, 0, 1);
    add("M64  N208 A:K  B:K  2 acc", 64, 208, 0, 0, 0, 2);
    add("M64  N64  A:K  B:MN 1 acc (attention_tc P V)", 64, 64, 0, 1, 0, 1);
    add("M64  N64  A:K  B:MN 2 acc", 64, 64, 0, 1, 0, 2);
    add("M64  N16  A:K  B:K32 1 acc", 64, 16, 0, 0, 0, 1, 0, 1);
    add("M64  N64  A:K  B:K  1 acc", 64, 64, 0, 0, 0, 1);
    add("M128 N64  A:TMEM B:MN 1 acc (P from TMEM)", 128, 64, 0, 1, 1, 1);
    add("M128 N64  A:TMEM B:K  1 acc", 128, 64, 0, 0, 1, 1);
    add("M128 N128 A:TMEM B:MN 1 acc", 128, 128, 0, 1, 1, 1);
    add("M64  N64  A:TMEM B:MN 1 acc", 64, 64, 0, 1, 1, 1);
    add("M64  N64  A:TMEM B:MN 2 acc", 64, 64, 0, 1, 1, 2);
    add("M64  N128 A:TMEM B:MN 1 acc", 64, 128, 0, 1, 1, 1);
    add("M64  N16  A:TMEM B:K32 1 acc", 64, 16, 0, 0, 1, 1, 0, 1);
    add("M128 N16  A:K  B:K  1 acc", 128, 16, 0, 0, 0, 1);
    add("M128 N32  A:K  B:K  1 acc", 128, 32, 0, 0, 0, 1);
    Case* d_cases;
    float* d_out;
    cudaMalloc(&d_cases, cs.size() * sizeof(Case));
    cudaMalloc(&d_out, cs.size() * sizeof(float));
    cudaMemcpy(d_cases, cs.data(), cs.size() * sizeof(Case), cudaMemcpyHostToDevice);
    const int smem = 131072 + 64 + 1024;
    cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("| case | cycles / MMA (K = 16) | floor max(M,128)*N/256 |\n|---|---:|---:|\n");
    for (size_t i = 0; i < cs.size(); ++i) {          // one launch per case (an illegal combination only loses the cases after it)
        for (int pass = 0; pass < 2; ++pass) {        // second pass = warm
            bench_kernel<<<1, 128, smem>>>(d_cases + i, 1, d_out + i);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("| %s | CUDA error: %s | |\n", names[i], cudaGetErrorString(e)); return 1; }
        }
        float v = 0.f;
        cudaMemcpy(&v, d_out + i, sizeof(float), cudaMemcpyDeviceToHost);
        printf("| %s | %.1f | %d |\n", names[i], v, (cs[i].m > 128 ? cs[i].m : 128) * cs[i].n / 256);
        fflush(stdout);
    }
    return 0;
}
