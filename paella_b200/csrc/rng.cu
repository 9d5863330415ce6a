// Random streams + the resample step of the sampling loop, reproducing PyTorch's CUDA
// generator (Philox4x32-10) element for element.
//   torch.randint        ref/src/utils.py:37
//   torch.multinomial    ref/src/utils.py:49-50   (q = exponential_(1); argmax(p / q))
//   CFG/temp/softmax     ref/src/utils.py:45-47
//   Paella.add_noise     ref/src/modules.py:277-283
// PyTorch-side arithmetic: ATen/native/cuda/DistributionTemplates.h, ATen/core/TransformationHelper.h.
#include "common.cuh"
#include "paella_b200.h"

#include <float.h>

namespace pb {

TorchPhilox make_torch_philox(uint64_t seed, uint64_t offset, long numel) {
    const long block = 256;
    long grid = (numel + block - 1) / block;
    const long cap = (long)sm_count() * (max_threads_per_sm() / block);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    TorchPhilox s;
    s.seed = seed;
    s.offset4 = offset / 4;
    s.stride = (uint32_t)(block * grid);
    return s;
}

static int64_t offset_increment(int64_t numel) {
    if (numel <= 0) return 0;
    TorchPhilox s = make_torch_philox(0, 0, numel);
    return ((numel - 1) / ((int64_t)s.stride * 4) + 1) * 4;
}

// ------------------------------------------------------------------ randint / rand
__global__ void randint_kernel(int64_t* __restrict__ out, int64_t numel, uint32_t range, TorchPhilox s) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= numel) return;
    out[e] = (int64_t)(torch_philox_u32(s, (uint64_t)e) % range);
}

__global__ void rand_kernel(float* __restrict__ out, int64_t numel, TorchPhilox s) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= numel) return;
    const float u = u32_to_uniform(torch_philox_u32(s, (uint64_t)e));
    out[e] = (u == 1.0f) ? 0.0f : u;     // uniform_kernel: value == to ? from : value
}

// ------------------------------------------------------------------ add_noise
__global__ void add_noise_kernel(const int64_t* __restrict__ x, const int64_t* __restrict__ random_x,
                                 const float* __restrict__ t, int64_t batch, int64_t hw, uint32_t num_labels,
                                 TorchPhilox s_mask, TorchPhilox s_rx, int64_t* __restrict__ out,
                                 int64_t* __restrict__ mask_out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * hw) return;
    float u = u32_to_uniform(torch_philox_u32(s_mask, (uint64_t)e));
    u = (u == 1.0f) ? 0.0f : u;
    const bool m = u <= t[e / hw];
    const int64_t rx = random_x ? random_x[e] : (int64_t)(torch_philox_u32(s_rx, (uint64_t)e) % num_labels);
    out[e] = m ? rx : x[e];
    if (mask_out) mask_out[e] = m ? 1 : 0;
}

// ------------------------------------------------------------------ multinomial(p, 1): argmax p/q, first index on ties
struct ArgBest {
    float v;
    int idx;
};
__device__ __forceinline__ ArgBest better(ArgBest a, ArgBest b) {
    // torch argmax: larger value wins, lower index on ties (inputs are NaN-free: torch asserts p valid)
    const bool take_b = (b.v > a.v) || (b.v == a.v && b.idx < a.idx);
    return take_b ? b : a;
}
__device__ __forceinline__ ArgBest warp_best(ArgBest a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgBest b;
        b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
        b.idx = __shfl_xor_sync(0xffffffffu, a.idx, o);
        a = better(a, b);
    }
    return a;
}

__global__ void __launch_bounds__(256) multinomial_kernel(const float* __restrict__ p, int64_t rows, int k, TorchPhilox s,
                                                          int64_t* __restrict__ out) {
    const int64_t row = blockIdx.x;
    const float* pr = p + row * k;
    ArgBest best{-INFINITY, 0x7fffffff};
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const uint64_t e = (uint64_t)row * (uint64_t)k + (uint64_t)j;
        const float q = torch_exponential1(u32_to_uniform(torch_philox_u32(s, e)));
        const float v = __fdiv_rn(pr[j], q);
        best = better(best, ArgBest{v, j});
    }
    __shared__ float sv[8];
    __shared__ int si[8];
    best = warp_best(best);
    if ((threadIdx.x & 31) == 0) {
        sv[threadIdx.x >> 5] = best.v;
        si[threadIdx.x >> 5] = best.idx;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        ArgBest b{-INFINITY, 0x7fffffff};
        if (threadIdx.x < (blockDim.x >> 5)) b = ArgBest{sv[threadIdx.x], si[threadIdx.x]};
        b = warp_best(b);
        if (threadIdx.x == 0) out[row] = b.idx;
    }
}

// ------------------------------------------------------------------ resample on reference-layout logits
// One CTA = 32 consecutive positions p of one sample (coalesced along HW) x all K, 8 warps stride K.
//   l  = lc*cfg + lu*(1-cfg)          (two rounded products and a rounded sum, like the three torch kernels)
//   l' = l * (1/T)                    (torch's div-by-CPU-scalar fast path multiplies by the reciprocal)
//   p  = exp(l' - max) / sum          (softmax dim=1)
//   tok = argmax_k p_k / q_k
template <int MODE>
__global__ void __launch_bounds__(256) resample_logits_kernel(const float* __restrict__ lc, const float* __restrict__ lu,
                                                              int k, int64_t hw, float cfg, float one_minus_cfg,
                                                              float inv_t, TorchPhilox s, int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = blockIdx.y;
    const int64_t pos = (int64_t)blockIdx.x * 32 + lane;
    const bool valid = pos < hw;
    const float* c_ptr = lc + b * (int64_t)k * hw + (valid ? pos : 0);
    const float* u_ptr = lu ? lu + b * (int64_t)k * hw + (valid ? pos : 0) : nullptr;
    __shared__ float red[8][33];
    __shared__ int redi[8][33];

    auto mixed = [&](int j) -> float {
        float l = c_ptr[(int64_t)j * hw];
        if (u_ptr) l = __fadd_rn(__fmul_rn(l, cfg), __fmul_rn(u_ptr[(int64_t)j * hw], one_minus_cfg));
        return l;
    };

    if (MODE == 1) {   // argmax of the (guided) logits
        ArgBest best{-INFINITY, 0x7fffffff};
        for (int j = warp; j < k; j += 8) best = better(best, ArgBest{mixed(j), j});
        red[warp][lane] = best.v;
        redi[warp][lane] = best.idx;
        __syncthreads();
        if (warp == 0) {
            ArgBest bb{red[0][lane], redi[0][lane]};
            for (int w = 1; w < 8; ++w) bb = better(bb, ArgBest{red[w][lane], redi[w][lane]});
            if (valid) out[b * hw + pos] = bb.idx;
        }
        return;
    }

    // pass 1: max
    float m = -INFINITY;
    for (int j = warp; j < k; j += 8) m = fmaxf(m, __fmul_rn(mixed(j), inv_t));
    red[warp][lane] = m;
    __syncthreads();
    m = red[0][lane];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w][lane]);
    __syncthreads();
    // pass 2: sum of exp
    float sum = 0.f;
    for (int j = warp; j < k; j += 8) sum += expf(__fmul_rn(mixed(j), inv_t) - m);
    red[warp][lane] = sum;
    __syncthreads();
    sum = red[0][lane];
    for (int w = 1; w < 8; ++w) sum += red[w][lane];
    __syncthreads();
    // pass 3: argmax p/q
    ArgBest best{-INFINITY, 0x7fffffff};
    if (valid) {
        const uint64_t row = (uint64_t)(b * hw + pos);
        for (int j = warp; j < k; j += 8) {
            const float pj = __fdiv_rn(expf(__fmul_rn(mixed(j), inv_t) - m), sum);
            const float q = torch_exponential1(u32_to_uniform(torch_philox_u32(s, row * (uint64_t)k + (uint64_t)j)));
            best = better(best, ArgBest{__fdiv_rn(pj, q), j});
        }
    }
    red[warp][lane] = best.v;
    redi[warp][lane] = best.idx;
    __syncthreads();
    if (warp == 0 && valid) {
        ArgBest bb{red[0][lane], redi[0][lane]};
        for (int w = 1; w < 8; ++w) bb = better(bb, ArgBest{red[w][lane], redi[w][lane]});
        out[b * hw + pos] = bb.idx;
    }
}

// ------------------------------------------------------------------ `quant` sampling mode (notebook cell 3, ref/src_distributed/train.py:155-156)
//   e = softmax(l / T) @ codebook  [C values per token];  token = nearest code of e   (deterministic, no draw)
// Same CTA shape as resample_logits_kernel: 32 positions x 8 warps striding the labels.
template <int C>
__global__ void __launch_bounds__(256) resample_quant_kernel(const float* __restrict__ lc, const float* __restrict__ lu, int k,
                                                             int64_t hw, float cfg, float one_minus_cfg, float inv_t,
                                                             const float* __restrict__ codebook, int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = blockIdx.y;
    const int64_t pos = (int64_t)blockIdx.x * 32 + lane;
    const bool valid = pos < hw;
    const float* c_ptr = lc + b * (int64_t)k * hw + (valid ? pos : 0);
    const float* u_ptr = lu ? lu + b * (int64_t)k * hw + (valid ? pos : 0) : nullptr;
    __shared__ float red[8][33];
    __shared__ float redc[C][8][33];
    __shared__ int redi[8][33];
    auto mixed = [&](int j) -> float {
        float l = c_ptr[(int64_t)j * hw];
        if (u_ptr) l = __fadd_rn(__fmul_rn(l, cfg), __fmul_rn(u_ptr[(int64_t)j * hw], one_minus_cfg));
        return __fmul_rn(l, inv_t);
    };
    float m = -INFINITY;
    for (int j = warp; j < k; j += 8) m = fmaxf(m, mixed(j));
    red[warp][lane] = m;
    __syncthreads();
    m = red[0][lane];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w][lane]);
    __syncthreads();
    float sum = 0.f, acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int j = warp; j < k; j += 8) {
        const float e = expf(mixed(j) - m);
        sum += e;
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(e, __ldg(codebook + (int64_t)j * C + c), acc[c]);
    }
    red[warp][lane] = sum;
#pragma unroll
    for (int c = 0; c < C; ++c) redc[c][warp][lane] = acc[c];
    __syncthreads();
    sum = 0.f;
    float x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = 0.f;
    for (int w = 0; w < 8; ++w) {
        sum += red[w][lane];
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] += redc[c][w][lane];
    }
    float x2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { x[c] = __fdiv_rn(x[c], sum); x2 = fmaf(x[c], x[c], x2); }
    // nearest code (same arithmetic as vq_nearest_kernel), codes strided over the 8 warps
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int j = warp; j < k; j += 8) {
        float c2 = 0.f, dot = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float cv = __ldg(codebook + (int64_t)j * C + c);
            c2 = fmaf(cv, cv, c2);
            dot = fmaf(x[c], cv, dot);
        }
        const float d = fmaf(-2.0f, dot, __fadd_rn(c2, x2));
        if (d < best) { best = d; bi = j; }
    }
    __syncthreads();
    red[warp][lane] = best;
    redi[warp][lane] = bi;
    __syncthreads();
    if (warp == 0 && valid) {
        float bd = red[0][lane];
        int bj = redi[0][lane];
        for (int w = 1; w < 8; ++w)
            if (red[w][lane] < bd || (red[w][lane] == bd && redi[w][lane] < bj)) { bd = red[w][lane]; bj = redi[w][lane]; }
        out[b * hw + pos] = bj;
    }
}

}  // namespace pb

using namespace pb;

extern "C" {

int64_t pb200_philox_offset_increment(int64_t numel) { return offset_increment(numel); }

int pb200_randint(int64_t* out, int64_t numel, int64_t num_labels, uint64_t seed, uint64_t offset, void* stream) {
    PB_CHECK(num_labels > 0 && num_labels < (1ll << 28), "randint: range %lld needs the 64-bit path (unsupported)",
             (long long)num_labels);
    PB_CHECK(offset % 4 == 0, "philox offset must be a multiple of 4");
    if (numel == 0) return 0;
    TorchPhilox s = make_torch_philox(seed, offset, numel);
    randint_kernel<<<ceil_div(numel, 256), 256, 0, (cudaStream_t)stream>>>(out, numel, (uint32_t)num_labels, s);
    PB_LAUNCH_CHECK();
    return 0;
}

int pb200_rand(float* out, int64_t numel, uint64_t seed, uint64_t offset, void* stream) {
    PB_CHECK(offset % 4 == 0, "philox offset must be a multiple of 4");
    if (numel == 0) return 0;
    TorchPhilox s = make_torch_philox(seed, offset, numel);
    rand_kernel<<<ceil_div(numel, 256), 256, 0, (cudaStream_t)stream>>>(out, numel, s);
    PB_LAUNCH_CHECK();
    return 0;
}

int pb200_multinomial(const float* p, int64_t rows, int64_t k, uint64_t seed, uint64_t offset, int64_t* out,
                      void* stream) {
    PB_CHECK(offset % 4 == 0, "philox offset must be a multiple of 4");
    PB_CHECK(k > 0 && k < (1ll << 30), "multinomial: bad category count");
    PB_CHECK(rows * k < (1ll << 31), "multinomial: rows*k >= 2^31 would split the torch kernel (unsupported)");
    if (rows == 0) return 0;
    TorchPhilox s = make_torch_philox(seed, offset, rows * k);
    multinomial_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(p, rows, (int)k, s, out);
    PB_LAUNCH_CHECK();
    return 0;
}

int pb200_resample_logits(const float* logits_c, const float* logits_u, int64_t batch, int64_t k, int64_t hw,
                          double cfg, double temperature, int mode, uint64_t seed, uint64_t offset, int64_t* out,
                          void* stream) {
    PB_CHECK(offset % 4 == 0, "philox offset must be a multiple of 4");
    PB_CHECK(mode == 0 || mode == 1, "resample: mode must be 0 (multinomial) or 1 (argmax)");
    PB_CHECK(batch * hw * k < (1ll << 31), "resample: B*HW*K >= 2^31 would split the torch kernel (unsupported)");
    if (batch == 0 || hw == 0) return 0;
    TorchPhilox s = make_torch_philox(seed, offset, batch * hw * k);
    // python scalars: `logits * cfg` and `(1 - cfg)` are doubles cast to the fp32 op-math type;
    // `temperatures[i]` is an fp32 0-dim CPU tensor and torch multiplies by its fp32 reciprocal.
    const float cfg_f = (float)cfg;
    const float one_minus = (float)(1.0 - cfg);
    const float inv_t = 1.0f / (float)temperature;
    dim3 grid(ceil_div(hw, 32), (unsigned)batch);
    if (mode == 0)
        resample_logits_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(logits_c, logits_u, (int)k, hw, cfg_f, one_minus,
                                                                          inv_t, s, out);
    else
        resample_logits_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(logits_c, logits_u, (int)k, hw, cfg_f, one_minus,
                                                                          inv_t, s, out);
    PB_LAUNCH_CHECK();
    return 0;
}

int pb200_resample_quant(const float* logits_c, const float* logits_u, int64_t batch, int64_t k, int64_t hw, double cfg,
                         double temperature, const float* codebook, int c_latent, int64_t* out, void* stream) {
    PB_CHECK(c_latent >= 1 && c_latent <= 8, "resample_quant: c_latent %d unsupported (1..8)", c_latent);
    if (batch == 0 || hw == 0) return 0;
    const float cfg_f = (float)cfg, one_minus = (float)(1.0 - cfg), inv_t = 1.0f / (float)temperature;
    dim3 grid(ceil_div(hw, 32), (unsigned)batch);
    cudaStream_t st = (cudaStream_t)stream;
    switch (c_latent) {
#define PB_RQ_CASE(C) \
    case C: resample_quant_kernel<C><<<grid, 256, 0, st>>>(logits_c, logits_u, (int)k, hw, cfg_f, one_minus, inv_t, codebook, out); break;
        PB_RQ_CASE(1) PB_RQ_CASE(2) PB_RQ_CASE(3) PB_RQ_CASE(4) PB_RQ_CASE(5) PB_RQ_CASE(6) PB_RQ_CASE(7) PB_RQ_CASE(8)
#undef PB_RQ_CASE
    }
    PB_LAUNCH_CHECK();
    return 0;
}

int pb200_add_noise(const int64_t* x, const int64_t* random_x, const float* t, int64_t batch, int64_t hw,
                    int64_t num_labels, uint64_t seed, uint64_t offset, int64_t* out, int64_t* mask_out,
                    void* stream) {
    PB_CHECK(offset % 4 == 0, "philox offset must be a multiple of 4");
    PB_CHECK(num_labels > 0 && num_labels < (1ll << 28), "add_noise: bad num_labels");
    const int64_t n = batch * hw;
    if (n == 0) return 0;
    TorchPhilox s_mask = make_torch_philox(seed, offset, n);
    TorchPhilox s_rx = make_torch_philox(seed, offset + offset_increment(n), n);
    add_noise_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, random_x, t, batch, hw, (uint32_t)num_labels,
                                                                         s_mask, s_rx, out, mask_out);
    PB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
