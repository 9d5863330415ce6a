// Vector quantiser: nearest-code search and index->vector gather.
// Replaces torchtools.nn.VectorQuantize.forward / idx2vq as called at ref/src/vqgan.py:94,104 and in
// the notebook's `quant` sampling mode.  Arithmetic (bit-for-bit the same as oracle/vq_nearest.c):
//   dot = fma chain over j;  c2, x2 = fma chains;  s = c2 + x2;  d = fma(-2, dot, s);  first minimum.
// HBM-bound on paper (16 B in, 8 B out per vector) but FFMA-bound in practice: 8192 codes x C fmas per
// vector.  The codebook is staged through shared memory in chunks and read as warp-wide broadcasts.
#include "common.cuh"
#include "paella_b200.h"

namespace pb {

constexpr int VQ_THREADS = 256;
constexpr int VQ_VPT = 2;          // vectors per thread (ILP)
constexpr int VQ_CHUNK = 1024;     // codes per shared-memory chunk

template <int C>
__global__ void __launch_bounds__(VQ_THREADS) vq_nearest_kernel(const float* __restrict__ x, int64_t n,
                                                                const float* __restrict__ cb, int k,
                                                                int64_t* __restrict__ idx) {
    __shared__ float s_cb[VQ_CHUNK * C];
    __shared__ float s_c2[VQ_CHUNK];
    float xv[VQ_VPT][C], x2[VQ_VPT], best[VQ_VPT];
    int bi[VQ_VPT];
    int64_t vid[VQ_VPT];
#pragma unroll
    for (int v = 0; v < VQ_VPT; ++v) {
        vid[v] = ((int64_t)blockIdx.x * VQ_VPT + v) * VQ_THREADS + threadIdx.x;
        x2[v] = 0.f;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            xv[v][j] = vid[v] < n ? x[vid[v] * C + j] : 0.f;
            x2[v] = fmaf(xv[v][j], xv[v][j], x2[v]);
        }
        best[v] = INFINITY;
        bi[v] = 0;
    }
    for (int k0 = 0; k0 < k; k0 += VQ_CHUNK) {
        const int kc = min(VQ_CHUNK, k - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kc * C; i += VQ_THREADS) s_cb[i] = cb[(int64_t)k0 * C + i];
        __syncthreads();
        for (int i = threadIdx.x; i < kc; i += VQ_THREADS) {
            float c2 = 0.f;
#pragma unroll
            for (int j = 0; j < C; ++j) c2 = fmaf(s_cb[i * C + j], s_cb[i * C + j], c2);
            s_c2[i] = c2;
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < kc; ++i) {
            float cv[C];
#pragma unroll
            for (int j = 0; j < C; ++j) cv[j] = s_cb[i * C + j];
            const float c2 = s_c2[i];
#pragma unroll
            for (int v = 0; v < VQ_VPT; ++v) {
                float dot = 0.f;
#pragma unroll
                for (int j = 0; j < C; ++j) dot = fmaf(xv[v][j], cv[j], dot);
                const float s = __fadd_rn(c2, x2[v]);
                const float d = fmaf(-2.0f, dot, s);
                if (d < best[v]) {
                    best[v] = d;
                    bi[v] = k0 + i;
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VQ_VPT; ++v)
        if (vid[v] < n) idx[vid[v]] = bi[v];
}

__global__ void vq_gather_kernel(const int64_t* __restrict__ idx, int64_t n, const float* __restrict__ cb, int k, int c,
                                 float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * c) return;
    const int64_t i = e / c;
    const int j = (int)(e - i * c);
    int64_t code = idx[i];
    code = code < 0 ? 0 : (code >= k ? k - 1 : code);
    out[e] = cb[code * c + j];
}

}  // namespace pb

using namespace pb;

extern "C" {

int pb200_vq_nearest(const float* x, int64_t n, int c, const float* codebook, int k, int64_t* idx, void* stream) {
    PB_CHECK(c >= 1 && c <= 8, "vq_nearest: c_latent %d unsupported (1..8)", c);
    PB_CHECK(k >= 1, "vq_nearest: empty codebook");
    if (n == 0) return 0;
    const int grid = ceil_div(n, VQ_THREADS * VQ_VPT);
    cudaStream_t st = (cudaStream_t)stream;
    switch (c) {
#define PB_VQ_CASE(C) \
    case C: vq_nearest_kernel<C><<<grid, VQ_THREADS, 0, st>>>(x, n, codebook, k, idx); break;
        PB_VQ_CASE(1) PB_VQ_CASE(2) PB_VQ_CASE(3) PB_VQ_CASE(4) PB_VQ_CASE(5) PB_VQ_CASE(6) PB_VQ_CASE(7) PB_VQ_CASE(8)
#undef PB_VQ_CASE
    }
    PB_LAUNCH_CHECK();
    return 0;
}

int pb200_vq_gather(const int64_t* idx, int64_t n, const float* codebook, int k, int c, float* out, void* stream) {
    if (n == 0) return 0;
    vq_gather_kernel<<<ceil_div(n * c, 256), 256, 0, (cudaStream_t)stream>>>(idx, n, codebook, k, c, out);
    PB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
