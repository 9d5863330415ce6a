/* CPU oracle (plain C) for the vector-quantiser nearest-code search.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/libvq_oracle.so.
 *
 * Restates the nearest-codebook lookup of torchtools.nn.VectorQuantize as used at
 * ref/src/vqgan.py:94 (third-party, unpinned, absent from /root/reference ->
 * PARITY UNPINNED; see oracle/vqgan_oracle.py).  The arithmetic below is the exact
 * operation order of paella_b200/csrc/vq.cu so that indices can be compared
 * bit-for-bit:
 *     dot = fma chain over j;  c2, x2 = fma chains;  s = c2 + x2;  d = fma(-2, dot, s)
 *     index = first minimum (strict <).
 * Compile with -ffp-contract=off so only the explicit fmaf() calls fuse.
 */
#include <math.h>
#include <stdint.h>

void vq_nearest_f32(const float* x, int64_t n, int c, const float* codebook, int k, int64_t* idx_out,
                    float* dist_out /* may be NULL: [n] best distance */) {
    for (int64_t i = 0; i < n; ++i) {
        const float* xi = x + i * c;
        float x2 = 0.f;
        for (int j = 0; j < c; ++j) x2 = fmaf(xi[j], xi[j], x2);
        float best = INFINITY;
        int64_t bi = 0;
        for (int kk = 0; kk < k; ++kk) {
            const float* ck = codebook + (int64_t)kk * c;
            float c2 = 0.f, dot = 0.f;
            for (int j = 0; j < c; ++j) c2 = fmaf(ck[j], ck[j], c2);
            for (int j = 0; j < c; ++j) dot = fmaf(xi[j], ck[j], dot);
            float s = c2 + x2;
            float d = fmaf(-2.0f, dot, s);
            if (d < best) { best = d; bi = kk; }
        }
        idx_out[i] = bi;
        if (dist_out) dist_out[i] = best;
    }
}

/* Philox4x32-10, same constants as curand_philox4x32_x.h — used to cross-check oracle/philox.py. */
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
