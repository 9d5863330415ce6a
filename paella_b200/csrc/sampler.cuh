// Fused out_mapper + Gumbel-max draw (see sampler.cu).
#pragma once
#include "common.cuh"

namespace pb {
// rows the fp16 feature buffer must be able to hold for R token rows (whole 4*rs-row Philox blocks)
int64_t fused_sampler_rows_padded(int64_t R, int NL);
// a16: fp16 [R, Kc] guided features; w16: fp16 [NL, Kc] out_mapper weight; out: int64 [R]
int launch_fused_sampler(const __half* a16, int64_t R, int Kc, const __half* w16, int NL, float inv_t, uint64_t seed,
                         uint64_t offset, int64_t* out, cudaStream_t st);
}  // namespace pb
