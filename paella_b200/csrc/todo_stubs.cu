// Entry points declared in include/paella_b200.h whose implementation has not landed yet.
// Each reports an error (there is no fallback); this file shrinks to nothing as the path is completed.
#include "common.cuh"
#include "paella_b200.h"

#define PB_TODO(name) pb::set_error(std::string(name) + ": not implemented yet"); return 1

extern "C" {
int pb200_vqgan_create(const pb200_vqgan_config*, pb200_vqgan**) { PB_TODO("pb200_vqgan_create"); }
void pb200_vqgan_destroy(pb200_vqgan*) {}
int64_t pb200_vqgan_weight_bytes(const pb200_vqgan*) { return 0; }
int pb200_vqgan_bind_weights(pb200_vqgan*, void*) { PB_TODO("pb200_vqgan_bind_weights"); }
int pb200_vqgan_num_params(const pb200_vqgan*) { return 0; }
const char* pb200_vqgan_param_name(const pb200_vqgan*, int) { return ""; }
int64_t pb200_vqgan_param_numel(const pb200_vqgan*, int) { return 0; }
int pb200_vqgan_load_param(pb200_vqgan*, const char*, const float*, int64_t, void*) { PB_TODO("pb200_vqgan_load_param"); }
int64_t pb200_vqgan_workspace_bytes(const pb200_vqgan*, int, int, int) { return 0; }
int pb200_vqgan_encode(pb200_vqgan*, const float*, int, int, int, float*, float*, int64_t*, void*, int64_t, void*) { PB_TODO("pb200_vqgan_encode"); }
int pb200_vqgan_decode(pb200_vqgan*, const int64_t*, const float*, int, int, int, float*, void*, int64_t, void*) { PB_TODO("pb200_vqgan_decode"); }
}
