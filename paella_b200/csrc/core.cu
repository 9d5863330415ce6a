// Error plumbing and device queries shared by the whole library.
#include "common.cuh"
#include "paella_b200.h"

#include <mutex>

namespace pb {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }

static int g_sm = -1, g_tpsm = -1;
static void query_device() {
    if (g_sm >= 0) return;
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
        g_sm = p.multiProcessorCount;
        g_tpsm = p.maxThreadsPerMultiProcessor;
    } else {
        (void)cudaGetLastError();
        g_sm = 0;
        g_tpsm = 2048;
    }
}
int sm_count() { query_device(); return g_sm; }
int max_threads_per_sm() { query_device(); return g_tpsm; }

}  // namespace pb

extern "C" {
const char* pb200_last_error(void) { return pb::last_error(); }
int pb200_abi_version(void) { return PB200_ABI_VERSION; }
int pb200_device_info(int* sms, int* tpsm) {
    PB_CHECK(pb::sm_count() > 0, "no CUDA device");
    if (sms) *sms = pb::sm_count();
    if (tpsm) *tpsm = pb::max_threads_per_sm();
    return 0;
}
}
