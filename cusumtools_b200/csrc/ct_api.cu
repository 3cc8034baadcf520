// C-ABI plumbing: error string, version, launch counter, device query.
#include "ct_common.cuh"
#include "cusumtools_b200.h"
#include <stdarg.h>

unsigned long long g_ct_launches = 0;
static thread_local char g_err[512] = "";

void ct_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ct_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ct_set_error("%s: %s", what, cudaGetErrorString(e));
        return CT_ERR_CUDA;
    }
    return CT_OK;
}

static int g_sm = 0, g_smem = 0;
static void query() {
    if (g_sm) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (g_sm <= 0) g_sm = 148;
}
int ct_sm_count() { query(); return g_sm; }
int ct_max_smem_optin() { query(); return g_smem; }

extern "C" {
int ct_version(void) { return CT_ABI_VERSION; }
const char* ct_last_error(void) { return g_err; }
uint64_t ct_launch_count(void) { return g_ct_launches; }
void ct_launch_count_reset(void) { g_ct_launches = 0; }
int ct_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* smem_optin) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { ct_set_error("no CUDA device"); return CT_ERR_CUDA; }
    query();
    if (sm_count) *sm_count = g_sm;
    if (cc_major) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev); *cc_major = v; }
    if (cc_minor) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev); *cc_minor = v; }
    if (smem_optin) *smem_optin = g_smem;
    return CT_OK;
}
}
