// Handle life cycle of libboxfusion_sm100.so.  No CPU fallback: creation fails unless the device is sm_100.
#include "bf_common.cuh"
#include <stdlib.h>

extern "C" int bf_version(void) { return 100; }
extern "C" int bf_fusion_cap(void) { return BF_FUSION_CAP; }

static char g_create_err[512] = "";

extern "C" int bf_create(int device, bf_handle** out) {
    if (!out) return BF_ERR_INVALID_ARG;
    *out = nullptr;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
        return BF_ERR_CUDA;
    }
    if (prop.major != 10) {
        snprintf(g_create_err, sizeof(g_create_err), "device %d is sm_%d%d; this library is built for sm_100a only",
                 device, prop.major, prop.minor);
        return BF_ERR_CUDA;
    }
    bf_handle* h = (bf_handle*)calloc(1, sizeof(bf_handle));
    if (!h) return BF_ERR_INVALID_ARG;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    *out = h;
    return BF_OK;
}

extern "C" void bf_destroy(bf_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int i = 0; i < BF_SCRATCH_SLOTS; ++i)
        if (h->buf[i]) cudaFree(h->buf[i]);
    free(h);
}

extern "C" const char* bf_last_error(bf_handle* h) { return h ? h->err : g_create_err; }
