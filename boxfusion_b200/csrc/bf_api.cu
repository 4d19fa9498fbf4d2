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
    { const char* t = getenv("BF_REFINE_TIMING"); h->refine_timing = (t && (t[0] == '1' || t[0] == '2')) ? (t[0] - '0') : 0; }
    { const char* t = getenv("BF_REFINE_PERSISTENT"); h->refine_force_persistent = t ? atoi(t) : 0; }
    { const char* t = getenv("BF_REFINE_SHAPE"); h->refine_force_variant = -1; h->refine_force_mode = -1; if (t) sscanf(t, "%d,%d,%d,%d", &h->refine_force_c, &h->refine_force_t, &h->refine_force_variant, &h->refine_force_mode); }
    *out = h;
    return BF_OK;
}

extern "C" int bf_set_option(bf_handle* h, int key, int value) {
    if (!h) return BF_ERR_INVALID_ARG;
    if (key == BF_OPT_REFINE_CONCURRENT) { h->refine_concurrent = value ? 1 : 0; return BF_OK; }
    if (key == BF_OPT_REFINE_PERSISTENT) { h->refine_force_persistent = value > 0 ? value : 0; return BF_OK; }
    return bf_fail(h, BF_ERR_INVALID_ARG, "bf_set_option", "unknown key");
}

extern "C" void bf_destroy(bf_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int i = 0; i < BF_SCRATCH_SLOTS; ++i)
        if (h->buf[i]) cudaFree(h->buf[i]);
    free(h);
}

extern "C" const char* bf_last_error(bf_handle* h) { return h ? h->err : g_create_err; }

// ---- diagnostic: measured FP32 FMA throughput of this device (the denominator of the FP32-pipe roofline;
// "measure with an FMA micro-benchmark on the box").  Synchronous.
__global__ void __launch_bounds__(256) bf_fma_probe_kernel(float* __restrict__ out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 0.001f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int bf_probe_fp32(bf_handle* h, int iters, double* tflops_out /*host*/, float* ms_out /*host*/) {
    if (!h || iters < 1 || !tflops_out) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_probe_fp32", "bad argument");
    bf_device_guard guard(h);
    const int blocks = h->sm_count * 8, threads = 256;
    void* p;
    int rc = bf_scratch(h, BF_SCRATCH_MISC, sizeof(float) * (size_t)blocks * threads, &p);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    BF_CUDA(h, cudaEventCreate(&e0));
    BF_CUDA(h, cudaEventCreate(&e1));
    bf_fma_probe_kernel<<<blocks, threads>>>((float*)p, iters);          // warm-up
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        BF_CUDA(h, cudaEventRecord(e0));
        bf_fma_probe_kernel<<<blocks, threads>>>((float*)p, iters);
        BF_CUDA(h, cudaEventRecord(e1));
        BF_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0.f;
        BF_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return BF_OK;
}

__global__ void __launch_bounds__(256) bf_dfma_probe_kernel(double* __restrict__ out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double b = 0.999, c = 0.001;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int bf_probe_fp64(bf_handle* h, int iters, double* tflops_out /*host*/, float* ms_out /*host*/) {
    if (!h || iters < 1 || !tflops_out) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_probe_fp64", "bad argument");
    bf_device_guard guard(h);
    const int blocks = h->sm_count * 8, threads = 256;
    void* p;
    int rc = bf_scratch(h, BF_SCRATCH_MISC, sizeof(double) * (size_t)blocks * threads, &p);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    BF_CUDA(h, cudaEventCreate(&e0));
    BF_CUDA(h, cudaEventCreate(&e1));
    bf_dfma_probe_kernel<<<blocks, threads>>>((double*)p, iters);          // warm-up
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        BF_CUDA(h, cudaEventRecord(e0));
        bf_dfma_probe_kernel<<<blocks, threads>>>((double*)p, iters);
        BF_CUDA(h, cudaEventRecord(e1));
        BF_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0.f;
        BF_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return BF_OK;
}
