// Shared host-side plumbing of libboxfusion_sm100.so: the opaque handle, scratch arenas, error text.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/boxfusion_b200.h"

enum {
    BF_SCRATCH_PLANES_A = 0,  // double [M,12,4]
    BF_SCRATCH_PLANES_B,      // double [N,12,4]
    BF_SCRATCH_AABB_A,        // float  [M,6]
    BF_SCRATCH_AABB_B,        // float  [N,6]
    BF_SCRATCH_WORK,          // int2   work list of gate-passing pairs
    BF_SCRATCH_COUNTERS,      // int64  [8]
    BF_SCRATCH_MASK,          // uint32 NMS bit mask [N, W]
    BF_SCRATCH_RANK,          // int32  [N]
    BF_SCRATCH_IOU,           // double per work item
    BF_SCRATCH_MISC,
    BF_SCRATCH_REFINE,        // refine fitness / contributions
    BF_SCRATCH_EDGES,         // uint64 over-threshold edges (rank_lo << 32 | rank_hi) of bf_nms3d
    BF_SCRATCH_SLOTS
};

struct bf_handle {
    int device;
    int sm_count;
    char err[512];
    void* buf[BF_SCRATCH_SLOTS];
    size_t cap[BF_SCRATCH_SLOTS];
    int last_refine_cluster;    // cluster size * 1000 + block size of the last bf_refine launch (diagnostic)
    int refine_occ[20];         // cached cudaOccupancyMaxActiveClusters answers per (cluster size, block size)
    long long refine_occ_smem[20];   // dynamic shared memory (+1) the cached answer was computed for
    int refine_force_mode;      // 4th field of BF_REFINE_SHAPE: -1 = automatic, 0 = one (view, particle) term per thread pass, 1 = one thread per particle
    int refine_force_c, refine_force_t, refine_force_variant;   // BF_REFINE_SHAPE="C,T[,variant]" in the environment: force the launch shape / kernel instantiation (tuning sweeps)
    int refine_concurrent;      // BF_OPT_REFINE_CONCURRENT
    int refine_timing;          // BF_REFINE_TIMING=1 in the environment: bf_refine's trace carries per-iteration phase cycle counts (diagnostic)
    int frozen;                 // scratch may not grow any more: a captured CUDA graph (bf_engine) holds the pointers
    int refine_force_persistent;   // BF_REFINE_PERSISTENT=G in the environment: stand-alone bf_refine launches G persistent clusters (the engine's shape; tests)
};

// A size that is either known on the host or read from device memory by the kernel itself (the graph-captured
// engine step: kernel parameters never change, so every size that varies per keyframe lives in HBM).
struct bf_dimref { int host; const int32_t* dev; };
static inline bf_dimref bf_dim_host(int v) { bf_dimref d; d.host = v; d.dev = nullptr; return d; }
static inline bf_dimref bf_dim_dev(const int32_t* p, int cap) { bf_dimref d; d.host = cap; d.dev = p; return d; }
#ifdef __CUDACC__
__device__ __forceinline__ int bf_dim(const bf_dimref d) { return d.dev ? *d.dev : d.host; }
#endif

// Every entry point runs on the handle's device whatever the caller's current device is (and restores it).
struct bf_device_guard {
    int prev;
    bool switched;
    explicit bf_device_guard(const bf_handle* h) : prev(-1), switched(false) {
        if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->device) switched = (cudaSetDevice(h->device) == cudaSuccess);
    }
    ~bf_device_guard() { if (switched) cudaSetDevice(prev); }
};

static inline int bf_fail(bf_handle* h, int code, const char* what, const char* detail) {
    if (h) snprintf(h->err, sizeof(h->err), "%s: %s", what, detail ? detail : "");
    return code;
}

#define BF_CUDA(h, expr)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) return bf_fail((h), BF_ERR_CUDA, #expr, cudaGetErrorString(e__)); \
    } while (0)

#define BF_LAUNCH_CHECK(h, name)                                                           \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) return bf_fail((h), BF_ERR_CUDA, name, cudaGetErrorString(e__)); \
    } while (0)

// Grow-only scratch.  Growing frees the old block (cudaFree synchronises the device, so no kernel
// can still be reading it).
static inline int bf_scratch(bf_handle* h, int slot, size_t bytes, void** out) {
    if (bytes < (4u << 20)) bytes = 4u << 20;           // 4 MB floor: maps of a few hundred boxes never regrow (a regrow is a device-wide sync)
    if (h->cap[slot] < bytes) {
        if (h->frozen) return bf_fail(h, BF_ERR_CAPACITY, "bf_scratch", "scratch of a graph-captured engine cannot grow (raise the engine capacities)");
        if (h->buf[slot]) BF_CUDA(h, cudaFree(h->buf[slot]));
        h->buf[slot] = nullptr;
        h->cap[slot] = 0;
        size_t want = 2 * bytes + 65536;                 // geometric growth: a growing map reallocates O(log N) times
        BF_CUDA(h, cudaMalloc(&h->buf[slot], want));
        h->cap[slot] = want;
    }
    *out = h->buf[slot];
    return BF_OK;
}

static inline int bf_blocks(long long n, int threads) { return (int)((n + threads - 1) / threads); }
