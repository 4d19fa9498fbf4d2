// Diagnostics only (tools/eval_profile.py): where the cycles of one (particle, view) evaluation go.
//
// The kernel below runs bf_refine_kernel's mode-A inner loop - one evaluation per thread pass, view-major, views staged in
// shared memory - for ONE box and ONE optimiser state, with BF_EVAL_PROFILE ticks compiled into bf_refine_eval.cuh: lane 0 of
// every warp accumulates the cycles between section boundaries.  Same flags as bf_refine.cu (-fmad=false), same source, so
// the per-section shares carry over; nothing in the product path calls this.
#define BF_EVAL_PROFILE 1
#include "bf_internal.cuh"
#include "bf_refine_eval.cuh"

#define BF_DBG_MAX_VIEWS 32

template <bool ROLL>
__global__ void __launch_bounds__(512, 1)
bf_eval_profile_kernel(const float* __restrict__ pst, int P, int PB, const float* __restrict__ state /* box6[6] search[6] rot[9] */,
                       const float* __restrict__ poses, const float* __restrict__ uv, int V, const float* __restrict__ intr,
                       int reps, float* __restrict__ out, long long* __restrict__ cycles) {
    __shared__ bf_view views[BF_DBG_MAX_VIEWS];
    __shared__ float st[21];
    const int tid = threadIdx.x, T = blockDim.x;
    const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3], img_w = intr[4], img_h = intr[5];
    for (int v = tid; v < V; v += T) bf_view_stage(views[v], poses + 16 * v, uv + 16 * v, img_w, img_h);
    if (tid < 21) st[tid] = state[tid];
    __syncthreads();
    const int w0 = (blockIdx.x * (T >> 5) + (tid >> 5)) & 63;
    const int wl = tid >> 5;
    if ((tid & 31) == 0) {
        for (int k = 0; k < BF_PROF_SECTIONS; ++k) bf_prof_acc[wl * BF_PROF_SECTIONS + k] = 0;
    }
    __syncthreads();
    int overflow = 0;
    const int p_lo = blockIdx.x * PB;
    const int items = PB * V;
    for (int r = 0; r < reps; ++r) {
        if ((tid & 31) == 0) bf_prof_last[wl] = clock64();
        for (int w = tid; w < items; w += T) {
            const int v = w / PB, p = p_lo + (w - v * PB);
            if (p < P) {
                float pst6[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) pst6[k] = __ldg(pst + 6 * p + k);
                float c[8][3];
                bf_particle_corners(st, pst6, st + 6, st + 12, c);
                BF_TICK(0)
                out[(size_t)blockIdx.x * items + w] = bf_eval_view<ROLL>(c, views[v], fx, cx, fy, cy, img_w, img_h, &overflow, nullptr);
                BF_TICK(11)
            }
        }
        __syncthreads();
    }
    if ((tid & 31) == 0)
        for (int k = 0; k < BF_PROF_SECTIONS; ++k) cycles[w0 * BF_PROF_SECTIONS + k] = bf_prof_acc[wl * BF_PROF_SECTIONS + k];
}

// cycles: [64][16] int64 (per warp slot, per section), summed over `reps` repetitions of the same pass.  grid * T/32 <= 64.
extern "C" int bf_debug_eval_profile(const float* pst, int P, int PB, const float* state21, const float* poses, const float* uv, int V,
                                     const float* intr6, int grid, int threads, int roll, int reps, float* out, long long* cycles,
                                     void* stream) {
    if (V > BF_DBG_MAX_VIEWS || threads > 512 || grid * (threads / 32) > 64) return BF_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (roll) bf_eval_profile_kernel<true><<<grid, threads, 0, st>>>(pst, P, PB, state21, poses, uv, V, intr6, reps, out, cycles);
    else bf_eval_profile_kernel<false><<<grid, threads, 0, st>>>(pst, P, PB, state21, poses, uv, V, intr6, reps, out, cycles);
    return cudaGetLastError() == cudaSuccess ? BF_OK : BF_ERR_CUDA;
}
