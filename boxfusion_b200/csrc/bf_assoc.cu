// K2 - detection-to-map association: 3-D NMS as an over-threshold bit mask (in score-rank space) plus
// a greedy matching-and-merge kernel that reproduces nms_3d (instances.py:22-101) together with the
// fusion-list bookkeeping of BoxManager.record (box_manager.py:40-88), and the 2-D correspondence
// scoring for small objects (instances.py:446-468, 643-717).
#include "bf_common.cuh"

int bf_iou3d_run(bf_handle* h, const float* cornersA, int M, const float* cornersB, int N, int triangle, int mode,
                 double* iou, int32_t* counts, int64_t* stats, double thr, const int32_t* rank, uint32_t* mask,
                 uint32_t* rowany, int W, unsigned long long* edges, int edge_cap, cudaStream_t st);
int bf_iou3d_overflowed(bf_handle* h, cudaStream_t st, int* overflow);

__global__ void bf_rank_kernel(const int32_t* __restrict__ order, int N, int32_t* __restrict__ rank) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < N) rank[order[r]] = r;
}

// ---- score order of nms_3d (instances.py:52) ---------------------------------------------------------------------
// 64-bit keys (descending score, ascending index) sorted ascending by a bitonic network in shared memory.
#define BF_ORDER_MAX 4096
__global__ void __launch_bounds__(1024)
bf_score_order_kernel(const float* __restrict__ scores, int N, int n_pad, int32_t* __restrict__ order) {
    extern __shared__ unsigned long long bf_keys[];
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        unsigned long long k = ~0ull;                                   // padding sorts last
        if (i < N) {
            float f = scores[i];
            unsigned u;
            if (f != f) u = 0xffffffffu;                                // NaN: greater than everything (torch's order)
            else {
                if (f == 0.0f) f = 0.0f;                                // -0 == +0
                const unsigned b = __float_as_uint(f);
                u = (b & 0x80000000u) ? ~b : (b | 0x80000000u);         // monotone float -> unsigned
            }
            k = ((unsigned long long)(~u) << 32) | (unsigned)i;        // descending score, then ascending index
        }
        bf_keys[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= n_pad; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = bf_keys[lo], b = bf_keys[hi];
                if ((a > b) == up) { bf_keys[lo] = b; bf_keys[hi] = a; }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < N; i += blockDim.x) order[i] = (int32_t)(bf_keys[i] & 0xffffffffull);
}

extern "C" int bf_score_order(bf_handle* h, const float* scores, int N, int32_t* order, void* stream) {
    if (!h || N < 0 || N > BF_ORDER_MAX || (N > 0 && (!scores || !order))) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_score_order", "bad argument");
    if (N == 0) return BF_OK;
    int n_pad = 2;
    while (n_pad < N) n_pad <<= 1;
    const int threads = n_pad / 2 < 1024 ? (n_pad / 2 < 32 ? 32 : n_pad / 2) : 1024;
    bf_score_order_kernel<<<1, threads, sizeof(unsigned long long) * (size_t)n_pad, (cudaStream_t)stream>>>(scores, N, n_pad, order);
    BF_LAUNCH_CHECK(h, "bf_score_order_kernel");
    return BF_OK;
}

// box_manager.py:188-215 with the test of :55 / :71
__device__ __forceinline__ bool bf_views_differ(const float* __restrict__ p1, const float* __restrict__ p2,
                                                float translation_gap, float rotation_gap, bool use_center,
                                                float center_dis, float center_gap) {
    const float dx = p2[3] - p1[3], dy = p2[7] - p1[7], dz = p2[11] - p1[11];
    const float baseline = sqrtf(dx * dx + dy * dy + dz * dz);
    float tr = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) tr += p2[4 * r] * p1[4 * r] + p2[4 * r + 1] * p1[4 * r + 1] + p2[4 * r + 2] * p1[4 * r + 2];
    const float c = fminf(fmaxf((tr - 1.f) * 0.5f, -1.f), 1.f);
    const float angle = acosf(c) * 180.f / 3.14159265358979323846f;
    return (baseline > translation_gap || angle > rotation_gap) || (use_center && center_dis > center_gap);
}

// sorted insert of `val` into list[0..len) (ascending; duplicates kept, like list.sort())
__device__ __forceinline__ void bf_sorted_insert(int32_t* list, int& len, int32_t val) {
    int k = len;
    while (k > 0 && list[k - 1] > val) { list[k] = list[k - 1]; --k; }
    list[k] = val;
    ++len;
}

// BoxManager.record for one suppressed box `idx` of head `cur` (box_manager.py:48-86).  lc/len_c: the head's list.
struct bf_record_ctx {
    const int32_t* order; const int32_t* init_id; const float* poses; const float* centers;
    int32_t* fl; int32_t* flen; int32_t* fflag; int32_t* keep; int32_t* status;
    float translation_gap, rotation_gap, center_gap;
};

__device__ __forceinline__ void bf_record_one(const bf_record_ctx& c, int cur, int idx, int32_t* lc, int& len_c,
                                              bool& cur_in_keep) {
    const float* cc = c.centers + 3 * cur;
    const float* ci = c.centers + 3 * idx;
    const float ex = cc[0] - ci[0], ey = cc[1] - ci[1], ez = cc[2] - ci[2];
    const float cdis = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
    const int len_i = c.flen[idx];
    const int32_t* li = c.fl + (size_t)idx * BF_FUSION_CAP;
    if (len_i == 1) {                                     // box_manager.py:50-62
        const float* pi = c.poses + 16 * (size_t)c.init_id[idx];
        int cnt = 0;
        for (int k = 0; k < len_c; ++k)
            cnt += bf_views_differ(c.poses + 16 * (size_t)lc[k], pi, c.translation_gap, c.rotation_gap, true, cdis, c.center_gap);
        if (cnt == len_c && len_c < 5) {
            if (len_c + 1 > BF_FUSION_CAP) c.status[0] = BF_ERR_CAPACITY;
            else bf_sorted_insert(lc, len_c, c.init_id[idx]);
        }
    } else {                                              // box_manager.py:65-86
        const float* pc = c.poses + 16 * (size_t)c.init_id[cur];
        int cnt = 0;
        for (int k = 0; k < len_i; ++k)
            cnt += bf_views_differ(c.poses + 16 * (size_t)li[k], pc, c.translation_gap, c.rotation_gap, true, cdis, c.center_gap);
        if (cnt == len_i && len_i < 5) {
            if (len_c + len_i > BF_FUSION_CAP) c.status[0] = BF_ERR_CAPACITY;
            else for (int k = 0; k < len_i; ++k) bf_sorted_insert(lc, len_c, li[k]);
        } else if (cur_in_keep) {                         // swap: keep.remove(cur); keep.append(idx)
            cur_in_keep = false;
            c.keep[idx] = 1;                              // forced keep
            c.keep[cur] = -1;                             // dropped
        }
        if (c.fflag[idx] == 1) c.fflag[cur] = 1;
    }
}

// Sparse greedy matching (the common case: few over-threshold pairs).  The IoU kernels emit one edge
// (rank_lo << 32 | rank_hi) per over-threshold pair.  One CTA sorts the edges, thread 0 walks them in order to
// decide which (head, partner) pairs are live - nms_3d's loop (instances.py:58-97) touches nothing else - and then
// record() runs in parallel over heads: calls of different heads write disjoint lists/flags and only read lists of
// suppressed boxes, which never change.  Falls through (returns) when the edge list overflowed; the dense kernel
// below handles that case.
#define BF_EDGE_CAP 8192
__global__ void __launch_bounds__(1024)
bf_greedy_edges_kernel(const unsigned long long* __restrict__ edges, const unsigned long long* __restrict__ counters,
                       int N, int W, bf_record_ctx ctx, int32_t* __restrict__ success) {
    extern __shared__ unsigned long long s_keys[];
    const unsigned long long E64 = counters[6];
    if (E64 > BF_EDGE_CAP) return;
    const int E = (int)E64, tid = threadIdx.x, T = blockDim.x;
    int n2 = 1;
    while (n2 < E) n2 <<= 1;
    uint32_t* remaining = (uint32_t*)(s_keys + n2);
    unsigned char* act = (unsigned char*)(remaining + W);
    for (int i = tid; i < N; i += T) { ctx.keep[i] = 0; success[i] = 0; }
    if (tid == 0) ctx.status[0] = 0;
    for (int i = tid; i < n2; i += T) s_keys[i] = (i < E) ? edges[i] : ~0ULL;
    for (int w = tid; w < W; w += T) {
        const int base = w << 5;
        remaining[w] = (base + 32 <= N) ? 0xffffffffu : ((base < N) ? ((1u << (N - base)) - 1u) : 0u);
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)                      // bitonic sort, ascending
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += T) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = s_keys[i], b = s_keys[p];
                    if (((i & k) == 0) ? (a > b) : (a < b)) { s_keys[i] = b; s_keys[p] = a; }
                }
            }
            __syncthreads();
        }
    if (tid == 0) {                                        // the serial part of nms_3d: who is a head, who is suppressed
        int cur = -1;
        bool alive = false;
        for (int e = 0; e < E; ++e) {
            const int r0 = (int)(s_keys[e] >> 32), r1 = (int)(s_keys[e] & 0xffffffffu);
            if (r0 != cur) { cur = r0; alive = (remaining[r0 >> 5] >> (r0 & 31)) & 1u; }
            const bool live = alive && ((remaining[r1 >> 5] >> (r1 & 31)) & 1u);
            act[e] = live ? 1 : 0;
            if (live) remaining[r1 >> 5] &= ~(1u << (r1 & 31));
        }
    }
    __syncthreads();
    for (int e = tid; e < E; e += T) {                     // one thread per head: record() over its live partners, in order
        const int r0 = (int)(s_keys[e] >> 32);
        if (e > 0 && (int)(s_keys[e - 1] >> 32) == r0) continue;
        const int cur = ctx.order[r0];
        int32_t* lc = ctx.fl + (size_t)cur * BF_FUSION_CAP;
        int len_c = ctx.flen[cur];
        bool cur_in_keep = true, any = false;
        for (int g = e; g < E && (int)(s_keys[g] >> 32) == r0; ++g) {
            if (!act[g]) continue;
            any = true;
            bf_record_one(ctx, cur, ctx.order[(int)(s_keys[g] & 0xffffffffu)], lc, len_c, cur_in_keep);
        }
        if (any) { success[cur] = 1; ctx.flen[cur] = len_c; }
    }
    __syncthreads();
    for (int r = tid; r < N; r += T) {                     // keep = never suppressed, minus dropped heads, plus forced keeps
        const int i = ctx.order[r];
        const bool rem = (remaining[r >> 5] >> (r & 31)) & 1u;
        const int k = ctx.keep[i];
        ctx.keep[i] = (k == 1) ? 1 : ((k == -1) ? 0 : (rem ? 1 : 0));
    }
}

// Dense fallback (edge list overflowed).  One warp.  Walks, in ascending score rank, the heads that have at least one over-threshold partner;
// every other box is kept untouched.  Lane 0 performs record(); the other lanes help with the bit-mask rows.
__global__ void __launch_bounds__(32)
bf_greedy_kernel(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ rowany, int N, int W,
                 const int32_t* __restrict__ order, const int32_t* __restrict__ init_id,
                 const float* __restrict__ poses, int M, const float* __restrict__ centers,
                 int32_t* __restrict__ fl, int32_t* __restrict__ flen, int32_t* __restrict__ fflag,
                 float translation_gap, float rotation_gap, float center_gap,
                 int32_t* __restrict__ keep, int32_t* __restrict__ success, int32_t* __restrict__ status,
                 const unsigned long long* __restrict__ counters) {
    extern __shared__ uint32_t s_mem[];
    if (counters[6] <= BF_EDGE_CAP) return;              // bf_greedy_edges_kernel handled this call
    uint32_t* remaining = s_mem;            // [W] ranks not yet suppressed
    uint32_t* sup = s_mem + W;              // [W] scratch: suppressed by the current head
    const int lane = threadIdx.x;
    for (int w = lane; w < W; w += 32) {
        const int base = w << 5;
        remaining[w] = (base + 32 <= N) ? 0xffffffffu : ((base < N) ? ((1u << (N - base)) - 1u) : 0u);
    }
    for (int i = lane; i < N; i += 32) { keep[i] = 0; success[i] = 0; }
    if (lane == 0) status[0] = 0;
    __syncwarp();
    for (int hw = 0; hw < W; ++hw) {
        uint32_t heads = rowany[hw];
        while (heads) {
            const int bit = __ffs(heads) - 1;
            heads &= heads - 1;
            const int r = (hw << 5) + bit;
            if (!((remaining[hw] >> bit) & 1u)) continue;                  // suppressed earlier: never a head
            bool any = false;
            for (int w = lane; w < W; w += 32) {
                const uint32_t s = mask[(size_t)r * W + w] & remaining[w];
                sup[w] = s;
                any |= (s != 0);
            }
            any = __any_sync(0xffffffffu, any);
            __syncwarp();
            if (!any) continue;
            const int cur = order[r];
            if (lane == 0) {
                success[cur] = 1;                                            // instances.py:72-83
                bool cur_in_keep = true;
                int32_t* lc = fl + (size_t)cur * BF_FUSION_CAP;
                int len_c = flen[cur];
                const float* cc = centers + 3 * cur;
                for (int w = 0; w < W; ++w) {
                    uint32_t s = sup[w];
                    while (s) {
                        const int b2 = __ffs(s) - 1;
                        s &= s - 1;
                        const int idx = order[(w << 5) + b2];                 // descending score order (:85)
                        const float* ci = centers + 3 * idx;
                        const float ex = cc[0] - ci[0], ey = cc[1] - ci[1], ez = cc[2] - ci[2];
                        const float cdis = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
                        const int len_i = flen[idx];
                        const int32_t* li = fl + (size_t)idx * BF_FUSION_CAP;
                        if (len_i == 1) {                                     // box_manager.py:50-62
                            const float* pi = poses + 16 * (size_t)init_id[idx];
                            int cnt = 0;
                            for (int k = 0; k < len_c; ++k)
                                cnt += bf_views_differ(poses + 16 * (size_t)lc[k], pi, translation_gap, rotation_gap, true, cdis, center_gap);
                            if (cnt == len_c && len_c < 5) {
                                if (len_c + 1 > BF_FUSION_CAP) status[0] = BF_ERR_CAPACITY;
                                else bf_sorted_insert(lc, len_c, init_id[idx]);
                            }
                        } else {                                              // box_manager.py:65-86
                            const float* pc = poses + 16 * (size_t)init_id[cur];
                            int cnt = 0;
                            for (int k = 0; k < len_i; ++k)
                                cnt += bf_views_differ(poses + 16 * (size_t)li[k], pc, translation_gap, rotation_gap, true, cdis, center_gap);
                            if (cnt == len_i && len_i < 5) {
                                if (len_c + len_i > BF_FUSION_CAP) status[0] = BF_ERR_CAPACITY;
                                else for (int k = 0; k < len_i; ++k) bf_sorted_insert(lc, len_c, li[k]);
                            } else if (cur_in_keep) {                         // swap: keep.remove(cur); keep.append(idx)
                                cur_in_keep = false;
                                keep[idx] = 1;                                // forced keep
                                keep[cur] = -1;                               // dropped
                            }
                            if (fflag[idx] == 1) fflag[cur] = 1;
                        }
                    }
                }
                flen[cur] = len_c;
            }
            __syncwarp();
            for (int w = lane; w < W; w += 32) remaining[w] &= ~sup[w];
            __syncwarp();
        }
    }
    __syncwarp();
    // keep = never-suppressed ranks, minus dropped heads, plus forced keeps
    for (int r = lane; r < N; r += 32) {
        const int i = order[r];
        const bool rem = (remaining[r >> 5] >> (r & 31)) & 1u;
        const int k = keep[i];
        keep[i] = (k == 1) ? 1 : ((k == -1) ? 0 : (rem ? 1 : 0));
    }
}

extern "C" int bf_nms3d(bf_handle* h, const float* corners, const float* centers, int N, const int32_t* order,
                        const int32_t* init_id, const float* poses, int M, int32_t* fusion_list, int32_t* fusion_len,
                        int32_t* fusion_flag, double iou_threshold, float translation_gap, float rotation_gap_deg,
                        float center_gap, int mode, int32_t* keep, int32_t* success, int32_t* status, void* stream) {
    if (!h || N < 0 || M < 0) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_nms3d", "bad size");
    if (N == 0) return BF_OK;
    if (!corners || !centers || !order || !init_id || !poses || !fusion_list || !fusion_len || !fusion_flag || !keep ||
        !success || !status)
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_nms3d", "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int W = (N + 31) / 32;
    if ((size_t)2 * W * sizeof(uint32_t) > 200 * 1024) return bf_fail(h, BF_ERR_CAPACITY, "bf_nms3d", "N too large for the greedy kernel");
    void* p;
    int rc;
    if ((rc = bf_scratch(h, BF_SCRATCH_MASK, sizeof(uint32_t) * ((size_t)N * W + W), &p))) return rc;
    uint32_t* mask = (uint32_t*)p;
    uint32_t* rowany = mask + (size_t)N * W;
    if ((rc = bf_scratch(h, BF_SCRATCH_RANK, sizeof(int32_t) * (size_t)N, &p))) return rc;
    int32_t* rank = (int32_t*)p;
    if ((rc = bf_scratch(h, BF_SCRATCH_EDGES, sizeof(unsigned long long) * BF_EDGE_CAP, &p))) return rc;
    unsigned long long* edges = (unsigned long long*)p;
    for (int attempt = 0; attempt < 2; ++attempt) {
        BF_CUDA(h, cudaMemsetAsync(mask, 0, sizeof(uint32_t) * ((size_t)N * W + W), st));
        bf_rank_kernel<<<bf_blocks(N, 128), 128, 0, st>>>(order, N, rank);
        BF_LAUNCH_CHECK(h, "bf_rank_kernel");
        if ((rc = bf_iou3d_run(h, corners, N, corners, N, 1, mode, nullptr, nullptr, nullptr, iou_threshold, rank, mask,
                               rowany, W, edges, BF_EDGE_CAP, st)))
            return rc;
        if ((long long)N * N <= (long long)(h->cap[BF_SCRATCH_WORK] / 8)) break;
        int ovf = 0;
        if ((rc = bf_iou3d_overflowed(h, st, &ovf))) return rc;
        if (!ovf) break;
        if (attempt == 1) return bf_fail(h, BF_ERR_CAPACITY, "bf_nms3d", "work list overflow");
        if ((rc = bf_scratch(h, BF_SCRATCH_WORK, 8 * (size_t)N * N / 2, &p))) return rc;
    }
    const unsigned long long* counters = (const unsigned long long*)h->buf[BF_SCRATCH_COUNTERS];
    bf_record_ctx ctx;
    ctx.order = order; ctx.init_id = init_id; ctx.poses = poses; ctx.centers = centers; ctx.fl = fusion_list;
    ctx.flen = fusion_len; ctx.fflag = fusion_flag; ctx.keep = keep; ctx.status = status;
    ctx.translation_gap = translation_gap; ctx.rotation_gap = rotation_gap_deg; ctx.center_gap = center_gap;
    // sparse path: sorted edge list in shared memory (keys + remaining bit set + live flags)
    const size_t smem_e = sizeof(unsigned long long) * BF_EDGE_CAP + sizeof(uint32_t) * (size_t)W + BF_EDGE_CAP + 16;
    BF_CUDA(h, cudaFuncSetAttribute(bf_greedy_edges_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
    bf_greedy_edges_kernel<<<1, 1024, smem_e, st>>>(edges, counters, N, W, ctx, success);
    BF_LAUNCH_CHECK(h, "bf_greedy_edges_kernel");
    // dense fallback: returns immediately unless the edge list overflowed
    const size_t smem = sizeof(uint32_t) * 2 * (size_t)W;
    if (smem > 48 * 1024)
        BF_CUDA(h, cudaFuncSetAttribute(bf_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bf_greedy_kernel<<<1, 32, smem, st>>>(mask, rowany, N, W, order, init_id, poses, M, centers, fusion_list, fusion_len,
                                          fusion_flag, translation_gap, rotation_gap_deg, center_gap, keep, success, status,
                                          counters);
    BF_LAUNCH_CHECK(h, "bf_greedy_kernel");
    return BF_OK;
}

// ------------------------------------------------------------------------------------------------
// correspondence_association scoring (instances.py:446-468): one thread per map box projects its 8
// corners (float64, like numpy), then one warp per small detection scans the G IoUs for the first max.
__global__ void bf_corr_project_kernel(const float* __restrict__ corners, int G, const float* __restrict__ pinv,
                                       double fx, double fy, double cx, double cy, double W, double H,
                                       double* __restrict__ boxes2d) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    bool any_valid = false, any_z = false;
    double umin = 0, vmin = 0, umax = 0, vmax = 0;
    for (int i = 0; i < 8; ++i) {
        const double x = corners[24 * g + 3 * i], y = corners[24 * g + 3 * i + 1], z = corners[24 * g + 3 * i + 2];
        // np.dot(boxes_homo, pose_inv.T): row i of pose_inv (float32 -> float64)
        const double X = x * (double)pinv[0] + y * (double)pinv[1] + z * (double)pinv[2] + (double)pinv[3];
        const double Y = x * (double)pinv[4] + y * (double)pinv[5] + z * (double)pinv[6] + (double)pinv[7];
        const double Z = x * (double)pinv[8] + y * (double)pinv[9] + z * (double)pinv[10] + (double)pinv[11];
        const double u = (fx * X / Z) + cx, v = (fy * Y / Z) + cy;
        any_valid |= (Z > 0) && (u > 0) && (u < W) && (v > 0) && (v < H);     // instances.py:693
        if (Z > 0 && Z < 8) {                                                 // instances.py:702
            const double uc = fmin(fmax(u, 0.0), W), vc = fmin(fmax(v, 0.0), H);
            if (!any_z) { umin = umax = uc; vmin = vmax = vc; any_z = true; }
            else { umin = fmin(umin, uc); umax = fmax(umax, uc); vmin = fmin(vmin, vc); vmax = fmax(vmax, vc); }
        }
    }
    const bool ok = any_valid && any_z;
    boxes2d[4 * g + 0] = ok ? umin : 0.0;
    boxes2d[4 * g + 1] = ok ? vmin : 0.0;
    boxes2d[4 * g + 2] = ok ? umax : 0.0;
    boxes2d[4 * g + 3] = ok ? vmax : 0.0;
}

__global__ void bf_corr_match_kernel(const double* __restrict__ boxes2d, const int32_t* __restrict__ small_mask, int G,
                                     const float* __restrict__ det, int n_small, int32_t* __restrict__ best,
                                     double* __restrict__ best_iou) {
    const int d = blockIdx.x;
    const int lane = threadIdx.x;
    if (d >= n_small) return;
    const double ax0 = det[4 * d], ay0 = det[4 * d + 1], ax1 = det[4 * d + 2], ay1 = det[4 * d + 3];
    const double areaA = (ax1 - ax0) * (ay1 - ay0);
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int g = lane; g < G; g += 32) {
        const double bx0 = boxes2d[4 * g], by0 = boxes2d[4 * g + 1], bx1 = boxes2d[4 * g + 2], by1 = boxes2d[4 * g + 3];
        const double areaB = (bx1 - bx0) * (by1 - by0);
        const double iw = fmax(0.0, fmin(ax1, bx1) - fmax(ax0, bx0));
        const double ih = fmax(0.0, fmin(ay1, by1) - fmax(ay0, by0));
        const double inter = iw * ih;
        double v = inter / (areaA + areaB - inter + 1e-6);                    // instances.py:643-668
        v = v * (small_mask[g] ? 1.0 : 0.0);                                  // instances.py:460-461
        if (v > bv) { bv = v; bi = g; }                                       // ascending g per lane: first max
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }           // np.argmax: first occurrence
    }
    if (lane == 0) { best[d] = (G > 0) ? bi : -1; best_iou[d] = (G > 0) ? bv : 0.0; }
}

extern "C" int bf_corr2d(bf_handle* h, const float* map_corners, const int32_t* small_mask, int G, const float* pose_inv,
                         float fx, float fy, float cx, float cy, float W, float H, const float* det_xyxy, int n_small,
                         double* boxes2d, int32_t* best, double* best_iou, void* stream) {
    if (!h || G < 0 || n_small < 0) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_corr2d", "bad size");
    if (n_small == 0) return BF_OK;
    if (!best || !best_iou || !det_xyxy || !pose_inv || (G > 0 && (!map_corners || !small_mask)))
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_corr2d", "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    double* b2 = boxes2d;
    if (!b2) {
        void* p;
        int rc = bf_scratch(h, BF_SCRATCH_MISC, sizeof(double) * 4 * (size_t)(G > 0 ? G : 1), &p);
        if (rc) return rc;
        b2 = (double*)p;
    }
    if (G > 0) {
        bf_corr_project_kernel<<<bf_blocks(G, 128), 128, 0, st>>>(map_corners, G, pose_inv, (double)fx, (double)fy,
                                                                 (double)cx, (double)cy, (double)W, (double)H, b2);
        BF_LAUNCH_CHECK(h, "bf_corr_project_kernel");
    }
    bf_corr_match_kernel<<<n_small, 32, 0, st>>>(b2, small_mask, G, det_xyxy, n_small, best, best_iou);
    BF_LAUNCH_CHECK(h, "bf_corr_match_kernel");
    return BF_OK;
}
