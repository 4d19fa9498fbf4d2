// K2 - detection-to-map association: 3-D NMS as an over-threshold bit mask (in score-rank space) plus
// a greedy matching-and-merge kernel that reproduces nms_3d (instances.py:22-101) together with the
// fusion-list bookkeeping of BoxManager.record (box_manager.py:40-88), and the 2-D correspondence
// scoring for small objects (instances.py:446-468, 643-717).
#include "bf_common.cuh"

#include "bf_internal.cuh"
#include "bf_record.cuh"

__global__ void bf_rank_kernel(const int32_t* __restrict__ order, const bf_dimref Nd, int32_t* __restrict__ rank) {
    const int N = bf_dim(Nd);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) rank[order[r]] = r;
}

// ---- score order of nms_3d (instances.py:52) ---------------------------------------------------------------------
// 64-bit keys (descending score, ascending index), unique per box.  N <= 8192 (incl. the 4 352-box stress map of BASELINE
// configs[2]): block 0 sorts them ascending with a bitonic network in shared memory.  Larger N (engine maps up to
// 65 536 rows): every block ranks its boxes by counting the smaller keys, tile by tile through shared memory -
// O(N^2) compares like the pair stage of the NMS it feeds, no multi-pass global sort, fixed launch shape.
#define BF_ORDER_SMEM 8192
__device__ __forceinline__ unsigned long long bf_score_key(float f, int i) {
    unsigned u;
    if (f != f) u = 0xffffffffu;                                // NaN: greater than everything (torch's order)
    else {
        if (f == 0.0f) f = 0.0f;                                // -0 == +0
        const unsigned b = __float_as_uint(f);
        u = (b & 0x80000000u) ? ~b : (b | 0x80000000u);         // monotone float -> unsigned
    }
    return ((unsigned long long)(~u) << 32) | (unsigned)i;      // descending score, then ascending index
}

__global__ void __launch_bounds__(1024)
bf_score_order_kernel(const float* __restrict__ scores, const bf_dimref Nd, int32_t* __restrict__ order, int32_t* __restrict__ rank) {
    extern __shared__ unsigned long long bf_keys[];      // BF_ORDER_SMEM keys (64 KB)
    const int N = bf_dim(Nd);
    if (N <= BF_ORDER_SMEM) {
        if (blockIdx.x != 0 || N <= 0) return;
        int n_pad = 2;
        while (n_pad < N) n_pad <<= 1;
        for (int i = threadIdx.x; i < n_pad; i += blockDim.x) bf_keys[i] = (i < N) ? bf_score_key(scores[i], i) : ~0ull;   // padding sorts last
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
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const int idx = (int)(bf_keys[i] & 0xffffffffull);
            order[i] = idx;
            if (rank) rank[idx] = i;
        }
        return;
    }
    // rank by counting: box i goes to position #{j : key_j < key_i}
    const int per_pass = gridDim.x * blockDim.x;
    for (int base = 0; base < N; base += per_pass) {                 // block-uniform trip count
        const int i = base + blockIdx.x * blockDim.x + threadIdx.x;
        const unsigned long long ki = (i < N) ? bf_score_key(scores[i], i) : 0ull;
        int cnt = 0;
        for (int t0 = 0; t0 < N; t0 += BF_ORDER_SMEM) {
            __syncthreads();
            for (int j = threadIdx.x; j < BF_ORDER_SMEM; j += blockDim.x) bf_keys[j] = (t0 + j < N) ? bf_score_key(scores[t0 + j], t0 + j) : ~0ull;
            __syncthreads();
            const int lim = min(BF_ORDER_SMEM, N - t0);
#pragma unroll 8
            for (int j = 0; j < lim; ++j) cnt += (bf_keys[j] < ki) ? 1 : 0;
        }
        if (i < N) { order[cnt] = i; if (rank) rank[i] = cnt; }
    }
}

int bf_score_order_run(bf_handle* h, const float* scores, bf_dimref Nd, int32_t* order, int32_t* rank, cudaStream_t st) {
    const int blocks = Nd.host <= BF_ORDER_SMEM ? 1 : (bf_blocks(Nd.host, 1024) < h->sm_count ? bf_blocks(Nd.host, 1024) : h->sm_count);
    const size_t smem = sizeof(unsigned long long) * BF_ORDER_SMEM;
    BF_CUDA(h, cudaFuncSetAttribute(bf_score_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bf_score_order_kernel<<<blocks, 1024, smem, st>>>(scores, Nd, order, rank);
    BF_LAUNCH_CHECK(h, "bf_score_order_kernel");
    return BF_OK;
}

extern "C" int bf_score_order(bf_handle* h, const float* scores, int N, int32_t* order, void* stream) {
    bf_device_guard guard(h);
    if (!h || N < 0 || N > 65536 || (N > 0 && (!scores || !order))) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_score_order", "bad argument");
    if (N == 0) return BF_OK;
    return bf_score_order_run(h, scores, bf_dim_host(N), order, nullptr, (cudaStream_t)stream);
}

// record() of one head over its live partners, in descending score order (instances.py:72-85).  `keys` is a list of
// (head rank << 32 | partner rank) sorted ascending, `first` the head's first entry; entries with act[g] == 0 are skipped.
__device__ __forceinline__ void bf_record_head(const bf_record_ctx& ctx, const unsigned long long* keys, const unsigned char* act,
                                               int first, int E, int32_t* __restrict__ success) {
    const int r0 = (int)(keys[first] >> 32);
    const int cur = ctx.order[r0];
    int32_t* lc = ctx.fl + (size_t)cur * BF_FUSION_CAP;
    int len_c = ctx.flen[cur];
    bool cur_in_keep = true, any = false;
    for (int g = first; g < E && (int)(keys[g] >> 32) == r0; ++g) {
        if (act && !act[g]) continue;
        any = true;
        const int idx = ctx.order[(int)(keys[g] & 0xffffffffu)];
        if (bf_record_one(ctx, cur, idx, lc, len_c) && cur_in_keep) {   // swap: keep.remove(cur); keep.append(idx)
            cur_in_keep = false;
            ctx.keep[idx] = 1;                            // forced keep
            ctx.keep[cur] = -1;                           // dropped
        }
    }
    if (any) { success[cur] = 1; ctx.flen[cur] = len_c; }
}

// Greedy matching.  The IoU kernels emit one edge (rank_lo << 32 | rank_hi) per over-threshold pair (and set the pair's
// bit in the dense rank-space mask when there is one).  One CTA.
//   sparse path (the common case, <= BF_EDGE_CAP edges): sort the edges in shared memory, thread 0 walks them in order to
//     decide which (head, partner) pairs are live - nms_3d's loop (instances.py:58-97) touches nothing else;
//   dense path (edge list overflowed): the mask read row by row IS the sorted edge list.  Warp 0 walks, in ascending score
//     rank, the heads that have at least one over-threshold partner and writes the LIVE (head, partner) pairs - at most one
//     per box, a box is suppressed once - to `live`;
// then record() runs in parallel over heads in both cases: calls of different heads write disjoint lists/flags and only read
// lists of suppressed boxes, which never change.  status is sticky (never cleared here): the caller zeroes it.
#define BF_EDGE_CAP 8192
__global__ void __launch_bounds__(1024)
bf_greedy_kernel(const unsigned long long* __restrict__ edges, const unsigned long long* __restrict__ counters,
                 const bf_dimref Nd, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ rowany,
                 unsigned long long* __restrict__ live, bf_record_ctx ctx, int32_t* __restrict__ success, int n_slots) {
    extern __shared__ unsigned long long s_keys[];
    __shared__ int s_E;
    const int N = bf_dim(Nd);
    const int W = (N + 31) >> 5;
    // n_slots >= 0: `edges` is a gathered list of that many slots, empty ones hold ~0 (they sort last); else the IoU stage's counters
    const unsigned long long E64 = n_slots >= 0 ? (unsigned long long)n_slots : counters[6];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    if (tid == 0 && n_slots < 0 && counters[5]) atomicExch(ctx.status, BF_ERR_CAPACITY);   // the IoU stage's candidate list overflowed (fixed-size scratch)
    const bool dense = E64 > BF_EDGE_CAP;
    if (dense && !mask) {
        if (tid == 0) atomicExch(ctx.status, BF_ERR_CAPACITY);                   // no dense mask for maps this large
        return;
    }
    int E_sparse = dense ? 0 : (int)E64;
    int n2 = 1;
    while (n2 < E_sparse) n2 <<= 1;
    // shared memory: [sorted keys (sparse path only)] [remaining bit set] [live flags (sparse path only)]
    uint32_t* remaining = (uint32_t*)(s_keys + (dense ? 0 : n2));
    unsigned char* act = (unsigned char*)(remaining + W);
    for (int i = tid; i < N; i += T) { ctx.keep[i] = 0; success[i] = 0; }
    for (int w = tid; w < W; w += T) {
        const int base = w << 5;
        remaining[w] = (base + 32 <= N) ? 0xffffffffu : ((base < N) ? ((1u << (N - base)) - 1u) : 0u);
    }
    const unsigned long long* keys;
    const unsigned char* actp;
    int E;
    if (!dense) {
        if (tid == 0) s_E = 0;
        __syncthreads();
        int valid = 0;
        for (int i = tid; i < n2; i += T) {
            const unsigned long long key = (i < E_sparse) ? edges[i] : ~0ULL;
            s_keys[i] = key;
            valid += (key != ~0ULL) ? 1 : 0;
        }
        if (valid) atomicAdd(&s_E, valid);
        __syncthreads();
        E_sparse = s_E;                                        // real edges (gathered lists carry empty slots)
        if (n2 <= T) {
            // a keyframe has a few dozen edges: rank every key by counting the smaller ones (keys are unique; empty slots
            // are all ~0 and rank by position) - two block barriers instead of one per stage of a sorting network
            unsigned long long key = ~0ULL;
            int rnk = 0;
            if (tid < n2) {
                key = s_keys[tid];
                for (int j = 0; j < n2; ++j) { const unsigned long long o = s_keys[j]; rnk += (o < key || (o == key && j < tid)) ? 1 : 0; }
            }
            __syncthreads();
            if (tid < n2) s_keys[rnk] = key;
            __syncthreads();
        } else
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
            for (int e = 0; e < E_sparse; ++e) {
                const int r0 = (int)(s_keys[e] >> 32), r1 = (int)(s_keys[e] & 0xffffffffu);
                if (r0 != cur) { cur = r0; alive = (remaining[r0 >> 5] >> (r0 & 31)) & 1u; }
                const bool lv = alive && ((remaining[r1 >> 5] >> (r1 & 31)) & 1u);
                act[e] = lv ? 1 : 0;
                if (lv) remaining[r1 >> 5] &= ~(1u << (r1 & 31));
            }
        }
        __syncthreads();
        keys = s_keys; actp = act; E = E_sparse;
    } else {
        __syncthreads();
        if (tid < 32) {
            int En = 0;
            for (int hw = 0; hw < W; ++hw) {
                uint32_t heads = rowany[hw];
                while (heads) {
                    const int bit = __ffs(heads) - 1;
                    heads &= heads - 1;
                    const int r = (hw << 5) + bit;
                    if (!((remaining[hw] >> bit) & 1u)) continue;                  // suppressed earlier: never a head
                    for (int w0 = 0; w0 < W; w0 += 32) {                            // partners in ascending rank = descending score (:85)
                        const int w = w0 + lane;
                        const uint32_t sb = (w < W) ? (mask[(size_t)r * W + w] & remaining[w]) : 0u;
                        const int cpop = __popc(sb);
                        int incl = cpop;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
                        int pos = En + incl - cpop;
                        uint32_t q = sb;
                        while (q) {
                            const int b2 = __ffs(q) - 1;
                            q &= q - 1;
                            live[pos++] = ((unsigned long long)r << 32) | (unsigned long long)((w << 5) + b2);
                        }
                        if (w < W) remaining[w] &= ~sb;
                        En += __shfl_sync(0xffffffffu, incl, 31);
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) s_E = En;
        }
        __threadfence_block();
        __syncthreads();
        keys = live; actp = nullptr; E = s_E;
    }
    for (int e = tid; e < E; e += T) {                     // one thread per head: record() over its live partners, in order
        if (e > 0 && (int)(keys[e - 1] >> 32) == (int)(keys[e] >> 32)) continue;
        bf_record_head(ctx, keys, actp, e, E, success);
    }
    __syncthreads();
    for (int r = tid; r < N; r += T) {                     // keep = never suppressed, minus dropped heads, plus forced keeps
        const int i = ctx.order[r];
        const bool rem = (remaining[r >> 5] >> (r & 31)) & 1u;
        const int k = ctx.keep[i];
        ctx.keep[i] = (k == 1) ? 1 : ((k == -1) ? 0 : (rem ? 1 : 0));
        if (ctx.valid_num && success[i]) ctx.valid_num[i] += 1.f;          // instances.py:72-73
    }
}

#define BF_DENSE_MASK_MAX_N 16384      // the dense rank-space mask (N x N/32 words) is kept for maps up to this size

// nms_3d + record() from corners / centres / score order.  Nd: the box count on the host, or its bound on the host and the
// count itself in device memory.  All scratch comes from the handle (sized for Nd.host).  status is OR-ed into, never cleared.
int bf_nms3d_run(bf_handle* h, const float* corners, const float* centers, bf_dimref Nd, const int32_t* order, int32_t* rank_or_null,
                 const int32_t* init_id, const float* poses, int32_t* fusion_list, int32_t* fusion_len, int32_t* fusion_flag,
                 double iou_threshold, float translation_gap, float rotation_gap_deg, float center_gap, int mode,
                 int32_t* keep, int32_t* success, int32_t* status, float* valid_num_or_null, cudaStream_t st) {
    const int N = Nd.host;
    const int W = (N + 31) / 32;
    const bool dense = N <= BF_DENSE_MASK_MAX_N;
    void* p;
    int rc;
    uint32_t *mask = nullptr, *rowany = nullptr;
    if (dense) {
        if ((rc = bf_scratch(h, BF_SCRATCH_MASK, sizeof(uint32_t) * ((size_t)N * W + W) + sizeof(unsigned long long) * (size_t)(N + 1), &p))) return rc;
        mask = (uint32_t*)p;
    }
    int32_t* rank = rank_or_null;
    if (!rank) {
        if ((rc = bf_scratch(h, BF_SCRATCH_RANK, sizeof(int32_t) * (size_t)N, &p))) return rc;
        rank = (int32_t*)p;
    }
    if ((rc = bf_scratch(h, BF_SCRATCH_EDGES, sizeof(unsigned long long) * BF_EDGE_CAP, &p))) return rc;
    unsigned long long* edges = (unsigned long long*)p;
    unsigned long long* live = nullptr;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (dense) {
            // the mask rows are laid out with the stride of the ACTUAL box count (<= the bound); rowany and the live list sit
            // behind the rows of the bound, so their addresses do not depend on the actual count
            BF_CUDA(h, cudaMemsetAsync(mask, 0, sizeof(uint32_t) * ((size_t)N * W + W), st));
            rowany = mask + (size_t)N * W;
            live = (unsigned long long*)(mask + (((size_t)N * W + W + 1) & ~(size_t)1));
        }
        if (!rank_or_null) {
            bf_rank_kernel<<<bf_blocks(N, 128) < 64 ? bf_blocks(N, 128) : 64, 128, 0, st>>>(order, Nd, rank);
            BF_LAUNCH_CHECK(h, "bf_rank_kernel");
        }
        if ((rc = bf_iou3d_run(h, corners, Nd, corners, Nd, 1, 0, mode, nullptr, nullptr, nullptr, iou_threshold, rank, mask,
                               rowany, edges, BF_EDGE_CAP, st)))
            return rc;
        if (Nd.dev || (long long)N * N <= (long long)(h->cap[BF_SCRATCH_WORK] / 8)) break;   // cannot overflow / checked by the engine's status word
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
    ctx.valid_num = valid_num_or_null;
    // sorted edge list in shared memory (keys + remaining bit set + live flags); the dense path needs the bit set only
    const size_t smem_e = sizeof(unsigned long long) * BF_EDGE_CAP + sizeof(uint32_t) * (size_t)W + BF_EDGE_CAP + 16;
    BF_CUDA(h, cudaFuncSetAttribute(bf_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
    bf_greedy_kernel<<<1, 1024, smem_e, st>>>(edges, counters, Nd, dense ? mask : nullptr, rowany, live, ctx, success, -1);
    BF_LAUNCH_CHECK(h, "bf_greedy_kernel");
    return BF_OK;
}

// ---- the two halves of bf_nms3d as separate entries (SURVEY.md section 8(e) axis 3): the over-threshold pairs of a block
// of rows of the pair triangle, and the greedy scan + record() over a (gathered) edge list -------------------------------
__global__ void bf_edges_finish_kernel(unsigned long long* __restrict__ edges, int cap, const unsigned long long* __restrict__ counters,
                                       int32_t* __restrict__ status) {
    const unsigned long long E = counters[6];
    if (blockIdx.x == 0 && threadIdx.x == 0 && (E > (unsigned long long)cap || counters[5])) atomicExch(status, BF_ERR_CAPACITY);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x)
        if ((unsigned long long)i >= E) edges[i] = ~0ULL;          // empty slots sort last
}

extern "C" int bf_nms3d_edges(bf_handle* h, const float* corners, int N, const int32_t* order, int row_begin, int row_end,
                              double iou_threshold, int mode, unsigned long long* edges, int edge_cap, int32_t* status, void* stream) {
    bf_device_guard guard(h);
    if (!h || N < 0 || N > 65536 || row_begin < 0 || row_end < row_begin || row_end > N || edge_cap < 1 || edge_cap > BF_EDGE_CAP)
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_nms3d_edges", "bad size");
    if (!corners || !order || !edges || !status) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_nms3d_edges", "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    void* p;
    int rc;
    if ((rc = bf_scratch(h, BF_SCRATCH_RANK, sizeof(int32_t) * (size_t)(N > 0 ? N : 1), &p))) return rc;
    int32_t* rank = (int32_t*)p;
    if ((rc = bf_scratch(h, BF_SCRATCH_COUNTERS, sizeof(unsigned long long) * 8, &p))) return rc;
    const int M = row_end - row_begin;
    if (N > 0) {
        bf_rank_kernel<<<bf_blocks(N, 128) < 64 ? bf_blocks(N, 128) : 64, 128, 0, st>>>(order, bf_dim_host(N), rank);
        BF_LAUNCH_CHECK(h, "bf_rank_kernel");
    }
    if (M > 0) {
        if ((rc = bf_scratch(h, BF_SCRATCH_WORK, 8 * ((size_t)M * N < (1u << 20) ? (size_t)M * N : (size_t)(1u << 20) + 64 * (size_t)(M + N)), &p))) return rc;
        if ((rc = bf_iou3d_run(h, corners + 24 * (size_t)row_begin, bf_dim_host(M), corners, bf_dim_host(N), 1, row_begin, mode, nullptr, nullptr,
                               nullptr, iou_threshold, rank, nullptr, nullptr, edges, edge_cap, st)))
            return rc;
    } else {
        BF_CUDA(h, cudaMemsetAsync(h->buf[BF_SCRATCH_COUNTERS], 0, sizeof(unsigned long long) * 8, st));
    }
    bf_edges_finish_kernel<<<8, 256, 0, st>>>(edges, edge_cap, (const unsigned long long*)h->buf[BF_SCRATCH_COUNTERS], status);
    BF_LAUNCH_CHECK(h, "bf_edges_finish_kernel");
    return BF_OK;
}

extern "C" int bf_nms3d_greedy(bf_handle* h, const unsigned long long* edges, int n_slots, const float* centers, int N, const int32_t* order,
                               const int32_t* init_id, const float* poses, int32_t* fusion_list, int32_t* fusion_len, int32_t* fusion_flag,
                               float translation_gap, float rotation_gap_deg, float center_gap, int32_t* keep, int32_t* success,
                               int32_t* status, void* stream) {
    bf_device_guard guard(h);
    if (!h || N < 0 || N > 65536 || n_slots < 0 || n_slots > BF_EDGE_CAP) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_nms3d_greedy", "bad size");
    if (N == 0) return BF_OK;
    if (!edges || !centers || !order || !init_id || !poses || !fusion_list || !fusion_len || !fusion_flag || !keep || !success || !status)
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_nms3d_greedy", "null pointer");
    bf_record_ctx ctx;
    ctx.order = order; ctx.init_id = init_id; ctx.poses = poses; ctx.centers = centers; ctx.fl = fusion_list;
    ctx.flen = fusion_len; ctx.fflag = fusion_flag; ctx.keep = keep; ctx.status = status;
    ctx.translation_gap = translation_gap; ctx.rotation_gap = rotation_gap_deg; ctx.center_gap = center_gap;
    ctx.valid_num = nullptr;
    const int W = (N + 31) / 32;
    const size_t smem_e = sizeof(unsigned long long) * BF_EDGE_CAP + sizeof(uint32_t) * (size_t)W + BF_EDGE_CAP + 16;
    BF_CUDA(h, cudaFuncSetAttribute(bf_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
    bf_greedy_kernel<<<1, 1024, smem_e, (cudaStream_t)stream>>>(edges, nullptr, bf_dim_host(N), nullptr, nullptr, nullptr, ctx, success, n_slots);
    BF_LAUNCH_CHECK(h, "bf_greedy_kernel");
    return BF_OK;
}

extern "C" int bf_nms3d(bf_handle* h, const float* corners, const float* centers, int N, const int32_t* order,
                        const int32_t* init_id, const float* poses, int M, int32_t* fusion_list, int32_t* fusion_len,
                        int32_t* fusion_flag, double iou_threshold, float translation_gap, float rotation_gap_deg,
                        float center_gap, int mode, int32_t* keep, int32_t* success, int32_t* status, void* stream) {
    bf_device_guard guard(h);
    if (!h || N < 0 || M < 0 || N > 65536) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_nms3d", "bad size");
    if (N == 0) return BF_OK;
    if (!corners || !centers || !order || !init_id || !poses || !fusion_list || !fusion_len || !fusion_flag || !keep ||
        !success || !status)
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_nms3d", "null pointer");
    return bf_nms3d_run(h, corners, centers, bf_dim_host(N), order, nullptr, init_id, poses, fusion_list, fusion_len, fusion_flag,
                        iou_threshold, translation_gap, rotation_gap_deg, center_gap, mode, keep, success, status, nullptr, (cudaStream_t)stream);
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
    bf_device_guard guard(h);
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
