// K3 - IoU-guided particle refinement.
//
// Reference: the PyCUDA kernel `compute_iou_value` and helpers (box_fusion.py:68-405) plus the Python
// optimiser around it (evaluate_iou :413-461, cal_transform :475-535, update_PST :537-562,
// init_opt_params :566-600, boxfusion loop :651-721).  The reference launches one tiny kernel per
// optimiser iteration per box with 13 blocking copies each; here ONE launch refines every box: one
// CTA per map box, all iterations inside the kernel, particles in registers, per-view observation
// hulls precomputed once in shared memory, the ordered "first 200 better particles" rule done with
// a ballot/popc prefix selection and the float32 sums accumulated in the reference's index order.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: every float expression below is evaluated
// with the same IEEE operations, in the same order, as the reference kernel compiled without
// contraction (the CPU oracle), so fitness values, accept/reject decisions and fused boxes are
// bit-identical to the oracle - not merely within tolerance.  Doubles appear exactly where the
// reference uses them (line_intersection, the 0.00001 literal, the float64 box state).
#include "bf_common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#include "bf_refine_eval.cuh"

#define BF_REFINE_THREADS 512    // upper bound of the block size; the launch picks 128..512 per call

// numpy pairwise_sum for float32 (n <= 128), see oracle/refine_oracle.c
__device__ float bf_np_pairwise_sum(const float* a, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    float r[8];
    int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

struct bf_refine_params {
    const float* pst; int P;
    const float* per_xyzlhw; const float* per_R; const float* per_scores; const float* per_uv; const float* per_poses;
    const int32_t* view_offsets; const int32_t* view_index; int B;
    bf_refine_cfg cfg;
    float* out_xyzlhw; int32_t* out_updated; int32_t* out_iters; float* trace; int32_t* status;
};

struct bf_refine_state {     // optimiser state of one box; the cluster leader's copy is authoritative
    double g[6];
    float box6[6];
    float rot[9];
    float search[6], prev[6];
    int previous_success, fail, need_update, done;
    float acc[8];
};

// ---- shared-memory layout (identical in every CTA of a cluster so that DSMEM offsets match) ------------
struct __align__(16) bf_refine_smem {
    bf_refine_state S;
    int warp_cnt[32];
    float vbox[6 * BF_MAX_VIEWS];        // gathered view boxes (init_opt_params)
    float vscore[BF_MAX_VIEWS];
    float col[3 * BF_MAX_VIEWS];
    int overflow;
    // followed by: bf_view views[max_views]; float fit[P]; int sel[max_hits]; float terms[8][max_hits]; float contrib[pair_cap]
};

extern __shared__ __align__(16) unsigned char bf_refine_smem_raw[];

// One thread-block CLUSTER per map box.  Work items of an optimiser iteration are spread over all CTAs of
// the cluster; every CTA writes its results straight into the leader's shared memory (DSMEM), the leader
// reduces in the reference's order and publishes the new state, two cluster barriers per iteration.
// One work item = one (view, particle); contributions are stored view-major and summed per particle in ascending
// view order by the leader (the reference's host order of the atomicAdd sum).
__global__ void __launch_bounds__(BF_REFINE_THREADS, 1)
bf_refine_kernel(const bf_refine_params prm, int pair_cap, int max_views, float* __restrict__ gcontrib, int timing) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned C = cluster.num_blocks();
    const unsigned crank = cluster.block_rank();
    const int b = blockIdx.x / C;
    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int v0 = prm.view_offsets[b];
    const int V = prm.view_offsets[b + 1] - v0;
    const bf_refine_cfg& cfg = prm.cfg;
    bf_refine_smem* sm = (bf_refine_smem*)bf_refine_smem_raw;
    bf_refine_state* S = &sm->S;
    bf_view* views = (bf_view*)(sm + 1);
    float* fit = (float*)(views + max_views);
    int* sel = (int*)(fit + prm.P);
    float* terms = (float*)(sel + cfg.max_hits);          // [8][max_hits] addends of cal_transform
    float* contrib = terms + 8 * cfg.max_hits;
    if (tid == 0) sm->overflow = 0;
    if (V < 1 || V > max_views) {                        // flagged in status by bf_check_views_kernel2 (cluster-uniform)
        if (tid == 0 && crank == 0) { prm.out_updated[b] = 0; prm.out_iters[b] = 0; }
        return;
    }
    // leader's buffers as seen from this CTA
    float* l_fit = cluster.map_shared_rank(fit, 0);
    float* l_contrib = cluster.map_shared_rank(contrib, 0);
    const bf_refine_state* l_S = cluster.map_shared_rank(S, 0);

    // ---- stage the views in every CTA: pose rows, observation hull (:367,375) and its area (:389) -------
    for (int v = tid; v < V; v += T) {
        const int m = prm.view_index[v0 + v];
        bf_view_stage(views[v], prm.per_poses + 16 * (size_t)m, prm.per_uv + 16 * (size_t)m, cfg.img_w, cfg.img_h);
#pragma unroll
        for (int k = 0; k < 6; ++k) sm->vbox[6 * v + k] = prm.per_xyzlhw[6 * (size_t)m + k];
        sm->vscore[v] = prm.per_scores[m];
    }
    __syncthreads();

    // ---- init_opt_params (:566-600) + init_searchsize (:468-472): thread 0 of every CTA (same result) ----
    if (tid == 0) {
        const float* vbox = sm->vbox; const float* vscore = sm->vscore; float* col = sm->col;
        int best = 0;
        for (int v = 1; v < V; ++v) if (vscore[v] > vscore[best]) best = v;
        for (int k = 0; k < 3; ++k) {
            float acc = vbox[k];
            for (int v = 1; v < V; ++v) acc += vbox[6 * v + k];
            S->g[k] = (double)(acc / (float)V);
        }
        const float* bd = vbox + 6 * best + 3;
        int order[3] = {0, 1, 2}, rank[3];
        for (int i = 1; i < 3; ++i) { const int k = order[i]; int j = i - 1; while (j >= 0 && bd[order[j]] > bd[k]) { order[j + 1] = order[j]; --j; } order[j + 1] = k; }
        for (int i = 0; i < 3; ++i) rank[order[i]] = i;
        for (int v = 0; v < V; ++v) {
            float d[3] = {vbox[6 * v + 3], vbox[6 * v + 4], vbox[6 * v + 5]};
            for (int i = 1; i < 3; ++i) { const float k = d[i]; int j = i - 1; while (j >= 0 && d[j] > k) { d[j + 1] = d[j]; --j; } d[j + 1] = k; }
            for (int k = 0; k < 3; ++k) col[k * V + v] = d[rank[k]];
        }
        for (int k = 0; k < 3; ++k) S->g[3 + k] = (double)(bf_np_pairwise_sum(col + k * V, V) / (float)V);
        const int mb = prm.view_index[v0 + best];
        for (int k = 0; k < 9; ++k) S->rot[k] = prm.per_R[9 * (size_t)mb + k];
        for (int k = 0; k < 3; ++k) { S->search[k] = cfg.center_init; S->search[3 + k] = cfg.shape_init; S->prev[k] = 0.f; S->prev[3 + k] = 0.f; }
        S->previous_success = 0; S->fail = 0; S->need_update = 0; S->done = 0;
        for (int k = 0; k < 6; ++k) S->box6[k] = (float)S->g[k];
    }
    __syncthreads();

    const int n_eval = min(32 * (cfg.pst_size / 32), prm.P);
    // contributions |1-iou| of every (view, particle), view-major.  Small problems keep them in the leader's shared
    // memory (written through DSMEM); large ones (C4: 4096 x 32) use an L2-resident global scratch slab of this box.
    const bool in_smem = (long long)n_eval * V <= (long long)pair_cap;
    float* wcontrib = in_smem ? l_contrib : gcontrib + (size_t)v0 * n_eval;        // where this CTA writes
    const float* rcontrib = in_smem ? contrib : gcontrib + (size_t)v0 * n_eval;    // where the leader reads
    const float beta = (float)cfg.beta, omb = (float)(1.0 - cfg.beta);
    int overflow = 0;
    int it = 0;
    cluster.sync();                                      // every CTA's shared memory is initialised
    for (int n = 0; n < cfg.iters; ++n) {
        long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
        if (timing) tc0 = clock64();
        // ---- evaluate_iou (:413-461): one work item = one (view, particle), spread over the whole cluster ----
        const int items = n_eval * V;
        for (int w = crank * T + tid; w < items; w += C * T) {
            const int v = w / n_eval, p = w - v * n_eval;              // view-major: a warp works on one view
            float pst6[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) pst6[k] = __ldg(prm.pst + 6 * (size_t)p + k);
            float c[8][3];
            bf_particle_corners(S->box6, pst6, S->search, S->rot, c);
            wcontrib[w] = bf_eval_view(c, views[v], cfg.fx, cfg.cx, cfg.fy, cfg.cy, cfg.img_w, cfg.img_h, &overflow, nullptr);
        }
        ++it;
        if (timing) tc1 = clock64();
        cluster.sync();
        if (timing) tc2 = clock64();
        if (crank == 0) {
            // ---- fitness per particle, views summed in ascending order (:400-401, :454) ---------------------
            for (int p = tid; p < n_eval; p += T) {
                float value = 0.0f, count = 0.0f;
                for (int v = 0; v < V; ++v) { value += rcontrib[v * n_eval + p]; count += 1; }
                fit[p] = value / (count + 1e-6f);
            }
            for (int p = n_eval + tid; p < prm.P; p += T) fit[p] = 0.0f / (0.0f + 1e-6f);   // never launched
            __syncthreads();
            // ---- cal_transform (:475-535): first `max_hits` particles j >= 1 with fit[j] < fit[0], index order
            const float origin = fit[0];
            int total = 0;
            for (int base = 0; base < prm.P && total < cfg.max_hits; base += T) {
                const int j = base + tid;
                const bool hit = (j >= 1 && j < prm.P) && (fit[j] < origin);
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                const int lane = tid & 31, warp = tid >> 5;
                if (lane == 0) sm->warp_cnt[warp] = __popc(bal);
                __syncthreads();
                int before = total, all = 0;
                const int nw = T >> 5;
                for (int w = 0; w < nw; ++w) { const int cw = sm->warp_cnt[w]; if (w < warp) before += cw; all += cw; }
                const int pos = before + __popc(bal & ((1u << lane) - 1u));
                if (hit && pos < cfg.max_hits) sel[pos] = j;
                total += all;
                __syncthreads();
            }
            const int hits = min(total, cfg.max_hits);
            // the eight addends of every selected particle (6 x PST*w, w, fitness*w) are formed in parallel ...
            for (int q = tid; q < hits; q += T) {
                const int j = sel[q];
                const float w = origin - fit[j];
#pragma unroll
                for (int k = 0; k < 6; ++k) terms[k * cfg.max_hits + q] = __ldg(prm.pst + 6 * (size_t)j + k) * w;
                terms[6 * cfg.max_hits + q] = w;
                terms[7 * cfg.max_hits + q] = fit[j] * w;
            }
            __syncthreads();
            // ... and accumulated sequentially in index order in float32, like the reference's Python loop (:490-515)
            if (tid < 8) {
                float acc = 0.0f;
                const float* tq = terms + tid * cfg.max_hits;
                for (int q = 0; q < hits; ++q) acc += tq[q];
                S->acc[tid] = acc;
            }
            __syncthreads();
            if (tid == 0) {
                int success;
                float min_iou, mt[6] = {0, 0, 0, 0, 0, 0};
                if (hits <= 0) { success = 0; min_iou = origin; }
                else {
                    success = 1;
                    const float sw = S->acc[6];
                    min_iou = S->acc[7] / sw;
                    for (int k = 0; k < 6; ++k) mt[k] = (S->acc[k] / sw) * S->search[k];
                }
                // update_PST (:537-562)
                const float ms = 1e-3f;
                float s[6];
                for (int k = 0; k < 6; ++k) s[k] = fabsf(mt[k]) + ms;
                float n2 = s[0] * s[0];
                for (int k = 1; k < 6; ++k) n2 = n2 + s[k] * s[k];
                const float nrm = sqrtf(n2);
                for (int k = 3; k < 6; ++k) S->search[k] = cfg.shape_scale * min_iou * (s[k] / nrm) + ms;
                for (int k = 0; k < 3; ++k) S->search[k] = cfg.center_scale * min_iou * (s[k] / nrm) + ms;
                if (S->previous_success && success)                                                  // :685-691
                    for (int k = 0; k < 6; ++k) S->search[k] = beta * S->search[k] + omb * S->prev[k];
                if (success) {                                                                        // :694-706
                    S->need_update = 1; S->previous_success = 1; S->fail = 0;
                    for (int k = 0; k < 6; ++k) { S->g[k] += (double)mt[k]; S->prev[k] = S->search[k]; }
                } else { S->fail += 1; S->previous_success = 0; }
                if (prm.trace) {
                    float* tr = prm.trace + ((size_t)b * cfg.iters + n) * 8;
                    tr[0] = (float)success; tr[1] = min_iou;
                    for (int k = 0; k < 6; ++k) tr[2 + k] = S->search[k];
                }
                for (int k = 0; k < 6; ++k) S->box6[k] = (float)S->g[k];
                S->done = (cfg.early_stop && S->fail >= 3) ? 1 : 0;                                   // :713
            }
        }
        if (timing) tc3 = clock64();
        cluster.sync();                                  // leader state published
        if (crank != 0) {
            if (tid < 6) { S->box6[tid] = l_S->box6[tid]; S->search[tid] = l_S->search[tid]; }
            if (tid == 6) S->done = l_S->done;
        }
        __syncthreads();
        if (timing && prm.trace && crank == 0 && tid == 0) {   // diagnostic: cycles of {own evaluations, wait for the cluster, leader phase, publish}
            float* tr = prm.trace + ((size_t)b * cfg.iters + n) * 8;
            tr[2] = (float)(tc1 - tc0); tr[3] = (float)(tc2 - tc1); tr[4] = (float)(tc3 - tc2); tr[5] = (float)(clock64() - tc3);
        }
        if (S->done) break;
    }
    if (overflow) atomicExch(&sm->overflow, 1);
    cluster.sync();                                      // nobody reads the leader's shared memory after this
    if (tid == 0) {
        if (sm->overflow) atomicExch(prm.status, BF_ERR_CAPACITY);
        if (crank == 0) {
            prm.out_iters[b] = it;
            prm.out_updated[b] = S->need_update;
            if (S->need_update) {                                                                     // :716-721
                for (int k = 3; k < 6; ++k) if (S->g[k] < 0.01) S->g[k] = 0.01;
                for (int k = 0; k < 6; ++k) prm.out_xyzlhw[6 * (size_t)b + k] = (float)S->g[k];
            } else {
                for (int k = 0; k < 6; ++k) prm.out_xyzlhw[6 * (size_t)b + k] = 0.0f;
            }
        }
    }
}

#define BF_PAIR_CAP 8192       // (view, particle) contributions the leader holds in shared memory: 32 KB

static size_t bf_refine_smem_bytes(int P, int max_hits, int pair_cap, int max_views) {
    return sizeof(bf_refine_smem) + sizeof(bf_view) * (size_t)max_views + sizeof(float) * (size_t)P +
           sizeof(int) * (size_t)max_hits + sizeof(float) * 8 * (size_t)max_hits + sizeof(float) * (size_t)pair_cap + 16;
}

__global__ void bf_check_views_kernel(const int32_t* __restrict__ off, int B, int32_t* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) status[0] = 0;
}
__global__ void bf_check_views_kernel2(const int32_t* __restrict__ off, int B, int max_views, int32_t* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { const int V = off[b + 1] - off[b]; if (V < 1 || V > max_views) atomicExch(status, BF_ERR_CAPACITY); }
}

extern "C" int bf_refine(bf_handle* h, const float* pst, int P, const float* per_xyzlhw, const float* per_R,
                         const float* per_scores, const float* per_uv, const float* per_poses, int M,
                         const int32_t* view_offsets, const int32_t* view_index, int B, const bf_refine_cfg* cfg,
                         float* out_xyzlhw, int32_t* out_updated, int32_t* out_iters, float* trace, int32_t* status,
                         void* stream) {
    if (!h || !cfg || B < 0 || P < 1 || P > BF_MAX_PARTICLES) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_refine", "bad size");
    if (!status) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_refine", "null status");
    cudaStream_t st = (cudaStream_t)stream;
    bf_check_views_kernel<<<1, 32, 0, st>>>(view_offsets, B, status);
    if (B == 0) return BF_OK;
    if (!pst || !per_xyzlhw || !per_R || !per_scores || !per_uv || !per_poses || !view_offsets || !view_index ||
        !out_xyzlhw || !out_updated || !out_iters)
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_refine", "null pointer");
    if (cfg->max_hits < 1 || cfg->max_hits > 4096 || cfg->iters < 1) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_refine", "bad cfg");
    const int max_views = (cfg->max_views > 0 && cfg->max_views <= BF_MAX_VIEWS) ? cfg->max_views : BF_MAX_VIEWS;
    bf_check_views_kernel2<<<bf_blocks(B, 128), 128, 0, st>>>(view_offsets, B, max_views, status);
    bf_refine_params prm;
    prm.pst = pst; prm.P = P; prm.per_xyzlhw = per_xyzlhw; prm.per_R = per_R; prm.per_scores = per_scores;
    prm.per_uv = per_uv; prm.per_poses = per_poses; prm.view_offsets = view_offsets; prm.view_index = view_index;
    prm.B = B; prm.cfg = *cfg; prm.out_xyzlhw = out_xyzlhw; prm.out_updated = out_updated; prm.out_iters = out_iters;
    prm.trace = trace; prm.status = status;
    // shared memory is sized per launch (what is not shared memory is L1 for the per-thread polygon buffers):
    // contributions live in the leader's shared memory only when every box of the call fits BF_PAIR_CAP
    const int n_eval0 = (32 * (cfg->pst_size / 32) < P) ? 32 * (cfg->pst_size / 32) : P;
    const int pair_cap = ((long long)n_eval0 * max_views <= BF_PAIR_CAP) ? n_eval0 * max_views : 0;
    // rounded up to 8 KB so that the cached occupancy answers below are reused across keyframes
    const size_t smem = (bf_refine_smem_bytes(P, cfg->max_hits, pair_cap, max_views) + 8191) / 8192 * 8192;
    // global contribution scratch: sum(V) * n_eval floats (only touched by boxes that do not fit shared memory)
    void* gscratch = nullptr;
    {
        const long long n_eval_h = (32 * (cfg->pst_size / 32) < P) ? 32 * (cfg->pst_size / 32) : P;
        const long long views_total = cfg->views_total > 0 ? (long long)cfg->views_total : (long long)B * BF_MAX_VIEWS;
        int rc = bf_scratch(h, BF_SCRATCH_REFINE, sizeof(float) * (size_t)(views_total * n_eval_h), &gscratch);
        if (rc) return rc;
    }
    BF_CUDA(h, cudaFuncSetAttribute(bf_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BF_CUDA(h, cudaFuncSetAttribute(bf_refine_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // Launch shape.  A box's optimiser iteration has items = n_eval * V independent evaluations followed by a short
    // leader phase, so its latency is passes = ceil(items / (C*T)) evaluations; B boxes need waves = ceil(B / clusters
    // that fit the machine).  Pick the (cluster size C, block size T) that minimises waves * passes, preferring fewer
    // threads on ties.  Occupancy answers are cached in the handle.
    const int n_eval = n_eval0;
    const long long items = (long long)n_eval * (cfg->views_total > 0 ? (cfg->views_total + B - 1) / B : max_views);
    static const int Cs[5] = {16, 8, 4, 2, 1};
    static const int Ts[4] = {512, 384, 256, 128};
    int bestC = 1, bestT = 256;
    double best_cost = 1e300;
    // throughput regime (the call alone fills the machine): 256-thread CTAs, two per SM so that one box's leader phase
    // and cluster barriers overlap another box's evaluations; cluster just large enough for ~2 items per thread
    const bool saturated = (double)B * (double)items >= (double)h->sm_count * 512.0;
    if (saturated) {
        bestT = 256;
        while (bestC < 16 && (long long)bestC * bestT * 2 < items) bestC *= 2;
        while (bestC > 1 && (long long)B * bestC > 4LL * h->sm_count) bestC /= 2;
        best_cost = 0.0;
    }
    for (int ci = 0; ci < 5 && !saturated; ++ci)
        for (int ti = 0; ti < 4; ++ti) {
            const int C = Cs[ci], T = Ts[ti];
            int active = 0;
            const int slot = ci * 4 + ti;
            if (h->refine_occ_smem[slot] == (long long)smem + 1) active = h->refine_occ[slot];
            else {
                cudaLaunchConfig_t lc = {};
                lc.gridDim = dim3((unsigned)(C * 64)); lc.blockDim = dim3(T); lc.dynamicSmemBytes = smem; lc.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                lc.attrs = at; lc.numAttrs = 1;
                if (cudaOccupancyMaxActiveClusters(&active, bf_refine_kernel, &lc) != cudaSuccess) { cudaGetLastError(); active = 0; }
                h->refine_occ[slot] = active; h->refine_occ_smem[slot] = (long long)smem + 1;
            }
            if (active < 1) continue;
            const long long passes = (items + (long long)C * T - 1) / ((long long)C * T);
            const long long waves = (B + active - 1) / active;
            const double cost = (double)waves * (double)passes + 1e-6 * C * T;      // ties -> fewer threads
            if (cost < best_cost) { best_cost = cost; bestC = C; bestT = T; }
        }
    {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)(B * bestC)); lc.blockDim = dim3(bestT); lc.dynamicSmemBytes = smem; lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = bestC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&lc, bf_refine_kernel, prm, pair_cap, max_views, (float*)gscratch, h->refine_timing);
        if (e != cudaSuccess) return bf_fail(h, BF_ERR_CUDA, "bf_refine_kernel", cudaGetErrorString(e));
        h->last_refine_cluster = bestC * 1000 + bestT;
    }
    return BF_OK;
}

// ---- evaluate_iou as a stand-alone entry (tests / diagnostics) -------------------------------------
__global__ void __launch_bounds__(BF_REFINE_THREADS)
bf_evaluate_kernel(const float* __restrict__ pst, int P, const float* __restrict__ box6, const float* __restrict__ rot9,
                   const float* __restrict__ uv, const float* __restrict__ poses, int V, const float* __restrict__ search6,
                   const bf_refine_cfg cfg, float* __restrict__ fitness) {
    __shared__ bf_view views[BF_MAX_VIEWS];
    __shared__ bf_refine_state S;
    const int tid = threadIdx.x;
    for (int v = tid; v < V; v += blockDim.x) {
        bf_view_stage(views[v], poses + 16 * v, uv + 16 * v, cfg.img_w, cfg.img_h);
    }
    if (tid == 0) {
        for (int k = 0; k < 6; ++k) { S.box6[k] = box6[k]; S.search[k] = search6[k]; }
        for (int k = 0; k < 9; ++k) S.rot[k] = rot9[k];
    }
    __syncthreads();
    int overflow = 0;
    const int n_eval = min(32 * (cfg.pst_size / 32), P);
    // one CTA per slice of particles: shift the particle loop by blockIdx
    for (int p = blockIdx.x * blockDim.x + tid; p < P; p += gridDim.x * blockDim.x) {
        float value = 0.0f, count = 0.0f;
        if (p < n_eval) {
            float pst6[6];
            for (int k = 0; k < 6; ++k) pst6[k] = pst[6 * (size_t)p + k];
            float c[8][3];
            bf_particle_corners(S.box6, pst6, S.search, S.rot, c);
            for (int v = 0; v < V; ++v) {
                value += bf_eval_view(c, views[v], cfg.fx, cfg.cx, cfg.fy, cfg.cy, cfg.img_w, cfg.img_h, &overflow, nullptr);
                count += 1;
            }
        }
        fitness[p] = value / (count + 1e-6f);
    }
}

extern "C" int bf_evaluate_iou(bf_handle* h, const float* pst, int P, const float* box6, const float* rot9, const float* uv,
                               const float* poses, int V, const float* search6, const bf_refine_cfg* cfg, float* fitness,
                               void* stream) {
    if (!h || !cfg || P < 1 || V < 1 || V > BF_MAX_VIEWS || !pst || !box6 || !rot9 || !uv || !poses || !search6 || !fitness)
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_evaluate_iou", "bad argument");
    const int grid = bf_blocks(P, BF_REFINE_THREADS);
    bf_evaluate_kernel<<<grid, BF_REFINE_THREADS, 0, (cudaStream_t)stream>>>(pst, P, box6, rot9, uv, poses, V, search6, *cfg, fitness);
    BF_LAUNCH_CHECK(h, "bf_evaluate_kernel");
    return BF_OK;
}
