// K3 - IoU-guided particle refinement.
//
// Reference: the PyCUDA kernel `compute_iou_value` and helpers (box_fusion.py:68-405) plus the Python
// optimiser around it (evaluate_iou :413-461, cal_transform :475-535, update_PST :537-562,
// init_opt_params :566-600, boxfusion loop :651-721).  The reference launches one tiny kernel per
// optimiser iteration per box with 13 blocking copies each; here ONE launch refines every box: one
// thread-block cluster per map box, all iterations inside the kernel, one (view, particle) evaluation per
// work item (bf_refine_eval.cuh), per-view observation hulls and edge lines precomputed once in shared
// memory, contributions gathered in the cluster leader's shared memory through DSMEM, the ordered
// "first 200 better particles" rule done as one block-wide ballot/prefix scan and the float32 sums
// accumulated in the reference's index order.
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
#define BF_PST_SMEM_MAX 2048      // particle templates up to this size are staged in shared memory
#define BF_CNT_SLOTS 192          // ceil(P/T) * T/32 + 1 <= (BF_MAX_PARTICLES + BF_REFINE_THREADS) / 32 + 1 = 145

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

// value = sum over views of one particle's contributions in ascending view order (:400-401 with the host's grid order);
// loads are issued U at a time (4 from shared memory, 16 from the L2-resident global slab, whose latency is what the
// saturated regime's leader phase waits for), the additions stay sequential.  32-bit indexing and rolled loops on
// purpose: the leader phase is short and runs once per iteration, its cost is instruction fetch more than arithmetic.
template <int U>
__device__ __forceinline__ float bf_sum_views(const float* __restrict__ c, int stride, int V) {
    float value = 0.0f;
    int v = 0;
#pragma unroll 1
    for (; v + U <= V; v += U) {
        float t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = c[(v + u) * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) value += t[u];
    }
#pragma unroll 1
    for (; v < V; ++v) value += c[v * stride];
    return value;
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
    int dbg[2];                           // BF_REFINE_TIMING: max / sum of the cluster's per-warp evaluation cycles
    // followed by: bf_view views[max_views]; float fit[P]; int cnt[BF_CNT_SLOTS]; float terms[8][max_hits]; float spst[6*pst_cap]; float contrib[pair_cap]
};

extern __shared__ __align__(16) unsigned char bf_refine_smem_raw[];

// One thread-block CLUSTER per map box.  Work items of an optimiser iteration are spread over all CTAs of
// the cluster; every CTA writes its results straight into the leader's shared memory (DSMEM), the leader
// reduces in the reference's order and publishes the new state, two cluster barriers per iteration.
// One work item = one (view, particle); contributions are stored view-major and summed per particle in ascending
// view order by the leader (the reference's host order of the atomicAdd sum).
//
// Three instantiations of the same source, chosen per call by the problem size (measured on B200, tools/sweep_shapes.sh):
//   <false, 512, 1>  latency regime (a handful of boxes, the bench's per-keyframe call): fully unrolled evaluation,
//                    up to 128 registers, one CTA of up to 512 threads per SM, clusters of 16;
//   <false, 256, 3>  the call fills the machine a few times over (C1): same code held to 80 registers, three CTAs per SM;
//   <true,  256, 4>  saturated (C4): compact rolled loops (the unrolled evaluation is ~50 KB of SASS, more than the 32 KB
//                    L1.5 instruction cache; `no_instruction` was the second largest stall), 64 registers, four CTAs per SM.
template <bool ROLL, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
bf_refine_kernel(const bf_refine_params prm, int pair_cap, int max_views, int pst_cap, float* __restrict__ gcontrib, int timing) {
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
    int* cnt = (int*)(fit + prm.P);                       // per (round, warp) hit counts -> exclusive prefixes (+ total)
    float* terms = (float*)(cnt + BF_CNT_SLOTS);          // [8][max_hits] addends of cal_transform
    float* spst = terms + 8 * cfg.max_hits;               // particle template staged in shared memory when it fits (pst_cap = P)
    float* contrib = spst + 6 * pst_cap;
    if (tid == 0) { sm->overflow = 0; sm->dbg[0] = 0; sm->dbg[1] = 0; }
    if (V < 1 || V > max_views) {                        // flagged in status by bf_check_views_kernel2 (cluster-uniform)
        if (tid == 0 && crank == 0) { prm.out_updated[b] = 0; prm.out_iters[b] = 0; }
        return;
    }
    // leader's buffers as seen from this CTA
    float* l_fit = cluster.map_shared_rank(fit, 0);
    float* l_contrib = cluster.map_shared_rank(contrib, 0);
    const bf_refine_state* l_S = cluster.map_shared_rank(S, 0);

    // ---- stage the particle template and the views in every CTA: pose rows, observation hull (:367,375), area (:389)
    for (int k = tid; k < 6 * pst_cap; k += T) spst[k] = __ldg(prm.pst + k);
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
    const float beta = (float)cfg.beta, omb = (float)(1.0 - cfg.beta);
    int overflow = 0;
    int it = 0;
    cluster.sync();                                      // every CTA's shared memory is initialised
    for (int n = 0; n < cfg.iters; ++n) {
        long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0, tcA = 0, tcB = 0, tcP1 = 0, tcP2 = 0;
        if (timing) tc0 = clock64();
        // ---- evaluate_iou (:413-461): one work item = one (view, particle), spread over the whole cluster ----
        const int items = n_eval * V;
        for (int w = crank * T + tid; w < items; w += C * T) {
            const int v = w / n_eval, p = w - v * n_eval;              // view-major: a warp works on one view
            float pst6[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) pst6[k] = pst_cap ? spst[6 * p + k] : __ldg(prm.pst + 6 * p + k);
            float c[8][3];
            bf_particle_corners(S->box6, pst6, S->search, S->rot, c);
            wcontrib[w] = bf_eval_view<ROLL>(c, views[v], cfg.fx, cfg.cx, cfg.fy, cfg.cy, cfg.img_w, cfg.img_h, &overflow, nullptr);
        }
        ++it;
        if (timing) {
            tc1 = clock64();
            if ((tid & 31) == 0) {
                int* l_dbg = cluster.map_shared_rank(sm->dbg, 0);
                atomicMax(l_dbg, (int)(tc1 - tc0)); atomicAdd(l_dbg + 1, (int)((tc1 - tc0) >> 6));
            }
        }
        cluster.sync();
        if (timing) tc2 = clock64();
        if (crank == 0) {
            // ---- fitness per particle, views summed in ascending order (:400-401, :454), and cal_transform's ordered
            //      selection (:475-535): the first `max_hits` particles j >= 1 with fit[j] < fit[0] in index order.
            //      Particle j = r*T + tid (conflict-free shared-memory reads); its rank among the hits = hits of earlier
            //      (round, warp) groups + hits of lower lanes.  Pass 1: fitness + per-group ballot counts; warp 0 turns
            //      the counts into exclusive prefixes; pass 2: ranks and the eight addends of every selected particle.
            const int nw = T >> 5, lane = tid & 31, warp = tid >> 5;
            const int rounds = (prm.P + T - 1) / T;
            const float denom = (float)V + 1e-6f;                 // count += 1 per view, then count + 1e-6 (:454)
            const float unlaunched = 0.0f / (0.0f + 1e-6f);       // particles beyond 32*int(pst_size/32)
            float origin = unlaunched;
            if (n_eval >= 1) {
                origin = (in_smem ? bf_sum_views<4>(contrib, n_eval, V) : bf_sum_views<16>(gcontrib + (size_t)v0 * n_eval, n_eval, V)) / denom;
            }
#pragma unroll 1
            for (int r = 0; r < rounds; ++r) {
                const int j = r * T + tid;
                float f = unlaunched;
                if (j < n_eval) {
                    f = (in_smem ? bf_sum_views<4>(contrib + j, n_eval, V) : bf_sum_views<16>(gcontrib + (size_t)v0 * n_eval + j, n_eval, V)) / denom;
                }
                if (j < prm.P) fit[j] = f;
                const bool hit = (j >= 1 && j < prm.P) && (f < origin);
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) cnt[r * nw + warp] = __popc(bal);
            }
            __syncthreads();
            if (timing) tcP1 = clock64();
            if (warp == 0) {
                const int ng = rounds * nw;
                int carry = 0;
#pragma unroll 1
                for (int base = 0; base < ng; base += 32) {
                    const int x = (base + lane < ng) ? cnt[base + lane] : 0;
                    int incl = x;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
                    if (base + lane < ng) cnt[base + lane] = carry + incl - x;
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) cnt[ng] = carry;
            }
            __syncthreads();
            if (timing) tcP2 = clock64();
            const int hits = min(cnt[rounds * nw], cfg.max_hits);
#pragma unroll 1
            for (int r = 0; r < rounds; ++r) {
                if (cnt[r * nw] >= cfg.max_hits) break;             // block-uniform: every later rank is beyond the cap
                const int j = r * T + tid;
                const float f = (j < prm.P) ? fit[j] : 0.0f;
                const bool hit = (j >= 1 && j < prm.P) && (f < origin);
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                const int pos = cnt[r * nw + warp] + __popc(bal & ((1u << lane) - 1u));
                if (hit && pos < cfg.max_hits) {
                    const float w = origin - f;
#pragma unroll
                    for (int k = 0; k < 6; ++k) terms[k * cfg.max_hits + pos] = (pst_cap ? spst[6 * j + k] : __ldg(prm.pst + 6 * j + k)) * w;
                    terms[6 * cfg.max_hits + pos] = w;
                    terms[7 * cfg.max_hits + pos] = f * w;
                }
            }
            __syncthreads();
            if (timing) tcA = clock64();
            // ... accumulated sequentially in index order in float32, like the reference's Python loop (:490-515):
            // lanes 0..7 of warp 0 own one sum each and hand it to lane 0 by shuffle
            float acc8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (warp == 0) {
                float acc = 0.0f;
                if (lane < 8) {
                    const float* tq = terms + lane * cfg.max_hits;
                    int q = 0;
#pragma unroll 1
                    for (; q + 8 <= hits; q += 8) {                   // loads batched, additions in index order
                        float t8[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) t8[u] = tq[q + u];
#pragma unroll
                        for (int u = 0; u < 8; ++u) acc += t8[u];
                    }
#pragma unroll 1
                    for (; q < hits; ++q) acc += tq[q];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc8[k] = __shfl_sync(0xffffffffu, acc, k);
            }
            if (timing) tcB = clock64();
            if (tid == 0) {
                int success;
                float min_iou, mt[6] = {0, 0, 0, 0, 0, 0};
                float search[6], prev[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) { search[k] = S->search[k]; prev[k] = S->prev[k]; }
                const int previous_success = S->previous_success;
                if (hits <= 0) { success = 0; min_iou = origin; }
                else {
                    success = 1;
                    const float sw = acc8[6];
                    min_iou = acc8[7] / sw;
#pragma unroll
                    for (int k = 0; k < 6; ++k) mt[k] = (acc8[k] / sw) * search[k];
                }
                // update_PST (:537-562)
                const float ms = 1e-3f;
                float s[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) s[k] = fabsf(mt[k]) + ms;
                float n2 = s[0] * s[0];
#pragma unroll
                for (int k = 1; k < 6; ++k) n2 = n2 + s[k] * s[k];
                const float nrm = sqrtf(n2);
#pragma unroll
                for (int k = 3; k < 6; ++k) search[k] = cfg.shape_scale * min_iou * (s[k] / nrm) + ms;
#pragma unroll
                for (int k = 0; k < 3; ++k) search[k] = cfg.center_scale * min_iou * (s[k] / nrm) + ms;
                if (previous_success && success) {                                                    // :685-691
#pragma unroll
                    for (int k = 0; k < 6; ++k) search[k] = beta * search[k] + omb * prev[k];
                }
#pragma unroll
                for (int k = 0; k < 6; ++k) S->search[k] = search[k];
                if (success) {                                                                        // :694-706
                    S->need_update = 1; S->previous_success = 1; S->fail = 0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) { const double g = S->g[k] + (double)mt[k]; S->g[k] = g; S->box6[k] = (float)g; S->prev[k] = search[k]; }
                    S->done = 0;
                } else {
                    const int fail = S->fail + 1;
                    S->fail = fail; S->previous_success = 0;
                    S->done = (cfg.early_stop && fail >= 3) ? 1 : 0;                                  // :713
                }
                if (prm.trace) {
                    float* tr = prm.trace + ((size_t)b * cfg.iters + n) * 8;
                    tr[0] = (float)success; tr[1] = min_iou;
#pragma unroll
                    for (int k = 0; k < 6; ++k) tr[2 + k] = search[k];
                }
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
            tr[6] = (float)(tcA - tc2);                                   // leader phase: selection part
            if (timing == 2) { tr[2] = (float)(tcP1 - tc2); tr[3] = (float)(tcP2 - tcP1); tr[5] = (float)(tcA - tcP2); }   // finer split of the selection
            tr[7] = (float)sm->dbg[0];                                    // slowest warp's evaluation cycles in the cluster
            tr[1] = (float)sm->dbg[1] * 64.0f / (float)(C * (T >> 5));    // mean warp evaluation cycles (overwrites min_iou in timing mode)
            sm->dbg[0] = 0; sm->dbg[1] = 0; (void)tcB;
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

static size_t bf_refine_smem_bytes(int P, int max_hits, int pair_cap, int max_views, int pst_cap) {
    return sizeof(bf_refine_smem) + sizeof(bf_view) * (size_t)max_views + sizeof(float) * (size_t)P +
           sizeof(int) * (size_t)BF_CNT_SLOTS + sizeof(float) * 8 * (size_t)max_hits + sizeof(float) * (size_t)pair_cap + sizeof(float) * 6 * (size_t)pst_cap + 16;
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
    // global contribution scratch: sum(V) * n_eval floats (only touched by boxes that do not fit shared memory)
    void* gscratch = nullptr;
    {
        const long long n_eval_h = (32 * (cfg->pst_size / 32) < P) ? 32 * (cfg->pst_size / 32) : P;
        const long long views_total = cfg->views_total > 0 ? (long long)cfg->views_total : (long long)B * BF_MAX_VIEWS;
        int rc = bf_scratch(h, BF_SCRATCH_REFINE, sizeof(float) * (size_t)(views_total * n_eval_h), &gscratch);
        if (rc) return rc;
    }
    typedef void (*bf_refine_fn)(const bf_refine_params, int, int, int, float*, int);
    static const bf_refine_fn kernels[3] = {bf_refine_kernel<false, 512, 1>, bf_refine_kernel<false, 256, 3>, bf_refine_kernel<true, 256, 4>};
    static const int kernel_max_t[3] = {512, 256, 256};
    // Launch shape.  A box's optimiser iteration has items = n_eval * V independent evaluations followed by a short
    // leader phase, so its latency is passes = ceil(items / (C*T)) evaluations; B boxes need waves = ceil(B / clusters
    // that fit the machine).  Latency regime: pick the (cluster size C, block size T) that minimises waves * passes,
    // preferring fewer threads on ties; occupancy answers are cached in the handle.
    const int n_eval = n_eval0;
    const long long items = (long long)n_eval * (cfg->views_total > 0 ? (cfg->views_total + B - 1) / B : max_views);
    // the call returns when its slowest box does: latency is set by the box with the most views (max_views is the caller's bound)
    const long long items_max = (long long)n_eval * max_views;
    static const int Cs[5] = {16, 8, 4, 2, 1};
    static const int Ts[4] = {512, 384, 256, 128};
    int bestC = 1, bestT = 256, variant = 0;
    double best_cost = 1e300;
    // throughput regimes (the call alone fills the machine): 256-thread CTAs, several per SM so that one box's leader phase
    // and cluster barriers overlap another box's evaluations; cluster just large enough for ~2 items per thread
    const bool saturated = (double)B * (double)items >= (double)h->sm_count * 512.0;
    if (saturated) {
        variant = ((double)B * (double)items >= (double)h->sm_count * 8192.0) ? 2 : 1;
        bestT = 256;
        while (bestC < 16 && (long long)bestC * bestT * 2 < items) bestC *= 2;
        while (bestC > 1 && (long long)B * bestC > 8LL * h->sm_count) bestC /= 2;
        best_cost = 0.0;
    }
    if (!saturated && h->refine_concurrent) {                      // BF_OPT_REFINE_CONCURRENT: leave room for the other streams' kernels
        variant = 1; bestT = 256; bestC = 16;
        while (bestC > 1 && (long long)(bestC / 2) * bestT >= items_max) bestC /= 2;
        best_cost = 0.0;
    }
    if (h->refine_force_c > 0 && h->refine_force_t > 0) {          // BF_REFINE_SHAPE (tuning sweeps)
        variant = (h->refine_force_variant >= 0 && h->refine_force_variant < 3) ? h->refine_force_variant : variant;
        if (h->refine_force_t <= kernel_max_t[variant]) { bestC = h->refine_force_c; bestT = h->refine_force_t; best_cost = 0.0; }
    }
    const bf_refine_fn kern = kernels[variant];
    // the particle template itself in shared memory when it is small (24 KB at P = 1024) - latency regime only: it shortens
    // the leader phase, but in the throughput regimes the space is worth more as resident CTAs and L1
    const int pst_cap = (variant == 0 && P <= BF_PST_SMEM_MAX) ? P : 0;
    // rounded up to 8 KB so that the cached occupancy answers below are reused across keyframes
    const size_t smem = (bf_refine_smem_bytes(P, cfg->max_hits, pair_cap, max_views, pst_cap) + 8191) / 8192 * 8192;
    BF_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BF_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int ci = 0; ci < 5 && best_cost > 0.0; ++ci)
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
                if (cudaOccupancyMaxActiveClusters(&active, kern, &lc) != cudaSuccess) { cudaGetLastError(); active = 0; }
                h->refine_occ[slot] = active; h->refine_occ_smem[slot] = (long long)smem + 1;
            }
            if (active < 1) continue;
            const long long waves = (B + active - 1) / active;
            const long long passes = (items_max + (long long)C * T - 1) / ((long long)C * T);
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
        cudaError_t e = cudaLaunchKernelEx(&lc, kern, prm, pair_cap, max_views, pst_cap, (float*)gscratch, h->refine_timing);
        if (e != cudaSuccess) return bf_fail(h, BF_ERR_CUDA, "bf_refine_kernel", cudaGetErrorString(e));
        h->last_refine_cluster = variant * 1000000 + bestC * 1000 + bestT;
    }
    return BF_OK;
}

extern "C" int bf_refine_last_launch(bf_handle* h) { return h ? h->last_refine_cluster : 0; }

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
                value += bf_eval_view<false>(c, views[v], cfg.fx, cfg.cx, cfg.fy, cfg.cy, cfg.img_w, cfg.img_h, &overflow, nullptr);
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
