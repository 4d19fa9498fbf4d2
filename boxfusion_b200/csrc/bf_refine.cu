// K3 - IoU-guided particle refinement.
//
// Reference: the PyCUDA kernel `compute_iou_value` and helpers (box_fusion.py:68-405) plus the Python
// optimiser around it (evaluate_iou :413-461, cal_transform :475-535, update_PST :537-562,
// init_opt_params :566-600, boxfusion loop :651-721).  The reference launches one tiny kernel per
// optimiser iteration per box with 13 blocking copies each; here ONE launch refines every box: one
// thread-block cluster per map box, all iterations inside the kernel, one (view, particle) evaluation per
// work item (bf_refine_eval.cuh), per-view observation hulls and edge lines precomputed once in shared
// memory, every CTA owning a block of particles with all their views (view sums stay CTA-local), fitness
// values exchanged between the CTAs of the cluster through DSMEM, the ordered "first 200 better particles"
// rule done as one block-wide ballot/prefix scan and the float32 sums accumulated in the reference's index order.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: every float expression below is evaluated
// with the same IEEE operations, in the same order, as the reference kernel compiled without
// contraction (the CPU oracle), so fitness values, accept/reject decisions and fused boxes are
// bit-identical to the oracle - not merely within tolerance.  Doubles appear exactly where the
// reference uses them (line_intersection, the 0.00001 literal, the float64 box state).
#include "bf_internal.cuh"
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
// loads are issued U at a time from CTA-local shared memory, the additions stay sequential.  32-bit indexing and rolled
// loops on purpose: this runs once per iteration, its cost is instruction fetch more than arithmetic.
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
    const int32_t* B_dev;        // engine step: the box count lives in device memory (B is then the bound)
    const float* intr_dev;       // engine step: fx, fy, cx, cy, W, H of the current keyframe in device memory
};

struct bf_refine_state {     // optimiser state of one box; every CTA of the cluster keeps an identical copy
    double g[6];
    float box6[6];
    float rot[9];
    float search[6], prev[6];
    int previous_success, fail, need_update, done;
};

// ---- shared-memory layout ---------------------------------------------------------------------------------------
struct __align__(16) bf_refine_smem {
    bf_refine_state S;
    float vbox[6 * BF_MAX_VIEWS];        // gathered view boxes (init_opt_params)
    float vscore[BF_MAX_VIEWS];
    float col[3 * BF_MAX_VIEWS];
    int dbg[2];                           // BF_REFINE_TIMING: max / sum of this CTA's per-warp evaluation cycles
    // followed by: bf_view views[max_views]; float fit[2][P]; int cnt[BF_CNT_SLOTS]; float terms[8][max_hits];
    //              float spst[6*pst_cap]; float contrib[contrib_cap]
};

extern __shared__ __align__(16) unsigned char bf_refine_smem_raw[];

// One thread-block CLUSTER per map box (persistent: cluster g of G takes boxes g, g+G, ...; a stand-alone call launches
// one cluster per box).  Work decomposition of one optimiser iteration (round 2):
//   * CTA c of the cluster OWNS the particle block [c*PB, (c+1)*PB) (PB a multiple of 32) with ALL V views of the box, so
//     the per-particle sum over views - ascending view order, the reference's host order of its atomicAdd sum - never
//     leaves the CTA: either every (view, particle) term goes to CTA-local shared memory (view-major, one term per thread
//     pass: the latency regime, every thread busy) and PB threads sum their column, or one thread keeps a particle
//     for all its views and accumulates in a register (throughput regime: no contribution storage at all, the particle's
//     corners are generated once instead of once per view).  Round 1 wrote every term to the leader's shared memory or to
//     a global slab (1.5 GB of DRAM writes per C4 launch) and had the leader re-read V terms per particle.
//   * every CTA then stores its PB fitness values into the fit[] vector of EVERY CTA of the cluster (DSMEM), one
//     cluster barrier, and every CTA runs cal_transform's ordered first-200 selection, update_PST and the state update
//     itself - deterministic, so all copies of the state stay identical.  One cluster barrier per iteration instead of
//     two, no leader phase the other CTAs wait for, no publish step.  fit[] is double-buffered by iteration parity: a CTA
//     can run at most one barrier ahead of the slowest one.
//
// Three instantiations of the same source, chosen per call by the problem size (measured on B200, tools/sweep_shapes.sh):
//   <false, 512, 1>  latency regime (a handful of boxes, the bench's per-keyframe call): fully unrolled evaluation,
//                    up to 128 registers, one CTA of up to 512 threads per SM, clusters of 16;
//   <false, 256, 3>  the call fills the machine a few times over (C1): same code held to 80 registers, three CTAs per SM;
//   <true,  256, 4>  saturated (C4): compact rolled loops (the unrolled evaluation is ~50 KB of SASS, more than the 32 KB
//                    L1.5 instruction cache; `no_instruction` was the second largest stall), 64 registers, four CTAs per SM.
template <bool ROLL, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
bf_refine_kernel(const bf_refine_params prm, int contrib_cap, int max_views, int pst_cap, int force_mode_b, int fit_dist, int timing) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned C = cluster.num_blocks();
    const unsigned crank = cluster.block_rank();
    const int G = gridDim.x / C;                       // clusters in flight
    const int cid = blockIdx.x / C;
    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const int P = prm.P;
    bf_refine_cfg cfg = prm.cfg;
    if (prm.intr_dev) {                                  // engine step: intrinsics of the current keyframe
        cfg.fx = prm.intr_dev[0]; cfg.fy = prm.intr_dev[1]; cfg.cx = prm.intr_dev[2]; cfg.cy = prm.intr_dev[3];
        cfg.img_w = prm.intr_dev[4]; cfg.img_h = prm.intr_dev[5];
    }
    const int B = prm.B_dev ? min(*prm.B_dev, prm.B) : prm.B;
    bf_refine_smem* sm = (bf_refine_smem*)bf_refine_smem_raw;
    bf_refine_state* S = &sm->S;
    bf_view* views = (bf_view*)(sm + 1);
    const int n_eval = min(32 * (cfg.pst_size / 32), P);
    int PB = (n_eval + (int)C - 1) / (int)C;
    PB = (PB + 31) & ~31;                                // a warp never straddles two views
    // fitness vector, double-buffered: a full copy [2][P] in every CTA (latency regime: the selection reads local shared
    // memory), or - fit_dist, throughput regime - only the CTA's own block [2][PB], read by the other CTAs through DSMEM
    // (shared memory is worth more as L1 for the polygon buffers there, and the selection is a vanishing share of an iteration)
    const int fit_len = fit_dist ? PB : P;
    float* fit = (float*)(views + max_views);
    int* cnt = (int*)(fit + 2 * fit_len);                  // per (round, warp) hit counts -> exclusive prefixes (+ total)
    float* terms = (float*)(cnt + BF_CNT_SLOTS);           // [8][max_hits] addends of cal_transform
    float* spst = terms + 8 * cfg.max_hits;                // particle template staged in shared memory when it fits (pst_cap = P)
    float* contrib = spst + 6 * pst_cap;                   // [V][PB] terms of this CTA's particle block (mode A)
    if (cid >= B) return;                                // cluster-uniform: nothing to do (before any barrier)
    const int p_lo = (int)crank * PB;
    const float beta = (float)cfg.beta, omb = (float)(1.0 - cfg.beta);
    const int nw = T >> 5, lane = tid & 31, warp = tid >> 5;
    const int rounds = (P + T - 1) / T;
    int overflow = 0;
    unsigned git = 0;                                    // iterations executed by this cluster (parity of the fit buffer)

    for (int k = tid; k < 6 * pst_cap; k += T) spst[k] = __ldg(prm.pst + k);
    if (tid == 0) { sm->dbg[0] = 0; sm->dbg[1] = 0; }

    for (int b = cid; b < B; b += G) {
        const int v0 = prm.view_offsets[b];
        const int V = prm.view_offsets[b + 1] - v0;
        if (V < 1 || V > max_views) {                    // cluster-uniform
            if (tid == 0 && crank == 0) { prm.out_updated[b] = 0; prm.out_iters[b] = 0; atomicExch(prm.status, BF_ERR_CAPACITY); }
            continue;
        }
        __syncthreads();                                 // previous box: everybody is done with views / state
        // ---- stage the views in every CTA: pose rows, observation hull (:367,375), area (:389)
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

        // mode A: one (view, particle) term per thread pass, column sums from CTA-local shared memory;
        // mode B: one thread per particle, views accumulated in a register (no storage, corners generated once)
        const bool mode_b = force_mode_b || (long long)PB * V > (long long)contrib_cap;
        const float denom = (float)V + 1e-6f;                 // count += 1 per view, then count + 1e-6 (:454)
        const float unlaunched = 0.0f / (0.0f + 1e-6f);       // particles beyond 32*int(pst_size/32)
        int it = 0;
        for (int n = 0; n < cfg.iters; ++n) {
            long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
            if (timing) tc0 = clock64();
            float* fitb = fit + (git & 1u) * fit_len;
            ++git;
            // fitness of particle j as the selection sees it
            auto fit_at = [&](int j) -> float {
                if (!fit_dist) return fitb[j];
                const int r = j / PB;
                return cluster.map_shared_rank(fitb, r)[j - r * PB];
            };
            // ---- evaluate_iou (:413-461) for this CTA's particle block ----
            if (!mode_b) {
                const int items = PB * V;
                for (int w = tid; w < items; w += T) {
                    const int v = w / PB, p = p_lo + (w - v * PB);            // view-major: a warp works on one view
                    if (p < n_eval) {
                        float pst6[6];
#pragma unroll
                        for (int k = 0; k < 6; ++k) pst6[k] = pst_cap ? spst[6 * p + k] : __ldg(prm.pst + 6 * p + k);
                        float c[8][3];
                        bf_particle_corners(S->box6, pst6, S->search, S->rot, c);
                        contrib[w] = bf_eval_view<ROLL>(c, views[v], cfg.fx, cfg.cx, cfg.fy, cfg.cy, cfg.img_w, cfg.img_h, &overflow, nullptr);
                    }
                }
                if (timing) {
                    tc1 = clock64();
                    if (lane == 0) { atomicMax(&sm->dbg[0], (int)(tc1 - tc0)); atomicAdd(&sm->dbg[1], (int)((tc1 - tc0) >> 6)); }
                }
                __syncthreads();
                // value = sum over views in ascending order (:400-401 with the host's grid order); fitness (:454) to every CTA
                for (int lp = tid; lp < PB; lp += T) {
                    const int p = p_lo + lp;
                    if (p < n_eval) {
                        const float f = bf_sum_views<4>(contrib + lp, PB, V) / denom;
                        if (fit_dist) fitb[lp] = f;
                        else for (unsigned r = 0; r < C; ++r) cluster.map_shared_rank(fitb, r)[p] = f;
                    }
                }
            } else {
                for (int lp = tid; lp < PB; lp += T) {
                    const int p = p_lo + lp;
                    if (p < n_eval) {
                        float pst6[6];
#pragma unroll
                        for (int k = 0; k < 6; ++k) pst6[k] = pst_cap ? spst[6 * p + k] : __ldg(prm.pst + 6 * p + k);
                        float c[8][3];
                        bf_particle_corners(S->box6, pst6, S->search, S->rot, c);
                        float value = 0.0f;
#pragma unroll 1
                        for (int v = 0; v < V; ++v)
                            value += bf_eval_view<ROLL>(c, views[v], cfg.fx, cfg.cx, cfg.fy, cfg.cy, cfg.img_w, cfg.img_h, &overflow, nullptr);
                        const float f = value / denom;
                        if (fit_dist) fitb[lp] = f;
                        else for (unsigned r = 0; r < C; ++r) cluster.map_shared_rank(fitb, r)[p] = f;
                    }
                }
                if (timing) {
                    tc1 = clock64();
                    if (lane == 0) { atomicMax(&sm->dbg[0], (int)(tc1 - tc0)); atomicAdd(&sm->dbg[1], (int)((tc1 - tc0) >> 6)); }
                }
            }
            ++it;
            if (C > 1) cluster.sync(); else __syncthreads();       // every block of fit[] has arrived in every CTA
            if (timing) tc2 = clock64();
            // ---- cal_transform's ordered selection (:475-535), replicated in every CTA: the first `max_hits` particles
            //      j >= 1 with fit[j] < fit[0] in index order.  Particle j = r*T + tid; its rank among the hits = hits of
            //      earlier (round, warp) groups + hits of lower lanes.  Pass 1: per-group ballot counts; warp 0 turns the
            //      counts into exclusive prefixes; pass 2: ranks and the eight addends of every selected particle.
            const float origin = (n_eval >= 1) ? fit_at(0) : unlaunched;
#pragma unroll 1
            for (int r = 0; r < rounds; ++r) {
                const int j = r * T + tid;
                const float f = (j < n_eval) ? fit_at(j) : unlaunched;
                const bool hit = (j >= 1 && j < P) && (f < origin);
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) cnt[r * nw + warp] = __popc(bal);
            }
            __syncthreads();
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
            const int hits = min(cnt[rounds * nw], cfg.max_hits);
#pragma unroll 1
            for (int r = 0; r < rounds; ++r) {
                if (cnt[r * nw] >= cfg.max_hits) break;             // block-uniform: every later rank is beyond the cap
                const int j = r * T + tid;
                const float f = (j < n_eval) ? fit_at(j) : unlaunched;
                const bool hit = (j >= 1 && j < P) && (f < origin);
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
            // ... accumulated sequentially in index order in float32, like the reference's Python loop (:490-515):
            // lanes 0..7 of warp 0 own one sum each and hand it to lane 0 by shuffle
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
                float acc8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc8[k] = __shfl_sync(0xffffffffu, acc, k);
                if (lane == 0) {
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
                    if (prm.trace && crank == 0) {
                        float* tr = prm.trace + ((size_t)b * cfg.iters + n) * 8;
                        tr[0] = (float)success; tr[1] = min_iou;
#pragma unroll
                        for (int k = 0; k < 6; ++k) tr[2 + k] = search[k];
                    }
                }
            }
            __syncthreads();                                 // the new state is visible to the whole CTA
            if (timing) tc3 = clock64();
            if (timing && prm.trace && crank == 0 && tid == 0) {   // diagnostic: cycles of {own evaluations, wait for the cluster, selection + update}
                float* tr = prm.trace + ((size_t)b * cfg.iters + n) * 8;
                tr[2] = (float)(tc1 - tc0); tr[3] = (float)(tc2 - tc1); tr[4] = (float)(tc3 - tc2); tr[5] = 0.f; tr[6] = 0.f;
                tr[7] = (float)sm->dbg[0];                                    // slowest warp's evaluation cycles in the leader CTA
                tr[1] = (float)sm->dbg[1] * 64.0f / (float)nw;                // mean warp evaluation cycles (overwrites min_iou in timing mode)
                sm->dbg[0] = 0; sm->dbg[1] = 0;
            }
            if (S->done) break;
        }
        if (tid == 0 && crank == 0) {
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
    if (overflow) atomicExch(prm.status, BF_ERR_CAPACITY);
    if (fit_dist && C > 1) cluster.sync();               // nobody reads this CTA's fitness block any more
}

#define BF_CONTRIB_CAP 8192    // (view, particle) terms a CTA holds in shared memory in mode A: 32 KB

static size_t bf_refine_smem_bytes(int fit_len, int max_hits, int contrib_cap, int max_views, int pst_cap) {
    return sizeof(bf_refine_smem) + sizeof(bf_view) * (size_t)max_views + sizeof(float) * 2 * (size_t)fit_len +
           sizeof(int) * (size_t)BF_CNT_SLOTS + sizeof(float) * 8 * (size_t)max_hits + sizeof(float) * (size_t)contrib_cap + sizeof(float) * 6 * (size_t)pst_cap + 16;
}

static int bf_particle_block(int n_eval, int C) {
    int PB = (n_eval + C - 1) / C;
    return (PB + 31) & ~31;
}

typedef void (*bf_refine_fn)(const bf_refine_params, int, int, int, int, int, int);

static int bf_refine_occupancy(bf_handle* h, bf_refine_fn kern, int variant, int ci, int ti, int C, int T, size_t smem, cudaStream_t st) {
    const int slot = ci * 4 + ti;
    const long long key = (long long)smem * 4 + variant + 1;
    if (h->refine_occ_smem[slot] == key) return h->refine_occ[slot];
    int active = 0;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(C * 64)); lc.blockDim = dim3(T); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&active, kern, &lc) != cudaSuccess) { cudaGetLastError(); active = 0; }
    h->refine_occ[slot] = active; h->refine_occ_smem[slot] = key;
    return active;
}

int bf_refine_run(bf_handle* h, const float* pst, int P, const float* per_xyzlhw, const float* per_R,
                  const float* per_scores, const float* per_uv, const float* per_poses,
                  const int32_t* view_offsets, const int32_t* view_index, int B, const bf_refine_cfg* cfg,
                  float* out_xyzlhw, int32_t* out_updated, int32_t* out_iters, float* trace, int32_t* status,
                  const bf_refine_dev_args* dev, cudaStream_t st) {
    const int max_views = (cfg->max_views > 0 && cfg->max_views <= BF_MAX_VIEWS) ? cfg->max_views : BF_MAX_VIEWS;
    bf_refine_params prm;
    prm.pst = pst; prm.P = P; prm.per_xyzlhw = per_xyzlhw; prm.per_R = per_R; prm.per_scores = per_scores;
    prm.per_uv = per_uv; prm.per_poses = per_poses; prm.view_offsets = view_offsets; prm.view_index = view_index;
    prm.B = B; prm.cfg = *cfg; prm.out_xyzlhw = out_xyzlhw; prm.out_updated = out_updated; prm.out_iters = out_iters;
    prm.trace = trace; prm.status = status;
    prm.B_dev = dev ? dev->B_dev : nullptr; prm.intr_dev = dev ? dev->intr_dev : nullptr;
    static const bf_refine_fn kernels[3] = {bf_refine_kernel<false, 512, 1>, bf_refine_kernel<false, 256, 3>, bf_refine_kernel<true, 256, 4>};
    static const int kernel_max_t[3] = {512, 256, 256};
    const int n_eval = (32 * (cfg->pst_size / 32) < P) ? 32 * (cfg->pst_size / 32) : P;
    // Launch shape.  A box's optimiser iteration costs passes = ceil(PB * V / T) evaluations per thread (PB = particle block
    // of one CTA) followed by the short selection; B boxes need waves = ceil(B / clusters that fit the machine).
    const long long avg_views = cfg->views_total > 0 ? (cfg->views_total + B - 1) / (B > 0 ? B : 1) : max_views;
    const long long items = (long long)n_eval * avg_views;              // evaluations of an average box per iteration
    static const int Cs[5] = {16, 8, 4, 2, 1};
    static const int Ts[4] = {512, 384, 256, 128};
    int bestC = 1, bestT = 256, variant = 0, clusters = B, mode_b = 0;
    double best_cost = 1e300;
    const bool persistent = (dev && dev->B_dev) || h->refine_force_persistent > 0;
    // throughput regimes (the call alone fills the machine): 256-thread CTAs, several per SM so that one box's selection
    // and cluster barrier overlap another box's evaluations
    const bool saturated = !persistent && (double)B * (double)items >= (double)h->sm_count * 512.0;
    if (saturated) {
        variant = ((double)B * (double)items >= (double)h->sm_count * 8192.0) ? 2 : 1;
        bestT = 256;
        const long long slots = (long long)h->sm_count * (variant == 2 ? 4 : 3);
        // saturated: about two waves of CTAs - boxes that stop early leave their slots to the second wave (measured on C4,
        // round 2: clusters of 2 / 4 / 8 / 16 -> 74 / 63 / 59 / 62 ms forced, 46 / 34 / 31 / 30 ms with early stop);
        // mid regime: fill the resident slots once (C1: clusters of 8 -> 0.64 ms, of 16 -> 0.75 ms)
        const long long fill = (variant == 2) ? 2 * slots : slots + slots / 8;
        while (bestC < 16 && (long long)B * bestC * 2 <= fill) bestC *= 2;
        best_cost = 0.0;
    }
    if (!saturated && !persistent && h->refine_concurrent) {       // BF_OPT_REFINE_CONCURRENT: leave room for the other streams' kernels
        variant = 1; bestT = 256; bestC = 16;
        while (bestC > 1 && (long long)bf_particle_block(n_eval, bestC / 2) * max_views <= bestT) bestC /= 2;
        best_cost = 0.0;
    }
    if (persistent) {
        // one fixed shape for the captured engine step.  Latency shape: 16 x 512 threads, one CTA per SM.  With
        // BF_OPT_REFINE_CONCURRENT (many engines share the GPU, bench.py c5): 16 x 256 threads of the 80-register
        // instantiation, three CTAs per SM, so that the idle phases of one sequence's clusters (cluster barrier, selection)
        // are filled by another sequence's evaluations.
        variant = h->refine_concurrent ? 1 : 0; bestT = h->refine_concurrent ? 256 : 512; best_cost = 0.0;
        // cluster size: 64 particles per CTA (8 views fill 512 threads in one pass); smaller templates leave room for more clusters
        bestC = 1;
        while (bestC < 16 && bestC * 64 < n_eval) bestC *= 2;
    }
    if (h->refine_force_c > 0 && h->refine_force_t > 0) {          // BF_REFINE_SHAPE (tuning sweeps)
        variant = (h->refine_force_variant >= 0 && h->refine_force_variant < 3) ? h->refine_force_variant : variant;
        if (h->refine_force_t <= kernel_max_t[variant]) { bestC = h->refine_force_c; bestT = h->refine_force_t; best_cost = 0.0; }
    }
    // one thread per particle with the views in a register once every thread has several evaluations anyway
    if (saturated) mode_b = ((long long)bf_particle_block(n_eval, bestC) * avg_views >= 4LL * bestT) ? 1 : 0;
    if (h->refine_force_mode >= 0) mode_b = h->refine_force_mode;
    const bf_refine_fn kern = kernels[variant];
    // the particle template itself in shared memory when it is small (24 KB at P = 1024) - latency regime only: it shortens
    // the selection, but in the throughput regimes the space is worth more as resident CTAs and L1
    const int pst_cap = (variant == 0 && P <= BF_PST_SMEM_MAX) ? P : 0;
    // terms of the CTA's particle block: sized for the smallest block any candidate shape uses (C = 16), capped
    auto contrib_for = [&](int C) {
        const long long need = mode_b ? 0 : (long long)bf_particle_block(n_eval, C) * max_views;
        return (int)(need <= BF_CONTRIB_CAP ? need : BF_CONTRIB_CAP);
    };
    int contrib_cap = contrib_for(best_cost > 0.0 ? 1 : bestC);
    // rounded up to 8 KB so that the cached occupancy answers below are reused across keyframes
    const int fit_dist = (saturated && mode_b && bestC > 1) ? 1 : 0;
    size_t smem = bf_refine_smem_bytes(fit_dist ? bf_particle_block(n_eval, bestC) : P, cfg->max_hits, contrib_cap, max_views, pst_cap);
    smem = (variant == 0) ? (smem + 8191) / 8192 * 8192 : (smem + 1023) / 1024 * 1024;
    BF_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BF_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // latency regime: pick the (cluster size C, block size T) that minimises waves * passes, preferring fewer threads on ties
    for (int ci = 0; ci < 5 && best_cost > 0.0; ++ci)
        for (int ti = 0; ti < 4; ++ti) {
            const int C = Cs[ci], T = Ts[ti];
            const int active = bf_refine_occupancy(h, kern, variant, ci, ti, C, T, smem, st);
            if (active < 1) continue;
            const long long waves = (B + active - 1) / active;
            const long long blk = (long long)bf_particle_block(n_eval, C) * max_views;    // the slowest box sets the latency
            const long long passes = (blk + T - 1) / T;
            const double cost = (double)waves * (double)passes + 1e-6 * C * T;      // ties -> fewer threads
            if (cost < best_cost) { best_cost = cost; bestC = C; bestT = T; }
        }
    if (persistent) {
        // as many clusters as the machine holds at once, every cluster loops over boxes cluster_id, cluster_id + G, ...
        int ci = 0;
        while (ci < 4 && Cs[ci] != bestC) ++ci;
        int active = bf_refine_occupancy(h, kern, variant, ci, h->refine_concurrent ? 2 : 0, bestC, bestT, smem, st);
        if (h->refine_concurrent && active > 8) active = 8;   // a keyframe refines a handful of boxes; the rest of the machine is other engines'
        if (h->refine_force_persistent > 0 && h->refine_force_persistent < active) active = h->refine_force_persistent;
        if (active < 1) return bf_fail(h, BF_ERR_CUDA, "bf_refine", "no resident cluster for the persistent launch shape");
        clusters = active;
        if (dev && dev->max_boxes > 0 && clusters > dev->max_boxes) clusters = dev->max_boxes;
        if (!(dev && dev->B_dev) && clusters > B) clusters = B;
    }
    {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)(clusters * bestC)); lc.blockDim = dim3(bestT); lc.dynamicSmemBytes = smem; lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = bestC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&lc, kern, prm, contrib_cap, max_views, pst_cap, mode_b, fit_dist, h->refine_timing);
        if (e != cudaSuccess) return bf_fail(h, BF_ERR_CUDA, "bf_refine_kernel", cudaGetErrorString(e));
        h->last_refine_cluster = variant * 1000000 + bestC * 1000 + bestT;
    }
    return BF_OK;
}

// status is OR-ed into (sticky), never cleared: the caller zeroes it.  Boxes whose view count is outside
// [1, max_views] are skipped and reported in status by the kernel itself.
extern "C" int bf_refine(bf_handle* h, const float* pst, int P, const float* per_xyzlhw, const float* per_R,
                         const float* per_scores, const float* per_uv, const float* per_poses, int M,
                         const int32_t* view_offsets, const int32_t* view_index, int B, const bf_refine_cfg* cfg,
                         float* out_xyzlhw, int32_t* out_updated, int32_t* out_iters, float* trace, int32_t* status,
                         void* stream) {
    bf_device_guard guard(h);
    if (!h || !cfg || B < 0 || P < 1 || P > BF_MAX_PARTICLES) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_refine", "bad size");
    if (!status) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_refine", "null status");
    if (B == 0) return BF_OK;
    if (!pst || !per_xyzlhw || !per_R || !per_scores || !per_uv || !per_poses || !view_offsets || !view_index ||
        !out_xyzlhw || !out_updated || !out_iters)
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_refine", "null pointer");
    if (cfg->max_hits < 1 || cfg->max_hits > 4096 || cfg->iters < 1) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_refine", "bad cfg");
    (void)M;
    return bf_refine_run(h, pst, P, per_xyzlhw, per_R, per_scores, per_uv, per_poses, view_offsets, view_index, B, cfg,
                         out_xyzlhw, out_updated, out_iters, trace, status, nullptr, (cudaStream_t)stream);
}

extern "C" int bf_refine_last_launch(bf_handle* h) { return h ? h->last_refine_cluster : 0; }

// Diagnostics: evaluations that were redone with plain divisions because an operand left the window of the branch-free ones
// (bf_refine_eval.cuh, bf_fdiv) since the last reset.  Synchronises the device.
extern "C" long long bf_debug_cold_redos(bf_handle* h, int reset) {
    bf_device_guard guard(h);
    if (!h) return -1;
    unsigned int v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(&v, bf_cold_redos, sizeof(v)) != cudaSuccess) return -1;
    if (reset) { const unsigned int z = 0; if (cudaMemcpyToSymbol(bf_cold_redos, &z, sizeof(z)) != cudaSuccess) return -1; }
    return (long long)v;
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
    bf_device_guard guard(h);
    if (!h || !cfg || P < 1 || V < 1 || V > BF_MAX_VIEWS || !pst || !box6 || !rot9 || !uv || !poses || !search6 || !fitness)
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_evaluate_iou", "bad argument");
    const int grid = bf_blocks(P, BF_REFINE_THREADS);
    bf_evaluate_kernel<<<grid, BF_REFINE_THREADS, 0, (cudaStream_t)stream>>>(pst, P, box6, rot9, uv, poses, V, search6, *cfg, fitness);
    BF_LAUNCH_CHECK(h, "bf_evaluate_kernel");
    return BF_OK;
}
