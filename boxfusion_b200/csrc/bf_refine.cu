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

struct P2 { float x, y; };

#define BF_CAND_MAX 48          // reference buffer: 36 (box_fusion.py:378); overflow is reported, not UB
#define BF_REFINE_THREADS 256

__device__ __forceinline__ float bf_cross(const P2 o, const P2 a, const P2 b) {          // :74-76
    return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x);
}

// :95-145  sort by (x,y), monotone chain popping on cross <= 0, output lower[:-1] + upper[:-1]
template <int NMAX>
__device__ __forceinline__ int bf_hull2d(P2* __restrict__ p, int n, P2* __restrict__ out) {
    if (n == 0) return 0;
    for (int i = 1; i < n; ++i) {
        const P2 k = p[i];
        int j = i - 1;
        while (j >= 0 && (p[j].x > k.x || (p[j].x == k.x && p[j].y > k.y))) { p[j + 1] = p[j]; --j; }
        p[j + 1] = k;
    }
    P2 up[NMAX];
    int nl = 0, nu = 0;
    for (int i = 0; i < n; ++i) {                       // lower chain is built directly in `out`
        while (nl >= 2 && bf_cross(out[nl - 2], out[nl - 1], p[i]) <= 0) --nl;
        out[nl++] = p[i];
    }
    for (int i = n - 1; i >= 0; --i) {
        while (nu >= 2 && bf_cross(up[nu - 2], up[nu - 1], p[i]) <= 0) --nu;
        up[nu++] = p[i];
    }
    --nl; --nu;
    for (int i = 0; i < nu; ++i) out[nl + i] = up[i];
    return nl + nu;
}

__device__ __forceinline__ float bf_shoelace(const P2* __restrict__ q, int n) {          // :148-156
    float a = 0.0f;
    for (int i = 0; i < n; ++i) {
        const P2 p1 = q[i], p2 = q[(i + 1 == n) ? 0 : i + 1];
        a += p1.x * p2.y - p2.x * p1.y;
    }
    return fabsf(a) * 0.5f;                              // fabs(area)/2.0 is exact either way
}

__device__ __forceinline__ bool bf_seg_intersect(const P2 a1, const P2 a2, const P2 b1, const P2 b2, P2* out) {  // :159-177
    const double dx1 = a2.x - a1.x, dy1 = a2.y - a1.y;
    const double dx2 = b2.x - b1.x, dy2 = b2.y - b1.y;
    const double den = dx1 * dy2 - dy1 * dx2;
    if (fabs(den) < 1e-8) return false;
    const double e1 = a1.y - b1.y, e2 = b1.x - a1.x;     // float differences widened to double
    const double t = (dx2 * e1 + dy2 * e2) / den;
    const double s = (dx1 * e1 + dy1 * e2) / den;
    if (t >= -1e-8 && t <= 1.00000001 && s >= -1e-8 && s <= 1.00000001) {
        out->x = (float)(a1.x + t * dx1);
        out->y = (float)(a1.y + t * dy1);
        return true;
    }
    return false;
}

__device__ __forceinline__ bool bf_inside_poly(const P2 p, const P2* __restrict__ q, int n) {   // :180-199
    bool in = false;
    for (int i = 0; i < n; ++i) {
        const P2 p1 = q[i], p2 = q[(i + 1 == n) ? 0 : i + 1];
        if ((p1.y > p.y) != (p2.y > p.y)) {
            const float xi = ((p.y - p1.y) * (p2.x - p1.x) / (p2.y - p1.y)) + p1.x;
            if (p.x < xi) in = !in;
        }
    }
    return in;
}

// IoU of the particle's projected hull h0 against the view's observation hull ht (:380-398).
__device__ __forceinline__ float bf_hull_iou(const P2* __restrict__ h0, int n0, const P2* __restrict__ ht, int nt,
                                             float area_t, int* overflow) {
    P2 cand[BF_CAND_MAX], hi[BF_CAND_MAX];
    int nc = 0;
    for (int i = 0; i < n0; ++i) if (bf_inside_poly(h0[i], ht, nt)) { if (nc < BF_CAND_MAX) cand[nc] = h0[i]; ++nc; }
    for (int i = 0; i < nt; ++i) if (bf_inside_poly(ht[i], h0, n0)) { if (nc < BF_CAND_MAX) cand[nc] = ht[i]; ++nc; }
    for (int i = 0; i < n0; ++i) {
        const P2 a1 = h0[i], a2 = h0[(i + 1 == n0) ? 0 : i + 1];
        for (int j = 0; j < nt; ++j) {
            P2 x;
            if (bf_seg_intersect(a1, a2, ht[j], ht[(j + 1 == nt) ? 0 : j + 1], &x)) { if (nc < BF_CAND_MAX) cand[nc] = x; ++nc; }
        }
    }
    if (nc > BF_CAND_MAX) { *overflow = 1; nc = BF_CAND_MAX; }
    const int ni = bf_hull2d<BF_CAND_MAX>(cand, nc, hi);
    const float ai = bf_shoelace(hi, ni), a0 = bf_shoelace(h0, n0);
    const float uni = a0 + area_t - ai;
    float iou = 0;
    if (uni > 0) iou = (float)((double)ai / ((double)uni + 0.00001));
    return iou;
}

struct bf_view {            // per-view constants staged in shared memory
    float pose[12];         // rows 0..2 of the camera->world 4x4
    P2 hull[8];
    int nt;
    float area_t;
};

// One (particle, view) term |1 - iou| from the particle's world corners (:345-400).
__device__ __forceinline__ float bf_eval_view(const float (*c)[3], const bf_view& vw, float fx, float cx, float fy,
                                              float cy, float img_w, float img_h, int* overflow) {
    P2 uv[8], h0[8];
    const float* ps = vw.pose;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float vx = c[j][0] - ps[3], vy = c[j][1] - ps[7], vz = c[j][2] - ps[11];
        const float camx = ps[0] * vx + ps[4] * vy + ps[8] * vz;
        const float camy = ps[1] * vx + ps[5] * vy + ps[9] * vz;
        const float camz = ps[2] * vx + ps[6] * vy + ps[10] * vz;
        const float px = ((camx * fx) / camz + cx);
        const float py = ((camy * fy) / camz + cy);
        uv[j].x = (px > img_w) ? img_w : (px < 0) ? 0 : px;
        uv[j].y = (py > img_h) ? img_h : (py < 0) ? 0 : py;
    }
    const int n0 = bf_hull2d<8>(uv, 8, h0);
    const float iou = bf_hull_iou(h0, n0, vw.hull, vw.nt, vw.area_t, overflow);
    return fabsf(1 - iou);
}

// Particle -> 8 world corners (:289-331).
__device__ __forceinline__ void bf_particle_corners(const float* box6, const float* pst6, const float* search,
                                                    const float* rot, float (*c)[3]) {
    float x3d = box6[0], y3d = box6[1], z3d = box6[2];
    float w3d = box6[5], h3d = box6[4], l3d = box6[3];
    x3d = x3d + pst6[0] * search[0];
    y3d = y3d + pst6[1] * search[1];
    z3d = z3d + pst6[2] * search[2];
    w3d = w3d + pst6[5] * search[5];
    h3d = h3d + pst6[4] * search[4];
    l3d = l3d + pst6[3] * search[3];
    const float xyz[3] = {x3d, y3d, z3d};
    w3d = fmaxf(w3d, 0.01f); h3d = fmaxf(h3d, 0.01f); l3d = fmaxf(l3d, 0.01f);
    const float hl = l3d / 2, hh = h3d / 2, hw = w3d / 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float vx = ((i & 1) ^ ((i >> 1) & 1)) ? hl : -hl;
        const float vy = (i & 2) ? hh : -hh;
        const float vz = (i & 4) ? hw : -hw;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
            acc += rot[j * 3 + 0] * vx;
            acc += rot[j * 3 + 1] * vy;
            acc += rot[j * 3 + 2] * vz;
            acc += xyz[j];
            c[i][j] = acc;
        }
    }
}

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
struct bf_refine_smem {
    bf_refine_state S;
    bf_view views[BF_MAX_VIEWS];
    int warp_cnt[32];
    float vbox[6 * BF_MAX_VIEWS];        // gathered view boxes (init_opt_params)
    float vscore[BF_MAX_VIEWS];
    float col[3 * BF_MAX_VIEWS];
    int overflow;
    // followed by: float fit[P]; int sel[max_hits]; float contrib[pair_cap]
};

extern __shared__ __align__(16) unsigned char bf_refine_smem_raw[];

// One thread-block CLUSTER per map box.  Work items of an optimiser iteration are spread over all CTAs of
// the cluster; every CTA writes its results straight into the leader's shared memory (DSMEM), the leader
// reduces in the reference's order and publishes the new state, two cluster barriers per iteration.
//   pair mode     (n_eval*V <= pair_cap): one work item = one (particle, view); contributions are stored
//                 view-major and summed per particle in ascending view order by the leader
//   particle mode (larger problems): one work item = one particle, views summed sequentially in registers
__global__ void __launch_bounds__(BF_REFINE_THREADS, 2)
bf_refine_kernel(const bf_refine_params prm, int pair_cap) {
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
    bf_view* views = sm->views;
    float* fit = (float*)(sm + 1);
    int* sel = (int*)(fit + prm.P);
    float* contrib = (float*)(sel + cfg.max_hits);
    if (tid == 0) sm->overflow = 0;
    if (V < 1 || V > BF_MAX_VIEWS) {                     // flagged in status by bf_check_views_kernel2 (cluster-uniform)
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
        bf_view& vw = views[v];
#pragma unroll
        for (int k = 0; k < 12; ++k) vw.pose[k] = prm.per_poses[16 * (size_t)m + k];
        P2 t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { t[k].x = prm.per_uv[16 * (size_t)m + 2 * k]; t[k].y = prm.per_uv[16 * (size_t)m + 2 * k + 1]; }
        P2 ht[8];
        vw.nt = bf_hull2d<8>(t, 8, ht);
        for (int k = 0; k < 8; ++k) vw.hull[k] = ht[k < vw.nt ? k : 0];
        vw.area_t = bf_shoelace(ht, vw.nt);
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
    const bool pair_mode = (long long)n_eval * V <= (long long)pair_cap;
    const float beta = (float)cfg.beta, omb = (float)(1.0 - cfg.beta);
    int overflow = 0;
    int it = 0;
    cluster.sync();                                      // every CTA's shared memory is initialised
    for (int n = 0; n < cfg.iters; ++n) {
        // ---- evaluate_iou (:413-461) spread over the cluster ------------------------------------------------
        if (pair_mode) {
            const int items = n_eval * V;
            for (int w = crank * T + tid; w < items; w += C * T) {
                const int v = w / n_eval, p = w - v * n_eval;          // view-major: a warp works on one view
                float pst6[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) pst6[k] = __ldg(prm.pst + 6 * (size_t)p + k);
                float c[8][3];
                bf_particle_corners(S->box6, pst6, S->search, S->rot, c);
                l_contrib[w] = bf_eval_view(c, views[v], cfg.fx, cfg.cx, cfg.fy, cfg.cy, cfg.img_w, cfg.img_h, &overflow);
            }
        } else {
            for (int p = crank * T + tid; p < n_eval; p += C * T) {
                float pst6[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) pst6[k] = __ldg(prm.pst + 6 * (size_t)p + k);
                float c[8][3];
                bf_particle_corners(S->box6, pst6, S->search, S->rot, c);
                float value = 0.0f, count = 0.0f;
                for (int v = 0; v < V; ++v) {           // ascending views: the host order of the atomicAdd sum (:400)
                    value += bf_eval_view(c, views[v], cfg.fx, cfg.cx, cfg.fy, cfg.cy, cfg.img_w, cfg.img_h, &overflow);
                    count += 1;
                }
                l_fit[p] = value / (count + 1e-6f);     // :454
            }
        }
        ++it;
        cluster.sync();
        if (crank == 0) {
            // ---- fitness per particle, views summed in ascending order (:400-401, :454) ---------------------
            if (pair_mode) {
                for (int p = tid; p < n_eval; p += T) {
                    float value = 0.0f, count = 0.0f;
                    for (int v = 0; v < V; ++v) { value += contrib[v * n_eval + p]; count += 1; }
                    fit[p] = value / (count + 1e-6f);
                }
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
            // sequential float32 accumulation in index order: 6 PST sums, weight sum, weighted-fitness sum
            if (tid < 8) {
                float acc = 0.0f;
                for (int q = 0; q < hits; ++q) {
                    const int j = sel[q];
                    const float w = origin - fit[j];
                    const float term = (tid < 6) ? __ldg(prm.pst + 6 * (size_t)j + tid) * w : ((tid == 6) ? w : fit[j] * w);
                    acc += term;
                }
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
        cluster.sync();                                  // leader state published
        if (crank != 0) {
            if (tid < 6) { S->box6[tid] = l_S->box6[tid]; S->search[tid] = l_S->search[tid]; }
            if (tid == 6) S->done = l_S->done;
        }
        __syncthreads();
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

#define BF_PAIR_CAP 16384      // (particle, view) contributions the leader can hold: 64 KB

static size_t bf_refine_smem_bytes(int P, int max_hits, int pair_cap) {
    return sizeof(bf_refine_smem) + sizeof(float) * (size_t)P + sizeof(int) * (size_t)max_hits + sizeof(float) * (size_t)pair_cap + 16;
}

__global__ void bf_check_views_kernel(const int32_t* __restrict__ off, int B, int32_t* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) status[0] = 0;
}
__global__ void bf_check_views_kernel2(const int32_t* __restrict__ off, int B, int32_t* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { const int V = off[b + 1] - off[b]; if (V < 1 || V > BF_MAX_VIEWS) atomicExch(status, BF_ERR_CAPACITY); }
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
    bf_check_views_kernel2<<<bf_blocks(B, 128), 128, 0, st>>>(view_offsets, B, status);
    bf_refine_params prm;
    prm.pst = pst; prm.P = P; prm.per_xyzlhw = per_xyzlhw; prm.per_R = per_R; prm.per_scores = per_scores;
    prm.per_uv = per_uv; prm.per_poses = per_poses; prm.view_offsets = view_offsets; prm.view_index = view_index;
    prm.B = B; prm.cfg = *cfg; prm.out_xyzlhw = out_xyzlhw; prm.out_updated = out_updated; prm.out_iters = out_iters;
    prm.trace = trace; prm.status = status;
    const int pair_cap = BF_PAIR_CAP;
    const size_t smem = bf_refine_smem_bytes(P, cfg->max_hits, pair_cap);
    BF_CUDA(h, cudaFuncSetAttribute(bf_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BF_CUDA(h, cudaFuncSetAttribute(bf_refine_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // cluster size: enough CTAs that a box's work items are ~2 per thread, as long as the whole launch still
    // fits the machine about twice over (2 CTAs of 256 threads per SM)
    const int n_eval = (32 * (cfg->pst_size / 32) < P) ? 32 * (cfg->pst_size / 32) : P;
    const long long hint_items = (long long)n_eval * (cfg->views_hint > 0 ? cfg->views_hint : 8);
    int C = 1;
    while (C < 16 && (long long)C * BF_REFINE_THREADS * 2 < hint_items) C *= 2;
    while (C > 1 && (long long)B * C > 4LL * h->sm_count) C /= 2;
    for (;; C /= 2) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)(B * C)); lc.blockDim = dim3(BF_REFINE_THREADS); lc.dynamicSmemBytes = smem; lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        int nclusters = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, bf_refine_kernel, &lc);
        if ((e != cudaSuccess || nclusters < 1) && C > 1) { cudaGetLastError(); continue; }
        e = cudaLaunchKernelEx(&lc, bf_refine_kernel, prm, pair_cap);
        if (e != cudaSuccess) {
            if (C > 1) { cudaGetLastError(); continue; }
            return bf_fail(h, BF_ERR_CUDA, "bf_refine_kernel", cudaGetErrorString(e));
        }
        h->last_refine_cluster = C;
        break;
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
        bf_view& vw = views[v];
        for (int k = 0; k < 12; ++k) vw.pose[k] = poses[16 * v + k];
        P2 t[8], ht[8];
        for (int k = 0; k < 8; ++k) { t[k].x = uv[16 * v + 2 * k]; t[k].y = uv[16 * v + 2 * k + 1]; }
        vw.nt = bf_hull2d<8>(t, 8, ht);
        for (int k = 0; k < 8; ++k) vw.hull[k] = ht[k < vw.nt ? k : 0];
        vw.area_t = bf_shoelace(ht, vw.nt);
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
                value += bf_eval_view(c, views[v], cfg.fx, cfg.cx, cfg.fy, cfg.cy, cfg.img_w, cfg.img_h, &overflow);
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
