/* TEST INFRASTRUCTURE ONLY - CPU restatement (plain C) of the reference's particle refinement.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product path (boxfusion_b200/) never does.
 *
 * What it restates (all paths relative to /root/reference):
 *   - CUDA kernel `compute_iou_value` and its helpers         boxfusion/box_fusion.py:68-405
 *   - BoxFusion.evaluate_iou  (fitness = value/(count+1e-6))  boxfusion/box_fusion.py:413-461
 *   - BoxFusion.cal_transform (ordered first-200 rule)        boxfusion/box_fusion.py:475-535
 *   - BoxFusion.update_PST                                     boxfusion/box_fusion.py:537-562
 *   - BoxFusion.init_opt_params                                boxfusion/box_fusion.py:566-600
 *   - the per-box optimiser loop of BoxFusion.boxfusion        boxfusion/box_fusion.py:651-721
 *
 * Pinning: the reference has no tests or golden vectors for this path.  This file
 * is pinned against the reference itself executed in the build container: (i) the kernel string
 * compiled verbatim for the host (oracle/_ref/libref_kernel.so, recipe oracle/ref_harness.py) and
 * (ii) the reference's own Python optimiser loop run over that kernel; tests/test_oracle_golden.py
 * and tests/golden/make_golden.py hold the comparisons (bit-exact on every compared float).
 *
 * Numeric model: float32 exactly where the reference uses float, double where it uses double
 * (line_intersection, the `+0.00001` literal, NumPy-2 promotion rules for the host loop);
 * build with -ffp-contract=off so no FMA is formed (the host-compiled reference has none either).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* threads used by the OpenMP loops of this library (1 = the single-threaded restatement; bench.py's faithful CPU arm) */
void bfo_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

typedef struct { float x, y; } pt2;

#define BFO_MAX_CAND 96
#define BFO_MAX_HULL 96

/* diagnostics the parity tests assert on (fixed buffers in the reference) */
static int g_max_cand = 0, g_max_inter_hull = 0;
void bfo_reset_stats(void) { g_max_cand = 0; g_max_inter_hull = 0; }
void bfo_get_stats(int* max_cand, int* max_inter_hull) { *max_cand = g_max_cand; *max_inter_hull = g_max_inter_hull; }

/* box_fusion.py:74-76 */
static float cross3(pt2 o, pt2 a, pt2 b) {
    return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x);
}

/* box_fusion.py:95-145: sort by (x, y), Andrew monotone chain popping on cross <= 0,
 * output = lower chain (without its last point) followed by upper chain (without its last). */
static int hull2d(pt2* p, int n, pt2* out) {
    if (n == 0) return 0;
    for (int i = 1; i < n; ++i) {              /* insertion sort: same total order as :103-112 */
        pt2 k = p[i];
        int j = i - 1;
        while (j >= 0 && (p[j].x > k.x || (p[j].x == k.x && p[j].y > k.y))) { p[j + 1] = p[j]; --j; }
        p[j + 1] = k;
    }
    pt2 lo[BFO_MAX_HULL], up[BFO_MAX_HULL];
    int nl = 0, nu = 0;
    for (int i = 0; i < n; ++i) {
        while (nl >= 2 && cross3(lo[nl - 2], lo[nl - 1], p[i]) <= 0) --nl;
        lo[nl++] = p[i];
    }
    for (int i = n - 1; i >= 0; --i) {
        while (nu >= 2 && cross3(up[nu - 2], up[nu - 1], p[i]) <= 0) --nu;
        up[nu++] = p[i];
    }
    --nl; --nu;
    for (int i = 0; i < nl; ++i) out[i] = lo[i];
    for (int i = 0; i < nu; ++i) out[nl + i] = up[i];
    return nl + nu;
}

/* box_fusion.py:148-156 */
static float shoelace(const pt2* q, int n) {
    float a = 0.0f;
    for (int i = 0; i < n; ++i) {
        const pt2 p1 = q[i], p2 = q[(i + 1) % n];
        a += p1.x * p2.y - p2.x * p1.y;
    }
    return (float)(fabs((double)a) / 2.0);
}

/* box_fusion.py:159-177 (double arithmetic on float differences) */
static int seg_intersect(pt2 a1, pt2 a2, pt2 b1, pt2 b2, pt2* out) {
    double dx1 = a2.x - a1.x, dy1 = a2.y - a1.y;
    double dx2 = b2.x - b1.x, dy2 = b2.y - b1.y;
    double den = dx1 * dy2 - dy1 * dx2;
    if (fabs(den) < 1e-8) return 0;
    double t = (dx2 * (a1.y - b1.y) + dy2 * (b1.x - a1.x)) / den;
    double s = (dx1 * (a1.y - b1.y) + dy1 * (b1.x - a1.x)) / den;
    if (t >= -1e-8 && t <= 1.00000001 && s >= -1e-8 && s <= 1.00000001) {
        out->x = (float)(a1.x + t * dx1);
        out->y = (float)(a1.y + t * dy1);
        return 1;
    }
    return 0;
}

/* box_fusion.py:180-199 (even-odd ray cast, float) */
static int inside_poly(pt2 p, const pt2* q, int n) {
    int in = 0;
    for (int i = 0; i < n; ++i) {
        const pt2 p1 = q[i], p2 = q[(i + 1) % n];
        if ((p1.y > p.y) != (p2.y > p.y)) {
            const float xi = ((p.y - p1.y) * (p2.x - p1.x) / (p2.y - p1.y)) + p1.x;
            if (p.x < xi) in = !in;
        }
    }
    return in;
}

/* box_fusion.py:202-236; the atan2 ordering at :239-260 is dropped because the caller re-hulls
 * the candidates (:384) and hull2d sorts them itself - the result does not depend on input order. */
static int poly_intersection(const pt2* a, int na, const pt2* b, int nb, pt2* cand) {
    int nc = 0;
    for (int i = 0; i < na; ++i) if (inside_poly(a[i], b, nb)) cand[nc++] = a[i];
    for (int i = 0; i < nb; ++i) if (inside_poly(b[i], a, na)) cand[nc++] = b[i];
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) {
            pt2 x;
            if (seg_intersect(a[i], a[(i + 1) % na], b[j], b[(j + 1) % nb], &x)) cand[nc++] = x;
        }
    return nc;
}

/* Hull of the 8 observed (target) corners of one view, box_fusion.py:367,375. */
int bfo_target_hull(const float* t_c16, float* hull_xy /*[16]*/) {
    pt2 p[8], h[BFO_MAX_HULL];
    for (int k = 0; k < 8; ++k) { p[k].x = t_c16[2 * k]; p[k].y = t_c16[2 * k + 1]; }
    int n = hull2d(p, 8, h);
    for (int k = 0; k < n; ++k) { hull_xy[2 * k] = h[k].x; hull_xy[2 * k + 1] = h[k].y; }
    return n;
}

static float iou_points(pt2* uv, pt2* tgt);

/* One (particle, view) evaluation: box_fusion.py:289-398.  Returns iou (float). */
float bfo_eval_particle_view(const float* box6, const float* t_c16, const float* pst6, const float* rot9,
                             const float* pose16, float fx, float cx, float fy, float cy,
                             const float* search6, float img_h, float img_w) {
    float x3d = box6[0], y3d = box6[1], z3d = box6[2];
    float w3d = box6[5], h3d = box6[4], l3d = box6[3];
    x3d = x3d + pst6[0] * search6[0];
    y3d = y3d + pst6[1] * search6[1];
    z3d = z3d + pst6[2] * search6[2];
    w3d = w3d + pst6[5] * search6[5];
    h3d = h3d + pst6[4] * search6[4];
    l3d = l3d + pst6[3] * search6[3];
    const float xyz[3] = {x3d, y3d, z3d};
    w3d = fmaxf(w3d, 0.01f); h3d = fmaxf(h3d, 0.01f); l3d = fmaxf(l3d, 0.01f);
    const float hl = l3d / 2, hh = h3d / 2, hw = w3d / 2;
    const float verts[8][3] = {{-hl, -hh, -hw}, {hl, -hh, -hw}, {hl, hh, -hw}, {-hl, hh, -hw},
                               {-hl, -hh, hw},  {hl, -hh, hw},  {hl, hh, hw},  {-hl, hh, hw}};
    pt2 uv[8];
    for (int i = 0; i < 8; ++i) {
        float c[3];
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 3; ++k) acc += rot9[j * 3 + k] * verts[i][k];
            acc += xyz[j];
            c[j] = acc;
        }
        const float vx = c[0] - pose16[3], vy = c[1] - pose16[7], vz = c[2] - pose16[11];
        const float camx = pose16[0] * vx + pose16[4] * vy + pose16[8] * vz;
        const float camy = pose16[1] * vx + pose16[5] * vy + pose16[9] * vz;
        const float camz = pose16[2] * vx + pose16[6] * vy + pose16[10] * vz;
        const float px = ((camx * fx) / camz + cx);
        const float py = ((camy * fy) / camz + cy);
        uv[i].x = (px > img_w) ? img_w : (px < 0) ? 0 : px;
        uv[i].y = (py > img_h) ? img_h : (py < 0) ? 0 : py;
    }
    pt2 tgt[8];
    for (int k = 0; k < 8; ++k) { tgt[k].x = t_c16[2 * k]; tgt[k].y = t_c16[2 * k + 1]; }
    return iou_points(uv, tgt);
}

/* box_fusion.py:380-398 on two raw 8-point sets (exported for the degenerate-polygon tests of the kernel's
 * evaluation core: points snapped to a grid, clamped onto the image border, duplicated). */
float bfo_iou_points(const float* a16, const float* b16) {
    pt2 a[8], b[8];
    for (int k = 0; k < 8; ++k) { a[k].x = a16[2 * k]; a[k].y = a16[2 * k + 1]; b[k].x = b16[2 * k]; b[k].y = b16[2 * k + 1]; }
    return iou_points(a, b);
}

static float iou_points(pt2* uv, pt2* tgt) {
    pt2 h0[BFO_MAX_HULL], ht[BFO_MAX_HULL], cand[BFO_MAX_CAND], hi[BFO_MAX_HULL];
    const int n0 = hull2d(uv, 8, h0);
    const int nt = hull2d(tgt, 8, ht);
    const int nc = poly_intersection(h0, n0, ht, nt, cand);
    const int ni = hull2d(cand, nc, hi);
    if (nc > g_max_cand) g_max_cand = nc;              /* diagnostics only; benign race under OpenMP */
    if (ni > g_max_inter_hull) g_max_inter_hull = ni;
    const float ai = shoelace(hi, ni), a0 = shoelace(h0, n0), at = shoelace(ht, nt);
    const float uni = a0 + at - ai;
    float iou = 0;
    if (uni > 0) iou = (float)((double)ai / ((double)uni + 0.00001));
    return iou;
}

/* evaluate_iou: box_fusion.py:413-461 + kernel accumulation :400-401 with the host grid order
 * (views ascending).  n_eval = 32*int(pst_size/32) particles are evaluated (launch shape :450-451,
 * SURVEY H5); rows n_eval..P-1 keep value=count=0 -> fitness 0. */
void bfo_evaluate(const float* box6, const float* t_c /*[V,16]*/, const float* pst /*[P,6]*/, int P, int n_eval,
                  const float* rot9, const float* poses /*[V,16]*/, int V, const float* K16,
                  const float* search6, float img_h, float img_w, float* fitness /*[P]*/) {
    const float fx = K16[0], cx = K16[2], fy = K16[5], cy = K16[6];
    /* particles are independent: OpenMP only changes who computes which fitness value (bench.py's multi-core C baseline;
     * OMP_NUM_THREADS=1 is the single-threaded restatement used for pinning) */
#pragma omp parallel for schedule(static) if (P >= 256)
    for (int p = 0; p < P; ++p) {
        float value = 0.0f, count = 0.0f;
        if (p < n_eval)
            for (int v = 0; v < V; ++v) {
                const float iou = bfo_eval_particle_view(box6, t_c + 16 * v, pst + 6 * p, rot9, poses + 16 * v,
                                                         fx, cx, fy, cy, search6, img_h, img_w);
                value += fabsf(1 - iou);
                count += 1;
            }
        fitness[p] = value / (count + 1e-6f);
    }
}

/* cal_transform: box_fusion.py:475-535 under NumPy-2 promotion (float32 sequential sums). */
int bfo_cal_transform(const float* fitness, const float* pst, int P, const float* search6,
                      float* min_iou, float* mean_transform6) {
    float s[6] = {0, 0, 0, 0, 0, 0}, sw = 0.0f, si = 0.0f;
    const float origin = fitness[0];
    int hits = 0;
    for (int k = 0; k < 6; ++k) mean_transform6[k] = 0.0f;
    for (int j = 1; j < P; ++j) {
        if (fitness[j] < origin) {
            const float w = origin - fitness[j];
            for (int k = 0; k < 6; ++k) s[k] += pst[6 * j + k] * w;
            sw += w;
            si += fitness[j] * w;
            if (++hits == 200) break;
        }
    }
    if (hits <= 0) { *min_iou = origin; return 0; }
    *min_iou = si / sw;
    for (int k = 0; k < 6; ++k) mean_transform6[k] = (s[k] / sw) * search6[k];
    return 1;
}

/* update_PST: box_fusion.py:537-562 (float32 throughout). */
void bfo_update_pst(float iou, const float* mt6, float center_scale, float shape_scale, float* search6) {
    const float ms = 1e-3f;
    float s[6];
    for (int k = 0; k < 6; ++k) s[k] = fabsf(mt6[k]) + ms;
    float n2 = s[0] * s[0];
    for (int k = 1; k < 6; ++k) n2 = n2 + s[k] * s[k];
    const float nrm = sqrtf(n2);
    for (int k = 3; k < 6; ++k) search6[k] = shape_scale * iou * (s[k] / nrm) + ms;
    for (int k = 0; k < 3; ++k) search6[k] = center_scale * iou * (s[k] / nrm) + ms;
}

/* numpy/_core/src/umath/loops_utils.h.src pairwise_sum (float32), as used by np.add.reduce on a
 * contiguous axis. */
static float np_pairwise_sum(const float* a, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        int i;
        for (i = 0; i < 8; ++i) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

/* init_opt_params: box_fusion.py:566-600. boxes [V,6] f32, scores [V] -> mean_xyzlwh (double[6]), best view. */
int bfo_init_opt_params(const float* boxes, const float* scores, int V, double* mean6) {
    int best = 0;
    for (int v = 1; v < V; ++v) if (scores[v] > scores[best]) best = v;      /* np.argmax: first max */
    for (int k = 0; k < 3; ++k) {                                             /* np.mean axis=0, f32 */
        float acc = boxes[k];
        for (int v = 1; v < V; ++v) acc += boxes[6 * v + k];
        mean6[k] = (double)(acc / (float)V);
    }
    /* rank of each best-view dim (argsort, then position of k in the sorted order) */
    int rank[3];
    const float* bd = boxes + 6 * best + 3;
    int order[3] = {0, 1, 2};
    for (int i = 1; i < 3; ++i) {                                              /* stable insertion sort */
        int k = order[i], j = i - 1;
        while (j >= 0 && bd[order[j]] > bd[k]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = k;
    }
    for (int i = 0; i < 3; ++i) rank[order[i]] = i;
    /* B_sorted[:, get_indices] is F-ordered, so np.mean(axis=0) reduces along contiguous memory and
     * NumPy's pairwise summation applies (8 accumulators for n >= 8, recursion above 128). */
    float* col = (float*)malloc(sizeof(float) * 3 * (size_t)V);
    for (int v = 0; v < V; ++v) {
        float d[3] = {boxes[6 * v + 3], boxes[6 * v + 4], boxes[6 * v + 5]};
        for (int i = 1; i < 3; ++i) { float k = d[i]; int j = i - 1; while (j >= 0 && d[j] > k) { d[j + 1] = d[j]; --j; } d[j + 1] = k; }
        for (int k = 0; k < 3; ++k) col[k * V + v] = d[rank[k]];
    }
    for (int k = 0; k < 3; ++k) mean6[3 + k] = (double)(np_pairwise_sum(col + k * V, V) / (float)V);
    free(col);
    return best;
}

typedef struct {
    int iters;                 /* cfg box_fusion.iters (20)                         */
    int pst_size;              /* cfg box_fusion.pst_size -> n_eval = 32*(pst_size/32) */
    float center_init, shape_init, center_scale, shape_scale;
    double beta;               /* 0.9 (Python float), box_fusion.py:622              */
    float img_h, img_w;
    int early_stop;            /* 1 = reference behaviour (break after 3 failures)   */
} bfo_cfg;

/* Per-box optimiser loop: box_fusion.py:651-721.  Returns need_update; out6 = float32-rounded
 * fused (x,y,z,l,h,w) as written back at :721; n_iters = evaluate_iou calls made.
 * trace (optional, [iters*8]): per iteration {success, min_iou, search_size[6] after update}. */
int bfo_refine_box(const float* view_boxes /*[V,6]*/, const float* view_R /*[V,9]*/, const float* view_scores,
                   const float* t_c /*[V,16]*/, const float* poses /*[V,16]*/, int V,
                   const float* pst, int P, const float* K16, const bfo_cfg* cfg,
                   float* out6, int* n_iters, float* trace) {
    double g[6];
    const int best = bfo_init_opt_params(view_boxes, view_scores, V, g);
    const float* rot = view_R + 9 * best;
    float search[6], prev[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 3; ++k) { search[k] = cfg->center_init; search[3 + k] = cfg->shape_init; }
    int need_update = 0, previous_success = 0, fail = 0, it = 0;
    const int n_eval = 32 * (cfg->pst_size / 32) < P ? 32 * (cfg->pst_size / 32) : P;
    float* fitness = (float*)malloc(sizeof(float) * (size_t)P);
    const float omb = (float)(1.0 - cfg->beta);     /* (1-beta) evaluated in Python double, then weak-cast to f32 */
    const float beta = (float)cfg->beta;
    for (int n = 0; n < cfg->iters; ++n) {
        float box6[6];
        for (int k = 0; k < 6; ++k) box6[k] = (float)g[k];
        bfo_evaluate(box6, t_c, pst, P, n_eval, rot, poses, V, K16, search, cfg->img_h, cfg->img_w, fitness);
        ++it;
        float min_iou, mt[6];
        const int success = bfo_cal_transform(fitness, pst, P, search, &min_iou, mt);
        bfo_update_pst(min_iou, mt, cfg->center_scale, cfg->shape_scale, search);
        if (previous_success && success)
            for (int k = 0; k < 6; ++k) search[k] = beta * search[k] + omb * prev[k];
        if (success) {
            need_update = 1; previous_success = 1; fail = 0;
            for (int k = 0; k < 6; ++k) { g[k] += (double)mt[k]; prev[k] = search[k]; }
        } else { ++fail; previous_success = 0; }
        if (trace) {
            trace[8 * n + 0] = (float)success; trace[8 * n + 1] = min_iou;
            for (int k = 0; k < 6; ++k) trace[8 * n + 2 + k] = search[k];
        }
        if (cfg->early_stop && fail >= 3) break;
    }
    free(fitness);
    *n_iters = it;
    if (need_update) {
        for (int k = 3; k < 6; ++k) if (g[k] < 0.01) g[k] = 0.01;
        for (int k = 0; k < 6; ++k) out6[k] = (float)g[k];
    }
    return need_update;
}
