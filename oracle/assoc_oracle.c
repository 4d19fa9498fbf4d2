/* TEST INFRASTRUCTURE ONLY - CPU restatement (plain C) of the reference's sampled oriented-3D IoU.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product path (boxfusion_b200/) never does.
 *
 * Restates (paths relative to /root/reference):
 *   - Instances3D.augment_vertices        boxfusion/instances.py:493-512
 *   - Instances3D.check_intersection      boxfusion/instances.py:514-557   (the "gate")
 *   - Instances3D.batch_in_convex_hull_3d boxfusion/instances.py:561-571
 *   - Instances3D.obb_iou                 boxfusion/instances.py:573-613   (25^3 grid estimate)
 *
 * Third-party arithmetic that is not in the reference tree: scipy.spatial.ConvexHull (Qhull;
 * reference pins scipy==1.15.3, build container has 1.18.1) supplies the half-space equations at
 * instances.py:532-533,563.  For 8 box corners in float32 the hull is unique: every box face is
 * split into two triangles along the convex diagonal and Qhull reports one unit-normal plane per
 * triangle (12 rows).  `bfo_hull_planes` recomputes exactly those 12 planes in double from the
 * float32 corners; tests/test_oracle_golden.py checks them against scipy's `equations` (agreement
 * <= 1e-14) and oracle/port.py keeps a slow scipy-based twin (obb_counts_scipy) for pinning against the
 * reference itself.  np.linspace on float32 end points is float32 under NumPy 2:
 * x_i = fl(fl(i*step)+start), step = fl(fl(stop-start)/24), x_24 = stop.
 */
#include <math.h>
#include <stdint.h>

#define NS 25

/* faces in cyclic vertex order (vertex numbering of boxes.py:756-766) */
static const int FACES[6][4] = {{0, 3, 7, 4}, {1, 2, 6, 5}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
/* instances.py:495-499 */
static const int EDGES[12][2] = {{0, 1}, {0, 4}, {1, 5}, {4, 5}, {2, 3}, {2, 6}, {6, 7}, {3, 7}, {0, 3}, {4, 7}, {1, 2}, {5, 6}};

static void plane3(const double* p0, const double* p1, const double* p2, const double* inner, double* out) {
    const double ux = p1[0] - p0[0], uy = p1[1] - p0[1], uz = p1[2] - p0[2];
    const double vx = p2[0] - p0[0], vy = p2[1] - p0[1], vz = p2[2] - p0[2];
    double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    const double nrm = sqrt(nx * nx + ny * ny + nz * nz);
    nx /= nrm; ny /= nrm; nz /= nrm;
    double d = -(p0[0] * nx + p0[1] * ny + p0[2] * nz);
    if (inner[0] * nx + inner[1] * ny + inner[2] * nz + d > 0) { nx = -nx; ny = -ny; nz = -nz; d = -d; }
    out[0] = nx; out[1] = ny; out[2] = nz; out[3] = d;
}

/* 12 outward unit-normal half-spaces n.p + d <= 0 of the convex hull of 8 float32 corners. */
void bfo_hull_planes(const float* c24, double* planes /*[12][4]*/) {
    double c[8][3], cen[3] = {0, 0, 0};
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) { c[i][k] = (double)c24[3 * i + k]; cen[k] += c[i][k]; }
    for (int k = 0; k < 3; ++k) cen[k] /= 8.0;
    for (int f = 0; f < 6; ++f) {
        const double *a = c[FACES[f][0]], *b = c[FACES[f][1]], *cc = c[FACES[f][2]], *d = c[FACES[f][3]];
        double P[4];
        plane3(a, b, cc, cen, P);
        if (P[0] * d[0] + P[1] * d[1] + P[2] * d[2] + P[3] <= 0) {       /* d below abc: fold along a-c */
            for (int k = 0; k < 4; ++k) planes[(2 * f) * 4 + k] = P[k];
            plane3(a, cc, d, cen, planes + (2 * f + 1) * 4);
        } else {                                                          /* fold along b-d */
            plane3(a, b, d, cen, planes + (2 * f) * 4);
            plane3(b, cc, d, cen, planes + (2 * f + 1) * 4);
        }
    }
}

static int inside12(const double* pl, double x, double y, double z) {
    for (int f = 0; f < 12; ++f)
        if (!(x * pl[4 * f] + y * pl[4 * f + 1] + z * pl[4 * f + 2] + pl[4 * f + 3] <= 1e-6)) return 0;
    return 1;
}

static void augment(const float* c24, float* out60) {       /* 8 corners + 12 float32 edge midpoints */
    for (int i = 0; i < 24; ++i) out60[i] = c24[i];
    for (int e = 0; e < 12; ++e)
        for (int k = 0; k < 3; ++k) out60[24 + 3 * e + k] = (c24[3 * EDGES[e][0] + k] + c24[3 * EDGES[e][1] + k]) / 2;
}

/* check_intersection, instances.py:514-557 */
int bfo_gate(const float* c1, const float* c2, const double* pl1, const double* pl2) {
    float a1[60], a2[60];
    augment(c1, a1); augment(c2, a2);
    int s = 0;
    for (int i = 0; i < 20; ++i) {
        s += inside12(pl2, a1[3 * i], a1[3 * i + 1], a1[3 * i + 2]);
        s += inside12(pl1, a2[3 * i], a2[3 * i + 1], a2[3 * i + 2]);
    }
    return s > 0;
}

static void linspace25(float lo, float hi, float* out) {    /* np.linspace(float32, float32, 25) */
    const float delta = hi - lo;
    const float step = delta / 24;
    for (int i = 0; i < NS; ++i) {
        float y = (float)i;
        if (step == 0) { y = y / 24; y = y * delta; } else y = y * step;
        out[i] = y + lo;
    }
    out[NS - 1] = hi;
}

/* obb_iou, instances.py:573-613.  counts = {count1, count2, common}; returns the gate. */
int bfo_obb_counts(const float* c1, const float* c2, int32_t* counts) {
    double pl1[48], pl2[48];
    bfo_hull_planes(c1, pl1); bfo_hull_planes(c2, pl2);
    counts[0] = counts[1] = counts[2] = 0;
    if (!bfo_gate(c1, c2, pl1, pl2)) return 0;
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = hi[k] = c1[k];
        for (int i = 0; i < 8; ++i) {
            lo[k] = fminf(lo[k], fminf(c1[3 * i + k], c2[3 * i + k]));
            hi[k] = fmaxf(hi[k], fmaxf(c1[3 * i + k], c2[3 * i + k]));
        }
    }
    float xs[NS], ys[NS], zs[NS];
    linspace25(lo[0], hi[0], xs); linspace25(lo[1], hi[1], ys); linspace25(lo[2], hi[2], zs);
    int32_t n1 = 0, n2 = 0, n12 = 0;
    for (int i = 0; i < NS; ++i)
        for (int j = 0; j < NS; ++j)
            for (int k = 0; k < NS; ++k) {
                const int in1 = inside12(pl1, xs[i], ys[j], zs[k]);
                const int in2 = inside12(pl2, xs[i], ys[j], zs[k]);
                n1 += in1; n2 += in2; n12 += in1 & in2;
            }
    counts[0] = n1; counts[1] = n2; counts[2] = n12;
    return 1;
}

double bfo_iou_from_counts(const int32_t* c) {              /* instances.py:608 (float64) */
    return (double)c[2] / ((double)(c[0] + c[1] - c[2]) + 1e-6);
}

/* many pairs of an [N,8,3] corner array; OpenMP over pairs */
void bfo_obb_counts_pairs(const float* corners, const int32_t* ia, const int32_t* ib, int npairs,
                          int32_t* counts /*[npairs,3]*/, int32_t* gate /*[npairs]*/) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int p = 0; p < npairs; ++p)
        gate[p] = bfo_obb_counts(corners + 24 * ia[p], corners + 24 * ib[p], counts + 3 * p);
}

/* boxes.py:725-778 (torch CPU bmm on 3x3 @ 3x8 rounds as ((r0*v0 + r1*v1) + r2*v2) + t, float32) */
void bfo_corners(const float* tensor /*[N,6]*/, const float* R /*[N,9]*/, int N, float* out /*[N,8,3]*/) {
    static const float SX[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, SY[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, SZ[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
    for (int n = 0; n < N; ++n) {
        const float* t = tensor + 6 * n; const float* r = R + 9 * n;
        const float hl = t[3] / 2, hh = t[4] / 2, hw = t[5] / 2;
        for (int i = 0; i < 8; ++i) {
            const float vx = SX[i] * hl, vy = SY[i] * hh, vz = SZ[i] * hw;
            for (int j = 0; j < 3; ++j)
                out[24 * n + 3 * i + j] = ((r[3 * j] * vx + r[3 * j + 1] * vy) + r[3 * j + 2] * vz) + t[j];
        }
    }
}

/* calculate_obb_iou (instances.py:106-125): one box against k others, IoU in float64; OpenMP over the others */
void bfo_obb_iou_one_vs_many(const float* c1, const float* others /*[k,8,3]*/, int k, double* iou /*[k]*/) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < k; ++i) {
        int32_t cnt[3];
        iou[i] = bfo_obb_counts(c1, others + 24 * i, cnt) ? bfo_iou_from_counts(cnt) : 0.0;
    }
}
