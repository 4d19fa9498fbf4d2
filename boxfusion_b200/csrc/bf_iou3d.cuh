// Device-side building blocks of the oriented-3D IoU (K1), shared by the IoU-matrix entry and the
// NMS entry.  Reference semantics: instances.py:493-613 (see include/boxfusion_b200.h).
//
// SAMPLED_REF.  The reference takes its half-spaces from Qhull.  For the 8 float32 corners of a
// box the convex hull is unique: each face is folded along its convex diagonal into two triangles
// and Qhull reports one unit-normal plane per triangle.  bf_face_planes() rebuilds those 12 planes
// in float64 from the float32 corners (agreement with scipy <= 1e-14, tests/test_oracle_golden.py),
// so the 25^3 inside-counts are the reference's.
//
// Counting.  For a grid row (fixed y_j, z_k) the float64 plane value g(i) = fma(nx, x_i, b) is a
// monotone function of i because x_i is non-decreasing and rounding is monotone; the points of the
// row inside a box therefore form one index interval, found per plane by bisection on the exact
// predicate g(i) <= 1e-6 (<= 5 evaluations instead of 25).  No approximation is involved.
#pragma once
#include "bf_common.cuh"

#define BF_NS 25
#define BF_INSIDE_EPS 1e-6

__constant__ int c_bf_faces[6][4] = {{0, 3, 7, 4}, {1, 2, 6, 5}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
__constant__ int c_bf_edges[12][2] = {{0, 1}, {0, 4}, {1, 5}, {4, 5}, {2, 3}, {2, 6}, {6, 7}, {3, 7}, {0, 3}, {4, 7}, {1, 2}, {5, 6}};

// plane through p0,p1,p2 (float64), unit normal oriented away from `inner`; no FMA so that it is
// the same arithmetic as the CPU restatement (oracle/assoc_oracle.c).
__device__ __forceinline__ void bf_plane3(const double* p0, const double* p1, const double* p2, const double* inner,
                                          double* out) {
    const double ux = __dsub_rn(p1[0], p0[0]), uy = __dsub_rn(p1[1], p0[1]), uz = __dsub_rn(p1[2], p0[2]);
    const double vx = __dsub_rn(p2[0], p0[0]), vy = __dsub_rn(p2[1], p0[1]), vz = __dsub_rn(p2[2], p0[2]);
    double nx = __dsub_rn(__dmul_rn(uy, vz), __dmul_rn(uz, vy));
    double ny = __dsub_rn(__dmul_rn(uz, vx), __dmul_rn(ux, vz));
    double nz = __dsub_rn(__dmul_rn(ux, vy), __dmul_rn(uy, vx));
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
    nx = __ddiv_rn(nx, nrm); ny = __ddiv_rn(ny, nrm); nz = __ddiv_rn(nz, nrm);
    double d = -__dadd_rn(__dadd_rn(__dmul_rn(p0[0], nx), __dmul_rn(p0[1], ny)), __dmul_rn(p0[2], nz));
    const double s = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(inner[0], nx), __dmul_rn(inner[1], ny)), __dmul_rn(inner[2], nz)), d);
    if (s > 0) { nx = -nx; ny = -ny; nz = -nz; d = -d; }
    out[0] = nx; out[1] = ny; out[2] = nz; out[3] = d;
}

// The two outward half-spaces n.p + d <= 0 of face f of hull(8 float32 corners); planes[2][4].
__device__ inline void bf_face_planes(const float* __restrict__ c24, int f, double* __restrict__ planes) {
    double c[8][3], cen[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) { c[i][k] = (double)c24[3 * i + k]; cen[k] = __dadd_rn(cen[k], c[i][k]); }
#pragma unroll
    for (int k = 0; k < 3; ++k) cen[k] = cen[k] * 0.125;
    const double *a = c[c_bf_faces[f][0]], *b = c[c_bf_faces[f][1]], *cc = c[c_bf_faces[f][2]], *d = c[c_bf_faces[f][3]];
    double P[4];
    bf_plane3(a, b, cc, cen, P);
    const double s = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[0], d[0]), __dmul_rn(P[1], d[1])), __dmul_rn(P[2], d[2])), P[3]);
    if (s <= 0) {                                          // d below abc: the face folds along a-c
#pragma unroll
        for (int k = 0; k < 4; ++k) planes[k] = P[k];
        bf_plane3(a, cc, d, cen, planes + 4);
    } else {                                               // folds along b-d
        bf_plane3(a, b, d, cen, planes);
        bf_plane3(b, cc, d, cen, planes + 4);
    }
}

__device__ __forceinline__ bool bf_inside12(const double* __restrict__ pl, double x, double y, double z) {
#pragma unroll 4
    for (int f = 0; f < 12; ++f) {
        const double v = fma(x, pl[4 * f], fma(y, pl[4 * f + 1], fma(z, pl[4 * f + 2], pl[4 * f + 3])));
        if (!(v <= BF_INSIDE_EPS)) return false;
    }
    return true;
}

// np.linspace(float32 lo, float32 hi, 25) under NumPy 2 (float32): x_i = fl(fl(i*step)+lo), x_24 = hi.
__device__ __forceinline__ float bf_linspace25(float lo, float hi, int i) {
    if (i == BF_NS - 1) return hi;
    const float delta = __fsub_rn(hi, lo);
    const float step = __fdiv_rn(delta, 24.0f);
    float y = (float)i;
    if (step == 0.0f) y = __fmul_rn(__fdiv_rn(y, 24.0f), delta); else y = __fmul_rn(y, step);
    return __fadd_rn(y, lo);
}

// Index interval [lo,hi] (inclusive; empty when lo > hi) of the row points inside the 12 half-spaces.
// xs: 25 non-decreasing abscissae (float64 copies of the float32 grid); pl: [12][4].
__device__ __forceinline__ void bf_row_interval(const double* __restrict__ pl, const double* __restrict__ xs,
                                                double y, double z, int& lo_out, int& hi_out) {
    int lo = 0, hi = BF_NS - 1;
    for (int f = 0; f < 12; ++f) {
        const double nx = pl[4 * f];
        const double b = fma(y, pl[4 * f + 1], fma(z, pl[4 * f + 2], pl[4 * f + 3]));
        const bool in_lo = fma(xs[lo], nx, b) <= BF_INSIDE_EPS;
        const bool in_hi = fma(xs[hi], nx, b) <= BF_INSIDE_EPS;
        if (in_lo && in_hi) continue;                 // monotone: the whole interval satisfies this plane
        if (!in_lo && !in_hi) { lo = 1; hi = 0; break; }
        int a = lo, c = hi;                            // exactly one end inside: bisect for the flip
        if (in_lo) {                                   // inside on [lo..t], outside after
            while (c - a > 1) { const int m = (a + c) >> 1; if (fma(xs[m], nx, b) <= BF_INSIDE_EPS) a = m; else c = m; }
            hi = a;
        } else {                                       // outside before t, inside on [t..hi]
            while (c - a > 1) { const int m = (a + c) >> 1; if (fma(xs[m], nx, b) <= BF_INSIDE_EPS) c = m; else a = m; }
            lo = c;
        }
    }
    lo_out = lo; hi_out = hi;
}
