// K1 - batched oriented-3D IoU and its use by the NMS entry.
//
// Pipeline (all on one stream, no host round trip):
//   bf_planes_kernel     per box: 12 float64 hull planes + float32 AABB as two float4 (6 threads per box)
//   bf_pairs_kernel      tiles of 8 x 128 pairs: the A boxes' AABBs staged in shared memory, every thread keeps its B box's
//                        AABB in registers (two coalesced float4 loads); exact AABB reject -> analytic IoU (ANALYTIC mode,
//                        co-axial pairs; corners read as float4, Sutherland-Hodgman clip in registers) or append to a
//                        candidate list
//   bf_count_kernel      per candidate, one CTA: containment gate (40 points over the threads), then the
//                        25^3 inside-counts by per-row bisection (625 rows over the threads), IoU in
//                        float64; writes the dense matrix and/or NMS mask bits       (persistent grid)
//
// Bound: FP64/FP32 CUDA-core issue, not HBM (inputs are KBs).
#include "bf_iou3d.cuh"

struct bf_work_item { int a, b; };

// ------------------------------------------------------------------------------------------------
// One thread per (box, face): the two triangle planes of that face; the first thread of a box also writes the AABB.
__global__ void bf_planes_kernel(const float* __restrict__ corners, const bf_dimref Nd, double* __restrict__ planes,
                                 float* __restrict__ aabb) {
    const int N = bf_dim(Nd);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 6 * N; t += gridDim.x * blockDim.x) {
    const int n = t / 6, f = t - 6 * n;
    float c[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) c[k] = corners[24 * n + k];
    double pl[8];
    bf_face_planes(c, f, pl);
#pragma unroll
    for (int k = 0; k < 8; ++k) planes[48 * (size_t)n + 8 * f + k] = pl[k];
    if (f == 0) {
        float lo[3] = {c[0], c[1], c[2]}, hi[3] = {c[0], c[1], c[2]};
#pragma unroll
        for (int i = 1; i < 8; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], c[3 * i + k]); hi[k] = fmaxf(hi[k], c[3 * i + k]); }
#pragma unroll
        for (int k = 0; k < 3; ++k) { aabb[8 * n + k] = lo[k]; aabb[8 * n + 4 + k] = hi[k]; }
        aabb[8 * n + 3] = 0.f; aabb[8 * n + 7] = 0.f;         // two float4 per box: (lo.xyz, 0), (hi.xyz, 0)
    }
    }
}

// ------------------------------------------------------------------------------------------------
// Analytic IoU of two boxes that share an axis (gravity-aligned boxes): BEV Sutherland-Hodgman clip of
// B's footprint against A's rectangle x overlap along the shared axis, float64.  Returns false when
// no pair of box axes is parallel within 1-1e-6 (caller falls back to the sampled estimator).
// Round 2: the 24 corner floats of each box arrive as six float4 loads, and the clip polygon (at most 8 vertices) lives in
// registers - every loop below is fully unrolled over static indices, the append position is a predicate, not an address.
// Work on the straight-line path (published as SURVEY section 8(d) asks): frames 2 x 66, axis search <= 9 x 5, footprint 44,
// clip 4 edges x 8 slots x (compare + interpolate 10) = 320, shoelace 8 x 4, volumes and ratio 14: ~600 flop per co-axial pair
// (float64), 48 B per box in, 8 B out.
struct bf_frame { double c[3]; double ax[3][3]; double half[3]; };

__device__ __forceinline__ void bf_load_corners(const float* __restrict__ corners, int n, float (&c)[24]) {
    const float4* q = reinterpret_cast<const float4*>(corners + 24 * (size_t)n);     // 96 B per box: 16-byte aligned rows
#pragma unroll
    for (int k = 0; k < 6; ++k) { const float4 v = __ldg(q + k); c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w; }
}

__device__ __forceinline__ void bf_frame_from_corners(const float (&c24)[24], bf_frame& F) {
    // v1-v0 = l along X, v3-v0 = h along Y, v4-v0 = w along Z (boxes.py:756-766)
#pragma unroll
    for (int k = 0; k < 3; ++k) F.c[k] = 0.5 * ((double)c24[k] + (double)c24[18 + k]);     // (v0+v6)/2
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int other = (a == 0) ? 1 : (a == 1 ? 3 : 4);
        double e[3], n2 = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) { e[k] = (double)c24[3 * other + k] - (double)c24[k]; n2 += e[k] * e[k]; }
        const double len = sqrt(n2);
        F.half[a] = 0.5 * len;
#pragma unroll
        for (int k = 0; k < 3; ++k) F.ax[a][k] = len > 0 ? e[k] / len : 0.0;
    }
}

// row a of a 3x3 / 3-vector held in registers, selected without an address (a is data-dependent)
__device__ __forceinline__ double bf_sel3(const double (&v)[3], int a) { return a == 0 ? v[0] : (a == 1 ? v[1] : v[2]); }
__device__ __forceinline__ void bf_row3(const double (&m)[3][3], int a, double (&r)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = a == 0 ? m[0][k] : (a == 1 ? m[1][k] : m[2][k]);
}

__device__ __forceinline__ bool bf_analytic_iou(const float (&ca)[24], const float (&cb)[24], double* iou_out) {
    bf_frame A, B;
    bf_frame_from_corners(ca, A);
    bf_frame_from_corners(cb, B);
    int ia = -1, ib = -1;
    // prefer the gravity axis (local Y, index 1) of both boxes
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int a = (i == 0) ? 1 : (i == 1 ? 0 : 2), b = (j == 0) ? 1 : (j == 1 ? 0 : 2);
            const double d = A.ax[a][0] * B.ax[b][0] + A.ax[a][1] * B.ax[b][1] + A.ax[a][2] * B.ax[b][2];
            if (ia < 0 && fabs(d) >= 1.0 - 1e-6) { ia = a; ib = b; }
        }
    if (ia < 0) return false;
    double u[3], P[3], Q[3], Bm[3], Bn[3];
    const int pa = (ia + 1) % 3, qa = (ia + 2) % 3;       // A's in-plane axes
    const int mb = (ib + 1) % 3, nb = (ib + 2) % 3;       // B's in-plane axes
    bf_row3(A.ax, ia, u); bf_row3(A.ax, pa, P); bf_row3(A.ax, qa, Q); bf_row3(B.ax, mb, Bm); bf_row3(B.ax, nb, Bn);
    const double hA = bf_sel3(A.half, ia), hB = bf_sel3(B.half, ib);
    // height overlap along u
    const double dc[3] = {B.c[0] - A.c[0], B.c[1] - A.c[1], B.c[2] - A.c[2]};
    const double hb = dc[0] * u[0] + dc[1] * u[1] + dc[2] * u[2];
    const double top = fmin(hA, hb + hB), bot = fmax(-hA, hb - hB);
    const double oh = top - bot;
    const double volA = 8.0 * A.half[0] * A.half[1] * A.half[2], volB = 8.0 * B.half[0] * B.half[1] * B.half[2];
    if (oh <= 0) { *iou_out = 0.0; return true; }
    // B footprint in A's (p,q) coordinates
    const double cp = dc[0] * P[0] + dc[1] * P[1] + dc[2] * P[2], cq = dc[0] * Q[0] + dc[1] * Q[1] + dc[2] * Q[2];
    const double hm = bf_sel3(B.half, mb), hn = bf_sel3(B.half, nb);
    const double mp = (Bm[0] * P[0] + Bm[1] * P[1] + Bm[2] * P[2]) * hm;
    const double mq = (Bm[0] * Q[0] + Bm[1] * Q[1] + Bm[2] * Q[2]) * hm;
    const double np_ = (Bn[0] * P[0] + Bn[1] * P[1] + Bn[2] * P[2]) * hn;
    const double nq = (Bn[0] * Q[0] + Bn[1] * Q[1] + Bn[2] * Q[2]) * hn;
    double px[8], py[8];
    px[0] = cp - mp - np_; py[0] = cq - mq - nq;
    px[1] = cp + mp - np_; py[1] = cq + mq - nq;
    px[2] = cp + mp + np_; py[2] = cq + mq + nq;
    px[3] = cp - mp + np_; py[3] = cq - mq + nq;
#pragma unroll
    for (int k = 4; k < 8; ++k) { px[k] = 0.0; py[k] = 0.0; }
    int n = 4;
    const double ha = bf_sel3(A.half, pa), hq = bf_sel3(A.half, qa);
    // Sutherland-Hodgman against x<=ha, x>=-ha, y<=hq, y>=-hq; a rectangle clipped by four half-planes has <= 8 vertices
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double lim = (e < 2) ? ha : hq;
        const double sgn = (e & 1) ? -1.0 : 1.0;          // inside: sgn*coord <= lim
        double qx[8], qy[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { qx[k] = 0.0; qy[k] = 0.0; }
        int m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < n) {
                // successor of vertex i (vertex 0 after the last one): static index, dynamic wrap
                double jx = px[0], jy = py[0];
                if (i + 1 < n) { jx = px[(i + 1) & 7]; jy = py[(i + 1) & 7]; }
                const double ci = sgn * ((e < 2) ? px[i] : py[i]), cj = sgn * ((e < 2) ? jx : jy);
                const bool ini = ci <= lim, inj = cj <= lim;
                if (ini) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (k == m) { qx[k] = px[i]; qy[k] = py[i]; }
                    ++m;
                }
                if (ini != inj) {
                    const double t = (lim - ci) / (cj - ci);
                    const double nx = px[i] + t * (jx - px[i]), ny = py[i] + t * (jy - py[i]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (k == m) { qx[k] = nx; qy[k] = ny; }
                    ++m;
                }
            }
        }
        n = m < 8 ? m : 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) { px[k] = qx[k]; py[k] = qy[k]; }
    }
    double area = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < n) {
            double jx = px[0], jy = py[0];
            if (i + 1 < n) { jx = px[(i + 1) & 7]; jy = py[(i + 1) & 7]; }
            area += px[i] * jy - jx * py[i];
        }
    area = 0.5 * fabs(area);
    const double vi = area * oh;
    *iou_out = vi / (volA + volB - vi);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Tiles of BF_TILE_A x 128 pairs (grid-stride over the tiles).  triangle != 0: A is rows [a_off, a_off + M) of B and only
// pairs with a_off + a < b are evaluated (NMS).  Outputs: dense iou/counts zero-filled (when given), work list of gate-passing
// pairs, stats.
// counters: [0] work items, [1] pairs, [2] AABB-passing, [3] gate-passing, [4] analytic, [5] overflow, [6] NMS edges
// M / N are host values or read from device memory (bf_dimref); the mask row stride is W = ceil(N/32) of the actual N.
// Data movement: the A boxes' AABBs of a tile are staged in shared memory as float4 (SoA: lo, hi), every thread keeps the AABB
// of its B box in registers (two float4 loads, 32 contiguous bytes per thread: coalesced), the zero fill of the dense outputs
// is coalesced over b for a fixed a.
#define BF_TILE_A 8
__global__ void __launch_bounds__(128)
bf_pairs_kernel(const float* __restrict__ cornersA, const float* __restrict__ aabbA,
                                const double* __restrict__ planesA, const bf_dimref Md, const float* __restrict__ cornersB,
                                const float* __restrict__ aabbB, const double* __restrict__ planesB, const bf_dimref Nd,
                                int triangle, int a_off, int mode, double* __restrict__ iou, int32_t* __restrict__ counts,
                                bf_work_item* __restrict__ work, int work_cap, unsigned long long* __restrict__ counters,
                                // NMS outputs (ANALYTIC hits are thresholded here)
                                double thr, const int32_t* __restrict__ rank, uint32_t* __restrict__ mask,
                                uint32_t* __restrict__ rowany, unsigned long long* __restrict__ edges, int edge_cap) {
    __shared__ float4 s_lo[BF_TILE_A], s_hi[BF_TILE_A];
    const int M = bf_dim(Md), N = bf_dim(Nd);
    const long long total = (long long)M * N;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        counters[1] = triangle ? (unsigned long long)M * (unsigned long long)(2LL * (N - a_off) - M - 1 > 0 ? 2LL * (N - a_off) - M - 1 : 0) / 2ULL : (unsigned long long)total;
    const int tiles_a = (M + BF_TILE_A - 1) / BF_TILE_A, tiles_b = (N + 127) / 128;
    const float4* bb4 = reinterpret_cast<const float4*>(aabbB);
    const float4* ba4 = reinterpret_cast<const float4*>(aabbA);
    const float m = 1e-4f;
    unsigned local_aabb = 0;
    for (long long tile = blockIdx.x; tile < (long long)tiles_a * tiles_b; tile += gridDim.x) {
        const int ta = (int)(tile / tiles_b), tb = (int)(tile - (long long)ta * tiles_b);
        const int a0 = ta * BF_TILE_A, b = tb * 128 + threadIdx.x;
        if (triangle && a0 + a_off >= tb * 128 + 127) {               // the whole tile lies on or below the diagonal
            if (!iou && !counts) continue;
        }
        __syncthreads();
        if (threadIdx.x < BF_TILE_A && a0 + threadIdx.x < M) { s_lo[threadIdx.x] = __ldg(ba4 + 2 * (a0 + threadIdx.x)); s_hi[threadIdx.x] = __ldg(ba4 + 2 * (a0 + threadIdx.x) + 1); }
        __syncthreads();
        if (b >= N) continue;
        const float4 blo = __ldg(bb4 + 2 * b), bhi = __ldg(bb4 + 2 * b + 1);
#pragma unroll
        for (int k = 0; k < BF_TILE_A; ++k) {
            const int a = a0 + k;
            if (a >= M) break;
            const long long p = (long long)a * N + b;
            if (iou) iou[p] = 0.0;
            if (counts) { counts[3 * p] = 0; counts[3 * p + 1] = 0; counts[3 * p + 2] = 0; }
            if (triangle && a + a_off >= b) continue;
            // exact reject: a point within 1e-6 of every face plane of a box lies within its AABB grown by 1e-4
            const float4 alo = s_lo[k], ahi = s_hi[k];
            if (alo.x > bhi.x + m || blo.x > ahi.x + m || alo.y > bhi.y + m || blo.y > ahi.y + m || alo.z > bhi.z + m ||
                blo.z > ahi.z + m)
                continue;
            ++local_aabb;
            const unsigned long long slot = atomicAdd(&counters[0], 1ULL);      // candidate: gate + counts in bf_count_kernel
            if (slot < (unsigned long long)work_cap) { work[slot].a = a; work[slot].b = b; }
            else atomicExch(&counters[5], 1ULL);
        }
    }
    if (local_aabb) atomicAdd(&counters[2], (unsigned long long)local_aabb);
}

// ------------------------------------------------------------------------------------------------
// ANALYTIC mode: one thread per AABB-passing candidate (grid-stride over the work list).  Co-axial pairs get their IoU here
// and leave the list (a = -1); the others stay for the sampled estimator below.
__global__ void __launch_bounds__(64)
bf_analytic_kernel(const float* __restrict__ cornersA, const float* __restrict__ cornersB, const bf_dimref Nd, int a_off,
                   bf_work_item* __restrict__ work, int work_cap, unsigned long long* __restrict__ counters,
                   double* __restrict__ iou, double thr, const int32_t* __restrict__ rank, uint32_t* __restrict__ mask,
                   uint32_t* __restrict__ rowany, unsigned long long* __restrict__ edges, int edge_cap) {
    const int N = bf_dim(Nd);
    const int W = (N + 31) >> 5;
    unsigned long long nwork = counters[0];
    if (nwork > (unsigned long long)work_cap) nwork = work_cap;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwork; w += (unsigned long long)gridDim.x * blockDim.x) {
        const int a = work[w].a, b = work[w].b;
        float ca[24], cb[24];
        bf_load_corners(cornersA, a, ca);
        bf_load_corners(cornersB, b, cb);
        double v;
        if (!bf_analytic_iou(ca, cb, &v)) continue;
        work[w].a = -1;                                                   // done: bf_count_kernel skips it
        atomicAdd(&counters[4], 1ULL);
        if (iou) iou[(size_t)a * N + b] = v;
        if (rank && v > thr) {
            const int ra = rank[a + a_off], rb = rank[b];
            const int r0 = min(ra, rb), r1 = max(ra, rb);
            if (mask) {
                atomicOr(&mask[(size_t)r0 * W + (r1 >> 5)], 1u << (r1 & 31));
                atomicOr(&rowany[r0 >> 5], 1u << (r0 & 31));
            }
            const unsigned long long e = atomicAdd(&counters[6], 1ULL);
            if (e < (unsigned long long)edge_cap) edges[e] = ((unsigned long long)r0 << 32) | (unsigned long long)r1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One CTA per candidate pair (persistent, CTA-stride over the work list): gate, then counts.
#define BF_COUNT_THREADS 640     // 625 grid rows: one per thread (a keyframe has a few dozen candidate pairs: latency, not throughput)
__global__ void __launch_bounds__(BF_COUNT_THREADS)
bf_count_kernel(const float* __restrict__ cornersA, const float* __restrict__ aabbA, const double* __restrict__ planesA,
                const float* __restrict__ cornersB, const float* __restrict__ aabbB, const double* __restrict__ planesB,
                const bf_dimref Nd, int a_off, const bf_work_item* __restrict__ work, int work_cap, unsigned long long* __restrict__ counters,
                double* __restrict__ iou, int32_t* __restrict__ counts, double thr, const int32_t* __restrict__ rank,
                uint32_t* __restrict__ mask, uint32_t* __restrict__ rowany, unsigned long long* __restrict__ edges,
                int edge_cap) {
    const int N = bf_dim(Nd);
    const int W = (N + 31) >> 5;
    __shared__ double s_pl[2][48];
    __shared__ double s_grid[3][BF_NS];
    __shared__ float s_c[2][24];
    __shared__ int s_red[3][BF_COUNT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long nwork = counters[0];
    if (nwork > (unsigned long long)work_cap) nwork = work_cap;
    for (unsigned long long w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int a = work[w].a, b = work[w].b;
        if (a < 0) continue;                                              // ANALYTIC mode settled this pair (block-uniform)
        __syncthreads();
        if (tid < 48) { s_pl[0][tid] = planesA[48 * (size_t)a + tid]; s_pl[1][tid] = planesB[48 * (size_t)b + tid]; }
        if (tid >= 64 && tid < 88) { s_c[0][tid - 64] = cornersA[24 * (size_t)a + tid - 64]; s_c[1][tid - 64] = cornersB[24 * (size_t)b + tid - 64]; }
        if (tid >= 96 && tid < 96 + 3 * BF_NS) {
            const int ax = (tid - 96) / BF_NS, i = (tid - 96) - ax * BF_NS;
            const float lo = fminf(aabbA[8 * a + ax], aabbB[8 * b + ax]), hi = fmaxf(aabbA[8 * a + 4 + ax], aabbB[8 * b + 4 + ax]);   // instances.py:581-582
            s_grid[ax][i] = (double)bf_linspace25(lo, hi, i);
        }
        __syncthreads();
        // ---- containment gate (instances.py:514-557): 2 x 20 points, one per thread -------------------------
        int inside = 0;
        if (tid < 40) {
            const int which = tid / 20, k = tid - 20 * which;          // points of box `which` against the other hull
            const float* c = s_c[which];
            float px, py, pz;
            if (k < 8) { px = c[3 * k]; py = c[3 * k + 1]; pz = c[3 * k + 2]; }
            else {
                const int i0 = c_bf_edges[k - 8][0], i1 = c_bf_edges[k - 8][1];
                px = __fadd_rn(c[3 * i0], c[3 * i1]) * 0.5f; py = __fadd_rn(c[3 * i0 + 1], c[3 * i1 + 1]) * 0.5f;
                pz = __fadd_rn(c[3 * i0 + 2], c[3 * i1 + 2]) * 0.5f;
            }
            inside = bf_inside12(s_pl[1 - which], px, py, pz) ? 1 : 0;
        }
        if (!__syncthreads_or(inside)) continue;                        // gate failed: IoU stays 0
        if (tid == 0) atomicAdd(&counters[3], 1ULL);
        // ---- 25^3 counts: 625 rows over the threads ---------------------------------------------------------
        int n1 = 0, n2 = 0, n12 = 0;
        for (int r = tid; r < BF_NS * BF_NS; r += BF_COUNT_THREADS) {
            const double y = s_grid[1][r / BF_NS], z = s_grid[2][r % BF_NS];
            int lo1, hi1, lo2, hi2;
            bf_row_interval(s_pl[0], s_grid[0], y, z, lo1, hi1);
            bf_row_interval(s_pl[1], s_grid[0], y, z, lo2, hi2);
            const int c1 = max(0, hi1 - lo1 + 1), c2 = max(0, hi2 - lo2 + 1);
            n1 += c1; n2 += c2;
            if (c1 > 0 && c2 > 0) n12 += max(0, min(hi1, hi2) - max(lo1, lo2) + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n1 += __shfl_xor_sync(0xffffffffu, n1, o);
            n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            n12 += __shfl_xor_sync(0xffffffffu, n12, o);
        }
        if (lane == 0) { s_red[0][warp] = n1; s_red[1][warp] = n2; s_red[2][warp] = n12; }
        __syncthreads();
        if (tid == 0) {
            n1 = n2 = n12 = 0;
            for (int k = 0; k < BF_COUNT_THREADS / 32; ++k) { n1 += s_red[0][k]; n2 += s_red[1][k]; n12 += s_red[2][k]; }
            const double v = (double)n12 / ((double)(n1 + n2 - n12) + 1e-6);                  // instances.py:608
            const size_t p = (size_t)a * N + b;
            if (iou) iou[p] = v;
            if (counts) { counts[3 * p] = n1; counts[3 * p + 1] = n2; counts[3 * p + 2] = n12; }
            if (rank && v > thr) {
                const int ra = rank[a + a_off], rb = rank[b];
                const int r0 = min(ra, rb), r1 = max(ra, rb);
                if (mask) {
                    atomicOr(&mask[(size_t)r0 * W + (r1 >> 5)], 1u << (r1 & 31));
                    atomicOr(&rowany[r0 >> 5], 1u << (r0 & 31));
                }
                const unsigned long long e = atomicAdd(&counters[6], 1ULL);
                if (e < (unsigned long long)edge_cap) edges[e] = ((unsigned long long)r0 << 32) | (unsigned long long)r1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Host driver shared by bf_iou3d_matrix, bf_nms3d and the engine step.  Md / Nd: sizes on the host, or upper bounds on
// the host + the actual sizes in device memory (the captured engine step: scratch is sized for the bounds, the grids are
// fixed and the kernels stride).  The counters are zeroed here.
int bf_iou3d_run(bf_handle* h, const float* cornersA, bf_dimref Md, const float* cornersB, bf_dimref Nd, int triangle, int a_off, int mode,
                 double* iou, int32_t* counts, int64_t* stats, double thr, const int32_t* rank, uint32_t* mask,
                 uint32_t* rowany, unsigned long long* edges, int edge_cap, cudaStream_t st) {
    double *plA = nullptr, *plB = nullptr;
    float *bbA = nullptr, *bbB = nullptr;
    int rc;
    void* p;
    const int M = Md.host, N = Nd.host;                   // sizing values (bounds when the real sizes are on the device)
    const bool dev_sized = Md.dev || Nd.dev;
    const int stride_grid = h->sm_count * 8;
    if ((rc = bf_scratch(h, BF_SCRATCH_PLANES_A, sizeof(double) * 48 * (size_t)M, &p))) return rc; plA = (double*)p;
    if ((rc = bf_scratch(h, BF_SCRATCH_AABB_A, sizeof(float) * 8 * (size_t)M, &p))) return rc; bbA = (float*)p;
    {
        const int g = bf_blocks(6LL * M, 96);
        bf_planes_kernel<<<(dev_sized && g > stride_grid) ? stride_grid : g, 96, 0, st>>>(cornersA, Md, plA, bbA);
    }
    BF_LAUNCH_CHECK(h, "bf_planes_kernel");
    if (cornersB == cornersA && Nd.host == Md.host && Nd.dev == Md.dev) { plB = plA; bbB = bbA; }
    else {
        if ((rc = bf_scratch(h, BF_SCRATCH_PLANES_B, sizeof(double) * 48 * (size_t)N, &p))) return rc; plB = (double*)p;
        if ((rc = bf_scratch(h, BF_SCRATCH_AABB_B, sizeof(float) * 8 * (size_t)N, &p))) return rc; bbB = (float*)p;
        const int g = bf_blocks(6LL * N, 96);
        bf_planes_kernel<<<(dev_sized && g > stride_grid) ? stride_grid : g, 96, 0, st>>>(cornersB, Nd, plB, bbB);
        BF_LAUNCH_CHECK(h, "bf_planes_kernel");
    }
    const long long total = (long long)M * N;
    // work-list capacity: every pair for small problems, else 64 candidates per box (grown on overflow)
    long long cap = total < (1LL << 20) ? total : (1LL << 20) + 64LL * (M + N);
    if ((long long)(h->cap[BF_SCRATCH_WORK] / sizeof(bf_work_item)) > cap || h->frozen) cap = h->cap[BF_SCRATCH_WORK] / sizeof(bf_work_item);
    if (cap > 0x7fffffffLL) cap = 0x7fffffffLL;
    if ((rc = bf_scratch(h, BF_SCRATCH_WORK, sizeof(bf_work_item) * (size_t)cap, &p))) return rc;
    bf_work_item* work = (bf_work_item*)p;
    if ((rc = bf_scratch(h, BF_SCRATCH_COUNTERS, sizeof(unsigned long long) * 8, &p))) return rc;
    unsigned long long* counters = (unsigned long long*)p;
    BF_CUDA(h, cudaMemsetAsync(counters, 0, sizeof(unsigned long long) * 8, st));
    {
        // one CTA per tile of 8 x 128 pairs; fixed, machine-filling grid when the sizes live on the device, capped otherwise
        // (the kernel strides over the tiles)
        long long g = (long long)((M + BF_TILE_A - 1) / BF_TILE_A) * ((N + 127) / 128);
        const long long gcap = (long long)h->sm_count * (dev_sized ? 16 : 64);
        if (g > gcap) g = gcap;
        if (g < 1) g = 1;
        bf_pairs_kernel<<<(unsigned)g, 128, 0, st>>>(cornersA, bbA, plA, Md, cornersB, bbB, plB, Nd, triangle, a_off, mode,
                                                     iou, counts, work, (int)cap, counters, thr, rank, mask, rowany, edges, edge_cap);
    }
    BF_LAUNCH_CHECK(h, "bf_pairs_kernel");
    if (mode == BF_IOU_ANALYTIC) {
        bf_analytic_kernel<<<h->sm_count * 2, 64, 0, st>>>(cornersA, cornersB, Nd, a_off, work, (int)cap, counters, iou, thr, rank, mask,
                                                           rowany, edges, edge_cap);
        BF_LAUNCH_CHECK(h, "bf_analytic_kernel");
    }
    const int grid = h->sm_count * 3;
    bf_count_kernel<<<grid, BF_COUNT_THREADS, 0, st>>>(cornersA, bbA, plA, cornersB, bbB, plB, Nd, a_off, work, (int)cap, counters,
                                                       iou, counts, thr, rank, mask, rowany, edges, edge_cap);
    BF_LAUNCH_CHECK(h, "bf_count_kernel");
    if (stats)   // pairs, AABB-passing, gate-passing, analytic
        BF_CUDA(h, cudaMemcpyAsync(stats, counters + 1, sizeof(int64_t) * 4, cudaMemcpyDeviceToDevice, st));
    return BF_OK;
}

// Reads the overflow flag (synchronises the stream).  Returns 1 when the work list overflowed.
int bf_iou3d_overflowed(bf_handle* h, cudaStream_t st, int* overflow) {
    unsigned long long flag = 0;
    BF_CUDA(h, cudaMemcpyAsync(&flag, (unsigned long long*)h->buf[BF_SCRATCH_COUNTERS] + 5, sizeof(flag),
                               cudaMemcpyDeviceToHost, st));
    BF_CUDA(h, cudaStreamSynchronize(st));
    *overflow = flag != 0;
    return BF_OK;
}

extern "C" int bf_iou3d_matrix(bf_handle* h, const float* cornersA, int M, const float* cornersB, int N, int mode,
                               double* iou, int32_t* counts, int64_t* stats, void* stream) {
    bf_device_guard guard(h);
    if (!h || M < 0 || N < 0 || (mode != BF_IOU_SAMPLED_REF && mode != BF_IOU_ANALYTIC))
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_iou3d_matrix", "bad argument");
    if (M == 0 || N == 0) return BF_OK;
    if (!cornersA || !cornersB || !iou) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_iou3d_matrix", "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    for (int attempt = 0; attempt < 2; ++attempt) {
        int rc = bf_iou3d_run(h, cornersA, bf_dim_host(M), cornersB, bf_dim_host(N), 0, 0, mode, iou, counts, stats, 0.0, nullptr,
                              nullptr, nullptr, nullptr, 0, st);
        if (rc) return rc;
        if ((long long)M * N <= (long long)(h->cap[BF_SCRATCH_WORK] / sizeof(bf_work_item))) break;   // cannot overflow
        int ovf = 0;
        if ((rc = bf_iou3d_overflowed(h, st, &ovf))) return rc;
        if (!ovf) break;
        if (attempt == 1) return bf_fail(h, BF_ERR_CAPACITY, "bf_iou3d_matrix", "work list overflow");
        void* p;   // grow to the full pair count and redo
        if ((rc = bf_scratch(h, BF_SCRATCH_WORK, sizeof(bf_work_item) * (size_t)M * N, &p))) return rc;
    }
    return BF_OK;
}

// ------------------------------------------------------------------------------------------------
// Instances3D.batch_in_convex_hull_3d (instances.py:559-571): points against the 12 hull planes of one box.
__global__ void bf_points_in_hull_kernel(const float* __restrict__ corners, const double* __restrict__ points, int n,
                                         uint8_t* __restrict__ inside) {
    __shared__ double pl[48];
    if (threadIdx.x < 6) {
        float c[24];
#pragma unroll
        for (int k = 0; k < 24; ++k) c[k] = corners[k];
        bf_face_planes(c, threadIdx.x, pl + 8 * threadIdx.x);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inside[i] = bf_inside12(pl, points[3 * (size_t)i], points[3 * (size_t)i + 1], points[3 * (size_t)i + 2]) ? 1 : 0;
}

extern "C" int bf_points_in_hull(bf_handle* h, const float* corners, const double* points, int n, uint8_t* inside, void* stream) {
    bf_device_guard guard(h);
    if (!h || n < 0 || (n > 0 && (!corners || !points || !inside))) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_points_in_hull", "bad argument");
    if (n == 0) return BF_OK;
    bf_points_in_hull_kernel<<<bf_blocks(n, 128), 128, 0, (cudaStream_t)stream>>>(corners, points, n, inside);
    BF_LAUNCH_CHECK(h, "bf_points_in_hull_kernel");
    return BF_OK;
}
