// K1 - batched oriented-3D IoU and its use by the NMS entry.
//
// Pipeline (all on one stream, no host round trip):
//   bf_planes_kernel     per box: 12 float64 hull planes + float32 AABB            (N threads)
//   bf_pairs_kernel      per pair: exact AABB reject -> analytic IoU (ANALYTIC mode, co-axial pairs)
//                        or append to a candidate list                             (M*N threads)
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
        for (int k = 0; k < 3; ++k) { aabb[6 * n + k] = lo[k]; aabb[6 * n + 3 + k] = hi[k]; }
    }
    }
}

// ------------------------------------------------------------------------------------------------
// Analytic IoU of two boxes that share an axis (gravity-aligned boxes): BEV Sutherland-Hodgman clip of
// B's footprint against A's rectangle x overlap along the shared axis, float64.  Returns false when
// no pair of box axes is parallel within 1-1e-6 (caller falls back to the sampled estimator).
struct bf_frame { double c[3]; double ax[3][3]; double half[3]; };

__device__ inline void bf_frame_from_corners(const float* __restrict__ c24, bf_frame& F) {
    // v1-v0 = l along X, v3-v0 = h along Y, v4-v0 = w along Z (boxes.py:756-766)
    const int other[3] = {1, 3, 4};
    for (int k = 0; k < 3; ++k) F.c[k] = 0.5 * ((double)c24[k] + (double)c24[18 + k]);     // (v0+v6)/2
    for (int a = 0; a < 3; ++a) {
        double e[3], n2 = 0;
        for (int k = 0; k < 3; ++k) { e[k] = (double)c24[3 * other[a] + k] - (double)c24[k]; n2 += e[k] * e[k]; }
        const double len = sqrt(n2);
        F.half[a] = 0.5 * len;
        for (int k = 0; k < 3; ++k) F.ax[a][k] = len > 0 ? e[k] / len : 0.0;
    }
}

__device__ inline bool bf_analytic_iou(const float* __restrict__ ca, const float* __restrict__ cb, double* iou_out) {
    bf_frame A, B;
    bf_frame_from_corners(ca, A);
    bf_frame_from_corners(cb, B);
    int ia = -1, ib = -1;
    // prefer the gravity axis (local Y, index 1) of both boxes
    const int pref[3] = {1, 0, 2};
    for (int i = 0; i < 3 && ia < 0; ++i)
        for (int j = 0; j < 3; ++j) {
            const int a = pref[i], b = pref[j];
            const double d = A.ax[a][0] * B.ax[b][0] + A.ax[a][1] * B.ax[b][1] + A.ax[a][2] * B.ax[b][2];
            if (fabs(d) >= 1.0 - 1e-6) { ia = a; ib = b; break; }
        }
    if (ia < 0) return false;
    const double* u = A.ax[ia];
    const int pa = (ia + 1) % 3, qa = (ia + 2) % 3;       // A's in-plane axes
    const int mb = (ib + 1) % 3, nb = (ib + 2) % 3;       // B's in-plane axes
    // height overlap along u
    double dc[3] = {B.c[0] - A.c[0], B.c[1] - A.c[1], B.c[2] - A.c[2]};
    const double hb = dc[0] * u[0] + dc[1] * u[1] + dc[2] * u[2];
    const double top = fmin(A.half[ia], hb + B.half[ib]), bot = fmax(-A.half[ia], hb - B.half[ib]);
    const double oh = top - bot;
    const double volA = 8.0 * A.half[0] * A.half[1] * A.half[2], volB = 8.0 * B.half[0] * B.half[1] * B.half[2];
    if (oh <= 0) { *iou_out = 0.0; return true; }
    // B footprint in A's (p,q) coordinates
    const double* P = A.ax[pa]; const double* Q = A.ax[qa];
    const double cp = dc[0] * P[0] + dc[1] * P[1] + dc[2] * P[2], cq = dc[0] * Q[0] + dc[1] * Q[1] + dc[2] * Q[2];
    const double mp = (B.ax[mb][0] * P[0] + B.ax[mb][1] * P[1] + B.ax[mb][2] * P[2]) * B.half[mb];
    const double mq = (B.ax[mb][0] * Q[0] + B.ax[mb][1] * Q[1] + B.ax[mb][2] * Q[2]) * B.half[mb];
    const double np_ = (B.ax[nb][0] * P[0] + B.ax[nb][1] * P[1] + B.ax[nb][2] * P[2]) * B.half[nb];
    const double nq = (B.ax[nb][0] * Q[0] + B.ax[nb][1] * Q[1] + B.ax[nb][2] * Q[2]) * B.half[nb];
    double px[10], py[10], qx[10], qy[10];
    px[0] = cp - mp - np_; py[0] = cq - mq - nq;
    px[1] = cp + mp - np_; py[1] = cq + mq - nq;
    px[2] = cp + mp + np_; py[2] = cq + mq + nq;
    px[3] = cp - mp + np_; py[3] = cq - mq + nq;
    int n = 4;
    const double ha = A.half[pa], hq = A.half[qa];
    // Sutherland-Hodgman against x<=ha, x>=-ha, y<=hq, y>=-hq
    for (int e = 0; e < 4 && n > 0; ++e) {
        const double lim = (e < 2) ? ha : hq;
        const double sgn = (e & 1) ? -1.0 : 1.0;          // inside: sgn*coord <= lim
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1 == n) ? 0 : i + 1;
            const double ci = sgn * ((e < 2) ? px[i] : py[i]), cj = sgn * ((e < 2) ? px[j] : py[j]);
            const bool ini = ci <= lim, inj = cj <= lim;
            if (ini) { qx[m] = px[i]; qy[m] = py[i]; ++m; }
            if (ini != inj) {
                const double t = (lim - ci) / (cj - ci);
                qx[m] = px[i] + t * (px[j] - px[i]);
                qy[m] = py[i] + t * (py[j] - py[i]);
                ++m;
            }
        }
        n = m;
        for (int i = 0; i < n; ++i) { px[i] = qx[i]; py[i] = qy[i]; }
    }
    double area = 0;
    for (int i = 0; i < n; ++i) { const int j = (i + 1 == n) ? 0 : i + 1; area += px[i] * py[j] - px[j] * py[i]; }
    area = 0.5 * fabs(area);
    const double vi = area * oh;
    *iou_out = vi / (volA + volB - vi);
    return true;
}

// ------------------------------------------------------------------------------------------------
// One thread per pair (grid-stride).  triangle != 0: A and B are the same set and only a < b is evaluated (NMS).
// Outputs: dense iou/counts zero-filled (when given), work list of gate-passing pairs, stats.
// counters: [0] work items, [1] pairs, [2] AABB-passing, [3] gate-passing, [4] analytic, [5] overflow, [6] NMS edges
// M / N are host values or read from device memory (bf_dimref); the mask row stride is W = ceil(N/32) of the actual N.
__global__ void bf_pairs_kernel(const float* __restrict__ cornersA, const float* __restrict__ aabbA,
                                const double* __restrict__ planesA, const bf_dimref Md, const float* __restrict__ cornersB,
                                const float* __restrict__ aabbB, const double* __restrict__ planesB, const bf_dimref Nd,
                                int triangle, int a_off, int mode, double* __restrict__ iou, int32_t* __restrict__ counts,
                                bf_work_item* __restrict__ work, int work_cap, unsigned long long* __restrict__ counters,
                                // NMS outputs (ANALYTIC hits are thresholded here)
                                double thr, const int32_t* __restrict__ rank, uint32_t* __restrict__ mask,
                                uint32_t* __restrict__ rowany, unsigned long long* __restrict__ edges, int edge_cap) {
    const int M = bf_dim(Md), N = bf_dim(Nd);
    const int W = (N + 31) >> 5;
    const long long total = (long long)M * N;
    const long long p0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // triangle: A is rows [a_off, a_off + M) of B and only pairs with a_off + a < b are evaluated
    if (p0 == 0) counters[1] = triangle ? (unsigned long long)M * (unsigned long long)(2LL * (N - a_off) - M - 1 > 0 ? 2LL * (N - a_off) - M - 1 : 0) / 2ULL : (unsigned long long)total;
    for (long long p = p0; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(p / N), b = (int)(p % N);
        if (iou) iou[p] = 0.0;
        if (counts) { counts[3 * p] = 0; counts[3 * p + 1] = 0; counts[3 * p + 2] = 0; }
        if (triangle && a + a_off >= b) continue;
        // exact reject: a point within 1e-6 of every face plane of a box lies within its AABB grown by 1e-4
        const float* ba = aabbA + 6 * a;
        const float* bb = aabbB + 6 * b;
        const float m = 1e-4f;
        if (ba[0] > bb[3] + m || bb[0] > ba[3] + m || ba[1] > bb[4] + m || bb[1] > ba[4] + m || ba[2] > bb[5] + m ||
            bb[2] > ba[5] + m)
            continue;
        atomicAdd(&counters[2], 1ULL);
        const float* ca = cornersA + 24 * a;
        const float* cb = cornersB + 24 * b;
        if (mode == BF_IOU_ANALYTIC) {
            double v;
            if (bf_analytic_iou(ca, cb, &v)) {
                atomicAdd(&counters[4], 1ULL);
                if (iou) iou[p] = v;
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
                continue;
            }
        }
        const unsigned long long slot = atomicAdd(&counters[0], 1ULL);      // candidate: gate + counts in bf_count_kernel
        if (slot < (unsigned long long)work_cap) { work[slot].a = a; work[slot].b = b; }
        else atomicExch(&counters[5], 1ULL);
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
        __syncthreads();
        if (tid < 48) { s_pl[0][tid] = planesA[48 * (size_t)a + tid]; s_pl[1][tid] = planesB[48 * (size_t)b + tid]; }
        if (tid >= 64 && tid < 88) { s_c[0][tid - 64] = cornersA[24 * (size_t)a + tid - 64]; s_c[1][tid - 64] = cornersB[24 * (size_t)b + tid - 64]; }
        if (tid >= 96 && tid < 96 + 3 * BF_NS) {
            const int ax = (tid - 96) / BF_NS, i = (tid - 96) - ax * BF_NS;
            const float lo = fminf(aabbA[6 * a + ax], aabbB[6 * b + ax]), hi = fmaxf(aabbA[6 * a + 3 + ax], aabbB[6 * b + 3 + ax]);   // instances.py:581-582
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
    if ((rc = bf_scratch(h, BF_SCRATCH_AABB_A, sizeof(float) * 6 * (size_t)M, &p))) return rc; bbA = (float*)p;
    {
        const int g = bf_blocks(6LL * M, 96);
        bf_planes_kernel<<<(dev_sized && g > stride_grid) ? stride_grid : g, 96, 0, st>>>(cornersA, Md, plA, bbA);
    }
    BF_LAUNCH_CHECK(h, "bf_planes_kernel");
    if (cornersB == cornersA && Nd.host == Md.host && Nd.dev == Md.dev) { plB = plA; bbB = bbA; }
    else {
        if ((rc = bf_scratch(h, BF_SCRATCH_PLANES_B, sizeof(double) * 48 * (size_t)N, &p))) return rc; plB = (double*)p;
        if ((rc = bf_scratch(h, BF_SCRATCH_AABB_B, sizeof(float) * 6 * (size_t)N, &p))) return rc; bbB = (float*)p;
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
        // fixed, machine-filling grid when the sizes live on the device; exact grid otherwise (capped: the kernel strides)
        long long g = (total + 127) / 128;
        const long long gcap = (long long)h->sm_count * (dev_sized ? 16 : 64);
        if (g > gcap) g = gcap;
        if (g < 1) g = 1;
        bf_pairs_kernel<<<(unsigned)g, 128, 0, st>>>(cornersA, bbA, plA, Md, cornersB, bbB, plB, Nd, triangle, a_off, mode,
                                                     iou, counts, work, (int)cap, counters, thr, rank, mask, rowany, edges, edge_cap);
    }
    BF_LAUNCH_CHECK(h, "bf_pairs_kernel");
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
