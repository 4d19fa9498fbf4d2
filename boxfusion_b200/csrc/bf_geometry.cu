// Box geometry kernels: corner generation (A1), camera->world lift (A2), observation projection (A15),
// pose disparity (A8).  All float32 with the reference's rounding order made explicit through
// __fmul_rn/__fadd_rn (torch's CPU bmm on 3x3 operands rounds as ((a0*b0 + a1*b1) + a2*b2), no FMA),
// so results are bit-identical to the reference's CPU tensors.
#include "bf_internal.cuh"

__device__ __forceinline__ float dot3_seq(float a0, float b0, float a1, float b1, float a2, float b2) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// boxes.py:725-778.  One thread per box: 15 floats in, 24 (+3) out.
__global__ void bf_corners_kernel(const float* __restrict__ xyzlhw, const float* __restrict__ R, const bf_dimref Nd,
                                  float* __restrict__ corners, float* __restrict__ centers) {
    const int N = bf_dim(Nd);
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float* t = xyzlhw + 6 * n;
    const float* r = R + 9 * n;
    const float hl = t[3] * 0.5f, hh = t[4] * 0.5f, hw = t[5] * 0.5f;   // x/2 is exact
    float r9[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) r9[k] = r[k];
    const float tx = t[0], ty = t[1], tz = t[2];
    float sum[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const float vx = ((v & 1) ^ ((v >> 1) & 1)) ? hl : -hl;        // v1,v2,v5,v6 -> +l/2
        const float vy = (v & 2) ? hh : -hh;                            // v2,v3,v6,v7 -> +h/2
        const float vz = (v & 4) ? hw : -hw;                            // v4..v7      -> +w/2
        const float c0 = __fadd_rn(dot3_seq(r9[0], vx, r9[1], vy, r9[2], vz), tx);
        const float c1 = __fadd_rn(dot3_seq(r9[3], vx, r9[4], vy, r9[5], vz), ty);
        const float c2 = __fadd_rn(dot3_seq(r9[6], vx, r9[7], vy, r9[8], vz), tz);
        corners[24 * n + 3 * v + 0] = c0;
        corners[24 * n + 3 * v + 1] = c1;
        corners[24 * n + 3 * v + 2] = c2;
        sum[0] = v ? __fadd_rn(sum[0], c0) : c0;                        // np.mean(axis=1): sequential f32
        sum[1] = v ? __fadd_rn(sum[1], c1) : c1;
        sum[2] = v ? __fadd_rn(sum[2], c2) : c2;
    }
    if (centers) {
        centers[3 * n + 0] = sum[0] * 0.125f;
        centers[3 * n + 1] = sum[1] * 0.125f;
        centers[3 * n + 2] = sum[2] * 0.125f;
    }
    }
}

int bf_box_corners_run(bf_handle* h, const float* xyzlhw, const float* R, bf_dimref Nd, float* corners, float* centers, cudaStream_t st) {
    int g = bf_blocks(Nd.host, 128);
    if (g > h->sm_count * 4) g = h->sm_count * 4;
    if (g < 1) g = 1;
    bf_corners_kernel<<<g, 128, 0, st>>>(xyzlhw, R, Nd, corners, centers);
    BF_LAUNCH_CHECK(h, "bf_corners_kernel");
    return BF_OK;
}

extern "C" int bf_box_corners(bf_handle* h, const float* xyzlhw, const float* R, int N, float* corners,
                              float* centers, void* stream) {
    bf_device_guard guard(h);
    if (!h || N < 0 || (N > 0 && (!xyzlhw || !R || !corners))) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_box_corners", "bad argument");
    if (N == 0) return BF_OK;
    return bf_box_corners_run(h, xyzlhw, R, bf_dim_host(N), corners, centers, (cudaStream_t)stream);
}

// boxes.py:825-833, in place.
__global__ void bf_transform2world_kernel(float* __restrict__ xyzlhw, float* __restrict__ R,
                                          const float* __restrict__ poses, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* p = poses + 16 * n;
    float* t = xyzlhw + 6 * n;
    float* r = R + 9 * n;
    const float c0 = t[0], c1 = t[1], c2 = t[2];
    float rb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) rb[k] = r[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a0 = p[4 * i], a1 = p[4 * i + 1], a2 = p[4 * i + 2];
        t[i] = __fadd_rn(dot3_seq(a0, c0, a1, c1, a2, c2), p[4 * i + 3]);
#pragma unroll
        for (int j = 0; j < 3; ++j) r[3 * i + j] = dot3_seq(a0, rb[j], a1, rb[3 + j], a2, rb[6 + j]);
    }
}

extern "C" int bf_transform2world(bf_handle* h, float* xyzlhw, float* R, const float* poses, int N, void* stream) {
    bf_device_guard guard(h);
    if (!h || N < 0 || (N > 0 && (!xyzlhw || !R || !poses))) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_transform2world", "bad argument");
    if (N == 0) return BF_OK;
    bf_transform2world_kernel<<<bf_blocks(N, 128), 128, 0, (cudaStream_t)stream>>>(xyzlhw, R, poses, N);
    BF_LAUNCH_CHECK(h, "bf_transform2world_kernel");
    return BF_OK;
}

// instances.py:333-369.  One thread per corner.  pose_inv = inverse camera pose (world->camera).
__global__ void bf_project_kernel(const float* __restrict__ corners, const float* __restrict__ pose_inv, int N,
                                  float fx, float fy, float cx, float cy, float W, float H, float* __restrict__ uv) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * 8) return;
    const int n = idx >> 3;
    const float* p = pose_inv + 16 * n;
    const float x = corners[3 * idx], y = corners[3 * idx + 1], z = corners[3 * idx + 2];
    const float X = __fadd_rn(dot3_seq(p[0], x, p[1], y, p[2], z), p[3]);
    const float Y = __fadd_rn(dot3_seq(p[4], x, p[5], y, p[6], z), p[7]);
    const float Z = __fadd_rn(dot3_seq(p[8], x, p[9], y, p[10], z), p[11]);
    float u = __fadd_rn(__fdiv_rn(__fmul_rn(fx, X), Z), cx);
    float v = __fadd_rn(__fdiv_rn(__fmul_rn(fy, Y), Z), cy);
    u = fminf(fmaxf(u, 0.f), W);                                          // torch.clamp(u, 0, W)
    v = fminf(fmaxf(v, 0.f), H);
    uv[2 * idx] = u;
    uv[2 * idx + 1] = v;
}

extern "C" int bf_project_boxes(bf_handle* h, const float* corners, const float* pose_inv, int N, float fx, float fy,
                                float cx, float cy, float W, float H, float* uv, void* stream) {
    bf_device_guard guard(h);
    if (!h || N < 0 || (N > 0 && (!corners || !pose_inv || !uv))) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_project_boxes", "bad argument");
    if (N == 0) return BF_OK;
    bf_project_kernel<<<bf_blocks(8LL * N, 128), 128, 0, (cudaStream_t)stream>>>(corners, pose_inv, N, fx, fy, cx, cy, W, H, uv);
    BF_LAUNCH_CHECK(h, "bf_project_kernel");
    return BF_OK;
}

// The two calls demo.py:220-221 makes per keyframe - transform2world(cam_pose) and project_3d_boxes(K, H, W) - when every
// detection of the keyframe carries the same camera pose (it always does there: the pose is np.repeat-ed): the pose
// travels as a kernel parameter (no host->device copy of [n,16] floats) and corners + projection are one kernel.
// Same arithmetic, in the same order, as bf_transform2world_kernel / bf_corners_kernel / bf_project_kernel.
struct bf_pose16 { float m[16]; };

__global__ void bf_transform2world_pose_kernel(float* __restrict__ xyzlhw, float* __restrict__ R, const bf_pose16 pose, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* p = pose.m;
    float* t = xyzlhw + 6 * n;
    float* r = R + 9 * n;
    const float c0 = t[0], c1 = t[1], c2 = t[2];
    float rb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) rb[k] = r[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a0 = p[4 * i], a1 = p[4 * i + 1], a2 = p[4 * i + 2];
        t[i] = __fadd_rn(dot3_seq(a0, c0, a1, c1, a2, c2), p[4 * i + 3]);
#pragma unroll
        for (int j = 0; j < 3; ++j) r[3 * i + j] = dot3_seq(a0, rb[j], a1, rb[3 + j], a2, rb[6 + j]);
    }
}

extern "C" int bf_transform2world_pose(bf_handle* h, float* xyzlhw, float* R, const float* pose_host, int N, void* stream) {
    bf_device_guard guard(h);
    if (!h || N < 0 || !pose_host || (N > 0 && (!xyzlhw || !R))) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_transform2world_pose", "bad argument");
    if (N == 0) return BF_OK;
    bf_pose16 p;
    memcpy(p.m, pose_host, sizeof(p.m));
    bf_transform2world_pose_kernel<<<bf_blocks(N, 128), 128, 0, (cudaStream_t)stream>>>(xyzlhw, R, p, N);
    BF_LAUNCH_CHECK(h, "bf_transform2world_pose_kernel");
    return BF_OK;
}

__global__ void bf_project_boxes_pose_kernel(const float* __restrict__ xyzlhw, const float* __restrict__ R, int N, const bf_pose16 pinv,
                                             float fx, float fy, float cx, float cy, float W, float H, float* __restrict__ uv) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * 8) return;
    const int n = idx >> 3, v = idx & 7;
    const float* t = xyzlhw + 6 * n;
    const float* r = R + 9 * n;
    const float hl = t[3] * 0.5f, hh = t[4] * 0.5f, hw = t[5] * 0.5f;
    const float vx = ((v & 1) ^ ((v >> 1) & 1)) ? hl : -hl, vy = (v & 2) ? hh : -hh, vz = (v & 4) ? hw : -hw;
    const float x = __fadd_rn(dot3_seq(r[0], vx, r[1], vy, r[2], vz), t[0]);
    const float y = __fadd_rn(dot3_seq(r[3], vx, r[4], vy, r[5], vz), t[1]);
    const float z = __fadd_rn(dot3_seq(r[6], vx, r[7], vy, r[8], vz), t[2]);
    const float* p = pinv.m;
    const float X = __fadd_rn(dot3_seq(p[0], x, p[1], y, p[2], z), p[3]);
    const float Y = __fadd_rn(dot3_seq(p[4], x, p[5], y, p[6], z), p[7]);
    const float Z = __fadd_rn(dot3_seq(p[8], x, p[9], y, p[10], z), p[11]);
    float u = __fadd_rn(__fdiv_rn(__fmul_rn(fx, X), Z), cx);
    float w = __fadd_rn(__fdiv_rn(__fmul_rn(fy, Y), Z), cy);
    uv[2 * idx] = fminf(fmaxf(u, 0.f), W);
    uv[2 * idx + 1] = fminf(fmaxf(w, 0.f), H);
}

extern "C" int bf_project_boxes_pose(bf_handle* h, const float* xyzlhw, const float* R, int N, const float* pose_inv_host, float fx,
                                     float fy, float cx, float cy, float W, float H, float* uv, void* stream) {
    bf_device_guard guard(h);
    if (!h || N < 0 || !pose_inv_host || (N > 0 && (!xyzlhw || !R || !uv))) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_project_boxes_pose", "bad argument");
    if (N == 0) return BF_OK;
    bf_pose16 p;
    memcpy(p.m, pose_inv_host, sizeof(p.m));
    bf_project_boxes_pose_kernel<<<bf_blocks(8LL * N, 128), 128, 0, (cudaStream_t)stream>>>(xyzlhw, R, N, p, fx, fy, cx, cy, W, H, uv);
    BF_LAUNCH_CHECK(h, "bf_project_boxes_pose_kernel");
    return BF_OK;
}

// box_manager.py:168-186
__global__ void bf_pose_disparity_kernel(const float* __restrict__ poses, const int32_t* __restrict__ ia,
                                         const int32_t* __restrict__ ib, int n, float* __restrict__ baseline,
                                         float* __restrict__ angle) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p1 = poses + 16 * ia[i];
    const float* p2 = poses + 16 * ib[i];
    const float dx = p2[3] - p1[3], dy = p2[7] - p1[7], dz = p2[11] - p1[11];
    baseline[i] = sqrtf(dx * dx + dy * dy + dz * dz);
    float tr = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) tr += p2[4 * r] * p1[4 * r] + p2[4 * r + 1] * p1[4 * r + 1] + p2[4 * r + 2] * p1[4 * r + 2];
    const float c = fminf(fmaxf((tr - 1.f) * 0.5f, -1.f), 1.f);
    angle[i] = acosf(c) * 180.f / 3.14159265358979323846f;
}

extern "C" int bf_pose_disparity(bf_handle* h, const float* poses, const int32_t* ia, const int32_t* ib, int n,
                                 float* baseline, float* angle_deg, void* stream) {
    bf_device_guard guard(h);
    if (!h || n < 0 || (n > 0 && (!poses || !ia || !ib || !baseline || !angle_deg))) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_pose_disparity", "bad argument");
    if (n == 0) return BF_OK;
    bf_pose_disparity_kernel<<<bf_blocks(n, 128), 128, 0, (cudaStream_t)stream>>>(poses, ia, ib, n, baseline, angle_deg);
    BF_LAUNCH_CHECK(h, "bf_pose_disparity_kernel");
    return BF_OK;
}

// Detection pre-filters of demo.py:138-148 in one pass: score threshold, check_uv_bounds
// (box_manager.py:217-225), check_floor_mask (:227-237), check_large_mask (:239-245).  flags bit0 = score below
// threshold, bit1 = centre outside the shrunken image, bit2 = floor-like, bit3 = too large; keep[i] = (flags == 0).
__global__ void bf_detection_filter_kernel(const float* __restrict__ xyzlhw, const float* __restrict__ proj_xy,
                                           const float* __restrict__ scores, int n, float score_thresh, int use_uv,
                                           float gap_w, float gap_h, float W, float H, int use_floor, float ratio,
                                           int use_large, float size_max, int32_t* __restrict__ flags, int32_t* __restrict__ keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = 0;
    if (!(scores[i] >= score_thresh)) f |= 1;
    if (use_uv) {
        const float u = proj_xy[2 * i], v = proj_xy[2 * i + 1];
        if (!((u > gap_w) && (u < (W - gap_w)) && (v > gap_h) && (v < (H - gap_h)))) f |= 2;
    }
    const float a = xyzlhw[6 * i + 3], b = xyzlhw[6 * i + 4], c = xyzlhw[6 * i + 5];
    const float mx = fmaxf(a, fmaxf(b, c)), mn = fminf(a, fminf(b, c));
    const float second = fmaxf(fminf(a, b), fminf(fmaxf(a, b), c));            // median of three
    if (use_floor) {
        const float half = ratio * 0.5f;
        const bool m1 = __fdiv_rn(mx, mn) > ratio;
        const bool m2 = (__fdiv_rn(mx, mn) > half) && (__fdiv_rn(mx, second) > half) && (__fdiv_rn(second, mn) < 2.0f) &&
                        (second < 0.15f) && (mn < 0.15f);
        if (m1 || m2) f |= 4;
    }
    if (use_large && mx > size_max) f |= 8;
    flags[i] = f;
    keep[i] = (f == 0) ? 1 : 0;
}

extern "C" int bf_detection_filter(bf_handle* h, const float* xyzlhw, const float* proj_xy, const float* scores, int n,
                                   float score_thresh, int use_uv, double uv_ratio, float W, float H, int use_floor, float floor_ratio,
                                   int use_large, float size_max, int32_t* flags, int32_t* keep, void* stream) {
    bf_device_guard guard(h);
    if (!h || n < 0 || (n > 0 && (!xyzlhw || !proj_xy || !scores || !flags || !keep)))
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_detection_filter", "bad argument");
    if (n == 0) return BF_OK;
    // gap = int((1 - ratio) * size), evaluated in double like the Python expression (box_manager.py:218-219)
    const float gap_w = (float)(int)((1.0 - uv_ratio) * (double)W), gap_h = (float)(int)((1.0 - uv_ratio) * (double)H);
    bf_detection_filter_kernel<<<bf_blocks(n, 128), 128, 0, (cudaStream_t)stream>>>(xyzlhw, proj_xy, scores, n, score_thresh, use_uv,
                                                                                gap_w, gap_h, W, H, use_floor, floor_ratio, use_large,
                                                                                size_max, flags, keep);
    BF_LAUNCH_CHECK(h, "bf_detection_filter_kernel");
    return BF_OK;
}
