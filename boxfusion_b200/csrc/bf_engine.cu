// Device-resident map store: the per-keyframe state that demo.py keeps in Python
// containers (all_pred_box, per_frame_ins, BoxManager.fusion_list / fusion_flag / already_fusion) stays in HBM across
// keyframes, so that a fusion step needs one H2D (the detections) and one 32-byte D2H.  The arithmetic is the same
// as the reference-shaped entry points (bf_transform2world, bf_project_boxes, bf_nms3d, bf_corr2d, bf_refine); the
// kernels here replace the host bookkeeping around them:
//   bf_engine_ingest   demo.py:216-221, 243/248, 253-254   lift + project + append to map and per-frame store
//   bf_engine_corr     instances.py:411-490, box_manager.py:90-129   small-object correspondence incl. its sequential tail
//   bf_engine_compact  `all_pred_box[keep_idx]`, `box_manager.update(keep_idx)` (demo.py:292, 325-327)
//   bf_engine_select   box_fusion.py:631-635   which map boxes get refined -> CSR for bf_refine
//   bf_engine_apply    box_fusion.py:716-724   write fused rows, flags, already_fusion
#include "bf_common.cuh"

// bf_map_buffers / bf_store_buffers / bf_fused_table: see include/boxfusion_b200.h

__device__ __forceinline__ float e_dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// ---------------------------------------------------------------------------------------------------------------
// ingest: one thread per detection.  in[] = packed keyframe: tensor_cam[n,6] R_cam[n,9] scores[n] box2d[n,4]
// projxy[n,2] pose[16] pose_inv[16]  (pose_inv = torch.linalg.inv(pose), computed by the host like the reference)
__global__ void bf_engine_ingest_kernel(const float* __restrict__ in, int n, float fx, float fy, float cx, float cy,
                                        float W, float H, int frame_id, int box_count, int N, int M, int D,
                                        bf_map_buffers mp, bf_store_buffers st, int32_t* __restrict__ fflag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* tc = in + 6 * i;
    const float* rc = in + 6 * n + 9 * i;
    const float score = in[15 * n + i];
    const float* b2 = in + 16 * n + 4 * i;
    const float* pxy = in + 20 * n + 2 * i;
    const float* pose = in + 22 * n;
    const float* pinv = pose + 16;
    float t[6], r[9];
    // boxes.py:825-833
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float p0 = pose[4 * a], p1 = pose[4 * a + 1], p2 = pose[4 * a + 2];
        t[a] = __fadd_rn(e_dot3(p0, tc[0], p1, tc[1], p2, tc[2]), pose[4 * a + 3]);
#pragma unroll
        for (int j = 0; j < 3; ++j) r[3 * a + j] = e_dot3(p0, rc[j], p1, rc[3 + j], p2, rc[6 + j]);
    }
    t[3] = tc[3]; t[4] = tc[4]; t[5] = tc[5];
    // boxes.py:725-778 + instances.py:333-369
    const float hl = t[3] * 0.5f, hh = t[4] * 0.5f, hw = t[5] * 0.5f;
    float uv[16];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const float vx = ((v & 1) ^ ((v >> 1) & 1)) ? hl : -hl, vy = (v & 2) ? hh : -hh, vz = (v & 4) ? hw : -hw;
        const float x = __fadd_rn(e_dot3(r[0], vx, r[1], vy, r[2], vz), t[0]);
        const float y = __fadd_rn(e_dot3(r[3], vx, r[4], vy, r[5], vz), t[1]);
        const float z = __fadd_rn(e_dot3(r[6], vx, r[7], vy, r[8], vz), t[2]);
        const float X = __fadd_rn(e_dot3(pinv[0], x, pinv[1], y, pinv[2], z), pinv[3]);
        const float Y = __fadd_rn(e_dot3(pinv[4], x, pinv[5], y, pinv[6], z), pinv[7]);
        const float Z = __fadd_rn(e_dot3(pinv[8], x, pinv[9], y, pinv[10], z), pinv[11]);
        uv[2 * v] = fminf(fmaxf(__fadd_rn(__fdiv_rn(__fmul_rn(fx, X), Z), cx), 0.f), W);
        uv[2 * v + 1] = fminf(fmaxf(__fadd_rn(__fdiv_rn(__fmul_rn(fy, Y), Z), cy), 0.f), H);
    }
    const size_t m = (size_t)N + i, s = (size_t)M + i;
#pragma unroll
    for (int k = 0; k < 6; ++k) { mp.tensor[6 * m + k] = t[k]; st.tensor[6 * s + k] = t[k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) { mp.R[9 * m + k] = r[k]; st.R[9 * s + k] = r[k]; }
#pragma unroll
    for (int k = 0; k < 16; ++k) { mp.uv[16 * m + k] = uv[k]; st.uv[16 * s + k] = uv[k]; mp.pose[16 * m + k] = pose[k]; st.pose[16 * s + k] = pose[k]; }
    mp.scores[m] = score; st.scores[s] = score;
#pragma unroll
    for (int k = 0; k < 4; ++k) mp.box2d[4 * m + k] = b2[k];
    mp.projxy[2 * m] = pxy[0]; mp.projxy[2 * m + 1] = pxy[1];
    mp.valid[m] = 0.f;
    mp.init_id[m] = box_count + i;                        // demo.py:218
    mp.frame_id[m] = frame_id;
    mp.fl[m * BF_FUSION_CAP] = M + i;                     // box_manager.py:24-28
    mp.flen[m] = 1;
    fflag[D + i] = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// pose predicate of record_corr (box_manager.py:100-102): no centre term
__device__ __forceinline__ bool e_views_differ(const float* __restrict__ p1, const float* __restrict__ p2, float tgap, float rgap) {
    const float dx = p2[3] - p1[3], dy = p2[7] - p1[7], dz = p2[11] - p1[11];
    const float baseline = sqrtf(dx * dx + dy * dy + dz * dz);
    float tr = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) tr += p2[4 * r] * p1[4 * r] + p2[4 * r + 1] * p1[4 * r + 1] + p2[4 * r + 2] * p1[4 * r + 2];
    const float c = fminf(fmaxf((tr - 1.f) * 0.5f, -1.f), 1.f);
    const float angle = acosf(c) * 180.f / 3.14159265358979323846f;
    return angle > rgap || baseline > tgap;
}

__device__ __forceinline__ void e_sorted_insert(int32_t* list, int& len, int32_t val) {
    int k = len;
    while (k > 0 && list[k - 1] > val) { list[k] = list[k - 1]; --k; }
    list[k] = val;
    ++len;
}

// BoxManager.record_corr(cur, [idx]) on keep FLAGS (box_manager.py:90-129)
__device__ void e_record_corr(int cur, int idx, const bf_map_buffers& mp, const float* __restrict__ sposes,
                              int32_t* __restrict__ fflag, int32_t* __restrict__ keep, float tgap, float rgap,
                              int32_t* __restrict__ status) {
    int32_t* lc = mp.fl + (size_t)cur * BF_FUSION_CAP;
    const int32_t* li = mp.fl + (size_t)idx * BF_FUSION_CAP;
    int len_c = mp.flen[cur];
    const int len_i = mp.flen[idx];
    if (len_i == 1) {
        const float* pi = sposes + 16 * (size_t)mp.init_id[idx];
        int cnt = 0;
        for (int k = 0; k < len_c; ++k) cnt += e_views_differ(sposes + 16 * (size_t)lc[k], pi, tgap, rgap);
        if (cnt == len_c && len_c < 5) {
            if (len_c + 1 > BF_FUSION_CAP) status[0] = BF_ERR_CAPACITY; else e_sorted_insert(lc, len_c, mp.init_id[idx]);
        }
    } else {
        const float* pc = sposes + 16 * (size_t)mp.init_id[cur];
        int cnt = 0;
        for (int k = 0; k < len_i; ++k) cnt += e_views_differ(sposes + 16 * (size_t)li[k], pc, tgap, rgap);
        if (cnt == len_i && len_i < 5) {
            if (len_c + len_i > BF_FUSION_CAP) status[0] = BF_ERR_CAPACITY;
            else for (int k = 0; k < len_i; ++k) e_sorted_insert(lc, len_c, li[k]);
        } else if (keep[cur]) { keep[cur] = 0; keep[idx] = 1; }          // keep[keep == cur_id] = idx
        if (fflag[idx] == 1) fflag[cur] = 1;
    }
    mp.flen[cur] = len_c;
}

// correspondence_association (instances.py:411-490) for one keyframe, one CTA.
//   keep/success: flags over the N_glo + n boxes after nms_3d (keep is edited in place)
//   pinv: np.linalg.inv(pose) of the current keyframe (float32, host)
//   info[0] = 1 if any new box survived nms_3d (demo.py:269), else 0
#define BF_ECORR_THREADS 256
__global__ void __launch_bounds__(BF_ECORR_THREADS)
bf_engine_corr_kernel(bf_map_buffers mp, const float* __restrict__ sposes, int32_t* __restrict__ fflag, int N_glo, int n,
                      int32_t* __restrict__ keep, const int32_t* __restrict__ success,
                      const float* __restrict__ pinv, double fx, double fy, double cx, double cy, double W, double H,
                      float small_size, float small_plus, double threshold, float tgap, float rgap, double* __restrict__ boxes2d,
                      int32_t* __restrict__ glo_keep_snapshot, int32_t* __restrict__ info, int32_t* __restrict__ status) {
    __shared__ int s_any_new, s_any_small;
    __shared__ double s_bv[BF_ECORR_THREADS / 32];
    __shared__ int s_bi[BF_ECORR_THREADS / 32];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_any_new = 0; s_any_small = 0; }
    __syncthreads();
    // valid_num += 1 for every head that suppressed something (instances.py:72-73)
    for (int i = tid; i < N_glo + n; i += T) if (success[i]) mp.valid[i] += 1.f;
    for (int j = tid; j < n; j += T) {
        if (keep[N_glo + j]) {
            s_any_new = 1;
            const float* d = mp.tensor + 6 * (size_t)(N_glo + j) + 3;
            if (!(fmaxf(d[0], fmaxf(d[1], d[2])) > small_size) && !success[N_glo + j]) s_any_small = 1;
        }
    }
    __syncthreads();
    if (tid == 0) info[0] = s_any_new;
    if (!s_any_new || !s_any_small || N_glo == 0) return;
    // snapshot of the kept map boxes (global_keep_idx, :424) and their clipped 2-D boxes in the current view (:670-717)
    for (int g = tid; g < N_glo; g += T) {
        glo_keep_snapshot[g] = keep[g];
        if (!keep[g]) continue;
        const float hl = mp.tensor[6 * (size_t)g + 3] * 0.5f, hh = mp.tensor[6 * (size_t)g + 4] * 0.5f, hw = mp.tensor[6 * (size_t)g + 5] * 0.5f;
        const float* r = mp.R + 9 * (size_t)g;
        const float* t = mp.tensor + 6 * (size_t)g;
        bool any_valid = false, any_z = false;
        double umin = 0, vmin = 0, umax = 0, vmax = 0;
        for (int v = 0; v < 8; ++v) {
            const float vx = ((v & 1) ^ ((v >> 1) & 1)) ? hl : -hl, vy = (v & 2) ? hh : -hh, vz = (v & 4) ? hw : -hw;
            const double x = __fadd_rn(e_dot3(r[0], vx, r[1], vy, r[2], vz), t[0]);
            const double y = __fadd_rn(e_dot3(r[3], vx, r[4], vy, r[5], vz), t[1]);
            const double z = __fadd_rn(e_dot3(r[6], vx, r[7], vy, r[8], vz), t[2]);
            const double X = x * (double)pinv[0] + y * (double)pinv[1] + z * (double)pinv[2] + (double)pinv[3];
            const double Y = x * (double)pinv[4] + y * (double)pinv[5] + z * (double)pinv[6] + (double)pinv[7];
            const double Z = x * (double)pinv[8] + y * (double)pinv[9] + z * (double)pinv[10] + (double)pinv[11];
            const double u = (fx * X / Z) + cx, vv = (fy * Y / Z) + cy;
            any_valid |= (Z > 0) && (u > 0) && (u < W) && (vv > 0) && (vv < H);
            if (Z > 0 && Z < 8) {
                const double uc = fmin(fmax(u, 0.0), W), vc = fmin(fmax(vv, 0.0), H);
                if (!any_z) { umin = umax = uc; vmin = vmax = vc; any_z = true; }
                else { umin = fmin(umin, uc); umax = fmax(umax, uc); vmin = fmin(vmin, vc); vmax = fmax(vmax, vc); }
            }
        }
        const bool ok = any_valid && any_z;
        boxes2d[4 * (size_t)g] = ok ? umin : 0.0; boxes2d[4 * (size_t)g + 1] = ok ? vmin : 0.0;
        boxes2d[4 * (size_t)g + 2] = ok ? umax : 0.0; boxes2d[4 * (size_t)g + 3] = ok ? vmax : 0.0;
    }
    __syncthreads();
    // small new boxes in index order; scoring in parallel, decision by thread 0 (:446-483)
    for (int j = 0; j < n; ++j) {
        const int cur_new = N_glo + j;
        const float* dn = mp.tensor + 6 * (size_t)cur_new + 3;
        // membership in cur_keep_idx / cur_success_nms refers to the state right after nms_3d (:428-435): keep flags of
        // NEW boxes are only cleared by this loop for the box itself or set back by a swap, so test the nms result
        const bool cand = glo_keep_snapshot[N_glo + j] && !(fmaxf(dn[0], fmaxf(dn[1], dn[2])) > small_size) && !success[cur_new];
        if (!cand) continue;                                               // uniform: all threads read the same flags
        const float* a = mp.box2d + 4 * (size_t)cur_new;
        const double ax0 = a[0], ay0 = a[1], ax1 = a[2], ay1 = a[3];
        const double areaA = (ax1 - ax0) * (ay1 - ay0);
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int g = tid; g < N_glo; g += T) {
            if (!glo_keep_snapshot[g]) continue;
            const double bx0 = boxes2d[4 * (size_t)g], by0 = boxes2d[4 * (size_t)g + 1], bx1 = boxes2d[4 * (size_t)g + 2], by1 = boxes2d[4 * (size_t)g + 3];
            const double areaB = (bx1 - bx0) * (by1 - by0);
            const double iw = fmax(0.0, fmin(ax1, bx1) - fmax(ax0, bx0)), ih = fmax(0.0, fmin(ay1, by1) - fmax(ay0, by0));
            const double inter = iw * ih;
            double v = inter / (areaA + areaB - inter + 1e-6);
            const float* dg = mp.tensor + 6 * (size_t)g + 3;
            v = v * ((fmaxf(dg[0], fmaxf(dg[1], dg[2])) < small_plus) ? 1.0 : 0.0);             // (:460-461)
            if (v > bv) { bv = v; bi = g; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_bv[warp] = bv; s_bi[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < T / 32; ++w) if (s_bv[w] > bv || (s_bv[w] == bv && s_bi[w] < bi)) { bv = s_bv[w]; bi = s_bi[w]; }
            if (bi != 0x7fffffff && bv > threshold) {
                const int cidx = bi;
                if (mp.scores[cidx] < mp.scores[cur_new]) {                // the new box wins (:471-477)
                    keep[cidx] = 0;
                    mp.valid[cur_new] += 1.f;
                    e_record_corr(cur_new, cidx, mp, sposes, fflag, keep, tgap, rgap, status);
                } else {                                                  // the old box wins (:478-483)
                    keep[cur_new] = 0;
                    mp.valid[cidx] += 1.f;
                    e_record_corr(cidx, cur_new, mp, sposes, fflag, keep, tgap, rgap, status);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// compaction: one CTA scans the keep flags -> src[] (old index of every new row) and info[1] = N_new;
// a second kernel gathers every field.
#define BF_ESCAN_THREADS 1024
__global__ void __launch_bounds__(BF_ESCAN_THREADS)
bf_engine_scan_kernel(const int32_t* __restrict__ keep, int N, int32_t* __restrict__ src, int32_t* __restrict__ info) {
    __shared__ int s_w[BF_ESCAN_THREADS / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (N + BF_ESCAN_THREADS - 1) / BF_ESCAN_THREADS;
    const int lo = tid * per, hi = min(N, lo + per);
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += keep[i] ? 1 : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = s_w[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        s_w[lane] = v;
        if (lane == 31) s_total = v;
    }
    __syncthreads();
    int pos = incl - cnt + (warp ? s_w[warp - 1] : 0);
    for (int i = lo; i < hi; ++i) if (keep[i]) src[pos++] = i;
    if (tid == 0) info[1] = s_total;
}

__global__ void bf_engine_gather_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ info, bf_map_buffers a,
                                        bf_map_buffers b) {
    const int n_new = info[1];
    const int k = blockIdx.x;                        // new row
    if (k >= n_new) return;
    const size_t s = (size_t)src[k], d = (size_t)k;
    const int t = threadIdx.x;                       // 64 threads: field elements
    if (t < 6) b.tensor[6 * d + t] = a.tensor[6 * s + t];
    if (t < 9) b.R[9 * d + t] = a.R[9 * s + t];
    if (t < 16) { b.pose[16 * d + t] = a.pose[16 * s + t]; b.uv[16 * d + t] = a.uv[16 * s + t]; }
    if (t < 4) b.box2d[4 * d + t] = a.box2d[4 * s + t];
    if (t < 2) b.projxy[2 * d + t] = a.projxy[2 * s + t];
    if (t < BF_FUSION_CAP) b.fl[d * BF_FUSION_CAP + t] = a.fl[s * BF_FUSION_CAP + t];
    if (t == 0) {
        b.scores[d] = a.scores[s]; b.valid[d] = a.valid[s]; b.init_id[d] = a.init_id[s]; b.frame_id[d] = a.frame_id[s];
        b.flen[d] = a.flen[s];
    }
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long e_list_hash(const int32_t* l, int len) {
    unsigned long long h = 1469598103934665603ULL ^ (unsigned long long)len;
    for (int k = 0; k < len; ++k) { h ^= (unsigned long long)(unsigned)l[k]; h *= 1099511628211ULL; }
    return h;
}

__device__ __forceinline__ bool e_in_fused(const bf_fused_table& ft, int F, const int32_t* l, int len, unsigned long long h) {
    for (int f = 0; f < F; ++f) {
        if (ft.hash[f] != h || ft.len[f] != len) continue;
        const int32_t* q = ft.lists + (size_t)f * BF_FUSION_CAP;
        bool same = true;
        for (int k = 0; k < len; ++k) same &= (q[k] == l[k]);
        if (same) return true;
    }
    return false;
}

// which map boxes are refined this keyframe (box_fusion.py:631-635) -> todo[], CSR offsets / view_index;
// info[2] = B, info[3] = sum V, info[4] = max V.  One CTA.
__global__ void __launch_bounds__(BF_ESCAN_THREADS)
bf_engine_select_kernel(bf_map_buffers mp, bf_fused_table ft, int32_t* __restrict__ info, int32_t* __restrict__ todo,
                        int32_t* __restrict__ offsets, int32_t* __restrict__ view_index, int max_views) {
    __shared__ int s_w[BF_ESCAN_THREADS / 32];
    __shared__ int s_wv[BF_ESCAN_THREADS / 32];
    __shared__ int s_maxv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = info[1], F = ft.count[0];
    const bool active = info[0] != 0;                         // demo.py:269: fusion only when a new box survived
    if (tid == 0) s_maxv = 0;
    __syncthreads();
    const int per = (N + BF_ESCAN_THREADS - 1) / BF_ESCAN_THREADS;
    const int lo = tid * per, hi = min(N, lo + per);
    int cnt = 0, vsum = 0, vmax = 0;
    unsigned long long picks = 0ull;                          // per <= 64 rows per thread (N <= 65536)
    for (int i = lo; i < hi && active; ++i) {
        const int len = mp.flen[i];
        if (len < 3) continue;
        const int32_t* l = mp.fl + (size_t)i * BF_FUSION_CAP;
        if (e_in_fused(ft, F, l, len, e_list_hash(l, len))) continue;
        if (i - lo < 64) picks |= 1ull << (i - lo);
        ++cnt; vsum += len; vmax = max(vmax, len);
    }
    int incl = cnt, vincl = vsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o), u = __shfl_up_sync(0xffffffffu, vincl, o);
        if (lane >= o) { incl += v; vincl += u; }
    }
    if (lane == 31) { s_w[warp] = incl; s_wv[warp] = vincl; }
    atomicMax(&s_maxv, vmax);
    __syncthreads();
    if (warp == 0) {
        int v = s_w[lane], u = s_wv[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v2 = __shfl_up_sync(0xffffffffu, v, o), u2 = __shfl_up_sync(0xffffffffu, u, o);
            if (lane >= o) { v += v2; u += u2; }
        }
        s_w[lane] = v; s_wv[lane] = u;
    }
    __syncthreads();
    int pos = incl - cnt + (warp ? s_w[warp - 1] : 0);
    int vpos = vincl - vsum + (warp ? s_wv[warp - 1] : 0);
    for (int i = lo; i < hi; ++i) {
        if (!((picks >> (i - lo)) & 1ull)) continue;
        const int len = mp.flen[i];
        todo[pos] = i; offsets[pos] = vpos;
        for (int k = 0; k < len; ++k) view_index[vpos + k] = mp.fl[(size_t)i * BF_FUSION_CAP + k];
        ++pos; vpos += len;
    }
    if (tid == BF_ESCAN_THREADS - 1) {
        const int B = s_w[BF_ESCAN_THREADS / 32 - 1], SV = s_wv[BF_ESCAN_THREADS / 32 - 1];
        offsets[B] = SV;
        info[2] = B; info[3] = SV; info[4] = s_maxv;
        if (s_maxv > max_views) info[5] = BF_ERR_CAPACITY;
    }
}

// write back fused boxes (box_fusion.py:716-724), sequentially in map order like the reference's loop
__global__ void bf_engine_apply_kernel(bf_map_buffers mp, bf_fused_table ft, int32_t* __restrict__ fflag,
                                       const int32_t* __restrict__ info, const int32_t* __restrict__ todo,
                                       const float* __restrict__ out, const int32_t* __restrict__ upd, int32_t* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int B = info[2];
    int F = ft.count[0];
    for (int k = 0; k < B; ++k) {
        const int i = todo[k];
        const int len = mp.flen[i];
        const int32_t* l = mp.fl + (size_t)i * BF_FUSION_CAP;
        const unsigned long long h = e_list_hash(l, len);
        if (e_in_fused(ft, F, l, len, h)) continue;           // fused earlier in this very call (check_if_fusion, :634)
        if (!upd[k]) continue;
        for (int c = 0; c < 6; ++c) mp.tensor[6 * (size_t)i + c] = out[6 * (size_t)k + c];
        fflag[i] = 1;
        if (F >= ft.cap) { status[0] = BF_ERR_CAPACITY; continue; }
        for (int c = 0; c < len; ++c) ft.lists[(size_t)F * BF_FUSION_CAP + c] = l[c];
        ft.len[F] = len; ft.hash[F] = h;
        ++F;
    }
    ft.count[0] = F;
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" int bf_engine_ingest(bf_handle* h, const float* packed, int n, float fx, float fy, float cx, float cy, float W,
                                float H, int frame_id, int box_count, int N, int M, int D, const bf_map_buffers* mp,
                                const bf_store_buffers* st, int32_t* fflag, void* stream) {
    if (!h || !packed || !mp || !st || !fflag || n < 0) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_engine_ingest", "bad argument");
    if (n == 0) return BF_OK;
    bf_engine_ingest_kernel<<<bf_blocks(n, 64), 64, 0, (cudaStream_t)stream>>>(packed, n, fx, fy, cx, cy, W, H, frame_id,
                                                                           box_count, N, M, D, *mp, *st, fflag);
    BF_LAUNCH_CHECK(h, "bf_engine_ingest_kernel");
    return BF_OK;
}

extern "C" int bf_engine_corr(bf_handle* h, const bf_map_buffers* mp, const float* store_poses, int32_t* fflag, int N_glo,
                              int n, int32_t* keep, const int32_t* success, const float* pose_inv_np, float fx, float fy,
                              float cx, float cy, float W, float H, float small_size, float small_plus, double threshold,
                              float translation_gap, float rotation_gap, int32_t* info, int32_t* status, void* stream) {
    if (!h || !mp || !keep || !success || !info || !status || N_glo < 0 || n < 0)
        return bf_fail(h, BF_ERR_INVALID_ARG, "bf_engine_corr", "bad argument");
    void* p;
    int rc = bf_scratch(h, BF_SCRATCH_MISC, sizeof(double) * 4 * (size_t)(N_glo + 1) + sizeof(int32_t) * (size_t)(N_glo + n + 1), &p);
    if (rc) return rc;
    double* boxes2d = (double*)p;
    int32_t* snap = (int32_t*)(boxes2d + 4 * (size_t)(N_glo + 1));
    // the snapshot also covers the new boxes' keep flags as nms_3d left them
    BF_CUDA(h, cudaMemcpyAsync(snap + N_glo, keep + N_glo, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    bf_engine_corr_kernel<<<1, BF_ECORR_THREADS, 0, (cudaStream_t)stream>>>(*mp, store_poses, fflag, N_glo, n, keep, success,
                                                                          pose_inv_np, (double)fx, (double)fy, (double)cx,
                                                                          (double)cy, (double)W, (double)H, small_size, small_plus,
                                                                          threshold, translation_gap, rotation_gap, boxes2d, snap,
                                                                          info, status);
    BF_LAUNCH_CHECK(h, "bf_engine_corr_kernel");
    return BF_OK;
}

extern "C" int bf_engine_compact(bf_handle* h, const int32_t* keep, int N, const bf_map_buffers* from, const bf_map_buffers* to,
                                 int32_t* info, void* stream) {
    if (!h || !keep || !from || !to || !info || N < 0 || N > 65536) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_engine_compact", "bad argument");
    void* p;
    int rc = bf_scratch(h, BF_SCRATCH_RANK, sizeof(int32_t) * (size_t)(N + 1), &p);
    if (rc) return rc;
    bf_engine_scan_kernel<<<1, BF_ESCAN_THREADS, 0, (cudaStream_t)stream>>>(keep, N, (int32_t*)p, info);
    BF_LAUNCH_CHECK(h, "bf_engine_scan_kernel");
    if (N > 0) {
        bf_engine_gather_kernel<<<N, 64, 0, (cudaStream_t)stream>>>((const int32_t*)p, info, *from, *to);
        BF_LAUNCH_CHECK(h, "bf_engine_gather_kernel");
    }
    return BF_OK;
}

extern "C" int bf_engine_select(bf_handle* h, const bf_map_buffers* mp, const bf_fused_table* ft, int32_t* info, int32_t* todo,
                                int32_t* offsets, int32_t* view_index, void* stream) {
    if (!h || !mp || !ft || !info || !todo || !offsets || !view_index) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_engine_select", "bad argument");
    bf_engine_select_kernel<<<1, BF_ESCAN_THREADS, 0, (cudaStream_t)stream>>>(*mp, *ft, info, todo, offsets, view_index, BF_MAX_VIEWS);
    BF_LAUNCH_CHECK(h, "bf_engine_select_kernel");
    return BF_OK;
}

extern "C" int bf_engine_apply(bf_handle* h, const bf_map_buffers* mp, const bf_fused_table* ft, int32_t* fflag,
                               const int32_t* info, const int32_t* todo, const float* out, const int32_t* upd, int32_t* status,
                               void* stream) {
    if (!h || !mp || !ft || !fflag || !info || !todo || !out || !upd || !status) return bf_fail(h, BF_ERR_INVALID_ARG, "bf_engine_apply", "bad argument");
    bf_engine_apply_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*mp, *ft, fflag, info, todo, out, upd, status);
    BF_LAUNCH_CHECK(h, "bf_engine_apply_kernel");
    return BF_OK;
}
