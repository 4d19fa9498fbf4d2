// Device-resident fusion engine: the per-keyframe state that demo.py keeps in Python
// containers (all_pred_box, per_frame_ins, BoxManager.fusion_list / fusion_flag / already_fusion) stays in HBM across
// keyframes and ONE call (bf_engine_step) runs a whole keyframe of demo.py:200-327.
//
// Design (round 2).  Nothing that varies per keyframe is a kernel parameter: the row counts, the detection count, the
// intrinsics and poses live in device memory (bf_engine_state + the packed keyframe the step copies in), every kernel
// reads them itself and runs on a fixed launch shape (grid-stride loops / persistent clusters).  The launch sequence is
// therefore identical for every keyframe and is captured once per engine into three CUDA graphs:
//   phase a  ingest (demo.py:216-221, 243/248, 253-254)  ->  corners -> score order -> planes -> pairs -> gate+counts
//            -> greedy matching with record()              (Instances3D.spatial_association, demo.py:262)
//   phase b  correspondence association incl. record_corr  (demo.py:273-289)
//   phase c  all_pred_box[keep_idx] + box_manager.update (demo.py:292), check_valid_num (:297-298), selection of the
//            boxes to fuse (box_fusion.py:631-635), bf_refine, write-back (:716-724), counters for the next keyframe.
// bf_engine_step(.., phases = 7) replays all three back to back (three graph launches, no host decision in between);
// the reference-shaped API replays them one by one because spatial_association / correspondence_association hand the
// keep indices back to the caller.  The arithmetic is that of the stand-alone entries (same device functions).
#include "bf_internal.cuh"
#include "bf_record.cuh"

struct bf_engine {
    bf_handle* h;                     // private: its scratch is frozen once the graphs exist
    bf_engine_cfg cfg;
    bf_engine_buffers bufs;
    // engine-owned device memory
    float* in_dev;                    // packed keyframe (header + rows)
    bf_engine_state* state;
    int32_t *keep, *success, *todo, *offsets, *view_index, *order, *rank, *src, *snap, *upd, *its;
    float *corners, *centers, *out;
    double* boxes2d;
    // host staging ring for the keyframe copy
    enum { RING = 8 };
    float* stage[RING];
    cudaEvent_t stage_evt[RING];
    int stage_next;
    size_t in_floats;
    // graphs
    cudaStream_t cap_stream;
    cudaGraphExec_t exec[9];          // one per phase bit, [7] the whole keyframe, [8] everything after the correspondence phase
    int have_graph;
    int launches[9];
    // run-ahead (the reference-shaped API): rollback snapshot of the state right after the NMS phase, pinned flag buffers
    uint32_t* snapshot;               // [ncap * SNAP_ROW + 2 * ncap + 64] words, allocated on first use
    int32_t* hflags[2];               // pinned host: keep[ncap], success[ncap], state[32]
    cudaEvent_t flag_evt[2];
    cudaGraphExec_t ra_exec[2];       // [0] NMS phase + flags to slot 0, [1] snapshot + correspondence phase + flags to slot 1
    char err[512];
};

static int e_fail(bf_engine* e, int code, const char* what, const char* detail) {
    if (e) snprintf(e->err, sizeof(e->err), "%s: %s", what, detail ? detail : "");
    return code;
}
#define E_CUDA(e, expr)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) return e_fail((e), BF_ERR_CUDA, #expr, cudaGetErrorString(e__)); \
    } while (0)

// everything the engine kernels need, by value (kernel parameters never change -> graph replay)
struct bf_engine_ctx {
    bf_map_buffers map[2];
    bf_store_buffers store;
    int32_t* fflag;
    bf_fused_table fused;
    const float* in;
    bf_engine_state* st;
    int32_t *keep, *success, *todo, *offsets, *view_index, *src, *snap, *upd;
    const float* out;
    double* boxes2d;
    int ncap, mcap, max_det;
    float tgap, rgap, small_size, small_plus;
    double small_threshold;
    int check_valid, gap, use_fusion;
};

__device__ __forceinline__ float e_dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// ---------------------------------------------------------------------------------------------------------------
// phase a, kernel 1 - ingest: one thread per detection.  in[] = header (BF_KF_HEADER words) then tensor_cam[n,6] R_cam[n,9]
// scores[n] box2d[n,4] projxy[n,2].  Appends rows [N, N+n) of the current map and [M, M+n) of the store; thread 0 derives the
// keyframe's sizes for every later kernel.
// sizes of the keyframe in flight, derived once by thread 0 of the ingest kernel for every later kernel
__device__ __forceinline__ int e_begin_keyframe(const bf_engine_ctx& c, int i, int& N, int& M) {
    bf_engine_state* st = c.st;
    int n = __float_as_int(c.in[0]);
    N = st->N; M = st->M;
    bool fits = true;
    if (n < 0 || n > c.max_det || N + n > c.ncap || M + n > c.mcap) { fits = false; n = 0; }
    if (i == 0) {
        if (!fits) st->status[4] = BF_ERR_CAPACITY;
        st->n = n; st->Nall = N + n; st->first = (N == 0) ? 1 : 0;
        st->Nnms = (N == 0) ? 0 : N + n;
        st->any_new = 0; st->B = 0; st->SV = 0; st->maxV = 0;
        st->Nnew = N + n;
        st->pad[1] = 0;
    }
    return n;
}

__global__ void bf_engine_ingest_kernel(bf_engine_ctx c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float* in = c.in;
    const int frame_id = __float_as_int(in[1]);
    int N, M;
    const int n = e_begin_keyframe(c, i, N, M);
    if (i >= n) return;
    const bf_map_buffers& mp = c.map[0];
    const bf_store_buffers& sb = c.store;
    const float fx = in[2], fy = in[3], cx = in[4], cy = in[5], W = in[6], H = in[7];
    const float* pose = in + 8;
    const float* pinv = in + 24;
    const float* rows = in + BF_KF_HEADER;
    const float* tc = rows + 6 * i;
    const float* rc = rows + 6 * n + 9 * i;
    const float score = rows[15 * n + i];
    const float* b2 = rows + 16 * n + 4 * i;
    const float* pxy = rows + 20 * n + 2 * i;
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
    for (int k = 0; k < 6; ++k) { mp.tensor[6 * m + k] = t[k]; sb.tensor[6 * s + k] = t[k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) { mp.R[9 * m + k] = r[k]; sb.R[9 * s + k] = r[k]; }
#pragma unroll
    for (int k = 0; k < 16; ++k) { mp.uv[16 * m + k] = uv[k]; sb.uv[16 * s + k] = uv[k]; mp.pose[16 * m + k] = pose[k]; sb.pose[16 * s + k] = pose[k]; }
    mp.scores[m] = score; sb.scores[s] = score;
#pragma unroll
    for (int k = 0; k < 4; ++k) mp.box2d[4 * m + k] = b2[k];
    mp.projxy[2 * m] = pxy[0]; mp.projxy[2 * m + 1] = pxy[1];
    mp.valid[m] = 0.f;
    mp.init_id[m] = M + i;                                // demo.py:218 (box_count = rows of the store)
    mp.frame_id[m] = frame_id;                            // demo.py:217
    mp.fl[m * BF_FUSION_CAP] = M + i;                     // box_manager.py:24-28
    mp.flen[m] = 1;
    c.fflag[M + i] = 0;
}

// The same append for detections that were lifted and projected already (the reference-shaped API: the caller ran
// transform2world / project_3d_boxes itself, demo.py:220-221): plain row copies from the caller's tensors.
struct bf_world_rows { const float *tensor, *R, *scores, *box2d, *projxy, *uv; };
__global__ void bf_engine_ingest_world_kernel(bf_engine_ctx c, bf_world_rows w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float* in = c.in;
    const int frame_id = __float_as_int(in[1]);
    int N, M;
    const int n = e_begin_keyframe(c, i, N, M);
    if (i >= n) return;
    const bf_map_buffers& mp = c.map[0];
    const bf_store_buffers& sb = c.store;
    const float* pose = in + 8;
    const size_t m = (size_t)N + i, s = (size_t)M + i;
#pragma unroll
    for (int k = 0; k < 6; ++k) { const float v = w.tensor[6 * (size_t)i + k]; mp.tensor[6 * m + k] = v; sb.tensor[6 * s + k] = v; }
#pragma unroll
    for (int k = 0; k < 9; ++k) { const float v = w.R[9 * (size_t)i + k]; mp.R[9 * m + k] = v; sb.R[9 * s + k] = v; }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float v = w.uv[16 * (size_t)i + k];
        mp.uv[16 * m + k] = v; sb.uv[16 * s + k] = v; mp.pose[16 * m + k] = pose[k]; sb.pose[16 * s + k] = pose[k];
    }
    const float score = w.scores[i];
    mp.scores[m] = score; sb.scores[s] = score;
#pragma unroll
    for (int k = 0; k < 4; ++k) mp.box2d[4 * m + k] = w.box2d[4 * (size_t)i + k];
    mp.projxy[2 * m] = w.projxy[2 * (size_t)i]; mp.projxy[2 * m + 1] = w.projxy[2 * (size_t)i + 1];
    mp.valid[m] = 0.f;
    mp.init_id[m] = M + i;
    mp.frame_id[m] = frame_id;
    mp.fl[m * BF_FUSION_CAP] = M + i;
    mp.flen[m] = 1;
    c.fflag[M + i] = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// phase b - correspondence_association (instances.py:411-490) for one keyframe, one CTA; the sequential accept /
// replace tail with BoxManager.record_corr (box_manager.py:90-129) runs on thread 0 through the shared bf_record_one.
//   keep/success: flags over the N_glo + n boxes after nms_3d (keep is edited in place)
//   st->any_new = 1 if any new box survived nms_3d (demo.py:269)

#define BF_ECORR_THREADS 256
__global__ void __launch_bounds__(BF_ECORR_THREADS)
bf_engine_corr_kernel(bf_engine_ctx c) {
    __shared__ int s_any_new, s_any_small;
    __shared__ double s_bv[BF_ECORR_THREADS / 32];
    __shared__ int s_bi[BF_ECORR_THREADS / 32];
    bf_engine_state* st = c.st;
    if (st->first) return;
    const bf_map_buffers& mp = c.map[0];
    const int N_glo = st->N, n = st->n;
    int32_t* keep = c.keep;
    const int32_t* success = c.success;
    const float* sposes = c.store.pose;
    const float* in = c.in;
    const double fx = (double)in[2], fy = (double)in[3], cx = (double)in[4], cy = (double)in[5], W = (double)in[6], H = (double)in[7];
    const float* pinv = in + 40;                           // np.linalg.inv(pose), float32
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_any_new = 0; s_any_small = 0; }
    __syncthreads();
    // (valid_num += 1 for every head that suppressed something, instances.py:72-73, was applied by the NMS phase)
    for (int j = tid; j < n; j += T) {
        if (keep[N_glo + j]) {
            s_any_new = 1;
            const float* d = mp.tensor + 6 * (size_t)(N_glo + j) + 3;
            if (!(fmaxf(d[0], fmaxf(d[1], d[2])) > c.small_size) && !success[N_glo + j]) s_any_small = 1;
        }
    }
    // membership in cur_keep_idx refers to the state right after nms_3d (:428-435): snapshot the new boxes' flags
    for (int j = tid; j < n; j += T) c.snap[N_glo + j] = keep[N_glo + j];
    __syncthreads();
    if (tid == 0) st->any_new = s_any_new;
    if (!s_any_new || !s_any_small || N_glo == 0) return;
    double* boxes2d = c.boxes2d;
    int32_t* glo_keep_snapshot = c.snap;
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
    bf_record_ctx rc;
    rc.order = nullptr; rc.init_id = mp.init_id; rc.poses = sposes; rc.centers = nullptr; rc.fl = mp.fl; rc.flen = mp.flen;
    rc.fflag = c.fflag; rc.keep = keep; rc.status = &st->status[1]; rc.valid_num = nullptr;
    rc.translation_gap = c.tgap; rc.rotation_gap = c.rgap; rc.center_gap = 0.f;
    // small new boxes in index order; scoring in parallel, decision by thread 0 (:446-483)
    for (int j = 0; j < n; ++j) {
        const int cur_new = N_glo + j;
        const float* dn = mp.tensor + 6 * (size_t)cur_new + 3;
        const bool cand = glo_keep_snapshot[N_glo + j] && !(fmaxf(dn[0], fmaxf(dn[1], dn[2])) > c.small_size) && !success[cur_new];
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
            v = v * ((fmaxf(dg[0], fmaxf(dg[1], dg[2])) < c.small_plus) ? 1.0 : 0.0);           // (:460-461)
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
            if (bi != 0x7fffffff && bv > c.small_threshold) {
                const int cidx = bi;
                int cur, idx;
                if (mp.scores[cidx] < mp.scores[cur_new]) { cur = cur_new; idx = cidx; }   // the new box wins (:471-477)
                else { cur = cidx; idx = cur_new; }                                        // the old box wins (:478-483)
                keep[idx] = 0;                                                             // keep_idx = keep_idx[keep_idx != idx]
                mp.valid[cur] += 1.f;
                int len_c = mp.flen[cur];
                if (bf_record_one(rc, cur, idx, mp.fl + (size_t)cur * BF_FUSION_CAP, len_c) && keep[cur]) {
                    keep[cur] = 0; keep[idx] = 1;                                          // keep[keep == cur_id] = idx (box_manager.py:124)
                }
                mp.flen[cur] = len_c;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// phase c - compaction: one CTA scans the keep flags -> src[] (old index of every new row), st->Nnew, and flips the current
// map buffer set; a second kernel gathers every field.  `stale` != 0: the flags are formed here from
// BoxManager.check_valid_num's rule (box_manager.py:151-166) over the already compacted map.
#define BF_ESCAN_THREADS 1024
__global__ void __launch_bounds__(BF_ESCAN_THREADS)
bf_engine_scan_kernel(bf_engine_ctx c, int stale) {
    __shared__ int s_w[BF_ESCAN_THREADS / 32];
    __shared__ int s_total;
    bf_engine_state* st = c.st;
    // first keyframe: every detection becomes a map row as is; check_valid_num only runs when a new box survived (demo.py:269, 297)
    if (st->first || (stale && !st->any_new)) { if (threadIdx.x == 0) st->pad[1] = 0; return; }
    const bf_map_buffers& mp = c.map[0];
    const int frame_id = __float_as_int(c.in[1]);
    const int N = stale ? st->Nnew : st->Nall;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (N + BF_ESCAN_THREADS - 1) / BF_ESCAN_THREADS;
    const int lo = min(N, tid * per), hi = min(N, lo + per);
    const int thr = frame_id - c.gap;
    int cnt = 0;
    for (int i = lo; i < hi; ++i) {
        int k;
        if (stale) { k = !((mp.valid[i] == 0.f) && (mp.frame_id[i] < thr)); c.keep[i] = k; }
        else k = c.keep[i] ? 1 : 0;
        cnt += k;
    }
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
    for (int i = lo; i < hi; ++i) if (c.keep[i]) c.src[pos++] = i;
    __syncthreads();
    if (tid == 0) { st->Nnew = s_total; st->pad[1] = 1; }    // rows to gather; a gather is pending
}

// gather the kept rows of map[0] into map[1] (dir = 0), then copy them back to map[0] (dir = 1): map[0] is always the
// current buffer set, so the stand-alone kernels (corners, score order, NMS) can take its plain pointers.
__global__ void bf_engine_gather_kernel(bf_engine_ctx c, int dir) {
    bf_engine_state* st = c.st;
    if (!st->pad[1]) return;
    const int n_new = st->Nnew;
    const bf_map_buffers& a = c.map[dir ? 1 : 0];
    const bf_map_buffers& b = c.map[dir ? 0 : 1];
    const int t = threadIdx.x;                       // 64 threads: field elements
    for (int k = blockIdx.x; k < n_new; k += gridDim.x) {
        const size_t s = dir ? (size_t)k : (size_t)c.src[k], d = (size_t)k;
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
// st->B, st->SV, st->maxV.  One CTA.  The "view set fused before?" test (BoxManager.check_if_fusion, box_manager.py:34-38)
// walks the already_fusion hashes in shared-memory tiles: every thread compares its candidate rows against the whole
// tile, only hash hits touch the lists themselves.
#define BF_ESEL_TILE 2048
__global__ void __launch_bounds__(BF_ESCAN_THREADS)
bf_engine_select_kernel(bf_engine_ctx c) {
    __shared__ int s_w[BF_ESCAN_THREADS / 32];
    __shared__ int s_wv[BF_ESCAN_THREADS / 32];
    __shared__ int s_maxv;
    __shared__ unsigned long long s_hash[BF_ESEL_TILE];
    bf_engine_state* st = c.st;
    const bf_map_buffers& mp = c.map[0];
    const bf_fused_table& ft = c.fused;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = st->Nnew, F = ft.count[0];
    const bool active = st->any_new != 0 && !st->first && c.use_fusion;   // demo.py:269: fusion only when a new box survived
    if (tid == 0) s_maxv = 0;
    __syncthreads();
    const int per = (N + BF_ESCAN_THREADS - 1) / BF_ESCAN_THREADS;
    const int lo = min(N, tid * per), hi = min(N, lo + per);
    int cnt = 0, vsum = 0, vmax = 0;
    unsigned long long picks = 0ull;                          // per <= 64 rows per thread (N <= 65536): candidates, then survivors
    if (active)
        for (int i = lo; i < hi; ++i) {
            const int len = mp.flen[i];
            if (len < 3) continue;
            if (len > BF_MAX_VIEWS) { st->status[5] = BF_ERR_CAPACITY; continue; }
            picks |= 1ull << (i - lo);
        }
    if (active)
        for (int f0 = 0; f0 < F; f0 += BF_ESEL_TILE) {        // block-uniform trip count
            const int nf = min(BF_ESEL_TILE, F - f0);
            __syncthreads();
            for (int f = tid; f < nf; f += BF_ESCAN_THREADS) s_hash[f] = ft.hash[f0 + f];
            __syncthreads();
            unsigned long long rest = picks;
            while (rest) {
                const int k = __ffsll((long long)rest) - 1;
                rest &= rest - 1;
                const int i = lo + k;
                const int len = mp.flen[i];
                const int32_t* l = mp.fl + (size_t)i * BF_FUSION_CAP;
                const unsigned long long h = e_list_hash(l, len);
                for (int f = 0; f < nf; ++f) {
                    if (s_hash[f] != h || ft.len[f0 + f] != len) continue;
                    const int32_t* q = ft.lists + (size_t)(f0 + f) * BF_FUSION_CAP;
                    bool same = true;
                    for (int j = 0; j < len; ++j) same &= (q[j] == l[j]);
                    if (same) { picks &= ~(1ull << k); break; }
                }
            }
        }
    for (unsigned long long rest = picks; rest; rest &= rest - 1) {
        const int len = mp.flen[lo + __ffsll((long long)rest) - 1];
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
        c.todo[pos] = i; c.offsets[pos] = vpos;
        for (int k = 0; k < len; ++k) c.view_index[vpos + k] = mp.fl[(size_t)i * BF_FUSION_CAP + k];
        ++pos; vpos += len;
    }
    if (tid == BF_ESCAN_THREADS - 1) {
        const int B = s_w[BF_ESCAN_THREADS / 32 - 1], SV = s_wv[BF_ESCAN_THREADS / 32 - 1];
        c.offsets[B] = SV;
        st->B = B; st->SV = SV; st->maxV = s_maxv;
        st->pad[2] = F;                                       // already_fusion entries before this keyframe's write-back
        st->refine_boxes_total += B; st->refine_views_total += SV;
    }
}

// write back fused boxes (box_fusion.py:716-724), sequentially in map order like the reference's loop.  The rows in todo[]
// were not in already_fusion when they were selected; a row can only have been fused "earlier in this very call"
// (check_if_fusion at :634 runs inside the reference's loop) by one of the entries this call appended, so only those
// are compared.  `finish` != 0: also close the keyframe (bf_engine_finish_kernel's work; one launch less in the
// whole-keyframe graph).
__global__ void bf_engine_apply_kernel(bf_engine_ctx c, int finish) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    bf_engine_state* st = c.st;
    const bf_map_buffers& mp = c.map[0];
    const bf_fused_table& ft = c.fused;
    const int B = st->B;
    const int F0 = ft.count[0];
    int F = F0;
    for (int k = 0; k < B; ++k) {
        if (!c.upd[k]) continue;
        const int i = c.todo[k];
        const int len = mp.flen[i];
        const int32_t* l = mp.fl + (size_t)i * BF_FUSION_CAP;
        const unsigned long long h = e_list_hash(l, len);
        bool dup = false;
        for (int f = F0; f < F && !dup; ++f) {
            if (ft.hash[f] != h || ft.len[f] != len) continue;
            const int32_t* q = ft.lists + (size_t)f * BF_FUSION_CAP;
            bool same = true;
            for (int j = 0; j < len; ++j) same &= (q[j] == l[j]);
            dup = same;
        }
        if (dup) continue;                                    // fused earlier in this very call (check_if_fusion, :634)
        for (int q = 0; q < 6; ++q) mp.tensor[6 * (size_t)i + q] = c.out[6 * (size_t)k + q];
        c.fflag[i] = 1;
        if (F >= ft.cap) { st->status[3] = BF_ERR_CAPACITY; continue; }
        for (int q = 0; q < len; ++q) ft.lists[(size_t)F * BF_FUSION_CAP + q] = l[q];
        ft.len[F] = len; ft.hash[F] = h;
        ++F;
    }
    ft.count[0] = F;
    if (finish) { st->N = st->Nnew; st->M = st->M + st->n; st->steps += 1; st->n = 0; }
}

// close the keyframe: the counters the next keyframe starts from
__global__ void bf_engine_finish_kernel(bf_engine_ctx c) {
    bf_engine_state* st = c.st;
    st->N = st->Nnew;
    st->M = st->M + st->n;
    st->steps += 1;
    st->n = 0;
}


// ---------------------------------------------------------------------------------------------------------------
static bf_engine_ctx e_ctx(const bf_engine* e) {
    bf_engine_ctx c;
    c.map[0] = e->bufs.map[0]; c.map[1] = e->bufs.map[1]; c.store = e->bufs.store; c.fflag = e->bufs.fusion_flag; c.fused = e->bufs.fused;
    c.in = e->in_dev; c.st = e->state;
    c.keep = e->keep; c.success = e->success; c.todo = e->todo; c.offsets = e->offsets; c.view_index = e->view_index;
    c.src = e->src; c.snap = e->snap; c.upd = e->upd; c.out = e->out; c.boxes2d = e->boxes2d;
    c.ncap = e->cfg.map_capacity; c.mcap = e->cfg.store_capacity; c.max_det = e->cfg.max_det;
    c.tgap = e->cfg.translation_gap; c.rgap = e->cfg.rotation_gap; c.small_size = e->cfg.small_size; c.small_plus = e->cfg.small_plus;
    c.small_threshold = e->cfg.small_threshold; c.check_valid = e->cfg.check_valid; c.gap = e->cfg.gap; c.use_fusion = e->cfg.use_fusion;
    return c;
}

#define E_LAUNCH_CHECK(e, name)                                                               \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) return e_fail((e), BF_ERR_CUDA, name, cudaGetErrorString(e__)); \
    } while (0)
#define E_SUB(e, expr)                                                                        \
    do {                                                                                      \
        int rc__ = (expr);                                                                    \
        if (rc__) { snprintf((e)->err, sizeof((e)->err), "%s", (e)->h->err); return rc__; }   \
    } while (0)

// ---- the launch sequences (issued eagerly or under stream capture); each returns the number of kernels it launches ----
// phase bits of bf_engine_step (include/boxfusion_b200.h)
enum { PH_INGEST = 0, PH_NMS, PH_CORR, PH_COMPACT, PH_VALID, PH_FUSE, PH_FINISH, PH_COUNT };

static int e_ph_ingest(bf_engine* e, cudaStream_t st, int* L) {
    const bf_engine_ctx c = e_ctx(e);
    bf_engine_ingest_kernel<<<bf_blocks(e->cfg.max_det, 64), 64, 0, st>>>(c);
    E_LAUNCH_CHECK(e, "bf_engine_ingest_kernel");
    *L = 1;
    return BF_OK;
}

// Instances3D.spatial_association (demo.py:262): corners / centres of the Nall rows, score order, NMS with record()
static int e_ph_nms(bf_engine* e, cudaStream_t st, int* L) {
    const bf_engine_ctx c = e_ctx(e);
    const bf_engine_cfg& g = e->cfg;
    const bf_map_buffers& mp = e->bufs.map[0];
    bf_handle* h = e->h;
    const bf_dimref Nd = bf_dim_dev(&e->state->Nnms, g.map_capacity);
    E_SUB(e, bf_box_corners_run(h, mp.tensor, mp.R, Nd, e->corners, e->centers, st));
    E_SUB(e, bf_score_order_run(h, mp.scores, Nd, e->order, e->rank, st));
    E_SUB(e, bf_nms3d_run(h, e->corners, e->centers, Nd, e->order, e->rank, mp.init_id, e->bufs.store.pose, mp.fl, mp.flen,
                          e->bufs.fusion_flag, g.nms_threshold, g.translation_gap, g.rotation_gap, g.center_gap, g.iou_mode,
                          e->keep, e->success, &e->state->status[0], mp.valid, st));
    (void)c;
    // corners, order, planes, pairs, count, greedy
    *L = 6;
    return BF_OK;
}

static int e_ph_corr(bf_engine* e, cudaStream_t st, int* L) {
    const bf_engine_ctx c = e_ctx(e);
    bf_engine_corr_kernel<<<1, BF_ECORR_THREADS, 0, st>>>(c);
    E_LAUNCH_CHECK(e, "bf_engine_corr_kernel");
    *L = 1;
    return BF_OK;
}

static int e_compaction(bf_engine* e, cudaStream_t st, int stale, int* L) {
    const bf_engine_ctx c = e_ctx(e);
    const int ggrid = e->cfg.map_capacity < 1024 ? e->cfg.map_capacity : 1024;
    bf_engine_scan_kernel<<<1, BF_ESCAN_THREADS, 0, st>>>(c, stale);
    E_LAUNCH_CHECK(e, "bf_engine_scan_kernel");
    bf_engine_gather_kernel<<<ggrid, 64, 0, st>>>(c, 0);
    E_LAUNCH_CHECK(e, "bf_engine_gather_kernel");
    bf_engine_gather_kernel<<<ggrid, 64, 0, st>>>(c, 1);
    E_LAUNCH_CHECK(e, "bf_engine_gather_kernel");
    *L = 3;
    return BF_OK;
}
// all_pred_box[keep_idx] + box_manager.update(keep_idx) (demo.py:292 / 325-327)
static int e_ph_compact(bf_engine* e, cudaStream_t st, int* L) { return e_compaction(e, st, 0, L); }
// BoxManager.check_valid_num (box_manager.py:151-166, demo.py:297-298)
static int e_ph_valid(bf_engine* e, cudaStream_t st, int* L) { return e_compaction(e, st, 1, L); }

// BoxFusion.boxfusion (demo.py:304-305): selection, refinement, write-back
static int e_fuse(bf_engine* e, cudaStream_t st, int finish, int* L) {
    const bf_engine_ctx c = e_ctx(e);
    const bf_engine_cfg& g = e->cfg;
    bf_engine_select_kernel<<<1, BF_ESCAN_THREADS, 0, st>>>(c);
    E_LAUNCH_CHECK(e, "bf_engine_select_kernel");
    bf_refine_cfg rc = g.refine;
    rc.views_total = 0; rc.max_views = 0;
    bf_refine_dev_args dv;
    dv.B_dev = &e->state->B; dv.intr_dev = e->in_dev + 2; dv.max_boxes = g.map_capacity;
    const bf_store_buffers& sb = e->bufs.store;
    E_SUB(e, bf_refine_run(e->h, g.pst, g.P, sb.tensor, sb.R, sb.scores, sb.uv, sb.pose, e->offsets, e->view_index,
                           g.map_capacity, &rc, e->out, e->upd, e->its, nullptr, &e->state->status[2], &dv, st));
    bf_engine_apply_kernel<<<1, 32, 0, st>>>(c, finish);
    E_LAUNCH_CHECK(e, "bf_engine_apply_kernel");
    *L = 3;
    return BF_OK;
}
static int e_ph_fuse(bf_engine* e, cudaStream_t st, int* L) { return e_fuse(e, st, 0, L); }

static int e_ph_finish(bf_engine* e, cudaStream_t st, int* L) {
    const bf_engine_ctx c = e_ctx(e);
    bf_engine_finish_kernel<<<1, 1, 0, st>>>(c);
    E_LAUNCH_CHECK(e, "bf_engine_finish_kernel");
    *L = 1;
    return BF_OK;
}

typedef int (*e_phase_fn)(bf_engine*, cudaStream_t, int*);
static const e_phase_fn e_phases[PH_COUNT] = {e_ph_ingest, e_ph_nms, e_ph_corr, e_ph_compact, e_ph_valid, e_ph_fuse, e_ph_finish};

static int e_issue(bf_engine* e, int phases, cudaStream_t st, int* launches) {
    int total = 0;
    const bool fused_finish = (phases & (1 << PH_FUSE)) && (phases & (1 << PH_FINISH));   // bf_engine_apply closes the keyframe itself
    for (int p = 0; p < PH_COUNT; ++p) {
        if (!(phases & (1 << p)) || (p == PH_FINISH && fused_finish)) continue;
        int L = 0;
        const int rc = (p == PH_FUSE && fused_finish) ? e_fuse(e, st, 1, &L) : e_phases[p](e, st, &L);
        if (rc) return rc;
        total += L;
    }
    if (launches) *launches = total;
    return BF_OK;
}

static int e_capture_one(bf_engine* e, int phases, cudaGraphExec_t* exec, int* launches) {
    cudaGraph_t graph = nullptr;
    E_CUDA(e, cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeRelaxed));
    const int rc = e_issue(e, phases, e->cap_stream, launches);
    const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return e_fail(e, BF_ERR_CUDA, "cudaStreamEndCapture", cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return e_fail(e, BF_ERR_CUDA, "cudaGraphInstantiate", cudaGetErrorString(ie));
    return BF_OK;
}

// the whole keyframe as the engine's own configuration runs it
static int e_full_mask(const bf_engine* e) {
    return (1 << PH_INGEST) | (1 << PH_NMS) | (1 << PH_CORR) | (1 << PH_COMPACT) | (e->cfg.check_valid ? (1 << PH_VALID) : 0) |
           (e->cfg.use_fusion ? (1 << PH_FUSE) : 0) | (1 << PH_FINISH);
}

extern "C" const char* bf_engine_last_error(bf_engine* e) { return e ? e->err : "null engine"; }

extern "C" void bf_engine_destroy(bf_engine* e) {
    if (!e) return;
    if (e->h) cudaSetDevice(e->h->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < PH_COUNT + 2; ++p) if (e->exec[p]) cudaGraphExecDestroy(e->exec[p]);
    for (int i = 0; i < 2; ++i) if (e->ra_exec[i]) cudaGraphExecDestroy(e->ra_exec[i]);
    if (e->snapshot) cudaFree(e->snapshot);
    for (int i = 0; i < 2; ++i) { if (e->hflags[i]) cudaFreeHost(e->hflags[i]); if (e->flag_evt[i]) cudaEventDestroy(e->flag_evt[i]); }
    if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
    for (int i = 0; i < bf_engine::RING; ++i) { if (e->stage[i]) cudaFreeHost(e->stage[i]); if (e->stage_evt[i]) cudaEventDestroy(e->stage_evt[i]); }
    void* bufs[] = {e->in_dev, e->state, e->keep, e->success, e->todo, e->offsets, e->view_index, e->order, e->rank, e->src, e->snap,
                    e->upd, e->its, e->corners, e->centers, e->out, e->boxes2d};
    for (void* b : bufs) if (b) cudaFree(b);
    if (e->h) bf_destroy(e->h);
    free(e);
}

extern "C" int bf_engine_create(int device, const bf_engine_cfg* cfg, const bf_engine_buffers* bufs, bf_engine** out) {
    if (!out) return BF_ERR_INVALID_ARG;
    *out = nullptr;
    if (!cfg || !bufs || cfg->map_capacity < 1 || cfg->map_capacity > 65536 || cfg->store_capacity < 1 || cfg->max_det < 1 ||
        cfg->max_det > cfg->map_capacity || !cfg->pst || cfg->P < 1 || cfg->P > BF_MAX_PARTICLES || cfg->refine.iters < 1 ||
        cfg->refine.max_hits < 1 || cfg->refine.max_hits > 4096)
        return BF_ERR_INVALID_ARG;
    bf_engine* e = (bf_engine*)calloc(1, sizeof(bf_engine));
    if (!e) return BF_ERR_INVALID_ARG;
    int rc = bf_create(device, &e->h);
    if (rc) { free(e); return rc; }
    bf_device_guard guard(e->h);
    e->cfg = *cfg; e->bufs = *bufs;
    e->h->refine_concurrent = cfg->concurrent ? 1 : 0;
    const size_t ncap = (size_t)cfg->map_capacity;
    e->in_floats = BF_KF_HEADER + (size_t)BF_KF_ROW * cfg->max_det;
#define E_ALLOC(ptr, bytes) do { cudaError_t ce = cudaMalloc((void**)&(ptr), (bytes)); if (ce != cudaSuccess) { bf_engine_destroy(e); return BF_ERR_CUDA; } cudaMemset((ptr), 0, (bytes)); } while (0)
    E_ALLOC(e->in_dev, sizeof(float) * e->in_floats);
    E_ALLOC(e->state, sizeof(bf_engine_state));
    E_ALLOC(e->keep, sizeof(int32_t) * ncap); E_ALLOC(e->success, sizeof(int32_t) * ncap); E_ALLOC(e->todo, sizeof(int32_t) * ncap);
    E_ALLOC(e->offsets, sizeof(int32_t) * (ncap + 1)); E_ALLOC(e->view_index, sizeof(int32_t) * ncap * BF_FUSION_CAP);
    E_ALLOC(e->order, sizeof(int32_t) * ncap); E_ALLOC(e->rank, sizeof(int32_t) * ncap); E_ALLOC(e->src, sizeof(int32_t) * (ncap + 1));
    E_ALLOC(e->snap, sizeof(int32_t) * (ncap + 1)); E_ALLOC(e->upd, sizeof(int32_t) * ncap); E_ALLOC(e->its, sizeof(int32_t) * ncap);
    E_ALLOC(e->corners, sizeof(float) * 24 * ncap); E_ALLOC(e->centers, sizeof(float) * 3 * ncap); E_ALLOC(e->out, sizeof(float) * 6 * ncap);
    E_ALLOC(e->boxes2d, sizeof(double) * 4 * (ncap + 1));
#undef E_ALLOC
    for (int i = 0; i < bf_engine::RING; ++i) {
        if (cudaMallocHost((void**)&e->stage[i], sizeof(float) * e->in_floats) != cudaSuccess ||
            cudaEventCreateWithFlags(&e->stage_evt[i], cudaEventDisableTiming) != cudaSuccess) { bf_engine_destroy(e); return BF_ERR_CUDA; }
    }
    if (cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking) != cudaSuccess) { bf_engine_destroy(e); return BF_ERR_CUDA; }
    *out = e;                                             // from here on the caller reads bf_engine_last_error and destroys on failure
    // one eager pass over an empty keyframe: allocates every scratch block at its bound and validates the launches
    rc = e_issue(e, (1 << PH_COUNT) - 1, e->cap_stream, nullptr);
    if (rc) return rc;
    if (cudaStreamSynchronize(e->cap_stream) != cudaSuccess) return e_fail(e, BF_ERR_CUDA, "bf_engine_create", cudaGetErrorString(cudaGetLastError()));
    cudaMemset(e->state, 0, sizeof(bf_engine_state));
    e->h->frozen = 1;
    if (cfg->use_graph) {
        for (int p = 0; p < PH_COUNT; ++p)
            if ((rc = e_capture_one(e, 1 << p, &e->exec[p], &e->launches[p]))) return rc;
        if ((rc = e_capture_one(e, e_full_mask(e), &e->exec[PH_COUNT], &e->launches[PH_COUNT]))) return rc;
        if ((rc = e_capture_one(e, e_full_mask(e) & ~((1 << PH_INGEST) | (1 << PH_NMS) | (1 << PH_CORR)), &e->exec[PH_COUNT + 1], &e->launches[PH_COUNT + 1]))) return rc;
        e->have_graph = 1;
    } else {
        // launch counts of the eager sequences (same kernels)
        e->launches[PH_INGEST] = 1; e->launches[PH_NMS] = 6; e->launches[PH_CORR] = 1;
        e->launches[PH_COMPACT] = 3; e->launches[PH_VALID] = 3; e->launches[PH_FUSE] = 3; e->launches[PH_FINISH] = 1;
        int tot = 0;
        for (int p = 0; p < PH_COUNT; ++p) if (e_full_mask(e) & (1 << p)) tot += e->launches[p];
        if (cfg->use_fusion) tot -= 1;                       // apply closes the keyframe itself
        e->launches[PH_COUNT] = tot;
    }
    return BF_OK;
}

extern "C" int bf_engine_reset(bf_engine* e, void* stream) {
    if (!e) return BF_ERR_INVALID_ARG;
    bf_device_guard guard(e->h);
    cudaStream_t st = (cudaStream_t)stream;
    E_CUDA(e, cudaMemsetAsync(e->state, 0, sizeof(bf_engine_state), st));
    E_CUDA(e, cudaMemsetAsync(e->bufs.fused.count, 0, sizeof(int32_t), st));
    return BF_OK;
}

static int e_run(bf_engine* e, int phases, cudaStream_t st) {
    if (!e->have_graph) return e_issue(e, phases, st, nullptr);
    if (phases == e_full_mask(e)) { E_CUDA(e, cudaGraphLaunch(e->exec[PH_COUNT], st)); return BF_OK; }
    if (phases == (e_full_mask(e) & ~((1 << PH_INGEST) | (1 << PH_NMS) | (1 << PH_CORR)))) { E_CUDA(e, cudaGraphLaunch(e->exec[PH_COUNT + 1], st)); return BF_OK; }
    for (int p = 0; p < PH_COUNT; ++p)
        if (phases & (1 << p)) E_CUDA(e, cudaGraphLaunch(e->exec[p], st));
    return BF_OK;
}

// host -> device copy of `floats` words into in_dev through the pinned staging ring
static int e_upload(bf_engine* e, const float* src, size_t floats, cudaStream_t st) {
    const int slot = e->stage_next;
    e->stage_next = (slot + 1) % bf_engine::RING;
    E_CUDA(e, cudaEventSynchronize(e->stage_evt[slot]));       // the copy issued RING keyframes ago has left this slot
    memcpy(e->stage[slot], src, sizeof(float) * floats);
    E_CUDA(e, cudaMemcpyAsync(e->in_dev, e->stage[slot], sizeof(float) * floats, cudaMemcpyHostToDevice, st));
    E_CUDA(e, cudaEventRecord(e->stage_evt[slot], st));
    return BF_OK;
}

extern "C" int bf_engine_step(bf_engine* e, const float* packed, int n, int phases, void* stream) {
    if (!e || n < 0 || n > e->cfg.max_det) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_step", "bad argument");
    bf_device_guard guard(e->h);
    cudaStream_t st = (cudaStream_t)stream;
    if (phases == 0) phases = e_full_mask(e);
    if (phases & (1 << PH_INGEST)) {
        if (!packed) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_step", "null keyframe");
        const int rc = e_upload(e, packed, BF_KF_HEADER + (size_t)BF_KF_ROW * n, st);
        if (rc) return rc;
    }
    return e_run(e, phases, st);
}

extern "C" int bf_engine_step_device(bf_engine* e, const float* packed_dev, int n, int phases, void* stream) {
    if (!e || n < 0 || n > e->cfg.max_det) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_step_device", "bad argument");
    bf_device_guard guard(e->h);
    cudaStream_t st = (cudaStream_t)stream;
    if (phases == 0) phases = e_full_mask(e);
    if (phases & (1 << PH_INGEST)) {
        if (!packed_dev) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_step_device", "null keyframe");
        E_CUDA(e, cudaMemcpyAsync(e->in_dev, packed_dev, sizeof(float) * (BF_KF_HEADER + (size_t)BF_KF_ROW * n), cudaMemcpyDeviceToDevice, st));
    }
    return e_run(e, phases, st);
}

extern "C" int bf_engine_ingest_world(bf_engine* e, const float* header /*host, BF_KF_HEADER floats*/, const float* tensor_w,
                                      const float* R_w, const float* scores, const float* box2d, const float* projxy, const float* uv,
                                      int n, void* stream) {
    if (!e || !header || n < 0 || n > e->cfg.max_det || (n > 0 && (!tensor_w || !R_w || !scores || !box2d || !projxy || !uv)))
        return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_ingest_world", "bad argument");
    bf_device_guard guard(e->h);
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = e_upload(e, header, BF_KF_HEADER, st);
    if (rc) return rc;
    bf_world_rows w;
    w.tensor = tensor_w; w.R = R_w; w.scores = scores; w.box2d = box2d; w.projxy = projxy; w.uv = uv;
    bf_engine_ingest_world_kernel<<<bf_blocks(e->cfg.max_det, 64), 64, 0, st>>>(e_ctx(e), w);
    E_LAUNCH_CHECK(e, "bf_engine_ingest_world_kernel");
    return BF_OK;
}

__global__ void bf_engine_set_counts_kernel(bf_engine_ctx c, int N, int M) {
    bf_engine_state* st = c.st;
    st->N = N; st->M = M; st->n = 0; st->Nall = N; st->Nnms = 0; st->Nnew = N; st->first = 0; st->any_new = 0; st->B = 0; st->pad[1] = 0;
}

extern "C" int bf_engine_set_counts(bf_engine* e, int N, int M, void* stream) {
    if (!e || N < 0 || M < 0 || N > e->cfg.map_capacity || M > e->cfg.store_capacity) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_set_counts", "bad argument");
    bf_device_guard guard(e->h);
    bf_engine_set_counts_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(e_ctx(e), N, M);
    E_LAUNCH_CHECK(e, "bf_engine_set_counts_kernel");
    return BF_OK;
}

// ---- run-ahead for the reference-shaped API -------------------------------------------------------------------------------------
// spatial_association has to hand the keep / success indices to the caller, but nothing the caller does with them feeds back
// into the keyframe: bf_engine_run_ahead issues the NMS phase, an asynchronous copy of the flags to pinned memory, a snapshot
// of everything the later phases overwrite, the correspondence phase, a second copy of the keep flags, and the rest of the
// keyframe - so the GPU refines while the host is still inside the caller's Python.  If the caller then strays from demo.py's
// sequence, bf_engine_rollback restores the state of right after the NMS phase and the phases are re-issued one by one:
// running ahead is invisible.
#define BF_SNAP_ROW 90                 // words per map row: 6 + 9 + 1 + 4 + 2 + 16 + 16 + 1 + 1 + 1 + 32 + 1
__global__ void bf_engine_snapshot_kernel(bf_engine_ctx c, uint32_t* __restrict__ snap, int restore) {
    bf_engine_state* st = c.st;
    uint32_t* s_keep = snap + (size_t)c.ncap * BF_SNAP_ROW;
    uint32_t* s_flag = s_keep + c.ncap;
    uint32_t* s_state = s_flag + c.ncap;                  // 32 words of state, then the fused count
    const int N = restore ? ((const bf_engine_state*)s_state)->Nall : st->Nall;
    const bf_map_buffers& mp = c.map[0];
    const int t = threadIdx.x;                            // 64 threads per row
    for (int k = blockIdx.x; k < N; k += gridDim.x) {
        uint32_t* row = snap + (size_t)k * BF_SNAP_ROW;
        const size_t d = (size_t)k;
#define BF_SNAP_FIELD(ptr, width, off)                                                                              \
        if (t < (width)) { uint32_t* q = (uint32_t*)(ptr) + (width) * d + t; if (restore) *q = row[(off) + t]; else row[(off) + t] = *q; }
        BF_SNAP_FIELD(mp.tensor, 6, 0) BF_SNAP_FIELD(mp.R, 9, 6) BF_SNAP_FIELD(mp.scores, 1, 15) BF_SNAP_FIELD(mp.box2d, 4, 16)
        BF_SNAP_FIELD(mp.projxy, 2, 20) BF_SNAP_FIELD(mp.pose, 16, 22) BF_SNAP_FIELD(mp.uv, 16, 38) BF_SNAP_FIELD(mp.valid, 1, 54)
        BF_SNAP_FIELD(mp.init_id, 1, 55) BF_SNAP_FIELD(mp.frame_id, 1, 56) BF_SNAP_FIELD(mp.fl, BF_FUSION_CAP, 57)
        BF_SNAP_FIELD(mp.flen, 1, 89)
#undef BF_SNAP_FIELD
        if (t == 63) {
            if (restore) { c.keep[k] = (int32_t)s_keep[k]; c.fflag[k] = (int32_t)s_flag[k]; }
            else { s_keep[k] = (uint32_t)c.keep[k]; s_flag[k] = (uint32_t)c.fflag[k]; }
        }
    }
    if (blockIdx.x == 0 && t < 32) {
        if (restore) ((uint32_t*)st)[t] = s_state[t]; else s_state[t] = ((const uint32_t*)st)[t];
        if (t == 0) { if (restore) c.fused.count[0] = (int32_t)s_state[32]; else s_state[32] = (uint32_t)c.fused.count[0]; }
    }
}

// the keep / success flags of the rows in play and the state block, written straight into pinned host memory (device-
// accessible under unified addressing): one kernel instead of three copies, and it can sit inside a captured graph
__global__ void bf_engine_publish_kernel(bf_engine_ctx c, int32_t* __restrict__ hkeep, int32_t* __restrict__ hsucc,
                                         uint32_t* __restrict__ hstate) {
    const int N = c.st->Nall;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) {
        hkeep[k] = c.keep[k];
        if (hsucc) hsucc[k] = c.success[k];
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) hstate[threadIdx.x] = ((const uint32_t*)c.st)[threadIdx.x];
}

static int e_publish(bf_engine* e, int slot, int with_success, cudaStream_t st) {
    int32_t* hp = e->hflags[slot];
    const size_t ncap = (size_t)e->cfg.map_capacity;
    const int grid = e->cfg.map_capacity < 4096 ? (e->cfg.map_capacity + 255) / 256 : 16;
    bf_engine_publish_kernel<<<grid, 256, 0, st>>>(e_ctx(e), hp, with_success ? hp + ncap : nullptr, (uint32_t*)(hp + 2 * ncap));
    E_LAUNCH_CHECK(e, "bf_engine_publish_kernel");
    return BF_OK;
}

static int e_snapshot(bf_engine* e, cudaStream_t st) {
    const int sgrid = e->cfg.map_capacity < 1024 ? e->cfg.map_capacity : 1024;
    bf_engine_snapshot_kernel<<<sgrid, 64, 0, st>>>(e_ctx(e), e->snapshot, 0);
    E_LAUNCH_CHECK(e, "bf_engine_snapshot_kernel");
    return BF_OK;
}

// part 0: NMS phase, flags -> slot 0.  part 1: snapshot of what the later phases overwrite, correspondence phase, flags -> slot 1.
static int e_ra_issue(bf_engine* e, int part, cudaStream_t st) {
    int rc;
    if (part == 0) {
        if ((rc = e_issue(e, 1 << PH_NMS, st, nullptr))) return rc;
        return e_publish(e, 0, 1, st);
    }
    if ((rc = e_snapshot(e, st))) return rc;
    if ((rc = e_issue(e, 1 << PH_CORR, st, nullptr))) return rc;
    return e_publish(e, 1, 0, st);
}

static int e_ra_capture(bf_engine* e, int part) {
    cudaGraph_t graph = nullptr;
    E_CUDA(e, cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeRelaxed));
    const int rc = e_ra_issue(e, part, e->cap_stream);
    const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return e_fail(e, BF_ERR_CUDA, "cudaStreamEndCapture", cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&e->ra_exec[part], graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return e_fail(e, BF_ERR_CUDA, "cudaGraphInstantiate", cudaGetErrorString(ie));
    return BF_OK;
}

static int e_tail_mask(const bf_engine* e) { return e_full_mask(e) & ~((1 << PH_INGEST) | (1 << PH_NMS) | (1 << PH_CORR)); }

// Five host calls per keyframe: graph (NMS + flags), event, graph (snapshot + correspondence + flags), event, graph (the rest).
extern "C" int bf_engine_run_ahead(bf_engine* e, int rows, void* stream) {
    if (!e || rows < 0 || rows > e->cfg.map_capacity) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_run_ahead", "bad argument");
    bf_device_guard guard(e->h);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ncap = (size_t)e->cfg.map_capacity;
    int rc;
    if (!e->snapshot) {
        E_CUDA(e, cudaMalloc((void**)&e->snapshot, sizeof(uint32_t) * (ncap * BF_SNAP_ROW + 2 * ncap + 64)));
        for (int i = 0; i < 2; ++i) {
            E_CUDA(e, cudaMallocHost((void**)&e->hflags[i], sizeof(int32_t) * (2 * ncap + 32)));
            E_CUDA(e, cudaEventCreateWithFlags(&e->flag_evt[i], cudaEventDisableTiming));
        }
        if (e->have_graph)
            for (int part = 0; part < 2; ++part)
                if ((rc = e_ra_capture(e, part))) return rc;
    }
    for (int part = 0; part < 2; ++part) {
        if (e->ra_exec[part]) E_CUDA(e, cudaGraphLaunch(e->ra_exec[part], st));
        else if ((rc = e_ra_issue(e, part, st))) return rc;
        E_CUDA(e, cudaEventRecord(e->flag_evt[part], st));
    }
    return e_run(e, e_tail_mask(e), st);
}

extern "C" int bf_engine_wait_flags(bf_engine* e, int slot, int32_t** keep, int32_t** success, bf_engine_state** state) {
    if (!e || slot < 0 || slot > 1 || !e->hflags[slot]) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_wait_flags", "bad argument");
    bf_device_guard guard(e->h);
    E_CUDA(e, cudaEventSynchronize(e->flag_evt[slot]));
    const size_t ncap = (size_t)e->cfg.map_capacity;
    if (keep) *keep = e->hflags[slot];
    if (success) *success = e->hflags[slot] + ncap;
    if (state) *state = (bf_engine_state*)(e->hflags[slot] + 2 * ncap);
    return BF_OK;
}

extern "C" int bf_engine_rollback(bf_engine* e, void* stream) {
    if (!e || !e->snapshot) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_rollback", "nothing to roll back");
    bf_device_guard guard(e->h);
    const int g = e->cfg.map_capacity < 1024 ? e->cfg.map_capacity : 1024;
    bf_engine_snapshot_kernel<<<g, 64, 0, (cudaStream_t)stream>>>(e_ctx(e), e->snapshot, 1);
    E_LAUNCH_CHECK(e, "bf_engine_snapshot_kernel");
    return BF_OK;
}

extern "C" int bf_engine_read_state(bf_engine* e, bf_engine_state* out, void* stream) {
    if (!e || !out) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_read_state", "bad argument");
    bf_device_guard guard(e->h);
    cudaStream_t st = (cudaStream_t)stream;
    E_CUDA(e, cudaMemcpyAsync(out, e->state, sizeof(bf_engine_state), cudaMemcpyDeviceToHost, st));
    E_CUDA(e, cudaStreamSynchronize(st));
    return BF_OK;
}

extern "C" int bf_engine_read_flags(bf_engine* e, int32_t* keep, int32_t* success, int count, bf_engine_state* state, void* stream) {
    if (!e || count < 0 || count > e->cfg.map_capacity) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_read_flags", "bad argument");
    bf_device_guard guard(e->h);
    cudaStream_t st = (cudaStream_t)stream;
    if (keep && count) E_CUDA(e, cudaMemcpyAsync(keep, e->keep, sizeof(int32_t) * (size_t)count, cudaMemcpyDeviceToHost, st));
    if (success && count) E_CUDA(e, cudaMemcpyAsync(success, e->success, sizeof(int32_t) * (size_t)count, cudaMemcpyDeviceToHost, st));
    if (state) E_CUDA(e, cudaMemcpyAsync(state, e->state, sizeof(bf_engine_state), cudaMemcpyDeviceToHost, st));
    E_CUDA(e, cudaStreamSynchronize(st));
    return BF_OK;
}

// diagnostic read-back of engine-owned per-keyframe results: which = 0 refine iterations per box, 1 CSR view offsets,
// 2 rows selected for refinement, 3 refine `updated` flags (synchronises)
extern "C" int bf_engine_read_i32(bf_engine* e, int which, int32_t* out, int count, void* stream) {
    if (!e || !out || count < 0 || count > e->cfg.map_capacity + 1) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_read_i32", "bad argument");
    bf_device_guard guard(e->h);
    const int32_t* src = which == 0 ? e->its : which == 1 ? e->offsets : which == 2 ? e->todo : which == 3 ? e->upd : nullptr;
    if (!src) return e_fail(e, BF_ERR_INVALID_ARG, "bf_engine_read_i32", "bad selector");
    cudaStream_t st = (cudaStream_t)stream;
    if (count) E_CUDA(e, cudaMemcpyAsync(out, src, sizeof(int32_t) * (size_t)count, cudaMemcpyDeviceToHost, st));
    E_CUDA(e, cudaStreamSynchronize(st));
    return BF_OK;
}

extern "C" int bf_engine_pointers(bf_engine* e, int32_t** keep, int32_t** success, bf_engine_state** state_dev, int32_t** refine_iters,
                                  int32_t** todo) {
    if (!e) return BF_ERR_INVALID_ARG;
    if (keep) *keep = e->keep;
    if (success) *success = e->success;
    if (state_dev) *state_dev = e->state;
    if (refine_iters) *refine_iters = e->its;
    if (todo) *todo = e->todo;
    return BF_OK;
}

extern "C" int bf_engine_launch_counts(bf_engine* e, int32_t* counts /*[8]: per phase bit, then the whole keyframe*/) {
    if (!e || !counts) return BF_ERR_INVALID_ARG;
    for (int p = 0; p < PH_COUNT + 1; ++p) counts[p] = e->launches[p];
    return BF_OK;
}
