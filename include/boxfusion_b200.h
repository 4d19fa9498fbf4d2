/* boxfusion_b200 - C ABI of the B200-native (sm_100a) multi-view box-fusion hot path.
 *
 * One shared library, `libboxfusion_sm100.so`; plain pointers and sizes only (no torch types).
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * tree pliam1105/BoxFusion; see SURVEY.md section 8(a) for the row ids A1..A22).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers unless the parameter is documented "host".
 *   - Callers own every buffer; the library borrows them for the duration of the call and keeps
 *     only its own scratch inside the opaque handle.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous with respect to the host
 *     unless documented otherwise.  One handle per device; a handle is not re-entrant.
 *   - Return value: BF_OK (0) or a negative BF_ERR_*; bf_last_error(h) gives the text.
 *   - There is no CPU fallback: bf_create fails (BF_ERR_CUDA) when the device is not sm_100.
 *   - Row-major float32 unless stated; boxes are (x,y,z,l,h,w) with l<->X, h<->Y, w<->Z of the box
 *     frame and R[9] row-major (boxes.py:725-778); poses are camera->world 4x4 row-major
 *     (box_fusion.py:348-354).
 */
#ifndef BOXFUSION_B200_H
#define BOXFUSION_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bf_handle bf_handle;

enum {
    BF_OK = 0,
    BF_ERR_INVALID_ARG = -1,
    BF_ERR_CUDA = -2,
    BF_ERR_CAPACITY = -3   /* a fixed-capacity structure overflowed (fusion list, polygon buffer, work list) */
};

/* IoU estimator of bf_iou3d_*  */
enum {
    BF_IOU_SAMPLED_REF = 0, /* reference-exact: containment gate + 25^3 grid counts (instances.py:514-613) */
    BF_IOU_ANALYTIC = 1     /* gravity-aligned BEV Sutherland-Hodgman clip x height overlap; pairs without a
                               shared box axis fall back (on the GPU) to BF_IOU_SAMPLED_REF */
};

#define BF_FUSION_CAP 32    /* max observations per fusion list held on the device (box_manager.py:13) */
#define BF_MAX_VIEWS 64     /* max views of one box in bf_refine */
#define BF_MAX_PARTICLES 4096

int bf_version(void);
int bf_create(int device, bf_handle** out);
void bf_destroy(bf_handle* h);
const char* bf_last_error(bf_handle* h);
int bf_fusion_cap(void);

/* ---- A1  GeneralInstance3DBoxes.corners (boxes.py:725-778) ------------------------------------
 * corners[n][v][3], vertex order v0..v7 of the reference; centers (optional) = mean of the 8 corners
 * as nms_3d computes them (instances.py:49). */
int bf_box_corners(bf_handle* h, const float* xyzlhw /*[N,6]*/, const float* R /*[N,9]*/, int N,
                   float* corners /*[N,8,3]*/, float* centers /*[N,3] or NULL*/, void* stream);

/* ---- A2  GeneralInstance3DBoxes.transform2world (boxes.py:825-833), in place -------------------- */
int bf_transform2world(bf_handle* h, float* xyzlhw /*[N,6]*/, float* R /*[N,9]*/, const float* poses /*[N,16]*/,
                       int N, void* stream);

/* ---- A15 Instances3D.project_3d_boxes (instances.py:333-369) -----------------------------------
 * Observation corners in each detection's own camera, clamped to [0,W]x[0,H].  pose_inv is the
 * world->camera matrix; the host binding obtains it with the reference's own torch.linalg.inv call
 * (:350) so that the projected corners are the reference's to the last bit of the einsum (:352). */
int bf_project_boxes(bf_handle* h, const float* corners /*[N,8,3]*/, const float* pose_inv /*[N,16]*/, int N,
                     float fx, float fy, float cx, float cy, float W, float H, float* uv /*[N,8,2]*/, void* stream);

/* ---- A3/A4  Instances3D.obb_iou / calculate_obb_iou (instances.py:106-125, 573-613) -------------
 * IoU of every (a in A) x (b in B) from corner arrays.  iou is float64 like the reference's;
 * counts (optional) receives {count1,count2,common} of the 25^3 estimator (zeros when the gate
 * fails or the analytic path was taken).  stats (optional, 4 x int64, device): pairs, AABB-passing
 * pairs, gate-passing pairs, analytic pairs. */
int bf_iou3d_matrix(bf_handle* h, const float* cornersA /*[M,8,3]*/, int M, const float* cornersB /*[N,8,3]*/, int N,
                    int mode, double* iou /*[M,N]*/, int32_t* counts /*[M,N,3] or NULL*/, int64_t* stats /*[4] or NULL*/,
                    void* stream);

/* ---- A5/A6/A7/A8  nms_3d + BoxManager.record (instances.py:22-101, box_manager.py:40-88,188-215)
 * Greedy score-ordered 3-D NMS over N boxes with the fusion-list bookkeeping of record().
 *   order      [N] int32: box indices by descending score (scores.argsort()[::-1])
 *   init_id    [N] int32: per-frame observation index of each box (instances.py:53)
 *   poses      [M,16]: per-frame camera poses indexed by the values stored in the fusion lists
 *   fusion_list[N,BF_FUSION_CAP] / fusion_len[N] / fusion_flag[N] (int32): in/out, rows kept sorted
 * outputs (int32, device): keep[N] 0/1, success[N] 0/1 (heads that suppressed something; valid_num += 1),
 *   status[1]: 0 or BF_ERR_CAPACITY if a fusion list overflowed BF_FUSION_CAP.
 * IoU uses `mode`; suppression is `iou > iou_threshold` evaluated in float64. */
int bf_nms3d(bf_handle* h, const float* corners /*[N,8,3]*/, const float* centers /*[N,3]*/, int N,
             const int32_t* order, const int32_t* init_id, const float* poses, int M,
             int32_t* fusion_list, int32_t* fusion_len, int32_t* fusion_flag,
             double iou_threshold, float translation_gap, float rotation_gap_deg, float center_gap, int mode,
             int32_t* keep, int32_t* success, int32_t* status, void* stream);

/* ---- A9-A12  correspondence_association core (instances.py:446-483, 643-717; box_manager.py:90-129)
 * For each of n_small detections (2-D boxes det_xyxy, float32) project the G candidate map boxes
 * (corners, float32) with pose_inv (row-major 4x4 float32, already inverted by the caller exactly as
 * the reference does with np.linalg.inv) and K, form the clipped 2-D AABB of corners with 0<Z<8
 * (float64), IoU against the detection (float64, +1e-6), zero rows whose map box is not small
 * (small_mask[G] int32), and return the first arg-max and its IoU per detection. */
int bf_corr2d(bf_handle* h, const float* map_corners /*[G,8,3]*/, const int32_t* small_mask /*[G]*/, int G,
              const float* pose_inv /*[16]*/, float fx, float fy, float cx, float cy, float W, float H,
              const float* det_xyxy /*[n_small,4]*/, int n_small,
              double* boxes2d /*[G,4] or NULL*/, int32_t* best /*[n_small]*/, double* best_iou /*[n_small]*/, void* stream);

/* ---- A3  Instances3D.batch_in_convex_hull_3d (instances.py:559-571) -------------------------------------------
 * inside[i] = 1 when point i satisfies n.p + d <= 1e-6 for the 12 hull half-spaces of the box (the same planes and
 * predicate bf_iou3d_matrix uses for the containment gate and the 25^3 counts).  Helper entry for the drop-in's
 * check_intersection / batch_in_convex_hull_3d; the association path itself never calls it. */
int bf_points_in_hull(bf_handle* h, const float* corners /*[8,3]*/, const double* points /*[n,3]*/, int n,
                      uint8_t* inside /*[n]*/, void* stream);

/* ---- A5  the score order of nms_3d, `order = scores.argsort()[::-1]` (instances.py:52), on the device --------
 * order[r] = index of the r-th highest score; equal scores keep ascending index (a stable descending sort, what the
 * Python binding obtains from torch.argsort(descending=True, stable=True); NumPy's own argsort is unstable for exact
 * ties).  NaN scores sort first, -0 == +0.  One CTA, bitonic network in shared memory: N <= 4096
 * (BF_ERR_INVALID_ARG beyond; callers with larger maps sort elsewhere). */
int bf_score_order(bf_handle* h, const float* scores /*[N]*/, int N, int32_t* order /*[N]*/, void* stream);

/* ---- A8  BoxManager.compute_pose_disparity (box_manager.py:168-186), batched ------------------- */
int bf_pose_disparity(bf_handle* h, const float* poses /*[M,16]*/, const int32_t* ia, const int32_t* ib, int n,
                      float* baseline /*[n]*/, float* angle_deg /*[n]*/, void* stream);

/* ---- A16-A22  BoxFusion.boxfusion optimiser (box_fusion.py:264-405, 413-600, 651-721) ----------
 * Refines B map boxes independently; box b fuses the observations view_index[view_offsets[b] ..
 * view_offsets[b+1]) (CSR, int32) of the per-frame store.  All `iters` optimiser iterations run
 * inside one launch.  Float32 arithmetic with the same operation order as the reference kernel
 * (compiled without FMA contraction), line intersections in float64, global box state in float64.
 */
typedef struct {
    int32_t iters;            /* box_fusion.iters (20)                                   */
    int32_t pst_size;         /* box_fusion.pst_size; particles >= 32*(pst_size/32) score 0 */
    float center_init, shape_init;      /* random_opt.*_init_size                          */
    float center_scale, shape_scale;    /* random_opt.*_scaling_coefficient                */
    double beta;              /* 0.9                                                     */
    float img_h, img_w;       /* BoxFusion.H / .W after update_intrinsics                */
    float fx, cx, fy, cy;     /* K[0],K[2],K[5],K[6] of the flattened 4x4 (box_fusion.py:356-357) */
    int32_t max_hits;         /* 200                                                     */
    int32_t early_stop;       /* 1: stop after 3 consecutive failures (reference); 0: run all iters */
    int32_t views_total;      /* entries of view_index (= view_offsets[B]); 0 = unknown (scratch is sized for BF_MAX_VIEWS per box) */
    int32_t max_views;        /* largest number of views of any box in this call; 0 = unknown (BF_MAX_VIEWS) */
} bf_refine_cfg;

int bf_refine(bf_handle* h, const float* pst /*[P,6]*/, int P,
              const float* per_xyzlhw /*[M,6]*/, const float* per_R /*[M,9]*/, const float* per_scores /*[M]*/,
              const float* per_uv /*[M,16]*/, const float* per_poses /*[M,16]*/, int M,
              const int32_t* view_offsets /*[B+1]*/, const int32_t* view_index /*[sum V]*/, int B,
              const bf_refine_cfg* cfg /*host*/,
              float* out_xyzlhw /*[B,6]*/, int32_t* out_updated /*[B]*/, int32_t* out_iters /*[B]*/,
              float* trace /*[B,iters,8] or NULL: success,min_iou,search[6]*/, int32_t* status /*[1]*/, void* stream);

/* Per-handle options.  BF_OPT_REFINE_CONCURRENT = 1: the caller drives several handles concurrently on their own streams
 * (independent sequences, bench.py --workload c5); small bf_refine calls then use 256-thread CTAs of the 80-register
 * instantiation, which can share an SM with another stream's kernels, instead of the shape that minimises the latency of
 * a lone call (+11 % keyframes/s with 8 concurrent sequences per GPU).  Results are identical.  No reference counterpart. */
enum { BF_OPT_REFINE_CONCURRENT = 1 };
int bf_set_option(bf_handle* h, int key, int value);

/* Diagnostic: how the last bf_refine call of this handle was launched: kernel instantiation * 1000000 (0 = latency
 * regime, 1 = mid, 2 = saturated / compact code) + cluster size * 1000 + block size.  No reference counterpart. */
int bf_refine_last_launch(bf_handle* h);

/* BoxFusion.evaluate_iou (box_fusion.py:413-461): one fitness vector for one box (test/diagnostic entry). */
int bf_evaluate_iou(bf_handle* h, const float* pst /*[P,6]*/, int P, const float* box6 /*[6]*/, const float* rot9,
                    const float* uv /*[V,16]*/, const float* poses /*[V,16]*/, int V, const float* search6,
                    const bf_refine_cfg* cfg /*host*/, float* fitness /*[P]*/, void* stream);

/* ---- Device-resident map store ------------------------------------------------
 * The state demo.py keeps in Python containers across keyframes (all_pred_box, per_frame_ins, BoxManager's
 * fusion_list / fusion_flag / already_fusion; demo.py:72-83, 243-327) held in caller-owned device buffers, with the
 * host bookkeeping around bf_nms3d / bf_refine done by kernels.  boxfusion_b200/engine.py drives them. */
typedef struct {            /* one row per map box (`all_pred_box`), capacity rows each */
    float* tensor; float* R; float* scores; float* box2d; float* projxy; float* pose; float* uv; float* valid;
    int32_t* init_id; int32_t* frame_id; int32_t* fl /*[cap,BF_FUSION_CAP]*/; int32_t* flen;
} bf_map_buffers;
typedef struct {            /* one row per observation ever made (`per_frame_ins`) */
    float* tensor; float* R; float* scores; float* uv; float* pose;
} bf_store_buffers;
typedef struct {            /* BoxManager.already_fusion */
    int32_t* lists /*[cap,BF_FUSION_CAP]*/; int32_t* len; unsigned long long* hash; int32_t* count /*[1]*/; int32_t cap;
} bf_fused_table;

/* demo.py:216-221, 243/248, 253-254: lift + project the keyframe's n detections and append them to map rows [N,N+n)
 * and store rows [M,M+n).  packed = tensor_cam[n,6] R_cam[n,9] scores[n] box2d[n,4] projxy[n,2] pose[16] pose_inv[16]. */
int bf_engine_ingest(bf_handle* h, const float* packed, int n, float fx, float fy, float cx, float cy, float W, float H,
                     int frame_id, int box_count, int N, int M, int D, const bf_map_buffers* mp /*host*/,
                     const bf_store_buffers* st /*host*/, int32_t* fflag, void* stream);
/* instances.py:411-490 + box_manager.py:90-129 on the keep/success flags bf_nms3d produced; also applies valid_num += 1
 * of nms_3d (instances.py:72-73).  info[0] <- 1 if any new box survived nms_3d (demo.py:269). */
int bf_engine_corr(bf_handle* h, const bf_map_buffers* mp /*host*/, const float* store_poses, int32_t* fflag, int N_glo, int n,
                   int32_t* keep, const int32_t* success, const float* pose_inv_np /*[16]*/, float fx, float fy, float cx,
                   float cy, float W, float H, float small_size, float small_plus, double threshold, float translation_gap,
                   float rotation_gap, int32_t* info, int32_t* status, void* stream);
/* `all_pred_box[keep_idx]` + `box_manager.update(keep_idx)` (demo.py:292, 325-327): stable compaction of every map
 * field from `from` into `to`; info[1] <- new row count. */
int bf_engine_compact(bf_handle* h, const int32_t* keep, int N, const bf_map_buffers* from /*host*/,
                      const bf_map_buffers* to /*host*/, int32_t* info, void* stream);
/* box_fusion.py:631-635: rows with >= 3 views whose view set was not fused before -> todo[], CSR for bf_refine;
 * info[2] = B, info[3] = sum of views, info[4] = max views, info[5] = status. */
int bf_engine_select(bf_handle* h, const bf_map_buffers* mp /*host*/, const bf_fused_table* ft /*host*/, int32_t* info,
                     int32_t* todo, int32_t* offsets, int32_t* view_index, void* stream);
/* box_fusion.py:716-724: write bf_refine's rows into the map, set fusion flags, extend already_fusion. */
int bf_engine_apply(bf_handle* h, const bf_map_buffers* mp /*host*/, const bf_fused_table* ft /*host*/, int32_t* fflag,
                    const int32_t* info, const int32_t* todo, const float* out, const int32_t* upd, int32_t* status, void* stream);

/* ---- Detection pre-filters in one pass (demo.py:138-148) -------------------------
 * score >= score_thresh, BoxManager.check_uv_bounds (box_manager.py:217-225, uv_ratio is the Python double),
 * check_floor_mask (:227-237), check_large_mask (:239-245).  flags: bit0 score, bit1 uv, bit2 floor, bit3 large;
 * keep[i] = (flags[i] == 0). */
int bf_detection_filter(bf_handle* h, const float* xyzlhw /*[n,6]*/, const float* proj_xy /*[n,2]*/, const float* scores /*[n]*/,
                        int n, float score_thresh, int use_uv, double uv_ratio, float W, float H, int use_floor, float floor_ratio,
                        int use_large, float size_max, int32_t* flags /*[n]*/, int32_t* keep /*[n]*/, void* stream);

/* Diagnostic: measured FP32 FMA throughput (TFLOP/s) of the device - the denominator of the FP32-pipe
 * roofline bench.py reports.  Synchronous; outputs are HOST pointers. */
int bf_probe_fp32(bf_handle* h, int iters, double* tflops_out /*host*/, float* ms_out /*host or NULL*/);

#ifdef __cplusplus
}
#endif
#endif /* BOXFUSION_B200_H */
