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

/* A2 / A15 for a keyframe whose detections share ONE camera pose (demo.py:216 np.repeat-s it): the pose (A2) or its inverse
 * (A15, torch.linalg.inv of the pose, taken by the host binding like the reference does at instances.py:350) is HOST memory,
 * 16 floats, passed to the kernel by value - no [n,16] upload - and A15 generates the corners itself (one kernel instead of
 * bf_box_corners + bf_project_boxes).  Bit-identical to the per-row entries above. */
int bf_transform2world_pose(bf_handle* h, float* xyzlhw /*[N,6]*/, float* R /*[N,9]*/, const float* pose /*host [16]*/, int N, void* stream);
int bf_project_boxes_pose(bf_handle* h, const float* xyzlhw /*[N,6] world*/, const float* R /*[N,9]*/, int N,
                          const float* pose_inv /*host [16]*/, float fx, float fy, float cx, float cy, float W, float H,
                          float* uv /*[N,8,2]*/, void* stream);

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
 *   status[1]: set to BF_ERR_CAPACITY if a fusion list overflowed BF_FUSION_CAP (or, for N > 16384, more than 8192
 *   over-threshold pairs were found); STICKY - never cleared by the library, the caller zeroes it.
 * IoU uses `mode`; suppression is `iou > iou_threshold` evaluated in float64. */
int bf_nms3d(bf_handle* h, const float* corners /*[N,8,3]*/, const float* centers /*[N,3]*/, int N,
             const int32_t* order, const int32_t* init_id, const float* poses, int M,
             int32_t* fusion_list, int32_t* fusion_len, int32_t* fusion_flag,
             double iou_threshold, float translation_gap, float rotation_gap_deg, float center_gap, int mode,
             int32_t* keep, int32_t* success, int32_t* status, void* stream);

/* The two halves of bf_nms3d as separate entries, for the row-sharded NMS of several GPUs (SURVEY.md section 8(e) axis 3):
 * bf_nms3d_edges finds the over-threshold pairs (a, b), a in [row_begin, row_end), a < b, of the pair triangle and writes one
 * key (rank_lo << 32 | rank_hi, ranks in score order) per pair into edges[0..edge_cap) (unused slots = ~0; more pairs than
 * edge_cap <= 8192 -> status BF_ERR_CAPACITY); the ranks all_gather their edge lists (a few KB) and every rank runs
 * bf_nms3d_greedy - the serial greedy scan of nms_3d plus record() - over the concatenation (n_slots <= 8192 slots, empty
 * ones ignored).  Together they compute exactly bf_nms3d. */
int bf_nms3d_edges(bf_handle* h, const float* corners /*[N,8,3]*/, int N, const int32_t* order, int row_begin, int row_end,
                   double iou_threshold, int mode, unsigned long long* edges /*[edge_cap]*/, int edge_cap, int32_t* status, void* stream);
int bf_nms3d_greedy(bf_handle* h, const unsigned long long* edges /*[n_slots]*/, int n_slots, const float* centers /*[N,3]*/, int N,
                    const int32_t* order, const int32_t* init_id, const float* poses, int32_t* fusion_list, int32_t* fusion_len,
                    int32_t* fusion_flag, float translation_gap, float rotation_gap_deg, float center_gap,
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
 * order[r] = index of the r-th highest score; equal scores keep ascending index (a stable descending sort, what
 * torch.argsort(descending=True, stable=True) gives; NumPy's own argsort is unstable for exact ties).
 * NaN scores sort first, -0 == +0.  N <= 8192: one CTA, bitonic network in shared memory; larger N (up to 65536):
 * every box is ranked by counting the smaller keys (tiles through shared memory, all SMs). */
int bf_score_order(bf_handle* h, const float* scores /*[N]*/, int N, int32_t* order /*[N]*/, void* stream);

/* ---- A8  BoxManager.compute_pose_disparity (box_manager.py:168-186), batched ------------------- */
int bf_pose_disparity(bf_handle* h, const float* poses /*[M,16]*/, const int32_t* ia, const int32_t* ib, int n,
                      float* baseline /*[n]*/, float* angle_deg /*[n]*/, void* stream);

/* ---- A16-A22  BoxFusion.boxfusion optimiser (box_fusion.py:264-405, 413-600, 651-721) ----------
 * Refines B map boxes independently; box b fuses the observations view_index[view_offsets[b] ..
 * view_offsets[b+1]) (CSR, int32) of the per-frame store.  All `iters` optimiser iterations run
 * inside one launch.  Float32 arithmetic with the same operation order as the reference kernel
 * (compiled without FMA contraction), line intersections in float64, global box state in float64.
 * status[1] is set to BF_ERR_CAPACITY when a box has fewer than 1 / more than max_views views (that box is skipped)
 * or an intersection polygon exceeded the reference's own 36-candidate buffer; sticky, the caller zeroes it.
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
enum { BF_OPT_REFINE_CONCURRENT = 1,
       /* value G > 0: stand-alone bf_refine calls use the engine's launch shape - at most G persistent clusters of 16 CTAs
        * x 512 threads, each looping over boxes - instead of one cluster per box (tests; 0 restores the default) */
       BF_OPT_REFINE_PERSISTENT = 2 };
int bf_set_option(bf_handle* h, int key, int value);

/* Diagnostic: how the last bf_refine call of this handle was launched: kernel instantiation * 1000000 (0 = latency
 * regime, 1 = mid, 2 = saturated / compact code) + cluster size * 1000 + block size.  No reference counterpart. */
int bf_refine_last_launch(bf_handle* h);

/* Diagnostic: number of (particle, view) evaluations since the last reset that the latency instantiations of bf_refine /
 * bf_evaluate_iou redid with plain IEEE divisions because a division operand left the exponent window of their branch-free
 * ones (zero on real data; tests construct such operands).  Synchronises the device; -1 on error.  No reference counterpart. */
long long bf_debug_cold_redos(bf_handle* h, int reset);

/* Diagnostic (tools/eval_profile.py, csrc/bf_debug.cu): runs bf_refine's one-evaluation-per-thread loop for ONE box and ONE
 * optimiser state with clock64() ticks compiled into the evaluation and returns, per warp, the cycles spent in each of its
 * sections.  state21 = box6[6], search[6], rot[9]; cycles = [64][16] int64; grid * threads / 32 <= 64; roll != 0 selects the
 * compact instantiation.  All pointers are device pointers.  No reference counterpart. */
int bf_debug_eval_profile(const float* pst, int P, int PB, const float* state21, const float* poses, const float* uv, int V,
                          const float* intr6, int grid, int threads, int roll, int reps, float* out, long long* cycles,
                          void* stream);

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

/* The engine: one keyframe of demo.py:200-327 per call, everything on the device.
 *
 * Round 2 replaced the per-kernel engine entries (ingest / corr / compact / select / apply, driven from Python with a
 * mid-step 32-byte read-back) by ONE call per keyframe.  Every size that changes per keyframe (map rows N, store rows M,
 * detections n, boxes to refine B, the intrinsics) lives in device memory (`bf_engine_state`, the keyframe header), the
 * kernels read it themselves and run on fixed launch shapes, so the whole keyframe is ONE CUDA graph captured once per
 * engine - ingest, 3-D NMS with record(), small-object correspondence, compaction, check_valid_num, selection,
 * particle refinement, write-back - replayed with identical parameters: no host decision, no read-back and no torch
 * call inside a keyframe.  The reference-shaped API (boxfusion_b200/instances.py etc.) replays the same phases as
 * separate graphs and reads the keep flags in between, because spatial_association / correspondence_association
 * return them to the caller (demo.py:262-289).
 *
 * The engine BORROWS the caller's map / store / fused buffers for its lifetime (they stay readable by the caller
 * between steps: snapshot, export) and owns everything else, including a private bf_handle whose scratch is frozen
 * once the graphs are captured.  One engine per sequence; engines are independent (own stream, own scratch). */
typedef struct bf_engine bf_engine;

typedef struct {
    int32_t map_capacity;       /* rows of each map buffer set (<= 65536)                                   */
    int32_t store_capacity;     /* rows of the per-frame store / fusion_flag                                */
    int32_t max_det;            /* detections per keyframe bound (demo.py: topk_per_image = 100)            */
    int32_t iou_mode;           /* BF_IOU_SAMPLED_REF | BF_IOU_ANALYTIC                                      */
    double nms_threshold;       /* box_fusion.nms_threshold                                                 */
    double small_threshold;     /* association.small_threshold                                              */
    float translation_gap, rotation_gap, center_gap;    /* association.* , 0.5 (box_manager.py:55)          */
    float small_size, small_plus;                         /* box_fusion.small_size and float32(small_size + 0.1) */
    int32_t use_fusion;         /* box_fusion.use                                                           */
    int32_t check_valid;        /* box_fusion.check_valid (demo.py:297-298)                                 */
    int32_t gap;                /* data.gap in FRAMES, compared with frame ids (box_manager.py:151-166)     */
    int32_t use_graph;          /* 1: replay captured CUDA graphs; 0: issue the same launches eagerly       */
    int32_t concurrent;         /* 1: several engines share the GPU on their own streams: the refinement uses 256-thread CTAs
                                   (three per SM) instead of the shape that minimises the latency of a lone keyframe          */
    bf_refine_cfg refine;       /* optimiser constants; the intrinsics fields are ignored (per keyframe)    */
    const float* pst;           /* particle template [P,6], device, borrowed                                */
    int32_t P;
} bf_engine_cfg;

typedef struct {                /* caller-owned device buffers, borrowed for the engine's lifetime */
    bf_map_buffers map[2];      /* ping-pong for the compactions; state.cur says which one is current */
    bf_store_buffers store;
    int32_t* fusion_flag;       /* [store_capacity] BoxManager.fusion_flag (never re-indexed, like the reference) */
    bf_fused_table fused;
} bf_engine_buffers;

typedef struct {                /* the engine's counters, device-resident; bf_engine_read_state copies them out */
    int32_t N;                  /* map rows (len(all_pred_box))                                             */
    int32_t M;                  /* store rows (= box_count = len(per_frame_ins) = len(fusion_flag))         */
    int32_t cur;                /* which map buffer set is current                                          */
    int32_t steps;              /* keyframes with detections processed                                      */
    int32_t n;                  /* detections of the keyframe in flight / last processed                    */
    int32_t Nall;               /* N + n while the keyframe is in flight                                    */
    int32_t Nnms;               /* Nall, or 0 on the first keyframe (no association, demo.py:228-243)       */
    int32_t first;              /* the keyframe in flight is the first one                                  */
    int32_t any_new;            /* a new box survived nms_3d (demo.py:269)                                  */
    int32_t Nnew;               /* rows after the last compaction                                           */
    int32_t B, SV, maxV;        /* refinement of the last keyframe: boxes, sum of views, max views          */
    int32_t status[8];          /* sticky BF_ERR_*: [0] nms lists [1] corr lists [2] refine [3] fused table [4] capacity of map/store
                                   [5] views > BF_MAX_VIEWS [6] IoU work list overflow [7] reserved */
    int32_t refine_boxes_total; /* running sums for accounting */
    int32_t refine_views_total;
    int32_t pad[9];
} bf_engine_state;              /* 32 x int32 */

/* Keyframe as the host hands it over: one packed float32 buffer
 *   [0] n (as int32 bits) [1] frame id (int32 bits; demo.py:217) [2..7] fx fy cx cy W H
 *   [8..23] pose (camera->world) [24..39] torch.linalg.inv(pose) (instances.py:350) [40..55] np.linalg.inv(pose) (instances.py:680)
 *   then tensor_cam[n,6] R_cam[n,9] scores[n] box2d[n,4] projxy[n,2]. */
#define BF_KF_HEADER 56
#define BF_KF_ROW 22

int bf_engine_create(int device, const bf_engine_cfg* cfg, const bf_engine_buffers* bufs, bf_engine** out);
void bf_engine_destroy(bf_engine* e);
const char* bf_engine_last_error(bf_engine* e);
/* new sequence in the same buffers */
int bf_engine_reset(bf_engine* e, void* stream);
/* One keyframe, asynchronous.  `phases` selects what runs (0 = the whole keyframe as the engine's configuration defines
 * it: one graph launch):
 *   bit 0  ingest: copy `packed` (HOST memory, BF_KF_HEADER + BF_KF_ROW * n floats; consumed before the call returns) to
 *          the device, lift + project + append                       (demo.py:216-221, 243/248, 253-254)
 *   bit 1  spatial association: corners, score order, 3-D NMS with record()        (demo.py:262)
 *   bit 2  correspondence association incl. record_corr                           (demo.py:273-289)
 *   bit 3  all_pred_box[keep_idx] + box_manager.update                            (demo.py:292, 325-327)
 *   bit 4  BoxManager.check_valid_num                                             (demo.py:297-298)
 *   bit 5  BoxFusion.boxfusion: selection, particle refinement, write-back         (demo.py:304-305)
 *   bit 6  close the keyframe (row counters for the next one)
 * The bits of one keyframe must be issued in ascending order, on one stream per engine.  The reference-shaped API issues
 * them call by call and reads the keep flags in between (bf_engine_read_flags); the engine's own step passes 0. */
int bf_engine_step(bf_engine* e, const float* packed /*host*/, int n, int phases, void* stream);
/* Same with the packed keyframe already in device memory (no host copy). */
int bf_engine_step_device(bf_engine* e, const float* packed_dev, int n, int phases, void* stream);
/* Bit 0 for detections the caller lifted and projected itself (the reference-shaped API: transform2world and
 * project_3d_boxes are separate calls there): row copies from the caller's device tensors; `header` (HOST,
 * BF_KF_HEADER floats) carries n, frame id, intrinsics, pose and np.linalg.inv(pose). */
int bf_engine_ingest_world(bf_engine* e, const float* header /*host*/, const float* tensor_w /*[n,6]*/, const float* R_w /*[n,9]*/,
                           const float* scores /*[n]*/, const float* box2d /*[n,4]*/, const float* projxy /*[n,2]*/,
                           const float* uv /*[n,16]*/, int n, void* stream);
/* The caller filled the first N map rows / M store rows (and the fused table) itself - state imported from the
 * reference-shaped containers: set the row counters accordingly (between keyframes only). */
int bf_engine_set_counts(bf_engine* e, int N, int M, void* stream);
/* Copies the state words to host memory and waits for the stream. */
int bf_engine_read_state(bf_engine* e, bf_engine_state* out /*host*/, void* stream);
/* keep / success flags of the keyframe in flight (first `count` rows) and the state words to HOST memory in one
 * synchronisation (any pointer may be NULL). */
int bf_engine_read_flags(bf_engine* e, int32_t* keep /*host*/, int32_t* success /*host*/, int count, bf_engine_state* state /*host*/,
                         void* stream);
/* Diagnostic read-back (synchronises) of engine-owned per-keyframe results: which = 0 refine iterations per box,
 * 1 CSR view offsets of the refined boxes, 2 map rows selected for refinement, 3 their `updated` flags. */
int bf_engine_read_i32(bf_engine* e, int which, int32_t* out /*host*/, int count, void* stream);
/* Run-ahead for the reference-shaped API (bits 1..6 of a keyframe whose bit 0 was issued): the NMS phase, an asynchronous copy
 * of the keep / success flags of the first `rows` map rows + the state words to pinned memory (slot 0), a rollback snapshot, the
 * correspondence phase, a copy of the keep flags + state (slot 1), then the rest of the keyframe - all queued at once, so the
 * GPU works through the keyframe while the host is inside the caller's code between spatial_association and boxfusion.
 * bf_engine_wait_flags waits for a slot's copy and returns HOST pointers into the engine's pinned buffers (valid until the
 * next run-ahead).  bf_engine_rollback restores the state of right after the NMS phase (the caller strayed from demo.py's
 * sequence): bits 2..6 can then be issued again one by one. */
int bf_engine_run_ahead(bf_engine* e, int rows, void* stream);
int bf_engine_wait_flags(bf_engine* e, int slot, int32_t** keep /*host*/, int32_t** success /*host*/, bf_engine_state** state /*host*/);
int bf_engine_rollback(bf_engine* e, void* stream);
/* Device pointers of engine-owned per-keyframe results (valid for the engine's lifetime). */
int bf_engine_pointers(bf_engine* e, int32_t** keep, int32_t** success, bf_engine_state** state_dev, int32_t** refine_iters,
                       int32_t** todo);
/* kernels launched per phase bit (counts[0..6]) and by the whole keyframe (counts[7]) */
int bf_engine_launch_counts(bf_engine* e, int32_t* counts /*[8]*/);

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
/* The same for the FP64 pipe (the sampled IoU's plane tests are float64): denominator of the `iou` block's roofline. */
int bf_probe_fp64(bf_handle* h, int iters, double* tflops_out /*host*/, float* ms_out /*host or NULL*/);

#ifdef __cplusplus
}
#endif
#endif /* BOXFUSION_B200_H */
