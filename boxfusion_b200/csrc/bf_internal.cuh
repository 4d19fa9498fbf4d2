// Cross-translation-unit plumbing of libboxfusion_sm100.so: the host-side launch sequences that several entry points
// share (the stand-alone C-ABI entries and the graph-captured engine step).  Not part of the public interface.
#pragma once
#include "bf_common.cuh"

// context of BoxManager.record / record_corr as replayed on the device (bf_assoc.cu: bf_record_one)
struct bf_record_ctx {
    const int32_t* order; const int32_t* init_id; const float* poses;
    const float* centers;                 // nullptr: record_corr (no centre-distance term, box_manager.py:100-102)
    int32_t* fl; int32_t* flen; int32_t* fflag; int32_t* keep; int32_t* status;
    float translation_gap, rotation_gap, center_gap;
    float* valid_num;                     // optional: `valid_num[i] += 1` for every head that suppressed something (instances.py:72-73)
};

// K1 (bf_iou3d.cu)
int bf_iou3d_run(bf_handle* h, const float* cornersA, bf_dimref Md, const float* cornersB, bf_dimref Nd, int triangle, int a_off, int mode,
                 double* iou, int32_t* counts, int64_t* stats, double thr, const int32_t* rank, uint32_t* mask,
                 uint32_t* rowany, unsigned long long* edges, int edge_cap, cudaStream_t st);
int bf_iou3d_overflowed(bf_handle* h, cudaStream_t st, int* overflow);

// K2 (bf_assoc.cu)
int bf_score_order_run(bf_handle* h, const float* scores, bf_dimref Nd, int32_t* order, int32_t* rank, cudaStream_t st);
int bf_nms3d_run(bf_handle* h, const float* corners, const float* centers, bf_dimref Nd, const int32_t* order, int32_t* rank_or_null,
                 const int32_t* init_id, const float* poses, int32_t* fusion_list, int32_t* fusion_len, int32_t* fusion_flag,
                 double iou_threshold, float translation_gap, float rotation_gap_deg, float center_gap, int mode,
                 int32_t* keep, int32_t* success, int32_t* status, float* valid_num_or_null, cudaStream_t st);

// geometry (bf_geometry.cu)
int bf_box_corners_run(bf_handle* h, const float* xyzlhw, const float* R, bf_dimref Nd, float* corners, float* centers, cudaStream_t st);

// K3 (bf_refine.cu): the optimiser with the box count B and the intrinsics either in `cfg` / `B` (host) or in device
// memory (B_dev; intr_dev = fx, fy, cx, cy, W, H), for the captured engine step.
struct bf_refine_dev_args { const int32_t* B_dev; const float* intr_dev; int max_boxes; };
int bf_refine_run(bf_handle* h, const float* pst, int P, const float* per_xyzlhw, const float* per_R,
                  const float* per_scores, const float* per_uv, const float* per_poses,
                  const int32_t* view_offsets, const int32_t* view_index, int B, const bf_refine_cfg* cfg,
                  float* out_xyzlhw, int32_t* out_updated, int32_t* out_iters, float* trace, int32_t* status,
                  const bf_refine_dev_args* dev, cudaStream_t st);
