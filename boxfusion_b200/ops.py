"""Thin functional layer over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Every function here is a single call into libboxfusion_sm100.so (see include/boxfusion_b200.h for the
reference interface each entry replaces).  Inputs living on the host are copied to the device; there
is no alternative code path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import IOU_ANALYTIC, IOU_SAMPLED_REF, RefineCfg, handle, ptr

FUSION_CAP = 32
MAX_VIEWS = 64

# kernels launched by one call of each entry point (the library's own __global__ functions)
KERNELS_PER_CALL = {"bf_box_corners": 1, "bf_transform2world": 1, "bf_project_boxes": 1, "bf_iou3d_matrix": 4,
                    "bf_nms3d": 5, "bf_corr2d": 2, "bf_pose_disparity": 1, "bf_refine": 1, "bf_evaluate_iou": 1,
                    "bf_detection_filter": 1, "bf_score_order": 1, "bf_points_in_hull": 1, "bf_engine_ingest_world": 1,
                    "bf_transform2world_pose": 1, "bf_project_boxes_pose": 1,
                    "bf_nms3d_edges": 5, "bf_nms3d_greedy": 1}


class Profile:
    """Launch counting (always on) and optional CUDA-event timing of every C-ABI call (bench.py)."""
    launches = 0
    calls = {}
    h2d_bytes = 0        # host->device bytes moved by the API (counted from the tensors copied)
    d2h_bytes = 0        # device->host bytes read back by the API
    timing = False
    timing_names = None  # None = time every entry; else the set of entry names to time
    events = {}          # name -> list of (start_event, end_event, meta)

    @classmethod
    def reset(cls, timing=False, names=None):
        cls.launches, cls.calls, cls.events, cls.timing, cls.timing_names = 0, {}, {}, timing, names
        cls.h2d_bytes = cls.d2h_bytes = 0

    @classmethod
    def elapsed_ms(cls):
        """name -> list of (ms, meta); call after torch.cuda.synchronize()."""
        return {k: [(a.elapsed_time(b), m) for a, b, m in v] for k, v in cls.events.items()}


def _call(h, name, fn, *args, meta=None):
    Profile.launches += KERNELS_PER_CALL[name]
    Profile.calls[name] = Profile.calls.get(name, 0) + 1
    if Profile.timing and (Profile.timing_names is None or name in Profile.timing_names):
        st = torch.cuda.current_stream(h.device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        rc = fn(*args)
        b.record(st)
        Profile.events.setdefault(name, []).append((a, b, meta))
    else:
        rc = fn(*args)
    h.check(rc, name)


def dev_tensor(x, dtype, device) -> torch.Tensor:
    """Contiguous tensor of `dtype` on `device`; host data is copied (and counted), CUDA data is used in place."""
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if x.is_cuda:
        if x.dtype == dtype and x.device == device and x.is_contiguous():
            return x                                      # the common case: already resident, no torch dispatch at all
    else:
        Profile.h2d_bytes += x.numel() * x.element_size()
    return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def to_host(t: torch.Tensor) -> np.ndarray:
    """Device->host read of a result (synchronises the stream); counted for the e2e bench."""
    if t.is_cuda:
        Profile.d2h_bytes += t.numel() * t.element_size()
    return t.cpu().numpy()


def _dev(device=None) -> torch.device:
    if device is None or (isinstance(device, torch.device) and device.type != "cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("boxfusion_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _pick_device(*tensors) -> torch.device:
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return _dev()


def box_corners(xyzlhw, R, want_centers: bool = False):
    """GeneralInstance3DBoxes.corners (boxes.py:725-778) -> [N,8,3] (and nms_3d's centres, instances.py:49)."""
    dev = _pick_device(xyzlhw, R)
    t = dev_tensor(xyzlhw, torch.float32, dev).reshape(-1, 6)
    r = dev_tensor(R, torch.float32, dev).reshape(-1, 9)
    n = t.shape[0]
    corners = torch.empty((n, 8, 3), dtype=torch.float32, device=dev)
    centers = torch.empty((n, 3), dtype=torch.float32, device=dev) if want_centers else None
    h = handle(dev)
    _call(h, "bf_box_corners", h.lib.bf_box_corners, h.h, ptr(t), ptr(r), n, ptr(corners), ptr(centers), h.stream())
    return (corners, centers) if want_centers else corners


def transform2world_(xyzlhw: torch.Tensor, R: torch.Tensor, poses) -> None:
    """GeneralInstance3DBoxes.transform2world (boxes.py:825-833), in place on CUDA tensors."""
    assert xyzlhw.is_cuda and R.is_cuda and xyzlhw.is_contiguous() and R.is_contiguous()
    p = dev_tensor(poses, torch.float32, xyzlhw.device).reshape(-1, 16)
    h = handle(xyzlhw.device)
    _call(h, "bf_transform2world", h.lib.bf_transform2world, h.h, ptr(xyzlhw), ptr(R), ptr(p), xyzlhw.shape[0], h.stream())


def shared_pose(cam_pose):
    """The single 4x4 (contiguous float32 numpy, host) when every row of a HOST cam_pose [n,4,4] is the same matrix, else None."""
    if isinstance(cam_pose, torch.Tensor):
        if cam_pose.is_cuda or cam_pose.dtype != torch.float32:
            return None
        cam_pose = cam_pose.numpy()
    a = np.asarray(cam_pose)
    if a.ndim != 3 or a.shape[0] < 1 or a.dtype != np.float32 or not (a == a[0]).all():
        return None
    return np.ascontiguousarray(a[0])


def transform2world_pose_(xyzlhw: torch.Tensor, R: torch.Tensor, pose16: np.ndarray) -> None:
    """transform2world for detections that share one pose: the pose is a kernel parameter (no upload)."""
    h = handle(xyzlhw.device)
    Profile.h2d_bytes += 64
    _call(h, "bf_transform2world_pose", h.lib.bf_transform2world_pose, h.h, ptr(xyzlhw), ptr(R), pose16.ctypes.data, xyzlhw.shape[0], h.stream())


def project_boxes_pose(xyzlhw: torch.Tensor, R: torch.Tensor, pose_inv16: np.ndarray, K, W: float, H: float) -> torch.Tensor:
    """project_3d_boxes for detections that share one pose: corners + projection in one kernel, inverse pose by value."""
    n = xyzlhw.shape[0]
    uv = torch.empty((n, 8, 2), dtype=torch.float32, device=xyzlhw.device)
    h = handle(xyzlhw.device)
    Profile.h2d_bytes += 64
    _call(h, "bf_project_boxes_pose", h.lib.bf_project_boxes_pose, h.h, ptr(xyzlhw), ptr(R), n, pose_inv16.ctypes.data,
          float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]), float(W), float(H), ptr(uv), h.stream())
    return uv


def project_boxes(corners, pose_inv, K, W: float, H: float) -> torch.Tensor:
    """Instances3D.project_3d_boxes (instances.py:333-369) given the inverted poses -> [N,8,2]."""
    dev = _pick_device(corners)
    c = dev_tensor(corners, torch.float32, dev).reshape(-1, 8, 3)
    pi = dev_tensor(pose_inv, torch.float32, dev).reshape(-1, 16)
    n = c.shape[0]
    uv = torch.empty((n, 8, 2), dtype=torch.float32, device=dev)
    h = handle(dev)
    _call(h, "bf_project_boxes", h.lib.bf_project_boxes, h.h, ptr(c), ptr(pi), n, float(K[0][0]), float(K[1][1]), float(K[0][2]),
                                   float(K[1][2]), float(W), float(H), ptr(uv), h.stream())
    return uv


def iou3d_matrix(cornersA, cornersB, mode: int = IOU_SAMPLED_REF, want_counts: bool = False, want_stats: bool = False):
    """calculate_obb_iou / Instances3D.obb_iou for every pair (instances.py:106-125, 573-613) -> float64 [M,N]."""
    dev = _pick_device(cornersA, cornersB)
    a = dev_tensor(cornersA, torch.float32, dev).reshape(-1, 8, 3)
    b = a if cornersB is cornersA else dev_tensor(cornersB, torch.float32, dev).reshape(-1, 8, 3)
    M, N = a.shape[0], b.shape[0]
    iou = torch.empty((M, N), dtype=torch.float64, device=dev)
    counts = torch.empty((M, N, 3), dtype=torch.int32, device=dev) if want_counts else None
    stats = torch.zeros(4, dtype=torch.int64, device=dev) if want_stats else None
    h = handle(dev)
    _call(h, "bf_iou3d_matrix", h.lib.bf_iou3d_matrix, h.h, ptr(a), M, ptr(b), N, int(mode), ptr(iou), ptr(counts), ptr(stats), h.stream())
    out = (iou,)
    if want_counts:
        out += (counts,)
    if want_stats:
        out += (stats,)
    return out if len(out) > 1 else iou


def nms3d(corners, centers, order, init_id, poses, fusion_list, fusion_len, fusion_flag, iou_threshold: float,
          translation_gap: float, rotation_gap: float, center_gap: float = 0.5, mode: int = IOU_SAMPLED_REF):
    """nms_3d + BoxManager.record (instances.py:22-101, box_manager.py:40-88).  All tensors on one CUDA device;
    fusion_list [N,FUSION_CAP] / fusion_len [N] / fusion_flag [N] int32 are updated in place.
    Returns (keep[N] int32 0/1, success[N] int32 0/1, status[1] int32) on the device."""
    dev = corners.device
    n = corners.shape[0]
    keep = torch.empty(n, dtype=torch.int32, device=dev)
    success = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    h = handle(dev)
    _call(h, "bf_nms3d", h.lib.bf_nms3d, h.h, ptr(corners), ptr(centers), n, ptr(order), ptr(init_id), ptr(poses), poses.shape[0],
                           ptr(fusion_list), ptr(fusion_len), ptr(fusion_flag), float(iou_threshold),
                           float(translation_gap), float(rotation_gap), float(center_gap), int(mode),
                           ptr(keep), ptr(success), ptr(status), h.stream())
    return keep, success, status


EDGE_CAP = 8192


def nms3d_edges(corners, order, row_begin: int, row_end: int, iou_threshold: float, mode: int = IOU_SAMPLED_REF, edge_cap: int = EDGE_CAP):
    """Over-threshold pairs of rows [row_begin, row_end) of the pair triangle -> (edges[edge_cap] int64 keys rank_lo << 32 | rank_hi,
    unused slots -1; status[1] int32), on the device (first half of bf_nms3d, for the row-sharded NMS)."""
    dev = corners.device
    n = corners.shape[0]
    edges = torch.empty(edge_cap, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    h = handle(dev)
    _call(h, "bf_nms3d_edges", h.lib.bf_nms3d_edges, h.h, ptr(corners), n, ptr(order), int(row_begin), int(row_end), float(iou_threshold),
          int(mode), ptr(edges), int(edge_cap), ptr(status), h.stream())
    return edges, status


def nms3d_greedy(edges, centers, order, init_id, poses, fusion_list, fusion_len, fusion_flag, translation_gap: float,
                 rotation_gap: float, center_gap: float = 0.5):
    """Greedy scan of nms_3d + BoxManager.record over an edge list (second half of bf_nms3d) -> (keep, success, status)."""
    dev = centers.device
    n = centers.shape[0]
    keep = torch.empty(n, dtype=torch.int32, device=dev)
    success = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    h = handle(dev)
    _call(h, "bf_nms3d_greedy", h.lib.bf_nms3d_greedy, h.h, ptr(edges), int(edges.shape[0]), ptr(centers), n, ptr(order), ptr(init_id),
          ptr(poses), ptr(fusion_list), ptr(fusion_len), ptr(fusion_flag), float(translation_gap), float(rotation_gap),
          float(center_gap), ptr(keep), ptr(success), ptr(status), h.stream())
    return keep, success, status


def corr2d(map_corners, small_mask, pose_inv, K, W: float, H: float, det_xyxy, want_boxes: bool = False):
    """correspondence_association scoring (instances.py:446-468, 643-717) -> (best[n] int32, best_iou[n] f64)."""
    dev = _pick_device(map_corners, det_xyxy)
    mc = dev_tensor(map_corners, torch.float32, dev).reshape(-1, 8, 3)
    sm = dev_tensor(small_mask, torch.int32, dev).reshape(-1)
    pi = dev_tensor(pose_inv, torch.float32, dev).reshape(16)
    det = dev_tensor(det_xyxy, torch.float32, dev).reshape(-1, 4)
    G, n = mc.shape[0], det.shape[0]
    best = torch.empty(n, dtype=torch.int32, device=dev)
    best_iou = torch.empty(n, dtype=torch.float64, device=dev)
    boxes2d = torch.empty((G, 4), dtype=torch.float64, device=dev) if want_boxes else None
    h = handle(dev)
    _call(h, "bf_corr2d", h.lib.bf_corr2d, h.h, ptr(mc), ptr(sm), G, ptr(pi), float(K[0][0]), float(K[1][1]), float(K[0][2]),
                            float(K[1][2]), float(W), float(H), ptr(det), n, ptr(boxes2d), ptr(best), ptr(best_iou),
                            h.stream())
    return (best, best_iou, boxes2d) if want_boxes else (best, best_iou)


def pose_disparity(poses, ia, ib):
    """BoxManager.compute_pose_disparity (box_manager.py:168-186), batched -> (baseline[n], angle_deg[n])."""
    dev = _pick_device(poses)
    p = dev_tensor(poses, torch.float32, dev).reshape(-1, 16)
    a = dev_tensor(ia, torch.int32, dev).reshape(-1)
    b = dev_tensor(ib, torch.int32, dev).reshape(-1)
    n = a.shape[0]
    base = torch.empty(n, dtype=torch.float32, device=dev)
    ang = torch.empty(n, dtype=torch.float32, device=dev)
    h = handle(dev)
    _call(h, "bf_pose_disparity", h.lib.bf_pose_disparity, h.h, ptr(p), ptr(a), ptr(b), n, ptr(base), ptr(ang), h.stream())
    return base, ang


def make_refine_cfg(cfg: dict, K16, img_h: float, img_w: float, beta: float = 0.9, early_stop: bool = True,
                    max_hits: int = 200, iters: Optional[int] = None) -> RefineCfg:
    bf = cfg["box_fusion"]
    ro = bf["random_opt"]
    K16 = np.asarray(K16, dtype=np.float32).reshape(-1)
    return RefineCfg(int(bf["iters"] if iters is None else iters), int(bf["pst_size"]),
                     float(ro["center_init_size"]), float(ro["shape_init_size"]),
                     float(ro["center_scaling_coefficient"]), float(ro["shape_scaling_coefficient"]),
                     float(beta), float(img_h), float(img_w),
                     float(K16[0]), float(K16[2]), float(K16[5]), float(K16[6]), int(max_hits), int(bool(early_stop)), 0, 0)


def refine(pst, per_xyzlhw, per_R, per_scores, per_uv, per_poses, view_offsets, view_index, rcfg: RefineCfg,
           want_trace: bool = False, max_views: int = 0):
    """BoxFusion.boxfusion optimiser for B boxes in one launch (box_fusion.py:651-721).
    Returns (out_xyzlhw[B,6] f32, updated[B] i32, iters[B] i32, trace|None, status[1] i32), all on the device."""
    dev = _pick_device(per_xyzlhw, pst)
    pst = dev_tensor(pst, torch.float32, dev).reshape(-1, 6)
    t = dev_tensor(per_xyzlhw, torch.float32, dev).reshape(-1, 6)
    r = dev_tensor(per_R, torch.float32, dev).reshape(-1, 9)
    s = dev_tensor(per_scores, torch.float32, dev).reshape(-1)
    uv = dev_tensor(per_uv, torch.float32, dev).reshape(-1, 16)
    po = dev_tensor(per_poses, torch.float32, dev).reshape(-1, 16)
    off = dev_tensor(view_offsets, torch.int32, dev).reshape(-1)
    idx = dev_tensor(view_index, torch.int32, dev).reshape(-1)
    B = off.shape[0] - 1
    rcfg.views_total = int(idx.shape[0])
    rcfg.max_views = int(max_views)
    out = torch.empty((B, 6), dtype=torch.float32, device=dev)
    upd = torch.empty(B, dtype=torch.int32, device=dev)
    its = torch.empty(B, dtype=torch.int32, device=dev)
    trace = torch.zeros((B, rcfg.iters, 8), dtype=torch.float32, device=dev) if want_trace else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    h = handle(dev)
    _call(h, "bf_refine", h.lib.bf_refine, h.h, ptr(pst), pst.shape[0], ptr(t), ptr(r), ptr(s), ptr(uv), ptr(po), t.shape[0],
                            ptr(off), ptr(idx), B, ctypes.byref(rcfg), ptr(out), ptr(upd), ptr(its), ptr(trace),
                            ptr(status), h.stream())
    return out, upd, its, trace, status


def points_in_hull(points, corners) -> torch.Tensor:
    """Instances3D.batch_in_convex_hull_3d (instances.py:559-571): bool [n], points against the hull of 8 corners."""
    dev = _pick_device(points, corners)
    c = dev_tensor(np.asarray(corners, dtype=np.float32) if not isinstance(corners, torch.Tensor) else corners, torch.float32, dev).reshape(8, 3)
    p = dev_tensor(np.asarray(points, dtype=np.float64) if not isinstance(points, torch.Tensor) else points, torch.float64, dev).reshape(-1, 3)
    out = torch.empty(p.shape[0], dtype=torch.uint8, device=dev)
    h = handle(dev)
    _call(h, "bf_points_in_hull", h.lib.bf_points_in_hull, h.h, ptr(c), ptr(p), p.shape[0], ptr(out), h.stream())
    return out.to(torch.bool)


ORDER_MAX = 65536


def score_order(scores) -> torch.Tensor:
    """`scores.argsort()[::-1]` of nms_3d (instances.py:52) as a stable descending sort on the device -> int32 [N]."""
    dev = _pick_device(scores)
    s = dev_tensor(scores, torch.float32, dev).reshape(-1)
    n = s.shape[0]
    if n > ORDER_MAX:
        raise RuntimeError(f"score_order: {n} boxes; the device sort handles up to {ORDER_MAX}")
    order = torch.empty(n, dtype=torch.int32, device=dev)
    h = handle(dev)
    _call(h, "bf_score_order", h.lib.bf_score_order, h.h, ptr(s), n, ptr(order), h.stream())
    return order


def last_refine_launch(device=None) -> dict:
    """Diagnostic: kernel instantiation / cluster size / block size of the handle's last bf_refine launch."""
    h = handle(device)
    v = int(h.lib.bf_refine_last_launch(h.h))
    return {"variant": ("latency", "mid", "saturated")[v // 1000000], "cluster": (v % 1000000) // 1000, "threads": v % 1000}


def cold_redos(device=None, reset: bool = True) -> int:
    """Diagnostic: evaluations redone with plain divisions since the last reset (bf_debug_cold_redos)."""
    h = handle(device)
    v = int(h.lib.bf_debug_cold_redos(h.h, int(reset)))
    if v < 0:
        raise RuntimeError("bf_debug_cold_redos failed")
    return v


def evaluate_iou(pst, box6, rot9, uv, poses, search6, rcfg: RefineCfg) -> torch.Tensor:
    """BoxFusion.evaluate_iou (box_fusion.py:413-461) -> fitness[P] float32 on the device."""
    dev = _pick_device(pst, uv)
    pst = dev_tensor(pst, torch.float32, dev).reshape(-1, 6)
    b = dev_tensor(np.asarray(box6, dtype=np.float32) if not isinstance(box6, torch.Tensor) else box6, torch.float32, dev).reshape(6)
    r = dev_tensor(rot9, torch.float32, dev).reshape(9)
    u = dev_tensor(uv, torch.float32, dev).reshape(-1, 16)
    p = dev_tensor(poses, torch.float32, dev).reshape(-1, 16)
    s = dev_tensor(search6, torch.float32, dev).reshape(6)
    fit = torch.empty(pst.shape[0], dtype=torch.float32, device=dev)
    h = handle(dev)
    _call(h, "bf_evaluate_iou", h.lib.bf_evaluate_iou, h.h, ptr(pst), pst.shape[0], ptr(b), ptr(r), ptr(u), ptr(p), u.shape[0], ptr(s),
                                  ctypes.byref(rcfg), ptr(fit), h.stream())
    return fit


def probe_fp32(iters: int = 4096, device=None) -> float:
    """Measured FP32 FMA throughput (TFLOP/s) of the device: denominator of the FP32-pipe roofline."""
    h = handle(_dev(device))
    tf = ctypes.c_double(0.0)
    ms = ctypes.c_float(0.0)
    h.check(h.lib.bf_probe_fp32(h.h, int(iters), ctypes.byref(tf), ctypes.byref(ms)), "bf_probe_fp32")
    return float(tf.value)


def probe_fp64(iters: int = 2048, device=None) -> float:
    """Measured FP64 FMA throughput (TFLOP/s) of the device: denominator of the sampled-IoU roofline."""
    h = handle(_dev(device))
    tf = ctypes.c_double(0.0)
    ms = ctypes.c_float(0.0)
    h.check(h.lib.bf_probe_fp64(h.h, int(iters), ctypes.byref(tf), ctypes.byref(ms)), "bf_probe_fp64")
    return float(tf.value)


def detection_filter(xyzlhw, proj_xy, scores, W: float, H: float, score_thresh: float, uv_ratio: Optional[float] = None,
                     floor_ratio: Optional[float] = None, size_max: Optional[float] = None):
    """demo.py:138-148 in one kernel: (keep[n] bool, flags[n] int32: bit0 score, bit1 uv bounds, bit2 floor, bit3 large)."""
    dev = _pick_device(xyzlhw, proj_xy, scores)
    t = dev_tensor(xyzlhw, torch.float32, dev).reshape(-1, 6)
    p = dev_tensor(proj_xy, torch.float32, dev).reshape(-1, 2)
    s = dev_tensor(scores, torch.float32, dev).reshape(-1)
    n = t.shape[0]
    flags = torch.empty(n, dtype=torch.int32, device=dev)
    keep = torch.empty(n, dtype=torch.int32, device=dev)
    h = handle(dev)
    _call(h, "bf_detection_filter", h.lib.bf_detection_filter, h.h, ptr(t), ptr(p), ptr(s), n, float(score_thresh),
          int(uv_ratio is not None), float(uv_ratio or 0.0), float(W), float(H), int(floor_ratio is not None),
          float(floor_ratio or 0.0), int(bool(size_max)), float(size_max or 0.0), ptr(flags), ptr(keep), h.stream())
    return keep.bool(), flags
