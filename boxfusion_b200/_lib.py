"""ctypes binding of libboxfusion_sm100.so (include/boxfusion_b200.h).

There is no fallback of any kind: if the shared library is missing, or no sm_100 device is
present, importing callers get a RuntimeError the first time they touch the hot path.
PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BOXFUSION_B200_LIB: developer knob for kernel-tuning builds of the same library (tools/build_variants.py)
LIB_PATH = os.environ.get("BOXFUSION_B200_LIB") or os.path.join(_HERE, "lib", "libboxfusion_sm100.so")

BF_OK, BF_ERR_INVALID_ARG, BF_ERR_CUDA, BF_ERR_CAPACITY = 0, -1, -2, -3
IOU_SAMPLED_REF, IOU_ANALYTIC = 0, 1
OPT_REFINE_CONCURRENT = 1
OPT_REFINE_PERSISTENT = 2
_ERR = {-1: "BF_ERR_INVALID_ARG", -2: "BF_ERR_CUDA", -3: "BF_ERR_CAPACITY"}

_vp, _i32, _f32, _f64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double


class RefineCfg(ctypes.Structure):
    """bf_refine_cfg (include/boxfusion_b200.h)."""
    _fields_ = [("iters", ctypes.c_int32), ("pst_size", ctypes.c_int32),
                ("center_init", _f32), ("shape_init", _f32), ("center_scale", _f32), ("shape_scale", _f32),
                ("beta", _f64), ("img_h", _f32), ("img_w", _f32),
                ("fx", _f32), ("cx", _f32), ("fy", _f32), ("cy", _f32),
                ("max_hits", ctypes.c_int32), ("early_stop", ctypes.c_int32), ("views_total", ctypes.c_int32), ("max_views", ctypes.c_int32)]


class MapBuffers(ctypes.Structure):
    """bf_map_buffers"""
    _fields_ = [(k, _vp) for k in ("tensor", "R", "scores", "box2d", "projxy", "pose", "uv", "valid", "init_id", "frame_id",
                                   "fl", "flen")]


class StoreBuffers(ctypes.Structure):
    """bf_store_buffers"""
    _fields_ = [(k, _vp) for k in ("tensor", "R", "scores", "uv", "pose")]


class FusedTable(ctypes.Structure):
    """bf_fused_table"""
    _fields_ = [("lists", _vp), ("len", _vp), ("hash", _vp), ("count", _vp), ("cap", ctypes.c_int32)]


class EngineCfg(ctypes.Structure):
    """bf_engine_cfg"""
    _fields_ = [("map_capacity", ctypes.c_int32), ("store_capacity", ctypes.c_int32), ("max_det", ctypes.c_int32),
                ("iou_mode", ctypes.c_int32), ("nms_threshold", _f64), ("small_threshold", _f64),
                ("translation_gap", _f32), ("rotation_gap", _f32), ("center_gap", _f32),
                ("small_size", _f32), ("small_plus", _f32),
                ("use_fusion", ctypes.c_int32), ("check_valid", ctypes.c_int32), ("gap", ctypes.c_int32), ("use_graph", ctypes.c_int32),
                ("concurrent", ctypes.c_int32), ("refine", RefineCfg), ("pst", _vp), ("P", ctypes.c_int32)]


class EngineBuffers(ctypes.Structure):
    """bf_engine_buffers"""
    _fields_ = [("map", MapBuffers * 2), ("store", StoreBuffers), ("fusion_flag", _vp), ("fused", FusedTable)]


class EngineState(ctypes.Structure):
    """bf_engine_state (32 x int32)"""
    _fields_ = [(k, ctypes.c_int32) for k in ("N", "M", "cur", "steps", "n", "Nall", "Nnms", "first", "any_new", "Nnew", "B", "SV", "maxV")] + \
               [("status", ctypes.c_int32 * 8), ("refine_boxes_total", ctypes.c_int32), ("refine_views_total", ctypes.c_int32),
                ("pad", ctypes.c_int32 * 9)]


KF_HEADER, KF_ROW = 56, 22
PH_INGEST, PH_NMS, PH_CORR, PH_COMPACT, PH_VALID, PH_FUSE, PH_FINISH = (1 << i for i in range(7))
_ESP = ctypes.POINTER(EngineState)

# name -> (restype, argtypes); every symbol declared in include/boxfusion_b200.h
PROTOTYPES = {
    "bf_version": (_i32, []),
    "bf_create": (_i32, [_i32, ctypes.POINTER(_vp)]),
    "bf_destroy": (None, [_vp]),
    "bf_last_error": (ctypes.c_char_p, [_vp]),
    "bf_fusion_cap": (_i32, []),
    "bf_box_corners": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "bf_transform2world": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "bf_project_boxes": (_i32, [_vp, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp]),
    "bf_transform2world_pose": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "bf_project_boxes_pose": (_i32, [_vp, _vp, _vp, _i32, _vp, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp]),
    "bf_iou3d_matrix": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "bf_nms3d": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _f64, _f32, _f32, _f32, _i32,
                        _vp, _vp, _vp, _vp]),
    "bf_nms3d_edges": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _f64, _i32, _vp, _i32, _vp, _vp]),
    "bf_nms3d_greedy": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "bf_corr2d": (_i32, [_vp, _vp, _vp, _i32, _vp, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "bf_points_in_hull": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "bf_score_order": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "bf_pose_disparity": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "bf_refine": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, ctypes.POINTER(RefineCfg),
                         _vp, _vp, _vp, _vp, _vp, _vp]),
    "bf_engine_create": (_i32, [_i32, ctypes.POINTER(EngineCfg), ctypes.POINTER(EngineBuffers), ctypes.POINTER(_vp)]),
    "bf_engine_destroy": (None, [_vp]),
    "bf_engine_last_error": (ctypes.c_char_p, [_vp]),
    "bf_engine_reset": (_i32, [_vp, _vp]),
    "bf_engine_step": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "bf_engine_step_device": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "bf_engine_ingest_world": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "bf_engine_set_counts": (_i32, [_vp, _i32, _i32, _vp]),
    "bf_engine_read_state": (_i32, [_vp, _ESP, _vp]),
    "bf_engine_read_flags": (_i32, [_vp, _vp, _vp, _i32, _ESP, _vp]),
    "bf_engine_run_ahead": (_i32, [_vp, _i32, _vp]),
    "bf_engine_wait_flags": (_i32, [_vp, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "bf_engine_rollback": (_i32, [_vp, _vp]),
    "bf_engine_read_i32": (_i32, [_vp, _i32, _vp, _i32, _vp]),
    "bf_engine_pointers": (_i32, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "bf_engine_launch_counts": (_i32, [_vp, ctypes.POINTER(ctypes.c_int32)]),
    "bf_detection_filter": (_i32, [_vp, _vp, _vp, _vp, _i32, _f32, _i32, _f64, _f32, _f32, _i32, _f32, _i32, _f32, _vp, _vp, _vp]),
    "bf_probe_fp32": (_i32, [_vp, _i32, ctypes.POINTER(_f64), ctypes.POINTER(_f32)]),
    "bf_probe_fp64": (_i32, [_vp, _i32, ctypes.POINTER(_f64), ctypes.POINTER(_f32)]),
    "bf_set_option": (_i32, [_vp, _i32, _i32]),
    "bf_refine_last_launch": (_i32, [_vp]),
    "bf_debug_cold_redos": (ctypes.c_longlong, [_vp, _i32]),
    "bf_debug_eval_profile": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "bf_evaluate_iou": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, ctypes.POINTER(RefineCfg), _vp, _vp]),
}

_LIB: Optional[ctypes.CDLL] = None
_HANDLES: Dict[int, "Handle"] = {}


def load_library() -> ctypes.CDLL:
    """dlopen the in-tree library and bind every exported symbol (no GPU needed for this step)."""
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m boxfusion_b200.build` "
                "(boxfusion_b200 has no CPU or PyTorch fallback for the fusion hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export the symbol
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


class Handle:
    """A bf_handle (device + scratch).  `handle(device)` shares one per device; engines that run concurrently on their
    own streams create private ones."""

    def __init__(self, device: int):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("boxfusion_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        self.device = device
        torch.cuda.init()
        with torch.cuda.device(device):
            torch.zeros(1, device=f"cuda:{device}")    # make sure the primary context exists
            h = _vp()
            rc = self.lib.bf_create(device, ctypes.byref(h))
            if rc != BF_OK:
                raise RuntimeError(f"bf_create({device}) failed: {_ERR.get(rc, rc)}: "
                                   f"{self.lib.bf_last_error(None).decode()}")
        self.h = h

    def check(self, rc: int, what: str):
        if rc != BF_OK:
            raise RuntimeError(f"{what}: {_ERR.get(rc, rc)}: {self.lib.bf_last_error(self.h).decode()}")

    def stream(self) -> int:
        return _raw_stream(self.device)


# torch's current stream of a device as a raw cudaStream_t (the private accessor costs ~1 us, the Stream object ~10 us; this is
# on every call of the per-keyframe path)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (lambda d: torch.cuda.current_stream(d).cuda_stream)


def handle(device=None) -> Handle:
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    device = int(device)
    if device not in _HANDLES:
        _HANDLES[device] = Handle(device)
    return _HANDLES[device]


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device pointer arguments must be contiguous CUDA tensors"
    return t.data_ptr()
