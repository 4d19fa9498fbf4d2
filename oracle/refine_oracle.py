"""TEST INFRASTRUCTURE ONLY - ctypes view of oracle/refine_oracle.c plus the list-driven outer loop
of `BoxFusion.boxfusion` (reference boxfusion/box_fusion.py:622-724).

Pinned against the reference executed in the build container (tests/test_oracle_golden.py,
tests/golden/make_golden.py).  Never imported by the product path.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Sequence

import numpy as np

from . import build as _build

_FP = ctypes.POINTER(ctypes.c_float)
_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int)


class _Cfg(ctypes.Structure):
    _fields_ = [("iters", ctypes.c_int), ("pst_size", ctypes.c_int),
                ("center_init", ctypes.c_float), ("shape_init", ctypes.c_float),
                ("center_scale", ctypes.c_float), ("shape_scale", ctypes.c_float),
                ("beta", ctypes.c_double), ("img_h", ctypes.c_float), ("img_w", ctypes.c_float),
                ("early_stop", ctypes.c_int)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(_build.build_oracle())
        _LIB.bfo_eval_particle_view.restype = ctypes.c_float
        _LIB.bfo_eval_particle_view.argtypes = [_FP] * 5 + [ctypes.c_float] * 4 + [_FP, ctypes.c_float, ctypes.c_float]
        _LIB.bfo_evaluate.restype = None
        _LIB.bfo_evaluate.argtypes = [_FP, _FP, _FP, ctypes.c_int, ctypes.c_int, _FP, _FP, ctypes.c_int, _FP, _FP,
                                      ctypes.c_float, ctypes.c_float, _FP]
        _LIB.bfo_cal_transform.restype = ctypes.c_int
        _LIB.bfo_cal_transform.argtypes = [_FP, _FP, ctypes.c_int, _FP, _FP, _FP]
        _LIB.bfo_update_pst.restype = None
        _LIB.bfo_update_pst.argtypes = [ctypes.c_float, _FP, ctypes.c_float, ctypes.c_float, _FP]
        _LIB.bfo_init_opt_params.restype = ctypes.c_int
        _LIB.bfo_init_opt_params.argtypes = [_FP, _FP, ctypes.c_int, _DP]
        _LIB.bfo_refine_box.restype = ctypes.c_int
        _LIB.bfo_refine_box.argtypes = [_FP, _FP, _FP, _FP, _FP, ctypes.c_int, _FP, ctypes.c_int, _FP,
                                        ctypes.POINTER(_Cfg), _FP, _IP, _FP]
        _LIB.bfo_target_hull.restype = ctypes.c_int
        _LIB.bfo_target_hull.argtypes = [_FP, _FP]
        _LIB.bfo_reset_stats.restype = None
        _LIB.bfo_get_stats.restype = None
        _LIB.bfo_get_stats.argtypes = [_IP, _IP]
        _LIB.bfo_set_threads.restype = None
        _LIB.bfo_set_threads.argtypes = [ctypes.c_int]
        _LIB.bfo_set_threads(1)           # single-threaded unless a caller asks otherwise (set_threads)
    return _LIB


def set_threads(n: int) -> None:
    """OpenMP threads of the C oracle (IoU pairs, particles); 1 = faithful single-threaded restatement."""
    lib().bfo_set_threads(int(n))


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_FP)


def K16_from_K3(K3) -> np.ndarray:
    """4x4 flattened intrinsics as BoxFusion keeps them (box_fusion.py:37-40, 463-466)."""
    K = np.eye(4, dtype=np.float64)
    K[:3, :3] = np.asarray(K3, dtype=np.float64)
    return K.reshape(-1).astype(np.float32)


def evaluate(box6, t_c, pst, rot, poses, K16, search, img_h, img_w, pst_size=None) -> np.ndarray:
    """evaluate_iou (box_fusion.py:413-461): fitness[P] float32."""
    pst_a, pst_p = _f(pst)
    P = pst_a.shape[0]
    pst_size = P if pst_size is None else pst_size
    n_eval = min(32 * (pst_size // 32), P)
    t_a, t_p = _f(np.asarray(t_c).reshape(-1, 16))
    V = t_a.shape[0]
    b_a, b_p = _f(box6); r_a, r_p = _f(rot); po_a, po_p = _f(np.asarray(poses).reshape(V, 16))
    k_a, k_p = _f(K16); s_a, s_p = _f(search)
    out = np.zeros(P, dtype=np.float32)
    lib().bfo_evaluate(b_p, t_p, pst_p, P, n_eval, r_p, po_p, V, k_p, s_p, float(img_h), float(img_w),
                       out.ctypes.data_as(_FP))
    return out


def make_cfg_struct(cfg: dict, img_h: float, img_w: float, beta: float = 0.9, early_stop: bool = True) -> _Cfg:
    bf = cfg["box_fusion"]
    ro = bf["random_opt"]
    return _Cfg(int(bf["iters"]), int(bf["pst_size"]), float(ro["center_init_size"]), float(ro["shape_init_size"]),
                float(ro["center_scaling_coefficient"]), float(ro["shape_scaling_coefficient"]),
                float(beta), float(img_h), float(img_w), int(bool(early_stop)))


def refine_box(view_boxes, view_R, view_scores, t_c, poses, pst, K16, cstruct: _Cfg, want_trace=False):
    """Optimiser loop for one map box (box_fusion.py:651-721)."""
    vb_a, vb_p = _f(view_boxes); V = vb_a.shape[0]
    vr_a, vr_p = _f(np.asarray(view_R).reshape(V, 9)); vs_a, vs_p = _f(view_scores)
    t_a, t_p = _f(np.asarray(t_c).reshape(V, 16)); po_a, po_p = _f(np.asarray(poses).reshape(V, 16))
    pst_a, pst_p = _f(pst); k_a, k_p = _f(K16)
    out6 = np.zeros(6, dtype=np.float32)
    n_it = ctypes.c_int(0)
    trace = np.zeros((cstruct.iters, 8), dtype=np.float32) if want_trace else None
    upd = lib().bfo_refine_box(vb_p, vr_p, vs_p, t_p, po_p, V, pst_p, pst_a.shape[0], k_p, ctypes.byref(cstruct),
                               out6.ctypes.data_as(_FP), ctypes.byref(n_it),
                               trace.ctypes.data_as(_FP) if want_trace else None)
    return bool(upd), out6, n_it.value, trace


def boxfusion_port(map_tensor: np.ndarray, fusion_list: List[List[int]], already_fusion: List[List[int]],
                   fusion_flag: List[int], per_tensor, per_R, per_scores, per_proj, per_pose,
                   pst, K16, cstruct: _Cfg):
    """Outer loop of BoxFusion.boxfusion (box_fusion.py:631-724) on plain arrays/lists; edits in place.

    Returns (list of fused map indices, total evaluate_iou calls)."""
    fused, calls = [], 0
    for i in range(map_tensor.shape[0]):
        fl = fusion_list[i]
        if len(fl) < 3 or (fl in already_fusion):
            continue
        idx = np.asarray(fl, dtype=np.int64)
        upd, out6, n_it, _ = refine_box(per_tensor[idx], per_R[idx], per_scores[idx], per_proj[idx], per_pose[idx],
                                        pst, K16, cstruct)
        calls += n_it
        if upd:
            map_tensor[i] = out6
            fusion_flag[i] = 1
            already_fusion.append(list(fl))
            fused.append(i)
    return fused, calls
