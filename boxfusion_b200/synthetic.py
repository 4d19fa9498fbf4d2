"""Deterministic synthetic scenes shaped like the reference's CA-1M / ScanNet inputs.

The reference ships no data, tests or fixtures, so every parity
test, golden vector and bench line in this repo is driven by this generator.
It produces, per keyframe, exactly the tensors the reference's `demo.py` hands
to the fusion hot path (demo.py:216-221): camera-frame boxes `(x,y,z,l,h,w)` +
`R`, scores, 2-D boxes and the camera->world pose.  Shapes follow SURVEY.md
section 8(d); camera / box conventions follow boxes.py:725-778 (box frame:
l<->X, h<->Y, w<->Z) and box_fusion.py:348-357 (pose = camera->world, z-forward,
y-down pinhole).

Only numpy `RandomState` is used so that a seed reproduces the same scene on
every machine.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# ----------------------------------------------------------------------------
# configs (same nested-dict schema the reference reads from config/*.yaml)
# ----------------------------------------------------------------------------

_CA1M = {
    "dataset": "online",  # avoids the K_depth.txt file read (box_fusion.py:36,45)
    "data": {"datadir": "synthetic/ca1m", "start": 0, "output_dir": None, "gap": 1},
    "cam": {"H": 512, "W": 384, "fx": 425.0, "fy": 425.0, "cx": 192.0, "cy": 256.0,
            "png_depth_scale": 1000.0},
    "detection": {"score_thresh": 0.4, "uv_bound": True, "uv_bound_value": 0.9,
                  "floor_mask": True, "floor_ratio": 15, "scale_box": 1.5,
                  "size_max_thres": 0, "class_sim_thres": 0.0},
    "association": {"small_threshold": 0.2, "rotation_gap": 30, "translation_gap": 0.8},
    "box_fusion": {"use": True, "iters": 20, "pst_path": None, "pst_size": 1024,
                   "random_opt": {"center_init_size": 0.1, "center_scaling_coefficient": 0.1,
                                  "shape_init_size": 0.5, "shape_scaling_coefficient": 0.5},
                   "check_valid": False, "nms_threshold": 0.1, "small_size": 0.5,
                   "clip_sim_coeff": 0.0},
    "vis": {"rerun": False, "show_class": False, "show_label": False, "trajectory": False},
    "eval": False,
}

_SCANNET = copy.deepcopy(_CA1M)
_SCANNET["data"]["datadir"] = "synthetic/scannet"
_SCANNET["cam"] = {"H": 480, "W": 640, "fx": 574.540771, "fy": 577.583740,
                   "cx": 322.522827, "cy": 238.558853, "png_depth_scale": 1000.0}
_SCANNET["detection"]["score_thresh"] = 0.5
_SCANNET["association"]["small_threshold"] = 0.1
_SCANNET["box_fusion"]["small_size"] = 0.35


def make_cfg(shape: str = "ca1m", pst_path: Optional[str] = None, pst_size: int = 1024) -> dict:
    """cfg dict with the keys the hot path reads."""
    cfg = copy.deepcopy(_CA1M if shape == "ca1m" else _SCANNET)
    cfg["box_fusion"]["pst_path"] = pst_path
    cfg["box_fusion"]["pst_size"] = pst_size
    return cfg


def make_pst(P: int, seed: int = 0) -> np.ndarray:
    """Particle-swarm template [P,6] f32: row 0 = 0, rows 1.. uniform in the unit 6-ball.

    The reference ships a fixed 1024-row template (box_fusion.py:31-32);
    for other sizes SURVEY section 8(d) prescribes g/|g| * u^(1/6).
    """
    rs = np.random.RandomState(seed)
    g = rs.standard_normal((P, 6))
    u = rs.uniform(0.0, 1.0, (P, 1))
    pst = g / np.linalg.norm(g, axis=1, keepdims=True) * u ** (1.0 / 6.0)
    pst[0] = 0.0
    return np.ascontiguousarray(pst.astype(np.float32))


# ----------------------------------------------------------------------------
# geometry helpers
# ----------------------------------------------------------------------------

def box_rotation(yaw: np.ndarray, roll: Optional[np.ndarray] = None,
                 pitch: Optional[np.ndarray] = None) -> np.ndarray:
    """World rotation of a gravity-aligned box: local X,Z horizontal, local Y = world -Z (down).

    Mirrors the reference's box frame where `h` (local Y) is the gravity axis
    (cubify_transformer.py:597-600 builds R = R_c2w * T_gravity * R_y(theta)).
    """
    yaw = np.asarray(yaw, dtype=np.float64)
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.zeros(yaw.shape + (3, 3))
    R[..., 0, 0], R[..., 1, 0], R[..., 2, 0] = c, s, 0.0          # local X
    R[..., 0, 1], R[..., 1, 1], R[..., 2, 1] = 0.0, 0.0, -1.0     # local Y (down)
    R[..., 0, 2], R[..., 1, 2], R[..., 2, 2] = -s, c, 0.0         # local Z = X x Y
    if roll is not None:
        cr, sr = np.cos(roll), np.sin(roll)
        Rx = np.zeros_like(R)
        Rx[..., 0, 0] = 1.0
        Rx[..., 1, 1], Rx[..., 1, 2], Rx[..., 2, 1], Rx[..., 2, 2] = cr, -sr, sr, cr
        R = Rx @ R
    if pitch is not None:
        cp, sp = np.cos(pitch), np.sin(pitch)
        Ry = np.zeros_like(R)
        Ry[..., 1, 1] = 1.0
        Ry[..., 0, 0], Ry[..., 0, 2], Ry[..., 2, 0], Ry[..., 2, 2] = cp, sp, -sp, cp
        R = Ry @ R
    return R


def look_at_pose(eye: np.ndarray, target: np.ndarray) -> np.ndarray:
    """camera->world 4x4 (z forward, y down, x right), world Z up."""
    fwd = target - eye
    fwd = fwd / np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    n = np.linalg.norm(right)
    right = np.array([1.0, 0.0, 0.0]) if n < 1e-9 else right / n
    down = np.cross(fwd, right)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, down, fwd, eye
    return T


def corners_from_boxes(tensor: np.ndarray, R: np.ndarray) -> np.ndarray:
    """float64 helper (vertex order of boxes.py:756-766); not the parity path."""
    l, h, w = tensor[:, 3], tensor[:, 4], tensor[:, 5]
    sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1]) * 0.5
    sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1]) * 0.5
    sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1]) * 0.5
    v = np.stack([l[:, None] * sx, h[:, None] * sy, w[:, None] * sz], axis=-1)  # [N,8,3]
    return np.einsum("nij,nkj->nki", R, v) + tensor[:, None, :3]


# ----------------------------------------------------------------------------
# scenes
# ----------------------------------------------------------------------------

@dataclass
class Keyframe:
    """What demo.py:138-148 leaves in `pred_instances` for one keyframe (camera frame)."""
    frame_id: int
    pose: np.ndarray            # [4,4] f32 camera->world
    K: np.ndarray               # [3,3] f32
    image_size: tuple           # (W, H)
    tensor_cam: np.ndarray      # [n,6] f32 (x,y,z,l,h,w) in the camera frame
    R_cam: np.ndarray           # [n,3,3] f32
    scores: np.ndarray          # [n] f32
    pred_boxes: np.ndarray      # [n,4] f32 xyxy
    pred_proj_xy: np.ndarray    # [n,2] f32
    gt_index: np.ndarray        # [n] int64, -1 = spurious detection


@dataclass
class SyntheticScene:
    """A room of `n_objects` boxes observed by a moving camera.

    shape        "ca1m" (384x512 portrait) or "scannet" (640x480), see make_cfg.
    max_det      upper bound on detections per keyframe (reference: topk 100, typically <= 50).
    tilt_noise   sigma (rad) of per-detection roll/pitch noise; 0 keeps every box
                 exactly gravity-aligned (ANALYTIC IoU mode requires a shared axis).
    """
    n_objects: int = 200
    seed: int = 0
    shape: str = "ca1m"
    max_det: int = 50
    new_frac: float = 0.1
    tilt_noise: float = 0.0
    room_scale: float = 1.2
    cfg: dict = field(init=False)

    def __post_init__(self):
        self.cfg = make_cfg(self.shape)
        rs = np.random.RandomState(self.seed)
        n = self.n_objects
        s = self.room_scale * np.sqrt(max(n, 1))
        self.side = s
        self.centers = np.stack([rs.uniform(-s / 2, s / 2, n), rs.uniform(-s / 2, s / 2, n),
                                 rs.uniform(0.2, 2.0, n)], axis=1)
        self.dims = np.clip(np.exp(rs.normal(np.log(0.6), 0.5, (n, 3))), 0.08, 2.5)
        self.yaw = rs.uniform(-np.pi, np.pi, n)
        cam = self.cfg["cam"]
        self.K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]],
                          dtype=np.float32)
        self.W, self.H = int(cam["W"]), int(cam["H"])

    # camera path: orbit (radius 2.5 m) around a target that sweeps the room
    def _pose(self, k: int, rs: np.random.RandomState) -> np.ndarray:
        s = self.side
        t = k * 0.37
        target = np.array([0.35 * s * np.sin(0.23 * t), 0.35 * s * np.cos(0.31 * t + 0.5), 0.9])
        ang = 0.9 * k + rs.uniform(-0.2, 0.2)
        eye = target + np.array([2.5 * np.cos(ang), 2.5 * np.sin(ang), 0.0])
        eye[2] = rs.uniform(0.3, 1.2) + 0.4
        return look_at_pose(eye, target)

    def keyframe(self, k: int) -> Keyframe:
        rs = np.random.RandomState((self.seed * 1000003 + 7919 * (k + 1)) % (2 ** 31 - 1))
        T = self._pose(k, rs)
        Rcw, tcw = T[:3, :3], T[:3, 3]
        fx, fy, cx, cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        pc = (self.centers - tcw) @ Rcw              # camera-frame centres
        z = pc[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = fx * pc[:, 0] / z + cx
            v = fy * pc[:, 1] / z + cy
        vis = (z > 0.4) & (z < 6.0) & (u > 0) & (u < self.W) & (v > 0) & (v < self.H)
        idx = np.nonzero(vis)[0]
        n_new = int(round(self.new_frac * min(len(idx), self.max_det)))
        idx = idx[np.argsort(z[idx])][: self.max_det - n_new]
        n_obs = len(idx)
        # noisy re-observations
        c = self.centers[idx] + rs.normal(0, 0.05, (n_obs, 3))
        d = self.dims[idx] * np.exp(rs.normal(0, 0.1, (n_obs, 3)))
        yaw = self.yaw[idx] + rs.normal(0, 0.05, n_obs)
        # spurious / brand-new boxes somewhere in the frustum
        zc = rs.uniform(0.8, 4.5, n_new)
        uc = rs.uniform(0.1 * self.W, 0.9 * self.W, n_new)
        vc = rs.uniform(0.1 * self.H, 0.9 * self.H, n_new)
        pn = np.stack([(uc - cx) / fx * zc, (vc - cy) / fy * zc, zc], axis=1) @ Rcw.T + tcw
        c = np.concatenate([c, pn], axis=0)
        d = np.concatenate([d, np.clip(np.exp(rs.normal(np.log(0.5), 0.5, (n_new, 3))), 0.08, 2.5)], axis=0)
        yaw = np.concatenate([yaw, rs.uniform(-np.pi, np.pi, n_new)])
        n = n_obs + n_new
        roll = pitch = None
        if self.tilt_noise > 0:
            roll = rs.normal(0, self.tilt_noise, n)
            pitch = rs.normal(0, self.tilt_noise, n)
        Rw = box_rotation(yaw, roll, pitch)
        scores = rs.uniform(0.4, 1.0, n) + np.arange(n) * 1e-7   # no exact ties
        # camera frame (what the detector emits; demo.py:220 lifts it back to world)
        R_cam = np.einsum("ji,njk->nik", Rcw, Rw)
        c_cam = (c - tcw) @ Rcw
        tensor_cam = np.concatenate([c_cam, d], axis=1)
        corners_cam = corners_from_boxes(tensor_cam, R_cam)
        zz = np.maximum(corners_cam[..., 2], 1e-3)
        uu = np.clip(fx * corners_cam[..., 0] / zz + cx, 0, self.W)
        vv = np.clip(fy * corners_cam[..., 1] / zz + cy, 0, self.H)
        pred_boxes = np.stack([uu.min(1), vv.min(1), uu.max(1), vv.max(1)], axis=1)
        proj_xy = np.stack([fx * c_cam[:, 0] / np.maximum(c_cam[:, 2], 1e-3) + cx,
                            fy * c_cam[:, 1] / np.maximum(c_cam[:, 2], 1e-3) + cy], axis=1)
        return Keyframe(
            frame_id=k, pose=T.astype(np.float32), K=self.K.copy(), image_size=(self.W, self.H),
            tensor_cam=tensor_cam.astype(np.float32), R_cam=R_cam.astype(np.float32),
            scores=scores.astype(np.float32), pred_boxes=pred_boxes.astype(np.float32),
            pred_proj_xy=proj_xy.astype(np.float32),
            gt_index=np.concatenate([idx, -np.ones(n_new, dtype=np.int64)]))


def random_boxes(n: int, seed: int, side: Optional[float] = None, tilt_noise: float = 0.0):
    """n world-frame boxes (tensor[n,6] f32, R[n,3,3] f32) at the section-8(d) density."""
    rs = np.random.RandomState(seed)
    s = side if side is not None else 1.2 * np.sqrt(max(n, 1))
    c = np.stack([rs.uniform(-s / 2, s / 2, n), rs.uniform(-s / 2, s / 2, n), rs.uniform(0.2, 2.0, n)], 1)
    d = np.clip(np.exp(rs.normal(np.log(0.6), 0.5, (n, 3))), 0.08, 2.5)
    yaw = rs.uniform(-np.pi, np.pi, n)
    roll = pitch = None
    if tilt_noise > 0:
        roll, pitch = rs.normal(0, tilt_noise, n), rs.normal(0, tilt_noise, n)
    R = box_rotation(yaw, roll, pitch)
    return (np.concatenate([c, d], 1).astype(np.float32), R.astype(np.float32))


def map_and_detections(n_map: int, n_det: int, seed: int, tilt_noise: float = 0.0):
    """C1/C3-style stress input: a map and detections that re-observe 70 % of random map boxes.

    Returns world-frame (tensor, R, scores) for map and detections.
    """
    rs = np.random.RandomState(seed + 17)
    mt, mR = random_boxes(n_map, seed, tilt_noise=tilt_noise)
    side = 1.2 * np.sqrt(max(n_map, 1))
    n_re = int(round(0.7 * n_det))
    pick = rs.randint(0, n_map, n_re)
    # recover yaw of picked map boxes from R (local X axis)
    yaw = np.arctan2(mR[pick, 1, 0], mR[pick, 0, 0]) + rs.normal(0, 0.05, n_re)
    c = mt[pick, :3] + rs.normal(0, 0.05, (n_re, 3))
    d = mt[pick, 3:] * np.exp(rs.normal(0, 0.1, (n_re, 3)))
    nt, nR = random_boxes(n_det - n_re, seed + 101, side=side, tilt_noise=tilt_noise)
    roll = pitch = None
    if tilt_noise > 0:
        roll, pitch = rs.normal(0, tilt_noise, n_re), rs.normal(0, tilt_noise, n_re)
    dt = np.concatenate([np.concatenate([c, d], 1).astype(np.float32), nt], 0)
    dR = np.concatenate([box_rotation(yaw, roll, pitch).astype(np.float32), nR], 0)
    ms = (rs.uniform(0.4, 1.0, n_map) + np.arange(n_map) * 1e-7).astype(np.float32)
    ds = (rs.uniform(0.4, 1.0, n_det) + np.arange(n_det) * 1e-7).astype(np.float32)
    return (mt, mR, ms), (dt, dR, ds)


def refine_problem(n_boxes: int, n_views: int, seed: int, shape: str = "ca1m"):
    """C4-style input for the particle refinement: per box, `n_views` noisy observations.

    Returns dict with per-view world boxes [B,V,6], R [B,V,3,3], scores [B,V],
    poses [B,V,4,4] (camera->world) and K, (W,H).  Observation corners are NOT
    included: callers project with the implementation under test (instances.py:333-369).
    """
    rs = np.random.RandomState(seed)
    cfg = make_cfg(shape)
    cam = cfg["cam"]
    K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], dtype=np.float32)
    B, V = n_boxes, n_views
    c = np.stack([rs.uniform(-3, 3, B), rs.uniform(-3, 3, B), rs.uniform(0.3, 1.5, B)], 1)
    d = np.clip(np.exp(rs.normal(np.log(0.6), 0.4, (B, 3))), 0.1, 2.0)
    yaw = rs.uniform(-np.pi, np.pi, B)
    poses = np.zeros((B, V, 4, 4))
    for b in range(B):
        a0 = rs.uniform(0, 2 * np.pi)
        for v in range(V):
            ang = a0 + v * (2 * np.pi / max(V, 3)) * rs.uniform(0.6, 1.0)
            eye = c[b] + np.array([2.5 * np.cos(ang), 2.5 * np.sin(ang), 0.0])
            eye[2] = rs.uniform(0.3, 1.2) + 0.4
            poses[b, v] = look_at_pose(eye, c[b] + rs.normal(0, 0.15, 3))
    oc = c[:, None, :] + rs.normal(0, 0.05, (B, V, 3))
    od = d[:, None, :] * np.exp(rs.normal(0, 0.1, (B, V, 3)))
    oy = yaw[:, None] + rs.normal(0, 0.05, (B, V))
    R = box_rotation(oy)
    scores = rs.uniform(0.4, 1.0, (B, V)) + np.arange(V)[None, :] * 1e-6
    return {"tensor": np.concatenate([oc, od], -1).astype(np.float32), "R": R.astype(np.float32),
            "scores": scores.astype(np.float32), "poses": poses.astype(np.float32), "K": K,
            "size": (int(cam["W"]), int(cam["H"]))}
