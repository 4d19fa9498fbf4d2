"""TEST INFRASTRUCTURE ONLY - CPU restatement ("port") of the reference's fusion hot path behind the
reference's own class names, so the synthetic replay driver (boxfusion_b200/driver.py) can run it
exactly like the reference or like the CUDA product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path never does.

Restates (paths relative to /root/reference; every function cites its lines):
  boxfusion/boxes.py:656-943      GeneralInstance3DBoxes (corners, dims, transform2world, cat, getitem)
  boxfusion/instances.py:22-125   nms_3d, calculate_obb_iou
  boxfusion/instances.py:128-331  Instances3D container
  boxfusion/instances.py:333-717  project_3d_boxes, spatial_/correspondence_association, obb_iou,
                                  check_intersection, IoU_2D_box, project_3d_to_2d_box
  boxfusion/box_manager.py:9-245  BoxManager
  boxfusion/box_fusion.py:27-724  BoxFusion (kernel + optimiser in oracle/refine_oracle.c)

Pinning: no reference tests/golden vectors exist for this path; the port is pinned
by running the unmodified reference in the build container on the same synthetic sequences and
comparing every mutated field bit for bit (tests/golden/make_golden.py -> tests/golden/*.npz,
tests/test_oracle_golden.py).  Environment of record: numpy 2.3.5, scipy 1.18.1 (Qhull), torch 2.11
(SURVEY F8).

`IOU_BACKEND`:
  "scipy" - half-spaces from scipy.spatial.ConvexHull exactly as the reference (slow; the CPU baseline)
  "c"     - half-spaces recomputed by oracle/assoc_oracle.c (fast; identical counts, see its header)
  "c_batch" - as "c", one OpenMP-parallel C call per NMS head (bench.py's multi-core C baseline)
"""
from __future__ import annotations

import copy
import ctypes
import itertools
from typing import Any, Dict, List, Tuple

import numpy as np
import torch
from scipy.spatial import ConvexHull

from . import build as _build
from . import refine_oracle as _ro

IOU_BACKEND = "scipy"

_FP = ctypes.POINTER(ctypes.c_float)
_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int32)
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(_build.build_oracle())
        _LIB.bfo_hull_planes.argtypes = [_FP, _DP]
        _LIB.bfo_obb_counts.argtypes = [_FP, _FP, _IP]
        _LIB.bfo_obb_counts.restype = ctypes.c_int
        _LIB.bfo_obb_counts_pairs.argtypes = [_FP, _IP, _IP, ctypes.c_int, _IP, _IP]
        _LIB.bfo_corners.argtypes = [_FP, _FP, ctypes.c_int, _FP]
        _LIB.bfo_obb_iou_one_vs_many.argtypes = [_FP, _FP, ctypes.c_int, _DP]
    return _LIB


# --------------------------------------------------------------------------------------------
# sampled oriented-3D IoU
# --------------------------------------------------------------------------------------------

def hull_planes_c(corners: np.ndarray) -> np.ndarray:
    c = np.ascontiguousarray(corners, dtype=np.float32)
    out = np.zeros((12, 4), dtype=np.float64)
    _lib().bfo_hull_planes(c.ctypes.data_as(_FP), out.ctypes.data_as(_DP))
    return out


def obb_counts_c(c1: np.ndarray, c2: np.ndarray) -> Tuple[int, np.ndarray]:
    a = np.ascontiguousarray(c1, dtype=np.float32)
    b = np.ascontiguousarray(c2, dtype=np.float32)
    cnt = np.zeros(3, dtype=np.int32)
    g = _lib().bfo_obb_counts(a.ctypes.data_as(_FP), b.ctypes.data_as(_FP), cnt.ctypes.data_as(_IP))
    return int(g), cnt


def obb_counts_pairs_c(corners: np.ndarray, ia: np.ndarray, ib: np.ndarray):
    c = np.ascontiguousarray(corners, dtype=np.float32)
    ia = np.ascontiguousarray(ia, dtype=np.int32)
    ib = np.ascontiguousarray(ib, dtype=np.int32)
    cnt = np.zeros((len(ia), 3), dtype=np.int32)
    gate = np.zeros(len(ia), dtype=np.int32)
    _lib().bfo_obb_counts_pairs(c.ctypes.data_as(_FP), ia.ctypes.data_as(_IP), ib.ctypes.data_as(_IP), len(ia),
                                cnt.ctypes.data_as(_IP), gate.ctypes.data_as(_IP))
    return gate, cnt


def iou_from_counts(cnt) -> np.ndarray:
    cnt = np.asarray(cnt, dtype=np.int64)
    return cnt[..., 2] / (cnt[..., 0] + cnt[..., 1] - cnt[..., 2] + 1e-6)      # instances.py:608


_EDGES = [[0, 1], [0, 4], [1, 5], [4, 5], [2, 3], [2, 6], [6, 7], [3, 7], [0, 3], [4, 7], [1, 2], [5, 6]]


def _augment(corners):                                   # instances.py:493-512
    mids = [(corners[a] + corners[b]) / 2 for a, b in _EDGES]
    return np.vstack([corners, mids])


def _inside(points, eq):                                 # instances.py:561-571
    return np.all(np.dot(points, eq[:, :3].T) + eq[:, 3] <= 1e-6, axis=1)


def obb_counts_scipy(c1: np.ndarray, c2: np.ndarray):
    """instances.py:514-613 with Qhull half-spaces; returns (gate, counts[3])."""
    e1, e2 = ConvexHull(c1).equations, ConvexHull(c2).equations
    a1, a2 = _augment(c1), _augment(c2)
    if np.sum(_inside(a1, e2)) + np.sum(_inside(a2, e1)) <= 0:
        return 0, np.zeros(3, dtype=np.int64)
    allc = np.concatenate([c1, c2], axis=0)
    lo, hi = np.min(allc, axis=0), np.max(allc, axis=0)
    xs, ys, zs = (np.linspace(lo[k], hi[k], 25) for k in range(3))
    xx, yy, zz = np.meshgrid(xs, ys, zs, indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    m1 = _inside(pts, ConvexHull(c1).equations)        # the reference rebuilds both hulls here (:600-601)
    m2 = _inside(pts, ConvexHull(c2).equations)
    return 1, np.array([m1.sum(), m2.sum(), (m1 & m2).sum()], dtype=np.int64)


def obb_iou(c1, c2):
    gate, cnt = (obb_counts_scipy if IOU_BACKEND == "scipy" else obb_counts_c)(np.asarray(c1), np.asarray(c2))
    return float(iou_from_counts(cnt)) if gate else 0.0


def calculate_obb_iou(corners1, corners_others):         # instances.py:106-125
    if IOU_BACKEND == "c_batch":                          # same values as "c", one call (OpenMP inside) per head
        a = np.ascontiguousarray(corners1, dtype=np.float32)
        b = np.ascontiguousarray(corners_others, dtype=np.float32)
        out = np.zeros(b.shape[0], dtype=np.float64)
        if b.shape[0]:
            _lib().bfo_obb_iou_one_vs_many(a.ctypes.data_as(_FP), b.ctypes.data_as(_FP), b.shape[0], out.ctypes.data_as(_DP))
        return out
    return np.asarray([obb_iou(corners1, corners_others[i]) for i in range(corners_others.shape[0])])


# --------------------------------------------------------------------------------------------
# NMS
# --------------------------------------------------------------------------------------------

def nms_3d(instance_lists, box_manager, boxes, scores, init_id, cam_poses, box_size, iou_threshold=0.5,
           merge_upper=0.7, merge_lower=0.3):
    """instances.py:22-101."""
    centers = np.mean(boxes, axis=1)
    order = scores.argsort()[::-1]
    order_init_id = init_id.tolist()
    keep: List[int] = []
    success: List[int] = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        rest = order[1:]
        ious = calculate_obb_iou(boxes[i], boxes[rest])
        hit = np.where(ious > iou_threshold)[0]
        if hit.shape[0] >= 1:
            instance_lists.valid_num[i] += 1
            success.append(i)
            keep = box_manager.record(i, [j for j in rest[hit]], order_init_id, cam_poses, box_size, keep, centers)
        order = rest[np.where(ious <= iou_threshold)[0]]
        if order.size == 1:
            keep.append(order[0])
            break
    keep.sort()
    success.sort()
    return np.array(keep), np.array(success)


# --------------------------------------------------------------------------------------------
# containers
# --------------------------------------------------------------------------------------------

class GeneralInstance3DBoxes:
    """boxes.py:656-943 (the parts the hot path touches)."""

    def __init__(self, xyzlhw, R, **_):
        dev = xyzlhw.device if isinstance(xyzlhw, torch.Tensor) else torch.device("cpu")
        self.tensor = torch.as_tensor(xyzlhw, dtype=torch.float32, device=dev).clone()
        self.R = torch.as_tensor(R, dtype=torch.float32, device=dev).clone()

    @property
    def dims(self):
        return self.tensor[:, 3:6]

    @property
    def gravity_center(self):
        return self.tensor[:, :3]

    @property
    def device(self):
        return self.tensor.device

    @property
    def corners(self):                                   # boxes.py:725-778
        t = np.ascontiguousarray(self.tensor.detach().cpu().numpy(), dtype=np.float32)
        r = np.ascontiguousarray(self.R.detach().cpu().numpy().reshape(-1, 9), dtype=np.float32)
        out = np.zeros((t.shape[0], 8, 3), dtype=np.float32)
        if t.shape[0]:
            _lib().bfo_corners(t.ctypes.data_as(_FP), r.ctypes.data_as(_FP), t.shape[0], out.ctypes.data_as(_FP))
        return torch.from_numpy(out).to(self.tensor.device)

    def transform2world(self, cam_pose):                 # boxes.py:825-833
        if not isinstance(cam_pose, torch.Tensor):
            cam_pose = torch.from_numpy(cam_pose)
        cam_pose = cam_pose.to(self.tensor.device)
        c = (cam_pose[:, :3, :3] @ self.tensor[:, :3].unsqueeze(-1) + cam_pose[:, :3, 3:]).squeeze()
        self.tensor[:, :3] = c
        self.R = cam_pose[:, :3, :3] @ self.R

    def __getitem__(self, item):
        if isinstance(item, int):
            return GeneralInstance3DBoxes(self.tensor[item].view(1, -1), self.R[item].view(1, 3, 3))
        return GeneralInstance3DBoxes(self.tensor[item], self.R[item])

    def __len__(self):
        return self.tensor.shape[0]

    @classmethod
    def cat(cls, boxes_list):
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0), torch.cat([b.R for b in boxes_list], dim=0))

    def to(self, device):
        return GeneralInstance3DBoxes(self.tensor.to(device), self.R.to(device))

    def clone(self):
        return GeneralInstance3DBoxes(self.tensor.clone(), self.R.clone())


class Instances3D:
    """instances.py:128-331 container + :333-717 hot-path methods."""

    def __init__(self, image_size: Tuple[int, int] = (0, 0), **kwargs: Any):
        self._image_size = image_size
        self._fields: Dict[str, Any] = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances3D!".format(name))
        return self._fields[name]

    def set(self, name, value):
        if len(self._fields):
            assert len(self) == len(value), "field length mismatch"
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __len__(self):
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances3D does not support __len__!")

    def __getitem__(self, item):
        if type(item) == int:
            if item >= len(self) or item < -len(self):
                raise IndexError("Instances3D index out of range!")
            item = slice(item, None, len(self))
        ret = Instances3D(image_size=self.image_size)
        for k, v in self._fields.items():
            if isinstance(v, (torch.Tensor, np.ndarray)) or hasattr(v, "tensor"):
                if isinstance(v, np.ndarray) and isinstance(item, torch.Tensor):
                    ret.set(k, v[item.cpu().numpy()])
                else:
                    ret.set(k, v[item])
            elif hasattr(v, "__iter__"):
                if isinstance(item, np.ndarray) and item.dtype == np.bool_:
                    ret.set(k, [v_ for i_, v_ in enumerate(v) if item[i_]])
                elif isinstance(item, torch.Tensor) and item.dtype == torch.bool:
                    ret.set(k, [v_ for i_, v_ in enumerate(v) if item[i_].item()])
                elif isinstance(item, torch.Tensor) and item.dtype == torch.int64:
                    ret.set(k, [v[i_.item()] for i_ in item])
                elif isinstance(item, slice):
                    ret.set(k, v[item])
                else:
                    raise ValueError("Expected Bool or Long Tensor")
            else:
                raise ValueError("Not supported!")
        return ret

    @staticmethod
    def cat(instance_lists):
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        ret = Instances3D(image_size=instance_lists[0]._image_size)
        for k in instance_lists[0]._fields.keys():
            values = [i.get(k) for i in instance_lists]
            v0 = values[0]
            if isinstance(v0, torch.Tensor):
                values = torch.cat(values, dim=0)
            elif isinstance(v0, np.ndarray):
                values = np.concatenate(values, axis=0)
            elif isinstance(v0, list):
                values = list(itertools.chain(*values))
            elif hasattr(type(v0), "cat"):
                values = type(v0).cat(values)
            else:
                raise ValueError("Unsupported type {} for concatenation".format(type(v0)))
            ret.set(k, values)
        return ret

    # ---- hot path ---------------------------------------------------------------------------
    def project_3d_boxes(self, K, H=480, W=640):          # instances.py:333-369
        corners = self.get("pred_boxes_3d").corners
        cam_pose = self.cam_pose
        N = corners.shape[0]
        homo = torch.cat([corners, torch.ones((N, 8, 1), device=corners.device)], dim=2)
        pose_inv = torch.linalg.inv(cam_pose).to(corners.device)
        cam = torch.einsum("nij,nkj->nki", pose_inv, homo)
        X, Y, Z = cam[..., 0], cam[..., 1], cam[..., 2]
        u = (K[0, 0] * X / Z) + K[0, 2]
        v = (K[1, 1] * Y / Z) + K[1, 2]
        self.projected_boxes = torch.stack([torch.clamp(u, 0, W), torch.clamp(v, 0, H)], dim=-1)

    def spatial_association(instance_lists, threshold, box_manager, cam_poses):    # instances.py:372-397
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists
        b = instance_lists.get("pred_boxes_3d")
        keep, success = nms_3d(instance_lists, box_manager, b.corners.cpu().numpy(), instance_lists.scores.cpu().numpy(),
                               instance_lists.init_id.cpu().numpy(), cam_poses, b.dims.cpu().numpy(),
                               iou_threshold=threshold)
        return sorted(keep), sorted(success)

    def correspondence_association(cfg, box_manager, cur_keep_idx, cur_success_nms, pred_instances, global_pred_box,
                                   all_pred_box, all_poses, per_frame_ins_cam_pose, frame_id, mask, intrinsic,
                                   all_kf_pose, threshold=0.33, H=480, W=640):      # instances.py:411-490
        N_glo = len(global_pred_box)
        cur_2d = pred_instances.pred_boxes.cpu().numpy()
        cur_scores = pred_instances.scores.cpu().numpy()
        glo_scores = global_pred_box.scores.cpu().numpy()
        pred_size = pred_instances.get("pred_boxes_3d").dims.cpu().numpy()
        init_id = all_pred_box.init_id.cpu().numpy()
        keep_idx = copy.deepcopy(np.asarray(mask))
        glo_keep = keep_idx[keep_idx < N_glo]
        small = [idx for idx in cur_keep_idx
                 if not (np.max(pred_size[idx, :3]) > cfg["box_fusion"]["small_size"] or idx in cur_success_nms)]
        if len(small) > 0:
            cur_pose = all_kf_pose[frame_id]
            gb = global_pred_box.get("pred_boxes_3d")
            for idx in small:
                b3 = gb.corners.cpu().numpy()[glo_keep, ...]
                b2 = Instances3D.project_3d_to_2d_box(b3, intrinsic.cpu().numpy(), cur_pose, H, W, frame_id=frame_id)
                if len(b2) == 0:
                    continue
                iou = Instances3D.IoU_2D_box(cur_2d[idx], b2)
                gdims = gb.dims.cpu().numpy()[glo_keep, ...]
                iou = iou * (np.max(gdims, axis=1) < cfg["box_fusion"]["small_size"] + 0.1)
                j = np.argmax(iou)
                if iou[j] > threshold:
                    cidx = glo_keep[j]
                    if glo_scores[cidx] < cur_scores[idx]:
                        keep_idx = keep_idx[keep_idx != cidx]
                        all_pred_box.valid_num[idx + N_glo] += 1
                        keep_idx = box_manager.record_corr(idx + N_glo, [cidx], init_id, per_frame_ins_cam_pose, keep_idx)
                    else:
                        keep_idx = keep_idx[keep_idx != (idx + N_glo)]
                        all_pred_box.valid_num[cidx] += 1
                        keep_idx = box_manager.record_corr(cidx, [idx + N_glo], init_id, per_frame_ins_cam_pose, keep_idx)
        keep_idx = np.sort(keep_idx)
        return all_pred_box[keep_idx], all_poses[keep_idx], keep_idx

    obb_iou = staticmethod(obb_iou)

    def IoU_2D_box(A, B):                                  # instances.py:643-668
        A = A.astype(np.float64)
        area_A = (A[2] - A[0]) * (A[3] - A[1])
        area_B = (B[:, 2] - B[:, 0]) * (B[:, 3] - B[:, 1])
        iw = np.maximum(0, np.minimum(A[2], B[:, 2]) - np.maximum(A[0], B[:, 0]))
        ih = np.maximum(0, np.minimum(A[3], B[:, 3]) - np.maximum(A[1], B[:, 1]))
        inter = iw * ih
        return inter / (area_A + area_B - inter + 1e-6)

    def project_3d_to_2d_box(boxes_3d, K, pose, H, W, frame_id=None):    # instances.py:670-717
        N = boxes_3d.shape[0]
        out = np.zeros((N, 4))
        homo = np.concatenate([boxes_3d, np.ones((N, 8, 1))], axis=2)
        cam = np.dot(homo, np.linalg.inv(pose).T)
        X, Y, Z = cam[..., 0], cam[..., 1], cam[..., 2]
        u = (K[0, 0] * X / Z) + K[0, 2]
        v = (K[1, 1] * Y / Z) + K[1, 2]
        valid = (Z > 0) * (u > 0) * (u < W) * (v > 0) * (v < H)
        zwin = (Z > 0) * (Z < 8)
        for i in range(N):
            if valid[i].sum() == 0:
                continue
            uu, vv = u[i][zwin[i]], v[i][zwin[i]]
            if len(uu) == 0:
                continue
            uu, vv = np.clip(uu, 0, W), np.clip(vv, 0, H)
            out[i] = [np.min(uu), np.min(vv), np.max(uu), np.max(vv)]
        return out


# --------------------------------------------------------------------------------------------
# BoxManager
# --------------------------------------------------------------------------------------------

class BoxManager:
    """box_manager.py:9-245."""

    def __init__(self, cfg):
        self.fusion_list: List[List[int]] = []
        self.last_fusion_frame: List[List[int]] = []
        self.fusion_flag: List[int] = []
        self.already_fusion: List[List[int]] = []
        self.num_record: Dict[int, int] = {}
        self.cfg = cfg
        self.rotation_gap = cfg["association"]["rotation_gap"]
        self.translation_gap = cfg["association"]["translation_gap"]
        self.small_size = cfg["box_fusion"]["small_size"]
        self.merge_log: List[Dict] = []

    def init_new_predictions(self, box_num, all_num):     # :24-28
        for i in range(box_num):
            self.fusion_list.append([i + all_num])
            self.last_fusion_frame.append([0])
            self.fusion_flag.append(0)

    def add_fusion_ind(self, idx_list):                   # :31-32
        self.already_fusion.append(copy.deepcopy(idx_list))

    def check_if_fusion(self, idx_list):                  # :34-38
        return idx_list in self.already_fusion

    def _differs(self, p1, p2, with_center, c1=None, c2=None):
        if with_center:
            b, r, _, cd = self.compute_pose_center_disparity(p1, p2, c1, c2)
            return bool((b > self.translation_gap or r > self.rotation_gap) or cd > 0.5)
        b, r, _ = self.compute_pose_disparity(p1, p2)
        return bool(r > self.rotation_gap or b > self.translation_gap)

    def record(self, cur_id, fusion_inds, init_id, cam_poses, box_size, keep, box_centers):    # :40-88
        fl = self.fusion_list
        for idx in fusion_inds:
            if len(fl[idx]) == 1:
                cnt = sum(self._differs(cam_poses[i], cam_poses[init_id[idx]], True, box_centers[cur_id], box_centers[idx])
                          for i in fl[cur_id])
                if cnt == len(fl[cur_id]) and len(fl[cur_id]) < 5:
                    fl[cur_id] += [init_id[idx]]
                    fl[cur_id].sort()
            else:
                cnt = sum(self._differs(cam_poses[i], cam_poses[init_id[cur_id]], True, box_centers[cur_id], box_centers[idx])
                          for i in fl[idx])
                if cnt == len(fl[idx]) and len(fl[idx]) < 5:
                    fl[cur_id] += fl[idx]
                    fl[cur_id].sort()
                elif cur_id in keep:
                    keep.remove(cur_id)
                    keep.append(idx)
                if self.fusion_flag[idx] == 1:
                    self.fusion_flag[cur_id] = 1
        return keep

    def record_corr(self, cur_id, fusion_inds, init_id, cam_poses, keep):    # :90-129
        fl = self.fusion_list
        for idx in fusion_inds:
            if len(fl[idx]) == 1:
                cnt = sum(self._differs(cam_poses[i], cam_poses[init_id[idx]], False) for i in fl[cur_id])
                if cnt == len(fl[cur_id]) and len(fl[cur_id]) < 5:
                    fl[cur_id] += [init_id[idx]]
                    fl[cur_id].sort()
            else:
                cnt = sum(self._differs(cam_poses[i], cam_poses[init_id[cur_id]], False) for i in fl[idx])
                if cnt == len(fl[idx]) and len(fl[idx]) < 5:
                    fl[cur_id] += fl[idx]
                    fl[cur_id].sort()
                elif cur_id in keep:
                    keep[keep == cur_id] = idx
                if self.fusion_flag[idx] == 1:
                    self.fusion_flag[cur_id] = 1
        return keep

    def update(self, keep_idx):                           # :131-133
        self.fusion_list = [self.fusion_list[i] for i in keep_idx]

    def update_fusion_flag(self, idx):
        self.fusion_flag[idx] = 1

    def get_fusion_idx(self):
        return [i for i in range(len(self.fusion_flag)) if self.fusion_flag[i] == 1]

    def get_nofusion_idx(self):
        return [i for i in range(len(self.fusion_flag)) if self.fusion_flag[i] == 0]

    def check_valid_num(self, all_pred_box, count, gap):  # :151-166
        zero = torch.where((all_pred_box.valid_num == 0) & (all_pred_box.frame_id < (count - gap)))[0]
        valid = torch.arange(len(all_pred_box))
        for idx in zero:
            valid = valid[valid != idx]
        self.fusion_list = [self.fusion_list[int(i)] for i in valid]
        return all_pred_box[valid]

    def compute_pose_disparity(self, pose1, pose2):       # :168-186
        R1, t1, R2, t2 = pose1[:3, :3], pose1[:3, 3], pose2[:3, :3], pose2[:3, 3]
        baseline = torch.norm(t2 - t1, p=2)
        trace = torch.clamp((torch.trace(R2 @ R1.T) - 1) / 2, min=-1.0, max=1.0)
        angle = torch.arccos(trace) * 180 / torch.pi
        return baseline, angle, 0.6 * baseline + 0.4 * angle

    def compute_pose_center_disparity(self, pose1, pose2, center1, center2):    # :188-215
        b, a, s = self.compute_pose_disparity(pose1, pose2)
        return b, a, s, np.sqrt(np.sum((center1 - center2) ** 2))

    def check_uv_bounds(self, uv, W, H, ratio=1.0):       # :217-225
        gw, gh = int((1 - ratio) * W), int((1 - ratio) * H)
        u, v = uv[:, 0], uv[:, 1]
        return (u > gw) & (u < (W - gw)) & (v > gh) & (v < (H - gh))

    def check_floor_mask(self, box_3d, ratio=20):         # :227-237
        s = box_3d[:, 3:]
        mx, mn = torch.amax(s, dim=1), torch.amin(s, dim=1)
        sec = torch.sort(s, dim=1, descending=True)[0][:, 1]
        second = (mx / mn > ratio / 2) & (mx / sec > ratio / 2) & (sec / mn < 2.0) & (sec < 0.15) & (mn < 0.15)
        return (mx / mn > ratio) | second

    def check_large_mask(self, box_3d, thres=0.5):        # :239-245
        return torch.amax(box_3d[:, 3:], dim=1) > thres


# --------------------------------------------------------------------------------------------
# BoxFusion
# --------------------------------------------------------------------------------------------

class BoxFusion:
    """box_fusion.py:27-724; the kernel and the per-box optimiser live in oracle/refine_oracle.c."""

    def __init__(self, cfg):
        self.cfg = cfg
        path = cfg["box_fusion"]["pst_path"]
        if isinstance(path, np.ndarray):
            self.PST = np.ascontiguousarray(path, dtype=np.float32)
        elif str(path).endswith(".npy"):
            self.PST = np.ascontiguousarray(np.load(path), dtype=np.float32)
        else:
            import cv2
            self.PST = np.ascontiguousarray(cv2.imread(path, -1))
        cam = cfg["cam"]
        self.K = np.array([[cam["fx"], 0.0, cam["cx"], 0.0], [0.0, cam["fy"], cam["cy"], 0.0],
                           [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
        self.H, self.W = cam["H"], cam["W"]
        self.update_K_flag = False
        self.fusion_iters = cfg["box_fusion"]["iters"]
        self.pst_size = cfg["box_fusion"]["pst_size"]

    def update_intrinsics(self, size, K):                  # :463-466
        self.H, self.W = size[1], size[0]
        self.K[:3, :3] = K

    def evaluate_iou(self, box_3d, corners_2d, box_rot, scores_box, camera_poses, search_size, num_of_boxes, verbose=False):
        return _ro.evaluate(box_3d, corners_2d, self.PST, box_rot, camera_poses[:num_of_boxes],
                            self.K.reshape(-1).astype(np.float32), search_size, self.H, self.W, self.pst_size)

    def boxfusion(self, all_pred_box, per_frame_box, box_manager, beta=0.9, verbose=False):    # :622-724
        per_pose = per_frame_box.cam_pose.cpu().numpy()
        per_t = per_frame_box.pred_boxes_3d.tensor.cpu().numpy()
        per_R = per_frame_box.get("pred_boxes_3d").R.cpu().numpy()
        per_s = per_frame_box.scores.cpu().numpy()
        per_c = per_frame_box.projected_boxes.cpu().numpy()
        cs = _ro.make_cfg_struct(self.cfg, self.H, self.W, beta=beta)
        K16 = self.K.reshape(-1).astype(np.float32)
        self.n_eval_calls = 0
        for i in range(len(all_pred_box)):
            fl = box_manager.fusion_list[i]
            if len(fl) < 3 or box_manager.check_if_fusion(fl):
                continue
            upd, out6, n_it, _ = _ro.refine_box(per_t[fl], per_R[fl], per_s[fl], per_c[fl], per_pose[fl], self.PST, K16, cs)
            self.n_eval_calls += n_it
            if upd:
                all_pred_box.pred_boxes_3d.tensor[i] = torch.from_numpy(out6).to(all_pred_box.pred_boxes_3d.tensor.device)
                box_manager.update_fusion_flag(i)
                box_manager.add_fusion_ind(fl)
