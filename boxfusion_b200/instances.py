"""`boxfusion.instances` for the B200 path (reference: boxfusion/instances.py).

`Instances3D` keeps the reference's Detectron2-style field container semantics (set/get, indexing
by int / slice / bool / long / ndarray, cat) because `demo.py` and the hot path mutate it in place;
the association entry points run on the GPU through libboxfusion_sm100.so:

  spatial_association         -> bf_box_corners + bf_nms3d   (instances.py:22-101, 372-397; box_manager.py:40-88)
  correspondence_association  -> bf_box_corners + bf_corr2d  (instances.py:411-490, 643-717)
  project_3d_boxes            -> bf_box_corners + bf_project_boxes (instances.py:333-369)
  obb_iou / calculate_obb_iou -> bf_iou3d_matrix             (instances.py:106-125, 573-613)

`IOU_MODE` selects the oriented-3D IoU estimator for association:
`ops.IOU_SAMPLED_REF` (default; reference-exact sampled estimator) or `ops.IOU_ANALYTIC`.
"""
from __future__ import annotations

import copy
import itertools
from typing import Any, Dict, List, Tuple, Union

import numpy as np
import torch

from . import fastpath, ops

IOU_MODE = ops.IOU_SAMPLED_REF


def _as_numpy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def calculate_obb_iou(corners1, corners_others):
    """One-vs-many oriented-3D IoU (instances.py:106-125): corners [8,3] vs [k,8,3] -> float64 [k] (numpy)."""
    c1 = np.asarray(corners1, dtype=np.float32).reshape(1, 8, 3)
    co = np.asarray(corners_others, dtype=np.float32).reshape(-1, 8, 3)
    if co.shape[0] == 0:
        return np.zeros(0, dtype=np.float64)
    return ops.iou3d_matrix(c1, co, IOU_MODE)[0].cpu().numpy()


def nms_3d(instance_lists, box_manager, boxes, scores, init_id, cam_poses, box_size, iou_threshold=0.5,
           merge_upper=0.7, merge_lower=0.3):
    """Greedy score-ordered 3-D NMS with fusion-list bookkeeping (instances.py:22-101) on the GPU.

    Same arguments as the reference: `boxes` are corners [N,8,3]; `box_manager.fusion_list/_flag` and
    `instance_lists.valid_num` are updated in place.  Returns (keep, success_nms) sorted int arrays."""
    dev = ops._pick_device(boxes, scores)
    corners = ops.dev_tensor(boxes, torch.float32, dev).reshape(-1, 8, 3)
    n = corners.shape[0]
    # nms_3d's centres are the float32 mean of the corners (instances.py:49): 8 sequential adds, /8
    c = corners[:, 0]
    for v in range(1, 8):
        c = c + corners[:, v]
    centers = (c * 0.125).contiguous()
    return _nms_device(instance_lists, box_manager, corners, centers, scores, init_id, cam_poses, iou_threshold, dev)


def _nms_device(instance_lists, box_manager, corners, centers, scores, init_id, cam_poses, iou_threshold, dev):
    n = corners.shape[0]
    s = ops.dev_tensor(scores, torch.float32, dev).reshape(-1)
    order = ops.score_order(s)                                                     # scores.argsort()[::-1], stable (bf_score_order)
    iid = ops.dev_tensor(init_id, torch.int64, dev).to(torch.int32).reshape(-1)
    poses = ops.dev_tensor(cam_poses, torch.float32, dev).reshape(-1, 16)
    fl_h, ln_h, flag_h = box_manager.pack_lists(n)
    packed = ops.dev_tensor(np.concatenate([fl_h.reshape(-1), ln_h, flag_h]), torch.int32, dev)
    cap = ops.FUSION_CAP
    fl, ln, flag = packed[: n * cap].view(n, cap), packed[n * cap: n * cap + n], packed[n * cap + n:]
    keep, success, status = ops.nms3d(corners, centers, order, iid, poses, fl, ln, flag, float(iou_threshold),
                                      float(box_manager.translation_gap), float(box_manager.rotation_gap), 0.5, IOU_MODE)
    # the call's single D2H also carries what correspondence_association needs on the host right afterwards
    # (largest box dimension, scores, init ids) so that it does not have to synchronise again
    dims = instance_lists.get("pred_boxes_3d").tensor[:, 3:6] if instance_lists.has("pred_boxes_3d") else None
    extra = None
    if dims is not None and dims.is_cuda:
        extra = torch.cat([torch.amax(dims, dim=1), s]).view(torch.int32)
    out = ops.to_host(torch.cat([packed, keep, success, status] + ([extra] if extra is not None else [])))
    if extra is not None:
        tail = out[n * cap + 4 * n + 1:].view(np.float32)
        instance_lists._bf_host = {"max_dim": tail[:n].copy(), "scores": tail[n:2 * n].copy(), "n": n}
    fl_o = out[: n * cap].reshape(n, cap)
    ln_o, flag_o = out[n * cap: n * cap + n], out[n * cap + n: n * cap + 2 * n]
    keep_o, succ_o = out[n * cap + 2 * n: n * cap + 3 * n], out[n * cap + 3 * n: n * cap + 4 * n]
    if out[n * cap + 4 * n] != 0:
        raise RuntimeError(f"bf_nms3d: a fusion list exceeded the device capacity ({cap})")
    box_manager.apply_lists(fl_o, ln_o, flag_o, ln_h)
    keep_idx, succ_idx = np.nonzero(keep_o)[0], np.nonzero(succ_o)[0]
    if len(succ_idx):
        instance_lists.valid_num[torch.as_tensor(succ_idx, device=instance_lists.valid_num.device)] += 1
    return keep_idx, succ_idx


class Instances3D:
    """Field container for instances in the world frame (reference: instances.py:128-331)."""

    def __init__(self, image_size: Tuple[int, int] = (0, 0), **kwargs: Any):
        self._image_size = image_size
        self._fields: Dict[str, Any] = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self) -> Tuple[int, int]:
        return self._image_size

    def __setattr__(self, name: str, val: Any) -> None:
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name: str) -> Any:
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances3D!".format(name))
        return self._fields[name]

    def set(self, name: str, value: Any) -> None:
        if self._fields:
            n_new = value.shape[0] if isinstance(value, (torch.Tensor, np.ndarray)) else len(value)
            assert len(self) == n_new, \
                "Adding a field of length {} to a Instances3D of length {}".format(n_new, len(self))
        self._fields[name] = value

    def has(self, name: str) -> bool:
        return name in self._fields

    def remove(self, name: str) -> None:
        del self._fields[name]

    def get(self, name: str) -> Any:
        return self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def to(self, *args: Any, **kwargs: Any) -> "Instances3D":
        ret = Instances3D(image_size=self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v.to(*args, **kwargs) if hasattr(v, "to") else v)
        return ret

    def __len__(self) -> int:
        if isinstance(self._fields, fastpath.EngineFields):
            return self._fields.n_rows()                 # a view of the engine state: the row count is known without touching it
        for v in self._fields.values():
            return v.shape[0] if isinstance(v, (torch.Tensor, np.ndarray)) else v.__len__()
        raise NotImplementedError("Empty Instances3D does not support __len__!")

    def __iter__(self):
        raise NotImplementedError("`Instances3D` object is not iterable!")

    def __getitem__(self, item: Union[int, slice, torch.Tensor, np.ndarray, list]) -> "Instances3D":
        if isinstance(self._fields, fastpath.EngineFields) and fastpath.ENABLED:
            fast = self._fields.sess.try_getitem(self, item)          # `all_pred_box[mask]` of demo.py:325 on the engine
            if fast is not None:
                return fast
        if type(item) == int:
            if item >= len(self) or item < -len(self):
                raise IndexError("Instances3D index out of range!")
            item = slice(item, None, len(self))
        ret = Instances3D(image_size=self.image_size)
        dev_item = {}                       # integer index arrays are uploaded once per device, not once per field

        def on(v):
            t = v.tensor if hasattr(v, "tensor") else v
            if not (isinstance(item, np.ndarray) and item.dtype.kind in "iu" and isinstance(t, torch.Tensor) and t.is_cuda):
                return item
            if t.device not in dev_item:
                dev_item[t.device] = torch.from_numpy(np.ascontiguousarray(item, dtype=np.int64)).to(t.device, non_blocking=True)
            return dev_item[t.device]

        for k, v in self._fields.items():
            if isinstance(v, (torch.Tensor, np.ndarray)) or hasattr(v, "tensor"):
                if isinstance(v, np.ndarray) and isinstance(item, torch.Tensor):
                    ret.set(k, v[item.cpu().numpy()])
                else:
                    ret.set(k, v[on(v)])
            elif hasattr(v, "__iter__"):
                if isinstance(item, np.ndarray) and item.dtype == np.bool_:
                    ret.set(k, [x for i, x in enumerate(v) if item[i]])
                elif isinstance(item, torch.Tensor) and item.dtype == torch.bool:
                    ret.set(k, [x for i, x in enumerate(v) if item[i].item()])
                elif isinstance(item, torch.Tensor) and item.dtype == torch.int64:
                    ret.set(k, [v[i.item()] for i in item])
                elif isinstance(item, slice):
                    ret.set(k, v[item])
                else:
                    raise ValueError("Expected Bool or Long Tensor")
            else:
                raise ValueError("Not supported!")
        return ret

    def split(self, split_size_or_sections):
        return [self[s] for s in torch.split(torch.arange(len(self)), split_size_or_sections)]

    def clone(self):
        ret = Instances3D(image_size=self._image_size)
        for k, v in self._fields.items():
            if hasattr(v, "clone"):
                v = v.clone()
            elif isinstance(v, np.ndarray):
                v = np.copy(v)
            elif isinstance(v, (str, list, tuple)):
                v = copy.copy(v)
            else:
                raise NotImplementedError
            ret.set(k, v)
        return ret

    @staticmethod
    def cat(instance_lists: List["Instances3D"]) -> "Instances3D":
        assert all(isinstance(i, Instances3D) for i in instance_lists)
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        sess = fastpath.find_session(*instance_lists)
        if sess is not None:                                          # demo.py:253-254 on the engine: row copies, lazy views
            fast = sess.try_cat(instance_lists)
            if fast is not None:
                return fast
            if sess.bm() is not None:
                sess.detach(sess.bm())
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

    def translate(self, translation):
        for field in self._fields.values():
            if hasattr(field, "translate"):
                field.translate(translation)

    def __str__(self) -> str:
        return "Instances3D(num_instances={}, fields=[{}])".format(len(self), ", ".join(self._fields.keys()))

    __repr__ = __str__

    # ------------------------------------------------------------------------------------------------
    # hot path
    # ------------------------------------------------------------------------------------------------
    def project_3d_boxes(self, K, H=480, W=640):
        """Observation corners of every detection in its own camera (instances.py:333-369) -> projected_boxes."""
        boxes = self.get("pred_boxes_3d")
        cam_pose = self.cam_pose
        dev = ops._pick_device(boxes.tensor)
        K = _as_numpy(K)
        one = ops.shared_pose(cam_pose) if (boxes.tensor.is_cuda and boxes.tensor.is_contiguous() and boxes.R.is_contiguous()) else None
        if one is not None:
            # a keyframe's detections share one pose (demo.py:216): the reference's own inverse call (:350) on that one matrix
            # (batched LU treats every matrix independently, so the values are identical), handed to the kernel by value
            pose_inv = torch.linalg.inv_ex(torch.from_numpy(one)[None], check_errors=False).inverse[0].numpy()
            self.projected_boxes = ops.project_boxes_pose(boxes.tensor, boxes.R, np.ascontiguousarray(pose_inv), K, float(W), float(H))
            self._bf_proj = (K, H, W, one)
            return
        corners = ops.box_corners(boxes.tensor, boxes.R)
        # same call as the reference (:350), on cam_pose's device
        if cam_pose.shape[0] > 1 and bool((cam_pose == cam_pose[:1]).all()):
            pose_inv = torch.linalg.inv(cam_pose[:1]).expand(cam_pose.shape[0], 4, 4)
        else:
            pose_inv = torch.linalg.inv(cam_pose)
        uv = ops.project_boxes(corners, pose_inv.to(dev), K, float(W), float(H))
        self.projected_boxes = uv if boxes.tensor.is_cuda else uv.to(boxes.tensor.device)
        self._bf_proj = (K, H, W, cam_pose[0])              # what the engine-backed path (fastpath.py) needs to know about this keyframe

    def spatial_association(instance_lists, threshold, box_manager, cam_poses):
        """3-D NMS association of map + new detections (instances.py:372-397) -> (keep, success) sorted lists."""
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists                              # reference quirk (:381-382), preserved
        sess = fastpath.session_of(box_manager)
        if sess is not None:
            fast = sess.try_nms(instance_lists, threshold, box_manager)
            if fast is not None:
                return fast
            sess.detach(box_manager)
        boxes = instance_lists.get("pred_boxes_3d")
        dev = ops._pick_device(boxes.tensor)
        corners, centers = ops.box_corners(boxes.tensor, boxes.R, want_centers=True)
        keep, success = _nms_device(instance_lists, box_manager, corners, centers, instance_lists.scores,
                                    instance_lists.init_id, cam_poses, threshold, dev)
        return [int(i) for i in keep], [int(i) for i in success]

    def correspondence_association(cfg, box_manager, cur_keep_idx, cur_success_nms, pred_instances, global_pred_box,
                                   all_pred_box, all_poses, per_frame_ins_cam_pose, frame_id, mask, intrinsic,
                                   all_kf_pose, threshold=0.33, H=480, W=640):
        """2-D correspondence association for small objects (instances.py:411-490)."""
        sess = fastpath.session_of(box_manager)
        if sess is not None:
            fast = sess.try_corr(cfg, box_manager, pred_instances, all_pred_box, all_poses, frame_id, mask, intrinsic, threshold, H, W)
            if fast is not None:
                return fast
            sess.detach(box_manager)
        N_glo = len(global_pred_box)
        keep_idx = copy.deepcopy(np.asarray(mask))
        small_size = cfg["box_fusion"]["small_size"]
        host = all_pred_box.__dict__.pop("_bf_host", None)   # left by spatial_association's download for exactly this call
        if host is not None and host["n"] == len(all_pred_box):
            pred_max = host["max_dim"][N_glo:]
        else:
            host = None
            pred_max = np.max(_as_numpy(pred_instances.get("pred_boxes_3d").dims)[:, :3], axis=1)
        success = set(int(i) for i in cur_success_nms)
        small_idx = [int(i) for i in cur_keep_idx if not (pred_max[i] > small_size or int(i) in success)]
        glo_keep = keep_idx[keep_idx < N_glo]
        if len(small_idx) > 0 and len(glo_keep) > 0:
            gb = global_pred_box.get("pred_boxes_3d")
            dev = ops._pick_device(gb.tensor)
            sel = torch.as_tensor(glo_keep, device=gb.tensor.device)
            g_t, g_R = gb.tensor[sel], gb.R[sel]
            corners = ops.box_corners(g_t, g_R)
            small_mask = (torch.amax(g_t[:, 3:6], dim=1) < small_size + 0.1).to(torch.int32)      # (:460)
            pose_inv = np.linalg.inv(np.asarray(all_kf_pose[frame_id]))                          # as the reference (:680)
            det = pred_instances.pred_boxes[torch.as_tensor(small_idx, device=pred_instances.pred_boxes.device)]
            best, best_iou = ops.corr2d(corners, small_mask, pose_inv.astype(np.float32), _as_numpy(intrinsic),
                                        float(W), float(H), det)
            best, best_iou = ops.to_host(best), ops.to_host(best_iou)
            if host is not None:
                glo_scores, cur_scores = host["scores"][:N_glo], host["scores"][N_glo:]
            else:
                cur_scores = _as_numpy(pred_instances.scores)
                glo_scores = _as_numpy(global_pred_box.scores)
            init_id = _as_numpy(all_pred_box.init_id)
            for k, idx in enumerate(small_idx):                                                  # sequential tail (:463-483)
                if not (best_iou[k] > threshold):
                    continue
                cidx = glo_keep[best[k]]
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

    @staticmethod
    def obb_iou(corners1, corners2):
        """Oriented-3D IoU of two boxes given as corners (instances.py:573-613) -> float."""
        a = np.asarray(corners1, dtype=np.float32).reshape(1, 8, 3)
        b = np.asarray(corners2, dtype=np.float32).reshape(1, 8, 3)
        return float(ops.iou3d_matrix(a, b, IOU_MODE)[0, 0].item())

    @staticmethod
    def augment_vertices(corners):
        """8 corners + the 12 edge mid-points, in the reference's edge order (instances.py:493-512) -> [20,3]."""
        c = np.asarray(corners)
        edges = ((0, 1), (0, 4), (1, 5), (4, 5), (2, 3), (2, 6), (6, 7), (3, 7), (0, 3), (4, 7), (1, 2), (5, 6))
        return np.vstack([c, [(c[a] + c[b]) / 2 for a, b in edges]])

    @staticmethod
    def batch_in_convex_hull_3d(points, corners):
        """Which points satisfy every hull half-space of the box within 1e-6 (instances.py:559-571) -> bool [n] (numpy)."""
        return ops.points_in_hull(points, corners).cpu().numpy()

    @staticmethod
    def check_intersection(corners1, corners2):
        """The containment gate of obb_iou (instances.py:514-557): any of the 20 augmented points of one box inside the other."""
        a1, a2 = Instances3D.augment_vertices(corners1), Instances3D.augment_vertices(corners2)
        return bool(ops.points_in_hull(a1, corners2).any().item() or ops.points_in_hull(a2, corners1).any().item())

    @staticmethod
    def IoU_2D(A, B):
        """AABB of the point set A against boxes B (instances.py:616-641) -> (iou, overlap_A), float64; unused by the
        reference's own pipeline, trivial host arithmetic kept for API completeness."""
        A = np.asarray(A).astype(np.float64)
        B = np.asarray(B)
        x0, y0 = np.min(A, axis=0)
        x1, y1 = np.max(A, axis=0)
        area_a = (x1 - x0) * (y1 - y0)
        area_b = (B[:, 2] - B[:, 0]) * (B[:, 3] - B[:, 1])
        iw = np.maximum(0, np.minimum(x1, B[:, 2]) - np.maximum(x0, B[:, 0]))
        ih = np.maximum(0, np.minimum(y1, B[:, 3]) - np.maximum(y0, B[:, 1]))
        inter = iw * ih
        return inter / (area_a + area_b - inter + 1e-6), inter / (area_a + 1e-6)

    @staticmethod
    def modify_instance(ind_old, ind_new, old_ins, new_ins):
        """Overwrite one instance with another's fields in place (instances.py:400-408; unused by demo.py)."""
        for f in ("scores", "pred_classes", "pred_boxes", "pred_logits", "object_desc", "pred_proj_xy"):
            old_ins.get(f)[ind_old] = new_ins.get(f)[ind_new]
        old_ins.pred_boxes_3d.tensor[ind_old] = new_ins.pred_boxes_3d.tensor[ind_new]
        old_ins.pred_boxes_3d.R[ind_old] = new_ins.pred_boxes_3d.R[ind_new]

    @staticmethod
    def project_3d_to_2d_box(boxes_3d, K, pose, H, W, frame_id=None):
        """Clipped 2-D AABB of map boxes in the current view (instances.py:670-717) -> float64 [N,4] (numpy)."""
        b = np.asarray(boxes_3d, dtype=np.float32).reshape(-1, 8, 3)
        if b.shape[0] == 0:
            return np.zeros((0, 4))
        pose_inv = np.linalg.inv(np.asarray(pose)).astype(np.float32)
        det = np.zeros((1, 4), dtype=np.float32)
        _, _, boxes2d = ops.corr2d(b, np.ones(b.shape[0], dtype=np.int32), pose_inv, _as_numpy(K), float(W), float(H),
                                   det, want_boxes=True)
        return boxes2d.cpu().numpy()

    @staticmethod
    def IoU_2D_box(A, B):
        """Axis-aligned one-vs-many 2-D IoU (instances.py:643-668), float64; trivial host arithmetic kept for
        API completeness (the association path scores inside bf_corr2d)."""
        A = np.asarray(A).astype(np.float64)
        B = np.asarray(B, dtype=np.float64)
        iw = np.maximum(0, np.minimum(A[2], B[:, 2]) - np.maximum(A[0], B[:, 0]))
        ih = np.maximum(0, np.minimum(A[3], B[:, 3]) - np.maximum(A[1], B[:, 1]))
        inter = iw * ih
        return inter / ((A[2] - A[0]) * (A[3] - A[1]) + (B[:, 2] - B[:, 0]) * (B[:, 3] - B[:, 1]) - inter + 1e-6)
