"""`boxfusion.box_manager.BoxManager` for the B200 path (reference: boxfusion/box_manager.py:9-245).

The manager owns the host-visible bookkeeping the reference API exposes (`fusion_list`, `fusion_flag`,
`already_fusion`, `num_record`) as plain Python lists, exactly like the reference, because `demo.py`
prints and re-indexes them.  The decisions that fill them are taken on the GPU:

  * during `Instances3D.spatial_association` the whole of `record()` runs inside the greedy NMS
    kernel (bf_nms3d); the manager only packs its lists into the kernel's fixed-capacity layout
    (`pack_lists`) and applies the result (`apply_lists`);
  * `record()` / `record_corr()` called directly (the reference's public methods; `record_corr` is
    used by correspondence_association) evaluate all pose-disparity predicates of the call in one
    bf_pose_disparity launch and then replay the reference's list logic on the host.
"""
from __future__ import annotations

import copy
import itertools
from typing import Dict, List

import numpy as np
import torch

from . import ops


class BoxManager:

    def __init__(self, cfg):
        self._session = None                        # fastpath.Session while the manager's state lives in a FusionEngine
        self._fast_engine, self._fast_strikes = None, 0
        self._fusion_list: List[List[int]] = []    # per map box: per-frame observation indices supporting it
        self.last_fusion_frame: List[List[int]] = []
        self._fusion_flag: List[int] = []
        self._already_fusion: List[List[int]] = []
        self._fused_set, self._fused_n = set(), 0
        self.num_record: Dict[int, int] = {}
        self.cfg = cfg
        self.rotation_gap = self.cfg["association"]["rotation_gap"]
        self.translation_gap = self.cfg["association"]["translation_gap"]
        self.small_size = self.cfg["box_fusion"]["small_size"]
        self.merge_log: List[Dict] = []

    # ---- the public lists (box_manager.py:13-16).  While a fastpath.Session runs they live on the device and are downloaded
    #      when somebody looks; assigning them, or any of the list-editing methods below, ends the session first ----------
    def _lists(self, name):
        if self._session is not None:
            self._session.lists_for_caller(self)
        return getattr(self, name)

    def _leave_fast_path(self):
        if self._session is not None:
            self._session.detach(self)

    fusion_list = property(lambda self: self._lists("_fusion_list"))
    fusion_flag = property(lambda self: self._lists("_fusion_flag"))
    already_fusion = property(lambda self: self._lists("_already_fusion"))

    @fusion_list.setter
    def fusion_list(self, v):
        self._leave_fast_path()
        self._fusion_list = v

    @fusion_flag.setter
    def fusion_flag(self, v):
        self._leave_fast_path()
        self._fusion_flag = v

    @already_fusion.setter
    def already_fusion(self, v):
        self._leave_fast_path()
        self._already_fusion = v

    # ---- bookkeeping (box_manager.py:24-38, 131-166) ---------------------------------------------
    def init_new_predictions(self, box_num, all_num):
        s = self._session
        if s is not None:
            if s.stage in (s.IDLE, s.CORR) and int(all_num) == s.M + (s.n if s.stage == s.CORR else 0):
                s.pending_new = (int(box_num), int(all_num))      # the engine creates these rows when the detections are ingested
                s.lists_host = False
                return
            s.detach(self)
        for i in range(box_num):
            self.fusion_list.append([i + all_num])
            self.last_fusion_frame.append([0])
            self.fusion_flag.append(0)

    def add_fusion_ind(self, idx_list):
        self._leave_fast_path()
        if self._fused_n == len(self.already_fusion):        # keep the lookup set in step with the public list
            self._fused_set.add(tuple(idx_list))
            self._fused_n += 1
        self.already_fusion.append(copy.deepcopy(idx_list))

    def check_if_fusion(self, idx_list):
        """`idx_list in self.already_fusion` (box_manager.py:34-38) through a set of tuples kept in step with the
        public list (rebuilt if a caller edited `already_fusion` directly)."""
        if self._fused_n != len(self.already_fusion):
            self._fused_set = {tuple(l) for l in self.already_fusion}      # int and numpy integers hash/compare alike
            self._fused_n = len(self.already_fusion)
        return tuple(idx_list) in self._fused_set

    def update(self, keep_idx):
        s = self._session
        if s is not None:
            if s.stage == s.CORR and len(keep_idx) == s.N:
                return                                            # the engine compacted the lists together with the map rows
            s.detach(self)
        self.fusion_list = [self.fusion_list[i] for i in keep_idx]

    def update_fusion_flag(self, idx):
        self._leave_fast_path()
        self.fusion_flag[idx] = 1

    def get_fusion_idx(self):
        return [i for i in range(len(self.fusion_flag)) if self.fusion_flag[i] == 1]

    def get_nofusion_idx(self):
        return [i for i in range(len(self.fusion_flag)) if self.fusion_flag[i] == 0]

    def check_valid_num(self, all_pred_box, count, gap):
        s = self._session
        if s is not None:
            fast = s.try_check_valid(self, all_pred_box, count, gap)
            if fast is not None:
                return fast
            s.detach(self)
        zero = torch.where((all_pred_box.valid_num == 0) & (all_pred_box.frame_id < (count - gap)))[0]
        valid = torch.arange(len(all_pred_box))
        if zero.shape[0] > 0:
            drop = torch.zeros(len(all_pred_box), dtype=torch.bool)
            drop[zero.cpu()] = True
            valid = valid[~drop]
        self.fusion_list = [self.fusion_list[int(i)] for i in valid]
        return all_pred_box[valid]

    # ---- device layout of the lists (include/boxfusion_b200.h: bf_nms3d) -------------------------
    def pack_lists(self, n: int):
        """fusion_list/fusion_flag of the first n boxes -> (list[n,CAP] i32, len[n] i32, flag[n] i32) numpy."""
        cap = ops.FUSION_CAP
        lists = self._fusion_list[:n]
        ln = np.fromiter(map(len, lists), dtype=np.int32, count=n)
        if n and ln.max() > cap:
            raise RuntimeError(f"a fusion list has {int(ln.max())} entries; device capacity is {cap}")
        fl = np.zeros((n, cap), dtype=np.int32)
        total = int(ln.sum())
        flat = np.fromiter(itertools.chain.from_iterable(lists), dtype=np.int64, count=total)
        rows = np.repeat(np.arange(n), ln)
        cols = np.arange(total) - np.repeat(np.cumsum(ln) - ln, ln)
        fl[rows, cols] = flat
        flag = np.asarray(self._fusion_flag[:n], dtype=np.int32)
        return fl, ln, flag

    def apply_lists(self, fl: np.ndarray, ln: np.ndarray, flag: np.ndarray, old_len: np.ndarray):
        """Write back rows the kernel changed (in place, like `fusion_list[cur] += ...; .sort()`)."""
        self._leave_fast_path()
        for i in np.nonzero(ln != old_len)[0]:
            self.fusion_list[i][:] = [int(x) for x in fl[i, :ln[i]]]
        for i in np.nonzero(flag != np.asarray(self.fusion_flag[:len(flag)], dtype=np.int32))[0]:
            self.fusion_flag[i] = int(flag[i])

    # ---- pose disparity (box_manager.py:168-215) ---------------------------------------------------
    def compute_pose_disparity(self, pose1, pose2):
        poses = torch.stack([torch.as_tensor(pose1, dtype=torch.float32), torch.as_tensor(pose2, dtype=torch.float32)])
        b, a = ops.pose_disparity(poses.reshape(2, 16), [0], [1])
        b, a = b[0].cpu(), a[0].cpu()
        return b, a, 0.6 * b + 0.4 * a

    def compute_pose_center_disparity(self, pose1, pose2, center1, center2):
        b, a, s = self.compute_pose_disparity(pose1, pose2)
        return b, a, s, self.euclidean_distance_3d(center1, center2)

    def euclidean_distance_3d(self, point1, point2):
        return np.sqrt(np.sum((np.asarray(point1) - np.asarray(point2)) ** 2))

    def _differs_batch(self, cam_poses, ia, ib):
        """One launch for all (view, view) predicates of a record call: baseline/angle vs the gaps."""
        if len(ia) == 0:
            return np.zeros(0, dtype=bool)
        base, ang = ops.pose_disparity(cam_poses, ia, ib)
        res = (base > self.translation_gap) | (ang > self.rotation_gap)
        return res.cpu().numpy()

    def record(self, cur_id, fusion_inds, init_id, cam_poses, box_size, keep, box_centers):
        """box_manager.py:40-88 (host replay; predicates from bf_pose_disparity)."""
        self._leave_fast_path()
        fl = self.fusion_list
        for idx in fusion_inds:
            cdis = self.euclidean_distance_3d(box_centers[cur_id], box_centers[idx]) > 0.5
            if len(fl[idx]) == 1:
                views, other = list(fl[cur_id]), int(init_id[idx])
            else:
                views, other = list(fl[idx]), int(init_id[cur_id])
            d = self._differs_batch(cam_poses, [int(v) for v in views], [other] * len(views))
            count = int(np.sum(d | cdis))
            if len(fl[idx]) == 1:
                if count == len(fl[cur_id]) and len(fl[cur_id]) < 5:
                    fl[cur_id] += [init_id[idx]]
                    fl[cur_id].sort()
            else:
                if count == len(fl[idx]) and len(fl[idx]) < 5:
                    fl[cur_id] += fl[idx]
                    fl[cur_id].sort()
                elif cur_id in keep:
                    keep.remove(cur_id)
                    keep.append(idx)
                if self.fusion_flag[idx] == 1:
                    self.fusion_flag[cur_id] = 1
        return keep

    def record_corr(self, cur_id, fusion_inds, init_id, cam_poses, keep):
        """box_manager.py:90-129 (host replay; predicates from bf_pose_disparity)."""
        self._leave_fast_path()
        fl = self.fusion_list
        for idx in fusion_inds:
            if len(fl[idx]) == 1:
                views, other = list(fl[cur_id]), int(init_id[idx])
            else:
                views, other = list(fl[idx]), int(init_id[cur_id])
            count = int(np.sum(self._differs_batch(cam_poses, [int(v) for v in views], [other] * len(views))))
            if len(fl[idx]) == 1:
                if count == len(fl[cur_id]) and len(fl[cur_id]) < 5:
                    fl[cur_id] += [init_id[idx]]
                    fl[cur_id].sort()
            else:
                if count == len(fl[idx]) and len(fl[idx]) < 5:
                    fl[cur_id] += fl[idx]
                    fl[cur_id].sort()
                elif cur_id in keep:
                    keep[keep == cur_id] = idx
                if self.fusion_flag[idx] == 1:
                    self.fusion_flag[cur_id] = 1
        return keep

    def filter_detections(self, pred_instances, W, H, score_thresh=0.0):
        """demo.py:138-148 fused (score threshold + the three masks below) -> bool keep mask, one kernel.  The individual
        check_* methods below keep the reference's signatures."""
        det = self.cfg["detection"]
        keep, _ = ops.detection_filter(pred_instances.pred_boxes_3d.tensor, pred_instances.pred_proj_xy, pred_instances.scores,
                                       W, H, float(score_thresh), det["uv_bound_value"] if det.get("uv_bound") else None,
                                       det["floor_ratio"] if det.get("floor_mask") else None, det.get("size_max_thres") or None)
        return keep

    # ---- detection pre-filters (box_manager.py:217-245): stay in torch on the detector's device -----
    def check_uv_bounds(self, uv_coords, W, H, ratio=1.0):
        gap_W, gap_H = int((1 - ratio) * W), int((1 - ratio) * H)
        u, v = uv_coords[:, 0], uv_coords[:, 1]
        return (u > gap_W) & (u < (W - gap_W)) & (v > gap_H) & (v < (H - gap_H))

    def check_floor_mask(self, box_3d, ratio=20):
        size = box_3d[:, 3:]
        mx, mn = torch.amax(size, dim=1), torch.amin(size, dim=1)
        second = torch.sort(size, dim=1, descending=True)[0][:, 1]
        thin = (mx / mn > ratio / 2) & (mx / second > ratio / 2) & (second / mn < 2.0) & (second < 0.15) & (mn < 0.15)
        return (mx / mn > ratio) | thin

    def check_large_mask(self, box_3d, thres=0.5):
        return torch.amax(box_3d[:, 3:], dim=1) > thres
