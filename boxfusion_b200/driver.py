"""Synthetic replay of the reference's per-keyframe caller (demo.py:200-327).

`demo.py` itself cannot run (missing SAMCLIP import, detector
checkpoints), so "drops in behind demo.py" is verified by replaying exactly the
calls its loop makes into the hot path, in the same order and with the same
in-place mutation contracts, against any implementation namespace `impl` that
exposes `Instances3D`, `GeneralInstance3DBoxes`, `BoxManager`, `BoxFusion`:

  * the reference itself (oracle/ref_harness.load_reference(), CPU) - this is how
    tests/golden/ is produced;
  * this package (`boxfusion_b200.api`) - the CUDA path under test.

Nothing here is hot-path code; it is the caller.
"""
from __future__ import annotations

import contextlib
import io
from typing import List, Optional

import numpy as np
import torch

from .synthetic import Keyframe


class FusionSession:
    """State that demo.py:run() keeps across keyframes (demo.py:72-83)."""

    def __init__(self, impl, cfg: dict, device: str = "cpu", quiet: bool = True, frame_stride: int = 1):
        """frame_stride: how far demo.py's frame counter `count` advances between two keyframes (cfg data.gap there: the
        detector runs on every gap-th frame, demo.py:134-136, 200); frame ids and check_valid_num's `count - gap` are in frames."""
        self.frame_stride = int(frame_stride)
        self.impl = impl
        self.cfg = cfg
        self.device = torch.device(device)
        self.quiet = quiet
        self.count = 0
        self.all_pred_box = None
        self.all_poses = None
        self.all_kf_pose = {}
        self.per_frame_ins = None
        self.box_count = 0
        self.box_manager = impl.BoxManager(cfg)
        self.box_fuser = impl.BoxFusion(cfg)
        self.last_keep_idx: Optional[np.ndarray] = None
        self.last_mask: Optional[List[int]] = None
        self.last_success: Optional[List[int]] = None

    # demo.py:216-221 ---------------------------------------------------------
    def make_pred_instances(self, kf: Keyframe):
        impl, dev = self.impl, self.device
        n = kf.tensor_cam.shape[0]
        ins = impl.Instances3D((kf.image_size[1], kf.image_size[0]))
        ins.scores = torch.from_numpy(kf.scores.copy()).to(dev)
        ins.pred_boxes = torch.from_numpy(kf.pred_boxes.copy()).to(dev)
        ins.pred_proj_xy = torch.from_numpy(kf.pred_proj_xy.copy()).to(dev)
        ins.pred_boxes_3d = impl.GeneralInstance3DBoxes(
            torch.from_numpy(kf.tensor_cam.copy()).to(dev), torch.from_numpy(kf.R_cam.copy()).to(dev))
        pose_np = np.repeat(kf.pose[None], repeats=n, axis=0)
        ins.cam_pose = torch.from_numpy(pose_np)
        ins.frame_id = torch.tensor([self.count]).repeat(n)
        ins.init_id = self.box_count + torch.arange(n)
        ins.valid_num = torch.zeros(n)
        ins.pred_boxes_3d.transform2world(ins.cam_pose)
        ins.project_3d_boxes(kf.K, H=kf.image_size[1], W=kf.image_size[0])
        return ins, pose_np

    # same container, but from tensors that were already lifted/projected (golden replays feed the
    # product exactly the post-demo.py:221 tensors the reference saw)
    def pred_instances_from_world(self, kf: Keyframe, tensor_w, R_w, projected):
        impl, dev = self.impl, self.device
        n = tensor_w.shape[0]
        ins = impl.Instances3D((kf.image_size[1], kf.image_size[0]))
        ins.scores = torch.from_numpy(kf.scores.copy()).to(dev)
        ins.pred_boxes = torch.from_numpy(kf.pred_boxes.copy()).to(dev)
        ins.pred_proj_xy = torch.from_numpy(kf.pred_proj_xy.copy()).to(dev)
        ins.pred_boxes_3d = impl.GeneralInstance3DBoxes(
            torch.from_numpy(np.ascontiguousarray(tensor_w)).to(dev), torch.from_numpy(np.ascontiguousarray(R_w)).to(dev))
        pose_np = np.repeat(kf.pose[None], repeats=n, axis=0)
        ins.cam_pose = torch.from_numpy(pose_np)
        ins.frame_id = torch.tensor([self.count]).repeat(n)
        ins.init_id = self.box_count + torch.arange(n)
        ins.valid_num = torch.zeros(n)
        ins.projected_boxes = torch.from_numpy(np.ascontiguousarray(projected)).to(dev)
        return ins, pose_np

    def step(self, kf: Keyframe, pred_instances=None, pose_np=None):
        """One keyframe through demo.py:117-118 and :200-327 (gap handling left to the caller)."""
        out = io.StringIO() if self.quiet else None
        with (contextlib.redirect_stdout(out) if self.quiet else contextlib.nullcontext()):
            return self._step(kf, pred_instances, pose_np)

    def _step(self, kf, pred_instances, pose_np):
        cfg, impl, bm = self.cfg, self.impl, self.box_manager
        count = self.count
        if self.box_fuser.update_K_flag is False:                               # demo.py:117-118
            self.box_fuser.update_intrinsics(kf.image_size, kf.K)
        self.all_kf_pose[count] = kf.pose
        if pred_instances is None:
            n = kf.tensor_cam.shape[0]
            if n == 0:                                                          # demo.py:206-212
                bm.num_record[count] = self.box_count
                self.count += self.frame_stride
                return None
            pred_instances, pose_np = self.make_pred_instances(kf)
        self.box_count += len(pred_instances)
        bm.num_record[count] = self.box_count
        keep_idx = None
        if self.all_pred_box is None:                                           # demo.py:228-243
            self.all_pred_box = pred_instances
            self.all_poses = pose_np
            self.per_frame_ins = pred_instances
            bm.init_new_predictions(len(pred_instances), 0)
        else:                                                                   # demo.py:246-327
            bm.init_new_predictions(len(pred_instances), len(self.per_frame_ins))
            num_before_cat = len(self.all_pred_box)
            cur_global = self.all_pred_box
            all_pred_box = impl.Instances3D.cat([self.all_pred_box, pred_instances])
            self.per_frame_ins = impl.Instances3D.cat([self.per_frame_ins, pred_instances])
            all_poses = np.concatenate((self.all_poses, pose_np), axis=0)
            mask, success_mask = impl.Instances3D.spatial_association(
                all_pred_box, cfg["box_fusion"]["nms_threshold"], bm, self.per_frame_ins.cam_pose)
            cur_keep_idx = [i - num_before_cat for i in mask if i >= num_before_cat]
            cur_success_nms = [i - num_before_cat for i in success_mask if i >= num_before_cat]
            keep_idx = np.asarray(mask)
            self.last_mask, self.last_success = mask, success_mask          # index lists as returned (kept for the parity tests)
            if len(cur_keep_idx) > 0:
                all_pred_box, all_poses, keep_idx = impl.Instances3D.correspondence_association(
                    cfg, bm, cur_keep_idx, cur_success_nms, pred_instances, cur_global,
                    all_pred_box, all_poses, self.per_frame_ins.cam_pose, count, mask,
                    torch.from_numpy(kf.K), self.all_kf_pose,
                    threshold=cfg["association"]["small_threshold"],
                    H=kf.image_size[1], W=kf.image_size[0])
                bm.update(keep_idx)
                if cfg["box_fusion"]["check_valid"]:
                    all_pred_box = bm.check_valid_num(all_pred_box, count, cfg["data"]["gap"])
                if cfg["box_fusion"]["use"]:
                    self.box_fuser.boxfusion(all_pred_box, self.per_frame_ins, bm)
            else:
                all_pred_box = all_pred_box[mask]
                all_poses = all_poses[mask]
                bm.update(keep_idx)
            self.all_pred_box, self.all_poses = all_pred_box, all_poses
        self.last_keep_idx = None if keep_idx is None else np.asarray(keep_idx).copy()
        self.count += self.frame_stride
        return self.last_keep_idx

    # snapshot of everything the API mutates, for parity comparison -----------
    def snapshot(self) -> dict:
        apb, bm = self.all_pred_box, self.box_manager
        fl = bm.fusion_list
        flat = np.array([x for l in fl for x in l], dtype=np.int64)
        off = np.cumsum([0] + [len(l) for l in fl]).astype(np.int64)
        af = bm.already_fusion
        return {
            "tensor": apb.pred_boxes_3d.tensor.detach().cpu().numpy().copy(),
            "R": apb.pred_boxes_3d.R.detach().cpu().numpy().copy(),
            "scores": apb.scores.detach().cpu().numpy().copy(),
            "valid_num": apb.valid_num.detach().cpu().numpy().copy(),
            "init_id": apb.init_id.detach().cpu().numpy().copy(),
            "fusion_flat": flat, "fusion_off": off,
            "fusion_flag": np.asarray(bm.fusion_flag, dtype=np.int64),
            "already_flat": np.array([x for l in af for x in l], dtype=np.int64),
            "already_off": np.cumsum([0] + [len(l) for l in af]).astype(np.int64),
            "keep_idx": (np.zeros(0, np.int64) if self.last_keep_idx is None
                         else self.last_keep_idx.astype(np.int64)),
        }
