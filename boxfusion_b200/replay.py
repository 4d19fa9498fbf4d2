"""Recorder / player for hot-path inputs.

`KeyframeRecorder` dumps, per keyframe, exactly what `demo.py` hands to the fusion path after the detector and its
pre-filters (demo.py:138-148, 200-221): camera-frame boxes, R, scores, 2-D boxes, projected centres, the camera pose,
K and the image size - into one compressed `.npz` per sequence.  A machine that has the detector records once; parity
and performance of the fusion path can then be checked anywhere with `load_sequence()` + `FusionSession` / `FusionEngine`,
without the detector in the loop."""
from __future__ import annotations

from typing import List

import numpy as np

from .synthetic import Keyframe

_FIELDS = ("tensor_cam", "R_cam", "scores", "pred_boxes", "pred_proj_xy")


class KeyframeRecorder:
    def __init__(self):
        self._frames: List[Keyframe] = []

    def add(self, frame_id: int, pose, K, image_size, tensor_cam, R_cam, scores, pred_boxes, pred_proj_xy) -> None:
        """Call right before demo.py:216 with `pred_instances` fields (torch tensors or arrays, any device)."""
        def arr(x, dt=np.float32):
            x = x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
            return np.ascontiguousarray(x, dtype=dt)
        n = arr(scores).shape[0]
        self._frames.append(Keyframe(
            frame_id=int(frame_id), pose=arr(pose).reshape(4, 4), K=arr(K).reshape(3, 3), image_size=(int(image_size[0]), int(image_size[1])),
            tensor_cam=arr(tensor_cam).reshape(n, 6), R_cam=arr(R_cam).reshape(n, 3, 3), scores=arr(scores).reshape(n),
            pred_boxes=arr(pred_boxes).reshape(n, 4), pred_proj_xy=arr(pred_proj_xy).reshape(n, 2), gt_index=-np.ones(n, dtype=np.int64)))

    def add_keyframe(self, kf: Keyframe) -> None:
        self._frames.append(kf)

    def save(self, path: str) -> None:
        out = {"n_frames": np.array(len(self._frames))}
        for i, kf in enumerate(self._frames):
            out[f"f{i}_frame_id"] = np.array(kf.frame_id)
            out[f"f{i}_pose"], out[f"f{i}_K"], out[f"f{i}_size"] = kf.pose, kf.K, np.array(kf.image_size)
            for k in _FIELDS:
                out[f"f{i}_{k}"] = getattr(kf, k)
        np.savez_compressed(path, **out)


def load_sequence(path: str) -> List[Keyframe]:
    g = np.load(path)
    frames = []
    for i in range(int(g["n_frames"])):
        n = g[f"f{i}_scores"].shape[0]
        frames.append(Keyframe(frame_id=int(g[f"f{i}_frame_id"]), pose=g[f"f{i}_pose"], K=g[f"f{i}_K"],
                               image_size=tuple(int(x) for x in g[f"f{i}_size"]),
                               **{k: g[f"f{i}_{k}"] for k in _FIELDS}, gt_index=-np.ones(n, dtype=np.int64)))
    return frames
