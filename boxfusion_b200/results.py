"""Result / wire formats of the reference, so an external evaluator gets identical artefacts.

  post_process(boxes, threshold)   tools/utils.py:302-317   ScanNet size filter on corner arrays [N,8,3]
  global_save_list / framewise_save_list / save_box / load_data
                                   demo.py:369-387, tools/utils.py:322-340   pickles of [[(class, corners[8,3], feature)]]

Corners come from the CUDA library (bf_box_corners); the rest is file formatting."""
from __future__ import annotations

import pickle
from typing import List, Sequence

import numpy as np
import torch

from . import ops


def post_process(boxes, threshold: float = 0.3):
    """Keep boxes whose axis-aligned extent is >= threshold on every axis (tools/utils.py:302-317).
    Accepts numpy or torch [N,8,3]; returns the same kind."""
    if isinstance(boxes, torch.Tensor):
        ranges = boxes.amax(dim=1) - boxes.amin(dim=1)
        return boxes[(ranges >= threshold).all(dim=1)]
    ranges = np.max(boxes, axis=1) - np.min(boxes, axis=1)
    return boxes[(ranges[:, 0] >= threshold) & (ranges[:, 1] >= threshold) & (ranges[:, 2] >= threshold)]


def map_corners(all_pred_box) -> np.ndarray:
    """`all_pred_box.pred_boxes_3d.corners.cpu().numpy()` (demo.py:373) through bf_box_corners."""
    b = all_pred_box.pred_boxes_3d
    if len(b) == 0:
        return np.zeros((0, 8, 3), dtype=np.float32)
    c = ops.to_host(ops.box_corners(b.tensor, b.R))
    # the reference's corners are the transpose(1, 2) VIEW of a [N,3,8] tensor (boxes.py:767-778), so every [8,3] array it
    # pickles is Fortran-ordered; same memory layout here -> the pickles are byte-identical, not just equal in value
    return np.ascontiguousarray(c.transpose(0, 2, 1)).transpose(0, 2, 1)


def global_save_list(all_pred_box, dataset: str = "CA1M") -> list:
    """demo.py:373-378: [[(0, corners[8,3], 1.0) per map box]] (ScanNet: after post_process)."""
    boxes = map_corners(all_pred_box)
    if dataset == "scannet":
        boxes = post_process(boxes)
    return [[(int(0), boxes[n], 1.0) for n in range(boxes.shape[0])]]


def framewise_save_list(per_frame_ins, class_index: Sequence[int], features: Sequence) -> list:
    """demo.py:383-386: [[(class_idx[n], corners[8,3], feature[n]) per observation]]."""
    boxes = map_corners(per_frame_ins)
    return [[(class_index[n], boxes[n], features[n]) for n in range(boxes.shape[0])]]


def save_box(data, filename) -> None:
    """tools/utils.py:322-332."""
    with open(filename, "wb") as f:
        pickle.dump(data, f, protocol=pickle.HIGHEST_PROTOCOL)


def load_data(filename):
    """tools/utils.py:335-340."""
    with open(filename, "rb") as f:
        return pickle.load(f)
